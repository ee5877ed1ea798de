"""Here (no GPU): condense `ncu -i X.ncu-rep --page raw --csv` (one kernel) into the metric,unit,value table kept under profiles/.
    python tools/ncu_summary.py gpurun_out/X_raw.csv profiles/X_ncu_summary.csv"""
import csv
import re
import sys

KEEP = re.compile(r"^(ID|Kernel Name|Block Size|Grid Size|dram__bytes_(read|write)\.sum(\.|$)|gpc__cycles_elapsed\.max|gpu__time_duration\.sum|"
                  r"l1tex__throughput\.avg|launch__(cluster|grid_size|registers|shared_mem)|lts__t_sector_hit_rate|lts__throughput\.avg|"
                  r"lts__t_sectors\.sum(\.|$)|lts__t_bytes\.sum(\.|$)|sm__cycles_elapsed\.avg|sm__inst_executed_pipe_tensor|sm__pipe_tensor_cycles_active\.avg|"
                  r"sm__throughput\.avg|sm__warps_active\.avg|smsp__average_warps_issue_stalled_.*_per_issue_active|smsp__inst_executed\.sum$|"
                  r"smsp__issue_active\.avg|sm__inst_executed_pipe_(alu|fma|fp16|lsu|uniform|xu)\.(sum|avg\.pct))")
csv.field_size_limit(10 ** 9)
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
names, units, vals = rows[0], rows[1], rows[2]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    for n, u, v in zip(names, units, vals):
        short = n.split(".", 2)[-1] if n.split(".")[0].isupper() and "TriageCompute" in n else n
        if KEEP.match(short) and (v != "" or short in ("ID",)):
            w.writerow([short, u, v])
