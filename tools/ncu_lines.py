"""Here (no GPU): aggregate an ncu --import-source capture by CUDA source line.
    python tools/ncu_lines.py gpurun_out/X.ncu-rep [top_n]
Prints the source lines of mlp_tc.cu with the most warp-stall samples, their dominant stall reasons and instruction counts."""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
csv.field_size_limit(10 ** 9)
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if "Source" in r and any("Sampling" in c for c in r))
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
samp = col.get("Warp Stall Sampling (All Samples)") or col.get("# Samples")
stall_cols = [(n, i) for n, i in col.items() if n.startswith("stall_") and "Not Issued" not in n]
inst = col.get("Instructions Executed")
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
cur_src = None
total = 0
for r in rows[hdr_i + 1:]:
    if len(r) <= samp:
        continue
    src = r[col["Source"]]
    try:
        s = int(float(r[samp] or 0))
    except ValueError:
        continue
    key = src.strip()[:150]
    a = agg[key]
    a[0] += s
    total += s
    if inst is not None:
        try:
            a[1] += int(float(r[inst] or 0))
        except ValueError:
            pass
    for n, i in stall_cols:
        try:
            a[2][n] += int(float(r[i] or 0))
        except (ValueError, IndexError):
            pass
print("total samples", total, "rows", len(rows) - hdr_i - 1)
for key, (s, ins, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    reasons = ", ".join("%s %d%%" % (n[6:], 100 * v // max(1, s)) for n, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print("%6.2f%%  inst %10d  %-40s | %s" % (100.0 * s / max(1, total), ins, reasons, key))
