#!/bin/bash
# round 2 profiling pass (1 GPU): ncu captures exported to CSV on the box (the .ncu-rep files stay there: 64 MiB return limit)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
COMMON="--steps 1 --warmup 3 --no-variants --no-cpu-baseline --no-extras"
cap() {  # name, timeout, bench args...
  name=$1; to=$2; shift 2
  timeout $to ncu --set full --clock-control none --import-source on -k regex:point_mlp_tc -s 7 -c 1 -o /tmp/ncu/$name python bench.py "$@" $COMMON > gpurun_out/$name.log 2>&1; echo "ncu $name rc=$?"
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page source --csv --print-source cuda > gpurun_out/${name}_source_cuda.csv 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
  ls -la /tmp/ncu/$name.ncu-rep gpurun_out/${name}_*
}
cap r2_prof_fp16_table 900 --precision fp16 --latent-table 1 --rays 60000
cap r2_prof_fp32tc_table 900 --precision fp32tc --latent-table 1 --rays 60000
cap r2_prof_fp32tc_dense 1500 --precision fp32tc
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_default.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-variants --no-extras > gpurun_out/r2_ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
du -sh gpurun_out
