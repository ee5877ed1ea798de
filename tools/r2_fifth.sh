#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
timeout 900 python -m pytest tests/test_gpu_preproj.py -q -m gpu -s -x -k "adversarial or kitti_s128 or full_size" 2>&1 | grep -E "max-abs-err|Error|error|assert|FAILED|passed|failed|watchdog" | tail -30 > gpurun_out/r2e_preproj_tests.log
for prec in fp16 fp32tc; do
SRF_TC_PROF=1 timeout 600 python bench.py --precision $prec --latent-table 1 --steps 3 --warmup 3 --no-variants --no-cpu-baseline --no-extras > gpurun_out/r2e_bench_${prec}_table.json 2> gpurun_out/r2e_bench_${prec}_table.err
done
COMMON="--steps 1 --warmup 3 --no-variants --no-cpu-baseline --no-extras"
cap() { name=$1; shift
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:point_mlp_tc -s 7 -c 1 -o /tmp/ncu/$name python bench.py "$@" $COMMON > gpurun_out/$name.log 2>&1; echo "ncu $name rc=$?"
  ncu -i /tmp/ncu/$name.ncu-rep --page source --csv > gpurun_out/${name}_sass.csv 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null; }
cap r2e_prof_fp16_table --precision fp16 --latent-table 1 --rays 60000
cap r2e_prof_fp16_dense --precision fp16 --rays 60000
tail -12 gpurun_out/r2e_preproj_tests.log
for prec in fp16 fp32tc; do cut -c1-160 gpurun_out/r2e_bench_${prec}_table.json; grep prof gpurun_out/r2e_bench_${prec}_table.err | sort | uniq -c | sort -rn | head -3 | cut -c1-330; done
du -sh gpurun_out
