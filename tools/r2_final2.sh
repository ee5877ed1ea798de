#!/bin/bash
# round 2, final single-GPU validation of the committed kernels: whole GPU suite, smoke, default bench + reference arm,
# ncu captures of the three main regimes (CSV exported on the box), secondary workloads
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_final_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/r2_final_smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "default rc=$?"; cut -c1-260 gpurun_out/r2_bench_default.json; tail -2 gpurun_out/r2_bench_default.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "reference rc=$?"; cut -c1-260 gpurun_out/r2_bench_reference.json
COMMON="--steps 1 --warmup 3 --no-variants --no-cpu-baseline --no-extras"
cap() {  # name, timeout, bench args...
  name=$1; to=$2; shift 2
  timeout $to ncu --set full --clock-control none --import-source on -k regex:point_mlp_tc -s 7 -c 1 -o /tmp/ncu/$name python bench.py "$@" $COMMON > gpurun_out/$name.log 2>&1; echo "ncu $name rc=$?"
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
}
cap r2f_prof_fp16_dense 300 --precision fp16 --rays 60000
cap r2f_prof_fp16_table 300 --precision fp16 --latent-table 1 --rays 60000
cap r2f_prof_fp32tc_dense 400 --precision fp32tc --rays 60000
timeout 300 python bench.py --workload A --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_A.json 2> gpurun_out/r2_bench_A.err; echo "A rc=$?"; cut -c1-200 gpurun_out/r2_bench_A.json
timeout 300 python bench.py --workload C --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_C.json 2> gpurun_out/r2_bench_C.err; echo "C rc=$?"; cut -c1-200 gpurun_out/r2_bench_C.json
timeout 300 python bench.py --workload sweep --precision fp16 --sweep-table 1 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_sweep_fp16_table.json 2> gpurun_out/r2_sweep_fp16_table.err; echo "sweep rc=$?"; cut -c1-200 gpurun_out/r2_sweep_fp16_table.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_default.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-variants --no-extras > gpurun_out/r2_ncu_launch_bench.log 2>&1; echo "ncu launches rc=$?"
du -sh gpurun_out
