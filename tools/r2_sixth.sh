#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fp32tc.py tests/test_gpu_tc_layers.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/r2f_tests_a.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_preproj.py -q -m gpu -x -k "not full_size" 2>&1 | tail -5 > gpurun_out/r2f_tests_b.log
for cfgs in "fp16 0" "fp16 1" "fp32tc 0" "fp32tc 1"; do set -- $cfgs
SRF_TC_PROF=1 timeout 600 python bench.py --precision $1 --latent-table $2 --steps 3 --warmup 3 --no-variants --no-cpu-baseline --no-extras > gpurun_out/r2f_bench_$1_$2.json 2> gpurun_out/r2f_bench_$1_$2.err
done
cat gpurun_out/r2f_tests_a.log gpurun_out/r2f_tests_b.log
for cfgs in "fp16 0" "fp16 1" "fp32tc 0" "fp32tc 1"; do set -- $cfgs; echo "== $1 table=$2"; cut -c1-150 gpurun_out/r2f_bench_$1_$2.json; grep prof gpurun_out/r2f_bench_$1_$2.err | sort | uniq -c | sort -rn | head -2 | cut -c1-330; done
