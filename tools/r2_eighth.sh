#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(time timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -15) > gpurun_out/r2h_all_gpu_tests.log 2>&1
timeout 300 python -m pytest tests/test_decoder.py -q -m gpu -s 2>&1 | grep -E "^1_|passed|failed" > gpurun_out/r2h_decoder.log
for cfgs in "fp16 0" "fp16 1" "fp32tc 0" "fp32tc 1"; do set -- $cfgs
SRF_TC_PROF=1 timeout 600 python bench.py --precision $1 --latent-table $2 --steps 3 --warmup 3 --no-variants --no-cpu-baseline --no-extras > gpurun_out/r2h_bench_$1_$2.json 2> gpurun_out/r2h_bench_$1_$2.err
done
cat gpurun_out/r2h_all_gpu_tests.log gpurun_out/r2h_decoder.log
for cfgs in "fp16 0" "fp16 1" "fp32tc 0" "fp32tc 1"; do set -- $cfgs; echo "== $1 table=$2"; cut -c1-150 gpurun_out/r2h_bench_$1_$2.json; grep prof gpurun_out/r2h_bench_$1_$2.err | sort | uniq -c | sort -rn | head -1 | cut -c1-330; done
