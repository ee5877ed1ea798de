#!/bin/bash
# round 2, first GPU pass: the split (fp32tc) kernel -- layer tests, parity, and a first config-B timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fp32tc.py -x -q -m gpu -s 2>&1 | tail -60 > gpurun_out/r2_fp32tc_tests.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc_layers.py -q -m gpu -s 2>&1 | tail -120 > gpurun_out/r2_parity_tests.log
SRF_TC_PROF=1 timeout 600 python bench.py --precision fp32tc --steps 3 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/r2_bench_fp32tc.json 2> gpurun_out/r2_bench_fp32tc.err
timeout 600 python bench.py --precision fp16 --steps 3 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/r2_bench_fp16.json 2> gpurun_out/r2_bench_fp16.err
tail -5 gpurun_out/r2_fp32tc_tests.log; tail -5 gpurun_out/r2_parity_tests.log; cat gpurun_out/r2_bench_fp32tc.json | cut -c1-600; tail -3 gpurun_out/r2_bench_fp32tc.err; cut -c1-300 gpurun_out/r2_bench_fp16.json
