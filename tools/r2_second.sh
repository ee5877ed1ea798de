#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fp32tc.py -q -m gpu -s 2>&1 | tail -40 > gpurun_out/r2b_fp32tc_tests.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "fp32tc or full_size" 2>&1 | grep -E "max-abs-err|RaySOM|Error|assert|FAILED|passed|failed" > gpurun_out/r2b_parity_tests.log
SRF_TC_PROF=1 timeout 600 python bench.py --precision fp32tc --steps 2 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/r2b_bench_fp32tc.json 2> gpurun_out/r2b_bench_fp32tc.err
SRF_TC_PROF=1 timeout 600 python bench.py --precision fp32tc --skip-zero-chunks 1 --steps 2 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/r2b_bench_fp32tc_skip.json 2> gpurun_out/r2b_bench_fp32tc_skip.err
tail -5 gpurun_out/r2b_fp32tc_tests.log; tail -25 gpurun_out/r2b_parity_tests.log; cut -c1-200 gpurun_out/r2b_bench_fp32tc.json; tail -2 gpurun_out/r2b_bench_fp32tc.err; cut -c1-200 gpurun_out/r2b_bench_fp32tc_skip.json; tail -1 gpurun_out/r2b_bench_fp32tc_skip.err
