#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder.py -q -m gpu -s 2>&1 | tail -25 > gpurun_out/r2g_decoder_tests.log
(time timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -8) > gpurun_out/r2g_all_gpu_tests.log 2>&1
for cfgs in "fp16 0" "fp16 1" "fp32tc 1"; do set -- $cfgs
timeout 600 python bench.py --precision $1 --latent-table $2 --steps 3 --warmup 3 --no-variants --no-cpu-baseline --no-extras > gpurun_out/r2g_bench_$1_$2.json 2> gpurun_out/r2g_bench_$1_$2.err
done
cat gpurun_out/r2g_decoder_tests.log; cat gpurun_out/r2g_all_gpu_tests.log
for cfgs in "fp16 0" "fp16 1" "fp32tc 1"; do set -- $cfgs; echo "== $1 table=$2"; cut -c1-150 gpurun_out/r2g_bench_$1_$2.json; done
