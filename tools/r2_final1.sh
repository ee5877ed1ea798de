#!/bin/bash
# round 2, final single-GPU measurements (the JSON lines are copied to profiles/)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --workload decoder --steps 5 --warmup 3 > gpurun_out/r2_decoder_1gpu.json 2> gpurun_out/r2_decoder_1gpu.err; echo "decoder rc=$?"; cut -c1-400 gpurun_out/r2_decoder_1gpu.json; tail -2 gpurun_out/r2_decoder_1gpu.err
timeout 300 python bench.py --workload train --train-matmul tf32 --steps 10 --warmup 3 > gpurun_out/r2_train_tf32_1gpu.json 2> gpurun_out/r2_train_tf32_1gpu.err; echo "train rc=$?"
timeout 600 python bench.py --workload A --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_A.json 2> gpurun_out/r2_bench_A.err; echo "A rc=$?"; cut -c1-200 gpurun_out/r2_bench_A.json
timeout 600 python bench.py --workload C --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_C.json 2> gpurun_out/r2_bench_C.err; echo "C rc=$?"; cut -c1-200 gpurun_out/r2_bench_C.json
timeout 1200 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "default rc=$?"; cut -c1-300 gpurun_out/r2_bench_default.json; tail -3 gpurun_out/r2_bench_default.err
timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "reference rc=$?"; cut -c1-300 gpurun_out/r2_bench_reference.json
