#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
cut -c1-250 gpurun_out/r2_bench_8gpu.json; tail -3 gpurun_out/r2_bench_8gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_8gpu.json 2> gpurun_out/r2_bench_reference_8gpu.err
cut -c1-200 gpurun_out/r2_bench_reference_8gpu.json
