#!/bin/bash
# round 2: multi-GPU line (weak frame-per-GPU + strong + workloads D/E) and the NCCL equality test on N GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nccl.py tests/test_sweep.py -q -m gpu -s -k "two_gpus or sharded" 2>&1 | tail -12 > gpurun_out/r2_nccl_tests_${N}gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
cat gpurun_out/r2_nccl_tests_${N}gpu.log; cut -c1-250 gpurun_out/r2_bench_${N}gpu.json; tail -3 gpurun_out/r2_bench_${N}gpu.err
