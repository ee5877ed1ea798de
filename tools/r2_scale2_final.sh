#!/bin/bash
# round 2, final kernels on 2 GPUs: NCCL equality tests + the torchrun bench line (weak + strong + workloads D/E), short form
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_nccl.py -q -m gpu -s 2>&1 | tail -8 > gpurun_out/r2_final_nccl_tests_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/r2_bench_2gpu_final.json 2> gpurun_out/r2_bench_2gpu_final.err
cat gpurun_out/r2_final_nccl_tests_2gpu.log; cut -c1-250 gpurun_out/r2_bench_2gpu_final.json; tail -3 gpurun_out/r2_bench_2gpu_final.err
