#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c_gpus.txt; lscpu | head -20 >> gpurun_out/r2c_gpus.txt; free -g >> gpurun_out/r2c_gpus.txt
timeout 900 python -m pytest tests/test_gpu_nccl.py -q -m gpu -s 2>&1 | tail -15 > gpurun_out/r2c_nccl_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2c_bench_2gpu.json 2> gpurun_out/r2c_bench_2gpu.err
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_1gpu.json 2> gpurun_out/r2c_bench_1gpu.err
timeout 600 python tools/ref_probe.py > gpurun_out/r2c_ref_probe.log 2>&1
cat gpurun_out/r2c_nccl_test.log; cut -c1-300 gpurun_out/r2c_bench_2gpu.json; tail -3 gpurun_out/r2c_bench_2gpu.err; cut -c1-300 gpurun_out/r2c_bench_1gpu.json; tail -3 gpurun_out/r2c_bench_1gpu.err; cat gpurun_out/r2c_ref_probe.log
