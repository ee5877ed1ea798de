mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/run15_bench_2gpu.json 2> gpurun_out/run15_bench_2gpu.err; echo "2gpu rc=$?"
cat gpurun_out/run15_bench_2gpu.json | cut -c1-1500; tail -5 gpurun_out/run15_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/run15_ref.json 2> gpurun_out/run15_ref.err; echo "ref rc=$?"; cat gpurun_out/run15_ref.json | cut -c1-900; tail -3 gpurun_out/run15_ref.err
