mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/run3_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/run3_pytest.log
tail -5 gpurun_out/run3_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/run3_bench.json 2> gpurun_out/run3_bench.err; echo "bench rc=$?"
cat gpurun_out/run3_bench.json; tail -5 gpurun_out/run3_bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --skip-zero-chunks 1 --no-cpu-baseline > gpurun_out/run3_bench_skip.json 2> gpurun_out/run3_bench_skip.err; echo "bench rc=$?"
cat gpurun_out/run3_bench_skip.json; tail -5 gpurun_out/run3_bench_skip.err
