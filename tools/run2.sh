mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_layers.py -x -q -s > gpurun_out/run2_layers.log 2>&1; echo "rc=$?" >> gpurun_out/run2_layers.log
tail -40 gpurun_out/run2_layers.log
