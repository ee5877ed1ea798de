mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_layers.py -x -q -s > gpurun_out/run5_layers.log 2>&1; echo "layers rc=$?"; tail -4 gpurun_out/run5_layers.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/run5_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/run5_pytest.log
for cg in 2 1; do
  export SRF_TC_CTA_GROUP=$cg
  for skip in 0 1; do
    timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-zero-chunks $skip > gpurun_out/run5_bench_cg${cg}_s$skip.json 2> gpurun_out/run5_bench_cg${cg}_s$skip.err; echo "cg=$cg skip=$skip rc=$?"
    python -c "
import json;d=json.load(open('gpurun_out/run5_bench_cg${cg}_s$skip.json'));print('cg=$cg skip=$skip', round(d['value']), round(d['ms_per_step'],1), round(d['roofline']['frac'],3), d['clocks'])" || tail -5 gpurun_out/run5_bench_cg${cg}_s$skip.err
  done
done
