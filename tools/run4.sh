mkdir -p gpurun_out
for cs in 2 4 1; do
  export SRF_TC_CLUSTER=$cs
  timeout 300 python -m pytest tests/test_gpu_tc_layers.py -x -q > gpurun_out/run4_layers_cs$cs.log 2>&1; echo "cs=$cs layers rc=$?"; tail -2 gpurun_out/run4_layers_cs$cs.log
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/run4_bench_cs$cs.json 2> gpurun_out/run4_bench_cs$cs.err; echo "cs=$cs bench rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/run4_bench_cs$cs.json'));print('cs=$cs', d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])" || tail -5 gpurun_out/run4_bench_cs$cs.err
done
