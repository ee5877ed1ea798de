mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_layers.py -x -q -s > gpurun_out/run14_layers.log 2>&1; rc=$?; echo "layers rc=$rc"; tail -8 gpurun_out/run14_layers.log | cut -c1-300
python - <<'PY'
import sys; sys.path.insert(0,'.')
from scenerf_b200 import _lib
print("watchdog flag 0x%x" % (_lib.load().srf_debug_watchdog_flag() & 0xffffffff))
PY
if [ $rc -ne 0 ]; then SRF_TC_CTA_GROUP=1 timeout 300 python -m pytest tests/test_gpu_tc_layers.py -x -q -s 2>&1 | tail -8 | cut -c1-300; exit 0; fi
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/run14_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/run14_pytest.log
SRF_TC_CTA_GROUP=1 timeout 300 python -m pytest tests/test_gpu_tc_layers.py -x -q > gpurun_out/run14_layers_cg1.log 2>&1; echo "cg1 layers rc=$?"; tail -2 gpurun_out/run14_layers_cg1.log
export SRF_TC_PROF=1
for skip in 0 1; do
  echo "== prof cg=2 skip=$skip"
  timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --rays 60000 --skip-zero-chunks $skip 2>&1 >/dev/null | grep -E "srf tc prof|Error|error" | tail -1 | cut -c1-420
done
unset SRF_TC_PROF
for skip in 0 1; do
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-zero-chunks $skip > gpurun_out/run14_bench_s$skip.json 2> gpurun_out/run14_bench_s$skip.err; echo "bench skip=$skip rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/run14_bench_s$skip.json'));print('skip=$skip', round(d['value']), round(d['ms_per_step'],1), round(d['roofline']['frac'],3), d['clocks'])" || tail -5 gpurun_out/run14_bench_s$skip.err
done
