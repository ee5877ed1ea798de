mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_layers.py -x -q -s > gpurun_out/run18_layers.log 2>&1; rc=$?; echo "layers rc=$rc"; tail -4 gpurun_out/run18_layers.log | cut -c1-300
python - <<'PY'
import sys; sys.path.insert(0,'.')
from scenerf_b200 import _lib
print("watchdog flag 0x%x" % (_lib.load().srf_debug_watchdog_flag() & 0xffffffff))
PY
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/run18_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/run18_pytest.log
export SRF_TC_PROF=1
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --rays 60000 2>&1 >/dev/null | grep -E "srf tc prof|Error|error" | grep "CTA=406" | tail -2 | cut -c1-420
unset SRF_TC_PROF
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/run18_bench.json 2> gpurun_out/run18_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/run18_bench.json'));print(round(d['value']), round(d['ms_per_step'],1), round(d['roofline']['frac'],3), d['clocks'], {k:round(v.get('ms_per_step',0),1) for k,v in d['variants'].items() if isinstance(v,dict)})" || tail -5 gpurun_out/run18_bench.err
