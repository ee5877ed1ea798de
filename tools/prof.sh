# usage: bash tools/prof.sh <tag>  -- layer tests, per-tile cycle accounting (dense + skip), full bench
tag=$1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_layers.py -x -q -s > gpurun_out/${tag}_layers.log 2>&1; rc=$?; echo "layers rc=$rc"; tail -3 gpurun_out/${tag}_layers.log | cut -c1-300
python - <<'PY'
import sys; sys.path.insert(0,'.')
from scenerf_b200 import _lib
print("watchdog flag 0x%x" % (_lib.load().srf_debug_watchdog_flag() & 0xffffffff))
PY
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest.log
export SRF_TC_PROF=1
for skip in 0 1; do
  echo "== prof skip=$skip"
  timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-variants --rays 60000 --skip-zero-chunks $skip 2>&1 >/dev/null | grep -E "srf tc prof|Error|error" | grep "CTA=406" | tail -1 | cut -c1-420
done
unset SRF_TC_PROF
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/${tag}_bench.json'));print(round(d['value']), round(d['ms_per_step'],1), round(d['roofline']['frac'],3), d['clocks'], {k:round(v.get('ms_per_step',0),1) for k,v in d['variants'].items() if isinstance(v,dict)})" || tail -5 gpurun_out/${tag}_bench.err
