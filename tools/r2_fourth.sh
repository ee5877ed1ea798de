#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_preproj.py -q -m gpu -s -x 2>&1 | grep -E "max-abs-err|RaySOM|Error|error|assert|FAILED|passed|failed|watchdog" | tail -60 > gpurun_out/r2d_preproj_tests.log
SRF_TC_PROF=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
tail -40 gpurun_out/r2d_preproj_tests.log; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2d_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"]); print(json.dumps(d["variants"], indent=0)[:3000])
PY
grep "prof" gpurun_out/r2d_bench.err | tail -12
