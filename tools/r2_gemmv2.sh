#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_backward.py -q -m gpu -x 2>&1 | tail -6
python - <<'PY'
import sys; sys.path.insert(0,'.')
from scenerf_b200 import _lib
print("tf32 watchdog flag: (via error text on failure only)")
PY
for v in v2 v1; do
if [ $v = v1 ]; then export SRF_TF32_V1=1; else unset SRF_TF32_V1; fi
timeout 300 python bench.py --workload train --train-matmul tf32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_train_tf32_$v.json 2> gpurun_out/r2_train_tf32_$v.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_train_tf32_$v.json').read().strip().splitlines()[-1]);print('$v: ms/step %.2f fwd %.2f bwd %.2f launches %d loss %.6f'%(d['ms_per_step'],d['forward_ms'],d['backward_ms'],d['gpu_launches'],d['loss']))" || tail -5 gpurun_out/r2_train_tf32_$v.err
done
unset SRF_TF32_V1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
