#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_layers.py tests/test_gpu_parity.py tests/test_gpu_preproj.py -q -m gpu -x -k "not full_size and not fp32tc" 2>&1 | tail -3
for tab in 1 0; do for which in base epiv1; do
if [ $which = base ]; then unset SCENERF_B200_LIB; else export SCENERF_B200_LIB=$PWD/scenerf_b200/libscenerf_b200_$which.so; fi
SRF_TC_PROF=1 timeout 600 python bench.py --precision fp16 --latent-table $tab --steps 3 --warmup 3 --no-variants --no-cpu-baseline --no-extras > gpurun_out/ab_$which.json 2> gpurun_out/ab_$which.err
python -c "
import json;d=json.loads(open('gpurun_out/ab_$which.json').read().strip().splitlines()[-1]);print('table=$tab $which: %.1f ms  %.0f rays/s'%(d['ms_per_step'],d['value']))"
grep prof gpurun_out/ab_$which.err | grep -E "CTA=3065" | tail -1 | cut -c40-330
done; done
