#!/bin/bash
# A/B of an experimental library build against the product build (same box, same run): per-tile cycle accounting + rays/s
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
VAR=$1
SCENERF_B200_LIB=$PWD/scenerf_b200/libscenerf_b200_$VAR.so timeout 600 python -m pytest tests/test_gpu_preproj.py -q -m gpu -k "adversarial or kitti_s128" 2>&1 | tail -2
for rep in 1 2; do for which in base $VAR; do
if [ $which = base ]; then unset SCENERF_B200_LIB; else export SCENERF_B200_LIB=$PWD/scenerf_b200/libscenerf_b200_$VAR.so; fi
SRF_TC_PROF=1 timeout 600 python bench.py --precision fp16 --latent-table ${TAB:-1} --steps 3 --warmup 3 --no-variants --no-cpu-baseline --no-extras > gpurun_out/ab_$which.json 2> gpurun_out/ab_$which.err
python -c "
import json;d=json.loads(open('gpurun_out/ab_$which.json').read().strip().splitlines()[-1]);print('$which rep $rep: %.1f ms  %.0f rays/s'%(d['ms_per_step'],d['value']))"
grep prof gpurun_out/ab_$which.err | grep -E "CTA=(3065|6130)" | tail -1 | cut -c40-330
done; done
