#!/bin/bash
# A/B of the lean split-mode (fp32tc) epilogue against the first version (variant library built with -DSRF_EPI_V1)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fp32tc.py tests/test_gpu_tc_layers.py tests/test_gpu_parity.py tests/test_gpu_preproj.py -q -m gpu -x -k "not full_size" 2>&1 | tail -3
for tab in 1 0; do for which in base epiv1; do
if [ $which = base ]; then unset SCENERF_B200_LIB; else export SCENERF_B200_LIB=$PWD/scenerf_b200/libscenerf_b200_$which.so; fi
timeout 600 python bench.py --precision fp32tc --latent-table $tab --steps 3 --warmup 3 --no-variants --no-cpu-baseline --no-extras > gpurun_out/ab3_${which}_$tab.json 2> gpurun_out/ab3_${which}_$tab.err
python -c "
import json;d=json.loads(open('gpurun_out/ab3_${which}_$tab.json').read().strip().splitlines()[-1]);print('table=$tab $which: %.1f ms  %.0f rays/s'%(d['ms_per_step'],d['value']))"
grep prof gpurun_out/ab3_${which}_$tab.err | grep -E "CTA=3065" | tail -1 | cut -c40-330
done; done
unset SCENERF_B200_LIB
for tab in 1 0; do
timeout 600 python bench.py --precision fp16 --latent-table $tab --steps 3 --warmup 3 --no-variants --no-cpu-baseline --no-extras > gpurun_out/ab3_fp16_$tab.json 2> gpurun_out/ab3_fp16_$tab.err
python -c "
import json;d=json.loads(open('gpurun_out/ab3_fp16_$tab.json').read().strip().splitlines()[-1]);print('fp16 table=$tab: %.1f ms  %.0f rays/s'%(d['ms_per_step'],d['value']))"
done
