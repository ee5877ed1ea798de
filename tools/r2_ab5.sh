#!/bin/bash
# A/B of the code diet (out-of-line spin loops, rolled positional-encoding loop) against the previous library
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fp32tc.py tests/test_gpu_tc_layers.py tests/test_gpu_parity.py tests/test_gpu_preproj.py -q -m gpu -x -k "not full_size" 2>&1 | tail -3
for prec in fp16 fp32tc; do for tab in 1 0; do for which in base prev; do
if [ $which = base ]; then unset SCENERF_B200_LIB; else export SCENERF_B200_LIB=$PWD/scenerf_b200/libscenerf_b200_$which.so; fi
timeout 600 python bench.py --precision $prec --latent-table $tab --steps 3 --warmup 3 --no-variants --no-cpu-baseline --no-extras > gpurun_out/ab5_${prec}_${which}_$tab.json 2> gpurun_out/ab5_${prec}_${which}_$tab.err
python -c "
import json;d=json.loads(open('gpurun_out/ab5_${prec}_${which}_$tab.json').read().strip().splitlines()[-1]);print('$prec table=$tab $which: %.1f ms  %.0f rays/s'%(d['ms_per_step'],d['value']))"
done; done; done
unset SCENERF_B200_LIB
for tab in 1 0; do
SRF_TC_PROF=1 timeout 600 python bench.py --precision fp16 --latent-table $tab --steps 2 --warmup 3 --no-variants --no-cpu-baseline --no-extras > gpurun_out/ab5_prof_$tab.json 2> gpurun_out/ab5_prof_$tab.err
grep prof gpurun_out/ab5_prof_$tab.err | grep -E "CTA=3065" | tail -1 | cut -c40-330
done
