#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_preproj.py tests/test_gpu_fp32tc.py tests/test_gpu_tc_layers.py tests/test_backward.py -q -m gpu -x -k "not full_size" 2>&1 | tail -4 > gpurun_out/r2i_tests.log
for cfgs in "fp16 0" "fp16 1" "fp32tc 0" "fp32tc 1"; do set -- $cfgs
SRF_TC_PROF=1 timeout 600 python bench.py --precision $1 --latent-table $2 --steps 3 --warmup 3 --no-variants --no-cpu-baseline --no-extras > gpurun_out/r2i_bench_$1_$2.json 2> gpurun_out/r2i_bench_$1_$2.err
done
cat gpurun_out/r2i_tests.log
for cfgs in "fp16 0" "fp16 1" "fp32tc 0" "fp32tc 1"; do set -- $cfgs; echo "== $1 table=$2"; cut -c1-150 gpurun_out/r2i_bench_$1_$2.json; grep prof gpurun_out/r2i_bench_$1_$2.err | sort | uniq -c | sort -rn | head -1 | cut -c1-330; done
echo "== training step"
for ch in 98304 9472; do for mm in tf32; do
SRF_TRAIN_CHUNK=$ch timeout 300 python bench.py --workload train --train-matmul $mm --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_train_${mm}_$ch.json 2> gpurun_out/r2i_train_${mm}_$ch.err
python -c "
import json;d=json.loads(open('gpurun_out/r2i_train_${mm}_$ch.json').read().strip().splitlines()[-1]);print('chunk $ch $mm: ms/step %.2f fwd %.2f bwd %.2f launches %d loss %.6f'%(d['ms_per_step'],d['forward_ms'],d['backward_ms'],d['gpu_launches'],d['loss']))" || tail -3 gpurun_out/r2i_train_${mm}_$ch.err
done; done
