#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
(time timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -15) > gpurun_out/r2h_all_gpu_tests.log 2>&1
timeout 300 python -m pytest tests/test_decoder.py -q -m gpu -s 2>&1 | grep -E "^1_|passed|failed" > gpurun_out/r2h_decoder.log
for cfgs in "fp16 0" "fp16 1" "fp32tc 0" "fp32tc 1"; do set -- $cfgs
SRF_TC_PROF=1 timeout 600 python bench.py --precision $1 --latent-table $2 --steps 3 --warmup 3 --no-variants --no-cpu-baseline --no-extras > gpurun_out/r2h_bench_$1_$2.json 2> gpurun_out/r2h_bench_$1_$2.err
done
cat gpurun_out/r2h_all_gpu_tests.log gpurun_out/r2h_decoder.log
for cfgs in "fp16 0" "fp16 1" "fp32tc 0" "fp32tc 1"; do set -- $cfgs; echo "== $1 table=$2"; cut -c1-150 gpurun_out/r2h_bench_$1_$2.json; grep prof gpurun_out/r2h_bench_$1_$2.err | sort | uniq -c | sort -rn | head -1 | cut -c1-330; done
echo "== training step: whole-call chunk (default) vs round-1 chunk"
timeout 600 python -m pytest tests/test_backward.py -q -m gpu 2>&1 | tail -3
for ch in 98304 9472; do for mm in tf32 fp32; do
SRF_TRAIN_CHUNK=$ch timeout 300 python bench.py --workload train --train-matmul $mm --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_train_${mm}_$ch.json 2> gpurun_out/r2h_train_${mm}_$ch.err
python -c "
import json;d=json.loads(open('gpurun_out/r2h_train_${mm}_$ch.json').read().strip().splitlines()[-1]);print('chunk $ch $mm: ms/step %.2f fwd %.2f bwd %.2f launches %d loss %.6f'%(d['ms_per_step'],d['forward_ms'],d['backward_ms'],d['gpu_launches'],d['loss']))" || tail -3 gpurun_out/r2h_train_${mm}_$ch.err
done; done
