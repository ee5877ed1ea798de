#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_sweep.py -q -m gpu 2>&1 | tail -3
for prec in fp16 fp32tc; do
timeout 600 python bench.py --workload sweep --precision $prec --sweep-table 1 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_sweep_${prec}_table.json 2> gpurun_out/r2_sweep_${prec}_table.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_sweep_${prec}_table.json').read().strip().splitlines()[-1]);print('$prec table: ms/frame %.1f rays/s %.0f launches %d'%(d['ms_per_frame'],d['rays_per_sec'],d['gpu_launches']))" || tail -4 gpurun_out/r2_sweep_${prec}_table.err
done
timeout 600 python bench.py --workload sweep --precision fp16 --sweep-table 0 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_sweep_fp16_dense.json 2> gpurun_out/r2_sweep_fp16_dense.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_sweep_fp16_dense.json').read().strip().splitlines()[-1]);print('fp16 dense: ms/frame %.1f rays/s %.0f'%(d['ms_per_frame'],d['rays_per_sec']))"
