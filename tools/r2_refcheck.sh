#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "cgroup cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>/dev/null) | cfs_quota: $(cat /sys/fs/cgroup/cpu/cpu.cfs_quota_us 2>/dev/null) | nproc $(nproc) | loadavg $(cat /proc/loadavg)"
python -c "
import sys; sys.path.insert(0,'.'); import bench; print(bench.effective_cpus())"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "extremes or stale" 2>&1 | tail -4
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_b.json 2> gpurun_out/r2_bench_reference_b.err; cut -c1-200 gpurun_out/r2_bench_reference_b.json
python -c "
import json;d=json.loads(open('gpurun_out/r2_bench_reference_b.json').read().strip().splitlines()[-1]);print(d['cpu_baseline'])"
echo "loadavg after: $(cat /proc/loadavg)"; top -b -n 1 | head -15
