"""Probe (run on the GPU box's host): which process/thread/chunk layout gives the reference's CPU renderer its best and most
stable rays/s on this host?  Prints one line per configuration; bench.py's CPU arm uses the winner as a FIXED layout."""
import os
import sys
import time

for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
    os.environ.pop(k, None)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def _worker(args):
    threads, n, reps, seed = args
    import torch
    torch.set_num_threads(threads)
    import bench
    from oracle import ref_runner
    cfg, pix, _ = bench.workload("B")
    tm = ref_runner.ReferenceTimer(cfg, pix, bench.make_cpu_pyramid(cfg), threads=threads, seed=seed)
    tm.step(n)
    t0 = time.perf_counter()
    ts = [tm.step(n) for _ in range(reps)]
    return n * reps, time.perf_counter() - t0, ts


def main():
    import multiprocessing as mp
    import torch
    cores = os.cpu_count()
    print("cpu_count", cores, "torch default threads", torch.get_num_threads(), flush=True)
    ctx = mp.get_context("spawn")
    for procs, threads, n, reps in ((1, min(cores, 128), 256, 4), (1, min(cores, 128), 1024, 2), (1, min(cores, 64), 1024, 2),
                                    (4, max(1, cores // 4), 256, 4), (8, max(1, cores // 8), 256, 4), (16, max(1, cores // 16), 256, 3),
                                    (8, max(1, cores // 8), 1024, 2)):
        t0 = time.perf_counter()
        with ctx.Pool(procs) as pool:
            res = pool.map(_worker, [(threads, n, reps, i) for i in range(procs)])
        rays = sum(r[0] for r in res)
        span = max(r[1] for r in res)
        allt = [t for r in res for t in r[2]]
        print("procs %2d x threads %3d, %4d rays/call: %8.1f rays/s  (call times min %.2f max %.2f s; wall incl. setup %.1f s)"
              % (procs, threads, n, rays / span, min(allt), max(allt), time.perf_counter() - t0), flush=True)


if __name__ == "__main__":
    main()
