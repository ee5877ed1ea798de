"""torchrun worker of tests/test_sweep.py::test_sweep_pose_sharded_two_gpus: the poses of the golden sweep are sharded
over the ranks (NCCL all-gather of the volumes + srf_tsdf_merge) and must equal the single-GPU sequential result on
every rank: distances and weights bit for bit, colours up to float32-exact distance ties (see below)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from cases import load_golden                       # noqa: E402
from test_sweep import sweep_cfg, N_POSES           # noqa: E402
from helpers import make_renderer                   # noqa: E402
from scenerf_b200 import synth, sweep               # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    g = load_golden("sweep_kitti")
    cfg, scale, pose_kw, pyr_seed = sweep_cfg()
    r = make_renderer(cfg, "fp32", device=dev)
    x_rgb = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_pyramid(pyr_seed, cfg.sphere_W, cfg.sphere_H).items()}
    sw = sweep.NovelDepthSweep(r, torch.from_numpy(cfg.K).to(dev), x_rgb, img_size=(244, 74), scale=scale)
    poses = sweep.sample_rel_poses(**pose_kw)
    noises = [(torch.from_numpy(g["noise_u%d" % i]), torch.from_numpy(g["noise_n%d" % i])) for i in range(N_POSES)]
    sharded = sw.reconstruct(poses, g["T_velo2cam"], g["vol_bnds"], noises=noises, rank=rank, world=world)
    single = sw.reconstruct(poses, g["T_velo2cam"], g["vol_bnds"], noises=noises)
    # distances and weights: bit-exact.  Colour: the sequential fold compares the stored float32 distance with the new
    # float64 one, so two observations whose distances round to the SAME float32 may resolve either way; after the
    # merge the later shard wins such a tie.  Both are "an observation of minimal |distance|"; allow <= 1e-4 of voxels.
    col_diff = float((sharded.get_volume()[1] != single.get_volume()[1]).mean())
    ok = (np.array_equal(sharded.get_volume()[0], single.get_volume()[0])
          and np.array_equal(sharded.get_weight(), single.get_weight()) and col_diff <= 1e-4)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SWEEP_DIST_OK" if int(flag.item()) == 1 else "SWEEP_DIST_MISMATCH", "world", world)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
