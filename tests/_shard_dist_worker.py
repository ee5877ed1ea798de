"""torchrun worker of tests/test_gpu_nccl.py: ONE frame rendered ray-sharded over the ranks
(scenerf_b200.dist.render_frame_sharded: contiguous ray ranges, one NCCL all-gather of the packed depth+rgb) must equal
the single-GPU render of the same frame BIT FOR BIT on every rank -- rays are independent and the in-kernel Philox noise
is keyed on (seed, index of the ray in the frame).  Checked for the three precision modes, a ray count that does not
divide by the world size, and for the frame-per-GPU gather (gather_frames)."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from cases import RENDER_CASES                      # noqa: E402
from helpers import make_renderer                   # noqa: E402
from scenerf_b200 import synth                      # noqa: E402
from scenerf_b200 import dist as sdist              # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg, seed = RENDER_CASES["kitti_s128"]
    x_rgb = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_pyramid(seed, cfg.sphere_W, cfg.sphere_H).items()}
    K, T = torch.from_numpy(cfg.K), torch.from_numpy(cfg.T)
    pix = torch.from_numpy(synth.random_pixels(5, 3001, cfg.img_W, cfg.img_H)).to(dev)      # 3001: ragged shards
    ok = True
    report = []
    for prec in ("fp32tc", "fp16", "fp32"):
        r = make_renderer(cfg, prec, device=dev, rng="philox")
        fn = lambda p, off: (lambda o: (o["depth"], o["color"]))(
            r.render_rays_batch(K, T, x_rgb, sampled_pixels=p, outputs="minimal", ray_offset=off, seed=4242))
        d, c = sdist.render_frame_sharded(fn, pix)
        one = r.render_rays_batch(K, T, x_rgb, sampled_pixels=pix, outputs="minimal", seed=4242)
        same = bool(torch.equal(d, one["depth"]) and torch.equal(c, one["color"]))
        # frame-per-GPU gather: every rank contributes the frame of its own pose
        Tr = torch.from_numpy(synth.yaw_translate(cfg.yaw_deg, cfg.tz + 0.5 * rank))
        mine = r.render_rays_batch(K, Tr, x_rgb, sampled_pixels=pix, outputs="minimal", seed=99)
        frames = sdist.gather_frames(mine["depth"], mine["color"])
        for q in range(world):
            Tq = torch.from_numpy(synth.yaw_translate(cfg.yaw_deg, cfg.tz + 0.5 * q))
            o = r.render_rays_batch(K, Tq, x_rgb, sampled_pixels=pix, outputs="minimal", seed=99)
            same = same and bool(torch.equal(frames[q, :, 0], o["depth"]) and torch.equal(frames[q, :, 1:], o["color"]))
        report.append("%s:%s" % (prec, "equal" if same else "MISMATCH max|d| %.3e" % float((d - one["depth"]).abs().max())))
        ok = ok and same
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print("rank %d of %d (%s): %s" % (rank, world, torch.cuda.get_device_name(dev), " ".join(report)), flush=True)
    if rank == 0:
        print("SHARD_DIST_OK" if int(flag.item()) == 1 else "SHARD_DIST_MISMATCH", "world", world, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
