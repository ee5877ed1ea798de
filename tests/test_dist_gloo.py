"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: ray-range sharding with padding, packing,
all-gather order, frame-per-GPU gather.  The renderer itself is replaced by a deterministic stand-in so that the
result of the sharded path can be compared with the un-sharded one bit for bit."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scenerf_b200 import dist as sdist


def _fake_render(pix):
    depth = pix[:, 0] * 0.5 + pix[:, 1] * 0.25 + 1.0
    color = torch.stack([pix[:, 0] * 0.001, pix[:, 1] * 0.002, pix[:, 0] * 0.0 + 0.5], dim=1)
    return depth, color


def _worker(rank, world, port, n_rays, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pix = torch.arange(n_rays * 2, dtype=torch.float32).reshape(n_rays, 2)
        d, c = sdist.render_frame_sharded(_fake_render, pix)
        d0, c0 = _fake_render(pix)
        ok = torch.equal(d, d0) and torch.equal(c, c0)
        fd, fc = _fake_render(pix + rank)
        frames = sdist.gather_frames(fd, fc)
        for r in range(world):
            ed, ec = _fake_render(pix + r)
            ok = ok and torch.equal(frames[r, :, 0], ed) and torch.equal(frames[r, :, 1:], ec)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,n_rays", [(2, 1000), (2, 7), (3, 10), (2, 1)])
def test_sharded_render_equals_unsharded(world, n_rays):
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), n_rays, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_shard_range_covers_everything():
    for n in (0, 1, 5, 128, 453620):
        for w in (1, 2, 3, 4, 8):
            spans = [sdist.shard_range(n, r, w) for r in range(w)]
            assert sum(b - a for a, b, _ in spans) == n
            assert all(spans[i][1] == spans[i + 1][0] or spans[i + 1][0] == n for i in range(w - 1))
            assert len({p for _, _, p in spans}) == 1 and spans[0][2] * w >= n
