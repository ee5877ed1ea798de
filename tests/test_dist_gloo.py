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


def _fake_render(pix, ray_offset=0):
    # depends on the ray's index inside the frame like the in-kernel Philox noise does: a shard rendered with the wrong
    # offset cannot reproduce the unsharded result
    idx = torch.arange(pix.shape[0], dtype=torch.float32) + float(ray_offset)
    depth = pix[:, 0] * 0.5 + pix[:, 1] * 0.25 + 1.0 + idx * 0.125
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


class _FakeLatticeRenderer:
    """Stand-in with the two members density_lattice touches: a device and predict(..., "density")."""
    device = torch.device("cpu")
    last_launches = 1

    def predict(self, mlp, pts, x_rgb, cam_K, T_cam2velo, viewdir, output_type):
        dens = pts[..., 0] * 0.5 + pts[..., 1] * 3.0 + pts[..., 2] * 0.125
        return dens, torch.stack([dens, dens * 2, dens * 3], dim=-1)


def _lattice_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from scenerf_b200 import lattice
        x, y, z = (-1.0, 0.5, 5), (-0.2, 0.1, 4), (0.2, 0.2, 7)          # 7 z-planes over 2 or 3 ranks: ragged slabs
        r = _FakeLatticeRenderer()
        full, fcol = lattice.density_lattice(r, None, None, x, y, z, cols_per_call=6, with_color=True)
        shard, scol = lattice.density_lattice(r, None, None, x, y, z, cols_per_call=6, rank=rank, world=world, with_color=True)
        ret[rank] = bool(torch.equal(full, shard) and torch.equal(fcol, scol) and tuple(shard.shape) == (5, 4, 7))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_density_lattice_z_slabs_equal_unsharded(world):
    """z-slab sharding + all-gather + reassembly of scenerf_b200.lattice on CPU over gloo."""
    ret = mp.Manager().dict()
    mp.spawn(_lattice_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
