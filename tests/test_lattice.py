"""Density-lattice query.  CPU: lattice generation and slab partition; GPU: densities against the
oracle's `predict` (scenerf.py:505-547 restatement)."""
import numpy as np
import pytest

from cases import PREDICT_CASES, params_for, pyramid_for
from scenerf_b200 import synth


def test_lattice_columns_and_slabs():
    import torch
    from scenerf_b200 import lattice
    from scenerf_b200.dist import shard_range
    x, y, z = (-1.0, 0.5, 5), (-0.2, 0.1, 4), (0.2, 0.2, 7)
    xa, ya, za = lattice.axis(*x, "cpu"), lattice.axis(*y, "cpu"), lattice.axis(*z, "cpu")
    pts = lattice.lattice_columns(xa, ya, za, 0, 20).numpy()
    gx, gy, gz = np.meshgrid(xa.numpy(), ya.numpy(), za.numpy(), indexing="ij")
    assert np.array_equal(pts.reshape(5, 4, 7, 3), np.stack([gx, gy, gz], -1))
    assert np.allclose(lattice.axis(*lattice.DEFAULT_Z, "cpu").numpy()[[0, -1]], [0.2, 51.2])
    assert lattice.axis(*lattice.DEFAULT_Y, "cpu").numel() == 256
    # z slabs of a 3-rank split cover every plane once
    seen = []
    for r in range(3):
        lo, hi, per = shard_range(7, r, 3)
        seen += list(range(lo, hi))
        assert hi - lo <= per
    assert seen == list(range(7))


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("fp32", 2e-4), ("fp16", 1e-2)])
def test_density_lattice_vs_oracle(prec, tol):
    import torch
    from helpers import make_renderer, torch_pyramid
    from oracle.scenerf_oracle import OracleRenderer
    from scenerf_b200 import lattice
    cfg, seed = PREDICT_CASES["predict_adversarial_kitti"]
    r = make_renderer(cfg, prec)
    x, y, z = (-6.0, 1.0, 12), (-1.5, 0.5, 6), (0.5, 1.5, 10)
    dens, col = lattice.density_lattice(r, torch_pyramid(cfg, seed), torch.from_numpy(cfg.K), x, y, z, cols_per_call=32,
                                        with_color=True)
    assert tuple(dens.shape) == (12, 6, 10) and r.last_lattice_launches > 0
    xa, ya, za = (np.arange(n, dtype=np.float32) * np.float32(st) + np.float32(s0) for s0, st, n in (x, y, z))
    gx, gy, gz = np.meshgrid(xa, ya, za, indexing="ij")
    pts = np.stack([gx, gy, gz], -1).reshape(-1, 10, 3).astype(np.float32)
    orc = OracleRenderer(cfg, *params_for(cfg))
    d_ref, c_ref = orc.predict(orc.pm, pts, pyramid_for(cfg, seed), cfg.K, np.tile(np.float32([[0, 0, 1]]), (pts.shape[0], 1)))
    d = dens.cpu().numpy().reshape(-1, 10)
    assert np.abs(d - d_ref).max() <= tol * max(1.0, np.abs(d_ref).max())
    assert np.abs(col.cpu().numpy().reshape(-1, 10, 3) - c_ref).max() <= max(tol, 2e-3)
    # slab decomposition (single process): two z-slabs computed separately equal the full query
    a = lattice.density_lattice(r, torch_pyramid(cfg, seed), torch.from_numpy(cfg.K), x, y, (0.5, 1.5, 5), cols_per_call=32)
    b = lattice.density_lattice(r, torch_pyramid(cfg, seed), torch.from_numpy(cfg.K), x, y, (0.5 + 5 * 1.5, 1.5, 5), cols_per_call=32)
    assert np.abs(torch.cat([a, b], 2).cpu().numpy() - dens.cpu().numpy()).max() <= tol * max(1.0, np.abs(d_ref).max())
