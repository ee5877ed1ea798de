"""Density-lattice query: `SceneRF.predict(..., output_type="density")`
(/root/reference/scenerf/models/scenerf.py:505-547) evaluated on a regular lattice in the infer-camera frame -- the
direct-density stand-in for scene reconstruction that BASELINE.json lists (256^3 points: x in [-25.6, 25.6) step 0.2,
y in [-3.2, 3.2) step 0.025, z in [0.2, 51.4) step 0.2; one (x, y) column of 256 z-samples per "ray", viewdir (0,0,1)).

Multi-GPU: every rank evaluates a contiguous slab of z-planes and the slabs are all-gathered."""
from __future__ import annotations

from typing import Tuple

import torch

from .dist import shard_range

DEFAULT_X = (-25.6, 0.2, 256)        # start, step, count
DEFAULT_Y = (-3.2, 0.025, 256)
DEFAULT_Z = (0.2, 0.2, 256)


def axis(start: float, step: float, count: int, device) -> torch.Tensor:
    return (torch.arange(count, dtype=torch.float32, device=device) * step + start).contiguous()


def lattice_columns(x_ax: torch.Tensor, y_ax: torch.Tensor, z_ax: torch.Tensor, col_lo: int, col_hi: int) -> torch.Tensor:
    """Points of columns [col_lo, col_hi) (column index = ix * ny + iy) as (n_cols, nz, 3)."""
    ny = y_ax.numel()
    cols = torch.arange(col_lo, col_hi, device=x_ax.device)
    xs, ys = x_ax[cols // ny], y_ax[cols % ny]
    nz = z_ax.numel()
    pts = torch.empty((cols.numel(), nz, 3), dtype=torch.float32, device=x_ax.device)
    pts[:, :, 0] = xs[:, None]
    pts[:, :, 1] = ys[:, None]
    pts[:, :, 2] = z_ax[None, :]
    return pts


def density_lattice(renderer, x_rgb, cam_K, x=DEFAULT_X, y=DEFAULT_Y, z=DEFAULT_Z, cols_per_call: int = 16384,
                    rank: int = 0, world: int = 1, group=None, with_color: bool = False):
    """-> density (nx, ny, nz) float32 on the device (and colour (nx, ny, nz, 3) if asked).  With world > 1 each rank
    evaluates z-planes [z_lo, z_hi) of every column and the slabs are all-gathered (every rank returns the full lattice)."""
    dev = renderer.device
    x_ax, y_ax, z_ax = axis(*x, dev), axis(*y, dev), axis(*z, dev)
    nx, ny, nz = x_ax.numel(), y_ax.numel(), z_ax.numel()
    z_lo, z_hi, per = shard_range(nz, rank, world)
    z_loc = z_ax[z_lo:z_hi]
    n_cols = nx * ny
    dens = torch.zeros((n_cols, per), dtype=torch.float32, device=dev)
    col = torch.zeros((n_cols, per, 3), dtype=torch.float32, device=dev) if with_color else None
    launches = 0
    if z_hi > z_lo:
        viewdir_one = torch.tensor([[0.0, 0.0, 1.0]], device=dev)
        for c0 in range(0, n_cols, cols_per_call):
            c1 = min(n_cols, c0 + cols_per_call)
            pts = lattice_columns(x_ax, y_ax, z_loc, c0, c1)
            d, c = renderer.predict("mlp", pts, x_rgb, cam_K, None, viewdir_one.expand(c1 - c0, 3).contiguous(), "density")
            launches += renderer.last_launches
            dens[c0:c1, :z_hi - z_lo] = d
            if with_color:
                col[c0:c1, :z_hi - z_lo] = c
    renderer.last_lattice_launches = launches
    if world > 1:
        import torch.distributed as dist
        parts = [torch.empty_like(dens) for _ in range(world)]
        dist.all_gather(parts, dens, group=group)
        dens = torch.cat(parts, dim=1)[:, :nz]
        if with_color:
            cparts = [torch.empty_like(col) for _ in range(world)]
            dist.all_gather(cparts, col, group=group)
            col = torch.cat(cparts, dim=1)[:, :nz]
    dens = dens.reshape(nx, ny, -1)
    return (dens, col.reshape(nx, ny, -1, 3)) if with_color else dens
