"""Multi-GPU plumbing of the ray renderer: one process per GPU (torch.distributed, NCCL over NVLink), rays are
independent so the path shards with no data-path collective; the only exchange is ONE all-gather of the packed
per-ray results (depth + rgb = 16 B/ray) -- SURVEY.md 8(e).

Two layouts, both used by the reference's callers:
  * ray-sharded frame  (render_frame_sharded): contiguous ray range per rank of one frame; every rank ends up with
    the full (R,4) result in input order.  Strong scaling of one frame.
  * frame per GPU      (gather_frames): every rank renders its own frame/pose (novel-depth eval batch, pose sweep of
    generate_novel_depths.py:103-122); results are gathered to (world, R, 4).  Weak scaling -- what bench.py times.
The functions only need `render_fn(pixels) -> (depth (r,), color (r,3))`, so the host logic is testable on CPU with
the gloo backend and a stand-in render_fn (tests/test_dist_gloo.py); on the GPU render_fn is B200Renderer.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_range(n_rays: int, rank: int, world: int) -> Tuple[int, int, int]:
    """Contiguous, padded-equal split: returns (start, stop, per_rank) with per_rank = ceil(n/world);
    the last ranks may get fewer (or zero) real rays; all ranks exchange per_rank rows."""
    per = (n_rays + world - 1) // world
    start = min(n_rays, rank * per)
    stop = min(n_rays, start + per)
    return start, stop, per


def pack_result(depth: torch.Tensor, color: torch.Tensor, rows: int) -> torch.Tensor:
    """(r,) + (r,3) -> (rows,4) [depth, r, g, b], zero padded to `rows`."""
    r = depth.shape[0]
    out = torch.empty((rows, 4), dtype=torch.float32, device=depth.device)
    if r:
        torch.cat([depth.reshape(r, 1), color.reshape(r, 3)], dim=1, out=out[:r])
    if r < rows:
        out[r:].zero_()
    return out


def _all_gather(packed: torch.Tensor, world: int) -> torch.Tensor:
    full = torch.empty((world * packed.shape[0], packed.shape[1]), dtype=packed.dtype, device=packed.device)
    if hasattr(dist, "all_gather_into_tensor") and packed.is_cuda:
        dist.all_gather_into_tensor(full, packed)
    else:
        parts = list(full.chunk(world, dim=0))
        dist.all_gather(parts, packed)
    return full


def render_frame_sharded(render_fn: Callable, sampled_pixels: torch.Tensor):
    """Ray-range shard of one frame + one all-gather.  Returns (depth (R,), color (R,3)) on every rank.
    render_fn(pixels, ray_offset) -> (depth (r,), color (r,3)): ray_offset is the index of the shard's first ray in
    the frame (B200Renderer.render_rays_batch(ray_offset=..., seed=...) keys its Philox noise on it, so that the
    sharded frame equals the single-GPU frame bit for bit)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    R = int(sampled_pixels.shape[0])
    start, stop, per = shard_range(R, rank, world)
    depth, color = render_fn(sampled_pixels[start:stop], start)
    full = _all_gather(pack_result(depth, color, per), world)[:R]
    return full[:, 0].contiguous(), full[:, 1:].contiguous()


def gather_frames(depth: torch.Tensor, color: torch.Tensor) -> torch.Tensor:
    """Frame-per-GPU layout: every rank contributes its own (R,) depth and (R,3) colour; returns (world, R, 4)."""
    world = dist.get_world_size()
    R = int(depth.shape[0])
    return _all_gather(pack_result(depth, color, R), world).reshape(world, R, 4)
