"""TEST INFRASTRUCTURE ONLY (see oracle/scenerf_oracle.py header): numpy restatement of the novel-view sweep glue.

Follows /root/reference:
  scenerf/models/utils.py:6-49                      sample_rel_poses_bf / sample_rel_poses
  scenerf/scripts/reconstruction/generate_novel_depths.py:103-152   pixel grid, render, transpose + F.interpolate
  scenerf/scripts/reconstruction/depth2tsdf.py:19-26,87-103         PNG read-back, TSDF integration of the sweep
and ATen's CPU upsample_bilinear2d (align_corners=False) for F.interpolate.  Pinned against outputs of the reference
code itself (tests/golden/sweep_kitti.npz, made by tests/golden/make_goldens.py)."""
import math

import numpy as np

f32 = np.float32


def _rot_y_pose(step, angle):
    rad = angle / 180 * math.pi
    rel = np.eye(4, dtype=f32)
    rel[2, 3] = f32(rel[2, 3] + f32(step))
    rot = np.eye(4, dtype=f32)
    rot[:3, :3] = np.array([[math.cos(rad), 0, math.sin(rad)], [0, 1, 0], [-math.sin(rad), 0, math.cos(rad)]], dtype=f32)
    out = np.zeros((4, 4), dtype=f32)
    for i in range(4):                       # fp32 matmul, k ascending (only one non-trivial product per entry)
        for j in range(4):
            acc = f32(0)
            for k in range(4):
                acc = f32(acc + f32(rot[i, k] * rel[k, j]))
            out[i, j] = acc
    return out


def _arange(end, step):
    n = int(math.ceil(end / step))            # torch.arange(start=0, end, step): size and values computed in double
    return [f32(i * step) for i in range(n)]


def sample_rel_poses(step=0.5, angle=0, max_distance=10.1):
    """utils.py:29-49 -- dict keyed (step, angle) in insertion order; angles [0, +a, -a]."""
    angles = [0] + ([angle, -angle] if angle != 0 else [])
    return {(float(s), a): _rot_y_pose(s, a) for s in _arange(max_distance, step) for a in angles}


def sample_rel_poses_bf(angle=0, max_distance=2.1, step=0.2):
    """utils.py:6-26 -- angles [0, -a, +a]."""
    angles = [0] + ([-angle, angle] if angle != 0 else [])
    return {(float(s), a): _rot_y_pose(s, a) for s in _arange(max_distance, step) for a in angles}


def pixel_grid(img_size, scale):
    """generate_novel_depths.py:103-112: x-major grid of (x, y) float pixels, stride `scale`."""
    xs = np.arange(0, img_size[0], scale, dtype=f32)
    ys = np.arange(0, img_size[1], scale, dtype=f32)
    gx, gy = np.meshgrid(xs, ys, indexing="ij")
    return np.stack([gx, gy], -1).reshape(-1, 2), gx.shape


def _taps(out_size, in_size):
    scale = f32(f32(in_size) / f32(out_size))
    dst = np.arange(out_size, dtype=f32)
    src = (scale * (dst + f32(0.5)) - f32(0.5)).astype(f32)
    src = np.maximum(src, f32(0))
    i0 = np.minimum(src.astype(np.int64), in_size - 1)
    i1 = i0 + (i0 < in_size - 1)
    w1 = np.clip((src - i0.astype(f32)).astype(f32), 0, 1).astype(f32)
    return i0, i1, (f32(1) - w1).astype(f32), w1


def upsample_bilinear(img, out_h, out_w):
    """F.interpolate(img[None,None], size=(out_h,out_w), mode='bilinear') for a (h,w) float32 image."""
    img = np.asarray(img, dtype=f32)
    y0, y1, wy0, wy1 = _taps(out_h, img.shape[0])
    x0, x1, wx0, wx1 = _taps(out_w, img.shape[1])
    r0 = (img[y0][:, x0] * wx0[None, :] + img[y0][:, x1] * wx1[None, :]).astype(f32)
    r1 = (img[y1][:, x0] * wx0[None, :] + img[y1][:, x1] * wx1[None, :]).astype(f32)
    return (r0 * wy0[:, None] + r1 * wy1[:, None]).astype(f32)


def to_images(depth_rays, color_rays, grid_shape, img_size, scale):
    """generate_novel_depths.py:125-147: rays (x-major) -> depth (H,W), colour (H,W,3) clamped to [0,1]."""
    W, H = img_size
    d = depth_rays.reshape(grid_shape).T
    c = color_rays.reshape(grid_shape + (3,)).transpose(2, 1, 0)
    if scale != 1:
        d = upsample_bilinear(d, H, W)
        c = np.stack([upsample_bilinear(c[k], H, W) for k in range(3)], 0)
    return d.astype(f32), np.clip(c, 0, 1).transpose(1, 2, 0).astype(f32)


def png_roundtrip(color01):
    """plt.imsave -> (x*255).astype(uint8); depth2tsdf.py:19-26,98: float32(u8)/255.0*255.0."""
    u8 = (np.asarray(color01, dtype=f32) * f32(255)).astype(np.uint8)
    return ((u8.astype(f32) / f32(255.0)) * f32(255.0)).astype(f32)
