"""CPU oracle for the SceneRF ray-render hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy/float32 restatement of the reference algorithm behind `SceneRF.render_rays_batch`
(/root/reference/scenerf/models/scenerf.py:392-748 and its helpers).  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import this
module; the product path (scenerf_b200/) never does and fails loudly without its CUDA library.

Pinning: the reference ships no tests or golden vectors.  This oracle is pinned against
  (1) the reference's only known-answer check -- scripts/determine_angles.py <-> the FOV constants hard-coded
      at scenerf.py:84-87 and scenerf_bf.py:84-87 -- and
  (2) outputs of the reference itself, executed in the build container by tests/golden/make_goldens.py on the
      deterministic inputs of scenerf_b200.synth and committed as tests/golden/*.npz
(tests/test_oracle.py).  Every function cites the reference lines it follows.

All arithmetic is float32 with the reference's operation order wherever the order is visible in the source
(linspace halves, non-fused x*f+phase in the positional encoding, the vectorised-CPU grid_sample un-normalise
`(g+1)*(size/2)-0.5`, ...).  Transcendentals (acos/atan2/sin/exp) come from numpy's libm and can differ from
torch's by an ulp; the consequence -- a sphere pixel that rounds the other way for ~1e-5 of the points -- is
measured by the tests rather than hidden.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32

SCALES = (1, 2, 4, 8, 16)
SCALE_KEYS = tuple("1_%d" % s for s in SCALES)


# ----------------------------------------------------------------------------------------------------------------
# small helpers reproducing torch semantics
# ----------------------------------------------------------------------------------------------------------------
def torch_linspace(start: float, end: float, steps: int) -> np.ndarray:
    """torch.linspace in float32: ATen fills the lower half as start+step*i and the upper half as
    end-step*(steps-1-i) (used at utils.py:78-80 and scenerf.py:556-560)."""
    start, end = f32(start), f32(end)
    if steps == 1:
        return np.array([start], dtype=f32)
    step = f32((end - start) / f32(steps - 1))
    i = np.arange(steps)
    lo = (start + step * i.astype(f32)).astype(f32)
    hi = (end - step * (steps - 1 - i).astype(f32)).astype(f32)
    return np.where(i < steps // 2, lo, hi).astype(f32)


def _mm3(M: np.ndarray, v: np.ndarray) -> np.ndarray:
    """(3x3) @ (n,3) rows -> (n,3), float32, left-to-right accumulation."""
    M = M.astype(f32)
    out = np.empty_like(v, dtype=f32)
    for r in range(3):
        out[:, r] = (M[r, 0] * v[:, 0] + M[r, 1] * v[:, 1]) + M[r, 2] * v[:, 2]
    return out


def cam_pts_2_cam_pts(p: np.ndarray, T: np.ndarray) -> np.ndarray:
    """utils.py:272-282 -- rigid transform of (n,3) points by the 4x4 T (homogeneous 1 appended)."""
    T = T.astype(f32)
    out = np.empty_like(p, dtype=f32)
    for r in range(3):
        out[:, r] = ((T[r, 0] * p[:, 0] + T[r, 1] * p[:, 1]) + T[r, 2] * p[:, 2]) + T[r, 3]
    return out


def compute_direction_from_pixels(pix: np.ndarray, inv_K: np.ndarray):
    """utils.py:177-182.  Returns (unit_direction, un-normalised direction)."""
    homo = np.concatenate([pix.astype(f32), np.ones((pix.shape[0], 1), f32)], axis=1)
    d = _mm3(inv_K[:3, :3], homo)
    n = np.sqrt((d * d).sum(axis=1, dtype=f32)).astype(f32)
    unit = d / np.maximum(n, f32(1e-12))[:, None]
    return unit.astype(f32), d


# ----------------------------------------------------------------------------------------------------------------
# sampling
# ----------------------------------------------------------------------------------------------------------------
def sample_rays_viewdir(inv_K, T, pix, n_pts, max_sample_depth, noise_u):
    """utils.py:112-173 with sampling_method="uniform" -> utils.py:75-90.
    noise_u: (R, n_pts) the U[0,1) draw of torch.rand_like (utils.py:84)."""
    unit, viewdir = compute_direction_from_pixels(pix, inv_K)
    d_min, d_max = 0.2, max_sample_depth
    step = f32((d_max - d_min) / n_pts)
    lin = torch_linspace(d_min, d_max, n_pts)
    t = (lin[None, :] + noise_u.astype(f32) * step).astype(f32)            # sensor distance (R,U)
    cam = (t[:, :, None] * unit[:, None, :]).astype(f32)                   # source frame
    depth = cam[:, :, 2].copy()
    pts = cam_pts_2_cam_pts(cam.reshape(-1, 3), T).reshape(cam.shape)
    viewdir_infer = _mm3(T[:3, :3], viewdir)                               # NOT normalised (utils.py:135,170)
    return pts, depth, t, viewdir_infer, unit


def sample_rays_gaussian(T, unit, means, stds, n_pts_per_gaussian, noise_n):
    """utils.py:186-229.  noise_n: (R, G*P) the N(0,1) draw of torch.normal (utils.py:208-211)."""
    t = np.repeat(means, n_pts_per_gaussian, axis=1).astype(f32)
    s = np.repeat(stds, n_pts_per_gaussian, axis=1).astype(f32)
    t = (t + noise_n.astype(f32) * s).astype(f32)
    t[t < f32(0.1)] = f32(0.1)
    cam = (t[:, :, None] * unit[:, None, :]).astype(f32)
    depth = cam[:, :, 2].copy()
    pts = cam_pts_2_cam_pts(cam.reshape(-1, 3), T).reshape(cam.shape)
    return pts, depth, t


# ----------------------------------------------------------------------------------------------------------------
# projection, spherical mapping, positional encoding, feature gather
# ----------------------------------------------------------------------------------------------------------------
def cam_pts_2_pix(p: np.ndarray, K: np.ndarray) -> np.ndarray:
    """utils.py:298-315: K @ p, divide where z>0, else the (-1,-1) sentinel."""
    h = _mm3(K, p)
    mask = h[:, 2] > 0
    pix = np.full((p.shape[0], 2), -1.0, dtype=f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        pix[mask] = (h[mask, :2] / h[mask, 2:3]).astype(f32)
    return pix


def sphere_coords_from_pixels(pix, inv_K, angles, sphere_W, sphere_H):
    """spherical_mapping.py:80-93 -> :8-18 -> :104-115 -> :95-102.  Returns int64 (n,2) rounded coords and the
    un-rounded float coords."""
    v_min, v_max, h_min, h_max = angles
    h_fov, v_fov = abs(h_max - h_min), abs(v_max - v_min)
    homo = np.concatenate([pix.astype(f32), np.ones((pix.shape[0], 1), f32)], axis=1)
    c = _mm3(inv_K, homo)                                                   # depth == 1
    x, y, z = c[:, 0], c[:, 1], c[:, 2]
    dist = np.sqrt(((x * x + y * y) + z * z).astype(f32)).astype(f32)
    with np.errstate(invalid="ignore", divide="ignore"):
        v = (np.arccos((-y / dist).astype(f32)).astype(f32) / f32(math.pi) * f32(180)).astype(f32)
    hh = (f32(180) - (np.arctan2(z, x).astype(f32) / f32(math.pi) * f32(180)).astype(f32)).astype(f32)
    proj_x = ((hh - f32(h_min)) / f32(h_fov)).astype(f32)
    proj_y = ((v - f32(v_min)) / f32(v_fov)).astype(f32)
    fx = (proj_x * f32(sphere_W - 1)).astype(f32)
    fy = (proj_y * f32(sphere_H - 1)).astype(f32)
    fl = np.stack([fx, fy], axis=1)
    return np.rint(fl).astype(np.int64), fl                                 # torch.round == half-to-even


_PI_F32 = f32(np.pi)


def positional_encoding(p: np.ndarray, num_freqs: int = 6) -> np.ndarray:
    """pe.py:13-43: [x, sin(x*f_k + phase)] with phases 0, pi/2 interleaved; x*f and +phase are separate
    float32 roundings (torch.addcmul is not fused on the CPU path).  Output (n, 3 + 3*2*num_freqs)."""
    freqs = (f32(np.pi) * (f32(2.0) ** np.arange(num_freqs, dtype=f32))).astype(f32)
    freqs = np.repeat(freqs, 2)                                             # f1 f1 f2 f2 ...
    phases = np.zeros(2 * num_freqs, dtype=f32)
    phases[1::2] = f32(np.pi * 0.5)
    arg = (phases[None, :, None] + (p[:, None, :].astype(f32) * freqs[None, :, None]).astype(f32)).astype(f32)
    emb = np.sin(arg).astype(f32).reshape(p.shape[0], -1)
    return np.concatenate([p.astype(f32), emb], axis=1)


def sample_feats_2d(fmap: np.ndarray, coords: np.ndarray, norm_size) -> np.ndarray:
    """utils.py:232-247: integer coords / norm_size * 2 - 1 -> F.grid_sample(bilinear, zeros,
    align_corners=False).  fmap (C,H,W) float32; coords (n,2) int64; returns (n,C).
    Un-normalise and interpolation follow ATen's vectorised CPU kernel: ix=(g+1)*(W/2)-0.5, w=ix-floor(ix),
    e=1-w, out = nw*v_nw + ne*v_ne + sw*v_sw + se*v_se with zero for out-of-range taps."""
    C, H, W = fmap.shape
    gx = ((coords[:, 0].astype(f32) / f32(norm_size[0])).astype(f32) * f32(2) - f32(1)).astype(f32)
    gy = ((coords[:, 1].astype(f32) / f32(norm_size[1])).astype(f32) * f32(2) - f32(1)).astype(f32)
    ix = ((gx + f32(1)) * f32(W / 2.0) - f32(0.5)).astype(f32)
    iy = ((gy + f32(1)) * f32(H / 2.0) - f32(0.5)).astype(f32)
    x_w, y_n = np.floor(ix), np.floor(iy)
    w = (ix - x_w).astype(f32)
    e = (f32(1) - w).astype(f32)
    n = (iy - y_n).astype(f32)
    s = (f32(1) - n).astype(f32)
    x0, y0 = x_w.astype(np.int64), y_n.astype(np.int64)
    out = np.zeros((coords.shape[0], C), dtype=f32)
    flat = fmap.reshape(C, H * W)
    for dx, dy, wt in ((0, 0, s * e), (1, 0, s * w), (0, 1, n * e), (1, 1, n * w)):
        xx, yy = x0 + dx, y0 + dy
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        idx = np.where(ok, yy * W + xx, 0)
        val = flat[:, idx].T * ok[:, None].astype(f32)
        out = (out + val * wt.astype(f32)[:, None]).astype(f32)
    return out


def gather_latent(x_rgb: dict, coords: np.ndarray, sphere_W: int, sphere_H: int) -> np.ndarray:
    """scenerf.py:522-527: scale 1 normalised by (W,H); scale s by (W//s, H//s) with FULL-res coords (quirk Q2)."""
    feats = [sample_feats_2d(x_rgb["1_1"], coords, (sphere_W, sphere_H))]
    for s in (2, 4, 8, 16):
        feats.append(sample_feats_2d(x_rgb["1_%d" % s], coords, (sphere_W // s, sphere_H // s)))
    return np.concatenate(feats, axis=1)


# ----------------------------------------------------------------------------------------------------------------
# MLP
# ----------------------------------------------------------------------------------------------------------------
def resnetfc(params: dict, z: np.ndarray, x: np.ndarray, n_blocks: int = 3) -> np.ndarray:
    """resnetfc.py:133-164 (+ block :54-63): h=lin_in(x); per block h+=lin_z(z); h=h+fc_1(relu(fc_0(relu(h))));
    out=lin_out(relu(h))."""
    h = x @ params["lin_in.weight"].T + params["lin_in.bias"]
    for b in range(n_blocks):
        h = h + (z @ params["lin_z.%d.weight" % b].T + params["lin_z.%d.bias" % b])
        net = np.maximum(h, 0) @ params["blocks.%d.fc_0.weight" % b].T + params["blocks.%d.fc_0.bias" % b]
        dx = np.maximum(net, 0) @ params["blocks.%d.fc_1.weight" % b].T + params["blocks.%d.fc_1.bias" % b]
        h = h + dx
    return (np.maximum(h, 0) @ params["lin_out.weight"].T + params["lin_out.bias"]).astype(f32)


def sigmoid(x):
    return (f32(1) / (f32(1) + np.exp(-x).astype(f32))).astype(f32)


def softplus(x, threshold=20.0):
    """nn.Softplus(beta=1): x if x>threshold else log1p(exp(x))  (scenerf.py:473-481)."""
    with np.errstate(over="ignore"):
        return np.where(x > threshold, x, np.log1p(np.exp(np.minimum(x, threshold))).astype(f32)).astype(f32)


# ----------------------------------------------------------------------------------------------------------------
# the renderer
# ----------------------------------------------------------------------------------------------------------------
class OracleRenderer:
    """Mirrors the part of SceneRF / SceneRF(bf) that the hot path reads: hyper-parameters, two ResnetFC
    parameter dicts and the method names predict / render_rays_batch (scenerf.py:392,505)."""

    def __init__(self, cfg, params_main: dict, params_gauss: dict):
        self.cfg = cfg
        self.pm = {k: np.ascontiguousarray(v, dtype=f32) for k, v in params_main.items()}
        self.pg = {k: np.ascontiguousarray(v, dtype=f32) for k, v in params_gauss.items()}
        self.add_const = f32(1.5 if cfg.dataset == "kitti" else 0.5)   # scenerf.py:592,594 / scenerf_bf.py:606,608
        self.debug = {}

    # -- scenerf.py:505-547 -------------------------------------------------------------------------------------
    def predict(self, params, cam_pts, x_rgb, K, viewdir, output_type="density", keep=None):
        cfg = self.cfg
        shp = cam_pts.shape
        p = cam_pts.reshape(-1, 3).astype(f32)
        pix = cam_pts_2_pix(p, K)
        inv_K = np.linalg.inv(K.astype(f32)).astype(f32) if not hasattr(self, "_inv_K") else self._inv_K
        coords, _ = sphere_coords_from_pixels(pix, inv_K, cfg.angles(), cfg.sphere_W, cfg.sphere_H)
        if keep is not None:
            self.debug[keep + "_sphere"] = coords
        pe = positional_encoding(p)
        z = gather_latent(x_rgb, coords, cfg.sphere_W, cfg.sphere_H)
        vd = np.repeat(viewdir.astype(f32), shp[1], axis=0)
        x = np.concatenate([pe, vd], axis=1)
        out = resnetfc(params, z, x)
        if output_type == "density":
            color = sigmoid(out[:, :3]).reshape(shp[0], shp[1], 3)
            density = softplus(out[:, 3] - f32(1)).reshape(shp[0], shp[1])
            return density, color
        return out.reshape(shp[0], shp[1], -1)

    # -- scenerf.py:549-596 -------------------------------------------------------------------------------------
    def predict_gaussian_means_and_stds(self, T, unit, x_rgb, K, viewdir):
        cfg = self.cfg
        G = cfg.n_gaussians
        step = cfg.max_sample_depth * 1.0 / cfg.n_gaussians
        m0 = torch_linspace(step / 2, cfg.max_sample_depth - step / 2, G)
        pts = (m0[None, :, None] * unit[:, None, :]).astype(f32)
        pts_infer = cam_pts_2_cam_pts(pts.reshape(-1, 3), T).reshape(pts.shape)
        self.debug["gauss_pts"] = pts_infer
        out = self.predict(self.pg, pts_infer, x_rgb, K, viewdir, output_type="offset", keep="gauss")
        self.debug["gauss_offset"] = out
        means = np.maximum((m0[None, :] + out[:, :, 0]).astype(f32), 0) + self.add_const
        stds = np.maximum((out[:, :, 1] + f32(cfg.std)).astype(f32), 0) + self.add_const
        return means.astype(f32), stds.astype(f32)

    # -- scenerf.py:704-748 -------------------------------------------------------------------------------------
    @staticmethod
    def render_depth_and_color(density, dist, depth_volume, colors):
        dist = np.where(dist < 0, f32(0), dist).astype(f32)
        deltas = np.empty_like(dist)
        deltas[:, 0] = dist[:, 0]
        deltas[:, 1:] = dist[:, 1:] - dist[:, :-1]
        alphas = (f32(1) - np.exp(-deltas * density).astype(f32)).astype(f32)
        shifted = np.concatenate([np.ones_like(alphas[:, :1]), (f32(1) - alphas) + f32(1e-10)], axis=1).astype(f32)
        T_alphas = np.cumprod(shifted, axis=1, dtype=f32)
        weights = (alphas * T_alphas[:, :-1]).astype(f32)
        depth = (weights * depth_volume).sum(axis=1, dtype=f32)
        color = (weights[:, :, None] * colors).sum(axis=1, dtype=f32)
        absdiff = np.abs(depth[:, None] - depth_volume)
        idx = absdiff.argmin(axis=1)
        srt = np.sort(absdiff, axis=1)
        argmin_margin = srt[:, 1] - srt[:, 0]
        closest = np.take_along_axis(absdiff, idx[:, None], 1)[:, 0]
        w_at = np.take_along_axis(weights, idx[:, None], 1)[:, 0]
        return dict(depth=depth, color=color, alphas=alphas, weights=weights, closest_pts_to_depth=closest,
                    weights_at_depth=w_at, argmin_margin=argmin_margin)

    # -- ray_som_kl.py:10-92 ------------------------------------------------------------------------------------
    def ray_som(self, means, stds, dist, alphas):
        sig = self.cfg.som_sigma
        R, G = means.shape
        distances = np.abs(means[:, None, :] - dist[:, :, None]).astype(f32)            # R,S,G
        rel = np.exp(-((means[:, :, None] - means[:, None, :]) ** 2) / f32(2 * sig ** 2)).astype(f32)  # [r,c2,c1]
        p_c1_c2 = (rel / rel.sum(axis=2, keepdims=True)).astype(f32)
        var = (stds ** 2).astype(f32)
        p_z_c1 = (np.exp(-distances ** 2 / (f32(2) * var[:, None, :])) /
                  (f32(math.sqrt(2 * math.pi)) * stds[:, None, :]) + f32(1e-5)).astype(f32)
        dens = (alphas + f32(1e-8)).astype(f32)
        p_z_c1 = (p_z_c1 * dens[:, :, None] + f32(1e-8)).astype(f32)
        temp = (p_z_c1[:, :, None, :] * p_c1_c2[:, None, :, :] + f32(1e-8)).astype(f32)
        p_z_c2 = temp.sum(axis=-1, dtype=f32)
        best = p_z_c2.argmax(axis=2)
        p_best = np.take_along_axis(p_z_c2, best[:, :, None], 2)[:, :, 0]
        # Relative gap between the best and the second-best prototype of every sample.  Samples far from all
        # gaussians have p(z|c2) equal up to rounding, so the reference's BMU choice there is decided by the last
        # ulp of exp(); tests use this margin to leave such rays out of the loss_kl / som_vars comparison.
        srt = np.sort(p_z_c2, axis=2)
        self.debug["som_margin"] = (((srt[:, :, -1] - srt[:, :, -2]) / srt[:, :, -1]).min(axis=1) if srt.shape[2] > 1
                                    else np.full(srt.shape[0], np.inf, f32))          # one prototype: no decision to make
        new_means = np.zeros_like(means)
        new_vars = np.zeros_like(stds)
        for r in range(G):
            rel_w = np.take_along_axis(rel[:, r, :], best, 1)
            w = (rel_w * p_z_c1[:, :, r] / p_best + f32(1e-5)).astype(f32)
            wsum = w.sum(axis=1, dtype=f32)
            new_means[:, r] = (w * dist).sum(axis=1, dtype=f32) / wsum
            new_vars[:, r] = (w * (dist - new_means[:, r:r + 1]) ** 2).sum(axis=1, dtype=f32) / wsum
        mean_diffs = np.abs(means - new_means)
        var_diffs = np.abs(np.sqrt(var) - np.sqrt(new_vars))
        mask = ((mean_diffs > 0.1) & (new_vars > 0)) & ((var_diffs > 0.1) & (new_vars > 0))
        s2 = np.sqrt(new_vars).astype(f32)
        s2 = np.where(s2 < f32(1.5), f32(1.5), s2)                                       # kl_gauss :83
        std_err = np.log(s2 / stds + f32(1e-8))
        mean_err = (stds ** 2 + (means - new_means) ** 2) / (f32(2) * s2 ** 2)
        kl = (std_err + mean_err - f32(0.5)).astype(f32)
        loss = (kl * mask.astype(f32)).mean(axis=1, dtype=f32)
        return loss.astype(f32), new_means.astype(f32), new_vars.astype(f32)

    # -- scenerf.py:598-700 -------------------------------------------------------------------------------------
    def batchify_depth_and_color(self, T, x_rgb, pix, K, inv_K, noise_u, noise_n):
        cfg = self.cfg
        self._inv_K = inv_K
        n_uni = cfg.n_pts_uni
        if cfg.dataset == "bf" and n_uni <= 0:
            n_uni = 2                                                                    # scenerf_bf.py:623-626
        pts_u, depth_u, t_u, viewdir, unit = sample_rays_viewdir(inv_K, T, pix, n_uni, cfg.max_sample_depth, noise_u)
        means, stds = self.predict_gaussian_means_and_stds(T, unit, x_rgb, K, viewdir)
        pts_g, depth_g, t_g = sample_rays_gaussian(T, unit, means, stds, cfg.n_pts_per_gaussian, noise_n)
        if cfg.n_pts_uni > 0:
            pts = np.concatenate([pts_u, pts_g], axis=1)
            depth = np.concatenate([depth_u, depth_g], axis=1)
            t = np.concatenate([t_u, t_g], axis=1)
        elif cfg.n_pts_per_gaussian == 1:
            pts, depth, t = pts_u, depth_u, t_u
        else:
            pts, depth, t = pts_g, depth_g, t_g
        order = np.argsort(t, axis=1, kind="stable")
        t = np.take_along_axis(t, order, 1)
        depth = np.take_along_axis(depth, order, 1)
        pts = np.take_along_axis(pts, order[:, :, None], 1)
        self.debug["main_pts"] = pts
        self.debug["viewdir"] = viewdir
        density, colors = self.predict(self.pm, pts, x_rgb, K, viewdir, keep="main")
        self.debug["main_color"] = colors
        ro = self.render_depth_and_color(density, t, depth, colors)
        self.debug["argmin_margin"] = ro["argmin_margin"]
        self.debug["sorted_dist"] = t
        loss_kl, som_means, som_vars = self.ray_som(means, stds, np.where(t < 0, f32(0), t), ro["alphas"])
        return dict(depth=ro["depth"], color=ro["color"], gaussian_means=means, gaussian_stds=stds,
                    weights_at_depth=ro["weights_at_depth"], closest_pts_to_depths=ro["closest_pts_to_depth"],
                    loss_kl=loss_kl, alphas=ro["alphas"], som_vars=som_vars, densities=density,
                    weights=ro["weights"], depth_volumes=depth)

    # -- scenerf.py:392-471 -------------------------------------------------------------------------------------
    def render_rays_batch(self, K, T, x_rgb, sampled_pixels, ray_batch_size, noise_u, noise_n, inv_K=None):
        """Same chunking as the reference.  noise_u (R,U) / noise_n (R,G*P) replace the two torch RNG draws."""
        K = K.astype(f32)
        T = T.astype(f32)
        if inv_K is None:
            inv_K = np.linalg.inv(K).astype(f32)
        outs = []
        for s in range(0, sampled_pixels.shape[0], ray_batch_size):
            e = s + ray_batch_size
            outs.append(self.batchify_depth_and_color(T, x_rgb, sampled_pixels[s:e].astype(f32), K, inv_K,
                                                      noise_u[s:e], noise_n[s:e]))
        return {k: np.concatenate([o[k] for o in outs], axis=0) for k in outs[0]}
