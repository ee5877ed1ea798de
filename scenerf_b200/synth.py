"""Deterministic synthetic inputs for the SceneRF ray-render hot path.

Everything here is built from an integer hash (splitmix64) followed by an exact
integer->float32 conversion, so the same arrays are reproduced bit-for-bit on any
machine / numpy version.  The golden-vector generator (tests/golden/make_goldens.py,
which runs the *reference* renderer), the CPU oracle tests, the GPU parity tests and
bench.py all draw their weights / feature pyramids / rays from this module, which is why
no 40 MB weight blobs need to be committed as fixtures.

Shapes follow the reference:
  * ResnetFC parameters: scenerf/models/resnetfc.py:66-131 (lin_in, lin_z.{0,1,2},
    blocks.{0,1,2}.fc_{0,1}, lin_out), instantiated at scenerf/models/scenerf.py:100-114.
  * feature pyramid x_rgb: dict "1_1","1_2","1_4","1_8","1_16" of CHW fp32 maps with
    80/160/320/640/1280 channels (scenerf/models/unet2d_sphere.py:84-88,200-206); spatial
    size of scale s is (round(H/s), round(W/s)) (unet2d_sphere.py:139, python round()).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

SCALES = (1, 2, 4, 8, 16)
SCALE_KEYS = tuple("1_%d" % s for s in SCALES)
CHANNELS = (80, 160, 320, 640, 1280)
D_LATENT = sum(CHANNELS)          # 2480
D_PE = 39
D_VIEW = 3
D_IN = D_PE + D_VIEW              # 42
D_HIDDEN = 512
N_BLOCKS = 3

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def hash_uniform(seed: int, n: int) -> np.ndarray:
    """n float32 values in [-1, 1), exactly reproducible (splitmix64 on the index)."""
    with np.errstate(over="ignore"):
        x = np.arange(n, dtype=np.uint64) + np.uint64((seed * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    # top 24 bits -> [0, 2^24) -> exact in float32
    u = (x >> np.uint64(40)).astype(np.float32)
    return u * np.float32(2.0 ** -23) - np.float32(1.0)


def hash_unit(seed: int, n: int) -> np.ndarray:
    """n float32 values in [0, 1)."""
    return (hash_uniform(seed, n) + np.float32(1.0)) * np.float32(0.5)


def hash_normalish(seed: int, n: int) -> np.ndarray:
    """Zero-mean, unit-variance, bell-shaped float32 values (sum of 4 uniforms, exact ops)."""
    a = hash_uniform(seed * 4 + 0, n)
    b = hash_uniform(seed * 4 + 1, n)
    c = hash_uniform(seed * 4 + 2, n)
    d = hash_uniform(seed * 4 + 3, n)
    # var of U[-1,1) = 1/3 ; sum of four -> 4/3 ; scale to 1
    return (a + b + c + d) * np.float32(math.sqrt(3.0) / 2.0)


def pyramid_shapes(sphere_W: int, sphere_H: int):
    """(C, H_s, W_s) per scale; python round() == banker's rounding like the reference."""
    return [(c, round(sphere_H / s), round(sphere_W / s)) for c, s in zip(CHANNELS, SCALES)]


def make_pyramid(seed: int, sphere_W: int, sphere_H: int, std: float = 0.5):
    """dict key -> (C,H,W) float32 array, values ~ zero-mean with the given std."""
    out = {}
    for i, (key, (c, h, w)) in enumerate(zip(SCALE_KEYS, pyramid_shapes(sphere_W, sphere_H))):
        out[key] = (hash_normalish(seed * 8 + i, c * h * w) * np.float32(std)).reshape(c, h, w)
    return out


def mlp_param_shapes(d_out: int):
    """Ordered (name, shape) list == ResnetFC.state_dict() order is irrelevant; names are."""
    shp = [("lin_in.weight", (D_HIDDEN, D_IN)), ("lin_in.bias", (D_HIDDEN,)),
           ("lin_out.weight", (d_out, D_HIDDEN)), ("lin_out.bias", (d_out,))]
    for b in range(N_BLOCKS):
        shp += [("blocks.%d.fc_0.weight" % b, (D_HIDDEN, D_HIDDEN)), ("blocks.%d.fc_0.bias" % b, (D_HIDDEN,)),
                ("blocks.%d.fc_1.weight" % b, (D_HIDDEN, D_HIDDEN)), ("blocks.%d.fc_1.bias" % b, (D_HIDDEN,))]
    for b in range(N_BLOCKS):
        shp += [("lin_z.%d.weight" % b, (D_HIDDEN, D_LATENT)), ("lin_z.%d.bias" % b, (D_HIDDEN,))]
    return shp


def make_mlp_params(seed: int, d_out: int, out_scale: float = 1.0, out_bias=None):
    """Synthetic ResnetFC parameters (float32 numpy), kaiming-like fan-in scaling.

    Unlike the reference init (resnetfc.py:37-40: fc_1 zero, all biases zero) every tensor
    is non-trivial so that a kernel that dropped a bias or an fc_1 would be caught.
    `out_scale` rescales lin_out so that density*delta stays O(0.01..1).
    """
    params = {}
    for i, (name, shape) in enumerate(mlp_param_shapes(d_out)):
        n = int(np.prod(shape))
        if name.endswith(".weight"):
            fan_in = shape[1]
            std = math.sqrt(2.0 / fan_in)
            if ".fc_1." in name:
                std *= 0.5
            if name.startswith("lin_out"):
                std *= out_scale
            w = hash_normalish(seed * 64 + i, n) * np.float32(std)
        else:
            w = hash_uniform(seed * 64 + i, n) * np.float32(0.1)
            if name.startswith("lin_out"):
                w = w * np.float32(out_scale)
                if out_bias is not None:
                    w = w + np.asarray(out_bias, dtype=np.float32)
        params[name] = np.ascontiguousarray(w.reshape(shape), dtype=np.float32)
    return params


def yaw_translate(yaw_deg: float, tz: float, tx: float = 0.0, ty: float = 0.0) -> np.ndarray:
    """T = R_y(yaw) @ translate  (the composition used by utils.py:29-49 sample_rel_poses)."""
    rad = yaw_deg / 180.0 * math.pi
    rel = np.eye(4, dtype=np.float64)
    rel[0, 3] += tx
    rel[1, 3] += ty
    rel[2, 3] += tz
    rot = np.eye(4, dtype=np.float64)
    rot[:3, :3] = [[math.cos(rad), 0, math.sin(rad)], [0, 1, 0], [-math.sin(rad), 0, math.cos(rad)]]
    return (rot @ rel).astype(np.float32)


KITTI_K = np.array([[707.0912, 0.0, 601.8873], [0.0, 707.0912, 183.1104], [0.0, 0.0, 1.0]], dtype=np.float32)
BF_K = np.array([[583.0, 0.0, 320.0], [0.0, 583.0, 240.0], [0.0, 0.0, 1.0]], dtype=np.float32)


@dataclass
class SceneConfig:
    """Scalar hyper-parameters of one renderer instance (mirrors SceneRF.__init__ kwargs,
    scenerf.py:23-43 / scenerf_bf.py:27-50) plus the synthetic camera / pose."""
    name: str
    dataset: str = "kitti"              # "kitti" | "bf"
    img_W: int = 1220
    img_H: int = 370
    sphere_W: int = 1500
    sphere_H: int = 452
    n_pts_uni: int = 32
    n_gaussians: int = 4
    n_pts_per_gaussian: int = 8
    std: float = 2.0
    max_sample_depth: float = 100.0
    add_fov_hor: float = 20.0
    add_fov_ver: float = 8.0
    som_sigma: float = 2.0
    yaw_deg: float = 0.0
    tz: float = 1.0
    K: np.ndarray = field(default_factory=lambda: KITTI_K.copy())

    @property
    def S(self):
        return self.n_pts_uni + self.n_gaussians * self.n_pts_per_gaussian

    @property
    def T(self):
        return yaw_translate(self.yaw_deg, self.tz)

    def angles(self):
        """(v_min, v_max, h_min, h_max) as the reference constructor computes them
        (scenerf.py:83-88, scenerf_bf.py:84-88)."""
        if self.dataset == "kitti":
            return (75.4815 - self.add_fov_ver, 104.7294 + self.add_fov_ver,
                    49.5950 - self.add_fov_hor, 131.1128 + self.add_fov_hor)
        return (67.6248 - self.add_fov_ver, 112.2911 + self.add_fov_ver,
                61.2383 - self.add_fov_hor, 118.6861 + self.add_fov_hor)


def config_A(**kw):
    d = dict(name="A")
    d.update(kw)
    return SceneConfig(**d)


def config_B(**kw):
    d = dict(name="B", img_W=1226, img_H=370, sphere_W=1226, sphere_H=370, n_pts_uni=64,
             n_pts_per_gaussian=16, yaw_deg=10.0, tz=2.0)
    d.update(kw)
    return SceneConfig(**d)


def config_C(**kw):
    d = dict(name="C", dataset="bf", img_W=640, img_H=480, sphere_W=640, sphere_H=480, n_pts_uni=64,
             n_pts_per_gaussian=8, std=0.1, max_sample_depth=12.0, add_fov_hor=14.0, add_fov_ver=11.0,
             som_sigma=0.02, yaw_deg=30.0, tz=0.4, K=BF_K.copy())
    d.update(kw)
    return SceneConfig(**d)


def random_pixels(seed: int, n: int, W: int, H: int) -> np.ndarray:
    """(n,2) float32 (x,y) non-integer pixel coordinates in [0,W)x[0,H)."""
    px = hash_unit(seed * 2 + 0, n) * np.float32(W)
    py = hash_unit(seed * 2 + 1, n) * np.float32(H)
    return np.stack([px, py], axis=1).astype(np.float32)


def grid_pixels(W: int, H: int, stride: int = 1) -> np.ndarray:
    """x-major integer pixel grid like render_colors.py:102-111 (meshgrid 'ij' of xs, ys)."""
    xs = np.arange(0, W, stride, dtype=np.float32)
    ys = np.arange(0, H, stride, dtype=np.float32)
    gx, gy = np.meshgrid(xs, ys, indexing="ij")
    return np.stack([gx.reshape(-1), gy.reshape(-1)], axis=1).astype(np.float32)


def make_model_params(cfg: "SceneConfig", seed: int = 11):
    """(main, gaussian) ResnetFC parameter dicts for a config.  lin_out is rescaled / re-biased so that the
    rendered transmittance decays over the whole ray (sigma*delta ~ 0.01..0.3) and the gaussian offsets move the
    proposals by O(1) -- otherwise parity tests would only exercise the first few samples of a ray."""
    dens_bias = -3.7 if cfg.dataset == "kitti" else -0.7
    main = make_mlp_params(seed, 4, out_scale=0.1, out_bias=[0.0, -2.0, 0.9, dens_bias])
    gauss = make_mlp_params(seed + 1, 2, out_scale=0.1, out_bias=[3.3, 2.6])
    return main, gauss


# ---- spherical decoder (producer of the pyramid): scenerf/models/unet2d_sphere.py:60-135 ------------------------------------------
DECODER_SKIP_CHANNELS = {16: 224, 8: 80, 4: 48, 2: 32, 1: 3}       # EfficientNet block widths the decoder concatenates (unet2d_sphere.py:91-105)


def decoder_level_channels(num_features: int):
    """up-level -> (Cin of its first conv, Cout), as DecoderSphere.__init__ wires them (unet2d_sphere.py:84-135)."""
    f = int(num_features)
    out = {16: f // 2, 8: f // 4, 4: f // 8, 2: f // 16, 1: f // 32}
    prev = {16: f, 8: out[16], 4: out[8], 2: out[4], 1: out[2]}
    return {s: (prev[s] + DECODER_SKIP_CHANNELS[s], out[s]) for s in (16, 8, 4, 2, 1)}


def make_decoder_params(num_features: int, bottleneck_features: int, seed: int = 21):
    """Deterministic DecoderSphere parameters + BatchNorm running statistics (state_dict names of unet2d_sphere.py), for the
    modules `forward` uses: conv2 and up16/up8/up4/up2/up1 (`_net.0` conv, `_net.1-3` BasicBlocks)."""
    p = {}
    k = [0]

    def nxt():
        k[0] += 1
        return seed * 1000 + k[0]

    def conv(name, cout, cin, ks):
        n = cout * cin * ks * ks
        p[name + ".weight"] = (hash_normalish(nxt(), n) * np.float32(math.sqrt(2.0 / (cin * ks * ks)))).reshape(cout, cin, ks, ks).astype(np.float32)
        p[name + ".bias"] = (hash_uniform(nxt(), cout) * np.float32(0.1)).astype(np.float32)

    def bn(name, c):
        p[name + ".weight"] = (np.float32(1.0) + hash_uniform(nxt(), c) * np.float32(0.3)).astype(np.float32)
        p[name + ".bias"] = (hash_uniform(nxt(), c) * np.float32(0.2)).astype(np.float32)
        p[name + ".running_mean"] = (hash_uniform(nxt(), c) * np.float32(0.2)).astype(np.float32)
        p[name + ".running_var"] = (np.float32(1.0) + hash_uniform(nxt(), c) * np.float32(0.4)).astype(np.float32)

    conv("conv2", int(num_features), int(bottleneck_features), 1)
    for s, (cin, cout) in decoder_level_channels(num_features).items():
        pre = "up%d._net." % s
        conv(pre + "0", cout, cin, 3)
        for blk in (1, 2, 3):
            for cb in (1, 2):
                conv(pre + "%d.conv_block%d.0" % (blk, cb), cout, cout, 3)
                bn(pre + "%d.conv_block%d.1" % (blk, cb), cout)
    return p
