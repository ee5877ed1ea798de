"""Runs the reference's OWN renderer class on the host CPU (bench.py's CPU arm, `cpu_baseline.kind == "reference"`).

TEST INFRASTRUCTURE ONLY (see oracle/build_ref.py): imports the unmodified reference sources staged under
oracle/_ref/ (or /root/reference when present) through the SURVEY 8c shim --
  * `pytorch_lightning` (absent here) -> a stand-in module whose LightningModule is nn.Module,
  * `UNet2DSphere.build` (torch.hub, needs network) -> nn.Identity; the feature pyramid is an input of the path --
builds `SceneRF` (scenerf.py:22 / scenerf_bf.py:27) with the synthetic weights of scenerf_b200.synth and calls
`render_rays_batch` exactly as the reference's scripts do (render_colors.py:113-119, save_depth_metrics.py:105-118).
"""
from __future__ import annotations

import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_STAGED = os.path.join(HERE, "_ref")


def reference_root():
    """Directory to put on sys.path, or None: the staged copy first (identical on every machine), else /root/reference."""
    for cand in (REF_STAGED, os.environ.get("SCENERF_REFERENCE", "/root/reference")):
        if cand and os.path.exists(os.path.join(cand, "scenerf", "models", "scenerf.py")):
            return cand
    return None


def available() -> bool:
    return reference_root() is not None


def _install_shims():
    import torch.nn as nn
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class _LM(nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

            def log(self, *a, **k):
                pass

            @property
            def device(self):
                return next(self.parameters()).device

        pl.LightningModule = _LM
        sys.modules["pytorch_lightning"] = pl
    root = reference_root()
    if root not in sys.path:
        sys.path.insert(0, root)
    import scenerf.models.unet2d_sphere as U
    U.UNet2DSphere.build = classmethod(lambda cls, **kw: nn.Identity())


def build_model(cfg, seed: int = 11):
    """The reference module for a synth.SceneConfig, synthetic ResnetFC weights loaded, eval mode, CPU."""
    import torch
    from scenerf_b200 import synth
    _install_shims()
    if cfg.dataset == "kitti":
        from scenerf.models.scenerf import SceneRF
    else:
        from scenerf.models.scenerf_bf import SceneRF
    m = SceneRF(som_sigma=cfg.som_sigma, std=cfg.std, img_size=(cfg.img_W, cfg.img_H),
                max_sample_depth=cfg.max_sample_depth, n_gaussians=cfg.n_gaussians, n_pts_uni=cfg.n_pts_uni,
                n_pts_per_gaussian=cfg.n_pts_per_gaussian, add_fov_hor=cfg.add_fov_hor, add_fov_ver=cfg.add_fov_ver,
                sphere_H=cfg.sphere_H, sphere_W=cfg.sphere_W).eval()
    pm, pg = synth.make_model_params(cfg, seed)
    m.mlp.load_state_dict({k: torch.from_numpy(v) for k, v in pm.items()})
    m.mlp_gaussian.load_state_dict({k: torch.from_numpy(v) for k, v in pg.items()})
    return m


def render(model, cfg, x_rgb, pixels, ray_batch_size: int):
    """One reference call: the 12-key dict of scenerf.py:456-469 for `pixels` (R,2) float32 CPU tensor."""
    import torch
    K, T = torch.from_numpy(cfg.K), torch.from_numpy(cfg.T)
    with torch.no_grad():
        if cfg.dataset == "kitti":
            return model.render_rays_batch(K, T, x_rgb, ray_batch_size=ray_batch_size, sampled_pixels=pixels)
        return model.render_rays_batch(K, T, x_rgb, sampled_pixels=pixels, ray_batch_size=ray_batch_size)


def cpu_model_name() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


class ReferenceTimer:
    """Times `render_rays_batch` of the reference on bounded samples of a workload's rays.  One process, all host
    threads torch gives it (intra-op OpenMP/MKL): the way the reference itself would run on this host."""

    def __init__(self, cfg, pixels_np, pyramid_np, threads: int | None = None, seed: int = 0):
        import numpy as np
        import torch
        if threads:
            torch.set_num_threads(int(threads))
        self.threads = torch.get_num_threads()
        self.cfg = cfg
        self.model = build_model(cfg)
        self.x_rgb = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in pyramid_np.items()}
        self.pix = pixels_np
        self.order = np.random.default_rng(seed).permutation(pixels_np.shape[0])
        self.pos = 0
        torch.manual_seed(seed)

    def step(self, n_rays: int):
        """One reference call on the next n_rays rays of a fixed random order (wraps around); returns seconds."""
        import numpy as np
        import torch
        idx = self.order[(self.pos + np.arange(n_rays)) % self.order.shape[0]]
        self.pos += n_rays
        pix = torch.from_numpy(np.ascontiguousarray(self.pix[idx]))
        t0 = time.perf_counter()
        out = render(self.model, self.cfg, self.x_rgb, pix, ray_batch_size=n_rays)
        self.checksum = float(out["depth"].sum())
        return time.perf_counter() - t0
