"""Shared test helpers: build a B200Renderer from a synth.SceneConfig and compare dicts against goldens."""
import numpy as np

from cases import params_for, pyramid_for


def hp_from_cfg(cfg):
    v_min, v_max, h_min, h_max = cfg.angles()
    return dict(dataset=cfg.dataset, n_pts_uni=cfg.n_pts_uni, n_gaussians=cfg.n_gaussians,
                n_pts_per_gaussian=cfg.n_pts_per_gaussian, std=cfg.std, max_sample_depth=cfg.max_sample_depth,
                out_img_W=cfg.sphere_W, out_img_H=cfg.sphere_H, som_sigma=cfg.som_sigma, v_angle_min=v_min,
                v_angle_max=v_max, h_angle_min=h_min, h_angle_max=h_max)


def make_renderer(cfg, precision, device="cuda:0", **kw):
    import torch
    from scenerf_b200.renderer import B200Renderer
    pm, pg = params_for(cfg)
    to = lambda d: {k: torch.from_numpy(v) for k, v in d.items()}
    return B200Renderer(hp_from_cfg(cfg), to(pm), to(pg), device=device, precision=precision, **kw)


def torch_pyramid(cfg, seed, device="cuda:0"):
    import torch
    return {k: torch.from_numpy(v).to(device) for k, v in pyramid_for(cfg, seed).items()}


def max_err(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max()) if np.size(a) else 0.0
