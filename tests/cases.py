"""The golden cases: name -> SceneConfig + how their inputs were drawn (must mirror tests/golden/make_goldens.py)."""
import os

import numpy as np

from scenerf_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

RENDER_CASES = {
    "kitti_mini": (synth.config_A(name="kitti_mini", sphere_W=300, sphere_H=90, yaw_deg=10.0, tz=1.0), 31),
    "kitti_s128": (synth.config_B(name="kitti_s128", sphere_W=306, sphere_H=92), 32),
    "bf_mini": (synth.config_C(name="bf_mini", sphere_W=160, sphere_H=120, n_pts_uni=32), 33),
    "bf_s96": (synth.config_C(name="bf_s96", sphere_W=160, sphere_H=120), 38),
    "kitti_identity": (synth.config_A(name="kitti_identity", sphere_W=300, sphere_H=90, yaw_deg=0.0, tz=0.0), 34),
}

PREDICT_CASES = {
    "predict_adversarial_kitti": (synth.config_A(name="adv_kitti", sphere_W=300, sphere_H=90), 35),
    "predict_adversarial_kitti_full": (synth.config_A(name="adv_kitti_full"), 36),
    "predict_adversarial_bf": (synth.config_C(name="adv_bf", sphere_W=160, sphere_H=120), 37),
}


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as g:
        return {k: g[k] for k in g.files}


_cache = {}


def pyramid_for(cfg, seed):
    key = (seed, cfg.sphere_W, cfg.sphere_H)
    if key not in _cache:
        _cache[key] = synth.make_pyramid(seed, cfg.sphere_W, cfg.sphere_H)
    return _cache[key]


def params_for(cfg):
    key = ("params", cfg.dataset)
    if key not in _cache:
        _cache[key] = synth.make_model_params(cfg)
    return _cache[key]
