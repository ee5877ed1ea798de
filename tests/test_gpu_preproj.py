"""GPU: pre-projected latents (B200Renderer(preproject=True), srf_build_latent_table; SURVEY 7 hard part 3b).

SphericalMapping.from_pixels rounds the sphere coordinates (spherical_mapping.py:115), so lin_z[b](z) (resnetfc.py:148-150)
is a function of the integer sphere pixel; the table holds it for every pixel with a valid tap and the kernel adds table rows
instead of running the three lin_z GEMM passes of the main network.  Exact in real arithmetic -- checked here against
  * the dense path of the same precision mode on adversarial points (corners of every scale, zero-padding boundary,
    behind-camera sentinel, out-of-grid points: the zero row),
  * the reference's goldens at the unchanged tolerances of the mode (small grids and the full-size config-B grid)."""
import numpy as np
import pytest

from cases import FULL_CASES, PREDICT_CASES, RENDER_CASES, load_golden, pyramid_for
from helpers import make_renderer, torch_pyramid
from test_gpu_parity import _compare_with_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec,tol", [("fp32tc", 3e-5), ("fp16", 1e-2)])
@pytest.mark.parametrize("name", sorted(PREDICT_CASES))
def test_table_vs_dense_on_adversarial_points(name, prec, tol):
    import torch
    cfg, seed = PREDICT_CASES[name]
    g = load_golden(name)
    x_rgb = torch_pyramid(cfg, seed)
    K = torch.from_numpy(cfg.K)
    args = (torch.from_numpy(g["cam_pts"]), x_rgb, K, None, torch.from_numpy(g["viewdir"]))
    dense = make_renderer(cfg, prec).predict("mlp", *args, output_type="offset").cpu().numpy()
    r = make_renderer(cfg, prec, preproject=True)
    tab = r.predict("mlp", *args, output_type="offset").cpu().numpy()
    assert r.last_pack_launches == 36                       # per network: 15 GEMMs + 3 blend kernels built its table
    err = float(np.abs(tab - dense).max())
    mag = float(max(1.0, np.abs(dense).max()))
    print("%s/%s: table vs dense raw-output max-abs-err %.3e (max |out| %.3e)" % (name, prec, err, mag))
    assert np.isfinite(tab).all() and err <= tol * mag
    # the gaussian-proposal network has its own table (its own lin_z weights)
    og = r.predict("mlp_gaussian", *args, output_type="offset").cpu().numpy()
    od = make_renderer(cfg, prec).predict("mlp_gaussian", *args, output_type="offset").cpu().numpy()
    assert np.abs(og - od).max() <= tol * float(max(1.0, np.abs(od).max()))


@pytest.mark.parametrize("prec", ["fp32tc", "fp16"])
@pytest.mark.parametrize("name", sorted(RENDER_CASES))
def test_render_with_table_vs_reference_golden(name, prec):
    cfg, seed = RENDER_CASES[name]
    _compare_with_golden(name + "+table", cfg, load_golden(name), prec, pyramid_for(cfg, seed), preproject=True)


def test_full_size_B_with_table_vs_reference_golden():
    """config B at full size (1226x370 grid: 455 k table rows, 2.8 GB fp32 / 1.4 GB fp16) against the reference's outputs."""
    from scenerf_b200 import synth
    cfg, seed = FULL_CASES["full_B"]
    g = load_golden("full_B")
    pyr = synth.make_pyramid(seed, cfg.sphere_W, cfg.sphere_H)
    for prec in ("fp32tc", "fp16"):
        _compare_with_golden("full_B+table", cfg, g, prec, pyr, preproject=True)
