"""CPU tests: the numpy oracle (oracle/scenerf_oracle.py) against what pins it --
  * the reference's only known-answer check (FOV constants, scripts/determine_angles.py <-> scenerf.py:84-87,
    scenerf_bf.py:84-87), and
  * golden vectors produced by the reference itself (tests/golden/make_goldens.py).
Tolerances are float32 round-off of a different libm / BLAS summation order, stated per quantity."""
import numpy as np
import pytest

from cases import RENDER_CASES, PREDICT_CASES, FULL_CASES, load_golden, pyramid_for, params_for
from oracle import scenerf_oracle as orc
from scenerf_b200 import synth

# abs tolerance = ATOL + RTOL * max(1, max|golden|) (logit-level round-off is O(1)-scaled) ; float32 pipeline with ~10 chained 512..2480-long dot products
RTOL = 1e-4
ATOL = 2e-6


def _close(a, b, name, rtol=RTOL, atol=ATOL, scale=None):
    """scale: magnitude the round-off is proportional to (defaults to max|golden|); e.g. |depth - z| inherits the
    round-off of depth, not of its own (small) value."""
    tol = atol + rtol * float(max(1.0, np.abs(b).max()) if scale is None else scale)
    err = float(np.abs(a - b).max())
    assert err <= tol, "%s: max-abs-err %.3e > tol %.3e" % (name, err, tol)


def test_angles_known_answer():
    """spherical_mapping.py:8-18,95-102 on every pixel must reproduce the constants the reference hard-codes."""
    g = load_golden("angles_kat")
    expect = {"kitti": (75.4815, 104.7294, 49.5950, 131.1128), "bf": (67.6248, 112.2911, 61.2383, 118.6861)}
    for name, K, W, H in (("kitti", synth.KITTI_K, 1220, 370), ("bf", synth.BF_K, 640, 480)):
        inv_K = np.linalg.inv(K).astype(np.float32)
        pix = synth.grid_pixels(W, H)
        # un-rounded coords with a unit mapping (min=0, fov=1, size=2 -> value == angle)
        _, fl = orc.sphere_coords_from_pixels(pix, inv_K, (0.0, 1.0, 0.0, 1.0), 2, 2)
        got = np.array([fl[:, 1].min(), fl[:, 1].max(), fl[:, 0].min(), fl[:, 0].max()])
        np.testing.assert_allclose(got, np.array(expect[name]), atol=2e-4)       # constants printed to 4 decimals
        np.testing.assert_allclose(got, g[name], atol=2e-5)                      # vs the reference run here


def test_torch_linspace_matches_reference_distances():
    """uniform sample distances (utils.py:75-90) are recoverable from the golden depth_volumes only after the
    sort, so check the linspace restatement directly against values torch produced (recorded in the golden
    gaussian initial means: scenerf.py:556-560)."""
    g = load_golden("kitti_mini")
    cfg = RENDER_CASES["kitti_mini"][0]
    m0 = orc.torch_linspace(12.5, 87.5, 4)
    unit, _ = orc.compute_direction_from_pixels(g["pixels"], np.linalg.inv(cfg.K).astype(np.float32))
    pts = orc.cam_pts_2_cam_pts((m0[None, :, None] * unit[:, None, :]).reshape(-1, 3), cfg.T).reshape(-1, 4, 3)
    _close(pts, g["gauss_pts"], "gauss_pts", rtol=1e-6)


@pytest.mark.parametrize("name", sorted(RENDER_CASES))
def test_render_rays_batch_vs_reference(name):
    cfg, seed = RENDER_CASES[name]
    g = load_golden(name)
    pm, pg = params_for(cfg)
    o = orc.OracleRenderer(cfg, pm, pg)
    out = o.render_rays_batch(cfg.K, cfg.T, pyramid_for(cfg, seed), g["pixels"], g["pixels"].shape[0],
                              g["noise_u"], g["noise_n"])
    assert set(out) == {"depth", "color", "gaussian_means", "gaussian_stds", "weights_at_depth",
                        "closest_pts_to_depths", "loss_kl", "alphas", "som_vars", "densities", "weights",
                        "depth_volumes"}
    # discrete decisions
    S = cfg.S
    flips_main = (o.debug["main_sphere"] != g["main_sphere"]).any(axis=1).reshape(-1, S).any(axis=1)
    flips_gauss = (o.debug["gauss_sphere"] != g["gauss_sphere"]).any(axis=1).reshape(-1, cfg.n_gaussians).any(axis=1)
    clean = ~(flips_main | flips_gauss)
    assert clean.mean() > 0.98, "too many sphere-pixel rounding flips: %d rays" % (~clean).sum()
    _close(o.debug["gauss_pts"], g["gauss_pts"], "gauss_pts", rtol=1e-6)
    _close(o.debug["main_pts"][clean], g["main_pts"][clean], "main_pts")
    _close(o.debug["viewdir"], g["viewdir"], "viewdir", rtol=1e-6)
    for k in ("depth", "color", "gaussian_means", "gaussian_stds", "alphas", "densities", "weights",
              "depth_volumes"):
        _close(out[k][clean], g[k][clean], k)
    _close(out["closest_pts_to_depths"][clean], g["closest_pts_to_depths"][clean], "closest_pts_to_depths",
           scale=np.abs(g["depth_volumes"]).max())
    amin_ok = clean & (o.debug["argmin_margin"] > 1e-3)
    _close(out["weights_at_depth"][amin_ok], g["weights_at_depth"][amin_ok], "weights_at_depth")
    _check_som(out, g, clean, o.debug["som_margin"])


@pytest.mark.parametrize("name", sorted(FULL_CASES))
def test_render_rays_batch_full_size_vs_reference(name):
    """The oracle at the FULL sizes of BASELINE.json configs B / B' / C (256 rays of the frame) against the reference's
    own outputs: pins the restatement where bench.py's in-run parity check and the GPU full-size tests use it."""
    cfg, seed = FULL_CASES[name]
    g = load_golden(name)
    pm, pg = params_for(cfg)
    o = orc.OracleRenderer(cfg, pm, pg)
    out = o.render_rays_batch(cfg.K, cfg.T, synth.make_pyramid(seed, cfg.sphere_W, cfg.sphere_H), g["pixels"],
                              g["pixels"].shape[0], g["noise_u"], g["noise_n"])
    flips_main = (o.debug["main_sphere"] != g["main_sphere"]).any(axis=1).reshape(-1, cfg.S).any(axis=1)
    flips_gauss = (o.debug["gauss_sphere"] != g["gauss_sphere"]).any(axis=1).reshape(-1, cfg.n_gaussians).any(axis=1)
    clean = ~(flips_main | flips_gauss)
    assert clean.mean() > 0.98, "too many sphere-pixel rounding flips: %d rays" % (~clean).sum()
    for k in ("depth", "color", "gaussian_means", "gaussian_stds", "alphas", "densities", "weights", "depth_volumes"):
        _close(out[k][clean], g[k][clean], k)
    # 256 rays x 96-128 samples: one BundleFusion ray (som_sigma = 0.02, exponents of ~1e3) sits at a relative gap of
    # 5e-4 and moves loss_kl by 1e-3 relative -- the near-tie bound is 1e-3 here
    _check_som(out, g, clean, o.debug["som_margin"], margin_thr=1e-3)


def _check_som(out, g, clean, margin, rtol=1e-4, margin_thr=1e-5):
    """loss_kl / som_vars (ray_som_kl.py:10-78) hinge on a per-sample arg-max over prototypes.  For samples far from
    every gaussian all candidates are equal up to the last ulp of exp(), so the reference's own choice there is
    round-off; a ray may disagree only if it contains such a (near-)tie, and only a small fraction may."""
    bad = np.zeros(clean.shape, bool)
    for k in ("loss_kl", "som_vars"):
        a, b = out[k], g[k]
        tol = ATOL + rtol * max(1.0, float(np.abs(b).max()))
        err = np.abs(a - b).reshape(a.shape[0], -1).max(axis=1)
        bad |= err > tol
    bad &= clean
    assert not (bad & (margin > margin_thr)).any(), "SOM outputs differ on rays without an arg-max near-tie"
    assert bad.mean() <= 0.1, "SOM outputs differ on %d of %d rays" % (bad.sum(), bad.size)


@pytest.mark.parametrize("name", sorted(PREDICT_CASES))
def test_predict_adversarial_vs_reference(name):
    """Points crafted to hit the (W//s,H//s) corners of scales 2..16, the zero-padding boundary, the behind-camera
    pixel sentinel and the outside of the sphere grid."""
    cfg, seed = PREDICT_CASES[name]
    g = load_golden(name)
    pm, pg = params_for(cfg)
    o = orc.OracleRenderer(cfg, pm, pg)
    x_rgb = pyramid_for(cfg, seed)
    density, color = o.predict(pm, g["cam_pts"], x_rgb, cfg.K, g["viewdir"], keep="adv")
    off = o.predict(pg, g["cam_pts"], x_rgb, cfg.K, g["viewdir"], output_type="offset")
    same = (o.debug["adv_sphere"] == g["sphere"]).all(axis=1).reshape(density.shape)
    assert same.mean() > 0.97, "sphere coords differ for %d points" % (~same).sum()
    # the case must really exercise the coarse scales and the out-of-grid region
    sx, sy = g["sphere"][:, 0], g["sphere"][:, 1]
    assert ((sx <= cfg.sphere_W // 16) & (sy <= cfg.sphere_H // 16) & (sx >= 0) & (sy >= 0)).sum() >= 8
    assert ((sx < -1) | (sx > cfg.sphere_W + 1)).sum() >= 8
    _close(density[same], g["density"][same], "density")
    _close(color[same], g["color"][same], "color")
    _close(off[same], g["offset"][same], "offset")


def test_gather_quirk_q1_box_filter():
    """Integer sphere coords at scale 1 sample exactly between pixels: the gather is a 2x2 box average."""
    cfg = RENDER_CASES["kitti_mini"][0]
    fmap = pyramid_for(cfg, 31)["1_1"]
    coords = np.array([[5, 7], [0, 0], [cfg.sphere_W, cfg.sphere_H], [17, 1]], dtype=np.int64)
    got = orc.sample_feats_2d(fmap, coords, (cfg.sphere_W, cfg.sphere_H))
    C, H, W = fmap.shape
    pad = np.zeros((C, H + 2, W + 2), np.float32)
    pad[:, 1:-1, 1:-1] = fmap
    for i, (x, y) in enumerate(coords):
        box = 0.25 * (pad[:, y, x] + pad[:, y, x + 1] + pad[:, y + 1, x] + pad[:, y + 1, x + 1])
        np.testing.assert_allclose(got[i], box, atol=2e-4)


def test_composite_properties():
    rng = np.random.default_rng(0)
    R, S = 16, 64
    dist = np.sort(rng.uniform(0.2, 100, (R, S)).astype(np.float32), axis=1)
    dens = rng.uniform(0, 0.2, (R, S)).astype(np.float32)
    col = rng.uniform(0, 1, (R, S, 3)).astype(np.float32)
    ro = orc.OracleRenderer.render_depth_and_color(dens, dist, dist * 0.9, col)
    assert (ro["weights"] >= 0).all() and (ro["weights"].sum(1) <= 1 + 1e-5).all()
    assert (ro["depth"] <= dist.max(1) * 0.9 + 1e-4).all()
    assert (ro["alphas"] >= 0).all() and (ro["alphas"] < 1).all()
