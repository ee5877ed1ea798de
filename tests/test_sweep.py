"""Novel-view sweep glue (hot-path contract row 8f-4): pose tables, pixel grid, transpose + bilinear upsample, PNG
round trip, sweep -> TSDF.  Golden = the reference's own functions run on a small image (tests/golden/make_goldens.py
`sweep_kitti`).  CPU tests pin the oracle restatement; GPU tests compare the CUDA path with the same golden."""
import numpy as np
import pytest

from cases import load_golden, params_for
from oracle import sweep_oracle as so
from oracle.tsdf_oracle import TSDFVolumeOracle
from scenerf_b200 import synth

N_POSES = 6


def sweep_cfg():
    cfg = synth.config_A(name="sweep_kitti", sphere_W=300, sphere_H=90)
    cfg.img_W, cfg.img_H = 244, 74
    cfg.K = synth.KITTI_K.copy()
    cfg.K[:2] /= 5.0
    return cfg, 4, dict(step=1.0, angle=10, max_distance=1.1), 39


@pytest.fixture(scope="module")
def g():
    return load_golden("sweep_kitti")


def test_pose_tables_match_reference(g):
    from scenerf_b200 import sweep
    for mod in (so, sweep):
        for fn, kw, key in ((mod.sample_rel_poses, dict(step=1.0, angle=10, max_distance=1.1), ""),
                            (mod.sample_rel_poses, dict(step=0.5, angle=10, max_distance=10.1), "full_"),
                            (mod.sample_rel_poses_bf, dict(angle=15, max_distance=0.7, step=0.2), "bf_")):
            poses = fn(**kw)
            keys = np.array([[s, a] for s, a in poses.keys()], dtype=np.float64)
            assert np.array_equal(keys, g[key + "pose_keys"]), (mod.__name__, key)
            assert np.array_equal(np.stack([np.asarray(v) for v in poses.values()]), g[key + "poses"]), (mod.__name__, key)
    assert len(so.sample_rel_poses(step=0.5, angle=10, max_distance=10.1)) == 63       # the "63 poses per frame" of SURVEY 8f


def test_pixel_grid_matches_reference(g):
    pix, shape = so.pixel_grid((244, 74), 4)
    assert shape == (61, 19) and np.array_equal(pix, g["pixels"])
    from scenerf_b200 import sweep
    p2, grid = sweep.pixel_grid((244, 74), 4, "cpu")
    assert grid == (61, 19) and np.array_equal(p2.numpy(), g["pixels"])


def test_upsample_restatement(g):
    up = so.upsample_bilinear(g["interp_src"].T, 23, 50)
    assert np.abs(up - g["interp_dst"]).max() <= 1e-6 * max(1.0, np.abs(g["interp_dst"]).max())
    for i in (0, 4):
        d, c = so.to_images(g["depth_rays%d" % i], g["color_rays%d" % i], (61, 19), (244, 74), 4)
        assert np.abs(d - g["depth%d" % i]).max() <= 1e-5
        assert np.abs(c - g["color%d" % i].astype(np.float32)).max() <= 1e-3        # stored as float16
        u8 = (c * np.float32(255)).astype(np.uint8)
        assert (u8 != g["rgb_tsdf%d" % i]).mean() <= 1e-4                           # 1-ulp blends can cross an integer


def test_sweep_to_tsdf_restatement(g):
    """Golden rays -> images -> PNG round trip -> TSDF oracle == the volume the reference pipeline built."""
    vol = TSDFVolumeOracle(g["vol_bnds"], 0.2, 10)
    inv = np.linalg.inv(g["T_velo2cam"])
    for i in range(N_POSES):
        d, c = so.to_images(g["depth_rays%d" % i], g["color_rays%d" % i], (61, 19), (244, 74), 4)
        vol.integrate(so.png_roundtrip(c), d, g["K"], inv @ g["poses"][i].astype(np.float64), 1.0)
    assert np.array_equal(vol.weight, g["tsdf_weight"])
    assert np.abs(vol.tsdf - g["tsdf"]).max() <= 1e-5
    assert (vol.color != g["tsdf_color"]).mean() <= 1e-3


def test_oracle_render_of_one_sweep_pose(g):
    from oracle.scenerf_oracle import OracleRenderer
    cfg, scale, _, pyr_seed = sweep_cfg()
    orc = OracleRenderer(cfg, *params_for(cfg))
    sel = np.arange(0, g["pixels"].shape[0], 9)
    out = orc.render_rays_batch(cfg.K, g["poses"][1], synth.make_pyramid(pyr_seed, cfg.sphere_W, cfg.sphere_H),
                                g["pixels"][sel], len(sel), g["noise_u1"][sel], g["noise_n1"][sel])
    assert np.abs(out["depth"] - g["depth_rays1"][sel]).max() <= 1e-4 * cfg.max_sample_depth
    assert np.abs(out["color"] - g["color_rays1"][sel]).max() <= 1e-4


# ------------------------------------------------------------------------------------------------ GPU

def frac_within(a, b, tol):
    """Rendered values are compared per pixel with a small allowance: a last-ulp difference in the rounded sphere pixel
    of one sample (~1e-4 of the points) changes that ray by centimetres in the reference itself."""
    return float((np.abs(np.asarray(a, dtype=np.float64) - b) <= tol).mean())


def _renderer(precision):
    import torch
    from helpers import make_renderer
    cfg, scale, pose_kw, pyr_seed = sweep_cfg()
    r = make_renderer(cfg, precision=precision)
    x_rgb = {k: torch.from_numpy(v).cuda() for k, v in synth.make_pyramid(pyr_seed, cfg.sphere_W, cfg.sphere_H).items()}
    return cfg, r, x_rgb


@pytest.mark.gpu
def test_rays_to_images_cuda(g):
    import torch
    from scenerf_b200 import sweep
    for i in (0, 4):
        dr, cr = torch.from_numpy(g["depth_rays%d" % i]).cuda(), torch.from_numpy(g["color_rays%d" % i]).cuda()
        d, c = sweep.rays_to_images(dr, cr, (61, 19), (244, 74), sweep.COLOR_CLAMP)
        assert np.abs(d.cpu().numpy() - g["depth%d" % i]).max() <= 1e-5
        assert np.abs(c.cpu().numpy() - g["color%d" % i].astype(np.float32)).max() <= 1e-3
        _, q = sweep.rays_to_images(None, cr, (61, 19), (244, 74), sweep.COLOR_PNG)
        q = q.cpu().numpy()
        ref = (g["rgb_tsdf%d" % i].astype(np.float32) / np.float32(255.0)) * np.float32(255.0)
        assert (q != ref).mean() <= 1e-4
        # bit-exact against the restatement (same arithmetic order, no FMA contraction)
        od, oc = so.to_images(g["depth_rays%d" % i], g["color_rays%d" % i], (61, 19), (244, 74), 4)
        assert np.array_equal(d.cpu().numpy(), od) and np.array_equal(c.cpu().numpy(), oc)
        assert np.array_equal(q, so.png_roundtrip(oc))
    # non-integer ratio + scale-1 transpose
    src = torch.from_numpy(np.ascontiguousarray(g["interp_src"])).cuda()
    d, _ = sweep.rays_to_images(src.reshape(-1), None, (17, 8), (50, 23))
    assert np.abs(d.cpu().numpy() - g["interp_dst"]).max() <= 1e-6 * max(1.0, np.abs(g["interp_dst"]).max())
    d, _ = sweep.rays_to_images(src.reshape(-1), None, (17, 8), (17, 8))
    assert np.array_equal(d.cpu().numpy(), g["interp_src"].T)


@pytest.mark.gpu
def test_sweep_reconstruct_fp32_matches_reference_pipeline(g):
    import torch
    from scenerf_b200 import sweep
    cfg, r, x_rgb = _renderer("fp32")
    sw = sweep.NovelDepthSweep(r, torch.from_numpy(cfg.K).cuda(), x_rgb, img_size=(244, 74), scale=4)
    assert np.array_equal(sw.pixels.cpu().numpy(), g["pixels"])
    poses = sweep.sample_rel_poses(step=1.0, angle=10, max_distance=1.1)
    noises = [(torch.from_numpy(g["noise_u%d" % i]), torch.from_numpy(g["noise_n%d" % i])) for i in range(N_POSES)]
    d1, c1 = sw.render(list(poses.values())[4].cuda(), sweep.COLOR_CLAMP, noises[4])
    assert frac_within(d1.cpu().numpy(), g["depth4"], 2e-4 * cfg.max_sample_depth) >= 0.97
    assert np.abs(d1.cpu().numpy() - g["depth4"]).max() <= 1.0
    assert frac_within(c1.cpu().numpy(), g["color4"].astype(np.float32), 1.5e-3) >= 0.97
    vol = sw.reconstruct(poses, g["T_velo2cam"], g["vol_bnds"], noises=noises)
    tsdf, color = vol.get_volume()
    assert np.array_equal(vol.get_weight(), g["tsdf_weight"]) or (vol.get_weight() != g["tsdf_weight"]).mean() <= 1e-4
    same = vol.get_weight() == g["tsdf_weight"]
    assert frac_within(tsdf[same], g["tsdf"][same], 2e-4 * cfg.max_sample_depth) >= 0.97
    assert np.abs(tsdf - g["tsdf"])[same].max() <= 1.0
    assert sw.launches > 0


@pytest.mark.gpu
def test_volume_merge_equals_sequential_integration(g):
    """Poses split over two volumes and merged in order == all poses in one volume (the multi-GPU rule)."""
    import torch
    from scenerf_b200 import sweep
    from scenerf_b200.tsdf import TSDFVolume
    inv = np.linalg.inv(g["T_velo2cam"])
    frames = []
    for i in range(N_POSES):
        dr, cr = torch.from_numpy(g["depth_rays%d" % i]).cuda(), torch.from_numpy(g["color_rays%d" % i]).cuda()
        d, q = sweep.rays_to_images(dr, cr, (61, 19), (244, 74), sweep.COLOR_PNG)
        frames.append((q, d, inv @ g["poses"][i].astype(np.float64)))

    def fuse(idx):
        v = TSDFVolume(g["vol_bnds"].copy(), voxel_size=0.2)
        for i in idx:
            v.integrate(frames[i][0], frames[i][1], g["K"], frames[i][2])
        return v
    seq = fuse(range(N_POSES))
    for cut in (1, 3, 5):
        a, b = fuse(range(cut)), fuse(range(cut, N_POSES))
        a.merge_(b._tsdf, b._weight, b._color)
        assert np.array_equal(a.get_volume()[0], seq.get_volume()[0])
        assert np.array_equal(a.get_weight(), seq.get_weight())
        # colour: exact except where two observations' distances round to the same float32 (csrc/image_ops.cu)
        assert (a.get_volume()[1] != seq.get_volume()[1]).mean() <= 1e-4
    # and the sequential device volume is the reference pipeline's volume
    assert np.array_equal(seq.get_weight(), g["tsdf_weight"])
    assert np.abs(seq.get_volume()[0] - g["tsdf"]).max() <= 1e-5


@pytest.mark.gpu
def test_sweep_fp16_default_path(g):
    import torch
    from scenerf_b200 import sweep
    cfg, r, x_rgb = _renderer("fp16")
    sw = sweep.NovelDepthSweep(r, torch.from_numpy(cfg.K).cuda(), x_rgb, img_size=(244, 74), scale=4)
    poses = sweep.sample_rel_poses(step=1.0, angle=10, max_distance=1.1)
    noises = [(torch.from_numpy(g["noise_u%d" % i]), torch.from_numpy(g["noise_n%d" % i])) for i in range(N_POSES)]
    d, c = sw.render(list(poses.values())[0].cuda(), sweep.COLOR_CLAMP, noises[0])
    assert frac_within(d.cpu().numpy(), g["depth0"], 1.5e-3 * cfg.max_sample_depth) >= 0.97   # stated fp16 tolerance
    assert frac_within(c.cpu().numpy(), g["color0"].astype(np.float32), 3e-3) >= 0.97
    vol = sw.reconstruct(poses, g["T_velo2cam"], g["vol_bnds"], noises=noises)
    assert (vol.get_weight() != g["tsdf_weight"]).mean() <= 2e-3
    same = vol.get_weight() == g["tsdf_weight"]
    assert frac_within(vol.get_volume()[0][same], g["tsdf"][same], 1.5e-3 * cfg.max_sample_depth) >= 0.97


@pytest.mark.gpu
def test_sweep_pose_sharded_two_gpus():
    """Multi-GPU path of the sweep (one process per GPU, NCCL): needs >= 2 GPUs, skipped on a 1-GPU box."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(here, "_sweep_dist_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "SWEEP_DIST_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


def test_volume_merge_rule_on_the_oracle(g):
    """The fold rule behind srf_tsdf_merge, on the CPU oracle: fusing contiguous pose ranges separately and folding the later
    volume into the earlier one (keep A where |A| < |B|, weights add) reproduces the sequential integration -- distances and
    weights exactly; colours except on float32-exact distance ties (see csrc/image_ops.cu)."""
    inv = np.linalg.inv(g["T_velo2cam"])
    frames = []
    for i in range(N_POSES):
        d, c = so.to_images(g["depth_rays%d" % i], g["color_rays%d" % i], (61, 19), (244, 74), 4)
        frames.append((so.png_roundtrip(c), d, inv @ g["poses"][i].astype(np.float64)))

    def fuse(idx):
        v = TSDFVolumeOracle(g["vol_bnds"], 0.2, 10)
        for i in idx:
            v.integrate(frames[i][0], frames[i][1], g["K"], frames[i][2], 1.0)
        return v
    seq = fuse(range(N_POSES))
    for cut in (2, 3):
        a, b = fuse(range(cut)), fuse(range(cut, N_POSES))
        t, w, c = np.full_like(seq.tsdf, 255), np.zeros_like(seq.tsdf), np.zeros_like(seq.tsdf)
        for p in (a, b):
            obs = p.weight != 0
            take = obs & ~(np.abs(t) < np.abs(p.tsdf))
            w = w + np.where(obs, p.weight, 0)
            t, c = np.where(take, p.tsdf, t), np.where(take, p.color, c)
        assert np.array_equal(t, seq.tsdf) and np.array_equal(w, seq.weight)
        assert (c != seq.color).mean() <= 1e-4
