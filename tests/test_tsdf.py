"""TSDF fusion (hot-path contract row 8f-2): CPU test = oracle vs the golden produced by the reference's own TSDFVolume;
GPU test = the CUDA kernel (through the C ABI) vs the same golden -- bit-exact, the arithmetic that decides pixels and
merges is float64 / exact float32 on both sides."""
import numpy as np
import pytest

from cases import load_golden
from oracle.tsdf_oracle import TSDFVolumeOracle, fold_color


def test_tsdf_oracle_matches_reference_golden():
    g = load_golden("tsdf_fusion")
    o = TSDFVolumeOracle(g["vol_bnds"], 0.2, 10)
    for i in range(3):
        o.integrate(g["rgb%d" % i], g["depth%d" % i], g["K"], g["pose%d" % i], 1.0)
    assert (g["tsdf"] != 255).mean() > 0.3          # the case really touches the volume
    assert np.array_equal(o.tsdf, g["tsdf"]) and np.array_equal(o.weight, g["weight"]) and np.array_equal(o.color, g["color"])


def test_fold_color_is_exact_for_8bit_images():
    rng = np.random.default_rng(0)
    im = rng.integers(0, 256, (5, 7, 3)).astype(np.float32)
    f = fold_color(im)
    assert np.array_equal(f, im[..., 2] * 65536 + im[..., 1] * 256 + im[..., 0])


@pytest.mark.gpu
@pytest.mark.parametrize("as_u8", [False, True])
def test_tsdf_cuda_matches_reference_golden(as_u8):
    import torch
    from scenerf_b200.tsdf import TSDFVolume
    g = load_golden("tsdf_fusion")
    vol = TSDFVolume(g["vol_bnds"].copy(), voxel_size=0.2, trunc_margin=10)
    for i in range(3):
        rgb = torch.from_numpy(g["rgb%d" % i])
        rgb = rgb.to(torch.uint8) if as_u8 else rgb
        vol.integrate(rgb.cuda(), torch.from_numpy(g["depth%d" % i]).cuda(), g["K"], g["pose%d" % i], obs_weight=1.)
    tsdf, color = vol.get_volume()
    assert np.array_equal(tsdf, g["tsdf"])
    assert np.array_equal(vol.get_weight(), g["weight"])
    assert np.array_equal(color, g["color"])


@pytest.mark.gpu
def test_tsdf_properties_full_kitti_volume():
    """256x256x32 volume of depth2tsdf.py:87-93 with 370x1220 renders: idempotence (integrating the same frame twice
    changes weights only) and order independence without ties."""
    import torch
    from scenerf_b200.tsdf import TSDFVolume
    from scenerf_b200 import synth
    vol_bnds = np.zeros((3, 2)); vol_bnds[:, 0] = [0, -25.6, -2]; vol_bnds[:, 1] = vol_bnds[:, 0] + [51.2, 51.2, 6.4]
    T_velo2cam = np.array([[0.0, -1.0, 0.0, 0.0], [0.0, 0.0, -1.0, -0.08], [1.0, 0.0, 0.0, -0.27], [0, 0, 0, 1.0]])
    H, W = 370, 1220
    frames = []
    for i, (yaw, tz) in enumerate(((0.0, 0.0), (10.0, 2.0), (-10.0, 4.0))):
        depth = torch.from_numpy((5.0 + 30.0 * synth.hash_unit(80 + i, H * W)).reshape(H, W).astype(np.float32)).cuda()
        rgb = torch.from_numpy(np.floor(synth.hash_unit(90 + i, H * W * 3) * 256).reshape(H, W, 3).astype(np.float32)).cuda()
        frames.append((rgb, depth, np.linalg.inv(T_velo2cam) @ synth.yaw_translate(yaw, tz).astype(np.float64)))

    def run(order):
        v = TSDFVolume(vol_bnds.copy(), voxel_size=0.2)
        for j in order:
            v.integrate(frames[j][0], frames[j][1], synth.KITTI_K, frames[j][2])
        return v
    a, b, c = run([0, 1, 2]), run([2, 0, 1]), run([0, 1, 2, 1])
    ta, tb, tc = a.get_volume()[0], b.get_volume()[0], c.get_volume()[0]
    assert ta.shape == (256, 256, 32) and (ta != 255).mean() > 0.05
    assert np.array_equal(np.abs(ta), np.abs(tb))          # min-|distance| merge: order can only matter for +-ties
    assert np.array_equal(ta, tc)
    assert np.array_equal(a.get_weight() + (c.get_weight() - a.get_weight()), c.get_weight())
    assert (c.get_weight() >= a.get_weight()).all()
