"""Producer tail of the pyramid (hot-path contract row 8f-3): `DecoderSphere.forward` (unet2d_sphere.py:167-206).
Golden = the reference module's own outputs on deterministic weights (tests/golden/decoder_sphere.npz)."""
import numpy as np
import pytest

from cases import load_golden
from scenerf_b200 import synth

CASE = dict(num_features=128, bottleneck=48, W=128, H=64, oW=150, oH=46, seed=21)      # == make_goldens.DECODER_CASE


def inputs():
    chans = {1: 3, 2: 32, 4: 48, 8: 80, 16: 224, 32: CASE["bottleneck"]}
    feats = {}
    for s, ch in chans.items():
        h, w = -(-CASE["H"] // s), -(-CASE["W"] // s)
        feats[s] = synth.hash_normalish(500 + s, ch * h * w).reshape(ch, h, w).astype(np.float32)
    return feats


def test_decoder_oracle_vs_reference():
    from oracle.decoder_oracle import decoder_forward
    g = load_golden("decoder_sphere")
    p = synth.make_decoder_params(CASE["num_features"], CASE["bottleneck"], CASE["seed"])
    out = decoder_forward(p, inputs(), g["pix"], g["pix_sphere"], CASE["oW"], CASE["oH"])
    for k in ("1_1", "1_2", "1_4", "1_8", "1_16"):
        assert out[k].shape == g[k].shape, k
        err = float(np.abs(out[k] - g[k]).max())
        assert err <= 2e-5 * max(1.0, float(np.abs(g[k]).max())), (k, err)
    assert g["1_1"].shape == (4, 46, 150) and g["1_16"].shape == (64, 3, 9)


@pytest.mark.gpu
def test_decoder_cuda_vs_reference_and_feeds_the_renderer():
    import torch
    from scenerf_b200.decoder import SphereDecoderB200
    g = load_golden("decoder_sphere")
    p = {k: torch.from_numpy(v) for k, v in synth.make_decoder_params(CASE["num_features"], CASE["bottleneck"], CASE["seed"]).items()}
    dec = SphereDecoderB200(p, CASE["oW"], CASE["oH"], device="cuda:0")
    feats = inputs()
    features = [None] * 12
    for idx, s in ((0, 1), (4, 2), (5, 4), (6, 8), (8, 16), (11, 32)):
        features[idx] = torch.from_numpy(feats[s])[None].cuda()
    pyr = dec(features, torch.from_numpy(g["pix"]).cuda(), torch.from_numpy(g["pix_sphere"]).cuda())
    torch.cuda.synchronize()
    assert dec.launches == 5 * 8                                          # per level: upsample+concat and 7 convolutions
    x_rgb = pyr.as_x_rgb()
    for i, k in enumerate(("1_1", "1_2", "1_4", "1_8", "1_16")):
        got, ref = x_rgb[k].cpu().numpy(), g[k]
        assert got.shape == ref.shape, k
        err, mag = float(np.abs(got - ref).max()), float(np.abs(ref).max())
        rel_l2 = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
        print("%s: max-abs-err %.3e (max |ref| %.3e), rel L2 %.2e" % (k, err, mag, rel_l2))
        # tf32 operands (10-bit mantissa, rounded to nearest) through up to 35 chained convolutions: stated tolerance 6e-3 of the
        # map's magnitude, 4e-3 in L2 (a numpy emulation of RN-tf32 operands gives 2.4e-3 / 1.9e-3 on the finest map)
        assert err <= 6e-3 * max(1.0, mag) and rel_l2 <= 4e-3, k
        h16 = pyr.view16(i).float().permute(2, 0, 1).cpu().numpy()
        assert np.abs(h16 - got).max() <= 2e-3 * max(1.0, mag)            # the fp16 copy is the rounded fp32 map


@pytest.mark.gpu
def test_packed_pyramid_renders_like_the_chw_dict():
    """A PackedPyramid handed to the renderer (no srf_pack_pyramid pass) gives bit-identical results to the same values
    passed as the reference's dict of CHW tensors."""
    import torch
    from cases import RENDER_CASES
    from helpers import make_renderer, torch_pyramid
    from scenerf_b200.decoder import PackedPyramid
    cfg, seed = RENDER_CASES["kitti_mini"]
    g = load_golden("kitti_mini")
    x_rgb = torch_pyramid(cfg, seed)
    shapes = [tuple(x_rgb[k].shape) for k in synth.SCALE_KEYS]
    pp = PackedPyramid(shapes, torch.device("cuda:0"), True)
    for i, k in enumerate(synth.SCALE_KEYS):
        pp.view32(i).copy_(x_rgb[k].permute(1, 2, 0))
        pp.view16(i).copy_(x_rgb[k].permute(1, 2, 0).half())
    K, T = torch.from_numpy(cfg.K), torch.from_numpy(cfg.T)
    noise = (torch.from_numpy(g["noise_u"]), torch.from_numpy(g["noise_n"]))
    for prec in ("fp32tc", "fp16", "fp32"):
        a = make_renderer(cfg, prec).render_rays_batch(K, T, x_rgb, sampled_pixels=torch.from_numpy(g["pixels"]), noise=noise)
        b = make_renderer(cfg, prec).render_rays_batch(K, T, pp, sampled_pixels=torch.from_numpy(g["pixels"]), noise=noise)
        for k in ("depth", "color", "alphas", "loss_kl"):
            assert torch.equal(a[k], b[k]), (prec, k)
