"""get_sphere_feature (hot-path contract row 8f-3, producer side): golden = outputs of the reference's own method."""
import numpy as np
import pytest

from cases import load_golden


@pytest.mark.parametrize("scale", [1, 2, 4])
def test_sphere_feature_oracle(scale):
    from oracle.sphere_feature_oracle import get_sphere_feature
    g = load_golden("sphere_feature")
    W, H, oW, oH = g["dims"]
    o = get_sphere_feature(g["x_%d" % scale], g["pix"], g["pix_sphere"], scale, oW, oH)
    assert o.shape == g["feat_%d" % scale].shape and (g["feat_%d" % scale] != 0).mean() > 0.3
    assert np.abs(o - g["feat_%d" % scale]).max() <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("scale", [1, 2, 4])
def test_sphere_feature_cuda(scale):
    import torch
    from oracle.sphere_feature_oracle import get_sphere_feature as oracle_fn
    from scenerf_b200.sphere_feature import get_sphere_feature, sphere_dims
    g = load_golden("sphere_feature")
    W, H, oW, oH = (int(v) for v in g["dims"])
    x = torch.from_numpy(g["x_%d" % scale])[None].cuda()
    pix, ps = torch.from_numpy(g["pix"]).cuda(), torch.from_numpy(g["pix_sphere"]).cuda()
    out = get_sphere_feature(x, pix, ps, scale, oW, oH)
    ref = g["feat_%d" % scale]
    assert tuple(out.shape[1:]) == ref.shape and sphere_dims(oW, oH, scale) == (ref.shape[2], ref.shape[1])
    assert np.abs(out[0].cpu().numpy() - ref).max() <= 1e-6
    assert np.array_equal(out[0].cpu().numpy(), oracle_fn(g["x_%d" % scale], g["pix"], g["pix_sphere"], scale, oW, oH))   # same op order
    hwc = get_sphere_feature(torch.cat([x, 2 * x]), pix, ps, scale, oW, oH, channels_last=True)
    assert torch.equal(hwc[0].permute(2, 0, 1), out[0]) and torch.equal(hwc[1], 2 * hwc[0])
