"""TEST INFRASTRUCTURE ONLY: numpy restatement of the convolutional tail of the spherical decoder,
`DecoderSphere.forward` (/root/reference/scenerf/models/unet2d_sphere.py:167-206) with `UpSampleBN` (:37-57) and `BasicBlock`
(:9-34) in eval mode -- the producer of the x_rgb pyramid.  Pinned by tests/golden/decoder_sphere.npz (outputs of the reference
module on deterministic weights from scenerf_b200.synth.make_decoder_params)."""
import numpy as np

from .sphere_feature_oracle import get_sphere_feature

f32 = np.float32
BN_EPS = 1e-5            # nn.BatchNorm2d default
LRELU = 0.01             # nn.LeakyReLU default


def conv2d(x, w, b, dil=1, pad=None):
    """x (Cin,H,W), w (Cout,Cin,k,k), stride 1, padding = `pad` (default dil*(k//2)): float32 accumulation per tap."""
    cout, cin, k, _ = w.shape
    pad = dil * (k // 2) if pad is None else pad
    H, W = x.shape[1] + 2 * pad - dil * (k - 1), x.shape[2] + 2 * pad - dil * (k - 1)
    xp = np.zeros((cin, x.shape[1] + 2 * pad, x.shape[2] + 2 * pad), f32)
    xp[:, pad:pad + x.shape[1], pad:pad + x.shape[2]] = x
    out = np.zeros((cout, H, W), np.float64)
    for ky in range(k):
        for kx in range(k):
            win = xp[:, ky * dil:ky * dil + H, kx * dil:kx * dil + W]
            out += np.einsum("oc,chw->ohw", w[:, :, ky, kx].astype(np.float64), win.astype(np.float64))
    return (out + b.astype(np.float64)[:, None, None]).astype(f32)


def bn_eval(x, p, name):
    s = p[name + ".weight"] / np.sqrt(p[name + ".running_var"] + f32(BN_EPS))
    return ((x - p[name + ".running_mean"][:, None, None]) * s[:, None, None] + p[name + ".bias"][:, None, None]).astype(f32)


def lrelu(x):
    return np.where(x > 0, x, x * f32(LRELU)).astype(f32)


def basic_block(x, p, pre, dil):
    y = lrelu(bn_eval(conv2d(x, p[pre + "conv_block1.0.weight"], p[pre + "conv_block1.0.bias"], dil), p, pre + "conv_block1.1"))
    y = bn_eval(conv2d(y, p[pre + "conv_block2.0.weight"], p[pre + "conv_block2.0.bias"], dil), p, pre + "conv_block2.1")
    return lrelu(y + x)


def upsample_bilinear_ac(x, H, W):
    """F.interpolate(x, size=(H,W), mode='bilinear', align_corners=True) for x (C,h,w)."""
    C, h, w = x.shape
    sy = f32(h - 1) / f32(H - 1) if H > 1 else f32(0)
    sx = f32(w - 1) / f32(W - 1) if W > 1 else f32(0)
    fy = (sy * np.arange(H, dtype=f32)).astype(f32)
    fx = (sx * np.arange(W, dtype=f32)).astype(f32)
    y0, x0 = fy.astype(np.int64), fx.astype(np.int64)
    y1, x1 = np.minimum(y0 + 1, h - 1), np.minimum(x0 + 1, w - 1)
    ly, lx = (fy - y0).astype(f32)[None, :, None], (fx - x0).astype(f32)[None, None, :]
    hy, hx = f32(1) - ly, f32(1) - lx
    v00, v01 = x[:, y0][:, :, x0], x[:, y0][:, :, x1]
    v10, v11 = x[:, y1][:, :, x0], x[:, y1][:, :, x1]
    return (hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11)).astype(f32)


def upsample_bn(x, skip, p, pre):
    f = np.concatenate([upsample_bilinear_ac(x, skip.shape[1], skip.shape[2]), skip], axis=0)
    y = conv2d(f, p[pre + "0.weight"], p[pre + "0.bias"], 1)
    for blk, dil in ((1, 1), (2, 2), (3, 3)):
        y = basic_block(y, p, pre + "%d." % blk, dil)
    return y


def decoder_forward(p, feats, pix, pix_sphere, out_img_W, out_img_H):
    """feats: dict scale -> (C,h,w) encoder map (scale 32 = the bottleneck before conv2); returns dict "1_s" -> (C,H_s,W_s)."""
    x32 = conv2d(feats[32], p["conv2.weight"], p["conv2.bias"], 1, pad=1)          # kernel 1, padding 1 (unet2d_sphere.py:79-81)
    sph = {32: get_sphere_feature(x32, pix, pix_sphere, 32, out_img_W, out_img_H)}
    for s in (16, 8, 4, 2, 1):
        sph[s] = get_sphere_feature(feats[s], pix, pix_sphere, s, out_img_W, out_img_H)
    out = {}
    x = sph[32]
    for s in (16, 8, 4, 2, 1):
        x = upsample_bn(x, sph[s], p, "up%d._net." % s)
        out["1_%d" % s] = x
    return out
