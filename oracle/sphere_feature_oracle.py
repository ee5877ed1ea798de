"""TEST INFRASTRUCTURE ONLY: numpy restatement of DecoderSphere.get_sphere_feature
(/root/reference/scenerf/models/unet2d_sphere.py:138-166) -- the producer side of the feature pyramid: an image-plane
feature map (C,h,w) is resampled onto the (out_H,out_W) sphere grid through the pixel -> sphere-pixel table of
SphericalMapping.from_pixels.  Pinned by tests/golden/sphere_feature.npz (outputs of the reference method)."""
import numpy as np

f32 = np.float32


def py_round(x):
    return int(round(x))                      # python round(): half to even, as unet2d_sphere.py:139


def get_sphere_feature(x, pix, pix_sphere, scale, out_img_W, out_img_H):
    """x (C,h,w) float32; pix (n,2) float32 image pixels; pix_sphere (n,2) int64 -> (C,out_H,out_W)."""
    C, h, w = x.shape
    oW, oH = py_round(out_img_W / scale), py_round(out_img_H / scale)
    ps = np.rint(pix_sphere.astype(f32) / f32(scale)).astype(np.int64)           # torch.round(pix_sphere / scale)
    ps[:, 0] = np.clip(ps[:, 0], 0, oW - 1)
    ps[:, 1] = np.clip(ps[:, 1], 0, oH - 1)
    p = np.floor(pix.astype(f32) / f32(scale)).astype(f32)                       # pix // scale
    m = np.full((oW, oH, 2), -10.0, dtype=f32)
    m[ps[:, 0], ps[:, 1]] = p                                                    # numpy fancy assignment: last wins
    m = m.reshape(-1, 2)
    gx = ((m[:, 0] / f32(w)).astype(f32) * f32(2) - f32(1)).astype(f32)
    gy = ((m[:, 1] / f32(h)).astype(f32) * f32(2) - f32(1)).astype(f32)
    ix = ((gx + f32(1)) * f32(w / 2.0) - f32(0.5)).astype(f32)
    iy = ((gy + f32(1)) * f32(h / 2.0) - f32(0.5)).astype(f32)
    x_w, y_n = np.floor(ix), np.floor(iy)
    ww = (ix - x_w).astype(f32)
    e = (f32(1) - ww).astype(f32)
    n = (iy - y_n).astype(f32)
    s = (f32(1) - n).astype(f32)
    x0, y0 = x_w.astype(np.int64), y_n.astype(np.int64)
    out = np.zeros((m.shape[0], C), dtype=f32)
    flat = x.reshape(C, h * w)
    for dx, dy, wt in ((0, 0, s * e), (1, 0, s * ww), (0, 1, n * e), (1, 1, n * ww)):
        xx, yy = x0 + dx, y0 + dy
        ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
        idx = np.where(ok, yy * w + xx, 0)
        out = (out + flat[:, idx].T * ok[:, None].astype(f32) * wt.astype(f32)[:, None]).astype(f32)
    return out.reshape(oW, oH, C).transpose(2, 1, 0).copy()
