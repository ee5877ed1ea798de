"""Producer side of the feature pyramid: `DecoderSphere.get_sphere_feature`
(/root/reference/scenerf/models/unet2d_sphere.py:138-166) on the device, with the option of emitting the channels-last
layout the render path gathers from.  The convolutional part of the decoder (UpSampleBN, unet2d_sphere.py:12-41,168-206)
stays PyTorch's."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def sphere_dims(out_img_W: int, out_img_H: int, scale: int):
    lib = _lib.load()
    w, h = C.c_int(0), C.c_int(0)
    lib.srf_sphere_feature_dims(int(out_img_W), int(out_img_H), int(scale), C.byref(w), C.byref(h))
    return w.value, h.value


def get_sphere_feature(x: torch.Tensor, pix: torch.Tensor, pix_sphere: torch.Tensor, scale: int, out_img_W: int, out_img_H: int,
                       channels_last: bool = False) -> torch.Tensor:
    """x (B,C,h,w) float32 CUDA; pix (n,2) float32; pix_sphere (n,2) int64 -> (B,C,out_H,out_W) (or (B,out_H,out_W,C))."""
    if x.device.type != "cuda":
        raise RuntimeError("scenerf_b200.sphere_feature is the device path (no CPU fallback)")
    if x.requires_grad and torch.is_grad_enabled():
        # the reference method is differentiable through F.grid_sample into the decoder; this kernel has no backward --
        # used inside a training graph it would silently cut every gradient to the U-Net
        raise RuntimeError("scenerf_b200.get_sphere_feature is inference-only (no backward): call it under torch.no_grad() "
                           "or keep the reference's DecoderSphere.get_sphere_feature for training")
    lib = _lib.load()
    x = x.detach().to(torch.float32).contiguous()
    pix = pix.detach().to(device=x.device, dtype=torch.float32).contiguous()
    ps = pix_sphere.detach().to(device=x.device, dtype=torch.int64).contiguous()
    B, Cc, h, w = x.shape
    oW, oH = sphere_dims(out_img_W, out_img_H, scale)
    out = torch.empty((B, oH, oW, Cc) if channels_last else (B, Cc, oH, oW), dtype=torch.float32, device=x.device)
    ws = torch.empty(oW * oH, dtype=torch.int32, device=x.device)
    st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    for b in range(B):
        _lib.check(lib.srf_sphere_feature(x[b].data_ptr(), Cc, h, w, pix.data_ptr(), ps.data_ptr(), int(pix.shape[0]), int(scale),
                                          int(out_img_W), int(out_img_H), out[b].data_ptr(), 1 if channels_last else 0,
                                          ws.data_ptr(), ws.numel() * 4, st))
    return out
