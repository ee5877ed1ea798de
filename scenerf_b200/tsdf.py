"""Device-side TSDF fusion with the interface of the reference's `TSDFVolume`
(/root/reference/scenerf/data/utils/fusion.py:20-58 constructor, :219-324 integrate, :326-330 get_volume), so that the
scene-reconstruction script (scripts/reconstruction/depth2tsdf.py:87-103) can consume rendered depth / colour tensors
straight from device memory instead of the .npy / .png round trip.  Semantics: the reference's CPU (numba) path.
Marching cubes / mesh export (fusion.py:332-380, skimage) stay with the caller: `get_volume()` returns numpy arrays."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class TSDFVolume:
    def __init__(self, vol_bnds, voxel_size, trunc_margin=10, use_gpu=True, device="cuda:0"):
        vol_bnds = np.asarray(vol_bnds, dtype=np.float64).copy()
        assert vol_bnds.shape == (3, 2), "[!] `vol_bnds` should be of shape (3, 2)."
        if not use_gpu or not torch.cuda.is_available():
            raise RuntimeError("scenerf_b200.tsdf.TSDFVolume is the device path (no CPU fallback)")
        self.lib = _lib.load()
        self.device = torch.device(device)
        self._voxel_size = float(voxel_size)
        self._trunc_margin = trunc_margin
        self._vol_dim = np.ceil((vol_bnds[:, 1] - vol_bnds[:, 0]) / self._voxel_size).copy(order="C").astype(int)
        vol_bnds[:, 1] = vol_bnds[:, 0] + self._vol_dim * self._voxel_size
        self._vol_bnds = vol_bnds
        self._vol_origin = vol_bnds[:, 0].copy(order="C").astype(np.float32)
        shape = tuple(int(d) for d in self._vol_dim)
        self._tsdf = torch.empty(shape, dtype=torch.float32, device=self.device)
        self._weight = torch.empty(shape, dtype=torch.float32, device=self.device)
        self._color = torch.empty(shape, dtype=torch.float32, device=self.device)
        self._dims = (C.c_int * 3)(*shape)
        _lib.check(self.lib.srf_tsdf_reset(self._tsdf.data_ptr(), self._weight.data_ptr(), self._color.data_ptr(), self._dims,
                                           C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def integrate(self, color_im, depth_im, cam_intr, cam_pose, obs_weight=1.):
        """color_im (H,W,3), depth_im (H,W): numpy arrays or torch tensors (device tensors are used in place);
        cam_intr (3,3), cam_pose (4,4): numpy / tensors (tiny, read on the host like the reference does)."""
        depth = torch.as_tensor(depth_im).to(device=self.device, dtype=torch.float32).contiguous()
        col = torch.as_tensor(color_im)
        is_u8 = col.dtype == torch.uint8
        col = col.to(device=self.device, dtype=torch.uint8 if is_u8 else torch.float32).contiguous()
        im_h, im_w = depth.shape
        if tuple(col.shape) != (im_h, im_w, 3):
            raise ValueError("color_im must be (H,W,3) matching depth_im, got %s" % (tuple(col.shape),))
        pose = np.asarray(torch.as_tensor(cam_pose).detach().cpu().numpy(), dtype=np.float64)
        inv_pose = np.ascontiguousarray(np.linalg.inv(pose))           # float64, as fusion.py:265
        intr = np.ascontiguousarray(np.asarray(torch.as_tensor(cam_intr).detach().cpu().numpy()).astype(np.float32))
        origin = (C.c_float * 3)(*[float(v) for v in self._vol_origin])
        _lib.check(self.lib.srf_tsdf_integrate(
            self._tsdf.data_ptr(), self._weight.data_ptr(), self._color.data_ptr(), self._dims, origin, self._voxel_size,
            inv_pose.ctypes.data_as(C.POINTER(C.c_double)), intr.ctypes.data_as(C.POINTER(C.c_float)), depth.data_ptr(),
            col.data_ptr(), 1 if is_u8 else 0, int(im_h), int(im_w), float(self._trunc_margin), float(obs_weight),
            C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def merge_(self, tsdf, weight, color):
        """Fold another volume's state (later observations) into this one: include/scenerf_b200.h srf_tsdf_merge."""
        t, w, c = (x.to(device=self.device, dtype=torch.float32).contiguous() for x in (tsdf, weight, color))
        if tuple(t.shape) != tuple(self._tsdf.shape):
            raise ValueError("volume shape mismatch %s vs %s" % (tuple(t.shape), tuple(self._tsdf.shape)))
        _lib.check(self.lib.srf_tsdf_merge(self._tsdf.data_ptr(), self._weight.data_ptr(), self._color.data_ptr(), t.data_ptr(),
                                           w.data_ptr(), c.data_ptr(), self._dims,
                                           C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return self

    def get_volume(self):
        return self._tsdf.cpu().numpy(), self._color.cpu().numpy()

    def get_weight(self):
        return self._weight.cpu().numpy()
