"""CPU oracle of the TSDF integration that consumes the rendered depth sweeps -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates the de-facto (CPU / numba) path of the reference's `TSDFVolume.integrate`
(/root/reference/scenerf/data/utils/fusion.py:219-324 with helpers :174-217, :382-387), as called by
scripts/reconstruction/depth2tsdf.py:87-103.  (The reference's inline PyCUDA kernel, fusion.py:72-145, implements a
DIFFERENT rule -- weighted running average of min(1, diff/trunc) -- and is not the one its published pipeline runs
without pycuda; SURVEY.md Appendix B.)  Pinned by tests/golden/tsdf_fusion.npz, produced by running the reference's own
class (tests/golden/make_goldens.py: tsdf_fusion).

Arithmetic notes that matter for parity: voxel centres are float32 (`origin + size*index` rounded to float32,
fusion.py:181-184); the world->camera transform and the projection run in float64 because `np.linalg.inv(cam_pose)` is
float64 (fusion.py:265, :382-387, :196-197); pixels are `np.round` (half-to-even) of `x*fx/z + cx`; the merge keeps the
observation of smaller |distance| (fusion.py:212-216) and the folded colour of that observation.
"""
import numpy as np


def vol_dims(vol_bnds, voxel_size):
    b = np.asarray(vol_bnds, dtype=np.float64)
    return np.ceil((b[:, 1] - b[:, 0]) / float(voxel_size)).astype(int)


def fold_color(color_im):
    """fusion.py:231-233: BGR-style fold into one float32 channel."""
    c = np.asarray(color_im).astype(np.float32)
    return np.floor(c[..., 2] * np.float32(256 * 256) + c[..., 1] * np.float32(256) + c[..., 0]).astype(np.float32)


class TSDFVolumeOracle:
    def __init__(self, vol_bnds, voxel_size, trunc_margin=10):
        self.voxel_size = float(voxel_size)
        self.trunc = trunc_margin
        self.dim = vol_dims(vol_bnds, voxel_size)
        self.origin = np.asarray(vol_bnds, dtype=np.float64)[:, 0].astype(np.float32)
        self.tsdf = np.zeros(self.dim, np.float32) + np.float32(255)
        self.weight = np.zeros(self.dim, np.float32)
        self.color = np.zeros(self.dim, np.float32)
        xv, yv, zv = np.meshgrid(range(self.dim[0]), range(self.dim[1]), range(self.dim[2]), indexing="ij")
        self.vox = np.stack([xv.reshape(-1), yv.reshape(-1), zv.reshape(-1)], axis=1).astype(int)

    def integrate(self, color_im, depth_im, cam_intr, cam_pose, obs_weight=1.0):
        im_h, im_w = depth_im.shape
        folded = fold_color(color_im)
        # fusion.py:181-184: float32(origin + float64(size) * float32(index))
        world = (self.origin[None, :].astype(np.float64) + self.voxel_size * self.vox.astype(np.float32).astype(np.float64)).astype(np.float32)
        inv_pose = np.linalg.inv(np.asarray(cam_pose, dtype=np.float64))
        h = np.hstack([world, np.ones((len(world), 1), np.float32)])
        cam = (inv_pose @ h.T.astype(np.float64)).T[:, :3]
        intr = np.asarray(cam_intr).astype(np.float32)
        fx, fy, cx, cy = (np.float64(intr[0, 0]), np.float64(intr[1, 1]), np.float64(intr[0, 2]), np.float64(intr[1, 2]))
        z = cam[:, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            px = np.round(cam[:, 0] * fx / z + cx)
            py = np.round(cam[:, 1] * fy / z + cy)
        ok = np.isfinite(px) & np.isfinite(py)
        pxi = np.where(ok, px, -1).astype(np.int64)
        pyi = np.where(ok, py, -1).astype(np.int64)
        valid_pix = ok & (pxi >= 0) & (pxi < im_w) & (pyi >= 0) & (pyi < im_h) & (z > 0)
        depth_val = np.zeros(len(world))
        depth_val[valid_pix] = depth_im[pyi[valid_pix], pxi[valid_pix]]
        diff = depth_val - z
        valid = (depth_val > 0) & (diff >= -self.trunc)
        vx, vy, vz = self.vox[valid, 0], self.vox[valid, 1], self.vox[valid, 2]
        old = self.tsdf[vx, vy, vz]
        dist = diff[valid]
        take = ~(np.abs(old) < np.abs(dist))                       # fusion.py:212-216
        self.weight[vx, vy, vz] = (self.weight[vx, vy, vz] + obs_weight).astype(np.float32)
        self.tsdf[vx, vy, vz] = np.where(take, dist, old).astype(np.float32)
        newc = folded[pyi[valid], pxi[valid]]
        self.color[vx, vy, vz] = np.where(take, newc, self.color[vx, vy, vz]).astype(np.float32)

    def get_volume(self):
        return self.tsdf, self.color
