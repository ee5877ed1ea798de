// TSDF integration of rendered depth / colour sweeps on the device ("next" row 8f-2 of the hot-path contract): the
// consumer of the novel-depth renders in the reference's scene-reconstruction pipeline
// (scripts/reconstruction/generate_novel_depths.py -> depth2tsdf.py:87-103 -> TSDFVolume.integrate).
//
// Semantics = the reference's CPU / numba path, fusion.py:219-324 (NOT its PyCUDA kernel, which implements a different
// rule): voxel centre float32(origin + size*index) (:181-184); world->camera and projection in float64 (:265,:382-387,
// :196-197); pixel = round-half-even(x*fx/z + cx); valid when inside the image, z > 0, depth > 0 and
// depth - z >= -trunc; keep the observation with the smaller |distance| (:212-216) together with its folded colour;
// weight += obs_weight.  One thread per voxel, float64 only for the 12 multiply-adds of the projection: HBM-bound
// (12 B read + 12 B written per touched voxel, 4 B of depth + 12 B of colour per hit).
#include "kernels.cuh"

namespace srf {

struct TsdfParams {
  int dx, dy, dz;             // volume dims, C-order [x][y][z]
  float origin[3];
  double voxel_size;
  double inv_pose[12];        // first 3 rows of inverse(cam_pose), float64
  double fx, fy, cx, cy;      // float32 intrinsics promoted to float64 (fusion.py:192-197)
  int im_h, im_w;
  double trunc;
  float obs_weight;
  int color_is_u8;            // colour image layout: (H,W,3) float32 or uint8
};

__global__ void tsdf_integrate_kernel(const __grid_constant__ TsdfParams q, float* __restrict__ tsdf, float* __restrict__ weight,
                                      float* __restrict__ color, const float* __restrict__ depth,
                                      const void* __restrict__ color_im) {
  const long long n = (long long)q.dx * q.dy * q.dz;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int z = (int)(i % q.dz), y = (int)((i / q.dz) % q.dy), x = (int)(i / ((long long)q.dz * q.dy));
  // float32(origin + float64(size) * float32(index))
  const double wx = (double)(float)((double)q.origin[0] + q.voxel_size * (double)x);
  const double wy = (double)(float)((double)q.origin[1] + q.voxel_size * (double)y);
  const double wz = (double)(float)((double)q.origin[2] + q.voxel_size * (double)z);
  const double* M = q.inv_pose;
  const double cxm = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M[0], wx), __dmul_rn(M[1], wy)), __dmul_rn(M[2], wz)), M[3]);
  const double cym = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M[4], wx), __dmul_rn(M[5], wy)), __dmul_rn(M[6], wz)), M[7]);
  const double czm = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(M[8], wx), __dmul_rn(M[9], wy)), __dmul_rn(M[10], wz)), M[11]);
  if (!(czm > 0.0)) return;
  const double pxd = rint(__dadd_rn(__ddiv_rn(__dmul_rn(cxm, q.fx), czm), q.cx));
  const double pyd = rint(__dadd_rn(__ddiv_rn(__dmul_rn(cym, q.fy), czm), q.cy));
  if (!(pxd >= 0.0 && pxd < (double)q.im_w && pyd >= 0.0 && pyd < (double)q.im_h)) return;
  const int px = (int)pxd, py = (int)pyd;
  const double dval = (double)depth[(size_t)py * q.im_w + px];
  const double diff = dval - czm;
  if (!(dval > 0.0 && diff >= -q.trunc)) return;
  const float old = tsdf[i];
  weight[i] = weight[i] + q.obs_weight;
  if (!(fabs((double)old) < fabs(diff))) {
    tsdf[i] = (float)diff;
    float c0, c1, c2;
    if (q.color_is_u8) {
      const unsigned char* c = reinterpret_cast<const unsigned char*>(color_im) + ((size_t)py * q.im_w + px) * 3;
      c0 = c[0]; c1 = c[1]; c2 = c[2];
    } else {
      const float* c = reinterpret_cast<const float*>(color_im) + ((size_t)py * q.im_w + px) * 3;
      c0 = c[0]; c1 = c[1]; c2 = c[2];
    }
    // fusion.py:231-233: floor(c2*65536 + c1*256 + c0) in float32, left to right
    color[i] = floorf(__fadd_rn(__fadd_rn(__fmul_rn(c2, 65536.0f), __fmul_rn(c1, 256.0f)), c0));
  }
}

__global__ void tsdf_reset_kernel(float* tsdf, float* weight, float* color, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  tsdf[i] = 255.0f;           // fusion.py:55
  weight[i] = 0.0f;
  color[i] = 0.0f;
}

void launch_tsdf_reset(float* tsdf, float* weight, float* color, long long n, cudaStream_t st) {
  tsdf_reset_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tsdf, weight, color, n);
}

void launch_tsdf_integrate(const int* dims, const float* origin, double voxel_size, const double* inv_pose, const float* intr,
                           int im_h, int im_w, double trunc, float obs_weight, int color_is_u8, float* tsdf, float* weight,
                           float* color, const float* depth, const void* color_im, cudaStream_t st) {
  TsdfParams q;
  q.dx = dims[0]; q.dy = dims[1]; q.dz = dims[2];
  for (int k = 0; k < 3; ++k) q.origin[k] = origin[k];
  q.voxel_size = voxel_size;
  for (int k = 0; k < 12; ++k) q.inv_pose[k] = inv_pose[k];
  q.fx = (double)intr[0]; q.fy = (double)intr[4]; q.cx = (double)intr[2]; q.cy = (double)intr[5];
  q.im_h = im_h; q.im_w = im_w; q.trunc = trunc; q.obs_weight = obs_weight; q.color_is_u8 = color_is_u8;
  const long long n = (long long)q.dx * q.dy * q.dz;
  tsdf_integrate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(q, tsdf, weight, color, depth, color_im);
}

}  // namespace srf
