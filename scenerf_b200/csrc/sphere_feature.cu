// Producer side of the feature pyramid ("next" row 8f-3, first piece): DecoderSphere.get_sphere_feature
// (scenerf/models/unet2d_sphere.py:138-166).  An image-plane feature map (C,h,w) is resampled onto the sphere grid:
//   1. every image pixel i writes its (floor(px/scale), floor(py/scale)) into the sphere cell
//      (clamp(round(sx_i/scale)), clamp(round(sy_i/scale))) of a (out_W,out_H) table initialised to -10; where several
//      pixels hit one cell the reference's index_put_ keeps the last one in index order (CPU, one thread) -- here the
//      winner is found with atomicMax over the pixel index, so the result is deterministic and equal to that rule;
//   2. grid_sample(bilinear, zeros, align_corners=False) of x at table/(w,h)*2-1 (ATen CPU arithmetic); untouched cells
//      (-10) sample zero padding.
// Output (C,out_H,out_W) like the reference, or channels-last (out_H,out_W,C) -- the layout the render path gathers from.
// HBM-bound: 4 taps x C x 4 B read (neighbouring cells share taps through L2) + C x 4 B written per sphere cell.
#include "kernels.cuh"

namespace srf {

__global__ void sphere_winner_kernel(const long long* __restrict__ pix_sphere, int n, int scale, int oW, int oH, int* __restrict__ winner) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // torch.round(pix_sphere / scale).long(): int64 -> float32 true division, half-to-even
  int cx = __float2int_rn(__fdiv_rn((float)pix_sphere[2 * (size_t)i + 0], (float)scale));
  int cy = __float2int_rn(__fdiv_rn((float)pix_sphere[2 * (size_t)i + 1], (float)scale));
  cx = min(max(cx, 0), oW - 1);
  cy = min(max(cy, 0), oH - 1);
  atomicMax(&winner[cx * oH + cy], i);
}

template <bool HWC>
__global__ void sphere_resample_kernel(const float* __restrict__ x, int C, int h, int w, const float* __restrict__ pix, int scale,
                                       int oW, int oH, const int* __restrict__ winner, float* __restrict__ out) {
  // one thread per (cell, channel); cells ordered oy-major so that CHW stores coalesce along ox
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long cells = (long long)oW * oH;
  if (t >= cells * C) return;
  int c, ox, oy;
  if (HWC) { c = (int)(t % C); const long long cell = t / C; ox = (int)(cell % oW); oy = (int)(cell / oW); }
  else { ox = (int)(t % oW); oy = (int)((t / oW) % oH); c = (int)(t / cells); }
  const int win = winner[ox * oH + oy];
  float mx = -10.0f, my = -10.0f;
  if (win >= 0) {
    mx = floorf(__fdiv_rn(pix[2 * (size_t)win + 0], (float)scale));      // pix // scale
    my = floorf(__fdiv_rn(pix[2 * (size_t)win + 1], (float)scale));
  }
  const float gx = fsub(fmul(fdiv(mx, (float)w), 2.0f), 1.0f);
  const float gy = fsub(fmul(fdiv(my, (float)h), 2.0f), 1.0f);
  const float ix = fsub(fmul(fadd(gx, 1.0f), (float)(w / 2.0)), 0.5f);
  const float iy = fsub(fmul(fadd(gy, 1.0f), (float)(h / 2.0)), 0.5f);
  const float xw = floorf(ix), yn = floorf(iy);
  const float fw = fsub(ix, xw), fe = fsub(1.0f, fw), fn = fsub(iy, yn), fs = fsub(1.0f, fn);
  const int x0 = (int)xw, y0 = (int)yn;
  const float wt[4] = {fmul(fs, fe), fmul(fs, fw), fmul(fn, fe), fmul(fn, fw)};
  const float* plane = x + (size_t)c * h * w;
  float acc = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int xx = x0 + (k & 1), yy = y0 + (k >> 1);
    const float v = (xx >= 0 && xx < w && yy >= 0 && yy < h) ? __ldg(plane + (size_t)yy * w + xx) : 0.0f;
    acc = fadd(acc, fmul(v, wt[k]));
  }
  if (HWC) out[((size_t)oy * oW + ox) * C + c] = acc;
  else out[((size_t)c * oH + oy) * oW + ox] = acc;
}

void launch_sphere_feature(const float* x, int C, int h, int w, const float* pix, const long long* pix_sphere, int n, int scale,
                           int oW, int oH, int* winner, float* out, int out_hwc, cudaStream_t st) {
  cudaMemsetAsync(winner, 0xFF, (size_t)oW * oH * sizeof(int), st);        // -1
  sphere_winner_kernel<<<(n + 255) / 256, 256, 0, st>>>(pix_sphere, n, scale, oW, oH, winner);
  const long long total = (long long)oW * oH * C;
  if (out_hwc) sphere_resample_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, C, h, w, pix, scale, oW, oH, winner, out);
  else sphere_resample_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, C, h, w, pix, scale, oW, oH, winner, out);
}

}  // namespace srf
