// Pre-projected latent table (SURVEY 7 hard part 3b): because SphericalMapping.from_pixels ROUNDS the sphere coordinates
// (scenerf/models/spherical_mapping.py:115), the 2480-channel latent a sample point gathers (scenerf.py:522-525,
// utils.py:232-247) depends only on its integer sphere pixel -- and so does every lin_z[b](z) of ResnetFC.forward
// (resnetfc.py:148-150).  The bilinear gather is linear, so with Q_{b,s}[texel] = W_z,b[:, scale s] . feat_s[texel]
//     lin_z[b](z)(pixel) - bias = sum_s sum_{4 taps} w_tap(pixel, s) * Q_{b,s}[tap]
// exactly in real arithmetic (rounding differs at the 1e-7 level: the 2480-long dot product is re-associated).
// Once per (image, network):  15 float32 GEMMs (feat_s [T_s x C_s] times W^T, 0.2 TFLOP in all, SIMT fp32 FMA) and one
// blend kernel per block write  table[(sy, sx)][b][512]  for every sphere pixel that can have a valid tap,
// sx in [0, W], sy in [0, H], plus one row for everything outside.  The cumulative bias of the block's epilogue is
// folded into every row (the outside row is that bias alone), so the epilogue loads no bias vector.  The point-MLP kernel (mlp_tc.cu) then
// replaces the three lin_z GEMM passes (70.5 % of the per-point FLOPs) by adding table rows in its E1 epilogues.
#include <cuda_fp16.h>
#include "kernels.cuh"

namespace srf {

constexpr int kPreBlocks = SRF_NUM_BLOCKS;          // 3
constexpr int kPreRowVals = kPreBlocks * kHidden;   // 1536 values per sphere pixel

size_t preproj_rows(int sphere_W, int sphere_H) { return (size_t)(sphere_W + 1) * (sphere_H + 1) + 1; }
size_t preproj_table_bytes(int sphere_W, int sphere_H, int fp16) {
  return preproj_rows(sphere_W, sphere_H) * kPreRowVals * (fp16 ? 2 : 4) + 256;
}
size_t preproj_workspace_bytes(const int* H, const int* W) {
  size_t b = 0;
  for (int s = 0; s < kScales; ++s) b += (((size_t)H[s] * W[s] * kHidden * sizeof(float)) + 255) & ~(size_t)255;
  return b + 256;
}

struct BlendArgs {
  const float* bias_a;          // cumulative bias of the block's E1 epilogue, c_b = bias_a + bias_b (mlp_tc.cu pack_header_kernel:
  const float* bias_b;          //   c_0 = lin_in.bias + lin_z.0.bias, c_b = blocks.(b-1).fc_1.bias + lin_z.b.bias), folded into every row
  const float* Q[kScales];      // [T_s][512] float32: W_z,b[:, scale s] . feat_s[texel]
  int b;                        // block 0..2
  int W1, H1;                   // sphere_W + 1, sphere_H + 1
  void* table;
  int fp16;
};

// one CTA (128 threads x 4 columns) per sphere pixel (sx, sy); same tap arithmetic and blend order as the gather of the
// dense path: ((q_nw*w_nw + q_ne*w_ne) + q_sw*w_sw) + q_se*w_se per scale, separate roundings; scales added 1/1 .. 1/16
__global__ void __launch_bounds__(128) preproj_blend_kernel(const __grid_constant__ DevParams p, const __grid_constant__ BlendArgs a) {
  const int row = blockIdx.x;                       // the last row (index W1*H1) is the "outside the grid" row: no taps
  const bool outside = row >= a.W1 * a.H1;
  const int sy = row / a.W1, sx = row - sy * a.W1;
  const int c4 = threadIdx.x;                       // columns 4*c4 .. 4*c4+3
  float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int s = 0; s < kScales; ++s) {
    if (outside) break;
    const Taps tp = scale_taps(p, s, sx, sy);
    if (!tp.any) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tp.off[t] >= 0) q = __ldg(reinterpret_cast<const float4*>(a.Q[s] + (size_t)(tp.off[t] / p.C[s]) * kHidden) + c4);
      const float w = tp.w[t];
      if (t == 0) { acc.x = fmul(q.x, w); acc.y = fmul(q.y, w); acc.z = fmul(q.z, w); acc.w = fmul(q.w, w); }
      else { acc.x = fadd(acc.x, fmul(q.x, w)); acc.y = fadd(acc.y, fmul(q.y, w)); acc.z = fadd(acc.z, fmul(q.z, w)); acc.w = fadd(acc.w, fmul(q.w, w)); }
    }
    tot.x = fadd(tot.x, acc.x); tot.y = fadd(tot.y, acc.y); tot.z = fadd(tot.z, acc.z); tot.w = fadd(tot.w, acc.w);
  }
  {
    const float4 ba = __ldg(reinterpret_cast<const float4*>(a.bias_a) + c4), bb = __ldg(reinterpret_cast<const float4*>(a.bias_b) + c4);
    tot.x = fadd(tot.x, fadd(ba.x, bb.x)); tot.y = fadd(tot.y, fadd(ba.y, bb.y));
    tot.z = fadd(tot.z, fadd(ba.z, bb.z)); tot.w = fadd(tot.w, fadd(ba.w, bb.w));
  }
  const size_t o = (size_t)row * kPreRowVals + (size_t)a.b * kHidden + 4 * c4;
  if (a.fp16) {
    __half2* dst = reinterpret_cast<__half2*>(reinterpret_cast<__half*>(a.table) + o);
    dst[0] = __floats2half2_rn(tot.x, tot.y);
    dst[1] = __floats2half2_rn(tot.z, tot.w);
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.table) + o) = tot;
  }
}

// p: DevParams of the image (fp32 HWC pyramid).  Returns the number of kernel launches, or -1 (workspace too small).
int run_preproject(const DevParams& p, const srf_mlp_weights& w, int fp16, void* table, size_t table_bytes, void* workspace,
                   size_t ws_bytes, cudaStream_t st) {
  if (p.feat_fp16) return -2;
  if (table_bytes < preproj_table_bytes(p.sphere_W, p.sphere_H, fp16)) return -1;
  if (ws_bytes < preproj_workspace_bytes(p.H, p.W)) return -1;
  int launches = 0;
  BlendArgs a;
  unsigned char* wp = reinterpret_cast<unsigned char*>(workspace);
  for (int s = 0; s < kScales; ++s) {
    a.Q[s] = reinterpret_cast<const float*>(wp);
    wp += (((size_t)p.H[s] * p.W[s] * kHidden * sizeof(float)) + 255) & ~(size_t)255;
  }
  a.W1 = p.sphere_W + 1; a.H1 = p.sphere_H + 1; a.table = table; a.fp16 = fp16;
  const size_t rows = preproj_rows(p.sphere_W, p.sphere_H);
  for (int b = 0; b < kPreBlocks; ++b) {
    for (int s = 0; s < kScales; ++s) {
      GemmArgs g;
      g.A = reinterpret_cast<const float*>(p.feat[s]); g.lda = p.C[s];
      g.B = w.lin_z_w[b] + p.ch_off[s]; g.ldb = w.d_latent; g.bt = true;
      g.C = const_cast<float*>(a.Q[s]); g.ldc = kHidden;
      g.M = p.H[s] * p.W[s]; g.N = kHidden; g.K = p.C[s];
      if (launch_gemm(g, st)) return -2;
      ++launches;
    }
    a.b = b;
    a.bias_a = b == 0 ? w.lin_in_b : w.fc1_b[b - 1];
    a.bias_b = w.lin_z_b[b];
    preproj_blend_kernel<<<(unsigned)rows, 128, 0, st>>>(p, a);
    ++launches;
  }
  return launches;
}

}  // namespace srf
