// Per-ray kernels around the point MLP: ray set-up + gaussian-proposal points, probabilistic sampling + sort,
// activation + alpha compositing + RaySOM.  One warp per ray; all float32.
//
// Reference being replaced (file:line relative to /root/reference/scenerf/models):
//   ray_setup_kernel      utils.py:177-182, utils.py:134-138,170 ; scenerf.py:549-572
//   sample_sort_kernel    scenerf.py:585-594 ; utils.py:75-90 ; utils.py:186-229 ; scenerf.py:636-659
//   composite_som_kernel  scenerf.py:533-536,473-481 ; scenerf.py:704-748 ; ray_som_kl.py:10-92
#include "kernels.cuh"

namespace srf {


// ---------------------------------------------------------------------------------------------------------------
__global__ void ray_setup_kernel(const __grid_constant__ DevParams p, const float* __restrict__ pixels, int R,
                                 float* __restrict__ unit_out, float* __restrict__ viewdir_out,
                                 float* __restrict__ gauss_pts) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float2 px = reinterpret_cast<const float2*>(pixels)[r];
  float d[3], u[3];
  pixel_direction(p, px.x, px.y, d, u);
  unit_out[r * 3 + 0] = u[0]; unit_out[r * 3 + 1] = u[1]; unit_out[r * 3 + 2] = u[2];
  // viewdir_infer = T[:3,:3] @ d   (un-normalised on purpose, utils.py:135,170)
  const float R3[9] = {p.T[0], p.T[1], p.T[2], p.T[4], p.T[5], p.T[6], p.T[8], p.T[9], p.T[10]};
  float vx, vy, vz;
  mat3_mul(R3, d[0], d[1], d[2], vx, vy, vz);
  viewdir_out[r * 3 + 0] = vx; viewdir_out[r * 3 + 1] = vy; viewdir_out[r * 3 + 2] = vz;
  for (int g = 0; g < p.G; ++g) {
    const float m0 = linspace_at(p.g_start, p.g_end, p.G, g);
    float x, y, z;
    rigid_transform(p.T, fmul(m0, u[0]), fmul(m0, u[1]), fmul(m0, u[2]), x, y, z);
    float* o = gauss_pts + ((size_t)r * p.G + g) * 3;
    o[0] = x; o[1] = y; o[2] = z;
  }
}

// ---------------------------------------------------------------------------------------------------------------
constexpr int kWarpsPerBlock = 4;
constexpr int kMaxS = 256;


__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sample_sort_kernel(const __grid_constant__ DevParams p, int R, const float* __restrict__ unit,
                   const float* __restrict__ gauss_raw,      // (R,G,2) mlp_gaussian output
                   const float* __restrict__ noise_u,        // (R,U) or null
                   const float* __restrict__ noise_n,        // (R,G*P) or null
                   float* __restrict__ means_out, float* __restrict__ stds_out,   // (R,G)
                   float* __restrict__ t_sorted,             // (R,S)
                   float* __restrict__ depth_volume,         // (R,S)
                   float* __restrict__ pts) {                // (R*S,3)
  __shared__ float keys[kWarpsPerBlock][kMaxS];
  __shared__ float gms[kWarpsPerBlock][2 * kMaxGaussians];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * kWarpsPerBlock + warp;
  if (r >= R) return;                                        // whole warp exits together
  float* k = keys[warp];
  float* gm = gms[warp];
  const int U = p.U, G = p.G, P = p.P, S = p.S;
  int n2 = 32;
  while (n2 < S) n2 <<= 1;

  // gaussian means / stds (scenerf.py:585-594)
  float mean_l = 0.f, std_l = 0.f;
  if (lane < G) {
    const float m0 = linspace_at(p.g_start, p.g_end, G, lane);
    const float o0 = gauss_raw[((size_t)r * G + lane) * 2 + 0];
    const float o1 = gauss_raw[((size_t)r * G + lane) * 2 + 1];
    mean_l = fadd(fmaxf(fadd(m0, o0), 0.0f), p.add_const);
    std_l = fadd(fmaxf(fadd(o1, p.base_std), 0.0f), p.add_const);
    means_out[(size_t)r * G + lane] = mean_l;
    stds_out[(size_t)r * G + lane] = std_l;
    gm[lane] = mean_l;
    gm[kMaxGaussians + lane] = std_l;
  }
  __syncwarp();
  // uniform samples (utils.py:75-90)
  for (int j = lane; j < U; j += 32) {
    const float u = noise_u ? noise_u[(size_t)r * U + j] : philox_uniform(p.seed, (uint32_t)r + p.ray0, (uint32_t)j, 1u);
    k[j] = fadd(linspace_at(0.2f, p.max_depth, U, j), fmul(u, p.uni_step));
  }
  // gaussian samples (utils.py:204-214)
  for (int j = lane; j < G * P; j += 32) {
    const int g = j / P;
    const float m = gm[g], s = gm[kMaxGaussians + g];
    const float e = noise_n ? noise_n[(size_t)r * G * P + j] : philox_normal(p.seed, (uint32_t)r + p.ray0, (uint32_t)j);
    float t = fadd(m, fmul(e, s));
    if (t < 0.1f) t = 0.1f;
    k[U + j] = t;
  }
  for (int j = S + lane; j < n2; j += 32) k[j] = __int_as_float(0x7f800000);   // +inf padding
  __syncwarp();
  // bitonic sort of n2 keys (scenerf.py:652-655; keys only: depth and points are recomputed from the sorted
  // distance with the same multiplications the reference applied before its gather, so payloads are bit-equal)
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = lane; i < (n2 >> 1); i += 32) {
        const int lo = ((i / stride) * stride * 2) + (i % stride);
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const float a = k[lo], b = k[hi];
        if ((a > b) == up) { k[lo] = b; k[hi] = a; }
      }
      __syncwarp();
    }
  }
  const float ux = unit[r * 3 + 0], uy = unit[r * 3 + 1], uz = unit[r * 3 + 2];
  for (int j = lane; j < S; j += 32) {
    const float t = k[j];
    const float cx = fmul(t, ux), cy = fmul(t, uy), cz = fmul(t, uz);
    t_sorted[(size_t)r * S + j] = t;
    depth_volume[(size_t)r * S + j] = cz;
    float x, y, z;
    rigid_transform(p.T, cx, cy, cz, x, y, z);
    float* o = pts + ((size_t)r * S + j) * 3;
    o[0] = x; o[1] = y; o[2] = z;
  }
}

// ---------------------------------------------------------------------------------------------------------------

struct CompositeSmem {
  float t[kMaxS], dv[kMaxS], sg[kMaxS], al[kMaxS], w[kMaxS], c0[kMaxS], c1[kMaxS], c2[kMaxS];
};

template <int kMaxG>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_som_kernel(const __grid_constant__ DevParams p, int R, const float* __restrict__ raw,   // (R*S,4)
                     const float* __restrict__ t_sorted, const float* __restrict__ depth_volume,
                     const float* __restrict__ means, const float* __restrict__ stds, srf_outputs out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CompositeSmem& sm = reinterpret_cast<CompositeSmem*>(smem_raw)[threadIdx.x >> 5];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * kWarpsPerBlock + warp;
  if (r >= R) return;
  const int S = p.S, G = p.G;
  const size_t base = (size_t)r * S;

  // ---- activation (scenerf.py:533-536) ----
  for (int j = lane; j < S; j += 32) {
    const float4 o = reinterpret_cast<const float4*>(raw)[base + j];
    sm.c0[j] = sigmoidf_ref(o.x); sm.c1[j] = sigmoidf_ref(o.y); sm.c2[j] = sigmoidf_ref(o.z);
    const float sg = softplusf_ref(fsub(o.w, 1.0f));
    sm.sg[j] = sg;
    sm.t[j] = fmaxf(t_sorted[base + j], 0.0f);              // scenerf.py:707
    sm.dv[j] = depth_volume[base + j];
    if (out.densities) out.densities[base + j] = sg;
  }
  __syncwarp();
  // ---- alphas (scenerf.py:708-711) ----
  for (int j = lane; j < S; j += 32) {
    const float delta = (j == 0) ? sm.t[0] : fsub(sm.t[j], sm.t[j - 1]);
    sm.al[j] = fsub(1.0f, expf(-fmul(delta, sm.sg[j])));
  }
  __syncwarp();
  // ---- transmittance: exclusive prefix product of (1 - alpha + 1e-10) (scenerf.py:718-723).
  //      Each lane owns a contiguous segment; segment products are scanned with warp shuffles. ----
  const int spt = (S + 31) >> 5;
  const int j0 = lane * spt, j1 = min(S, j0 + spt);
  float seg = 1.0f;
  for (int j = j0; j < j1; ++j) seg *= fadd(fsub(1.0f, sm.al[j]), 1e-10f);
  float incl = seg;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= v;
  }
  float Tacc = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) Tacc = 1.0f;
  float d_acc = 0.f, r_acc = 0.f, g_acc = 0.f, b_acc = 0.f;
  for (int j = j0; j < j1; ++j) {
    const float a = sm.al[j];
    const float w = fmul(a, Tacc);
    sm.w[j] = w;
    d_acc += w * sm.dv[j];
    r_acc += w * sm.c0[j]; g_acc += w * sm.c1[j]; b_acc += w * sm.c2[j];
    Tacc *= fadd(fsub(1.0f, a), 1e-10f);
  }
  const float depth = warp_sum(d_acc);
  r_acc = warp_sum(r_acc); g_acc = warp_sum(g_acc); b_acc = warp_sum(b_acc);
  __syncwarp();
  // ---- closest sample to the rendered depth (scenerf.py:730-735), first index wins ties ----
  float best = __int_as_float(0x7f800000);
  int best_j = 0x7fffffff;
  for (int j = lane; j < S; j += 32) {
    const float d = fabsf(fsub(depth, sm.dv[j]));
    if (d < best) { best = d; best_j = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
    if (ob < best || (ob == best && oj < best_j)) { best = ob; best_j = oj; }
  }
  if (lane == 0) {
    if (out.depth) out.depth[r] = depth;
    if (out.color) { out.color[r * 3 + 0] = r_acc; out.color[r * 3 + 1] = g_acc; out.color[r * 3 + 2] = b_acc; }
    if (out.closest_pts_to_depths) out.closest_pts_to_depths[r] = best;
    if (out.weights_at_depth) out.weights_at_depth[r] = sm.w[best_j];
  }
  for (int j = lane; j < S; j += 32) {
    if (out.alphas) out.alphas[base + j] = sm.al[j];
    if (out.weights) out.weights[base + j] = sm.w[j];
  }
  if (!(out.loss_kl || out.som_vars || out.som_means)) return;

  // ---- RaySOM (ray_som_kl.py:10-78) ----
  float m[kMaxG], sd[kMaxG], var[kMaxG], rel[kMaxG][kMaxG], pc[kMaxG][kMaxG];
#pragma unroll
  for (int g = 0; g < kMaxG; ++g) {
    m[g] = (g < G) ? means[(size_t)r * G + g] : 0.f;
    sd[g] = (g < G) ? stds[(size_t)r * G + g] : 1.f;
    var[g] = fmul(sd[g], sd[g]);
  }
#pragma unroll
  for (int c2 = 0; c2 < kMaxG; ++c2) {
    float s = 0.f;
#pragma unroll
    for (int c1 = 0; c1 < kMaxG; ++c1) {
      const float d = fsub(m[c2], m[c1]);
      rel[c2][c1] = (c1 < G && c2 < G) ? expf(fdiv(-fmul(d, d), p.two_sig2)) : 0.f;
      if (c1 < G) s = (c1 == 0) ? rel[c2][c1] : fadd(s, rel[c2][c1]);
    }
#pragma unroll
    for (int c1 = 0; c1 < kMaxG; ++c1) pc[c2][c1] = (c1 < G && c2 < G) ? fdiv(rel[c2][c1], s) : 0.f;
  }
  const float kSqrt2Pi = 2.50662827463100024f;
  float sw[kMaxG], swt[kMaxG];
#pragma unroll
  for (int g = 0; g < kMaxG; ++g) { sw[g] = 0.f; swt[g] = 0.f; }
  // pass 1: weights w[r][sample] (kept in smem: reuse c0.. arrays are still needed? no -> reuse c0,c1,c2,sg,dv,al? keep al)
  // we store per-sample the BMU index and p_best, recompute p(z|c1) in pass 2 (cheap) instead of storing G values.
  for (int j = lane; j < S; j += 32) {
    const float t = sm.t[j];
    const float dens = fadd(sm.al[j], 1e-8f);
    float pz[kMaxG];
#pragma unroll
    for (int g = 0; g < kMaxG; ++g) {
      const float d = fabsf(fsub(m[g], t));
      const float e = fadd(fdiv(expf(fdiv(-fmul(d, d), fmul(2.0f, var[g]))), fmul(kSqrt2Pi, sd[g])), 1e-5f);
      pz[g] = fadd(fmul(e, dens), 1e-8f);
    }
    float pbest = -1.f;
    int bi = 0;
#pragma unroll
    for (int c2 = 0; c2 < kMaxG; ++c2) {
      if (c2 < G) {
        float s = 0.f;
#pragma unroll
        for (int c1 = 0; c1 < kMaxG; ++c1)
          if (c1 < G) {
            const float term = fadd(fmul(pz[c1], pc[c2][c1]), 1e-8f);
            s = (c1 == 0) ? term : fadd(s, term);
          }
        if (s > pbest) { pbest = s; bi = c2; }
      }
    }
#pragma unroll
    for (int g = 0; g < kMaxG; ++g) {
      if (g < G) {
        float relw = 0.f;
#pragma unroll
        for (int c = 0; c < kMaxG; ++c) relw = (c == bi) ? rel[g][c] : relw;
        const float w = fadd(fdiv(fmul(relw, pz[g]), pbest), 1e-5f);
        sw[g] += w;
        swt[g] += w * t;
      }
    }
    sm.c0[j] = pbest;
    sm.c1[j] = __int_as_float(bi);
  }
  float nm[kMaxG], nv[kMaxG];
#pragma unroll
  for (int g = 0; g < kMaxG; ++g) {
    sw[g] = warp_sum(sw[g]);
    swt[g] = warp_sum(swt[g]);
    nm[g] = fdiv(swt[g], sw[g]);
    nv[g] = 0.f;
  }
  for (int j = lane; j < S; j += 32) {
    const float t = sm.t[j];
    const float dens = fadd(sm.al[j], 1e-8f);
    const float pbest = sm.c0[j];
    const int bi = __float_as_int(sm.c1[j]);
#pragma unroll
    for (int g = 0; g < kMaxG; ++g) {
      if (g < G) {
        const float d = fabsf(fsub(m[g], t));
        const float e = fadd(fdiv(expf(fdiv(-fmul(d, d), fmul(2.0f, var[g]))), fmul(kSqrt2Pi, sd[g])), 1e-5f);
        const float pz = fadd(fmul(e, dens), 1e-8f);
        float relw = 0.f;
#pragma unroll
        for (int c = 0; c < kMaxG; ++c) relw = (c == bi) ? rel[g][c] : relw;
        const float w = fadd(fdiv(fmul(relw, pz), pbest), 1e-5f);
        const float dd = fsub(t, nm[g]);
        nv[g] += w * fmul(dd, dd);
      }
    }
  }
  float kl_sum = 0.f;
#pragma unroll
  for (int g = 0; g < kMaxG; ++g) {
    nv[g] = fdiv(warp_sum(nv[g]), sw[g]);
    if (g < G) {
      const float mean_diff = fabsf(fsub(m[g], nm[g]));
      const float var_diff = fabsf(fsub(sqrtf(var[g]), sqrtf(nv[g])));
      const bool mask = (mean_diff > 0.1f) && (var_diff > 0.1f) && (nv[g] > 0.0f);
      float s2 = sqrtf(nv[g]);
      if (s2 < 1.5f) s2 = 1.5f;                                        // ray_som_kl.py:83
      const float std_err = logf(fadd(fdiv(s2, sd[g]), 1e-8f));
      const float dm = fsub(m[g], nm[g]);
      const float mean_err = fdiv(fadd(fmul(sd[g], sd[g]), fmul(dm, dm)), fmul(2.0f, fmul(s2, s2)));
      const float kl = fsub(fadd(std_err, mean_err), 0.5f);
      kl_sum = fadd(kl_sum, mask ? kl : 0.0f);
    }
  }
  if (lane == 0) {
    if (out.loss_kl) out.loss_kl[r] = fdiv(kl_sum, (float)G);
    for (int g = 0; g < G; ++g) {
      if (out.som_vars) out.som_vars[(size_t)r * G + g] = nv[g];
      if (out.som_means) out.som_means[(size_t)r * G + g] = nm[g];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
void launch_ray_setup(const DevParams& p, const float* pixels, int R, float* unit, float* viewdir, float* gauss_pts,
                      cudaStream_t st) {
  ray_setup_kernel<<<(R + 127) / 128, 128, 0, st>>>(p, pixels, R, unit, viewdir, gauss_pts);
}
void launch_sample_sort(const DevParams& p, int R, const float* unit, const float* gauss_raw, const float* noise_u,
                        const float* noise_n, float* means, float* stds, float* t_sorted, float* depth_volume,
                        float* pts, cudaStream_t st) {
  sample_sort_kernel<<<(R + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, st>>>(
      p, R, unit, gauss_raw, noise_u, noise_n, means, stds, t_sorted, depth_volume, pts);
}
void launch_composite_som(const DevParams& p, int R, const float* raw, const float* t_sorted,
                          const float* depth_volume, const float* means, const float* stds, const srf_outputs& out,
                          cudaStream_t st) {
  const size_t smem = sizeof(CompositeSmem) * kWarpsPerBlock;   // 32 KB: below the 48 KB default limit
  const dim3 grid((R + kWarpsPerBlock - 1) / kWarpsPerBlock), block(kWarpsPerBlock * 32);
  if (p.G <= 4)
    composite_som_kernel<4><<<grid, block, smem, st>>>(p, R, raw, t_sorted, depth_volume, means, stds, out);
  else
    composite_som_kernel<8><<<grid, block, smem, st>>>(p, R, raw, t_sorted, depth_volume, means, stds, out);
}

}  // namespace srf
