// 3x3 (dilated) convolution of the spherical decoder on tensor cores, channels-last in and out -- the producer tail of the feature
// pyramid.  Reference: scenerf/models/unet2d_sphere.py:9-57 (`BasicBlock`, `UpSampleBN`: Conv2d 3x3 with
// padding = dilation, BatchNorm2d in eval mode, LeakyReLU(0.01), residual add), applied five times by `DecoderSphere.forward`
// (:167-206); its outputs "1_1".."1_16" ARE the x_rgb pyramid the ray renderer gathers from (scenerf.py:522-525).
//
// Implicit GEMM, no im2col buffer:  out[p, co] = sum_{tap} sum_{ci} in[p + d*(tap - 1), ci] * W[tap][co][ci].
//   M tile = 128 consecutive pixels of one image row, N tile = 128 output channels, K loop = 9 taps x Cin/32 blocks.
//   A tiles come from a 3-D tensor map over the [H][W][C] input: box {32 ch, 128 px, 1 row} at (c0, x0 + dx, y + dy) --
//   TMA's out-of-bounds zero fill IS the convolution's zero padding (negative / too large coordinates), and the box lands
//   in shared memory as the same 128-row x 128-byte swizzled tile a plain GEMM would stage.
//   B tiles: 2-D map over the weights repacked [tap][co][ci] (ci padded to a multiple of 4), box {32, 128} at (c0, tap*Cout + n0).
//   tcgen05.mma kind::tf32 (fp32 operands read in place, 10-bit mantissa -- the regime of the reference's own cuDNN default
//   `allow_tf32=True` on Ampere-class GPUs), fp32 accumulator in TMEM, 3-stage mbarrier ring, one elected issuer thread.
//   The tensor core TRUNCATES the 13 low mantissa bits of what it reads; through the decoder's 35 chained convolutions that
//   bias compounds (1.3 % relative L2 on the finest map of the test network).  So every tensor that feeds a convolution is
//   stored already ROUNDED TO NEAREST tf32 (weights at pack time, the concat buffer, intermediate activations: `round_out`),
//   which makes the truncation exact: 0.2 % relative L2, 6.7x better, at no cost.  The pyramid maps themselves stay unrounded.
//   Epilogue (4 warps, tcgen05.ld): y = acc*scale[co] + shift[co] (conv bias + eval-mode BatchNorm folded on the host),
//   (+ residual[p, co]), LeakyReLU, store fp32 [H][W][C] and/or fp16 [H][W][C] -- i.e. straight into the packed pyramid
//   layout of srf_pyramid (no CHW -> HWC pass).
// Roofline: tensor (tf32 = half the kind::f16 rate); 2*9*Cin*Cout flops per pixel.
#include <cuda.h>
#include <cuda_fp16.h>
#include "kernels.cuh"

namespace srf {
namespace conv {

constexpr int kBM = 128, kBN = 128, kBK = 32, kStages = 3;
constexpr uint32_t kTileBytes = kBM * kBK * 4;            // 16 KB
constexpr int kThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err) {     // bounded: a protocol bug must not hang the GPU
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 2000000000LL) { if (err) atomicExch(err, (int)(0x43000000u | (bar & 0xFFFFFF))); __trap(); }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (same encoding as mlp_tc.cu: make_desc_sw128)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: c_format F32 (1) at [4,6), a/b format TF32 (2) at [7,10) / [10,13), K-major, N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

struct ConvArgs {
  int H, W, Cin, Cout;            // Cin as stored (channel stride of the input, multiple of 4; padded channels hold zeros)
  int dil;                        // dilation = padding
  const float* scale;             // (Cout) folded BatchNorm scale, or all ones
  const float* shift;             // (Cout) folded conv bias / BatchNorm shift
  const float* residual;          // [H][W][ld_res] or null
  int ld_res;
  float slope;                    // LeakyReLU negative slope; 1.0 = no activation
  int round_out;                  // store out32 rounded to the nearest tf32 (it feeds another convolution)
  float* out32; int ld32;         // [H][W][ld32] or null
  __half* out16; int ld16;        // [H][W][ld16] or null
  int* err;
};

// nearest value with a 10-bit mantissa (ties away from zero; finite inputs)
__device__ __forceinline__ float round_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

__global__ void __launch_bounds__(kThreads)
conv3x3_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ ConvArgs a) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;            // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t sA = base, sB = base + kStages * kTileBytes;
  const uint32_t bars = sB + kStages * kTileBytes;                        // full[kStages], empty[kStages], acc
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int xt = (a.W + kBM - 1) / kBM;
  const int y = blockIdx.y / xt, x0 = (blockIdx.y % xt) * kBM;
  const int n0 = blockIdx.x * kBN;
  const int kb = (a.Cin + kBK - 1) / kBK;                                 // channel blocks per tap
  const int nk = 9 * kb;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int s = 0; s < kStages; ++s) { mbar_init(bars + 8u * s, 1); mbar_init(bars + 8u * (kStages + s), 1); }
    mbar_init(bars + 8u * (2 * kStages), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      for (int j = 0; j < nk; ++j) {
        const int tap = j / kb, cb = j - tap * kb;
        const int dy = (tap / 3 - 1) * a.dil, dx = (tap % 3 - 1) * a.dil;
        const int s = j % kStages;
        mbar_wait(bars + 8u * (kStages + s), (((uint32_t)(j / kStages)) & 1u) ^ 1u, a.err);
        mbar_arrive_expect_tx(bars + 8u * s, 2 * kTileBytes);
        tma_load_3d(sA + s * kTileBytes, &tmA, cb * kBK, x0 + dx, y + dy, bars + 8u * s);      // OOB -> zeros = the conv padding
        tma_load_2d(sB + s * kTileBytes, &tmB, cb * kBK, tap * a.Cout + n0, bars + 8u * s);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int j = 0; j < nk; ++j) {
        const int s = j % kStages;
        mbar_wait(bars + 8u * s, ((uint32_t)(j / kStages)) & 1u, a.err);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int k4 = 0; k4 < kBK / 8; ++k4)
          umma_tf32(tmem, make_desc_sw128(sA + s * kTileBytes + k4 * 32), make_desc_sw128(sB + s * kTileBytes + k4 * 32), kIdesc,
                    (j > 0 || k4 > 0) ? 1u : 0u);
        umma_commit(bars + 8u * (kStages + s));
      }
      umma_commit(bars + 8u * (2 * kStages));
    }
  } else {
    const int q = warp & 3;                                 // TMEM lane quarter this warp may read
    const int px = x0 + q * 32 + lane;
    const size_t pix = (size_t)y * a.W + px;
    mbar_wait(bars + 8u * (2 * kStages), 0, a.err);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
    for (int c = 0; c < kBN / 32; ++c) {
      if (n0 + c * 32 >= a.Cout) break;                     // warp-uniform
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (px < a.W) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const int co = n0 + c * 32 + j4 * 4;
          if (co >= a.Cout) break;                          // Cout % 4 == 0
          const float4 sc = __ldg(reinterpret_cast<const float4*>(a.scale + co)), sh = __ldg(reinterpret_cast<const float4*>(a.shift + co));
          float4 o = make_float4(fmaf(__uint_as_float(v[j4 * 4]), sc.x, sh.x), fmaf(__uint_as_float(v[j4 * 4 + 1]), sc.y, sh.y),
                                 fmaf(__uint_as_float(v[j4 * 4 + 2]), sc.z, sh.z), fmaf(__uint_as_float(v[j4 * 4 + 3]), sc.w, sh.w));
          if (a.residual) {
            const float4 t = *reinterpret_cast<const float4*>(a.residual + pix * a.ld_res + co);
            o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
          }
          o.x = o.x > 0.f ? o.x : o.x * a.slope; o.y = o.y > 0.f ? o.y : o.y * a.slope;
          o.z = o.z > 0.f ? o.z : o.z * a.slope; o.w = o.w > 0.f ? o.w : o.w * a.slope;
          if (a.out16) {
            __half2* d = reinterpret_cast<__half2*>(a.out16 + pix * a.ld16 + co);
            d[0] = __floats2half2_rn(o.x, o.y);
            d[1] = __floats2half2_rn(o.z, o.w);
          }
          if (a.out32) {
            if (a.round_out) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
            *reinterpret_cast<float4*>(a.out32 + pix * a.ld32 + co) = o;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
  }
}

// UpSampleBN front end (unet2d_sphere.py:47-56): F.interpolate(x, size=(H, W), bilinear, align_corners=True) of the coarser
// map, concatenated in front of the skip map: out[y][x] = [ up(x)(Cx) | skip(Cs) | zero padding to ld ], all channels-last.
__global__ void upsample_concat_kernel(const float* __restrict__ x, int h, int w, int Cx, int ldx, const float* __restrict__ skip, int Cs,
                                       int lds, int H, int W, float* __restrict__ out, int ld) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)H * W * ld;
  if (idx >= total) return;
  const int c = (int)(idx % ld);
  const size_t pix = idx / ld;
  const int ox = (int)(pix % W), oy = (int)(pix / W);
  float v = 0.f;
  if (c < Cx) {
    // ATen area_pixel_compute_source_index, align_corners=True: src = dst * (in - 1) / (out - 1)  (scale computed in float)
    const float sy = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
    const float fy = __fmul_rn(sy, (float)oy), fx = __fmul_rn(sx, (float)ox);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = __fsub_rn(fy, (float)y0), lx = __fsub_rn(fx, (float)x0);
    const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
    const float v00 = x[((size_t)y0 * w + x0) * ldx + c], v01 = x[((size_t)y0 * w + x1) * ldx + c];
    const float v10 = x[((size_t)y1 * w + x0) * ldx + c], v11 = x[((size_t)y1 * w + x1) * ldx + c];
    // ATen upsample_bilinear2d: h0lambda * (w0lambda * v00 + w1lambda * v01) + h1lambda * (w0lambda * v10 + w1lambda * v11)
    v = __fadd_rn(__fmul_rn(hy, __fadd_rn(__fmul_rn(hx, v00), __fmul_rn(lx, v01))), __fmul_rn(ly, __fadd_rn(__fmul_rn(hx, v10), __fmul_rn(lx, v11))));
  } else if (c < Cx + Cs) {
    v = skip[pix * lds + (c - Cx)];
  }
  out[idx] = round_tf32(v);            // this buffer only feeds the level's first convolution
}

}  // namespace conv

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn conv_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int* g_conv_err = nullptr;
int conv_watchdog_flag() { return g_conv_err ? *reinterpret_cast<volatile int*>(g_conv_err) : 0; }

// in: [H][W][Cin] float32 (Cin = channel stride, multiple of 4, 16-byte aligned); w9: [9][Cout][Cin] float32 (tap = ky*3 + kx).
// Returns 0, -1 (shape / alignment not expressible as tensor maps), -2 (driver entry point missing).
int launch_conv3x3_tf32(const float* in, int H, int W, int Cin, const float* w9, int Cout, int dil, const float* scale, const float* shift,
                        const float* residual, int ld_res, float slope, int round_out, float* out32, int ld32, void* out16, int ld16,
                        cudaStream_t st) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (H < 1 || W < 1 || Cin < 4 || (Cin % 4) || Cout < 4 || (Cout % 4) || dil < 1 || !al16(in) || !al16(w9) || !al16(scale) || !al16(shift)) return -1;
  if ((out32 && (!al16(out32) || ld32 % 4)) || (out16 && ((reinterpret_cast<uintptr_t>(out16) & 7) || ld16 % 4)) || (residual && (!al16(residual) || ld_res % 4)))
    return -1;
  EncodeTiledFn fn = conv_encode_fn();
  if (!fn) return -2;
  CUtensorMap tmA, tmB;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t strides[2] = {(cuuint64_t)Cin * 4, (cuuint64_t)W * Cin * 4};
    const cuuint32_t box[3] = {(cuuint32_t)conv::kBK, (cuuint32_t)conv::kBM, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (fn(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -1;
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)9 * Cout};
    const cuuint64_t strides[1] = {(cuuint64_t)Cin * 4};
    const cuuint32_t box[2] = {(cuuint32_t)conv::kBK, (cuuint32_t)conv::kBN};
    const cuuint32_t estr[2] = {1, 1};
    if (fn(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(w9), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -1;
  }
  if (!g_conv_err) {
    int* h = nullptr;
    if (cudaHostAlloc(&h, sizeof(int), cudaHostAllocMapped) == cudaSuccess) { *h = 0; cudaHostGetDevicePointer(&g_conv_err, h, 0); }
  }
  static bool attr = false;
  const size_t smem = 2 * conv::kStages * conv::kTileBytes + 1024 + 256;
  if (!attr) { cudaFuncSetAttribute(conv::conv3x3_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  conv::ConvArgs a;
  a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.dil = dil; a.scale = scale; a.shift = shift; a.residual = residual; a.ld_res = ld_res;
  a.slope = slope; a.round_out = round_out; a.out32 = out32; a.ld32 = ld32; a.out16 = reinterpret_cast<__half*>(out16); a.ld16 = ld16; a.err = g_conv_err;
  const dim3 grid((Cout + conv::kBN - 1) / conv::kBN, (unsigned)(H * ((W + conv::kBM - 1) / conv::kBM)));
  conv::conv3x3_tf32_kernel<<<grid, conv::kThreads, smem, st>>>(tmA, tmB, a);
  return 0;
}

void launch_upsample_concat(const float* x, int h, int w, int Cx, int ldx, const float* skip, int Cs, int lds, int H, int W, float* out, int ld,
                            cudaStream_t st) {
  const size_t total = (size_t)H * W * ld;
  conv::upsample_concat_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, h, w, Cx, ldx, skip, Cs, lds, H, W, out, ld);
}

}  // namespace srf
