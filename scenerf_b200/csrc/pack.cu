// One-time layout change of the feature pyramid: CHW planar fp32 (what UNet2DSphere emits,
// scenerf/models/unet2d_sphere.py:200-206) -> HWC channels-last fp32, so that each bilinear tap of
// utils.py:239-245 reads C contiguous floats instead of C planes.
#include <cuda_fp16.h>
#include "kernels.cuh"

namespace srf {

// tiled transpose: src [C][HW] -> dst [HW][C]
template <typename T>
__global__ void chw_to_hwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int C, int HW) {
  __shared__ float tile[32][33];
  const int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, hw = hw0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && hw < HW) ? src[(size_t)c * HW + hw] : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int hw = hw0 + i, c = c0 + threadIdx.x;
    if (hw < HW && c < C) {
      if constexpr (sizeof(T) == 2) dst[(size_t)hw * C + c] = __float2half_rn(tile[threadIdx.x][i]);
      else dst[(size_t)hw * C + c] = tile[threadIdx.x][i];
    }
  }
}

void launch_chw_to_hwc(const float* src, void* dst, int C, int H, int W, bool fp16, cudaStream_t st) {
  const int HW = H * W;
  dim3 grid((HW + 31) / 32, (C + 31) / 32), block(32, 8);
  if (fp16) chw_to_hwc_kernel<__half><<<grid, block, 0, st>>>(src, reinterpret_cast<__half*>(dst), C, HW);
  else chw_to_hwc_kernel<float><<<grid, block, 0, st>>>(src, reinterpret_cast<float*>(dst), C, HW);
}

}  // namespace srf
