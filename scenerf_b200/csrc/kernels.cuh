// Host-side launchers of the scenerf_b200 kernels (definitions in the .cu files of this directory).
#pragma once
#include "common.cuh"

namespace srf {

// ray_kernels.cu
void launch_ray_setup(const DevParams& p, const float* pixels, int R, float* unit, float* viewdir, float* gauss_pts,
                      cudaStream_t st);
void launch_sample_sort(const DevParams& p, int R, const float* unit, const float* gauss_raw, const float* noise_u,
                        const float* noise_n, float* means, float* stds, float* t_sorted, float* depth_volume,
                        float* pts, cudaStream_t st);
void launch_composite_som(const DevParams& p, int R, const float* raw, const float* t_sorted,
                          const float* depth_volume, const float* means, const float* stds, const srf_outputs& out,
                          cudaStream_t st);

// pack.cu
void launch_chw_to_hwc(const float* src, void* dst, int C, int H, int W, bool fp16, cudaStream_t st);

// tsdf.cu : TSDF integration of rendered depth sweeps (reference: data/utils/fusion.py:219-324, CPU path)
void launch_tsdf_reset(float* tsdf, float* weight, float* color, long long n, cudaStream_t st);
void launch_tsdf_integrate(const int* dims, const float* origin, double voxel_size, const double* inv_pose, const float* intr,
                           int im_h, int im_w, double trunc, float obs_weight, int color_is_u8, float* tsdf, float* weight,
                           float* color, const float* depth, const void* color_im, cudaStream_t st);

// image_ops.cu : resampling of x-major renders into images, TSDF volume merge
void launch_upsample_render(const float* depth_xm, const float* color_xm, int gw, int gh, int H, int W, float* depth_out,
                            float* color_out, int color_mode, cudaStream_t st);
void launch_tsdf_merge(float* tsdf_a, float* weight_a, float* color_a, const float* tsdf_b, const float* weight_b,
                       const float* color_b, long long n, cudaStream_t st);

// mlp_simt.cu : float32 point MLP (gather + positional encoding + ResnetFC), n points in chunks.
//   pts (n,3) infer-frame points; viewdir (n/n_per,3); raw_out (n,d_out).  Returns number of kernel launches.
size_t simt_workspace_bytes(int d_latent, int n_points);
int run_point_mlp_simt(const DevParams& p, const srf_mlp_weights& w, const float* pts, const float* viewdir, int n,
                       int n_per, float* raw_out, int32_t* dbg_sphere, void* workspace, size_t ws_bytes,
                       cudaStream_t st);

// sphere_feature.cu : image-plane feature map -> sphere grid (unet2d_sphere.py:138-166)
void launch_sphere_feature(const float* x, int C, int h, int w, const float* pix, const long long* pix_sphere, int n, int scale,
                           int oW, int oH, int* winner, float* out, int out_hwc, cudaStream_t st);

// gemm.cu : float32 SIMT GEMM (see the file header for operand layouts and the epilogue)
struct GemmArgs {
  const float* A = nullptr; int lda = 0; bool at = false; bool relu_a = false;
  const float* B = nullptr; int ldb = 0; bool bt = false; bool relu_b = false;
  float* C = nullptr; int ldc = 0; int M = 0, N = 0, K = 0;
  const float* bias = nullptr; const float* mask = nullptr; int ldm = 0; const float* R = nullptr; int ldr = 0; int accumulate = 0;
  const int* skip_if_zero = nullptr;   // device flag: when *flag == 0 the launch does nothing (an all-zero K-segment of the latent)
  // tf32 kernel only: the latent axis (K when seg_mode == 1, N when seg_mode == 2) is the concatenation of the 5 pyramid
  // scales [seg_off[s], seg_off[s+1]); k-blocks / column tiles that lie entirely in scales with seg_flags[s] == 0 are
  // skipped on the device (their inputs are exact zeros / their outputs are never read).  One launch instead of five.
  const int* seg_flags = nullptr; int seg_mode = 0; int seg_off[6] = {0, 0, 0, 0, 0, 0};
  float* splitk_ws = nullptr; size_t splitk_ws_floats = 0;   // optional scratch: enables deterministic split-K for long-K, few-tile shapes
  float* relu_out = nullptr; int ld_relu = 0;   // tf32 kernel only: the epilogue also stores max(result, 0) here (the next GEMM's A operand)
};
int launch_gemm(const GemmArgs& g, cudaStream_t st);
// per-thread count of kernels launched by the float32 / tf32 MLP paths (gemm.cu, gemm_tf32.cu, mlp_simt.cu, backward.cu);
// the run_* entry points report the difference as their launch count (srf_last_launch_count)
int& launch_counter();    // 0, or -1 for an operand-layout combination that is not instantiated

// gemm_tf32.cu : the same contract on tensor cores (tcgen05 kind::tf32, float32 operands read in place); NT layout only
// (at=false, bt=true, no operand ReLU).  Returns 0, or -1 when the shape cannot be expressed as TMA tensor maps.
int launch_gemm_tf32(const GemmArgs& g, cudaStream_t st);
int tf32_watchdog_flag();

// one warp per point: X[i] = [ gathered latent (d_latent) | positional encoding (39) | viewdir (3) | 0-pad ], row stride ld
//   scale_any (5 ints, or NULL): set to 1 for every scale at which some point of the chunk has an in-range bilinear tap.
//   Scales whose flag stays 0 contribute exact zeros to x_in (quirk Q2: out-of-range normalised coordinates), so the
//   lin_z GEMMs skip their K-segment -- bit-identical results.
void launch_build_xin(const DevParams& p, const float* pts, const float* viewdir, int m, int n_per, int point0, float* X, int ld,
                      int32_t* dbg_sphere, int* scale_any, cudaStream_t st);

// out (M,d_out) = lin_out(relu(Hh))   (resnetfc.py:163; one warp per row)
void launch_lin_out(const float* Hh, const float* W, const float* bias, float* out, int M, int d_out, cudaStream_t st);

// backward.cu : float32 backward of the path (reference: torch.autograd through scenerf.py:392-748)
size_t mlp_backward_workspace_bytes(int d_latent, int n_points);
size_t mlp_saved_bytes(int d_latent, int n_points);            // activation store of one pass (SRF_FLAG_SAVE_ACTIVATIONS)
size_t mlp_forward_save_scratch_bytes(int n_points);           // scratch of run_point_mlp_forward_save (tf32 mode: two ReLU'd operand buffers)
int run_point_mlp_forward_save(const DevParams& p, const srf_mlp_weights& w, const float* pts, const float* viewdir, int n, int n_per,
                               float* raw_out, int32_t* dbg_sphere, void* saved_base, int tf32_matmul, void* scratch, size_t scratch_bytes,
                               cudaStream_t st);
int run_point_mlp_backward_simt(const DevParams& p, const srf_mlp_weights& w, const srf_mlp_weights& gw, float* const* grad_pyr_chw,
                                const float* pts, const float* viewdir, int n, int n_per, const float* g_raw, const void* saved_base,
                                int tf32_matmul, void* workspace, size_t ws_bytes, cudaStream_t st);
void launch_ray_backward(const DevParams& p, int R, const float* raw, const float* t_sorted, const float* unit,
                         const float* gauss_raw, const float* noise_n, const srf_outputs& fwd, const srf_outputs& cot,
                         float* graw_main, float* graw_gauss, cudaStream_t st);

// mlp_tc.cu : tcgen05 tensor-core point MLP.
//   split != 0: the fp32-grade layout (hi + lo fp16 images of W 2^s, see mlp_tc.cu) read through w.tc_split_packed
size_t tc_weights_bytes(int d_out, int d_latent, int split);
int pack_weights_tc(const srf_mlp_weights& w, void* dst, size_t bytes, int split, cudaStream_t st);
constexpr int kTcFlagPreproj = 1 << 29; // internal bit: this pass may use DevParams::preproj (the table of ITS network)
constexpr int kTcFlagSplit = 1 << 30;   // internal bit of the `flags` argument of run_point_mlp_tc*: split (fp32-grade) mode
size_t tc_workspace_bytes(int d_latent, int n_points);
int run_point_mlp_tc(const DevParams& p, const srf_mlp_weights& w, const float* pts, const float* viewdir, int n,
                     int n_per, float* raw_out, int32_t* dbg_sphere, int flags, void* workspace, size_t ws_bytes,
                     cudaStream_t st);

// conv_tf32.cu : the spherical decoder's 3x3 (dilated) convolutions as a tcgen05 kind::tf32 implicit GEMM on channels-last maps
//   (unet2d_sphere.py:9-57) + the UpSampleBN front end (bilinear align_corners=True upsample of the coarser map, concat with the skip map)
int launch_conv3x3_tf32(const float* in, int H, int W, int Cin, const float* w9, int Cout, int dil, const float* scale, const float* shift,
                        const float* residual, int ld_res, float slope, int round_out, float* out32, int ld32, void* out16, int ld16,
                        cudaStream_t st);
void launch_upsample_concat(const float* x, int h, int w, int Cx, int ldx, const float* skip, int Cs, int lds, int H, int W, float* out, int ld,
                            cudaStream_t st);
int conv_watchdog_flag();

// preproj.cu : pre-projected latent table  table[(sy,sx)][block][512] = lin_z[block].weight . z(sphere pixel)  (see file header)
size_t preproj_rows(int sphere_W, int sphere_H);
size_t preproj_table_bytes(int sphere_W, int sphere_H, int fp16);
size_t preproj_workspace_bytes(const int* H, const int* W);
int run_preproject(const DevParams& p, const srf_mlp_weights& w, int fp16, void* table, size_t table_bytes, void* workspace,
                   size_t ws_bytes, cudaStream_t st);

// non-zero after the kernel's mbarrier watchdog fired: 0x40000000 | warp<<24 | (barrier smem offset)<<4 | parity
int tc_watchdog_flag();
// diagnostic: stop every tile after `debug_layer` (1,2,4,5,7,8,9,10 -- see the tile program in mlp_tc.cu) and dump the
// raw fp32 accumulator (n_tiles*128, 512) to debug_acc
int run_point_mlp_tc_debug(const DevParams& p, const srf_mlp_weights& w, const float* pts, const float* viewdir, int n,
                           int n_per, float* raw_out, int32_t* dbg_sphere, int flags, void* workspace, size_t ws_bytes,
                           int debug_layer, float* debug_acc, cudaStream_t st);

}  // namespace srf
