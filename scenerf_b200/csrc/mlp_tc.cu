// placeholder until the tcgen05 kernel lands (next commit)
#include "kernels.cuh"
namespace srf {
size_t tc_weights_bytes(int, int) { return 256; }
int pack_weights_tc(const srf_mlp_weights&, void*, size_t, cudaStream_t) { return 1; }
size_t tc_workspace_bytes(int, int) { return 256; }
int run_point_mlp_tc(const DevParams&, const srf_mlp_weights&, const float*, const float*, int, int, float*, int32_t*,
                     int, void*, size_t, cudaStream_t) { return -1; }
}  // namespace srf
