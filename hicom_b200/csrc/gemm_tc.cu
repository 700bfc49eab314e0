// placeholder until the tcgen05 kernels land: everything routes to the SIMT path
#include "gemm_tc.cuh"
namespace hicom {
bool tc_linear_supported(int, int, int, int, int, long long, long long, long long, const void*, const void*, const void*) { return false; }
int launch_tc_linear(const TcLinearParams&, cudaStream_t) { set_error("tcgen05 linear not built"); return 1; }
bool tc_global_selected(int, int, int, int) { return false; }
size_t tc_global_workspace_bytes(int, int, int, int, int, int, int) { return 0; }
int launch_tc_global(const void*, const float*, const float*, const float*, const void*, float*, float*, float*, int, int, int, int, int, int, int, void*, cudaStream_t) { set_error("tcgen05 global not built"); return 1; }
}  // namespace hicom
