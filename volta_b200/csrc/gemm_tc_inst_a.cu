// tcgen05 GEMM instantiations: VB_EPI_NONE, VB_EPI_BIAS
#include "gemm_tc_kernel.cuh"

namespace vb {
VB_DEFINE_TC_LAUNCH(VB_EPI_NONE)
VB_DEFINE_TC_LAUNCH(VB_EPI_BIAS)
}  // namespace vb
