// tcgen05 GEMM instantiations: VB_EPI_BIAS_DROP_RES, VB_EPI_MUL
#include "gemm_tc_kernel.cuh"

namespace vb {
VB_DEFINE_TC_LAUNCH(VB_EPI_BIAS_DROP_RES)
VB_DEFINE_TC_LAUNCH(VB_EPI_MUL)
}  // namespace vb
