// tcgen05 GEMM instantiations: VB_EPI_ACCUM, VB_EPI_BIAS_ADD
#include "gemm_tc_kernel.cuh"

namespace vb {
VB_DEFINE_TC_LAUNCH(VB_EPI_ACCUM)
VB_DEFINE_TC_LAUNCH(VB_EPI_BIAS_ADD)
}  // namespace vb
