// tcgen05 GEMM instantiations: VB_EPI_BIAS_GELU, VB_EPI_BIAS_RELU
#include "gemm_tc_kernel.cuh"

namespace vb {
VB_DEFINE_TC_LAUNCH(VB_EPI_BIAS_GELU)
VB_DEFINE_TC_LAUNCH(VB_EPI_BIAS_RELU)
}  // namespace vb
