// tcgen05 GEMM instantiations: VB_EPI_ADD, VB_EPI_DRELU
#include "gemm_tc_kernel.cuh"

namespace vb {
VB_DEFINE_TC_LAUNCH(VB_EPI_ADD)
VB_DEFINE_TC_LAUNCH(VB_EPI_DRELU)
}  // namespace vb
