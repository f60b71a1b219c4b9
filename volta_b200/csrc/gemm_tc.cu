// Host-side dispatch of the tcgen05 GEMM (kernel: gemm_tc_kernel.cuh; instantiations: gemm_tc_inst*.cu).
#include "gemm_common.cuh"

namespace vb {

#define VB_DECL(E) int gemm_tc_launch_##E(const vb_gemm_args& a, cudaStream_t stream);
VB_DECL(VB_EPI_NONE) VB_DECL(VB_EPI_BIAS) VB_DECL(VB_EPI_BIAS_GELU) VB_DECL(VB_EPI_BIAS_RELU) VB_DECL(VB_EPI_BIAS_DROP_RES)
VB_DECL(VB_EPI_MUL) VB_DECL(VB_EPI_ADD) VB_DECL(VB_EPI_DRELU) VB_DECL(VB_EPI_ACCUM) VB_DECL(VB_EPI_BIAS_ADD)
#undef VB_DECL

bool gemm_tc_supported(const vb_gemm_args& a, const char** why) {
  auto fail = [&](const char* w) { if (why) *why = w; return false; };
  if (a.ab_dtype != VB_BF16) return fail("operands are not bf16");
  if (!aligned16(a.A) || !aligned16(a.B)) return fail("operand base not 16-byte aligned");
  if ((a.lda % 8) != 0 || (a.ldb % 8) != 0) return fail("operand pitch not a multiple of 8 elements");
  if (a.M < 1 || a.N < 1 || a.K < 1) return fail("empty problem");
  const int esz = a.d_dtype == VB_BF16 ? 2 : 4;
  if (!aligned16(a.D) || (a.ldd * esz) % 16 != 0) return fail("D not 16-byte aligned");
  if (epi_uses_aux_in(a.epilogue) && (!aligned16(a.aux_in) || (a.ld_aux * esz) % 16 != 0)) return fail("aux_in alignment");
  if (a.epilogue == VB_EPI_BIAS_GELU && (!aligned16(a.aux_out) || (a.ld_aux * esz) % 16 != 0)) return fail("aux_out alignment");
  if (epi_uses_bias(a.epilogue) && !aligned16(a.bias)) return fail("bias alignment");
  if (a.epilogue == VB_EPI_ACCUM && a.d_dtype != VB_F32) return fail("ACCUM needs fp32 D");
  if (a.B2 != nullptr && a.split_M > 0) {
    if (a.epilogue == VB_EPI_ACCUM) return fail("row-split launches are for forward / dgrad GEMMs");
    if ((a.split_M % 256) != 0 || a.split_M >= a.M || !aligned16(a.B2)) return fail("split_M must be a multiple of 256 below M, B2 16-byte aligned");
  }
  if (a.D2 != nullptr && a.k_seg > 0) {
    if (a.epilogue != VB_EPI_ACCUM) return fail("two reduction segments (k_seg / D2) are for ACCUM (weight-gradient) GEMMs");
    if ((a.k_seg % 64) != 0 || a.k_seg >= a.K || !aligned16(a.D2)) return fail("k_seg must be a multiple of 64 below K, D2 16-byte aligned");
  }
  return true;
}

int gemm_tc_launch(const vb_gemm_args& a, cudaStream_t stream) {
  const char* why = nullptr;
  VB_REQUIRE(gemm_tc_supported(a, &why), VB_ERR_UNSUPPORTED, "tcgen05 GEMM cannot run this problem: %s", why);
  switch (a.epilogue) {
#define VB_CASE(E) case E: return gemm_tc_launch_##E(a, stream);
    VB_CASE(VB_EPI_NONE) VB_CASE(VB_EPI_BIAS) VB_CASE(VB_EPI_BIAS_GELU) VB_CASE(VB_EPI_BIAS_RELU) VB_CASE(VB_EPI_BIAS_DROP_RES)
    VB_CASE(VB_EPI_MUL) VB_CASE(VB_EPI_ADD) VB_CASE(VB_EPI_DRELU) VB_CASE(VB_EPI_ACCUM) VB_CASE(VB_EPI_BIAS_ADD)
#undef VB_CASE
  }
  set_error("gemm: bad epilogue %d", a.epilogue);
  return VB_ERR_ARG;
}

}  // namespace vb
