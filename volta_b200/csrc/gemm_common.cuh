// Epilogue description shared by the tcgen05 GEMM (gemm_tc.cu) and the SIMT GEMM (gemm_simt.cu),
// so both back-ends compute bit-for-bit the same epilogue arithmetic on the fp32 accumulator.
#pragma once
#include "common.cuh"

namespace vb {

struct EpiParams {
  void* D;
  int64_t ldd;
  const float* bias;
  const void* aux_in;
  void* aux_out;
  int64_t ld_aux;
  int64_t M, N;
  float alpha;
  float drop_scale;      // 1/(1-p)
  uint32_t drop_thresh;  // p * 2^16
  uint64_t seed, offset;
  const uint64_t* seed_ptr;  // see vb_set_seed_pointer
  int32_t epilogue;      // VB_EPI_*
  int32_t d_bf16;        // 1: D/aux are bf16, 0: fp32
};

// Scalar epilogue for one accumulator element (m, n).  `aux` is aux_in[m, n] already converted
// (0 when unused).  Returns the value to store in D; `pre` receives the aux_out value.
__device__ __forceinline__ float epilogue_scalar(const EpiParams& p, float acc, float bias, float aux,
                                                 int64_t m, int64_t n, float& pre) {
  acc *= p.alpha;
  switch (p.epilogue) {
    case VB_EPI_NONE: return acc;
    case VB_EPI_BIAS: return acc + bias;
    case VB_EPI_BIAS_GELU: { const float x = acc + bias; pre = dgelu_erf(x); return gelu_erf(x); }
    case VB_EPI_BIAS_RELU: return fmaxf(acc + bias, 0.0f);
    case VB_EPI_BIAS_DROP_RES: {
      float v = acc + bias;
      if (p.drop_thresh != 0u) {
        const bool keep = dropout_keep(effective_seed(p.seed, p.seed_ptr), p.offset, (uint64_t)m * (uint64_t)p.N + (uint64_t)n, p.drop_thresh);
        v = keep ? v * p.drop_scale : 0.0f;
      }
      return v + aux;
    }
    case VB_EPI_MUL: return acc * aux;
    case VB_EPI_ADD: return acc + aux;
    case VB_EPI_DRELU: return aux > 0.0f ? acc : 0.0f;
    case VB_EPI_BIAS_ADD: return acc + bias + aux;
    default: return acc;  // VB_EPI_ACCUM handled by the caller (atomic add)
  }
}

__host__ __device__ __forceinline__ bool epi_uses_bias(int e) {
  return e == VB_EPI_BIAS || e == VB_EPI_BIAS_GELU || e == VB_EPI_BIAS_RELU || e == VB_EPI_BIAS_DROP_RES ||
         e == VB_EPI_BIAS_ADD;
}
__host__ __device__ __forceinline__ bool epi_uses_aux_in(int e) {
  return e == VB_EPI_BIAS_DROP_RES || e == VB_EPI_MUL || e == VB_EPI_ADD || e == VB_EPI_DRELU ||
         e == VB_EPI_BIAS_ADD;
}

int gemm_tc_launch(const vb_gemm_args& a, cudaStream_t stream);
int gemm_simt_launch(const vb_gemm_args& a, cudaStream_t stream);
bool gemm_tc_supported(const vb_gemm_args& a, const char** why);
int fill_epi_params(const vb_gemm_args& a, EpiParams& p);

}  // namespace vb
