// SIMT GEMM: the fp32 validation mode and the shapes TMA cannot take (K = 5 box projection, N = 1/2
// heads, unaligned pitches).  Same epilogues as the tcgen05 kernel (gemm_common.cuh).  fp32 FMA
// accumulation in k order, 64x64 tile, 4x4 outputs per thread.
#include "gemm_common.cuh"

namespace vb {
namespace simt {

constexpr int TM = 64, TN = 64, TK = 16;

struct Operand {
  const void* ptr;
  int64_t stride_mn, stride_k;  // element strides
};

template <typename T>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(Operand A, Operand B, int64_t M, int64_t N, int64_t K, EpiParams p, const int32_t* m_limit,
                 const int32_t* k_limit, const int32_t* m_limit2, int64_t seg_M) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const T* a = reinterpret_cast<const T*>(A.ptr);
  const T* b = reinterpret_cast<const T*>(B.ptr);
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * TM, n0 = (int64_t)blockIdx.x * TN;
  // device-side extents (vb_gemm_args::m_limit / k_limit / m_limit2)
  // (whole 256-row groups, like the tcgen05 back-end: rows up to the padded count are computed from the zero rows there)
  if (seg_M > 0 && m0 >= seg_M) {
    if (m_limit2 && m0 - seg_M >= (((int64_t)*m_limit2 + 255) & ~(int64_t)255)) return;
  } else if (m_limit && m0 >= (((int64_t)*m_limit + 255) & ~(int64_t)255)) {
    return;
  }
  if (k_limit) K = min(K, (int64_t)max(*k_limit, 0));
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // choose the loading order that walks the contiguous dimension with consecutive threads
  const bool a_k_fast = A.stride_k == 1, b_k_fast = B.stride_k == 1;
  for (int64_t k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
    for (int i = 0; i < (TM * TK) / 256; ++i) {
      const int e = tid + i * 256;
      const int mm = a_k_fast ? e / TK : e % TM, kk = a_k_fast ? e % TK : e / TM;
      const int64_t gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < K) ? to_f<T>(a[gm * A.stride_mn + gk * A.stride_k]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < (TN * TK) / 256; ++i) {
      const int e = tid + i * 256;
      const int nn = b_k_fast ? e / TK : e % TN, kk = b_k_fast ? e % TK : e / TN;
      const int64_t gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < N && gk < K) ? to_f<T>(b[gn * B.stride_mn + gk * B.stride_k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int e = p.epilogue;
  const bool use_bias = epi_uses_bias(e), use_aux = epi_uses_aux_in(e);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= N) continue;
      if (e == VB_EPI_ACCUM) {
        atomicAdd(reinterpret_cast<float*>(p.D) + m * p.ldd + n, acc[i][j] * p.alpha);
        continue;
      }
      const float bias = use_bias ? p.bias[n] : 0.f;
      float aux = 0.f, pre = 0.f;
      if (use_aux)
        aux = p.d_bf16 ? to_f<bf16>(reinterpret_cast<const bf16*>(p.aux_in)[m * p.ld_aux + n])
                       : reinterpret_cast<const float*>(p.aux_in)[m * p.ld_aux + n];
      const float o = epilogue_scalar(p, acc[i][j], bias, aux, m, n, pre);
      if (p.d_bf16) {
        reinterpret_cast<bf16*>(p.D)[m * p.ldd + n] = from_f<bf16>(o);
        if (e == VB_EPI_BIAS_GELU) reinterpret_cast<bf16*>(p.aux_out)[m * p.ld_aux + n] = from_f<bf16>(pre);
      } else {
        reinterpret_cast<float*>(p.D)[m * p.ldd + n] = o;
        if (e == VB_EPI_BIAS_GELU) reinterpret_cast<float*>(p.aux_out)[m * p.ld_aux + n] = pre;
      }
    }
  }
}

}  // namespace simt

int fill_epi_params(const vb_gemm_args& a, EpiParams& p) {
  VB_REQUIRE(a.A && a.B && a.D, VB_ERR_ARG, "gemm: null operand");
  VB_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, VB_ERR_ARG, "gemm: empty problem M=%lld N=%lld K=%lld", (long long)a.M,
             (long long)a.N, (long long)a.K);
  VB_REQUIRE(a.epilogue >= VB_EPI_NONE && a.epilogue <= VB_EPI_BIAS_ADD, VB_ERR_ARG, "gemm: bad epilogue %d", a.epilogue);
  VB_REQUIRE(!epi_uses_bias(a.epilogue) || a.bias, VB_ERR_ARG, "gemm: epilogue %d needs bias", a.epilogue);
  VB_REQUIRE(!epi_uses_aux_in(a.epilogue) || a.aux_in, VB_ERR_ARG, "gemm: epilogue %d needs aux_in", a.epilogue);
  VB_REQUIRE(a.epilogue != VB_EPI_BIAS_GELU || a.aux_out, VB_ERR_ARG, "gemm: BIAS_GELU needs aux_out");
  VB_REQUIRE(a.epilogue != VB_EPI_ACCUM || a.d_dtype == VB_F32, VB_ERR_ARG, "gemm: ACCUM needs fp32 D");
  VB_REQUIRE(a.dropout_p >= 0.f && a.dropout_p < 1.f, VB_ERR_ARG, "gemm: dropout_p %f out of range", a.dropout_p);
  VB_REQUIRE(a.ldd >= a.N, VB_ERR_ARG, "gemm: ldd < N");
  p.D = a.D; p.ldd = a.ldd; p.bias = a.bias; p.aux_in = a.aux_in; p.aux_out = a.aux_out; p.ld_aux = a.ld_aux;
  p.M = a.M; p.N = a.N;
  p.alpha = a.alpha == 0.f ? 1.f : a.alpha;
  p.drop_thresh = (a.epilogue == VB_EPI_BIAS_DROP_RES && a.dropout_p > 0.f) ? dropout_thresh(a.dropout_p) : 0u;
  p.drop_scale = 1.f / (1.f - a.dropout_p);
  p.seed = a.seed; p.offset = a.offset; p.seed_ptr = seed_pointer();
  p.epilogue = a.epilogue;
  p.d_bf16 = a.d_dtype == VB_BF16;
  return 0;
}

namespace simt {
// N <= 8 output columns (the ITM / task logit heads: [B, 1024] x [2, 1024]^T): one warp per output row, the lanes
// stride over K, all N dot products in registers.  The tiled kernel above would run such a problem on 4 blocks.
template <typename T>
__global__ void __launch_bounds__(256)
gemm_skinny_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ B, int64_t ldb, int64_t M, int N, int64_t K,
                   EpiParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * 8 + warp;
  if (m >= M) return;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t k = lane; k < K; k += 32) {
    const float av = to_f<T>(A[m * lda + k]);
#pragma unroll
    for (int n = 0; n < 8; ++n)
      if (n < N) acc[n] = fmaf(av, to_f<T>(B[n * ldb + k]), acc[n]);
  }
#pragma unroll
  for (int n = 0; n < 8; ++n) acc[n] = warp_sum(acc[n]);
  if (lane == 0) {
    for (int n = 0; n < N; ++n) {
      float pre = 0.f;
      const float o = epilogue_scalar(p, acc[n], p.epilogue == VB_EPI_BIAS ? p.bias[n] : 0.f, 0.f, m, n, pre);
      if (p.d_bf16) reinterpret_cast<bf16*>(p.D)[m * p.ldd + n] = from_f<bf16>(o);
      else reinterpret_cast<float*>(p.D)[m * p.ldd + n] = o;
    }
  }
}
}  // namespace simt

int gemm_simt_launch(const vb_gemm_args& a, cudaStream_t stream) {
  VB_REQUIRE(a.B2 == nullptr || a.split_M <= 0, VB_ERR_UNSUPPORTED, "gemm: row-split launches need the tcgen05 back-end");
  EpiParams p;
  int rc = fill_epi_params(a, p);
  if (rc) return rc;
  if (a.N <= 8 && a.a_kmajor && a.b_kmajor && (a.epilogue == VB_EPI_NONE || a.epilogue == VB_EPI_BIAS) && !a.m_limit &&
      !a.k_limit && !a.m_limit2) {
    const unsigned grid = (unsigned)ceil_div(a.M, 8);
    if (a.ab_dtype == VB_BF16)
      simt::gemm_skinny_kernel<bf16><<<grid, 256, 0, stream>>>((const bf16*)a.A, a.lda, (const bf16*)a.B, a.ldb, a.M, (int)a.N, a.K, p);
    else
      simt::gemm_skinny_kernel<float><<<grid, 256, 0, stream>>>((const float*)a.A, a.lda, (const float*)a.B, a.ldb, a.M, (int)a.N, a.K, p);
    return check_launch("gemm_skinny_kernel");
  }
  simt::Operand A{a.A, a.a_kmajor ? a.lda : 1, a.a_kmajor ? 1 : a.lda};
  simt::Operand B{a.B, a.b_kmajor ? a.ldb : 1, a.b_kmajor ? 1 : a.ldb};
  dim3 grid((unsigned)ceil_div(a.N, simt::TN), (unsigned)ceil_div(a.M, simt::TM));
  VB_REQUIRE(grid.y <= 65535u, VB_ERR_UNSUPPORTED, "gemm_simt: M too large (%lld)", (long long)a.M);
  if (a.ab_dtype == VB_BF16)
    simt::gemm_simt_kernel<bf16><<<grid, 256, 0, stream>>>(A, B, a.M, a.N, a.K, p, a.m_limit, a.k_limit, a.m_limit2, (a.limit_seg_M > 0 && a.limit_seg_M < a.M) ? a.limit_seg_M : 0);
  else
    simt::gemm_simt_kernel<float><<<grid, 256, 0, stream>>>(A, B, a.M, a.N, a.K, p, a.m_limit, a.k_limit, a.m_limit2, (a.limit_seg_M > 0 && a.limit_seg_M < a.M) ? a.limit_seg_M : 0);
  return check_launch("gemm_simt_kernel");
}

}  // namespace vb
