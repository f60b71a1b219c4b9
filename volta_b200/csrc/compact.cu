// Row compaction with a device-side count - the masked-LM head only needs the rows whose label is not ignored.
//
// The reference computes vocabulary logits for ALL B*T positions (volta/encoders.py:686-696, 772) and lets
// CrossEntropyLoss(ignore_index=-1) (encoders.py:1030, 1095-1099) drop ~90 % of them: rows with label -1 add nothing
// to the loss and receive a zero gradient.  Here the labelled rows are packed to the front of a buffer, and the
// decoder GEMMs / cross-entropy / bias gradient run on `count` rows only - with the same loss and the same gradients.
// `count` lives in device memory: launches (and CUDA-graph nodes) stay static, kernels read the limit themselves
// (vb_gemm_args::m_limit / k_limit, row_limit of vb_xent_* / vb_colsum).
#include "common.cuh"

namespace vb {
namespace compact {

constexpr int PLAN_THREADS = 1024;

// One block.  Thread t owns the contiguous range [t*per, (t+1)*per): local count, block-wide exclusive scan, then the
// order-preserving scatter of indices / labels.  rows is ~1e4: a few microseconds.
__global__ void __launch_bounds__(PLAN_THREADS)
plan_kernel(const int64_t* __restrict__ labels, int64_t rows, int32_t pad, int32_t* __restrict__ limits,
            int32_t* __restrict__ src_index, int32_t* __restrict__ pos, int64_t* __restrict__ labels_c) {
  __shared__ int32_t warp_tot[PLAN_THREADS / 32];
  __shared__ int32_t total_s;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int64_t per = (rows + PLAN_THREADS - 1) / PLAN_THREADS;
  const int64_t r0 = min(rows, (int64_t)t * per), r1 = min(rows, r0 + per);
  int32_t mine = 0;
  for (int64_t r = r0; r < r1; ++r) mine += labels[r] >= 0 ? 1 : 0;
  int32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int32_t w = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t v = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += v;
    }
    warp_tot[lane] = w;                 // inclusive totals over warps
    if (lane == 31) total_s = w;
  }
  __syncthreads();
  int32_t at = incl - mine + (warp > 0 ? warp_tot[warp - 1] : 0);
  for (int64_t r = r0; r < r1; ++r) {
    const int64_t lab = labels[r];
    if (lab >= 0) {
      pos[r] = at;
      src_index[at] = (int32_t)r;
      labels_c[at] = lab;
      ++at;
    } else {
      pos[r] = -1;
    }
  }
  const int32_t total = total_s;
  for (int64_t i = total + t; i < rows; i += PLAN_THREADS) { src_index[i] = -1; labels_c[i] = -1; }
  if (t == 0) {
    limits[0] = total;
    const int64_t padded = ((int64_t)total + pad - 1) / pad * pad;
    limits[1] = (int32_t)min(rows, padded);
  }
}

// dst[i, :] = src[src_index[i], :] for i < limits[0]; zero rows up to limits[1]; rows beyond are not touched.
template <typename T>
__global__ void __launch_bounds__(256)
gather_kernel(const T* __restrict__ src, int64_t ld_src, const int32_t* __restrict__ src_index,
              const int32_t* __restrict__ limits, T* __restrict__ dst, int64_t ld_dst, int64_t rows, int cols, bool vec) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + warp;
  if (i >= rows || i >= limits[1]) return;
  const bool live = i < limits[0];
  const T* s = live ? src + (int64_t)src_index[i] * ld_src : nullptr;
  T* d = dst + i * ld_dst;
  if (vec) {
    constexpr int E = 16 / sizeof(T);
    for (int c = lane * E; c < cols; c += 32 * E) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (live) v = *reinterpret_cast<const uint4*>(s + c);
      *reinterpret_cast<uint4*>(d + c) = v;
    }
  } else {
    for (int c = lane; c < cols; c += 32) d[c] = live ? s[c] : from_f<T>(0.f);
  }
}

// dst[r, :] = pos[r] >= 0 ? src_c[pos[r], :] : 0   (the adjoint of the gather, written densely: no atomics, no memset)
template <typename T>
__global__ void __launch_bounds__(256)
scatter_kernel(const T* __restrict__ src_c, int64_t ld_src, const int32_t* __restrict__ pos, T* __restrict__ dst,
               int64_t ld_dst, int64_t rows, int cols, bool vec) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + warp;
  if (r >= rows) return;
  const int32_t p = pos[r];
  const T* s = p >= 0 ? src_c + (int64_t)p * ld_src : nullptr;
  T* d = dst + r * ld_dst;
  if (vec) {
    constexpr int E = 16 / sizeof(T);
    for (int c = lane * E; c < cols; c += 32 * E) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (p >= 0) v = *reinterpret_cast<const uint4*>(s + c);
      *reinterpret_cast<uint4*>(d + c) = v;
    }
  } else {
    for (int c = lane; c < cols; c += 32) d[c] = p >= 0 ? s[c] : from_f<T>(0.f);
  }
}

// rows [row_offset + *row_begin, row_offset + *row_end) := 0
template <typename T>
__global__ void __launch_bounds__(256)
zero_rows_kernel(T* __restrict__ buf, int64_t ld, int cols, const int32_t* __restrict__ row_begin,
                 const int32_t* __restrict__ row_end, int64_t row_offset, const int32_t* __restrict__ row_begin2,
                 const int32_t* __restrict__ row_end2, int64_t row_offset2, bool vec) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (blockIdx.y == 1) { row_begin = row_begin2; row_end = row_end2; row_offset = row_offset2; }   // second row range
  const int64_t r = (int64_t)*row_begin + (int64_t)blockIdx.x * 8 + warp;
  if (r >= *row_end) return;
  T* d = buf + (row_offset + r) * ld;
  if (vec) {
    constexpr int E = 16 / sizeof(T);
    for (int c = lane * E; c < cols; c += 32 * E) *reinterpret_cast<uint4*>(d + c) = make_uint4(0u, 0u, 0u, 0u);
  } else {
    for (int c = lane; c < cols; c += 32) d[c] = from_f<T>(0.f);
  }
}

template <typename T>
static bool vec_ok(const void* a, int64_t lda, const void* b, int64_t ldb, int cols) {
  constexpr int E = 16 / sizeof(T);
  return aligned16(a) && aligned16(b) && lda % E == 0 && ldb % E == 0 && cols % E == 0;
}

}  // namespace compact
}  // namespace vb

extern "C" {

int vb_compact_plan(const int64_t* labels, int64_t rows, int32_t pad, int32_t* limits, int32_t* src_index, int32_t* pos,
                    int64_t* labels_c, void* stream) {
  VB_REQUIRE(labels && limits && src_index && pos && labels_c, VB_ERR_ARG, "compact_plan: null pointer");
  VB_REQUIRE(rows >= 0 && rows < (int64_t)1 << 31 && pad > 0, VB_ERR_ARG, "compact_plan: rows=%lld pad=%d", (long long)rows, pad);
  vb::compact::plan_kernel<<<1, vb::compact::PLAN_THREADS, 0, (cudaStream_t)stream>>>(labels, rows, pad, limits, src_index, pos, labels_c);
  return vb::check_launch("compact_plan_kernel");
}

int vb_compact_gather(const void* src, int64_t ld_src, const int32_t* src_index, const int32_t* limits, void* dst,
                      int64_t ld_dst, int64_t rows, int32_t cols, int32_t dtype, void* stream) {
  VB_REQUIRE(src && src_index && limits && dst && cols > 0 && ld_src >= cols && ld_dst >= cols, VB_ERR_ARG, "compact_gather: bad arguments");
  if (rows == 0) return 0;
  const unsigned grid = (unsigned)vb::ceil_div(rows, 8);
  using namespace vb::compact;
  if (dtype == VB_BF16)
    gather_kernel<vb::bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const vb::bf16*)src, ld_src, src_index, limits, (vb::bf16*)dst,
                                                                  ld_dst, rows, cols, vec_ok<vb::bf16>(src, ld_src, dst, ld_dst, cols));
  else
    gather_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src, ld_src, src_index, limits, (float*)dst, ld_dst,
                                                               rows, cols, vec_ok<float>(src, ld_src, dst, ld_dst, cols));
  return vb::check_launch("compact_gather_kernel");
}

int vb_compact_scatter(const void* src_c, int64_t ld_src, const int32_t* pos, void* dst, int64_t ld_dst, int64_t rows,
                       int32_t cols, int32_t dtype, void* stream) {
  VB_REQUIRE(src_c && pos && dst && cols > 0 && ld_src >= cols && ld_dst >= cols, VB_ERR_ARG, "compact_scatter: bad arguments");
  if (rows == 0) return 0;
  const unsigned grid = (unsigned)vb::ceil_div(rows, 8);
  using namespace vb::compact;
  if (dtype == VB_BF16)
    scatter_kernel<vb::bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const vb::bf16*)src_c, ld_src, pos, (vb::bf16*)dst, ld_dst, rows,
                                                                   cols, vec_ok<vb::bf16>(src_c, ld_src, dst, ld_dst, cols));
  else
    scatter_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src_c, ld_src, pos, (float*)dst, ld_dst, rows, cols,
                                                                vec_ok<float>(src_c, ld_src, dst, ld_dst, cols));
  return vb::check_launch("compact_scatter_kernel");
}

int vb_zero_rows(void* buf, int64_t ld, int32_t cols, const int32_t* row_begin, const int32_t* row_end, int64_t row_offset,
                 const int32_t* row_begin2, const int32_t* row_end2, int64_t row_offset2, int32_t max_rows, int32_t dtype,
                 void* stream) {
  VB_REQUIRE(buf && row_begin && row_end && cols > 0 && ld >= cols && max_rows > 0, VB_ERR_ARG, "zero_rows: bad arguments");
  VB_REQUIRE((row_begin2 == nullptr) == (row_end2 == nullptr), VB_ERR_ARG, "zero_rows: second range needs both bounds");
  const dim3 grid((unsigned)vb::ceil_div((int64_t)max_rows, 8), row_begin2 ? 2u : 1u);
  using namespace vb::compact;
  if (dtype == VB_BF16)
    zero_rows_kernel<vb::bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((vb::bf16*)buf, ld, cols, row_begin, row_end, row_offset,
                                                                     row_begin2, row_end2, row_offset2,
                                                                     vec_ok<vb::bf16>(buf, ld, buf, ld, cols));
  else
    zero_rows_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float*)buf, ld, cols, row_begin, row_end, row_offset,
                                                                  row_begin2, row_end2, row_offset2,
                                                                  vec_ok<float>(buf, ld, buf, ld, cols));
  return vb::check_launch("zero_rows_kernel");
}

}  // extern "C"
