// Retrieval ranking on the device.
//
// The reference keeps the 5000 x 1000 caption-image score matrix in host numpy and, per caption, argsorts the row to
// find where the matching image landed (eval_retrieval.py:201-208, top-20 list :222); for text retrieval it argsorts
// every column and takes the best-placed matching caption (:248-256).  The position of an element in a descending
// sort is a COUNT - elements that beat it - so no sort is needed: one pass per row / column.
// Ties: numpy's default argsort is not stable, so the reference's result is unspecified when scores are equal; here
// equal scores keep index order (the stable descending order).
#include <float.h>

#include "common.cuh"

namespace vb {
namespace retrieval {

constexpr int THREADS = 128;
constexpr int MAX_TOPK = 64;

__device__ __forceinline__ int block_min(int v, int* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  int r = scratch[0];
#pragma unroll
  for (int k = 1; k < THREADS / 32; ++k) r = min(r, scratch[k]);
  return r;
}

__device__ __forceinline__ int block_sum(int v, int* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  int r = 0;
#pragma unroll
  for (int k = 0; k < THREADS / 32; ++k) r += scratch[k];
  return r;
}

// (score, index) precedes (s2, i2) in the stable descending order
__device__ __forceinline__ bool before(float s, int i, float s2, int i2) { return s > s2 || (s == s2 && i < i2); }

// One block per caption row: rank of the first matching image, optionally the first k entries of the descending order.
__global__ void __launch_bounds__(THREADS)
row_rank_kernel(const float* __restrict__ scores, int64_t ld, const uint8_t* __restrict__ targets, int64_t ld_t, int I,
                int32_t* __restrict__ rank, int32_t* __restrict__ topk, int k) {
  __shared__ int scratch[THREADS / 32];
  __shared__ float best_s[THREADS / 32];
  __shared__ int best_i[THREADS / 32];
  const int c = blockIdx.x, t = threadIdx.x;
  const float* s = scores + (int64_t)c * ld;
  const uint8_t* tg = targets + (int64_t)c * ld_t;
  int first = I;
  for (int i = t; i < I; i += THREADS)
    if (tg[i] == 1) { first = i; break; }
  first = block_min(first, scratch);
  int r = I;                                       // no matching image: the reference's initial rank (:166)
  if (first < I) {
    const float st = s[first];
    int beats = 0;
    for (int i = t; i < I; i += THREADS) beats += before(s[i], i, st, first) ? 1 : 0;
    r = block_sum(beats, scratch);
  }
  if (t == 0) rank[c] = r;
  if (topk == nullptr) return;
  float last_s = FLT_MAX;
  int last_i = -1;
  for (int round = 0; round < k; ++round) {
    float bs = -FLT_MAX;
    int bi = I;                                    // I = nothing left
    for (int i = t; i < I; i += THREADS) {
      const float v = s[i];
      const bool after_last = last_i < 0 || before(last_s, last_i, v, i);
      if (after_last && (bi == I || before(v, i, bs, bi))) { bs = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != I && (bi == I || before(os, oi, bs, bi))) { bs = os; bi = oi; }
    }
    __syncthreads();
    if ((t & 31) == 0) { best_s[t >> 5] = bs; best_i[t >> 5] = bi; }
    __syncthreads();
    bs = best_s[0]; bi = best_i[0];
#pragma unroll
    for (int w = 1; w < THREADS / 32; ++w)
      if (best_i[w] != I && (bi == I || before(best_s[w], best_i[w], bs, bi))) { bs = best_s[w]; bi = best_i[w]; }
    if (t == 0) topk[(int64_t)c * k + round] = bi == I ? -1 : bi;
    last_s = bs; last_i = bi;
    if (bi == I) {                                 // fewer than k entries: pad with -1
      for (int rest = round + 1 + t; rest < k; rest += THREADS) topk[(int64_t)c * k + rest] = -1;
      break;
    }
  }
}

// Image columns: a block owns 32 adjacent columns (a warp reads 128 contiguous bytes of a row) and splits the captions
// over its 32 warps; per column the best-placed matching caption is reduced through shared memory, then the captions
// that beat it are counted the same way.
constexpr int COL_ROWS = 32;
__global__ void __launch_bounds__(32 * COL_ROWS)
col_rank_kernel(const float* __restrict__ scores, int64_t ld, const uint8_t* __restrict__ targets, int64_t ld_t, int C, int I,
                int32_t* __restrict__ rank) {
  __shared__ float s_best[COL_ROWS][33];
  __shared__ int s_idx[COL_ROWS][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i = blockIdx.x * 32 + tx;
  const bool live = i < I;
  float bs = 0.f;
  int bc = C;
  if (live) {
    for (int c = ty; c < C; c += COL_ROWS) {
      if (targets[(int64_t)c * ld_t + i] == 1) {
        const float v = scores[(int64_t)c * ld + i];
        if (bc == C || before(v, c, bs, bc)) { bs = v; bc = c; }
      }
    }
  }
  s_best[ty][tx] = bs;
  s_idx[ty][tx] = bc;
  __syncthreads();
  bs = s_best[0][tx]; bc = s_idx[0][tx];
#pragma unroll 4
  for (int k = 1; k < COL_ROWS; ++k) {
    const float os = s_best[k][tx];
    const int oc = s_idx[k][tx];
    if (oc != C && (bc == C || before(os, oc, bs, bc))) { bs = os; bc = oc; }
  }
  __syncthreads();
  int beats = 0;
  if (live && bc < C)
    for (int c = ty; c < C; c += COL_ROWS) beats += before(scores[(int64_t)c * ld + i], c, bs, bc) ? 1 : 0;
  s_idx[ty][tx] = beats;
  __syncthreads();
  if (ty == 0 && live) {
    int r = 0;
#pragma unroll 4
    for (int k = 0; k < COL_ROWS; ++k) r += s_idx[k][tx];
    rank[i] = bc < C ? r : C;
  }
}

}  // namespace retrieval
}  // namespace vb

extern "C" int vb_retrieval_ranks(const float* scores, int64_t ld, const uint8_t* targets, int64_t ld_t, int32_t C, int32_t I,
                                  int32_t axis, int32_t* rank, int32_t* topk, int32_t k, void* stream) {
  using namespace vb;
  VB_REQUIRE(scores && targets && rank, VB_ERR_ARG, "vb_retrieval_ranks: null pointer");
  VB_REQUIRE(C >= 0 && I >= 0 && ld >= I && ld_t >= I, VB_ERR_ARG, "vb_retrieval_ranks: C=%d I=%d ld=%lld ld_t=%lld", C, I,
             (long long)ld, (long long)ld_t);
  VB_REQUIRE(axis == 0 || axis == 1, VB_ERR_ARG, "vb_retrieval_ranks: axis=%d (0 = per caption row, 1 = per image column)", axis);
  VB_REQUIRE(topk == nullptr || (axis == 0 && k >= 1 && k <= retrieval::MAX_TOPK), VB_ERR_ARG,
             "vb_retrieval_ranks: top-k list needs axis 0 and 1 <= k <= %d (k=%d)", retrieval::MAX_TOPK, k);
  if (C == 0 || I == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (axis == 0)
    retrieval::row_rank_kernel<<<C, retrieval::THREADS, 0, st>>>(scores, ld, targets, ld_t, I, rank, topk, k);
  else
    retrieval::col_rank_kernel<<<(unsigned)ceil_div(I, 32), dim3(32, retrieval::COL_ROWS), 0, st>>>(scores, ld, targets, ld_t, C, I,
                                                                                                   rank);
  return check_launch("vb_retrieval_ranks");
}
