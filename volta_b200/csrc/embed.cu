// Embedding kernels: the gather-sum of word/position/token-type rows (volta/embeddings.py:55-65, 345-348,
// 442-445) and the K = num_locs box projection (embeddings.py:135,141; a 5-wide Linear is FMAs, not a GEMM).
// The LayerNorm that follows is vb_layernorm_fwd; the 2048-d region projection is vb_gemm.
#include "common.cuh"

namespace vb {
namespace emb {

constexpr int MAX_TABLES = 4;
struct Tables {
  const float* table[MAX_TABLES];
  const int64_t* idx[MAX_TABLES];
  int64_t const_row[MAX_TABLES];
  int n;
};
struct GradTables {
  float* table[MAX_TABLES];
  const int64_t* idx[MAX_TABLES];
  int64_t const_row[MAX_TABLES];
  int64_t skip[MAX_TABLES];
  int n;
};

template <typename T>
__global__ void __launch_bounds__(256)
gather_fwd_kernel(Tables t, T* __restrict__ out, int64_t ld_out, int64_t rows, int H, int accum) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + warp;
  if (r >= rows) return;
  int64_t src[MAX_TABLES];
#pragma unroll
  for (int j = 0; j < MAX_TABLES; ++j) src[j] = (j < t.n) ? (t.idx[j] ? t.idx[j][r] : t.const_row[j]) : 0;
  for (int c = lane * 8; c < H; c += 256) {
    float acc[8];
    if (accum) Vec8<T>::load(out + r * ld_out + c, acc);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < MAX_TABLES; ++j) {
      if (j < t.n) {
        float v[8];
        Vec8<float>::load(t.table[j] + src[j] * H + c, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += v[k];
      }
    }
    Vec8<T>::store(out + r * ld_out + c, acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
gather_bwd_kernel(GradTables t, const T* __restrict__ dout, int64_t ld, int64_t rows, int H) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + warp;
  if (r >= rows) return;
  int64_t dst[MAX_TABLES];
#pragma unroll
  for (int j = 0; j < MAX_TABLES; ++j) dst[j] = (j < t.n) ? (t.idx[j] ? t.idx[j][r] : t.const_row[j]) : 0;
  for (int c = lane * 8; c < H; c += 256) {
    float v[8];
    Vec8<T>::load(dout + r * ld + c, v);
#pragma unroll
    for (int j = 0; j < MAX_TABLES; ++j) {
      if (j < t.n && t.table[j] != nullptr && dst[j] != t.skip[j]) {
        float* d = t.table[j] + dst[j] * H + c;
        if ((reinterpret_cast<uintptr_t>(d) & 15u) == 0) {  // two 16-byte vector reductions instead of eight scalar atomics
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) atomicAdd(d + k, v[k]);
        }
      }
    }
  }
}

// out[key[r], c] += a[r, c]: a thread owns a column and keeps one partial per key, so no atomics inside a block.
template <typename T, int NK>
__global__ void __launch_bounds__(256)
colsum_by_key_kernel(const T* __restrict__ a, int64_t lda, int64_t rows, int64_t cols, const int64_t* __restrict__ keys,
                     int nkeys, float* __restrict__ out, int64_t rows_per_block) {
  const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  if (c >= cols) return;
  float acc[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) acc[k] = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    const int key = (int)keys[r];
    const float v = to_f<T>(a[r * lda + c]);
#pragma unroll
    for (int k = 0; k < NK; ++k) acc[k] += (key == k) ? v : 0.f;
  }
#pragma unroll
  for (int k = 0; k < NK; ++k)
    if (k < nkeys && acc[k] != 0.f) atomicAdd(out + (int64_t)k * cols + c, acc[k]);
}

template <typename T>
__global__ void __launch_bounds__(256)
vlbert_prep_kernel(const float* __restrict__ feat, const float* __restrict__ mask_emb, T* __restrict__ out, int64_t ld_out,
                   int64_t* __restrict__ masked, int64_t* __restrict__ kind, int64_t rows, int F, int K, int mvrc, int end_kind) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + warp;
  if (r >= rows) return;
  const float* f = feat + r * F;
  bool nz = false;
  for (int c = lane; c < F; c += 32) nz |= (f[c] != 0.f);
  const bool is_masked = !__any_sync(0xffffffffu, nz);
  for (int c = lane; c < F; c += 32) out[r * ld_out + c] = from_f<T>(is_masked ? mask_emb[c] : f[c]);
  if (lane == 0) {
    masked[r] = is_masked ? 1 : 0;
    kind[r] = (r % K == K - 1) ? end_kind : ((is_masked && mvrc) ? 1 : 0);
  }
}

// single block: B is a few hundred at most
__global__ void __launch_bounds__(1024)
vlbert_index_kernel(const int64_t* __restrict__ ids, int B, int T, int K, int64_t* __restrict__ tpos,
                    int64_t* __restrict__ opos, int64_t* __restrict__ text_end, int64_t* __restrict__ last_row) {
  __shared__ int s_min;
  if (threadIdx.x == 0) s_min = T;
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int n = 0;
    for (int t = 0; t < T; ++t) n += ids[(int64_t)b * T + t] != 0;
    text_end[b] = n;
    last_row[b] = (int64_t)b * K + K - 1;
    atomicMin(&s_min, n);
  }
  __syncthreads();
  const int mn = s_min;
  for (int i = threadIdx.x; i < B * T; i += blockDim.x) {
    const int t = i % T;
    tpos[i] = t >= mn ? t + K : t;
  }
  for (int i = threadIdx.x; i < B * K; i += blockDim.x) {
    const int b = i / K, k = i % K;
    opos[i] = text_end[b] + (k == K - 1 ? 1 : 0);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
batched_colsum_kernel(const T* __restrict__ a, int64_t rows, int64_t cols, float* __restrict__ out) {
  const int64_t b = blockIdx.y;
  const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  const T* p = a + b * rows * cols + c;
  float s = 0.f;
  for (int64_t r = 0; r < rows; ++r) s += to_f<T>(p[r * cols]);
  out[b * cols + c] = s;
}

template <typename T>
__global__ void __launch_bounds__(256)
box_fwd_kernel(const float* __restrict__ loc, int nl, const float* __restrict__ W, const float* __restrict__ b,
               T* __restrict__ out, int64_t ld_out, int64_t rows, int H, int accum) {
  const int64_t total = rows * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / H;
    const int c = (int)(i - r * H);
    float s = b ? b[c] : 0.f;
    for (int j = 0; j < nl; ++j) s = fmaf(loc[r * nl + j], W[c * nl + j], s);
    if (accum) s += to_f<T>(out[r * ld_out + c]);
    out[r * ld_out + c] = from_f<T>(s);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
box_bwd_kernel(const float* __restrict__ loc, int nl, const T* __restrict__ dout, int64_t ld, float* __restrict__ dW,
               float* __restrict__ db, int64_t rows, int H, int64_t rows_per_block) {
  __shared__ float part[8][9][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float acc[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) acc[j] = 0.f;
  if (c < H) {
    for (int64_t r = r0 + warp; r < r1; r += 8) {
      const float g = to_f<T>(dout[r * ld + c]);
      acc[8] += g;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nl) acc[j] = fmaf(g, loc[r * nl + j], acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 9; ++j) part[warp][j][lane] = acc[j];
  __syncthreads();
  if (warp == 0 && c < H) {
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += part[w][j][lane];
      if (j < nl) atomicAdd(dW + (int64_t)c * nl + j, s);
      else if (j == 8 && db) atomicAdd(db + c, s);
    }
  }
}

}  // namespace emb
}  // namespace vb

extern "C" {

int vb_embed_gather_fwd(const float* const* tables, const int64_t* const* idx, const int64_t* const_rows,
                        int32_t n_tables, void* out, int64_t ld_out, int64_t rows, int32_t H, int32_t accum,
                        int32_t dtype, void* stream) {
  VB_REQUIRE(tables && out && n_tables >= 1 && n_tables <= vb::emb::MAX_TABLES, VB_ERR_ARG, "embed_gather_fwd: bad table list");
  VB_REQUIRE(H % 8 == 0 && vb::aligned16(out) && (ld_out % 8) == 0, VB_ERR_ARG, "embed_gather_fwd: H/ld must be multiples of 8");
  if (rows == 0) return 0;
  vb::emb::Tables t{};
  t.n = n_tables;
  for (int j = 0; j < n_tables; ++j) {
    VB_REQUIRE(tables[j] && vb::aligned16(tables[j]), VB_ERR_ARG, "embed_gather_fwd: table %d null/unaligned", j);
    t.table[j] = tables[j];
    t.idx[j] = idx ? idx[j] : nullptr;
    t.const_row[j] = const_rows ? const_rows[j] : 0;
  }
  const unsigned grid = (unsigned)vb::ceil_div(rows, 8);
  if (dtype == VB_BF16)
    vb::emb::gather_fwd_kernel<vb::bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(t, (vb::bf16*)out, ld_out, rows, H, accum);
  else
    vb::emb::gather_fwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(t, (float*)out, ld_out, rows, H, accum);
  return vb::check_launch("embed_gather_fwd_kernel");
}

int vb_embed_gather_bwd(float* const* dtables, const int64_t* const* idx, const int64_t* const_rows,
                        const int64_t* skip_idx, int32_t n_tables, const void* dout, int64_t ld_dout, int64_t rows,
                        int32_t H, int32_t dtype, void* stream) {
  VB_REQUIRE(dtables && dout && n_tables >= 1 && n_tables <= vb::emb::MAX_TABLES, VB_ERR_ARG, "embed_gather_bwd: bad table list");
  VB_REQUIRE(H % 8 == 0 && vb::aligned16(dout) && (ld_dout % 8) == 0, VB_ERR_ARG, "embed_gather_bwd: H/ld must be multiples of 8");
  if (rows == 0) return 0;
  vb::emb::GradTables t{};
  t.n = n_tables;
  for (int j = 0; j < n_tables; ++j) {
    t.table[j] = dtables[j];  // may be NULL: that table gets no gradient (frozen)
    t.idx[j] = idx ? idx[j] : nullptr;
    t.const_row[j] = const_rows ? const_rows[j] : 0;
    t.skip[j] = skip_idx ? skip_idx[j] : -1;
  }
  const unsigned grid = (unsigned)vb::ceil_div(rows, 8);
  if (dtype == VB_BF16)
    vb::emb::gather_bwd_kernel<vb::bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(t, (const vb::bf16*)dout, ld_dout, rows, H);
  else
    vb::emb::gather_bwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(t, (const float*)dout, ld_dout, rows, H);
  return vb::check_launch("embed_gather_bwd_kernel");
}

int vb_colsum_by_key(const void* a, int32_t dtype, int64_t lda, int64_t rows, int64_t cols, const int64_t* keys,
                     int32_t nkeys, float* out, void* stream) {
  VB_REQUIRE(a && keys && out && nkeys >= 1 && nkeys <= 4, VB_ERR_ARG, "colsum_by_key: needs 1..4 keys (got %d)", nkeys);
  if (rows * cols == 0) return 0;
  const int64_t strips = vb::ceil_div(cols, 256);
  int64_t slabs = std::max<int64_t>(1, std::min<int64_t>(vb::ceil_div(rows, 32), (148 * 8) / strips));
  const int64_t rpb = vb::ceil_div(rows, slabs);
  slabs = vb::ceil_div(rows, rpb);
  dim3 grid((unsigned)strips, (unsigned)slabs);
  if (dtype == VB_BF16)
    vb::emb::colsum_by_key_kernel<vb::bf16, 4><<<grid, 256, 0, (cudaStream_t)stream>>>((const vb::bf16*)a, lda, rows, cols, keys, nkeys, out, rpb);
  else
    vb::emb::colsum_by_key_kernel<float, 4><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)a, lda, rows, cols, keys, nkeys, out, rpb);
  return vb::check_launch("colsum_by_key_kernel");
}

int vb_vlbert_prep(const float* feat, const float* mask_emb, void* out, int64_t ld_out, int64_t* masked, int64_t* kind,
                   int64_t rows, int32_t F, int32_t K, int32_t mvrc, int32_t end_kind, int32_t dtype, void* stream) {
  VB_REQUIRE(feat && mask_emb && out && masked && kind && F > 0 && K > 0, VB_ERR_ARG, "vlbert_prep: bad arguments");
  if (rows == 0) return 0;
  const unsigned grid = (unsigned)vb::ceil_div(rows, 8);
  if (dtype == VB_BF16)
    vb::emb::vlbert_prep_kernel<vb::bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(feat, mask_emb, (vb::bf16*)out, ld_out, masked, kind, rows, F, K, mvrc, end_kind);
  else
    vb::emb::vlbert_prep_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(feat, mask_emb, (float*)out, ld_out, masked, kind, rows, F, K, mvrc, end_kind);
  return vb::check_launch("vlbert_prep_kernel");
}

int vb_vlbert_index(const int64_t* ids, int32_t B, int32_t T, int32_t K, int64_t* tpos, int64_t* opos, int64_t* text_end,
                    int64_t* last_row, void* stream) {
  VB_REQUIRE(ids && tpos && opos && text_end && last_row && B > 0 && T > 0 && K > 0, VB_ERR_ARG, "vlbert_index: bad arguments");
  vb::emb::vlbert_index_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(ids, B, T, K, tpos, opos, text_end, last_row);
  return vb::check_launch("vlbert_index_kernel");
}

int vb_batched_colsum(const void* a, int32_t dtype, int64_t batch, int64_t rows, int64_t cols, float* out, void* stream) {
  VB_REQUIRE(a && out && batch > 0 && batch <= 65535, VB_ERR_ARG, "batched_colsum: bad arguments");
  if (rows * cols == 0) return 0;
  dim3 grid((unsigned)vb::ceil_div(cols, 256), (unsigned)batch);
  if (dtype == VB_BF16)
    vb::emb::batched_colsum_kernel<vb::bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const vb::bf16*)a, rows, cols, out);
  else
    vb::emb::batched_colsum_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)a, rows, cols, out);
  return vb::check_launch("batched_colsum_kernel");
}

int vb_box_project_fwd(const float* loc, int32_t nl, const float* W, const float* b, void* out, int64_t ld_out,
                       int64_t rows, int32_t H, int32_t accum, int32_t dtype, void* stream) {
  VB_REQUIRE(loc && W && out && nl >= 1 && nl <= 8, VB_ERR_ARG, "box_project_fwd: bad arguments (nl=%d)", nl);
  if (rows == 0) return 0;
  const unsigned g = (unsigned)std::min<int64_t>(vb::ceil_div(rows * H, 256), 148 * 16);
  if (dtype == VB_BF16)
    vb::emb::box_fwd_kernel<vb::bf16><<<g, 256, 0, (cudaStream_t)stream>>>(loc, nl, W, b, (vb::bf16*)out, ld_out, rows, H, accum);
  else
    vb::emb::box_fwd_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>(loc, nl, W, b, (float*)out, ld_out, rows, H, accum);
  return vb::check_launch("box_fwd_kernel");
}

int vb_box_project_bwd(const float* loc, int32_t nl, const void* dout, int64_t ld_dout, float* dW, float* db,
                       int64_t rows, int32_t H, int32_t dtype, void* stream) {
  VB_REQUIRE(loc && dout && dW && nl >= 1 && nl <= 8, VB_ERR_ARG, "box_project_bwd: bad arguments (nl=%d)", nl);
  if (rows == 0) return 0;
  const int64_t strips = vb::ceil_div(H, 32);
  int64_t slabs = std::max<int64_t>(1, std::min<int64_t>(vb::ceil_div(rows, 64), (148 * 4) / strips));
  const int64_t rpb = vb::ceil_div(rows, slabs);
  slabs = vb::ceil_div(rows, rpb);
  dim3 grid((unsigned)strips, (unsigned)slabs);
  if (dtype == VB_BF16)
    vb::emb::box_bwd_kernel<vb::bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(loc, nl, (const vb::bf16*)dout, ld_dout, dW, db, rows, H, rpb);
  else
    vb::emb::box_bwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(loc, nl, (const float*)dout, ld_dout, dW, db, rows, H, rpb);
  return vb::check_launch("box_bwd_kernel");
}

}  // extern "C"
