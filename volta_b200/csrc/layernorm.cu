// LayerNorm forward / backward - replaces apex FusedLayerNorm (apex/csrc/layer_norm_cuda_kernel.cu:
// cuApplyLayerNorm :279-322, cuComputeGradInput :522-637, cuCompute(Part)GradGammaBeta :403-520) and the
// Python BertLayerNorm (volta/encoders.py:48-61).  HBM-bound: one warp per row, 16-byte vector loads, the
// row stays in registers between the statistics and the normalisation (one read, one write per element),
// fp32 statistics, warp-shuffle reductions.  The backward fuses the residual blocks' extras: the dropout-
// masked copy of dx that feeds the preceding dense layer's dgrad/wgrad, and the dense bias gradient.
#include "common.cuh"

namespace vb {
namespace ln {

constexpr int WARPS = 8;

// vb_row_seqs by value: which rows are trailing padding of their sample.  Such rows are zeros on output and contribute
// nothing: as attention keys they are masked (weight exactly 0), as queries the attention kernel skips them
// (vb_attn_args::len_*), their gradients are exactly zero, and no loss reads them - so every VALID row is bit-identical
// with and without the skip, and the kernel saves the read of a third of the rows.
struct Seqs {
  const int32_t* len_a; int32_t L_a; int64_t rows_a;
  const int32_t* len_b; int32_t L_b;
};
__device__ __forceinline__ bool row_dead(const Seqs& q, int64_t row) {
  if (q.len_a == nullptr) return false;
  const int32_t* len = q.len_a;
  int L = q.L_a;
  int64_t r = row;
  if (row >= q.rows_a) {
    if (q.len_b == nullptr) return false;
    len = q.len_b; L = q.L_b; r = row - q.rows_a;
  }
  const int64_t b = r / L;
  const int v = len[b];
  return v > 0 && (int)(r - b * L) >= v;
}

template <typename T, int NV>
__global__ void __launch_bounds__(WARPS * 32)
ln_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              T* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows, int H,
              float eps, float out_scale, int accum, Seqs seqs, const float* __restrict__ gamma2,
              const float* __restrict__ beta2, int64_t split_row) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS + warp;
  if (row >= rows) return;
  if (row >= split_row) { gamma = gamma2; beta = beta2; }   // second stream of a joint buffer: its own affine parameters
  if (!accum && row_dead(seqs, row)) {   // trailing padding of its sample: zeros, nothing read
    if (lane == 0) {
      if (mean_out) mean_out[row] = 0.f;
      if (rstd_out) rstd_out[row] = 0.f;
    }
    const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + 32 * i) * 8;
      if (c < H) Vec8<T>::store(y + row * H + c, z);
    }
    return;
  }
  const T* xr = x + row * H;
  float v[NV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 8;
    if (c < H) {
      Vec8<T>::load(xr + c, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    }
  }
  const float mean = warp_sum(s) / (float)H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 8;
    if (c < H) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; q += d * d; }
    }
  }
  const float var = warp_sum(q) / (float)H;       // biased variance
  const float rstd = rsqrtf(var + eps);           // eps inside the sqrt
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  T* yr = y + row * H;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 8;
    if (c < H) {
      float g[8], b[8], o[8];
      Vec8<float>::load(gamma + c, g);
      Vec8<float>::load(beta + c, b);
      if (accum) Vec8<T>::load(yr + c, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = (g[j] * ((v[i][j] - mean) * rstd) + b[j]) * out_scale;
        o[j] = accum ? o[j] + t : t;
      }
      Vec8<T>::store(yr + c, o);
    }
  }
}

// Backward.  Each warp walks rows (row = warp_global + k * total_warps).  The per-column partial sums for
// dgamma / dbeta (/ dbias) live in a per-warp slice of shared memory (conflict-free 16-byte read-modify-write,
// no block barrier inside the row loop) instead of ~70 registers per thread: the kernel fits two 8-warp blocks per
// SM, which is what a bandwidth-bound kernel needs.  The block reduces the slices and issues one atomicAdd per column.
template <typename T, int NV, bool EXTRA>
__global__ void __launch_bounds__(WARPS * 32, 2)
ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
              const float* __restrict__ rstd, const float* __restrict__ gamma, T* __restrict__ dx,
              float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows, int H, float out_scale,
              T* __restrict__ dx_drop, uint32_t drop_thresh, float drop_scale, uint64_t seed_in, uint64_t offset,
              float* __restrict__ dbias, const uint64_t* __restrict__ seed_ptr, int early_fetch, Seqs seqs,
              const float* __restrict__ gamma2, float* __restrict__ dgamma2, float* __restrict__ dbeta2,
              float* __restrict__ dbias2, int64_t split_row, int blocks_a) {
  pdl_enter();
  const uint64_t seed = effective_seed(seed_in, seed_ptr);
  extern __shared__ __align__(16) float red[];  // [3 or 2][WARPS][H] partial sums, then gamma [H]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Two streams in one launch (rows [0, split_row) and [split_row, rows) with their own gamma and gradient sums): the
  // first blocks_a blocks walk the first range, the others the second, so a block's column sums belong to one stream.
  int64_t row_begin = 0, row_end = rows;
  int blk = blockIdx.x, nblk = gridDim.x;
  if (split_row < rows) {
    if (blk < blocks_a) { row_end = split_row; nblk = blocks_a; }
    else {
      row_begin = split_row; blk -= blocks_a; nblk -= blocks_a;
      gamma = gamma2; dgamma = dgamma2; dbeta = dbeta2; dbias = dbias2;
    }
  }
  const int64_t warp_global = row_begin + (int64_t)blk * WARPS + warp;
  const int64_t total_warps = (int64_t)nblk * WARPS;
  rows = row_end;
  float* sg = red + warp * H;
  float* sb = red + (WARPS + warp) * H;
  float* sx = red + (2 * WARPS + warp) * H;
  float* sgam = red + (EXTRA ? 3 : 2) * WARPS * H;   // gamma, shared by the block (frees 8*NV registers per thread)
  for (int c = threadIdx.x * 4; c < H; c += WARPS * 32 * 4)
    *reinterpret_cast<float4*>(sgam + c) = *reinterpret_cast<const float4*>(gamma + c);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 8;
    if (c < H) {
      const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      Vec8<float>::store(sg + c, z);
      Vec8<float>::store(sb + c, z);
      if constexpr (EXTRA) Vec8<float>::store(sx + c, z);
    }
  }
  __syncthreads();
  // Software pipeline over the rows of this warp: the (still packed) dy / x vectors of the NEXT row are requested
  // before the current row is processed, so the HBM latency of a row hides under the ~1000 instructions of the
  // previous one instead of stalling all warps of the SM in the same phase.
  using Raw = typename Vec8<T>::Raw;
  Raw nd[NV], nx[NV];
  float nmu = 0.f, nrs = 0.f;
  bool ndead = false;
  auto fetch = [&](int64_t r) {
    ndead = row_dead(seqs, r);
    if (ndead) return;
    nmu = mean[r]; nrs = rstd[r];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + 32 * i) * 8;
      if (c < H) {
        nd[i] = Vec8<T>::load_raw(dy + r * H + c);
        nx[i] = Vec8<T>::load_raw(x + r * H + c);
      }
    }
  };
  int64_t row = warp_global;
  if (row < rows) fetch(row);
  for (; row < rows; row += total_warps) {
    if (ndead) {   // trailing padding: dy is exactly zero there, so are dx and every partial sum
      if (row + total_warps < rows) fetch(row + total_warps);
      const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (lane + 32 * i) * 8;
        if (c < H) {
          Vec8<T>::store(dx + row * H + c, z);
          if constexpr (EXTRA) {
            if (dx_drop != nullptr) Vec8<T>::store(dx_drop + row * H + c, z);
          }
        }
      }
      continue;
    }
    const float mu = nmu, rs = nrs;
    float xh[NV][8], g[NV][8];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + 32 * i) * 8;
      if (c < H) {
        float d[8], xv[8], a8[8], b8[8], gm[8];
        Vec8<T>::unpack(nd[i], d);
        Vec8<T>::unpack(nx[i], xv);
        Vec8<float>::load(sgam + c, gm);
        Vec8<float>::load(sg + c, a8);
        Vec8<float>::load(sb + c, b8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dyo = d[j] * out_scale;
          xh[i][j] = (xv[j] - mu) * rs;
          g[i][j] = dyo * gm[j];
          c1 += g[i][j];
          c2 += g[i][j] * xh[i][j];
          a8[j] = fmaf(dyo, xh[i][j], a8[j]);
          b8[j] += dyo;
        }
        Vec8<float>::store(sg + c, a8);
        Vec8<float>::store(sb + c, b8);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { xh[i][j] = 0.f; g[i][j] = 0.f; }
      }
    }
    if (early_fetch && row + total_warps < rows) fetch(row + total_warps);   // nd / nx are dead here: reuse them for the next row
    c1 = warp_sum(c1) / (float)H;
    c2 = warp_sum(c2) / (float)H;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + 32 * i) * 8;
      if (c < H) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rs * (g[i][j] - c1 - xh[i][j] * c2);
        Vec8<T>::store(dx + row * H + c, o);
        if constexpr (EXTRA) {
          float e8[8];
          Vec8<float>::load(sx + c, e8);
          if (dx_drop != nullptr) {
            float od[8];
#pragma unroll
            for (int h = 0; h < 8; h += 2) {
              bool k0, k1;
              dropout_keep2(seed, offset, (uint64_t)row * (uint64_t)H + (uint64_t)(c + h), drop_thresh, k0, k1);
              od[h] = k0 ? o[h] * drop_scale : 0.f;
              od[h + 1] = k1 ? o[h + 1] * drop_scale : 0.f;
            }
            Vec8<T>::store(dx_drop + row * H + c, od);
            // the bias gradient sees what the dense layer's output gradient is: the rounded, masked copy
#pragma unroll
            for (int j = 0; j < 8; ++j) e8[j] += to_f<T>(from_f<T>(od[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) e8[j] += to_f<T>(from_f<T>(o[j]));
          }
          Vec8<float>::store(sx + c, e8);
        }
      }
    }
    if (!early_fetch && row + total_warps < rows) fetch(row + total_warps);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += WARPS * 32) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      s0 += red[w * H + c];
      s1 += red[(WARPS + w) * H + c];
      if constexpr (EXTRA) s2 += red[(2 * WARPS + w) * H + c];
    }
    if (dgamma) atomicAdd(dgamma + c, s0);
    if (dbeta) atomicAdd(dbeta + c, s1);
    if constexpr (EXTRA) {
      if (dbias) atomicAdd(dbias + c, s2);
    }
  }
}

template <typename T>
static int fwd_dispatch(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                        int64_t rows, int H, float eps, float out_scale, int accum, Seqs seqs, const float* gamma2,
                        const float* beta2, int64_t split_row, cudaStream_t st) {
  const int nv = (H + 255) / 256;
  const unsigned grid = (unsigned)ceil_div(rows, WARPS);
#define VB_LN_FWD(NV_)                                                                                             \
  launch_pdl(ln_fwd_kernel<T, NV_>, dim3(grid), dim3(WARPS * 32), 0, st, (const T*)x, gamma, beta, (T*)y, mean, rstd, rows, H, \
             eps, out_scale, accum, seqs, gamma2, beta2, split_row)
  switch (nv) {
    case 1: VB_LN_FWD(1); break;
    case 2: VB_LN_FWD(2); break;
    case 3: VB_LN_FWD(3); break;
    case 4: VB_LN_FWD(4); break;
    case 5: case 6: VB_LN_FWD(6); break;
    case 7: case 8: VB_LN_FWD(8); break;
    default: set_error("layernorm: H=%d > 2048 unsupported", H); return VB_ERR_UNSUPPORTED;
  }
#undef VB_LN_FWD
  return check_launch("ln_fwd_kernel");
}

template <typename T>
static int bwd_dispatch(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                        void* dx, float* dgamma, float* dbeta, int64_t rows, int H, float out_scale, void* dx_drop,
                        float p, uint64_t seed, uint64_t offset, float* dbias, Seqs seqs, const float* gamma2,
                        float* dgamma2, float* dbeta2, float* dbias2, int64_t split_row, cudaStream_t st) {
  const int nv = (H + 255) / 256;
  const bool extra = (dx_drop != nullptr) || (dbias != nullptr);
  static int sms = 0;
  if (!sms) { int dev; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  // enough rows per warp to amortise the column reduction, enough blocks to fill the machine
  int64_t grid = std::min<int64_t>(ceil_div(rows, WARPS), (int64_t)sms * 2);  // two resident blocks per SM
  int blocks_a = (int)grid;
  if (split_row < rows) {   // blocks in proportion to the two row ranges, at least one each
    grid = std::max<int64_t>(grid, 2);
    blocks_a = (int)std::min<int64_t>(grid - 1, std::max<int64_t>(1, (grid * split_row + rows / 2) / rows));
  }
  const size_t smem = ((size_t)(extra ? 3 : 2) * WARPS + 1) * H * sizeof(float);
  const uint32_t thresh = dx_drop ? dropout_thresh(p) : 0u;
  const float scale = 1.f / (1.f - p);
  const char* ef = getenv("VOLTA_B200_LN_EARLY_FETCH");
  const int early = (ef && ef[0] == '0') ? 0 : 1;
#define VB_LN_BWD(NV_, EX_)                                                                                      \
  do {                                                                                                           \
    if (smem > 48 * 1024)                                                                                        \
      cudaFuncSetAttribute(ln_bwd_kernel<T, NV_, EX_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    launch_pdl(ln_bwd_kernel<T, NV_, EX_>, dim3((unsigned)grid), dim3(WARPS * 32), smem, st,                    \
        (const T*)dy, (const T*)x, mean, rstd, gamma, (T*)dx, dgamma, dbeta, rows, H, out_scale, (T*)dx_drop,    \
        thresh, scale, seed, offset, dbias, seed_pointer(), early, seqs, gamma2, dgamma2, dbeta2, dbias2,  \
        split_row, blocks_a);                                                                     \
  } while (0)
#define VB_LN_BWD2(NV_) do { if (extra) VB_LN_BWD(NV_, true); else VB_LN_BWD(NV_, false); } while (0)
  switch (nv) {
    case 1: VB_LN_BWD2(1); break;
    case 2: VB_LN_BWD2(2); break;
    case 3: VB_LN_BWD2(3); break;
    case 4: VB_LN_BWD2(4); break;
    case 5: case 6: VB_LN_BWD2(6); break;
    case 7: case 8: VB_LN_BWD2(8); break;
    default: set_error("layernorm: H=%d > 2048 unsupported", H); return VB_ERR_UNSUPPORTED;
  }
#undef VB_LN_BWD2
#undef VB_LN_BWD
  return check_launch("ln_bwd_kernel");
}

static Seqs make_seqs(const vb_row_seqs* q) {
  Seqs s = {nullptr, 1, 0, nullptr, 1};
  if (q != nullptr && q->len_a != nullptr && q->L_a > 0) {
    s.len_a = q->len_a; s.L_a = q->L_a; s.rows_a = q->rows_a;
    if (q->len_b != nullptr && q->L_b > 0) { s.len_b = q->len_b; s.L_b = q->L_b; }
  }
  return s;
}

}  // namespace ln
}  // namespace vb

extern "C" {

int vb_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                     int64_t rows, int32_t H, float eps, float out_scale, int32_t accum, const vb_row_seqs* seqs,
                     const float* gamma2, const float* beta2, int64_t split_row, int32_t dtype, void* stream) {
  VB_REQUIRE(x && gamma && beta && y, VB_ERR_ARG, "layernorm_fwd: null pointer");
  VB_REQUIRE(rows >= 0 && H > 0 && H % 8 == 0, VB_ERR_ARG, "layernorm_fwd: H=%d must be a positive multiple of 8", H);
  VB_REQUIRE(vb::aligned16(x) && vb::aligned16(y) && vb::aligned16(gamma) && vb::aligned16(beta), VB_ERR_ARG,
             "layernorm_fwd: pointers must be 16-byte aligned");
  if (rows == 0) return 0;
  if (!gamma2 || !beta2 || split_row <= 0 || split_row >= rows) { split_row = rows; gamma2 = gamma; beta2 = beta; }
  if (dtype == VB_BF16)
    return vb::ln::fwd_dispatch<vb::bf16>(x, gamma, beta, y, mean, rstd, rows, H, eps, out_scale, accum, vb::ln::make_seqs(seqs), gamma2, beta2, split_row, (cudaStream_t)stream);
  VB_REQUIRE(dtype == VB_F32, VB_ERR_ARG, "layernorm_fwd: bad dtype %d", dtype);
  return vb::ln::fwd_dispatch<float>(x, gamma, beta, y, mean, rstd, rows, H, eps, out_scale, accum, vb::ln::make_seqs(seqs), gamma2, beta2, split_row, (cudaStream_t)stream);
}

int vb_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
                     float* dgamma, float* dbeta, int64_t rows, int32_t H, float out_scale, void* dx_drop,
                     float dropout_p, uint64_t seed, uint64_t offset, float* dbias, const vb_row_seqs* seqs,
                     const float* gamma2, float* dgamma2, float* dbeta2, float* dbias2, int64_t split_row, int32_t dtype,
                     void* stream) {
  VB_REQUIRE(dy && x && mean && rstd && gamma && dx, VB_ERR_ARG, "layernorm_bwd: null pointer");
  VB_REQUIRE(rows >= 0 && H > 0 && H % 8 == 0, VB_ERR_ARG, "layernorm_bwd: H=%d must be a positive multiple of 8", H);
  VB_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, VB_ERR_ARG, "layernorm_bwd: dropout_p out of range");
  VB_REQUIRE(vb::aligned16(dy) && vb::aligned16(x) && vb::aligned16(dx) && vb::aligned16(gamma) &&
                 (!dx_drop || vb::aligned16(dx_drop)), VB_ERR_ARG, "layernorm_bwd: pointers must be 16-byte aligned");
  if (rows == 0) return 0;
  if (!gamma2 || split_row <= 0 || split_row >= rows) split_row = rows;
  if (dropout_p == 0.f) dx_drop = nullptr;
  if (dtype == VB_BF16)
    return vb::ln::bwd_dispatch<vb::bf16>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, rows, H, out_scale, dx_drop,
                                          dropout_p, seed, offset, dbias, vb::ln::make_seqs(seqs), gamma2, dgamma2, dbeta2, dbias2, split_row, (cudaStream_t)stream);
  VB_REQUIRE(dtype == VB_F32, VB_ERR_ARG, "layernorm_bwd: bad dtype %d", dtype);
  return vb::ln::bwd_dispatch<float>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, rows, H, out_scale, dx_drop,
                                     dropout_p, seed, offset, dbias, vb::ln::make_seqs(seqs), gamma2, dgamma2, dbeta2, dbias2, split_row, (cudaStream_t)stream);
}

}  // extern "C"
