// Element-wise, gather and reduction glue kernels: every small ATen op the reference issues around the
// GEMMs (relu / mul / dropout on pooled vectors, bias-gradient column sums, row gathers, fp32->bf16 weight
// casts, VL-BERT sin/cos box embedding).  All are HBM-bound; 16-byte vector paths where alignment allows.
#include "common.cuh"

namespace vb {
namespace elt {

__device__ __forceinline__ float load_any(const void* p, int dtype, int64_t i) {
  return dtype == VB_BF16 ? to_f<bf16>(reinterpret_cast<const bf16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void store_any(void* p, int dtype, int64_t i, float v) {
  if (dtype == VB_BF16) reinterpret_cast<bf16*>(p)[i] = from_f<bf16>(v);
  else reinterpret_cast<float*>(p)[i] = v;
}

__device__ __forceinline__ float apply(int op, float a, float b, float o, float alpha, float scale, bool keep) {
  switch (op) {
    case VB_ELT_COPY: return a;
    case VB_ELT_ADD: return a + b;
    case VB_ELT_MUL: return a * b;
    case VB_ELT_RELU: return fmaxf(a, 0.f);
    case VB_ELT_DRELU: return b > 0.f ? a : 0.f;
    case VB_ELT_GELU: return gelu_erf(a);
    case VB_ELT_DGELU: return a * dgelu_erf(b);
    case VB_ELT_SCALE: return a * alpha;
    case VB_ELT_DROPOUT: return keep ? a * scale : 0.f;
    case VB_ELT_AXPY: return o + alpha * a;
    case VB_ELT_SWISH: return a / (1.f + __expf(-a));                               // x * sigmoid(x), volta/encoders.py:153-154
    case VB_ELT_DSWISH: { const float sg = 1.f / (1.f + __expf(-a)); return sg * (1.f + a * (1.f - sg)); }   // d/dx of the above
    default: return a;
  }
}

__global__ void __launch_bounds__(256)
eltwise_kernel(int op, const void* a, int a_dt, int64_t lda, const void* b, int b_dt, int64_t ldb, void* out, int o_dt,
               int64_t ldo, int64_t rows, int64_t cols, float alpha, uint32_t thresh, float scale, uint64_t seed_in,
               uint64_t offset, const uint64_t* seed_ptr) {
  pdl_enter();
  const uint64_t seed = effective_seed(seed_in, seed_ptr);
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    const float av = load_any(a, a_dt, r * lda + c);
    const float bv = b ? load_any(b, b_dt, r * ldb + c) : 0.f;
    const float ov = (op == VB_ELT_AXPY) ? load_any(out, o_dt, r * ldo + c) : 0.f;
    const bool keep = (op == VB_ELT_DROPOUT && thresh) ? dropout_keep(seed, offset, (uint64_t)i, thresh) : true;
    store_any(out, o_dt, r * ldo + c, apply(op, av, bv, ov, alpha, scale, keep));
  }
}

// Vector path: 8 elements (one 16-byte load of bf16) per thread and step.  The scalar kernel above moves 14 MB in
// ~34 us (0.8 TB/s); the pass-through copies, dropouts and casts of a step go through this one instead.
template <typename TA, typename TO>
__global__ void __launch_bounds__(256)
eltwise_vec_kernel(int op, const TA* __restrict__ a, int64_t lda, const TA* __restrict__ b, int64_t ldb, TO* out, int64_t ldo,
                   int64_t rows, int64_t cols, float alpha, uint32_t thresh, float scale, uint64_t seed_in, uint64_t offset,
                   const uint64_t* seed_ptr) {
  pdl_enter();
  const uint64_t seed = effective_seed(seed_in, seed_ptr);
  const int64_t cv = cols >> 3, total = rows * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cv, c = (i - r * cv) << 3;
    float av[8], bv[8], ov[8], res[8];
    Vec8<TA>::load(a + r * lda + c, av);
    if (b) Vec8<TA>::load(b + r * ldb + c, bv);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) bv[j] = 0.f;
    }
    if (op == VB_ELT_AXPY) Vec8<TO>::load(out + r * ldo + c, ov);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) ov[j] = 0.f;
    }
    bool keep[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) keep[j] = true;
    if (op == VB_ELT_DROPOUT && thresh) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) dropout_keep2(seed, offset, (uint64_t)(r * cols + c + j), thresh, keep[j], keep[j + 1]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) res[j] = apply(op, av[j], bv[j], ov[j], alpha, scale, keep[j]);
    Vec8<TO>::store(out + r * ldo + c, res);
  }
}

// column sums: block handles a 32-column strip x a slab of rows; lanes walk columns (coalesced), warps walk rows.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ a, int64_t lda, int64_t rows, int64_t cols, float* __restrict__ out, int64_t rows_per_block,
              const int32_t* __restrict__ row_limit) {
  __shared__ float part[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + lane;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  if (row_limit) rows = min(rows, (int64_t)*row_limit);   // device-side row count
  if (r0 >= rows) return;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  if (c < cols)
    for (int64_t r = r0 + warp; r < r1; r += 8) s += to_f<T>(a[r * lda + c]);
  part[warp][lane] = s;
  __syncthreads();
  if (warp == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w][lane];
    if (c < cols) atomicAdd(out + c, t);
  }
}

// Vector version (the bias gradients of the big activations: rows ~ 9e3, cols 768..3072): every lane owns 8
// adjacent columns (one 16-byte load of bf16), a warp covers 256 columns of a row, the 8 warps of the block walk
// the slab's rows with 4 loads in flight each; partials meet in shared memory, one atomicAdd per column and block.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const T* __restrict__ a, int64_t lda, int64_t rows, int64_t cols, float* __restrict__ out, int64_t rows_per_block,
                  const int32_t* __restrict__ row_limit) {
  pdl_enter();
  __shared__ float part[8][32][9];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t c = ((int64_t)blockIdx.x * 32 + lane) * 8;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  if (row_limit) rows = min(rows, (int64_t)*row_limit);
  if (r0 >= rows) return;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < cols) {
    int64_t r = r0 + warp;
    for (; r + 24 < r1; r += 32) {
      float v0[8], v1[8], v2[8], v3[8];
      Vec8<T>::load(a + r * lda + c, v0);
      Vec8<T>::load(a + (r + 8) * lda + c, v1);
      Vec8<T>::load(a + (r + 16) * lda + c, v2);
      Vec8<T>::load(a + (r + 24) * lda + c, v3);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += (v0[j] + v1[j]) + (v2[j] + v3[j]);
    }
    for (; r < r1; r += 8) {
      float v0[8];
      Vec8<T>::load(a + r * lda + c, v0);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v0[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) part[warp][lane][j] = acc[j];
  __syncthreads();
  const int l2 = threadIdx.x >> 3, j2 = threadIdx.x & 7;
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += part[w][l2][j2];
  const int64_t oc = ((int64_t)blockIdx.x * 32 + l2) * 8 + j2;
  if (oc < cols) atomicAdd(out + oc, t);
}

template <typename T, bool SCATTER>
__global__ void __launch_bounds__(256)
rows_kernel(const T* __restrict__ a, int64_t lda, const int64_t* __restrict__ idx, T* __restrict__ out, int64_t ldo,
            int64_t rows, int64_t cols) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    const int64_t s = idx[r];
    if (SCATTER) {
      // out[idx[r], c] += a[r, c]; rows of `idx` may repeat, fp32 only uses atomics, bf16 requires unique idx
      if (sizeof(T) == 4) atomicAdd(reinterpret_cast<float*>(out) + s * ldo + c, to_f<T>(a[r * lda + c]));
      else out[s * ldo + c] = from_f<T>(to_f<T>(out[s * ldo + c]) + to_f<T>(a[r * lda + c]));
    } else {
      out[r * ldo + c] = a[s * lda + c];
    }
  }
}

__global__ void __launch_bounds__(256)
multi_cast_kernel(const vb_cast_item* __restrict__ items, int n_items) {
  const vb_cast_item it = items[blockIdx.y];
  const int64_t n8 = it.n >> 3;
  bf16* dst = reinterpret_cast<bf16*>(it.dst);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
    Vec8<float>::load(it.src + i * 8, v);
    Vec8<bf16>::store(dst + i * 8, v);
  }
  if (blockIdx.x == 0 && threadIdx.x < (it.n & 7)) {
    const int64_t i = (n8 << 3) + threadIdx.x;
    dst[i] = from_f<bf16>(it.src[i]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
box_sincos_kernel(const float* __restrict__ boxes, int nl, int dim, T* __restrict__ out, int64_t ldo, int64_t rows) {
  // out[r, q*2*dim + f] = sin(pos_q / 1000^(f/dim)),  out[r, q*2*dim + dim + f] = cos(...)   (embeddings.py:102-124)
  const int64_t total = rows * 4 * dim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (4 * dim);
    const int q = (int)((i / dim) & 3), f = (int)(i % dim);
    const float* bx = boxes + r * nl;
    float pos;
    if (q == 0) pos = (bx[0] + bx[2]) / 2.f * 100.f;
    else if (q == 1) pos = (bx[1] + bx[3]) / 2.f * 100.f;
    else if (q == 2) pos = (bx[2] - bx[0]) * 100.f;
    else pos = (bx[3] - bx[1]) * 100.f;
    const float ang = pos / powf(1000.f, (float)f / (float)dim);
    out[r * ldo + q * 2 * dim + f] = from_f<T>(sinf(ang));
    out[r * ldo + q * 2 * dim + dim + f] = from_f<T>(cosf(ang));
  }
}

__global__ void fill_kernel(int4* buf, int64_t n16, int value) {
  const int4 v = make_int4(value, value, value, value);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) buf[i] = v;
}

static unsigned grid_for(int64_t total, int threads = 256) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, threads), 148 * 16));
}

}  // namespace elt
}  // namespace vb

extern "C" {

int vb_eltwise(int32_t op, const void* a, int32_t a_dtype, int64_t lda, const void* b, int32_t b_dtype, int64_t ldb,
               void* out, int32_t out_dtype, int64_t ldo, int64_t rows, int64_t cols, float alpha, float p, uint64_t seed,
               uint64_t offset, void* stream) {
  VB_REQUIRE(op >= VB_ELT_COPY && op <= VB_ELT_DSWISH, VB_ERR_ARG, "eltwise: bad op %d", op);
  VB_REQUIRE(a && out, VB_ERR_ARG, "eltwise: null pointer");
  const bool needs_b = (op == VB_ELT_ADD || op == VB_ELT_MUL || op == VB_ELT_DRELU || op == VB_ELT_DGELU);
  VB_REQUIRE(!needs_b || b, VB_ERR_ARG, "eltwise: op %d needs a second operand", op);
  VB_REQUIRE(p >= 0.f && p < 1.f, VB_ERR_ARG, "eltwise: dropout p out of range");
  if (rows * cols == 0) return 0;
  const uint32_t thresh = (op == VB_ELT_DROPOUT && p > 0.f) ? vb::dropout_thresh(p) : 0u;
  const bool vec_ok = (cols % 8) == 0 && (lda % 8) == 0 && (ldo % 8) == 0 && (!needs_b || ((ldb % 8) == 0 && b_dtype == a_dtype)) &&
                      vb::aligned16(a) && vb::aligned16(out) && (!needs_b || vb::aligned16(b)) && rows * cols >= 4096;
  if (vec_ok) {
    const unsigned g = vb::elt::grid_for(rows * cols / 8);
    const void* bb = needs_b ? b : nullptr;
    const float sc = 1.f / (1.f - p);
#define VB_ELT_VEC(TA_, TO_)                                                                                              \
  vb::launch_pdl(vb::elt::eltwise_vec_kernel<TA_, TO_>, dim3(g), dim3(256), 0, (cudaStream_t)stream, op, (const TA_*)a, lda, \
                 (const TA_*)bb, ldb, (TO_*)out, ldo, rows, cols, alpha, thresh, sc, seed, offset, vb::seed_pointer())
    if (a_dtype == VB_BF16 && out_dtype == VB_BF16) VB_ELT_VEC(vb::bf16, vb::bf16);
    else if (a_dtype == VB_F32 && out_dtype == VB_BF16) VB_ELT_VEC(float, vb::bf16);
    else if (a_dtype == VB_BF16 && out_dtype == VB_F32) VB_ELT_VEC(vb::bf16, float);
    else VB_ELT_VEC(float, float);
#undef VB_ELT_VEC
    return vb::check_launch("eltwise_vec_kernel");
  }
  vb::launch_pdl(vb::elt::eltwise_kernel, dim3(vb::elt::grid_for(rows * cols)), dim3(256), 0, (cudaStream_t)stream,
                 op, a, a_dtype, lda, needs_b ? b : nullptr, b_dtype, ldb, out, out_dtype, ldo, rows, cols, alpha, thresh,
                 1.f / (1.f - p), seed, offset, vb::seed_pointer());
  return vb::check_launch("eltwise_kernel");
}

int vb_colsum(const void* a, int32_t dtype, int64_t lda, int64_t rows, int64_t cols, float* out, const int32_t* row_limit,
              void* stream) {
  VB_REQUIRE(a && out, VB_ERR_ARG, "colsum: null pointer");
  if (rows * cols == 0) return 0;
  const int64_t es = dtype == VB_BF16 ? 2 : 4;
  if (cols % 8 == 0 && lda % 8 == 0 && vb::aligned16(a) && rows >= 64) {
    const int64_t strips = vb::ceil_div(cols, 256);
    int64_t slabs = std::max<int64_t>(1, std::min<int64_t>(vb::ceil_div(rows, 32), (148 * 4) / strips));
    const int64_t rpb = vb::ceil_div(vb::ceil_div(rows, slabs), 8) * 8;
    slabs = vb::ceil_div(rows, rpb);
    dim3 grid((unsigned)strips, (unsigned)slabs);
    if (dtype == VB_BF16)
      vb::launch_pdl(vb::elt::colsum_vec_kernel<vb::bf16>, grid, dim3(256), 0, (cudaStream_t)stream, (const vb::bf16*)a, lda, rows, cols, out, rpb, row_limit);
    else
      vb::launch_pdl(vb::elt::colsum_vec_kernel<float>, grid, dim3(256), 0, (cudaStream_t)stream, (const float*)a, lda, rows, cols, out, rpb, row_limit);
    (void)es;
    return vb::check_launch("colsum_vec_kernel");
  }
  const int64_t strips = vb::ceil_div(cols, 32);
  int64_t slabs = std::max<int64_t>(1, std::min<int64_t>(vb::ceil_div(rows, 64), (148 * 8) / std::max<int64_t>(1, strips)));
  const int64_t rpb = vb::ceil_div(rows, slabs);
  slabs = vb::ceil_div(rows, rpb);
  dim3 grid((unsigned)strips, (unsigned)slabs);
  if (dtype == VB_BF16)
    vb::elt::colsum_kernel<vb::bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const vb::bf16*)a, lda, rows, cols, out, rpb, row_limit);
  else
    vb::elt::colsum_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)a, lda, rows, cols, out, rpb, row_limit);
  return vb::check_launch("colsum_kernel");
}

int vb_gather_rows(const void* a, int64_t lda, const int64_t* idx, void* out, int64_t ldo, int64_t rows, int64_t cols,
                   int32_t dtype, void* stream) {
  VB_REQUIRE(a && idx && out, VB_ERR_ARG, "gather_rows: null pointer");
  if (rows * cols == 0) return 0;
  const unsigned g = vb::elt::grid_for(rows * cols);
  if (dtype == VB_BF16)
    vb::elt::rows_kernel<vb::bf16, false><<<g, 256, 0, (cudaStream_t)stream>>>((const vb::bf16*)a, lda, idx, (vb::bf16*)out, ldo, rows, cols);
  else
    vb::elt::rows_kernel<float, false><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)a, lda, idx, (float*)out, ldo, rows, cols);
  return vb::check_launch("gather_rows_kernel");
}

int vb_scatter_rows_add(const void* a, int64_t lda, const int64_t* idx, void* out, int64_t ldo, int64_t rows,
                        int64_t cols, int32_t dtype, void* stream) {
  VB_REQUIRE(a && idx && out, VB_ERR_ARG, "scatter_rows_add: null pointer");
  if (rows * cols == 0) return 0;
  const unsigned g = vb::elt::grid_for(rows * cols);
  if (dtype == VB_BF16)
    vb::elt::rows_kernel<vb::bf16, true><<<g, 256, 0, (cudaStream_t)stream>>>((const vb::bf16*)a, lda, idx, (vb::bf16*)out, ldo, rows, cols);
  else
    vb::elt::rows_kernel<float, true><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)a, lda, idx, (float*)out, ldo, rows, cols);
  return vb::check_launch("scatter_rows_kernel");
}

int vb_multi_cast_bf16(const vb_cast_item* items_dev, int32_t n_items, int64_t max_n, void* stream) {
  VB_REQUIRE(items_dev && n_items > 0, VB_ERR_ARG, "multi_cast: empty item table");
  VB_REQUIRE(n_items <= 65535, VB_ERR_ARG, "multi_cast: too many items");
  const int64_t cap = std::max<int64_t>(64, (148 * 16) / n_items);
  const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(vb::ceil_div(max_n / 8, 256), cap));
  vb::elt::multi_cast_kernel<<<dim3(gx, (unsigned)n_items), 256, 0, (cudaStream_t)stream>>>(items_dev, n_items);
  return vb::check_launch("multi_cast_kernel");
}

int vb_box_sincos(const float* boxes, int32_t nl, int32_t dim, void* out, int64_t ldo, int64_t rows, int32_t dtype,
                  void* stream) {
  VB_REQUIRE(boxes && out && nl >= 4 && dim > 0, VB_ERR_ARG, "box_sincos: bad arguments");
  if (rows == 0) return 0;
  const unsigned g = vb::elt::grid_for(rows * 4 * dim);
  if (dtype == VB_BF16)
    vb::elt::box_sincos_kernel<vb::bf16><<<g, 256, 0, (cudaStream_t)stream>>>(boxes, nl, dim, (vb::bf16*)out, ldo, rows);
  else
    vb::elt::box_sincos_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>(boxes, nl, dim, (float*)out, ldo, rows);
  return vb::check_launch("box_sincos_kernel");
}

int vb_fill(void* buf, int64_t bytes, int32_t value, void* stream) {
  VB_REQUIRE(buf && bytes % 16 == 0 && vb::aligned16(buf), VB_ERR_ARG, "fill: buffer must be 16-byte aligned and sized");
  if (bytes == 0) return 0;
  vb::elt::fill_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>((int4*)buf, bytes / 16, value);
  return vb::check_launch("fill_kernel");
}

}  // extern "C"
