// Region batch assembly on the device - the step of the reference's input pipeline that touches the big tensor.
//
// The reference assembles every pre-training batch on the host with numpy: box area + normalisation
// (volta/datasets/concept_cap_dataset.py:450-461), zeroing of the regions picked for prediction and exclusion of the
// regions that overlap them (IoU > 0.4, :31-68, :661-690), padding masks / labels (:562-568), and the mean-pooled
// global feature, its box and its mask prepended or appended to every sample (:237-264).  At 20 k samples/s that is
// 6 GB/s of fp32 features through numpy; here the detector output is uploaded once (fp32 or bf16) and one kernel
// writes image_feat / image_loc / image_mask / image_label in the layout and dtype the encoder consumes.
//
// Arithmetic is the reference's, operation for operation (fp32, no contraction), so fp32 outputs are bit-identical:
// the feature sum runs row after row, the mean is one correctly rounded division (the reference divides in fp64 and
// casts back to fp32, which rounds to the same value), IoU and box terms use explicit round-to-nearest intrinsics.
#include "common.cuh"

namespace vb {
namespace pipeline {

constexpr int THREADS = 128;
constexpr int MAX_REGIONS = 1024;
constexpr int U = 10;  // row loads in flight per thread: the per-thread chain is latency-bound; 10 keeps the fp32 kernels at 72 registers (one wave)

// W consecutive elements <-> W floats; W = 16 bytes of the INPUT type, so every load is a 16-byte vector.
template <typename T, int W> struct Pack;
template <> struct Pack<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 r = *reinterpret_cast<const float4*>(p);
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Pack<float, 8> {
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Pack<bf16, 4> {
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[4]) {
    uint2 r;
    r.x = pack_bf16x2(v[0], v[1]);
    r.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = r;
  }
};
template <> struct Pack<bf16, 8> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) { Vec8<bf16>::load(p, v); }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) { Vec8<bf16>::store(p, v); }
};
template <typename T> struct Pack<T, 1> {
  static __device__ __forceinline__ void load(const T* p, float (&v)[1]) { v[0] = to_f<T>(*p); }
  static __device__ __forceinline__ void store(T* p, const float (&v)[1]) { *p = from_f<T>(v[0]); }
};

// ONE launch, grid (column groups, samples).  A thread owns W columns of one sample and walks the live regions (a
// block-wide list built once): copy the row, add it to the running sum in the reference's order; the other rows are
// written as zeros (adding them would not change the sum).  While those loads are in flight nothing else is
// needed; the divisor of the mean (regions that do not overlap a picked one) is computed by the block afterwards from
// the sample's boxes in shared memory, and the block of the first column group also writes the sample's boxes, mask
// and labels.  Rows that are zeroed or beyond the box count are not read.
template <typename TIn, typename TOut, int W>
__global__ void __launch_bounds__(THREADS, W == 4 ? 7 : 4)   // fp32 input: 1024 blocks at B = 256 stay ONE wave on 148 SMs
region_batch_kernel(const TIn* __restrict__ raw, const float* __restrict__ raw_box, const float* __restrict__ image_hw,
                    const int32_t* __restrict__ num_boxes, const uint8_t* __restrict__ sel, int R, int F, int num_locs,
                    int global_pos, TOut* __restrict__ out, float* __restrict__ loc, int64_t* __restrict__ mask,
                    int64_t* __restrict__ label, int32_t* __restrict__ count) {
  __shared__ float4 box_s[MAX_REGIONS];
  __shared__ float area_s[MAX_REGIONS];
  __shared__ uint8_t sel_s[MAX_REGIONS];
  __shared__ uint16_t rows_s[MAX_REGIONS];
  __shared__ int32_t warp_cnt[THREADS / 32];
  __shared__ int32_t n_live_s;
  const int b = blockIdx.y, t = threadIdx.x;
  const int nb = max(0, min(num_boxes[b], R));
  const int V = R + (global_pos != 0 ? 1 : 0), off = global_pos == 1 ? 1 : 0;
  // the picks decide which rows are read at all; the boxes are only needed after the feature pass - request this
  // thread's first box now, park it in a register, and fill shared memory once the feature loads have been issued
  for (int j = t; j < R; j += THREADS) sel_s[j] = j < nb ? sel[(int64_t)b * R + j] : (uint8_t)0;
  float4 box0 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t < nb) box0 = *reinterpret_cast<const float4*>(raw_box + ((int64_t)b * R + t) * 4);
  __syncthreads();
  // row lists (warp 0, ballot compaction, ascending): `rows_s[0, n_live)` are read and summed, the rows from the end
  // of the array downwards (`rows_s[R-1-k]`, k < R - n_live) are written as zeros - the feature loop then carries no
  // per-row decisions
  if (t < 32) {
    int n_l = 0, n_d = 0;
    for (int base = 0; base < R; base += 32) {
      const int j = base + t;
      const bool in = j < R;
      const bool lv = in && j < nb && sel_s[j] != 1;
      const unsigned ml = __ballot_sync(0xffffffffu, lv), md = __ballot_sync(0xffffffffu, in && !lv);
      const unsigned below = (1u << t) - 1u;
      if (lv) rows_s[n_l + __popc(ml & below)] = (uint16_t)j;
      else if (in) rows_s[R - 1 - (n_d + __popc(md & below))] = (uint16_t)j;
      n_l += __popc(ml);
      n_d += __popc(md);
    }
    if (t == 0) n_live_s = n_l;
  }
  __syncthreads();
  const int n_live = n_live_s;

  // ---- features: copy the live rows with a running sum (the reference's order), zero the others ----
  const int c = (blockIdx.x * THREADS + t) * W;
  const bool active = c < F;
  float sum[W];
#pragma unroll
  for (int w = 0; w < W; ++w) sum[w] = 0.f;
  if (active) {
    const TIn* src = raw + (int64_t)b * R * F + c;
    TOut* dst = out + ((int64_t)b * V + off) * F + c;
    int k0 = 0;
    for (; k0 + U <= n_live; k0 += U) {            // full rounds: U loads in flight, no predicates
      float v[U][W];
#pragma unroll
      for (int u = 0; u < U; ++u) Pack<TIn, W>::load(src + (int64_t)rows_s[k0 + u] * F, v[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int w = 0; w < W; ++w) sum[w] = __fadd_rn(sum[w], v[u][w]);
        Pack<TOut, W>::store(dst + (int64_t)rows_s[k0 + u] * F, v[u]);
      }
    }
    if (k0 < n_live) {                             // last, partial round
      float v[U][W];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (k0 + u < n_live) Pack<TIn, W>::load(src + (int64_t)rows_s[k0 + u] * F, v[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k0 + u < n_live) {
#pragma unroll
          for (int w = 0; w < W; ++w) sum[w] = __fadd_rn(sum[w], v[u][w]);
          Pack<TOut, W>::store(dst + (int64_t)rows_s[k0 + u] * F, v[u]);
        }
      }
    }
    float zero[W];
#pragma unroll
    for (int w = 0; w < W; ++w) zero[w] = 0.f;
    for (int k = n_live; k < R; ++k) Pack<TOut, W>::store(dst + (int64_t)rows_s[k] * F, zero);
  }

  for (int j = t; j < R; j += THREADS) {
    float4 bx = box0;
    if (j != t) {
      bx = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < nb) bx = *reinterpret_cast<const float4*>(raw_box + ((int64_t)b * R + j) * 4);
    }
    box_s[j] = bx;
    area_s[j] = __fmul_rn(__fadd_rn(__fsub_rn(bx.z, bx.x), 1.f), __fadd_rn(__fsub_rn(bx.w, bx.y), 1.f));
  }
  __syncthreads();

  // ---- region j leaves the mean's divisor when it overlaps (IoU > 0.4) any picked region, itself included ----
  int32_t mine = 0;
  for (int j = t; j < R; j += THREADS) {
    bool masked = false;
    if (j < nb) {
      const float4 bj = box_s[j];
      const float aj = area_s[j];
      for (int i = 0; i < nb && !masked; ++i) {
        if (sel_s[i] == 0) continue;
        const float4 bi = box_s[i];
        float iw = __fadd_rn(__fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)), 1.f);
        float ih = __fadd_rn(__fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y)), 1.f);
        iw = iw < 0.f ? 0.f : iw;
        ih = ih < 0.f ? 0.f : ih;
        const float inter = __fmul_rn(iw, ih);
        const float ua = __fsub_rn(__fadd_rn(area_s[i], aj), inter);
        masked = __fdiv_rn(inter, ua) > 0.4f;
      }
    }
    mine += masked ? 0 : 1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((t & 31) == 0) warp_cnt[t >> 5] = mine;
  __syncthreads();
  int32_t n_keep = 0;
#pragma unroll
  for (int k = 0; k < THREADS / 32; ++k) n_keep += warp_cnt[k];
  if (n_keep == 0) n_keep = 1;

  if (active && global_pos != 0) {
    const float n = (float)n_keep;
    float g[W];
#pragma unroll
    for (int w = 0; w < W; ++w) g[w] = __fdiv_rn(sum[w], n);
    Pack<TOut, W>::store(out + ((int64_t)b * V + (global_pos == 1 ? 0 : R)) * F + c, g);
  }

  // ---- boxes, padding mask, labels: once per sample ----
  if (blockIdx.x != 0) return;
  const float h = image_hw[2 * b], wd = image_hw[2 * b + 1];
  const float wh = __fmul_rn(wd, h);
  float* loc_b = loc + (int64_t)b * V * num_locs;
  int64_t* mask_b = mask + (int64_t)b * V;
  for (int j = t; j < R; j += THREADS) {
    const float4 bx = box_s[j];
    float* o = loc_b + (int64_t)(j + off) * num_locs;
    o[0] = __fdiv_rn(bx.x, wd);
    o[1] = __fdiv_rn(bx.y, h);
    o[2] = __fdiv_rn(bx.z, wd);
    o[3] = __fdiv_rn(bx.w, h);
    if (num_locs == 5) o[4] = __fdiv_rn(__fmul_rn(__fsub_rn(bx.w, bx.y), __fsub_rn(bx.z, bx.x)), wh);
    mask_b[j + off] = j < nb ? 1 : 0;
    label[(int64_t)b * R + j] = sel_s[j] != 0 ? 1 : -1;
  }
  if (t == 0) {
    count[b] = n_keep;
    if (global_pos != 0) {
      const int g = global_pos == 1 ? 0 : R;
      float* o = loc_b + (int64_t)g * num_locs;
      o[0] = 0.f; o[1] = 0.f; o[2] = 1.f; o[3] = 1.f;
      for (int k = 4; k < num_locs; ++k) o[k] = 1.f;
      mask_b[g] = 1;
    }
  }
}

template <typename TIn, typename TOut>
static void launch(const void* raw, const float* raw_box, const float* image_hw, const int32_t* num_boxes, const uint8_t* sel,
                   int B, int R, int F, int num_locs, int global_pos, void* out, float* loc, int64_t* mask, int64_t* label,
                   int32_t* count, cudaStream_t st) {
  constexpr int WV = 16 / (int)sizeof(TIn);
  const bool vec = F % WV == 0 && aligned16(raw) && aligned16(out);
  const int per_block = THREADS * (vec ? WV : 1);
  dim3 grid((unsigned)ceil_div(F, per_block), (unsigned)B);
  if (vec)
    region_batch_kernel<TIn, TOut, WV><<<grid, THREADS, 0, st>>>((const TIn*)raw, raw_box, image_hw, num_boxes, sel, R, F, num_locs,
                                                                 global_pos, (TOut*)out, loc, mask, label, count);
  else
    region_batch_kernel<TIn, TOut, 1><<<grid, THREADS, 0, st>>>((const TIn*)raw, raw_box, image_hw, num_boxes, sel, R, F, num_locs,
                                                                global_pos, (TOut*)out, loc, mask, label, count);
}

}  // namespace pipeline
}  // namespace vb

extern "C" int vb_region_batch(const void* raw_feat, int32_t in_dtype, const float* raw_box, const float* image_hw,
                               const int32_t* num_boxes, const uint8_t* sel, int32_t B, int32_t R, int32_t F,
                               int32_t num_locs, int32_t global_pos, void* feat, int32_t out_dtype, float* loc,
                               int64_t* mask, int64_t* label, int32_t* count, void* stream) {
  using namespace vb;
  VB_REQUIRE(raw_feat && raw_box && image_hw && num_boxes && sel && feat && loc && mask && label && count, VB_ERR_ARG,
             "vb_region_batch: null pointer");
  VB_REQUIRE(B >= 0 && R >= 1 && R <= pipeline::MAX_REGIONS && F >= 1, VB_ERR_ARG,
             "vb_region_batch: B=%d R=%d (1..%d) F=%d", B, R, pipeline::MAX_REGIONS, F);
  VB_REQUIRE(B <= 65535, VB_ERR_ARG, "vb_region_batch: B=%d exceeds 65535 samples per call", B);
  VB_REQUIRE(num_locs == 4 || num_locs == 5, VB_ERR_UNSUPPORTED, "vb_region_batch: num_locs=%d (4 or 5)", num_locs);
  VB_REQUIRE(global_pos >= 0 && global_pos <= 2, VB_ERR_ARG, "vb_region_batch: global_pos=%d (0 none, 1 first, 2 last)", global_pos);
  VB_REQUIRE((in_dtype == VB_F32 || in_dtype == VB_BF16) && (out_dtype == VB_F32 || out_dtype == VB_BF16), VB_ERR_ARG,
             "vb_region_batch: dtypes %d -> %d", in_dtype, out_dtype);
  VB_REQUIRE(aligned16(raw_box), VB_ERR_ARG, "vb_region_batch: raw_box must be 16-byte aligned");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
#define VB_PIPE(TI_, TO_) \
  pipeline::launch<TI_, TO_>(raw_feat, raw_box, image_hw, num_boxes, sel, B, R, F, num_locs, global_pos, feat, loc, mask, label, count, st)
  if (in_dtype == VB_F32 && out_dtype == VB_F32) VB_PIPE(float, float);
  else if (in_dtype == VB_F32) VB_PIPE(float, bf16);
  else if (out_dtype == VB_F32) VB_PIPE(bf16, float);
  else VB_PIPE(bf16, bf16);
#undef VB_PIPE
  return check_launch("vb_region_batch");
}
