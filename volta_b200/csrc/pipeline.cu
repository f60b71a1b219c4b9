// Region batch assembly on the device - the step of the reference's input pipeline that touches the big tensor.
//
// The reference assembles every pre-training batch on the host with numpy: box area + normalisation
// (volta/datasets/concept_cap_dataset.py:450-461), zeroing of the regions picked for prediction and exclusion of the
// regions that overlap them (IoU > 0.4, :31-68, :661-690), padding masks / labels (:562-568), and the mean-pooled
// global feature, its box and its mask prepended or appended to every sample (:237-264).  At 20 k samples/s that is
// 6 GB/s of fp32 features through numpy; here the detector output is uploaded once (fp32 or bf16) and two kernels
// write image_feat / image_loc / image_mask / image_label in the layout and dtype the encoder consumes.
//
// Arithmetic is the reference's, operation for operation (fp32, no contraction), so fp32 outputs are bit-identical:
// the feature sum runs row after row, the mean is one correctly rounded division (the reference divides in fp64 and
// casts back to fp32, which rounds to the same value), IoU and box terms use explicit round-to-nearest intrinsics.
#include "common.cuh"

namespace vb {
namespace pipeline {

constexpr int META_THREADS = 128;
constexpr int FEAT_THREADS = 128;
constexpr int MAX_REGIONS = 1024;

// One block per sample: boxes -> image_loc, padding mask, labels, and the divisor of the global feature.
__global__ void __launch_bounds__(META_THREADS)
meta_kernel(const float* __restrict__ raw_box, const float* __restrict__ image_hw, const int32_t* __restrict__ num_boxes,
            const uint8_t* __restrict__ sel, int R, int num_locs, int global_pos, float* __restrict__ loc,
            int64_t* __restrict__ mask, int64_t* __restrict__ label, int32_t* __restrict__ count) {
  __shared__ float4 box_s[MAX_REGIONS];
  __shared__ float area_s[MAX_REGIONS];
  __shared__ uint8_t sel_s[MAX_REGIONS];
  __shared__ int32_t warp_cnt[META_THREADS / 32];
  const int b = blockIdx.x, t = threadIdx.x;
  const int nb = max(0, min(num_boxes[b], R));
  const int V = R + (global_pos != 0 ? 1 : 0), off = global_pos == 1 ? 1 : 0;
  const float h = image_hw[2 * b], w = image_hw[2 * b + 1];
  const float wh = __fmul_rn(w, h);
  float* loc_b = loc + (int64_t)b * V * num_locs;
  int64_t* mask_b = mask + (int64_t)b * V;
  for (int j = t; j < R; j += META_THREADS) {
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < nb) bx = *reinterpret_cast<const float4*>(raw_box + ((int64_t)b * R + j) * 4);
    box_s[j] = bx;
    area_s[j] = __fmul_rn(__fadd_rn(__fsub_rn(bx.z, bx.x), 1.f), __fadd_rn(__fsub_rn(bx.w, bx.y), 1.f));
    const uint8_t s = j < nb ? sel[(int64_t)b * R + j] : (uint8_t)0;
    sel_s[j] = s;
    float* o = loc_b + (int64_t)(j + off) * num_locs;
    o[0] = __fdiv_rn(bx.x, w);
    o[1] = __fdiv_rn(bx.y, h);
    o[2] = __fdiv_rn(bx.z, w);
    o[3] = __fdiv_rn(bx.w, h);
    if (num_locs == 5) o[4] = __fdiv_rn(__fmul_rn(__fsub_rn(bx.w, bx.y), __fsub_rn(bx.z, bx.x)), wh);
    mask_b[j + off] = j < nb ? 1 : 0;
    label[(int64_t)b * R + j] = s != 0 ? 1 : -1;
  }
  if (global_pos != 0 && t == 0) {
    const int g = global_pos == 1 ? 0 : R;
    float* o = loc_b + (int64_t)g * num_locs;
    o[0] = 0.f; o[1] = 0.f; o[2] = 1.f; o[3] = 1.f;
    for (int k = 4; k < num_locs; ++k) o[k] = 1.f;
    mask_b[g] = 1;
  }
  __syncthreads();
  // region j leaves the mean's divisor when it overlaps (IoU > 0.4) any region picked for prediction, itself included
  int32_t mine = 0;
  for (int j = t; j < R; j += META_THREADS) {
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
  if (t == 0) {
    int32_t c = 0;
    for (int k = 0; k < META_THREADS / 32; ++k) c += warp_cnt[k];
    count[b] = c != 0 ? c : 1;
  }
}

template <typename T> struct Quad;   // 4 consecutive elements
template <> struct Quad<float> {
  static __device__ __forceinline__ float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Quad<bf16> {
  static __device__ __forceinline__ float4 load(const bf16* p) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    const float2 a = unpack_bf16x2(r.x), c = unpack_bf16x2(r.y);
    return make_float4(a.x, a.y, c.x, c.y);
  }
  static __device__ __forceinline__ void store(bf16* p, float4 v) {
    uint2 r;
    r.x = pack_bf16x2(v.x, v.y);
    r.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = r;
  }
};

// grid (column groups, samples); a thread owns 4 columns (W = 4) or 1 (W = 1, any F) of one sample and walks the
// regions: copy (or zero) the row, add it to the running sum in the reference's order, finish with the global row.
// Rows that are zeroed or beyond the box count are not read.
template <typename TIn, typename TOut, int W>
__global__ void __launch_bounds__(FEAT_THREADS)
feat_kernel(const TIn* __restrict__ raw, const int32_t* __restrict__ num_boxes, const uint8_t* __restrict__ sel,
            const int32_t* __restrict__ count, int R, int F, int global_pos, TOut* __restrict__ out) {
  __shared__ uint8_t live_s[MAX_REGIONS];
  const int b = blockIdx.y;
  const int nb = max(0, min(num_boxes[b], R));
  for (int j = threadIdx.x; j < R; j += FEAT_THREADS) live_s[j] = (j < nb && sel[(int64_t)b * R + j] != 1) ? 1 : 0;
  __syncthreads();
  const int c = (blockIdx.x * FEAT_THREADS + threadIdx.x) * W;
  if (c >= F) return;
  const int V = R + (global_pos != 0 ? 1 : 0), off = global_pos == 1 ? 1 : 0;
  const TIn* src = raw + (int64_t)b * R * F + c;
  TOut* dst = out + ((int64_t)b * V + off) * F + c;
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int U = 6;   // loads in flight per thread (36 regions = 6 rounds)
  for (int r0 = 0; r0 < R; r0 += U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = r0 + u;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < R && live_s[r]) {
        if (W == 4) v[u] = Quad<TIn>::load(src + (int64_t)r * F);
        else v[u].x = to_f<TIn>(src[(int64_t)r * F]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = r0 + u;
      if (r >= R) break;
      sum.x = __fadd_rn(sum.x, v[u].x);
      if (W == 4) {
        sum.y = __fadd_rn(sum.y, v[u].y); sum.z = __fadd_rn(sum.z, v[u].z); sum.w = __fadd_rn(sum.w, v[u].w);
        Quad<TOut>::store(dst + (int64_t)r * F, v[u]);
      } else {
        dst[(int64_t)r * F] = from_f<TOut>(v[u].x);
      }
    }
  }
  if (global_pos != 0) {
    const float n = (float)count[b];
    TOut* g = out + ((int64_t)b * V + (global_pos == 1 ? 0 : R)) * F + c;
    if (W == 4) {
      Quad<TOut>::store(g, make_float4(__fdiv_rn(sum.x, n), __fdiv_rn(sum.y, n), __fdiv_rn(sum.z, n), __fdiv_rn(sum.w, n)));
    } else {
      g[0] = from_f<TOut>(__fdiv_rn(sum.x, n));
    }
  }
}

template <typename TIn, typename TOut>
static void launch_feat(const void* raw, const int32_t* num_boxes, const uint8_t* sel, const int32_t* count, int B, int R,
                        int F, int global_pos, void* out, cudaStream_t st) {
  const bool vec = F % 4 == 0 && aligned16(raw) && aligned16(out);
  const int per_block = FEAT_THREADS * (vec ? 4 : 1);
  dim3 grid((unsigned)ceil_div(F, per_block), (unsigned)B);
  if (vec)
    feat_kernel<TIn, TOut, 4><<<grid, FEAT_THREADS, 0, st>>>((const TIn*)raw, num_boxes, sel, count, R, F, global_pos, (TOut*)out);
  else
    feat_kernel<TIn, TOut, 1><<<grid, FEAT_THREADS, 0, st>>>((const TIn*)raw, num_boxes, sel, count, R, F, global_pos, (TOut*)out);
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
  pipeline::meta_kernel<<<B, pipeline::META_THREADS, 0, st>>>(raw_box, image_hw, num_boxes, sel, R, num_locs, global_pos,
                                                             loc, mask, label, count);
  if (int rc = check_launch("vb_region_batch(meta)")) return rc;
  if (in_dtype == VB_F32 && out_dtype == VB_F32)
    pipeline::launch_feat<float, float>(raw_feat, num_boxes, sel, count, B, R, F, global_pos, feat, st);
  else if (in_dtype == VB_F32)
    pipeline::launch_feat<float, bf16>(raw_feat, num_boxes, sel, count, B, R, F, global_pos, feat, st);
  else if (out_dtype == VB_F32)
    pipeline::launch_feat<bf16, float>(raw_feat, num_boxes, sel, count, B, R, F, global_pos, feat, st);
  else
    pipeline::launch_feat<bf16, bf16>(raw_feat, num_boxes, sel, count, B, R, F, global_pos, feat, st);
  return check_launch("vb_region_batch(feat)");
}
