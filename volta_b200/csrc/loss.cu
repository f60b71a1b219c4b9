// Pre-training criteria on device: CrossEntropyLoss(ignore_index=-1) for MLM / ITM
// (volta/encoders.py:1030, 1095-1109) and the masked-region KL `kl_1601` (volta/losses.py:16-22).
// One block per logits row; rows whose label is ignored are never read (their gradient is zero).
// The backward recomputes the softmax from the saved row log-sum-exp and may run in place over the logits.
#include "common.cuh"

namespace vb {
namespace loss {

constexpr int THREADS = 256;

__device__ __forceinline__ float block_max(float v, float* sh) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
#pragma unroll
  for (int w = 1; w < THREADS / 32; ++w) r = fmaxf(r, sh[w]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < THREADS / 32; ++w) r += sh[w];
  __syncthreads();
  return r;
}

// log-sum-exp of one row (block-cooperative).  Vector path when the row is 16-byte aligned.
template <typename T>
__device__ __forceinline__ float row_lse(const T* __restrict__ x, int C, bool vec_ok, float* sh) {
  float mx = -INFINITY;
  const int C8 = vec_ok ? (C & ~7) : 0;
  for (int c = threadIdx.x * 8; c < C8; c += THREADS * 8) {
    float v[8];
    Vec8<T>::load(x + c, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) mx = fmaxf(mx, v[j]);
  }
  for (int c = C8 + threadIdx.x; c < C; c += THREADS) mx = fmaxf(mx, to_f<T>(x[c]));
  mx = block_max(mx, sh);
  float s = 0.f;
  for (int c = threadIdx.x * 8; c < C8; c += THREADS * 8) {
    float v[8];
    Vec8<T>::load(x + c, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += __expf(v[j] - mx);
  }
  for (int c = C8 + threadIdx.x; c < C; c += THREADS) s += __expf(to_f<T>(x[c]) - mx);
  s = block_sum(s, sh);
  return mx + __logf(s);
}

template <typename T>
__global__ void __launch_bounds__(THREADS)
xent_fwd_kernel(const T* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels, int C,
                float* __restrict__ lse_out, float* __restrict__ loss_sum, float* __restrict__ count, bool vec_ok,
                const int32_t* __restrict__ row_limit) {
  __shared__ float sh[THREADS / 32];
  const int64_t r = blockIdx.x;
  if (row_limit && r >= *row_limit) return;   // device-side row count (compacted rows): nothing beyond it exists
  const int64_t lab = labels[r];
  if (lab < 0 || lab >= C) {  // ignore_index = -1
    if (threadIdx.x == 0) lse_out[r] = 0.f;
    return;
  }
  const T* x = logits + r * ld;
  const float lse = row_lse<T>(x, C, vec_ok, sh);
  if (threadIdx.x == 0) {
    lse_out[r] = lse;
    atomicAdd(loss_sum, lse - to_f<T>(x[lab]));
    atomicAdd(count, 1.0f);
  }
}

template <typename T>
__global__ void __launch_bounds__(THREADS)
xent_bwd_kernel(const T* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels, int C,
                const float* __restrict__ lse, const float* __restrict__ count, const float* __restrict__ dloss,
                T* __restrict__ dlogits, bool vec_ok, const int32_t* __restrict__ row_limit) {
  const int64_t r = blockIdx.x;
  if (row_limit && r >= *row_limit) return;
  const int64_t lab = labels[r];
  const T* x = logits + r * ld;
  T* d = dlogits + r * ld;
  const bool live = (lab >= 0 && lab < C);
  const float g = live ? (*dloss) / fmaxf(*count, 1.0f) : 0.f;
  const float l = live ? lse[r] : 0.f;
  const int C8 = vec_ok ? (C & ~7) : 0;
  for (int c = threadIdx.x * 8; c < C8; c += THREADS * 8) {
    float v[8], o[8];
    if (live) {
      Vec8<T>::load(x + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (__expf(v[j] - l) - ((c + j) == lab ? 1.f : 0.f)) * g;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
    }
    Vec8<T>::store(d + c, o);
  }
  for (int c = C8 + threadIdx.x; c < C; c += THREADS)
    d[c] = from_f<T>(live ? (__expf(to_f<T>(x[c]) - l) - (c == lab ? 1.f : 0.f)) * g : 0.f);
}

// ---- kl_1601 ---------------------------------------------------------------------
// prediction row r (0 <= r < rows = B*per) lives at logits row (r/per)*per_ld + r%per + skip.
template <typename T>
__global__ void __launch_bounds__(THREADS)
kl_fwd_kernel(const T* __restrict__ logits, int64_t ld, int per, int per_ld, int skip, const int64_t* __restrict__ label,
              const float* __restrict__ target, int C, float weight, float* __restrict__ lse_out,
              float* __restrict__ loss_sum, float* __restrict__ count, bool vec_ok) {
  __shared__ float sh[THREADS / 32];
  const int64_t r = blockIdx.x;
  if (label[r] != 1) {
    if (threadIdx.x == 0) lse_out[r] = 0.f;
    return;
  }
  const T* x = logits + ((r / per) * per_ld + (r % per) + skip) * ld;
  const float* t = target + r * (int64_t)C;
  const float lse = row_lse<T>(x, C, vec_ok, sh);
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += THREADS) {
    const float tv = t[c];
    if (tv > 0.f) s += tv * (__logf(tv) - (to_f<T>(x[c]) - lse));  // KLDivLoss(log_softmax(pred), target)
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) {
    lse_out[r] = lse;
    atomicAdd(loss_sum, weight * s);
    atomicAdd(count, 1.0f);
  }
}

// grid over ALL logits rows (B*per_ld) so that skipped / unpicked rows get an explicit zero gradient.
template <typename T>
__global__ void __launch_bounds__(THREADS)
kl_bwd_kernel(const T* __restrict__ logits, int64_t ld, int per, int per_ld, int skip, const int64_t* __restrict__ label,
              const float* __restrict__ target, int C, float weight, const float* __restrict__ lse,
              const float* __restrict__ count, const float* __restrict__ dloss, T* __restrict__ dlogits) {
  __shared__ float sh[THREADS / 32];
  const int64_t lr = blockIdx.x;
  const int64_t b = lr / per_ld;
  const int within = (int)(lr % per_ld) - skip;
  const bool is_pred = within >= 0 && within < per;
  const int64_t r = b * per + within;
  const bool live = is_pred && label[r] == 1;
  T* d = dlogits + lr * ld;
  if (!live) {
    for (int c = threadIdx.x; c < C; c += THREADS) d[c] = from_f<T>(0.f);
    return;
  }
  const T* x = logits + lr * ld;
  const float* t = target + r * (int64_t)C;
  float ts = 0.f;
  for (int c = threadIdx.x; c < C; c += THREADS) ts += t[c];
  ts = block_sum(ts, sh);
  const float g = weight * (*dloss) / fmaxf(*count, 1.0f);
  const float l = lse[r];
  for (int c = threadIdx.x; c < C; c += THREADS)
    d[c] = from_f<T>(g * (__expf(to_f<T>(x[c]) - l) * ts - t[c]));
}

// The other masked-region criteria of volta/losses.py, over the same row mapping as kl_1601:
//   VB_REGION_XENT : confidence-weighted cross-entropy against a hard label - xent_1600 / xent_400 (with conf,
//                    :87-107) and xent_1601 (conf == NULL, :125-133):  weight * sum_{label==1} CE_r * conf_r / max(#,1)
//   VB_REGION_MSE / VB_REGION_HUBER : feature regression - mse_2048 (:25-33) / huber_2048 (:110-122); the caller
//                    passes weight / C so that the shared denominator max(#label==1, 1) gives the reference's
//                    mean over masked regions AND feature dimensions.
// `add_count`: only one of the criteria of a step accumulates the region count (they share the denominator).
template <typename T>
__global__ void __launch_bounds__(THREADS)
region_fwd_kernel(int kind, const T* __restrict__ logits, int64_t ld, int per, int per_ld, int skip,
                  const int64_t* __restrict__ label, const int64_t* __restrict__ tgt_idx, const float* __restrict__ conf,
                  const float* __restrict__ tgt_vec, int64_t ld_t, int C, float weight, float* __restrict__ lse_out,
                  float* __restrict__ loss_sum, float* __restrict__ count, int add_count, bool vec_ok) {
  __shared__ float sh[THREADS / 32];
  const int64_t r = blockIdx.x;
  if (label[r] != 1) {
    if (threadIdx.x == 0 && lse_out) lse_out[r] = 0.f;
    return;
  }
  const T* x = logits + ((r / per) * per_ld + (r % per) + skip) * ld;
  float s = 0.f;
  if (kind == VB_REGION_XENT) {
    const float lse = row_lse<T>(x, C, vec_ok, sh);
    const int64_t t = tgt_idx[r];
    if (threadIdx.x == 0) {
      lse_out[r] = lse;
      s = (lse - to_f<T>(x[t])) * (conf ? conf[r] : 1.f);
    }
  } else {
    const float* t = tgt_vec + r * ld_t;
    for (int c = threadIdx.x; c < C; c += THREADS) {
      const float d = to_f<T>(x[c]) - t[c];
      const float a = fabsf(d);
      s += (kind == VB_REGION_MSE) ? d * d : (a < 1.f ? 0.5f * d * d : a - 0.5f);   // SmoothL1Loss, beta = 1
    }
    s = block_sum(s, sh);
  }
  if (threadIdx.x == 0) {
    atomicAdd(loss_sum, weight * s);
    if (add_count) atomicAdd(count, 1.0f);
  }
}

template <typename T>
__global__ void __launch_bounds__(THREADS)
region_bwd_kernel(int kind, const T* __restrict__ logits, int64_t ld, int per, int per_ld, int skip,
                  const int64_t* __restrict__ label, const int64_t* __restrict__ tgt_idx, const float* __restrict__ conf,
                  const float* __restrict__ tgt_vec, int64_t ld_t, int C, float weight, const float* __restrict__ lse,
                  const float* __restrict__ count, const float* __restrict__ dloss, T* __restrict__ dlogits) {
  const int64_t lr = blockIdx.x;
  const int64_t b = lr / per_ld;
  const int within = (int)(lr % per_ld) - skip;
  const bool is_pred = within >= 0 && within < per;
  const int64_t r = b * per + within;
  const bool live = is_pred && label[r] == 1;
  T* d = dlogits + lr * ld;
  if (!live) {
    for (int c = threadIdx.x; c < C; c += THREADS) d[c] = from_f<T>(0.f);
    return;
  }
  const T* x = logits + lr * ld;
  const float g = weight * (*dloss) / fmaxf(*count, 1.0f);
  if (kind == VB_REGION_XENT) {
    const float gc = g * (conf ? conf[r] : 1.f), l = lse[r];
    const int64_t t = tgt_idx[r];
    for (int c = threadIdx.x; c < C; c += THREADS)
      d[c] = from_f<T>(gc * (__expf(to_f<T>(x[c]) - l) - (c == t ? 1.f : 0.f)));
  } else {
    const float* t = tgt_vec + r * ld_t;
    for (int c = threadIdx.x; c < C; c += THREADS) {
      const float e = to_f<T>(x[c]) - t[c];
      d[c] = from_f<T>(g * ((kind == VB_REGION_MSE) ? 2.f * e : fminf(fmaxf(e, -1.f), 1.f)));
    }
  }
}

// ---- task criteria (volta/task_utils.py:24-27 LossMap; used at :103-137 and :237-282) -------------------------------
// BCEWithLogitsLoss(reduction="mean") over a [rows, C] logit matrix with fp32 soft targets (VQA / GQA answer scores,
// RefCOCO region scores):  l = max(x, 0) - x t + log(1 + exp(-|x|))  - the stable form PyTorch evaluates.
// One block per row writes the row's partial sum; a single block adds the partial sums in a fixed order (deterministic,
// no atomics) and scales:  loss = scale * sum  (scale = 1 / (rows C) for "mean").
template <typename T>
__global__ void __launch_bounds__(THREADS)
bce_fwd_kernel(const T* __restrict__ logits, int64_t ld, const float* __restrict__ target, int64_t ld_t, int C,
               float* __restrict__ row_loss) {
  __shared__ float sh[THREADS / 32];
  const int64_t r = blockIdx.x;
  const T* x = logits + r * ld;
  const float* t = target + r * ld_t;
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += THREADS) {
    const float v = to_f<T>(x[c]);
    s += fmaxf(v, 0.f) - v * t[c] + log1pf(__expf(-fabsf(v)));
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) row_loss[r] = s;
}
__global__ void __launch_bounds__(THREADS)
sum_rows_kernel(const float* __restrict__ row_loss, int64_t rows, float scale, float* __restrict__ out) {
  __shared__ float sh[THREADS / 32];
  float s = 0.f;
  for (int64_t r = threadIdx.x; r < rows; r += THREADS) s += row_loss[r];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) out[0] = s * scale;
}
// dx = (sigmoid(x) - t) * dloss * scale
template <typename T>
__global__ void __launch_bounds__(THREADS)
bce_bwd_kernel(const T* __restrict__ logits, int64_t ld, const float* __restrict__ target, int64_t ld_t, int C,
               const float* __restrict__ dloss, float scale, T* __restrict__ dlogits, int64_t ld_d) {
  const int64_t r = blockIdx.x;
  const T* x = logits + r * ld;
  const float* t = target + r * ld_t;
  T* d = dlogits + r * ld_d;
  const float g = (*dloss) * scale;
  for (int c = threadIdx.x; c < C; c += THREADS) {
    const float v = to_f<T>(x[c]);
    const float e = __expf(-fabsf(v));
    const float sg = v >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
    d[c] = from_f<T>((sg - t[c]) * g);
  }
}
// Row-wise argmax (the smallest index among equal maxima) and the label at that position: `torch.max(logits, 1)[1]`
// followed by the one-hot scatter * labels of compute_score_with_logits (volta/task_utils.py:429-434), and the
// `preds` / `select_idx` of the VL-logit / V-logit branches (:113-114, :120-121).
template <typename T>
__global__ void __launch_bounds__(THREADS)
argmax_pick_kernel(const T* __restrict__ logits, int64_t ld, const float* __restrict__ labels, int64_t ld_l, int C,
                   int64_t* __restrict__ arg, float* __restrict__ pick) {
  __shared__ float shv[THREADS / 32];
  __shared__ int shi[THREADS / 32];
  const int64_t r = blockIdx.x;
  const T* x = logits + r * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = threadIdx.x; c < C; c += THREADS) {
    const float v = to_f<T>(x[c]);
    if (bi == 0x7fffffff || v > best) { best = v; bi = c; }   // c increases: the first of equal maxima stays
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { shv[threadIdx.x >> 5] = best; shi[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < THREADS / 32; ++w)
      if (shi[w] != 0x7fffffff && (bi == 0x7fffffff || shv[w] > best || (shv[w] == best && shi[w] < bi))) { best = shv[w]; bi = shi[w]; }
    if (bi == 0x7fffffff) bi = 0;
    if (arg) arg[r] = bi;
    if (pick) pick[r] = labels ? labels[r * ld_l + bi] : best;
  }
}

__global__ void finalize_kernel(const float* __restrict__ loss_sum, const float* __restrict__ count, float* __restrict__ out, int n) {
  const int i = threadIdx.x;
  if (i < n) out[i] = loss_sum[i] / fmaxf(count[i], 1.0f);
}

template <typename T> static bool row_vec_ok(const void* p, int64_t ld) {
  return aligned16(p) && ((ld * (int64_t)sizeof(T)) % 32 == 0 || (sizeof(T) == 2 && (ld % 8) == 0));
}

}  // namespace loss
}  // namespace vb

extern "C" {

int vb_xent_fwd(const void* logits, int64_t ld, const int64_t* labels, int64_t rows, int32_t C, float* lse,
                float* loss_sum, float* count, const int32_t* row_limit, int32_t dtype, void* stream) {
  VB_REQUIRE(logits && labels && lse && loss_sum && count && C > 0 && ld >= C, VB_ERR_ARG, "xent_fwd: bad arguments");
  if (rows == 0) return 0;
  using namespace vb::loss;
  if (dtype == VB_BF16)
    xent_fwd_kernel<vb::bf16><<<(unsigned)rows, THREADS, 0, (cudaStream_t)stream>>>(
        (const vb::bf16*)logits, ld, labels, C, lse, loss_sum, count, row_vec_ok<vb::bf16>(logits, ld), row_limit);
  else
    xent_fwd_kernel<float><<<(unsigned)rows, THREADS, 0, (cudaStream_t)stream>>>(
        (const float*)logits, ld, labels, C, lse, loss_sum, count, row_vec_ok<float>(logits, ld), row_limit);
  return vb::check_launch("xent_fwd_kernel");
}

int vb_xent_bwd(const void* logits, int64_t ld, const int64_t* labels, int64_t rows, int32_t C, const float* lse,
                const float* count, const float* dloss, void* dlogits, const int32_t* row_limit, int32_t dtype, void* stream) {
  VB_REQUIRE(logits && labels && lse && count && dloss && dlogits && C > 0 && ld >= C, VB_ERR_ARG, "xent_bwd: bad arguments");
  if (rows == 0) return 0;
  using namespace vb::loss;
  if (dtype == VB_BF16)
    xent_bwd_kernel<vb::bf16><<<(unsigned)rows, THREADS, 0, (cudaStream_t)stream>>>(
        (const vb::bf16*)logits, ld, labels, C, lse, count, dloss, (vb::bf16*)dlogits,
        row_vec_ok<vb::bf16>(logits, ld) && row_vec_ok<vb::bf16>(dlogits, ld), row_limit);
  else
    xent_bwd_kernel<float><<<(unsigned)rows, THREADS, 0, (cudaStream_t)stream>>>(
        (const float*)logits, ld, labels, C, lse, count, dloss, (float*)dlogits,
        row_vec_ok<float>(logits, ld) && row_vec_ok<float>(dlogits, ld), row_limit);
  return vb::check_launch("xent_bwd_kernel");
}

int vb_kl_fwd(const void* logits, int64_t ld, int32_t per, int32_t per_ld, int32_t skip, const int64_t* label,
              const float* target, int64_t rows, int32_t C, float weight, float* lse, float* loss_sum, float* count,
              int32_t dtype, void* stream) {
  VB_REQUIRE(logits && label && target && lse && loss_sum && count && C > 0 && ld >= C && per > 0 && per_ld >= per + skip,
             VB_ERR_ARG, "kl_fwd: bad arguments");
  if (rows == 0) return 0;
  using namespace vb::loss;
  if (dtype == VB_BF16)
    kl_fwd_kernel<vb::bf16><<<(unsigned)rows, THREADS, 0, (cudaStream_t)stream>>>(
        (const vb::bf16*)logits, ld, per, per_ld, skip, label, target, C, weight, lse, loss_sum, count,
        row_vec_ok<vb::bf16>(logits, ld));
  else
    kl_fwd_kernel<float><<<(unsigned)rows, THREADS, 0, (cudaStream_t)stream>>>(
        (const float*)logits, ld, per, per_ld, skip, label, target, C, weight, lse, loss_sum, count,
        row_vec_ok<float>(logits, ld));
  return vb::check_launch("kl_fwd_kernel");
}

int vb_kl_bwd(const void* logits, int64_t ld, int32_t per, int32_t per_ld, int32_t skip, const int64_t* label,
              const float* target, int64_t rows, int32_t C, float weight, const float* lse, const float* count,
              const float* dloss, void* dlogits, int32_t dtype, void* stream) {
  VB_REQUIRE(logits && label && target && lse && count && dloss && dlogits && C > 0 && ld >= C && per > 0 &&
                 per_ld >= per + skip, VB_ERR_ARG, "kl_bwd: bad arguments");
  if (rows == 0) return 0;
  using namespace vb::loss;
  const int64_t logit_rows = (rows / per) * per_ld;
  if (dtype == VB_BF16)
    kl_bwd_kernel<vb::bf16><<<(unsigned)logit_rows, THREADS, 0, (cudaStream_t)stream>>>(
        (const vb::bf16*)logits, ld, per, per_ld, skip, label, target, C, weight, lse, count, dloss, (vb::bf16*)dlogits);
  else
    kl_bwd_kernel<float><<<(unsigned)logit_rows, THREADS, 0, (cudaStream_t)stream>>>(
        (const float*)logits, ld, per, per_ld, skip, label, target, C, weight, lse, count, dloss, (float*)dlogits);
  return vb::check_launch("kl_bwd_kernel");
}

int vb_region_loss_fwd(int32_t kind, const void* logits, int64_t ld, int32_t per, int32_t per_ld, int32_t skip,
                       const int64_t* label, const int64_t* tgt_idx, const float* conf, const float* tgt_vec, int64_t ld_t,
                       int64_t rows, int32_t C, float weight, float* lse, float* loss_sum, float* count, int32_t add_count,
                       int32_t dtype, void* stream) {
  VB_REQUIRE(logits && label && loss_sum && count && C > 0 && ld >= C && per > 0 && per_ld >= per + skip, VB_ERR_ARG,
             "region_loss_fwd: bad arguments");
  VB_REQUIRE(kind == VB_REGION_XENT ? (tgt_idx && lse) : ((kind == VB_REGION_MSE || kind == VB_REGION_HUBER) && tgt_vec && ld_t >= C),
             VB_ERR_ARG, "region_loss_fwd: kind %d lacks its target", kind);
  if (rows == 0) return 0;
  using namespace vb::loss;
  if (dtype == VB_BF16)
    region_fwd_kernel<vb::bf16><<<(unsigned)rows, THREADS, 0, (cudaStream_t)stream>>>(
        kind, (const vb::bf16*)logits, ld, per, per_ld, skip, label, tgt_idx, conf, tgt_vec, ld_t, C, weight, lse, loss_sum, count,
        add_count, row_vec_ok<vb::bf16>(logits, ld));
  else
    region_fwd_kernel<float><<<(unsigned)rows, THREADS, 0, (cudaStream_t)stream>>>(
        kind, (const float*)logits, ld, per, per_ld, skip, label, tgt_idx, conf, tgt_vec, ld_t, C, weight, lse, loss_sum, count,
        add_count, row_vec_ok<float>(logits, ld));
  return vb::check_launch("region_fwd_kernel");
}

int vb_region_loss_bwd(int32_t kind, const void* logits, int64_t ld, int32_t per, int32_t per_ld, int32_t skip,
                       const int64_t* label, const int64_t* tgt_idx, const float* conf, const float* tgt_vec, int64_t ld_t,
                       int64_t rows, int32_t C, float weight, const float* lse, const float* count, const float* dloss,
                       void* dlogits, int32_t dtype, void* stream) {
  VB_REQUIRE(logits && label && count && dloss && dlogits && C > 0 && ld >= C && per > 0 && per_ld >= per + skip, VB_ERR_ARG,
             "region_loss_bwd: bad arguments");
  VB_REQUIRE(kind == VB_REGION_XENT ? (tgt_idx && lse) : ((kind == VB_REGION_MSE || kind == VB_REGION_HUBER) && tgt_vec),
             VB_ERR_ARG, "region_loss_bwd: kind %d lacks its target", kind);
  if (rows == 0) return 0;
  using namespace vb::loss;
  const int64_t logit_rows = (rows / per) * per_ld;
  if (dtype == VB_BF16)
    region_bwd_kernel<vb::bf16><<<(unsigned)logit_rows, THREADS, 0, (cudaStream_t)stream>>>(
        kind, (const vb::bf16*)logits, ld, per, per_ld, skip, label, tgt_idx, conf, tgt_vec, ld_t, C, weight, lse, count, dloss,
        (vb::bf16*)dlogits);
  else
    region_bwd_kernel<float><<<(unsigned)logit_rows, THREADS, 0, (cudaStream_t)stream>>>(
        kind, (const float*)logits, ld, per, per_ld, skip, label, tgt_idx, conf, tgt_vec, ld_t, C, weight, lse, count, dloss,
        (float*)dlogits);
  return vb::check_launch("region_bwd_kernel");
}

int vb_bce_logits_fwd(const void* logits, int64_t ld, const float* target, int64_t ld_t, int64_t rows, int32_t C,
                      float* row_loss, float scale, float* loss_out, int32_t dtype, void* stream) {
  VB_REQUIRE(logits && target && row_loss && loss_out && rows > 0 && C > 0, VB_ERR_ARG, "bce_logits_fwd: bad arguments");
  VB_REQUIRE(ld >= C && ld_t >= C, VB_ERR_ARG, "bce_logits_fwd: leading dimension smaller than C");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VB_BF16)
    vb::loss::bce_fwd_kernel<vb::bf16><<<(unsigned)rows, vb::loss::THREADS, 0, st>>>((const vb::bf16*)logits, ld, target, ld_t, C, row_loss);
  else if (dtype == VB_F32)
    vb::loss::bce_fwd_kernel<float><<<(unsigned)rows, vb::loss::THREADS, 0, st>>>((const float*)logits, ld, target, ld_t, C, row_loss);
  else { vb::set_error("bce_logits_fwd: bad dtype %d", dtype); return VB_ERR_ARG; }
  vb::loss::sum_rows_kernel<<<1, vb::loss::THREADS, 0, st>>>(row_loss, rows, scale, loss_out);
  return vb::check_launch("bce_fwd_kernel");
}

int vb_bce_logits_bwd(const void* logits, int64_t ld, const float* target, int64_t ld_t, int64_t rows, int32_t C,
                      const float* dloss, float scale, void* dlogits, int64_t ld_d, int32_t dtype, void* stream) {
  VB_REQUIRE(logits && target && dloss && dlogits && rows > 0 && C > 0, VB_ERR_ARG, "bce_logits_bwd: bad arguments");
  VB_REQUIRE(ld >= C && ld_t >= C && ld_d >= C, VB_ERR_ARG, "bce_logits_bwd: leading dimension smaller than C");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VB_BF16)
    vb::loss::bce_bwd_kernel<vb::bf16><<<(unsigned)rows, vb::loss::THREADS, 0, st>>>((const vb::bf16*)logits, ld, target, ld_t, C, dloss, scale, (vb::bf16*)dlogits, ld_d);
  else if (dtype == VB_F32)
    vb::loss::bce_bwd_kernel<float><<<(unsigned)rows, vb::loss::THREADS, 0, st>>>((const float*)logits, ld, target, ld_t, C, dloss, scale, (float*)dlogits, ld_d);
  else { vb::set_error("bce_logits_bwd: bad dtype %d", dtype); return VB_ERR_ARG; }
  return vb::check_launch("bce_bwd_kernel");
}

int vb_argmax_pick(const void* logits, int64_t ld, const float* labels, int64_t ld_l, int64_t rows, int32_t C, int64_t* arg,
                   float* pick, int32_t dtype, void* stream) {
  VB_REQUIRE(logits && (arg || pick) && rows > 0 && C > 0, VB_ERR_ARG, "argmax_pick: bad arguments");
  VB_REQUIRE(ld >= C && (!labels || ld_l >= C), VB_ERR_ARG, "argmax_pick: leading dimension smaller than C");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VB_BF16)
    vb::loss::argmax_pick_kernel<vb::bf16><<<(unsigned)rows, vb::loss::THREADS, 0, st>>>((const vb::bf16*)logits, ld, labels, ld_l, C, arg, pick);
  else if (dtype == VB_F32)
    vb::loss::argmax_pick_kernel<float><<<(unsigned)rows, vb::loss::THREADS, 0, st>>>((const float*)logits, ld, labels, ld_l, C, arg, pick);
  else { vb::set_error("argmax_pick: bad dtype %d", dtype); return VB_ERR_ARG; }
  return vb::check_launch("argmax_pick_kernel");
}

int vb_loss_finalize(const float* loss_sum, const float* count, float* out, int32_t n, void* stream) {
  VB_REQUIRE(loss_sum && count && out && n > 0 && n <= 32, VB_ERR_ARG, "loss_finalize: bad arguments");
  vb::loss::finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(loss_sum, count, out, n);
  return vb::check_launch("finalize_kernel");
}

}  // extern "C"
