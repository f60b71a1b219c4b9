// Optimizer step over the flat parameter / gradient buffers: AdamW with the semantics of
// pytorch_transformers.AdamW as the reference calls it (train_concap.py:199-227: per-parameter groups with their own lr
// and weight decay, correct_bias=True) and the global-norm gradient clip of train_concap.py:305-307
// (torch.nn.utils.clip_grad_norm_).  pytorch_transformers 1.1.0 is not in the image (requirements.txt:39); the update
// rule is restated from its published algorithm:
//     m = b1 m + (1-b1) g ;  v = b2 v + (1-b2) g^2 ;  p -= lr sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps) ;  p -= lr wd p
// One pass: reads p, g, m, v (fp32), writes p, m, v and the bf16 tensor-core copy of p - the separate fp32->bf16 refresh
// pass and ~500 per-tensor optimizer launches of a torch optimizer do not exist.  HBM-bound: 30 B per parameter.
#include "common.cuh"

namespace vb {
namespace optim {

constexpr int CHUNK = 64;  // elements; parameters start on 64-element boundaries of the flat buffer (store.py ALIGN)

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  __shared__ float sh[8];
  float s = 0.f;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = n4 << 2; i < n; ++i) s += x[i] * x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(out, t);
  }
}

// group_hp[g] = {lr, weight_decay, step_size, active}; chunk_group[c] = group of elements [64c, 64c+64) or 0xFFFF (not optimised)
// RULE 1 = RAdam as the reference ships it (volta/optimization.py:9-93, train_task.py:227-228): the moments as above, then
//     p -= (wd lr) p ;   N_sma >= 5:  p -= step_size m / (sqrt(v) + eps)     else:  p -= step_size m
// with group_hp[g] = {wd * lr, rectified (N_sma >= 5) ? 1 : 0, step_size, active}; the host forms N_sma and step_size.
template <int RULE>
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             bf16* __restrict__ p16, int64_t n, const uint16_t* __restrict__ chunk_group, const float4* __restrict__ group_hp,
             float beta1, float beta2, float omb1, float omb2, float eps, const float* __restrict__ sumsq, float max_norm) {
  float coef = 1.f;
  if (sumsq != nullptr && max_norm > 0.f) coef = fminf(1.f, max_norm / (sqrtf(*sumsq) + 1e-6f));  // clip_grad_norm_
  const int64_t n8 = n >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i << 3;
    const uint16_t grp = chunk_group[e / CHUNK];
    if (grp == 0xFFFFu) continue;
    const float4 hp = group_hp[grp];
    if (hp.w == 0.f) continue;
    float pv[8], gv[8], mv[8], vv[8];
    Vec8<float>::load(p + e, pv);
    Vec8<float>::load(g + e, gv);
    Vec8<float>::load(m + e, mv);
    Vec8<float>::load(v + e, vv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gg = gv[j] * coef;
      mv[j] = beta1 * mv[j] + omb1 * gg;          // omb = 1 - beta formed in double on the host, as the reference's
      vv[j] = beta2 * vv[j] + omb2 * gg * gg;     // Python scalars are (1.f - 0.999f is off by 5e-5 relative)
      float x = pv[j];
      if constexpr (RULE == 0) {
        x -= hp.z * mv[j] / (sqrtf(vv[j]) + eps);
        if (hp.y > 0.f) x -= hp.x * hp.y * x;
      } else {
        if (hp.x != 0.f) x -= hp.x * x;
        x -= hp.y != 0.f ? hp.z * mv[j] / (sqrtf(vv[j]) + eps) : hp.z * mv[j];
      }
      pv[j] = x;
    }
    Vec8<float>::store(p + e, pv);
    Vec8<float>::store(m + e, mv);
    Vec8<float>::store(v + e, vv);
    if (p16 != nullptr) Vec8<bf16>::store(p16 + e, pv);
  }
}

}  // namespace optim
}  // namespace vb

extern "C" {

int vb_sumsq(const float* x, int64_t n, float* out, void* stream) {
  VB_REQUIRE(x && out && n >= 0 && vb::aligned16(x), VB_ERR_ARG, "sumsq: bad arguments");
  if (n == 0) return 0;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(vb::ceil_div(n / 4, 256), 148 * 8));
  vb::optim::sumsq_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, out);
  return vb::check_launch("sumsq_kernel");
}

static int adam_flat(int rule, float* p, const float* g, float* m, float* v, void* p16, int64_t n, const uint16_t* chunk_group,
                     const float* group_hp, double beta1, double beta2, float eps, const float* sumsq, float max_norm, void* stream) {
  VB_REQUIRE(p && g && m && v && chunk_group && group_hp, VB_ERR_ARG, "adamw_flat: null pointer");
  VB_REQUIRE(n % vb::optim::CHUNK == 0, VB_ERR_ARG, "adamw_flat: n must be a multiple of 64 (flat store alignment)");
  VB_REQUIRE(vb::aligned16(p) && vb::aligned16(g) && vb::aligned16(m) && vb::aligned16(v) && vb::aligned16(group_hp) &&
                 (!p16 || vb::aligned16(p16)), VB_ERR_ARG, "adamw_flat: buffers must be 16-byte aligned");
  if (n == 0) return 0;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(vb::ceil_div(n / 8, 256), 148 * 16));
  if (rule == 0)
    vb::optim::adamw_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (vb::bf16*)p16, n, chunk_group,
                                                                        reinterpret_cast<const float4*>(group_hp), (float)beta1,
                                                                        (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), eps,
                                                                        sumsq, max_norm);
  else
    vb::optim::adamw_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (vb::bf16*)p16, n, chunk_group,
                                                                        reinterpret_cast<const float4*>(group_hp), (float)beta1,
                                                                        (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), eps,
                                                                        sumsq, max_norm);
  return vb::check_launch("adamw_kernel");
}

int vb_adamw_flat(float* p, const float* g, float* m, float* v, void* p16, int64_t n, const uint16_t* chunk_group,
                  const float* group_hp, double beta1, double beta2, float eps, const float* sumsq, float max_norm, void* stream) {
  return adam_flat(0, p, g, m, v, p16, n, chunk_group, group_hp, beta1, beta2, eps, sumsq, max_norm, stream);
}

int vb_radam_flat(float* p, const float* g, float* m, float* v, void* p16, int64_t n, const uint16_t* chunk_group,
                  const float* group_hp, double beta1, double beta2, float eps, const float* sumsq, float max_norm, void* stream) {
  return adam_flat(1, p, g, m, v, p16, n, chunk_group, group_hp, beta1, beta2, eps, sumsq, max_norm, stream);
}

}  // extern "C"
