// Shared device/host helpers for the volta_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/volta_b200.h"

namespace vb {

// ---- error plumbing ------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define VB_REQUIRE(cond, code, ...)      \
  do {                                   \
    if (!(cond)) {                       \
      vb::set_error(__VA_ARGS__);        \
      return (code);                     \
    }                                    \
  } while (0)

// ---- dtype helpers -------------------------------------------------------
using bf16 = __nv_bfloat16;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 p = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(p);
}

// 8 consecutive elements <-> 8 floats (16 B for bf16, 32 B for fp32).  Pointers must be
// 16-byte aligned for the vector path; callers check and fall back to `load8_scalar`.
template <typename T> struct Vec8;
template <> struct Vec8<bf16> {
  typedef uint4 Raw;   // 8 packed elements: keep a prefetched vector in 4 registers until it is needed
  static __device__ __forceinline__ Raw load_raw(const bf16* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[8]) {
    float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y), c = unpack_bf16x2(r.z), d = unpack_bf16x2(r.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) { unpack(load_raw(p), v); }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) {
    uint4 r;
    r.x = pack_bf16x2(v[0], v[1]); r.y = pack_bf16x2(v[2], v[3]);
    r.z = pack_bf16x2(v[4], v[5]); r.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = r;
  }
};
template <> struct Vec8<float> {
  struct Raw { float4 a, b; };
  static __device__ __forceinline__ Raw load_raw(const float* p) {
    Raw r;
    r.a = *reinterpret_cast<const float4*>(p); r.b = *reinterpret_cast<const float4*>(p + 4);
    return r;
  }
  static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[8]) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
  }
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) { unpack(load_raw(p), v); }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};

// ---- math ------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) {            // volta/encoders.py:136
  return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float dgelu_erf(float x) {           // d/dx of the above
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Fast erf for the tensor-core epilogues (Abramowitz & Stegun 7.1.26, |abs err| <= 1.5e-7 - far below bf16
// resolution): one MUFU.RCP + one MUFU.EX2 + ~10 FMAs instead of erff()'s ~40-instruction polynomial, so the
// GELU epilogue stays hidden under the MMA of the next tile.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = ex2_approx(-1.4426950408889634f * ax * ax);
  const float r = 1.0f - poly * t * e;
  return copysignf(r, x);
}
// GELU and its derivative from shared intermediates (exp(-x^2/2) serves both erf and the Gaussian pdf):
// 2 MUFU + ~16 FMA/ALU per element.
__device__ __forceinline__ void gelu_fwd_grad_fast(float x, float& y, float& dy) {
  const float az = fabsf(x) * 0.70710678118654752440f;
  const float t = rcp_approx(fmaf(0.3275911f, az, 1.0f));
  const float e = ex2_approx(-1.4426950408889634f * az * az);  // exp(-x^2/2)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erfv = copysignf(1.0f - poly * t * e, x);
  const float cdf = fmaf(0.5f, erfv, 0.5f);
  y = x * cdf;
  dy = fmaf(x, 0.39894228040143267794f * e, cdf);
}
__device__ __forceinline__ float gelu_fast(float x) { return x * 0.5f * (1.0f + erf_fast(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float dgelu_fast(float x) {
  const float cdf = 0.5f * (1.0f + erf_fast(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * ex2_approx(-0.72134752044448170368f * x * x);
  return fmaf(x, pdf, cdf);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- counter-based dropout RNG (masks are regenerated in the backward, never stored) -----------------
// One 32-bit hash (lowbias32 finaliser over a seeded multiply-mix of the 64-bit pair index) yields the keep
// decisions of TWO consecutive elements, 16 bits each: ~5 integer instructions per element.  An earlier
// version used Philox4x32-10 (~18 instructions per element): inside the GEMM epilogue and the LayerNorm
// backward that made the dropout mask, not the tensor core or HBM, the bottleneck.
__device__ __forceinline__ uint32_t rng_pair(uint64_t seed, uint64_t pair) {
  const uint32_t lo = (uint32_t)pair, hi = (uint32_t)(pair >> 32);
  uint32_t h = (lo * 0x9E3779B1u) ^ (uint32_t)seed;
  h ^= (hi + (uint32_t)(seed >> 32)) * 0x85EBCA77u;
  h ^= h >> 16; h *= 0x7FEB352Du;
  h ^= h >> 15; h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}
// threshold on 16 bits: drop when r16 < p * 65536
__host__ __device__ __forceinline__ uint32_t dropout_thresh(float p) {
  const double t = (double)p * 65536.0 + 0.5;
  return t >= 65535.0 ? 65535u : (uint32_t)t;
}
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t offset, uint64_t idx, uint32_t thresh) {
  idx += offset;
  const uint32_t h = rng_pair(seed, idx >> 1);
  const uint32_t r16 = (idx & 1) ? (h >> 16) : (h & 0xFFFFu);
  return r16 >= thresh;
}
// keep flags for 2 consecutive elements starting at an EVEN index: one hash
__device__ __forceinline__ void dropout_keep2(uint64_t seed, uint64_t offset, uint64_t idx, uint32_t thresh, bool& k0, bool& k1) {
  const uint32_t h = rng_pair(seed, (idx + offset) >> 1);
  k0 = (h & 0xFFFFu) >= thresh;
  k1 = (h >> 16) >= thresh;
}

// Device-resident seed base (vb_set_seed_pointer): when set, every dropout-bearing launch adds *ptr to its seed at run
// time, so a captured CUDA graph draws fresh masks on every replay.  NULL = seeds are plain by-value arguments.
const uint64_t* seed_pointer();
int gemm_sm_limit();  // vb_set_gemm_sm_limit: 0 = use every SM
__device__ __forceinline__ uint64_t effective_seed(uint64_t seed, const uint64_t* ptr) { return ptr ? seed + __ldg(ptr) : seed; }

// Programmatic dependent launch hooks for the short kernels between the GEMMs.  Host side: launch_pdl() can set the
// programmatic-serialization attribute; device side, the FIRST statement of the kernel waits until the kernel it
// depends on has completed and flushed (a no-op without the attribute).  The persistent GEMM releases its dependents
// at its start (it owns whole SMs, so a dependent block only lands where a GEMM block has exited).
// Compiled out by default (-DVB_PDL_SMALL_KERNELS enables it together with VOLTA_B200_PDL_SMALL=1): even without the
// launch attribute the griddepcontrol.wait at the top of every block cost the 1152-block LayerNorm forward 4 us
// (11.8 -> 15.6 us, measured).
__device__ __forceinline__ void pdl_enter() {
#ifdef VB_PDL_SMALL_KERNELS
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  // Off by default: A/B on a B200 (same box, 30-step runs) gave 19.27 / 19.38 ms without vs 19.51 / 19.49 ms with the
  // attribute on the short kernels - a waiting block still holds its slot.  VOLTA_B200_PDL_SMALL=1 re-enables it.
  static const int on = [] { const char* e = getenv("VOLTA_B200_PDL_SMALL"); return (e && e[0] == '1') ? 1 : 0; }();
  cfg.numAttrs = on;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace vb
