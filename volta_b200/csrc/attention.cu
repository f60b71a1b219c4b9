// Gated attention (volta/encoders.py:258-340) after the Q/K/V projections: per (sample, head), the
// text and vision rows form one sequence [text; vision]; the config's tt/tv/vt/vv flags enable the four
// (query-stream, key-stream) blocks and the softmax of a row runs JOINTLY over all its enabled key blocks.
// Sequences are short (<= 56 text + <= 101 regions), so a whole (sample, head) lives in one CTA's shared
// memory: no probability tensor ever reaches HBM; the backward recomputes P from the saved row lse.
//
// Two back-ends behind one C entry point:
//  * MMA  (bf16, d in {64,128}): warp-level mma.sync.m16n8k16 with ldmatrix operands, fp32 softmax.
//    (tcgen05 wants 128-row tiles; Lq = 36 would fill 28 % of one - see DESIGN.md.)
//  * SIMT (fp32 validation mode / any d): one warp per query (or key) row.
#include <algorithm>

#include "common.cuh"
#include "gemm_tc_kernel.cuh"   // PTX wrappers (mbarrier, TMA, tcgen05) shared with the GEMM

namespace vb {
namespace attn {

struct P {
  const void* qkv[2];
  int64_t ld[2];
  const int32_t* mask[2];
  void* ctx[2];
  int64_t ld_ctx[2];
  float* lse[2];
  const void* dctx[2];
  void* dqkv[2];
  int B, L[2], heads, d, A;  // A = heads*d
  int pair_mask;
  float scale;
  uint32_t thresh[2];        // per query stream
  float keep_scale[2];
  uint64_t seed;             // seed + offset
  const uint64_t* seed_ptr;  // see vb_set_seed_pointer
  int ngroups;               // MMA back-end: independent (query set, key set) groups per (sample, head)
  const int32_t* len[2];     // optional [B] per stream: 1 + index of the last position whose mask is set (vb_mask_lengths)
  const int32_t* cu[2];      // optional [B] per stream: PACKED rows - sample b owns rows [cu[b], cu[b] + len[b]) of the buffers
};

// Effective length of stream s in sample b: positions at or beyond it are padding for every head and every layer - as
// keys they have softmax weight exactly 0 (additive -10000, encoders.py:983-991; exp underflows to 0 in fp32 and bf16),
// as queries nothing downstream reads them (SURVEY.md 8a' item 3) - so whole 16-row tiles beyond it are neither loaded
// nor multiplied; their outputs are written as zeros.  Masked positions BELOW the length (arbitrary masks) take the
// ordinary path.  A sample with no valid position at all (len 0) is computed in full, like the reference does.
__device__ __forceinline__ int eff_len(const P& p, int s, int b) {
  int L = p.L[s];
  if (p.len[s] != nullptr) {
    const int v = p.len[s][b];
    if (p.cu[s] != nullptr) return min(max(v, 0), L);   // packed buffers hold exactly len[b] rows of this sample
    if (v > 0 && v < L) L = v;
  }
  return L;
}

__host__ __device__ __forceinline__ bool pair_on(int pair_mask, int qs, int ks) { return (pair_mask >> (qs * 2 + ks)) & 1; }

// Dropout on the attention probabilities (volta/encoders.py:294-299, 309-314).  The keep decision of the element
// (sample b, head h, global query row qg | key stream ks, key j) is a pure function of
//   row_id = (b*heads + h)*Ltot + qg   and   (ks, j >> 1):
// one 32-bit hash serves the two keys 2m, 2m+1 of a stream (16 bits each), and it splits into a per-row term and a
// per-column term, so in the mma.sync fragment layout (2 rows x 2 adjacent keys per thread and n-tile) the cost
// is ~5 integer instructions per probability.  Forward, backward and the SIMT validation back-end all use it.
__device__ __forceinline__ uint32_t attn_rowterm(uint64_t seed, uint32_t row_id) { return (row_id * 0x9E3779B1u) ^ (uint32_t)seed; }
__device__ __forceinline__ uint32_t attn_colterm(uint64_t seed, int ks, int j) {
  return ((((uint32_t)ks << 16) | ((uint32_t)j >> 1)) + (uint32_t)(seed >> 32)) * 0x85EBCA77u;
}
__device__ __forceinline__ uint32_t attn_mix(uint32_t h) {
  h ^= h >> 16; h *= 0x7FEB352Du;
  h ^= h >> 15; h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}
__device__ __forceinline__ bool attn_keep(uint64_t seed, uint32_t row_id, int ks, int j, uint32_t thresh) {
  const uint32_t h = attn_mix(attn_rowterm(seed, row_id) ^ attn_colterm(seed, ks, j));
  return ((j & 1) ? (h >> 16) : (h & 0xFFFFu)) >= thresh;
}

// =====================================================================================
// SIMT back-end
// =====================================================================================
template <typename T>
__global__ void __launch_bounds__(128)
simt_fwd_kernel(P p) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qs = blockIdx.y;
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  if (p.qkv[qs] == nullptr || p.ctx[qs] == nullptr) return;
  const int Ltot = p.L[0] + p.L[1];
  float* qbuf = sm + warp * (p.d + Ltot);
  float* pbuf = qbuf + p.d;
  const T* Q = reinterpret_cast<const T*>(p.qkv[qs]);
  for (int i = warp; i < p.L[qs]; i += 4) {
    const int64_t qrow = (int64_t)b * p.L[qs] + i;
    for (int dd = lane; dd < p.d; dd += 32) qbuf[dd] = to_f<T>(Q[qrow * p.ld[qs] + h * p.d + dd]);
    __syncwarp();
    float mx = -INFINITY;
    for (int kg = lane; kg < Ltot; kg += 32) {
      const int ks = kg < p.L[0] ? 0 : 1;
      const int j = ks ? kg - p.L[0] : kg;
      float s = -INFINITY;
      if (pair_on(p.pair_mask, qs, ks) && p.qkv[ks] != nullptr) {
        const T* K = reinterpret_cast<const T*>(p.qkv[ks]) + ((int64_t)b * p.L[ks] + j) * p.ld[ks] + p.A + h * p.d;
        float acc = 0.f;
        for (int dd = 0; dd < p.d; ++dd) acc = fmaf(qbuf[dd], to_f<T>(K[dd]), acc);
        s = acc * p.scale + (p.mask[ks][(int64_t)b * p.L[ks] + j] ? 0.f : -10000.0f);
      }
      pbuf[kg] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    if (mx == -INFINITY) mx = 0.f;
    float sum = 0.f;
    for (int kg = lane; kg < Ltot; kg += 32) sum += __expf(pbuf[kg] - mx);
    sum = warp_sum(sum);
    const float lse = mx + __logf(sum);
    if (lane == 0) p.lse[qs][((int64_t)b * p.heads + h) * p.L[qs] + i] = lse;
    const int qg = qs ? p.L[0] + i : i;
    const uint32_t row_id = ((uint32_t)b * p.heads + h) * Ltot + qg;
    for (int kg = lane; kg < Ltot; kg += 32) {
      float pr = __expf(pbuf[kg] - lse);
      if (p.thresh[qs]) {
        const int ks = kg < p.L[0] ? 0 : 1;
        pr = attn_keep(effective_seed(p.seed, p.seed_ptr), row_id, ks, ks ? kg - p.L[0] : kg, p.thresh[qs]) ? pr * p.keep_scale[qs] : 0.f;
      }
      pbuf[kg] = pr;
    }
    __syncwarp();
    T* O = reinterpret_cast<T*>(p.ctx[qs]) + qrow * p.ld_ctx[qs] + h * p.d;
    for (int dd = lane; dd < p.d; dd += 32) {
      float acc = 0.f;
      for (int kg = 0; kg < Ltot; ++kg) {
        const float pr = pbuf[kg];
        if (pr != 0.f) {
          const int ks = kg < p.L[0] ? 0 : 1;
          const int j = ks ? kg - p.L[0] : kg;
          acc = fmaf(pr, to_f<T>(reinterpret_cast<const T*>(p.qkv[ks])[((int64_t)b * p.L[ks] + j) * p.ld[ks] + 2 * p.A + h * p.d + dd]), acc);
        }
      }
      O[dd] = from_f<T>(acc);
    }
    __syncwarp();
  }
}

// dQ: one warp per query row.
template <typename T>
__global__ void __launch_bounds__(128)
simt_bwd_dq_kernel(P p) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qs = blockIdx.y;
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  if (p.qkv[qs] == nullptr || p.dqkv[qs] == nullptr) return;
  const int Ltot = p.L[0] + p.L[1];
  float* qbuf = sm + warp * (2 * p.d + Ltot);
  float* dobuf = qbuf + p.d;
  float* pbuf = dobuf + p.d;
  if (p.dctx[qs] == nullptr || p.ctx[qs] == nullptr) {  // stream is keys-only: its queries get no gradient
    for (int i = warp; i < p.L[qs]; i += 4)
      for (int dd = lane; dd < p.d; dd += 32)
        reinterpret_cast<T*>(p.dqkv[qs])[((int64_t)b * p.L[qs] + i) * p.ld[qs] + h * p.d + dd] = from_f<T>(0.f);
    return;
  }
  for (int i = warp; i < p.L[qs]; i += 4) {
    const int64_t qrow = (int64_t)b * p.L[qs] + i;
    float dl = 0.f;
    for (int dd = lane; dd < p.d; dd += 32) {
      qbuf[dd] = to_f<T>(reinterpret_cast<const T*>(p.qkv[qs])[qrow * p.ld[qs] + h * p.d + dd]);
      const float dov = to_f<T>(reinterpret_cast<const T*>(p.dctx[qs])[qrow * p.ld_ctx[qs] + h * p.d + dd]);
      dobuf[dd] = dov;
      dl += dov * to_f<T>(reinterpret_cast<const T*>(p.ctx[qs])[qrow * p.ld_ctx[qs] + h * p.d + dd]);
    }
    const float delta = warp_sum(dl);
    __syncwarp();
    const float lse = p.lse[qs][((int64_t)b * p.heads + h) * p.L[qs] + i];
    const int qg = qs ? p.L[0] + i : i;
    const uint32_t row_id = ((uint32_t)b * p.heads + h) * Ltot + qg;
    for (int kg = lane; kg < Ltot; kg += 32) {
      const int ks = kg < p.L[0] ? 0 : 1;
      const int j = ks ? kg - p.L[0] : kg;
      float ds = 0.f;
      if (pair_on(p.pair_mask, qs, ks) && p.qkv[ks] != nullptr) {
        const T* K = reinterpret_cast<const T*>(p.qkv[ks]) + ((int64_t)b * p.L[ks] + j) * p.ld[ks] + p.A + h * p.d;
        const T* V = K + p.A;
        float s = 0.f, dp = 0.f;
        for (int dd = 0; dd < p.d; ++dd) {
          s = fmaf(qbuf[dd], to_f<T>(K[dd]), s);
          dp = fmaf(dobuf[dd], to_f<T>(V[dd]), dp);
        }
        s = s * p.scale + (p.mask[ks][(int64_t)b * p.L[ks] + j] ? 0.f : -10000.0f);
        const float pr = __expf(s - lse);
        if (p.thresh[qs]) dp = attn_keep(effective_seed(p.seed, p.seed_ptr), row_id, ks, j, p.thresh[qs]) ? dp * p.keep_scale[qs] : 0.f;
        ds = pr * (dp - delta) * p.scale;
      }
      pbuf[kg] = ds;
    }
    __syncwarp();
    T* dQ = reinterpret_cast<T*>(p.dqkv[qs]) + qrow * p.ld[qs] + h * p.d;
    for (int dd = lane; dd < p.d; dd += 32) {
      float acc = 0.f;
      for (int kg = 0; kg < Ltot; ++kg) {
        const float ds = pbuf[kg];
        if (ds != 0.f) {
          const int ks = kg < p.L[0] ? 0 : 1;
          const int j = ks ? kg - p.L[0] : kg;
          acc = fmaf(ds, to_f<T>(reinterpret_cast<const T*>(p.qkv[ks])[((int64_t)b * p.L[ks] + j) * p.ld[ks] + p.A + h * p.d + dd]), acc);
        }
      }
      dQ[dd] = from_f<T>(acc);
    }
    __syncwarp();
  }
}

// dK, dV: one warp per key row, looping over every query row that attends it.
template <typename T>
__global__ void __launch_bounds__(128)
simt_bwd_dkv_kernel(P p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ks = blockIdx.y;
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  if (p.qkv[ks] == nullptr || p.dqkv[ks] == nullptr) return;
  const int Ltot = p.L[0] + p.L[1];
  constexpr int MAXR = 8;  // d <= 256
  for (int j = warp; j < p.L[ks]; j += 4) {
    const int64_t krow = (int64_t)b * p.L[ks] + j;
    const T* K = reinterpret_cast<const T*>(p.qkv[ks]) + krow * p.ld[ks] + p.A + h * p.d;
    const T* V = K + p.A;
    float kreg[MAXR], vreg[MAXR], dk[MAXR], dv[MAXR];
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
      const int dd = lane + 32 * r;
      kreg[r] = dd < p.d ? to_f<T>(K[dd]) : 0.f;
      vreg[r] = dd < p.d ? to_f<T>(V[dd]) : 0.f;
      dk[r] = 0.f; dv[r] = 0.f;
    }
    const float madd = p.mask[ks][krow] ? 0.f : -10000.0f;
    for (int qs = 0; qs < 2; ++qs) {
      if (!pair_on(p.pair_mask, qs, ks) || p.qkv[qs] == nullptr) continue;
      for (int i = 0; i < p.L[qs]; ++i) {
        const int64_t qrow = (int64_t)b * p.L[qs] + i;
        const T* Q = reinterpret_cast<const T*>(p.qkv[qs]) + qrow * p.ld[qs] + h * p.d;
        const T* dO = reinterpret_cast<const T*>(p.dctx[qs]) + qrow * p.ld_ctx[qs] + h * p.d;
        const T* O = reinterpret_cast<const T*>(p.ctx[qs]) + qrow * p.ld_ctx[qs] + h * p.d;
        float qreg[MAXR], doreg[MAXR];
        float s = 0.f, dp = 0.f, dl = 0.f;
#pragma unroll
        for (int r = 0; r < MAXR; ++r) {
          const int dd = lane + 32 * r;
          qreg[r] = dd < p.d ? to_f<T>(Q[dd]) : 0.f;
          doreg[r] = dd < p.d ? to_f<T>(dO[dd]) : 0.f;
          const float ov = dd < p.d ? to_f<T>(O[dd]) : 0.f;
          s = fmaf(qreg[r], kreg[r], s);
          dp = fmaf(doreg[r], vreg[r], dp);
          dl = fmaf(doreg[r], ov, dl);
        }
        s = warp_sum(s); dp = warp_sum(dp); dl = warp_sum(dl);
        const float lse = p.lse[qs][((int64_t)b * p.heads + h) * p.L[qs] + i];
        float pr = __expf(s * p.scale + madd - lse);
        float pd = pr;
        if (p.thresh[qs]) {
          const int qg = qs ? p.L[0] + i : i;
          const bool keep = attn_keep(effective_seed(p.seed, p.seed_ptr), ((uint32_t)b * p.heads + h) * Ltot + qg, ks, j, p.thresh[qs]);
          pd = keep ? pr * p.keep_scale[qs] : 0.f;
          dp = keep ? dp * p.keep_scale[qs] : 0.f;
        }
        const float ds = pr * (dp - dl) * p.scale;
#pragma unroll
        for (int r = 0; r < MAXR; ++r) {
          dk[r] = fmaf(ds, qreg[r], dk[r]);
          dv[r] = fmaf(pd, doreg[r], dv[r]);
        }
      }
    }
    T* dK = reinterpret_cast<T*>(p.dqkv[ks]) + krow * p.ld[ks] + p.A + h * p.d;
    T* dV = dK + p.A;
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
      const int dd = lane + 32 * r;
      if (dd < p.d) { dK[dd] = from_f<T>(dk[r]); dV[dd] = from_f<T>(dv[r]); }
    }
  }
}

// =====================================================================================
// MMA back-end (bf16)
// =====================================================================================
// Work decomposition.  The (query-stream, key-stream) blocks a sub-layer enables split a (sample, head) into one or
// two independent GROUPS: tt+vv (pair_mask 9) and tv+vt (6) give two groups with one query stream and one key
// stream each; every other mask is one group holding all its streams.  One CTA = one (sample, head, group):
//   forward : one warp per 16-query tile - S = Q K^T, joint softmax over the enabled key n-tiles (8 keys each,
//             so 36 keys cost 5 n-tiles, not 6), dropout, O = P V; O leaves through shared memory in 16-byte rows.
//   backward: one warp per 16-KEY tile - S^T = K Q^T and dP^T = V dO^T per 16-query block, P and dS recomputed
//             ONCE; dV += P^T dO and dK += dS^T Q accumulate in registers (every element written exactly once);
//             the warp also leaves its dS^T tile (bf16) in shared memory, and after one block barrier a second
//             phase (one warp per 16-QUERY tile) forms dQ = dS K with dS read back through ldmatrix.trans - no
//             atomics (red.shared.add.f32 is a CAS spin loop in SASS), no second softmax/dropout recomputation.
// An earlier version kept a fixed 6/10-block key loop per query tile and recomputed S, the softmax and the dropout
// hash in two separate backward phases: ncu showed both kernels issue-bound (25 M / 84 M warp instructions per
// launch at B=256 for ~100 MB of HBM traffic).
__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

template <int D> struct Tile {
  static constexpr int PITCH = D + 8;  // bf16 elements; 16-byte padding keeps ldmatrix conflict-free
  // A operand: 16x16 block of a row-major [row][D] tile at (row0, col0)
  static __device__ __forceinline__ void load_a(const bf16* base, int row0, int col0, uint32_t (&r)[4]) {
    const int lane = threadIdx.x & 31;
    ldsm_x4(smem_u32(base + (row0 + (lane & 15)) * PITCH + col0 + (lane >> 4) * 8), r);
  }
  // B operand with B[k][n] = M[n][k] (M row-major [n][D]): 16 n x 16 k block at (n0, k0):
  //   r0,r1 = b0b1,b2b3 of n-tile n0..n0+7 ; r2,r3 = same for n0+8..n0+15
  static __device__ __forceinline__ void load_b(const bf16* base, int n0, int k0, uint32_t (&r)[4]) {
    const int lane = threadIdx.x & 31;
    ldsm_x4(smem_u32(base + (n0 + (lane & 7) + (lane >> 4) * 8) * PITCH + k0 + ((lane >> 3) & 1) * 8), r);
  }
  // A operand with A[m][k] = M[k][m] (M row-major, pitch in elements): 16 m x 16 k block at (m0, k0), transposing load
  static __device__ __forceinline__ void load_at(const bf16* base, int pitch, int m0, int k0, uint32_t (&r)[4]) {
    const int lane = threadIdx.x & 31;
    ldsm_x4_t(smem_u32(base + (k0 + (lane & 7) + (lane >> 4) * 8) * pitch + m0 + ((lane >> 3) & 1) * 8), r);
  }
  // B operand with B[k][n] = M[k][n] (M row-major [k][D]): 16 k x 16 n block at (k0, n0), transposing load
  static __device__ __forceinline__ void load_bt(const bf16* base, int k0, int n0, uint32_t (&r)[4]) {
    const int lane = threadIdx.x & 31;
    ldsm_x4_t(smem_u32(base + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + n0 + (lane >> 4) * 8), r);
  }
};

// Geometry of one group: which streams supply queries / keys, and where their (16-row padded) tiles sit.
struct Geo {
  int qset, kset;          // bit s = stream s has query rows / key rows in this CTA
  int qrows[2], krows[2];  // padded rows (0 if absent); stream 0 first, stream 1 at offset rows[0]
  int RQ, RK;
};
__host__ __device__ __forceinline__ int attn_groups(int pair_mask) { return (pair_mask == 9 || pair_mask == 6) ? 2 : 1; }
__host__ __device__ __forceinline__ Geo make_geo(int pair_mask, int L0, int L1, int group) {
  Geo g;
  if (attn_groups(pair_mask) == 2) {
    g.qset = 1 << group;
    g.kset = pair_mask == 9 ? (1 << group) : (1 << (1 - group));
  } else {
    g.qset = ((pair_mask & 3) ? 1 : 0) | ((pair_mask & 12) ? 2 : 0);
    g.kset = ((pair_mask & 5) ? 1 : 0) | ((pair_mask & 10) ? 2 : 0);
  }
  const int Lp[2] = {(L0 + 15) & ~15, (L1 + 15) & ~15};
  for (int s = 0; s < 2; ++s) {
    g.qrows[s] = ((g.qset >> s) & 1) ? Lp[s] : 0;
    g.krows[s] = ((g.kset >> s) & 1) ? Lp[s] : 0;
  }
  g.RQ = g.qrows[0] + g.qrows[1];
  g.RK = g.krows[0] + g.krows[1];
  return g;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// Asynchronous copy of rows [0, L) of one sample's [L, D] head slice into a padded tile (rows [L, Lp) zero-filled).
template <int D>
__device__ __forceinline__ void load_seg(bf16* dst, const bf16* src, int64_t ld, int L, int Lp) {
  constexpr int CH = D / 8;
  // thread -> (row r0 + k*rstep, fixed 16-byte column chunk): pointer increments only, no div/mod per copy
  const int r0 = threadIdx.x / CH, c = (threadIdx.x % CH) * 8, rstep = blockDim.x / CH;
  uint32_t d = smem_u32(dst + r0 * Tile<D>::PITCH + c);
  const bf16* g = src + (int64_t)r0 * ld + c;
  const uint32_t dstep = rstep * Tile<D>::PITCH * 2;
  const int64_t gstep = (int64_t)rstep * ld;
  for (int r = r0; r < Lp; r += rstep, d += dstep, g += gstep) cp_async16(d, r < L ? (const void*)g : (const void*)src, r < L);
}
// additive key mask in the log2 domain: 0 / -10000*log2(e) (encoders.py:983-991), -inf on padding rows
__device__ __forceinline__ void load_madd(float* dst, const int32_t* mask, int L, int Lp) {
  for (int r = threadIdx.x; r < Lp; r += blockDim.x)
    dst[r] = r < L ? (mask[r] ? 0.f : -14426.950408889634f) : -INFINITY;
}
// 16 rows x D of a warp's own tile rows -> global, 16 bytes per lane (rows of the sample beyond L are dropped)
template <int D>
__device__ __forceinline__ void store_rows16(const bf16* tile_rows, bf16* dst, int64_t ld, int j0, int L) {
  constexpr int CH = D / 8;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 16 * CH / 32; ++i) {
    const int e = i * 32 + lane, r = e / CH, c = (e % CH) * 8;
    if (j0 + r < L)
      *reinterpret_cast<uint4*>(dst + (int64_t)(j0 + r) * ld + c) = *reinterpret_cast<const uint4*>(tile_rows + r * Tile<D>::PITCH + c);
  }
}
// C fragments (16 x D fp32, scaled per row half) -> bf16 in the warp's own 16 tile rows
template <int D>
__device__ __forceinline__ void frag_to_rows16(bf16* tile_rows, const float (&o)[D / 8][4], float s0, float s1) {
  const int lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int nd = 0; nd < D / 8; ++nd) {
    *reinterpret_cast<uint32_t*>(tile_rows + gq * Tile<D>::PITCH + nd * 8 + 2 * tq) = pack_bf16x2(o[nd][0] * s0, o[nd][1] * s0);
    *reinterpret_cast<uint32_t*>(tile_rows + (gq + 8) * Tile<D>::PITCH + nd * 8 + 2 * tq) = pack_bf16x2(o[nd][2] * s1, o[nd][3] * s1);
  }
}

// ---------------------------------------------------------------------------------------------- forward
// (A 3-warp variant capped at 64 registers for ten blocks per SM measured the same as this one: 23.6 / 38.5 us.)
template <int D, int NKB>
__global__ void __launch_bounds__(320)
mma_fwd_kernel(P p) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t smraw[];
  using TL = Tile<D>;
  const Geo g = make_geo(p.pair_mask, p.L[0], p.L[1], blockIdx.y);
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Ks = Qs + g.RQ * TL::PITCH;
  bf16* Vs = Ks + g.RK * TL::PITCH;
  float* madd = reinterpret_cast<float*>(Vs + g.RK * TL::PITCH);
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  const uint64_t seed_eff = effective_seed(p.seed, p.seed_ptr);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;

  const int Le[2] = {p.L[0] ? eff_len(p, 0, b) : 0, p.L[1] ? eff_len(p, 1, b) : 0};
  // first buffer row of this sample and how many of its rows exist there (packed: the valid ones only)
  const int64_t row0[2] = {p.cu[0] ? (int64_t)p.cu[0][b] : (int64_t)b * p.L[0], p.cu[1] ? (int64_t)p.cu[1][b] : (int64_t)b * p.L[1]};
  const int Ls[2] = {p.cu[0] ? Le[0] : p.L[0], p.cu[1] ? Le[1] : p.L[1]};
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const bf16* base = reinterpret_cast<const bf16*>(p.qkv[s]) + row0[s] * p.ld[s] + h * D;
    if (g.qrows[s]) load_seg<D>(Qs + (s ? g.qrows[0] : 0) * TL::PITCH, base, p.ld[s], Le[s], g.qrows[s]);
    if (g.krows[s]) {
      const int off = s ? g.krows[0] : 0;
      load_seg<D>(Ks + off * TL::PITCH, base + p.A, p.ld[s], Le[s], g.krows[s]);
      load_seg<D>(Vs + off * TL::PITCH, base + 2 * p.A, p.ld[s], Le[s], g.krows[s]);
      load_madd(madd + off, p.mask[s] + (int64_t)b * p.L[s], p.L[s], g.krows[s]);
    }
  }
  cp_async_wait_all();
  __syncthreads();
  if (warp * 16 >= g.RQ) return;

  const int qt = warp;
  const int qs = (qt * 16 < g.qrows[0]) ? 0 : 1;
  const int jq0 = qt * 16 - (qs ? g.qrows[0] : 0);  // first row of this tile inside its stream
  if (jq0 >= Le[qs]) {
    // padding-only query tile: its Q rows were zero-filled by the loader - they are the zero context rows to store
    const bf16* zrows = Qs + qt * 16 * TL::PITCH;
    store_rows16<D>(zrows, reinterpret_cast<bf16*>(p.ctx[qs]) + row0[qs] * p.ld_ctx[qs] + h * D, p.ld_ctx[qs], jq0, Ls[qs]);
    if (lane < 16 && jq0 + lane < p.L[qs]) p.lse[qs][((int64_t)b * p.heads + h) * p.L[qs] + jq0 + lane] = 0.f;
    return;
  }
  // key n-tiles (8 keys) this query stream attends: enabled key streams, rows below their length
  uint32_t ntmask = 0;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
    if (g.krows[ks] && pair_on(p.pair_mask, qs, ks))
      ntmask |= ((1u << ((Le[ks] + 7) >> 3)) - 1u) << ((ks ? g.krows[0] : 0) >> 3);
  const float sc2 = p.scale * 1.4426950408889634f;

  float s[2 * NKB][4];
#pragma unroll
  for (int i = 0; i < 2 * NKB; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < D / 16; ++kk) {
    uint32_t a[4];
    TL::load_a(Qs, qt * 16, kk * 16, a);
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb) {
      if ((ntmask >> (2 * kb)) & 3u) {
        uint32_t bb[4];
        TL::load_b(Ks, kb * 16, kk * 16, bb);
        mma16816(s[2 * kb], a, bb[0], bb[1]);
        if ((ntmask >> (2 * kb + 1)) & 1u) mma16816(s[2 * kb + 1], a, bb[2], bb[3]);
      }
    }
  }
  // scale + additive mask (log2 domain), row max over rows gq and gq+8
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int nt = 0; nt < 2 * NKB; ++nt) {
    if ((ntmask >> nt) & 1u) {
      const float2 md = *reinterpret_cast<const float2*>(madd + nt * 8 + 2 * tq);
      s[nt][0] = fmaf(s[nt][0], sc2, md.x); s[nt][1] = fmaf(s[nt][1], sc2, md.y);
      s[nt][2] = fmaf(s[nt][2], sc2, md.x); s[nt][3] = fmaf(s[nt][3], sc2, md.y);
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    if (mx[r] == -INFINITY) mx[r] = 0.f;
  }
  float sum[2] = {0.f, 0.f};
#pragma unroll
  for (int nt = 0; nt < 2 * NKB; ++nt) {
    if ((ntmask >> nt) & 1u) {
      s[nt][0] = ex2f(s[nt][0] - mx[0]); s[nt][1] = ex2f(s[nt][1] - mx[0]);
      s[nt][2] = ex2f(s[nt][2] - mx[1]); s[nt][3] = ex2f(s[nt][3] - mx[1]);
      sum[0] += s[nt][0] + s[nt][1];
      sum[1] += s[nt][2] + s[nt][3];
    }
  }
  float osc[2];  // 1/sum (and the dropout scale) is applied to O, not to the 16 x Lk probabilities
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
    sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
    const int j = jq0 + gq + r * 8;
    osc[r] = (j < Le[qs] && sum[r] > 0.f) ? p.keep_scale[qs] / sum[r] : 0.f;   // padding rows of a partly valid tile: zeros, like whole padding tiles
    if (tq == 0 && j < p.L[qs]) p.lse[qs][((int64_t)b * p.heads + h) * p.L[qs] + j] = (mx[r] + __log2f(sum[r])) * 0.6931471805599453f;
  }
  const uint32_t thresh = p.thresh[qs];
  if (thresh) {
    const int Ltot = p.L[0] + p.L[1];
    const uint32_t row0 = ((uint32_t)b * p.heads + h) * Ltot + (qs ? p.L[0] : 0) + jq0 + gq;
    const uint32_t rt0 = attn_rowterm(seed_eff, row0), rt1 = attn_rowterm(seed_eff, row0 + 8);
    const uint32_t thr_hi = thresh << 16;
#pragma unroll
    for (int nt = 0; nt < 2 * NKB; ++nt) {
      if ((ntmask >> nt) & 1u) {
        const int kr = nt * 8 + 2 * tq;
        const int ks = (kr < g.krows[0]) ? 0 : 1;
        const uint32_t ct = attn_colterm(seed_eff, ks, kr - (ks ? g.krows[0] : 0));
        const uint32_t h0 = attn_mix(rt0 ^ ct), h1 = attn_mix(rt1 ^ ct);
        if ((h0 << 16) < thr_hi) s[nt][0] = 0.f;   // low half: even key;  (r16 < thresh) <=> (r16 << 16 < thresh << 16)
        if (h0 < thr_hi) s[nt][1] = 0.f;           // high half: odd key
        if ((h1 << 16) < thr_hi) s[nt][2] = 0.f;
        if (h1 < thr_hi) s[nt][3] = 0.f;
      }
    }
  }
  // O = P . V
  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
#pragma unroll
  for (int kb = 0; kb < NKB; ++kb) {
    if ((ntmask >> (2 * kb)) & 3u) {
      uint32_t a[4];
      a[0] = pack_bf16x2(s[2 * kb][0], s[2 * kb][1]);
      a[1] = pack_bf16x2(s[2 * kb][2], s[2 * kb][3]);
      a[2] = pack_bf16x2(s[2 * kb + 1][0], s[2 * kb + 1][1]);   // zero when that n-tile is disabled
      a[3] = pack_bf16x2(s[2 * kb + 1][2], s[2 * kb + 1][3]);
#pragma unroll
      for (int nd = 0; nd < D / 16; ++nd) {
        uint32_t bb[4];
        TL::load_bt(Vs, kb * 16, nd * 16, bb);
        mma16816(o[2 * nd], a, bb[0], bb[1]);
        mma16816(o[2 * nd + 1], a, bb[2], bb[3]);
      }
    }
  }
  // the warp's 16 Q rows are dead (only this warp read them): stage O there, then 16-byte coalesced stores
  bf16* mine = Qs + qt * 16 * TL::PITCH;
  __syncwarp();
  frag_to_rows16<D>(mine, o, osc[0], osc[1]);
  __syncwarp();
  store_rows16<D>(mine, reinterpret_cast<bf16*>(p.ctx[qs]) + row0[qs] * p.ld_ctx[qs] + h * D, p.ld_ctx[qs], jq0, Ls[qs]);
}

// ---------------------------------------------------------------------------------------------- backward
// SMALL: the block has at most 3 warps (<= 48 padded query and key rows per group - every ctrl_* pre-training shape):
// capping the registers at 65536 / (96 * 6) lets six blocks share an SM (the kernel is issue-bound at 45 % with five).
template <int D, bool SMALL>
__global__ void __launch_bounds__(SMALL ? 96 : 320, SMALL ? 6 : 1)
mma_bwd_kernel(P p) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t smraw[];
  using TL = Tile<D>;
  constexpr int CH = D / 8;
  const Geo g = make_geo(p.pair_mask, p.L[0], p.L[1], blockIdx.y);
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* dOs = Qs + g.RQ * TL::PITCH;
  bf16* Ks = dOs + g.RQ * TL::PITCH;
  bf16* Vs = Ks + g.RK * TL::PITCH;
  float* madd = reinterpret_cast<float*>(Vs + g.RK * TL::PITCH);
  float* lse2 = madd + g.RK;                       // per query row: lse * log2(e); +inf on padding rows (P = 0)
  float* delta = lse2 + g.RQ;                      // rowsum(dO * O)
  uint32_t* rowt = reinterpret_cast<uint32_t*>(delta + g.RQ);  // dropout row terms
  bf16* dSt = reinterpret_cast<bf16*>(rowt + g.RQ);            // dS^T [RK][RQ + 8]
  bf16* Os = dSt;                                              // O tile [RQ][PITCH]: dead (delta done) before dS^T is written
  const int SP = g.RQ + 8;
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  const uint64_t seed_eff = effective_seed(p.seed, p.seed_ptr);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int Ltot = p.L[0] + p.L[1];

  const int Le[2] = {p.L[0] ? eff_len(p, 0, b) : 0, p.L[1] ? eff_len(p, 1, b) : 0};
  const int64_t row0[2] = {p.cu[0] ? (int64_t)p.cu[0][b] : (int64_t)b * p.L[0], p.cu[1] ? (int64_t)p.cu[1][b] : (int64_t)b * p.L[1]};
  const int Ls[2] = {p.cu[0] ? Le[0] : p.L[0], p.cu[1] ? Le[1] : p.L[1]};
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const bf16* base = reinterpret_cast<const bf16*>(p.qkv[s]) + row0[s] * p.ld[s] + h * D;
    if (g.qrows[s]) load_seg<D>(Qs + (s ? g.qrows[0] : 0) * TL::PITCH, base, p.ld[s], Le[s], g.qrows[s]);
    if (g.krows[s]) {
      const int off = s ? g.krows[0] : 0;
      load_seg<D>(Ks + off * TL::PITCH, base + p.A, p.ld[s], Le[s], g.krows[s]);
      load_seg<D>(Vs + off * TL::PITCH, base + 2 * p.A, p.ld[s], Le[s], g.krows[s]);
      load_madd(madd + off, p.mask[s] + (int64_t)b * p.L[s], p.L[s], g.krows[s]);
    }
  }
  // dO and O tiles (O only feeds delta; its tile is reused for nothing else), per-query-row scalars
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (!g.qrows[s]) continue;
    const int off = s ? g.qrows[0] : 0;
    const int64_t head0 = row0[s] * p.ld_ctx[s] + h * D;
    load_seg<D>(dOs + off * TL::PITCH, reinterpret_cast<const bf16*>(p.dctx[s]) + head0, p.ld_ctx[s], Le[s], g.qrows[s]);
    load_seg<D>(Os + off * TL::PITCH, reinterpret_cast<const bf16*>(p.ctx[s]) + head0, p.ld_ctx[s], Le[s], g.qrows[s]);
    for (int r = threadIdx.x; r < g.qrows[s]; r += blockDim.x) {
      lse2[off + r] = r < Le[s] ? p.lse[s][((int64_t)b * p.heads + h) * p.L[s] + r] * 1.4426950408889634f : INFINITY;
      rowt[off + r] = attn_rowterm(seed_eff, ((uint32_t)b * p.heads + h) * Ltot + (s ? p.L[0] : 0) + r);
    }
  }
  cp_async_wait_all();
  __syncthreads();
  // delta[q] = sum_d dO[q,d] * O[q,d] = diag(dO . O^T): one 16x16 product per query tile on the tensor core
  // (8 HMMA) instead of a per-element dot product; lane (gq, tq = gq/2) holds the diagonal entries of rows gq, gq+8.
  if (warp * 16 < g.RQ) {
    float dd[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      uint32_t a[4], bb[4];
      TL::load_a(dOs, warp * 16, kk * 16, a);
      TL::load_b(Os, warp * 16, kk * 16, bb);
      mma16816(dd[0], a, bb[0], bb[1]);
      mma16816(dd[1], a, bb[2], bb[3]);
    }
    if (tq == (gq >> 1)) {
      delta[warp * 16 + gq] = (gq & 1) ? dd[0][1] : dd[0][0];
      delta[warp * 16 + gq + 8] = (gq & 1) ? dd[1][3] : dd[1][2];
    }
  }
  __syncthreads();

  if (warp * 16 < g.RK) {
    const int kt = warp;
    const int ks = (kt * 16 < g.krows[0]) ? 0 : 1;
    const int jk0 = kt * 16 - (ks ? g.krows[0] : 0);
    const float sc2 = p.scale * 1.4426950408889634f;
    float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    const float md0 = madd[kt * 16 + gq], md1 = madd[kt * 16 + gq + 8];
    const uint32_t ct0 = attn_colterm(seed_eff, ks, jk0 + gq), ct1 = attn_colterm(seed_eff, ks, jk0 + gq + 8);
    const int hshl = (gq & 1) ? 0 : 16;  // keys jk0+gq and jk0+gq+8 have the parity of gq: which half of the pair hash
    const uint32_t thr_hi[2] = {p.thresh[0] << 16, p.thresh[1] << 16};
    const float kscs[2] = {p.keep_scale[0], p.keep_scale[1]};
    const bool k_live = jk0 < Le[ks];   // a padding-only key tile gets dK = dV = 0 (stored below)
#pragma unroll 1
    for (int qb = 0; k_live && qb * 16 < g.RQ; ++qb) {
      const int qs = (qb * 16 < g.qrows[0]) ? 0 : 1;
      if (!pair_on(p.pair_mask, qs, ks)) continue;
      if (qb * 16 - (qs ? g.qrows[0] : 0) >= Le[qs]) continue;   // padding-only query tile: P = 0, nothing to add
      const uint32_t thr = qs ? thr_hi[1] : thr_hi[0];   // (r16 < thresh)  <=>  ((h << hshl) < thresh << 16)
      const float ksc = qs ? kscs[1] : kscs[0];
      float st[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dpt[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        uint32_t ak[4], av[4], bq[4], bdo[4];
        TL::load_a(Ks, kt * 16, kk * 16, ak);
        TL::load_a(Vs, kt * 16, kk * 16, av);
        TL::load_b(Qs, qb * 16, kk * 16, bq);
        TL::load_b(dOs, qb * 16, kk * 16, bdo);
        mma16816(st[0], ak, bq[0], bq[1]);
        mma16816(st[1], ak, bq[2], bq[3]);
        mma16816(dpt[0], av, bdo[0], bdo[1]);
        mma16816(dpt[1], av, bdo[2], bdo[3]);
      }
      float pd[2][4], ds[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int qr = qb * 16 + nt * 8 + 2 * tq;  // query rows qr, qr+1 (tile columns)
        const float2 l2 = *reinterpret_cast<const float2*>(lse2 + qr);
        const float2 dl = *reinterpret_cast<const float2*>(delta + qr);
        const uint2 rt = *reinterpret_cast<const uint2*>(rowt + qr);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pr = ex2f(fmaf(st[nt][e], sc2, (e >> 1) ? md1 : md0) - ((e & 1) ? l2.y : l2.x));
          const uint32_t hh = attn_mix(((e & 1) ? rt.y : rt.x) ^ ((e >> 1) ? ct1 : ct0)) << hshl;
          const bool keep = hh >= thr;              // thr == 0 (no dropout) keeps everything
          pd[nt][e] = keep ? pr * ksc : 0.f;
          const float dpe = keep ? dpt[nt][e] * ksc : 0.f;
          ds[nt][e] = pr * (dpe - ((e & 1) ? dl.y : dl.x)) * p.scale;
        }
      }
      uint32_t apd[4], ads[4];
      apd[0] = pack_bf16x2(pd[0][0], pd[0][1]); apd[1] = pack_bf16x2(pd[0][2], pd[0][3]);
      apd[2] = pack_bf16x2(pd[1][0], pd[1][1]); apd[3] = pack_bf16x2(pd[1][2], pd[1][3]);
      ads[0] = pack_bf16x2(ds[0][0], ds[0][1]); ads[1] = pack_bf16x2(ds[0][2], ds[0][3]);
      ads[2] = pack_bf16x2(ds[1][0], ds[1][1]); ads[3] = pack_bf16x2(ds[1][2], ds[1][3]);
      {  // dS^T tile -> shared memory for the dQ phase
        uint32_t* r0 = reinterpret_cast<uint32_t*>(dSt + (kt * 16 + gq) * SP + qb * 16 + 2 * tq);
        uint32_t* r1 = reinterpret_cast<uint32_t*>(dSt + (kt * 16 + gq + 8) * SP + qb * 16 + 2 * tq);
        r0[0] = ads[0]; r1[0] = ads[1]; r0[4] = ads[2]; r1[4] = ads[3];
      }
#pragma unroll
      for (int nd = 0; nd < D / 16; ++nd) {
        uint32_t bdo[4], bq[4];
        TL::load_bt(dOs, qb * 16, nd * 16, bdo);
        TL::load_bt(Qs, qb * 16, nd * 16, bq);
        mma16816(dv[2 * nd], apd, bdo[0], bdo[1]);
        mma16816(dv[2 * nd + 1], apd, bdo[2], bdo[3]);
        mma16816(dk[2 * nd], ads, bq[0], bq[1]);
        mma16816(dk[2 * nd + 1], ads, bq[2], bq[3]);
      }
    }
    // dV, then dK: staged in the warp's own (now dead) V rows - its K rows are still read by the dQ phase -
    // and stored as 16-byte coalesced rows
    if (p.dqkv[ks] != nullptr) {
      bf16* dst = reinterpret_cast<bf16*>(p.dqkv[ks]) + row0[ks] * p.ld[ks] + h * D;
      bf16* mv = Vs + kt * 16 * TL::PITCH;
      __syncwarp();
      frag_to_rows16<D>(mv, dv, 1.f, 1.f);
      __syncwarp();
      store_rows16<D>(mv, dst + 2 * p.A, p.ld[ks], jk0, Ls[ks]);
      __syncwarp();
      frag_to_rows16<D>(mv, dk, 1.f, 1.f);
      __syncwarp();
      store_rows16<D>(mv, dst + p.A, p.ld[ks], jk0, Ls[ks]);
    }
  }
  __syncthreads();
  // ---- dQ = dS . K : one warp per 16-query tile, dS through ldmatrix.trans from the dS^T tile
  if (warp * 16 < g.RQ) {
    const int qt = warp;
    const int qs = (qt * 16 < g.qrows[0]) ? 0 : 1;
    const int jq0 = qt * 16 - (qs ? g.qrows[0] : 0);
    float dq[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }
    const bool q_live = jq0 < Le[qs];
    for (int kt = 0; q_live && kt * 16 < g.RK; ++kt) {
      const int ks = (kt * 16 < g.krows[0]) ? 0 : 1;
      if (!pair_on(p.pair_mask, qs, ks)) continue;
      if (kt * 16 - (ks ? g.krows[0] : 0) >= Le[ks]) continue;   // that dS^T tile was never written
      uint32_t a[4];
      TL::load_at(dSt, SP, qt * 16, kt * 16, a);
#pragma unroll
      for (int nd = 0; nd < D / 16; ++nd) {
        uint32_t bk[4];
        TL::load_bt(Ks, kt * 16, nd * 16, bk);
        mma16816(dq[2 * nd], a, bk[0], bk[1]);
        mma16816(dq[2 * nd + 1], a, bk[2], bk[3]);
      }
    }
    if (p.dqkv[qs] != nullptr) {
      bf16* mine = Qs + qt * 16 * TL::PITCH;  // nobody reads Q after the barrier
      frag_to_rows16<D>(mine, dq, 1.f, 1.f);
      __syncwarp();
      store_rows16<D>(mine, reinterpret_cast<bf16*>(p.dqkv[qs]) + row0[qs] * p.ld[qs] + h * D, p.ld[qs], jq0, Ls[qs]);
    }
  }
  // streams that only supply keys (queries) get zero dQ (dK, dV)
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (p.dqkv[s] == nullptr) continue;
    bf16* dst = reinterpret_cast<bf16*>(p.dqkv[s]) + row0[s] * p.ld[s] + h * D;
    if (g.qrows[s]) {
    } else if (g.krows[s] && attn_groups(p.pair_mask) == 1) {
      for (int e = threadIdx.x; e < Ls[s] * CH; e += blockDim.x)
        *reinterpret_cast<uint4*>(dst + (int64_t)(e / CH) * p.ld[s] + (e % CH) * 8) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (g.qrows[s] && !g.krows[s] && attn_groups(p.pair_mask) == 1) {
      for (int e = threadIdx.x; e < Ls[s] * CH * 2; e += blockDim.x) {
        const int r = e / (2 * CH), c = (e % (2 * CH)) * 8;  // K and V column blocks
        *reinterpret_cast<uint4*>(dst + (int64_t)r * p.ld[s] + p.A + (c / D) * p.A + (c % D)) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
}


// Attention probabilities on request (config.visualization + output_all_attention_masks, volta/encoders.py:342-356,
// 873-875): the path never materialises P, so this separate kernel recomputes it for inspection - one warp per query row,
// joint softmax over the enabled key streams, no dropout (the reference's visualisation is read in eval mode).
// out[qs]: fp32 [B, heads, L[qs], L[0] + L[1]], text keys first; columns of a key stream the row does not attend are 0.
template <typename T>
__global__ void __launch_bounds__(128)
simt_probs_kernel(P p, float* out_t, float* out_v) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qs = blockIdx.y;
  float* out = qs ? out_v : out_t;
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  if (p.qkv[qs] == nullptr || out == nullptr) return;
  const int Ltot = p.L[0] + p.L[1];
  float* qbuf = sm + warp * (p.d + Ltot);
  float* pbuf = qbuf + p.d;
  const T* Q = reinterpret_cast<const T*>(p.qkv[qs]);
  for (int i = warp; i < p.L[qs]; i += 4) {
    const int64_t qrow = (int64_t)b * p.L[qs] + i;
    for (int dd = lane; dd < p.d; dd += 32) qbuf[dd] = to_f<T>(Q[qrow * p.ld[qs] + h * p.d + dd]);
    __syncwarp();
    float mx = -INFINITY;
    for (int kg = lane; kg < Ltot; kg += 32) {
      const int ks = kg < p.L[0] ? 0 : 1;
      const int j = ks ? kg - p.L[0] : kg;
      float s = -INFINITY;
      if (pair_on(p.pair_mask, qs, ks) && p.qkv[ks] != nullptr) {
        const T* K = reinterpret_cast<const T*>(p.qkv[ks]) + ((int64_t)b * p.L[ks] + j) * p.ld[ks] + p.A + h * p.d;
        float acc = 0.f;
        for (int dd = 0; dd < p.d; ++dd) acc = fmaf(qbuf[dd], to_f<T>(K[dd]), acc);
        s = acc * p.scale + (p.mask[ks][(int64_t)b * p.L[ks] + j] ? 0.f : -10000.0f);
      }
      pbuf[kg] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    if (mx == -INFINITY) mx = 0.f;
    float sum = 0.f;
    for (int kg = lane; kg < Ltot; kg += 32) sum += __expf(pbuf[kg] - mx);
    sum = warp_sum(sum);
    float* orow = out + (((int64_t)b * p.heads + h) * p.L[qs] + i) * Ltot;
    for (int kg = lane; kg < Ltot; kg += 32) orow[kg] = sum > 0.f ? __expf(pbuf[kg] - mx) / sum : 0.f;
    __syncwarp();
  }
}

// =====================================================================================
// tcgen05 back-end (forward): S = Q K^T and O = P V on the 5th-generation tensor cores, accumulators in TMEM,
// operands brought in by TMA.
//
// A (sample, head) of this model is 20-101 rows - a quarter of the 128-row tcgen05 tile.  But consecutive samples are
// consecutive ROWS of the [rows, 3A] Q/K/V buffer (packed or not), so ONE 128-row TMA box at (row of the first sample,
// column of head h) brings in the Q (or K, or V) rows of G = floor(128 / L) samples at once.  The CTA computes the whole
// 128 x 128 score block of those G samples per key stream with d/16 MMAs, and the softmax - one thread per query row,
// reading its row of S straight out of TMEM - keeps only the columns of its own sample (block-diagonal) before P goes
// back to shared memory (bf16, the K-major swizzled layout the MMA reads) for O = P V.  Masked keys, the joint softmax
// over [text keys | vision keys], the dropout hash, the saved row log-sum-exp and the zero rows of padding queries are
// those of the warp-level kernel above, so the backward (which recomputes P from lse) is unchanged.
// One CTA = (group of G samples, head, query stream): 128 threads, thread 0 also issues the TMA loads and the MMAs.
// =====================================================================================
namespace tcf {
using namespace vb::tc;

constexpr int MAXG = 8;
constexpr int TILE = 128;

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes) {
  // cute::UMMA::SmemDescriptor, SWIZZLE_128B: start address, leading-dimension byte offset, stride byte offset 1024
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)(1024u >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}

struct Meta {
  int qoff[MAXG], qspan[MAXG], qlen[MAXG];   // tile rows of sample s: [qoff, qoff + qspan); valid queries: the first qlen
  int koff[2][MAXG], klen[2][MAXG];          // tile columns (per key stream) of the valid keys of sample s
  int kend[2];                               // last key column + 1 per key stream
  int G;
};

template <int D, int NKS>
__global__ void __launch_bounds__(128)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1, const P p, const int G,
                   const int qs_mask) {
  extern __shared__ uint8_t smraw[];
  constexpr uint32_t TILE_BYTES = TILE * D * 2;                 // one 128-row operand tile: D/64 k-blocks of 16 KB
  constexpr uint32_t P_BYTES = 2u * 16384u;                     // 128 rows x 128 key columns, bf16 (two 64-column k-blocks)
  constexpr uint32_t QK_BYTES = (1 + NKS) * TILE_BYTES;
  constexpr uint32_t QKP_BYTES = QK_BYTES > NKS * P_BYTES ? QK_BYTES : NKS * P_BYTES;   // P reuses the Q / K tiles
  const uint32_t base = (smem_u32(smraw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK0 = base + TILE_BYTES, sV0 = base + QKP_BYTES, sP0 = base;
  uint8_t* gen = smraw + (base - smem_u32(smraw));              // generic pointer to the aligned base
  uint8_t* tail = gen + QKP_BYTES + NKS * TILE_BYTES;
  float* madd = reinterpret_cast<float*>(tail);                 // [2][128]
  Meta* meta = reinterpret_cast<Meta*>(tail + 1024);
  const uint32_t bars = base + QKP_BYTES + NKS * TILE_BYTES + 1024 + 512;
  const uint32_t bar_tma = bars, bar_mma = bars + 8, tmem_slot = bars + 16;

  const int t = threadIdx.x, warp = t >> 5;
  // which query stream this CTA serves (blockIdx.y-th set bit of qs_mask)
  int qs = (qs_mask & 1) ? 0 : 1;
  if (blockIdx.y == 1) qs = 1;
  const int grp = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  const int b0 = grp * G;
  const int nG = min(G, p.B - b0);
  const bool ks_on[2] = {pair_on(p.pair_mask, qs, 0) && p.qkv[0] != nullptr, pair_on(p.pair_mask, qs, 1) && p.qkv[1] != nullptr};

  if (t == 0) {
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
    fence_proxy_async();
    tma_prefetch_desc(&map0);
    tma_prefetch_desc(&map1);
  }
  if (warp == 0) tmem_alloc<NKS * 128, 1>(tmem_slot);
  // ---- sample geometry of this group
  __shared__ int64_t base_row[2];
  if (t < 2) {
    const int st = t;
    int64_t r0 = 0;
    if (p.qkv[st] != nullptr) r0 = p.cu[st] ? (int64_t)p.cu[st][b0] : (int64_t)b0 * p.L[st];
    base_row[st] = r0;
  }
  if (t == 0) meta->G = nG;
  __syncthreads();
  if (t < nG) {
    const int b = b0 + t;
    for (int st = 0; st < 2; ++st) {
      if (p.qkv[st] == nullptr) { if (st == qs) { meta->qoff[t] = 0; meta->qspan[t] = 0; meta->qlen[t] = 0; } meta->koff[st][t] = 0; meta->klen[st][t] = 0; continue; }
      const int Le = eff_len(p, st, b);
      const int64_t r0 = p.cu[st] ? (int64_t)p.cu[st][b] : (int64_t)b * p.L[st];
      const int off = (int)(r0 - base_row[st]);
      if (st == qs) { meta->qoff[t] = off; meta->qspan[t] = p.cu[st] ? Le : p.L[st]; meta->qlen[t] = Le; }
      meta->koff[st][t] = off; meta->klen[st][t] = ks_on[st] ? Le : 0;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (t < 2) {
    int e = 0;
    for (int s_ = 0; s_ < nG; ++s_) e = max(e, meta->koff[t][s_] + meta->klen[t][s_]);
    meta->kend[t] = ks_on[t] ? min(e, TILE) : 0;
  }
  // ---- TMA: Q tile of the query stream, K and V tiles of every enabled key stream
  if (t == 0) {
    int nks = (ks_on[0] ? 1 : 0) + (ks_on[1] ? 1 : 0);
    mbar_expect_tx(bar_tma, (1 + 2 * nks) * TILE_BYTES);
    const CUtensorMap* mq = qs ? &map1 : &map0;
#pragma unroll
    for (int kb = 0; kb < D / 64; ++kb) tma_load_2d(sQ + kb * 16384, mq, bar_tma, h * D + kb * 64, (int32_t)base_row[qs]);
    int slot = 0;
    for (int ks = 0; ks < 2; ++ks) {
      if (!ks_on[ks]) continue;
      const CUtensorMap* mk = ks ? &map1 : &map0;
#pragma unroll
      for (int kb = 0; kb < D / 64; ++kb) {
        tma_load_2d(sK0 + slot * TILE_BYTES + kb * 16384, mk, bar_tma, p.A + h * D + kb * 64, (int32_t)base_row[ks]);
        tma_load_2d(sV0 + slot * TILE_BYTES + kb * 16384, mk, bar_tma, 2 * p.A + h * D + kb * 64, (int32_t)base_row[ks]);
      }
      ++slot;
    }
  }
  __syncthreads();   // kend visible
  // additive key masks (log2 domain) per tile column, while the loads are in flight
  for (int ks = 0; ks < 2; ++ks) {
    float v = -INFINITY;
    if (ks_on[ks]) {
      for (int s_ = 0; s_ < nG; ++s_) {
        const int j = t - meta->koff[ks][s_];
        if (j >= 0 && j < meta->klen[ks][s_]) v = p.mask[ks][(int64_t)(b0 + s_) * p.L[ks] + j] ? 0.f : -14426.950408889634f;
      }
    }
    madd[ks * TILE + t] = v;
  }
  __syncthreads();   // every thread reads other threads' mask entries
  mbar_wait(bar_tma, 0);
  // V rows between the last valid key and the end of its 16-key MMA step may be memory nobody wrote: zero them
  {
    int slot = 0;
    for (int ks = 0; ks < 2; ++ks) {
      if (!ks_on[ks]) continue;
      const int e0 = meta->kend[ks], e1 = min(TILE, (e0 + 15) & ~15);
      for (int i = t; i < (e1 - e0) * (D / 64) * 8; i += 128) {
        const int row = e0 + i / ((D / 64) * 8), rem = i % ((D / 64) * 8), kb = rem / 8, c16 = rem % 8;
        *reinterpret_cast<uint4*>(gen + (sV0 - base) + slot * TILE_BYTES + kb * 16384 + row * 128 + c16 * 16) = make_uint4(0u, 0u, 0u, 0u);
      }
      ++slot;
    }
  }
  // ---- S = Q K^T per key stream (thread 0), accumulators at TMEM columns [ks_slot * 128, +128)
  if (t == 0) {
    tcgen05_fence_after();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int slot = 0;
    for (int ks = 0; ks < 2; ++ks) {
      if (!ks_on[ks]) continue;
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        const uint32_t off = (kk / 4) * 16384 + (kk % 4) * 32;
        umma_bf16(tmem_base + slot * 128, smem_desc(sQ + off, 16), smem_desc(sK0 + slot * TILE_BYTES + off, 16), idesc, kk > 0 ? 1u : 0u);
      }
      ++slot;
    }
    umma_commit(bar_mma);
  }
  // ---- softmax: thread t = query row t of the tile
  int sidx = -1;
  for (int s_ = 0; s_ < nG; ++s_)
    if (t >= meta->qoff[s_] && t < meta->qoff[s_] + meta->qspan[s_]) sidx = s_;
  const int jq = sidx >= 0 ? t - meta->qoff[sidx] : 0;
  const bool row_valid = sidx >= 0 && jq < meta->qlen[sidx];
  // key slots of this CTA (compile-time indexed: no local-memory arrays): slot -> key stream, this row's column range
  // [lo, lo + n) inside the slot, the columns the P V MMAs read (kend rounded up to 16)
  int slot_ks[NKS], lo[NKS], n[NKS], kend16[NKS];
  {
    int sl = 0;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      if (ks_on[ks] && sl < NKS) {
        slot_ks[sl] = ks;
        lo[sl] = row_valid ? meta->koff[ks][sidx] : 0;
        n[sl] = row_valid ? min(TILE - lo[sl], meta->klen[ks][sidx]) : 0;
        kend16[sl] = min(TILE, (meta->kend[ks] + 15) & ~15);
        ++sl;
      }
    }
#pragma unroll
    for (int q = 0; q < NKS; ++q)
      if (q >= sl) { slot_ks[q] = 0; lo[q] = 0; n[q] = 0; kend16[q] = 0; }
  }
  // 32-column chunks this WARP needs per slot: the union of its 32 rows' ranges (consecutive rows belong to consecutive
  // samples, so the union is short); chunks outside it hold no probability of these rows and are written as zeros
  int ch0[NKS], ch1[NKS];
#pragma unroll
  for (int q = 0; q < NKS; ++q) {
    const int a0 = __reduce_min_sync(0xffffffffu, n[q] > 0 ? lo[q] : TILE);
    const int a1 = __reduce_max_sync(0xffffffffu, n[q] > 0 ? lo[q] + n[q] : 0);
    ch0[q] = a0 >> 5;
    ch1[q] = (a1 + 31) >> 5;          // empty when a1 == 0
    if (a1 == 0) ch0[q] = 0;
  }
  const float sc2 = p.scale * 1.4426950408889634f;
  mbar_wait(bar_mma, 0);
  tcgen05_fence_after();
  const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
  float mx = -INFINITY;
#pragma unroll
  for (int q = 0; q < NKS; ++q) {
    const float* md = madd + slot_ks[q] * TILE;
    for (int ch = ch0[q]; ch < ch1[q]; ++ch) {
      uint32_t raw[32];
      tmem_ld32(trow + q * 128 + ch * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = ch * 32 + j;
        const float v = fmaf(__uint_as_float(raw[j]), sc2, md[col]);
        mx = fmaxf(mx, ((unsigned)(col - lo[q]) < (unsigned)n[q]) ? v : -INFINITY);
      }
    }
  }
  if (mx == -INFINITY) mx = 0.f;
  float sum = 0.f;
  const uint32_t thresh = p.thresh[qs];
  const uint64_t seed_eff = effective_seed(p.seed, p.seed_ptr);
  const int Ltot = p.L[0] + p.L[1];
  const uint32_t rt = attn_rowterm(seed_eff, ((uint32_t)(b0 + max(sidx, 0)) * p.heads + h) * Ltot + (qs ? p.L[0] : 0) + jq);
#pragma unroll
  for (int q = 0; q < NKS; ++q) {
    const float* md = madd + slot_ks[q] * TILE;
    uint8_t* prow = gen + (sP0 - base) + q * P_BYTES + (t >> 3) * 1024 + (t & 7) * 128;
    const int nch = (kend16[q] + 31) >> 5;             // every column the P V MMAs will read gets written
    for (int ch = 0; ch < nch; ++ch) {
      uint32_t packed[16];
      if (ch >= ch0[q] && ch < ch1[q]) {               // warp-uniform
        uint32_t raw[32];
        tmem_ld32(trow + q * 128 + ch * 32, raw);
        tmem_ld_wait();
        float pr[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = ch * 32 + j;
          const float e = ex2f(fmaf(__uint_as_float(raw[j]), sc2, md[col]) - mx);
          pr[j] = ((unsigned)(col - lo[q]) < (unsigned)n[q]) ? e : 0.f;
          sum += pr[j];
        }
        if (thresh) {                                  // uniform: dropout on the probabilities of this query stream
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int jk = ch * 32 + j - lo[q];
            const uint32_t hsh = attn_mix(rt ^ attn_colterm(seed_eff, slot_ks[q], jk));
            if (((jk & 1) ? (hsh >> 16) : (hsh & 0xFFFFu)) < thresh) pr[j] = 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 32; j += 2) packed[j >> 1] = pack_bf16x2(pr[j], pr[j + 1]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) packed[j] = 0u;
      }
      // 32 columns = 4 chunks of 16 bytes in k-block (ch / 2), 128B-swizzled: chunk index ^= row % 8
      uint8_t* kbp = prow + (ch >> 1) * 16384;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int c16 = ((ch & 1) * 4 + q4) ^ (t & 7);
        *reinterpret_cast<uint4*>(kbp + c16 * 16) = make_uint4(packed[q4 * 4], packed[q4 * 4 + 1], packed[q4 * 4 + 2], packed[q4 * 4 + 3]);
      }
    }
  }
  // every thread's reads of S and writes of P (and of the zeroed V rows) are done before the second MMA overwrites S with O
  tcgen05_fence_before();
  fence_proxy_async();
  __syncthreads();
  if (t == 0) {
    tcgen05_fence_after();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int slot = 0;
    bool first = true;
    for (int ks = 0; ks < 2; ++ks) {
      if (!ks_on[ks]) continue;
      const int nk16 = (meta->kend[ks] + 15) >> 4;
      for (int k16 = 0; k16 < nk16; ++k16) {
        const uint64_t adesc = smem_desc(sP0 + slot * P_BYTES + (k16 / 4) * 16384 + (k16 % 4) * 32, 16);
        const uint64_t bdesc = smem_desc(sV0 + slot * TILE_BYTES + k16 * 2048, 16384);
        umma_bf16(tmem_base, adesc, bdesc, idesc, first ? 0u : 1u);
        first = false;
      }
      ++slot;
    }
    umma_commit(bar_mma);
  }
  // ---- epilogue: O row / sum -> bf16 -> ctx; row log-sum-exp
  const float osc = (row_valid && sum > 0.f) ? p.keep_scale[qs] / sum : 0.f;
  if (sidx >= 0 && jq < p.L[qs])
    p.lse[qs][((int64_t)(b0 + sidx) * p.heads + h) * p.L[qs] + jq] = row_valid ? (mx + __log2f(sum)) * 0.6931471805599453f : 0.f;
  mbar_wait(bar_mma, 1);
  tcgen05_fence_after();
  const bool any_keys = (meta->kend[0] + meta->kend[1]) > 0;
  bf16* out = reinterpret_cast<bf16*>(p.ctx[qs]) + (base_row[qs] + t) * p.ld_ctx[qs] + h * D;
#pragma unroll
  for (int ch = 0; ch < D / 32; ++ch) {
    uint32_t raw[32];
    tmem_ld32(trow + ch * 32, raw);     // warp-collective: every lane takes part, only rows of a sample are stored
    tmem_ld_wait();
    if (sidx >= 0) {
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint32_t o4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float a0 = any_keys ? __uint_as_float(raw[v * 8 + 2 * u]) * osc : 0.f;
          const float a1 = any_keys ? __uint_as_float(raw[v * 8 + 2 * u + 1]) * osc : 0.f;
          o4[u] = pack_bf16x2(osc != 0.f ? a0 : 0.f, osc != 0.f ? a1 : 0.f);
        }
        *reinterpret_cast<uint4*>(out + ch * 32 + v * 8) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc<NKS * 128, 1>(tmem_base);
  }
}

template <int D, int NKS> static size_t smem_bytes() {
  const size_t tile = (size_t)TILE * D * 2, pb = 2 * 16384;
  const size_t qk = (1 + NKS) * tile, qkp = qk > NKS * pb ? qk : NKS * pb;
  return 1024 + qkp + NKS * tile + 1024 + 512 + 64;
}

// Can the tcgen05 forward run this call?  Every stream's sample must fit a 128-row tile.
static bool ok(const vb_attn_args& a, const P& p) {
  if (a.dtype != VB_BF16 || (a.d != 64 && a.d != 128)) return false;
  for (int s = 0; s < 2; ++s) {
    if (!p.qkv[s]) continue;
    if (p.L[s] > TILE || !aligned16(p.qkv[s]) || (p.ld[s] % 8) != 0) return false;
    if (p.ctx[s] && (!aligned16(p.ctx[s]) || (p.ld_ctx[s] % 8) != 0)) return false;
  }
  return true;
}

static int launch(const vb_attn_args& a, const P& p, cudaStream_t st) {
  const int Lmax = std::max(p.L[0], p.L[1]);
  const int G = std::max(1, std::min(MAXG, TILE / Lmax));
  const int qs_mask = ((p.pair_mask & 3) ? 1 : 0) | ((p.pair_mask & 12) ? 2 : 0);
  const int nq = (qs_mask == 3) ? 2 : 1;
  int nks = 1;
  for (int qs = 0; qs < 2; ++qs)
    if ((qs_mask >> qs) & 1) nks = std::max(nks, (pair_on(p.pair_mask, qs, 0) ? 1 : 0) + (pair_on(p.pair_mask, qs, 1) ? 1 : 0));
  CUtensorMap maps[2];
  for (int s = 0; s < 2; ++s) {
    const int o = p.qkv[s] ? s : 1 - s;   // an absent stream gets a valid (unused) descriptor
    int rc = vb::tc::make_map(&maps[s], p.qkv[o], (int64_t)p.B * p.L[o], 3 * (int64_t)p.A, p.ld[o], TILE);
    if (rc) return rc;
  }
  const dim3 grid((unsigned)(((p.B + G - 1) / G) * p.heads), (unsigned)nq);
#define VB_ATC(D_, N_)                                                                                         \
  do {                                                                                                         \
    const size_t sm = smem_bytes<D_, N_>();                                                                    \
    cudaError_t e = cudaFuncSetAttribute(attn_tc_fwd_kernel<D_, N_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
    VB_REQUIRE(e == cudaSuccess, (int)e, "attention (tcgen05): cannot reserve %zu bytes of shared memory", sm); \
    attn_tc_fwd_kernel<D_, N_><<<grid, 128, sm, st>>>(maps[0], maps[1], p, G, qs_mask);                        \
  } while (0)
  if (a.d == 64) { if (nks == 1) VB_ATC(64, 1); else VB_ATC(64, 2); }
  else { if (nks == 1) VB_ATC(128, 1); else VB_ATC(128, 2); }
#undef VB_ATC
  return check_launch("attention attn_tc_fwd_kernel");
}
}  // namespace tcf

// ---- host ----------------------------------------------------------------------------
static int fill(const vb_attn_args& a, P& p, bool bwd, bool need_ctx = true) {
  VB_REQUIRE(a.B > 0 && a.heads > 0 && a.d > 0, VB_ERR_ARG, "attention: bad sizes B=%d heads=%d d=%d", a.B, a.heads, a.d);
  VB_REQUIRE(a.qkv_t || a.qkv_v, VB_ERR_ARG, "attention: no stream given");
  VB_REQUIRE((a.pair_mask & ~15) == 0 && a.pair_mask != 0, VB_ERR_ARG, "attention: bad pair_mask %d", a.pair_mask);
  const bool tq = a.pair_mask & 3, vq = a.pair_mask & 12;         // stream has queries
  const bool tk = a.pair_mask & 5, vk = a.pair_mask & 10;         // stream has keys
  VB_REQUIRE(!(tq || tk) || (a.qkv_t && a.mask_t && a.Lt > 0), VB_ERR_ARG, "attention: text stream used but not provided");
  VB_REQUIRE(!(vq || vk) || (a.qkv_v && a.mask_v && a.Lv > 0), VB_ERR_ARG, "attention: vision stream used but not provided");
  VB_REQUIRE(!need_ctx || !tq || (a.ctx_t && a.lse_t), VB_ERR_ARG, "attention: ctx_t/lse_t missing");
  VB_REQUIRE(!need_ctx || !vq || (a.ctx_v && a.lse_v), VB_ERR_ARG, "attention: ctx_v/lse_v missing");
  if (bwd) {
    VB_REQUIRE(!tq || a.dctx_t, VB_ERR_ARG, "attention_bwd: dctx_t missing");
    VB_REQUIRE(!vq || a.dctx_v, VB_ERR_ARG, "attention_bwd: dctx_v missing");
    VB_REQUIRE(!a.qkv_t || a.dqkv_t, VB_ERR_ARG, "attention_bwd: dqkv_t missing");
    VB_REQUIRE(!a.qkv_v || a.dqkv_v, VB_ERR_ARG, "attention_bwd: dqkv_v missing");
  }
  VB_REQUIRE(a.dropout_p_t >= 0.f && a.dropout_p_t < 1.f && a.dropout_p_v >= 0.f && a.dropout_p_v < 1.f, VB_ERR_ARG,
             "attention: dropout out of range");
  p.qkv[0] = (tq || tk) ? a.qkv_t : nullptr; p.qkv[1] = (vq || vk) ? a.qkv_v : nullptr;  // a stream no enabled pair touches is absent
  p.ld[0] = a.ld_t; p.ld[1] = a.ld_v;
  p.mask[0] = a.mask_t; p.mask[1] = a.mask_v;
  p.ctx[0] = tq ? a.ctx_t : nullptr; p.ctx[1] = vq ? a.ctx_v : nullptr;
  p.ld_ctx[0] = a.ld_ctx_t; p.ld_ctx[1] = a.ld_ctx_v;
  p.lse[0] = a.lse_t; p.lse[1] = a.lse_v;
  p.dctx[0] = tq ? a.dctx_t : nullptr; p.dctx[1] = vq ? a.dctx_v : nullptr;
  p.dqkv[0] = p.qkv[0] ? a.dqkv_t : nullptr; p.dqkv[1] = p.qkv[1] ? a.dqkv_v : nullptr;
  p.B = a.B; p.L[0] = p.qkv[0] ? a.Lt : 0; p.L[1] = p.qkv[1] ? a.Lv : 0;
  p.ngroups = attn_groups(a.pair_mask);
  p.heads = a.heads; p.d = a.d; p.A = a.heads * a.d;
  p.pair_mask = a.pair_mask;
  p.scale = a.scale;
  p.thresh[0] = a.dropout_p_t > 0.f ? dropout_thresh(a.dropout_p_t) : 0u;
  p.thresh[1] = a.dropout_p_v > 0.f ? dropout_thresh(a.dropout_p_v) : 0u;
  p.keep_scale[0] = 1.f / (1.f - a.dropout_p_t);
  p.keep_scale[1] = 1.f / (1.f - a.dropout_p_v);
  p.seed = a.seed + a.offset;
  p.seed_ptr = seed_pointer();
  p.len[0] = p.qkv[0] ? a.len_t : nullptr;
  p.len[1] = p.qkv[1] ? a.len_v : nullptr;
  p.cu[0] = p.qkv[0] ? a.cu_t : nullptr;
  p.cu[1] = p.qkv[1] ? a.cu_v : nullptr;
  VB_REQUIRE((!p.cu[0] || p.len[0]) && (!p.cu[1] || p.len[1]), VB_ERR_ARG, "attention: packed rows (cu_*) need the lengths (len_*)");
  return 0;
}

struct LaunchGeo { int warps_fwd, warps_bwd, nkb; size_t smem_fwd, smem_bwd; };
static LaunchGeo launch_geo(const P& p, int d) {
  LaunchGeo lg = {0, 0, 0, 0, 0};
  const size_t pitch = (size_t)(d + 8) * 2;
  for (int grp = 0; grp < p.ngroups; ++grp) {
    const Geo g = make_geo(p.pair_mask, p.L[0], p.L[1], grp);
    lg.warps_fwd = std::max(lg.warps_fwd, g.RQ / 16);
    lg.warps_bwd = std::max(lg.warps_bwd, std::max(g.RQ, g.RK) / 16);
    lg.nkb = std::max(lg.nkb, g.RK / 16);
    lg.smem_fwd = std::max(lg.smem_fwd, (size_t)(g.RQ + 2 * g.RK) * pitch + (size_t)g.RK * 4);
    lg.smem_bwd = std::max(lg.smem_bwd, (size_t)(2 * g.RQ + 2 * g.RK) * pitch + (size_t)(g.RK + 3 * g.RQ) * 4 +
                                            std::max((size_t)g.RK * (g.RQ + 8) * 2, (size_t)g.RQ * pitch));
  }
  return lg;
}

static bool mma_ok(const vb_attn_args& a, const P& p, bool bwd, const char** why) {
  auto fail = [&](const char* w) { if (why) *why = w; return false; };
  if (a.dtype != VB_BF16) return fail("dtype is not bf16");
  if (a.d != 64 && a.d != 128) return fail("head size is not 64 or 128");
  const LaunchGeo lg = launch_geo(p, a.d);
  if (lg.nkb > 10 || lg.warps_bwd > 10) return fail("more than 160 padded query or key rows per (sample, head)");
  if ((bwd ? lg.smem_bwd : lg.smem_fwd) > 227 * 1024) return fail("tiles exceed 227 KB of shared memory");
  for (int s = 0; s < 2; ++s) {
    if (!p.qkv[s]) continue;
    if (!aligned16(p.qkv[s]) || (p.ld[s] % 8) != 0) return fail("qkv not 16-byte aligned");
    if (p.ctx[s] && (!aligned16(p.ctx[s]) || (p.ld_ctx[s] % 8) != 0)) return fail("ctx not 16-byte aligned");
    if (p.dctx[s] && !aligned16(p.dctx[s])) return fail("dctx not 16-byte aligned");
    if (p.dqkv[s] && !aligned16(p.dqkv[s])) return fail("dqkv not 16-byte aligned");
  }
  return true;
}

template <typename K> static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    VB_REQUIRE(e == cudaSuccess, (int)e, "attention: cannot reserve %zu bytes of shared memory: %s", bytes, cudaGetErrorString(e));
  }
  return 0;
}

static int run(const vb_attn_args& a, cudaStream_t st, bool bwd) {
  P p;
  int rc = fill(a, p, bwd);
  if (rc) return rc;
  if (bwd) {  // a stream no enabled pair touches receives no gradient
    const size_t es = a.dtype == VB_BF16 ? 2 : 4, width = (size_t)3 * a.heads * a.d * es;
    if (a.dqkv_t && !p.qkv[0]) cudaMemset2DAsync(a.dqkv_t, (size_t)a.ld_t * es, 0, width, (size_t)a.B * a.Lt, st);
    if (a.dqkv_v && !p.qkv[1]) cudaMemset2DAsync(a.dqkv_v, (size_t)a.ld_v * es, 0, width, (size_t)a.B * a.Lv, st);
  }
  const char* why = nullptr;
  const bool can_mma = mma_ok(a, p, bwd, &why);
  VB_REQUIRE(a.backend != VB_ATTN_MMA || can_mma, VB_ERR_UNSUPPORTED, "attention: MMA back-end cannot run: %s", why);
  const bool use_mma = can_mma && a.backend != VB_ATTN_SIMT;
  VB_REQUIRE(use_mma || (!p.cu[0] && !p.cu[1]), VB_ERR_UNSUPPORTED, "attention: packed rows (cu_*) need the tensor-core back-end");
  const unsigned grid = (unsigned)(a.B * a.heads);
  // forward on the tcgen05 / TMEM / TMA kernel (VB_ATTN_TC, or AUTO unless VOLTA_B200_ATTN_TC=0); the backward keeps
  // the warp-level kernel, which recomputes P from the same lse / dropout hash
  // Measured (B200, isolated launches): at 30 x 101 rows and d = 128 (vilbert_base retrieval) it is the faster one; at
  // the 36/37-row pre-training shapes three samples share a tile, each thread's softmax walks columns of other samples,
  // and the warp-level kernel (scores kept in mma fragments) stays ahead - AUTO therefore takes it only when one sample
  // fills more than half a tile.  VOLTA_B200_ATTN_TC=1 forces it wherever it can run, =0 never.
  if (!bwd && a.backend != VB_ATTN_SIMT && a.backend != VB_ATTN_MMA && tcf::ok(a, p)) {
    const char* e = getenv("VOLTA_B200_ATTN_TC");
    const bool force = a.backend == VB_ATTN_TC || (e && e[0] == '1');
    const bool never = a.backend != VB_ATTN_TC && e && e[0] == '0';
    if (force || (!never && std::max(p.L[0], p.L[1]) > tcf::TILE / 2)) return tcf::launch(a, p, st);
  }
  VB_REQUIRE(a.backend != VB_ATTN_TC || bwd, VB_ERR_UNSUPPORTED, "attention: tcgen05 back-end cannot run this call (bf16, d 64/128, <= 128 rows per sample)");
  if (use_mma) {
    const LaunchGeo lg = launch_geo(p, a.d);
    const dim3 g2(grid, (unsigned)p.ngroups);
    if (!bwd) {
#define VB_FWD(D_, KB_)                                                              \
  do {                                                                               \
    if ((rc = set_smem(mma_fwd_kernel<D_, KB_>, lg.smem_fwd))) return rc;            \
    launch_pdl(mma_fwd_kernel<D_, KB_>, g2, dim3(lg.warps_fwd * 32), lg.smem_fwd, st, p); \
  } while (0)
      if (a.d == 64) { if (lg.nkb <= 3) VB_FWD(64, 3); else if (lg.nkb <= 6) VB_FWD(64, 6); else VB_FWD(64, 10); }
      else { if (lg.nkb <= 3) VB_FWD(128, 3); else if (lg.nkb <= 6) VB_FWD(128, 6); else VB_FWD(128, 10); }
#undef VB_FWD
      return check_launch("attention mma_fwd_kernel");
    }
#define VB_BWD(D_, S_)                                                                       \
  do {                                                                                       \
    if ((rc = set_smem(mma_bwd_kernel<D_, S_>, lg.smem_bwd))) return rc;                     \
    launch_pdl(mma_bwd_kernel<D_, S_>, g2, dim3(lg.warps_bwd * 32), lg.smem_bwd, st, p);     \
  } while (0)
    if (a.d == 64) { if (lg.warps_bwd <= 3) VB_BWD(64, true); else VB_BWD(64, false); }
    else VB_BWD(128, false);  // the 128-wide head spills heavily under the 6-blocks-per-SM register cap
#undef VB_BWD
    return check_launch("attention mma_bwd_kernel");
  }
  VB_REQUIRE(a.d <= 256, VB_ERR_UNSUPPORTED, "attention (SIMT): head size %d > 256", a.d);
  const int Ltot = p.L[0] + p.L[1];
  dim3 g2(grid, 2);
  if (!bwd) {
    const size_t smem = 4 * (size_t)(a.d + Ltot) * sizeof(float);
    if (a.dtype == VB_BF16) simt_fwd_kernel<bf16><<<g2, 128, smem, st>>>(p);
    else simt_fwd_kernel<float><<<g2, 128, smem, st>>>(p);
    return check_launch("attention simt_fwd_kernel");
  }
  const size_t smem = 4 * (size_t)(2 * a.d + Ltot) * sizeof(float);
  if (a.dtype == VB_BF16) {
    simt_bwd_dq_kernel<bf16><<<g2, 128, smem, st>>>(p);
    simt_bwd_dkv_kernel<bf16><<<g2, 128, 0, st>>>(p);
  } else {
    simt_bwd_dq_kernel<float><<<g2, 128, smem, st>>>(p);
    simt_bwd_dkv_kernel<float><<<g2, 128, 0, st>>>(p);
  }
  return check_launch("attention simt_bwd kernels");
}

}  // namespace attn
}  // namespace vb

namespace vb {
namespace attn {
// len[b] = 1 + index of the last position of sample b whose mask is set (0 if none): one warp per sample
__global__ void __launch_bounds__(256)
mask_len_kernel(const int32_t* __restrict__ mask, int B, int L, int32_t* __restrict__ len) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + warp;
  if (b >= B) return;
  int last = 0;
  for (int j = lane; j < L; j += 32)
    if (mask[(int64_t)b * L + j] != 0) last = j + 1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  if (lane == 0) len[b] = last;
}
}  // namespace attn
}  // namespace vb

extern "C" {
int vb_mask_lengths(const int32_t* mask, int32_t B, int32_t L, int32_t* len, void* stream) {
  VB_REQUIRE(mask && len && B > 0 && L > 0, VB_ERR_ARG, "mask_lengths: bad arguments");
  vb::attn::mask_len_kernel<<<(unsigned)((B + 7) / 8), 256, 0, (cudaStream_t)stream>>>(mask, B, L, len);
  return vb::check_launch("mask_len_kernel");
}
int vb_attention_fwd(const vb_attn_args* a, void* stream) {
  VB_REQUIRE(a != nullptr, VB_ERR_ARG, "attention_fwd: null args");
  return vb::attn::run(*a, (cudaStream_t)stream, false);
}
int vb_attention_probs(const vb_attn_args* a, float* probs_t, float* probs_v, void* stream) {
  VB_REQUIRE(a != nullptr, VB_ERR_ARG, "attention_probs: null args");
  VB_REQUIRE(a->cu_t == nullptr && a->cu_v == nullptr, VB_ERR_UNSUPPORTED, "attention_probs: padded layout only (no cu_*)");
  vb::attn::P p;
  int rc = vb::attn::fill(*a, p, false, /*need_ctx=*/false);   // this entry point writes neither ctx nor lse
  if (rc) return rc;
  VB_REQUIRE(a->d <= 256, VB_ERR_UNSUPPORTED, "attention_probs: head size %d > 256", a->d);
  VB_REQUIRE((!(a->pair_mask & 3) || probs_t) && (!(a->pair_mask & 12) || probs_v), VB_ERR_ARG, "attention_probs: output missing for a query stream");
  const int Ltot = p.L[0] + p.L[1];
  const size_t smem = 4 * (size_t)(a->d + Ltot) * sizeof(float);
  const dim3 grid((unsigned)(a->B * a->heads), 2);
  if (a->dtype == VB_BF16) vb::attn::simt_probs_kernel<vb::bf16><<<grid, 128, smem, (cudaStream_t)stream>>>(p, probs_t, probs_v);
  else vb::attn::simt_probs_kernel<float><<<grid, 128, smem, (cudaStream_t)stream>>>(p, probs_t, probs_v);
  return vb::check_launch("attention simt_probs_kernel");
}
int vb_attention_bwd(const vb_attn_args* a, void* stream) {
  VB_REQUIRE(a != nullptr, VB_ERR_ARG, "attention_bwd: null args");
  return vb::attn::run(*a, (cudaStream_t)stream, true);
}
}
