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
#include "common.cuh"

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
  uint64_t seed, offset;
};

__device__ __forceinline__ bool pair_on(int pair_mask, int qs, int ks) { return (pair_mask >> (qs * 2 + ks)) & 1; }

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
    const uint64_t base = (((uint64_t)b * p.heads + h) * Ltot + qg) * (uint64_t)Ltot;
    for (int kg = lane; kg < Ltot; kg += 32) {
      float pr = __expf(pbuf[kg] - lse);
      if (p.thresh[qs]) pr = dropout_keep(p.seed, p.offset, base + kg, p.thresh[qs]) ? pr * p.keep_scale[qs] : 0.f;
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
    const uint64_t base = (((uint64_t)b * p.heads + h) * Ltot + qg) * (uint64_t)Ltot;
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
        if (p.thresh[qs]) dp = dropout_keep(p.seed, p.offset, base + kg, p.thresh[qs]) ? dp * p.keep_scale[qs] : 0.f;
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
    const int kg = ks ? p.L[0] + j : j;
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
          const uint64_t idx = (((uint64_t)b * p.heads + h) * Ltot + qg) * (uint64_t)Ltot + kg;
          const bool keep = dropout_keep(p.seed, p.offset, idx, p.thresh[qs]);
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
  // B operand with B[k][n] = M[k][n] (M row-major [k][D]): 16 k x 16 n block at (k0, n0), transposing load
  static __device__ __forceinline__ void load_bt(const bf16* base, int k0, int n0, uint32_t (&r)[4]) {
    const int lane = threadIdx.x & 31;
    ldsm_x4_t(smem_u32(base + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + n0 + (lane >> 4) * 8), r);
  }
};

struct Geo {  // padded geometry of one (sample, head)
  int Lp[2];  // rows padded to 16 per stream (0 if the stream is absent)
  int R;      // Lp[0] + Lp[1]
};
__device__ __forceinline__ Geo make_geo(const P& p) {
  Geo g;
  g.Lp[0] = p.qkv[0] ? ((p.L[0] + 15) & ~15) : 0;
  g.Lp[1] = p.qkv[1] ? ((p.L[1] + 15) & ~15) : 0;
  g.R = g.Lp[0] + g.Lp[1];
  return g;
}

// Cooperative ASYNCHRONOUS load of one operand (column block: Q / K / V of the packed qkv, or a ctx-layout tensor)
// into its padded shared-memory tile with cp.async (LDGSTS): all 16-byte copies of all operands are in flight
// at once and rows beyond the sequence are zero-filled (src-size 0).  A register-staged loop (LDG -> STS per
// iteration) serialised ~18 global round trips per thread and dominated the kernel.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int D>
__device__ __forceinline__ void load_rows(bf16* dst, const P& p, const Geo& g, int b, int h, const void* const src[2],
                                          const int64_t ld[2], int col_off) {
  constexpr int CH = D / 8;
  const void* any = src[0] ? src[0] : src[1];
  for (int e = threadIdx.x; e < g.R * CH; e += blockDim.x) {
    const int r = e / CH, c = (e % CH) * 8;
    const int s = r < g.Lp[0] ? 0 : 1;
    const int j = s ? r - g.Lp[0] : r;
    const bool valid = j < p.L[s] && src[s] != nullptr;
    const void* from = valid ? (const void*)(reinterpret_cast<const bf16*>(src[s]) + ((int64_t)b * p.L[s] + j) * ld[s] + col_off + h * D + c)
                             : any;
    cp_async16(smem_u32(dst + r * Tile<D>::PITCH + c), from, valid);
  }
}

template <int D, int MAXKB>
__global__ void __launch_bounds__(128, (D == 64 && MAXKB <= 6) ? 4 : 2)
mma_fwd_kernel(P p) {
  extern __shared__ __align__(16) uint8_t smraw[];
  using TL = Tile<D>;
  const Geo g = make_geo(p);
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Ks = Qs + g.R * TL::PITCH;
  bf16* Vs = Ks + g.R * TL::PITCH;
  float* madd = reinterpret_cast<float*>(Vs + g.R * TL::PITCH);
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int Ltot = p.L[0] + p.L[1];

  load_rows<D>(Qs, p, g, b, h, p.qkv, p.ld, 0);
  load_rows<D>(Ks, p, g, b, h, p.qkv, p.ld, p.A);
  load_rows<D>(Vs, p, g, b, h, p.qkv, p.ld, 2 * p.A);
  for (int r = threadIdx.x; r < g.R; r += blockDim.x) {
    const int s = r < g.Lp[0] ? 0 : 1;
    const int j = s ? r - g.Lp[0] : r;
    madd[r] = j < p.L[s] ? (p.mask[s][(int64_t)b * p.L[s] + j] ? 0.f : -10000.0f) : -INFINITY;
  }
  cp_async_wait_all();
  __syncthreads();

  const int nkb = g.R / 16;
  for (int qt = warp; qt < nkb; qt += 4) {
    const int qs = (qt * 16 < g.Lp[0]) ? 0 : 1;
    if (p.ctx[qs] == nullptr) continue;
    float s[2 * MAXKB][4];
#pragma unroll
    for (int i = 0; i < 2 * MAXKB; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
      uint32_t a[4];
      TL::load_a(Qs, qt * 16, ks * 16, a);
#pragma unroll
      for (int kb = 0; kb < MAXKB; ++kb) {
        if (kb < nkb && pair_on(p.pair_mask, qs, (kb * 16 < g.Lp[0]) ? 0 : 1)) {
          uint32_t bb[4];
          TL::load_b(Ks, kb * 16, ks * 16, bb);
          mma16816(s[2 * kb], a, bb[0], bb[1]);
          mma16816(s[2 * kb + 1], a, bb[2], bb[3]);
        }
      }
    }
    // scale + masks, row max (rows gq and gq+8)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int kb = 0; kb < MAXKB; ++kb) {
      const bool on = kb < nkb && pair_on(p.pair_mask, qs, (kb * 16 < g.Lp[0]) ? 0 : 1);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = -INFINITY;
          if (on) v = s[2 * kb + nt][e] * p.scale + madd[kb * 16 + nt * 8 + 2 * tq + (e & 1)];
          s[2 * kb + nt][e] = v;
          mx[e >> 1] = fmaxf(mx[e >> 1], v);
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
    for (int i = 0; i < 2 * MAXKB; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float ex = __expf(s[i][e] - mx[e >> 1]);
        s[i][e] = ex;
        sum[e >> 1] += ex;
      }
    float inv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
      inv[r] = 1.f / sum[r];
      const int j = qt * 16 - (qs ? g.Lp[0] : 0) + gq + r * 8;
      if (tq == 0 && j < p.L[qs]) p.lse[qs][((int64_t)b * p.heads + h) * p.L[qs] + j] = mx[r] + __logf(sum[r]);
    }
    const uint32_t thresh = p.thresh[qs];
    const float ksc = p.keep_scale[qs];
#pragma unroll
    for (int kb = 0; kb < MAXKB; ++kb)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float pr = s[2 * kb + nt][e] * inv[e >> 1];
          if (thresh) {
            const int kr = kb * 16 + nt * 8 + 2 * tq + (e & 1);
            const int kstream = kr < g.Lp[0] ? 0 : 1;
            const int kg = kstream ? p.L[0] + (kr - g.Lp[0]) : kr;
            const int jq = qt * 16 - (qs ? g.Lp[0] : 0) + gq + (e >> 1) * 8;
            const int qg = qs ? p.L[0] + jq : jq;
            const uint64_t idx = (((uint64_t)b * p.heads + h) * Ltot + qg) * (uint64_t)Ltot + kg;
            if (pr != 0.f) pr = dropout_keep(p.seed, p.offset, idx, thresh) ? pr * ksc : 0.f;
          }
          s[2 * kb + nt][e] = pr;
        }
    // O = P . V
    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
#pragma unroll
    for (int kb = 0; kb < MAXKB; ++kb) {
      if (kb < nkb && pair_on(p.pair_mask, qs, (kb * 16 < g.Lp[0]) ? 0 : 1)) {
        uint32_t a[4];
        a[0] = pack_bf16x2(s[2 * kb][0], s[2 * kb][1]);
        a[1] = pack_bf16x2(s[2 * kb][2], s[2 * kb][3]);
        a[2] = pack_bf16x2(s[2 * kb + 1][0], s[2 * kb + 1][1]);
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
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int j = qt * 16 - (qs ? g.Lp[0] : 0) + gq + r * 8;
      if (j < p.L[qs]) {
        bf16* O = reinterpret_cast<bf16*>(p.ctx[qs]) + ((int64_t)b * p.L[qs] + j) * p.ld_ctx[qs] + h * D;
#pragma unroll
        for (int nd = 0; nd < D / 8; ++nd)
          *reinterpret_cast<uint32_t*>(O + nd * 8 + 2 * tq) = pack_bf16x2(o[nd][r * 2], o[nd][r * 2 + 1]);
      }
    }
  }
}

template <int D>
__global__ void __launch_bounds__(128, D == 64 ? 4 : 2)
mma_bwd_kernel(P p) {
  extern __shared__ __align__(16) uint8_t smraw[];
  using TL = Tile<D>;
  const Geo g = make_geo(p);
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Ks = Qs + g.R * TL::PITCH;
  bf16* Vs = Ks + g.R * TL::PITCH;
  bf16* dOs = Vs + g.R * TL::PITCH;
  bf16* Os = dOs + g.R * TL::PITCH;
  float* madd = reinterpret_cast<float*>(Os + g.R * TL::PITCH);
  float* lse_s = madd + g.R;
  float* delta_s = lse_s + g.R;
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int Ltot = p.L[0] + p.L[1];

  load_rows<D>(Qs, p, g, b, h, p.qkv, p.ld, 0);
  load_rows<D>(Ks, p, g, b, h, p.qkv, p.ld, p.A);
  load_rows<D>(Vs, p, g, b, h, p.qkv, p.ld, 2 * p.A);
  load_rows<D>(dOs, p, g, b, h, p.dctx, p.ld_ctx, 0);
  {
    const void* octx[2] = {p.ctx[0], p.ctx[1]};
    load_rows<D>(Os, p, g, b, h, octx, p.ld_ctx, 0);
  }
  for (int r = threadIdx.x; r < g.R; r += blockDim.x) {
    const int s = r < g.Lp[0] ? 0 : 1;
    const int j = s ? r - g.Lp[0] : r;
    const bool valid = j < p.L[s];
    madd[r] = valid ? (p.mask[s][(int64_t)b * p.L[s] + j] ? 0.f : -10000.0f) : -INFINITY;
    lse_s[r] = (valid && p.lse[s]) ? p.lse[s][((int64_t)b * p.heads + h) * p.L[s] + j] : INFINITY;  // pad rows: P = 0
    delta_s[r] = 0.f;
  }
  cp_async_wait_all();
  __syncthreads();
  {  // delta[r] = sum_d dO[r,d] * O[r,d]  (both tiles are in shared memory; padded rows are zero)
    constexpr int CH = D / 8;
    for (int e = threadIdx.x; e < g.R * CH; e += blockDim.x) {
      const int r = e / CH, c = (e % CH) * 8;
      float ov[8], dv[8];
      Vec8<bf16>::load(Os + r * TL::PITCH + c, ov);
      Vec8<bf16>::load(dOs + r * TL::PITCH + c, dv);
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc = fmaf(ov[k], dv[k], acc);
      atomicAdd(delta_s + r, acc);
    }
  }
  __syncthreads();

  const int nb = g.R / 16;
  // ---------------- phase 1: dQ (warp per 16-query tile) ----------------
  for (int qt = warp; qt < nb; qt += 4) {
    const int qs = (qt * 16 < g.Lp[0]) ? 0 : 1;
    uint32_t aq[D / 16][4], ado[D / 16][4];
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
      TL::load_a(Qs, qt * 16, ks * 16, aq[ks]);
      TL::load_a(dOs, qt * 16, ks * 16, ado[ks]);
    }
    float dq[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }
    const uint32_t thresh = p.thresh[qs];
    const float ksc = p.keep_scale[qs];
    for (int kb = 0; kb < nb; ++kb) {
      if (!pair_on(p.pair_mask, qs, (kb * 16 < g.Lp[0]) ? 0 : 1)) continue;
      float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks) {
        uint32_t bk[4], bv[4];
        TL::load_b(Ks, kb * 16, ks * 16, bk);
        TL::load_b(Vs, kb * 16, ks * 16, bv);
        mma16816(s[0], aq[ks], bk[0], bk[1]);
        mma16816(s[1], aq[ks], bk[2], bk[3]);
        mma16816(dp[0], ado[ks], bv[0], bv[1]);
        mma16816(dp[1], ado[ks], bv[2], bv[3]);
      }
      float ds[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int kr = kb * 16 + nt * 8 + 2 * tq + (e & 1);
          const int qr = qt * 16 + gq + (e >> 1) * 8;
          const float pr = __expf(s[nt][e] * p.scale + madd[kr] - lse_s[qr]);
          float dpe = dp[nt][e];
          if (thresh && pr != 0.f) {
            const int kstream = kr < g.Lp[0] ? 0 : 1;
            const int kg = kstream ? p.L[0] + (kr - g.Lp[0]) : kr;
            const int jq = qr - (qs ? g.Lp[0] : 0);
            const int qg = qs ? p.L[0] + jq : jq;
            const uint64_t idx = (((uint64_t)b * p.heads + h) * Ltot + qg) * (uint64_t)Ltot + kg;
            dpe = dropout_keep(p.seed, p.offset, idx, thresh) ? dpe * ksc : 0.f;
          }
          ds[nt][e] = pr == 0.f ? 0.f : pr * (dpe - delta_s[qr]) * p.scale;
        }
      uint32_t a[4];
      a[0] = pack_bf16x2(ds[0][0], ds[0][1]);
      a[1] = pack_bf16x2(ds[0][2], ds[0][3]);
      a[2] = pack_bf16x2(ds[1][0], ds[1][1]);
      a[3] = pack_bf16x2(ds[1][2], ds[1][3]);
#pragma unroll
      for (int nd = 0; nd < D / 16; ++nd) {
        uint32_t bb[4];
        TL::load_bt(Ks, kb * 16, nd * 16, bb);
        mma16816(dq[2 * nd], a, bb[0], bb[1]);
        mma16816(dq[2 * nd + 1], a, bb[2], bb[3]);
      }
    }
    if (p.dqkv[qs] != nullptr) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int j = qt * 16 - (qs ? g.Lp[0] : 0) + gq + r * 8;
        if (j < p.L[qs]) {
          bf16* dQ = reinterpret_cast<bf16*>(p.dqkv[qs]) + ((int64_t)b * p.L[qs] + j) * p.ld[qs] + h * D;
#pragma unroll
          for (int nd = 0; nd < D / 8; ++nd)
            *reinterpret_cast<uint32_t*>(dQ + nd * 8 + 2 * tq) = pack_bf16x2(dq[nd][r * 2], dq[nd][r * 2 + 1]);
        }
      }
    }
  }
  // ---------------- phase 2: dK, dV (warp per 16-key tile; S^T = K Q^T) ----------------
  for (int kt = warp; kt < nb; kt += 4) {
    const int kst = (kt * 16 < g.Lp[0]) ? 0 : 1;
    float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    for (int qb = 0; qb < nb; ++qb) {
      const int qs = (qb * 16 < g.Lp[0]) ? 0 : 1;
      if (!pair_on(p.pair_mask, qs, kst) || p.dctx[qs] == nullptr) continue;
      const uint32_t thresh = p.thresh[qs];
      const float ksc = p.keep_scale[qs];
      float st[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dpt[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks) {
        uint32_t ak[4], av[4], bq[4], bdo[4];
        TL::load_a(Ks, kt * 16, ks * 16, ak);
        TL::load_a(Vs, kt * 16, ks * 16, av);
        TL::load_b(Qs, qb * 16, ks * 16, bq);
        TL::load_b(dOs, qb * 16, ks * 16, bdo);
        mma16816(st[0], ak, bq[0], bq[1]);
        mma16816(st[1], ak, bq[2], bq[3]);
        mma16816(dpt[0], av, bdo[0], bdo[1]);
        mma16816(dpt[1], av, bdo[2], bdo[3]);
      }
      float pd[2][4], ds[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int kr = kt * 16 + gq + (e >> 1) * 8;              // key row (tile row)
          const int qr = qb * 16 + nt * 8 + 2 * tq + (e & 1);      // query row (tile column)
          const float pr = __expf(st[nt][e] * p.scale + madd[kr] - lse_s[qr]);
          float pdv = pr, dpe = dpt[nt][e];
          if (thresh && pr != 0.f) {
            const int kg = kst ? p.L[0] + (kr - g.Lp[0]) : kr;
            const int jq = qr - (qs ? g.Lp[0] : 0);
            const int qg = qs ? p.L[0] + jq : jq;
            const uint64_t idx = (((uint64_t)b * p.heads + h) * Ltot + qg) * (uint64_t)Ltot + kg;
            const bool keep = dropout_keep(p.seed, p.offset, idx, thresh);
            pdv = keep ? pr * ksc : 0.f;
            dpe = keep ? dpe * ksc : 0.f;
          }
          pd[nt][e] = pdv;
          ds[nt][e] = pr == 0.f ? 0.f : pr * (dpe - delta_s[qr]) * p.scale;
        }
      uint32_t apd[4], ads[4];
      apd[0] = pack_bf16x2(pd[0][0], pd[0][1]); apd[1] = pack_bf16x2(pd[0][2], pd[0][3]);
      apd[2] = pack_bf16x2(pd[1][0], pd[1][1]); apd[3] = pack_bf16x2(pd[1][2], pd[1][3]);
      ads[0] = pack_bf16x2(ds[0][0], ds[0][1]); ads[1] = pack_bf16x2(ds[0][2], ds[0][3]);
      ads[2] = pack_bf16x2(ds[1][0], ds[1][1]); ads[3] = pack_bf16x2(ds[1][2], ds[1][3]);
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
    if (p.dqkv[kst] != nullptr) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int j = kt * 16 - (kst ? g.Lp[0] : 0) + gq + r * 8;
        if (j < p.L[kst]) {
          bf16* dK = reinterpret_cast<bf16*>(p.dqkv[kst]) + ((int64_t)b * p.L[kst] + j) * p.ld[kst] + p.A + h * D;
          bf16* dV = dK + p.A;
#pragma unroll
          for (int nd = 0; nd < D / 8; ++nd) {
            *reinterpret_cast<uint32_t*>(dK + nd * 8 + 2 * tq) = pack_bf16x2(dk[nd][r * 2], dk[nd][r * 2 + 1]);
            *reinterpret_cast<uint32_t*>(dV + nd * 8 + 2 * tq) = pack_bf16x2(dv[nd][r * 2], dv[nd][r * 2 + 1]);
          }
        }
      }
    }
  }
}

// ---- host ----------------------------------------------------------------------------
static int fill(const vb_attn_args& a, P& p, bool bwd) {
  VB_REQUIRE(a.B > 0 && a.heads > 0 && a.d > 0, VB_ERR_ARG, "attention: bad sizes B=%d heads=%d d=%d", a.B, a.heads, a.d);
  VB_REQUIRE(a.qkv_t || a.qkv_v, VB_ERR_ARG, "attention: no stream given");
  VB_REQUIRE((a.pair_mask & ~15) == 0 && a.pair_mask != 0, VB_ERR_ARG, "attention: bad pair_mask %d", a.pair_mask);
  const bool tq = a.pair_mask & 3, vq = a.pair_mask & 12;         // stream has queries
  const bool tk = a.pair_mask & 5, vk = a.pair_mask & 10;         // stream has keys
  VB_REQUIRE(!(tq || tk) || (a.qkv_t && a.mask_t && a.Lt > 0), VB_ERR_ARG, "attention: text stream used but not provided");
  VB_REQUIRE(!(vq || vk) || (a.qkv_v && a.mask_v && a.Lv > 0), VB_ERR_ARG, "attention: vision stream used but not provided");
  VB_REQUIRE(!tq || (a.ctx_t && a.lse_t), VB_ERR_ARG, "attention: ctx_t/lse_t missing");
  VB_REQUIRE(!vq || (a.ctx_v && a.lse_v), VB_ERR_ARG, "attention: ctx_v/lse_v missing");
  if (bwd) {
    VB_REQUIRE(!tq || a.dctx_t, VB_ERR_ARG, "attention_bwd: dctx_t missing");
    VB_REQUIRE(!vq || a.dctx_v, VB_ERR_ARG, "attention_bwd: dctx_v missing");
    VB_REQUIRE(!a.qkv_t || a.dqkv_t, VB_ERR_ARG, "attention_bwd: dqkv_t missing");
    VB_REQUIRE(!a.qkv_v || a.dqkv_v, VB_ERR_ARG, "attention_bwd: dqkv_v missing");
  }
  VB_REQUIRE(a.dropout_p_t >= 0.f && a.dropout_p_t < 1.f && a.dropout_p_v >= 0.f && a.dropout_p_v < 1.f, VB_ERR_ARG,
             "attention: dropout out of range");
  p.qkv[0] = a.qkv_t; p.qkv[1] = a.qkv_v;
  p.ld[0] = a.ld_t; p.ld[1] = a.ld_v;
  p.mask[0] = a.mask_t; p.mask[1] = a.mask_v;
  p.ctx[0] = tq ? a.ctx_t : nullptr; p.ctx[1] = vq ? a.ctx_v : nullptr;
  p.ld_ctx[0] = a.ld_ctx_t; p.ld_ctx[1] = a.ld_ctx_v;
  p.lse[0] = a.lse_t; p.lse[1] = a.lse_v;
  p.dctx[0] = tq ? a.dctx_t : nullptr; p.dctx[1] = vq ? a.dctx_v : nullptr;
  p.dqkv[0] = a.dqkv_t; p.dqkv[1] = a.dqkv_v;
  p.B = a.B; p.L[0] = a.qkv_t ? a.Lt : 0; p.L[1] = a.qkv_v ? a.Lv : 0;
  p.heads = a.heads; p.d = a.d; p.A = a.heads * a.d;
  p.pair_mask = a.pair_mask;
  p.scale = a.scale;
  p.thresh[0] = a.dropout_p_t > 0.f ? dropout_thresh(a.dropout_p_t) : 0u;
  p.thresh[1] = a.dropout_p_v > 0.f ? dropout_thresh(a.dropout_p_v) : 0u;
  p.keep_scale[0] = 1.f / (1.f - a.dropout_p_t);
  p.keep_scale[1] = 1.f / (1.f - a.dropout_p_v);
  p.seed = a.seed; p.offset = a.offset;
  return 0;
}

static bool mma_ok(const vb_attn_args& a, const P& p, const char** why) {
  auto fail = [&](const char* w) { if (why) *why = w; return false; };
  if (a.dtype != VB_BF16) return fail("dtype is not bf16");
  if (a.d != 64 && a.d != 128) return fail("head size is not 64 or 128");
  const int R = (p.qkv[0] ? ((p.L[0] + 15) & ~15) : 0) + (p.qkv[1] ? ((p.L[1] + 15) & ~15) : 0);
  if (R > 160) return fail("more than 160 padded rows per (sample, head)");
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
  const char* why = nullptr;
  const bool can_mma = mma_ok(a, p, &why);
  VB_REQUIRE(a.backend != VB_ATTN_MMA || can_mma, VB_ERR_UNSUPPORTED, "attention: MMA back-end cannot run: %s", why);
  const bool use_mma = can_mma && a.backend != VB_ATTN_SIMT;
  const unsigned grid = (unsigned)(a.B * a.heads);
  if (use_mma) {
    const int R = (p.qkv[0] ? ((p.L[0] + 15) & ~15) : 0) + (p.qkv[1] ? ((p.L[1] + 15) & ~15) : 0);
    const size_t tile = (size_t)R * (a.d + 8) * 2;
    if (!bwd) {
      const size_t smem = 3 * tile + R * sizeof(float);
#define VB_FWD(D_, KB_)                                                   \
  do {                                                                    \
    if ((rc = set_smem(mma_fwd_kernel<D_, KB_>, smem))) return rc;        \
    mma_fwd_kernel<D_, KB_><<<grid, 128, smem, st>>>(p);                  \
  } while (0)
      if (a.d == 64) { if (R <= 96) VB_FWD(64, 6); else VB_FWD(64, 10); }
      else { if (R <= 96) VB_FWD(128, 6); else VB_FWD(128, 10); }
#undef VB_FWD
      return check_launch("attention mma_fwd_kernel");
    }
    const size_t smem = 5 * tile + 3 * R * sizeof(float);
    if (a.d == 64) {
      if ((rc = set_smem(mma_bwd_kernel<64>, smem))) return rc;
      mma_bwd_kernel<64><<<grid, 128, smem, st>>>(p);
    } else {
      if ((rc = set_smem(mma_bwd_kernel<128>, smem))) return rc;
      mma_bwd_kernel<128><<<grid, 128, smem, st>>>(p);
    }
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

extern "C" {
int vb_attention_fwd(const vb_attn_args* a, void* stream) {
  VB_REQUIRE(a != nullptr, VB_ERR_ARG, "attention_fwd: null args");
  return vb::attn::run(*a, (cudaStream_t)stream, false);
}
int vb_attention_bwd(const vb_attn_args* a, void* stream) {
  VB_REQUIRE(a != nullptr, VB_ERR_ARG, "attention_bwd: null args");
  return vb::attn::run(*a, (cudaStream_t)stream, true);
}
}
