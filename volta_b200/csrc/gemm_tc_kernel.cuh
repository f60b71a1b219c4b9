// tcgen05 / TMEM / TMA GEMM for sm_100a:  D[M,N] = epilogue( A . B^T ), bf16 operands, fp32 accumulation.
//
// One persistent, warp-specialised kernel serves nn.Linear forward, dgrad and wgrad
// (volta/encoders.py:204-206,376,463,518 ... and their autograd):
//   warp 0   : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier expect_tx)
//   warp 1   : MMA issuer     (one lane issues tcgen05.mma 128 x BN x 16, accumulators in TMEM)
//   warp 2   : TMEM allocator
//   warps 4-11: epilogue      (tcgen05.ld 32x32b -> bias/GELU/dropout/residual/... -> smem staging -> global)
// The TMEM accumulator is double buffered (2 x BN columns) so the epilogue of tile i overlaps the
// MMAs of tile i+1.  Eight epilogue warps (two per TMEM lane quarter, alternating 32-column chunks) keep two
// warps on every SM sub-partition so the epilogue's shared-memory / TMEM / global latencies overlap.
// BN = 192 divides every hidden width of the model (768, 2304, 3072) and gives 72*4 = 288 tiles for a
// [9216 x 768] output: 1.95 waves on 148 SMs instead of 1.46 with BN = 256.  Operands may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows]);
// the latter is what dgrad (B = W[out,in]) and wgrad (A = dY, B = X) need, and is expressed purely
// through the shared-memory matrix descriptors - no transposed copies are ever made.
//
// CG = 2 ("CTA pair", tcgen05 cta_group::2): two CTAs of a cluster on the two SMs of a TPC compute one 256 x BN tile.
// Each CTA stages its own 128 rows of A but only HALF of the B tile (BN/2 rows); the leader CTA's single MMA thread
// issues tcgen05.mma.cta_group::2 (M = 256), which reads both halves and writes each CTA's 128 x BN accumulator into
// its own TMEM.  ncu showed the CG = 1 kernel pinned at the chip's L2->SM operand bandwidth (~12 TB/s of TMA loads,
// tensor pipe ~50 % busy): a 128 x 192 tile pulls 40 KB per 64-deep k-block (79 FLOP/B); the pair's 256 x 256 tile
// pulls 32 KB per CTA (131 FLOP/B).  Protocol: every CTA's TMA signals the LEADER's full barrier (mapa address,
// .cta_group::2 copies); tcgen05.commit multicasts the smem-slot release and the accumulator-ready signal to both
// CTAs; the peer's epilogue warps release the accumulator stage on the leader's barrier through DSMEM.
#pragma once
#include <cuda.h>

#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "gemm_common.cuh"

namespace vb {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;               // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
// Epilogue warps: a multiple of 4 (TMEM lane quarter = warp % 4); the warps of a lane quarter take the 32-column
// chunks round-robin.  Measured: 12 warps do not help - the GELU epilogue (59.6 us vs 58.5 us with 8) is bound by
// the latency chain tcgen05.ld -> math -> staging -> store inside a warp, not by issue slots, and the residual /
// dropout epilogues spill under the 128-register cap of a 512-thread block.  The count stays a per-kernel constant
// so a future epilogue can choose differently.
template <int EPI, typename OutT> __host__ __device__ constexpr int epi_warps() {
  (void)sizeof(OutT);
  return 8;
}
constexpr int MAX_EPI_WARPS = 12;
constexpr int EPI_WARP0 = 4;
constexpr uint32_t A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr uint32_t MN_BLOCK_BYTES = 64 * BK * 2;  // one 64(mn) x 64(k) MN-major TMA box = 8 KB

template <int BN, int CG, int EW> struct Cfg {
  static_assert(BN == 128 || BN == 192 || BN == 256, "tile widths: 128, 192 or 256 columns");
  static_assert(CG == 1 || CG == 2, "cta_group::1 or ::2");
  static constexpr uint32_t B_ROWS = BN / CG;                      // B rows (tile columns) staged by ONE CTA
  static constexpr uint32_t B_STAGE_BYTES = B_ROWS * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr uint32_t TMEM_COLS = (BN <= 128) ? 256 : 512;    // 2 accumulator stages, power of two
  static constexpr int NUM_THREADS = 128 + EW * 32;
  static constexpr size_t FIXED_BYTES = 1024 /*align slack*/ + EW * 6144 /*epilogue staging*/ +
                                        2 * BN * 4 /*bias*/ + 384 /*barriers, work-queue responses*/;
  static constexpr int STAGES_FIT = (int)((227 * 1024 - FIXED_BYTES) / STAGE_BYTES);
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;    // CG1: 5 / 4 / 3,  CG2: 7 / 6 / 5 stages
  static_assert(STAGES >= 3, "operand ring too shallow");
  static constexpr size_t SMEM_BYTES = FIXED_BYTES + (size_t)STAGES * STAGE_BYTES;
};

struct Params {
  EpiParams epi;
  int32_t num_m_blocks, num_n_blocks;
  int32_t kb_total, kb_per_split, split_k;
  int32_t a_kmajor, b_kmajor;
  uint32_t idesc;
  int32_t split_m_blk;   // tiles with m_blk >= split_m_blk take the second B operand / bias (row-split launch); INT_MAX = none
  const float* bias2;
  int32_t clc;           // 1: work items come from cluster launch control (grid = one cluster per work item); 0: static stride
  // Device-side extents (vb_gemm_args::m_limit / k_limit): only the first *m_limit rows of the output / the first
  // *k_limit elements of the reduction hold data; whole tiles / k-blocks beyond them are not scheduled.
  const int32_t* m_limit;
  const int32_t* k_limit;
  const int32_t* m_limit2;   // second row segment (rows >= seg_m_blk tiles): its own device-side row count
  int32_t seg_m_blk;         // first tile row of the second segment; INT_MAX = one segment
  // Two reduction segments (ACCUM): k-blocks [0, kseg_kb) accumulate into epi.D, k-blocks [kseg_kb, kb_total) into D2
  // (vb_gemm_args::k_seg / D2 / k_limit2); kseg_kb = 0: one segment.
  float* D2;
  const int32_t* k_limit2;
  int32_t kseg_kb;
};

// ---- PTX wrappers ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("volta_b200 gemm_tc: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// ---- CTA-pair (cluster of 2) helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
// ---- cluster launch control: a running cluster asks the hardware for the grid index of a cluster that has not been
// launched yet and does its work itself (dynamic persistent scheduling without a global counter).  The 16-byte answer
// is written to `resp` (with MULTICAST: at the same offset in every CTA of the cluster) and counted on `bar` (ditto).
template <bool MULTICAST>
__device__ __forceinline__ void clc_try_cancel(uint32_t resp, uint32_t bar) {
  if constexpr (MULTICAST)
    asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 [%0], [%1];"
                 ::"r"(resp), "r"(bar) : "memory");
  else
    asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];"
                 ::"r"(resp), "r"(bar) : "memory");
}
// true + blockIdx.x of the first CTA of the cancelled cluster, or false when nothing was left to cancel
__device__ __forceinline__ bool clc_read(uint32_t resp, uint32_t& ctaid_x) {
  uint32_t ok, x = 0, y = 0, z = 0;
  asm volatile(
      "{\n\t.reg .pred p1;\n\t.reg .b128 r;\n\t"
      "ld.shared.b128 r, [%4];\n\t"
      "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n\t"
      "selp.u32 %3, 1, 0, p1;\n\t"
      "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, %1, %2, _}, r;\n\t}"
      : "+r"(x), "+r"(y), "+r"(z), "=r"(ok)
      : "r"(resp)
      : "memory");
  ctaid_x = x;
  return ok != 0;
}
// TMA load of a CTA-pair kernel: data lands in THIS CTA's shared memory, completion is counted on `bar`, a
// shared::cluster address that may name the leader CTA's barrier.
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* map, uint32_t bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// TMA store of a [32 rows x 64 B] staging tile (64B-swizzled) to the output matrix; rows / columns outside the matrix
// are clipped by the hardware, so ragged edges need no special path.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <uint32_t COLS, int CG> __device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {  // one warp of EACH CTA of the pair executes this
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <uint32_t COLS, int CG> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// mbarrier arrives when all tcgen05.mma issued so far by this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// pair version: the barrier at this offset in BOTH CTAs of the pair receives the arrive
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B.
//   K-major : rows of 128 B (64 bf16 of K); 8-row groups SBO = 1024 B apart; LBO unused.
//   MN-major: rows of 128 B (64 bf16 of M/N) per K index; 8-K groups SBO = 1024 B apart;
//             successive 64-element M/N blocks LBO = 8192 B apart (one TMA box each).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, bool kmajor) {
  const uint64_t lbo = kmajor ? 1ull : (uint64_t)(MN_BLOCK_BYTES >> 4);
  const uint64_t sbo = 1024ull >> 4;
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (2ull << 61);
}

// ---- epilogue ---------------------------------------------------------------------------------
// Each epilogue warp owns 32 accumulator rows (its TMEM lane quarter) and every other 32-column chunk.
// tcgen05.ld hands every thread ONE row, so writing straight from registers would touch 32 different cache lines
// per store instruction.  Instead the warp goes through a private, padded shared-memory staging tile
// (32 rows x 32 columns) so that every global load / store / red instruction covers whole 64/128-byte row
// segments: residual / activation-derivative reads, D writes and the split-K fp32 reductions are coalesced.
constexpr uint32_t STG_WARP_BYTES = 6144;  // bf16: 2 output buffers + 1 aux buffer of 32 rows x 64 B (dense, 64B-swizzled,
                                           // the layout TMA stores read); fp32: 32 rows x 144 B (padded)
constexpr uint32_t STG_BUF_BYTES = 2048;

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <typename OutT> struct Stg {
  static constexpr int CW = 32;                          // columns per chunk = one tcgen05.ld.x32
  static constexpr int EV = 16 / (int)sizeof(OutT);      // elements per 16-byte vector (8 bf16 / 4 fp32)
  static constexpr int VPR = CW / EV;                    // vectors per staged row (4 / 8)
  static constexpr int RPI = 32 / VPR;                   // rows covered by one warp-wide 16 B access (8 / 4)
  static constexpr int ITERS = 32 / RPI;                 // accesses per tile (4 / 8)
  static constexpr uint32_t PITCH = CW * sizeof(OutT) + 16;  // 80 B / 144 B: conflict-free for both access patterns

  // coalesced global -> registers (rows of this warp; zero fill out of range).  INTERIOR: the whole 32 x CW chunk is
  // inside the matrix - one base pointer, no per-vector bounds logic (the common case; ~25 instructions less per vector).
  template <bool INTERIOR>
  static __device__ __forceinline__ void load_regs(uint4 (&r)[ITERS], const OutT* g, int64_t ld, int64_t m0, int64_t n0,
                                                   int64_t M, int64_t N, int lane) {
    if constexpr (INTERIOR) {
      const OutT* src = g + (m0 + lane / VPR) * ld + n0 + (lane % VPR) * EV;
#pragma unroll
      for (int i = 0; i < ITERS; ++i) r[i] = *reinterpret_cast<const uint4*>(src + (int64_t)(i * RPI) * ld);
      return;
    }
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int row = i * RPI + lane / VPR, v = lane % VPR;
      const int64_t m = m0 + row, n = n0 + v * EV;
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (m < M) {
        const OutT* src = g + m * ld + n;
        if (n + EV <= N) {
          val = *reinterpret_cast<const uint4*>(src);
        } else if (n < N) {
          OutT tmp[EV];
#pragma unroll
          for (int j = 0; j < EV; ++j) tmp[j] = (n + j < N) ? src[j] : from_f<OutT>(0.f);
          val = *reinterpret_cast<uint4*>(tmp);
        }
      }
      r[i] = val;
    }
  }
  static __device__ __forceinline__ void regs_to_stage(const uint4 (&r)[ITERS], uint8_t* stg, int lane) {
#pragma unroll
    for (int i = 0; i < ITERS; ++i)
      *reinterpret_cast<uint4*>(stg + (i * RPI + lane / VPR) * PITCH + (lane % VPR) * 16) = r[i];
  }
  // coalesced staging -> global store
  template <bool INTERIOR>
  static __device__ __forceinline__ void store_tile(const uint8_t* stg, OutT* g, int64_t ld, int64_t m0, int64_t n0,
                                                    int64_t M, int64_t N, int lane) {
    if constexpr (INTERIOR) {
      OutT* dst = g + (m0 + lane / VPR) * ld + n0 + (lane % VPR) * EV;
      const uint8_t* src = stg + (lane / VPR) * PITCH + (lane % VPR) * 16;
#pragma unroll
      for (int i = 0; i < ITERS; ++i)
        *reinterpret_cast<uint4*>(dst + (int64_t)(i * RPI) * ld) = *reinterpret_cast<const uint4*>(src + i * RPI * PITCH);
      return;
    }
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int row = i * RPI + lane / VPR, v = lane % VPR;
      const int64_t m = m0 + row, n = n0 + v * EV;
      if (m < M && n < N) {
        const uint4 val = *reinterpret_cast<const uint4*>(stg + row * PITCH + v * 16);
        OutT* dst = g + m * ld + n;
        if (n + EV <= N) {
          *reinterpret_cast<uint4*>(dst) = val;
        } else {
          const OutT* tmp = reinterpret_cast<const OutT*>(&val);
#pragma unroll
          for (int j = 0; j < EV; ++j) if (n + j < N) dst[j] = tmp[j];
        }
      }
    }
  }
};

// coalesced staging -> global fp32 reduction (split-K / shared-weight accumulation)
template <bool INTERIOR>
__device__ __forceinline__ void red_tile(const uint8_t* stg, float* g, int64_t ld, int64_t m0, int64_t n0, int64_t M,
                                         int64_t N, int lane) {
  using S = Stg<float>;
  if constexpr (INTERIOR) {
    float* dst = g + (m0 + lane / S::VPR) * ld + n0 + (lane % S::VPR) * 4;
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && (ld & 3) == 0) {
      const uint8_t* src = stg + (lane / S::VPR) * S::PITCH + (lane % S::VPR) * 16;
#pragma unroll
      for (int i = 0; i < S::ITERS; ++i) {
        const float4 val = *reinterpret_cast<const float4*>(src + i * S::RPI * S::PITCH);
        red_add_v4(dst + (int64_t)(i * S::RPI) * ld, val.x, val.y, val.z, val.w);
      }
      return;
    }
  }
#pragma unroll
  for (int i = 0; i < S::ITERS; ++i) {
    const int row = i * S::RPI + lane / S::VPR, v = lane % S::VPR;
    const int64_t m = m0 + row, n = n0 + v * 4;
    if (m < M && n < N) {
      const float4 val = *reinterpret_cast<const float4*>(stg + row * S::PITCH + v * 16);
      float* dst = g + m * ld + n;
      if (n + 4 <= N && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
        red_add_v4(dst, val.x, val.y, val.z, val.w);
      } else {
        const float t[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) if (n + j < N) atomicAdd(dst + j, t[j]);
      }
    }
  }
}

// All chunks of one tile that belong to this warp.  EPI is a compile-time constant: the epilogue kind is
// dispatched ONCE per tile (a per-element switch made the unrolled body overflow the instruction cache).
// bf16 outputs leave through TMA: each warp writes its 32 x 32 chunk as dense 64-byte rows (64B swizzle: 16-byte slot
// ^= (row >> 1) & 3, conflict-free for the row-per-lane writes) into one of two 2 KB buffers, one lane issues
// cp.async.bulk.tensor to D (and to aux_out for GELU) and the warp moves on; before a buffer is rewritten the lane
// waits until the store that last used it has read it (bulk_wait_read).  Measured on the GELU GEMM: the LDS + STG
// store path it replaces cost 12.5 of 58.5 us.  fp32 outputs keep the padded staging + per-thread vector stores.
#ifndef VB_EPI_DEFER_WAIT
#define VB_EPI_DEFER_WAIT 0   // measured: no gain (12.34 / 12.25 vs 12.29 / 12.32 ms per step), +24 registers; kept for the record
#endif
// Epilogues that read a second tensor (residual, activation derivative): the warp's three 2 KB buffers are used as ONE
// output buffer + TWO aux buffers, so that the TMA loads of the aux chunks run two chunks ahead (the first two are
// issued before the accumulator is even complete) instead of one; the single output buffer is rewritten only after a
// wait placed behind the chunk's math (the store engine has long read it by then).
#ifndef VB_EPI_AUX2
#define VB_EPI_AUX2 1
#endif
template <typename OutT, int EPI> __host__ __device__ constexpr bool epi_aux2() {
  return VB_EPI_AUX2 != 0 && sizeof(OutT) == 2 &&
         (EPI == VB_EPI_BIAS_DROP_RES || EPI == VB_EPI_MUL || EPI == VB_EPI_ADD || EPI == VB_EPI_DRELU || EPI == VB_EPI_BIAS_ADD);
}
// the first two aux chunks of a tile (this warp's share: chunks group, group + NUM_EPI_GROUPS): buffers 1 and 2, barriers 0 and 1
template <typename OutT, int BN, int NUM_EPI_GROUPS>
__device__ __forceinline__ void epilogue_aux_prefetch(const EpiParams& p, uint8_t* stg, int64_t m0, int64_t nb0, int lane, int group,
                                                      const CUtensorMap* mapAuxIn, uint32_t aux_bar, uint32_t aux_bar2) {
  using S = Stg<OutT>;
  constexpr int NCH = BN / S::CW;
  if (lane == 0) {
    if (nb0 + group * S::CW < p.N) {
      mbar_expect_tx(aux_bar, STG_BUF_BYTES);
      tma_load_2d(smem_u32(stg + STG_BUF_BYTES), mapAuxIn, aux_bar, (int32_t)(nb0 + group * S::CW), (int32_t)m0);
    }
    if (group + NUM_EPI_GROUPS < NCH && nb0 + (group + NUM_EPI_GROUPS) * S::CW < p.N) {
      mbar_expect_tx(aux_bar2, STG_BUF_BYTES);
      tma_load_2d(smem_u32(stg + 2 * STG_BUF_BYTES), mapAuxIn, aux_bar2, (int32_t)(nb0 + (group + NUM_EPI_GROUPS) * S::CW), (int32_t)m0);
    }
  }
}
template <typename OutT, int BN, int EPI, bool INTERIOR, int NUM_EPI_GROUPS>
__device__ __forceinline__ void epilogue_tile(const EpiParams& p, void* d_out, uint32_t tmem_acc, uint8_t* stg, int64_t m0, int64_t nb0,
                                           int lane, int group, const float* bias_s, const CUtensorMap* mapD,
                                           const CUtensorMap* mapAuxOut, const CUtensorMap* mapAuxIn, uint32_t aux_bar,
                                           uint32_t aux_bar2, uint32_t& aux_phase, int& ob_io) {
  using S = Stg<OutT>;
  constexpr bool TMA_OUT = sizeof(OutT) == 2 && EPI != VB_EPI_ACCUM;
  constexpr bool DEFER = VB_EPI_DEFER_WAIT != 0;         // two output buffers alternating over the warp's whole life, wait before the rewrite
  constexpr bool AUX2 = epi_aux2<OutT, EPI>();           // one output buffer (0), two aux buffers (1, 2)
  constexpr bool REG_STAGE = TMA_OUT && (DEFER || AUX2); // the chunk's output waits in registers until its buffer is free
  const int swz = (lane >> 1) & 3;                       // 64B-swizzle term of this lane's row
  uint8_t* const abuf = stg + (AUX2 ? 1 : 2) * STG_BUF_BYTES;   // aux_in staging (TMA_OUT layout)
  int ob = (DEFER && !AUX2) ? ob_io : 0;                 // output buffer in use
  if constexpr (TMA_OUT && !REG_STAGE) {
    if (lane == 0) bulk_wait_read<0>();                  // stores of the previous tile have drained their buffers
    __syncwarp();
  }
  constexpr bool use_bias = EPI == VB_EPI_BIAS || EPI == VB_EPI_BIAS_GELU || EPI == VB_EPI_BIAS_RELU ||
                            EPI == VB_EPI_BIAS_DROP_RES || EPI == VB_EPI_BIAS_ADD;
  constexpr bool use_aux = EPI == VB_EPI_BIAS_DROP_RES || EPI == VB_EPI_MUL || EPI == VB_EPI_ADD ||
                           EPI == VB_EPI_DRELU || EPI == VB_EPI_BIAS_ADD;
  constexpr int NCH = BN / S::CW;
  uint8_t* my_row = stg + lane * S::PITCH;
  const int64_t m = m0 + lane;
  const OutT* aux = reinterpret_cast<const OutT*>(p.aux_in);
  uint64_t seed = 0;
  if constexpr (EPI == VB_EPI_BIAS_DROP_RES) seed = effective_seed(p.seed, p.seed_ptr);
  uint4 pre[S::ITERS];
  if constexpr (AUX2) {
    // the first two aux chunks were requested by epilogue_aux_prefetch() before the accumulator was awaited
  } else if constexpr (use_aux && TMA_OUT) {
    // residual / activation-derivative chunks arrive by TMA (32 x 32 box, zero fill outside the matrix) into the aux buffer;
    // the load of chunk c+1 is issued as soon as chunk c has been read out of it
    if (lane == 0 && nb0 + group * S::CW < p.N) {
      mbar_expect_tx(aux_bar, STG_BUF_BYTES);
      tma_load_2d(smem_u32(abuf), mapAuxIn, aux_bar, (int32_t)(nb0 + group * S::CW), (int32_t)m0);
    }
  } else if constexpr (use_aux) {
    if (nb0 + group * S::CW < p.N) S::template load_regs<INTERIOR>(pre, aux, p.ld_aux, m0, nb0 + group * S::CW, p.M, p.N, lane);
  }
  uint4 auxr[S::VPR];
  int kchunk = 0;                                        // this warp's chunk counter inside the tile (AUX2: buffer / barrier parity)
#pragma unroll 1
  for (int c = group; c < NCH; c += NUM_EPI_GROUPS) {
    const int64_t n0 = nb0 + c * S::CW;
    if (n0 >= p.N) break;  // warp-uniform
    uint32_t raw[32];
    tmem_ld32(tmem_acc + (uint32_t)(c * S::CW), raw);
    if constexpr (AUX2) {
      const int which = kchunk & 1;
      const uint32_t bar = which ? aux_bar2 : aux_bar;
      uint8_t* const ab = abuf + which * STG_BUF_BYTES;
      mbar_wait(bar, (aux_phase >> which) & 1u);
      aux_phase ^= 1u << which;
#pragma unroll
      for (int v = 0; v < S::VPR; ++v) auxr[v] = *reinterpret_cast<const uint4*>(ab + lane * 64 + ((v ^ swz) << 4));
      __syncwarp();                                      // every lane has read its row: the buffer may be refilled
      if (lane == 0 && c + 2 * NUM_EPI_GROUPS < NCH && n0 + 2 * NUM_EPI_GROUPS * S::CW < p.N) {   // two chunks ahead
        mbar_expect_tx(bar, STG_BUF_BYTES);
        tma_load_2d(smem_u32(ab), mapAuxIn, bar, (int32_t)(n0 + 2 * NUM_EPI_GROUPS * S::CW), (int32_t)m0);
      }
      ++kchunk;
    } else if constexpr (use_aux && TMA_OUT) {
      mbar_wait(aux_bar, aux_phase);
      aux_phase ^= 1u;
#pragma unroll
      for (int v = 0; v < S::VPR; ++v) auxr[v] = *reinterpret_cast<const uint4*>(abuf + lane * 64 + ((v ^ swz) << 4));
      __syncwarp();                                      // every lane has read its row: the buffer may be refilled
      if (lane == 0 && c + NUM_EPI_GROUPS < NCH && n0 + NUM_EPI_GROUPS * S::CW < p.N) {
        mbar_expect_tx(aux_bar, STG_BUF_BYTES);
        tma_load_2d(smem_u32(abuf), mapAuxIn, aux_bar, (int32_t)(n0 + NUM_EPI_GROUPS * S::CW), (int32_t)m0);
      }
    } else if constexpr (use_aux) {
      S::regs_to_stage(pre, stg, lane);
      __syncwarp();
      // prefetch the next chunk's residual / derivative tile while this chunk is processed
      if (c + NUM_EPI_GROUPS < NCH && n0 + NUM_EPI_GROUPS * S::CW < p.N)
        S::template load_regs<INTERIOR>(pre, aux, p.ld_aux, m0, n0 + NUM_EPI_GROUPS * S::CW, p.M, p.N, lane);
    }
    tmem_ld_wait();
    uint4 packed[S::VPR];   // REG_STAGE: the chunk's first output, packed, until its staging buffer is known to be free
    float acc[S::CW];
#pragma unroll
    for (int j = 0; j < S::CW; ++j) acc[j] = __uint_as_float(raw[j]) * p.alpha;
#pragma unroll
    for (int v = 0; v < S::VPR; ++v) {
      const int64_t n = n0 + v * S::EV;
      float b[S::EV], x[S::EV];
#pragma unroll
      for (int j = 0; j < S::EV; ++j) { b[j] = 0.f; x[j] = 0.f; }
      if constexpr (use_bias) {  // this tile's bias slice was staged in shared memory (zero beyond N): broadcast reads
#pragma unroll
        for (int j = 0; j < S::EV; j += 4) {
          const float4 t = *reinterpret_cast<const float4*>(bias_s + c * S::CW + v * S::EV + j);
          b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
        }
      }
      if constexpr (use_aux) {
        uint4 rawx;
        if constexpr (TMA_OUT) rawx = auxr[v];
        else rawx = *reinterpret_cast<const uint4*>(my_row + v * 16);
        const OutT* t = reinterpret_cast<const OutT*>(&rawx);
#pragma unroll
        for (int j = 0; j < S::EV; ++j) x[j] = to_f<OutT>(t[j]);
      }
#pragma unroll
      for (int j = 0; j < S::EV; ++j) {
        float a = acc[v * S::EV + j];
        if constexpr (EPI == VB_EPI_BIAS || EPI == VB_EPI_BIAS_GELU || EPI == VB_EPI_BIAS_DROP_RES) a = a + b[j];
        else if constexpr (EPI == VB_EPI_BIAS_RELU) a = fmaxf(a + b[j], 0.f);
        else if constexpr (EPI == VB_EPI_MUL) a = a * x[j];
        else if constexpr (EPI == VB_EPI_ADD) a = a + x[j];
        else if constexpr (EPI == VB_EPI_DRELU) a = x[j] > 0.f ? a : 0.f;
        else if constexpr (EPI == VB_EPI_BIAS_ADD) a = a + b[j] + x[j];
        acc[v * S::EV + j] = a;
      }
      if constexpr (EPI == VB_EPI_BIAS_DROP_RES) {
        if (p.drop_thresh != 0u) {
          if ((p.N & 1) == 0) {  // one hash per 2 consecutive columns
#pragma unroll
            for (int h = 0; h < S::EV; h += 2) {
              bool k0, k1;
              dropout_keep2(seed, p.offset, (uint64_t)m * (uint64_t)p.N + (uint64_t)(n + h), p.drop_thresh, k0, k1);
              acc[v * S::EV + h] = k0 ? acc[v * S::EV + h] * p.drop_scale : 0.f;
              acc[v * S::EV + h + 1] = k1 ? acc[v * S::EV + h + 1] * p.drop_scale : 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < S::EV; ++j)
              acc[v * S::EV + j] = dropout_keep(seed, p.offset, (uint64_t)m * (uint64_t)p.N + (uint64_t)(n + j), p.drop_thresh)
                                       ? acc[v * S::EV + j] * p.drop_scale : 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < S::EV; ++j) acc[v * S::EV + j] += x[j];
      }
      OutT o[S::EV];
      if constexpr (EPI == VB_EPI_BIAS_GELU) {
        // stage gelu'(x) first (aux_out); gelu(x) stays in the accumulator registers for the second store pass
#pragma unroll
        for (int j = 0; j < S::EV; ++j) {
          float y, dy;
          gelu_fwd_grad_fast(acc[v * S::EV + j], y, dy);
          acc[v * S::EV + j] = y;
          o[j] = from_f<OutT>(dy);
        }
      } else {
#pragma unroll
        for (int j = 0; j < S::EV; ++j) o[j] = from_f<OutT>(acc[v * S::EV + j]);
      }
      if constexpr (REG_STAGE) packed[v] = *reinterpret_cast<uint4*>(o);
      else if constexpr (TMA_OUT) *reinterpret_cast<uint4*>(stg + ob * STG_BUF_BYTES + lane * 64 + ((v ^ swz) << 4)) = *reinterpret_cast<uint4*>(o);
      else *reinterpret_cast<uint4*>(my_row + v * 16) = *reinterpret_cast<uint4*>(o);
    }
    if constexpr (TMA_OUT) {
      // first (or only) output of the chunk: gelu'(x) -> aux_out for GELU, the result -> D otherwise
      if constexpr (REG_STAGE) {
        if (lane == 0) {
          if constexpr (AUX2) bulk_wait_read<0>();   // the one output buffer: the previous chunk's store has read it
          else bulk_wait_read<1>();                  // the store before the previous one - the last user of buffer `ob`
        }
        __syncwarp();
#pragma unroll
        for (int v = 0; v < S::VPR; ++v)
          *reinterpret_cast<uint4*>(stg + ob * STG_BUF_BYTES + lane * 64 + ((v ^ swz) << 4)) = packed[v];
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(EPI == VB_EPI_BIAS_GELU ? mapAuxOut : mapD, smem_u32(stg + ob * STG_BUF_BYTES), (int32_t)n0, (int32_t)m0);
        bulk_commit();
        if constexpr (!REG_STAGE) bulk_wait_read<1>();          // the OTHER buffer (store before this one) is free again
      }
      if constexpr (!AUX2) ob ^= 1;
      if constexpr (!REG_STAGE) __syncwarp();
      if constexpr (EPI == VB_EPI_BIAS_GELU) {
        if constexpr (REG_STAGE) {
          if (lane == 0) bulk_wait_read<1>();   // only the gelu' store just issued may still be reading (the other buffer)
          __syncwarp();
        }
#pragma unroll
        for (int v = 0; v < S::VPR; ++v) {
          OutT o[S::EV];
#pragma unroll
          for (int j = 0; j < S::EV; ++j) o[j] = from_f<OutT>(acc[v * S::EV + j]);
          *reinterpret_cast<uint4*>(stg + ob * STG_BUF_BYTES + lane * 64 + ((v ^ swz) << 4)) = *reinterpret_cast<uint4*>(o);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(mapD, smem_u32(stg + ob * STG_BUF_BYTES), (int32_t)n0, (int32_t)m0);
          bulk_commit();
          if constexpr (!REG_STAGE) bulk_wait_read<1>();
        }
        ob ^= 1;
        if constexpr (!REG_STAGE) __syncwarp();
      }
      continue;
    }
    __syncwarp();
    if constexpr (EPI == VB_EPI_ACCUM) {
      if constexpr (sizeof(OutT) == 4) red_tile<INTERIOR>(stg, reinterpret_cast<float*>(d_out), p.ldd, m0, n0, p.M, p.N, lane);
    } else {
      if constexpr (EPI == VB_EPI_BIAS_GELU) {
        S::template store_tile<INTERIOR>(stg, reinterpret_cast<OutT*>(p.aux_out), p.ld_aux, m0, n0, p.M, p.N, lane);
        __syncwarp();
#pragma unroll
        for (int v = 0; v < S::VPR; ++v) {
          OutT o[S::EV];
#pragma unroll
          for (int j = 0; j < S::EV; ++j) o[j] = from_f<OutT>(acc[v * S::EV + j]);
          *reinterpret_cast<uint4*>(my_row + v * 16) = *reinterpret_cast<uint4*>(o);
        }
        __syncwarp();
      }
      S::template store_tile<INTERIOR>(stg, reinterpret_cast<OutT*>(d_out), p.ldd, m0, n0, p.M, p.N, lane);
    }
    __syncwarp();
  }
  ob_io = ob;
}

template <int BN, int EPI, int CG, typename OutT>
__global__ void __launch_bounds__(128 + 32 * epi_warps<EPI, OutT>(), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmD,
               const __grid_constant__ CUtensorMap tmAux, const __grid_constant__ CUtensorMap tmAuxIn, const Params p) {
  constexpr int NUM_EPI_WARPS = epi_warps<EPI, OutT>();
  using C = Cfg<BN, CG, NUM_EPI_WARPS>;
  extern __shared__ uint8_t smem_raw[];
  // identical carve-up in both CTAs of a pair: the MMA descriptors and the multicast commits address the peer's
  // shared memory by the same offsets (dynamic shared memory starts at the same window offset in every CTA).
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle-128B tiles need 1024 B alignment
  const uint32_t stg_u32 = smem_base + C::STAGES * C::STAGE_BYTES;
  uint8_t* stg_base = smem_raw + (stg_u32 - smem_u32(smem_raw));
  float* bias_smem = reinterpret_cast<float*>(stg_base + NUM_EPI_WARPS * STG_WARP_BYTES);  // [2][BN]
  const uint32_t bars = stg_u32 + NUM_EPI_WARPS * STG_WARP_BYTES + 2 * BN * 4;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + 2 + s); };
  const uint32_t tmem_slot = bars + 8u * (2 * C::STAGES + 4);
  auto aux_bar = [&](int w) { return bars + 8u * (2 * C::STAGES + 5 + w); };   // one per epilogue warp: aux-chunk TMA loads
  auto clc_full = [&](int s) { return bars + 8u * (2 * C::STAGES + 13 + s); };  // work queue: answer s has landed
  auto clc_empty = [&](int s) { return bars + 8u * (2 * C::STAGES + 15 + s); }; // (leader's copy) every reader is done with answer s
  auto clc_resp = [&](int s) { return bars + 8u * (2 * C::STAGES + 18) + 16u * s; };
  auto aux_bar2 = [&](int w) { return bars + 8u * (2 * C::STAGES + 22 + w); };  // second aux-chunk barrier per epilogue warp (epi_aux2)
  static_assert(8 * (2 * C::STAGES + 22 + NUM_EPI_WARPS) <= 384, "barrier area (Cfg::FIXED_BYTES) too small");
  auto smem_a = [&](int s) { return smem_base + s * C::STAGE_BYTES; };
  auto smem_b = [&](int s) { return smem_base + s * C::STAGE_BYTES + A_STAGE_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs)
  // Programmatic dependent launch, primary side: let the next kernel of the stream (if it was launched with the
  // programmatic-serialization attribute, as every GEMM here is) be scheduled as soon as SMs drain.  This kernel owns
  // a whole SM per CTA, so a dependent CTA starts only where one of ours has exited, runs its prologue (barriers,
  // TMEM allocation, descriptor prefetch) under our tail and then blocks in cudaGridDependencySynchronize().
  cudaTriggerProgrammaticLaunchCompletion();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmB2);
    tma_prefetch_desc(&tmD);
    tma_prefetch_desc(&tmAux);
    tma_prefetch_desc(&tmAuxIn);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), CG * NUM_EPI_WARPS); }
    for (int w = 0; w < NUM_EPI_WARPS; ++w) { mbar_init(aux_bar(w), 1); mbar_init(aux_bar2(w), 1); }
    // readers of a work-queue answer besides the leader's producer thread (which issues the queries): per CTA the
    // epilogue warps + one of {MMA thread (leader), producer thread (peer)}
    for (int s = 0; s < 2; ++s) { mbar_init(clc_full(s), 1); mbar_init(clc_empty(s), CG * (NUM_EPI_WARPS + 1)); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 2) tmem_alloc<C::TMEM_COLS, CG>(tmem_slot);
  tcgen05_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps
  // the tail of the previous kernel in the stream; global memory is only touched after this point.
  cudaGridDependencySynchronize();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  int num_m_blocks = p.num_m_blocks, kb_total = p.kb_total, split_k = p.split_k;
  // live tile rows: [0, live_a) of the first segment, then [seg_blk, seg_blk + live_b) of the second (packed text rows
  // and packed vision rows of a joint buffer); the work list enumerates live tile rows only
  const int seg_blk = min(p.seg_m_blk, num_m_blocks);
  int live_a = seg_blk, live_b = num_m_blocks - seg_blk;
  // contract (vb_gemm_args::m_limit): rows are produced in whole 256-row groups whatever the tile height - a 128-row
  // (CG = 1) tile grid must not stop at the first 128 rows, the consumers reduce up to the count padded to 256
  auto live_tiles = [](int lim) { return (((max(lim, 0) + 255) & ~255) + BM * CG - 1) / (BM * CG); };
  if (p.m_limit) live_a = min(live_a, live_tiles(*p.m_limit));
  if (p.m_limit2) live_b = min(live_b, live_tiles(*p.m_limit2));
  num_m_blocks = live_a + live_b;
  auto m_of = [&](int m_idx) { return m_idx < live_a ? m_idx : seg_blk + (m_idx - live_a); };
  int kb_per_split = p.kb_per_split;
  // reduction segments: A = k-blocks [0, kbA) -> D, B = k-blocks [kbB0, kbB0 + kbB) -> D2 (one segment: kbB = 0)
  int kbA = kb_total, kbB0 = 0, kbB = 0;
  if (p.kseg_kb > 0) { kbA = min(p.kseg_kb, kb_total); kbB0 = kbA; kbB = kb_total - kbA; }
  if (p.k_limit) kbA = min(kbA, (max(*p.k_limit, 0) + BK - 1) / BK);
  if (p.k_limit2) kbB = min(kbB, (max(*p.k_limit2, 0) + BK - 1) / BK);
  int splitA = split_k;
  if (p.k_limit || p.kseg_kb > 0) {
    // keep the number of slices the host planned for the machine, shorten them evenly to the live reduction length
    kb_per_split = max(1, (kbA + kbB + p.split_k - 1) / p.split_k);
    // two segments round up separately: stay within the number of slices the host planned (one more slice than CTA pairs
    // would cost a whole extra round)
    while (kbB > 0 && kb_per_split < kbA + kbB &&
           (kbA + kb_per_split - 1) / kb_per_split + (kbB + kb_per_split - 1) / kb_per_split > p.split_k) ++kb_per_split;
    splitA = (kbA + kb_per_split - 1) / kb_per_split;
    split_k = splitA + (kbB + kb_per_split - 1) / kb_per_split;
  }
  kb_total = kbA;
  // k-block range [kb0, kb1) and output segment of slice `split` of a tile
  auto slice_of = [&](int split, int& kb0, int& kb1) -> bool {
    if (split < splitA) { kb0 = split * kb_per_split; kb1 = min(kbA, kb0 + kb_per_split); return false; }
    kb0 = kbB0 + (split - splitA) * kb_per_split; kb1 = min(kbB0 + kbB, kb0 + kb_per_split); return true;
  };
  const int64_t tiles = (int64_t)num_m_blocks * p.num_n_blocks;   // tiles of (CG*128) x BN
  const int64_t total_work = tiles * split_k;
  const int64_t first = blockIdx.x / CG, step = gridDim.x / CG;     // one work item per CTA (pair) at a time
  // Work items.  Static mode: first, first + step, ... (grid = resident CTAs).  Cluster-launch-control mode (p.clc): the
  // grid has one cluster per work item; a resident cluster does its own item, then keeps cancelling clusters that have
  // not started and does theirs.  SMs held by another kernel (a collective, a side-stream reduction) then simply take
  // no part, instead of delaying a fixed share of the tiles until they are free.  The leader's producer thread issues
  // query q when it STARTS loading an item (at most one in flight, none after a failed one), everyone else reads
  // answer q from slot q & 1 and releases it on the leader's clc_empty barrier.
  auto next_work = [&](int64_t& w, int& q, bool arrive, bool whole_warp) -> bool {
    if (!p.clc) { w += step; return w < total_work; }
    const int slot = q & 1; const uint32_t par = (uint32_t)(q >> 1) & 1u; ++q;
    mbar_wait(clc_full(slot), par);
    uint32_t x;
    const bool ok = clc_read(clc_resp(slot), x);
    fence_proxy_async();   // our generic-proxy read is ordered before the async-proxy write of a later answer
    if (whole_warp) __syncwarp();   // every lane has read before lane 0 releases the slot
    if (arrive) {
      if constexpr (CG == 2) mbar_arrive_cluster(mapa_shared(clc_empty(slot), 0)); else mbar_arrive(clc_empty(slot));
    }
    w = (int64_t)(x / CG);
    return ok;
  };

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer (every CTA loads its own A rows and its share of B) =====================
    int stage = 0; uint32_t phase = 0;
    int q = 0;
    if (first < total_work) for (int64_t w = first;;) {
      if (p.clc && rank == 0) {
        const int slot = q & 1; const uint32_t par = (uint32_t)(q >> 1) & 1u;
        mbar_wait(clc_empty(slot), par ^ 1u);
        mbar_expect_tx(clc_full(slot), 16);
        if constexpr (CG == 2) mbar_expect_tx_cluster(mapa_shared(clc_full(slot), 1), 16);
        clc_try_cancel<CG == 2>(clc_resp(slot), clc_full(slot));
      }
      const int split = (int)(w % split_k);
      const int64_t tile = w / split_k;
      const int n_blk = (int)(tile % p.num_n_blocks), m_blk = m_of((int)(tile / p.num_n_blocks));
      const int m0 = (m_blk * CG + (int)rank) * BM;
      const int n0 = n_blk * BN + (int)rank * (int)C::B_ROWS;
      const CUtensorMap* mapB = (m_blk >= p.split_m_blk) ? &tmB2 : &tmB;   // row-split launch: second stream's weights
      int kb0, kb1;
      slice_of(split, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        uint32_t fb = full_bar(stage);
        if constexpr (CG == 2) {
          if (rank == 0) mbar_expect_tx(fb, 2 * C::STAGE_BYTES);  // both CTAs' copies complete on the leader's barrier
          fb = mapa_shared(fb, 0);
        } else {
          mbar_expect_tx(fb, C::STAGE_BYTES);
        }
        auto load = [&](uint32_t dst, const CUtensorMap* map, int c0, int c1) {
          if constexpr (CG == 2) tma_load_2d_cg2(dst, map, fb, c0, c1); else tma_load_2d(dst, map, fb, c0, c1);
        };
        if (p.a_kmajor) {
          load(smem_a(stage), &tmA, kb * BK, m0);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) load(smem_a(stage) + j * MN_BLOCK_BYTES, &tmA, m0 + j * 64, kb * BK);
        }
        if (p.b_kmajor) {
          load(smem_b(stage), mapB, kb * BK, n0);
        } else {
#pragma unroll
          for (int j = 0; j < (int)C::B_ROWS / 64; ++j) load(smem_b(stage) + j * MN_BLOCK_BYTES, mapB, n0 + j * 64, kb * BK);
        }
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
      if (!next_work(w, q, rank != 0, false)) break;
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ===================== MMA issuer (leader CTA only) =====================
    int stage = 0; uint32_t phase = 0; int it = 0;
    const uint32_t a_step = p.a_kmajor ? (UMMA_K * 2) : (UMMA_K * 128);  // bytes per UMMA_K step
    const uint32_t b_step = p.b_kmajor ? (UMMA_K * 2) : (UMMA_K * 128);
    int q = 0;
    if (first < total_work) for (int64_t w = first;; ++it) {
      const int split = (int)(w % split_k);
      int kb0, kb1;
      slice_of(split, kb0, kb1);
      const int as = it & 1; const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(tempty_bar(as), aphase ^ 1u);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t adesc = make_smem_desc(smem_a(stage) + k * a_step, p.a_kmajor);
          const uint64_t bdesc = make_smem_desc(smem_b(stage) + k * b_step, p.b_kmajor);
          if constexpr (CG == 2) umma_bf16_cg2(tmem_d, adesc, bdesc, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          else umma_bf16(tmem_d, adesc, bdesc, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        // smem slot reusable (in both CTAs) once these MMAs retire
        if constexpr (CG == 2) umma_commit_cg2(empty_bar(stage)); else umma_commit(empty_bar(stage));
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
      // accumulator ready for the epilogue warps (of both CTAs)
      if constexpr (CG == 2) umma_commit_cg2(tfull_bar(as)); else umma_commit(tfull_bar(as));
      if (!next_work(w, q, true, false)) break;
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue (every CTA drains its own 128 x BN accumulator) =====================
    const int ew = (warp - EPI_WARP0) & 3;    // == warp % 4: TMEM lane quarter this warp may access
    const int group = (warp - EPI_WARP0) >> 2;  // which share of the 32-column chunks
    uint8_t* stg = stg_base + (warp - EPI_WARP0) * STG_WARP_BYTES;
    const uint32_t my_aux_bar = aux_bar(warp - EPI_WARP0), my_aux_bar2 = aux_bar2(warp - EPI_WARP0);
    uint32_t aux_phase = 0;   // bit 0: parity of my_aux_bar, bit 1: of my_aux_bar2
    int out_buf = 0;       // which of the warp's two TMA-store staging buffers the next store uses
    int it = 0, q = 0;
    if (first < total_work) for (int64_t w = first;; ++it) {
      const int64_t tile = w / split_k;
      const int n_blk = (int)(tile % p.num_n_blocks), m_blk = m_of((int)(tile / p.num_n_blocks));
      const int as = it & 1; const uint32_t aphase = (it >> 1) & 1;
      // Stage this tile's bias slice in shared memory while the MMAs are still running.  Double buffered by
      // accumulator stage; the named barrier also guarantees nobody still reads the buffer of tile it-2.
      float* bias_cur = bias_smem + as * BN;
      const int64_t m0 = ((int64_t)m_blk * CG + rank) * BM + ew * 32;
      if constexpr (epi_aux2<OutT, EPI>()) {
        // the residual / derivative chunks of this tile do not depend on the accumulator: request the first two now
        if (m0 < p.epi.M)
          epilogue_aux_prefetch<OutT, BN, NUM_EPI_WARPS / 4>(p.epi, stg, m0, (int64_t)n_blk * BN, lane, group, &tmAuxIn, my_aux_bar, my_aux_bar2);
      }
      if constexpr (EPI == VB_EPI_BIAS || EPI == VB_EPI_BIAS_GELU || EPI == VB_EPI_BIAS_RELU || EPI == VB_EPI_BIAS_DROP_RES || EPI == VB_EPI_BIAS_ADD) {
        const float* bias_g = (m_blk >= p.split_m_blk) ? p.bias2 : p.epi.bias;
        for (int j = threadIdx.x - EPI_WARP0 * 32; j < BN; j += NUM_EPI_WARPS * 32) {
          const int64_t n = (int64_t)n_blk * BN + j;
          bias_cur[j] = n < p.epi.N ? __ldg(bias_g + n) : 0.f;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_WARPS * 32) : "memory");
      mbar_wait(tfull_bar(as), aphase);
      tcgen05_fence_after();
      // the answer for the next item was requested when this item's loads began: it is there by now, take it (and
      // release the slot) before the long part
      int64_t w_next = w;
      const bool more = next_work(w_next, q, lane == 0, true);
      const uint32_t tmem_acc = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN);
      void* d_out = p.epi.D;
      if constexpr (EPI == VB_EPI_ACCUM) {
        int kb0_, kb1_;
        if (slice_of((int)(w % split_k), kb0_, kb1_)) d_out = p.D2;   // second reduction segment: the other stream's gradient
      }
      if (m0 < p.epi.M) {  // warp-uniform
        const int64_t nb0 = (int64_t)n_blk * BN;
        const bool interior = m0 + 32 <= p.epi.M && nb0 + BN <= p.epi.N;  // warp-uniform
        // OutT is a kernel template parameter: one epilogue body per kernel (code size, register allocation)
        if (interior) epilogue_tile<OutT, BN, EPI, true, NUM_EPI_WARPS / 4>(p.epi, d_out, tmem_acc, stg, m0, nb0, lane, group, bias_cur, &tmD, &tmAux, &tmAuxIn, my_aux_bar, my_aux_bar2, aux_phase, out_buf);
        else epilogue_tile<OutT, BN, EPI, false, NUM_EPI_WARPS / 4>(p.epi, d_out, tmem_acc, stg, m0, nb0, lane, group, bias_cur, &tmD, &tmAux, &tmAuxIn, my_aux_bar, my_aux_bar2, aux_phase, out_buf);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(mapa_shared(tempty_bar(as), 0));  // the leader's MMA thread waits for both CTAs
        else mbar_arrive(tempty_bar(as));
      }
      if (!more) break;
      w = w_next;
    }
    if constexpr (sizeof(OutT) == 2 && EPI != VB_EPI_ACCUM) {
      if (lane == 0) bulk_wait_all();   // outstanding TMA stores must finish before the staging memory goes away
    }
  }

  tcgen05_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();  // the peer's shared memory / TMEM stay valid until both are done
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<C::TMEM_COLS, CG>(tmem_base);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

// 2-D bf16 row-major matrix [rows, cols] (pitch ld elements), box = 64 cols x box_rows, 128B swizzle.
inline int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  VB_REQUIRE(enc != nullptr, VB_ERR_DEVICE, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VB_REQUIRE(r == CUDA_SUCCESS, VB_ERR_ARG, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
             (long long)rows, (long long)cols, (long long)ld);
  return 0;
}

// bf16 output matrix [rows, cols] (pitch ld elements) for the epilogue's TMA stores: box = 32 columns x 32 rows, 64B swizzle.
inline int make_out_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn enc = get_encode();
  VB_REQUIRE(enc != nullptr, VB_ERR_DEVICE, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {32u, 32u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VB_REQUIRE(r == CUDA_SUCCESS, VB_ERR_ARG, "cuTensorMapEncodeTiled (output) failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
             (long long)rows, (long long)cols, (long long)ld);
  return 0;
}

inline bool use_clc() {
  const char* e = getenv("VOLTA_B200_GEMM_CLC");   // read per launch: tests switch it
  return e && e[0] == '1';
}

inline int sm_count() {
  static int g_sm_count = 0;
  if (g_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  // vb_set_gemm_sm_limit: the data-parallel backward leaves a few SMs to NCCL.  The persistent kernel assigns tiles
  // statically, so a CTA that cannot be placed (its SM runs a collective) would delay its tiles by the whole collective.
  const int lim = gemm_sm_limit();
  return (lim > 0 && lim < g_sm_count) ? std::max(2, lim & ~1) : g_sm_count;
}

template <int BN, int EPI, int CG, typename OutT>
static int launch_t(const vb_gemm_args& a, cudaStream_t stream) {
  using C = Cfg<BN, CG, epi_warps<EPI, OutT>()>;
  Params p;
  int rc = fill_epi_params(a, p.epi);
  if (rc) return rc;
  p.num_m_blocks = (int)ceil_div(a.M, BM * CG);
  p.num_n_blocks = (int)ceil_div(a.N, BN);
  p.kb_total = (int)ceil_div(a.K, BK);
  p.a_kmajor = a.a_kmajor;
  p.b_kmajor = a.b_kmajor;
  const int64_t tiles = (int64_t)p.num_m_blocks * p.num_n_blocks;
  const int64_t units = std::max(1, sm_count() / CG);  // CTAs (pairs) that run concurrently
  int split = 1;
  if (a.epilogue == VB_EPI_ACCUM) {
    split = a.split_k;
    if (split <= 0) {
      // Split-K for the weight gradients (few output tiles, long reduction).  Cost model in k-block units:
      // rounds over the SMs x (k-blocks per slice + epilogue of one fp32 red tile ~ 8 k-blocks); smallest wins.
      const int max_split = (int)std::max<int64_t>(1, p.kb_total / 4);
      int64_t best = -1;
      for (int s_ = 1; s_ <= max_split && s_ <= 64; ++s_) {
        const int64_t per = ceil_div(p.kb_total, s_);
        const int64_t eff = ceil_div(p.kb_total, per);
        const int64_t cost = ceil_div(tiles * eff, units) * (per + 8);
        if (best < 0 || cost < best) { best = cost; split = s_; }
      }
    }
    split = std::max(1, std::min(split, p.kb_total));
  }
  p.kb_per_split = (int)ceil_div(p.kb_total, split);
  p.split_k = (int)ceil_div(p.kb_total, p.kb_per_split);
  // instruction descriptor (cute::UMMA::InstrDescriptor): F32 accum, BF16 x BF16, majors, N>>3, M>>4
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((a.a_kmajor ? 0u : 1u) << 15) | ((a.b_kmajor ? 0u : 1u) << 16) |
            ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);

  CUtensorMap tmA, tmB;
  if (a.a_kmajor) rc = make_map(&tmA, a.A, a.M, a.K, a.lda, BM);
  else rc = make_map(&tmA, a.A, a.K, a.M, a.lda, BK);
  if (rc) return rc;
  if (a.b_kmajor) rc = make_map(&tmB, a.B, a.N, a.K, a.ldb, (int)C::B_ROWS);
  else rc = make_map(&tmB, a.B, a.K, a.N, a.ldb, BK);
  if (rc) return rc;
  CUtensorMap tmB2 = tmB;
  p.split_m_blk = INT32_MAX;
  p.bias2 = p.epi.bias;
  if (a.B2 != nullptr && a.split_M > 0 && a.split_M < a.M) {
    VB_REQUIRE(a.split_M % (BM * CG) == 0, VB_ERR_ARG, "gemm: split_M=%lld is not a multiple of the %d-row tile", (long long)a.split_M, BM * CG);
    if (a.b_kmajor) rc = make_map(&tmB2, a.B2, a.N, a.K, a.ldb, (int)C::B_ROWS);
    else rc = make_map(&tmB2, a.B2, a.K, a.N, a.ldb, BK);
    if (rc) return rc;
    p.split_m_blk = (int32_t)(a.split_M / (BM * CG));
    p.bias2 = a.bias2 ? a.bias2 : p.epi.bias;
  }

  CUtensorMap tmD = tmA, tmAux = tmA;   // placeholders unless the epilogue stores through TMA (bf16 outputs)
  if (sizeof(OutT) == 2 && EPI != VB_EPI_ACCUM) {
    if ((rc = make_out_map(&tmD, a.D, a.M, a.N, a.ldd))) return rc;
    tmAux = tmD;
    if (EPI == VB_EPI_BIAS_GELU && (rc = make_out_map(&tmAux, a.aux_out, a.M, a.N, a.ld_aux))) return rc;
  }
  CUtensorMap tmAuxIn = tmD;
  if (sizeof(OutT) == 2 && epi_uses_aux_in(EPI) && (rc = make_out_map(&tmAuxIn, a.aux_in, a.M, a.N, a.ld_aux))) return rc;

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI, CG, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
    VB_REQUIRE(e == cudaSuccess, (int)e, "cudaFuncSetAttribute(gemm_tc): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int64_t work = tiles * p.split_k;
  // default: classic persistent grid with a static stride.  VOLTA_B200_GEMM_CLC=1: cluster launch control - one cluster
  // per work item, resident clusters take over the ones not yet launched.  Measured (B200): 2-5 us slower per GEMM
  // in isolation (qkv 35.5 vs 31.8 us), equal inside the single-GPU step, 0.25 ms/step SLOWER with the NCCL
  // all-reduce running next to the backward (19.1 vs 18.85 ms at 2 GPUs) - so the collective does not starve
  // statically assigned tiles, and the static grid stays the default.
  p.m_limit = a.m_limit;
  p.k_limit = a.k_limit;
  p.m_limit2 = nullptr;
  p.seg_m_blk = INT32_MAX;
  if (a.limit_seg_M > 0 && a.limit_seg_M < a.M) {
    VB_REQUIRE(a.limit_seg_M % (BM * CG) == 0, VB_ERR_ARG, "gemm: limit_seg_M=%lld is not a multiple of the %d-row tile",
               (long long)a.limit_seg_M, BM * CG);
    p.seg_m_blk = (int32_t)(a.limit_seg_M / (BM * CG));
    p.m_limit2 = a.m_limit2;
  }
  p.D2 = nullptr; p.k_limit2 = nullptr; p.kseg_kb = 0;
  if (EPI == VB_EPI_ACCUM && a.D2 != nullptr && a.k_seg > 0 && a.k_seg < a.K) {
    p.D2 = reinterpret_cast<float*>(a.D2);
    p.k_limit2 = a.k_limit2;
    p.kseg_kb = (int32_t)(a.k_seg / BK);
  }
  p.clc = (use_clc() && work * CG < (int64_t)INT32_MAX && !a.m_limit && !a.k_limit && !p.m_limit2 && !p.kseg_kb) ? 1 : 0;
  const int grid = (int)(p.clc ? work : std::min<int64_t>(work, units)) * CG;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(C::NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.numAttrs = 1;
  if (CG == 2) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  cfg.attrs = attr;
  cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, EPI, CG, OutT>, tmA, tmB, tmB2, tmD, tmAux, tmAuxIn, p);
  VB_REQUIRE(le == cudaSuccess, (int)le, "gemm_tc_kernel: launch failed: %s", cudaGetErrorString(le));
  return check_launch("gemm_tc_kernel");
}

template <int BN, int EPI, int CG>
static int launch(const vb_gemm_args& a, cudaStream_t stream) {
  if constexpr (EPI == VB_EPI_ACCUM) return launch_t<BN, EPI, CG, float>(a, stream);
  else return a.d_dtype == VB_BF16 ? launch_t<BN, EPI, CG, bf16>(a, stream) : launch_t<BN, EPI, CG, float>(a, stream);
}

// Tile selection.  The kernel is bound by the L2 -> SM operand stream or by the tensor pipe, whichever is slower
// (cycles per 64-deep k-block of one CTA): bytes / ~44 B per cycle per SM  vs  2 * BN cycles of MMA, plus the
// epilogue tail of the last tile; total = rounds over the CTAs (pairs) x k-blocks x that.  CTA pairs (CG = 2) halve the
// B-operand bytes; they need M > 128 to be worth the idle half, and an MN-major B splits only in 64-column blocks.
template <int EPI>
static int launch_epi(const vb_gemm_args& a, cudaStream_t stream) {
  const int64_t sms = std::max(1, sm_count());
  struct Opt { int bn, cg; };
  const Opt opts[] = {{128, 1}, {192, 1}, {192, 2}, {256, 2}};
  double best = -1.0;
  Opt pick = {192, 1};
  // VOLTA_B200_GEMM_TILE=<BN>x<CG> pins the tile (tests exercise every instantiation with it)
  const char* force = getenv("VOLTA_B200_GEMM_TILE");
  int fbn = 0, fcg = 0;
  if (force && sscanf(force, "%dx%d", &fbn, &fcg) != 2) fbn = fcg = 0;
  for (const Opt& o : opts) {
    if (o.cg == 2 && !a.b_kmajor && ((o.bn / 2) % 64) != 0) continue;   // not expressible
    if (fbn) {
      if (o.bn == fbn && o.cg == fcg) { pick = o; best = 0.0; }
      continue;
    }
    if (o.cg == 2 && (a.M <= 128 || (sms % 2) != 0)) continue;
    if (a.B2 != nullptr && a.split_M > 0 && (a.split_M % (128 * o.cg)) != 0) continue;   // row split must fall on a tile edge
    if (a.limit_seg_M > 0 && (a.limit_seg_M % (128 * o.cg)) != 0) continue;
    if (o.bn > 128 && a.N <= 128 && o.cg == 1) continue;
    // rows that will really be produced: the caller's estimate of the device-side count when there is one
    const int64_t Mlive = (a.m_hint > 0 && (a.m_limit || a.m_limit2)) ? std::min<int64_t>(a.m_hint, a.M) : a.M;
    const int64_t tiles = ceil_div(Mlive, 128 * o.cg) * ceil_div(a.N, o.bn);
    const int64_t rounds = ceil_div(tiles, sms / o.cg);
    const double per_kb = std::max(2.0 * o.bn, (16384.0 + 128.0 * o.bn / o.cg) / 44.0);
    double cost = (double)rounds * per_kb;
    if (EPI == VB_EPI_ACCUM) cost = per_kb / o.bn;  // split-K evens out the rounds: cheapest bytes per column wins
    if (best < 0 || cost < best) { best = cost; pick = o; }
  }
  if (pick.cg == 2) return pick.bn == 256 ? launch<256, EPI, 2>(a, stream) : launch<192, EPI, 2>(a, stream);
  return pick.bn == 128 ? launch<128, EPI, 1>(a, stream) : launch<192, EPI, 1>(a, stream);
}

// one translation unit per group of epilogues (parallel build; each kernel gets its own register allocation)
#define VB_DEFINE_TC_LAUNCH(E) \
  int gemm_tc_launch_##E(const vb_gemm_args& a, cudaStream_t stream) { return tc::launch_epi<E>(a, stream); }

}  // namespace tc
}  // namespace vb
