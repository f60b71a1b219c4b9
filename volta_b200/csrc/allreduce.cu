// Gradient all-reduce over NVLink 5 / NVSwitch, written for the data-parallel backward of this library.
//
// Replaces, for the bucketed gradient exchange of apex DistributedDataParallel (apex/apex/parallel/distributed.py:425-475:
// flatten -> dist.all_reduce -> mul_(1/world) -> unflatten), the NCCL ring kernel.  Why a kernel of our own: the backward
// GEMMs are persistent kernels that own every SM with 227 KB of shared memory per CTA.  An NCCL kernel needs shared
// memory of its own, so each of its CTAs takes a whole SM away and the GEMM CTA assigned to that SM waits for the
// collective to end (measured in round 1: +2 ms on a 12 ms step at 8 GPUs).  This kernel uses NO shared memory, 128
// threads and <= 40 registers per block, so its blocks are co-resident with a GEMM CTA on the same SM, and it lets the
// NVSwitch do the reduction:
//
//   every rank owns 1/world of the bucket;  for its slice it issues   multimem.ld_reduce.add.v4.f32  on the MULTICAST
//   address of the bucket (the switch reads the 16 bytes from all ranks' HBM and returns their sum), scales by 1/world,
//   and writes the result back with  multimem.st  (the switch stores it into every rank's buffer).
//   Per GPU and bucket byte: one HBM read, one HBM write, one byte out and one byte in on NVLink - versus 2(N-1)/N bytes
//   each way plus staging copies for a ring.
//
// The buffers are symmetric memory (allocated and exchanged by torch.distributed._symmetric_memory: plumbing); the
// kernel, the cross-rank barriers on the signal pads and the launch plan are ours.  Without multicast support (no
// NVSwitch / fabric manager) the same kernel runs a two-shot exchange over plain peer pointers (P2P loads of the own
// slice from every rank, P2P stores of the result to every rank).
#include "common.cuh"

namespace vb {
namespace ar {

constexpr int THREADS = 128;
constexpr int MAX_BLOCKS = 304;   // up to two blocks per SM
constexpr int MAX_WORLD = 16;
constexpr int PAD_OFFSET_BYTES = 4096;   // our flags live in [4096, 4096 + MAX_BLOCKS * MAX_WORLD * 4) of the signal pad

struct Args {
  const uint64_t* buffer_ptrs;     // device array [world]: base address of the symmetric buffer on every rank (peer-mapped)
  uint64_t multicast_ptr;          // multicast (NVLS) address of the same buffer, 0 = not available
  const uint64_t* signal_pad_ptrs; // device array [world]: base address of every rank's signal pad
  int rank, world;
  int64_t offset, count;           // element range [offset, offset + count) of the fp32 buffer; both multiples of 4
  float scale;                     // applied to the sum (1/world = average)
};

__device__ __forceinline__ uint32_t cas_sys_relaxed(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.relaxed.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}

// Barrier between the blocks of the same index on every rank.  Flag (rank r -> rank p, block b) lives in p's signal pad;
// r raises it 0 -> 1, p lowers it 1 -> 0 when it has seen it, so consecutive barriers need no phase bookkeeping.
// Bounded spin: a protocol bug (ranks disagreeing on the bucket sequence) traps instead of hanging the GPU.
__device__ __forceinline__ void rank_barrier(const Args& a) {
  __syncthreads();
  if (threadIdx.x < a.world) {
    __threadfence_system();                                    // our earlier writes are visible before the flag
    const int peer = threadIdx.x;
    uint32_t* put = reinterpret_cast<uint32_t*>(a.signal_pad_ptrs[peer] + PAD_OFFSET_BYTES) + blockIdx.x * MAX_WORLD + a.rank;
    uint32_t* get = reinterpret_cast<uint32_t*>(a.signal_pad_ptrs[a.rank] + PAD_OFFSET_BYTES) + blockIdx.x * MAX_WORLD + peer;
    const long long t0 = clock64();
    while (cas_sys_relaxed(put, 0u, 1u) != 0u) {
      if (clock64() - t0 > 120000000000LL) { printf("volta_b200 allreduce: barrier (put) timed out, rank %d block %d peer %d\n", a.rank, blockIdx.x, peer); __trap(); }
    }
    while (cas_sys_relaxed(get, 1u, 0u) != 1u) {
      if (clock64() - t0 > 120000000000LL) { printf("volta_b200 allreduce: barrier (wait) timed out, rank %d block %d peer %d\n", a.rank, blockIdx.x, peer); __trap(); }
    }
    __threadfence_system();                                    // the peers' earlier writes are visible after the flag
  }
  __syncthreads();
}

__device__ __forceinline__ float4 mc_ld_reduce(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_sys(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// U = independent 16-byte vectors in flight per thread: 4 next to the backward GEMMs (40 registers: a block fits beside a
// GEMM CTA), 8 for the last bucket of a step, which runs alone and is latency-bound on the NVLink round trip.
template <bool MULTICAST, int U>
__global__ void __launch_bounds__(THREADS)
allreduce_kernel(const Args a) {
  // every rank's gradients of this bucket are final (written by kernels that precede this launch in their streams)
  rank_barrier(a);
  const int64_t n4 = a.count >> 2;                               // 16-byte vectors in the bucket
  const int64_t per = (n4 + a.world - 1) / a.world;
  const int64_t v0 = min(n4, (int64_t)a.rank * per), v1 = min(n4, v0 + per);   // this rank's slice
  const int64_t stride = (int64_t)gridDim.x * THREADS;
  int64_t i = v0 + (int64_t)blockIdx.x * THREADS + threadIdx.x;
  if constexpr (MULTICAST) {
    float* mc = reinterpret_cast<float*>(a.multicast_ptr) + a.offset;
    for (; i + (U - 1) * stride < v1; i += U * stride) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = mc_ld_reduce(mc + ((i + u * stride) << 2));
#pragma unroll
      for (int u = 0; u < U; ++u) {
        v[u].x *= a.scale; v[u].y *= a.scale; v[u].z *= a.scale; v[u].w *= a.scale;
        mc_st(mc + ((i + u * stride) << 2), v[u]);
      }
    }
    for (; i < v1; i += stride) {
      float4 v = mc_ld_reduce(mc + (i << 2));
      v.x *= a.scale; v.y *= a.scale; v.z *= a.scale; v.w *= a.scale;
      mc_st(mc + (i << 2), v);
    }
  } else {
    for (; i < v1; i += stride) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p = 0; p < a.world; ++p) {                       // rank order: every rank computes the same sum bit for bit
        const float4 v = ld_sys(reinterpret_cast<const float*>(a.buffer_ptrs[p]) + a.offset + (i << 2));
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      s.x *= a.scale; s.y *= a.scale; s.z *= a.scale; s.w *= a.scale;
      for (int p = 0; p < a.world; ++p) st_sys(reinterpret_cast<float*>(a.buffer_ptrs[p]) + a.offset + (i << 2), s);
    }
  }
  // every slice has been written everywhere before any rank's next kernel reads the bucket
  rank_barrier(a);
}

}  // namespace ar
}  // namespace vb

extern "C" {

int vb_allreduce_f32(const uint64_t* buffer_ptrs_dev, uint64_t multicast_ptr, const uint64_t* signal_pad_ptrs_dev,
                     int32_t rank, int32_t world, int64_t offset, int64_t count, float scale, int32_t blocks, void* stream) {
  using namespace vb::ar;
  VB_REQUIRE(buffer_ptrs_dev && signal_pad_ptrs_dev, VB_ERR_ARG, "allreduce: null pointer table");
  VB_REQUIRE(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world, VB_ERR_ARG, "allreduce: rank %d / world %d", rank, world);
  VB_REQUIRE(offset >= 0 && count >= 0 && (offset % 4) == 0 && (count % 4) == 0, VB_ERR_ARG,
             "allreduce: offset / count must be multiples of 4 elements (16 bytes)");
  if (count == 0) return 0;
  Args a;
  a.buffer_ptrs = buffer_ptrs_dev; a.multicast_ptr = multicast_ptr; a.signal_pad_ptrs = signal_pad_ptrs_dev;
  a.rank = rank; a.world = world; a.offset = offset; a.count = count; a.scale = scale;
  const int grid = std::max(1, std::min(blocks > 0 ? blocks : 148, MAX_BLOCKS));
  const bool deep = blocks > 148;   // the caller asks for more than one block per SM: nothing else is running
  if (multicast_ptr != 0) {
    if (deep) allreduce_kernel<true, 8><<<grid, THREADS, 0, (cudaStream_t)stream>>>(a);
    else allreduce_kernel<true, 4><<<grid, THREADS, 0, (cudaStream_t)stream>>>(a);
  } else {
    allreduce_kernel<false, 4><<<grid, THREADS, 0, (cudaStream_t)stream>>>(a);
  }
  return vb::check_launch("allreduce_kernel");
}

int vb_allreduce_pad_bytes(void) {
  return vb::ar::PAD_OFFSET_BYTES + vb::ar::MAX_BLOCKS * vb::ar::MAX_WORLD * 4;
}

}  // extern "C"
