// Gradient all-reduce over NVLink peer memory (data-parallel training), as plain load/store kernels on
// the flat gradient buffer of every rank -- the buffers are torch symmetric memory (CUDA VMM allocations mapped into every
// rank of the group, transformer.py::PeerReduce).
//
// Why not only NCCL: its all-reduce kernels need shared memory and whole SMs; the backward's persistent tcgen05 GEMMs own
// every SM (226 KB of shared memory each), so an NCCL bucket launched "under" the backward only gets SMs in the gaps
// between GEMMs and most of its time ends up exposed (2 x B200: +0.29 ms on a 4.44 ms step, and the exposure does not
// change with the bucket size, the gradient dtype or SMs set aside for NCCL -- profiles/r2_summary.md).  These kernels
// use NO shared memory and 64 threads x <= 140 registers per CTA (a GEMM CTA leaves 11.7 k registers of its SM free), one
// CTA per SM with 16 independent 16-byte loads in flight per thread (2.4 MB in flight chip-wide, NVLink's
// bandwidth-latency product): they are co-resident with the GEMM CTAs and stream the peers' slices while the tensor
// cores work.
//
// Two-shot, owner computes, race-free by construction.  For a bucket [lo, hi) cut into N chunks:
//   barrier                                   every rank's bucket is final (its backward has written it)
//   reduce-scatter   rank r:  own[chunk r] = sum_q peer_q[chunk r]       (fixed order q = 0..N-1: bit-identical everywhere)
//   barrier                                   every chunk is reduced at its owner
//   all-gather       rank r:  own[chunk q] = peer_q[chunk q]  for q != r
// plus one barrier after the last bucket of a step (nobody still reads a buffer that the next step zeroes).
// The barrier is its own one-warp kernel: lane q stores this launch's epoch into rank q's flag word (st.release.sys)
// and spins on its own flag word of rank q (ld.acquire.sys).  Epochs only grow, live in device memory (CUDA-graph
// replays keep counting) and a spin gives up after ~2 s of %globaltimer, raising an error word instead of hanging.
// Kernel boundaries separate the barrier from the transfers, so the peers' data is never cached across a barrier
// (L1 is invalidated at launch) and plain 16-byte loads / stores can be used.
#include "icap_common.cuh"

namespace {

constexpr int MAXR = 8;
struct Peers { void* p[MAXR]; };

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void p2p_barrier_kernel(Peers flags, int rank, int nranks, unsigned int* __restrict__ epoch, int* __restrict__ err) {
  pdl_prologue();                                // everything enqueued before on this stream has completed
  const int q = threadIdx.x;
  const unsigned int e = epoch[0] + 1u;
  __syncwarp();
  if (q < nranks) {
    __threadfence_system();
    unsigned int* theirs = reinterpret_cast<unsigned int*>(flags.p[q]) + rank;       // my word in rank q's flag array
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(e) : "memory");
    const unsigned int* mine = reinterpret_cast<const unsigned int*>(flags.p[rank]) + q;   // rank q's word in mine
    const unsigned long long t0 = gtime_ns();
    unsigned int seen;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
      if ((int)(seen - e) >= 0) break;
      if (gtime_ns() - t0 > 2000000000ull) { atomicExch(err, 1 + q); break; }         // a peer never arrived: report, do not hang
      __nanosleep(200);
    }
  }
  __syncwarp();
  if (q == 0) epoch[0] = e;
}

// own[i] = sum_q peer_q[i] for i in this rank's chunk [c0, c1) (float4 granularity).  NR = ranks (compile time: the
// loads of all ranks for U consecutive grid-stride elements are issued back to back -- NR * U 16-byte loads in flight per
// thread, which is what NVLink's ~3 us round trip needs to reach its bandwidth).
template <int NR, int U>
__global__ void __launch_bounds__(64)
p2p_reduce_scatter_kernel(Peers bufs, int rank, int64_t c0, int64_t c1) {
  pdl_prologue();
  float4* own = reinterpret_cast<float4*>(bufs.p[rank]);
  const int64_t n4 = (c1 - c0) >> 2, b4 = c0 >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * U) {
    float4 v[U][NR];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
#pragma unroll
      for (int q = 0; q < NR; ++q)
        if (i < n4) v[u][q] = reinterpret_cast<const float4*>(bufs.p[q])[b4 + i];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n4) {
        float4 a = v[u][0];
#pragma unroll
        for (int q = 1; q < NR; ++q) { a.x += v[u][q].x; a.y += v[u][q].y; a.z += v[u][q].z; a.w += v[u][q].w; }
        own[b4 + i] = a;
      }
    }
  }
}

// own[chunk q] = peer_q[chunk q] for every q != rank; chunk q = [lo + q*cs, min(hi, lo + (q+1)*cs))
template <int NR, int U>
__global__ void __launch_bounds__(64)
p2p_all_gather_kernel(Peers bufs, int rank, int64_t lo, int64_t hi, int64_t cs) {
  pdl_prologue();
  float4* own = reinterpret_cast<float4*>(bufs.p[rank]);
  const int64_t cs4 = cs >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < cs4; i0 += stride * U) {
    float4 v[U][NR];
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int q = 0; q < NR; ++q) {
        const int64_t e = lo + (int64_t)q * cs + 4 * (i0 + u * stride);
        if (q != rank && i0 + u * stride < cs4 && e < hi) v[u][q] = reinterpret_cast<const float4*>(bufs.p[q])[e >> 2];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int q = 0; q < NR; ++q) {
        const int64_t e = lo + (int64_t)q * cs + 4 * (i0 + u * stride);
        if (q != rank && i0 + u * stride < cs4 && e < hi) own[e >> 2] = v[u][q];
      }
    }
  }
}

// NVSwitch in-switch reduction (NVLS): the buffers of all ranks are bound to ONE multicast address range (torch symmetric
// memory sets it up).  `multimem.ld_reduce` on a multicast address returns the SUM over all ranks' copies, computed inside
// the switch; `multimem.st` writes all ranks' copies at once.  Rank r reduces + broadcasts chunk r of the bucket: each GPU
// pulls 1/N of the bucket (already reduced) and pushes 1/N -- an all-reduce in one pass with 2/N of the bucket per link
// direction instead of the ring's 2 (N-1)/N.  One barrier before (every rank's bucket is final), one after (every chunk
// has been broadcast).
template <int U>
__global__ void __launch_bounds__(128)
p2p_nvls_allreduce_kernel(float* __restrict__ mc, int64_t c0, int64_t c1) {
  pdl_prologue();
  const int64_t n4 = (c1 - c0) >> 2;
  float* base = mc + c0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n4)
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(base + 4 * i) : "memory");
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n4)
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(base + 4 * i), "f"(v[u].x),
                     "f"(v[u].y), "f"(v[u].z), "f"(v[u].w) : "memory");
    }
  }
}

}  // namespace

static int fill_peers(Peers& P, void* const* ptrs, int nranks) {
  for (int q = 0; q < MAXR; ++q) P.p[q] = q < nranks ? ptrs[q] : nullptr;
  for (int q = 0; q < nranks; ++q)
    if (!ptrs[q] || ((uintptr_t)ptrs[q] & 15)) return -1;
  return 0;
}

extern "C" int icap_p2p_barrier(void* const* flag_ptrs, int rank, int nranks, unsigned int* epoch, int* err, void* stream) {
  ICAP_ARG(flag_ptrs && epoch && err && nranks >= 1 && nranks <= MAXR && rank >= 0 && rank < nranks,
           "icap_p2p_barrier: bad argument (at most %d ranks)", MAXR);
  Peers P;
  ICAP_ARG(fill_peers(P, flag_ptrs, nranks) == 0, "icap_p2p_barrier: null / unaligned flag pointer");
  ICAP_CUDA(icap_launch(p2p_barrier_kernel, 1, 32, 0, (cudaStream_t)stream, P, rank, nranks, epoch, err));
  ICAP_LAUNCH_CHECK("icap_p2p_barrier");
  return 0;
}

// the two transfer phases of one bucket [lo, hi) (element offsets into the fp32 buffers, multiples of 4); the caller puts
// icap_p2p_barrier before each of them
extern "C" int icap_p2p_reduce_scatter(void* const* buf_ptrs, int rank, int nranks, int64_t lo, int64_t hi, int ctas,
                                       void* stream) {
  ICAP_ARG(buf_ptrs && nranks >= 1 && nranks <= MAXR && rank >= 0 && rank < nranks && lo >= 0 && hi > lo && lo % 4 == 0 &&
           hi % 4 == 0, "icap_p2p_reduce_scatter: bad argument");
  Peers P;
  ICAP_ARG(fill_peers(P, buf_ptrs, nranks) == 0, "icap_p2p_reduce_scatter: null / unaligned buffer pointer");
  const int64_t cs = (ceil_div64(hi - lo, nranks) + 3) / 4 * 4;
  const int64_t c0 = lo + rank * cs, c1 = c0 + cs < hi ? c0 + cs : hi;
  if (c1 <= c0) return 0;
  const unsigned grid = (unsigned)(ctas > 0 ? ctas : 148);
  cudaStream_t st = (cudaStream_t)stream;
  switch (nranks) {            // compile-time rank counts: NR * U = 16 loads of 16 bytes in flight per thread
    case 1: return 0;
    case 2: ICAP_CUDA(icap_launch(p2p_reduce_scatter_kernel<2, 8>, grid, 64, 0, st, P, rank, c0, c1)); break;
    case 4: ICAP_CUDA(icap_launch(p2p_reduce_scatter_kernel<4, 4>, grid, 64, 0, st, P, rank, c0, c1)); break;
    case 8: ICAP_CUDA(icap_launch(p2p_reduce_scatter_kernel<8, 2>, grid, 64, 0, st, P, rank, c0, c1)); break;
    default: icap_set_error("icap_p2p_reduce_scatter: 2, 4 or 8 ranks (got %d)", nranks); return -1;
  }
  ICAP_LAUNCH_CHECK("icap_p2p_reduce_scatter");
  return 0;
}

extern "C" int icap_p2p_all_gather(void* const* buf_ptrs, int rank, int nranks, int64_t lo, int64_t hi, int ctas,
                                   void* stream) {
  ICAP_ARG(buf_ptrs && nranks >= 1 && nranks <= MAXR && rank >= 0 && rank < nranks && lo >= 0 && hi > lo && lo % 4 == 0 &&
           hi % 4 == 0, "icap_p2p_all_gather: bad argument");
  Peers P;
  ICAP_ARG(fill_peers(P, buf_ptrs, nranks) == 0, "icap_p2p_all_gather: null / unaligned buffer pointer");
  if (nranks == 1) return 0;
  const int64_t cs = (ceil_div64(hi - lo, nranks) + 3) / 4 * 4;
  const unsigned grid = (unsigned)(ctas > 0 ? ctas : 148);
  cudaStream_t st = (cudaStream_t)stream;
  switch (nranks) {
    case 2: ICAP_CUDA(icap_launch(p2p_all_gather_kernel<2, 8>, grid, 64, 0, st, P, rank, lo, hi, cs)); break;
    case 4: ICAP_CUDA(icap_launch(p2p_all_gather_kernel<4, 4>, grid, 64, 0, st, P, rank, lo, hi, cs)); break;
    case 8: ICAP_CUDA(icap_launch(p2p_all_gather_kernel<8, 2>, grid, 64, 0, st, P, rank, lo, hi, cs)); break;
    default: icap_set_error("icap_p2p_all_gather: 2, 4 or 8 ranks (got %d)", nranks); return -1;
  }
  ICAP_LAUNCH_CHECK("icap_p2p_all_gather");
  return 0;
}

// One-pass all-reduce of [lo, hi) through the NVSwitch multicast mapping `mc_base` of the symmetric buffer (see the
// kernel).  Call sequence per bucket: icap_p2p_barrier; icap_p2p_allreduce_nvls; and icap_p2p_barrier once before the
// reduced values are read.
extern "C" int icap_p2p_allreduce_nvls(void* mc_base, int rank, int nranks, int64_t lo, int64_t hi, int ctas, void* stream) {
  ICAP_ARG(mc_base && ((uintptr_t)mc_base & 15) == 0 && nranks >= 1 && nranks <= MAXR && rank >= 0 && rank < nranks && lo >= 0 &&
           hi > lo && lo % 4 == 0 && hi % 4 == 0, "icap_p2p_allreduce_nvls: bad argument");
  const int64_t cs = (ceil_div64(hi - lo, nranks) + 3) / 4 * 4;
  const int64_t c0 = lo + rank * cs, c1 = c0 + cs < hi ? c0 + cs : hi;
  if (c1 <= c0) return 0;
  ICAP_CUDA(icap_launch(p2p_nvls_allreduce_kernel<8>, (unsigned)(ctas > 0 ? ctas : 148), 128, 0, (cudaStream_t)stream,
                        (float*)mc_base, c0, c1));
  ICAP_LAUNCH_CHECK("icap_p2p_allreduce_nvls");
  return 0;
}
