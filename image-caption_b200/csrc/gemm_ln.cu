// Projection + dropout + residual + LayerNorm in ONE tcgen05 kernel (bf16, sm_100a):
//
//   s = dropout(A[M,K] . W[N,K]^T + bias) + R          y = ((s - mean) * rstd * gamma + beta) * rowscale
//
// for the model widths N = d in {128, 256, 512, 1024}.  LayerNorm needs whole rows, a 128 x N fp32 accumulator per CTA
// would leave most SMs idle at these M, so the N columns of a 128-row block are split over a THREAD-BLOCK CLUSTER of
// CL = N / BN CTAs (BN = 128 or 256): every CTA runs its own TMA -> tcgen05.mma main loop on a 128 x BN tile
// (accumulator in TMEM), the epilogue threads (one row x BN/2 columns each, values kept in registers) reduce their row locally
// (mean, sum of squared deviations), publish the pair into the shared memory of every CTA of the cluster (DSMEM),
// and after one cluster barrier combine the 2*CL partials (Chan's parallel variance) and normalise.  The GEMM output
// never touches HBM: against icap_gemm + icap_add_ln_fwd this saves one launch, one write and one read of [M, N].
//
// Replaces `joint_linear -> Dropout -> LayerNorm(out + residual)` (modules.py:86-90) and
// `position_wise_2 -> Dropout -> LayerNorm(out + x)` [+ `*= non_pad_mask`] (modules.py:117-120,154-155,203-204).
// The dropout decisions are those of icap_add_ln_fwd (same counter hash, same element index), so
// icap_add_ln_bwd consumes the saved sum / mean / rstd unchanged.
#include <cuda.h>
#include <stdlib.h>
#include "icap_common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int A_TILE_BYTES = BM * BK * 2;               // 16 KB per k-block
constexpr int NTHREADS = 320, EPI_WARPS = 8;
constexpr int MAX_CL = 8;
constexpr int STATS_BYTES = MAX_CL * 2 * BM * 8;        // float2 {mean, M2} per (cluster rank, column half, row)
constexpr int PAR_BYTES = 3 * 256 * 4;                  // gamma | beta | bias slices of this CTA (BN <= 256)
// BN = columns per CTA: 128 (6-stage ring, more CTAs: small M) or 256 (4 stages, half the operand traffic per MAC)
template <int BN> struct LCfg {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STATS_BYTES + PAR_BYTES + 128 /*barriers*/ + 1024 /*align slack*/;
};

struct LnArgs {
  int M, N, K;
  const bf16* res;
  int64_t ldr;
  const float *bias, *gamma, *beta, *rowscale;
  bf16* y;
  int64_t ldy;
  bf16* s;             // nullable: pre-norm sum for the backward
  int64_t lds;
  float *mean, *rstd;  // nullable
  float eps, p_drop;
  uint32_t thresh;
  uint64_t seed;
  const int* seed_dev;
};

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_f2(uint32_t cluster_addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 t;
  __nv_bfloat162 h;
  h = __floats2bfloat162_rn(v[0], v[1]); t.x = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(v[2], v[3]); t.y = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(v[4], v[5]); t.z = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(v[6], v[7]); t.w = *reinterpret_cast<uint32_t*>(&h);
  return t;
}

template <int CL, int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmY,
               const __grid_constant__ CUtensorMap tmS, const LnArgs g) {
  using CF = LCfg<BN>;
  constexpr int STAGES = CF::STAGES, B_TILE_BYTES = CF::B_TILE_BYTES, STAGE_BYTES = CF::STAGE_BYTES;
  constexpr int NCOL = BN / 2;                                   // columns per epilogue thread
  constexpr int NB = NCOL / 64;                                  // 32-row x 64-column (4 KB, 128B-swizzled) boxes per warp
  constexpr int BOX = 4096, BPS = B_TILE_BYTES / BOX;            // boxes per B slot of the ring
  static_assert(2 * BPS == 8 * NB && STAGES >= 4 && STAGES * A_TILE_BYTES >= 8 * NB * BOX, "staging plan");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base, sB = sA + STAGES * A_TILE_BYTES;
  const uint32_t sStats = sB + STAGES * B_TILE_BYTES, sPar = sStats + STATS_BYTES;
  const uint32_t bars = sPar + PAR_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES, tfull_bar = bars + 16 * STAGES;
  const uint32_t rfull_bar = tfull_bar + 8, slot_addr = rfull_bar + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot_addr - smem_u32(smem_raw)));
  float* const par = reinterpret_cast<float*>(smem_raw + (sPar - smem_u32(smem_raw)));
  const float2* const stats = reinterpret_cast<const float2*>(smem_raw + (sStats - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;        // cluster (CL,1,1), gridDim.x == CL: rank = column slice
  const int n0 = (int)rank * BN, m0 = (int)blockIdx.y * BM;
  const int nkb = (g.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar + 8 * i, 1);
      mbar_init(empty_bar + 8 * i, 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(rfull_bar, 1);
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
    if (g.res) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
    if (g.s) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmS) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_addr), "n"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // phase 1 of the cluster barrier: "this CTA is running" -- waited for right before the first DSMEM store
  if constexpr (CL > 1) cluster_arrive();
  // PDL: the next kernel of the stream may become resident now (it waits for OUR completion itself); this kernel's own
  // grid dependency is resolved by the TMA producer below, AFTER it has requested the first stages of W -- the weights
  // (like bias / gamma / beta) are model parameters, never written by the kernel launched just before
  pdl_launch_dependents();

  float v[NCOL];     // epilogue threads: one row x NCOL columns of s
  const int q = warp & 3, half = warp >= 6 ? 1 : 0;              // TMEM lane quarter / column half of an epilogue warp
  const int r_in = q * 32 + lane, row = m0 + r_in;
  const int c0 = half * NCOL;                                   // first column inside this CTA's BN
  const bool live = warp >= 2 && row < g.M;

  // Staging plan (all boxes 32 rows x 64 bf16 columns, 128B-swizzled, 4 KB).  Box id b = (half * NB + j) * 4 + q belongs
  // to the epilogue warp (q, half), column box j.  The residual tile is fetched by the TMA producer as two extra "stages"
  // of the operand ring -- ring slots nkb, nkb+1 (B part), freed by the MMAs of k-blocks nkb-STAGES.. -- so it lands
  // while the last k-blocks are still being multiplied.  After the accumulator is complete the whole ring is free:
  // y is staged in the A part, the saved sum in the B parts of ring slots nkb+2, nkb+3.
  auto res_box = [&](int b) { return sB + (uint32_t)(((nkb + b / BPS) % STAGES) * B_TILE_BYTES + (b % BPS) * BOX); };
  auto sum_box = [&](int b) { return sB + (uint32_t)(((nkb + 2 + b / BPS) % STAGES) * B_TILE_BYTES + (b % BPS) * BOX); };
  auto y_box = [&](int b) { return sA + (uint32_t)(b * BOX); };

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      int s = 0, ph = 0;
      const int npre = nkb < STAGES ? nkb : STAGES;
      for (int i = 0; i < npre; ++i) {                                             // W tiles of the first ring round
        mbar_expect_tx(full_bar + 8 * i, STAGE_BYTES);
        tma_load_2d(sB + i * B_TILE_BYTES, &tmB, i * BK, n0, full_bar + 8 * i);    // box {64 k, BN n}
      }
      pdl_wait();                                                                  // A / residual come from the predecessor
      for (int i = 0; i < nkb; ++i) {
        if (i >= npre) {
          mbar_wait(empty_bar + 8 * s, ph ^ 1);
          mbar_expect_tx(full_bar + 8 * s, STAGE_BYTES);
          tma_load_2d(sB + s * B_TILE_BYTES, &tmB, i * BK, n0, full_bar + 8 * s);
        }
        tma_load_2d(sA + s * A_TILE_BYTES, &tmA, i * BK, m0, full_bar + 8 * s);    // box {64 k, 128 m}
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      if (g.res) {
        mbar_expect_tx(rfull_bar, BM * BN * 2);
        for (int e = 0; e < 2; ++e) {
          mbar_wait(empty_bar + 8 * s, ph ^ 1);                                    // ring slot nkb + e is free
          for (int bb = 0; bb < BPS; ++bb) {
            const int b = e * BPS + bb, qq = b & 3, hj = b >> 2;                   // hj = half * NB + j
            tma_load_2d(res_box(b), &tmR, n0 + hj * 64, m0 + qq * 32, rfull_bar);  // box {64 cols, 32 rows}
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer: D=f32, A=B=bf16 K-major, 128 x BN x 16
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int s = 0, ph = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(full_bar + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t aS = sA + s * A_TILE_BYTES, bS = sB + s * B_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_bf16(tmem_base, make_sdesc(aS + k * 32, 16, 1024), make_sdesc(bS + k * 32, 16, 1024), idesc,
                    (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(empty_bar + 8 * s);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit(tfull_bar);
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- epilogue, part 1: s and its local statistics
    for (int t = (int)threadIdx.x - 64; t < BN; t += 32 * EPI_WARPS) {    // stage gamma | beta | bias of this column slice
      par[t] = __ldg(g.gamma + n0 + t);
      par[BN + t] = __ldg(g.beta + n0 + t);
      par[2 * BN + t] = g.bias ? __ldg(g.bias + n0 + t) : 0.f;
    }
    const uint32_t sw = (uint32_t)(lane & 7);                     // 128B swizzle: 16-byte chunk c of row r sits at c ^ (r & 7)
    const uint32_t my_row = (uint32_t)lane * 128u;
    const int row0 = m0 + q * 32;
    asm volatile("bar.sync 1, 256;" ::: "memory");                // the staged parameters are visible to all epilogue warps
    mbar_wait(tfull_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
#pragma unroll
    for (int c = 0; c < NCOL / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tacc + (uint32_t)(32 * c), r);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[32 * c + j] = __uint_as_float(r[j]);
    }
    if (g.bias) {
#pragma unroll
      for (int j = 0; j < NCOL; j += 4) {
        const float4 b = lds_f4(sPar + (uint32_t)(2 * BN + c0 + j) * 4);
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
    if (g.p_drop > 0.f) {
      uint64_t seed = g.seed;
      if (g.seed_dev) seed += (uint64_t)(*g.seed_dev) * 0x9E3779B97F4A7C15ull;
      const uint32_t sf = seed_fold(seed);
      const float keep_scale = 1.f / (1.f - g.p_drop);
      const uint32_t e8 = (uint32_t)row * (uint32_t)(g.N >> 3) + (uint32_t)((n0 + c0) >> 3);
#pragma unroll
      for (int j = 0; j < NCOL / 8; ++j)
        dropout_apply8(*reinterpret_cast<float(*)[8]>(&v[8 * j]), sf, e8 + (uint32_t)j, g.thresh, keep_scale);
    }
    if (g.res) {
      mbar_wait(rfull_bar, 0);                                    // the residual tile has landed (rows >= M: zero filled)
#pragma unroll
      for (int jb = 0; jb < NB; ++jb) {
        const uint32_t box = res_box((half * NB + jb) * 4 + q) + my_row;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          uint4 rj;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rj.x), "=r"(rj.y), "=r"(rj.z), "=r"(rj.w)
                       : "r"(box + (((uint32_t)t ^ sw) << 4)));
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rj);
          const int c = jb * 64 + t * 8;
#pragma unroll
          for (int i = 0; i < 4; ++i) { v[c + 2 * i] += __low2float(h[i]); v[c + 2 * i + 1] += __high2float(h[i]); }
        }
      }
    }
    if (g.s) {                                                    // saved pre-norm sum: registers -> swizzled box -> TMA store
#pragma unroll
      for (int jb = 0; jb < NB; ++jb) {
        const uint32_t box = sum_box((half * NB + jb) * 4 + q);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint4 o = pack8(v + jb * 64 + t * 8);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(box + my_row + (((uint32_t)t ^ sw) << 4)), "r"(o.x),
                       "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && row0 < g.M) {
          tma_store_2d(&tmS, box, n0 + c0 + jb * 64, row0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NCOL; ++j) sum += v[j];
    const float lmean = sum * (1.f / (float)NCOL);
    float m2 = 0.f;
#pragma unroll
    for (int j = 0; j < NCOL; ++j) { const float c = v[j] - lmean; m2 += c * c; }
    // publish {mean, M2} of this NCOL-column piece into slot (rank, half) of EVERY CTA of the cluster
    const uint32_t slot = sStats + (uint32_t)(((int)rank * 2 + half) * BM + r_in) * 8u;
    if constexpr (CL > 1) {
      cluster_wait();                                             // every CTA of the cluster has started
#pragma unroll
      for (int dst = 0; dst < CL; ++dst) st_cluster_f2(map_to_cta(slot, (uint32_t)dst), lmean, m2);
    } else {
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(slot), "f"(lmean), "f"(m2) : "memory");
    }
  }
  if constexpr (CL > 1) {
    if (warp < 2) cluster_wait();                                 // phase 1 (the epilogue warps waited above)
    cluster_arrive();                                             // phase 2: all partial statistics are published
    cluster_wait();
  } else {
    __syncthreads();
  }
  if (warp >= 2) {
    // ---------------------------------------------------------------- epilogue, part 2: combine, normalise, store
    float mp[2 * CL], mean = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * CL; ++i) {
      const float2 p = stats[i * BM + r_in];
      mp[i] = p.x;
      mean += p.x;
      m2 += p.y;
    }
    mean *= 1.f / (float)(2 * CL);
#pragma unroll
    for (int i = 0; i < 2 * CL; ++i) { const float dlt = mp[i] - mean; m2 += (float)NCOL * dlt * dlt; }
    const float rstd = rsqrtf(m2 / (float)(BN * CL) + g.eps);
    if (live && g.mean && rank == 0 && half == 0) { g.mean[row] = mean; g.rstd[row] = rstd; }
    {
      const float rs = (live && g.rowscale) ? __ldg(g.rowscale + row) : 1.f;
      const uint32_t sw = (uint32_t)(lane & 7), my_row = (uint32_t)lane * 128u;
      const int row0 = m0 + q * 32;
#pragma unroll
      for (int j = 0; j < NCOL / 8; ++j) {
        const float4 g0 = lds_f4(sPar + (uint32_t)(c0 + 8 * j) * 4), g1 = lds_f4(sPar + (uint32_t)(c0 + 8 * j + 4) * 4);
        const float4 b0 = lds_f4(sPar + (uint32_t)(BN + c0 + 8 * j) * 4), b1 = lds_f4(sPar + (uint32_t)(BN + c0 + 8 * j + 4) * 4);
        float o[8];
        o[0] = ((v[8 * j + 0] - mean) * rstd * g0.x + b0.x) * rs;
        o[1] = ((v[8 * j + 1] - mean) * rstd * g0.y + b0.y) * rs;
        o[2] = ((v[8 * j + 2] - mean) * rstd * g0.z + b0.z) * rs;
        o[3] = ((v[8 * j + 3] - mean) * rstd * g0.w + b0.w) * rs;
        o[4] = ((v[8 * j + 4] - mean) * rstd * g1.x + b1.x) * rs;
        o[5] = ((v[8 * j + 5] - mean) * rstd * g1.y + b1.y) * rs;
        o[6] = ((v[8 * j + 6] - mean) * rstd * g1.z + b1.z) * rs;
        o[7] = ((v[8 * j + 7] - mean) * rstd * g1.w + b1.w) * rs;
        const uint4 o4 = pack8(o);
        const uint32_t box = y_box((half * NB + j / 8) * 4 + q);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(box + my_row + (((uint32_t)(j & 7) ^ sw) << 4)),
                     "r"(o4.x), "r"(o4.y), "r"(o4.z), "r"(o4.w) : "memory");
        if ((j & 7) == 7) {                                       // one 64-column box complete: hand it to the TMA engine
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0 && row0 < g.M) {
            tma_store_2d(&tmY, box, n0 + c0 + (j / 8) * 64, row0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
}


// ------------------------------------------------------------------------------------------------------------------
// Small-footprint variant for the latency-bound decode steps (no dropout, no saved sum / statistics): same cluster
// scheme (CL CTAs x 128 columns per 128-row block, statistics through DSMEM), but built like gemm_small.cu so that
// TWO CTAs fit an SM and consecutive launches of a stream overlap: 192 threads (warp 0 TMA, warp 1 MMA, warps 2-5
// epilogue), 3-stage 32 KB ring, 128 TMEM columns, `griddepcontrol.launch_dependents` right after the prologue and the
// first W stages fetched BEFORE `griddepcontrol.wait` (W, bias, gamma, beta are model parameters).  The epilogue makes
// TWO passes over the accumulator, which simply stays in TMEM: pass 1 row statistics (Chan-combined per 32-column
// chunk), pass 2 normalise + store -- 80 registers instead of a 64-value row slice per thread.
constexpr int S_STAGES = 3, S_BN = 128, S_THREADS = 192;
constexpr int S_STAGE_BYTES = A_TILE_BYTES + S_BN * BK * 2;
constexpr int S_STATS_BYTES = MAX_CL * BM * 8;
constexpr int S_PAR_BYTES = 3 * S_BN * 4;
constexpr int S_SMEM_BYTES = S_STAGES * S_STAGE_BYTES + S_STATS_BYTES + S_PAR_BYTES + 128 + 1024;

template <int CL>
__global__ void __launch_bounds__(S_THREADS, 2)
gemm_ln_small_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmY, const LnArgs g) {
  constexpr int BN = S_BN, STAGES = S_STAGES, B_TILE_BYTES = BN * BK * 2, STAGE_BYTES = S_STAGE_BYTES, BOX = 4096;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base, sB = sA + STAGES * A_TILE_BYTES;
  const uint32_t sStats = sB + STAGES * B_TILE_BYTES, sPar = sStats + S_STATS_BYTES, bars = sPar + S_PAR_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES, tfull_bar = bars + 16 * STAGES;
  const uint32_t rfull_bar = tfull_bar + 8, slot_addr = rfull_bar + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot_addr - smem_u32(smem_raw)));
  const float2* const stats = reinterpret_cast<const float2*>(smem_raw + (sStats - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
  const int n0 = (int)rank * BN, m0 = (int)blockIdx.y * BM;
  const int nkb = (g.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
    if (g.res) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar + 8 * i, 1);
      mbar_init(empty_bar + 8 * i, 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(rfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_addr), "n"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  if constexpr (CL > 1) cluster_arrive();          // phase 1: "this CTA is running" (waited for before the first DSMEM store)
  pdl_launch_dependents();

  // residual: ring "k-blocks" nkb, nkb+1 (B parts); box b = jb * 4 + q (jb: 64-column box, q: 32-row quarter)
  auto res_box = [&](int b) { return sB + (uint32_t)(((nkb + b / 4) % STAGES) * B_TILE_BYTES + (b % 4) * BOX); };
  auto y_box = [&](int b) { return sA + (uint32_t)(b * BOX); };

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      const int npre = nkb < STAGES ? nkb : STAGES;
      for (int i = 0; i < npre; ++i) {                // weights first: they do not depend on the preceding kernel
        mbar_expect_tx(full_bar + 8 * i, STAGE_BYTES);
        tma_load_2d(sB + i * B_TILE_BYTES, &tmB, i * BK, n0, full_bar + 8 * i);
      }
      pdl_wait();
      for (int i = 0; i < npre; ++i) tma_load_2d(sA + i * A_TILE_BYTES, &tmA, i * BK, m0, full_bar + 8 * i);
      for (int i = npre; i < nkb; ++i) {
        const int s = i % STAGES, r = i / STAGES;
        mbar_wait(empty_bar + 8 * s, (uint32_t)((r & 1) ^ 1));
        mbar_expect_tx(full_bar + 8 * s, STAGE_BYTES);
        tma_load_2d(sA + s * A_TILE_BYTES, &tmA, i * BK, m0, full_bar + 8 * s);
        tma_load_2d(sB + s * B_TILE_BYTES, &tmB, i * BK, n0, full_bar + 8 * s);
      }
      if (g.res) {
        mbar_expect_tx(rfull_bar, BM * BN * 2);
        for (int e = 0; e < 2; ++e) {
          const int i = nkb + e, s = i % STAGES, r = i / STAGES;
          mbar_wait(empty_bar + 8 * s, (uint32_t)((r & 1) ^ 1));
          for (int bb = 0; bb < 4; ++bb)
            tma_load_2d(res_box(e * 4 + bb), &tmR, n0 + e * 64, m0 + bb * 32, rfull_bar);     // box {64 cols, 32 rows}
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES, r = i / STAGES;
        mbar_wait(full_bar + 8 * s, (uint32_t)(r & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t aS = sA + s * A_TILE_BYTES, bS = sB + s * B_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_bf16(tmem_base, make_sdesc(aS + k * 32, 16, 1024), make_sdesc(bS + k * 32, 16, 1024), idesc,
                    (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(empty_bar + 8 * s);
      }
      umma_commit(tfull_bar);
    }
    __syncwarp();
  }

  const int q = warp & 3, r_in = q * 32 + lane, row = m0 + r_in;
  const uint32_t sw = (uint32_t)(lane & 7), my_row = (uint32_t)lane * 128u;
  const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16);
  // one 32-column chunk of s = acc + bias + residual for this thread's row
  auto load_chunk = [&](int c, float (&v)[32]) {
    uint32_t r[32];
    tmem_ld32(tacc + (uint32_t)(32 * c), r);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (g.bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = lds_f4(sPar + (uint32_t)(2 * BN + 32 * c + j) * 4);
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
    if (g.res) {
      const uint32_t box = res_box((c >> 1) * 4 + q) + my_row;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        uint4 rj;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rj.x), "=r"(rj.y), "=r"(rj.z), "=r"(rj.w)
                     : "r"(box + ((((uint32_t)((c & 1) * 4 + t)) ^ sw) << 4)));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rj);
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[t * 8 + 2 * i] += __low2float(h[i]); v[t * 8 + 2 * i + 1] += __high2float(h[i]); }
      }
    }
  };

  if (warp >= 2) {
    // ---------------------------------------------------------------- epilogue pass 1: row statistics of this CTA's 128 columns
    const int t128 = (int)threadIdx.x - 64;
    {
      float* const par = reinterpret_cast<float*>(smem_raw + (sPar - smem_u32(smem_raw)));
      par[t128] = __ldg(g.gamma + n0 + t128);
      par[BN + t128] = __ldg(g.beta + n0 + t128);
      par[2 * BN + t128] = g.bias ? __ldg(g.bias + n0 + t128) : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    mbar_wait(tfull_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (g.res) mbar_wait(rfull_bar, 0);
    float mean = 0.f, m2 = 0.f;
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
      float v[32];
      load_chunk(c, v);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) sum += v[j];
      const float cm = sum * (1.f / 32.f);
      float cm2 = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) { const float dlt = v[j] - cm; cm2 += dlt * dlt; }
      // Chan: combine (32*c values, mean, m2) with (32 values, cm, cm2)
      const float dlt = cm - mean, nn = (float)(32 * (c + 1));
      mean += dlt * (32.f / nn);
      m2 += cm2 + dlt * dlt * ((float)(32 * c) * 32.f / nn);
    }
    const uint32_t slot = sStats + (uint32_t)((int)rank * BM + r_in) * 8u;
    if constexpr (CL > 1) {
      cluster_wait();                                             // every CTA of the cluster has started
#pragma unroll
      for (int dst = 0; dst < CL; ++dst) st_cluster_f2(map_to_cta(slot, (uint32_t)dst), mean, m2);
    } else {
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(slot), "f"(mean), "f"(m2) : "memory");
    }
  }
  if constexpr (CL > 1) {
    if (warp < 2) cluster_wait();
    cluster_arrive();                                             // phase 2: all partial statistics are published
    cluster_wait();
  } else {
    __syncthreads();
  }
  if (warp >= 2) {
    // ---------------------------------------------------------------- epilogue pass 2: combine, normalise, store
    float mp[CL], mean = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < CL; ++i) {
      const float2 p = stats[i * BM + r_in];
      mp[i] = p.x;
      mean += p.x;
      m2 += p.y;
    }
    mean *= 1.f / (float)CL;
#pragma unroll
    for (int i = 0; i < CL; ++i) { const float dlt = mp[i] - mean; m2 += (float)BN * dlt * dlt; }
    const float rstd = rsqrtf(m2 / (float)(BN * CL) + g.eps);
    const float rs = (row < g.M && g.rowscale) ? __ldg(g.rowscale + row) : 1.f;
    const int row0 = m0 + q * 32;
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
      float v[32];
      load_chunk(c, v);
      const uint32_t box = y_box((c >> 1) * 4 + q);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float4 g0 = lds_f4(sPar + (uint32_t)(32 * c + 8 * t) * 4), g1 = lds_f4(sPar + (uint32_t)(32 * c + 8 * t + 4) * 4);
        const float4 b0 = lds_f4(sPar + (uint32_t)(BN + 32 * c + 8 * t) * 4), b1 = lds_f4(sPar + (uint32_t)(BN + 32 * c + 8 * t + 4) * 4);
        float o[8];
        o[0] = ((v[8 * t + 0] - mean) * rstd * g0.x + b0.x) * rs;
        o[1] = ((v[8 * t + 1] - mean) * rstd * g0.y + b0.y) * rs;
        o[2] = ((v[8 * t + 2] - mean) * rstd * g0.z + b0.z) * rs;
        o[3] = ((v[8 * t + 3] - mean) * rstd * g0.w + b0.w) * rs;
        o[4] = ((v[8 * t + 4] - mean) * rstd * g1.x + b1.x) * rs;
        o[5] = ((v[8 * t + 5] - mean) * rstd * g1.y + b1.y) * rs;
        o[6] = ((v[8 * t + 6] - mean) * rstd * g1.z + b1.z) * rs;
        o[7] = ((v[8 * t + 7] - mean) * rstd * g1.w + b1.w) * rs;
        const uint4 o4 = pack8(o);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(box + my_row + ((((uint32_t)((c & 1) * 4 + t)) ^ sw) << 4)),
                     "r"(o4.x), "r"(o4.y), "r"(o4.z), "r"(o4.w) : "memory");
      }
      if (c & 1) {                                                // one 64-column box complete: hand it to the TMA engine
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && row0 < g.M) {
          tma_store_2d(&tmY, box, n0 + (c >> 1) * 64, row0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
}

template <int CL>
int launch_ln_small(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tr, const CUtensorMap& ty,
                    const LnArgs& g, cudaStream_t st) {
  static bool attr_done = false;
  auto kern = gemm_ln_small_kernel<CL>;
  if (!attr_done) {
    ICAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM_BYTES));
    attr_done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)CL, (unsigned)ceil_div64(g.M, BM));
  cfg.blockDim = dim3(S_THREADS);
  cfg.dynamicSmemBytes = S_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CL > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CL;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (icap_g_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  ICAP_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tr, ty, g));
  return 0;
}

template <int CL, int BN>
int launch_ln(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tr, const CUtensorMap& ty,
              const CUtensorMap& ts, const LnArgs& g, cudaStream_t st) {
  static bool attr_done = false;
  auto kern = gemm_ln_kernel<CL, BN>;
  constexpr int SMEM_BYTES = LCfg<BN>::SMEM_BYTES;
  if (!attr_done) {
    ICAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)CL, (unsigned)ceil_div64(g.M, BM));
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CL > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CL;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (icap_g_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  ICAP_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tr, ty, ts, g));
  return 0;
}

}  // namespace

extern "C" int icap_gemm_ln(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* W, int64_t ldw,
                            const float* bias, const void* res, int64_t ldr, const float* gamma, const float* beta,
                            const float* rowscale, void* y, int64_t ldy, void* sum_out, int64_t lds, float* mean_out,
                            float* rstd_out, float eps, float p_drop, uint64_t seed, const int* seed_dev,
                            void* stream) {
  ICAP_ARG(M > 0 && K > 0 && A && W && gamma && beta && y, "icap_gemm_ln: null/empty argument");
  ICAP_ARG((mean_out == nullptr) == (rstd_out == nullptr), "icap_gemm_ln: mean_out and rstd_out go together");
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  const bool ok = (N == 128 || N == 256 || N == 512 || N == 1024) && lda % 8 == 0 && ldw % 8 == 0 && ldy % 8 == 0 &&
                  (res == nullptr || ldr % 8 == 0) && (sum_out == nullptr || lds % 8 == 0) && al16(A) && al16(W) &&
                  al16(y) && al16(res) && al16(sum_out) && M * (N / 8) < (1ll << 30) && p_drop >= 0.f && p_drop < 1.f;
  if (!ok) {
    icap_set_error("icap_gemm_ln: unsupported shape/alignment (N=%lld must be 128/256/512/1024, 16-byte aligned rows)",
                   (long long)N);
    return -2;       // the caller falls back to icap_gemm + icap_add_ln_fwd (same arithmetic, two launches)
  }
  // Columns per CTA: 128 (cluster of N/128) while all clusters of the launch fit the GPU at once -- more CTAs, shorter
  // main loops, the right shape for the small-M decode steps; otherwise 256 (cluster of N/256), one wave for the
  // training row counts.  ICAP_GEMM_LN_BN=128|256 forces it.
  const int64_t tiles_m = ceil_div64(M, BM);
  int bn = (N == 128 || tiles_m * (N / 128) <= (int64_t)(icap_num_sms() * 3) / 4) ? 128 : 256;
  static IcapEnv e_bn;
  if (const char* e = e_bn.get("ICAP_GEMM_LN_BN")) { const int f = atoi(e); if ((f == 128 || f == 256) && N % f == 0) bn = f; }
  // inference form at latency-bound row counts: optionally the small-footprint kernel (two CTAs per SM).  Measured on
  // B200 (beam-5 decode of 512 images): 14.8 ms against 13.8 ms with the 320-thread kernel below -- its two-pass,
  // four-warp epilogue sits on the critical path of every launch -- so it is OFF by default (ICAP_GEMM_LN_SMALL=1
  // automatic, =2 every eligible call).
  static IcapEnv e_small;
  const int small_mode = e_small.geti("ICAP_GEMM_LN_SMALL", 0);
  const bool small = small_mode != 0 && sum_out == nullptr && mean_out == nullptr && p_drop == 0.f && N / 128 <= MAX_CL &&
                     (small_mode == 2 || tiles_m * (N / 128) <= (int64_t)icap_num_sms() * 9 / 4);
  if (small) bn = 128;
  CUtensorMap ta, tb;
  int rc;
  if ((rc = icap_make_tmap_2d(&ta, A, M, K, lda, BM, ICAP_BF16))) return rc;
  if ((rc = icap_make_tmap_2d(&tb, W, N, K, ldw, bn, ICAP_BF16))) return rc;
  CUtensorMap ty, tr, ts;                                        // 32-row x 64-column boxes of the [M, N] tensors
  if ((rc = icap_make_tmap_2d(&ty, y, M, N, ldy, 32, ICAP_BF16))) return rc;
  tr = ty; ts = ty;
  if (res && (rc = icap_make_tmap_2d(&tr, res, M, N, ldr, 32, ICAP_BF16))) return rc;
  if (sum_out && (rc = icap_make_tmap_2d(&ts, sum_out, M, N, lds, 32, ICAP_BF16))) return rc;
  LnArgs g;
  g.M = (int)M; g.N = (int)N; g.K = (int)K;
  g.res = (const bf16*)res; g.ldr = ldr;
  g.bias = bias; g.gamma = gamma; g.beta = beta; g.rowscale = rowscale;
  g.y = (bf16*)y; g.ldy = ldy; g.s = (bf16*)sum_out; g.lds = lds;
  g.mean = mean_out; g.rstd = rstd_out;
  g.eps = eps; g.p_drop = p_drop; g.thresh = dropout_threshold(p_drop); g.seed = seed; g.seed_dev = seed_dev;
  cudaStream_t st = (cudaStream_t)stream;
  if (small) {
    switch (N / 128) {
      case 1: rc = launch_ln_small<1>(ta, tb, tr, ty, g, st); break;
      case 2: rc = launch_ln_small<2>(ta, tb, tr, ty, g, st); break;
      case 4: rc = launch_ln_small<4>(ta, tb, tr, ty, g, st); break;
      default: rc = launch_ln_small<8>(ta, tb, tr, ty, g, st); break;
    }
  } else if (bn == 128) {
    switch (N / 128) {
      case 1: rc = launch_ln<1, 128>(ta, tb, tr, ty, ts, g, st); break;
      case 2: rc = launch_ln<2, 128>(ta, tb, tr, ty, ts, g, st); break;
      case 4: rc = launch_ln<4, 128>(ta, tb, tr, ty, ts, g, st); break;
      default: rc = launch_ln<8, 128>(ta, tb, tr, ty, ts, g, st); break;
    }
  } else {
    switch (N / 256) {
      case 1: rc = launch_ln<1, 256>(ta, tb, tr, ty, ts, g, st); break;
      case 2: rc = launch_ln<2, 256>(ta, tb, tr, ty, ts, g, st); break;
      default: rc = launch_ln<4, 256>(ta, tb, tr, ty, ts, g, st); break;
    }
  }
  if (rc) return rc;
  ICAP_LAUNCH_CHECK("icap_gemm_ln");
  return 0;
}
