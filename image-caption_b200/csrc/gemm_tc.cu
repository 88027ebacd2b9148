// bf16 tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory ->
// tcgen05.mma (cta_group::1, kind::f16, 128x128x16) with the fp32 accumulator in TMEM ->
// tcgen05.ld epilogue (bias / ReLU / ReLU-mask / accumulate / split-K reduction).
//
//   C[M,N] (+)= op(A)[M,K] . op(B)[K,N]
//
// Same operand conventions as gemm_simt.cu (a_kmajor / b_kmajor); MN-major operands are fed to the
// tensor core directly through the UMMA descriptor major bits, so dgrad / wgrad need no transposes.
//
// CTA = 192 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (each owns the TMEM lane quarter warp_id % 4).  3-stage 32 KB ring so two
// CTAs are resident per SM: one CTA's epilogue overlaps the other's main loop.
//
// Replaces the reference's nn.Linear / torch.matmul calls (modules.py:72-77,86,113-116;
// model.py:93,295-306,433) in bf16 mode.
#include <cuda.h>
#include <stdlib.h>
#include "icap_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 3, UMMA_K = 16;
constexpr int TILE_BYTES = BM * BK * 2;          // 16 KB per operand per stage
constexpr int TMEM_COLS = 128;
constexpr int NTHREADS = 192;
constexpr int SMEM_BYTES = STAGES * 2 * TILE_BYTES + 1024 /*align slack*/ + 128 /*barriers*/;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (UMMA), SWIZZLE_128B, version 1
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;   // descriptor version (Blackwell)
  d |= 2ull << 61;   // SWIZZLE_128B
  return d;
}

template <typename TO> struct OutVec;
template <> struct OutVec<float> {
  static __device__ __forceinline__ void store8(float* p, const float* v) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  static __device__ __forceinline__ void load8(const float* p, float* v) {
    float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};
template <> struct OutVec<bf16> {
  static __device__ __forceinline__ void store8(bf16* p, const float* v) {
    uint4 t;
    __nv_bfloat162 h;
    h = __floats2bfloat162_rn(v[0], v[1]); t.x = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(v[2], v[3]); t.y = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(v[4], v[5]); t.z = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(v[6], v[7]); t.w = *reinterpret_cast<uint32_t*>(&h);
    *reinterpret_cast<uint4*>(p) = t;
  }
  static __device__ __forceinline__ void load8(const bf16* p, float* v) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
  }
};

template <bool A_KMAJOR, bool B_KMAJOR, typename TO>
__global__ void __launch_bounds__(NTHREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmX, int M, int N, int K,
               TO* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int epi, const TO* __restrict__ aux,
               int64_t ldaux, int accumulate, int kb_per_split, int vec_ok, int store_mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base, sB = smem_base + STAGES * TILE_BYTES;
  const uint32_t bars = sB + STAGES * TILE_BYTES;        // full[S], empty[S], tmem_full, tmem slot
  const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES, tfull_bar = bars + 16 * STAGES;
  const uint32_t slot_addr = tfull_bar + 8;
  const uint32_t aux_bar = tfull_bar + 16;                // 4 x 8 B: one per epilogue warp (aux tile loads)
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (slot_addr - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nkb_total = (K + BK - 1) / BK;
  const int kb0 = blockIdx.z * kb_per_split;
  const int nkb = min(kb_per_split, nkb_total - kb0);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar + 8 * i, 1);
      mbar_init(empty_bar + 8 * i, 1);
    }
    mbar_init(tfull_bar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(aux_bar + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_addr), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES, ph = (i / STAGES) & 1;
        mbar_wait(empty_bar + 8 * s, ph ^ 1);
        mbar_expect_tx(full_bar + 8 * s, 2 * TILE_BYTES);
        const int k0 = (kb0 + i) * BK;
        const uint32_t dA = sA + s * TILE_BYTES, dB = sB + s * TILE_BYTES;
        if (A_KMAJOR) tma_load_2d(dA, &tmA, k0, m0, full_bar + 8 * s);           // box {64 k, 128 m}
        else {                                                                    // 2 boxes {64 m, 64 k}
          tma_load_2d(dA, &tmA, m0, k0, full_bar + 8 * s);
          tma_load_2d(dA + TILE_BYTES / 2, &tmA, m0 + 64, k0, full_bar + 8 * s);
        }
        if (B_KMAJOR) tma_load_2d(dB, &tmB, k0, n0, full_bar + 8 * s);
        else {
          tma_load_2d(dB, &tmB, n0, k0, full_bar + 8 * s);
          tma_load_2d(dB + TILE_BYTES / 2, &tmB, n0 + 64, k0, full_bar + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer (one thread)
      // instruction descriptor: D=f32, A=B=bf16, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_KMAJOR ? 0u : 1u) << 15) |
                             ((B_KMAJOR ? 0u : 1u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES, ph = (i / STAGES) & 1;
        mbar_wait(full_bar + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t aS = sA + s * TILE_BYTES, bS = sB + s * TILE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // K-major : 8-row groups 1024 B apart (SBO), +32 B per 16-element k step inside the swizzle atom
          // MN-major: 64-element MN atoms 8192 B apart (LBO), 8-k groups 1024 B apart (SBO), +2048 B per k step
          const uint64_t ad = A_KMAJOR ? make_sdesc(aS + k * 32, 16, 1024) : make_sdesc(aS + k * 2048, 8192, 1024);
          const uint64_t bd = B_KMAJOR ? make_sdesc(bS + k * 32, 16, 1024) : make_sdesc(bS + k * 2048, 8192, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar + 8 * s);      // smem slot free once these MMAs have read it
      }
      umma_commit(tfull_bar);                // accumulator complete
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    mbar_wait(tfull_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const bool add_bias = (bias != nullptr) && (blockIdx.z == 0);
    if (store_mode != 0) {
      // ---- staged epilogue: TMEM -> registers -> 128B-swizzled smem slab (the drained pipeline stages are
      // reused) -> TMA store / TMA reduce-add.  One 32-row slab per warp, so no cross-warp barrier is needed.
      constexpr int ESZ = (int)sizeof(TO);
      constexpr int COLS_PER_BOX = 128 / ESZ;               // 64 bf16 or 32 fp32 columns = one 128 B swizzle row
      constexpr int NBOX = BN / COLS_PER_BOX;               // 2 or 4 boxes of 32 rows x 128 B = 4 KB
      const uint32_t slab = sA + (uint32_t)q * (NBOX * 4096);
      const uint32_t my_row = slab + (uint32_t)lane * 128;
      const uint32_t sw = (uint32_t)(lane & 7);
      if (epi == 2) {                                       // ReLU mask tile arrives by TMA into the same slab
        if (lane == 0) {
          mbar_expect_tx(aux_bar + 8 * q, NBOX * 4096);
          for (int j = 0; j < NBOX; ++j)
            tma_load_2d(slab + j * 4096, &tmX, n0 + j * COLS_PER_BOX, m0 + q * 32, aux_bar + 8 * q);
        }
        mbar_wait(aux_bar + 8 * q, 0);
      }
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
        const int col0 = n0 + c * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (add_bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (col0 + j < N) v[j] += __ldg(bias + col0 + j);
        }
        if (epi == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        const int box = (c * 32) / COLS_PER_BOX;
        const int chunk0 = ((c * 32) % COLS_PER_BOX) * ESZ / 16;     // first 16 B chunk of this 32-column group
        constexpr int NCH = 32 * ESZ / 16;                             // 4 (bf16) or 8 (fp32) chunks
        constexpr int EPC = 16 / ESZ;                                  // elements per chunk
#pragma unroll
        for (int t = 0; t < NCH; ++t) {
          const uint32_t addr = my_row + box * 4096 + (((uint32_t)(chunk0 + t) ^ sw) << 4);
          if (epi == 2) {
            uint4 a;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(addr));
            if constexpr (ESZ == 2) {
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (!(__low2float(h[e]) > 0.f)) v[t * 8 + 2 * e] = 0.f;
                if (!(__high2float(h[e]) > 0.f)) v[t * 8 + 2 * e + 1] = 0.f;
              }
            } else {
              const float* fa = reinterpret_cast<const float*>(&a);
#pragma unroll
              for (int e = 0; e < 4; ++e) if (!(fa[e] > 0.f)) v[t * 4 + e] = 0.f;
            }
          }
          uint4 o;
          if constexpr (ESZ == 2) {
            __nv_bfloat162 h;
            h = __floats2bfloat162_rn(v[t * 8 + 0], v[t * 8 + 1]); o.x = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2bfloat162_rn(v[t * 8 + 2], v[t * 8 + 3]); o.y = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2bfloat162_rn(v[t * 8 + 4], v[t * 8 + 5]); o.z = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2bfloat162_rn(v[t * 8 + 6], v[t * 8 + 7]); o.w = *reinterpret_cast<uint32_t*>(&h);
          } else {
            o.x = __float_as_uint(v[t * EPC + 0]); o.y = __float_as_uint(v[t * EPC + 1]);
            o.z = __float_as_uint(v[t * EPC + 2]); o.w = __float_as_uint(v[t * EPC + 3]);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0 && m0 + q * 32 < M) {
        for (int j = 0; j < NBOX; ++j) {
          if (n0 + j * COLS_PER_BOX >= N) break;
          if (store_mode == 2) tma_reduce_add_2d(&tmC, slab + j * 4096, n0 + j * COLS_PER_BOX, m0 + q * 32);
          else tma_store_2d(&tmC, slab + j * 4096, n0 + j * COLS_PER_BOX, m0 + q * 32);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    } else
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
      const int col0 = n0 + c * 32;
      if (row >= M || col0 >= N) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (add_bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (col0 + j < N) v[j] += __ldg(bias + col0 + j);
      }
      TO* crow = C + (int64_t)row * ldc + col0;
      const bool full = vec_ok && (col0 + 32 <= N);
      if (epi == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      } else if (epi == 2) {
        const TO* arow = aux + (int64_t)row * ldaux + col0;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float a[8];
            OutVec<TO>::load8(arow + j, a);
#pragma unroll
            for (int t = 0; t < 8; ++t) v[j + t] = a[t] > 0.f ? v[j + t] : 0.f;
          }
        } else {
          for (int j = 0; j < 32; ++j) if (col0 + j < N) v[j] = to_f32(arow[j]) > 0.f ? v[j] : 0.f;
        }
      }
      if (accumulate == 2) {
        if constexpr (sizeof(TO) == 4) {
          for (int j = 0; j < 32; ++j) if (col0 + j < N) atomicAdd(reinterpret_cast<float*>(crow) + j, v[j]);
        }
      } else if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (accumulate == 1) {
            float o[8];
            OutVec<TO>::load8(crow + j, o);
#pragma unroll
            for (int t = 0; t < 8; ++t) v[j + t] += o[t];
          }
          OutVec<TO>::store8(crow + j, v + j);
        }
      } else {
        for (int j = 0; j < 32; ++j) {
          if (col0 + j < N) {
            float o = v[j];
            if (accumulate == 1) o += to_f32(crow[j]);
            crow[j] = from_f32<TO>(o);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D row-major tensor [rows][cols] (ld elements) with a 128-byte-wide box of box_rows rows, 128B swizzle
int make_tmap(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows,
              int dtype = ICAP_BF16) {
  EncodeTiledFn fn = get_encode_fn();
  ICAP_ARG(fn != nullptr, "cuTensorMapEncodeTiled entry point not found (driver too old?)");
  const int esz = dtype == ICAP_BF16 ? 2 : 4;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(tm, dtype == ICAP_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ICAP_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): ptr=%p rows=%lld cols=%lld ld=%lld", (int)r, ptr,
           (long long)rows, (long long)cols, (long long)ld);
  return 0;
}

template <bool AK, bool BKM, typename TO>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tx, int M, int N,
           int K, void* C, int64_t ldc, const float* bias, int epi, const void* aux, int64_t ldaux, int accumulate,
           int kb_per_split, int splits, int vec_ok, int store_mode, cudaStream_t st) {
  static bool attr_done = false;
  auto kern = gemm_tc_kernel<AK, BKM, TO>;
  if (!attr_done) {
    ICAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_done = true;
  }
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), (unsigned)splits);
  kern<<<grid, NTHREADS, SMEM_BYTES, st>>>(ta, tb, tc, tx, M, N, K, (TO*)C, ldc, bias, epi, (const TO*)aux, ldaux,
                                           accumulate, kb_per_split, vec_ok, store_mode);
  ICAP_LAUNCH_CHECK("icap_gemm(bf16 tcgen05)");
  return 0;
}

}  // namespace

int icap_gemm_bf16_launch(int a_kmajor, int b_kmajor, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda,
                          const void* B, int64_t ldb, void* C, int64_t ldc, int c_dtype, const float* bias, int epi,
                          const void* aux, int64_t ldaux, int accumulate, int split_k, cudaStream_t st) {
  ICAP_ARG((uintptr_t)A % 16 == 0 && (uintptr_t)B % 16 == 0, "icap_gemm(bf16): A/B must be 16-byte aligned");
  ICAP_ARG(lda % 8 == 0 && ldb % 8 == 0, "icap_gemm(bf16): lda/ldb must be multiples of 8 (TMA 16-byte strides)");
  ICAP_ARG(!(a_kmajor == 0 && b_kmajor == 1), "icap_gemm(bf16): (A MN-major, B K-major) is not instantiated");
  CUtensorMap ta, tb;
  int rc;
  if (a_kmajor) rc = make_tmap(&ta, A, M, K, lda, BM); else rc = make_tmap(&ta, A, K, M, lda, 64);
  if (rc) return rc;
  if (b_kmajor) rc = make_tmap(&tb, B, N, K, ldb, BN); else rc = make_tmap(&tb, B, K, N, ldb, 64);
  if (rc) return rc;
  const int nkb = (int)ceil_div64(K, BK);
  if (split_k < 1) split_k = 1;
  if (split_k > nkb) split_k = nkb;
  int kb_per = (nkb + split_k - 1) / split_k;
  split_k = (nkb + kb_per - 1) / kb_per;
  if (split_k > 1) {
    ICAP_ARG(accumulate != 0 && epi == 0 && c_dtype == ICAP_F32,
             "icap_gemm(bf16): split_k>1 needs fp32 C, accumulate!=0 and no activation epilogue");
    accumulate = 2;
  }
  ICAP_ARG(accumulate != 2 || c_dtype == ICAP_F32, "icap_gemm(bf16): atomic accumulate needs fp32 C");
  const int esz = c_dtype == ICAP_F32 ? 4 : 2;
  int vec_ok = ((uintptr_t)C % 16 == 0) && ((ldc * esz) % 16 == 0);
  if (epi == 2) vec_ok = vec_ok && ((uintptr_t)aux % 16 == 0) && ((ldaux * esz) % 16 == 0);
  // staged TMA epilogue whenever C (and aux) satisfy the TMA alignment rules; accumulation = TMA reduce-add
  int store_mode = vec_ok ? (accumulate ? 2 : 1) : 0;
  if (getenv("ICAP_GEMM_DIRECT_EPILOGUE")) store_mode = 0;
  CUtensorMap tc = ta, tx = ta;
  if (store_mode) {
    if ((rc = make_tmap(&tc, C, M, N, ldc, 32, c_dtype))) return rc;
    if (epi == 2 && (rc = make_tmap(&tx, aux, M, N, ldaux, 32, c_dtype))) return rc;
  }
#define GO(AK, BKM)                                                                                                 \
  (c_dtype == ICAP_F32                                                                                              \
       ? launch<AK, BKM, float>(ta, tb, tc, tx, (int)M, (int)N, (int)K, C, ldc, bias, epi, aux, ldaux, accumulate,  \
                                kb_per, split_k, vec_ok, store_mode, st)                                            \
       : launch<AK, BKM, bf16>(ta, tb, tc, tx, (int)M, (int)N, (int)K, C, ldc, bias, epi, aux, ldaux, accumulate,   \
                               kb_per, split_k, vec_ok, store_mode, st))
  if (a_kmajor && b_kmajor) return GO(true, true);
  if (a_kmajor && !b_kmajor) return GO(true, false);
  return GO(false, false);
#undef GO
}
