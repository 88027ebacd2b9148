// Fused  y = (LayerNorm(A . W^T + bias + res) * gamma + beta) * rowscale   for SMALL row counts (decode steps).
//
// In the KV-cached decode loop every output projection (modules.py:86-90: joint_linear -> dropout(eval: identity) ->
// LayerNorm(out + residual)) is a [rows, 512] x [512, 512] GEMM followed by a LayerNorm over the same rows, with
// rows = batch * beam (2560 for 512 images x 5 beams).  As two launches (tcgen05 GEMM + add_ln) they cost ~8.5 + 5.6 us,
// almost all of it fixed per-launch latency.  Here one CTA owns 32 complete output rows (all N columns), so the
// LayerNorm statistics are CTA-local and the projection never goes to memory:
//   8 warps, warp w computes columns [w*N/8, (w+1)*N/8) of the 32 rows with mma.sync.m16n8k16 (bf16 -> fp32);
//   A (32 x 64) and W (N x 64) k-chunks are staged with cp.async into XOR-swizzled shared memory, double buffered;
//   epilogue: + bias + residual, two-pass mean / variance through an 8 x 32 shared-memory exchange, affine, row scale.
// rows / 32 CTAs (80 for rows = 2560): each streams W once from L2 (N*K*2 B), so this only pays for small row counts
// and K <= 1024; larger problems go through icap_gemm + icap_add_ln_fwd.
#include <stdlib.h>
#include "icap_common.cuh"

namespace {

constexpr int BMR = 32;           // rows per CTA (two m16 tiles)
constexpr int KC = 64;            // k-chunk (one 128-byte swizzle row of bf16)
constexpr int ROWB = KC * 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
  return base + row * ROWB + (((uint32_t)chunk ^ (uint32_t)(row & 7)) << 4);
}
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// N = 8 warps * NPW columns; NPW = 64 (N = 512) or 32 (N = 256) or 128 (N = 1024)
template <int NPW>
__global__ void __launch_bounds__(256, 1)
linear_res_ln_kernel(int M, int K, const bf16* __restrict__ A, int64_t lda, const bf16* __restrict__ W, int64_t ldw,
                     const float* __restrict__ bias, const bf16* __restrict__ res, int64_t ldr,
                     const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ rowscale,
                     bf16* __restrict__ y, int64_t ldy, float eps) {
  pdl_prologue();
  constexpr int N = 8 * NPW;
  constexpr int NT = NPW / 8;               // n8 tiles per warp
  constexpr int STAGE = (BMR + N) * ROWB;   // bytes per k-chunk stage: A tile then W tile
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ float red[8][BMR];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BMR;
  const int nkc = K / KC;

  auto load_chunk = [&](int kc, int stage) {
    const uint32_t sA = sbase + stage * STAGE, sW = sA + BMR * ROWB;
    const int c = threadIdx.x & 7;
    {                                              // A: 32 rows x 8 chunks = 256 copies, one per thread
      const int r = threadIdx.x >> 3;
      const uint32_t dst = tile_addr(sA, r, c);
      if (m0 + r < M)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(A + (int64_t)(m0 + r) * lda + kc * KC + c * 8) : "memory");
      else
        asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0u) : "memory");
    }
    const bf16* wp = W + (int64_t)(threadIdx.x >> 3) * ldw + kc * KC + c * 8;
    const int64_t wstep = 32 * ldw;
#pragma unroll
    for (int r = threadIdx.x >> 3; r < N; r += 32, wp += wstep)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tile_addr(sW, r, c)), "l"(wp) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[2][NT][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f; }

  load_chunk(0, 0);
  for (int kc = 0; kc < nkc; ++kc) {
    const int st = kc & 1;
    if (kc + 1 < nkc) {
      load_chunk(kc + 1, st ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint32_t sA = sbase + st * STAGE, sW = sA + BMR * ROWB;
#pragma unroll
    for (int kk = 0; kk < KC / 16; ++kk) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) ldsm4(tile_addr(sA, mt * 16 + (lane & 15), 2 * kk + (lane >> 4)), a[mt]);
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t b[4];
        ldsm4(tile_addr(sW, warp * NPW + np * 16 + (lane & 7) + ((lane >> 4) << 3), 2 * kk + ((lane >> 3) & 1)), b);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma16816(acc[mt][2 * np], a[mt], b[0], b[1]);
          mma16816(acc[mt][2 * np + 1], a[mt], b[2], b[3]);
        }
      }
    }
    __syncthreads();          // everyone is done with stage st before chunk kc + 2 overwrites it
  }

  // ---- epilogue: x = acc + bias + res ; LayerNorm over the N columns of each row (two-pass statistics)
  // thread: rows rl = mt*16 + (lane>>2) + 8*h (h = 0, 1), columns warp*NPW + nt*8 + 2*(lane&3) + {0, 1}
  float psum[4] = {0.f, 0.f, 0.f, 0.f};      // [mt*2 + h]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int col = warp * NPW + nt * 8 + 2 * (lane & 3);
      const float b0 = bias ? __ldg(bias + col) : 0.f, b1 = bias ? __ldg(bias + col + 1) : 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = m0 + mt * 16 + (lane >> 2) + 8 * h;
        float r0 = 0.f, r1 = 0.f;
        if (res && row < M) {
          const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(res + (int64_t)row * ldr + col);
          r0 = __low2float(t); r1 = __high2float(t);
        }
        acc[mt][nt][2 * h] += b0 + r0;
        acc[mt][nt][2 * h + 1] += b1 + r1;
        psum[mt * 2 + h] += acc[mt][nt][2 * h] + acc[mt][nt][2 * h + 1];
      }
    }
  float mean[4], rstd[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    psum[i] += __shfl_xor_sync(0xffffffffu, psum[i], 1);
    psum[i] += __shfl_xor_sync(0xffffffffu, psum[i], 2);
  }
  if ((lane & 3) == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) red[warp][(i >> 1) * 16 + (lane >> 2) + 8 * (i & 1)] = psum[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rl = (i >> 1) * 16 + (lane >> 2) + 8 * (i & 1);
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][rl];
    mean[i] = t / (float)N;
  }
  __syncthreads();
  float psq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float c0 = acc[mt][nt][2 * h] - mean[mt * 2 + h], c1 = acc[mt][nt][2 * h + 1] - mean[mt * 2 + h];
        psq[mt * 2 + h] += c0 * c0 + c1 * c1;
      }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    psq[i] += __shfl_xor_sync(0xffffffffu, psq[i], 1);
    psq[i] += __shfl_xor_sync(0xffffffffu, psq[i], 2);
  }
  if ((lane & 3) == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) red[warp][(i >> 1) * 16 + (lane >> 2) + 8 * (i & 1)] = psq[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rl = (i >> 1) * 16 + (lane >> 2) + 8 * (i & 1);
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][rl];
    rstd[i] = rsqrtf(t / (float)N + eps);
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int col = warp * NPW + nt * 8 + 2 * (lane & 3);
      const float g0 = __ldg(gamma + col), g1 = __ldg(gamma + col + 1), e0 = __ldg(beta + col), e1 = __ldg(beta + col + 1);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = m0 + mt * 16 + (lane >> 2) + 8 * h;
        if (row < M) {
          const float rs = rowscale ? rowscale[row] : 1.f;
          const float o0 = ((acc[mt][nt][2 * h] - mean[mt * 2 + h]) * rstd[mt * 2 + h] * g0 + e0) * rs;
          const float o1 = ((acc[mt][nt][2 * h + 1] - mean[mt * 2 + h]) * rstd[mt * 2 + h] * g1 + e1) * rs;
          *reinterpret_cast<__nv_bfloat162*>(y + (int64_t)row * ldy + col) = __floats2bfloat162_rn(o0, o1);
        }
      }
    }
}

template <int NPW>
int launch_lrl(int64_t M, int64_t K, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const void* res,
               int64_t ldr, const float* gamma, const float* beta, const float* rowscale, void* y, int64_t ldy, float eps,
               cudaStream_t st) {
  constexpr int N = 8 * NPW;
  const size_t smem = 2 * (size_t)(BMR + N) * ROWB;
  static bool attr_done = false;
  if (!attr_done) {
    ICAP_CUDA(cudaFuncSetAttribute(linear_res_ln_kernel<NPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  ICAP_CUDA(icap_launch(linear_res_ln_kernel<NPW>, dim3((unsigned)ceil_div64(M, BMR)), dim3(256), smem, st, (int)M, (int)K,
                        (const bf16*)A, lda, (const bf16*)W, ldw, bias, (const bf16*)res, ldr, gamma, beta, rowscale, (bf16*)y,
                        ldy, eps));
  return 0;
}

}  // namespace

// y[M,N] = (LayerNorm(A[M,K] . W[N,K]^T + bias + res[M,N]) * gamma + beta) * rowscale[row]   (bf16 in/out, fp32 math)
// Fused output projection + residual + LayerNorm of a decode step (modules.py:86-90 in eval mode).  N in {256, 512};
// K a multiple of 64.  Returns -2 (and sets the error text) for shapes it does not cover: callers fall back to
// icap_gemm + icap_add_ln_fwd.
extern "C" int icap_linear_res_ln(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* W, int64_t ldw,
                                  const float* bias, const void* res, int64_t ldr, const float* gamma, const float* beta,
                                  const float* rowscale, void* y, int64_t ldy, float eps, void* stream) {
  ICAP_ARG(M > 0 && A && W && gamma && beta && y, "icap_linear_res_ln: null/empty argument");
  const bool ok = (N == 256 || N == 512) && K % KC == 0 && K > 0 && lda % 8 == 0 && ldw % 8 == 0 && ldy % 2 == 0 &&
                  (res == nullptr || ldr % 2 == 0) && (uintptr_t)A % 16 == 0 && (uintptr_t)W % 16 == 0 &&
                  (uintptr_t)y % 4 == 0 && (uintptr_t)res % 4 == 0;
  if (!ok) {
    icap_set_error("icap_linear_res_ln: unsupported shape/alignment (N=%lld K=%lld)", (long long)N, (long long)K);
    return -2;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 512) return launch_lrl<64>(M, K, A, lda, W, ldw, bias, res, ldr, gamma, beta, rowscale, y, ldy, eps, st);
  return launch_lrl<32>(M, K, A, lda, W, ldw, bias, res, ldr, gamma, beta, rowscale, y, ldy, eps, st);
}
