// Fused (dropout) + residual add + LayerNorm (+ non-pad row mask), forward and backward.
//
//   s   = dropout_p(a[row]) + res[row % res_rows]          (s optionally written back over a)
//   y   = (LN(s) * gamma + beta) * rowscale[row]
//
// This is the post-LN tail of every sub-layer of the reference (modules.py:86-90,117-120), the
// embedding norms (model.py:307-309,433-436: res = positional table, res_rows = T) and the per-block
// `output *= non_pad_mask` (modules.py:154-155,203-204) folded in as rowscale.  HBM-bound: one warp
// per row, 16-byte vector accesses, warp-shuffle statistics, no shared memory.
#include "icap_common.cuh"

namespace {

constexpr int MAX_IT = 8;   // d <= MAX_IT * 32 * 4 = 1024 ; kernels are specialised on NIT = ceil(d / 128)

template <int NIT, typename TA, typename TR, typename TY>
__global__ void __launch_bounds__(256)
add_ln_fwd_kernel(int M, int d, TA* __restrict__ a, const TR* __restrict__ res, int res_rows,
                  const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ rowscale,
                  TY* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int write_sum,
                  float p_drop, uint32_t thresh, uint64_t seed, const int* __restrict__ seed_dev, float eps) {
  if (seed_dev) seed += (uint64_t)(*seed_dev) * 0x9E3779B97F4A7C15ull;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int nvec = d >> 2;
  float v[NIT][4];
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  TA* arow = a + (int64_t)row * d;
  const TR* rrow = res ? res + (int64_t)(row % res_rows) * d : nullptr;
  float sum = 0.f;
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int vi = it * 32 + lane;
    if (vi < nvec) {
      load4(arow + vi * 4, v[it]);
      if (p_drop > 0.f) {
        uint32_t keep = dropout_keep4(seed, (uint64_t)row * nvec + vi, thresh);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[it][j] = (keep >> j) & 1 ? v[it][j] * keep_scale : 0.f;
      }
      if (rrow) {
        float r[4];
        load4(rrow + vi * 4, r);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[it][j] += r[j];
      }
      if (write_sum) store4(arow + vi * 4, v[it]);
#pragma unroll
      for (int j = 0; j < 4; ++j) sum += v[it][j];
    }
  }
  const float mean = warp_sum(sum) / (float)d;
  float sq = 0.f;
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int vi = it * 32 + lane;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { float c = v[it][j] - mean; sq += c * c; }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)d + eps);
  const float rs = rowscale ? rowscale[row] : 1.f;
  TY* yrow = y + (int64_t)row * d;
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int vi = it * 32 + lane;
    if (vi < nvec) {
      float g[4], b[4], o[4];
      load4(gamma + vi * 4, g);
      load4(beta + vi * 4, b);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = ((v[it][j] - mean) * rstd * g[j] + b[j]) * rs;
      store4(yrow + vi * 4, o);
    }
  }
  if (lane == 0 && mean_out) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

// Backward.  dy = dy1 (+ dy2);  g = dy * rowscale
//   dgamma += sum_rows g * xhat ; dbeta += sum_rows g
//   ds = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat))      (-> residual branch)
//   da = dropout mask(ds) / (1-p)                                          (-> GEMM branch; == ds if p = 0)
//   dbias2 += sum_rows da   (optional: bias of the GEMM that produced a)
// Persistent over rows: each warp keeps per-lane column partials in registers, one smem + atomic
// reduction per block at the end.
template <int NIT, typename T>
__global__ void __launch_bounds__(256)
add_ln_bwd_kernel(int M, int d, const T* __restrict__ dy1, const T* __restrict__ dy2, const T* __restrict__ s,
                  const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                  const float* __restrict__ gamma, const float* __restrict__ rowscale, T* __restrict__ ds,
                  T* __restrict__ da, float* __restrict__ dgamma, float* __restrict__ dbeta,
                  float* __restrict__ dbias2, float p_drop, uint32_t thresh, uint64_t seed,
                  const int* __restrict__ seed_dev) {
  if (seed_dev) seed += (uint64_t)(*seed_dev) * 0x9E3779B97F4A7C15ull;
  extern __shared__ float red[];   // [3][d]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int nvec = d >> 2;
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float pg[NIT][4], pb[NIT][4], pc[NIT][4];
#pragma unroll
  for (int it = 0; it < NIT; ++it)
#pragma unroll
    for (int j = 0; j < 4; ++j) { pg[it][j] = 0.f; pb[it][j] = 0.f; pc[it][j] = 0.f; }
  float gam[NIT][4];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int vi = it * 32 + lane;
    if (vi < nvec) load4(gamma + vi * 4, gam[it]);
  }

  for (int row = blockIdx.x * wpb + wib; row < M; row += gridDim.x * wpb) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    const float rs = rowscale ? rowscale[row] : 1.f;
    float xh[NIT][4], gg[NIT][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int vi = it * 32 + lane;
      if (vi < nvec) {
        float g[4], x[4];
        load4(dy1 + (int64_t)row * d + vi * 4, g);
        if (dy2) {
          float g2[4];
          load4(dy2 + (int64_t)row * d + vi * 4, g2);
#pragma unroll
          for (int j = 0; j < 4; ++j) g[j] += g2[j];
        }
        load4(s + (int64_t)row * d + vi * 4, x);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float gr = g[j] * rs;
          const float xhat = (x[j] - mean) * rstd;
          pg[it][j] += gr * xhat;
          pb[it][j] += gr;
          const float gx = gr * gam[it][j];
          xh[it][j] = xhat;
          gg[it][j] = gx;
          s1 += gx;
          s2 += gx * xhat;
        }
      }
    }
    s1 = warp_sum(s1) / (float)d;
    s2 = warp_sum(s2) / (float)d;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int vi = it * 32 + lane;
      if (vi < nvec) {
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = rstd * (gg[it][j] - s1 - xh[it][j] * s2);
        if (ds) store4(ds + (int64_t)row * d + vi * 4, o);
        if (da) {
          if (p_drop > 0.f) {
            uint32_t keep = dropout_keep4(seed, (uint64_t)row * nvec + vi, thresh);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = (keep >> j) & 1 ? o[j] * keep_scale : 0.f;
          }
          store4(da + (int64_t)row * d + vi * 4, o);
        }
        if (dbias2) {
          if (!da && p_drop > 0.f) {
            uint32_t keep = dropout_keep4(seed, (uint64_t)row * nvec + vi, thresh);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = (keep >> j) & 1 ? o[j] * keep_scale : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) pc[it][j] += o[j];
        }
      }
    }
  }
  // block reduction of the column partials
  for (int i = threadIdx.x; i < 3 * d; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int vi = it * 32 + lane;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&red[vi * 4 + j], pg[it][j]);
        atomicAdd(&red[d + vi * 4 + j], pb[it][j]);
        if (dbias2) atomicAdd(&red[2 * d + vi * 4 + j], pc[it][j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, red[i]);
    if (dbeta) atomicAdd(dbeta + i, red[d + i]);
    if (dbias2) atomicAdd(dbias2 + i, red[2 * d + i]);
  }
}

}  // namespace

extern "C" int icap_add_ln_fwd(int a_dtype, int act_dtype, int64_t M, int64_t d, void* a, const void* res,
                               int64_t res_rows, const float* gamma, const float* beta, const float* rowscale,
                               void* y, float* mean_out, float* rstd_out, int write_sum, float p_drop,
                               uint64_t seed, const int* seed_dev, float eps, void* stream) {
  ICAP_ARG(d % 4 == 0 && d <= MAX_IT * 128, "icap_add_ln_fwd: d=%lld must be a multiple of 4 and <= %d", (long long)d,
           MAX_IT * 128);
  ICAP_ARG(M > 0 && a && y && gamma && beta, "icap_add_ln_fwd: null argument");
  if (res == nullptr) res_rows = 1;
  ICAP_ARG(res_rows > 0, "icap_add_ln_fwd: res_rows must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)ceil_div64(M, 8));
  const uint32_t th = dropout_threshold(p_drop);
#define GO1(NIT, TA, TR, TY)                                                                                      \
  add_ln_fwd_kernel<NIT, TA, TR, TY><<<grid, 256, 0, st>>>((int)M, (int)d, (TA*)a, (const TR*)res,               \
                                                            (int)res_rows, gamma, beta, rowscale, (TY*)y,        \
                                                            mean_out, rstd_out, write_sum, p_drop, th, seed, seed_dev, eps)
#define GO(TA, TR, TY)                                                                                            \
  do {                                                                                                            \
    if (d <= 128) GO1(1, TA, TR, TY);                                                                             \
    else if (d <= 256) GO1(2, TA, TR, TY);                                                                        \
    else if (d <= 512) GO1(4, TA, TR, TY);                                                                        \
    else GO1(8, TA, TR, TY);                                                                                      \
  } while (0)
  if (a_dtype == ICAP_F32 && act_dtype == ICAP_F32) GO(float, float, float);
  else if (a_dtype == ICAP_BF16 && act_dtype == ICAP_BF16) GO(bf16, bf16, bf16);
  else if (a_dtype == ICAP_F32 && act_dtype == ICAP_BF16) GO(float, bf16, bf16);
  else ICAP_ARG(false, "icap_add_ln_fwd: unsupported dtype combination a=%d act=%d", a_dtype, act_dtype);
#undef GO
#undef GO1
  ICAP_LAUNCH_CHECK("icap_add_ln_fwd");
  return 0;
}

extern "C" int icap_add_ln_bwd(int act_dtype, int64_t M, int64_t d, const void* dy1, const void* dy2, const void* s,
                               const float* mean, const float* rstd, const float* gamma, const float* rowscale,
                               void* ds, void* da, float* dgamma, float* dbeta, float* dbias2, float p_drop,
                               uint64_t seed, const int* seed_dev, void* stream) {
  ICAP_ARG(d % 4 == 0 && d <= MAX_IT * 128, "icap_add_ln_bwd: d=%lld must be a multiple of 4 and <= %d", (long long)d,
           MAX_IT * 128);
  ICAP_ARG(M > 0 && dy1 && s && mean && rstd && gamma, "icap_add_ln_bwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = ceil_div64(M, 8);
  if (blocks > 148 * 2) blocks = 148 * 2;   // few blocks: the column reductions end in one global atomic per block
  const uint32_t th = dropout_threshold(p_drop);
  size_t smem = 3 * d * sizeof(float);
#define GOB(NIT, T)                                                                                               \
  add_ln_bwd_kernel<NIT, T><<<(unsigned)blocks, 256, smem, st>>>(                                                 \
      (int)M, (int)d, (const T*)dy1, (const T*)dy2, (const T*)s, mean, rstd, gamma, rowscale, (T*)ds, (T*)da,     \
      dgamma, dbeta, dbias2, p_drop, th, seed, seed_dev)
#define GOBT(T)                                                                                                   \
  do {                                                                                                            \
    if (d <= 128) GOB(1, T);                                                                                      \
    else if (d <= 256) GOB(2, T);                                                                                 \
    else if (d <= 512) GOB(4, T);                                                                                 \
    else GOB(8, T);                                                                                               \
  } while (0)
  if (act_dtype == ICAP_F32) GOBT(float);
  else GOBT(bf16);
#undef GOBT
#undef GOB
  ICAP_LAUNCH_CHECK("icap_add_ln_bwd");
  return 0;
}
