// Small-dimension (D <= 32) variants of the three sample x component kernels of the SAMTRON iteration.
//
// The reference's own examples run at D = 20 (examples/5_samtron_20D_student-T.py) and D = 10 (examples/6_samtron_planar4.py):
// BASELINE configurations C1 / C2.  The tile engines built for D = 100 ... 256 (gemm_kernels.cu: 128 x 64 x 16 tiles; the
// tcgen05 kernels: 64-column operand blocks) pad such a problem by 4 - 40x -- the C2 iteration spent 1.4 ms in four
// log-density launches whose useful work is 90 MFMA.  Here the whole per-pair computation lives in one thread's
// registers:
//   * log-density / mixture gradient: one thread = one sample, x in registers, the component's factor (padded to DP x DP,
//     DP in {8, 16, 24, 32}) broadcast from shared memory as 128-bit loads (4 FMAs per load), components staged in chunks;
//   * Stein statistics: one CTA = (component, range of 128-sample blocks), one thread = up to 5 entries of the extended
//     matrix [x - mu | -1]^T [g] (the gradient sum -sum w g rides along as an extra row), weightless blocks skipped,
//     partial sums of the ranges added in a fixed order by a second tiny kernel (deterministic).
// Plain fp32 arithmetic in the reference's order of operations (x - mu first, then the triangular product).
#include "common.cuh"
#include <stdlib.h>

namespace gvi {
namespace sd {

constexpr int TPB = 128;          // samples per CTA (log-density, gradient)
constexpr int KC = 8;             // components staged per shared-memory chunk

__host__ __device__ constexpr int padded(int D) { return D <= 8 ? 8 : (D <= 16 ? 16 : (D <= 24 ? 24 : 32)); }

// stage `nk` matrices [D x D] (row-major, global) zero-padded to [DP x DP] into shared memory
template <int DP, bool LOWER>
__device__ __forceinline__ void stage_matrices(float* dst, const float* __restrict__ src, int nk, int D) {
  for (int e = threadIdx.x; e < nk * DP * DP; e += blockDim.x) {
    const int kk = e / (DP * DP), r = (e / DP) % DP, c = e % DP;
    dst[e] = (r < D && c < (LOWER ? r + 1 : D)) ? __ldg(src + ((long long)kk * D + r) * D + c) : 0.f;
  }
}
template <int DP>
__device__ __forceinline__ void stage_vectors(float* dst, const float* __restrict__ src, int nk, int D) {
  for (int e = threadIdx.x; e < nk * DP; e += blockDim.x) {
    const int kk = e / DP, c = e % DP;
    dst[e] = c < D ? __ldg(src + (long long)kk * D + c) : 0.f;
  }
}

// lq[k, n] = cst[k] - 1/2 | Linv_k (x_n - mu_k) |^2      (models/full_cov_gmm.py:56-62)
template <int DP>
__global__ void __launch_bounds__(TPB)
logdens_small_kernel(const float* __restrict__ X, int N, int D, const float* __restrict__ means,
                     const float* __restrict__ linv, const float* __restrict__ cst, int K, int kchunks_per_cta,
                     float* __restrict__ lq) {
  __shared__ __align__(16) float Ls[KC * DP * DP];
  __shared__ __align__(16) float Ms[KC * DP];
  __shared__ float Cs[KC];
  const int n = blockIdx.x * TPB + threadIdx.x;
  const bool live = n < N;
  float x[DP];
#pragma unroll
  for (int j = 0; j < DP; ++j) x[j] = (live && j < D) ? __ldg(X + (long long)n * D + j) : 0.f;
  const int k_begin = blockIdx.y * kchunks_per_cta * KC;
  const int k_end = min(K, k_begin + kchunks_per_cta * KC);
  for (int k0 = k_begin; k0 < k_end; k0 += KC) {
    const int nk = min(KC, k_end - k0);
    __syncthreads();
    stage_matrices<DP, true>(Ls, linv + (long long)k0 * D * D, nk, D);
    stage_vectors<DP>(Ms, means + (long long)k0 * D, nk, D);
    if (threadIdx.x < nk) Cs[threadIdx.x] = __ldg(cst + k0 + threadIdx.x);
    __syncthreads();
    if (!live) continue;
    for (int kk = 0; kk < nk; ++kk) {
      float d[DP];
#pragma unroll
      for (int j = 0; j < DP; ++j) d[j] = x[j] - Ms[kk * DP + j];
      const float* L = Ls + kk * DP * DP;
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < DP; ++i) {
        float z = 0.f;
#pragma unroll
        for (int q = 0; q <= i / 4; ++q) {
          const float4 l = *reinterpret_cast<const float4*>(L + i * DP + 4 * q);
          z = fmaf(l.x, d[4 * q], z);
          z = fmaf(l.y, d[4 * q + 1], z);
          z = fmaf(l.z, d[4 * q + 2], z);
          z = fmaf(l.w, d[4 * q + 3], z);
        }
        ss = fmaf(z, z, ss);
      }
      lq[(long long)(k0 + kk) * N + n] = Cs[kk] - 0.5f * ss;
    }
  }
}

// grad[n, :] = - sum_k r_kn P_k (x_n - mu_k),  r_kn = exp(lq[k, n] + logw[k] - logq[n])      (models/gmm.py:274-300)
// Components whose responsibility is below e^-60 for every sample of the warp are skipped (exact in fp32).
template <int DP>
__global__ void __launch_bounds__(TPB)
mixgrad_small_kernel(const float* __restrict__ X, int N, int D, const float* __restrict__ means,
                     const float* __restrict__ prec, const float* __restrict__ lq, const float* __restrict__ logw,
                     const float* __restrict__ logq, int K, float* __restrict__ grad) {
  __shared__ __align__(16) float Ps[KC * DP * DP];
  __shared__ __align__(16) float Ms[KC * DP];
  __shared__ float Ws[KC];
  const int n = blockIdx.x * TPB + threadIdx.x;
  const bool live = n < N;
  float x[DP], g[DP];
#pragma unroll
  for (int j = 0; j < DP; ++j) {
    x[j] = (live && j < D) ? __ldg(X + (long long)n * D + j) : 0.f;
    g[j] = 0.f;
  }
  const float lqn = live ? __ldg(logq + n) : 0.f;
  for (int k0 = 0; k0 < K; k0 += KC) {
    const int nk = min(KC, K - k0);
    __syncthreads();
    stage_matrices<DP, false>(Ps, prec + (long long)k0 * D * D, nk, D);
    stage_vectors<DP>(Ms, means + (long long)k0 * D, nk, D);
    if (threadIdx.x < nk) Ws[threadIdx.x] = __ldg(logw + k0 + threadIdx.x);
    __syncthreads();
    for (int kk = 0; kk < nk; ++kk) {
      const float lr = live ? __ldg(lq + (long long)(k0 + kk) * N + n) + Ws[kk] - lqn : -INFINITY;
      if (!__any_sync(0xffffffffu, lr > -60.f)) continue;
      const float r = expf(lr);
      float d[DP];
#pragma unroll
      for (int j = 0; j < DP; ++j) d[j] = x[j] - Ms[kk * DP + j];
      const float* P = Ps + kk * DP * DP;
#pragma unroll
      for (int i = 0; i < DP; ++i) {
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < DP / 4; ++q) {
          const float4 p = *reinterpret_cast<const float4*>(P + i * DP + 4 * q);
          v = fmaf(p.x, d[4 * q], v);
          v = fmaf(p.y, d[4 * q + 1], v);
          v = fmaf(p.z, d[4 * q + 2], v);
          v = fmaf(p.w, d[4 * q + 3], v);
        }
        g[i] = fmaf(-r, v, g[i]);
      }
    }
  }
  if (live) {
#pragma unroll
    for (int j = 0; j < DP; ++j)
      if (j < D) grad[(long long)n * D + j] = g[j];
  }
}

// ---- Stein statistics ---------------------------------------------------------------------------------------
// part[s][k][i][j] (i in [0, D], j in [0, D)) = sum over the 128-sample blocks of split s of
//     w_kn (x_ni - mu_ki) g_nj     for i < D,          - w_kn g_nj     for i = D   (the gradient sum).
constexpr int ST_THREADS = 256;
constexpr int ST_MAX_OUT = 5;       // (32 + 1) * 32 = 1056 <= 5 * 256

__global__ void __launch_bounds__(ST_THREADS)
stein_small_kernel(const float* __restrict__ X, int N, int D, const float* __restrict__ means,
                   const float* __restrict__ W, const uint8_t* __restrict__ active, const float* __restrict__ G, int S,
                   float* __restrict__ part) {
  __shared__ float xs[128 * 33];      // [n][i] pitch D + 1: w (x - mu), and -w in column D
  __shared__ float gs[128 * 32];      // [n][j] pitch D
  __shared__ float mu[32];
  const int k = blockIdx.x, s = blockIdx.y, K = gridDim.x;
  const int nblk = ceil_div(N, 128);
  const int b0 = (int)((long long)nblk * s / S), b1 = (int)((long long)nblk * (s + 1) / S);
  const int D1 = D + 1, nout = D1 * D;
  if (threadIdx.x < D) mu[threadIdx.x] = __ldg(means + (long long)k * D + threadIdx.x);
  float acc[ST_MAX_OUT];
  int oi[ST_MAX_OUT], oj[ST_MAX_OUT];
#pragma unroll
  for (int q = 0; q < ST_MAX_OUT; ++q) {
    acc[q] = 0.f;
    const int e = threadIdx.x + q * ST_THREADS;
    oi[q] = e < nout ? e / D : 0;
    oj[q] = e < nout ? e % D : 0;
  }
  const float* Wk = W + (long long)k * N;
  for (int b = b0; b < b1; ++b) {
    if (active != nullptr && active[(long long)k * nblk + b] == 0) continue;     // uniform per CTA
    const int n0 = b * 128, cnt = min(128, N - n0);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * D; e += ST_THREADS) {
      const int r = e / D, c = e % D;
      const float w = __ldg(Wk + n0 + r);
      xs[r * D1 + c] = w * (__ldg(X + (long long)(n0 + r) * D + c) - mu[c]);
      gs[r * D + c] = __ldg(G + (long long)(n0 + r) * D + c);
      if (c == 0) xs[r * D1 + D] = -w;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < ST_MAX_OUT; ++q) {
      if (threadIdx.x + q * ST_THREADS < nout) {
        float a = acc[q];
        const float* xp = xs + oi[q];
        const float* gp = gs + oj[q];
        for (int r = 0; r < cnt; ++r) a = fmaf(xp[r * D1], gp[r * D], a);
        acc[q] = a;
      }
    }
  }
  float* out = part + ((long long)s * K + k) * nout;
#pragma unroll
  for (int q = 0; q < ST_MAX_OUT; ++q) {
    const int e = threadIdx.x + q * ST_THREADS;
    if (e < nout) out[e] = acc[q];
  }
}

// M[k][i][j] = sum_s part[s][k][i][j] (i < D), gneg[k][j] = sum_s part[s][k][D][j]; fixed order of the splits
__global__ void stein_small_reduce_kernel(const float* __restrict__ part, int K, int D, int S, float* __restrict__ M,
                                          float* __restrict__ gneg) {
  const int nout = (D + 1) * D;
  const long long total = (long long)K * nout;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < S; ++s) a += part[(long long)s * total + e];
    const int k = (int)(e / nout), r = (int)(e % nout);
    if (r < D * D) M[(long long)k * D * D + r] = a;
    else gneg[(long long)k * D + (r - D * D)] = a;
  }
}

}  // namespace sd

// GMMVI_B200_SMALL_DIM=0 keeps the general tile engines (read per call: the tests compare both)
bool small_dim_supported(int D) {
  const char* e = getenv("GMMVI_B200_SMALL_DIM");
  return D >= 1 && D <= 32 && !(e != nullptr && e[0] == '0');
}

int launch_logdens_small(const float* X, int N, int D, const float* means, const float* linv, const float* cst, int K,
                         float* lq, cudaStream_t st) {
  using namespace sd;
  const int nb = ceil_div(N, TPB);
  // enough CTAs to fill the machine: split the components over blockIdx.y when there are few sample blocks
  const int kchunks = ceil_div(K, KC);
  int ysplit = min(kchunks, max(1, ceil_div(148 * 4, nb)));
  const int per = ceil_div(kchunks, ysplit);
  ysplit = ceil_div(kchunks, per);
  dim3 grid(nb, ysplit);
  switch (padded(D)) {
    case 8: logdens_small_kernel<8><<<grid, TPB, 0, st>>>(X, N, D, means, linv, cst, K, per, lq); break;
    case 16: logdens_small_kernel<16><<<grid, TPB, 0, st>>>(X, N, D, means, linv, cst, K, per, lq); break;
    case 24: logdens_small_kernel<24><<<grid, TPB, 0, st>>>(X, N, D, means, linv, cst, K, per, lq); break;
    default: logdens_small_kernel<32><<<grid, TPB, 0, st>>>(X, N, D, means, linv, cst, K, per, lq); break;
  }
  return check_launch("logdens_small_kernel");
}

int launch_mixgrad_small(const float* X, int N, int D, const float* means, const float* prec, const float* lq,
                         const float* logw, const float* logq, int K, float* grad, cudaStream_t st) {
  using namespace sd;
  dim3 grid(ceil_div(N, TPB));
  switch (padded(D)) {
    case 8: mixgrad_small_kernel<8><<<grid, TPB, 0, st>>>(X, N, D, means, prec, lq, logw, logq, K, grad); break;
    case 16: mixgrad_small_kernel<16><<<grid, TPB, 0, st>>>(X, N, D, means, prec, lq, logw, logq, K, grad); break;
    case 24: mixgrad_small_kernel<24><<<grid, TPB, 0, st>>>(X, N, D, means, prec, lq, logw, logq, K, grad); break;
    default: mixgrad_small_kernel<32><<<grid, TPB, 0, st>>>(X, N, D, means, prec, lq, logw, logq, K, grad); break;
  }
  return check_launch("mixgrad_small_kernel");
}

int stein_small_splits(int N, int K) {
  const int nblk = ceil_div(N, 128);
  return max(1, min(nblk, ceil_div(148 * 2, max(K, 1))));
}
size_t stein_small_workspace_floats(int N, int K, int D) {
  return (size_t)stein_small_splits(N, K) * K * (D + 1) * D;
}
// M [K, D, D] and gneg [K, D] (both written); ws: stein_small_workspace_floats(N, K, D) floats
int launch_stein_small(const float* X, int N, int D, const float* means, const float* W, const uint8_t* active,
                       const float* G, int K, float* M, float* gneg, float* ws, cudaStream_t st) {
  using namespace sd;
  const int S = stein_small_splits(N, K);
  dim3 grid(K, S);
  stein_small_kernel<<<grid, ST_THREADS, 0, st>>>(X, N, D, means, W, active, G, S, ws);
  int rc = check_launch("stein_small_kernel");
  if (rc) return rc;
  const long long total = (long long)K * (D + 1) * D;
  stein_small_reduce_kernel<<<(int)min((long long)1024, (total + 255) / 256), 256, 0, st>>>(ws, K, D, S, M, gneg);
  return check_launch("stein_small_reduce_kernel");
}

}  // namespace gvi
