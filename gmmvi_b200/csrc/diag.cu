// Diagonal-covariance family, second generation (BASELINE configuration C4-diagonal: K = 256, D = 200, N = 16384).
//
// The first kernels (elementwise.cu) spent 0.66 ms per log-density pass and 1.24 ms in the Stein statistics at that shape:
// a division per element and two global loads per FMA in the former, one thread per (component, dimension) walking all
// N samples serially in the latter.  Here
//   * log densities: one thread = one sample with 32 coordinates in registers at a time, the parameters of 8 components
//     staged as (mu, 1/sigma) pairs in shared memory and read as 128-bit broadcasts (3.5 instructions per element),
//     log-normalisers reduced by the staging warps;
//   * Stein statistics: -E[H]_kd = -(sum_n w_kn x_nd g_nd - mu_kd sum_n w_kn g_nd) / sigma_kd^2 and -E[g]_kd = -sum_n w_kn
//     g_nd are TWO matrix products of the weight matrix W [K, N] with [X o G | G] [N, 2D] (ng_estimator.py:177-180), run
//     as one split-K batched GEMM on the existing engines + a fixed-order reduction (deterministic).
#include "common.cuh"

namespace gvi {

int launch_gemm_auto(int transA, int transB, int batch, int M, int N, int Kd, float alpha, const float* A, int lda,
                     long long strideA, const float* B, int ldb, long long strideB, float* C, int ldc,
                     long long strideC, float* ws, size_t ws_floats, cudaStream_t st);
size_t tc_gemm_workspace_floats(int batch, int M, int N, int Kd);

namespace dg {

constexpr int TS = 128;      // samples per CTA
constexpr int CPC = 16;      // components per CTA (their accumulators live in registers across the coordinate chunks)
constexpr int DC = 32;       // coordinates held in registers at a time

// lq[k, n] = -D/2 log 2 pi - sum_d log sigma_kd - 1/2 sum_d ((mu_kd - x_nd) / sigma_kd)^2       (diagonal_gmm.py:31-34, 47-53)
// Loop order: coordinate chunk outside, component inside, so that a thread reads its 32 coordinates ONCE per chunk (a row
// per thread is an uncoalesced access: 32 L1 wavefronts per warp-wide load, and with the component chunk outside the kernel
// was bound by exactly those: ncu l1tex 85 %, profiles/r02_ncu_logdens_diag2.txt) and the 32 accumulators stay in registers.
__global__ void __launch_bounds__(TS)
logdens_diag2_kernel(const float* __restrict__ X, int N, int D, int Dp, const float* __restrict__ means,
                     const float* __restrict__ stds, int K, float* __restrict__ lq) {
  extern __shared__ __align__(16) float smem[];
  float2* par = reinterpret_cast<float2*>(smem);            // [CPC][Dp] (mu, 1 / sigma); (0, 0) past D
  float* cst = smem + 2 * CPC * Dp;                         // [CPC]
  const int n = blockIdx.x * TS + threadIdx.x;
  const bool live = n < N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool vec = (D % 4 == 0) && (reinterpret_cast<uintptr_t>(X) % 16 == 0);
  const int k0 = blockIdx.y * CPC;
  const int nk = min(CPC, K - k0);
  for (int kk = warp; kk < nk; kk += TS / 32) {
    const float* mu = means + (long long)(k0 + kk) * D;
    const float* sg = stds + (long long)(k0 + kk) * D;
    float ls = 0.f;
    for (int d = lane; d < Dp; d += 32) {
      float2 p = make_float2(0.f, 0.f);
      if (d < D) {
        const float s = __ldg(sg + d);
        p = make_float2(__ldg(mu + d), 1.f / s);
        ls += logf(s);
      }
      par[kk * Dp + d] = p;
    }
    ls = warp_sum(ls);
    if (lane == 0) cst[kk] = -0.5f * (float)D * kLog2Pi - ls;
  }
  __syncthreads();
  if (!live) return;
  float acc[CPC];
#pragma unroll
  for (int kk = 0; kk < CPC; ++kk) acc[kk] = 0.f;
  for (int d0 = 0; d0 < Dp; d0 += DC) {
    float x[DC];
    const float* xr = X + (long long)n * D + d0;
    if (vec) {
#pragma unroll
      for (int j = 0; j < DC; j += 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d0 + j < D) v = __ldg(reinterpret_cast<const float4*>(xr + j));
        x[j] = v.x; x[j + 1] = v.y; x[j + 2] = v.z; x[j + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < DC; ++j) x[j] = (d0 + j < D) ? __ldg(xr + j) : 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < CPC; ++kk) {
      if (kk < nk) {
        const float4* p4 = reinterpret_cast<const float4*>(par + kk * Dp + d0);
        float a0 = acc[kk], a1 = 0.f;                     // two chains: the FMA latency is exposed otherwise
#pragma unroll
        for (int j = 0; j < DC; j += 2) {
          const float4 p = p4[j >> 1];                    // (mu_j, isg_j, mu_j+1, isg_j+1)
          const float t0 = (p.x - x[j]) * p.y, t1 = (p.z - x[j + 1]) * p.w;
          a0 = fmaf(t0, t0, a0);
          a1 = fmaf(t1, t1, a1);
        }
        acc[kk] = a0 + a1;
      }
    }
  }
#pragma unroll
  for (int kk = 0; kk < CPC; ++kk)
    if (kk < nk) lq[(long long)(k0 + kk) * N + n] = cst[kk] - 0.5f * acc[kk];
}

// B[n] = [x_n o g_n | g_n]  (N x 2D)
__global__ void stein_diag_operand_kernel(const float* __restrict__ X, const float* __restrict__ G, long long N, int D,
                                          float* __restrict__ B) {
  const long long total = N * D;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / D;
    const int c = (int)(e % D);
    const float g = G[e];
    B[r * 2 * D + c] = X[e] * g;
    B[r * 2 * D + D + c] = g;
  }
}

// Hneg[k, d] = -(A1 - mu A2) / sigma^2, gneg[k, d] = -A2 with [A1 | A2] = sum over the split partial products (fixed order)
__global__ void stein_diag_finalize_kernel(const float* __restrict__ part, int S, int K, int D, const float* __restrict__ means,
                                           const float* __restrict__ stds, float* __restrict__ Hneg, float* __restrict__ gneg) {
  const long long total = (long long)K * D;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long k = e / D;
    const int d = (int)(e % D);
    float a1 = 0.f, a2 = 0.f;
    for (int s = 0; s < S; ++s) {
      const float* p = part + ((long long)s * K + k) * 2 * D;
      a1 += p[d];
      a2 += p[D + d];
    }
    const float sg = stds[e];
    Hneg[e] = -(a1 - means[e] * a2) / (sg * sg);
    gneg[e] = -a2;
  }
}

}  // namespace dg

int launch_logdens_diag2(const float* X, int N, int D, const float* means, const float* stds, int K, float* lq,
                         cudaStream_t st) {
  using namespace dg;
  const int Dp = ceil_div(D, DC) * DC;
  const size_t smem = (size_t)(2 * CPC * Dp + CPC) * sizeof(float);
  if (smem > 200 * 1024) return 1;                    // D > ~780: the caller keeps the first-generation kernel
  static unsigned long long attr_set_mask = 0;
  if (smem > 48 * 1024 && first_call_on_device(attr_set_mask))
    cudaFuncSetAttribute(logdens_diag2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  logdens_diag2_kernel<<<dim3(ceil_div(N, TS), ceil_div(K, CPC)), TS, smem, st>>>(X, N, D, Dp, means, stds, K, lq);
  return check_launch("logdens_diag2_kernel");
}

static void stein_diag_split(int N, int K, int D, int& S, int& Ns) {
  // enough (tile, split) CTAs for ~2 waves; split lengths a multiple of 32 samples
  const int tiles = ceil_div(K, 128) * ceil_div(2 * D, 64);
  S = max(1, min(ceil_div(N, 256), ceil_div(148 * 2, tiles)));
  Ns = ceil_div(ceil_div(N, S), 32) * 32;
  S = N / Ns;                                        // full splits; the remainder is one more partial product
}
size_t stein_diag_workspace_floats(int N, int K, int D) {
  if (N <= 0 || K <= 0) return 0;
  int S, Ns;
  stein_diag_split(N, K, D, S, Ns);
  const size_t b = (size_t)N * 2 * D, part = (size_t)(S + 1) * K * 2 * D;
  return (b + part + 63) / 64 * 64 + tc_gemm_workspace_floats(max(S, 1), K, 2 * D, Ns);
}
int launch_stein_diag2(const float* X, int N, int D, const float* means, const float* stds, const float* W, const float* G,
                       int K, float* Hneg, float* gneg, float* ws, cudaStream_t st) {
  using namespace dg;
  int S, Ns;
  stein_diag_split(N, K, D, S, Ns);
  float* B = ws;
  float* part = B + (size_t)N * 2 * D;
  float* tcws = ws + ((size_t)N * 2 * D + (size_t)(S + 1) * K * 2 * D + 63) / 64 * 64;
  const size_t tcws_floats = tc_gemm_workspace_floats(max(S, 1), K, 2 * D, Ns);
  const long long total = (long long)N * D;
  stein_diag_operand_kernel<<<(int)min((long long)148 * 8, (total + 255) / 256), 256, 0, st>>>(X, G, N, D, B);
  int rc = check_launch("stein_diag_operand_kernel");
  if (rc) return rc;
  int parts = 0;
  if (S > 0) {
    rc = launch_gemm_auto(0, 0, S, K, 2 * D, Ns, 1.f, W, N, Ns, B, 2 * D, (long long)Ns * 2 * D, part, 2 * D,
                          (long long)K * 2 * D, tcws, tcws_floats, st);
    if (rc) return rc;
    parts = S;
  }
  const int rem = N - S * Ns;
  if (rem > 0) {
    rc = launch_gemm_auto(0, 0, 1, K, 2 * D, rem, 1.f, W + (long long)S * Ns, N, 0, B + (long long)S * Ns * 2 * D, 2 * D, 0,
                          part + (size_t)parts * K * 2 * D, 2 * D, 0, nullptr, 0, st);
    if (rc) return rc;
    ++parts;
  }
  const long long kd = (long long)K * D;
  stein_diag_finalize_kernel<<<(int)min((long long)1024, (kd + 255) / 256), 256, 0, st>>>(part, parts, K, D, means, stds,
                                                                                         Hneg, gneg);
  return check_launch("stein_diag_finalize_kernel");
}

}  // namespace gvi
