// Stein sufficient statistics on the tcgen05 tensor cores (sm_100a), D <= 256:
//
//   M[k][j][i] = sum_n W[k,n] (x_nj - mu_kj) G[n,i]          (ng_estimator.py:173-188, the [N,D,D] outer products)
//
// Per component this is a [D x N] . [N x D] GEMM whose reduction dimension is the SAMPLE index, so both UMMA
// operands want the samples contiguous ("K-major"): the call first transposes X -> Xt[D][Np] (fp32) and
// G -> Gt[D][Np] (scaled by a power of two and split into fp16 hi / lo).  The kernel then runs, per component,
//   D[j, i] += A_hi G_hi^T + A_lo G_hi^T + A_hi G_lo^T,    A[j, n] = w_kn (x_nj - mu_kj) s_k  split into fp16 hi / lo,
// i.e. the same "2 x fp16" split precision as the log-density kernel (tc_logdens16.cu): the weights and the
// centring are fused into the A producer (formed in fp32 BEFORE the split), G_hi / G_lo tiles arrive by TMA.
// Samples are processed in stages of 32 (3-stage ring of 64 KB: two 128-row A tiles hi+lo, one 256-row G tile
// hi+lo, all K-major SWIZZLE_64B); 128-sample blocks whose weights are negligible (`active` mask of the
// importance-weight kernel) are skipped by every role.  The accumulator (2 x 128 x 256 fp32 = all 512 TMEM columns)
// is drained to global memory every FLUSH_BLOCKS blocks with round-to-nearest adds, because the tensor core's fp32
// accumulator truncates (~2^-24 per step) and a component can sum 10^4 steps.
// Warp roles (512 threads): 0 TMA (G tiles), 1 MMA issue (warp-uniform, elected lane), 2 TMEM allocator,
// 4-7 epilogue / drain, 8-15 A producers.  One work unit = (component, split of the block range); units with
// several splits write partial sums that a small kernel adds in a fixed order (deterministic).
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace gvi {
namespace st16 {
using namespace tcx;

constexpr int KS = 32;                  // samples per stage
constexpr int STAGES = 3;
constexpr int THREADS = 512;
constexpr int A_TILE = 128 * 64;        // 8 KB: 128 rows x 32 fp16
constexpr int B_TILE = 256 * 64;        // 16 KB: up to 256 rows x 32 fp16
constexpr int STAGE_BYTES = 4 * A_TILE + 2 * B_TILE;     // A_hi[0], A_hi[1], A_lo[0], A_lo[1], B_hi, B_lo = 64 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int FLUSH_BLOCKS = 32;        // 128-sample blocks between drains of the accumulator (measured, dense C5 size:
                                        // 16 -> 15.5 ms, error 0.9e-5; 32 -> 13.3 ms, 1.4e-5; 64 -> 12.2 ms, 2.5e-5 of the 1e-4 budget)

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t make_idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ float sub_f32_f16(float v, unsigned short h) {
  float d;
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(h), "h"((unsigned short)0xBC00), "f"(v));
  return d;
}
// Power of two s with b * s < 2^14 for the non-negative bound b (exponent clamped to +-40).
__device__ __forceinline__ float pow2_scale(float b) {
  const int e = (__float_as_int(b) >> 23) & 0xff;
  int se = 127 + 13 - (e - 127);
  se = se < 87 ? 87 : (se > 167 ? 167 : se);
  return __int_as_float(se << 23);
}

struct Barriers {
  uint64_t full_a[STAGES];
  uint64_t full_b[STAGES];
  uint64_t empty[STAGES];
  uint64_t acc_full;
  uint64_t acc_empty;
  uint32_t tmem_base;
};

// scal[0] = max |X|, scal[1] = max |G|
__global__ void __launch_bounds__(THREADS, 1)
stein_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                const float* __restrict__ Xt, int N, int Np, int D, int Dn, const float* __restrict__ means,
                const float* __restrict__ W, const uint8_t* __restrict__ active, const float* __restrict__ wmax,
                const float* __restrict__ minf, const float* __restrict__ scal, int K, int S, int flush_blocks,
                float* __restrict__ out /* [S][K][D][D] */) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Barriers* bars = reinterpret_cast<Barriers*>(smem + STAGES * STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = ceil_div(N, 128);
  const int mt = ceil_div(D, 128);                 // 128-row tiles of the output (1 or 2)
  const int units = K * S;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bars->full_a[s], 8);              // one elected arrive per producer warp
      mbar_init(&bars->full_b[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->acc_full, 1);
    mbar_init(&bars->acc_empty, 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  // every role walks the same sequence: units u = blockIdx.x, += gridDim.x; inside a unit the active blocks of
  // [b0, b1) in order, 4 stages per block
  auto unit_range = [&](int u, int& k, int& b0, int& b1) {
    k = u / S;
    const int s = u - k * S;
    b0 = (int)((long long)nblk * s / S);
    b1 = (int)((long long)nblk * (s + 1) / S);
  };
  auto is_active = [&](int k, int b) { return active == nullptr || active[(long long)k * nblk + b] != 0; };

  if (warp == 0) {
    // ---------------- TMA: G^T tiles (hi, lo), 32 samples x Dn rows ----------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int k, b0, b1;
        unit_range(u, k, b0, b1);
        for (int b = b0; b < b1; ++b) {
          if (!is_active(k, b)) continue;
          for (int q = 0; q < 128 / KS; ++q) {
            mbar_wait(&bars->empty[s], ph ^ 1);
            uint8_t* st = smem + s * STAGE_BYTES;
            mbar_arrive_expect_tx(&bars->full_b[s], 2u * (uint32_t)Dn * 64u);
            tma_load_2d(st + 4 * A_TILE, &map_hi, &bars->full_b[s], b * 128 + q * KS, 0);
            tma_load_2d(st + 4 * A_TILE + B_TILE, &map_lo, &bars->full_b[s], b * 128 + q * KS, 0);
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issue (whole warp, uniform; one elected lane issues) ----------------
    int s = 0;
    uint32_t ph = 0;
    uint32_t nflush = 0;
    const uint64_t desc0 = make_desc_sw64(smem_u32(smem));
    const uint32_t idesc = make_idesc_f16(Dn);
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int k, b0, b1;
      unit_range(u, k, b0, b1);
      int since = 0;             // blocks accumulated since the last drain
      bool fresh = true;         // accumulator empty (next MMA overwrites)
      int nact = 0;
      for (int b = b0; b < b1; ++b) nact += is_active(k, b) ? 1 : 0;
      int seen = 0;
      for (int b = b0; b < b1; ++b) {
        if (!is_active(k, b)) continue;
        ++seen;
        if (fresh) {             // wait until the epilogue has drained the previous contents
          mbar_wait(&bars->acc_empty, (nflush & 1) ^ 1);
          tc_fence_after();
        }
        for (int q = 0; q < 128 / KS; ++q) {
          mbar_wait(&bars->full_a[s], ph);
          mbar_wait(&bars->full_b[s], ph);
          tc_fence_after();
          const uint64_t st = desc0 + (uint64_t)((uint32_t)(s * STAGE_BYTES) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < KS / 16; ++ks) {
              const uint64_t b_hi = st + (uint64_t)((4 * A_TILE) >> 4) + (uint64_t)(ks * 2);
              const uint64_t b_lo = b_hi + (uint64_t)(B_TILE >> 4);
              for (int m = 0; m < mt; ++m) {
                const uint64_t a_hi = st + (uint64_t)((m * A_TILE) >> 4) + (uint64_t)(ks * 2);
                const uint64_t a_lo = a_hi + (uint64_t)((2 * A_TILE) >> 4);
                const uint32_t d = tmem_base + (uint32_t)(m * 256);
                umma_f16(d, a_hi, b_hi, idesc, (fresh && q == 0 && ks == 0) ? 0u : 1u);
                umma_f16(d, a_lo, b_hi, idesc, 1u);
                umma_f16(d, a_hi, b_lo, idesc, 1u);
              }
            }
            umma_commit(&bars->empty[s]);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        fresh = false;
        if (++since == flush_blocks || seen == nact) {
          if (elect_one()) umma_commit(&bars->acc_full);
          __syncwarp();
          ++nflush;
          since = 0;
          fresh = true;
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ---------------- epilogue: drain the accumulator (round-to-nearest adds in global memory) -----------------
    const int q = warp - 4;
    uint32_t nflush = 0;
    const float sg = pow2_scale(scal[1]);
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int k, b0, b1;
      unit_range(u, k, b0, b1);
      const int sidx = u - k * S;
      float* Mk = out + ((long long)sidx * K + k) * D * D;
      const float inv = 1.0f / (pow2_scale(wmax[k] * (scal[0] + minf[k])) * sg);
      int nact = 0;
      for (int b = b0; b < b1; ++b) nact += is_active(k, b) ? 1 : 0;
      const int ndrain = ceil_div(nact, flush_blocks);
      if (ndrain == 0) {         // nothing carries weight: the sum is zero
        for (int e = (q * 32 + lane); e < D * D; e += 128) Mk[e] = 0.f;
        continue;
      }
      for (int dr = 0; dr < ndrain; ++dr, ++nflush) {
        mbar_wait(&bars->acc_full, nflush & 1);
        tc_fence_after();
        for (int m = 0; m < mt; ++m) {
          const int j = m * 128 + 32 * q + lane;
          for (int cb = 0; cb < ceil_div(D, 32); ++cb) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(m * 256 + cb * 32), v);
            tmem_ld_wait();
            if (j < D) {
              float* dst = Mk + (long long)j * D + cb * 32;
              if ((D & 3) == 0) {      // 16-byte accesses: every lane owns a contiguous 128-byte piece of its row
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                  if (cb * 32 + e < D) {
                    float4 val = make_float4(__uint_as_float(v[e]) * inv, __uint_as_float(v[e + 1]) * inv,
                                             __uint_as_float(v[e + 2]) * inv, __uint_as_float(v[e + 3]) * inv);
                    if (dr != 0) {
                      const float4 old = *reinterpret_cast<const float4*>(dst + e);
                      val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w;
                    }
                    *reinterpret_cast<float4*>(dst + e) = val;
                  }
                }
              } else {
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                  if (cb * 32 + e < D) {
                    const float val = __uint_as_float(v[e]) * inv;
                    dst[e] = (dr == 0) ? val : dst[e] + val;
                  }
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->acc_empty);
      }
    }
  } else if (warp >= 8) {
    // ---------------- A producers: A[j, n] = w_kn (x_nj - mu_kj) s_k, fp16 hi / lo, K-major SWIZZLE_64B -----------
    // lane l: float4 c = l % 8 of the 32 samples of the stage, rows j = rsub + 32 q (rsub = 4 * (warp - 8) + l / 8)
    const int pw = warp - 8;
    const int c = lane & 7;
    const int rsub = 4 * pw + (lane >> 3);
    const uint32_t a_off = (uint32_t)(rsub * 64 + ((((c >> 1) ^ ((rsub >> 1) & 3)) << 4) + ((c & 1) << 3)));
    const uint32_t smem_base = smem_u32(smem);
    const int nq = mt * 4;                         // passes of 32 rows
    int s = 0;
    uint32_t ph = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int k, b0, b1;
      unit_range(u, k, b0, b1);
      const float sk = pow2_scale(wmax[k] * (scal[0] + minf[k]));
      float mu[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int j = rsub + 32 * q;
        mu[q] = (q < nq && j < D) ? __ldg(means + (long long)k * D + j) : 0.f;
      }
      const float* Wk = W + (long long)k * N;
      // the loads of the next stage are issued before the current one is converted (two register sets)
      struct Regs {
        float4 x[8];
        float w[4];
      };
      int b_ld = b0, q_ld = 0;                      // load cursor: next active block / stage inside it
      auto advance = [&]() {                        // move the cursor to the next stage of an active block
        while (b_ld < b1 && !is_active(k, b_ld)) ++b_ld;
      };
      auto issue = [&](Regs& R) {
        const int n0 = b_ld * 128 + q_ld * KS + 4 * c;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (q < nq) R.x[q] = __ldg(reinterpret_cast<const float4*>(Xt + (long long)(rsub + 32 * q) * Np + n0));
#pragma unroll
        for (int e = 0; e < 4; ++e) R.w[e] = (n0 + e < N) ? __ldg(Wk + n0 + e) : 0.f;
        if (++q_ld == 128 / KS) {
          q_ld = 0;
          ++b_ld;
          advance();
        }
      };
      auto emit = [&](const Regs& R) {
        float w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) w[e] = R.w[e] * sk;
        mbar_wait(&bars->empty[s], ph ^ 1);
        const uint32_t st = smem_base + (uint32_t)(s * STAGE_BYTES) + a_off;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (q < nq) {
            const float m = mu[q];
            const float v0 = (R.x[q].x - m) * w[0], v1 = (R.x[q].y - m) * w[1];
            const float v2 = (R.x[q].z - m) * w[2], v3 = (R.x[q].w - m) * w[3];
            const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
            const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01), u23 = *reinterpret_cast<const uint32_t*>(&h23);
            const __half2 l01 = __floats2half2_rn(sub_f32_f16(v0, (unsigned short)(u01 & 0xffffu)),
                                                  sub_f32_f16(v1, (unsigned short)(u01 >> 16)));
            const __half2 l23 = __floats2half2_rn(sub_f32_f16(v2, (unsigned short)(u23 & 0xffffu)),
                                                  sub_f32_f16(v3, (unsigned short)(u23 >> 16)));
            // row j = rsub + 32 q: tile q / 4, row (rsub + 32 (q % 4)) of the tile
            const uint32_t dst = st + (uint32_t)((q >> 2) * A_TILE + (q & 3) * 2048);
            sts64(dst, u01, u23);
            sts64(dst + 2 * A_TILE, *reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->full_a[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      };
      int nact = 0;
      for (int b = b0; b < b1; ++b) nact += is_active(k, b) ? 1 : 0;
      const int nst = nact * (128 / KS);
      advance();
      Regs Ra, Rb;
      if (nst > 0) issue(Ra);
      for (int i = 0; i < nst; i += 2) {
        if (i + 1 < nst) issue(Rb);
        emit(Ra);
        if (i + 1 < nst) {
          if (i + 2 < nst) issue(Ra);
          emit(Rb);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---- preparation kernels --------------------------------------------------------------------------------------
// out[0] = max |a|, accumulated with atomicMax on the float bits (values are non-negative)
__global__ void absmax_kernel(const float* __restrict__ a, long long n, float* __restrict__ out) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(a[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}
// wmax[k] = max_n W[k][n]
__global__ void __launch_bounds__(256) rowmax_kernel(const float* __restrict__ W, int N, float* __restrict__ wmax) {
  __shared__ float scratch[34];
  const float* row = W + (long long)blockIdx.x * N;
  float m = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) m = fmaxf(m, fabsf(row[n]));
  m = block_max(m, scratch);
  if (threadIdx.x == 0) wmax[blockIdx.x] = m;
}
// 32 x 32 tiled transposes: Xt[j][n] = X[n][j] (fp32, zero padded to [Dm][Np]);
// Gt_hi/lo[i][n] = split(G[n][i] * s), s = pow2_scale(scal[1]) (fp16, zero padded to [Dn][Np])
__global__ void __launch_bounds__(256)
transpose_x_kernel(const float* __restrict__ X, int N, int D, int Np, int Dm, float* __restrict__ Xt) {
  __shared__ float t[32][33];
  const int n0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int n = n0 + r, j = j0 + threadIdx.x;
    t[r][threadIdx.x] = (n < N && j < D) ? X[(long long)n * D + j] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int j = j0 + r, n = n0 + threadIdx.x;
    if (j < Dm && n < Np) Xt[(long long)j * Np + n] = t[threadIdx.x][r];
  }
}
__global__ void __launch_bounds__(256)
transpose_split_g_kernel(const float* __restrict__ G, int N, int D, int Np, int Dn, const float* __restrict__ scal,
                         __half* __restrict__ hi, __half* __restrict__ lo) {
  __shared__ float t[32][33];
  const float s = pow2_scale(scal[1]);
  const int n0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int n = n0 + r, i = i0 + threadIdx.x;
    t[r][threadIdx.x] = (n < N && i < D) ? G[(long long)n * D + i] * s : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = i0 + r, n = n0 + threadIdx.x;
    if (i < Dn && n < Np) {
      const float v = t[threadIdx.x][r];
      const __half h = __float2half_rn(v);
      hi[(long long)i * Np + n] = h;
      lo[(long long)i * Np + n] = __float2half_rn(v - __half2float(h));
    }
  }
}
// M[k] = sum_s part[s][k] in a fixed order
__global__ void reduce_splits_kernel(const float* __restrict__ part, long long per_split, int S, float* __restrict__ M) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < per_split; e += (long long)gridDim.x * blockDim.x) {
    float acc = part[e];
    for (int s = 1; s < S; ++s) acc += part[(long long)s * per_split + e];
    M[e] = acc;
  }
}

static int make_map(CUtensorMap* map, const void* base, int Dn, int Np) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) {
    set_last_error("cuTensorMapEncodeTiled is not available from the driver");
    return GVI_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)Np, (cuuint64_t)Dn};
  cuuint64_t gstride[1] = {(cuuint64_t)Np * 2};
  cuuint32_t box[2] = {KS, (cuuint32_t)Dn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled (G^T) failed with CUresult %d", (int)r);
    return GVI_ERR_CUDA;
  }
  return GVI_OK;
}

// 128-sample blocks between two drains of the accumulator (GMMVI_B200_STEIN_FLUSH overrides, read per call)
static int flush_blocks() {
  const char* e = getenv("GMMVI_B200_STEIN_FLUSH");
  const int v = e ? atoi(e) : 0;
  return v > 0 ? v : FLUSH_BLOCKS;
}
static int num_splits(int K, int nblk) {
  int S = 1;
  if (K < 148) S = (148 + K - 1) / K;
  if (S > nblk) S = nblk;
  return S < 1 ? 1 : S;
}

}  // namespace st16

// below D = 48 the 128-row A tile is mostly padding and the preparation launches outweigh the SIMT kernel
bool stein_tc_supported(int N, int D) { return D >= 48 && D <= 256 && N >= 1; }

// floats of workspace: Xt [Dm][Np] + Gt hi/lo [Dn][Np] halves + wmax[K] + minf[K] + scal[2] + partials
size_t stein_tc_workspace_floats(int N, int K, int D) {
  const long long Np = (long long)ceil_div(N, 128) * 128;
  const int Dm = ceil_div(D, 128) * 128, Dn = ceil_div(D, 16) * 16;
  const int S = st16::num_splits(K, ceil_div(N, 128));
  size_t f = (size_t)Dm * Np + (size_t)Dn * Np /* two fp16 arrays */ + 2 * (size_t)K + 64;
  if (S > 1) f += (size_t)S * K * D * D;
  return f + 64;
}

int launch_stein_stats_tc(const float* X, int N, int D, const float* means, const float* W, const uint8_t* active,
                          const float* G, int K, float* M, float* ws, cudaStream_t st) {
  using namespace st16;
  const int nblk = ceil_div(N, 128);
  const int Np = nblk * 128;
  const int Dm = ceil_div(D, 128) * 128, Dn = ceil_div(D, 16) * 16;
  const int S = num_splits(K, nblk);
  float* Xt = ws;
  __half* Ghi = reinterpret_cast<__half*>(Xt + (size_t)Dm * Np);
  __half* Glo = Ghi + (size_t)Dn * Np;
  float* wmaxp = reinterpret_cast<float*>(Glo + (size_t)Dn * Np);
  float* minf = wmaxp + K;
  float* scal = minf + K;
  float* part = scal + 64;
  part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(part) + 15) & ~uintptr_t(15));
  cudaMemsetAsync(scal, 0, 2 * sizeof(float), st);
  absmax_kernel<<<296, 256, 0, st>>>(X, (long long)N * D, scal);
  absmax_kernel<<<296, 256, 0, st>>>(G, (long long)N * D, scal + 1);
  rowmax_kernel<<<K, 256, 0, st>>>(W, N, wmaxp);
  rowmax_kernel<<<K, 256, 0, st>>>(means, D, minf);
  dim3 tb(32, 8);
  transpose_x_kernel<<<dim3(Np / 32, Dm / 32), tb, 0, st>>>(X, N, D, Np, Dm, Xt);
  transpose_split_g_kernel<<<dim3(Np / 32, ceil_div(Dn, 32)), tb, 0, st>>>(G, N, D, Np, Dn, scal, Ghi, Glo);
  int rc = check_launch("stein_tc preparation");
  if (rc) return rc;
  CUtensorMap map_hi, map_lo;
  if ((rc = make_map(&map_hi, Ghi, Dn, Np))) return rc;
  if ((rc = make_map(&map_lo, Glo, Dn, Np))) return rc;
  static int num_sms = 0;
  static unsigned long long dev_mask = 0;       // per device: the attributes below are per device
  if (first_call_on_device(dev_mask)) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(stein_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("stein_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      num_sms = 0;
      return GVI_ERR_CUDA;
    }
  }
  const int units = K * S;
  const int grid = units < num_sms ? units : num_sms;
  float* outp = (S > 1) ? part : M;
  stein_tc_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(map_hi, map_lo, Xt, N, Np, D, Dn, means, W, active, wmaxp, minf,
                                                    scal, K, S, flush_blocks(), outp);
  if ((rc = check_launch("stein_tc_kernel"))) return rc;
  if (S > 1) {
    const long long per = (long long)K * D * D;
    reduce_splits_kernel<<<(unsigned)((per + 255) / 256 < 2048 ? (per + 255) / 256 : 2048), 256, 0, st>>>(part, per, S, M);
    rc = check_launch("reduce_splits_kernel");
  }
  return rc;
}

}  // namespace gvi
