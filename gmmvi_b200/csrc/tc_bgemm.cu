// Batched GEMM on the tcgen05 tensor cores in 3xTF32 split precision (fp32-grade results):
//     C[b] (M x N) = alpha * A[b] (M x K) * B[b]^T (N x K),   operands pre-split into TF32 hi / lo parts.
// Used for the D x D x D products of the component update (B = L^T R L), the Stein finalisation (P M), the
// precision matrices (Linv^T Linv) and the MORE un-whitening; the reference issues these as tf.matmul on
// [D, D] tensors inside per-component Python loops (ng_based_component_updater.py:107-112, 176-180, 458;
// ng_estimator.py:183-186; least_squares.py:184).
//
// Persistent, warp-specialised: warp 0 = TMA producer (A hi/lo 128 x 32, B hi/lo NT x 32 per k-block, 128B swizzle,
// 2-stage ring), warp 1 = MMA issuer (3 tcgen05.mma per 8-wide k-step, fp32 accumulators in TMEM, double buffered),
// warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld -> alpha -> global).
//
// Long reductions (the MORE normal equations Phi^T W Phi reduce over thousands of samples, least_squares.py:60-75) are
// cut into segments of `kseg` k-blocks: the tensor core's fp32 accumulator truncates, so every segment starts a fresh
// accumulator (the two TMEM buffers alternate) and the epilogue adds it to C in global memory with round-to-nearest
// adds, the old values of C prefetched one 32-column chunk ahead.  `beta` = 1 adds to the existing C (trailing update of
// the blocked Cholesky), `lower_only` enumerates only the tiles that touch the lower triangle.
#include "tc_common.cuh"
#include <stdlib.h>
#include <algorithm>
#include <cuda_fp16.h>

namespace gvi {
namespace tcg {
using namespace tcx;

constexpr int TILE_M = 128;
constexpr int KBLK = 32;
constexpr int STAGES = 2;
constexpr int THREADS = 256;
constexpr int A_BYTES = TILE_M * 128;
constexpr int B_BYTES = 256 * 128;
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
constexpr int EPI_PITCH = 33;                                     // floats per staged accumulator row
constexpr int EPI_BYTES = 4 * 32 * EPI_PITCH * 4;                 // one 32 x 32 staging tile per epilogue warp
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + EPI_BYTES;
constexpr int TMEM_COLS = 512;
constexpr int ACC_COLS = 256;

struct Barriers {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

// kind::f16 with fp16 operands, fp32 accumulate, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc_h16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_h16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// H16 = false: TF32 hi / lo operands (32 floats per 128-byte k-block row, kind::tf32);
// H16 = true:  fp16 hi / lo operands of matrices pre-scaled by a power of two per batch entry (64 halves per k-block
//              row, kind::f16 at twice the TF32 rate and half the operand bytes); alpha_b[b] undoes the scales.
template <bool H16>
__global__ void __launch_bounds__(THREADS, 1)
tc_bgemm_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, int batch, int M,
                int N, int Kd, float alpha_all, const float* __restrict__ alpha_b, float* __restrict__ C, int ldc,
                long long strideC, float beta, int kseg, int lower_only, int tiles_per_batch) {
  constexpr int KELEMS = H16 ? 64 : 32;          // reduction elements per 128-byte row of a stage
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Barriers* bars = reinterpret_cast<Barriers*>(smem + STAGES * STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt_n = ceil_div(M, TILE_M), nt_n = ceil_div(N, ACC_COLS);
  const long long total = (long long)batch * tiles_per_batch;
  const int nkb = ceil_div(Kd, KELEMS);
  const int nseg = ceil_div(nkb, kseg);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->acc_full[b], 1);
      mbar_init(&bars->acc_empty[b], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  auto decode = [&](long long w, int& b, int& m0, int& n0, int& nt) {
    b = (int)(w / tiles_per_batch);
    int r = (int)(w % tiles_per_batch);
    if (lower_only) {
      // row mt of the tile grid keeps the column tiles with n0 <= m0 + TILE_M - 1: min(nt_n, mt / 2 + 1) of them
      int mt = 0;
      for (;; ++mt) {
        const int cnt = min(nt_n, (mt * TILE_M + TILE_M - 1) / ACC_COLS + 1);
        if (r < cnt) break;
        r -= cnt;
      }
      m0 = mt * TILE_M;
      n0 = r * ACC_COLS;
    } else {
      m0 = (r / nt_n) * TILE_M;
      n0 = (r % nt_n) * ACC_COLS;
    }
    nt = min(ACC_COLS, (N - n0 + 15) & ~15);      // MMA N: multiple of 16, rows beyond N are zero-filled by TMA
  };

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (long long w = blockIdx.x; w < total; w += gridDim.x) {
        int b, m0, n0, nt;
        decode(w, b, m0, n0, nt);
        const int nrows_b = (nt + 31) & ~31;     // TMA boxes are 32 rows high
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&bars->empty[s], ph ^ 1);
          uint8_t* st = smem + s * STAGE_BYTES;
          mbar_arrive_expect_tx(&bars->full[s], 2u * A_BYTES + 2u * nrows_b * 128u);
          tma_load_3d(st, &mapAh, &bars->full[s], kb * KELEMS, m0, b);
          tma_load_3d(st + A_BYTES, &mapAl, &bars->full[s], kb * KELEMS, m0, b);
          for (int r = 0; r < nrows_b; r += 32) {
            tma_load_3d(st + 2 * A_BYTES + r * 128, &mapBh, &bars->full[s], kb * KELEMS, n0 + r, b);
            tma_load_3d(st + 2 * A_BYTES + B_BYTES + r * 128, &mapBl, &bars->full[s], kb * KELEMS, n0 + r, b);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      long long it = 0;
      for (long long w = blockIdx.x; w < total; w += gridDim.x) {
        int b, m0, n0, nt;
        decode(w, b, m0, n0, nt);
        const uint32_t idesc = H16 ? make_idesc_h16(nt) : make_idesc(nt);
        for (int seg = 0; seg < nseg; ++seg, ++it) {
          const int buf = (int)(it & 1);
          const uint32_t use = (uint32_t)(it >> 1);
          mbar_wait(&bars->acc_empty[buf], (use & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * ACC_COLS);
          const int kb_end = min(nkb, (seg + 1) * kseg);
          for (int kb = seg * kseg; kb < kb_end; ++kb) {
            mbar_wait(&bars->full[s], ph);
            tc_fence_after();
            const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
            const int first = kb == seg * kseg;
#pragma unroll
            for (int ks = 0; ks < KBLK / 8; ++ks) {
              const uint64_t a_hi = make_desc(st + ks * 32);
              const uint64_t a_lo = make_desc(st + A_BYTES + ks * 32);
              const uint64_t b_hi = make_desc(st + 2 * A_BYTES + ks * 32);
              const uint64_t b_lo = make_desc(st + 2 * A_BYTES + B_BYTES + ks * 32);
              if (H16) {
                umma_h16(d_tmem, a_hi, b_hi, idesc, (first && ks == 0) ? 0u : 1u);
                umma_h16(d_tmem, a_lo, b_hi, idesc, 1u);
                umma_h16(d_tmem, a_hi, b_lo, idesc, 1u);
              } else {
                umma_tf32(d_tmem, a_hi, b_hi, idesc, (first && ks == 0) ? 0u : 1u);
                umma_tf32(d_tmem, a_lo, b_hi, idesc, 1u);
                umma_tf32(d_tmem, a_hi, b_lo, idesc, 1u);
              }
            }
            umma_commit(&bars->empty[s]);
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
          umma_commit(&bars->acc_full[buf]);
        }
      }
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    long long it = 0;
    const bool vec = (ldc % 4 == 0) && (reinterpret_cast<uintptr_t>(C) % 16 == 0) && (strideC % 4 == 0);
    // beta = 1: the old C tile of the NEXT work item is pulled into L2 while this one is processed (one bulk prefetch
    // per row and lane); the trailing updates of the MORE Cholesky stream C from HBM with few MMAs per tile, and the
    // one-chunk-ahead register prefetch of the read-modify-write alone covers only ~1/3 of the HBM latency.
    auto prefetch_c = [&](long long wq) {
      if (beta != 0.f && vec && wq < total) {
        int b2, m2, n2, nt2;
        decode(wq, b2, m2, n2, nt2);
        const int row = m2 + 32 * q + lane;
        const int ncols = min(nt2, N - n2) & ~3;
        if (row < M && ncols > 0) {
          const float* src = C + b2 * strideC + (long long)row * ldc + n2;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(ncols * 4) : "memory");
        }
      }
    };
    prefetch_c(blockIdx.x);
    for (long long w = blockIdx.x; w < total; w += gridDim.x) {
      int b, m0, n0, nt;
      decode(w, b, m0, n0, nt);
      prefetch_c(w + gridDim.x);
      const int m = m0 + 32 * q + lane;
      float* crow = C + b * strideC + (long long)m * ldc + n0;
      const float alpha = alpha_b ? alpha_all * __ldg(alpha_b + b) : alpha_all;
      for (int seg = 0; seg < nseg; ++seg, ++it) {
        const int buf = (int)(it & 1);
        const uint32_t use = (uint32_t)(it >> 1);
        const bool add = seg > 0 || beta != 0.f;       // C (+)= alpha * acc
        if (vec && n0 + nt <= N && (nt & 31) == 0) {
          // Coalesced path (all columns of the tile in bounds, 16-byte aligned rows): a tcgen05.ld gives a lane one
          // ROW of 32 accumulator columns, so stores straight from those registers touch 32 cache lines per
          // instruction (this, not the MMAs, bounded the short trailing updates of the MORE Cholesky).  The chunk goes
          // through a 32 x 33 staging tile in shared memory instead and leaves as 4 rows x 128 contiguous bytes per
          // instruction; the old values of C for the read-modify-write arrive the same way, one chunk ahead.
          float* stg = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256) + q * 32 * EPI_PITCH;
          const int rr = lane >> 3, c4 = (lane & 7) * 4;
          float* cbase = C + b * strideC + (long long)(m0 + 32 * q + rr) * ldc + n0 + c4;
          const float bscale = seg > 0 ? 1.f : beta;
          float4 old[2][8];
          auto fetch = [&](int c, float4 (&o)[8]) {
            if (add && c * 32 < nt) {
#pragma unroll
              for (int t = 0; t < 8; ++t)
                if (m0 + 32 * q + 4 * t + rr < M)
                  o[t] = *reinterpret_cast<const float4*>(cbase + (long long)(4 * t) * ldc + c * 32);
            }
          };
          fetch(0, old[0]);
          mbar_wait(&bars->acc_full[buf], use & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < ACC_COLS / 32; ++c) {
            if (c * 32 < nt) {
              fetch(c + 1, old[(c + 1) & 1]);
              uint32_t v[32];
              tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * ACC_COLS + c * 32), v);
              tmem_ld_wait();
              __syncwarp();                           // the previous chunk has been read out of the staging tile
#pragma unroll
              for (int i = 0; i < 32; ++i) stg[lane * EPI_PITCH + i] = __uint_as_float(v[i]);
              __syncwarp();
              const float4(&o)[8] = old[c & 1];
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                const float* sp = stg + (4 * t + rr) * EPI_PITCH + c4;
                float4 r = make_float4(alpha * sp[0], alpha * sp[1], alpha * sp[2], alpha * sp[3]);
                if (m0 + 32 * q + 4 * t + rr < M) {
                  if (add) {
                    r.x = fmaf(bscale, o[t].x, r.x);
                    r.y = fmaf(bscale, o[t].y, r.y);
                    r.z = fmaf(bscale, o[t].z, r.z);
                    r.w = fmaf(bscale, o[t].w, r.w);
                  }
                  *reinterpret_cast<float4*>(cbase + (long long)(4 * t) * ldc + c * 32) = r;
                }
              }
            }
          }
        } else if (!add) {
          mbar_wait(&bars->acc_full[buf], use & 1);
          tc_fence_after();
          for (int c = 0; c * 32 < nt; ++c) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * ACC_COLS + c * 32), v);
            tmem_ld_wait();
            if (m < M) {
              const int ncol = min(32, N - (n0 + c * 32));
              if (vec && ncol == 32) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                  *reinterpret_cast<float4*>(crow + c * 32 + i) =
                      make_float4(alpha * __uint_as_float(v[i]), alpha * __uint_as_float(v[i + 1]),
                                  alpha * __uint_as_float(v[i + 2]), alpha * __uint_as_float(v[i + 3]));
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (i < ncol) crow[c * 32 + i] = alpha * __uint_as_float(v[i]);
              }
            }
          }
        } else {
          // read-modify-write with the old values of C fetched one chunk ahead (the first chunk before the wait for the
          // accumulator); this thread wrote them itself in the previous segment, or a previous kernel did (beta)
          const float bscale = seg > 0 ? 1.f : beta;
          float4 old[2][8];
          auto fetch = [&](int c, float4 (&o)[8]) {
            if (m < M && c * 32 < nt) {
              const int ncol = min(32, N - (n0 + c * 32));
              if (vec && ncol == 32) {
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = *reinterpret_cast<const float4*>(crow + c * 32 + 4 * i);
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  o[i].x = 4 * i + 0 < ncol ? crow[c * 32 + 4 * i + 0] : 0.f;
                  o[i].y = 4 * i + 1 < ncol ? crow[c * 32 + 4 * i + 1] : 0.f;
                  o[i].z = 4 * i + 2 < ncol ? crow[c * 32 + 4 * i + 2] : 0.f;
                  o[i].w = 4 * i + 3 < ncol ? crow[c * 32 + 4 * i + 3] : 0.f;
                }
              }
            }
          };
          fetch(0, old[0]);
          mbar_wait(&bars->acc_full[buf], use & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < ACC_COLS / 32; ++c) {
            if (c * 32 < nt) {
              fetch(c + 1, old[(c + 1) & 1]);
              uint32_t v[32];
              tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * ACC_COLS + c * 32), v);
              tmem_ld_wait();
              if (m < M) {
                const float4(&o)[8] = old[c & 1];
                const int ncol = min(32, N - (n0 + c * 32));
                if (vec && ncol == 32) {
#pragma unroll
                  for (int i = 0; i < 8; ++i)
                    *reinterpret_cast<float4*>(crow + c * 32 + 4 * i) =
                        make_float4(fmaf(alpha, __uint_as_float(v[4 * i]), bscale * o[i].x),
                                    fmaf(alpha, __uint_as_float(v[4 * i + 1]), bscale * o[i].y),
                                    fmaf(alpha, __uint_as_float(v[4 * i + 2]), bscale * o[i].z),
                                    fmaf(alpha, __uint_as_float(v[4 * i + 3]), bscale * o[i].w));
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    if (4 * i + 0 < ncol) crow[c * 32 + 4 * i + 0] = fmaf(alpha, __uint_as_float(v[4 * i + 0]), bscale * o[i].x);
                    if (4 * i + 1 < ncol) crow[c * 32 + 4 * i + 1] = fmaf(alpha, __uint_as_float(v[4 * i + 1]), bscale * o[i].y);
                    if (4 * i + 2 < ncol) crow[c * 32 + 4 * i + 2] = fmaf(alpha, __uint_as_float(v[4 * i + 2]), bscale * o[i].z);
                    if (4 * i + 3 < ncol) crow[c * 32 + 4 * i + 3] = fmaf(alpha, __uint_as_float(v[4 * i + 3]), bscale * o[i].w);
                  }
                }
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&bars->acc_empty[buf]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// in [b][R][C] (row pitch ld) -> hi / lo [b][C][R]  (transposed split)
__global__ void split_tf32_transpose_kernel(const float* __restrict__ in, int R, int Cc, int ld, long long stride_in,
                                            float* __restrict__ hi, float* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const float* src = in + b * stride_in;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? src[(long long)r * ld + c] : 0.f;
  }
  __syncthreads();
  const long long ob = (long long)b * R * Cc;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < Cc && r < R) {
      const float x = tile[threadIdx.x][i];
      const float h = to_tf32(x);
      hi[ob + (long long)c * R + r] = h;
      lo[ob + (long long)c * R + r] = to_tf32(x - h);
    }
  }
}

__global__ void split_tf32_strided_kernel(const float* __restrict__ in, int R, int Cc, int ld, long long stride_in,
                                          float* __restrict__ hi, float* __restrict__ lo) {
  const int b = blockIdx.y;
  const float* src = in + b * stride_in;
  const long long ob = (long long)b * R * Cc;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)R * Cc;
       e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / Cc), c = (int)(e % Cc);
    const float x = src[(long long)r * ld + c];
    const float h = to_tf32(x);
    hi[ob + e] = h;
    lo[ob + e] = to_tf32(x - h);
  }
}

// Same for 16-byte aligned operands with Cc % 4 == 0, ld % 4 == 0 and R * Cc < 2^31: one float4 per thread and step,
// 32-bit index arithmetic (the generic kernel spends most of its time in two 64-bit divisions per element).
__global__ void __launch_bounds__(256)
split_tf32_strided_vec4_kernel(const float* __restrict__ in, int R, int C4, int ld4, long long stride_in,
                               float* __restrict__ hi, float* __restrict__ lo) {
  const int b = blockIdx.y;
  const float4* __restrict__ src = reinterpret_cast<const float4*>(in + b * stride_in);
  float4* __restrict__ h4 = reinterpret_cast<float4*>(hi + (long long)b * R * C4 * 4);
  float4* __restrict__ l4 = reinterpret_cast<float4*>(lo + (long long)b * R * C4 * 4);
  const int total = R * C4;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int r = e / C4, c = e - r * C4;
    const float4 x = __ldg(src + ((long long)r * ld4 + c));
    float4 h, l;
    h.x = to_tf32(x.x); h.y = to_tf32(x.y); h.z = to_tf32(x.z); h.w = to_tf32(x.w);
    l.x = to_tf32(x.x - h.x); l.y = to_tf32(x.y - h.y); l.z = to_tf32(x.z - h.z); l.w = to_tf32(x.w - h.w);
    h4[e] = h;
    l4[e] = l;
  }
}

// hi / lo split of the symmetric matrix read from the LOWER triangle of `in`: out[b][r][c] = in[b][max(r,c)][min(r,c)]
// (what tf.linalg.cholesky sees of a not exactly symmetric input).  32 x 32 tiles through shared memory so that both
// the direct and the mirrored reads are coalesced.
__global__ void split_tf32_symlower_kernel(const float* __restrict__ in, int D, float* __restrict__ hi,
                                           float* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const float* src = in + (long long)b * D * D;
  const long long ob = (long long)b * D * D;
  const bool upper = c0 > r0;                    // tile strictly above the diagonal: read the mirrored tile, transposed
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int rr = (upper ? c0 : r0) + i, cc = (upper ? r0 : c0) + threadIdx.x;
    tile[i][threadIdx.x] = (rr < D && cc < D) ? src[(long long)rr * D + cc] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < D && c < D) {
      float x;
      if (upper) x = tile[threadIdx.x][i];
      else if (c0 == r0 && c > r) x = tile[threadIdx.x][i];      // diagonal tile: mirror inside the tile
      else x = tile[i][threadIdx.x];
      const float h = to_tf32(x);
      hi[ob + (long long)r * D + c] = h;
      lo[ob + (long long)r * D + c] = to_tf32(x - h);
    }
  }
}

// ---- fp16 hi / lo split of dense batches under one power-of-two scale per batch entry (stand-alone entry point) ----
__global__ void absmax_batch_kernel(const float* __restrict__ in, long long per_batch, unsigned* __restrict__ out) {
  const float* src = in + (long long)blockIdx.y * per_batch;
  float m = 0.f;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < per_batch;
       e += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(src[e]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out + blockIdx.y, __float_as_uint(m));      // non-negative floats order as ints
}
__global__ void split_h16_batch_kernel(const float* __restrict__ in, long long per_batch,
                                       const unsigned* __restrict__ mx, __half* __restrict__ hi,
                                       __half* __restrict__ lo) {
  const long long ob = (long long)blockIdx.y * per_batch;
  const float s = h16_scale_of(__uint_as_float(mx[blockIdx.y]));
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < per_batch;
       e += (long long)gridDim.x * blockDim.x) {
    const float v = in[ob + e] * s;
    const __half h = __float2half_rn(v);
    hi[ob + e] = h;
    lo[ob + e] = __float2half_rn(v - __half2float(h));
  }
}
__global__ void alpha_from_max_kernel(const unsigned* __restrict__ ma, const unsigned* __restrict__ mb, int batch,
                                      float* __restrict__ alpha_b) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < batch)
    alpha_b[b] = 1.f / (h16_scale_of(__uint_as_float(ma[b])) * h16_scale_of(__uint_as_float(mb[b])));
}

static int num_sms_cached() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

}  // namespace tcg

bool tc_gemm_supported(int M, int N, int Kd) { return M > 0 && N > 0 && Kd >= 4 && (Kd % 4 == 0); }

// Split (optionally transposing) a batch of row-major matrices into TF32 hi / lo parts, densely packed.
//   trans == 0: out[b][R][C] = in[b][R][C];  trans == 1: out[b][C][R] = in[b][R][C]^T
int launch_split_tf32(const float* in, int batch, int R, int Cc, int ld, long long stride_in, int trans, float* hi,
                      float* lo, cudaStream_t st) {
  if (batch <= 0 || R <= 0 || Cc <= 0) return GVI_OK;
  if (trans) {
    dim3 grid(ceil_div(Cc, 32), ceil_div(R, 32), batch), block(32, 8);
    tcg::split_tf32_transpose_kernel<<<grid, block, 0, st>>>(in, R, Cc, ld, stride_in, hi, lo);
    return check_launch("split_tf32_transpose_kernel");
  }
  const bool vec4 = Cc % 4 == 0 && ld % 4 == 0 && stride_in % 4 == 0 && (long long)R * Cc < 2147483647LL &&
                    ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) % 16 == 0) &&
                    ((long long)R * Cc) % 4 == 0;
  if (vec4) {
    dim3 grid((unsigned)min((long long)64, ((long long)R * (Cc / 4) + 255) / 256), batch);
    tcg::split_tf32_strided_vec4_kernel<<<grid, 256, 0, st>>>(in, R, Cc / 4, ld / 4, stride_in, hi, lo);
    return check_launch("split_tf32_strided_vec4_kernel");
  }
  dim3 grid((unsigned)min((long long)1024, ((long long)R * Cc + 255) / 256), batch);
  tcg::split_tf32_strided_kernel<<<grid, 256, 0, st>>>(in, R, Cc, ld, stride_in, hi, lo);
  return check_launch("split_tf32_strided_kernel");
}

// C[b] = alpha * A[b] B[b]^T + beta * C[b] with A [b][M][Kd], B [b][N][Kd] densely packed hi / lo operands.
//   kseg_kblocks > 0: the reduction is cut into segments of that many 32-wide k-blocks, added to C with round-to-nearest
//                     adds (0 = one segment);  beta is 0 or 1;  lower_only: tiles strictly above the diagonal are skipped.
static int tiles_lower(int mt_n, int nt_n) {
  int t = 0;
  for (int mt = 0; mt < mt_n; ++mt) t += std::min(nt_n, (mt * tcg::TILE_M + tcg::TILE_M - 1) / tcg::ACC_COLS + 1);
  return t;
}
static int launch_tc_bgemm_any(bool h16, int batch, int M, int N, int Kd, float alpha, const float* alpha_b,
                               const void* Ah, const void* Al, const void* Bh, const void* Bl, float* C, int ldc,
                               long long strideC, float beta, int kseg_kblocks, int lower_only, cudaStream_t st,
                               long long opStrideA = 0, long long opStrideB = 0) {
  const long long sOa = opStrideA ? opStrideA : (long long)M * Kd, sOb = opStrideB ? opStrideB : (long long)N * Kd;
  if (batch <= 0 || M <= 0 || N <= 0) return GVI_OK;
  if (!tc_gemm_supported(M, N, Kd) || (h16 && Kd % 8 != 0)) {
    set_last_error("tc_bgemm: unsupported shape M=%d N=%d K=%d", M, N, Kd);
    return GVI_ERR_UNSUPPORTED;
  }
  CUtensorMap mAh, mAl, mBh, mBl;
  int rc;
  if (h16) {
    if ((rc = tcx::make_map_3d_h16(&mAh, Ah, Kd, M, batch, Kd, sOa, 128))) return rc;
    if ((rc = tcx::make_map_3d_h16(&mAl, Al, Kd, M, batch, Kd, sOa, 128))) return rc;
    if ((rc = tcx::make_map_3d_h16(&mBh, Bh, Kd, N, batch, Kd, sOb, 32))) return rc;
    if ((rc = tcx::make_map_3d_h16(&mBl, Bl, Kd, N, batch, Kd, sOb, 32))) return rc;
  } else {
    if ((rc = tcx::make_map_3d(&mAh, (const float*)Ah, Kd, M, batch, Kd, sOa, 128))) return rc;
    if ((rc = tcx::make_map_3d(&mAl, (const float*)Al, Kd, M, batch, Kd, sOa, 128))) return rc;
    if ((rc = tcx::make_map_3d(&mBh, (const float*)Bh, Kd, N, batch, Kd, sOb, 32))) return rc;
    if ((rc = tcx::make_map_3d(&mBl, (const float*)Bl, Kd, N, batch, Kd, sOb, 32))) return rc;
  }
  static unsigned long long attr_mask = 0;
  if (first_call_on_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(tcg::tc_bgemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         tcg::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(tcg::tc_bgemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("tc_bgemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return GVI_ERR_CUDA;
    }
  }
  const int mt_n = ceil_div(M, tcg::TILE_M), nt_n = ceil_div(N, tcg::ACC_COLS);
  const int tpb = lower_only ? tiles_lower(mt_n, nt_n) : mt_n * nt_n;
  const int nkb = ceil_div(Kd, h16 ? 64 : 32);
  const int kseg = (kseg_kblocks > 0 && kseg_kblocks < nkb) ? kseg_kblocks : nkb;
  const long long total = (long long)batch * tpb;
  const int grid = (int)min((long long)tcg::num_sms_cached(), total);
  if (h16)
    tcg::tc_bgemm_kernel<true><<<grid, tcg::THREADS, tcg::SMEM_BYTES, st>>>(
        mAh, mAl, mBh, mBl, batch, M, N, Kd, alpha, alpha_b, C, ldc, strideC, beta, kseg, lower_only ? 1 : 0, tpb);
  else
    tcg::tc_bgemm_kernel<false><<<grid, tcg::THREADS, tcg::SMEM_BYTES, st>>>(
        mAh, mAl, mBh, mBl, batch, M, N, Kd, alpha, alpha_b, C, ldc, strideC, beta, kseg, lower_only ? 1 : 0, tpb);
  return check_launch("tc_bgemm_kernel");
}
int launch_tc_bgemm_ex(int batch, int M, int N, int Kd, float alpha, const float* Ah, const float* Al, const float* Bh,
                       const float* Bl, float* C, int ldc, long long strideC, float beta, int kseg_kblocks,
                       int lower_only, cudaStream_t st) {
  return launch_tc_bgemm_any(false, batch, M, N, Kd, alpha, nullptr, Ah, Al, Bh, Bl, C, ldc, strideC, beta,
                             kseg_kblocks, lower_only, st);
}
// fp16 hi / lo operands ([b][M][Kd] / [b][N][Kd] halves, Kd % 8 == 0) of matrices scaled by powers of two;
// C[b] = alpha * alpha_b[b] * A[b] B[b]^T + beta * C[b]; kseg_kblocks counts blocks of 64.
int launch_tc_bgemm_h16_ex(int batch, int M, int N, int Kd, float alpha, const float* alpha_b, const void* Ah,
                           const void* Al, const void* Bh, const void* Bl, float* C, int ldc, long long strideC,
                           float beta, int kseg_kblocks, int lower_only, cudaStream_t st) {
  return launch_tc_bgemm_any(true, batch, M, N, Kd, alpha, alpha_b, Ah, Al, Bh, Bl, C, ldc, strideC, beta,
                             kseg_kblocks, lower_only, st);
}
// the same with explicit distances (in halves) between the operand matrices of two batch entries: the B operand may
// be the leading rows of a taller matrix
int launch_tc_bgemm_h16_strided(int batch, int M, int N, int Kd, float alpha, const float* alpha_b, const void* Ah,
                                const void* Al, long long opStrideA, const void* Bh, const void* Bl,
                                long long opStrideB, float* C, int ldc, long long strideC, float beta,
                                int kseg_kblocks, int lower_only, cudaStream_t st) {
  return launch_tc_bgemm_any(true, batch, M, N, Kd, alpha, alpha_b, Ah, Al, Bh, Bl, C, ldc, strideC, beta,
                             kseg_kblocks, lower_only, st, opStrideA, opStrideB);
}
int launch_tc_bgemm(int batch, int M, int N, int Kd, float alpha, const float* Ah, const float* Al, const float* Bh,
                    const float* Bl, float* C, int ldc, long long strideC, cudaStream_t st) {
  return launch_tc_bgemm_ex(batch, M, N, Kd, alpha, Ah, Al, Bh, Bl, C, ldc, strideC, 0.f, 0, 0, st);
}

// C[b] = alpha * opA(A[b]) opB(B[b]) for row-major fp32 inputs, like launch_bgemm; ws holds the split operands:
// 2 * batch * (M*Kd + N*Kd) floats.
size_t tc_gemm_workspace_floats(int batch, int M, int N, int Kd) {
  return (size_t)2 * batch * ((size_t)M * Kd + (size_t)N * Kd) + 256;
}
int launch_tc_gemm_ex(int transA, int transB, int batch, int M, int N, int Kd, float alpha, const float* A, int lda,
                      long long strideA, const float* B, int ldb, long long strideB, float* C, int ldc,
                      long long strideC, float beta, int kseg_kblocks, int lower_only, float* ws, cudaStream_t st) {
  float* Ah = ws;
  float* Al = Ah + (size_t)batch * M * Kd;
  float* Bh = Al + (size_t)batch * M * Kd;
  float* Bl = Bh + (size_t)batch * N * Kd;
  int rc;
  // A operand must be [M][Kd]: transA == 0 -> A is [M][Kd] already; transA == 1 -> A is [Kd][M], transpose it
  if (transA) rc = launch_split_tf32(A, batch, Kd, M, lda, strideA, 1, Ah, Al, st);
  else        rc = launch_split_tf32(A, batch, M, Kd, lda, strideA, 0, Ah, Al, st);
  if (rc) return rc;
  // B operand must be [N][Kd]: transB == 0 -> B is [Kd][N], transpose it; transB == 1 -> B is [N][Kd] already
  if (A == B && lda == ldb && strideA == strideB && M == N && (transA != 0) != (transB != 0)) {
    Bh = Ah;          // X^T X or X X^T: both operands are the same split
    Bl = Al;
  } else {
    if (transB) rc = launch_split_tf32(B, batch, N, Kd, ldb, strideB, 0, Bh, Bl, st);
    else        rc = launch_split_tf32(B, batch, Kd, N, ldb, strideB, 1, Bh, Bl, st);
    if (rc) return rc;
  }
  return launch_tc_bgemm_ex(batch, M, N, Kd, alpha, Ah, Al, Bh, Bl, C, ldc, strideC, beta, kseg_kblocks, lower_only,
                            st);
}
int launch_tc_gemm(int transA, int transB, int batch, int M, int N, int Kd, float alpha, const float* A, int lda,
                   long long strideA, const float* B, int ldb, long long strideB, float* C, int ldc,
                   long long strideC, float* ws, cudaStream_t st) {
  return launch_tc_gemm_ex(transA, transB, batch, M, N, Kd, alpha, A, lda, strideA, B, ldb, strideB, C, ldc, strideC,
                           0.f, 0, 0, ws, st);
}

int launch_bgemm(int transA, int transB, int batch, int M, int N, int Kd, float alpha, const float* A, int lda,
                 long long strideA, const float* B, int ldb, long long strideB, float* C, int ldc,
                 long long strideC, cudaStream_t st);

bool tc_gemm_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GMMVI_B200_TC_GEMM");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// Tensor-core GEMM when the shape allows it and the caller provided split workspace, else the SIMT tile engine.
int launch_gemm_auto(int transA, int transB, int batch, int M, int N, int Kd, float alpha, const float* A, int lda,
                     long long strideA, const float* B, int ldb, long long strideB, float* C, int ldc,
                     long long strideC, float* ws, size_t ws_floats, cudaStream_t st) {
  if (tc_gemm_enabled() && ws != nullptr && tc_gemm_supported(M, N, Kd) && M >= 64 && N >= 32 &&
      ws_floats >= tc_gemm_workspace_floats(batch, M, N, Kd) && (reinterpret_cast<uintptr_t>(ws) % 16 == 0))
    return launch_tc_gemm(transA, transB, batch, M, N, Kd, alpha, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, ws,
                          st);
  return launch_bgemm(transA, transB, batch, M, N, Kd, alpha, A, lda, strideA, B, ldb, strideB, C, ldc, strideC, st);
}

// The two D^3 products of the whitened update, T = Rlow L and B = L^T T with Rlow = the symmetric matrix in the lower
// triangle of R (update.cu), on the tensor cores with the operand splits shared: the mirror of R is folded into its
// split, and the transposed split of L serves as B operand of the first and as A operand of the second product
// (5 kernels instead of mirror + 4 splits + 2 GEMMs).  ws: 6 K D^2 floats.  Returns 1 when the shape / workspace does
// not allow the tensor-core path (the caller then uses the generic route), 0 on success, < 0 on error.
size_t tc_update_products_workspace_floats(int K, int D) { return (size_t)6 * K * D * D + 256; }
int launch_update_products_tc(const float* R, const float* L, int K, int D, float* T, float* Bm, float* ws,
                              size_t ws_floats, cudaStream_t st) {
  if (!(tc_gemm_enabled() && ws != nullptr && tc_gemm_supported(D, D, D) && D >= 64 &&
        ws_floats >= tc_update_products_workspace_floats(K, D) && (reinterpret_cast<uintptr_t>(ws) % 16 == 0)))
    return 1;
  const size_t n = (size_t)K * D * D;
  const long long DD = (long long)D * D;
  float *Rh = ws, *Rl = ws + n, *Lth = ws + 2 * n, *Ltl = ws + 3 * n, *Tth = ws + 4 * n, *Ttl = ws + 5 * n;
  dim3 grid(ceil_div(D, 32), ceil_div(D, 32), K), block(32, 8);
  tcg::split_tf32_symlower_kernel<<<grid, block, 0, st>>>(R, D, Rh, Rl);
  int rc = check_launch("split_tf32_symlower_kernel");
  if (rc) return rc;
  if ((rc = launch_split_tf32(L, K, D, D, D, DD, 1, Lth, Ltl, st))) return rc;            // L^T, [c][r]
  if ((rc = launch_tc_bgemm(K, D, D, D, 1.f, Rh, Rl, Lth, Ltl, T, D, DD, st))) return rc;     // T = Rlow L
  if ((rc = launch_split_tf32(T, K, D, D, D, DD, 1, Tth, Ttl, st))) return rc;            // T^T
  return launch_tc_bgemm(K, D, D, D, 1.f, Lth, Ltl, Tth, Ttl, Bm, D, DD, st);             // B = L^T T
}

}  // namespace gvi

using namespace gvi;

extern "C" int gvi_tc_bgemm_supported(int M, int N, int Kd) { return tc_gemm_supported(M, N, Kd) ? 1 : 0; }

extern "C" size_t gvi_tc_bgemm_workspace(int batch, int M, int N, int Kd) {
  return tc_gemm_workspace_floats(batch, M, N, Kd) * sizeof(float);
}

extern "C" int gvi_tc_bgemm_ex_f32(int transA, int transB, int batch, int M, int N, int Kd, float alpha,
                                   const float* A, int lda, long long strideA, const float* B, int ldb,
                                   long long strideB, float* C, int ldc, long long strideC, float beta,
                                   int kseg_kblocks, int lower_only, void* ws, size_t ws_bytes, void* stream) {
  GVI_REQUIRE(batch >= 0 && M >= 0 && N >= 0 && Kd >= 0 && kseg_kblocks >= 0, "gvi_tc_bgemm_ex_f32: bad sizes");
  GVI_REQUIRE(beta == 0.f || beta == 1.f, "gvi_tc_bgemm_ex_f32: beta must be 0 or 1");
  if (batch == 0 || M == 0 || N == 0) return GVI_OK;
  GVI_REQUIRE(A && B && C && ws, "gvi_tc_bgemm_ex_f32: null pointer");
  if (!tc_gemm_supported(M, N, Kd)) {
    set_last_error("gvi_tc_bgemm_ex_f32: unsupported shape (K must be a positive multiple of 4)");
    return GVI_ERR_UNSUPPORTED;
  }
  if (ws_bytes < gvi_tc_bgemm_workspace(batch, M, N, Kd)) {
    set_last_error("gvi_tc_bgemm_ex_f32: workspace %zu < %zu", ws_bytes, gvi_tc_bgemm_workspace(batch, M, N, Kd));
    return GVI_ERR_WORKSPACE;
  }
  return launch_tc_gemm_ex(transA, transB, batch, M, N, Kd, alpha, A, lda, strideA, B, ldb, strideB, C, ldc, strideC,
                           beta, kseg_kblocks, lower_only, (float*)ws, (cudaStream_t)stream);
}

extern "C" int gvi_tc_bgemm_f32(int transA, int transB, int batch, int M, int N, int Kd, float alpha, const float* A,
                                int lda, long long strideA, const float* B, int ldb, long long strideB, float* C,
                                int ldc, long long strideC, void* ws, size_t ws_bytes, void* stream) {
  GVI_REQUIRE(batch >= 0 && M >= 0 && N >= 0 && Kd >= 0, "gvi_tc_bgemm_f32: bad sizes");
  if (batch == 0 || M == 0 || N == 0) return GVI_OK;
  GVI_REQUIRE(A && B && C && ws, "gvi_tc_bgemm_f32: null pointer");
  if (!tc_gemm_supported(M, N, Kd)) {
    set_last_error("gvi_tc_bgemm_f32: unsupported shape (K must be a positive multiple of 4)");
    return GVI_ERR_UNSUPPORTED;
  }
  if (ws_bytes < gvi_tc_bgemm_workspace(batch, M, N, Kd)) {
    set_last_error("gvi_tc_bgemm_f32: workspace %zu < %zu", ws_bytes, gvi_tc_bgemm_workspace(batch, M, N, Kd));
    return GVI_ERR_WORKSPACE;
  }
  return launch_tc_gemm(transA, transB, batch, M, N, Kd, alpha, A, lda, strideA, B, ldb, strideB, C, ldc, strideC,
                        (float*)ws, (cudaStream_t)stream);
}

// Stand-alone 2 x fp16 split product (used by the tests; the MORE estimator writes its operands pre-split):
// C[b] = alpha * A[b] B[b]^T + beta * C[b] for dense A [b][M][Kd], B [b][N][Kd], C [b][M][N], Kd % 8 == 0.
extern "C" size_t gvi_tc_bgemm_h16_workspace(int batch, int M, int N, int Kd) {
  if (batch <= 0 || M <= 0 || N <= 0 || Kd <= 0) return 0;
  const size_t halves = (size_t)2 * batch * ((size_t)M + N) * Kd;
  return ((halves * 2 + 255) / 256) * 256 + (size_t)3 * batch * sizeof(float) + 256;
}
extern "C" int gvi_tc_bgemm_h16_f32(int batch, int M, int N, int Kd, float alpha, const float* A, const float* B,
                                    float* C, float beta, int kseg_kblocks, int lower_only, void* ws, size_t ws_bytes,
                                    void* stream) {
  GVI_REQUIRE(batch >= 0 && M >= 0 && N >= 0 && Kd >= 0 && kseg_kblocks >= 0, "gvi_tc_bgemm_h16_f32: bad sizes");
  GVI_REQUIRE(beta == 0.f || beta == 1.f, "gvi_tc_bgemm_h16_f32: beta must be 0 or 1");
  if (batch == 0 || M == 0 || N == 0) return GVI_OK;
  GVI_REQUIRE(A && B && C && ws, "gvi_tc_bgemm_h16_f32: null pointer");
  if (Kd % 8 != 0) {
    set_last_error("gvi_tc_bgemm_h16_f32: K must be a multiple of 8");
    return GVI_ERR_UNSUPPORTED;
  }
  if (ws_bytes < gvi_tc_bgemm_h16_workspace(batch, M, N, Kd) || reinterpret_cast<uintptr_t>(ws) % 16 != 0) {
    set_last_error("gvi_tc_bgemm_h16_f32: workspace too small or not 16-byte aligned");
    return GVI_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t na = (size_t)batch * M * Kd, nb = (size_t)batch * N * Kd;
  __half* Ah = (__half*)ws;
  __half* Al = Ah + na;
  __half* Bh = Al + na;
  __half* Bl = Bh + nb;
  unsigned* mx = (unsigned*)((char*)ws + ((2 * (na + nb) * 2 + 255) / 256) * 256);
  float* alpha_b = (float*)(mx + 2 * batch);
  const bool same = A == B && M == N;
  cudaMemsetAsync(mx, 0, (size_t)2 * batch * sizeof(unsigned), st);
  const long long pa = (long long)M * Kd, pb = (long long)N * Kd;
  dim3 ga((unsigned)std::min<long long>(256, (pa + 255) / 256), batch), gb((unsigned)std::min<long long>(256, (pb + 255) / 256), batch);
  tcg::absmax_batch_kernel<<<ga, 256, 0, st>>>(A, pa, mx);
  tcg::absmax_batch_kernel<<<gb, 256, 0, st>>>(B, pb, mx + batch);
  tcg::split_h16_batch_kernel<<<ga, 256, 0, st>>>(A, pa, mx, Ah, Al);
  if (!same) tcg::split_h16_batch_kernel<<<gb, 256, 0, st>>>(B, pb, mx + batch, Bh, Bl);
  tcg::alpha_from_max_kernel<<<ceil_div(batch, 128), 128, 0, st>>>(mx, mx + batch, batch, alpha_b);
  int rc = check_launch("gvi_tc_bgemm_h16_f32: split kernels");
  if (rc) return rc;
  return launch_tc_bgemm_h16_ex(batch, M, N, Kd, alpha, alpha_b, Ah, Al, same ? Ah : Bh, same ? Al : Bl, C, N,
                                (long long)M * N, beta, kseg_kblocks, lower_only, st);
}
