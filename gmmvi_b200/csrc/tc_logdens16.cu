// Full-covariance component log densities, tcgen05 kind::f16 in "2 x fp16" split precision with the
// component's inverse Cholesky factor RESIDENT in shared memory (sm_100a only).
//
//   lq[k, n] = cst[k] - 1/2 | Linv_k (x_n - mu_k) |^2            (models/full_cov_gmm.py:56-62)
//
// Why a second tensor-core kernel next to tc_logdens.cu (3xTF32): that kernel streams both operands from L2
// for every (component, 128-sample tile) work item -- 423 KB per item -- and ncu shows it limited by the
// L2 -> SM path and by the TF32 rate at the same time.  fp16 operands are half as wide and run at twice the
// TF32 rate, which (a) halves the tensor time and (b) lets the lower-triangular k-blocks of BOTH split
// halves of Linv_k (160 KB at D = 256) stay in shared memory while a CTA streams sample tiles past them, so
// the only per-item L2 traffic left is the 128 KB sample tile.
//
// Precision.  x = hi + lo with hi = rn_f16(x s), lo = rn_f16(x s - hi): 22 significand bits, the same as the
// TF32 split, PROVIDED nothing overflows or underflows fp16's 5-bit exponent.  Both operands are therefore
// scaled by exact powers of two: the A tile (x_n - mu_k, 128 rows) by s_tk chosen from the bound
// max_n |x_n|_inf + |mu_k|_inf so that |a| < 2^14, the factor Linv_k by t_k chosen from max |Linv_k| likewise; the epilogue multiplies the
// sum of squares by (s t)^-2.  An element 2^-17 of its row's bound still keeps 22 bits; smaller ones lose
// bits gradually (absolute error 2^-39 of the bound), which is below fp32's own rounding of the dot product.
// The products hi*hi, lo*hi, hi*lo are exact in the fp32 accumulator (11 x 11 bits).
//
// Triangular structure.  K-step jb (16 columns j of Linv) only feeds output columns i >= 16 jb, so the MMA
// of that step runs with N = Dp - 16 jb: 53 % of the dense MMA work at D = 256.
//
// Warp roles (768 threads, one persistent CTA per SM, each CTA owns a CONTIGUOUS range of the k-major work list):
//   warp 0      TMA: loads the hi / lo k-blocks of Linv_k (padded fp16 [K, Dp, Dp]) when the component changes
//   warp 1      MMA issuer (one thread): 3 tcgen05.mma per 16-column step
//   warp 2      TMEM allocator (2 x 256 fp32 columns, double buffered accumulators)
//   warps 4-7   epilogue: tcgen05.ld, row sums of squares, un-scaling, lq store
//   warps 8-23  A producers, two groups of 8 warps on alternate 32-column stages: x - mu_k in fp32 (cancellation
//               happens BEFORE the split), scale, split into fp16 hi / lo (FFMA2 / F2FP / FHFMA), write the K-major
//               SWIZZLE_64B operand; each thread keeps the global loads of its next two stages in flight
// Register budget by setmaxnreg: control warps 48, epilogue 80, producers 88 (the pool is the launch allocation, 80 x 768).
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

namespace gvi {
namespace h16 {
using namespace tcx;

constexpr int TILE_M = 128;
constexpr int KB = 64;                    // B: fp16 elements per 128-byte swizzle row = columns per resident k-block
constexpr int KA = 32;                    // A: columns per pipeline stage (64-byte rows, SWIZZLE_64B)
constexpr int STAGES = 4;
constexpr int THREADS = 768;              // 4 control + 4 epilogue + 2 x 8 producer warps
constexpr int A_BYTES = TILE_M * 64;      // 8 KB per hi / lo
constexpr int STAGE_BYTES = 2 * A_BYTES;  // 16 KB
constexpr int ACC_COLS = 256;
constexpr int TMEM_COLS = 512;

__host__ __device__ inline int padded_dim(int D) { return (D + KB - 1) / KB * KB; }
// rows of the packed lower-triangular block storage: block kb keeps rows 64 kb .. Dp-1
__host__ __device__ inline int block_row_offset(int Dp, int kb) { return kb * Dp - KB * (kb * (kb - 1) / 2); }
__host__ __device__ inline int packed_rows(int Dp) { return block_row_offset(Dp, Dp / KB); }

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
// kind::f16 with fp16 operands (a_format = b_format = 0), fp32 accumulate, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

// K-major SWIZZLE_64B shared-memory matrix descriptor: 8-row x 64-byte atoms, 512 B between atoms (SBO)
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

// Power of two s with b * s < 2^14 for the non-negative bound b (exponent clamped to +-40).
__host__ __device__ __forceinline__ float pow2_scale(float b) {
#ifdef __CUDA_ARCH__
  const int e = (__float_as_int(b) >> 23) & 0xff;
#else
  union { float f; int i; } u; u.f = b;
  const int e = (u.i >> 23) & 0xff;
#endif
  int se = 127 + 13 - (e - 127);
  se = se < 87 ? 87 : (se > 167 ? 167 : se);
#ifdef __CUDA_ARCH__
  return __int_as_float(se << 23);
#else
  union { float f; int i; } v; v.i = se << 23;
  return v.f;
#endif
}

__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
// fp32x2 FMA (one issue slot for two lanes of math) and the mixed-precision v - float(h) (FHFMA): sm_100 only
__device__ __forceinline__ float2 ffma2(float2 a, float b, float2 c) {
  uint64_t ra, rb, rc, rd;
  float2 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b), "f"(b));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 ffma2v(float2 a, float2 b, float2 c) {
  uint64_t ra, rb, rc, rd;
  float2 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float sub_f32_f16(float v, unsigned short h) {
  float d;
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(h), "h"((unsigned short)0xBC00), "f"(v));
  return d;
}
// Wait whose fast path is a single try_wait; the bounded spin (trap instead of hang) is kept out of line.
__device__ __forceinline__ void mbar_wait_slow(uint32_t addr, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!ok && clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void mbar_wait_fast(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  if (!ok) mbar_wait_slow(addr, parity);
}
// mbarrier wait that lets the hardware suspend the thread (up to `ns`) instead of spinning through issue slots
__device__ __forceinline__ void mbar_wait_sleepy(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(20000u)
        : "memory");
    if (ok) break;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
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
__device__ __forceinline__ void mbar_arrive_addr(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

struct Barriers {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint64_t b_full;
  uint64_t b_free;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(THREADS, 1)
logdens_h16_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                   const float* __restrict__ X, const float* __restrict__ tileinf, int N, int D, int Dp,
                   const float* __restrict__ means, const float* __restrict__ minf,
                   const float* __restrict__ tmax, const float* __restrict__ cst, int K, float* __restrict__ lq) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int brows = packed_rows(Dp);
  uint8_t* b_hi = smem;
  uint8_t* b_lo = smem + (size_t)brows * 128;
  uint8_t* a_base = smem + (size_t)brows * 256;
  Barriers* bars = reinterpret_cast<Barriers*>(a_base + STAGES * STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = ceil_div(N, TILE_M);
  const long long total = (long long)T * K;      // < 2^31 (checked by the host)
  const int nkb = Dp / KB;
  const int nst = Dp / KA;               // pipeline stages per work item
  const int w_begin = (int)(total * blockIdx.x / gridDim.x);
  const int w_end = (int)(total * (blockIdx.x + 1) / gridDim.x);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bars->full[s], 8);          // one elected arrive per producer warp
      mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->acc_full[b], 1);
      mbar_init(&bars->acc_empty[b], 4);     // one elected arrive per epilogue warp
    }
    mbar_init(&bars->b_full, 1);
    mbar_init(&bars->b_free, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  const int wg = warp >> 2;
  if (wg == 0) {
    reg_dec<56>();
    if (warp == 0) {
      // ---------------- TMA: Linv_k (hi, lo) becomes resident whenever the component changes ----------------
      if (lane == 0 && w_end > w_begin) {
        const int k_first = w_begin / T, k_last = (w_end - 1) / T;
        int nload = 0;
        for (int k = k_first; k <= k_last; ++k, ++nload) {
          if (nload > 0) mbar_wait_sleepy(&bars->b_free, (uint32_t)((nload - 1) & 1));
          mbar_arrive_expect_tx(&bars->b_full, 2u * (uint32_t)brows * 128u);
          for (int kb = 0; kb < nkb; ++kb) {
            const int row0 = block_row_offset(Dp, kb);
            for (int r = 0; r < Dp - kb * KB; r += 64) {
              tma_load_2d(b_hi + (size_t)(row0 + r) * 128, &map_hi, &bars->b_full, kb * KB, k * Dp + kb * KB + r);
              tma_load_2d(b_lo + (size_t)(row0 + r) * 128, &map_lo, &bars->b_full, kb * KB, k * Dp + kb * KB + r);
            }
          }
        }
      }
    } else if (warp == 1) {
      // ---------------- MMA issuer ----------------
      // The whole warp runs the loop with warp-uniform values (descriptors live in uniform registers); only the
      // tcgen05 instructions themselves are issued by one elected lane.  A lane-0 branch around the loop would
      // make ptxas move every descriptor through R2UR inside an ELECT loop: ~30 instructions per MMA.
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      int cur_k = -1, nload = 0;
      const uint64_t a_desc0 = make_desc_sw64(smem_u32(a_base));
      const uint64_t bhi_desc0 = make_desc(smem_u32(b_hi)), blo_desc0 = make_desc(smem_u32(b_lo));
      for (int w = w_begin; w < w_end; ++w, ++it) {
        const int k = w / T;
        if (k != cur_k) {
          mbar_wait_sleepy(&bars->b_full, (uint32_t)(nload & 1));
          ++nload;
          cur_k = k;
        }
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait_sleepy(&bars->acc_empty[buf], (use & 1) ^ 1);
        tc_fence_after();
        for (int si = 0; si < nst; ++si) {
          mbar_wait_fast(smem_u32(&bars->full[s]), ph);
          tc_fence_after();
          const int kb = si >> 1;
          // descriptor start-address fields are in 16-byte units; the sums stay below 2^14 (smem < 256 KB)
          const uint32_t boff16 = (uint32_t)block_row_offset(Dp, kb) * 8u;
          const uint64_t a_st = a_desc0 + (uint64_t)((uint32_t)(s * STAGE_BYTES) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < KA / 16; ++ks) {
              const int j0 = si * KA + ks * 16;
              const int kq = (si & 1) * 2 + ks;          // 16-column step inside B's 64-column block
              const uint32_t idesc = make_idesc_f16(Dp - j0);
              const uint32_t d_tmem = tmem_base + (uint32_t)(buf * ACC_COLS + j0);
              const uint64_t a_hi = a_st + (uint64_t)(ks * 2);
              const uint64_t a_lo = a_hi + (uint64_t)(A_BYTES >> 4);
              const uint64_t boffk = (uint64_t)(boff16 + (uint32_t)(kq * 16) * 8u + (uint32_t)(kq * 2));
              const uint64_t bd_hi = bhi_desc0 + boffk;
              const uint64_t bd_lo = blo_desc0 + boffk;
              umma_f16(d_tmem, a_hi, bd_hi, idesc, j0 != 0 ? 1u : 0u);
              umma_f16(d_tmem, a_lo, bd_hi, idesc, 1u);
              umma_f16(d_tmem, a_hi, bd_lo, idesc, 1u);
            }
            umma_commit(&bars->empty[s]);
            if (si == nst - 1) {
              umma_commit(&bars->acc_full[buf]);
              // last item of this component in my range: the resident factor may be overwritten once these
              // MMAs retire
              if (w + 1 < w_end && (w + 1) / T != k) umma_commit(&bars->b_free);
            }
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (wg == 1) {
    reg_dec<72>();
    // ---------------- epilogue: row sums of squares ----------------
    const int q = warp - 4;
    int it = 0;
    const int ncol32 = Dp / 32;
    for (int w = w_begin; w < w_end; ++w, ++it) {
      const int k = w / T, t = w - k * T;
      const int n = t * TILE_M + 32 * q + lane;
      // operands of the un-scaling factor 1 / (s t): the same inputs as the producers use
      const float xi = __ldg(tileinf + t), mi = __ldg(minf + k), tm = __ldg(tmax + k), c = __ldg(cst + k);
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait_sleepy(&bars->acc_full[buf], use & 1);
      tc_fence_after();
      float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
      for (int cb = 0; cb < ncol32; ++cb) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * ACC_COLS + cb * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float2 a = make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1]));
          const float2 b = make_float2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
          s01 = ffma2v(a, a, s01);
          s23 = ffma2v(b, b, s23);
        }
      }
      const float s0 = s01.x + s23.x, s1 = s01.y + s23.y;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty[buf]);
      const float f = 1.0f / (pow2_scale(xi + mi) * pow2_scale(tm));
      if (n < N) __stcs(lq + (long long)k * N + n, c - 0.5f * (((s0 + s1) * f) * f));     // streaming store: keep X in L2
    }
  } else {
    reg_inc<88>();
    // ---------------- A producers ----------------
    // A warp-wide 128-bit load covers four complete 128-byte row segments (32 fp32 columns): lane l reads the float4
    // c = l % 8 of row rsub + 32 q (rsub = 4 * producer warp + l / 8, q = 0..3) and writes its four fp16 hi / lo
    // values as one 8-byte store each into the 64B-swizzled operand.
    // Two groups of 8 warps convert alternate stages (group g owns the stages si = g, g + 2, ... of every work
    // item), so two stages are being converted at any time and a group's fence / barrier latency is covered by
    // the other group's math.  A warp-wide 128-bit load covers four complete 128-byte row segments (32 fp32
    // columns): lane l reads the float4 c = l % 8 of row rsub + 32 q (rsub = 4 * warp-in-group + l / 8, q = 0..3)
    // and writes its four fp16 hi / lo values as one 8-byte store each into the 64B-swizzled operand.
    const int grp = (warp - 8) >> 3;
    const int pw = (warp - 8) & 7;
    const int c = lane & 7;
    const int rsub = 4 * pw + (lane >> 3);
    // ((row >> 1) & 3) == ((rsub >> 1) & 3) for all my rows
    uint32_t a_smem = smem_u32(a_base) + (uint32_t)(grp * STAGE_BYTES) +
                      (uint32_t)(rsub * 64 + ((((c >> 1) ^ ((rsub >> 1) & 3)) << 4) + ((c & 1) << 3)));
    uint32_t empty_bar = smem_u32(&bars->empty[grp]), full_bar = smem_u32(&bars->full[grp]);
    const float4* __restrict__ X4 = reinterpret_cast<const float4*>(X);
    const float4* __restrict__ M4 = reinterpret_cast<const float4*>(means);
    const int D4 = D >> 2;
    const int hst = nst >> 1;                                  // my stages per work item
    const int nitems = (w_end - w_begin) * hst;
    const int stride4 = 32 * D4;                               // float4 units between my consecutive rows
    // keep the per-thread constants in registers (ptxas otherwise re-derives them from %tid in every stage)
    asm volatile("" : "+r"(a_smem), "+r"(empty_bar), "+r"(full_bar));

    // ---- load stream state ----
    int h_ld = 0;                                              // my stage inside the work item
    int k_ld = 0, t_ld = 0;
    if (w_end > w_begin) {
      k_ld = w_begin / T;
      t_ld = w_begin - k_ld * T;
    }
    // ---- store stream state ----
    int h_st = 0, jj_st = 0;
    float sc_cur = 0.f;
    struct Regs {
      float4 x[4];
      float4 m;
      float xi, mi;
    };
    // Loads are never predicated: columns past D (zero padding of the operand) are clamped to valid addresses and
    // multiplied by a zero scale.  Offsets are 32-bit float4 indices (the host checks N * D < 2^31).
    auto issue = [&](Regs& R) {
      const int col4 = (2 * h_ld + grp) * (KA / 4) + c;
      const int colc4 = min(col4, D4 - 1);
      const int row0 = t_ld * TILE_M + rsub;
      if (row0 + 96 < N) {                                     // all four of my rows exist
        const int i0 = row0 * D4 + colc4;
#pragma unroll
        for (int q = 0; q < 4; ++q) R.x[q] = __ldg(X4 + (i0 + q * stride4));
      } else {                                                 // tail of the sample range: clamp (never stored)
#pragma unroll
        for (int q = 0; q < 4; ++q) R.x[q] = __ldg(X4 + (min(row0 + 32 * q, N - 1) * D4 + colc4));
      }
      R.m = __ldg(M4 + (k_ld * D4 + colc4));
      if (h_ld == 0) {
        R.xi = __ldg(tileinf + t_ld);
        R.mi = __ldg(minf + k_ld);
      }
      if (++h_ld == hst) {
        h_ld = 0;
        if (++t_ld == T) { t_ld = 0; ++k_ld; }
      }
    };
    auto emit = [&](const Regs& R) {
      if (h_st == 0) sc_cur = pow2_scale(R.xi + R.mi);
      const bool colok = (2 * h_st + grp) * (KA / 4) + c < D4;
      const float sc = colok ? sc_cur : 0.f;
      if (++h_st == hst) h_st = 0;
      // (x - m) sc == fma(x, sc, -(m sc)) bit for bit: scaling by a power of two commutes with rounding
      const float2 nm01 = make_float2(-R.m.x * sc, -R.m.y * sc), nm23 = make_float2(-R.m.z * sc, -R.m.w * sc);
      const uint32_t sl = (uint32_t)(jj_st & 1);               // my slot: stage ring entry grp + 2 sl
      mbar_wait_fast(empty_bar + sl * 16, (uint32_t)(((jj_st >> 1) & 1) ^ 1));
      ++jj_st;
      const uint32_t st = a_smem + sl * (2 * STAGE_BYTES);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 v01 = ffma2(make_float2(R.x[q].x, R.x[q].y), sc, nm01);
        const float2 v23 = ffma2(make_float2(R.x[q].z, R.x[q].w), sc, nm23);
        const __half2 h01 = __floats2half2_rn(v01.x, v01.y), h23 = __floats2half2_rn(v23.x, v23.y);
        const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01), u23 = *reinterpret_cast<const uint32_t*>(&h23);
        const __half2 l01 = __floats2half2_rn(sub_f32_f16(v01.x, (unsigned short)(u01 & 0xffffu)),
                                              sub_f32_f16(v01.y, (unsigned short)(u01 >> 16)));
        const __half2 l23 = __floats2half2_rn(sub_f32_f16(v23.x, (unsigned short)(u23 & 0xffffu)),
                                              sub_f32_f16(v23.y, (unsigned short)(u23 >> 16)));
        sts64(st + q * 2048, u01, u23);
        sts64(st + A_BYTES + q * 2048, *reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_addr(full_bar + sl * 16);
    };
    // register ring of two of my stages: the global loads run one stage of this group (two stages of the
    // pipeline, 32 KB per SM counting both groups... plus the one being converted) ahead of the conversion
    Regs R0, R1;
    if (nitems > 0) issue(R0);
    for (int j = 0; j < nitems; j += 2) {
      if (j + 1 < nitems) issue(R1);
      emit(R0);
      if (j + 1 < nitems) { if (j + 2 < nitems) issue(R0); emit(R1); }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// =====================================================================================================
// Variant with the A operand in TENSOR MEMORY (default for Dp >= 128).
//
// ncu on the kernel above: the tensor pipe is active 54 % of the time and the shared-memory port is the limiter -- per
// work item the MMAs fetch 192 KB of A and 209 KB of B from shared memory and the producers store another 128 KB, and
// an SS MMA with N < 128 is bound by its operand fetch (4 KB of A + 32 N bytes of B at 128 B/clk; measured with
// tools/probe_tmem.cu: 48 cycles at N = 64 where the math needs 32).  Here the producers write the split (x - mu) tile
// straight into TMEM (tcgen05.st.16x256b: four lanes hold one row's 16 fp16 values, which is exactly one float4 of
// the fp32 source per lane) and the MMAs take A from there, so shared memory only serves the resident factor.
//
// TMEM (512 columns): [0, 256) ring of 8 A stages (32 K-columns each; per 16-column K-step 8 columns of packed hi, 8 of lo);
// [256, 384) and [384, 512) two accumulators of Dp/2 <= 128 columns.  A work item is issued as two sub-items so that
// the accumulators stay double buffered: first the output columns [0, Dp/2) (K-steps j < Dp/2 only, triangular), then
// [Dp/2, Dp) (all K-steps); the epilogue adds the two partial row sums.  An A stage is released after its second use.
//
// Producers: 16 warps = 4 groups of 4 warps (one per TMEM lane quadrant); group g converts the stages g, g + 4, ... of
// the global stage sequence.  Thread t of a warp owns rows t/4 + 8 a (a = 0..3) of its quadrant and the float4 columns
// t%4 and t%4 + 4 of the stage, i.e. a warp-wide load covers 8 rows x 64 contiguous bytes.
namespace h16t {
using namespace h16;

constexpr int RING = 8;
constexpr int ACC0 = 256, ACC1 = 384;

struct BarriersT {
  uint64_t full[RING];
  uint64_t empty[RING];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint64_t b_full;
  uint64_t b_free;
  uint32_t tmem_base;
};

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16 lanes x 16 columns: registers 4 i .. 4 i + 3 are the i-th 16x256b block (columns 8 i .. 8 i + 7); inside a block
// thread t holds (row t/4, columns 2 (t%4), +1) in registers 0, 1 and (row t/4 + 8, same columns) in registers 2, 3
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// The tcgen05.mma of one sub-item (SUB = 0: output columns [0, Dp/2), SUB = 1: [Dp/2, Dp)) of one work item, fully
// unrolled for NST = Dp / 32 stages so that every operand offset, N and accumulator column is an immediate.  The
// issuing warp's instruction stream is the critical path of the kernel: the MMA queue is shallow (the issuing thread
// blocks in UTCHMMA almost in lock step with the execution: tools/probe_tmem.cu T4), so whatever scalar work sits
// between two stages is exposed whenever the queued MMAs are short.  Hence (a) immediates instead of ~150 dependent
// scalar instructions per stage, and (b) one issuing warp per sub-item: the two accumulators are independent, which
// also hides the ~40-cycle latency between dependent small MMAs on one accumulator.
template <int NST, int SUB>
__device__ __forceinline__ void mma_sub_item(uint32_t tmem_base, BarriersT* bars, uint32_t g0, uint32_t it,
                                             uint64_t bhi_desc0, uint64_t blo_desc0, bool last_of_k) {
  constexpr int Dp = 32 * NST, H = 16 * NST;
  constexpr int ns = SUB == 0 ? NST / 2 : NST;
  // keep the operand addresses of an item out of registers: they are re-derived (one add with an immediate each)
  // instead of being hoisted out of the item loop
  asm volatile("" : "+l"(bhi_desc0), "+l"(blo_desc0), "+r"(tmem_base));
  mbar_wait_sleepy(&bars->acc_empty[SUB], (it & 1u) ^ 1u);
  tc_fence_after();
  const uint32_t acc = tmem_base + (uint32_t)(SUB == 0 ? ACC0 : ACC1);
#pragma unroll
  for (int si = 0; si < ns; ++si) {
    const uint32_t g = g0 + (uint32_t)si;
    const uint32_t slot = NST == RING ? (uint32_t)si : (g & (RING - 1));
    const uint32_t par = NST == RING ? (it & 1u) : ((g >> 3) & 1u);
    mbar_wait_fast(smem_u32(&bars->full[slot]), par);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t a0 = tmem_base + slot * 32u;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int c = 2 * si + ks;                 // K-step: columns 16 c .. 16 c + 15 of Linv
        const int kb = c >> 2, kq = c & 3;
        const int n = SUB == 0 ? H - 16 * c : (c < NST ? H : Dp - 16 * c);
        const int i0 = SUB == 0 ? 16 * c : (c < NST ? H : 16 * c);
        const int dcol = SUB == 0 ? 16 * c : (c < NST ? 0 : 16 * c - H);
        const uint32_t idesc = make_idesc_f16(n);
        const uint32_t d_tmem = acc + (uint32_t)dcol;
        const uint32_t a_hi = a0 + (uint32_t)(ks * 16);
        const uint32_t a_lo = a_hi + 8u;
        // descriptor start-address fields are in 16-byte units (smem < 256 KB)
        const uint64_t boffk = (uint64_t)((uint32_t)(block_row_offset(Dp, kb) + i0 - KB * kb) * 8u + (uint32_t)(kq * 2));
        const uint64_t bd_hi = bhi_desc0 + boffk, bd_lo = blo_desc0 + boffk;
        umma_f16_ts(d_tmem, a_hi, bd_hi, idesc, c != 0 ? 1u : 0u);
        umma_f16_ts(d_tmem, a_lo, bd_hi, idesc, 1u);
        umma_f16_ts(d_tmem, a_hi, bd_lo, idesc, 1u);
      }
      // a stage is free once both sub-items have read it; stages past the first half are only read by SUB 1
      umma_commit(&bars->empty[slot]);
      if (SUB == 1 && si >= NST / 2) umma_commit(&bars->empty[slot]);
      if (si == ns - 1) {
        umma_commit(&bars->acc_full[SUB]);
        if (last_of_k) umma_commit(&bars->b_free);
      }
    }
    __syncwarp();
  }
}

template <int SUB>
__device__ __forceinline__ void mma_issuer(uint32_t tmem_base, BarriersT* bars, int nst, int T, int w_begin, int w_end,
                                           uint64_t bhi_desc0, uint64_t blo_desc0) {
  int nload = 0;
  uint32_t it = 0, g0 = 0;                       // g0: global stage counter of the item's first stage
  int t = 0;
  bool new_k = true;
  if (w_end > w_begin) t = w_begin - (w_begin / T) * T;
  for (int w = w_begin; w < w_end; ++w, ++it, g0 += (uint32_t)nst) {
    if (new_k) {
      mbar_wait_sleepy(&bars->b_full, (uint32_t)(nload & 1));
      ++nload;
    }
    new_k = (t == T - 1);                        // the next item starts a new component
    t = new_k ? 0 : t + 1;
    const bool last_of_k = new_k && (w + 1 < w_end);
    if (nst == 8) mma_sub_item<8, SUB>(tmem_base, bars, g0, it, bhi_desc0, blo_desc0, last_of_k);
    else if (nst == 6) mma_sub_item<6, SUB>(tmem_base, bars, g0, it, bhi_desc0, blo_desc0, last_of_k);
    else mma_sub_item<4, SUB>(tmem_base, bars, g0, it, bhi_desc0, blo_desc0, last_of_k);
  }
}

__global__ void __launch_bounds__(THREADS, 1)
logdens_h16t_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                    const float* __restrict__ X, const float* __restrict__ tileinf, int N, int D, int Dp,
                    const float* __restrict__ means, const float* __restrict__ minf,
                    const float* __restrict__ tmax, const float* __restrict__ cst, int K, float* __restrict__ lq) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int brows = packed_rows(Dp);
  uint8_t* b_hi = smem;
  uint8_t* b_lo = smem + (size_t)brows * 128;
  BarriersT* bars = reinterpret_cast<BarriersT*>(smem + (size_t)brows * 256);
  // mean rows of the components in this CTA's range (zero padded to Dp): the producers need one float4 of the mean per
  // converted float4 and an L2 round trip for it would sit in front of every conversion
  float* mean_sm = reinterpret_cast<float*>(smem + (size_t)brows * 256 + 256);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = ceil_div(N, TILE_M);
  const long long total = (long long)T * K;      // < 2^31 (checked by the host)
  const int nkb = Dp / KB;
  const int nst = Dp / KA;               // A stages per work item: 4, 6 or 8
  const int H = Dp / 2;                  // output columns per sub-item
  const int w_begin = (int)(total * blockIdx.x / gridDim.x);
  const int w_end = (int)(total * (blockIdx.x + 1) / gridDim.x);

  if (threadIdx.x == 0) {
    for (int s = 0; s < RING; ++s) {
      mbar_init(&bars->full[s], 4);          // one elected arrive per producer warp of the owning group
      mbar_init(&bars->empty[s], 2);         // one commit per sub-item that reads the stage (see mma_sub_item)
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->acc_full[b], 1);
      mbar_init(&bars->acc_empty[b], 4);     // one elected arrive per epilogue warp
    }
    mbar_init(&bars->b_full, 1);
    mbar_init(&bars->b_free, 2);           // both issuers are done with the resident factor
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  const int k_first_cta = w_end > w_begin ? w_begin / T : 0;
  if (w_end > w_begin) {
    const int ncomp = (w_end - 1) / T - k_first_cta + 1;
    for (int i = threadIdx.x; i < ncomp * Dp; i += THREADS) {
      const int kk = i / Dp, j = i - kk * Dp;
      mean_sm[i] = j < D ? __ldg(means + (size_t)(k_first_cta + kk) * D + j) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  const int wg = warp >> 2;
  if (wg == 0) {
    reg_dec<56>();
    if (warp == 0) {
      // ---------------- TMA: Linv_k (hi, lo) becomes resident whenever the component changes ----------------
      if (lane == 0 && w_end > w_begin) {
        const int k_first = w_begin / T, k_last = (w_end - 1) / T;
        int nload = 0;
        for (int k = k_first; k <= k_last; ++k, ++nload) {
          if (nload > 0) mbar_wait_sleepy(&bars->b_free, (uint32_t)((nload - 1) & 1));
          mbar_arrive_expect_tx(&bars->b_full, 2u * (uint32_t)brows * 128u);
          for (int kb = 0; kb < nkb; ++kb) {
            const int row0 = block_row_offset(Dp, kb);
            for (int r = 0; r < Dp - kb * KB; r += 64) {
              tma_load_2d(b_hi + (size_t)(row0 + r) * 128, &map_hi, &bars->b_full, kb * KB, k * Dp + kb * KB + r);
              tma_load_2d(b_lo + (size_t)(row0 + r) * 128, &map_lo, &bars->b_full, kb * KB, k * Dp + kb * KB + r);
            }
          }
        }
      }
    } else if (warp == 1 || warp == 3) {
      // ---------------- MMA issuers: warp 1 the sub-items 0, warp 3 the sub-items 1 of every work item ----------------
      const uint64_t bhi_desc0 = make_desc(smem_u32(b_hi)), blo_desc0 = make_desc(smem_u32(b_lo));
      if (warp == 1) mma_issuer<0>(tmem_base, bars, nst, T, w_begin, w_end, bhi_desc0, blo_desc0);
      else mma_issuer<1>(tmem_base, bars, nst, T, w_begin, w_end, bhi_desc0, blo_desc0);
    }
  } else if (wg == 1) {
    reg_dec<72>();
    // ---------------- epilogue: row sums of squares of the two sub-items ----------------
    const int q = warp - 4;
    int it = 0;
    const int ncol32 = H / 32;
    for (int w = w_begin; w < w_end; ++w, ++it) {
      const int k = w / T, t = w - k * T;
      const int n = t * TILE_M + 32 * q + lane;
      const float xi = __ldg(tileinf + t), mi = __ldg(minf + k), tm = __ldg(tmax + k), c = __ldg(cst + k);
      float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
        mbar_wait_sleepy(&bars->acc_full[sub], (uint32_t)it & 1u);
        tc_fence_after();
        for (int cb = 0; cb < ncol32; ++cb) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)((sub == 0 ? ACC0 : ACC1) + cb * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float2 a = make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1]));
            const float2 b = make_float2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
            s01 = ffma2v(a, a, s01);
            s23 = ffma2v(b, b, s23);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->acc_empty[sub]);
      }
      const float s0 = s01.x + s23.x, s1 = s01.y + s23.y;
      const float f = 1.0f / (pow2_scale(xi + mi) * pow2_scale(tm));
      if (n < N) __stcs(lq + (long long)k * N + n, c - 0.5f * (((s0 + s1) * f) * f));     // streaming store: keep X in L2
    }
  } else {
    reg_inc<88>();
    // ---------------- A producers ----------------
    const int grp = (warp - 8) >> 2;                           // stage sequence g = grp, grp + 4, ...
    const int q = warp & 3;                                    // TMEM lane quadrant this warp may access
    const int c4 = lane & 3;
    const int rsub = 32 * q + (lane >> 2);                     // my rows: rsub + 8 a, a = 0..3
    const float4* __restrict__ X4 = reinterpret_cast<const float4*>(X);
    const int D4 = D >> 2;
    const long long nstage_total = (long long)(w_end - w_begin) * nst;
    // my half-stages: two per stage g = grp + 4 j
    const int nunits = nstage_total > grp ? (int)((nstage_total - grp + 3) / 4) * 2 : 0;
    const uint32_t taddr0 = tmem_base + ((uint32_t)(32 * q) << 16);

    struct Half {
      float4 x[4];                                             // [2 * (row a & 1) + b]: rows rsub + 16 h + 8 (a & 1)
      float xi, mi;
    };
    // ---- load stream state: (stage s of item (k, t)), half h ----
    int s_ld = grp, k_ld = 0, t_ld = 0, h_ld = 0;
    if (w_end > w_begin) {
      k_ld = w_begin / T;
      t_ld = w_begin - k_ld * T;
    }
    while (s_ld >= nst) {                                      // nst = 4 with grp < 4 never enters; kept for safety
      s_ld -= nst;
      if (++t_ld == T) { t_ld = 0; ++k_ld; }
    }
    // ---- store stream state ----
    int s_st = s_ld, t_st = t_ld, h_st = 0;
    uint32_t mrow_st = smem_u32(mean_sm);                      // cached mean row of the store stream's component
    uint32_t g_st = (uint32_t)grp;
    float sc_cur = 0.f;

    auto issue = [&](Half& R) {
      const int col4 = s_ld * 8 + c4;
      const int ca = min(col4, D4 - 1), cb = min(col4 + 4, D4 - 1);
      const int row0 = t_ld * TILE_M + rsub + 16 * h_ld;
      const int r0 = min(row0, N - 1), r1 = min(row0 + 8, N - 1);
      R.x[0] = __ldg(X4 + (r0 * D4 + ca));
      R.x[1] = __ldg(X4 + (r0 * D4 + cb));
      R.x[2] = __ldg(X4 + (r1 * D4 + ca));
      R.x[3] = __ldg(X4 + (r1 * D4 + cb));
      if (h_ld == 0) {
        R.xi = __ldg(tileinf + t_ld);
        R.mi = __ldg(minf + k_ld);
      }
      h_ld ^= 1;
      if (h_ld == 0) {
        s_ld += 4;
        if (s_ld >= nst) {
          s_ld -= nst;
          if (++t_ld == T) { t_ld = 0; ++k_ld; }
        }
      }
    };
    auto emit = [&](const Half& R) {
      const uint32_t slot = g_st & (RING - 1);
      if (h_st == 0) {
        sc_cur = pow2_scale(R.xi + R.mi);
        mbar_wait_sleepy(&bars->empty[slot], ((g_st >> 3) & 1u) ^ 1u);
        tc_fence_after();
      }
      const int col4 = s_st * 8 + c4;
#pragma unroll
      for (int b = 0; b < 2; ++b) {                            // K-step ks = b of the stage: float4 column col4 + 4 b
        const float sc = col4 + 4 * b < D4 ? sc_cur : 0.f;
        const float4 m = lds128(mrow_st + (uint32_t)(16 * (col4 + 4 * b)));   // zero padded to Dp columns
        // (x - m) sc == fma(x, sc, -(m sc)) bit for bit: scaling by a power of two commutes with rounding
        const float2 nm01 = make_float2(-m.x * sc, -m.y * sc), nm23 = make_float2(-m.z * sc, -m.w * sc);
        uint32_t r[8];                                         // blocks (hi, lo) x registers (row, row + 8) x 2
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const float4 x = R.x[2 * a + b];
          const float2 v01 = ffma2(make_float2(x.x, x.y), sc, nm01);
          const float2 v23 = ffma2(make_float2(x.z, x.w), sc, nm23);
          const __half2 h01 = __floats2half2_rn(v01.x, v01.y), h23 = __floats2half2_rn(v23.x, v23.y);
          const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01), u23 = *reinterpret_cast<const uint32_t*>(&h23);
          const __half2 l01 = __floats2half2_rn(sub_f32_f16(v01.x, (unsigned short)(u01 & 0xffffu)),
                                                sub_f32_f16(v01.y, (unsigned short)(u01 >> 16)));
          const __half2 l23 = __floats2half2_rn(sub_f32_f16(v23.x, (unsigned short)(u23 & 0xffffu)),
                                                sub_f32_f16(v23.y, (unsigned short)(u23 >> 16)));
          r[2 * a] = u01;
          r[2 * a + 1] = u23;
          r[4 + 2 * a] = *reinterpret_cast<const uint32_t*>(&l01);
          r[4 + 2 * a + 1] = *reinterpret_cast<const uint32_t*>(&l23);
        }
        tmem_st_16x256b_x2(taddr0 + ((uint32_t)(16 * h_st) << 16) + slot * 32u + (uint32_t)(16 * b), r);
      }
      h_st ^= 1;
      if (h_st == 0) {
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->full[slot]);
        g_st += 4;
        s_st += 4;
        if (s_st >= nst) {
          s_st -= nst;
          if (++t_st == T) { t_st = 0; mrow_st += (uint32_t)(Dp * 4); }
        }
      }
    };
    Half R0, R1;
    if (nunits > 0) issue(R0);
    for (int j = 0; j < nunits; j += 2) {
      if (j + 1 < nunits) issue(R1);
      emit(R0);
      if (j + 1 < nunits) {
        if (j + 2 < nunits) issue(R0);
        emit(R1);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

constexpr int MEANS_CACHED_MAX = 32;      // components whose mean rows are cached in shared memory (32 KB at Dp = 256)
static size_t smem_bytes_t(int Dp) {
  return (size_t)packed_rows(Dp) * 256 + 1024 /*alignment slack*/ + 256 /*barriers*/ + (size_t)MEANS_CACHED_MAX * Dp * 4;
}

}  // namespace h16t

// =====================================================================================================
// Mixture gradient on the same structure:  grad[n, :] = - sum_k r_kn P_k (x_n - mu_k)      (models/gmm.py:274-300)
//
// Work item = (128-sample tile t, component k) with some responsibility r_kn > e^-60 in the tile (bit mask from
// resp_mask_kernel); every role of the CTA walks the mask of its tiles in the same order.  Per item the producers
// write A = r_kn (x_n - mu_k) s (fp16 hi / lo, tensor memory, exactly like the log-density kernel plus the row
// factor), and the item is issued in two phases h = 0, 1 for the output columns [h Dp/2, (h + 1) Dp/2): the TMA warp
// loads that half of the fp16-split precision matrix P_k (rows = output columns; 128 KB at D = 256) into shared
// memory, the MMAs run over all K-steps into the half accumulator h, the epilogue adds -Z / (s t_k) into grad
// (vector reductions performed in L2: all items of a tile belong to the same CTA and are retired in order by the same
// threads, so the sum is deterministic).  The A stages stay in TMEM for both phases.  P_k is used once per item, so the
// load of the next half is exposed (single buffer); at C5 that is 1.5 items per tile and irrelevant, and in the dense
// worst case the kernel is still several times faster than the SIMT tile engine it replaces.
namespace mg {
using namespace h16t;

struct BarriersM {
  uint64_t full[RING];
  uint64_t empty[RING];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint64_t b_full;
  uint64_t b_free;
  uint32_t tmem_base;
};

// Fire-and-forget vector add into global memory (performed in L2).  A read-modify-write in the epilogue costs an L2
// round trip per 32-column chunk (~10 us per item, 9 % tensor activity); the reduction does not wait.  Every address is
// only ever updated by one thread, in program order, so the sum stays deterministic.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// the (tile, component) items of tiles [tile, tile_end) in mask order
struct Walker {
  const uint32_t* mask;
  int words, tile, tile_end, wd;
  uint32_t bits;
  __device__ __forceinline__ void init(const uint32_t* m, int w, int t0, int t1) {
    mask = m; words = w; tile = t0; tile_end = t1; wd = 0;
    bits = t0 < t1 ? __ldg(m + (long long)t0 * w) : 0u;
  }
  __device__ __forceinline__ bool next(int& t, int& k) {
    while (true) {
      if (bits) {
        k = wd * 32 + (__ffs(bits) - 1);
        bits &= bits - 1;
        t = tile;
        return true;
      }
      if (++wd == words) { wd = 0; ++tile; }
      if (tile >= tile_end) return false;
      bits = __ldg(mask + (long long)tile * words + wd);
    }
  }
};

template <int NST>
__global__ void __launch_bounds__(THREADS, 1)
mixgrad_h16_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                   const float* __restrict__ X, const float* __restrict__ tileinf, int N, int D,
                   const float* __restrict__ means, const float* __restrict__ minf, const float* __restrict__ tmaxp,
                   const float* __restrict__ lq, const float* __restrict__ logw, const float* __restrict__ logq, int K,
                   const uint32_t* __restrict__ mask, int words, float* __restrict__ grad) {
  constexpr int Dp = 32 * NST, Hh = 16 * NST, NKB = Dp / KB;
  constexpr uint32_t HALF_BYTES = (uint32_t)NKB * Hh * 128;        // one split half of the resident P rows
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_hi = smem;
  uint8_t* b_lo = smem + HALF_BYTES;
  BarriersM* bars = reinterpret_cast<BarriersM*>(smem + 2 * HALF_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = ceil_div(N, TILE_M);
  const int t_begin = (int)((long long)T * blockIdx.x / gridDim.x);
  const int t_end = (int)((long long)T * (blockIdx.x + 1) / gridDim.x);

  if (threadIdx.x == 0) {
    for (int s = 0; s < RING; ++s) {
      mbar_init(&bars->full[s], 4);
      mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->acc_full[b], 1);
      mbar_init(&bars->acc_empty[b], 4);
    }
    mbar_init(&bars->b_full, 1);
    mbar_init(&bars->b_free, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  const int wg = warp >> 2;
  if (wg == 0) {
    reg_dec<56>();
    if (warp == 0) {
      // ---------------- TMA: the half h of P_k (hi, lo) for every (item, h) ----------------
      if (lane == 0) {
        Walker wk;
        wk.init(mask, words, t_begin, t_end);
        int t, k;
        uint32_t p = 0;                                          // phase counter: (item, h) pairs
        while (wk.next(t, k)) {
          for (int h = 0; h < 2; ++h, ++p) {
            if (p > 0) mbar_wait_sleepy(&bars->b_free, (p - 1) & 1u);
            mbar_arrive_expect_tx(&bars->b_full, 2u * HALF_BYTES);
            for (int kb = 0; kb < NKB; ++kb)
              for (int r = 0; r < Hh; r += 64) {
                tma_load_2d(b_hi + (size_t)(kb * Hh + r) * 128, &map_hi, &bars->b_full, kb * KB, k * Dp + h * Hh + r);
                tma_load_2d(b_lo + (size_t)(kb * Hh + r) * 128, &map_lo, &bars->b_full, kb * KB, k * Dp + h * Hh + r);
              }
          }
        }
      }
    } else if (warp == 1) {
      // ---------------- MMA issuer: both phases of every item ----------------
      Walker wk;
      wk.init(mask, words, t_begin, t_end);
      int t, k;
      uint32_t it = 0, p = 0, g0 = 0;
      uint64_t bhi_desc0 = make_desc(smem_u32(b_hi)), blo_desc0 = make_desc(smem_u32(b_lo));
      while (wk.next(t, k)) {
        uint32_t tb = tmem_base;
        asm volatile("" : "+l"(bhi_desc0), "+l"(blo_desc0), "+r"(tb));      // re-derive operand addresses per item
#pragma unroll
        for (int h = 0; h < 2; ++h, ++p) {
          mbar_wait_sleepy(&bars->b_full, p & 1u);
          mbar_wait_sleepy(&bars->acc_empty[h], (it & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t acc = tb + (uint32_t)(h == 0 ? ACC0 : ACC1);
          const uint32_t idesc = make_idesc_f16(Hh);
#pragma unroll
          for (int si = 0; si < NST; ++si) {
            const uint32_t g = g0 + (uint32_t)si;
            const uint32_t slot = NST == RING ? (uint32_t)si : (g & (RING - 1));
            if (h == 0) {
              mbar_wait_fast(smem_u32(&bars->full[slot]), NST == RING ? (it & 1u) : ((g >> 3) & 1u));
              tc_fence_after();
            }
            if (elect_one()) {
              const uint32_t a0 = tb + slot * 32u;
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const int c = 2 * si + ks;
                const uint64_t boffk = (uint64_t)((uint32_t)((c >> 2) * Hh * 8) + (uint32_t)((c & 3) * 2));
                const uint64_t bd_hi = bhi_desc0 + boffk, bd_lo = blo_desc0 + boffk;
                const uint32_t a_hi = a0 + (uint32_t)(ks * 16), a_lo = a_hi + 8u;
                umma_f16_ts(acc, a_hi, bd_hi, idesc, c != 0 ? 1u : 0u);
                umma_f16_ts(acc, a_lo, bd_hi, idesc, 1u);
                umma_f16_ts(acc, a_hi, bd_lo, idesc, 1u);
              }
              if (h == 1) umma_commit(&bars->empty[slot]);
              if (si == NST - 1) {
                umma_commit(&bars->acc_full[h]);
                umma_commit(&bars->b_free);
              }
            }
            __syncwarp();
          }
        }
        ++it;
        g0 += NST;
      }
    }
  } else if (wg == 1) {
    reg_dec<72>();
    // ---------------- epilogue: grad[n, h Hh + c] -= Z[n, c] / (s t_k) ----------------
    const int q = warp - 4;
    Walker wk;
    wk.init(mask, words, t_begin, t_end);
    int t, k;
    uint32_t it = 0;
    while (wk.next(t, k)) {
      const int n = t * TILE_M + 32 * q + lane;
      const float f = -1.0f / (pow2_scale(__ldg(tileinf + t) + __ldg(minf + k)) * pow2_scale(__ldg(tmaxp + k)));
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        mbar_wait_sleepy(&bars->acc_full[h], it & 1u);
        tc_fence_after();
        for (int cb = 0; cb < Hh / 32; ++cb) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)((h == 0 ? ACC0 : ACC1) + cb * 32), v);
          tmem_ld_wait();
          const int col0 = h * Hh + cb * 32;
          if (n < N) {
            float* gp = grad + (long long)n * D + col0;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              if (col0 + i < D)                                    // D % 4 == 0: the whole float4 is inside
                red_add_v4(gp + i, f * __uint_as_float(v[i]), f * __uint_as_float(v[i + 1]),
                           f * __uint_as_float(v[i + 2]), f * __uint_as_float(v[i + 3]));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->acc_empty[h]);
      }
      ++it;
    }
  } else {
    reg_inc<88>();
    // ---------------- A producers: r_kn (x_n - mu_k) s, fp16 hi / lo, into tensor memory ----------------
    const int grp = (warp - 8) >> 2;
    const int q = warp & 3;
    const int c4 = lane & 3;
    const int rsub = 32 * q + (lane >> 2);
    const float4* __restrict__ X4 = reinterpret_cast<const float4*>(X);
    const float4* __restrict__ M4 = reinterpret_cast<const float4*>(means);
    const int D4 = D >> 2;
    const uint32_t taddr0 = tmem_base + ((uint32_t)(32 * q) << 16);
    struct Half {
      float4 x[4];
      float a0, a1;                                            // log responsibilities of my two rows
      float bound;                                             // tileinf + minf of the item
    };
    // load stream: stage s_ld of item (t_ld, k_ld), half h_ld; store stream likewise (two walkers over the same items)
    Walker wl, ws_;
    wl.init(mask, words, t_begin, t_end);
    ws_.init(mask, words, t_begin, t_end);
    int t_ld = 0, k_ld = 0, t_st = 0, k_st = 0;
    bool ld_ok = wl.next(t_ld, k_ld);
    ws_.next(t_st, k_st);
    int s_ld = grp, h_ld = 0, s_st = grp, h_st = 0;
    while (ld_ok && s_ld >= NST) {                             // NST = 4, grp < 4: never; kept for symmetry
      s_ld -= NST;
      ld_ok = wl.next(t_ld, k_ld);
    }
    uint32_t g_st = (uint32_t)grp;
    float sc_cur = 0.f;
    auto issue = [&](Half& R) -> bool {
      if (!ld_ok) return false;
      const int col4 = s_ld * 8 + c4;
      const int ca = min(col4, D4 - 1), cb = min(col4 + 4, D4 - 1);
      const int row0 = t_ld * TILE_M + rsub + 16 * h_ld;
      const int r0 = min(row0, N - 1), r1 = min(row0 + 8, N - 1);
      R.x[0] = __ldg(X4 + (r0 * D4 + ca));
      R.x[1] = __ldg(X4 + (r0 * D4 + cb));
      R.x[2] = __ldg(X4 + (r1 * D4 + ca));
      R.x[3] = __ldg(X4 + (r1 * D4 + cb));
      const float lw = __ldg(logw + k_ld);
      R.a0 = row0 < N ? __ldg(lq + (long long)k_ld * N + r0) + lw - __ldg(logq + r0) : -INFINITY;
      R.a1 = row0 + 8 < N ? __ldg(lq + (long long)k_ld * N + r1) + lw - __ldg(logq + r1) : -INFINITY;
      R.bound = __ldg(tileinf + t_ld) + __ldg(minf + k_ld);
      h_ld ^= 1;
      if (h_ld == 0) {
        s_ld += 4;
        if (s_ld >= NST) {
          s_ld -= NST;
          ld_ok = wl.next(t_ld, k_ld);
        }
      }
      return true;
    };
    auto emit = [&](const Half& R) {
      const uint32_t slot = g_st & (RING - 1);
      if (h_st == 0) {
        sc_cur = pow2_scale(R.bound);
        mbar_wait_sleepy(&bars->empty[slot], ((g_st >> 3) & 1u) ^ 1u);
        tc_fence_after();
      }
      const float scr0 = R.a0 > -60.f ? expf(R.a0) * sc_cur : 0.f;
      const float scr1 = R.a1 > -60.f ? expf(R.a1) * sc_cur : 0.f;
      const int col4 = s_st * 8 + c4;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const bool colok = col4 + 4 * b < D4;
        const float4 m = __ldg(M4 + (k_st * D4 + min(col4 + 4 * b, D4 - 1)));
        uint32_t r[8];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const float sc = colok ? (a ? scr1 : scr0) : 0.f;
          const float4 x = R.x[2 * a + b];
          const float2 v01 = ffma2(make_float2(x.x, x.y), sc, make_float2(-m.x * sc, -m.y * sc));
          const float2 v23 = ffma2(make_float2(x.z, x.w), sc, make_float2(-m.z * sc, -m.w * sc));
          const __half2 h01 = __floats2half2_rn(v01.x, v01.y), h23 = __floats2half2_rn(v23.x, v23.y);
          const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01), u23 = *reinterpret_cast<const uint32_t*>(&h23);
          const __half2 l01 = __floats2half2_rn(sub_f32_f16(v01.x, (unsigned short)(u01 & 0xffffu)),
                                                sub_f32_f16(v01.y, (unsigned short)(u01 >> 16)));
          const __half2 l23 = __floats2half2_rn(sub_f32_f16(v23.x, (unsigned short)(u23 & 0xffffu)),
                                                sub_f32_f16(v23.y, (unsigned short)(u23 >> 16)));
          r[2 * a] = u01;
          r[2 * a + 1] = u23;
          r[4 + 2 * a] = *reinterpret_cast<const uint32_t*>(&l01);
          r[4 + 2 * a + 1] = *reinterpret_cast<const uint32_t*>(&l23);
        }
        tmem_st_16x256b_x2(taddr0 + ((uint32_t)(16 * h_st) << 16) + slot * 32u + (uint32_t)(16 * b), r);
      }
      h_st ^= 1;
      if (h_st == 0) {
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->full[slot]);
        g_st += 4;
        s_st += 4;
        if (s_st >= NST) {
          s_st -= NST;
          ws_.next(t_st, k_st);
        }
      }
    };
    Half R0, R1;
    bool v0 = issue(R0);
    while (v0) {
      const bool v1 = issue(R1);
      emit(R0);
      if (!v1) break;
      v0 = issue(R0);
      emit(R1);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int NST>
static size_t smem_bytes_m() { return (size_t)2 * (32 * NST / KB) * (16 * NST) * 128 + 1024 + 256; }

}  // namespace mg

// One CTA per component: t = max |Linv_k| -> scale 2^e with |Linv| 2^e < 2^14; write zero-padded fp16 hi / lo.
// D % 4 == 0 (a condition of the fp16 kernels): float4 loads, 8-byte stores of four halves.
__global__ void __launch_bounds__(512)
split_h16_kernel(const float* __restrict__ linv, int D, int Dp, __half* __restrict__ hi, __half* __restrict__ lo,
                 float* __restrict__ tmax, int full) {
  __shared__ float scratch[34];
  const int k = blockIdx.x;
  const float4* __restrict__ src = reinterpret_cast<const float4*>(linv + (long long)k * D * D);
  const int D4 = D >> 2, Dp4 = Dp >> 2;
  float m = 0.f;
  for (int i = threadIdx.x; i < D * D4; i += blockDim.x) {
    const float4 v = __ldg(src + i);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  m = block_max(m, scratch);
  if (threadIdx.x == 0) tmax[k] = m;
  const float sc = pow2_scale(m);
  uint2* __restrict__ dh = reinterpret_cast<uint2*>(hi + (long long)k * Dp * Dp);
  uint2* __restrict__ dl = reinterpret_cast<uint2*>(lo + (long long)k * Dp * Dp);
  for (int i = threadIdx.x; i < Dp * Dp4; i += blockDim.x) {
    const int r = i / Dp4, c4 = i - r * Dp4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < D && c4 < D4 && (full || 4 * c4 <= r)) {          // triangular input: blocks above the diagonal stay zero
      v = __ldg(src + (r * D4 + c4));
      // triangular input: strictly-upper entries are structurally zero
      v.x *= sc;
      v.y = (full || 4 * c4 + 1 <= r) ? v.y * sc : 0.f;
      v.z = (full || 4 * c4 + 2 <= r) ? v.z * sc : 0.f;
      v.w = (full || 4 * c4 + 3 <= r) ? v.w * sc : 0.f;
    }
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
    dh[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    dl[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
  }
}

// out[g] = max |in[r, c]| over the rows r of group g (group consecutive rows: one contiguous chunk)
__global__ void __launch_bounds__(256)
group_absmax_kernel(const float* __restrict__ in, long long rows, int cols, int group, float* __restrict__ out) {
  __shared__ float scratch[34];
  const long long r0 = (long long)blockIdx.x * group;
  const long long nr = min((long long)group, rows - r0);
  const float* p = in + r0 * cols;
  const long long n = nr * cols;
  float m = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(__ldg(p + i)));
  m = block_max(m, scratch);
  if (threadIdx.x == 0) out[blockIdx.x] = m;
}

static int make_map_h16(CUtensorMap* map, const void* base, int K, int Dp) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) {
    set_last_error("cuTensorMapEncodeTiled is not available from the driver");
    return GVI_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)Dp, (cuuint64_t)K * (cuuint64_t)Dp};
  cuuint64_t gstride[1] = {(cuuint64_t)Dp * 2};
  cuuint32_t box[2] = {KB, 64};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled (fp16 Linv) failed with CUresult %d", (int)r);
    return GVI_ERR_CUDA;
  }
  return GVI_OK;
}

static size_t smem_bytes(int Dp) {
  return (size_t)packed_rows(Dp) * 256 + STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
}

}  // namespace h16
}  // namespace gvi

using namespace gvi;

extern "C" int gvi_logdens_full_h16_supported(int D) { return (D >= 4 && D <= 256 && D % 4 == 0) ? 1 : 0; }
extern "C" int gvi_h16_padded_dim(int D) { return h16::padded_dim(D); }

extern "C" int gvi_split_h16_f32(const float* linv, int K, int D, void* hi, void* lo, float* tmax, void* stream) {
  GVI_REQUIRE(K >= 0 && D > 0, "gvi_split_h16_f32: bad sizes");
  if (K == 0) return GVI_OK;
  GVI_REQUIRE(linv && hi && lo && tmax, "gvi_split_h16_f32: null pointer");
  GVI_REQUIRE(D % 4 == 0 && reinterpret_cast<uintptr_t>(linv) % 16 == 0 && reinterpret_cast<uintptr_t>(hi) % 8 == 0 &&
                  reinterpret_cast<uintptr_t>(lo) % 8 == 0,
              "gvi_split_h16_f32: needs D %% 4 == 0 and aligned operands");
  h16::split_h16_kernel<<<K, 512, 0, (cudaStream_t)stream>>>(linv, D, h16::padded_dim(D), (__half*)hi, (__half*)lo,
                                                             tmax, 0);
  return check_launch("split_h16_kernel");
}

// The same split for a FULL matrix per component (the precision matrices of the tensor-core mixture gradient).
extern "C" int gvi_split_h16_full_f32(const float* mat, int K, int D, void* hi, void* lo, float* tmax, void* stream) {
  GVI_REQUIRE(K >= 0 && D > 0, "gvi_split_h16_full_f32: bad sizes");
  if (K == 0) return GVI_OK;
  GVI_REQUIRE(mat && hi && lo && tmax, "gvi_split_h16_full_f32: null pointer");
  GVI_REQUIRE(D % 4 == 0 && reinterpret_cast<uintptr_t>(mat) % 16 == 0 && reinterpret_cast<uintptr_t>(hi) % 8 == 0 &&
                  reinterpret_cast<uintptr_t>(lo) % 8 == 0,
              "gvi_split_h16_full_f32: needs D %% 4 == 0 and aligned operands");
  h16::split_h16_kernel<<<K, 512, 0, (cudaStream_t)stream>>>(mat, D, h16::padded_dim(D), (__half*)hi, (__half*)lo, tmax,
                                                             1);
  return check_launch("split_h16_kernel");
}

namespace gvi {
int launch_resp_mask(const float* lq, const float* logw, const float* logq, int K, int N, uint32_t* mask, cudaStream_t st);
}

extern "C" int gvi_mixture_grad_full_h16_supported(int D) {
  return (D % 4 == 0 && ((D > 64 && D <= 128) || (D > 192 && D <= 256))) ? 1 : 0;      // Dp = 128 or 256
}

// grad[N, D] = - sum_k r_kn P_k (x_n - mu_k) on the tensor cores (mg::mixgrad_h16_kernel).  p_hi / p_lo / tmaxp: the
// fp16 split of prec[K, D, D] from gvi_split_h16_full_f32; tileinf / minf as for gvi_logdens_full_h16_f32; ws: the
// block mask, gvi_mixture_grad_full_workspace(N, K) bytes.
extern "C" int gvi_mixture_grad_full_h16_f32(const float* X, const float* tileinf, int N, int D, const float* means,
                                             const float* minf, const void* p_hi, const void* p_lo, const float* tmaxp,
                                             const float* lq, const float* logw, const float* logq, int K, float* grad,
                                             void* ws, size_t ws_bytes, void* stream) {
  GVI_REQUIRE(N >= 0 && D > 0 && K >= 0, "gvi_mixture_grad_full_h16_f32: bad sizes");
  if (!gvi_mixture_grad_full_h16_supported(D)) {
    set_last_error("gvi_mixture_grad_full_h16_f32: D=%d unsupported (D %% 4 == 0 and 64 < D <= 128 or 192 < D <= 256)", D);
    return GVI_ERR_UNSUPPORTED;
  }
  if (N == 0) return GVI_OK;
  GVI_REQUIRE(X && tileinf && means && minf && p_hi && p_lo && tmaxp && lq && logw && logq && grad && ws,
              "gvi_mixture_grad_full_h16_f32: null pointer");
  GVI_REQUIRE(reinterpret_cast<uintptr_t>(X) % 16 == 0 && reinterpret_cast<uintptr_t>(means) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(grad) % 16 == 0 && reinterpret_cast<uintptr_t>(p_hi) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(p_lo) % 16 == 0,
              "gvi_mixture_grad_full_h16_f32: operands must be 16-byte aligned");
  const int T = ceil_div(N, h16::TILE_M), words = ceil_div(K, 32);
  GVI_REQUIRE(ws_bytes >= (size_t)T * words * sizeof(uint32_t), "gvi_mixture_grad_full_h16_f32: workspace too small");
  const int Dp = h16::padded_dim(D);
  GVI_REQUIRE((long long)K * Dp < 2147483647LL && (long long)N * D < 2147483647LL,
              "gvi_mixture_grad_full_h16_f32: K*D or N*D too large (32-bit offsets)");
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* mask = (uint32_t*)ws;
  int rc = launch_resp_mask(lq, logw, logq, K, N, mask, st);
  if (rc) return rc;
  cudaError_t e = cudaMemsetAsync(grad, 0, (size_t)N * D * sizeof(float), st);
  if (e != cudaSuccess) {
    set_last_error("gvi_mixture_grad_full_h16_f32: memset: %s", cudaGetErrorString(e));
    return GVI_ERR_CUDA;
  }
  CUtensorMap map_hi, map_lo;
  if ((rc = h16::make_map_h16(&map_hi, p_hi, K, Dp))) return rc;
  if ((rc = h16::make_map_h16(&map_lo, p_lo, K, Dp))) return rc;
  static int num_sms = 0;
  static unsigned long long dev_mask = 0;       // per device: the attributes below are per device
  if (first_call_on_device(dev_mask)) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaFuncSetAttribute(h16::mg::mixgrad_h16_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)h16::mg::smem_bytes_m<8>());
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(h16::mg::mixgrad_h16_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)h16::mg::smem_bytes_m<4>());
    if (e != cudaSuccess) {
      set_last_error("gvi_mixture_grad_full_h16_f32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      num_sms = 0;
      return GVI_ERR_CUDA;
    }
  }
  const int grid = min(num_sms, T);
  if (Dp == 256)
    h16::mg::mixgrad_h16_kernel<8><<<grid, h16::THREADS, h16::mg::smem_bytes_m<8>(), st>>>(
        map_hi, map_lo, X, tileinf, N, D, means, minf, tmaxp, lq, logw, logq, K, mask, words, grad);
  else
    h16::mg::mixgrad_h16_kernel<4><<<grid, h16::THREADS, h16::mg::smem_bytes_m<4>(), st>>>(
        map_hi, map_lo, X, tileinf, N, D, means, minf, tmaxp, lq, logw, logq, K, mask, words, grad);
  return check_launch("mixgrad_h16_kernel");
}

extern "C" int gvi_group_absmax_f32(const float* in, long long rows, int cols, int group, float* out, void* stream) {
  GVI_REQUIRE(rows >= 0 && cols > 0 && group > 0, "gvi_group_absmax_f32: bad sizes");
  if (rows == 0) return GVI_OK;
  GVI_REQUIRE(in && out, "gvi_group_absmax_f32: null pointer");
  const long long groups = (rows + group - 1) / group;
  h16::group_absmax_kernel<<<(unsigned)groups, 256, 0, (cudaStream_t)stream>>>(in, rows, cols, group, out);
  return check_launch("group_absmax_kernel");
}

extern "C" int gvi_logdens_full_h16_f32(const float* X, const float* tileinf, int N, int D, const float* means,
                                        const float* minf, const void* linv_hi, const void* linv_lo,
                                        const float* tmax, const float* cst, int K, float* lq, void* stream) {
  GVI_REQUIRE(N >= 0 && D > 0 && K >= 0, "gvi_logdens_full_h16_f32: bad sizes");
  if (!gvi_logdens_full_h16_supported(D)) {
    set_last_error("gvi_logdens_full_h16_f32: D=%d unsupported (needs D %% 4 == 0, 4 <= D <= 256)", D);
    return GVI_ERR_UNSUPPORTED;
  }
  if (N == 0 || K == 0) return GVI_OK;
  GVI_REQUIRE(X && tileinf && means && minf && linv_hi && linv_lo && tmax && cst && lq,
              "gvi_logdens_full_h16_f32: null pointer");
  GVI_REQUIRE(reinterpret_cast<uintptr_t>(X) % 16 == 0 && reinterpret_cast<uintptr_t>(means) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(linv_hi) % 16 == 0 && reinterpret_cast<uintptr_t>(linv_lo) % 16 == 0,
              "gvi_logdens_full_h16_f32: operands must be 16-byte aligned");
  const int Dp = h16::padded_dim(D);
  GVI_REQUIRE((long long)K * Dp < 2147483647LL && (long long)N * D < 2147483647LL &&
                  (long long)ceil_div(N, h16::TILE_M) * K * 8 < 2147483647LL,
              "gvi_logdens_full_h16_f32: K*D, N*D or the work list too large (32-bit offsets)");
  CUtensorMap map_hi, map_lo;
  int rc = h16::make_map_h16(&map_hi, linv_hi, K, Dp);
  if (rc) return rc;
  rc = h16::make_map_h16(&map_lo, linv_lo, K, Dp);
  if (rc) return rc;
  static int num_sms = 0;
  static int a_in_tmem = 1;      // GMMVI_B200_H16_A=smem selects the shared-memory A operand for every size
  static unsigned long long dev_mask = 0;       // per device: the attributes below are per device
  if (first_call_on_device(dev_mask)) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(h16::logdens_h16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)h16::smem_bytes(256));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(h16::h16t::logdens_h16t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)h16::h16t::smem_bytes_t(256));
    if (e != cudaSuccess) {
      set_last_error("gvi_logdens_full_h16_f32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      num_sms = 0;
      return GVI_ERR_CUDA;
    }
    const char* env = getenv("GMMVI_B200_H16_A");
    a_in_tmem = (env != nullptr && strcmp(env, "smem") == 0) ? 0 : 1;
  }
  const long long total = (long long)ceil_div(N, h16::TILE_M) * K;
  const int grid = (int)min((long long)num_sms, total);
  // the TMEM-A kernel caches the mean rows of a CTA's components in shared memory: a contiguous item range touches at
  // most ceil(items / T) + 1 components
  const int T = ceil_div(N, h16::TILE_M);
  const bool means_fit = ((total + grid - 1) / grid + T - 1) / T + 1 <= h16::h16t::MEANS_CACHED_MAX;
  if (a_in_tmem && Dp >= 128 && means_fit) {
    h16::h16t::logdens_h16t_kernel<<<grid, h16::THREADS, h16::h16t::smem_bytes_t(Dp), (cudaStream_t)stream>>>(
        map_hi, map_lo, X, tileinf, N, D, Dp, means, minf, tmax, cst, K, lq);
    return check_launch("logdens_h16t_kernel");
  }
  h16::logdens_h16_kernel<<<grid, h16::THREADS, h16::smem_bytes(Dp), (cudaStream_t)stream>>>(
      map_hi, map_lo, X, tileinf, N, D, Dp, means, minf, tmax, cst, K, lq);
  return check_launch("logdens_h16_kernel");
}
