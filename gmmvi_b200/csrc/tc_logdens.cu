// Full-covariance component log densities on the 5th-generation tensor cores (sm_100a only).
//
//   lq[k, n] = cst[k] - 1/2 | Linv_k (x_n - mu_k) |^2            (models/full_cov_gmm.py:56-62)
//
// One persistent CTA per SM walks the (component k, 128-sample tile) work list, k-major so that the CTAs
// running concurrently share Linv_k in L2.  Per work item the contraction Z[128 x D] = Diff[128 x D] Linv_k^T
// runs as tcgen05.mma kind::tf32 with fp32 accumulation in TMEM, in "3xTF32" split precision:
//       Z += A_hi B_hi + A_lo B_hi + A_hi B_lo,   x = hi + lo with hi, lo exactly representable in TF32,
// which keeps ~22 mantissa bits per product (a single TF32 pass would not hold the 1e-5 tolerance).
// Warp roles (384 threads):
//   warp 0      TMA producer: Linv_k (pre-split hi / lo, [K*D, D] row-major) -> 128B-swizzled smem, only the
//               rows i >= 32*kb of k-block kb (Linv is lower triangular: the rest is structurally zero)
//   warp 1      MMA issuer (one thread): per k-block 4 x 3 tcgen05.mma with N = D - 32*kb
//   warp 2      TMEM allocator (512 columns = two 128 x 256 fp32 accumulators, double buffered)
//   warps 4-7   epilogue: tcgen05.ld the accumulator, row sums of squares, write lq
//   warps 8-11  A producers: x - mu_k formed in fp32 BEFORE the split (cancellation), written as hi / lo
//               into the K-major SWIZZLE_128B layout the UMMA descriptor expects
// Pipelines: smem full/empty ring (2 stages of {A_hi, A_lo, B_hi, B_lo}), TMEM full/empty pair.
#include "common.cuh"
#include <cuda.h>

namespace gvi {
namespace tc {

constexpr int TILE_M = 128;
constexpr int KBLK = 32;                 // fp32 elements per 128-byte swizzle row
constexpr int STAGES = 2;
constexpr int THREADS = 384;
constexpr int A_BYTES = TILE_M * 128;    // 16 KB per hi / lo
constexpr int B_BYTES = 256 * 128;       // 32 KB per hi / lo (N <= 256 rows)
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // 96 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int TMEM_COLS = 512;
constexpr int ACC_COLS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100 version 1):
// 8-row x 128-byte atoms, stride between atoms (SBO) 1024 B, LBO unused (1).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::tf32, fp32 accumulate, both operands K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

struct Barriers {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(THREADS, 1)
tc_logdens_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                  const float* __restrict__ X, int N, int D, const float* __restrict__ means,
                  const float* __restrict__ cst, int K, float* __restrict__ lq) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Barriers* bars = reinterpret_cast<Barriers*>(smem + STAGES * STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = ceil_div(N, TILE_M);
  const long long total = (long long)T * K;
  const int nkb = D / KBLK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bars->full[s], 128 + 1);
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

  if (warp == 0) {
    // ---------------- TMA producer for B = Linv_k (hi, lo) ----------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (long long w = blockIdx.x; w < total; w += gridDim.x) {
        const int k = (int)(w / T);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&bars->empty[s], ph ^ 1);
          uint8_t* st = smem + s * STAGE_BYTES;
          const int rows = D - kb * KBLK;
          mbar_arrive_expect_tx(&bars->full[s], 2u * rows * 128u);
          for (int r = 0; r < rows; r += 32) {
            tma_load_2d(st + 2 * A_BYTES + r * 128, &map_hi, &bars->full[s], kb * KBLK, k * D + kb * KBLK + r);
            tma_load_2d(st + 2 * A_BYTES + B_BYTES + r * 128, &map_lo, &bars->full[s], kb * KBLK,
                        k * D + kb * KBLK + r);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      long long it = 0;
      for (long long w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int buf = (int)(it & 1);
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(&bars->acc_empty[buf], (use & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&bars->full[s], ph);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t idesc = make_idesc(D - kb * KBLK);
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * ACC_COLS + kb * KBLK);
#pragma unroll
          for (int ks = 0; ks < KBLK / 8; ++ks) {
            const uint64_t a_hi = make_desc(st + ks * 32);
            const uint64_t a_lo = make_desc(st + A_BYTES + ks * 32);
            const uint64_t b_hi = make_desc(st + 2 * A_BYTES + ks * 32);
            const uint64_t b_lo = make_desc(st + 2 * A_BYTES + B_BYTES + ks * 32);
            umma_tf32(d_tmem, a_hi, b_hi, idesc, (kb | ks) != 0 ? 1u : 0u);
            umma_tf32(d_tmem, a_lo, b_hi, idesc, 1u);
            umma_tf32(d_tmem, a_hi, b_lo, idesc, 1u);
          }
          umma_commit(&bars->empty[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&bars->acc_full[buf]);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ---------------- epilogue: row sums of squares ----------------
    const int q = warp - 4;
    long long it = 0;
    for (long long w = blockIdx.x; w < total; w += gridDim.x, ++it) {
      const int k = (int)(w / T), t = (int)(w % T);
      const int buf = (int)(it & 1);
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(&bars->acc_full[buf], use & 1);
      tc_fence_after();
      float s0 = 0.f, s1 = 0.f;
      for (int c = 0; c < nkb; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * ACC_COLS + c * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float a = __uint_as_float(v[i]), b = __uint_as_float(v[i + 1]);
          s0 = fmaf(a, a, s0);
          s1 = fmaf(b, b, s1);
        }
      }
      tc_fence_before();
      mbar_arrive(&bars->acc_empty[buf]);
      const int n = t * TILE_M + 32 * q + lane;
      if (n < N) lq[(long long)k * N + n] = __ldg(cst + k) - 0.5f * (s0 + s1);
    }
  } else if (warp >= 8) {
    // ---------------- A producers: diff = x - mu_k, split into TF32 hi / lo ----------------
    // Thread p owns the 16-byte chunk c = p % 8 of rows r0, r0 + 16, ..., r0 + 112 (r0 = p / 8): a warp-wide
    // 128-bit load then covers 4 complete 128-byte row segments (4 L1 wavefronts instead of 32).
    const int p = threadIdx.x - 256;
    const int c = p & 7, r0 = p >> 3;
    const int swz = (c ^ (r0 & 7)) << 4;     // (row & 7) == (r0 & 7) for every row this thread touches
    // The (work item, k-block) sequence is flattened and the global loads of step j+1 are issued before step j is
    // converted, so that the L2 latency of X overlaps the split / shared-memory stores of the previous block.
    int s = 0;
    uint32_t ph = 0;
    const long long nwork = total > blockIdx.x ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long nitems = nwork * nkb;
    // cursors of the load stream (one step ahead) and of the store stream; divisions only once per work item
    long long w_ld = blockIdx.x, w_st = blockIdx.x;
    int kb_ld = 0, kb_st = 0;
    int k_ld = (int)(w_ld / T), nb_ld = (int)(w_ld % T) * TILE_M + r0, nb_st = nb_ld;
    auto issue = [&](float4 (&xv)[8], float4& m) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int n = nb_ld + 16 * q;
        xv[q] = n < N ? __ldg(reinterpret_cast<const float4*>(X + (long long)n * D) + kb_ld * 8 + c)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      m = __ldg(reinterpret_cast<const float4*>(means + (long long)k_ld * D) + kb_ld * 8 + c);
      if (++kb_ld == nkb) {
        kb_ld = 0;
        w_ld += gridDim.x;
        if (w_ld < total) {
          k_ld = (int)(w_ld / T);
          nb_ld = (int)(w_ld % T) * TILE_M + r0;
        }
      }
    };
    auto emit = [&](const float4 (&xv)[8], const float4& m) {
      mbar_wait(&bars->empty[s], ph ^ 1);
      uint8_t* st = smem + s * STAGE_BYTES;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const bool rowok = (nb_st + 16 * q) < N;
        float4 d;
        d.x = rowok ? xv[q].x - m.x : 0.f;
        d.y = rowok ? xv[q].y - m.y : 0.f;
        d.z = rowok ? xv[q].z - m.z : 0.f;
        d.w = rowok ? xv[q].w - m.w : 0.f;
        float4 hi, lo;
        hi.x = to_tf32(d.x); hi.y = to_tf32(d.y); hi.z = to_tf32(d.z); hi.w = to_tf32(d.w);
        lo.x = to_tf32(d.x - hi.x); lo.y = to_tf32(d.y - hi.y); lo.z = to_tf32(d.z - hi.z); lo.w = to_tf32(d.w - hi.w);
        const int off = (r0 + 16 * q) * 128 + swz;
        *reinterpret_cast<float4*>(st + off) = hi;
        *reinterpret_cast<float4*>(st + A_BYTES + off) = lo;
      }
      fence_proxy_async();
      mbar_arrive(&bars->full[s]);
      if (++s == STAGES) { s = 0; ph ^= 1; }
      if (++kb_st == nkb) {
        kb_st = 0;
        w_st += gridDim.x;
        if (w_st < total) nb_st = (int)(w_st % T) * TILE_M + r0;
      }
    };
    float4 xa[8], xb[8], ma, mb;
    if (nitems > 0) issue(xa, ma);
    for (long long j = 0; j < nitems; j += 2) {
      if (j + 1 < nitems) issue(xb, mb);
      emit(xa, ma);
      if (j + 1 < nitems) {
        if (j + 2 < nitems) issue(xa, ma);
        emit(xb, mb);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// x -> (hi, lo) with hi = rn_tf32(x), lo = rn_tf32(x - hi)
__global__ void split_tf32_kernel(const float* __restrict__ in, long long n, float* __restrict__ hi,
                                  float* __restrict__ lo) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = in[i];
    const float h = to_tf32(x);
    hi[i] = h;
    lo[i] = to_tf32(x - h);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
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

static int make_map(CUtensorMap* map, const float* base, int K, int D) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) {
    set_last_error("cuTensorMapEncodeTiled is not available from the driver");
    return GVI_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)K * (cuuint64_t)D};
  cuuint64_t gstride[1] = {(cuuint64_t)D * sizeof(float)};
  cuuint32_t box[2] = {KBLK, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return GVI_ERR_CUDA;
  }
  return GVI_OK;
}

}  // namespace tc
}  // namespace gvi

using namespace gvi;

extern "C" int gvi_split_tf32_f32(const float* in, long long n, float* hi, float* lo, void* stream) {
  GVI_REQUIRE(n >= 0, "gvi_split_tf32_f32: bad size");
  if (n == 0) return GVI_OK;
  GVI_REQUIRE(in && hi && lo, "gvi_split_tf32_f32: null pointer");
  const int blocks = (int)min((long long)148 * 8, (n + 255) / 256);
  tc::split_tf32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, n, hi, lo);
  return check_launch("split_tf32_kernel");
}

extern "C" int gvi_logdens_full_tc_supported(int D) { return (D % 32 == 0 && D >= 32 && D <= 256) ? 1 : 0; }

extern "C" int gvi_logdens_full_tc_f32(const float* X, int N, int D, const float* means, const float* linv_hi,
                                       const float* linv_lo, const float* cst, int K, float* lq, void* stream) {
  GVI_REQUIRE(N >= 0 && D > 0 && K >= 0, "gvi_logdens_full_tc_f32: bad sizes");
  if (!gvi_logdens_full_tc_supported(D)) {
    set_last_error("gvi_logdens_full_tc_f32: D=%d unsupported (needs D %% 32 == 0, 32 <= D <= 256)", D);
    return GVI_ERR_UNSUPPORTED;
  }
  if (N == 0 || K == 0) return GVI_OK;
  GVI_REQUIRE(X && means && linv_hi && linv_lo && cst && lq, "gvi_logdens_full_tc_f32: null pointer");
  GVI_REQUIRE(reinterpret_cast<uintptr_t>(X) % 16 == 0 && reinterpret_cast<uintptr_t>(means) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(linv_hi) % 16 == 0 && reinterpret_cast<uintptr_t>(linv_lo) % 16 == 0,
              "gvi_logdens_full_tc_f32: operands must be 16-byte aligned");
  GVI_REQUIRE((long long)K * D < 2147483647LL, "gvi_logdens_full_tc_f32: K*D too large");
  CUtensorMap map_hi, map_lo;
  int rc = tc::make_map(&map_hi, linv_hi, K, D);
  if (rc) return rc;
  rc = tc::make_map(&map_lo, linv_lo, K, D);
  if (rc) return rc;
  static int num_sms = 0;
  static unsigned long long dev_mask = 0;       // per device: the attributes below are per device
  if (first_call_on_device(dev_mask)) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(tc::tc_logdens_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         tc::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("gvi_logdens_full_tc_f32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      num_sms = 0;
      return GVI_ERR_CUDA;
    }
  }
  const long long total = (long long)ceil_div(N, tc::TILE_M) * K;
  const int grid = (int)min((long long)num_sms, total);
  tc::tc_logdens_kernel<<<grid, tc::THREADS, tc::SMEM_BYTES, (cudaStream_t)stream>>>(map_hi, map_lo, X, N, D, means,
                                                                                    cst, K, lq);
  return check_launch("tc_logdens_kernel");
}
