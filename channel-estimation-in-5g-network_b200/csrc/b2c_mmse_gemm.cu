// K4: dense Wiener filter at the pilots on the 5th-generation tensor cores.
//
//   out[c][i] = sum_j W[i][j] * in[c][j]        complex64, W [np][np], c over (slot, rx) columns
//
// replaces `mmse_matrix @ h_ls` of MMSEEstimator.estimate_at_pilots' known-covariance branch
// (src/baseline_estimators.py:181-190).  The complex product is run as ONE real GEMM on the
// interleaved (re, im) views:   D[i'][c] = sum_j' A[i'][j'] * X[c][j'],  i' = 2i+q, j' = 2j+p,
//   A[2i][2j] = Wr, A[2i][2j+1] = -Wi, A[2i+1][2j] = Wi, A[2i+1][2j+1] = Wr      (built on the fly)
// so X is `in` and D is `out` exactly as they lie in memory (no de-interleave pass).
//
// Tensor-core path: tcgen05.mma.cta_group::1.kind::tf32, M = 128 (rows i'), N = 128 (columns c),
// K = 8 per instruction, operands in shared memory in the canonical K-major no-swizzle layout
// (8 x 16-byte core matrices), fp32 accumulator in TMEM (128 lanes x 128 columns), read back with
// tcgen05.ld.  Plain TF32 (10-bit mantissa) cannot hold the 1e-4 parity bound, so each operand is
// split  v = hi + lo  (hi = v with the low 13 mantissa bits cleared, lo = v - hi, both exact) and
// three MMAs are issued per K step:  hi*hi + hi*lo + lo*hi   ("3xTF32", error ~2^-21).
#include <stdlib.h>

#include "b2c_common.cuh"

namespace b2c {

constexpr int TC_BM = 128;    // rows i' per CTA (TMEM lanes)
constexpr int TC_BN = 128;    // columns c per CTA (TMEM columns)
constexpr int TC_BK = 32;     // real k per stage = 8 x 16-byte chunks per row
constexpr int TC_THREADS = 128;
constexpr int TC_KCH = TC_BK / 4;                 // 16-byte chunks per row per stage
constexpr int TC_LBO_A = TC_BM * 16;              // bytes between consecutive K chunks (A tile)
constexpr int TC_LBO_B = TC_BN * 16;
constexpr int TC_SBO = 128;                       // bytes between consecutive 8-row groups
constexpr int TC_TILE_A = TC_BM * TC_BK * 4;      // bytes of one A tile (hi or lo)
constexpr int TC_TILE_B = TC_BN * TC_BK * 4;
constexpr int TC_SMEM = 2 * TC_TILE_A + 2 * TC_TILE_B + 1024;   // hi+lo for A and B, + alignment slack

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4, [16,30) leading byte offset>>4, [32,46) stride byte offset>>4, [46,48) version = 1.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

// kind::tf32 instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2),
// both K-major (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void split_tf32(float v, float &hi, float &lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  lo = v - hi;
}

__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok)
               : "r"(bar), "r"(parity)
               : "memory");
  return ok;
}

// REALW = false: W complex [np][np] (m = k = np), the interleaved embedding described above.
// REALW = true : W REAL [m][k] applied to the real and imaginary parts alike (the Clough-Tocher 'cubic'
//                interpolation as a dense linear map, [nsym*nsc] x [npilots]): column c of `in` becomes two
//                GEMM columns (re, im), so a CTA covers 64 complex columns and K runs over k, not 2k.
// Element (row r, real column kk) of an operand tile in the canonical K-major no-swizzle layout (bytes).
__host__ __device__ __forceinline__ int tile_off(int r, int kk, int lbo) { return (kk >> 2) * lbo + (r >> 3) * TC_SBO + (r & 7) * 16 + (kk & 3) * 4; }

// One-time operand preparation (per Wiener matrix / interpolation map): the A operand of every (row tile, K stage)
// is written to global memory already split into its TF32 hi / lo parts and already in the shared-memory tile
// layout, [hi 16 KB][lo 16 KB] per tile, so that the GEMM fetches a stage of A with ONE 32 KB bulk copy
// (cp.async.bulk -> mbarrier) and spends no CUDA-core work on it.
template <bool REALW>
__global__ void __launch_bounds__(256) dense_prepare_kernel(const float *__restrict__ Wf, int m, int k, float *__restrict__ prep,
                                                            int nstages) {
  const int tm = blockIdx.x, st = blockIdx.y;
  float *hi = prep + ((int64_t)tm * nstages + st) * (2 * TC_TILE_A / 4), *lo = hi + TC_TILE_A / 4;
  for (int e = threadIdx.x; e < TC_BM * TC_BK; e += 256) {
    const int r = e / TC_BK, kk = e - r * TC_BK;
    const int row = tm * TC_BM + r, col = st * TC_BK + kk;
    float v = 0.f;
    if (REALW) {
      if (row < m && col < k) v = __ldg(Wf + (int64_t)row * k + col);
    } else {
      const int i = row >> 1, j = col >> 1;
      if (i < k && j < k) {
        const float2 w = __ldg(reinterpret_cast<const float2 *>(Wf) + (int64_t)i * k + j);
        v = (row & 1) ? ((col & 1) ? w.x : w.y) : ((col & 1) ? -w.y : w.x);
      }
    }
    float h, l;
    split_tf32(v, h, l);
    const int o = tile_off(r, kk, TC_LBO_A) / 4;
    hi[o] = h;
    lo[o] = l;
  }
}

template <bool REALW, bool PREP>
__global__ void __launch_bounds__(TC_THREADS, 3) dense_tc_kernel(const float *__restrict__ Wf, int m, int k,
                                                              const float2 *__restrict__ in, float *__restrict__ out,
                                                              int64_t ncols, int64_t ld_in, int64_t ld_out) {
  const float2 *__restrict__ W = reinterpret_cast<const float2 *>(Wf);
  const int np = k;              // complex path: square matrix
  const int64_t ld = ld_in;
  extern __shared__ unsigned char smem_dyn[];
  __shared__ uint32_t tmem_base_sm;
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ __align__(8) uint64_t a_bar;          // PREP: completion of the bulk copy of this stage's A tiles

  // 1024-byte aligned operand tiles: [A hi][A lo][B hi][B lo]
  const uint32_t s0 = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char *sp = smem_dyn + (s0 - smem_u32(smem_dyn));
  unsigned char *sAh = sp, *sAl = sp + TC_TILE_A, *sBh = sp + 2 * TC_TILE_A, *sBl = sp + 2 * TC_TILE_A + TC_TILE_B;
  const uint32_t aAh = s0, aAl = s0 + TC_TILE_A, aBh = s0 + 2 * TC_TILE_A, aBl = s0 + 2 * TC_TILE_A + TC_TILE_B;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int i0c = blockIdx.x * (TC_BM / 2);          // complex path: first complex row of W in this tile
  const int e0 = blockIdx.x * TC_BM;                 // real path: first row of W in this tile
  const int64_t c0 = (int64_t)blockIdx.y * (REALW ? TC_BN / 2 : TC_BN);    // first (complex) column
  const int kreal = REALW ? k : 2 * np;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_sm)), "r"((uint32_t)TC_BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mma_bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&a_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_d = tmem_base_sm;

  // ---- per-thread load / store assignments -----------------------------------------------------
  // The canonical layout puts element (row r, chunk kc, byte e) at kc*LBO + (r>>3)*128 + (r&7)*16 + e,
  // so the shared-memory bank is decided by (r & 7, e) alone.  The maps below make the 16 lanes of
  // every half-warp store to 16 distinct 8-byte bank pairs (conflict-free) while the global loads
  // still consume whole 32-byte sectors.
  //   A (64 complex rows il x 16 complex cols jl of W per stage, 8 values per thread):
  //     il = 8*q' .. : il[1:0] = lane[1:0], il[2] = warp[1], il[5:3] = q ; jl[2:0] = lane[4:2], jl[3] = warp[0]
  //     each value feeds two real rows (2il, 2il+1); lane[3] picks which of them goes out first, so
  //     one store instruction covers even and odd rows (all 16 bank pairs) instead of only even ones.
  //   B (128 rows cl x 16 complex cols jl of `in` per stage, 16 values per thread):
  //     cl[2:0] = lane[2:0], cl[6:3] = q ; jl[0] = lane[3], jl[1] = lane[4], jl[3:2] = warp
  // DEEP (prepared complex operand: the A registers are free): two register sets for the data operand, so that its
  // global loads run two K stages ahead of their use (+2.4 %; the real-W variant is register-bound at 168 and loses 3.5 %)
  constexpr bool DEEP = PREP && !REALW;
  float2 ra[8], rb0[16], rb1[DEEP ? 16 : 1], ra_w[REALW ? 16 : 1];
  const int a_il_lo = (lane & 3) | ((warp >> 1) << 2), a_jl = ((lane >> 2) & 7) | ((warp & 1) << 3);
  const int b_cl_lo = lane & 7, b_jl = ((lane >> 3) & 3) | (warp << 2);
  // real-W maps (16 values per thread for A and for B):
  //   A (128 rows r x 16 float pairs jl): r[2:0] = lane[2:0], r[6:3] = q ; jl as for the complex B tile
  //   B (64 complex columns cc x 32 k): cc[1:0] = lane[1:0], cc[5:2] = q ; kk[2:0] = lane[4:2], kk[4:3] = warp;
  //     value (re, im) goes to rows 2cc / 2cc+1 as two 4-byte stores, lane[4] picks which first
  const int rb_cc_lo = lane & 3, rb_kk = ((lane >> 2) & 7) | (warp << 3);
  auto load_stage = [&](int k0, float2 (&rb)[16]) {     // k0: real k offset of the stage (multiple of 32)
    if (REALW) {
      if (!PREP) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int r = e0 + (q << 3) + b_cl_lo, j = k0 + 2 * b_jl;
          const float *src = Wf + (int64_t)r * k + j;
          ra_w[q] = make_float2((r < m && j < k) ? __ldg(src) : 0.f, (r < m && j + 1 < k) ? __ldg(src + 1) : 0.f);
        }
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int64_t c = c0 + (q << 2) + rb_cc_lo;
        const int j = k0 + rb_kk;
        rb[q] = (c < ncols && j < k) ? __ldg(in + c * ld_in + j) : make_float2(0.f, 0.f);
      }
      return;
    }
    const int jc0 = k0 >> 1;
    if (!PREP) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int i = i0c + (q << 3) + a_il_lo, j = jc0 + a_jl;
        ra[q] = (i < np && j < np) ? __ldg(W + (int64_t)i * np + j) : make_float2(0.f, 0.f);
      }
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int64_t c = c0 + (q << 3) + b_cl_lo;
      const int j = jc0 + b_jl;
      rb[q] = (c < ncols && j < np) ? __ldg(in + c * ld + j) : make_float2(0.f, 0.f);
    }
  };
  auto store_stage = [&](float2 (&rb)[16]) {
    if (REALW) {
      const int kc = b_jl >> 1, eo = (b_jl & 1) * 8;
      if (!PREP) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int r = (q << 3) + b_cl_lo;
          float x_h, x_l, y_h, y_l;
          split_tf32(ra_w[q].x, x_h, x_l);
          split_tf32(ra_w[q].y, y_h, y_l);
          const int o = kc * TC_LBO_A + (r >> 3) * TC_SBO + (r & 7) * 16 + eo;
          *reinterpret_cast<float2 *>(sAh + o) = make_float2(x_h, y_h);
          *reinterpret_cast<float2 *>(sAl + o) = make_float2(x_l, y_l);
        }
      }
      const int kc2 = rb_kk >> 2, eo2 = (rb_kk & 3) * 4;
      const bool im_first = (lane >> 4) & 1;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int cc = (q << 2) + rb_cc_lo;
        float xr_h, xr_l, xi_h, xi_l;
        split_tf32(rb[q].x, xr_h, xr_l);
        split_tf32(rb[q].y, xi_h, xi_l);
        const int r0 = 2 * cc, r1 = 2 * cc + 1;
        const int o0 = kc2 * TC_LBO_B + (r0 >> 3) * TC_SBO + (r0 & 7) * 16 + eo2;   // row 2c  : real part
        const int o1 = kc2 * TC_LBO_B + (r1 >> 3) * TC_SBO + (r1 & 7) * 16 + eo2;   // row 2c+1: imaginary part
        const int of = im_first ? o1 : o0, os = im_first ? o0 : o1;
        *reinterpret_cast<float *>(sBh + of) = im_first ? xi_h : xr_h;
        *reinterpret_cast<float *>(sBl + of) = im_first ? xi_l : xr_l;
        *reinterpret_cast<float *>(sBh + os) = im_first ? xr_h : xi_h;
        *reinterpret_cast<float *>(sBl + os) = im_first ? xr_l : xi_l;
      }
      return;
    }
    const int a_kc = a_jl >> 1, a_eo = (a_jl & 1) * 8;
    const bool odd_first = (lane >> 3) & 1;
#pragma unroll
    for (int q = 0; q < (PREP ? 0 : 8); ++q) {
      const int il = (q << 3) + a_il_lo;
      float wr_h, wr_l, wi_h, wi_l;
      split_tf32(ra[q].x, wr_h, wr_l);
      split_tf32(ra[q].y, wi_h, wi_l);
      const int r0 = 2 * il, r1 = 2 * il + 1;
      const int o0 = a_kc * TC_LBO_A + (r0 >> 3) * TC_SBO + (r0 & 7) * 16 + a_eo;   // row 2i  : ( Wr, -Wi)
      const int o1 = a_kc * TC_LBO_A + (r1 >> 3) * TC_SBO + (r1 & 7) * 16 + a_eo;   // row 2i+1: ( Wi,  Wr)
      const float2 e_h = make_float2(wr_h, -wi_h), e_l = make_float2(wr_l, -wi_l);
      const float2 d_h = make_float2(wi_h, wr_h), d_l = make_float2(wi_l, wr_l);
      const int of = odd_first ? o1 : o0, os = odd_first ? o0 : o1;
      *reinterpret_cast<float2 *>(sAh + of) = odd_first ? d_h : e_h;
      *reinterpret_cast<float2 *>(sAl + of) = odd_first ? d_l : e_l;
      *reinterpret_cast<float2 *>(sAh + os) = odd_first ? e_h : d_h;
      *reinterpret_cast<float2 *>(sAl + os) = odd_first ? e_l : d_l;
    }
    const int b_kc = b_jl >> 1, b_eo = (b_jl & 1) * 8;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int cl = (q << 3) + b_cl_lo;
      float xr_h, xr_l, xi_h, xi_l;
      split_tf32(rb[q].x, xr_h, xr_l);
      split_tf32(rb[q].y, xi_h, xi_l);
      const int o = b_kc * TC_LBO_B + (cl >> 3) * TC_SBO + (cl & 7) * 16 + b_eo;
      *reinterpret_cast<float2 *>(sBh + o) = make_float2(xr_h, xi_h);
      *reinterpret_cast<float2 *>(sBl + o) = make_float2(xr_l, xi_l);
    }
  };

  const int nstages = (kreal + TC_BK - 1) / TC_BK;
  uint32_t parity = 0;
  const unsigned char *prep = reinterpret_cast<const unsigned char *>(Wf) + (int64_t)blockIdx.x * nstages * (2 * TC_TILE_A);
  constexpr int AHEAD = DEEP ? 2 : 1;
  auto do_stage = [&](int st, float2 (&rb)[16]) {
    if (PREP && tid == 0) {
      // the MMAs of the previous stage have completed (waited below), so both A tiles are free: fetch this stage's
      // pre-split hi / lo pair (contiguous 32 KB, already in tile layout) while the threads stage B
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&a_bar)), "r"(2u * TC_TILE_A) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(aAh),
                   "l"(prep + (int64_t)st * (2 * TC_TILE_A)), "r"(2u * TC_TILE_A), "r"(smem_u32(&a_bar))
                   : "memory");
    }
    store_stage(rb);
    if (st + AHEAD < nstages) load_stage((st + AHEAD) * TC_BK, rb);   // registers are free again: the next stage's global loads fly
                                                          // under the barrier, the MMAs and the wait for them
    // generic-proxy smem writes -> visible to the tensor core (async proxy), then CTA barrier
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      if (PREP) {
        while (!mbar_try_wait(smem_u32(&a_bar), parity)) {
        }
      }
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
      for (int kk = 0; kk < TC_BK / 8; ++kk) {     // one instruction covers K = 8 = two 16-byte chunks
        const uint64_t dAh = umma_desc(aAh + kk * 2 * TC_LBO_A, TC_LBO_A, TC_SBO);
        const uint64_t dAl = umma_desc(aAl + kk * 2 * TC_LBO_A, TC_LBO_A, TC_SBO);
        const uint64_t dBh = umma_desc(aBh + kk * 2 * TC_LBO_B, TC_LBO_B, TC_SBO);
        const uint64_t dBl = umma_desc(aBl + kk * 2 * TC_LBO_B, TC_LBO_B, TC_SBO);
        umma_tf32(tmem_d, dAh, dBh, (st | kk) != 0);
        umma_tf32(tmem_d, dAh, dBl, 1u);
        umma_tf32(tmem_d, dAl, dBh, 1u);
      }
      // arrives on the mbarrier once every MMA issued so far has finished reading smem / writing TMEM
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mma_bar))
                   : "memory");
    }
    while (!mbar_try_wait(smem_u32(&mma_bar), parity)) {
    }
    parity ^= 1u;
  };
  load_stage(0, rb0);
  if constexpr (DEEP) {
    float2 (&rbB)[16] = reinterpret_cast<float2 (&)[16]>(rb1);
    if (nstages > 1) load_stage(TC_BK, rbB);
    for (int st = 0; st < nstages; st += 2) {
      do_stage(st, rb0);
      if (st + 1 < nstages) do_stage(st + 1, rbB);
    }
  } else {
    for (int st = 0; st < nstages; ++st) do_stage(st, rb0);
  }

  // ---- epilogue: TMEM -> registers -> global (row i' contiguous in memory) -----------------------
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const int ip = blockIdx.x * TC_BM + warp * 32 + lane;     // real output row of this thread (its TMEM lane)
  const int64_t ld2 = 2 * ld_out;
#pragma unroll 1
  for (int cb = 0; cb < TC_BN; cb += 32) {
    uint32_t r[32];
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    if (REALW) {
      if (ip < m) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {      // TMEM columns (2q, 2q+1) = (re, im) of complex column c
          const int64_t c = c0 + (cb >> 1) + q;
          if (c < ncols)
            *reinterpret_cast<float2 *>(out + c * ld2 + 2 * ip) = make_float2(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]));
        }
      }
    } else if (ip < kreal) {
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const int64_t c = c0 + cb + q;
        if (c < ncols) out[c * ld2 + ip] = __uint_as_float(r[q]);   // a warp writes 32 consecutive floats
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"((uint32_t)TC_BN));
}


// ---- N = 256 form of the prepared complex product --------------------------------------------------------------
// Same operands and arithmetic as dense_tc_kernel<false, true>, but one CTA (256 threads) owns a 128 x 256 tile: the
// A tiles (the 32 KB bulk copy and the 3 x 4 KB the MMAs read per K step) are amortised over twice the columns, which
// is what bounds the 128-column form (shared-memory bandwidth: operand reads + staging writes).  TMEM: 256 columns,
// two CTAs per SM use all 512.  Warps 0-3 / 4-7 stage and read back the column halves 0-127 / 128-255.
constexpr int T2_BN = 256;
constexpr int T2_THREADS = 256;
constexpr int T2_LBO_B = T2_BN * 16;
constexpr int T2_TILE_B = T2_BN * TC_BK * 4;
constexpr int T2_SMEM = 2 * TC_TILE_A + 2 * T2_TILE_B + 1024;
constexpr uint32_t T2_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(T2_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32_n256(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(T2_IDESC), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(T2_THREADS, 2) dense_tc256_kernel(const float *__restrict__ prepared, int np,
                                                                   const float2 *__restrict__ in, float *__restrict__ out,
                                                                   int64_t ncols, int64_t ld) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ uint32_t tmem_base_sm;
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ __align__(8) uint64_t a_bar;
  const uint32_t s0 = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char *sp = smem_dyn + (s0 - smem_u32(smem_dyn));
  unsigned char *sBh = sp + 2 * TC_TILE_A, *sBl = sBh + T2_TILE_B;
  const uint32_t aAh = s0, aAl = s0 + TC_TILE_A, aBh = s0 + 2 * TC_TILE_A, aBl = aBh + T2_TILE_B;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t c0 = (int64_t)blockIdx.y * T2_BN;
  const int kreal = 2 * np;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_sm)), "r"((uint32_t)T2_BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mma_bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&a_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_d = tmem_base_sm;

  // B staging map (16 values per thread): column c = half * 128 + 8 q + lane[2:0], half = warp[2];
  // complex j within the stage: jl[0] = lane[3], jl[1] = lane[4], jl[3:2] = warp[1:0]   (conflict-free, see above)
  const int b_cl_lo = ((warp >> 2) << 7) | (lane & 7), b_jl = ((lane >> 3) & 3) | ((warp & 3) << 2);
  const int b_kc = b_jl >> 1, b_eo = (b_jl & 1) * 8;
  float2 rb0[16], rb1[16];
  auto load_stage = [&](int k0, float2 (&rb)[16]) {
    const int j = (k0 >> 1) + b_jl;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int64_t c = c0 + (q << 3) + b_cl_lo;
      rb[q] = (c < ncols && j < np) ? __ldg(in + c * ld + j) : make_float2(0.f, 0.f);
    }
  };
  auto store_stage = [&](float2 (&rb)[16]) {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int cl = (q << 3) + b_cl_lo;
      float xr_h, xr_l, xi_h, xi_l;
      split_tf32(rb[q].x, xr_h, xr_l);
      split_tf32(rb[q].y, xi_h, xi_l);
      const int o = b_kc * T2_LBO_B + (cl >> 3) * TC_SBO + (cl & 7) * 16 + b_eo;
      *reinterpret_cast<float2 *>(sBh + o) = make_float2(xr_h, xi_h);
      *reinterpret_cast<float2 *>(sBl + o) = make_float2(xr_l, xi_l);
    }
  };

  const int nstages = (kreal + TC_BK - 1) / TC_BK;
  uint32_t parity = 0;
  const unsigned char *prep = reinterpret_cast<const unsigned char *>(prepared) + (int64_t)blockIdx.x * nstages * (2 * TC_TILE_A);
  auto do_stage = [&](int st, float2 (&rb)[16]) {
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&a_bar)), "r"(2u * TC_TILE_A) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(aAh),
                   "l"(prep + (int64_t)st * (2 * TC_TILE_A)), "r"(2u * TC_TILE_A), "r"(smem_u32(&a_bar))
                   : "memory");
    }
    store_stage(rb);
    if (st + 2 < nstages) load_stage((st + 2) * TC_BK, rb);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      while (!mbar_try_wait(smem_u32(&a_bar), parity)) {
      }
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
      for (int kk = 0; kk < TC_BK / 8; ++kk) {
        const uint64_t dAh = umma_desc(aAh + kk * 2 * TC_LBO_A, TC_LBO_A, TC_SBO);
        const uint64_t dAl = umma_desc(aAl + kk * 2 * TC_LBO_A, TC_LBO_A, TC_SBO);
        const uint64_t dBh = umma_desc(aBh + kk * 2 * T2_LBO_B, T2_LBO_B, TC_SBO);
        const uint64_t dBl = umma_desc(aBl + kk * 2 * T2_LBO_B, T2_LBO_B, TC_SBO);
        umma_tf32_n256(tmem_d, dAh, dBh, (st | kk) != 0);
        umma_tf32_n256(tmem_d, dAh, dBl, 1u);
        umma_tf32_n256(tmem_d, dAl, dBh, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mma_bar))
                   : "memory");
    }
    while (!mbar_try_wait(smem_u32(&mma_bar), parity)) {
    }
    parity ^= 1u;
  };
  load_stage(0, rb0);
  if (nstages > 1) load_stage(TC_BK, rb1);
  for (int st = 0; st < nstages; st += 2) {
    do_stage(st, rb0);
    if (st + 1 < nstages) do_stage(st + 1, rb1);
  }

  // epilogue: warp w reads TMEM lanes 32 (w & 3) .. + 31 (its hardware quarter), columns 128 (w >> 2) .. + 127
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const int ip = blockIdx.x * TC_BM + (warp & 3) * 32 + lane;
  const int chalf = (warp >> 2) * 128;
  const int64_t ld2 = 2 * ld;
#pragma unroll 1
  for (int cb = 0; cb < 128; cb += 32) {
    uint32_t r[32];
    const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(chalf + cb);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    if (ip < kreal) {
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const int64_t c = c0 + chalf + cb + q;
        if (c < ncols) out[c * ld2 + ip] = __uint_as_float(r[q]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"((uint32_t)T2_BN));
}


// ---- warp-specialised form of the prepared complex product ------------------------------------------------------
// One CTA per SM, 128 x 256 tile, two-stage shared-memory ring (stage = A hi/lo 32 KB + B hi 32 KB + B lo 32 KB).
//   warp 0, lane 0 : MMA issuer -- waits full barriers only, so the 12 MMAs of consecutive stages queue back to back
//   warp 1, lane 0 : A producer -- one 32 KB bulk copy of the pre-split tiles per stage
//   warps 2..9     : B producers -- global loads two stages ahead in registers, TF32 split, staging stores
//   then warps 2..9 read the accumulator back (TMEM -> global)
// Barriers per ring slot: fullA (transaction bytes), fullB (one arrival per producer warp), empty (tcgen05.commit).
constexpr int WS_THREADS = 320;
constexpr int WS_STAGE = 2 * TC_TILE_A + 2 * T2_TILE_B;      // 96 KB
constexpr int WS_SMEM = 2 * WS_STAGE + 1024;

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// REALW = false: complex W [np][np] embedded as a real 2np x 2np operand (m = k = np), a tile covers 256 columns.
// REALW = true : real W [m][k] applied to re / im alike, a tile covers 128 complex columns (= 256 GEMM columns).
// Grouped launch: several independent products (one Wiener matrix per (pattern, SNR) group of the batch, each applied to
// its own contiguous block of columns) share ONE grid, so that the small per-group tile counts (14 row tiles x a few
// column tiles) fill the 148 SMs together instead of one partial wave each.
constexpr int MAX_GROUPS = 32;
struct GroupTable {
  int ngroups;                          // 0: plain launch, blockIdx.x = row tile, blockIdx.y = column tile
  int tile_end[MAX_GROUPS];             // running sum of tiles (row tiles x column tiles) over the groups
  int tiles_m[MAX_GROUPS];
  int np[MAX_GROUPS];
  long long col0[MAX_GROUPS], ncols[MAX_GROUPS];
  const float *prepared[MAX_GROUPS];
};

template <bool REALW, bool GROUPED = false>
__global__ void __launch_bounds__(WS_THREADS, 1) dense_tc_ws_kernel(const float *__restrict__ prepared, int m, int k,
                                                                   const float2 *__restrict__ in, float *__restrict__ out,
                                                                   int64_t ncols, int64_t ld_in, int64_t ld_out,
                                                                   const __grid_constant__ GroupTable gt) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ uint32_t tmem_base_sm;
  __shared__ __align__(8) uint64_t full_a[2], full_b[2], empty[2], acc_bar;
  const uint32_t s0 = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char *sp = smem_dyn + (s0 - smem_u32(smem_dyn));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int tile_m = blockIdx.x, tile_n = blockIdx.y;
  if constexpr (GROUPED) {
    // which group owns linear tile blockIdx.x, and which (row, column) tile of that group it is; selects over the
    // unrolled table keep every access to the kernel parameters statically indexed
    const int tile = blockIdx.x;
    int g = 0;
#pragma unroll
    for (int i = 0; i < MAX_GROUPS; ++i)
      if (i < gt.ngroups - 1 && tile >= gt.tile_end[i]) g = i + 1;
    int begin = 0, tm_count = 1;
    long long col0 = 0;
#pragma unroll
    for (int i = 0; i < MAX_GROUPS; ++i) {
      if (i == g) {
        begin = i ? gt.tile_end[i - 1] : 0;
        tm_count = gt.tiles_m[i];
        m = k = gt.np[i];
        col0 = gt.col0[i];
        ncols = gt.ncols[i];
        prepared = gt.prepared[i];
      }
    }
    const int local = tile - begin;
    tile_n = local / tm_count;
    tile_m = local - tile_n * tm_count;
    in += col0 * ld_in;
    out += 2 * col0 * ld_out;
  }
  const int64_t c0 = (int64_t)tile_n * (REALW ? T2_BN / 2 : T2_BN);     // first (complex) column of the tile
  const int np = k, kreal = REALW ? k : 2 * k;
  const int mreal = REALW ? m : 2 * k;
  const int64_t ld = ld_in;
  const int nstages = (kreal + TC_BK - 1) / TC_BK;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_sm)), "r"((uint32_t)T2_BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&full_a[b])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;\n" ::"r"(smem_u32(&full_b[b])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&empty[b])));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&acc_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_d = tmem_base_sm;

  if (warp == 0) {
    if (lane == 0) {
      for (int st = 0; st < nstages; ++st) {
        const int b = st & 1;
        const uint32_t ph = (uint32_t)(st >> 1) & 1u;
        const uint32_t aAh = s0 + b * WS_STAGE, aAl = aAh + TC_TILE_A, aBh = aAh + 2 * TC_TILE_A, aBl = aBh + T2_TILE_B;
        mbar_wait(smem_u32(&full_a[b]), ph);
        mbar_wait(smem_u32(&full_b[b]), ph);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
        for (int kk = 0; kk < TC_BK / 8; ++kk) {
          const uint64_t dAh = umma_desc(aAh + kk * 2 * TC_LBO_A, TC_LBO_A, TC_SBO);
          const uint64_t dAl = umma_desc(aAl + kk * 2 * TC_LBO_A, TC_LBO_A, TC_SBO);
          const uint64_t dBh = umma_desc(aBh + kk * 2 * T2_LBO_B, T2_LBO_B, TC_SBO);
          const uint64_t dBl = umma_desc(aBl + kk * 2 * T2_LBO_B, T2_LBO_B, TC_SBO);
          umma_tf32_n256(tmem_d, dAh, dBh, (st | kk) != 0);
          umma_tf32_n256(tmem_d, dAh, dBl, 1u);
          umma_tf32_n256(tmem_d, dAl, dBh, 1u);
        }
        // ring slot b is free once these MMAs have read it
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&empty[b])) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&acc_bar)) : "memory");
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const unsigned char *prep = reinterpret_cast<const unsigned char *>(prepared) + (int64_t)tile_m * nstages * (2 * TC_TILE_A);
      for (int st = 0; st < nstages; ++st) {
        const int b = st & 1;
        if (st >= 2) mbar_wait(smem_u32(&empty[b]), (uint32_t)((st >> 1) - 1) & 1u);
        const uint32_t bar = smem_u32(&full_a[b]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(2u * TC_TILE_A) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                         s0 + b * WS_STAGE),
                     "l"(prep + (int64_t)st * (2 * TC_TILE_A)), "r"(2u * TC_TILE_A), "r"(bar)
                     : "memory");
      }
    }
  } else {
    // B producers: the staging maps of dense_tc_kernel / dense_tc256_kernel with producer warp pw = warp - 2 in 0..7
    // (pw >> 2 picks the tile half: rows 0-127 / 128-255 of the B tile)
    const int pw = warp - 2;
    const int b_cl_lo = ((pw >> 2) << 7) | (lane & 7), b_jl = ((lane >> 3) & 3) | ((pw & 3) << 2);
    const int b_kc = b_jl >> 1, b_eo = (b_jl & 1) * 8;
    // real-W map: complex column cc = 64 (pw >> 2) + 4 q + lane[1:0], k within the stage kk = lane[4:2] | (pw & 3) << 3
    const int rb_cc_lo = ((pw >> 2) << 6) | (lane & 3), rb_kk = ((lane >> 2) & 7) | ((pw & 3) << 3);
    float2 rb0[16], rb1[16];
    auto load_stage = [&](int st, float2 (&rb)[16]) {
      if (REALW) {
        const int j = st * TC_BK + rb_kk;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int64_t c = c0 + (q << 2) + rb_cc_lo;
          rb[q] = (c < ncols && j < k) ? __ldg(in + c * ld + j) : make_float2(0.f, 0.f);
        }
        return;
      }
      const int j = st * (TC_BK / 2) + b_jl;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int64_t c = c0 + (q << 3) + b_cl_lo;
        rb[q] = (c < ncols && j < np) ? __ldg(in + c * ld + j) : make_float2(0.f, 0.f);
      }
    };
    auto produce = [&](int st, float2 (&rb)[16]) {
      const int b = st & 1;
      if (st >= 2) mbar_wait(smem_u32(&empty[b]), (uint32_t)((st >> 1) - 1) & 1u);
      unsigned char *sBh = sp + b * WS_STAGE + 2 * TC_TILE_A, *sBl = sBh + T2_TILE_B;
      if (REALW) {
        const int kc2 = rb_kk >> 2, eo2 = (rb_kk & 3) * 4;
        const bool im_first = (lane >> 4) & 1;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int cc = (q << 2) + rb_cc_lo;
          float xr_h, xr_l, xi_h, xi_l;
          split_tf32(rb[q].x, xr_h, xr_l);
          split_tf32(rb[q].y, xi_h, xi_l);
          const int r0 = 2 * cc, r1 = 2 * cc + 1;
          const int o0 = kc2 * T2_LBO_B + (r0 >> 3) * TC_SBO + (r0 & 7) * 16 + eo2;   // row 2c  : real part
          const int o1 = kc2 * T2_LBO_B + (r1 >> 3) * TC_SBO + (r1 & 7) * 16 + eo2;   // row 2c+1: imaginary part
          const int of = im_first ? o1 : o0, os = im_first ? o0 : o1;
          *reinterpret_cast<float *>(sBh + of) = im_first ? xi_h : xr_h;
          *reinterpret_cast<float *>(sBl + of) = im_first ? xi_l : xr_l;
          *reinterpret_cast<float *>(sBh + os) = im_first ? xr_h : xi_h;
          *reinterpret_cast<float *>(sBl + os) = im_first ? xr_l : xi_l;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int cl = (q << 3) + b_cl_lo;
          float xr_h, xr_l, xi_h, xi_l;
          split_tf32(rb[q].x, xr_h, xr_l);
          split_tf32(rb[q].y, xi_h, xi_l);
          const int o = b_kc * T2_LBO_B + (cl >> 3) * TC_SBO + (cl & 7) * 16 + b_eo;
          *reinterpret_cast<float2 *>(sBh + o) = make_float2(xr_h, xi_h);
          *reinterpret_cast<float2 *>(sBl + o) = make_float2(xr_l, xi_l);
        }
      }
      if (st + 2 < nstages) load_stage(st + 2, rb);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // this thread's stores -> async proxy
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&full_b[b])) : "memory");
    };
    load_stage(0, rb0);
    if (nstages > 1) load_stage(1, rb1);
    for (int st = 0; st < nstages; st += 2) {
      produce(st, rb0);
      if (st + 1 < nstages) produce(st + 1, rb1);
    }

    // epilogue: TMEM lanes 32 (warp & 3) .. + 31 (this warp's hardware quarter), columns 128 (pw >> 2) .. + 127
    mbar_wait(smem_u32(&acc_bar), 0u);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const int ip = tile_m * TC_BM + (warp & 3) * 32 + lane;
    const int chalf = (pw >> 2) * 128;
    const int64_t ld2 = 2 * ld_out;
#pragma unroll 1
    for (int cb = 0; cb < 128; cb += 32) {
      uint32_t r[32];
      const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(chalf + cb);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      if (REALW) {
        if (ip < mreal) {
#pragma unroll
          for (int q = 0; q < 16; ++q) {      // TMEM columns (2q, 2q+1) = (re, im) of complex column c
            const int64_t c = c0 + ((chalf + cb) >> 1) + q;
            if (c < ncols)
              *reinterpret_cast<float2 *>(out + c * ld2 + 2 * ip) = make_float2(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]));
          }
        }
      } else if (ip < mreal) {
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const int64_t c = c0 + chalf + cb + q;
          if (c < ncols) out[c * ld2 + ip] = __uint_as_float(r[q]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"((uint32_t)T2_BN));
}


// ---- data operand in TMEM ("TS" form) ----------------------------------------------------------------------------------
// The warp-specialised SS kernel above is bound by the shared-memory pipe (ncu: L1/shared data pipe 79-81 %, tensor pipe
// 63 %): every K step the three MMAs read A (4 KB) and B (8 KB) from shared memory, and the data operand is first WRITTEN
// there by the CUDA cores (64 KB of hi / lo tiles per stage).  Here the roles are swapped:
//     D[c][i'] = sum_j' X[c][j'] * Wt[i'][j']        M = 128 data columns c,  N = 2 x 128 real output rows i'
//   * A = the data, straight from global memory into registers, split into TF32 hi / lo and stored to TENSOR MEMORY with
//     tcgen05.st (one lane per column c, 32 K-columns per stage): it never touches shared memory;
//   * B = the prepared Wiener tiles (b2c_dense_prepare's hi / lo pairs, unchanged), two 128-row tiles per CTA fetched by
//     cp.async.bulk into a three-stage ring: the only shared-memory traffic left is the bulk write and the MMAs' B reads
//     (107 B/clk/SM at full tensor rate instead of 159 of the 128 the pipe can carry).
//   * the accumulator row of a thread is one column c with 256 consecutive output reals: 16-byte stores.
// warp 0: MMA issuer, warp 1: bulk producer, warps 2-5: data producers (their TMEM lane quarter) and epilogue.
constexpr int TA_THREADS = 192;
constexpr int TA_NST = 3;
constexpr int TA_STAGE = 2 * (2 * TC_TILE_A);                 // two W row tiles, hi + lo each: 64 KB
constexpr int TA_SMEM = TA_NST * TA_STAGE + 1024;
constexpr int TA_BM = 128;                                    // data columns per CTA (TMEM lanes)
constexpr uint32_t TA_COL_D = 0, TA_COL_A = 256, TA_COLS = 512;

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(TC_IDESC), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}

// CL2 (opt-in, B2C_DENSE_CLUSTER=1): launched as clusters of two CTAs that share a pair of W row tiles and take adjacent
// column tiles.  Each CTA fetches ONE of the two W tiles per stage and multicasts it into both CTAs' rings
// (cp.async.bulk ... .multicast::cluster), so a stage costs an SM 32 KB of L2 reads for W instead of 64 KB.  A ring slot is reused once BOTH CTAs' MMAs have released it: the commits are multicast
// to both CTAs' empty barriers (count 2).  The grouped tile order is pair-major with an even column-tile count per group.
template <bool GROUPED, bool CL2 = false>
__global__ void __launch_bounds__(TA_THREADS, 1) dense_tc_ta_kernel(const float *__restrict__ prepared, int np,
                                                                   const float2 *__restrict__ in, float *__restrict__ out,
                                                                   int64_t ncols, int64_t ld,
                                                                   const __grid_constant__ GroupTable gt) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ uint32_t tmem_base_sm;
  __shared__ __align__(8) uint64_t full_a[TA_NST], full_b[TA_NST], empty[TA_NST], acc_bar;
  const uint32_t s0 = (smem_u32(smem_dyn) + 1023u) & ~1023u;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int pair = blockIdx.x, tile_n = blockIdx.y;
  if constexpr (GROUPED) {
    const int tile = blockIdx.x;
    int g = 0;
#pragma unroll
    for (int i = 0; i < MAX_GROUPS; ++i)
      if (i < gt.ngroups - 1 && tile >= gt.tile_end[i]) g = i + 1;
    int begin = 0, pairs = 1, end = 0;
    long long col0 = 0;
#pragma unroll
    for (int i = 0; i < MAX_GROUPS; ++i) {
      if (i == g) {
        begin = i ? gt.tile_end[i - 1] : 0;
        end = gt.tile_end[i];
        pairs = gt.tiles_m[i];                      // here: PAIRS of W row tiles
        np = gt.np[i];
        col0 = gt.col0[i];
        ncols = gt.ncols[i];
        prepared = gt.prepared[i];
      }
    }
    const int local = tile - begin;
    if constexpr (CL2) {                            // pair-major: the two CTAs of a cluster differ in the column tile only
      const int ntn = (end - begin) / pairs;
      pair = local / ntn;
      tile_n = local - pair * ntn;
    } else {
      tile_n = local / pairs;
      pair = local - tile_n * pairs;
    }
    in += col0 * ld;
    out += 2 * col0 * ld;
  }
  uint32_t crank = 0;
  if constexpr (CL2) asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(crank));
  const int kreal = 2 * np;                                   // K and the output rows are both the 2 np real indices
  const int nstages = (kreal + TC_BK - 1) / TC_BK;
  const int tiles_w = (kreal + TC_BM - 1) / TC_BM;
  const int tm0 = 2 * pair;
  const bool has2 = tm0 + 1 < tiles_w;
  const int64_t c0 = (int64_t)tile_n * TA_BM;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_sm)), "r"(TA_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < TA_NST; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;\n" ::"r"(smem_u32(&full_a[b])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&full_b[b])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&empty[b])), "r"(CL2 ? 2u : 1u));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&acc_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if constexpr (CL2) {      // the peer's barriers exist before anything of ours can land on them
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_sm;

  if (warp == 0) {
    if (lane == 0) {
      for (int st = 0; st < nstages; ++st) {
        const int b = st % TA_NST;
        const uint32_t ph = (uint32_t)(st / TA_NST) & 1u;
        mbar_wait(smem_u32(&full_a[b]), ph);
        mbar_wait(smem_u32(&full_b[b]), ph);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t aA = tmem + TA_COL_A + (uint32_t)b * 64u;              // hi: 32 columns, lo: the next 32
        const uint32_t sB = s0 + b * TA_STAGE;
#pragma unroll
        for (int kk = 0; kk < TC_BK / 8; ++kk) {
          // per accumulator the same three products in the same order as the SS kernels (W hi x hi, W hi x lo, W lo x hi);
          // the two accumulator halves alternate, so that consecutive MMAs never accumulate into the same TMEM tile
          const uint32_t b0 = sB, b1 = sB + 2 * TC_TILE_A;
          const uint64_t dBh0 = umma_desc(b0 + kk * 2 * TC_LBO_A, TC_LBO_A, TC_SBO), dBl0 = umma_desc(b0 + TC_TILE_A + kk * 2 * TC_LBO_A, TC_LBO_A, TC_SBO);
          const uint64_t dBh1 = umma_desc(b1 + kk * 2 * TC_LBO_A, TC_LBO_A, TC_SBO), dBl1 = umma_desc(b1 + TC_TILE_A + kk * 2 * TC_LBO_A, TC_LBO_A, TC_SBO);
          const uint32_t d0 = tmem + TA_COL_D, d1 = d0 + 128u, ah = aA + kk * 8, al = aA + 32 + kk * 8;
          const uint32_t acc = (st | kk) != 0;
          umma_tf32_ts(d0, ah, dBh0, acc);
          if (has2) umma_tf32_ts(d1, ah, dBh1, acc);
          umma_tf32_ts(d0, al, dBh0, 1u);
          if (has2) umma_tf32_ts(d1, al, dBh1, 1u);
          umma_tf32_ts(d0, ah, dBl0, 1u);
          if (has2) umma_tf32_ts(d1, ah, dBl1, 1u);
        }
        if constexpr (CL2)     // the slot is free once BOTH CTAs have read it (each multicasts W tiles into both rings)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                           smem_u32(&empty[b])), "h"((uint16_t)3)
                       : "memory");
        else
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&empty[b])) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&acc_bar)) : "memory");
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const unsigned char *prep = reinterpret_cast<const unsigned char *>(prepared);
      const uint32_t bytes = has2 ? 2u * (2u * TC_TILE_A) : 2u * TC_TILE_A;
      for (int st = 0; st < nstages; ++st) {
        const int b = st % TA_NST;
        if (st >= TA_NST) mbar_wait(smem_u32(&empty[b]), (uint32_t)(st / TA_NST - 1) & 1u);
        const uint32_t bar = smem_u32(&full_b[b]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
        if constexpr (CL2) {
          // this CTA fetches W tile h = its cluster rank (rank 0 alone when the pair has one tile) for BOTH CTAs
          const int h = (int)crank;
          if (h == 0 || has2)
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n" ::"r"(
                    s0 + b * TA_STAGE + h * (2 * TC_TILE_A)),
                "l"(prep + ((int64_t)(tm0 + h) * nstages + st) * (2 * TC_TILE_A)), "r"(2u * TC_TILE_A), "r"(bar), "h"((uint16_t)3)
                : "memory");
        } else {
          for (int h = 0; h < (has2 ? 2 : 1); ++h)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                             s0 + b * TA_STAGE + h * (2 * TC_TILE_A)),
                         "l"(prep + ((int64_t)(tm0 + h) * nstages + st) * (2 * TC_TILE_A)), "r"(2u * TC_TILE_A), "r"(bar)
                         : "memory");
        }
      }
    }
  } else {
    // data producers: this thread owns data column c (TMEM lane 32 q + lane of its warp's quarter q = warp % 4)
    const int q = warp & 3;
    const int64_t c = c0 + q * 32 + lane;
    const bool cvalid = c < ncols;
    const float *row = reinterpret_cast<const float *>(in) + c * 2 * ld;
    const int64_t row_floats = 2 * ld;
    float4 cur[8], nxt[8];
    auto load_stage = [&](int st, float4 (&r)[8]) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int e = st * TC_BK + 4 * i;
        r[i] = (cvalid && e + 4 <= row_floats && e < kreal) ? __ldg(reinterpret_cast<const float4 *>(row + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (e + 4 > kreal) {                       // the K tail: elements past 2 np never enter the product
          if (e + 0 >= kreal) r[i].x = 0.f;
          if (e + 1 >= kreal) r[i].y = 0.f;
          if (e + 2 >= kreal) r[i].z = 0.f;
          if (e + 3 >= kreal) r[i].w = 0.f;
        }
      }
    };
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    load_stage(0, cur);
    for (int st = 0; st < nstages; ++st) {
      const int b = st % TA_NST;
      if (st + 1 < nstages) load_stage(st + 1, nxt);
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float v[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float h, l;
          split_tf32(v[j], h, l);
          hi[4 * i + j] = __float_as_uint(h);
          lo[4 * i + j] = __float_as_uint(l);
        }
      }
      if (st >= TA_NST) {
        mbar_wait(smem_u32(&empty[b]), (uint32_t)(st / TA_NST - 1) & 1u);       // the MMAs that read this TMEM stage are done
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      }
      tmem_st32(lane_base + TA_COL_A + (uint32_t)b * 64u, hi);
      tmem_st32(lane_base + TA_COL_A + (uint32_t)b * 64u + 32u, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&full_a[b])) : "memory");
#pragma unroll
      for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
    }

    // epilogue: this thread's accumulator row = column c, 256 consecutive output reals
    mbar_wait(smem_u32(&acc_bar), 0u);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    float *orow = out + c * 2 * ld;
    const bool vec_ok = ((2 * ld) & 3) == 0;
#pragma unroll 1
    for (int cb = 0; cb < (has2 ? 256 : 128); cb += 32) {
      uint32_t r[32];
      const uint32_t taddr = lane_base + TA_COL_D + (uint32_t)cb;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      if (cvalid) {
        const int i0 = tm0 * TC_BM + cb;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const int ip = i0 + e;
          if (vec_ok && ip + 4 <= kreal) {
            *reinterpret_cast<float4 *>(orow + ip) = make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]), __uint_as_float(r[e + 2]),
                                                                 __uint_as_float(r[e + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (ip + j < kreal) orow[ip + j] = __uint_as_float(r[e + j]);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if constexpr (CL2) {      // neither CTA leaves while the other may still multicast into its ring or arrive on its barriers
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(TA_COLS));
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_mmse_dense(const float *W, int32_t np, const float *in, float *out, int64_t ncols, int64_t ld,
                              void *stream) {
  B2C_REQUIRE(W && in && out, B2C_E_ARG, "b2c_mmse_dense: null argument");
  B2C_REQUIRE(np >= 1 && ld >= np && ncols >= 0, B2C_E_ARG, "b2c_mmse_dense: np=%d ld=%lld ncols=%lld", np, (long long)ld,
              (long long)ncols);
  B2C_REQUIRE(in != out, B2C_E_ARG, "b2c_mmse_dense: in-place not supported");
  if (ncols == 0) return B2C_OK;
  dim3 grid((unsigned)((2 * np + TC_BM - 1) / TC_BM), (unsigned)((ncols + TC_BN - 1) / TC_BN));
  B2C_REQUIRE(grid.y <= 65535, B2C_E_UNSUPPORTED, "b2c_mmse_dense: ncols=%lld too large for one launch", (long long)ncols);
  B2C_CUDA((set_max_smem<dense_tc_kernel<false, false>>(TC_SMEM)));
  dense_tc_kernel<false, false><<<grid, TC_THREADS, TC_SMEM, (cudaStream_t)stream>>>(W, np, np, reinterpret_cast<const float2 *>(in),
                                                                                      out, ncols, ld, ld);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

extern "C" int b2c_dense_real_apply(const float *W, int32_t m, int32_t k, const float *in, float *out, int64_t ncols,
                                    int64_t ld_in, int64_t ld_out, void *stream) {
  B2C_REQUIRE(W && in && out, B2C_E_ARG, "b2c_dense_real_apply: null argument");
  B2C_REQUIRE(m >= 1 && k >= 1 && ld_in >= k && ld_out >= m && ncols >= 0, B2C_E_ARG,
              "b2c_dense_real_apply: m=%d k=%d ld_in=%lld ld_out=%lld ncols=%lld", m, k, (long long)ld_in, (long long)ld_out,
              (long long)ncols);
  B2C_REQUIRE(in != out, B2C_E_ARG, "b2c_dense_real_apply: in-place not supported");
  if (ncols == 0) return B2C_OK;
  dim3 grid((unsigned)((m + TC_BM - 1) / TC_BM), (unsigned)((ncols + TC_BN / 2 - 1) / (TC_BN / 2)));
  B2C_REQUIRE(grid.y <= 65535, B2C_E_UNSUPPORTED, "b2c_dense_real_apply: ncols=%lld too large for one launch", (long long)ncols);
  B2C_CUDA((set_max_smem<dense_tc_kernel<true, false>>(TC_SMEM)));
  dense_tc_kernel<true, false><<<grid, TC_THREADS, TC_SMEM, (cudaStream_t)stream>>>(W, m, k, reinterpret_cast<const float2 *>(in), out,
                                                                                     ncols, ld_in, ld_out);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

// ---- prepared-operand variants -------------------------------------------------------------------------------
static void prep_dims(int32_t m, int32_t k, int32_t is_complex, int &tiles_m, int &nstages) {
  const int mr = is_complex ? 2 * k : m, kr = is_complex ? 2 * k : k;
  tiles_m = (mr + TC_BM - 1) / TC_BM;
  nstages = (kr + TC_BK - 1) / TC_BK;
}

extern "C" int64_t b2c_dense_prepared_bytes(int32_t m, int32_t k, int32_t is_complex) {
  if (m < 1 || k < 1) return 0;
  int tiles_m, nstages;
  prep_dims(m, k, is_complex, tiles_m, nstages);
  return (int64_t)tiles_m * nstages * 2 * TC_TILE_A;
}

extern "C" int b2c_dense_prepare(const float *W, int32_t m, int32_t k, int32_t is_complex, void *prepared, void *stream) {
  B2C_REQUIRE(W && prepared && m >= 1 && k >= 1 && (!is_complex || m == k), B2C_E_ARG, "b2c_dense_prepare: m=%d k=%d", m, k);
  B2C_REQUIRE(((uintptr_t)prepared & 15) == 0, B2C_E_ARG, "b2c_dense_prepare: workspace must be 16-byte aligned");
  int tiles_m, nstages;
  prep_dims(m, k, is_complex, tiles_m, nstages);
  dim3 grid((unsigned)tiles_m, (unsigned)nstages);
  if (is_complex) dense_prepare_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(W, m, k, static_cast<float *>(prepared), nstages);
  else dense_prepare_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(W, m, k, static_cast<float *>(prepared), nstages);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

extern "C" int b2c_dense_apply_prepared(const void *prepared, int32_t m, int32_t k, int32_t is_complex, const float *in,
                                        float *out, int64_t ncols, int64_t ld_in, int64_t ld_out, void *stream) {
  B2C_REQUIRE(prepared && in && out && m >= 1 && k >= 1 && (!is_complex || m == k), B2C_E_ARG,
              "b2c_dense_apply_prepared: null argument or m=%d k=%d", m, k);
  B2C_REQUIRE(ld_in >= k && ld_out >= m && ncols >= 0 && in != out, B2C_E_ARG,
              "b2c_dense_apply_prepared: ld_in=%lld ld_out=%lld ncols=%lld", (long long)ld_in, (long long)ld_out, (long long)ncols);
  B2C_REQUIRE(((uintptr_t)prepared & 15) == 0, B2C_E_ARG, "b2c_dense_apply_prepared: workspace must be 16-byte aligned");
  if (ncols == 0) return B2C_OK;
  int tiles_m, nstages;
  prep_dims(m, k, is_complex, tiles_m, nstages);
  const float *P = static_cast<const float *>(prepared);
  if (is_complex) {
    B2C_REQUIRE(ld_in == ld_out, B2C_E_ARG, "b2c_dense_apply_prepared: complex form takes one leading dimension");
    // Two tile shapes, equal per-SM throughput: 128 x 128 (3 CTAs/SM) and 128 x 256 (2 CTAs/SM).  What differs is the
    // wave quantisation at this column count, so take the shape whose last wave is fuller.
    int sm_count = 0;
    B2C_CUDA((cudaError_t)sm_count_current(&sm_count));
    if (ncols >= TA_BM && (ld_in & 1) == 0 && ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0 && !getenv("B2C_DENSE_SS")) {
      // data operand in tensor memory (see dense_tc_ta_kernel): one CTA per pair of W row tiles x 128 columns
      dim3 grid((unsigned)((tiles_m + 1) / 2), (unsigned)((ncols + TA_BM - 1) / TA_BM));
      B2C_REQUIRE(grid.y <= 65535, B2C_E_UNSUPPORTED, "b2c_dense_apply_prepared: ncols=%lld too large for one launch", (long long)ncols);
      B2C_CUDA((set_max_smem<dense_tc_ta_kernel<false>>(TA_SMEM)));
      dense_tc_ta_kernel<false><<<grid, TA_THREADS, TA_SMEM, (cudaStream_t)stream>>>(P, k, reinterpret_cast<const float2 *>(in), out, ncols, ld_in,
                                                                                     GroupTable{});
      B2C_CUDA(cudaGetLastError());
      return B2C_OK;
    }
    auto wave_eff = [&](int64_t tiles, int per_sm) {
      const int64_t slots = (int64_t)sm_count * per_sm;
      return (double)tiles / (double)(((tiles + slots - 1) / slots) * slots);
    };
    const int64_t t128 = (int64_t)tiles_m * ((ncols + TC_BN - 1) / TC_BN), t256 = (int64_t)tiles_m * ((ncols + T2_BN - 1) / T2_BN);
    // Warp-specialised form (one CTA per SM, two-stage ring) whenever its waves are not much emptier than the best
    // of the two single-stage shapes; it is ~20 % faster per SM.
    const double eff_ws = wave_eff(t256, 1), eff_best = wave_eff(t256, 2) > wave_eff(t128, 3) ? wave_eff(t256, 2) : wave_eff(t128, 3);
    if (ncols >= T2_BN && 1.2 * eff_ws >= eff_best) {
      dim3 grid2((unsigned)tiles_m, (unsigned)((ncols + T2_BN - 1) / T2_BN));
      B2C_REQUIRE(grid2.y <= 65535, B2C_E_UNSUPPORTED, "b2c_dense_apply_prepared: ncols=%lld too large for one launch", (long long)ncols);
      B2C_CUDA((set_max_smem<dense_tc_ws_kernel<false>>(WS_SMEM)));
      dense_tc_ws_kernel<false><<<grid2, WS_THREADS, WS_SMEM, (cudaStream_t)stream>>>(P, k, k, reinterpret_cast<const float2 *>(in), out,
                                                                                      ncols, ld_in, ld_out, GroupTable{});
      B2C_CUDA(cudaGetLastError());
      return B2C_OK;
    }
    if (ncols >= T2_BN && wave_eff(t256, 2) > wave_eff(t128, 3) + 0.02) {
      dim3 grid2((unsigned)tiles_m, (unsigned)((ncols + T2_BN - 1) / T2_BN));
      B2C_REQUIRE(grid2.y <= 65535, B2C_E_UNSUPPORTED, "b2c_dense_apply_prepared: ncols=%lld too large for one launch", (long long)ncols);
      B2C_CUDA((set_max_smem<dense_tc256_kernel>(T2_SMEM)));
      dense_tc256_kernel<<<grid2, T2_THREADS, T2_SMEM, (cudaStream_t)stream>>>(P, k, reinterpret_cast<const float2 *>(in), out, ncols, ld_in);
      B2C_CUDA(cudaGetLastError());
      return B2C_OK;
    }
    dim3 grid((unsigned)tiles_m, (unsigned)((ncols + TC_BN - 1) / TC_BN));
    B2C_REQUIRE(grid.y <= 65535, B2C_E_UNSUPPORTED, "b2c_dense_apply_prepared: ncols=%lld too large for one launch", (long long)ncols);
    B2C_CUDA((set_max_smem<dense_tc_kernel<false, true>>(TC_SMEM)));
    dense_tc_kernel<false, true><<<grid, TC_THREADS, TC_SMEM, (cudaStream_t)stream>>>(P, k, k, reinterpret_cast<const float2 *>(in), out,
                                                                                       ncols, ld_in, ld_out);
  } else {
    if (ncols >= T2_BN / 2) {     // warp-specialised form: a tile covers 128 complex columns
      dim3 grid2((unsigned)tiles_m, (unsigned)((ncols + T2_BN / 2 - 1) / (T2_BN / 2)));
      B2C_REQUIRE(grid2.y <= 65535, B2C_E_UNSUPPORTED, "b2c_dense_apply_prepared: ncols=%lld too large for one launch", (long long)ncols);
      B2C_CUDA((set_max_smem<dense_tc_ws_kernel<true>>(WS_SMEM)));
      dense_tc_ws_kernel<true><<<grid2, WS_THREADS, WS_SMEM, (cudaStream_t)stream>>>(P, m, k, reinterpret_cast<const float2 *>(in), out, ncols,
                                                                                     ld_in, ld_out, GroupTable{});
      B2C_CUDA(cudaGetLastError());
      return B2C_OK;
    }
    dim3 grid((unsigned)tiles_m, (unsigned)((ncols + TC_BN / 2 - 1) / (TC_BN / 2)));
    B2C_REQUIRE(grid.y <= 65535, B2C_E_UNSUPPORTED, "b2c_dense_apply_prepared: ncols=%lld too large for one launch", (long long)ncols);
    B2C_CUDA((set_max_smem<dense_tc_kernel<true, true>>(TC_SMEM)));
    dense_tc_kernel<true, true><<<grid, TC_THREADS, TC_SMEM, (cudaStream_t)stream>>>(P, m, k, reinterpret_cast<const float2 *>(in), out,
                                                                                      ncols, ld_in, ld_out);
  }
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

extern "C" int b2c_dense_apply_grouped(const b2c_dense_group *groups_host, int32_t ngroups, const float *in, float *out, int64_t ld,
                                       void *stream) {
  B2C_REQUIRE(groups_host && in && out && in != out, B2C_E_ARG, "b2c_dense_apply_grouped: null argument or in-place");
  B2C_REQUIRE(ngroups >= 0 && ngroups <= MAX_GROUPS, B2C_E_UNSUPPORTED, "b2c_dense_apply_grouped: %d groups (at most %d per call)", ngroups,
              MAX_GROUPS);
  GroupTable gt = {};
  int tiles = 0;
  bool ta = (ld & 1) == 0 && ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0 && !getenv("B2C_DENSE_SS");
  for (int g = 0; g < ngroups; ++g) ta = ta && ((groups_host[g].col0 * ld) & 1) == 0;
  // Cluster-of-two form (W tiles multicast into both CTAs' rings): measured 221-223 TFLOP/s useful against 228-229 for
  // independent CTAs on 32768 columns -- it takes L2 throughput from 39 % to 32 % but the kernel is not L2-bound (it runs
  // at 0.94 of the TF32 rate cuBLAS reaches on the same box), so it is opt-in: B2C_DENSE_CLUSTER=1.
  const bool cl2 = ta && getenv("B2C_DENSE_CLUSTER") != nullptr;
  for (int g = 0; g < ngroups; ++g) {
    const b2c_dense_group &q = groups_host[g];
    B2C_REQUIRE(q.prepared && q.np >= 1 && q.ncols >= 0 && q.col0 >= 0 && ld >= q.np, B2C_E_ARG,
                "b2c_dense_apply_grouped: group %d: np=%d col0=%lld ncols=%lld ld=%lld", g, q.np, (long long)q.col0, (long long)q.ncols,
                (long long)ld);
    B2C_REQUIRE(((uintptr_t)q.prepared & 15) == 0, B2C_E_ARG, "b2c_dense_apply_grouped: prepared operand must be 16-byte aligned");
    if (q.ncols == 0) continue;                    // empty groups take no tiles
    int tm, nst;
    prep_dims(q.np, q.np, 1, tm, nst);
    if (ta) tm = (tm + 1) / 2;                     // TS kernel: a CTA takes a PAIR of W row tiles x 128 columns
    int64_t tn = ta ? (q.ncols + TA_BM - 1) / TA_BM : (q.ncols + T2_BN - 1) / T2_BN;
    if (ta && cl2) tn = (tn + 1) & ~1ll;           // clusters of two adjacent column tiles (a padding tile computes nothing)
    B2C_REQUIRE(tiles + tm * tn < (1ll << 30), B2C_E_UNSUPPORTED, "b2c_dense_apply_grouped: too many tiles");
    const int i = gt.ngroups++;
    tiles += (int)(tm * tn);
    gt.tile_end[i] = tiles;
    gt.tiles_m[i] = tm;
    gt.np[i] = q.np;
    gt.col0[i] = q.col0;
    gt.ncols[i] = q.ncols;
    gt.prepared[i] = static_cast<const float *>(q.prepared);
  }
  if (tiles == 0) return B2C_OK;
  if (ta && cl2) {
    B2C_CUDA((set_max_smem<dense_tc_ta_kernel<true, true>>(TA_SMEM)));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)tiles);
    cfg.blockDim = dim3(TA_THREADS);
    cfg.dynamicSmemBytes = TA_SMEM;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B2C_CUDA(cudaLaunchKernelEx(&cfg, dense_tc_ta_kernel<true, true>, (const float *)nullptr, 0, reinterpret_cast<const float2 *>(in), out,
                                (int64_t)0, ld, gt));
    return B2C_OK;
  }
  if (ta) {
    B2C_CUDA((set_max_smem<dense_tc_ta_kernel<true>>(TA_SMEM)));
    dense_tc_ta_kernel<true><<<(unsigned)tiles, TA_THREADS, TA_SMEM, (cudaStream_t)stream>>>(nullptr, 0, reinterpret_cast<const float2 *>(in), out,
                                                                                            0, ld, gt);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
  }
  B2C_CUDA((set_max_smem<dense_tc_ws_kernel<false, true>>(WS_SMEM)));
  dense_tc_ws_kernel<false, true><<<(unsigned)tiles, WS_THREADS, WS_SMEM, (cudaStream_t)stream>>>(
      nullptr, 0, 0, reinterpret_cast<const float2 *>(in), out, 0, ld, ld, gt);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}
