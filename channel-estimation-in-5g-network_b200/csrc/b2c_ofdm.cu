// K2: OFDM modulate / demodulate.  Replaces OFDMSystem.modulate / demodulate
// (src/channel_simulator.py:150-203): subcarrier map, (i)fftshift, 1024-point (I)FFT with the
// sqrt(N) scaling, cyclic-prefix insert / strip -- fused into one pass over HBM.
//
// One WARP per OFDM symbol row, no block barriers: the 1024-point transform is factored 32 x 32.
// Lane l loads x[32*n1 + l] (coalesced), runs a 32-point radix-2 FFT entirely in registers, applies
// the inter-stage twiddles W_1024^(l*k1) from a CTA-shared table, transposes through a padded
// per-warp shared-memory tile (the only shared-memory round trip), runs the second 32-point FFT in
// registers and stores X[l + 32*k2] (coalesced).  The subcarrier map / shift is index arithmetic
// on the loads (modulate) or stores (demodulate); the cyclic prefix is a second predicated store.
#include "b2c_common.cuh"

namespace b2c {

constexpr int FFT_N = 1024;
constexpr int OFDM_WARPS = 8;
constexpr int OFDM_THREADS = OFDM_WARPS * 32;
constexpr int OFDM_SMEM = (1024 + OFDM_WARPS * 32 * 33) * (int)sizeof(float2);

__device__ constexpr float C32[16] = {1.f, 0.98078528f, 0.923879533f, 0.831469612f, 0.707106781f, 0.555570233f,
                                      0.382683432f, 0.195090322f, 0.f, -0.195090322f, -0.382683432f, -0.555570233f,
                                      -0.707106781f, -0.831469612f, -0.923879533f, -0.98078528f};
__device__ constexpr float S32[16] = {0.f, 0.195090322f, 0.382683432f, 0.555570233f, 0.707106781f, 0.831469612f,
                                      0.923879533f, 0.98078528f, 1.f, 0.98078528f, 0.923879533f, 0.831469612f,
                                      0.707106781f, 0.555570233f, 0.382683432f, 0.195090322f};

__host__ __device__ constexpr int bitrev5(int i) {
  return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// 32-point radix-2 decimation-in-frequency FFT in registers (fully unrolled, compile-time twiddles).
// Output is left in bit-reversed order: natural bin k is x[bitrev5(k)].
template <bool INV>
__device__ __forceinline__ void fft32(float2 (&x)[32]) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
#pragma unroll
    for (int blk = 0; blk < 32; blk += 2 * h) {
#pragma unroll
      for (int i = 0; i < h; ++i) {
        const int tw = i * (16 / h);   // W_{2h}^i = W_32^tw
        const float2 a = x[blk + i], b = x[blk + i + h];
        x[blk + i] = make_float2(a.x + b.x, a.y + b.y);
        const float dx = a.x - b.x, dy = a.y - b.y;
        if (tw == 0) {
          x[blk + i + h] = make_float2(dx, dy);
        } else if (tw == 8) {   // multiply by -j (forward) / +j (inverse)
          x[blk + i + h] = INV ? make_float2(-dy, dx) : make_float2(dy, -dx);
        } else {
          const float c = C32[tw], s = INV ? S32[tw] : -S32[tw];
          x[blk + i + h] = make_float2(fmaf(dx, c, -dy * s), fmaf(dx, s, dy * c));
        }
      }
    }
  }
}

// 1024-point transform of the 32 values each lane holds (x[n1] = element 32*n1 + lane).  On return
// lane k1 holds, in bit-reversed register order, X[k1 + 32*k2] = x[bitrev5(k2)].
template <bool INV>
__device__ __forceinline__ void fft1024_warp(float2 (&x)[32], float2 (*tile)[33], const float2 (*twid)[32], int lane) {
  fft32<INV>(x);
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) {
    float2 v = x[bitrev5(k1)];
    if (k1 != 0) {
      float2 w = twid[k1][lane];        // exp(-j 2 pi lane k1 / 1024); conjugate for the inverse
      if (INV) w.y = -w.y;
      v = cmul(v, w);
    }
    tile[k1][lane] = v;
  }
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) x[n2] = tile[lane][n2];
  __syncwarp();
  fft32<INV>(x);
}

// used bin index k of natural FFT bin j (or -1): the occupied block is centred on DC with DC removed
// (src/channel_simulator.py:141-148) and (i)fftshift moves shifted index i to natural bin (i + N/2) mod N.
__device__ __forceinline__ int used_of_bin(int j, int half) {
  if (j >= 1 && j < half) return j + half - 1;
  if (j >= FFT_N - half) return j - (FFT_N - half);
  return -1;
}

template <bool MOD>
__global__ void __launch_bounds__(OFDM_THREADS, 2) ofdm_kernel(b2c_geom g, const float2 *__restrict__ in,
                                                               float2 *__restrict__ out, int64_t rows) {
  extern __shared__ __align__(16) float2 ofdm_smem[];
  float2(*twid)[32] = reinterpret_cast<float2(*)[32]>(ofdm_smem);                    // [32][32]
  float2(*tiles)[32][33] = reinterpret_cast<float2(*)[32][33]>(ofdm_smem + 1024);    // [warps][32][33]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 1024; i += OFDM_THREADS) {
    float s, c;
    sincospif(-2.0f * (float)((i >> 5) * (i & 31)) / (float)FFT_N, &s, &c);
    twid[i >> 5][i & 31] = make_float2(c, s);
  }
  __syncthreads();
  const int nsc = g.nsc, cp = g.cp_length, half = (g.nsc + 1) / 2;
  const float scale = rsqrtf((float)FFT_N);
  const int64_t stride_t = FFT_N + cp;
  for (int64_t r = (int64_t)blockIdx.x * OFDM_WARPS + warp; r < rows; r += (int64_t)gridDim.x * OFDM_WARPS) {
    float2 x[32];
    if (MOD) {
      const float2 *src = in + r * nsc;
#pragma unroll
      for (int n1 = 0; n1 < 32; ++n1) {
        const int k = used_of_bin(32 * n1 + lane, half);
        x[n1] = k >= 0 ? __ldg(src + k) : make_float2(0.f, 0.f);
      }
      fft1024_warp<true>(x, tiles[warp], twid, lane);
      float2 *dst = out + r * stride_t;
#pragma unroll
      for (int k2 = 0; k2 < 32; ++k2) {
        const int n = lane + 32 * k2;
        const float2 v = cscale(scale, x[bitrev5(k2)]);      // ifft * sqrt(N) = sum / sqrt(N)
        st_stream(dst + cp + n, v);
        if (n >= FFT_N - cp) st_stream(dst + n - (FFT_N - cp), v);   // cyclic prefix = last cp samples
      }
    } else {
      const float2 *src = in + r * stride_t + cp;               // strip the prefix
#pragma unroll
      for (int n1 = 0; n1 < 32; ++n1) x[n1] = __ldg(src + 32 * n1 + lane);
      fft1024_warp<false>(x, tiles[warp], twid, lane);
      float2 *dst = out + r * nsc;
#pragma unroll
      for (int k2 = 0; k2 < 32; ++k2) {
        const int k = used_of_bin(lane + 32 * k2, half);
        if (k >= 0) st_stream(dst + k, cscale(scale, x[bitrev5(k2)]));
      }
    }
  }
}

}  // namespace b2c

using namespace b2c;

static int ofdm_check(const b2c_geom *g, const void *in, void *out, int64_t rows, const char *who) {
  B2C_REQUIRE(g && in && out, B2C_E_ARG, "%s: null argument", who);
  B2C_REQUIRE(g->fft_size == FFT_N, B2C_E_UNSUPPORTED, "%s: fft_size=%d (only 1024 is built)", who, g->fft_size);
  B2C_REQUIRE(g->nsc >= 1 && g->nsc < FFT_N && (g->nsc & 1) == 1 && g->cp_length >= 0 && g->cp_length <= FFT_N, B2C_E_ARG,
              "%s: nsc=%d (odd, < 1024) cp=%d", who, g->nsc, g->cp_length);
  B2C_REQUIRE(rows >= 0, B2C_E_ARG, "%s: rows=%lld", who, (long long)rows);
  return B2C_OK;
}

static unsigned ofdm_grid(int64_t rows) {
  int64_t ctas = (rows + OFDM_WARPS - 1) / OFDM_WARPS;
  return (unsigned)(ctas < 148 * 2 ? ctas : 148 * 2);   // persistent: 2 CTAs per SM, grid-stride over rows
}

extern "C" int b2c_ofdm_modulate(const b2c_geom *g, const float *in, float *out, int64_t rows, void *stream) {
  if (int rc = ofdm_check(g, in, out, rows, "b2c_ofdm_modulate")) return rc;
  if (rows == 0) return B2C_OK;
  B2C_CUDA((set_max_smem<ofdm_kernel<true>>(OFDM_SMEM)));
  ofdm_kernel<true><<<ofdm_grid(rows), OFDM_THREADS, OFDM_SMEM, (cudaStream_t)stream>>>(
      *g, reinterpret_cast<const float2 *>(in), reinterpret_cast<float2 *>(out), rows);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

extern "C" int b2c_ofdm_demodulate(const b2c_geom *g, const float *in, float *out, int64_t rows, void *stream) {
  if (int rc = ofdm_check(g, in, out, rows, "b2c_ofdm_demodulate")) return rc;
  if (rows == 0) return B2C_OK;
  B2C_CUDA((set_max_smem<ofdm_kernel<false>>(OFDM_SMEM)));
  ofdm_kernel<false><<<ofdm_grid(rows), OFDM_THREADS, OFDM_SMEM, (cudaStream_t)stream>>>(
      *g, reinterpret_cast<const float2 *>(in), reinterpret_cast<float2 *>(out), rows);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}
