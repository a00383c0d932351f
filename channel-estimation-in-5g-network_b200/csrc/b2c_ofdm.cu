// K2: OFDM modulate / demodulate.  1024-point radix-4 Stockham FFT in shared memory fused with the
// subcarrier map, (i)fftshift, 1/sqrt(N) scaling and cyclic-prefix insert / strip.
// Replaces OFDMSystem.modulate / demodulate (src/channel_simulator.py:150-203).
#include "b2c_common.cuh"

namespace b2c {

constexpr int FFT_N = 1024;
constexpr int FFT_THREADS = FFT_N / 4;

// Shifted-domain index of used bin k (src/channel_simulator.py:141-148): a contiguous block
// centred on DC with DC itself removed.
__device__ __forceinline__ int used_bin(int k, int nsc) {
  int useful = nsc + 1;
  return FFT_N / 2 - useful / 2 + k + (k >= useful / 2 ? 1 : 0);
}

// One radix-4 Stockham pass.  tw[m] = exp(-j 2 pi m / N); INV conjugates.
template <bool INV>
__device__ __forceinline__ void stockham_pass(const float2 *__restrict__ src, float2 *__restrict__ dst,
                                              const float2 *__restrict__ tw, int Ns) {
  const int j = threadIdx.x;
  const int k = j & (Ns - 1);
  const int step = FFT_N / (4 * Ns);
  float2 v0 = src[j], v1 = src[j + FFT_N / 4], v2 = src[j + FFT_N / 2], v3 = src[j + 3 * FFT_N / 4];
  float2 w1 = tw[(k * step) & (FFT_N - 1)], w2 = tw[(2 * k * step) & (FFT_N - 1)],
         w3 = tw[(3 * k * step) & (FFT_N - 1)];
  if (INV) {
    w1.y = -w1.y;
    w2.y = -w2.y;
    w3.y = -w3.y;
  }
  v1 = cmul(v1, w1);
  v2 = cmul(v2, w2);
  v3 = cmul(v3, w3);
  float2 a = cadd(v0, v2), b = make_float2(v0.x - v2.x, v0.y - v2.y);
  float2 c = cadd(v1, v3), d = make_float2(v1.x - v3.x, v1.y - v3.y);
  // forward: multiply d by -j ; inverse: by +j
  float2 dj = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
  const int j0 = ((j - k) << 2) + k;
  dst[j0] = cadd(a, c);
  dst[j0 + Ns] = cadd(b, dj);
  dst[j0 + 2 * Ns] = make_float2(a.x - c.x, a.y - c.y);
  dst[j0 + 3 * Ns] = make_float2(b.x - dj.x, b.y - dj.y);
}

template <bool INV>
__device__ __forceinline__ float2 *fft1024(float2 *a, float2 *b, const float2 *tw) {
#pragma unroll
  for (int Ns = 1; Ns < FFT_N; Ns <<= 2) {
    __syncthreads();
    stockham_pass<INV>(a, b, tw, Ns);
    float2 *t = a;
    a = b;
    b = t;
  }
  __syncthreads();
  return a;   // buffer holding the result (5 passes: the second buffer)
}

__global__ void __launch_bounds__(FFT_THREADS) ofdm_modulate_kernel(b2c_geom g, const float2 *__restrict__ in,
                                                                    float2 *__restrict__ out, int64_t rows) {
  __shared__ float2 buf0[FFT_N], buf1[FFT_N], tw[FFT_N];
  for (int i = threadIdx.x; i < FFT_N; i += FFT_THREADS) {
    float s, c;
    sincospif(-2.0f * (float)i / (float)FFT_N, &s, &c);
    tw[i] = make_float2(c, s);
  }
  const int nsc = g.nsc, cp = g.cp_length;
  const float scale = rsqrtf((float)FFT_N);   // ifft (1/N) * sqrt(N)
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < FFT_N; i += FFT_THREADS) buf0[i] = make_float2(0.f, 0.f);
    __syncthreads();
    // freq_domain[used] = symbols; ifftshift: natural bin = (shifted + N/2) mod N
    for (int k = threadIdx.x; k < nsc; k += FFT_THREADS)
      buf0[(used_bin(k, nsc) + FFT_N / 2) & (FFT_N - 1)] = __ldg(in + r * nsc + k);
    float2 *t = fft1024<true>(buf0, buf1, tw);
    float2 *o = out + r * (FFT_N + cp);
    for (int i = threadIdx.x; i < FFT_N + cp; i += FFT_THREADS) {
      int n = i < cp ? FFT_N - cp + i : i - cp;   // cyclic prefix = last cp samples
      o[i] = cscale(scale, t[n]);
    }
  }
}

__global__ void __launch_bounds__(FFT_THREADS) ofdm_demodulate_kernel(b2c_geom g, const float2 *__restrict__ in,
                                                                      float2 *__restrict__ out, int64_t rows) {
  __shared__ float2 buf0[FFT_N], buf1[FFT_N], tw[FFT_N];
  for (int i = threadIdx.x; i < FFT_N; i += FFT_THREADS) {
    float s, c;
    sincospif(-2.0f * (float)i / (float)FFT_N, &s, &c);
    tw[i] = make_float2(c, s);
  }
  const int nsc = g.nsc, cp = g.cp_length;
  const float scale = rsqrtf((float)FFT_N);
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    __syncthreads();
    const float2 *x = in + r * (FFT_N + cp) + cp;   // strip the prefix
    for (int i = threadIdx.x; i < FFT_N; i += FFT_THREADS) buf0[i] = __ldg(x + i);
    float2 *t = fft1024<false>(buf0, buf1, tw);
    for (int k = threadIdx.x; k < nsc; k += FFT_THREADS)
      out[r * nsc + k] = cscale(scale, t[(used_bin(k, nsc) + FFT_N / 2) & (FFT_N - 1)]);
  }
}

}  // namespace b2c

using namespace b2c;

static int ofdm_check(const b2c_geom *g, const void *in, void *out, int64_t rows, const char *who) {
  B2C_REQUIRE(g && in && out, B2C_E_ARG, "%s: null argument", who);
  B2C_REQUIRE(g->fft_size == FFT_N, B2C_E_UNSUPPORTED, "%s: fft_size=%d (only 1024 is built)", who, g->fft_size);
  B2C_REQUIRE(g->nsc >= 1 && g->nsc < FFT_N && g->cp_length >= 0 && g->cp_length <= FFT_N, B2C_E_ARG,
              "%s: nsc=%d cp=%d", who, g->nsc, g->cp_length);
  B2C_REQUIRE(rows >= 0, B2C_E_ARG, "%s: rows=%lld", who, (long long)rows);
  return B2C_OK;
}

extern "C" int b2c_ofdm_modulate(const b2c_geom *g, const float *in, float *out, int64_t rows, void *stream) {
  if (int rc = ofdm_check(g, in, out, rows, "b2c_ofdm_modulate")) return rc;
  if (rows == 0) return B2C_OK;
  unsigned grid = (unsigned)(rows < 148 * 8 ? rows : 148 * 8);
  ofdm_modulate_kernel<<<grid, FFT_THREADS, 0, (cudaStream_t)stream>>>(
      *g, reinterpret_cast<const float2 *>(in), reinterpret_cast<float2 *>(out), rows);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

extern "C" int b2c_ofdm_demodulate(const b2c_geom *g, const float *in, float *out, int64_t rows, void *stream) {
  if (int rc = ofdm_check(g, in, out, rows, "b2c_ofdm_demodulate")) return rc;
  if (rows == 0) return B2C_OK;
  unsigned grid = (unsigned)(rows < 148 * 8 ? rows : 148 * 8);
  ofdm_demodulate_kernel<<<grid, FFT_THREADS, 0, (cudaStream_t)stream>>>(
      *g, reinterpret_cast<const float2 *>(in), reinterpret_cast<float2 *>(out), rows);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}
