// Link-level helpers either side of the estimation path:
//   b2c_equalize          -- equalize_channel, per-RE ZF / MMSE solve   (src/baseline_estimators.py:273-312)
//   b2c_qam_modulate      -- qam_modulation, QPSK / 16-QAM              (src/utils.py:71-108)
//   b2c_qam_demodulate    -- qam_demodulation, minimum distance         (src/utils.py:111-152)
//   b2c_count_bit_errors  -- calculate_ber numerator                    (src/utils.py:155-157)
//   b2c_pair00_moments    -- ChannelDataset._compute_normalization_stats (src/train.py:41-57)
//   b2c_ml_features       -- prepare_ml_inputs (src/dataset_generator.py:183-227) and
//                            ChannelDataset.__getitem__ (src/train.py:62-94): 5-channel real features
// All HBM-bound element-wise / small-reduction work; none of it is on the bench's timed path.
#include "b2c_common.cuh"

namespace b2c {

// The link-level entries work on any grid width (no bin pairing, no Philox lanes): only bound the shape.
static int check_dims(const b2c_geom *g) {
  B2C_REQUIRE(g && g->nsym >= 1 && g->nsc >= 1 && (int64_t)g->nsym * g->nsc < (1 << 28), B2C_E_UNSUPPORTED,
              "grid %dx%d unsupported", g ? g->nsym : 0, g ? g->nsc : 0);
  B2C_REQUIRE(g->ntx >= 1 && g->ntx <= B2C_MAX_ANT && g->nrx >= 1 && g->nrx <= B2C_MAX_ANT, B2C_E_UNSUPPORTED,
              "antenna counts %dx%d outside [1,%d]", g->ntx, g->nrx, B2C_MAX_ANT);
  return B2C_OK;
}

// ---- equalize_channel ------------------------------------------------------------------------------
// x = (H^H H + lambda I)^-1 H^H y per resource element.  The reference regularises ZF with 1e-8 and
// feeds tx-replicated (rank-1) estimates, so H^H H + lambda I reaches condition numbers ~1e9: the
// normal equations are accumulated and solved in fp64 (Cholesky, Hermitian positive definite)
// whatever the I/O type.
struct cd {
  double x, y;
};
__device__ __forceinline__ cd cd_ld(const float2 *p) {
  float2 v = __ldg(p);
  return cd{(double)v.x, (double)v.y};
}
__device__ __forceinline__ cd cd_ld(const double2 *p) {
  double2 v = __ldg(p);
  return cd{v.x, v.y};
}
__device__ __forceinline__ void cd_st(float2 *p, cd v) { *p = make_float2((float)v.x, (float)v.y); }
__device__ __forceinline__ void cd_st(double2 *p, cd v) { *p = make_double2(v.x, v.y); }
// a += conj(p) * q
__device__ __forceinline__ void cd_mac_conj(cd &a, cd p, cd q) {
  a.x = fma(p.x, q.x, fma(p.y, q.y, a.x));
  a.y = fma(p.x, q.y, fma(-p.y, q.x, a.y));
}
// a -= p * conj(q)
__device__ __forceinline__ void cd_msub_pcq(cd &a, cd p, cd q) {
  a.x -= fma(p.x, q.x, p.y * q.y);
  a.y -= fma(p.y, q.x, -p.x * q.y);
}
// a -= p * q
__device__ __forceinline__ void cd_msub(cd &a, cd p, cd q) {
  a.x -= fma(p.x, q.x, -p.y * q.y);
  a.y -= fma(p.x, q.y, p.y * q.x);
}

template <int NTX, typename CT>
__global__ void __launch_bounds__(128) equalize_kernel(int nrx, int nsc, int64_t total, const CT *__restrict__ rx,
                                                       const CT *__restrict__ H, CT *__restrict__ out, double lambda) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int64_t bs = gid / nsc;              // slot * nsym + symbol
  const int k = (int)(gid - bs * nsc);
  const CT *Hp = H + bs * nrx * NTX * (int64_t)nsc + k;
  const CT *yp = rx + bs * nrx * (int64_t)nsc + k;

  cd A[NTX][NTX], b[NTX];                    // lower triangle of H^H H, and H^H y
#pragma unroll
  for (int i = 0; i < NTX; ++i) {
    b[i] = cd{0.0, 0.0};
#pragma unroll
    for (int j = 0; j < NTX; ++j) A[i][j] = cd{0.0, 0.0};
  }
  for (int r = 0; r < nrx; ++r) {
    cd h[NTX];
#pragma unroll
    for (int t = 0; t < NTX; ++t) h[t] = cd_ld(Hp + (int64_t)(r * NTX + t) * nsc);
    const cd y = cd_ld(yp + (int64_t)r * nsc);
#pragma unroll
    for (int i = 0; i < NTX; ++i) {
      cd_mac_conj(b[i], h[i], y);
#pragma unroll
      for (int j = 0; j <= i; ++j) cd_mac_conj(A[i][j], h[i], h[j]);
    }
  }
  // in-place Cholesky A = L L^H (lower), then L z = b, L^H x = z
#pragma unroll
  for (int j = 0; j < NTX; ++j) {
    double d2 = A[j][j].x + lambda;
#pragma unroll
    for (int p = 0; p < j; ++p) d2 -= fma(A[j][p].x, A[j][p].x, A[j][p].y * A[j][p].y);
    const double d = sqrt(fmax(d2, 1e-300));
    const double inv = 1.0 / d;
    A[j][j] = cd{d, inv};                    // .y of the diagonal carries 1/d
#pragma unroll
    for (int i = j + 1; i < NTX; ++i) {
      cd v = A[i][j];
#pragma unroll
      for (int p = 0; p < j; ++p) cd_msub_pcq(v, A[i][p], A[j][p]);
      A[i][j] = cd{v.x * inv, v.y * inv};
    }
  }
#pragma unroll
  for (int i = 0; i < NTX; ++i) {
    cd v = b[i];
#pragma unroll
    for (int p = 0; p < i; ++p) cd_msub(v, A[i][p], b[p]);
    b[i] = cd{v.x * A[i][i].y, v.y * A[i][i].y};
  }
#pragma unroll
  for (int i = NTX - 1; i >= 0; --i) {
    cd v = b[i];
#pragma unroll
    for (int p = i + 1; p < NTX; ++p) cd_msub(v, cd{A[p][i].x, -A[p][i].y}, b[p]);
    b[i] = cd{v.x * A[i][i].y, v.y * A[i][i].y};
  }
  CT *op = out + bs * NTX * (int64_t)nsc + k;
#pragma unroll
  for (int t = 0; t < NTX; ++t) cd_st(op + (int64_t)t * nsc, b[t]);
}

template <typename CT>
static int launch_equalize(const b2c_geom *g, int64_t B, const void *rx, const void *H, void *out, double lambda,
                           cudaStream_t st) {
  const int64_t total = B * g->nsym * (int64_t)g->nsc;
  const unsigned grid = (unsigned)((total + 127) / 128);
  const CT *r = static_cast<const CT *>(rx), *h = static_cast<const CT *>(H);
  CT *o = static_cast<CT *>(out);
#define B2C_EQ(N)                                                                         \
  case N:                                                                                 \
    equalize_kernel<N, CT><<<grid, 128, 0, st>>>(g->nrx, g->nsc, total, r, h, o, lambda); \
    break;
  switch (g->ntx) {
    B2C_EQ(1) B2C_EQ(2) B2C_EQ(3) B2C_EQ(4) B2C_EQ(5) B2C_EQ(6) B2C_EQ(7) B2C_EQ(8)
    default:
      set_error("b2c_equalize: ntx=%d", g->ntx);
      return B2C_E_UNSUPPORTED;
  }
#undef B2C_EQ
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

// ---- QAM -------------------------------------------------------------------------------------------
// Constellation point `idx` as the reference lists it (src/utils.py:91-103):
//   QPSK   : (+1+1j, -1+1j, +1-1j, -1-1j) / sqrt(2)
//   16-QAM : idx = 4a + b, re = L[a], im = L[b], L = (-3, -1, +3, +1), / sqrt(10)
// and the position map gray[decimal] (:93, :102) applied between bit groups (MSB first) and points.
__device__ __forceinline__ double2 qam_point(int M, int idx) {
  if (M == 4) {
    const double a = 0.70710678118654752440;
    return make_double2((idx & 1) ? -a : a, (idx & 2) ? -a : a);
  }
  const double s = 0.31622776601683793320;   // 1/sqrt(10)
  const int a = idx >> 2, b = idx & 3;
  const double lr = (a == 0) ? -3.0 : (a == 1) ? -1.0 : (a == 2) ? 3.0 : 1.0;
  const double li = (b == 0) ? -3.0 : (b == 1) ? -1.0 : (b == 2) ? 3.0 : 1.0;
  return make_double2(lr * s, li * s);
}
// gray[] of the reference; both tables are involutions, so argsort(gray) == gray (:145-146)
__device__ __forceinline__ int qam_gray(int M, int d) {
  if (M == 4) return d ^ (d >> 1);
  // [0,1,3,2, 4,5,7,6, 12,13,15,14, 8,9,11,10]: Gray code of the high and of the low bit pair
  const int hi = d >> 2, lo = d & 3;
  return ((hi ^ (hi >> 1)) << 2) | (lo ^ (lo >> 1));
}

__global__ void __launch_bounds__(256) qam_mod_kernel(const uint8_t *__restrict__ bits, int64_t n, int M, int bps,
                                                      float2 *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int d = 0;
  for (int q = 0; q < bps; ++q) d = (d << 1) | (bits[i * bps + q] & 1);
  const double2 p = qam_point(M, qam_gray(M, d));
  out[i] = make_float2((float)p.x, (float)p.y);
}

template <typename CT>
__global__ void __launch_bounds__(256) qam_demod_kernel(const CT *__restrict__ sym, int64_t n, int M, int bps,
                                                        uint8_t *__restrict__ bits) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const CT v = sym[i];
  const double vx = v.x, vy = v.y;
  int best = 0;
  double dbest = 1e300;
  for (int c = 0; c < M; ++c) {              // first minimum wins, as np.argmin (:141-142)
    const double2 p = qam_point(M, c);
    const double dx = vx - p.x, dy = vy - p.y, d = fma(dx, dx, dy * dy);
    if (d < dbest) {
      dbest = d;
      best = c;
    }
  }
  const int dec = qam_gray(M, best);
  for (int q = 0; q < bps; ++q) bits[i * bps + q] = (uint8_t)((dec >> (bps - 1 - q)) & 1);
}

__global__ void __launch_bounds__(256) bit_errors_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b,
                                                         int64_t n, unsigned long long *__restrict__ count) {
  unsigned long long c = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    c += a[i] != b[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

// Bit errors per slot on the data REs: one CTA per slot, the pilot set of the slot's pattern as a bitmap in shared memory.
__global__ void __launch_bounds__(256) bit_errors_slot_kernel(b2c_geom g, b2c_patterns pat, const int32_t *__restrict__ pattern_id,
                                                              const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, int bps,
                                                              uint32_t *__restrict__ counts) {
  extern __shared__ uint32_t pilot_bits[];                 // [(nre + 31) / 32]
  __shared__ uint32_t wsum[8];
  const int nre = g.nsym * g.nsc, nwords = (nre + 31) >> 5;
  const int64_t s = blockIdx.x;
  const int pid = pattern_id[s];
  for (int i = threadIdx.x; i < nwords; i += 256) pilot_bits[i] = 0u;
  __syncthreads();
  const int np = pat.npilots[pid];
  const int *pre = pat.pilot_re + (int64_t)pid * pat.np_max;
  for (int j = threadIdx.x; j < np; j += 256) {
    const int e = __ldg(pre + j);
    atomicOr(&pilot_bits[e >> 5], 1u << (e & 31));
  }
  __syncthreads();
  const uint8_t *pa = a + s * (int64_t)nre * bps, *pb = b + s * (int64_t)nre * bps;
  uint32_t c = 0;
  for (int i = threadIdx.x; i < nre * bps; i += 256) {
    const int e = i / bps;
    if (!((pilot_bits[e >> 5] >> (e & 31)) & 1u)) c += pa[i] != pb[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < 8; ++w) t += wsum[w];
    counts[s] = t;
  }
}

// counts[0] += #NaN, counts[1] += #Inf among n elements (verify_phase3_datasets.py:98-111).  CPLX: elements are
// complex64 and one counts as NaN / Inf when either part is, as numpy.isnan / numpy.isinf do.
template <bool CPLX>
__global__ void __launch_bounds__(256) nonfinite_kernel(const float *__restrict__ x, int64_t n, unsigned long long *__restrict__ counts) {
  unsigned long long nn = 0, ni = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (CPLX) {
      const float2 v = __ldg(reinterpret_cast<const float2 *>(x) + i);
      nn += (isnan(v.x) || isnan(v.y)) ? 1 : 0;
      ni += (isinf(v.x) || isinf(v.y)) ? 1 : 0;
    } else {
      const float v = __ldg(x + i);
      nn += isnan(v) ? 1 : 0;
      ni += isinf(v) ? 1 : 0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nn += __shfl_xor_sync(0xffffffffu, nn, o);
    ni += __shfl_xor_sync(0xffffffffu, ni, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (nn) atomicAdd(counts, nn);
    if (ni) atomicAdd(counts + 1, ni);
  }
}

__device__ __forceinline__ double block_sum_d(double v, double *scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = lane < nwarp ? scratch[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

// sum[0] += sum_i |a_i - b_i| over n complex64 elements (compute_mae, run_phase5_evaluation.py:51-54)
__global__ void __launch_bounds__(256) abs_diff_kernel(const float2 *__restrict__ a, const float2 *__restrict__ b, int64_t n,
                                                       double *__restrict__ sum) {
  __shared__ double red[33];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 x = __ldg(a + i), y = __ldg(b + i);
    const float dx = x.x - y.x, dy = x.y - y.y;
    acc += (double)sqrtf(fmaf(dx, dx, dy * dy));
  }
  const double tot = block_sum_d(acc, red);
  if (threadIdx.x == 0) atomicAdd(sum, tot);
}

// ---- ML feature packing -----------------------------------------------------------------------------
struct PairView {
  const float2 *rx, *ls, *tr;          // antenna pair (0,0) rows of slot b: base + s * stride
  int rx_stride, ls_stride, tr_stride;
};
__device__ __forceinline__ PairView pair_view(const b2c_geom &g, int64_t b, const float2 *rx, const float2 *ls,
                                              const float2 *tr, int ls_sym_stride) {
  PairView v;
  const int P = g.pitch ? g.pitch : g.nsc;       // padded rows (b2c_geom.pitch) or contiguous
  v.rx_stride = g.nrx * P;
  v.ls_stride = ls_sym_stride;
  v.tr_stride = g.nrx * g.ntx * P;
  v.rx = rx + b * g.nsym * (int64_t)v.rx_stride;
  v.ls = ls + b * g.nsym * (int64_t)v.ls_stride;
  v.tr = tr + b * g.nsym * (int64_t)v.tr_stride;
  return v;
}


// Sum re, sum im, sum re^2, sum im^2 of the pair-(0,0) rows of rx, H_ls, H_true, accumulated into mom[3][4].
__global__ void __launch_bounds__(256) pair00_moments_kernel(b2c_geom g, const float2 *__restrict__ rx,
                                                             const float2 *__restrict__ ls, const float2 *__restrict__ tr,
                                                             int ls_sym_stride, double *__restrict__ mom) {
  __shared__ double red[33];
  const PairView v = pair_view(g, blockIdx.x, rx, ls, tr, ls_sym_stride);
  const int nre = g.nsym * g.nsc;
  double acc[3][4] = {};
  for (int e = threadIdx.x; e < nre; e += blockDim.x) {
    const int s = e / g.nsc, k = e - s * g.nsc;
    const float2 a[3] = {__ldg(v.rx + s * v.rx_stride + k), __ldg(v.ls + s * v.ls_stride + k),
                         __ldg(v.tr + s * v.tr_stride + k)};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      acc[q][0] += a[q].x;
      acc[q][1] += a[q].y;
      acc[q][2] += (double)a[q].x * a[q].x;
      acc[q][3] += (double)a[q].y * a[q].y;
    }
  }
#pragma unroll
  for (int q = 0; q < 3; ++q)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double t = block_sum_d(acc[q][c], red);
      if (threadIdx.x == 0) atomicAdd(mom + q * 4 + c, t);
    }
}

// Per-slot error sums of the pair-(0,0) rows: out[b] = {sum |L - H|^2, sum |alpha_b L - H|^2, sum |H|^2}
// (run_phase5_evaluation.py:283-296: LS and the alpha-scaled "MMSE" baseline, NMSE per sample).
__global__ void __launch_bounds__(256) pair00_errors_kernel(b2c_geom g, const float2 *__restrict__ ls,
                                                            const float2 *__restrict__ tr, int ls_sym_stride,
                                                            const float *__restrict__ alpha, double *__restrict__ out) {
  __shared__ double red[33];
  const int64_t b = blockIdx.x;
  const PairView v = pair_view(g, b, tr, ls, tr, ls_sym_stride);
  const int nre = g.nsym * g.nsc;
  const float a = alpha ? alpha[b] : 1.0f;
  double e0 = 0, e1 = 0, pw = 0;
  for (int e = threadIdx.x; e < nre; e += blockDim.x) {
    const int s = e / g.nsc, k = e - s * g.nsc;
    const float2 l = __ldg(v.ls + s * v.ls_stride + k), h = __ldg(v.tr + s * v.tr_stride + k);
    const float dx = l.x - h.x, dy = l.y - h.y, mx = fmaf(a, l.x, -h.x), my = fmaf(a, l.y, -h.y);
    e0 += (double)fmaf(dx, dx, dy * dy);
    e1 += (double)fmaf(mx, mx, my * my);
    pw += (double)fmaf(h.x, h.x, h.y * h.y);
  }
  e0 = block_sum_d(e0, red);
  e1 = block_sum_d(e1, red);
  pw = block_sum_d(pw, red);
  if (threadIdx.x == 0) out[b * 3 + 0] = e0, out[b * 3 + 1] = e1, out[b * 3 + 2] = pw;
}

// One CTA per slot.  norm_mode 0: raw; 1: prepare_ml_inputs' per-sample scaling (inputs[..., :4] by
// 1/(std+1e-8) over the four real channels jointly, targets by 1/(std+1e-8), :219-223); 2: the affine
// (v - mean) * scale of ChannelDataset.__getitem__ with norm = {rx_mean, rx_scale, ls_mean, ls_scale,
// true_mean, true_scale}.  layout 0: channel-last [nsym][nsc][5] / [nsym][nsc][2]; 1: channel-first
// [5][nsym][nsc] / [2][nsym][nsc].  Channel 4 is the pilot mask of the slot's pattern.
__global__ void __launch_bounds__(256) ml_features_kernel(b2c_geom g, b2c_patterns pat, const int32_t *__restrict__ pattern_id,
                                                          const float2 *__restrict__ rx, const float2 *__restrict__ ls,
                                                          const float2 *__restrict__ tr, int ls_sym_stride, int layout,
                                                          int norm_mode, const float *__restrict__ norm,
                                                          float *__restrict__ inputs, float *__restrict__ targets) {
  __shared__ double red[33];
  const int64_t b = blockIdx.x;
  const PairView v = pair_view(g, b, rx, ls, tr, ls_sym_stride);
  const int nre = g.nsym * g.nsc;
  float in_mean[2] = {0.f, 0.f}, in_scale[2] = {1.f, 1.f}, t_mean = 0.f, t_scale = 1.f;
  if (norm_mode == 1) {
    double s_in = 0, q_in = 0, s_t = 0, q_t = 0;
    for (int e = threadIdx.x; e < nre; e += blockDim.x) {
      const int s = e / g.nsc, k = e - s * g.nsc;
      const float2 a = __ldg(v.rx + s * v.rx_stride + k), c = __ldg(v.ls + s * v.ls_stride + k),
                   t = __ldg(v.tr + s * v.tr_stride + k);
      s_in += (double)a.x + a.y + c.x + c.y;
      q_in += (double)a.x * a.x + (double)a.y * a.y + (double)c.x * c.x + (double)c.y * c.y;
      s_t += (double)t.x + t.y;
      q_t += (double)t.x * t.x + (double)t.y * t.y;
    }
    s_in = block_sum_d(s_in, red);
    q_in = block_sum_d(q_in, red);
    s_t = block_sum_d(s_t, red);
    q_t = block_sum_d(q_t, red);
    const double n_in = 4.0 * nre, n_t = 2.0 * nre;
    const double sd_in = sqrt(fmax(q_in / n_in - (s_in / n_in) * (s_in / n_in), 0.0));
    const double sd_t = sqrt(fmax(q_t / n_t - (s_t / n_t) * (s_t / n_t), 0.0));
    in_scale[0] = in_scale[1] = (float)(1.0 / (sd_in + 1e-8));
    t_scale = (float)(1.0 / (sd_t + 1e-8));
  } else if (norm_mode == 2) {
    in_mean[0] = norm[0], in_scale[0] = norm[1], in_mean[1] = norm[2], in_scale[1] = norm[3];
    t_mean = norm[4], t_scale = norm[5];
  }
  float *ip = inputs + b * 5 * (int64_t)nre, *tp = targets + b * 2 * (int64_t)nre;
  for (int f = threadIdx.x; f < 5 * nre; f += blockDim.x) {
    const int e = layout ? f % nre : f / 5, c = layout ? f / nre : f - 5 * e;
    float val = 0.f;
    if (c < 4) {
      const int s = e / g.nsc, k = e - s * g.nsc;
      const float2 a = (c < 2) ? __ldg(v.rx + s * v.rx_stride + k) : __ldg(v.ls + s * v.ls_stride + k);
      val = (((c & 1) ? a.y : a.x) - (c < 2 ? in_mean[0] : in_mean[1])) * (c < 2 ? in_scale[0] : in_scale[1]);
    }
    ip[f] = val;
  }
  for (int f = threadIdx.x; f < 2 * nre; f += blockDim.x) {
    const int e = layout ? f % nre : f >> 1, c = layout ? f / nre : f & 1;
    const int s = e / g.nsc, k = e - s * g.nsc;
    const float2 a = __ldg(v.tr + s * v.tr_stride + k);
    tp[f] = ((c ? a.y : a.x) - t_mean) * t_scale;
  }
  __syncthreads();                           // mask zeros above are visible before the ones land
  const int p = pattern_id[b], np_ = pat.npilots[p];
  const int32_t *re = pat.pilot_re + (int64_t)p * pat.np_max;
  for (int j = threadIdx.x; j < np_; j += blockDim.x) {
    const int e = re[j];
    ip[layout ? 4 * nre + e : 5 * e + 4] = 1.0f;
  }
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_equalize(const b2c_geom *g, int64_t B, const void *rx, const void *H, void *out, double lambda,
                            int32_t fp64_io, void *stream) {
  int rc = b2c::check_dims(g);
  if (rc) return rc;
  B2C_REQUIRE(rx && H && out && B >= 0, B2C_E_ARG, "b2c_equalize: null argument or B < 0");
  B2C_REQUIRE(lambda >= 0.0, B2C_E_ARG, "b2c_equalize: lambda=%g", lambda);
  if (B == 0) return B2C_OK;
  return fp64_io ? launch_equalize<double2>(g, B, rx, H, out, lambda, (cudaStream_t)stream)
                 : launch_equalize<float2>(g, B, rx, H, out, lambda, (cudaStream_t)stream);
}

static int qam_bits(int32_t M) { return M == 4 ? 2 : M == 16 ? 4 : 0; }

extern "C" int b2c_qam_modulate(const uint8_t *bits, int64_t nsymbols, int32_t M, float *out, void *stream) {
  const int bps = qam_bits(M);
  B2C_REQUIRE(bps, B2C_E_UNSUPPORTED, "Modulation order %d not implemented", M);
  B2C_REQUIRE(bits && out && nsymbols >= 0, B2C_E_ARG, "b2c_qam_modulate: null argument");
  if (nsymbols == 0) return B2C_OK;
  qam_mod_kernel<<<(unsigned)((nsymbols + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bits, nsymbols, M, bps,
                                                                                        reinterpret_cast<float2 *>(out));
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

extern "C" int b2c_qam_demodulate(const void *symbols, int64_t nsymbols, int32_t M, int32_t fp64_in, uint8_t *bits,
                                  void *stream) {
  const int bps = qam_bits(M);
  B2C_REQUIRE(bps, B2C_E_UNSUPPORTED, "Demodulation order %d not implemented", M);
  B2C_REQUIRE(symbols && bits && nsymbols >= 0, B2C_E_ARG, "b2c_qam_demodulate: null argument");
  if (nsymbols == 0) return B2C_OK;
  const unsigned grid = (unsigned)((nsymbols + 255) / 256);
  if (fp64_in)
    qam_demod_kernel<double2><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const double2 *>(symbols), nsymbols, M, bps, bits);
  else
    qam_demod_kernel<float2><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const float2 *>(symbols), nsymbols, M, bps, bits);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

extern "C" int b2c_bit_errors_per_slot(const b2c_geom *g, const b2c_patterns *pat, const int32_t *pattern_id, int64_t B,
                                       const uint8_t *a, const uint8_t *b, int32_t bps, uint32_t *counts, void *stream) {
  B2C_REQUIRE(g && pat && pattern_id && a && b && counts, B2C_E_ARG, "b2c_bit_errors_per_slot: null argument");
  B2C_REQUIRE(pat->pilot_re && pat->npilots, B2C_E_ARG, "b2c_bit_errors_per_slot: incomplete pattern pool");
  B2C_REQUIRE(g->nsym >= 1 && g->nsc >= 1 && (int64_t)g->nsym * g->nsc <= 65535 * 4 && bps >= 1 && bps <= 8 && B >= 0 && B < (1ll << 31),
              B2C_E_ARG, "b2c_bit_errors_per_slot: grid %dx%d bps=%d B=%lld", g->nsym, g->nsc, bps, (long long)B);
  if (B == 0) return B2C_OK;
  const size_t smem = (size_t)((g->nsym * g->nsc + 31) / 32) * sizeof(uint32_t);
  bit_errors_slot_kernel<<<(unsigned)B, 256, smem, (cudaStream_t)stream>>>(*g, *pat, pattern_id, a, b, bps, counts);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

extern "C" int b2c_count_bit_errors(const uint8_t *a, const uint8_t *b, int64_t n, uint64_t *count, void *stream) {
  B2C_REQUIRE(a && b && count && n >= 0, B2C_E_ARG, "b2c_count_bit_errors: null argument");
  if (n == 0) return B2C_OK;
  const int64_t want = (n + 256 * 64 - 1) / (256 * 64);
  const unsigned grid = (unsigned)(want < 1 ? 1 : want > 148 * 8 ? 148 * 8 : want);
  bit_errors_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, n, reinterpret_cast<unsigned long long *>(count));
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

extern "C" int b2c_count_nonfinite(const float *x, int64_t n, int32_t is_complex, uint64_t *counts, void *stream) {
  B2C_REQUIRE(x && counts && n >= 0, B2C_E_ARG, "b2c_count_nonfinite: null argument");
  if (n == 0) return B2C_OK;
  const int64_t want = (n + 256 * 16 - 1) / (256 * 16);
  const unsigned grid = (unsigned)(want < 1 ? 1 : want > 148 * 8 ? 148 * 8 : want);
  unsigned long long *c = reinterpret_cast<unsigned long long *>(counts);
  if (is_complex) nonfinite_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, c);
  else nonfinite_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, c);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

extern "C" int b2c_abs_diff_sum(const float *a, const float *b, int64_t n, double *sum, void *stream) {
  B2C_REQUIRE(a && b && sum && n >= 0, B2C_E_ARG, "b2c_abs_diff_sum: null argument");
  if (n == 0) return B2C_OK;
  const int64_t want = (n + 256 * 8 - 1) / (256 * 8);
  const unsigned grid = (unsigned)(want < 1 ? 1 : want > 148 * 8 ? 148 * 8 : want);
  abs_diff_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2 *>(a), reinterpret_cast<const float2 *>(b), n, sum);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

static int check_pair00(const b2c_geom *g, int64_t B, const float *rx, const float *H_ls, const float *H_true,
                        int64_t ls_sym_stride) {
  int rc = b2c::check_dims(g);
  if (rc) return rc;
  B2C_REQUIRE(rx && H_ls && H_true && B >= 0, B2C_E_ARG, "pair-(0,0) view: null argument or B < 0");
  B2C_REQUIRE(g->pitch == 0 || g->pitch >= g->nsc, B2C_E_ARG, "pair-(0,0) view: pitch=%d < nsc=%d", g->pitch, g->nsc);
  B2C_REQUIRE(ls_sym_stride >= g->nsc && ls_sym_stride <= (int64_t)g->nrx * g->ntx * (g->pitch ? g->pitch : g->nsc), B2C_E_ARG,
              "pair-(0,0) view: ls_sym_stride=%lld", (long long)ls_sym_stride);
  return B2C_OK;
}

extern "C" int b2c_pair00_moments(const b2c_geom *g, int64_t B, const float *rx, const float *H_ls, const float *H_true,
                                  int64_t ls_sym_stride, double *moments, void *stream) {
  int rc = check_pair00(g, B, rx, H_ls, H_true, ls_sym_stride);
  if (rc) return rc;
  B2C_REQUIRE(moments, B2C_E_ARG, "b2c_pair00_moments: null moments");
  if (B == 0) return B2C_OK;
  pair00_moments_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(
      *g, reinterpret_cast<const float2 *>(rx), reinterpret_cast<const float2 *>(H_ls),
      reinterpret_cast<const float2 *>(H_true), (int)ls_sym_stride, moments);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

extern "C" int b2c_pair00_errors(const b2c_geom *g, int64_t B, const float *H_ls, const float *H_true,
                                 int64_t ls_sym_stride, const float *alpha, double *out, void *stream) {
  int rc = check_pair00(g, B, H_true, H_ls, H_true, ls_sym_stride);
  if (rc) return rc;
  B2C_REQUIRE(out, B2C_E_ARG, "b2c_pair00_errors: null out");
  if (B == 0) return B2C_OK;
  pair00_errors_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(*g, reinterpret_cast<const float2 *>(H_ls),
                                                                       reinterpret_cast<const float2 *>(H_true),
                                                                       (int)ls_sym_stride, alpha, out);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

extern "C" int b2c_ml_features(const b2c_geom *g, const b2c_patterns *pat, const int32_t *pattern_id, int64_t B,
                               const float *rx, const float *H_ls, const float *H_true, int64_t ls_sym_stride,
                               int32_t layout, int32_t norm_mode, const float *norm, float *inputs, float *targets,
                               void *stream) {
  int rc = check_pair00(g, B, rx, H_ls, H_true, ls_sym_stride);
  if (rc) return rc;
  B2C_REQUIRE(pat && pattern_id && inputs && targets, B2C_E_ARG, "b2c_ml_features: null argument");
  B2C_REQUIRE((layout == 0 || layout == 1) && norm_mode >= 0 && norm_mode <= 2 && (norm_mode != 2 || norm), B2C_E_ARG,
              "b2c_ml_features: layout=%d norm_mode=%d", layout, norm_mode);
  if (B == 0) return B2C_OK;
  ml_features_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(
      *g, *pat, pattern_id, reinterpret_cast<const float2 *>(rx), reinterpret_cast<const float2 *>(H_ls),
      reinterpret_cast<const float2 *>(H_true), (int)ls_sym_stride, layout, norm_mode, norm, inputs, targets);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}
