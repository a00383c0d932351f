// K3: LS pilot division + plan interpolation (+ default MMSE, + squared-error statistics) on
// caller-supplied received grids, and K5: folding per-slot statistics into per-bin accumulators.
#include "b2c_common.cuh"

namespace b2c {

constexpr int EST_THREADS = 320;

struct LsArgs {
  b2c_geom g;
  b2c_patterns pat;
  const int32_t *pattern_id;
  const float *snr_db;
  const float2 *rx, *pilots, *hp_in, *H_true;
  int64_t pilots_stride;
  int mmse_mode;
  float2 *H_ls, *H_mmse, *hp_out;
  double *stats;
  const int32_t *hp_col;   // row of (slot b, rx 0) in hp_in (NULL: b * nrx)
  int64_t hp_ld;           // row stride of hp_in
};

// mmse_mode 2 (hp_in = Wiener-filtered pilot estimates): the interpolated grid IS the MMSE estimate
__device__ __forceinline__ const float2 *hp_in_row(const LsArgs &a, int64_t b, int rx) {
  return a.hp_in + ((a.hp_col ? (int64_t)a.hp_col[b] : b * a.g.nrx) + rx) * a.hp_ld;
}

// One CTA per (slot, rx antenna).  The reference's rx_4d is rx replicated over tx
// (src/dataset_generator.py:63-64), so the LS/MMSE grids are computed once per (slot, rx) and
// written ntx times.
template <int NTX, bool EXACT, int NSC>
__global__ void __launch_bounds__(EST_THREADS, 2) ls_interp_kernel(LsArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *hp = reinterpret_cast<float2 *>(smem_raw);
  __shared__ float red[33];
  __shared__ float ssm[EST_THREADS / 32][6];

  const int nsc = NSC ? NSC : a.g.nsc, nsym = a.g.nsym, ntx = EXACT ? NTX : a.g.ntx, nrx = a.g.nrx;
  const int64_t b = blockIdx.x / nrx;
  const int rx = blockIdx.x - (int)b * nrx;
  const int pid = a.pattern_id[b];
  const int np = a.pat.npilots[pid];
  const int *pre = a.pat.pilot_re + (int64_t)pid * a.pat.np_max;

  // h_p = y_p / (x_p + 1e-12), row-major pilot order (src/baseline_estimators.py:109-110)
  const float2 *const hp_row = a.hp_in ? hp_in_row(a, b, rx) : nullptr;
  const bool m2 = a.mmse_mode == 2;
  float psum = 0.f;
  for (int j = threadIdx.x; j < np; j += EST_THREADS) {
    float2 h;
    if (a.hp_in) {
      h = __ldg(hp_row + j);
    } else {
      int e = __ldg(pre + j);
      int s = e / nsc, k = e - s * nsc;
      float2 y = __ldg(a.rx + ((b * nsym + s) * nrx + rx) * (int64_t)nsc + k);
      h = ls_divide(y, __ldg(a.pilots + b * a.pilots_stride + j));
    }
    hp[j] = h;
    if (j == 0) hp[a.pat.np_max] = make_float2(0.f, 0.f);   // zero slot for REs outside the hull
    if (a.hp_out) a.hp_out[(b * nrx + rx) * (int64_t)a.pat.np_max + j] = h;
    psum += cabs2(h);
  }
  float P = block_sum(psum, red) / (float)np;   // barrier inside: hp[] complete past this point
  float alpha = 0.f;
  if (a.mmse_mode == 1) {
    float sig2 = exp10f(-0.1f * a.snr_db[b]);    // noise_variance = 1/snr_linear (:174-175)
    alpha = P / (P + sig2);
  }
  if (m2) alpha = 1.f;

  float st[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};   // [0] antenna pair (rx, 0), [1] all tx of this rx

  // Main pass.  Thread t owns bins t and t + EST_THREADS (the grid has at most 2 * EST_THREADS used
  // bins); all loads of a symbol (2 plan entries, up to 2 * ntx true-channel values) are issued
  // before anything is consumed.  Per-slot 64-bit bases + 32-bit element offsets; with NSC fixed the
  // per-tx offsets are immediates.
  const int nre = nsym * nsc;
  const uint4 *plan = reinterpret_cast<const uint4 *>(a.pat.plan) + (int64_t)pid * (nre + 1);
  const int64_t slot_h = (int64_t)nsym * nrx * ntx * nsc;
  float2 *const Lb = (a.H_ls && !m2) ? a.H_ls + b * slot_h : nullptr;
  float2 *const Mb = a.H_mmse ? a.H_mmse + b * slot_h : nullptr;
  const float2 *const Tb = (a.H_true && a.stats) ? a.H_true + b * slot_h : nullptr;
  const int k0 = threadIdx.x;
  const bool v0 = k0 < nsc, v1 = k0 + EST_THREADS < nsc;
  int oP0 = v0 ? k0 : nre, oP1 = v1 ? k0 + EST_THREADS : nre;      // row nre = the all-outside entry
  const int dP0 = v0 ? nsc : 0, dP1 = v1 ? nsc : 0;
  int oH = rx * ntx * nsc + k0;
  const int dH = nrx * ntx * nsc;
  const float2 zero2 = make_float2(0.f, 0.f);
#pragma unroll 2
  for (int sy = 0; sy < nsym; ++sy) {
    const uint4 e0 = __ldg(plan + oP0), e1 = __ldg(plan + oP1);
    float2 h0[NTX], h1[NTX];
#pragma unroll
    for (int tx = 0; tx < NTX; ++tx) {
      const bool on = (EXACT || tx < ntx) && Tb != nullptr;
      h0[tx] = (on && v0) ? __ldg(Tb + oH + tx * nsc) : zero2;
      h1[tx] = (on && v1) ? __ldg(Tb + oH + tx * nsc + EST_THREADS) : zero2;
    }
    const float2 l0 = plan_apply(plan_decode(e0), hp), l1 = plan_apply(plan_decode(e1), hp);
    const float2 m0 = cscale(alpha, l0), m1 = cscale(alpha, l1);
#pragma unroll
    for (int tx = 0; tx < NTX; ++tx) {
      if (EXACT || tx < ntx) {
        const int o = oH + tx * nsc;
        if (Lb) {
          if (v0) st_stream(Lb + o, l0);
          if (v1) st_stream(Lb + o + EST_THREADS, l1);
        }
        if (Mb) {
          if (v0) st_stream(Mb + o, m0);
          if (v1) st_stream(Mb + o + EST_THREADS, m1);
        }
        if (Tb) {   // idle lanes hold h = l = 0 and add nothing
          const float e_ls = cabs2(make_float2(h0[tx].x - l0.x, h0[tx].y - l0.y)) + cabs2(make_float2(h1[tx].x - l1.x, h1[tx].y - l1.y));
          const float e_mm = cabs2(make_float2(h0[tx].x - m0.x, h0[tx].y - m0.y)) + cabs2(make_float2(h1[tx].x - m1.x, h1[tx].y - m1.y));
          const float pw = cabs2(h0[tx]) + cabs2(h1[tx]);
          st[1][0] += e_ls;
          st[1][1] += e_mm;
          st[1][2] += pw;
          if (tx == 0) {
            st[0][0] += e_ls;
            st[0][1] += e_mm;
            st[0][2] += pw;
          }
        }
      }
    }
    oP0 += dP0;
    oP1 += dP1;
    oH += dH;
  }

  if (a.stats) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float v = warp_sum(st[q][j]);
        if (lane == 0) ssm[warp][q * 3 + j] = v;
      }
    __syncthreads();
    if (threadIdx.x < 6) {
      double acc = 0.0;
      for (int w = 0; w < EST_THREADS / 32; ++w) acc += (double)ssm[w][threadIdx.x];
      if (!m2 || threadIdx.x % 3 == 1) a.stats[(b * nrx + rx) * 6 + threadIdx.x] = acc;
    }
  }
}


// Wide-store variant for padded rows (b2c_geom.pitch = 600 on the default 599-bin grid, the layout the slot
// pipeline's throughput configuration produces): thread t < 300 owns the ADJACENT bins (2t, 2t+1), so every
// row is read / written as one aligned 16-byte access per lane (see scripts/store_pattern_bench.cu: this
// store form reaches 6.8 TB/s where the 8-byte form stops at ~4.4 TB/s).  rx, H_true, H_ls and H_mmse all
// use the padded pitch; element 599 of each output row is padding.
template <int NTX, int PITCH>
__global__ void __launch_bounds__(EST_THREADS, 2) ls_interp_wide_kernel(LsArgs a) {
  constexpr int NSC = 599;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *hp = reinterpret_cast<float2 *>(smem_raw);
  __shared__ float red[33];
  __shared__ float ssm[EST_THREADS / 32][6];

  const int nsym = a.g.nsym, nrx = a.g.nrx;
  const int64_t b = blockIdx.x / nrx;
  const int rx = blockIdx.x - (int)b * nrx;
  const int pid = a.pattern_id[b];
  const int np = a.pat.npilots[pid];
  const int *pre = a.pat.pilot_re + (int64_t)pid * a.pat.np_max;

  const float2 *const hp_row = a.hp_in ? hp_in_row(a, b, rx) : nullptr;
  const bool m2 = a.mmse_mode == 2;
  float psum = 0.f;
  for (int j = threadIdx.x; j < np; j += EST_THREADS) {
    float2 h;
    if (a.hp_in) {
      h = __ldg(hp_row + j);
    } else {
      int e = __ldg(pre + j);
      int s = e / NSC, k = e - s * NSC;
      float2 y = __ldg(a.rx + ((b * nsym + s) * nrx + rx) * (int64_t)PITCH + k);
      h = ls_divide(y, __ldg(a.pilots + b * a.pilots_stride + j));
    }
    hp[j] = h;
    if (j == 0) hp[a.pat.np_max] = make_float2(0.f, 0.f);
    if (a.hp_out) a.hp_out[(b * nrx + rx) * (int64_t)a.pat.np_max + j] = h;
    psum += cabs2(h);
  }
  float P = block_sum(psum, red) / (float)np;
  float alpha = 0.f;
  if (a.mmse_mode == 1) alpha = P / (P + exp10f(-0.1f * a.snr_db[b]));
  if (m2) alpha = 1.f;

  float st[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  const int nre = nsym * NSC;
  const uint4 *plan = reinterpret_cast<const uint4 *>(a.pat.plan) + (int64_t)pid * (nre + 1);
  const int64_t slot_h = (int64_t)nsym * nrx * NTX * PITCH;
  float2 *const Lb = (a.H_ls && !m2) ? a.H_ls + b * slot_h : nullptr;
  float2 *const Mb = a.H_mmse ? a.H_mmse + b * slot_h : nullptr;
  const float2 *const Tb = (a.H_true && a.stats) ? a.H_true + b * slot_h : nullptr;
  const int t = threadIdx.x;
  const bool act = t < (NSC + 1) / 2, v1 = 2 * t + 1 < NSC;     // bin 599 (t = 299) is padding
  const int k0 = act ? 2 * t : 0;
  int oP0 = act ? k0 : nre, oP1 = (act && v1) ? k0 + 1 : nre;
  const int dP0 = act ? NSC : 0, dP1 = (act && v1) ? NSC : 0;
  int oH = rx * NTX * PITCH + k0;
  const int dH = nrx * NTX * PITCH;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
  for (int sy = 0; sy < nsym; ++sy) {
    const uint4 e0 = __ldg(plan + oP0), e1 = __ldg(plan + oP1);
    float4 h[NTX];
#pragma unroll
    for (int tx = 0; tx < NTX; ++tx)
      h[tx] = (Tb && act) ? __ldg(reinterpret_cast<const float4 *>(Tb + oH + tx * PITCH)) : zero4;
    const float2 l0 = plan_apply(plan_decode(e0), hp), l1 = plan_apply(plan_decode(e1), hp);
    const float4 l = make_float4(l0.x, l0.y, l1.x, l1.y);
    const float4 m = make_float4(alpha * l0.x, alpha * l0.y, alpha * l1.x, alpha * l1.y);
#pragma unroll
    for (int tx = 0; tx < NTX; ++tx) {
      const int o = oH + tx * PITCH;
      if (act) {
        if (Lb) __stcs(reinterpret_cast<float4 *>(Lb + o), l);
        if (Mb) __stcs(reinterpret_cast<float4 *>(Mb + o), m);
      }
      if (Tb) {
        // the padding element of H_true is not part of the statistics (selected away, not multiplied: it may hold anything)
        const float4 q = make_float4(h[tx].x, h[tx].y, v1 ? h[tx].z : 0.f, v1 ? h[tx].w : 0.f);
        const float e_ls = cabs2(make_float2(q.x - l.x, q.y - l.y)) + cabs2(make_float2(q.z - l.z, q.w - l.w));
        const float e_mm = cabs2(make_float2(q.x - m.x, q.y - m.y)) + cabs2(make_float2(q.z - m.z, q.w - m.w));
        const float pw = cabs2(make_float2(q.x, q.y)) + cabs2(make_float2(q.z, q.w));
        st[1][0] += e_ls;
        st[1][1] += e_mm;
        st[1][2] += pw;
        if (tx == 0) {
          st[0][0] += e_ls;
          st[0][1] += e_mm;
          st[0][2] += pw;
        }
      }
    }
    oP0 += dP0;
    oP1 += dP1;
    oH += dH;
  }

  if (a.stats) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float v = warp_sum(st[q][j]);
        if (lane == 0) ssm[warp][q * 3 + j] = v;
      }
    __syncthreads();
    if (threadIdx.x < 6) {
      double acc = 0.0;
      for (int w = 0; w < EST_THREADS / 32; ++w) acc += (double)ssm[w][threadIdx.x];
      if (!m2 || threadIdx.x % 3 == 1) a.stats[(b * nrx + rx) * 6 + threadIdx.x] = acc;
    }
  }
}

template <int NTX>
static int launch_ls_wide(const LsArgs &a, int64_t B, cudaStream_t stream) {
  size_t smem = (size_t)(a.pat.np_max + 1) * sizeof(float2);
  auto kern = ls_interp_wide_kernel<NTX, 600>;
  if (smem > 48 * 1024) B2C_CUDA((set_max_smem<ls_interp_wide_kernel<NTX, 600>>(smem)));
  kern<<<(unsigned)(B * a.g.nrx), EST_THREADS, smem, stream>>>(a);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

template <int NTX, bool EXACT, int NSC>
static int launch_ls(const LsArgs &a, int64_t B, cudaStream_t stream) {
  size_t smem = (size_t)(a.pat.np_max + 1) * sizeof(float2);
  auto kern = ls_interp_kernel<NTX, EXACT, NSC>;
  if (smem > 48 * 1024) B2C_CUDA((set_max_smem<ls_interp_kernel<NTX, EXACT, NSC>>(smem)));
  kern<<<(unsigned)(B * a.g.nrx), EST_THREADS, smem, stream>>>(a);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

// LS / default-MMSE on bare pilot vectors (estimate_at_pilots, src/baseline_estimators.py:23-42,155-196):
// out[v][:] = alpha_v * y[v][:] / (x[:] + 1e-12), alpha_v = 1 (mode 0) or P/(P + 10^(-snr/10)) (mode 1).
__global__ void __launch_bounds__(256) pilot_vec_kernel(const float2 *__restrict__ y, const float2 *__restrict__ x,
                                                        int n, float snr_db, int mode, float2 *__restrict__ out) {
  __shared__ float red[33];
  const int64_t v = blockIdx.x;
  float psum = 0.f;
  for (int j = threadIdx.x; j < n; j += 256) {
    float2 h = ls_divide(__ldg(y + v * n + j), __ldg(x + j));
    out[v * n + j] = h;
    psum += cabs2(h);
  }
  if (mode == 0) return;
  float P = block_sum(psum, red) / (float)n;
  float alpha = P / (P + exp10f(-0.1f * snr_db));
  for (int j = threadIdx.x; j < n; j += 256) out[v * n + j] = cscale(alpha, out[v * n + j]);
}

// ---- K5 ------------------------------------------------------------------------------------------
constexpr int BIN_THREADS = 1024;   // one CTA per bin scans every slot: latency-bound, so as many slots in flight as a CTA allows

__global__ void __launch_bounds__(BIN_THREADS)
stats_bins_kernel(b2c_geom g, const double *__restrict__ stats, const int32_t *__restrict__ bin_id,
                  const float *__restrict__ snr_db, int64_t B, double *__restrict__ bins) {
  __shared__ double sm[BIN_THREADS / 32][B2C_N_BINSTAT];
  const int bin = blockIdx.x;
  const double n_all = (double)g.nsym * g.nrx * g.ntx * g.nsc, n_pair = (double)g.nsym * g.nsc;
  double acc[B2C_N_BINSTAT];
#pragma unroll
  for (int j = 0; j < B2C_N_BINSTAT; ++j) acc[j] = 0.0;
  for (int64_t b = threadIdx.x; b < B; b += BIN_THREADS) {
    if (bin_id[b] != bin) continue;
    const double *s = stats + b * (int64_t)g.nrx * 6;   // [nrx][{pair (rx,0), all tx}][3]
    double e_ls = 0, e_mm = 0, pw = 0;
    for (int r = 0; r < g.nrx; ++r) {
      e_ls += s[r * 6 + 3];
      e_mm += s[r * 6 + 4];
      pw += s[r * 6 + 5];
    }
    // evaluate_estimator (src/baseline_estimators.py:326-331): means over the whole 4-D array
    double mse_ls = e_ls / n_all, mse_mm = e_mm / n_all, pmean = pw / n_all;
    double nm_ls = mse_ls / (pmean + 1e-12), nm_mm = mse_mm / (pmean + 1e-12);
    // compute_nmse on pair (0,0) (run_phase8_pilot_optimization.py:32-37,149-154)
    double p00 = s[2] / n_pair;
    double n00_ls = (s[0] / n_pair) / (p00 + 1e-10), n00_mm = (s[1] / n_pair) / (p00 + 1e-10);
    acc[0] += 1.0;
    acc[1] += mse_ls;
    acc[2] += mse_mm;
    acc[3] += nm_ls;
    acc[4] += nm_mm;
    acc[5] += nm_ls * nm_ls;
    acc[6] += nm_mm * nm_mm;
    acc[7] += pmean;
    acc[8] += n00_ls;
    acc[9] += n00_ls * n00_ls;
    acc[10] += n00_mm;
    acc[11] += n00_mm * n00_mm;
    if (snr_db) {
      // compute_ber_approximation (run_phase5_evaluation.py:57-68): QPSK BER proxy from the estimation NMSE,
      // effective_snr = snr / (1 + snr * nmse), ber = clip(0.5 exp(-effective_snr / 2), 1e-10, 0.5)
      const double snr_lin = pow(10.0, (double)snr_db[b] / 10.0);
      const double b_ls = 0.5 * exp(-0.5 * snr_lin / (1.0 + snr_lin * n00_ls));
      const double b_mm = 0.5 * exp(-0.5 * snr_lin / (1.0 + snr_lin * n00_mm));
      acc[12] += fmin(fmax(b_ls, 1e-10), 0.5);
      acc[13] += fmin(fmax(b_mm, 1e-10), 0.5);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < B2C_N_BINSTAT; ++j) {
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sm[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < B2C_N_BINSTAT) {
    double t = 0.0;
    for (int w = 0; w < BIN_THREADS / 32; ++w) t += sm[w][threadIdx.x];
    bins[bin * B2C_N_BINSTAT + threadIdx.x] += t;
  }
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_ls_interp(const b2c_geom *g, const b2c_patterns *pat, const int32_t *pattern_id,
                             const float *snr_db, int64_t B, const float *rx, const float *pilots,
                             int64_t pilots_stride, const float *hp_in, int32_t mmse_mode,
                             const float *H_true, float *H_ls, float *H_mmse, float *hp_out, double *stats,
                             const int32_t *hp_col, int64_t hp_ld, void *stream) {
  B2C_REQUIRE(g && pat && pattern_id, B2C_E_ARG, "b2c_ls_interp: null argument");
  B2C_REQUIRE(g->nsym >= 1 && g->nsc >= 1 && g->nsc <= 2 * EST_THREADS && g->ntx >= 1 &&
                  g->ntx <= B2C_MAX_ANT && g->nrx >= 1 && g->nrx <= B2C_MAX_ANT * B2C_MAX_ANT,
              B2C_E_UNSUPPORTED, "b2c_ls_interp: geometry %dx%d grid, %dx%d antennas unsupported", g->nsym, g->nsc,
              g->ntx, g->nrx);
  B2C_REQUIRE(pat->plan && pat->pilot_re && pat->npilots, B2C_E_ARG, "b2c_ls_interp: incomplete pattern pool");
  B2C_REQUIRE(hp_in || (rx && pilots), B2C_E_ARG, "b2c_ls_interp: need rx and pilots (or hp_in)");
  B2C_REQUIRE(mmse_mode == 0 || (mmse_mode == 1 && snr_db) || (mmse_mode == 2 && hp_in), B2C_E_ARG,
              "b2c_ls_interp: mmse_mode=%d invalid, or snr_db (mode 1) / hp_in (mode 2) missing", mmse_mode);
  B2C_REQUIRE(mmse_mode != 0 || !H_mmse, B2C_E_ARG, "b2c_ls_interp: H_mmse requested with mmse_mode=0");
  B2C_REQUIRE((!hp_col && hp_ld == 0) || (hp_in && (hp_ld == 0 || hp_ld >= pat->np_max)), B2C_E_ARG,
              "b2c_ls_interp: hp_col / hp_ld describe hp_in (hp_ld >= np_max)");
  B2C_REQUIRE(!stats || H_true, B2C_E_ARG, "b2c_ls_interp: stats need H_true");
  B2C_REQUIRE(B >= 0 && B * g->nrx < (1ll << 31), B2C_E_ARG, "b2c_ls_interp: B=%lld out of range", (long long)B);
  B2C_REQUIRE(pat->np_max >= 1 && pat->np_max <= 65534 && (size_t)pat->np_max * 8 <= 100 * 1024, B2C_E_UNSUPPORTED,
              "b2c_ls_interp: np_max=%d unsupported", pat->np_max);
  if (B == 0) return B2C_OK;
  LsArgs a = {};
  a.g = *g;
  a.pat = *pat;
  a.pattern_id = pattern_id;
  a.snr_db = snr_db;
  a.rx = reinterpret_cast<const float2 *>(rx);
  a.pilots = reinterpret_cast<const float2 *>(pilots);
  a.hp_in = reinterpret_cast<const float2 *>(hp_in);
  a.H_true = reinterpret_cast<const float2 *>(H_true);
  a.pilots_stride = pilots_stride;
  a.mmse_mode = mmse_mode;
  a.H_ls = reinterpret_cast<float2 *>(H_ls);
  a.H_mmse = reinterpret_cast<float2 *>(H_mmse);
  a.hp_out = reinterpret_cast<float2 *>(hp_out);
  a.stats = stats;
  a.hp_col = hp_col;
  a.hp_ld = hp_ld ? hp_ld : pat->np_max;
  cudaStream_t st = (cudaStream_t)stream;
  if (g->pitch != 0 && g->pitch != g->nsc) {   // padded rows (rx, H_true, H_ls, H_mmse alike): wide kernel
    B2C_REQUIRE(g->nsc == 599 && g->pitch == 600 && (g->ntx == 1 || g->ntx == 2 || g->ntx == 4 || g->ntx == 8), B2C_E_UNSUPPORTED,
                "b2c_ls_interp: pitch=%d needs the default grid (599 bins, pitch 600) and ntx in {1,2,4,8}", g->pitch);
    if (g->ntx == 1) return launch_ls_wide<1>(a, B, st);
    if (g->ntx == 2) return launch_ls_wide<2>(a, B, st);
    if (g->ntx == 4) return launch_ls_wide<4>(a, B, st);
    return launch_ls_wide<8>(a, B, st);
  }
  if (g->nsc == 599) {   // default grid, power-of-two TX count: compile-time offsets
    if (g->ntx == 1) return launch_ls<1, true, 599>(a, B, st);
    if (g->ntx == 2) return launch_ls<2, true, 599>(a, B, st);
    if (g->ntx == 4) return launch_ls<4, true, 599>(a, B, st);
    if (g->ntx == 8) return launch_ls<8, true, 599>(a, B, st);
  }
  if (g->ntx <= 1) return launch_ls<1, false, 0>(a, B, st);
  if (g->ntx <= 2) return launch_ls<2, false, 0>(a, B, st);
  if (g->ntx <= 4) return launch_ls<4, false, 0>(a, B, st);
  return launch_ls<8, false, 0>(a, B, st);
}

extern "C" int b2c_stats_bins(const b2c_geom *g, const double *stats, const int32_t *bin_id, const float *snr_db,
                              int64_t B, int32_t nbins, double *bins, void *stream) {
  B2C_REQUIRE(g && stats && bin_id && bins, B2C_E_ARG, "b2c_stats_bins: null argument");
  if (int rc = check_geom(g)) return rc;
  B2C_REQUIRE(nbins >= 1 && B >= 0, B2C_E_ARG, "b2c_stats_bins: nbins=%d B=%lld", nbins, (long long)B);
  if (B == 0) return B2C_OK;
  stats_bins_kernel<<<nbins, BIN_THREADS, 0, (cudaStream_t)stream>>>(*g, stats, bin_id, snr_db, B, bins);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

extern "C" int b2c_pilot_vectors(const float *y, const float *x, int64_t nvec, int32_t n, float snr_db,
                                 int32_t mmse_mode, float *out, void *stream) {
  B2C_REQUIRE(y && x && out, B2C_E_ARG, "b2c_pilot_vectors: null argument");
  B2C_REQUIRE(n >= 1 && nvec >= 0 && nvec < (1ll << 31) && (mmse_mode == 0 || mmse_mode == 1), B2C_E_ARG,
              "b2c_pilot_vectors: n=%d nvec=%lld mode=%d", n, (long long)nvec, mmse_mode);
  if (nvec == 0) return B2C_OK;
  pilot_vec_kernel<<<(unsigned)nvec, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2 *>(y), reinterpret_cast<const float2 *>(x), n, snr_db, mmse_mode,
      reinterpret_cast<float2 *>(out));
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}
