// Stand-alone public-API kernels that are not on the fused path:
//   b2c_apply_channel -- MIMOChannel.apply_channel for arbitrary tx / H (src/channel_simulator.py:313-345)
//   b2c_tdl_full      -- ChannelModel.generate_time_varying_channel (:84-127), every time sample
#include "b2c_common.cuh"
#include "b2c_rng.cuh"

namespace b2c {

constexpr int MAXT = B2C_MAX_TAPS;
constexpr int NOSC = B2C_N_OSC;

// ---- apply_channel ---------------------------------------------------------------------------------
// pass 1: y = H x per RE (:330-334) into rx, and the slot's sum |y|^2 (for :337).
__global__ void __launch_bounds__(256) apply_hx_kernel(b2c_geom g, const float2 *__restrict__ tx,
                                                       const float2 *__restrict__ H, float2 *__restrict__ rx,
                                                       double *__restrict__ power, int chunks) {
  __shared__ float red[33];
  const int64_t b = blockIdx.x / chunks;
  const int chunk = blockIdx.x - (int)b * chunks;
  const int per = g.nsym * g.nrx * g.nsc;
  const int e = chunk * 256 + threadIdx.x;
  float p = 0.f;
  if (e < per) {
    int k = e % g.nsc, q = e / g.nsc;
    int r = q % g.nrx, s = q / g.nrx;
    float2 acc = make_float2(0.f, 0.f);
    for (int t = 0; t < g.ntx; ++t) {
      float2 h = __ldg(H + (((b * g.nsym + s) * g.nrx + r) * g.ntx + t) * (int64_t)g.nsc + k);
      float2 x = __ldg(tx + ((b * g.nsym + s) * g.ntx + t) * (int64_t)g.nsc + k);
      acc = cadd(acc, cmul(h, x));
    }
    rx[b * per + e] = acc;
    p = cabs2(acc);
  }
  float tot = block_sum(p, red);
  if (threadIdx.x == 0) atomicAdd(power + b, (double)tot);
}

// pass 2: rx += noise_std * (n_re + j n_im)  (:338-343)
__global__ void __launch_bounds__(256) apply_noise_kernel(b2c_geom g, b2c_slots slots, b2c_inject inj, int has_inj,
                                                          float2 *__restrict__ rx, const double *__restrict__ power,
                                                          int chunks) {
  const int64_t b = blockIdx.x / chunks;
  const int chunk = blockIdx.x - (int)b * chunks;
  const int per = g.nsym * g.nrx * g.nsc;
  const int e = chunk * 256 + threadIdx.x;
  if (e >= per) return;
  double p_sig = power[b] / (double)per;
  float sigma = (float)sqrt(p_sig / pow(10.0, (double)slots.snr_db[b] / 10.0) / 2.0);
  float2 n;
  if (has_inj) {
    n = __ldg(reinterpret_cast<const float2 *>(inj.noise) + b * per + e);
  } else {
    int k = e % g.nsc, q = e / g.nsc;
    int r = q % g.nrx, s = q / g.nrx;
    const int half = (g.nsc + 1) >> 1, h = k >= half, l = h ? k - half : half - 1 - k;   // b2c.h: lane |f|-1, half f>0
    uint4 w = draw(make_key(slots.seed, slots.slot0 + b), STREAM_NOISE, (uint32_t)((s * g.nrx + r) * B2C_RNG_LANES + l));
    n = h ? normal_pair(w.z, w.w) : normal_pair(w.x, w.y);
  }
  float2 y = rx[b * per + e];
  rx[b * per + e] = make_float2(fmaf(sigma, n.x, y.x), fmaf(sigma, n.y, y.y));
}

// ---- tdl_full ---------------------------------------------------------------------------------------
struct TdlArgs {
  int ntaps, ntx, nrx, L;
  int tap_delay[MAXT];
  const int *tap_path;
  const float *tap_amp;
  const float *jakes_u;   // [npaths][ntx][nrx][2][20] or null
  PhiloxKey key;
  float doppler_hz, sample_period_s;
  int64_t num_samples;
  float2 *out;
};

// grid (time chunks, links); link = (tap, tx, rx)
__global__ void __launch_bounds__(256) tdl_full_kernel(TdlArgs a) {
  __shared__ float2 osc[NOSC];   // (phase turns, doppler shift in Hz)
  const int link = blockIdx.y;
  const int t = link / (a.ntx * a.nrx), rem = link - t * (a.ntx * a.nrx);
  const int tx = rem / a.nrx, rx = rem - tx * a.nrx;
  const int p = a.tap_path[t];
  if (threadIdx.x < NOSC / 2) {
    int pr = threadIdx.x;
    float ua0, up0, ua1, up1;
    if (a.jakes_u) {
      const float *ju = a.jakes_u + (((int64_t)p * a.ntx + tx) * a.nrx + rx) * (2 * NOSC);
      ua0 = ju[2 * pr];
      ua1 = ju[2 * pr + 1];
      up0 = ju[NOSC + 2 * pr];
      up1 = ju[NOSC + 2 * pr + 1];
    } else {
      uint4 w = draw(a.key, STREAM_JAKES, (uint32_t)(((p * a.ntx + tx) * a.nrx + rx) * (NOSC / 2) + pr));
      ua0 = u01(w.x);
      up0 = u01(w.y);
      ua1 = u01(w.z);
      up1 = u01(w.w);
    }
    osc[2 * pr] = make_float2(up0, a.doppler_hz * cospif(2.0f * ua0));
    osc[2 * pr + 1] = make_float2(up1, a.doppler_hz * cospif(2.0f * ua1));
  }
  __syncthreads();
  const int64_t n = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (n >= a.num_samples) return;
  // time in seconds as a double product rounded once: keeps the phase accurate for long records
  const float tsec = (float)((double)n * (double)a.sample_period_s);
  float ar = 0.f, ai = 0.f;
#pragma unroll 4
  for (int i = 0; i < NOSC; ++i) {
    float2 pd = osc[i];
    float2 z = cis_turns(fmaf(tsec, pd.y, pd.x));
    ar += z.x;
    ai += z.y;
  }
  const float amp = a.tap_amp[t];
  a.out[((n * a.nrx + rx) * a.ntx + tx) * (int64_t)a.L + a.tap_delay[t]] = make_float2(amp * ar, amp * ai);
}

// ---- time-domain TDL convolution (per-symbol, circular) ---------------------------------------------------------------
// y[s][rx][n] = sum_tx sum_taps g[rx][s][tx][t] * x[s][tx][(n - d_t) mod N]   on the N-sample symbol bodies, prefix re-attached.
// The reference never convolves in time: it multiplies the CFR of the symbol-start gains into the grid per bin
// (src/channel_simulator.py:274-345), which is EXACTLY a circular convolution of each symbol body with that symbol's
// taps (a linear one would leak the ETU taps at delay 77 > CP 72 into the next symbol and could not reproduce the
// reference).  modulate -> this kernel -> demodulate therefore lands on the frequency-domain rx (noise aside): the
// time-domain statement of north_star kernel 1 and an in-situ exercise of the OFDM modem (K2).
// One CTA per (slot, symbol, rx): the ntx symbol bodies staged in shared memory, 4 outputs per thread.
constexpr int TDC_THREADS = 256;
__global__ void __launch_bounds__(TDC_THREADS) tdl_circular_kernel(b2c_geom g, b2c_profiles prof, const int32_t *__restrict__ model_id,
                                                                  const float2 *__restrict__ gains, const float2 *__restrict__ x,
                                                                  float2 *__restrict__ y) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *xs = reinterpret_cast<float2 *>(smem_raw);          // [ntx][N]
  __shared__ float2 gs[B2C_MAX_ANT * MAXT];
  __shared__ int ds[MAXT];
  const int N = g.fft_size, cp = g.cp_length, T = N + cp;
  const int rx = blockIdx.x % g.nrx;
  const int64_t bs = blockIdx.x / g.nrx;                        // b * nsym + s
  const int64_t b = bs / g.nsym;
  const int s = (int)(bs - b * g.nsym);
  const int m = model_id[b];
  const int ntaps = prof.ntaps[m];
  for (int i = threadIdx.x; i < g.ntx * N; i += TDC_THREADS) {
    const int tx = i / N, n = i - tx * N;
    xs[i] = __ldg(x + (bs * g.ntx + tx) * (int64_t)T + cp + n);                 // symbol body (prefix stripped)
  }
  if (threadIdx.x < g.ntx * MAXT)
    gs[threadIdx.x] = __ldg(gains + (((b * g.nrx + rx) * g.nsym + s) * g.ntx) * (int64_t)MAXT + threadIdx.x);
  if (threadIdx.x < MAXT) ds[threadIdx.x] = threadIdx.x < ntaps ? prof.tap_delay[m * MAXT + threadIdx.x] : 0;
  __syncthreads();
  float2 *yr = y + (bs * g.nrx + rx) * (int64_t)T;
  for (int n = threadIdx.x; n < N; n += TDC_THREADS) {
    float2 acc = make_float2(0.f, 0.f);
    for (int tx = 0; tx < g.ntx; ++tx)
      for (int t = 0; t < ntaps; ++t) acc = cadd(acc, cmul(gs[tx * MAXT + t], xs[tx * N + ((n - ds[t]) & (N - 1))]));
    yr[cp + n] = acc;
    if (n >= N - cp) yr[n - (N - cp)] = acc;                                    // cyclic prefix = the last cp samples
  }
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_tdl_circular(const b2c_geom *g, const b2c_profiles *prof, const int32_t *model_id, int64_t B, const float *gains,
                                const float *x_time, float *y_time, void *stream) {
  B2C_REQUIRE(g && prof && model_id && gains && x_time && y_time, B2C_E_ARG, "b2c_tdl_circular: null argument");
  if (int rc = check_geom(g)) return rc;
  B2C_REQUIRE(prof->tap_delay, B2C_E_ARG, "b2c_tdl_circular: b2c_profiles.tap_delay missing");
  B2C_REQUIRE(g->fft_size >= 64 && (g->fft_size & (g->fft_size - 1)) == 0 && g->cp_length >= 0 && g->cp_length <= g->fft_size, B2C_E_UNSUPPORTED,
              "b2c_tdl_circular: fft_size=%d must be a power of two, cp=%d", g->fft_size, g->cp_length);
  B2C_REQUIRE(B >= 0 && B * g->nsym * g->nrx < (1ll << 31), B2C_E_ARG, "b2c_tdl_circular: B=%lld out of range", (long long)B);
  if (B == 0) return B2C_OK;
  const size_t smem = (size_t)g->ntx * g->fft_size * sizeof(float2);
  B2C_REQUIRE(smem <= 200 * 1024, B2C_E_UNSUPPORTED, "b2c_tdl_circular: %zu B shared memory needed", smem);
  if (smem > 48 * 1024) B2C_CUDA(set_max_smem<tdl_circular_kernel>(smem));
  tdl_circular_kernel<<<(unsigned)(B * g->nsym * g->nrx), TDC_THREADS, smem, (cudaStream_t)stream>>>(
      *g, *prof, model_id, reinterpret_cast<const float2 *>(gains), reinterpret_cast<const float2 *>(x_time),
      reinterpret_cast<float2 *>(y_time));
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

extern "C" int b2c_apply_channel(const b2c_geom *g, const b2c_slots *slots, const b2c_inject *inj, int64_t B,
                                 const float *tx, const float *H, float *rx, double *power_scratch, void *stream) {
  B2C_REQUIRE(g && slots && tx && H && rx && power_scratch && slots->snr_db, B2C_E_ARG, "b2c_apply_channel: null argument");
  if (int rc = check_geom(g)) return rc;
  B2C_REQUIRE(!inj || inj->noise, B2C_E_ARG, "b2c_apply_channel: inject struct without noise");
  if (B == 0) return B2C_OK;
  const int per = g->nsym * g->nrx * g->nsc;
  const int chunks = (per + 255) / 256;
  B2C_REQUIRE(B > 0 && B * chunks < (1ll << 31), B2C_E_ARG, "b2c_apply_channel: B=%lld out of range", (long long)B);
  cudaStream_t st = (cudaStream_t)stream;
  B2C_CUDA(cudaMemsetAsync(power_scratch, 0, sizeof(double) * B, st));
  apply_hx_kernel<<<(unsigned)(B * chunks), 256, 0, st>>>(*g, reinterpret_cast<const float2 *>(tx),
                                                          reinterpret_cast<const float2 *>(H),
                                                          reinterpret_cast<float2 *>(rx), power_scratch, chunks);
  B2C_CUDA(cudaGetLastError());
  b2c_inject ij = {};
  if (inj) ij = *inj;
  apply_noise_kernel<<<(unsigned)(B * chunks), 256, 0, st>>>(*g, *slots, ij, inj != nullptr,
                                                             reinterpret_cast<float2 *>(rx), power_scratch, chunks);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

extern "C" int b2c_tdl_full(const b2c_geom *g, const b2c_profiles *prof, int32_t model_id, float doppler_hz,
                            float sample_period_s, int64_t num_samples, int32_t L, const int32_t *tap_delay_host,
                            int32_t ntaps, const float *jakes_u, uint64_t seed, int64_t slot, float *out, void *stream) {
  B2C_REQUIRE(g && prof && tap_delay_host && out, B2C_E_ARG, "b2c_tdl_full: null argument");
  B2C_REQUIRE(g->ntx >= 1 && g->ntx <= B2C_MAX_ANT && g->nrx >= 1 && g->nrx <= B2C_MAX_ANT, B2C_E_UNSUPPORTED,
              "b2c_tdl_full: antenna counts %dx%d", g->ntx, g->nrx);
  B2C_REQUIRE(model_id >= 0 && model_id < prof->n_models, B2C_E_ARG, "b2c_tdl_full: model_id=%d", model_id);
  B2C_REQUIRE(num_samples >= 0 && L >= 1, B2C_E_ARG, "b2c_tdl_full: num_samples=%lld L=%d", (long long)num_samples, L);
  if (num_samples == 0) return B2C_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // ntaps comes from the host tables (it is prof->ntaps[model_id]): no device read-back, nothing synchronises
  B2C_REQUIRE(ntaps >= 1 && ntaps <= MAXT, B2C_E_ARG, "b2c_tdl_full: ntaps=%d", ntaps);
  TdlArgs a = {};
  a.ntaps = ntaps;
  a.ntx = g->ntx;
  a.nrx = g->nrx;
  a.L = L;
  for (int t = 0; t < ntaps; ++t) {
    B2C_REQUIRE(tap_delay_host[t] >= 0 && tap_delay_host[t] < L, B2C_E_ARG, "b2c_tdl_full: tap delay %d outside L=%d",
                tap_delay_host[t], L);
    a.tap_delay[t] = tap_delay_host[t];
  }
  a.tap_path = prof->tap_path + model_id * MAXT;
  a.tap_amp = prof->tap_amp + model_id * MAXT;
  a.jakes_u = jakes_u;
  a.key = make_key(seed, slot);
  a.doppler_hz = doppler_hz;
  a.sample_period_s = sample_period_s;
  a.num_samples = num_samples;
  a.out = reinterpret_cast<float2 *>(out);
  B2C_CUDA(cudaMemsetAsync(out, 0, sizeof(float2) * (size_t)num_samples * g->nrx * g->ntx * L, st));
  dim3 grid((unsigned)((num_samples + 255) / 256), (unsigned)(ntaps * g->ntx * g->nrx));
  tdl_full_kernel<<<grid, 256, 0, st>>>(a);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}
