// K1a tap gains + the fused per-slot kernel (CFR, grid, y = Hx + n, LS, interpolation, MMSE, stats).
//
// Work decomposition: the channel is sampled once per OFDM symbol (src/channel_simulator.py:300-302),
// so a slot's whole channel is nsym*nrx*ntx*ntaps complex tap gains (8 KB for 4x4 ETU).  K1a
// evaluates those and the slot's noise scale; the slot kernel then runs one CTA per (slot, rx
// antenna): it expands the gains to the CFR on the used bins by a direct ntaps-term DFT with
// tabulated twiddles, and because every other quantity of the slot (grid symbols, noise, LS
// pilots, interpolated estimates, errors) is a pure function of (gains, counters, plan) it never
// re-reads anything it wrote: HBM traffic is the output arrays only.
#include <stdlib.h>

#include "b2c_common.cuh"
#include "b2c_rng.cuh"

namespace b2c {

constexpr int MAXT = B2C_MAX_TAPS;
constexpr int NOSC = B2C_N_OSC;
constexpr int GAIN_THREADS = 288;   // 4x4 ETU: 9 taps x 16 links x 2 half-sums = 288 work items
constexpr int SLOT_THREADS = 320;   // thread t owns the mirror bins -(t+1), +(t+1): covers up to 639 used bins
constexpr int RNG_LANES = B2C_RNG_LANES;
constexpr int WIDE_PITCH = 600;     // padded row pitch of the wide-store kernel (599 used bins + 1)

// ------------------------------------------------------------------------------------------
// K1a: Jakes sum-of-sinusoids gains at the symbol-start instants (src/channel_simulator.py:102-125
// evaluated only at the samples :301-302 consumes) and the AWGN scale of :337-340.
// One CTA per slot.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GAIN_THREADS, 4)
tap_gains_kernel(b2c_geom g, b2c_profiles prof, b2c_slots slots, b2c_inject inj, int has_inj,
                 float2 *__restrict__ gains, float *__restrict__ noise_std) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *osc = reinterpret_cast<float2 *>(smem_raw);   // [link][20] (phase turns, turns/symbol)
  __shared__ double dred[GAIN_THREADS / 32];
  __shared__ uint32_t ltab[MAXT * B2C_MAX_ANT * B2C_MAX_ANT];   // link -> t | tx << 8 | rx << 16 | path << 24

  const int64_t b = blockIdx.x;
  const int m = slots.model_id[b];
  const int ntaps = prof.ntaps[m];
  const int ntx = g.ntx, nrx = g.nrx, nsym = g.nsym;
  const int *tap_path = prof.tap_path + m * MAXT;
  const float *tap_amp = prof.tap_amp + m * MAXT;
  const float fdT = slots.doppler_hz[b] * g.symbol_period_s;   // Doppler turns per symbol at cos = 1
  const PhiloxKey key = make_key(slots.seed, slots.slot0 + b);
  const int nlinks = ntaps * ntx * nrx;                       // link = (tap, tx, rx)

  // The (tap, tx, rx, path) decomposition of a link costs runtime integer divisions: done once per link here, read back
  // from shared memory by the three stages below (they were ~45 % of stage 1's instructions).
  for (int link = threadIdx.x; link < nlinks; link += blockDim.x) {
    const int t = link / (ntx * nrx), rem = link - t * (ntx * nrx);
    const int tx = rem / nrx, rx = rem - tx * nrx;
    ltab[link] = (uint32_t)t | ((uint32_t)tx << 8) | ((uint32_t)rx << 16) | ((uint32_t)tap_path[t] << 24);
  }
  __syncthreads();

  // stage 1: one (phase, per-symbol phase step) pair per oscillator
  for (int i = threadIdx.x; i < nlinks * (NOSC / 2); i += blockDim.x) {
    const int link = i / (NOSC / 2), pr = i - link * (NOSC / 2);
    const uint32_t lt = ltab[link];
    const int tx = (lt >> 8) & 0xff, rx = (lt >> 16) & 0xff, p = lt >> 24;
    float ua0, up0, ua1, up1;
    if (has_inj) {
      const float *ju = inj.jakes_u + ((((int64_t)b * inj.p_max + p) * ntx + tx) * nrx + rx) * (2 * NOSC);
      ua0 = ju[2 * pr];
      ua1 = ju[2 * pr + 1];
      up0 = ju[NOSC + 2 * pr];
      up1 = ju[NOSC + 2 * pr + 1];
    } else {
      uint4 w = draw(key, STREAM_JAKES, (uint32_t)(((p * ntx + tx) * nrx + rx) * (NOSC / 2) + pr));
      ua0 = u01(w.x);
      up0 = u01(w.y);
      ua1 = u01(w.z);
      up1 = u01(w.w);
    }
    // doppler_shift = fd * cos(theta), theta = 2 pi u  (:109, :119)
    // SFU cosine: its 4e-7 absolute error is scaled by fdT (~1e-2 turns per symbol) before it reaches a phase
    osc[link * NOSC + 2 * pr] = make_float2(up0, fdT * cis_turns(ua0).x);
    osc[link * NOSC + 2 * pr + 1] = make_float2(up1, fdT * cis_turns(ua1).x);
  }
  __syncthreads();

  // stage 2: gain[link][s] = amp * sum_n exp(j 2 pi (phase_n + s * step_n))
  float2 *gout = gains + b * (int64_t)(nrx * nsym * ntx * MAXT);   // [rx][s][tx][MAXT]
  const int nout = nrx * nsym * ntx * MAXT;
  for (int o = threadIdx.x; o < nout; o += blockDim.x)
    if ((o % MAXT) >= ntaps) gout[o] = make_float2(0.f, 0.f);
  // Two threads per link, ten oscillators each, held as FIVE PAIRS in structure-of-arrays form: zr = (Re z_a, Re z_b),
  // zi = (Im z_a, Im z_b).  exp(j 2 pi (phase + s step)) advances from symbol to symbol by one complex rotation
  // (rounding grows by <= ~1e-7 per step; the parity bound is 1e-4), which in this form is four packed operations per
  // pair with no operand splats; the symbol loop is outermost, so the only state is the 5 x (z, w) pairs.
  constexpr int NP2 = NOSC / 4;                                 // oscillator pairs per thread
  const int nwork = nlinks * 2;
  const bool small_step = fabsf(fdT) <= 0.1f;                   // per-symbol Doppler rotation below 0.1 turn (fd <= 1.4 kHz)
  for (int base = 0; base < nwork; base += blockDim.x) {       // warp-uniform trip count (shuffles below)
    const int i = base + threadIdx.x;
    const bool valid = i < nwork;
    const int link = valid ? i >> 1 : 0, q = i & 1;
    const float2 *oc = osc + link * NOSC + q * (NOSC / 2);
    float2 zr[NP2], zi[NP2], wr[NP2], wi[NP2], nwi[NP2];
#pragma unroll
    for (int n = 0; n < NP2; ++n) {
      float zc[2], zs[2], wc[2], ws[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float2 pd = valid ? oc[2 * n + h] : make_float2(0.f, 0.f);
        // the start phase takes the SFU sincos (its 4e-7 error enters once); the per-symbol rotation w is applied
        // nsym times, so it gets a full-precision evaluation
        const float2 z = cis_turns(pd.x);
        zc[h] = valid ? z.x : 0.f;
        zs[h] = valid ? z.y : 0.f;
        if (small_step) {
          // |x| = 2 pi |step| <= 0.63 rad: Taylor to x^9 / x^8 (truncation < 2e-9), cheaper than the general routine
          const float x = 6.283185307179586f * pd.y, x2 = x * x;
          ws[h] = x * fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 2.7557319e-6f, -1.9841270e-4f), 8.3333333e-3f), -1.6666667e-1f), 1.0f);
          wc[h] = fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 2.4801587e-5f, -1.3888889e-3f), 4.1666667e-2f), -0.5f), 1.0f);
        } else {
          sincospif(2.0f * pd.y, &ws[h], &wc[h]);
        }
      }
      zr[n] = make_float2(zc[0], zc[1]);
      zi[n] = make_float2(zs[0], zs[1]);
      wr[n] = make_float2(wc[0], wc[1]);
      wi[n] = make_float2(ws[0], ws[1]);
      nwi[n] = make_float2(-ws[0], -ws[1]);
    }
    const uint32_t lt = ltab[link];
    const int t = lt & 0xff, tx = (lt >> 8) & 0xff, rx = (lt >> 16) & 0xff;
    const float a = tap_amp[t];
    float2 *go = gout + ((rx * nsym) * ntx + tx) * MAXT + t;     // + s * ntx * MAXT
    for (int s = 0; s < nsym; ++s) {                             // nsym is uniform: no divergence around the shuffles
      float2 sr = zr[0], si = zi[0];
#pragma unroll
      for (int n = 1; n < NP2; ++n) {
        sr = __fadd2_rn(sr, zr[n]);
        si = __fadd2_rn(si, zi[n]);
      }
      float re = sr.x + sr.y, im = si.x + si.y;
      re += __shfl_xor_sync(0xffffffffu, re, 1);                 // the other half of this link's oscillators
      im += __shfl_xor_sync(0xffffffffu, im, 1);
      if (valid && (s & 1) == q) go[s * ntx * MAXT] = make_float2(a * re, a * im);
#pragma unroll
      for (int n = 0; n < NP2; ++n) {                            // z <- z w:  (zr wr - zi wi, zr wi + zi wr)
        const float2 pr_ = __fmul2_rn(zr[n], wr[n]), pi_ = __fmul2_rn(zr[n], wi[n]);
        zr[n] = __ffma2_rn(zi[n], nwi[n], pr_);
        zi[n] = __ffma2_rn(zi[n], wr[n], pi_);
      }
    }
  }
  __syncthreads();   // this CTA's global writes are visible to its own threads past the barrier

  // stage 3: mean |sum_tx H x|^2 over the slot as the quadratic form gs^H C gs (|x| = 1,
  // same grid on every TX, :402-404), then noise_std of :338-340.  Tiny, done in double.
  // Items run over (row = rx * nsym + s, tap) with the tap axis padded to MAXT: shifts instead of divisions by ntaps.
  float2 *gs = osc;   // reuse: [rx*nsym + s][MAXT] sum over tx
  for (int i = threadIdx.x; i < nrx * nsym * MAXT; i += blockDim.x) {
    const int t = i & (MAXT - 1), rs = i / MAXT;
    if (t < ntaps) {                                            // surviving taps only
      float sr = 0.f, si = 0.f;
      for (int tx = 0; tx < ntx; ++tx) {
        float2 v = gout[(rs * ntx + tx) * MAXT + t];
        sr += v.x;
        si += v.y;
      }
      gs[i] = make_float2(sr, si);
    }
  }
  __syncthreads();
  const float2 *corr = reinterpret_cast<const float2 *>(prof.tap_corr) + m * MAXT * MAXT;
  double part = 0.0;
  for (int i = threadIdx.x; i < nrx * nsym * MAXT; i += blockDim.x) {
    const int p = i & (MAXT - 1), rs = i / MAXT;
    if (p < ntaps) {
      float2 a = gs[rs * MAXT + p];
      float accr = 0.f;                 // <= 16 fp32 terms per row; the sum over rows is carried in double
      for (int q = 0; q < ntaps; ++q) {
        float2 c = gs[rs * MAXT + q], k = corr[p * MAXT + q];
        // Re( a * conj(c) * k )
        float zr_ = fmaf(a.x, c.x, a.y * c.y), zi_ = fmaf(a.y, c.x, -a.x * c.y);
        accr = fmaf(zr_, k.x, fmaf(-zi_, k.y, accr));
      }
      part += (double)accr;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) dred[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < GAIN_THREADS / 32; ++w) tot += dred[w];
    double p_sig = tot / ((double)nsym * nrx * g.nsc);
    double snr_lin = pow(10.0, (double)slots.snr_db[b] / 10.0);
    noise_std[b] = (float)sqrt(p_sig / snr_lin / 2.0);
  }
}

// ------------------------------------------------------------------------------------------
// Fused slot kernel.
// ------------------------------------------------------------------------------------------
struct SlotArgs {
  b2c_geom g;
  b2c_profiles prof;
  b2c_patterns pat;
  b2c_slots slots;
  b2c_inject inj;
  int has_inj;
  const float2 *gains;
  const float *noise_std;
  float2 *H_true, *rx, *tx, *H_ls, *H_mmse;
  double *stats;
  int compact;   // 1: tx-replicated outputs written once (H_ls/H_mmse [B][nsym][nrx][nsc], tx [B][nsym][nsc])
  uint32_t sym_and, sym_or;   // symbol-word mask: (w & and) | or; identity, or the QPSK quantisation of b2c_slots.qpsk
  float2 *hp_out;          // b2c_pilot_io: h_ls at the pilots, row hp_col[b] + rx (NULL: not written)
  const int32_t *hp_col;
  int64_t hp_ld;
};

struct SlotCtx {
  int64_t b;
  int rx, m, ntaps, pid;
  float sigma, alpha;
  PhiloxKey key;
  const float2 *gsp;   // smem [nsym][ntx][MAXT] tap gains of this rx
  const float2 *hp;    // smem [np] LS estimates at the pilots
  uint4 *pstage;       // smem, 16-byte aligned: [2][2][SLOT_THREADS] per-thread plan-entry staging of the wide kernels, or the
                       // bulk-copied ring (PLAN_RING rows of PLAN_ROW entries, then its mbarriers and warp counters)
};

// Symbol and noise draws for resource element (s, k) of this CTA's rx antenna (layout in b2c.h:
// bin k belongs to Philox lane l = |f| - 1 and word half h = (f > 0), so that the thread owning the
// mirror pair (-f, +f) gets both bins from one call).
__device__ __forceinline__ void rng_lane(int k, int nsc, int &l, int &h) {
  const int half = (nsc + 1) >> 1;
  h = k >= half;
  l = h ? k - half : half - 1 - k;
}
__device__ __forceinline__ float2 draw_symbol(const SlotArgs &a, const SlotCtx &c, int s, int k) {
  if (a.has_inj) return cis_turns(__ldg(a.inj.sym_turns + (c.b * a.g.nsym + s) * a.g.nsc + k));
  int l, h;
  rng_lane(k, a.g.nsc, l, h);
  uint4 w = draw(c.key, STREAM_SYMBOLS, (uint32_t)((s >> 1) * RNG_LANES + l));
  return cis_u01((pick(w, (s & 1) * 2 + h) & a.sym_and) | a.sym_or);
}
__device__ __forceinline__ float2 draw_noise(const SlotArgs &a, const SlotCtx &c, int s, int k) {
  if (a.has_inj)
    return __ldg(reinterpret_cast<const float2 *>(a.inj.noise) +
                 ((c.b * a.g.nsym + s) * a.g.nrx + c.rx) * a.g.nsc + k);
  int l, h;
  rng_lane(k, a.g.nsc, l, h);
  uint4 w = draw(c.key, STREAM_NOISE, (uint32_t)((s * a.g.nrx + c.rx) * RNG_LANES + l));
  return h ? normal_pair(w.z, w.w) : normal_pair(w.x, w.y);
}

// LS at the pilots: h_p = y_p / (x_p + 1e-12) (src/baseline_estimators.py:109-110), with y_p evaluated
// directly at the pilot REs from the tx-summed gains, then the default-MMSE shrinkage factor.
template <int T, int NSC, int NTHR = SLOT_THREADS>
__device__ __forceinline__ void pilot_phase(const SlotArgs &a, SlotCtx &c, const float2 *gs, float2 *hp, float *red) {
  const int nsc = NSC ? NSC : a.g.nsc;
  const int np = a.pat.npilots[c.pid];
  const int *pre = a.pat.pilot_re + (int64_t)c.pid * a.pat.np_max;
  const float2 *tw = reinterpret_cast<const float2 *>(a.prof.tap_tw) + (int64_t)c.m * MAXT * nsc;
  float psum = 0.f;
  if (threadIdx.x == 0) hp[a.pat.np_max] = make_float2(0.f, 0.f);   // the plan's "outside the hull" slot
  float2 *const hp_g = a.hp_out ? a.hp_out + ((a.hp_col ? (int64_t)a.hp_col[c.b] : c.b * a.g.nrx) + c.rx) * a.hp_ld : nullptr;
  // A pilot costs two dependent L2 round trips (its RE index, then its T twiddles).  Pilots are taken PB at a time per
  // thread -- all PB indices first, then all PB * T twiddles, then the arithmetic -- so a CTA pays two round trips per
  // PB * 320 pilots instead of two per 320 (838 pilots: 2 instead of 6).
  constexpr int PB = T <= 9 ? 3 : 1;      // PB * T twiddles live at once: keep the 16-tap instantiation at one
  for (int base = threadIdx.x; base < np; base += PB * NTHR) {
    int e[PB];
#pragma unroll
    for (int q = 0; q < PB; ++q) {
      const int j = base + q * NTHR;
      e[q] = j < np ? __ldg(pre + j) : 0;
    }
    float2 twk[PB][T];
#pragma unroll
    for (int q = 0; q < PB; ++q) {
      const int k = e[q] % nsc;
#pragma unroll
      for (int t = 0; t < T; ++t) twk[q][t] = __ldg(tw + t * nsc + k);
    }
#pragma unroll
    for (int q = 0; q < PB; ++q) {
      const int j = base + q * NTHR;
      if (j < np) {
        const int s = e[q] / nsc, k = e[q] - s * nsc;
        float2 hsum = make_float2(0.f, 0.f);
#pragma unroll
        for (int t = 0; t < T; ++t) hsum = cadd(hsum, cmul(gs[s * MAXT + t], twk[q][t]));
        const float2 x = draw_symbol(a, c, s, k), n = draw_noise(a, c, s, k);
        const float2 y = cmul(hsum, x);
        const float2 h = ls_divide(make_float2(fmaf(c.sigma, n.x, y.x), fmaf(c.sigma, n.y, y.y)), x);
        hp[j] = h;
        if (hp_g) hp_g[j] = h;
        psum += cabs2(h);
      }
    }
  }
  // default MMSE: R_h = P I  =>  W = P/(P + sigma^2) I  (src/baseline_estimators.py:174-190)
  const float P = block_sum(psum, red) / (float)np;
  const float sig2 = exp10f(-0.1f * a.slots.snr_db[c.b]);
  c.alpha = P / (P + sig2);
}

// Main loop of the slot kernel.
//
// Bin ownership.  The used bins are symmetric about DC: k < half  <->  f = k - half  in [-half, -1],
// k >= half  <->  f = k - half + 1  in [1, half-1]  (half = (nsc+1)/2).  Thread t owns the mirror pair
//   k0 = half-1-t  (f = -(t+1))   and   k1 = half+t  (f = +(t+1)),
// whose twiddles are complex conjugates: tw[tap][k0] = conj(tw[tap][k1]).  With
//   A = sum_t Re(g_t) * tw_t ,  B = sum_t Im(g_t) * tw_t        (tw_t of the +f bin, packed FFMA2)
// both bins come out of the SAME two complex sums:
//   H(+f) = (A.x - B.y, A.y + B.x)        H(-f) = (A.x + B.y, B.x - A.y)
// i.e. half the FMA work and half the twiddle registers of evaluating the two bins separately.
// A warp still stores 32 consecutive bins per instruction (ascending for k1, descending for k0).
//
//   T      compile-time tap count (>= the profile's surviving taps; absent taps have zero gain)
//   NTX    compile-time TX bound; EXACT = (ntx == NTX) removes the per-tx predicate
//   NSC    used bins when known at compile time (599 for the default grid), 0 = read from b2c_geom
//   FAST   the throughput configuration: Philox draws, every output requested, even nsym
// Output addressing: one 64-bit base per slot (uniform) + 32-bit per-thread element offsets.
template <int T, int NTX, bool EXACT, bool EST, int NSC, bool FAST>
__device__ __forceinline__ void slot_body(const SlotArgs &a, const SlotCtx &c, float2 (&st)[2][3]) {
  const int nsc = NSC ? NSC : a.g.nsc;
  const int half = (nsc + 1) >> 1;
  const int nsym = a.g.nsym, nrx = a.g.nrx;
  const int ntx = EXACT ? NTX : a.g.ntx;
  const int t_ = threadIdx.x;
  const bool v0 = t_ < half, v1 = t_ < half - 1;          // lower-half bin k0, upper-half bin k1
  const int k0 = v0 ? half - 1 - t_ : 0, k1 = v1 ? half + t_ : 0;   // idle lanes: clamped, store nothing
  const float m1 = v1 ? 1.f : 0.f;

  // Twiddles of the +f bin (conjugate of the -f bin's table row when only that one exists).  Idle
  // lanes get zeros, so their H, LS/MMSE values and error terms are exactly zero.
  const float2 *tw = reinterpret_cast<const float2 *>(a.prof.tap_tw) + (int64_t)c.m * MAXT * nsc;
  const float2 zero2 = make_float2(0.f, 0.f), neg1 = make_float2(-1.f, -1.f), nalpha = make_float2(-c.alpha, -c.alpha);
  float2 twp[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    float2 w = zero2;
    if (v1) w = __ldg(tw + t * nsc + k1);
    else if (v0) {
      w = __ldg(tw + t * nsc + k0);
      w.y = -w.y;
    }
    twp[t] = w;
  }

  const int64_t slot_h = (int64_t)nsym * nrx * ntx * nsc;
  const int nre = nsym * nsc;
  const bool compact = !FAST && a.compact;   // compact layout: generic instantiation only
  const int64_t slot_r = (int64_t)nsym * nrx * nsc;
  float2 *const Hb = (FAST || a.H_true) ? a.H_true + c.b * slot_h : nullptr;
  float2 *const Lb = (EST && (FAST || a.H_ls)) ? a.H_ls + c.b * (compact ? slot_r : slot_h) : nullptr;
  float2 *const Mb = (EST && (FAST || a.H_mmse)) ? a.H_mmse + c.b * (compact ? slot_r : slot_h) : nullptr;
  float2 *const Rb = (FAST || a.rx) ? a.rx + c.b * slot_r : nullptr;
  float2 *const Tb = ((FAST || a.tx) && c.rx == 0) ? a.tx + c.b * (int64_t)nsym * (compact ? 1 : ntx) * nsc : nullptr;
  const uint4 *plan = EST ? reinterpret_cast<const uint4 *>(a.pat.plan) + (int64_t)c.pid * (nre + 1) : nullptr;
  const bool inj = !FAST && a.has_inj;
  const float2 *inj_noise = inj ? reinterpret_cast<const float2 *>(a.inj.noise) + c.b * slot_r : nullptr;
  const float *inj_sym = inj ? a.inj.sym_turns + c.b * (int64_t)nsym * nsc : nullptr;

  // Running 64-bit store pointers per bin (advanced once per symbol; with NSC fixed the per-tx
  // offsets are immediates).  H_ls / H_mmse share H_true's layout: their addresses are H_true's plus a
  // uniform byte delta (compact layout: rx's).
  float2 *pH0 = Hb + (c.rx * ntx * nsc + k0), *pH1 = Hb + (c.rx * ntx * nsc + k1);   // H[s][rx][0][k]
  float2 *pR0 = Rb + (c.rx * nsc + k0), *pR1 = Rb + (c.rx * nsc + k1);               // rx[s][rx][k]
  float2 *pT0 = Tb + k0, *pT1 = Tb + k1;                                             // tx[s][0][k]
  const int64_t dL = (const char *)Lb - (const char *)(compact ? Rb : Hb);
  const int64_t dM = (const char *)Mb - (const char *)(compact ? Rb : Hb);
  int oI = c.rx * nsc;              // injected-noise row offset (generic path)
  int oP0 = v0 ? k0 : nre, oP1 = v1 ? k1 : nre;   // plan rows; row nre = "outside" for idle lanes
  const int dP0 = v0 ? nsc : 0, dP1 = v1 ? nsc : 0;
  const int dH = nrx * ntx * nsc, dR = nrx * nsc, dT = (compact ? 1 : ntx) * nsc;
  const float2 *gps = c.gsp;
  const bool need_draws = FAST || Rb != nullptr || a.tx != nullptr;   // H-only calls skip the draws

  static_assert(SLOT_THREADS == RNG_LANES, "thread t draws Philox lane t");
  uint4 ws = make_uint4(0, 0, 0, 0);
  for (int s2 = 0; s2 < nsym; s2 += 2) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int s = s2 + j;
      if (FAST || s < nsym) {
        // ---- LS interpolation for this RE (identical for every tx); MMSE = alpha * LS ----------------
        float2 l0 = zero2, l1 = zero2;
        if (EST) {
          l0 = plan_apply(plan_decode(__ldg(plan + oP0)), c.hp);
          l1 = plan_apply(plan_decode(__ldg(plan + oP1)), c.hp);
          oP0 += dP0;
          oP1 += dP1;
          // pull the next symbol's entries into L1 now: their L2 latency hides behind the tx loop and
          // costs no registers (the last iteration touches the row after the pattern's plan: in bounds,
          // the pool is padded by one pattern)
          prefetch_l1(plan + oP0);
          prefetch_l1(plan + oP1);
        }
        const float2 mm0 = cscale(c.alpha, l0), mm1 = cscale(c.alpha, l1);
        // ---- CFR per tx: H[s, rx, tx, +-f] from A = sum_t Re(g) tw, B = sum_t Im(g) tw ---------------
        float2 hs0 = zero2, hs1 = zero2;
#pragma unroll
        for (int tx = 0; tx < NTX; ++tx) {
          if (EXACT || tx < ntx) {
            const float2 *gp = gps + tx * MAXT;
            float2 A = zero2, B = zero2;
#pragma unroll
            for (int t = 0; t < T; ++t) {
              const float2 gq = gp[t];        // broadcast load; the (x,x) / (y,y) splats are register moves
              A = __ffma2_rn(make_float2(gq.x, gq.x), twp[t], A);
              B = __ffma2_rn(make_float2(gq.y, gq.y), twp[t], B);
            }
            const float2 h1 = make_float2(A.x - B.y, A.y + B.x);   // +f : sum_t g_t tw_t
            const float2 h0 = make_float2(A.x + B.y, B.x - A.y);   // -f : sum_t g_t conj(tw_t)
            hs0 = __fadd2_rn(hs0, h0);
            hs1 = __fadd2_rn(hs1, h1);
            if (FAST || Hb) {
              if (v0) st_stream(pH0 + tx * nsc, h0);
              if (v1) st_stream(pH1 + tx * nsc, h1);
            }
            if (EST) {
              // compact: one copy per (s, rx), at rx's offsets
              const float2 *q0 = compact ? pR0 : pH0 + tx * nsc, *q1 = compact ? pR1 : pH1 + tx * nsc;
              if ((FAST || Lb) && (!compact || tx == 0)) {
                if (v0) st_stream((float2 *)((char *)q0 + dL), l0);
                if (v1) st_stream((float2 *)((char *)q1 + dL), l1);
              }
              if ((FAST || Mb) && (!compact || tx == 0)) {
                if (v0) st_stream((float2 *)((char *)q0 + dM), mm0);
                if (v1) st_stream((float2 *)((char *)q1 + dM), mm1);
              }
              // squared errors, packed: d = h - l (one FFMA2), acc += d*d (one FFMA2) with the re^2 / im^2
              // halves summed at the end; idle lanes hold h = l = 0.  tx 0 goes to st[0] (antenna pair
              // (rx,0) for the pilot sweep), the other tx to st[1]; the flush adds st[0] into st[1].
              const float2 h1m = cscale(m1, h1);   // lane that owns only the unpaired bin -half: drop its +f value
              float2 (&acc)[3] = st[tx == 0 ? 0 : 1];
              float2 d = __ffma2_rn(l0, neg1, h0);
              acc[0] = __ffma2_rn(d, d, acc[0]);
              d = __ffma2_rn(l1, neg1, h1m);
              acc[0] = __ffma2_rn(d, d, acc[0]);
              d = __ffma2_rn(l0, nalpha, h0);
              acc[1] = __ffma2_rn(d, d, acc[1]);
              d = __ffma2_rn(l1, nalpha, h1m);
              acc[1] = __ffma2_rn(d, d, acc[1]);
              acc[2] = __ffma2_rn(h0, h0, acc[2]);
              acc[2] = __ffma2_rn(h1m, h1m, acc[2]);
            }
          }
        }
        // ---- draws (after the tx loop: keeps them out of its register budget) ------------------------
        float2 x0 = zero2, x1 = zero2, n0 = zero2, n1 = zero2;
        if (need_draws) {
          if (inj) {
            x0 = cis_turns(__ldg(inj_sym + s * nsc + k0));
            x1 = cis_turns(__ldg(inj_sym + s * nsc + k1));
            n0 = __ldg(inj_noise + oI + k0);
            n1 = __ldg(inj_noise + oI + k1);
          } else {
            if (j == 0) ws = draw(c.key, STREAM_SYMBOLS, (uint32_t)((s2 >> 1) * RNG_LANES + t_));
            x0 = cis_u01(((j ? ws.z : ws.x) & a.sym_and) | a.sym_or);
            x1 = cis_u01(((j ? ws.w : ws.y) & a.sym_and) | a.sym_or);
            const uint4 wn = draw(c.key, STREAM_NOISE, (uint32_t)((s * nrx + c.rx) * RNG_LANES + t_));
            n0 = normal_pair(wn.x, wn.y);
            n1 = normal_pair(wn.z, wn.w);
          }
        }
        // ---- y = (sum_tx H) x + sigma n  (:330-343) ------------------------------------------------
        if (FAST || Rb) {
          const float2 y0 = cmul(hs0, x0), y1 = cmul(hs1, x1);
          if (v0) st_stream(pR0, make_float2(fmaf(c.sigma, n0.x, y0.x), fmaf(c.sigma, n0.y, y0.y)));
          if (v1) st_stream(pR1, make_float2(fmaf(c.sigma, n1.x, y1.x), fmaf(c.sigma, n1.y, y1.y)));
        }
        if (Tb) {   // the rx-0 CTA writes the (tx-replicated) grid
#pragma unroll
          for (int tx = 0; tx < NTX; ++tx) {
            if ((EXACT || tx < ntx) && (!compact || tx == 0)) {
              if (v0) st_stream(pT0 + tx * nsc, x0);
              if (v1) st_stream(pT1 + tx * nsc, x1);
            }
          }
        }
        pH0 += dH;
        pH1 += dH;
        pR0 += dR;
        pR1 += dR;
        pT0 += dT;
        pT1 += dT;
        oI += dR;
        gps += ntx * MAXT;
      }
    }
  }
}


// Main loop, wide-store variant (throughput configuration with padded rows, b2c_geom.pitch = 600).
//
// Same mirror-bin arithmetic as slot_body, but every lane issues ONE 16-byte store per row instead of two
// 8-byte stores: lane t keeps the value of its bin K (even lanes: +f, k = 300+t; odd lanes: -f, k = 299-t;
// K is even either way) and hands the value of its other bin S to the neighbouring lane t^1, whose K+1 it is.
// Loading the twiddles of bin K makes "h of K" = (A.x - B.y, A.y + B.x) and "h of S" = (A.x + B.y, B.x - A.y)
// on every lane, so the exchange needs no selects: one SHFL.BFLY pair per distinct value.  Rows are `PITCH`
// complex apart (even, so K*8 is 16-byte aligned); element 599 of each row is padding.
// Measured on the bare store pattern (scripts/store_pattern_bench.cu): 6.8 TB/s against 4.4 TB/s for the
// 8-byte form -- the kernel's ceiling moves from the store path to its arithmetic.
// STORE = false: statistics-only sweeps (pilot-density / SNR curves, sharded statistics): no array is written, so the
// grid symbols, the noise and the lane exchanges that only feed stores are skipped altogether.
// ntx >= 4: the wide kernel takes the error sums over all tx from the tx-summed CFR instead of per-tx differences
__host__ __device__ constexpr bool wide_fold(int ntx) { return ntx >= 4; }

// Bulk-staged plan rows (see slot_body_wide2 for the scheme)
constexpr int PLAN_ROW = 600;       // entries per row of the bulk-staged plan ring: 599 bins + the all-outside entry
constexpr int PLAN_RING = 4;        // rows in the ring: a row is fetched PLAN_RING symbols before it is used
// The ring's mbarriers and warp counters live right behind its rows in dynamic shared memory and are addressed, like the rows,
// by 32-bit shared-window offsets from ONE base register (static __shared__ objects reached through generic pointers made
// the compiler re-derive the window base from SR_CgaCtaId inside the loop: an S2R with its latency twice per symbol).
constexpr uint32_t PLAN_RING_BYTES = PLAN_RING * PLAN_ROW * 16u + PLAN_RING * 8u + PLAN_RING * 4u;
__device__ __forceinline__ uint32_t ring_row(uint32_t ring, int b) { return ring + (uint32_t)b * (PLAN_ROW * 16u); }
__device__ __forceinline__ uint32_t ring_bar(uint32_t ring, int b) { return ring + PLAN_RING * PLAN_ROW * 16u + (uint32_t)b * 8u; }
__device__ __forceinline__ uint32_t ring_cnt(uint32_t ring, int b) { return ring + PLAN_RING * PLAN_ROW * 16u + PLAN_RING * 8u + (uint32_t)b * 4u; }
__device__ __forceinline__ void ring_fetch(uint32_t ring, int b, const uint4 *src) {       // one thread: 599 entries -> ring row b
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(ring_bar(ring, b)), "r"(599u * 16u) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(ring_row(ring, b)), "l"(src),
               "r"(599u * 16u), "r"(ring_bar(ring, b)) : "memory");
}
__device__ __forceinline__ int ring_count_in(uint32_t ring, int b) {                         // returns the count before this warp
  int old;
  // inc, not add: ptxas turns a predicated atom.add into a warp-aggregated one (VOTE, S2R lane id / lane mask, SHFL) although
  // only one lane runs it -- two S2R per symbol showed up as 6-7 % of the kernel's stall samples
  asm volatile("atom.shared.inc.u32 %0, [%1], 0x7fffffff;\n" : "=r"(old) : "r"(ring_cnt(ring, b)) : "memory");
  return old;
}
__device__ __forceinline__ void ring_count_reset(uint32_t ring, int b) {
  asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(ring_cnt(ring, b)), "r"(0) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}

// COMPACT: the tx-replicated outputs are written once, in rx's row layout: H_ls / H_mmse [B][nsym][nrx][PITCH] and
// tx [B][nsym][PITCH] (1 945 552 unique bytes per 4x4 slot instead of 3 756 928).
template <int T, int NTX, bool EST, int PITCH, bool STORE, bool COMPACT, bool BULK = false>
__device__ __forceinline__ void slot_body_wide(const SlotArgs &a, const SlotCtx &c, float2 (&st)[2][3]) {
  constexpr int NSC = 599, HALF = 300;
  const int nsym = a.g.nsym, nrx = a.g.nrx;
  const int t_ = threadIdx.x;
  const bool act = t_ < HALF, odd = t_ & 1;
  const int kp = HALF + t_, km = HALF - 1 - t_;                  // +f and -f bins of this lane
  const int K = act ? (odd ? km : kp) : 0;                      // kept (stored) bin: even
  const bool vS = act && !(odd && kp >= NSC);                   // lane 299's +f bin does not exist
  const int S = vS ? (odd ? kp : km) : 0;
  const float mS = vS ? 1.f : 0.f;

  const float2 *tw = reinterpret_cast<const float2 *>(a.prof.tap_tw) + (int64_t)c.m * MAXT * NSC;
  const float2 zero2 = make_float2(0.f, 0.f), neg1 = make_float2(-1.f, -1.f), nalpha = make_float2(-c.alpha, -c.alpha);
  constexpr bool FOLD = wide_fold(NTX);                          // error sums over all tx from the tx-summed CFR (below)
  float2 twp[T];
#pragma unroll
  for (int t = 0; t < T; ++t) twp[t] = act ? __ldg(tw + t * NSC + K) : zero2;

  const int64_t slot_h = (int64_t)nsym * nrx * NTX * PITCH, slot_r = (int64_t)nsym * nrx * PITCH;
  float2 *const Hb = STORE ? a.H_true + c.b * slot_h : nullptr;
  float2 *const Rb = STORE ? a.rx + c.b * slot_r : nullptr;
  constexpr int TXC = COMPACT ? 1 : NTX;                        // copies of the grid / of the estimates per (s, rx)
  float2 *const Tb = (STORE && c.rx == 0) ? a.tx + c.b * (int64_t)nsym * TXC * PITCH : nullptr;
  const uint4 *plan = EST ? reinterpret_cast<const uint4 *>(a.pat.plan) + (int64_t)c.pid * (nsym * NSC + 1) : nullptr;
  float2 *pH = Hb + (c.rx * NTX * PITCH + K), *pR = Rb + (c.rx * PITCH + K), *pT = Tb + K;
  // H_ls / H_mmse addresses are H_true's (compact: rx's) plus a uniform byte delta
  const char *const eb = COMPACT ? (const char *)Rb : (const char *)Hb;
  const int64_t slot_e = COMPACT ? slot_r : slot_h;
  const int64_t dL = (EST && STORE) ? (const char *)(a.H_ls + c.b * slot_e) - eb : 0;
  const bool mstore = EST && STORE && a.H_mmse != nullptr;     // dataset mode (H_true, rx, tx, H_ls) leaves H_mmse out
  const int64_t dM = mstore ? (const char *)(a.H_mmse + c.b * slot_e) - eb : 0;
  const int nre = nsym * NSC;
  int oPK = act ? K : nre, oPS = vS ? S : nre;                   // plan rows; row nre = "outside" for idle lanes
  const int dPK = act ? NSC : 0, dPS = vS ? NSC : 0;
  const int dH = nrx * NTX * PITCH, dR = nrx * PITCH, dT = TXC * PITCH;
  const float2 *gps = c.gsp;

  static_assert(SLOT_THREADS == RNG_LANES, "thread t draws Philox lane t");
  static_assert((PITCH & 1) == 0 && PITCH > NSC, "wide stores need an even, padded row pitch");
  uint4 ws = make_uint4(0, 0, 0, 0);
  // Plan entries are staged one symbol ahead with cp.async into two per-thread shared-memory slots: their L2 latency
  // (the kernel's top stall when they were plain loads behind an L1 prefetch) is spent under the previous symbol's work
  // and costs no registers.  A thread reads back only what it copied itself, so cp.async.wait_group is all the
  // synchronisation needed.  The last iteration stages the row after the pattern's plan: in bounds (padded pool).
  // shared-window address of this thread's staging slots: explicit ld.shared / cp.async on it (through the generic
  // pointer the compiler emitted generic LD.E.128, which pays the address-space check on every access)
  const uint32_t ps_s = (uint32_t)__cvta_generic_to_shared(c.pstage + t_);
  constexpr uint32_t PS_SLOT = SLOT_THREADS * sizeof(uint4);
  auto stage_plan = [&](int buf) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(ps_s + (buf * 2 + 0) * PS_SLOT), "l"(plan + oPK) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(ps_s + (buf * 2 + 1) * PS_SLOT), "l"(plan + oPS) : "memory");
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  auto staged = [&](int slot) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(ps_s + slot * PS_SLOT) : "memory");
    return v;
  };
  // BULK: plan rows arrive by cp.async.bulk in a ring of PLAN_RING rows (started by the kernel before the pilot phase).  Every
  // lane loads entry 300 + t and entry 299 - t (both conflict-free: 32 consecutive entries per warp) and takes K / S from
  // them by parity; entry 599 of a ring row is the all-outside entry (idle lanes, the missing bin).
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(c.pstage);
  const uint32_t oA = ((act && kp < NSC) ? kp : NSC) * 16u, oB = (act ? km : NSC) * 16u;
  auto ring_ld = [&](uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
  };
  if (EST && !BULK) stage_plan(0);
  auto xchg = [](float2 v) {     // value of this lane's S bin -> the neighbour that stores it as K+1
    return make_float2(__shfl_xor_sync(0xffffffffu, v.x, 1), __shfl_xor_sync(0xffffffffu, v.y, 1));
  };
  auto st16 = [](float2 *p, float2 lo, float2 hi) { __stcs(reinterpret_cast<float4 *>(p), make_float4(lo.x, lo.y, hi.x, hi.y)); };
  for (int s2 = 0; s2 < nsym; s2 += 2) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int s = s2 + j;
      float2 lK = zero2, lS = zero2;
      int ring_old = 0;
      if (EST && BULK) {
        const int rb = s & (PLAN_RING - 1);
        while (!mbar_try_wait(ring_bar(ring, rb), (uint32_t)(s / PLAN_RING) & 1u)) {}
        const uint4 eA = ring_ld(ring_row(ring, rb) + oA), eB = ring_ld(ring_row(ring, rb) + oB);
        if (s + PLAN_RING < nsym && (t_ & 31) == 0) ring_old = ring_count_in(ring, rb);      // see slot_body_wide2
        const uint4 eK = make_uint4(odd ? eB.x : eA.x, odd ? eB.y : eA.y, odd ? eB.z : eA.z, odd ? eB.w : eA.w);
        const uint4 eS = make_uint4(odd ? eA.x : eB.x, odd ? eA.y : eB.y, odd ? eA.z : eB.z, odd ? eA.w : eB.w);
        lK = plan_apply(plan_decode(eK), c.hp);
        lS = plan_apply(plan_decode(eS), c.hp);
      }
      if (EST && !BULK) {
        oPK += dPK;
        oPS += dPS;
        stage_plan(j ^ 1);                                   // next symbol's entries (s2 is even: buffer = s & 1 = j)
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        lK = plan_apply(plan_decode(staged(j * 2 + 0)), c.hp);
        lS = plan_apply(plan_decode(staged(j * 2 + 1)), c.hp);
      }
      if (EST && STORE) {
        // H_ls / H_mmse rows are the same for every tx: written here, so that only lK / lS stay live below
        const float2 lN = xchg(lS);
        const float2 mK = cscale(c.alpha, lK), mN = cscale(c.alpha, lN);
        if (act) {
          float2 *const pE = COMPACT ? pR : pH;
#pragma unroll
          for (int tx = 0; tx < TXC; ++tx) {
            st16((float2 *)((char *)(pE + tx * PITCH) + dL), lK, lN);
            if (mstore) st16((float2 *)((char *)(pE + tx * PITCH) + dM), mK, mN);
          }
        }
      }
      float2 hsK = zero2, hsS = zero2;
#pragma unroll
      for (int tx = 0; tx < NTX; ++tx) {
        const float2 *gp = gps + tx * MAXT;
        float2 A = zero2, B = zero2;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const float2 gq = gp[t];
          A = __ffma2_rn(make_float2(gq.x, gq.x), twp[t], A);
          B = __ffma2_rn(make_float2(gq.y, gq.y), twp[t], B);
        }
        const float2 hK = make_float2(A.x - B.y, A.y + B.x);                  // sum_t g_t tw_t(K)
        const float2 hS = cscale(mS, make_float2(A.x + B.y, B.x - A.y));      // mirror bin (0 where it does not exist)
        hsK = __fadd2_rn(hsK, hK);
        hsS = __fadd2_rn(hsS, hS);
        if (STORE) {
          const float2 hN = xchg(hS);
          if (act) st16(pH + tx * PITCH, hK, hN);
        }
        if (EST) {
          if (!FOLD) {
            float2 (&acc)[3] = st[tx == 0 ? 0 : 1];
            float2 d = __ffma2_rn(lK, neg1, hK);
            acc[0] = __ffma2_rn(d, d, acc[0]);
            d = __ffma2_rn(lS, neg1, hS);
            acc[0] = __ffma2_rn(d, d, acc[0]);
            d = __ffma2_rn(lK, nalpha, hK);
            acc[1] = __ffma2_rn(d, d, acc[1]);
            d = __ffma2_rn(lS, nalpha, hS);
            acc[1] = __ffma2_rn(d, d, acc[1]);
            acc[2] = __ffma2_rn(hK, hK, acc[2]);
            acc[2] = __ffma2_rn(hS, hS, acc[2]);
          } else {
            // folded form (ntx >= 4), see below: per tx only the power; tx 0 also its own cross term (pair (rx, 0))
            float2 &pw = st[tx == 0 ? 0 : 1][2];
            pw = __ffma2_rn(hK, hK, pw);
            pw = __ffma2_rn(hS, hS, pw);
            if (tx == 0) {
              st[0][0] = __ffma2_rn(lK, hK, st[0][0]);
              st[0][0] = __ffma2_rn(lS, hS, st[0][0]);
            }
          }
        }
      }
      if (EST && FOLD) {
        // sum_tx |h_tx - c l|^2 = sum_tx |h_tx|^2 - 2 c Re(conj(l) sum_tx h_tx) + ntx c^2 |l|^2   (c = 1: LS, c = alpha: MMSE).
        // l is the same for every tx, so the error sums over ALL tx need only the power, the cross term with the tx-SUMMED
        // CFR (hsK / hsS, which the received grid needs anyway) and |l|^2: 2 packed ops per tx + 6 per symbol instead
        // of 12 per tx.  The kernel accumulates the three raw sums; the flush combines them in double.  For ntx > 1 the
        // result is of the order of the power itself (the LS estimate is the tx-SUM, src/channel_simulator.py:402-404),
        // so nothing cancels.  Accumulators: st[0][0] = sum Re(conj(l) h_0) (packed re / im products), st[0][2] = |h_0|^2,
        // st[1][0] = sum Re(conj(l) sum_tx h), st[1][1] = sum |l|^2, st[1][2] = sum_{tx >= 1} |h_tx|^2.
        st[1][0] = __ffma2_rn(lK, hsK, st[1][0]);
        st[1][0] = __ffma2_rn(lS, hsS, st[1][0]);
        st[1][1] = __ffma2_rn(lK, lK, st[1][1]);
        st[1][1] = __ffma2_rn(lS, lS, st[1][1]);
      }
      if (EST && BULK && s + PLAN_RING < nsym && (t_ & 31) == 0 && ring_old == SLOT_THREADS / 32 - 1) {
        // last warp out of this ring slot: refill it with the row PLAN_RING symbols ahead
        const int rb = s & (PLAN_RING - 1);
        ring_count_reset(ring, rb);
        ring_fetch(ring, rb, plan + (s + PLAN_RING) * NSC);
      }
      if (!STORE) {
        gps += NTX * MAXT;
        continue;
      }
      // ---- draws: Philox lane t serves both bins of the mirror pair; word half h = (f > 0) -------------
      if (j == 0) ws = draw(c.key, STREAM_SYMBOLS, (uint32_t)((s2 >> 1) * RNG_LANES + t_));
      const uint32_t wm = j ? ws.z : ws.x, wp = j ? ws.w : ws.y;               // -f, +f
      const float2 xK = cis_u01(((odd ? wm : wp) & a.sym_and) | a.sym_or), xS = cis_u01(((odd ? wp : wm) & a.sym_and) | a.sym_or);
      const uint4 wn = draw(c.key, STREAM_NOISE, (uint32_t)((s * nrx + c.rx) * RNG_LANES + t_));
      const float2 nK = normal_pair(odd ? wn.x : wn.z, odd ? wn.y : wn.w);
      const float2 nS = normal_pair(odd ? wn.z : wn.x, odd ? wn.w : wn.y);
      const float2 yk = cmul(hsK, xK), ys = cmul(hsS, xS);
      const float2 yK = make_float2(fmaf(c.sigma, nK.x, yk.x), fmaf(c.sigma, nK.y, yk.y));
      const float2 yN = xchg(make_float2(fmaf(c.sigma, nS.x, ys.x), fmaf(c.sigma, nS.y, ys.y)));
      if (act) st16(pR, yK, yN);
      if (Tb) {   // the rx-0 CTA writes the (tx-replicated) grid
        const float2 xN = xchg(xS);
        if (act) {
#pragma unroll
          for (int tx = 0; tx < TXC; ++tx) st16(pT + tx * PITCH, xK, xN);
        }
      }
      pH += dH;
      pR += dR;
      pT += dT;
      gps += NTX * MAXT;
    }
  }
}

// One thread: barriers, counters, the all-outside entry of every ring row, and the first PLAN_RING plan rows on their way.
__device__ __forceinline__ void plan_ring_start(const SlotArgs &a, const SlotCtx &c) {
  const int nsym = a.g.nsym;
  const uint4 *plan = reinterpret_cast<const uint4 *>(a.pat.plan) + (int64_t)c.pid * (nsym * 599 + 1);
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(c.pstage);
  const uint4 outside = __ldg(plan + nsym * 599);
#pragma unroll
  for (int b = 0; b < PLAN_RING; ++b) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(ring_bar(ring, b)));
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"r"(ring_row(ring, b) + 599u * 16u), "r"(outside.x), "r"(outside.y), "r"(outside.z),
                 "r"(outside.w) : "memory");
    ring_count_reset(ring, b);
  }
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
#pragma unroll
  for (int b = 0; b < PLAN_RING; ++b)
    if (b < nsym) ring_fetch(ring, b, plan + b * 599);
}

template <int NTX, bool EXACT, bool EST, int NSC, bool FAST, int WIDE, bool STORE = true, bool COMPACT = false, bool BULK = false>
__global__ void __launch_bounds__(SLOT_THREADS, 2) slot_kernel(SlotArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nsc = NSC ? NSC : a.g.nsc;
  const int nsym = a.g.nsym, ntx = a.g.ntx, nrx = a.g.nrx;
  float2 *gsp = reinterpret_cast<float2 *>(smem_raw);                 // [nsym][ntx][MAXT]
  float2 *gs = reinterpret_cast<float2 *>(gsp + nsym * ntx * MAXT);   // [nsym][MAXT] sum over tx
  float2 *hp = gs + nsym * MAXT;                                      // [np_max + 1], last = 0
  __shared__ float red[33];
  __shared__ float ssm[SLOT_THREADS / 32][6];

  SlotCtx c;
  c.b = blockIdx.x / nrx;
  c.rx = blockIdx.x - (int)c.b * nrx;
  c.m = a.slots.model_id[c.b];
  c.ntaps = a.prof.ntaps[c.m];
  c.sigma = a.noise_std[c.b];
  c.key = make_key(a.slots.seed, a.slots.slot0 + c.b);
  c.gsp = gsp;
  c.hp = hp;
  c.pstage = reinterpret_cast<uint4 *>((reinterpret_cast<uintptr_t>(hp + (EST ? a.pat.np_max + 1 : 0)) + 15) & ~(uintptr_t)15);
  c.alpha = 0.f;
  c.pid = EST ? a.slots.pattern_id[c.b] : 0;
  if (WIDE != 0 && EST && BULK && threadIdx.x == 0) plan_ring_start(a, c);

  const float2 *gin = a.gains + (c.b * nrx + c.rx) * (int64_t)(nsym * ntx * MAXT);
  for (int i = threadIdx.x; i < nsym * ntx * MAXT; i += SLOT_THREADS) {
    gsp[i] = __ldg(gin + i);
  }
  __syncthreads();

  if (EST) {
    // LS at the pilots: h_p = y_p / (x_p + 1e-12)  (src/baseline_estimators.py:109-110), with y_p
    // evaluated directly at the pilot REs from the tx-summed gains.
    for (int i = threadIdx.x; i < nsym * MAXT; i += SLOT_THREADS) {
      int s = i / MAXT, t = i - s * MAXT;
      float sr = 0.f, si = 0.f;
      for (int tx = 0; tx < ntx; ++tx) {
        float2 v = gsp[(s * ntx + tx) * MAXT + t];
        sr += v.x;
        si += v.y;
      }
      gs[i] = make_float2(sr, si);
    }
    __syncthreads();
  }

  float2 st[2][3];   // packed (re^2, im^2) sums: [0] tx 0 (antenna pair (rx, 0)), [1] tx >= 1
#pragma unroll
  for (int q = 0; q < 2; ++q) st[q][0] = st[q][1] = st[q][2] = make_float2(0.f, 0.f);

  if (c.ntaps <= 5) {
    if (EST) pilot_phase<5, NSC>(a, c, gs, hp, red);
    if constexpr (WIDE) slot_body_wide<5, NTX, EST, WIDE, STORE, COMPACT, BULK>(a, c, st);
    else slot_body<5, NTX, EXACT, EST, NSC, FAST>(a, c, st);
  }
  else if (c.ntaps <= 8) {
    if (EST) pilot_phase<8, NSC>(a, c, gs, hp, red);
    if constexpr (WIDE) slot_body_wide<8, NTX, EST, WIDE, STORE, COMPACT, BULK>(a, c, st);
    else slot_body<8, NTX, EXACT, EST, NSC, FAST>(a, c, st);
  }
  else if (c.ntaps <= 9) {
    if (EST) pilot_phase<9, NSC>(a, c, gs, hp, red);
    if constexpr (WIDE) slot_body_wide<9, NTX, EST, WIDE, STORE, COMPACT, BULK>(a, c, st);
    else slot_body<9, NTX, EXACT, EST, NSC, FAST>(a, c, st);
  }
  else {
    if (EST) pilot_phase<MAXT, NSC>(a, c, gs, hp, red);
    if constexpr (WIDE) slot_body_wide<MAXT, NTX, EST, WIDE, STORE, COMPACT, BULK>(a, c, st);
    else slot_body<MAXT, NTX, EXACT, EST, NSC, FAST>(a, c, st);
  }

  if (EST && a.stats) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        // [0] = pair (rx, 0); [1] = all tx of this rx = tx 0 + the rest
        float v = st[0][j].x + st[0][j].y;
        if (q == 1) v += st[1][j].x + st[1][j].y;
        if constexpr (WIDE != 0 && wide_fold(NTX)) v = st[q][j].x + st[q][j].y;      // raw sums, combined below
        v = warp_sum(v);
        if (lane == 0) ssm[warp][q * 3 + j] = v;
      }
    __syncthreads();
    if constexpr (WIDE != 0 && wide_fold(NTX)) {
      // folded form of slot_body_wide: [0][0] = T0 = sum Re(conj(l) h_0), [0][2] = P0 = sum |h_0|^2, [1][0] = T = sum
      // Re(conj(l) sum_tx h), [1][1] = U = sum |l|^2, [1][2] = P1 = sum_{tx >= 1} |h|^2.
      //   pair (rx, 0): e_ls = P0 - 2 T0 + U          e_mmse = P0 - 2 alpha T0 + alpha^2 U            power = P0
      //   all tx      : e_ls = P - 2 T + ntx U        e_mmse = P - 2 alpha T + ntx alpha^2 U          power = P = P0 + P1
      if (threadIdx.x < 6) {
        double r[6];
        for (int i = 0; i < 6; ++i) {
          double acc = 0.0;
          for (int w = 0; w < SLOT_THREADS / 32; ++w) acc += (double)ssm[w][i];
          r[i] = acc;
        }
        const double T0 = r[0], P0 = r[2], T = r[3], U = r[4], P = r[2] + r[5], al = (double)c.alpha, n = (double)NTX;
        const int q = threadIdx.x / 3, j = threadIdx.x - 3 * q;
        const double cc = j == 0 ? 1.0 : al;
        double v;
        if (j == 2) v = q ? P : P0;
        else v = q ? P - 2.0 * cc * T + n * cc * cc * U : P0 - 2.0 * cc * T0 + cc * cc * U;
        a.stats[(c.b * nrx + c.rx) * 6 + threadIdx.x] = v;
      }
    } else if (threadIdx.x < 6) {
      double acc = 0.0;
      for (int w = 0; w < SLOT_THREADS / 32; ++w) acc += (double)ssm[w][threadIdx.x];
      a.stats[(c.b * nrx + c.rx) * 6 + threadIdx.x] = acc;
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------------
// Register-blocked form of the wide kernel: 160 threads per CTA, thread t owns TWO mirror pairs, the adjacent frequencies
// f1 = 2t+1 and f2 = 2t+2.  On the +f side that is the adjacent bin pair (300+2t, 301+2t), on the -f side (298-2t, 299-2t):
// both 16-byte aligned in a row of pitch 600, so every row is written with two 16-byte stores per thread and NO lane-pair
// exchange (the 320-thread form needs 12 shuffles + selects per symbol for that); a tap gain fetched from shared memory
// now feeds four packed FMAs instead of two (half the shared-memory wavefronts per bin) and the per-thread loop overhead
// is spread over four bins.  Same Philox counters (bin +-f draws from lane f-1), same operation order per bin:
// bit-identical to slot_body_wide.  t = 149: f2 = 300 has no +f bin (element 599 is the padding element, written as 0).
constexpr int SLOT2_THREADS = 160;
// BULK: the plan rows of a symbol (599 contiguous 16-byte entries) are brought into a two-row shared-memory ring by
// cp.async.bulk (one instruction of one thread per symbol, completion on an mbarrier) instead of one cp.async per entry
// and thread: the copy engine writes shared memory without passing through the LSU data pipe, which was the first limiter
// of the statistics kernels (ncu: 83-92 % busy, about a third of its wavefronts the per-thread staging copies).
template <int T, int NTX, bool EST, int PITCH, bool STORE, bool COMPACT, bool BULK = false>
__device__ __forceinline__ void slot_body_wide2(const SlotArgs &a, const SlotCtx &c, float2 (&st)[2][3]) {
  constexpr int NSC = 599, HALF = 300;
  const int nsym = a.g.nsym, nrx = a.g.nrx;
  const int t_ = threadIdx.x;
  const bool act = t_ < HALF / 2;
  const int l1 = 2 * t_, l2 = l1 + 1;                            // Philox lanes of f1, f2 (lane = f - 1)
  const int kp = act ? HALF + l1 : 0;                            // +f1 bin (even); +f2 = kp + 1
  const int km = act ? HALF - 2 - l1 : 0;                        // -f2 bin (even); -f1 = km + 1
  const bool vp2 = act && kp + 1 < NSC;                          // +f2 of the last thread is the padding element
  const float mp2 = vp2 ? 1.f : 0.f;

  const float2 *tw = reinterpret_cast<const float2 *>(a.prof.tap_tw) + (int64_t)c.m * MAXT * NSC;
  const float2 zero2 = make_float2(0.f, 0.f), neg1 = make_float2(-1.f, -1.f), nalpha = make_float2(-c.alpha, -c.alpha);
  constexpr bool FOLD = wide_fold(NTX);
  float2 tw1[T], tw2[T];                                          // twiddles of +f1, +f2 (conjugates serve -f1, -f2)
#pragma unroll
  for (int t = 0; t < T; ++t) {
    tw1[t] = act ? __ldg(tw + t * NSC + kp) : zero2;
    float2 w = zero2;
    if (vp2) w = __ldg(tw + t * NSC + kp + 1);
    else if (act) {                  // t = 149: -f2 (bin 0) exists although +f2 does not: conjugate of bin 0's own entry
      w = __ldg(tw + t * NSC + km);
      w.y = -w.y;
    }
    tw2[t] = w;
  }

  const int64_t slot_h = (int64_t)nsym * nrx * NTX * PITCH, slot_r = (int64_t)nsym * nrx * PITCH;
  constexpr int TXC = COMPACT ? 1 : NTX;
  float2 *const Hb = STORE ? a.H_true + c.b * slot_h : nullptr;
  float2 *const Rb = STORE ? a.rx + c.b * slot_r : nullptr;
  float2 *const Tb = (STORE && c.rx == 0) ? a.tx + c.b * (int64_t)nsym * TXC * PITCH : nullptr;
  const uint4 *plan = EST ? reinterpret_cast<const uint4 *>(a.pat.plan) + (int64_t)c.pid * (nsym * NSC + 1) : nullptr;
  float2 *pH = Hb + c.rx * NTX * PITCH, *pR = Rb + c.rx * PITCH, *pT = Tb;
  const char *const eb = COMPACT ? (const char *)Rb : (const char *)Hb;
  const int64_t slot_e = COMPACT ? slot_r : slot_h;
  const int64_t dL = (EST && STORE) ? (const char *)(a.H_ls + c.b * slot_e) - eb : 0;
  const bool mstore = EST && STORE && a.H_mmse != nullptr;
  const int64_t dM = mstore ? (const char *)(a.H_mmse + c.b * slot_e) - eb : 0;
  const int nre = nsym * NSC;
  // plan rows of the four bins: +f1, +f2, -f2, -f1 (row nre = the all-outside entry for idle lanes / the missing bin)
  int oP[4] = {act ? kp : nre, vp2 ? kp + 1 : nre, act ? km : nre, act ? km + 1 : nre};
  const int dP[4] = {act ? NSC : 0, vp2 ? NSC : 0, act ? NSC : 0, act ? NSC : 0};
  const int dH = nrx * NTX * PITCH, dR = nrx * PITCH, dT = TXC * PITCH;
  const float2 *gps = c.gsp;

  const uint32_t ps_s = (uint32_t)__cvta_generic_to_shared(c.pstage + t_);
  constexpr uint32_t PS_SLOT = SLOT2_THREADS * sizeof(uint4);
  auto stage_plan = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(ps_s + (buf * 4 + q) * PS_SLOT), "l"(plan + oP[q]) : "memory");
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  auto staged = [&](int slot) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(ps_s + slot * PS_SLOT) : "memory");
    return v;
  };
  auto st16 = [](float2 *p, float2 lo, float2 hi) { __stcs(reinterpret_cast<float4 *>(p), make_float4(lo.x, lo.y, hi.x, hi.y)); };
  // bulk form: ring of PLAN_RING rows of PLAN_ROW entries; entry 599 of each row is the all-outside entry (idle lanes / the
  // missing bin); this thread's four entries sit at fixed positions of a row
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(c.pstage);
  // thread t reads the adjacent entries (kp, kp + 1) and (km, km + 1): with every thread loading the lower entry first a
  // quarter-warp would touch 8 entries 32 bytes apart (two per bank group); threads 4..7 of every eight load the UPPER
  // entry first instead, which makes the eight 16-byte accesses of a quarter-warp fall into eight different bank groups
  const bool swp = (t_ >> 2) & 1;
  const uint32_t eraw[4] = {(act ? kp : NSC) * 16u, (vp2 ? kp + 1 : NSC) * 16u, (act ? km : NSC) * 16u, (act ? km + 1 : NSC) * 16u};
  const uint32_t eo[4] = {swp ? eraw[1] : eraw[0], swp ? eraw[0] : eraw[1], swp ? eraw[3] : eraw[2], swp ? eraw[2] : eraw[3]};
  auto ring_entry = [&](int buf, int q) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(ring_row(ring, buf) + eo[q]) : "memory");
    return v;
  };
  // (bulk form: the kernel has initialised the barriers and started rows 0 and 1 before the pilot phase)
  if (EST && !BULK) stage_plan(0);
  uint4 ws1 = make_uint4(0, 0, 0, 0), ws2 = ws1;
  for (int s2 = 0; s2 < nsym; s2 += 2) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int s = s2 + j;
      float2 lp1 = zero2, lp2 = zero2, lm2 = zero2, lm1 = zero2;      // LS estimates at +f1, +f2, -f2, -f1
      int ring_old = 0;
      if (EST && BULK) {
        // row s sits in ring[s % PLAN_RING], its (s / PLAN_RING)-th fill
        const int rb = s & (PLAN_RING - 1);
        while (!mbar_try_wait(ring_bar(ring, rb), (uint32_t)(s / PLAN_RING) & 1u)) {}
        const uint4 r0 = ring_entry(rb, 0), r1 = ring_entry(rb, 1), r2 = ring_entry(rb, 2), r3 = ring_entry(rb, 3);
        auto sel4 = [](bool c, uint4 x, uint4 y) { return make_uint4(c ? x.x : y.x, c ? x.y : y.y, c ? x.z : y.z, c ? x.w : y.w); };
        const uint4 e0 = sel4(swp, r1, r0), e1 = sel4(swp, r0, r1), e2 = sel4(swp, r3, r2), e3 = sel4(swp, r2, r3);
        // the ring slot is free once every warp holds its entries: the warps count themselves off on a shared-memory counter
        // and whichever comes last refills the slot with row s + PLAN_RING at the END of its symbol (the atomic's return
        // value is not needed before, so nobody waits for anybody; the copy has PLAN_RING - 1 symbols of work to land under).
        // A warp cannot count itself twice on one slot: its next visit (symbol s + PLAN_RING) waits for this very refill.
        if (s + PLAN_RING < nsym && (t_ & 31) == 0) ring_old = ring_count_in(ring, rb);
        lp1 = plan_apply(plan_decode(e0), c.hp);
        lp2 = plan_apply(plan_decode(e1), c.hp);
        lm2 = plan_apply(plan_decode(e2), c.hp);
        lm1 = plan_apply(plan_decode(e3), c.hp);
      }
      if (EST && !BULK) {
#pragma unroll
        for (int q = 0; q < 4; ++q) oP[q] += dP[q];
        stage_plan(j ^ 1);
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        lp1 = plan_apply(plan_decode(staged(j * 4 + 0)), c.hp);
        lp2 = plan_apply(plan_decode(staged(j * 4 + 1)), c.hp);
        lm2 = plan_apply(plan_decode(staged(j * 4 + 2)), c.hp);
        lm1 = plan_apply(plan_decode(staged(j * 4 + 3)), c.hp);
      }
      if (EST && STORE && act) {
        float2 *const pE = COMPACT ? pR : pH;
#pragma unroll
        for (int tx = 0; tx < TXC; ++tx) {
          char *q = (char *)(pE + tx * PITCH) + dL;
          st16((float2 *)q + kp, lp1, lp2);
          st16((float2 *)q + km, lm2, lm1);
          if (mstore) {
            char *qm = (char *)(pE + tx * PITCH) + dM;
            st16((float2 *)qm + kp, cscale(c.alpha, lp1), cscale(c.alpha, lp2));
            st16((float2 *)qm + km, cscale(c.alpha, lm2), cscale(c.alpha, lm1));
          }
        }
      }
      float2 sp1 = zero2, sp2 = zero2, sm2 = zero2, sm1 = zero2;      // sum over tx of the CFR at the four bins
#pragma unroll
      for (int tx = 0; tx < NTX; ++tx) {
        const float2 *gp = gps + tx * MAXT;
        float2 A1 = zero2, B1 = zero2, A2 = zero2, B2 = zero2;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const float2 gq = gp[t];
          const float2 gx = make_float2(gq.x, gq.x), gy = make_float2(gq.y, gq.y);
          A1 = __ffma2_rn(gx, tw1[t], A1);
          B1 = __ffma2_rn(gy, tw1[t], B1);
          A2 = __ffma2_rn(gx, tw2[t], A2);
          B2 = __ffma2_rn(gy, tw2[t], B2);
        }
        const float2 hp1 = make_float2(A1.x - B1.y, A1.y + B1.x);                // +f1
        const float2 hm1 = make_float2(A1.x + B1.y, B1.x - A1.y);                // -f1
        const float2 hp2 = cscale(mp2, make_float2(A2.x - B2.y, A2.y + B2.x));   // +f2 (0 where it does not exist)
        const float2 hm2 = make_float2(A2.x + B2.y, B2.x - A2.y);                // -f2
        sp1 = __fadd2_rn(sp1, hp1);
        sp2 = __fadd2_rn(sp2, hp2);
        sm2 = __fadd2_rn(sm2, hm2);
        sm1 = __fadd2_rn(sm1, hm1);
        if (STORE && act) {
          st16(pH + tx * PITCH + kp, hp1, hp2);
          st16(pH + tx * PITCH + km, hm2, hm1);
        }
        if (EST) {
          if (!FOLD) {
            float2 (&acc)[3] = st[tx == 0 ? 0 : 1];
            const float2 hh[4] = {hp1, hp2, hm2, hm1}, ll[4] = {lp1, lp2, lm2, lm1};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float2 d = __ffma2_rn(ll[q], neg1, hh[q]);
              acc[0] = __ffma2_rn(d, d, acc[0]);
              d = __ffma2_rn(ll[q], nalpha, hh[q]);
              acc[1] = __ffma2_rn(d, d, acc[1]);
              acc[2] = __ffma2_rn(hh[q], hh[q], acc[2]);
            }
          } else {
            float2 &pw = st[tx == 0 ? 0 : 1][2];
            pw = __ffma2_rn(hp1, hp1, pw);
            pw = __ffma2_rn(hp2, hp2, pw);
            pw = __ffma2_rn(hm2, hm2, pw);
            pw = __ffma2_rn(hm1, hm1, pw);
            if (tx == 0) {
              st[0][0] = __ffma2_rn(lp1, hp1, st[0][0]);
              st[0][0] = __ffma2_rn(lp2, hp2, st[0][0]);
              st[0][0] = __ffma2_rn(lm2, hm2, st[0][0]);
              st[0][0] = __ffma2_rn(lm1, hm1, st[0][0]);
            }
          }
        }
      }
      if (EST && FOLD) {
        st[1][0] = __ffma2_rn(lp1, sp1, st[1][0]);
        st[1][0] = __ffma2_rn(lp2, sp2, st[1][0]);
        st[1][0] = __ffma2_rn(lm2, sm2, st[1][0]);
        st[1][0] = __ffma2_rn(lm1, sm1, st[1][0]);
        st[1][1] = __ffma2_rn(lp1, lp1, st[1][1]);
        st[1][1] = __ffma2_rn(lp2, lp2, st[1][1]);
        st[1][1] = __ffma2_rn(lm2, lm2, st[1][1]);
        st[1][1] = __ffma2_rn(lm1, lm1, st[1][1]);
      }
      if (EST && BULK && s + PLAN_RING < nsym && (t_ & 31) == 0 && ring_old == SLOT2_THREADS / 32 - 1) {
        ring_count_reset(ring, s & (PLAN_RING - 1));
        ring_fetch(ring, s & (PLAN_RING - 1), plan + (s + PLAN_RING) * NSC);
      }
      if (!STORE) {
        gps += NTX * MAXT;
        continue;
      }
      // ---- draws: bin +-f comes from Philox lane f - 1; word half h = (f > 0) ---------------------------------------
      if (j == 0) {
        ws1 = draw(c.key, STREAM_SYMBOLS, (uint32_t)((s2 >> 1) * RNG_LANES + l1));
        ws2 = draw(c.key, STREAM_SYMBOLS, (uint32_t)((s2 >> 1) * RNG_LANES + l2));
      }
      const float2 xm1 = cis_u01(((j ? ws1.z : ws1.x) & a.sym_and) | a.sym_or), xp1 = cis_u01(((j ? ws1.w : ws1.y) & a.sym_and) | a.sym_or);
      const float2 xm2 = cis_u01(((j ? ws2.z : ws2.x) & a.sym_and) | a.sym_or), xp2 = cis_u01(((j ? ws2.w : ws2.y) & a.sym_and) | a.sym_or);
      const uint4 wn1 = draw(c.key, STREAM_NOISE, (uint32_t)((s * nrx + c.rx) * RNG_LANES + l1));
      const uint4 wn2 = draw(c.key, STREAM_NOISE, (uint32_t)((s * nrx + c.rx) * RNG_LANES + l2));
      const float2 nm1 = normal_pair(wn1.x, wn1.y), np1 = normal_pair(wn1.z, wn1.w);
      const float2 nm2 = normal_pair(wn2.x, wn2.y), np2 = normal_pair(wn2.z, wn2.w);
      auto rxv = [&](float2 hs, float2 x, float2 n) {
        const float2 y = cmul(hs, x);
        return make_float2(fmaf(c.sigma, n.x, y.x), fmaf(c.sigma, n.y, y.y));
      };
      if (act) {
        st16(pR + kp, rxv(sp1, xp1, np1), rxv(sp2, xp2, np2));
        st16(pR + km, rxv(sm2, xm2, nm2), rxv(sm1, xm1, nm1));
        if (Tb) {
#pragma unroll
          for (int tx = 0; tx < TXC; ++tx) {
            st16(pT + tx * PITCH + kp, xp1, xp2);
            st16(pT + tx * PITCH + km, xm2, xm1);
          }
        }
      }
      pH += dH;
      pR += dR;
      pT += dT;
      gps += NTX * MAXT;
    }
  }
}

// SCORE (dense statistics): the second pass of the array-free dense-Wiener pipeline.  The pilot phase is replaced by
// loading this (slot, rx)'s FILTERED pilot vector (row hp_col[b] + rx of a.hp_out, here an input) into shared memory;
// the body regenerates the true CFR from the tap gains, interpolates the filtered pilots and takes the error sums, of
// which only the MMSE fields (stats[..., 1]) are written -- no resource-grid array ever reaches HBM.
template <int NTX, bool EST, bool STORE, bool COMPACT, bool SCORE = false, bool BULK = false>
__global__ void __launch_bounds__(SLOT2_THREADS, 3) slot2_kernel(const __grid_constant__ SlotArgs a) {
  constexpr int WIDE = WIDE_PITCH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nsym = a.g.nsym, ntx = a.g.ntx, nrx = a.g.nrx;
  float2 *gsp = reinterpret_cast<float2 *>(smem_raw);                 // [nsym][ntx][MAXT]
  float2 *gs = reinterpret_cast<float2 *>(gsp + nsym * ntx * MAXT);   // [nsym][MAXT] sum over tx
  float2 *hp = gs + nsym * MAXT;                                      // [np_max + 1], last = 0
  __shared__ float red[33];
  __shared__ float ssm[SLOT2_THREADS / 32][6];

  SlotCtx c;
  c.b = blockIdx.x / nrx;
  c.rx = blockIdx.x - (int)c.b * nrx;
  c.m = a.slots.model_id[c.b];
  c.ntaps = a.prof.ntaps[c.m];
  c.sigma = SCORE ? 0.f : a.noise_std[c.b];
  c.key = make_key(a.slots.seed, a.slots.slot0 + c.b);
  c.gsp = gsp;
  c.hp = hp;
  c.pstage = reinterpret_cast<uint4 *>((reinterpret_cast<uintptr_t>(hp + (EST ? a.pat.np_max + 1 : 0)) + 15) & ~(uintptr_t)15);
  c.alpha = 0.f;
  c.pid = 0;

  const float2 *gin = a.gains + (c.b * nrx + c.rx) * (int64_t)(nsym * ntx * MAXT);
  if (EST) c.pid = a.slots.pattern_id[c.b];
  // plan ring: barriers, the all-outside entries and the first rows on their way before anything else happens (they land
  // under the gain staging and the pilot phase)
  if (EST && BULK && threadIdx.x == 0) plan_ring_start(a, c);
  for (int i = threadIdx.x; i < nsym * ntx * MAXT; i += SLOT2_THREADS) gsp[i] = __ldg(gin + i);
  if (SCORE) {
    c.alpha = 1.f;                     // the filtered pilots are interpolated as they are
    // the whole row (np_max entries; ld >= np_max): the plan only refers to the pattern's first npilots entries, and not
    // waiting for npilots[pid] takes one dependent round trip out of the prologue
    const int np = a.pat.np_max;
    const float2 *hm = a.hp_out + ((a.hp_col ? (int64_t)a.hp_col[c.b] : c.b * nrx) + c.rx) * a.hp_ld;
    for (int i = threadIdx.x; i < np; i += SLOT2_THREADS) hp[i] = __ldg(hm + i);
    if (threadIdx.x == 0) hp[a.pat.np_max] = make_float2(0.f, 0.f);
  }
  __syncthreads();
  if (EST && !SCORE) {
    for (int i = threadIdx.x; i < nsym * MAXT; i += SLOT2_THREADS) {
      int s = i / MAXT, t = i - s * MAXT;
      float sr = 0.f, si = 0.f;
      for (int tx = 0; tx < ntx; ++tx) {
        float2 v = gsp[(s * ntx + tx) * MAXT + t];
        sr += v.x;
        si += v.y;
      }
      gs[i] = make_float2(sr, si);
    }
    __syncthreads();
  }
  float2 st[2][3];
#pragma unroll
  for (int q = 0; q < 2; ++q) st[q][0] = st[q][1] = st[q][2] = make_float2(0.f, 0.f);

  if (c.ntaps <= 5) {
    if (EST && !SCORE) pilot_phase<5, 599, SLOT2_THREADS>(a, c, gs, hp, red);
    slot_body_wide2<5, NTX, EST, WIDE, STORE, COMPACT, BULK>(a, c, st);
  } else if (c.ntaps <= 8) {
    if (EST && !SCORE) pilot_phase<8, 599, SLOT2_THREADS>(a, c, gs, hp, red);
    slot_body_wide2<8, NTX, EST, WIDE, STORE, COMPACT, BULK>(a, c, st);
  } else if (c.ntaps <= 9) {
    if (EST && !SCORE) pilot_phase<9, 599, SLOT2_THREADS>(a, c, gs, hp, red);
    slot_body_wide2<9, NTX, EST, WIDE, STORE, COMPACT, BULK>(a, c, st);
  } else {
    if (EST && !SCORE) pilot_phase<MAXT, 599, SLOT2_THREADS>(a, c, gs, hp, red);
    slot_body_wide2<MAXT, NTX, EST, WIDE, STORE, COMPACT, BULK>(a, c, st);
  }

  if (EST && a.stats) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float v = st[0][j].x + st[0][j].y;
        if (q == 1) v += st[1][j].x + st[1][j].y;
        if constexpr (wide_fold(NTX)) v = st[q][j].x + st[q][j].y;
        v = warp_sum(v);
        if (lane == 0) ssm[warp][q * 3 + j] = v;
      }
    __syncthreads();
    if constexpr (wide_fold(NTX)) {
      if (threadIdx.x < 6) {
        double r[6];
        for (int i = 0; i < 6; ++i) {
          double acc = 0.0;
          for (int w = 0; w < SLOT2_THREADS / 32; ++w) acc += (double)ssm[w][i];
          r[i] = acc;
        }
        const double T0 = r[0], P0 = r[2], T = r[3], U = r[4], P = r[2] + r[5], al = (double)c.alpha, n = (double)NTX;
        const int q = threadIdx.x / 3, j = threadIdx.x - 3 * q;
        const double cc = j == 0 ? 1.0 : al;
        double v;
        if (j == 2) v = q ? P : P0;
        else v = q ? P - 2.0 * cc * T + n * cc * cc * U : P0 - 2.0 * cc * T0 + cc * cc * U;
        if (!SCORE || j == 1) a.stats[(c.b * nrx + c.rx) * 6 + threadIdx.x] = v;
      }
    } else if (threadIdx.x < 6) {
      double acc = 0.0;
      for (int w = 0; w < SLOT2_THREADS / 32; ++w) acc += (double)ssm[w][threadIdx.x];
      if (!SCORE || threadIdx.x % 3 == 1) a.stats[(c.b * nrx + c.rx) * 6 + threadIdx.x] = acc;
    }
  }
}

template <int NTX, bool EST, bool STORE, bool COMPACT, bool SCORE = false, bool BULK = false>
static int launch_slot2_form(const SlotArgs &a, int64_t B, size_t smem, cudaStream_t stream) {
  // plan staging: 2 buffers x 4 entries per thread, or (bulk) a ring of PLAN_RING rows of PLAN_ROW entries
  if (EST) smem += 16 + (BULK ? PLAN_RING_BYTES : 8 * SLOT2_THREADS * sizeof(uint4));
  if (smem > 48 * 1024) B2C_CUDA((set_max_smem<slot2_kernel<NTX, EST, STORE, COMPACT, SCORE, BULK>>(smem)));
  slot2_kernel<NTX, EST, STORE, COMPACT, SCORE, BULK><<<(unsigned)(B * a.g.nrx), SLOT2_THREADS, smem, stream>>>(a);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}
template <int NTX, bool EST, bool STORE, bool COMPACT, bool SCORE = false>
static int launch_slot2(const SlotArgs &a, int64_t B, size_t smem, cudaStream_t stream) {
  // Plan rows by bulk copy into the shared-memory ring, measured against the per-thread cp.async staging (2x2, 18 944 slots):
  // scoring pass 1.10 ms against 1.35 ms (its L1 data pipe was 92 % busy), first pass 1.74 against 1.81 ms.  The 4x4 first pass
  // spends its time in the pilot phase and the tx loop, latency-bound at 15 warps per SM: 0.883 against 0.875 ms per 4144
  // slots, so ntx >= 4 keeps the per-thread form there.  B2C_PLAN_BULK=1 / 0 forces either form (tests, A/B runs).
  const char *e = getenv("B2C_PLAN_BULK");
  const bool bulk = e ? e[0] != '0' : (SCORE || NTX <= 2);
  if (EST && !STORE && bulk) return launch_slot2_form<NTX, EST, STORE, COMPACT, SCORE, true>(a, B, smem, stream);
  return launch_slot2_form<NTX, EST, STORE, COMPACT, SCORE, false>(a, B, smem, stream);
}

static size_t slot_smem_bytes(const b2c_geom *g, int np_max) {
  return (size_t)g->nsym * g->ntx * MAXT * sizeof(float2) + (size_t)g->nsym * MAXT * sizeof(float2) +
         (size_t)(np_max + 1) * sizeof(float2);
}

template <int NTX, bool EXACT, bool EST, int NSC, bool FAST, int WIDE = 0, bool STORE = true, bool COMPACT = false, bool BULK = false>
static int launch_slot_form(const SlotArgs &a, int64_t B, size_t smem, cudaStream_t stream) {
  auto kern = slot_kernel<NTX, EXACT, EST, NSC, FAST, WIDE, STORE, COMPACT, BULK>;
  // plan-entry staging (+ alignment slack): per-thread slots, or the ring of bulk-copied rows
  if (WIDE && EST) smem += 16 + (BULK ? PLAN_RING_BYTES : 4 * SLOT_THREADS * sizeof(uint4));
  if (smem > 48 * 1024) B2C_CUDA((set_max_smem<slot_kernel<NTX, EXACT, EST, NSC, FAST, WIDE, STORE, COMPACT, BULK>>(smem)));
  kern<<<(unsigned)(B * a.g.nrx), SLOT_THREADS, smem, stream>>>(a);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}
template <int NTX, bool EXACT, bool EST, int NSC, bool FAST, int WIDE = 0, bool STORE = true, bool COMPACT = false>
static int launch_slot(const SlotArgs &a, int64_t B, size_t smem, cudaStream_t stream) {
  if constexpr (WIDE != 0 && EST && STORE) {
    // storing wide kernels: plan rows by bulk copy into the shared-memory ring (measured against the per-thread cp.async
    // staging on one box: full layout 2.56 vs 2.68 ms per 4144 4x4 slots = 0.93 vs 0.89 of the HBM peak, c5 dataset mode
    // +3.9 %, 2x2 +2.9 %, compact +1.3 %; bit-identical outputs).  B2C_PLAN_BULK=0 selects the per-thread form.
    const char *e = getenv("B2C_PLAN_BULK");
    if (!(e && e[0] == '0')) return launch_slot_form<NTX, EXACT, EST, NSC, FAST, WIDE, STORE, COMPACT, true>(a, B, smem, stream);
  }
  return launch_slot_form<NTX, EXACT, EST, NSC, FAST, WIDE, STORE, COMPACT, false>(a, B, smem, stream);
}

// Fast path: the throughput configuration -- default grid (599 used bins), power-of-two TX count,
// even symbol count, Philox draws, every output of the call's kind requested (simulate-only: H, rx, tx;
// with estimation: all five arrays + stats).  Everything else (parity runs with
// injected draws, partial outputs, other geometries) takes the generic instantiation.
template <bool EST>
static int launch_slot_ntx(const SlotArgs &a, int64_t B, size_t smem, cudaStream_t stream) {
  const int ntx = a.g.ntx;
  const bool fast = !a.compact && a.g.nsc == 599 && (a.g.nsym & 1) == 0 && !a.has_inj && a.H_true && a.rx && a.tx &&
                    (!EST || (a.H_ls && a.H_mmse && a.stats));
  const int pitch = a.g.pitch ? a.g.pitch : a.g.nsc;
  if constexpr (EST) {
    // statistics only (no array requested) on the default grid: the store-free instantiation of the wide kernel
    const bool stats_only = a.stats && !a.H_true && !a.rx && !a.tx && !a.H_ls && !a.H_mmse && !a.compact && a.g.nsc == 599 &&
                            (a.g.nsym & 1) == 0 && !a.has_inj;
    // statistics-only sweeps take the register-blocked form (measured on c4: 3.99 M vs 3.47 M slots/s; with stores the two
    // forms are within 1 % of each other and the 320-thread one keeps more warps in flight); B2C_NO_SLOT2=1 switches it off
    if (stats_only && !getenv("B2C_NO_SLOT2")) {
      if (ntx == 1) return launch_slot2<1, true, false, false>(a, B, smem, stream);
      if (ntx == 2) return launch_slot2<2, true, false, false>(a, B, smem, stream);
      if (ntx == 4) return launch_slot2<4, true, false, false>(a, B, smem, stream);
      if (ntx == 8) return launch_slot2<8, true, false, false>(a, B, smem, stream);
    }
    if (stats_only) {
      if (ntx == 1) return launch_slot<1, true, true, 599, true, WIDE_PITCH, false>(a, B, smem, stream);
      if (ntx == 2) return launch_slot<2, true, true, 599, true, WIDE_PITCH, false>(a, B, smem, stream);
      if (ntx == 4) return launch_slot<4, true, true, 599, true, WIDE_PITCH, false>(a, B, smem, stream);
      if (ntx == 8) return launch_slot<8, true, true, 599, true, WIDE_PITCH, false>(a, B, smem, stream);
    }
  }
  if (pitch != a.g.nsc) {
    // padded rows: the wide-store kernel of the throughput configuration only (H_mmse and stats are optional there:
    // H_true + rx + tx + H_ls is what the reference's generate_sample returns)
    const bool wide_ok = a.g.nsc == 599 && (a.g.nsym & 1) == 0 && !a.has_inj && a.H_true && a.rx && a.tx &&
                         (!EST || a.H_ls);
    B2C_REQUIRE(wide_ok && pitch == WIDE_PITCH, B2C_E_UNSUPPORTED,
                "b2c_slot_pipeline: pitch=%d needs the throughput configuration (599 bins, even nsym, Philox draws, "
                "H_true + rx + tx (+ H_ls when estimating) requested) and pitch == %d", pitch, WIDE_PITCH);
    if (a.compact) {
      if (ntx == 1) return launch_slot<1, true, EST, 599, true, WIDE_PITCH, true, true>(a, B, smem, stream);
      if (ntx == 2) return launch_slot<2, true, EST, 599, true, WIDE_PITCH, true, true>(a, B, smem, stream);
      if (ntx == 4) return launch_slot<4, true, EST, 599, true, WIDE_PITCH, true, true>(a, B, smem, stream);
      if (ntx == 8) return launch_slot<8, true, EST, 599, true, WIDE_PITCH, true, true>(a, B, smem, stream);
    }
    if (ntx == 1) return launch_slot<1, true, EST, 599, true, WIDE_PITCH>(a, B, smem, stream);
    if (ntx == 2) return launch_slot<2, true, EST, 599, true, WIDE_PITCH>(a, B, smem, stream);
    if (ntx == 4) return launch_slot<4, true, EST, 599, true, WIDE_PITCH>(a, B, smem, stream);
    if (ntx == 8) return launch_slot<8, true, EST, 599, true, WIDE_PITCH>(a, B, smem, stream);
    B2C_REQUIRE(false, B2C_E_UNSUPPORTED, "b2c_slot_pipeline: padded rows need ntx in {1, 2, 4, 8}, got %d", ntx);
  }
  if (fast) {
    if (ntx == 1) return launch_slot<1, true, EST, 599, true>(a, B, smem, stream);
    if (ntx == 2) return launch_slot<2, true, EST, 599, true>(a, B, smem, stream);
    if (ntx == 4) return launch_slot<4, true, EST, 599, true>(a, B, smem, stream);
    if (ntx == 8) return launch_slot<8, true, EST, 599, true>(a, B, smem, stream);
  }
  if (ntx <= 1) return launch_slot<1, false, EST, 0, false>(a, B, smem, stream);
  if (ntx <= 2) return launch_slot<2, false, EST, 0, false>(a, B, smem, stream);
  if (ntx <= 4) return launch_slot<4, false, EST, 0, false>(a, B, smem, stream);
  return launch_slot<8, false, EST, 0, false>(a, B, smem, stream);
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_tap_gains(const b2c_geom *g, const b2c_profiles *prof, const b2c_slots *slots,
                             const b2c_inject *inj, int64_t B, float *gains, float *noise_std,
                             void *stream) {
  B2C_REQUIRE(g && prof && slots && gains && noise_std, B2C_E_ARG, "b2c_tap_gains: null argument");
  if (int rc = check_geom(g)) return rc;
  B2C_REQUIRE(B >= 0 && B < (1ll << 31), B2C_E_ARG, "b2c_tap_gains: B=%lld out of range", (long long)B);
  B2C_REQUIRE(!inj || inj->jakes_u, B2C_E_ARG, "b2c_tap_gains: inject struct without jakes_u");
  if (B == 0) return B2C_OK;
  b2c_inject ij = {};
  if (inj) ij = *inj;
  // stage-1 oscillators for up to MAXT taps; stage 3 reuses the buffer for the tx-summed gains
  size_t osc = (size_t)MAXT * g->ntx * g->nrx * NOSC * sizeof(float2);
  size_t gsb = (size_t)g->nrx * g->nsym * MAXT * sizeof(float2);
  size_t smem = osc > gsb ? osc : gsb;
  B2C_REQUIRE(smem <= 200 * 1024, B2C_E_UNSUPPORTED, "b2c_tap_gains: %zu B shared memory needed", smem);
  if (smem > 48 * 1024) B2C_CUDA(set_max_smem<tap_gains_kernel>(smem));
  tap_gains_kernel<<<(unsigned)B, GAIN_THREADS, smem, (cudaStream_t)stream>>>(
      *g, *prof, *slots, ij, inj != nullptr, reinterpret_cast<float2 *>(gains), noise_std);
  B2C_CUDA(cudaGetLastError());
  return B2C_OK;
}

extern "C" int b2c_slot_pipeline(const b2c_geom *g, const b2c_profiles *prof, const b2c_patterns *pat,
                                 const b2c_slots *slots, const b2c_inject *inj, int64_t B,
                                 const float *gains, const float *noise_std, float *H_true, float *rx,
                                 float *tx, float *H_ls, float *H_mmse, double *stats, int32_t compact,
                                 const b2c_pilot_io *pilots_out, void *stream) {
  B2C_REQUIRE(g && prof && slots && gains && noise_std, B2C_E_ARG, "b2c_slot_pipeline: null argument");
  if (int rc = check_geom(g, /*allow_pitch=*/true)) return rc;
  B2C_REQUIRE(B >= 0 && B * g->nrx < (1ll << 31), B2C_E_ARG, "b2c_slot_pipeline: B=%lld out of range",
              (long long)B);
  const bool est = H_ls || H_mmse || stats || pilots_out;
  B2C_REQUIRE(!pilots_out || (pilots_out->hp && pat && pilots_out->ld >= pat->np_max), B2C_E_ARG,
              "b2c_slot_pipeline: pilots_out needs hp and ld >= np_max");
  B2C_REQUIRE(!est || (pat && pat->plan && pat->pilot_re && pat->npilots && slots->pattern_id), B2C_E_ARG,
              "b2c_slot_pipeline: estimation outputs requested without a pattern pool");
  B2C_REQUIRE(!inj || !(est || rx || tx) || (inj->sym_turns && inj->noise), B2C_E_ARG,
              "b2c_slot_pipeline: inject struct needs sym_turns and noise");
  B2C_REQUIRE(!est || pat->np_max <= 65534, B2C_E_UNSUPPORTED, "b2c_slot_pipeline: more than 65534 pilots");
  if (B == 0) return B2C_OK;
  SlotArgs a = {};
  a.g = *g;
  a.prof = *prof;
  if (pat) a.pat = *pat;
  a.slots = *slots;
  if (inj) a.inj = *inj;
  a.has_inj = inj != nullptr;
  a.gains = reinterpret_cast<const float2 *>(gains);
  a.noise_std = noise_std;
  a.H_true = reinterpret_cast<float2 *>(H_true);
  a.rx = reinterpret_cast<float2 *>(rx);
  a.tx = reinterpret_cast<float2 *>(tx);
  a.H_ls = reinterpret_cast<float2 *>(H_ls);
  a.H_mmse = reinterpret_cast<float2 *>(H_mmse);
  a.stats = stats;
  a.compact = compact != 0;
  // QPSK grid: keep the word's top two bits (the quadrant) and put the phase at the quadrant's centre: the 23-bit
  // mantissa becomes k 2^21 + 2^20, i.e. u = k/4 + 1/8 (+ 2^-24)
  a.sym_and = slots->qpsk ? 0xC0000000u : 0xFFFFFFFFu;
  a.sym_or = slots->qpsk ? 0x20000000u : 0u;
  if (pilots_out) {
    a.hp_out = reinterpret_cast<float2 *>(pilots_out->hp);
    a.hp_col = pilots_out->col;
    a.hp_ld = pilots_out->ld;
  }
  size_t smem = slot_smem_bytes(g, est ? pat->np_max : 0);
  B2C_REQUIRE(smem <= 100 * 1024, B2C_E_UNSUPPORTED, "b2c_slot_pipeline: %zu B shared memory needed", smem);
  return est ? launch_slot_ntx<true>(a, B, smem, (cudaStream_t)stream)
             : launch_slot_ntx<false>(a, B, smem, (cudaStream_t)stream);
}

extern "C" int b2c_dense_score(const b2c_geom *g, const b2c_profiles *prof, const b2c_patterns *pat,
                               const b2c_slots *slots, int64_t B, const float *gains, const float *hm,
                               const int32_t *hm_col, int64_t hm_ld, double *stats, void *stream) {
  B2C_REQUIRE(g && prof && pat && slots && gains && hm && stats, B2C_E_ARG, "b2c_dense_score: null argument");
  if (int rc = check_geom(g)) return rc;
  B2C_REQUIRE(B >= 0 && B * g->nrx < (1ll << 31), B2C_E_ARG, "b2c_dense_score: B=%lld out of range", (long long)B);
  B2C_REQUIRE(pat->plan && pat->npilots && slots->pattern_id && slots->model_id, B2C_E_ARG,
              "b2c_dense_score: pattern pool / per-slot ids missing");
  B2C_REQUIRE(hm_ld >= pat->np_max && pat->np_max <= 65534, B2C_E_ARG, "b2c_dense_score: hm_ld=%lld < np_max=%d",
              (long long)hm_ld, pat->np_max);
  const int ntx = g->ntx;
  B2C_REQUIRE(g->nsc == 599 && (g->nsym & 1) == 0 && (ntx == 1 || ntx == 2 || ntx == 4 || ntx == 8), B2C_E_UNSUPPORTED,
              "b2c_dense_score: needs the default grid (599 bins, even nsym) and ntx in {1, 2, 4, 8}");
  if (B == 0) return B2C_OK;
  SlotArgs a = {};
  a.g = *g;
  a.prof = *prof;
  a.pat = *pat;
  a.slots = *slots;
  a.gains = reinterpret_cast<const float2 *>(gains);
  a.stats = stats;
  a.hp_out = const_cast<float2 *>(reinterpret_cast<const float2 *>(hm));   // read only in SCORE mode
  a.hp_col = hm_col;
  a.hp_ld = hm_ld;
  const size_t smem = slot_smem_bytes(g, pat->np_max);
  B2C_REQUIRE(smem <= 100 * 1024, B2C_E_UNSUPPORTED, "b2c_dense_score: %zu B shared memory needed", smem);
  cudaStream_t st = (cudaStream_t)stream;
  if (ntx == 1) return launch_slot2<1, true, false, false, true>(a, B, smem, st);
  if (ntx == 2) return launch_slot2<2, true, false, false, true>(a, B, smem, st);
  if (ntx == 4) return launch_slot2<4, true, false, false, true>(a, B, smem, st);
  return launch_slot2<8, true, false, false, true>(a, B, smem, st);
}
