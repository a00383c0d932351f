/*
 * b2c.h -- C ABI of libb2c.so: batched MIMO-OFDM channel simulation and LS/MMSE channel
 * estimation on NVIDIA B200 (sm_100a).
 *
 * The reference (anish-dev09/CHANNEL-ESTIMATION-IN-5G-NETWORK) is pure Python and has no
 * FFI of its own; the "plugin interface" this library sits behind is the Python surface of
 * src/channel_simulator.py, src/baseline_estimators.py and src/dataset_generator.py.  Each entry
 * point below names the reference function(s) it replaces (file:line into the reference) and
 * is bound from Python with ctypes (see INTEGRATION.md and
 * channel-estimation-in-5g-network_b200/_b2c.py).
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer unless the
 *     parameter name ends in _host; complex64 arrays are interleaved (re, im) float pairs;
 *   - nothing allocates, nothing synchronises: kernels are enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - return value 0 = enqueued; negative = B2C_E_*; b2c_last_error_string() describes the
 *     last failure on the calling thread;
 *   - arrays are row-major with the reference's per-slot axis order, stacked over a leading
 *     slot axis B:   H[B][nsym][nrx][ntx][nsc], rx[B][nsym][nrx][nsc], tx[B][nsym][ntx][nsc].
 *
 * Random draws.  Every kernel that consumes randomness has two modes:
 *   injected  -- the caller supplies the draws (parity tests feed the reference's recorded
 *                numpy draws);
 *   Philox    -- Philox4x32-10 keyed by (seed, global slot index): results do not depend on
 *                batch size, launch geometry or the number of GPUs the slots are sharded over.
 *   Counter layout: key = (seed lo, seed hi); ctr = (index, stream, slot lo, slot hi)
 *     (used bin k has frequency offset f = k - half for k < half, f = k - half + 1 otherwise, half = (nsc+1)/2;
 *      it is drawn from Philox lane l = |f| - 1 and word half h = (f > 0): the kernels' thread l owns the
 *      mirror pair -f / +f, whose channel twiddles are complex conjugates)
 *     stream 0 SYMBOLS: index = (s>>1)*320 + l, word (s&1)*2 + h   -> phase of RE (s,k), in turns
 *     stream 1 JAKES  : index = ((p*ntx+tx)*nrx+rx)*10 + (n>>1), words (0,1) even n / (2,3) odd n
 *                                                                  -> (arrival angle, phase) of oscillator n
 *     stream 2 NOISE  : index = (s*nrx+rx)*320 + l, words (2h, 2h+1)
 *                                                                  -> Box-Muller (u1,u2) of rx[s][rx][k]
 *     uniform u = ((word>>9)+0.5)*2^-23.   oracle/philox.py is the bit-exact CPU twin.
 */
#ifndef B2C_H_
#define B2C_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2C_ABI_VERSION 3
#define B2C_MAX_TAPS 16     /* distinct sample delays per TDL profile (EPA 5, EVA 8, ETU 9)  */
#define B2C_MAX_ANT 8       /* ntx, nrx <= 8                                                  */
#define B2C_MAX_SYM 16      /* OFDM symbols per slot                                          */
#define B2C_RNG_LANES 320   /* Philox lanes per row (see "Random draws"); nsc <= 2*320 - 1, odd */
#define B2C_N_OSC 20        /* Jakes oscillators, src/channel_simulator.py:100                */
#define B2C_N_STAT 3        /* sum|H-H_ls|^2, sum|H-H_mmse|^2, sum|H|^2 ...                    */
#define B2C_N_STATGRP 2     /* ... per rx antenna for {antenna pair (rx,0), all tx of that rx} */

enum {
  B2C_OK = 0,
  B2C_E_ARG = -1,           /* invalid argument (null pointer, size out of range)             */
  B2C_E_CUDA = -2,          /* a CUDA runtime call failed; see b2c_last_error_string()        */
  B2C_E_UNSUPPORTED = -3    /* geometry beyond the compiled limits above                      */
};

/* Slot geometry.  OFDMConfig / MIMOConfig, src/channel_simulator.py:17-31; nsc is
 * len(OFDMSystem.used_indices) (:141-148), 599 for the default 600 useful subcarriers. */
typedef struct b2c_geom {
  int32_t nsym;             /* OFDM symbols per slot (14)                                     */
  int32_t nsc;              /* used subcarriers (599)                                         */
  int32_t ntx, nrx;
  int32_t fft_size;         /* 1024                                                           */
  int32_t cp_length;        /* 72                                                             */
  float symbol_period_s;    /* (fft_size+cp_length)/sampling_rate: spacing of the symbol-start
                               instants the channel is sampled at (:300-302)                  */
  int32_t pitch;            /* complex elements between consecutive rows of the [...][nsc] arrays;
                               0 or nsc = contiguous (every entry point).  b2c_slot_pipeline also
                               takes 600 in its throughput configuration: rows padded by one element
                               so that each lane writes 16 aligned bytes;
                               b2c_ml_features / b2c_pair00_* read either layout.               */
} b2c_geom;

/* TDL profile tables, built on the host once per (profile set, geometry) by
 * channel-estimation-in-5g-network_b200/_tables.py from ChannelModel.__init__ (:56-82) with the
 * duplicate-delay overwrite of :125 resolved.  All device pointers, M = number of profiles.  */
typedef struct b2c_profiles {
  int32_t n_models;
  const int32_t *ntaps;     /* [M]                    surviving taps                           */
  const int32_t *npaths;    /* [M]                    paths in the profile (RNG draw order)    */
  const int32_t *tap_path;  /* [M][B2C_MAX_TAPS]      path index that owns the tap             */
  const float *tap_amp;     /* [M][B2C_MAX_TAPS]      sqrt(P_path) / sqrt(2*20)                */
  const float *tap_tw;      /* [M][B2C_MAX_TAPS][nsc] complex: exp(-j 2 pi (i_k - N/2) d / N)  */
  const float *tap_corr;    /* [M][B2C_MAX_TAPS][B2C_MAX_TAPS] complex: sum_k T[p,k] conj(T[q,k]) */
  const int32_t *tap_delay; /* [M][B2C_MAX_TAPS]      sample delay of the tap (b2c_tdl_circular) */
} b2c_profiles;

/* Pool of pilot patterns with their interpolation plans.  PilotPattern
 * (src/channel_simulator.py:209-236) + the Delaunay/barycentric (or nearest) structure that
 * scipy.interpolate.griddata builds inside LSEstimator.interpolate_channel
 * (src/baseline_estimators.py:65-79); built once per pattern on the host.
 * plan: nsym*nsc + 1 entries of 16 bytes per pattern, row-major (sym, sc):
 *   uint16 i0, i1, i2, flags(bit0 = inside hull, informational); float w0, w1;   w2 = 1 - w0 - w1
 *   value = w0*h[i0] + w1*h[i1] + w2*h[i2] over the pilot vector h extended by h[np_max] = 0.
 *   Resource elements outside the pilots' convex hull carry i0 = i1 = i2 = np_max, w0 = 1, w1 = 0
 *   (griddata fill_value=0.0, exactly); the extra last entry is such an "outside" row.         */
typedef struct b2c_patterns {
  int32_t n_patterns;
  int32_t np_max;           /* row stride of pilot_re                                         */
  const int32_t *npilots;   /* [n_patterns]                                                   */
  const int32_t *pilot_re;  /* [n_patterns][np_max]  sorted flat RE index of pilot j          */
  const void *plan;         /* [n_patterns][nsym*nsc + 1] 16-byte plan entries                */
} b2c_patterns;

/* Per-slot parameters (device arrays of length B).  */
typedef struct b2c_slots {
  int64_t slot0;            /* global index of slot 0 of this batch (Philox counter)          */
  uint64_t seed;
  const int32_t *model_id;  /* index into b2c_profiles                                        */
  const float *doppler_hz;
  const float *snr_db;
  const int32_t *pattern_id;/* index into b2c_patterns                                        */
  int32_t qpsk;             /* Philox mode only.  0: every RE carries exp(j 2 pi u), u uniform -- the reference's grid
                               (src/channel_simulator.py:394-399).  1: the phase is quantised to the QPSK points
                               (2 floor(4u) + 1) pi / 4, i.e. the grid carries 2 random bits per RE, so that
                               equalise -> demap -> count gives a real bit-error rate (b2c_bit_errors_per_slot)   */
} b2c_slots;

/* Injected draws (all nullable as a group: NULL struct pointer = Philox mode).
 *   jakes_u  [B][P_max][ntx][nrx][2][20] float, raw U[0,1) (angles then phases), reference order
 *            src/channel_simulator.py:102-110; P_max = max npaths over the profile set
 *   sym_turns[B][nsym][nsc] float, phase/(2 pi) of each resource element's symbol (:394-399)
 *   noise    [B][nsym][nrx][nsc] complex float, the two randn blocks of :342 interleaved      */
typedef struct b2c_inject {
  const float *jakes_u;
  int32_t p_max;
  const float *sym_turns;
  const float *noise;
} b2c_inject;

/* Pilot-vector output of b2c_slot_pipeline for the dense-Wiener (known-covariance MMSE) pipeline
 * (src/baseline_estimators.py:169-190): h_ls at the pilots of (slot b, rx) is written to row col[b] + rx of
 *   hp [rows][ld] complex, elements j < npilots (the row's remaining elements are left as they are).
 * `col` lets the caller lay the vectors out grouped by Wiener matrix (one (pattern, SNR) group = one contiguous
 * block of rows = one GEMM); NULL = b * nrx.                                                             */
typedef struct b2c_pilot_io {
  float *hp;
  const int32_t *col;       /* [B] or NULL                                                    */
  int64_t ld;               /* complex elements between rows, >= np_max                       */
} b2c_pilot_io;

const char *b2c_last_error_string(void);
int b2c_abi_version(void);

/* K1a.  Jakes tap gains at the nsym symbol-start instants + the slot's AWGN standard deviation.
 * Replaces ChannelModel.generate_time_varying_channel as consumed by
 * MIMOChannel.generate_channel_frequency_response (src/channel_simulator.py:84-127, 285-302)
 * and the power/noise-scale arithmetic of apply_channel (:337-340).
 *   gains    [B][nrx][nsym][ntx][B2C_MAX_TAPS] complex float (out)
 *   noise_std[B] float (out): sqrt(mean|H x|^2 / 10^(snr/10) / 2); the mean is evaluated as the
 *            quadratic form g^H C g (all TX send the same unit-modulus grid, :402-404).        */
int b2c_tap_gains(const b2c_geom *g, const b2c_profiles *prof, const b2c_slots *slots,
                  const b2c_inject *inj, int64_t B, float *gains, float *noise_std, void *stream);

/* K1+K3+K5 fused.  One pass per slot: CFR from the tap gains, resource grid, y = Hx + n, LS at
 * the pilots, plan interpolation, default MMSE (alpha * LS) and squared-error statistics.
 * Replaces simulate_transmission (src/channel_simulator.py:348-421) followed by
 * LSEstimator('linear'|'nearest').estimate (src/baseline_estimators.py:83-117),
 * MMSEEstimator().estimate on its default branch (:232-270, :177-180) and evaluate_estimator
 * (:315-337).  Every output pointer is optional (NULL = not written); estimation is skipped
 * entirely when H_ls, H_mmse and stats are all NULL (then this is simulate_transmission alone).
 *   H_true [B][nsym][nrx][ntx][nsc], rx [B][nsym][nrx][nsc], tx [B][nsym][ntx][nsc] complex
 *   compact = 1: the tx-replicated outputs are written once -- H_ls, H_mmse [B][nsym][nrx][nsc] and
 *   tx [B][nsym][nsc] (every TX antenna sends the same grid, :402-404, and LS/MMSE never see tx) --
 *   for callers that expand them as stride-0 views (the host-buffer pipeline: half the PCIe bytes).
 *   H_ls, H_mmse like H_true;  stats [B][nrx][B2C_N_STATGRP][B2C_N_STAT] double
 *   g->pitch = 600 (throughput configuration only: nsc = 599, even nsym, ntx in {1,2,4,8}, Philox draws, H_true + rx +
 *   tx requested and H_ls when estimating -- H_mmse and stats stay optional --; B2C_E_UNSUPPORTED
 *   otherwise): the last axis of the five arrays
 *   is 600 elements apart in memory (element 599 is padding) and every lane writes 16 aligned bytes per row;
 *   pilots_out (optional, needs estimation): also hand out h_ls at the pilots, see b2c_pilot_io.
 *   same values as the contiguous layout, bit for bit.  compact = 1 combines with it: each unique value is written
 *   once (1 945 552 B per 4x4 slot instead of 3 756 928), in padded rows.                                  */
int b2c_slot_pipeline(const b2c_geom *g, const b2c_profiles *prof, const b2c_patterns *pat,
                      const b2c_slots *slots, const b2c_inject *inj, int64_t B,
                      const float *gains, const float *noise_std,
                      float *H_true, float *rx, float *tx, float *H_ls, float *H_mmse,
                      double *stats, int32_t compact, const b2c_pilot_io *pilots_out, void *stream);

/* K3.  LS pilot division + plan interpolation on caller-supplied received grids.
 * Replaces LSEstimator.estimate (src/baseline_estimators.py:83-117) and, with mmse_mode=1,
 * MMSEEstimator.estimate's default branch (:232-270).
 *   rx       [B][nsym][nrx][nsc] complex (the reference's rx_4d is this replicated over tx)
 *   pilots   [B or 1][np_max] complex transmitted pilot symbols; pilots_stride = np_max or 0
 *   hp_in    optional [B][nrx][np_max] complex: use these pilot-position values instead of
 *            rx/pilots (the dense-Wiener path feeds W h_ls here)
 *            hp_col / hp_ld (optional): row of (slot b, rx) in hp_in is hp_col[b] + rx and rows are hp_ld
 *            elements apart (NULL / 0: b * nrx + rx, np_max) -- the layout b2c_pilot_io describes
 *   mmse_mode 0: H_mmse not produced; 1: alpha = P/(P+10^(-snr/10)), P = mean|h_ls|^2 (:174-180);
 *            2: hp_in holds Wiener-filtered pilot estimates (known-covariance branch, :181-190): their
 *               interpolation is written to H_mmse, and of the statistics only the MMSE error sums (entry 1 of
 *               each group) are written -- the other entries keep what b2c_slot_pipeline put there
 *   H_true   optional, for stats.  hp_out optional [B][nrx][np_max]: h_ls at the pilots (:110).
 *   g->pitch = 600 (nsc = 599, ntx in {1,2,4,8}): rx, H_true, H_ls and H_mmse all have padded rows.       */
int b2c_ls_interp(const b2c_geom *g, const b2c_patterns *pat, const int32_t *pattern_id,
                  const float *snr_db, int64_t B, const float *rx, const float *pilots,
                  int64_t pilots_stride, const float *hp_in, int32_t mmse_mode,
                  const float *H_true, float *H_ls, float *H_mmse, float *hp_out, double *stats,
                  const int32_t *hp_col, int64_t hp_ld, void *stream);

/* LS / default-MMSE on bare pilot vectors: out[v][j] = alpha_v * y[v][j] / (x[j] + 1e-12).
 * Replaces LSEstimator.estimate_at_pilots (src/baseline_estimators.py:23-42) with mmse_mode=0 and
 * MMSEEstimator.estimate_at_pilots' default branch (:155-196) with mmse_mode=1
 * (alpha_v = P/(P+10^(-snr_db/10)), P = mean_j |y/x|^2).   y, out [nvec][n]; x [n] complex.       */
int b2c_pilot_vectors(const float *y, const float *x, int64_t nvec, int32_t n, float snr_db,
                      int32_t mmse_mode, float *out, void *stream);

/* K4.  Dense Wiener filter at the pilots: out[c][:] = W @ in[c][:], complex64, for ncols
 * columns (slot x rx).  Replaces `mmse_matrix @ h_ls` of MMSEEstimator.estimate_at_pilots on its
 * known-covariance branch (src/baseline_estimators.py:181-190); W = R (R + sigma^2 I)^-1 is
 * built once per (pattern, SNR) on the host.
 *   W [np][np] row-major complex; in/out [ncols][ld] complex.                                 */
int b2c_mmse_dense(const float *W, int32_t np, const float *in, float *out, int64_t ncols,
                   int64_t ld, void *stream);

/* K4b.  Dense REAL linear map applied to complex columns on the tensor cores (same tcgen05 / 3xTF32
 * machinery): out[c][e] = sum_j W[e][j] * in[c][j], W real [m][k] row-major, in [ncols][ld_in],
 * out [ncols][ld_out] complex.  Used for LSEstimator(interpolation_method='cubic')
 * (src/baseline_estimators.py:13-21, 65-79): SciPy's Clough-Tocher interpolant on a fixed pilot set is a
 * linear map of the pilot values (to 1e-6), built once per pattern on the host as W [nsym*nsc][npilots]. */
int b2c_dense_real_apply(const float *W, int32_t m, int32_t k, const float *in, float *out, int64_t ncols,
                         int64_t ld_in, int64_t ld_out, void *stream);

/* K4 / K4b with a prepared operand.  b2c_dense_prepare writes the A operand (the complex W of b2c_mmse_dense with
 * is_complex = 1, m = k = np; or the real W [m][k] of b2c_dense_real_apply) once, already split into TF32 hi / lo
 * parts and already in the GEMM's shared-memory tile layout, into a caller-owned workspace of
 * b2c_dense_prepared_bytes(m, k, is_complex) bytes (16-byte aligned); b2c_dense_apply_prepared then fetches each
 * K stage of A with one 32 KB bulk copy (cp.async.bulk + mbarrier) instead of staging it through registers.
 * Same results as the one-shot entry points, bit for bit.                                                     */
int64_t b2c_dense_prepared_bytes(int32_t m, int32_t k, int32_t is_complex);
int b2c_dense_prepare(const float *W, int32_t m, int32_t k, int32_t is_complex, void *prepared, void *stream);
int b2c_dense_apply_prepared(const void *prepared, int32_t m, int32_t k, int32_t is_complex, const float *in,
                             float *out, int64_t ncols, int64_t ld_in, int64_t ld_out, void *stream);

/* K4, grouped.  Several prepared complex products in ONE launch: group g applies its Wiener matrix (np_g x np_g, in
 * b2c_dense_prepare's form) to columns [col0_g, col0_g + ncols_g) of in / out [.][ld] -- the (pattern, SNR) groups of a
 * batch in the dense-Wiener pipeline (W = R (R + sigma^2 I)^-1 "built once, then applied to the batch",
 * src/baseline_estimators.py:181-190).  One grid over the tiles of all groups: the small per-group tile counts fill
 * the SMs together.  groups_host is a HOST array (read during the call); at most 32 groups per call.  Same results as
 * ngroups calls of b2c_dense_apply_prepared, bit for bit.                                                    */
typedef struct b2c_dense_group {
  const void *prepared;     /* b2c_dense_prepare(W, np, np, 1, ...) workspace (device)         */
  int64_t col0, ncols;      /* this group's block of columns                                  */
  int32_t np;
} b2c_dense_group;
int b2c_dense_apply_grouped(const b2c_dense_group *groups_host, int32_t ngroups, const float *in, float *out, int64_t ld,
                            void *stream);

/* Dense statistics, second pass.  For evaluation runs that want only the error statistics of the dense Wiener estimate
 * (the MSE / NMSE-versus-SNR curves of evaluate_estimator, src/baseline_estimators.py:326-337, with
 * MMSEEstimator(channel_cov=...) as the estimator) no resource-grid array has to exist in HBM at all:
 *   b2c_slot_pipeline(stats + pilots_out, no arrays)  ->  b2c_dense_apply_grouped  ->  b2c_dense_score.
 * This entry regenerates the true CFR of every (slot, rx) from the tap gains exactly as the slot kernel does,
 * interpolates the FILTERED pilot vector hm (row hm_col[b] + rx of [.][hm_ld], or b * nrx + rx when hm_col is NULL)
 * through the pattern's plan and writes sum |H_mmse - H_true|^2 into stats[b][rx][q][1] (q = 0: pair (rx, 0), q = 1:
 * all tx); the other fields of stats are left as the first pass wrote them.  Default grid only (599 bins, even nsym,
 * ntx in {1, 2, 4, 8}); gains as b2c_tap_gains wrote them for the same batch.                                     */
int b2c_dense_score(const b2c_geom *g, const b2c_profiles *prof, const b2c_patterns *pat, const b2c_slots *slots,
                    int64_t B, const float *gains, const float *hm, const int32_t *hm_col, int64_t hm_ld,
                    double *stats, void *stream);

/* K5.  Fold per-slot statistics into per-bin float64 accumulators (deterministic order).
 * Replaces the per-sample evaluate_estimator / compute_nmse + list aggregation of
 * src/baseline_estimators.py:326-337 and run_phase8_pilot_optimization.py:32-37,186-206.
 *   stats [B][nrx][2][3] (from the kernels above), bin_id [B] in [0, nbins) or <0 to skip
 *   bins  [nbins][B2C_N_BINSTAT] double, ACCUMULATED INTO (zero it first):
 *     0 count  1 sum mse_ls  2 sum mse_mmse  3 sum nmse_ls  4 sum nmse_mmse  5 sum nmse_ls^2
 *     6 sum nmse_mmse^2  7 sum mean|H|^2  8 sum nmse00_ls  9 sum nmse00_ls^2  10 sum nmse00_mmse
 *     11 sum nmse00_mmse^2       (nmse: /(pow+1e-12); nmse00: antenna pair (0,0), /(pow+1e-10))
 *     12 sum ber_proxy_ls  13 sum ber_proxy_mmse: compute_ber_approximation (run_phase5_evaluation.py:57-68) of the
 *     pair-(0,0) NMSE at the slot's SNR, the "BER curve" of the pilot-density sweep; only when snr_db [B] is given
 *     (NULL: entries 12, 13 are left untouched)                                                        */
#define B2C_N_BINSTAT 14
int b2c_stats_bins(const b2c_geom *g, const double *stats, const int32_t *bin_id, const float *snr_db, int64_t B,
                   int32_t nbins, double *bins, void *stream);

/* K2.  OFDM modulate / demodulate, batched over rows (one row = one OFDM symbol of one antenna).
 * Replaces OFDMSystem.modulate (src/channel_simulator.py:150-178): map 599 -> 1024 bins,
 * ifftshift, IFFT * sqrt(N), cyclic-prefix prepend; and OFDMSystem.demodulate (:180-203): CP
 * strip, FFT / sqrt(N), fftshift, gather used bins.  fft_size must be 1024.
 *   modulate:   in [rows][nsc] -> out [rows][fft_size+cp]      demodulate: the reverse          */
int b2c_ofdm_modulate(const b2c_geom *g, const float *in, float *out, int64_t rows, void *stream);
int b2c_ofdm_demodulate(const b2c_geom *g, const float *in, float *out, int64_t rows, void *stream);

/* a8 generic.  MIMOChannel.apply_channel (src/channel_simulator.py:313-345) for arbitrary tx
 * grids and channel responses: y = H x per RE, slot-mean power, AWGN.
 *   tx [B][nsym][ntx][nsc], H [B][nsym][nrx][ntx][nsc], snr_db [B], rx [B][nsym][nrx][nsc] (out)
 *   noise injected (inj->noise) or Philox stream 2; power_scratch [B] double (zeroed by call).  */
int b2c_apply_channel(const b2c_geom *g, const b2c_slots *slots, const b2c_inject *inj, int64_t B,
                      const float *tx, const float *H, float *rx, double *power_scratch,
                      void *stream);

/* a3 standalone.  ChannelModel.generate_time_varying_channel (src/channel_simulator.py:84-127):
 * the full (num_samples, nrx, ntx, max_delay+1) CIR of ONE realisation, sample n at t = n/fs.
 *   jakes_u [npaths][ntx][nrx][2][20] injected, or NULL for Philox (slot = slots->slot0)
 *   tap_delay_host [ntaps] HOST array of the surviving taps' sample delays, ntaps = prof->ntaps[model_id] as the host
 *   tables hold it (passed in so that the call reads nothing back from the device and never synchronises)
 *   out [num_samples][nrx][ntx][L] complex, L = max delay + 1 (written in full, zeros included) */
int b2c_tdl_full(const b2c_geom *g, const b2c_profiles *prof, int32_t model_id, float doppler_hz,
                 float sample_period_s, int64_t num_samples, int32_t L, const int32_t *tap_delay_host,
                 int32_t ntaps, const float *jakes_u, uint64_t seed, int64_t slot, float *out, void *stream);

/* Time-domain statement of the channel: per-symbol CIRCULAR convolution of the modulated symbols with the tap gains of
 * b2c_tap_gains, y[s][rx][n] = sum_tx sum_t g[rx][s][tx][t] x[s][tx][(n - d_t) mod N] on the N-sample bodies, cyclic
 * prefix re-attached.  b2c_ofdm_modulate -> this -> b2c_ofdm_demodulate equals the frequency-domain product
 * MIMOChannel.apply_channel forms per bin (src/channel_simulator.py:274-345, noise aside): the reference samples the
 * channel once per symbol and multiplies CFRs, which is a circular convolution per symbol (a linear one cannot
 * reproduce it: ETU's last tap, 77 samples, exceeds the 72-sample prefix).
 *   gains [B][nrx][nsym][ntx][B2C_MAX_TAPS] (b2c_tap_gains), x_time [B][nsym][ntx][fft+cp], y_time [B][nsym][nrx][fft+cp]  */
int b2c_tdl_circular(const b2c_geom *g, const b2c_profiles *prof, const int32_t *model_id, int64_t B, const float *gains,
                     const float *x_time, float *y_time, void *stream);

/* ---- "next" rows either side of the path ------------------------------- */

/* equalize_channel (src/baseline_estimators.py:273-312): x = (H^H H + lambda I)^-1 H^H y per resource
 * element; lambda = 1e-8 for 'zf' (:297), 0.01 for 'mmse' (:305-306).  Accumulated and solved in fp64.
 *   rx [B][nsym][nrx][nsc], H [B][nsym][nrx][ntx][nsc], out [B][nsym][ntx][nsc]
 *   fp64_io 0: complex64 buffers, 1: complex128 buffers (the drop-in keeps the reference's dtype)   */
int b2c_equalize(const b2c_geom *g, int64_t B, const void *rx, const void *H, void *out, double lambda,
                 int32_t fp64_io, void *stream);

/* qam_modulation / qam_demodulation (src/utils.py:71-108, 111-152), M in {4, 16} (else
 * B2C_E_UNSUPPORTED, the reference's NotImplementedError): bits are bytes 0/1, MSB first per symbol;
 * demodulation is minimum distance, first minimum wins.  symbols complex64 (complex128 in if fp64_in). */
int b2c_qam_modulate(const uint8_t *bits, int64_t nsymbols, int32_t M, float *symbols, void *stream);
int b2c_qam_demodulate(const void *symbols, int64_t nsymbols, int32_t M, int32_t fp64_in, uint8_t *bits,
                       void *stream);

/* calculate_ber numerator (src/utils.py:155-157): *count += #{i : a[i] != b[i]}.                       */
int b2c_count_bit_errors(const uint8_t *a, const uint8_t *b, int64_t n, uint64_t *count, void *stream);

/* Per-slot bit errors on the DATA resource elements: counts[b] = #{ i : a[b][e][i] != b[b][e][i] } over the resource
 * elements e of slot b that are not pilots of its pattern (pilots carry no payload).  a, b: [B][nsym*nsc][bps] bits
 * (bytes 0 / 1) as b2c_qam_demodulate writes them for [B][nsym][nsc] grids; the per-(density, SNR) BER curves of the
 * pilot sweep are sums of these counts (calculate_ber, src/utils.py:155-157, applied per cell).             */
int b2c_bit_errors_per_slot(const b2c_geom *g, const b2c_patterns *pat, const int32_t *pattern_id, int64_t B,
                            const uint8_t *a, const uint8_t *b, int32_t bps, uint32_t *counts, void *stream);

/* Dataset integrity check (verify_phase3_datasets.py:98-111): counts[0] += #NaN, counts[1] += #Inf among n
 * elements; is_complex: elements are complex64 and count when either part is NaN / Inf (numpy.isnan / isinf). */
int b2c_count_nonfinite(const float *x, int64_t n, int32_t is_complex, uint64_t *counts, void *stream);

/* compute_mae numerator (run_phase5_evaluation.py:51-54): *sum += sum_i |a_i - b_i| over n complex64 elements.  */
int b2c_abs_diff_sum(const float *a, const float *b, int64_t n, double *sum, void *stream);

/* Feature packing for the ML side from GPU-resident slots, antenna pair (0,0).
 *   rx [B][nsym][nrx][nsc]; H_true [B][nsym][nrx][ntx][nsc]; H_ls with ls_sym_stride complex elements
 *   between consecutive symbols (nrx*ntx*nsc for the full layout, nrx*nsc for the compact one).
 * b2c_pair00_moments: moments[3][4] += {sum re, sum im, sum re^2, sum im^2} of rx, H_ls, H_true rows
 *   (ChannelDataset._compute_normalization_stats, src/train.py:41-57).
 * b2c_ml_features: inputs (rx re, rx im, H_ls re, H_ls im, pilot mask), targets (H_true re, im), float32
 *   layout 0: [B][nsym][nsc][5] / [B][nsym][nsc][2]   (prepare_ml_inputs, src/dataset_generator.py:183-227)
 *   layout 1: [B][5][nsym][nsc] / [B][2][nsym][nsc]   (ChannelDataset.__getitem__, src/train.py:62-94)
 *   norm_mode 0: none; 1: per-slot 1/(std+1e-8) of prepare_ml_inputs (:219-223);
 *             2: (v - mean) * scale with norm = {rx_mean, rx_scale, ls_mean, ls_scale, true_mean, true_scale} */
/* b2c_pair00_errors: out[b] = {sum |L-H|^2, sum |alpha_b L - H|^2, sum |H|^2} over the pair-(0,0) rows of
 *   slot b (alpha NULL = 1): the per-sample LS / alpha-scaled NMSE of run_phase5_evaluation.py:283-296. */
int b2c_pair00_errors(const b2c_geom *g, int64_t B, const float *H_ls, const float *H_true,
                      int64_t ls_sym_stride, const float *alpha, double *out, void *stream);
int b2c_pair00_moments(const b2c_geom *g, int64_t B, const float *rx, const float *H_ls, const float *H_true,
                       int64_t ls_sym_stride, double *moments, void *stream);
int b2c_ml_features(const b2c_geom *g, const b2c_patterns *pat, const int32_t *pattern_id, int64_t B,
                    const float *rx, const float *H_ls, const float *H_true, int64_t ls_sym_stride,
                    int32_t layout, int32_t norm_mode, const float *norm, float *inputs, float *targets,
                    void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B2C_H_ */
