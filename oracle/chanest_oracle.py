"""
CPU oracle for the MIMO-OFDM simulate + LS/MMSE hot path.  TEST INFRASTRUCTURE ONLY.

This file is a float64 NumPy restatement of the reference algorithm.  It is the
checker for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package never does (and fails loudly when its CUDA library is missing).

Parity status: PINNED.  ``oracle/make_golden.py`` runs the real reference
(`/root/reference/src`, importable in the build container) with a recorder
wrapped around ``numpy.random`` and stores draws + outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` asserts this restatement
reproduces those outputs to <= 1e-12 from the recorded draws.  The reference's
own tests hold no golden vectors, so the fixtures are
"outputs of the reference itself run here".

Every function cites the reference lines it restates (paths are into
/root/reference).  All randomness is *injected*: functions take the draws as
arrays (recorded from the reference, or produced by oracle/philox.py in the
counter layout the CUDA kernels use).

Two cost profiles are provided where the reference's own cost structure differs
from the cheapest way to get the same numbers:
  * fast  -- evaluates the Jakes sum only at the 14 symbol-start instants and
             builds one Delaunay plan per pilot pattern (used by the tests);
  * faithful -- keeps the reference's work: full 15344-sample time vector per
             (path, tx, rx), two griddata calls per antenna pair, one Np x Np
             inverse per antenna pair (used only to time the CPU baseline).
Both produce the same arrays (tests/test_oracle_golden.py checks it).
"""

from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------
# Tapped-delay-line profiles (3GPP TS 36.104 Annex B.2 EPA/EVA/ETU; the
# reference tabulates them at src/channel_simulator.py:41-54).
# --------------------------------------------------------------------------
TDL_NS = {
    "EPA": (0, 30, 70, 90, 110, 190, 410),
    "EVA": (0, 30, 150, 310, 370, 710, 1090, 1730, 2510),
    "ETU": (0, 50, 120, 200, 230, 500, 1600, 2300, 5000),
}
TDL_DB = {
    "EPA": (0.0, -1.0, -2.0, -3.0, -8.0, -17.2, -20.8),
    "EVA": (0.0, -1.5, -1.4, -3.6, -0.6, -9.1, -7.0, -12.0, -16.9),
    "ETU": (-1.0, -1.0, -1.0, 0.0, 0.0, 0.0, -3.0, -5.0, -7.0),
}
N_OSC = 20  # sum-of-sinusoids oscillators, src/channel_simulator.py:100


def tdl_profile(model: str, sampling_rate: float):
    """Normalised linear path powers and integer sample delays.

    src/channel_simulator.py:67-82.  `model` is upper-cased as there; unknown
    names raise KeyError like the reference's dict lookup (:72).
    """
    key = model.upper()
    tau = np.asarray(TDL_NS[key], dtype=np.float64) * 1e-9
    p_db = np.asarray(TDL_DB[key], dtype=np.float64)
    p_lin = 10.0 ** (p_db / 10.0)
    p_lin = p_lin / p_lin.sum()
    d = np.round(tau * sampling_rate).astype(int)
    return {"delays": tau, "powers_db": p_db, "powers_linear": p_lin, "delay_samples": d}


def surviving_taps(delay_samples):
    """Which path owns each tap after the reference's overwrite semantics.

    The tap write at src/channel_simulator.py:125 is an assignment, so when two
    paths round to the same sample delay the later path replaces the earlier.
    Returns (tap_delay[T], owner_path[T]) sorted by delay.
    """
    owner = {}
    for p, d in enumerate(delay_samples):
        owner[int(d)] = p
    delays = sorted(owner)
    return np.asarray(delays, dtype=np.int64), np.asarray([owner[d] for d in delays], dtype=np.int64)


def used_bins(fft_size: int, useful: int):
    """Shifted-domain indices of the occupied subcarriers, DC removed.

    src/channel_simulator.py:141-148.
    """
    dc = fft_size // 2
    idx = np.arange(dc - useful // 2, dc + useful // 2)
    return idx[idx != dc]


# --------------------------------------------------------------------------
# Pilot pattern and resource grid (src/channel_simulator.py:209-256)
# --------------------------------------------------------------------------
def pilot_layout(perm, nsc: int, nsym: int, density: float):
    """From the shuffled arange (the one RNG draw at :227-228) to indices/mask.

    Returns (pilot_indices[Np] int64 sorted, (sym_idx, sc_idx), mask[nsym,nsc] bool).
    """
    total = nsc * nsym
    n_p = int(total * density)
    chosen = np.sort(np.asarray(perm)[:n_p])
    pos = np.unravel_index(chosen, (nsym, nsc))
    mask = np.zeros((nsym, nsc), dtype=bool)
    mask[pos] = True
    return chosen, pos, mask


def build_grid(mask, pilot_symbols, data_symbols):
    """Row-major fill of pilots then data, src/channel_simulator.py:249-252."""
    g = np.zeros(mask.shape, dtype=complex)
    g[mask] = pilot_symbols
    g[~mask] = data_symbols
    return g


# --------------------------------------------------------------------------
# Jakes sum-of-sinusoids fading (src/channel_simulator.py:84-127)
# --------------------------------------------------------------------------
def jakes_cir(model: str, doppler_hz: float, sampling_rate: float, jakes_u,
              sample_idx, ntx: int, nrx: int):
    """CIR at the requested sample instants (fast profile).

    jakes_u[P, ntx, nrx, 2, 20] holds the raw U[0,1) draws in the reference's
    order (path-major, then tx, then rx; angles before phases, :102-110).
    Returns h[len(sample_idx), nrx, ntx, max_delay+1] complex128.
    """
    prof = tdl_profile(model, sampling_rate)
    d = prof["delay_samples"]
    amp = np.sqrt(prof["powers_linear"])
    t = np.asarray(sample_idx, dtype=np.float64) / sampling_rate
    out = np.zeros((t.size, nrx, ntx, int(d.max()) + 1), dtype=complex)
    ju = np.asarray(jakes_u, dtype=np.float64)
    for p in range(len(d)):
        theta = 2 * np.pi * ju[p, :, :, 0, :]          # [ntx, nrx, 20]
        phi = 2 * np.pi * ju[p, :, :, 1, :]
        shift = doppler_hz * np.cos(theta)
        arg = 2 * np.pi * shift[None] * t[:, None, None, None] + phi[None]
        g = (np.cos(arg).sum(-1) + 1j * np.sin(arg).sum(-1)) / np.sqrt(2 * N_OSC)
        out[:, :, :, d[p]] = amp[p] * np.transpose(g, (0, 2, 1))   # assignment: later path wins
    return out


def jakes_cir_faithful(model, doppler_hz, sampling_rate, jakes_u, num_samples, ntx, nrx):
    """Same numbers as jakes_cir(..., arange(num_samples)) with the reference's
    cost structure: one 20-term accumulation over the whole time vector for each
    (path, tx, rx) triple (:102-125).  Only used to time the CPU baseline."""
    prof = tdl_profile(model, sampling_rate)
    d = prof["delay_samples"]
    amp = np.sqrt(prof["powers_linear"])
    out = np.zeros((num_samples, nrx, ntx, int(d.max()) + 1), dtype=complex)
    t = np.arange(num_samples) / sampling_rate
    ju = np.asarray(jakes_u, dtype=np.float64)
    for p in range(len(d)):
        for a in range(ntx):
            for b in range(nrx):
                theta = 2 * np.pi * ju[p, a, b, 0]
                phi = 2 * np.pi * ju[p, a, b, 1]
                re = np.zeros(num_samples)
                im = np.zeros(num_samples)
                for n in range(N_OSC):
                    w = 2 * np.pi * (doppler_hz * np.cos(theta[n])) * t + phi[n]
                    re += np.cos(w)
                    im += np.sin(w)
                out[:, b, a, d[p]] = amp[p] * ((re + 1j * im) / np.sqrt(2 * N_OSC))
    return out


def cfr_from_cir(h_sym, fft_size: int, used):
    """CIR per symbol -> CFR on the used bins.

    src/channel_simulator.py:300-309: zero-padded FFT, fftshift, gather.
    h_sym[nsym, nrx, ntx, L] -> H[nsym, nrx, ntx, len(used)].
    """
    spec = np.fft.fftshift(np.fft.fft(h_sym, n=fft_size, axis=-1), axes=-1)
    return spec[..., used]


def channel_frequency_response(model, doppler_hz, cfg, jakes_u, ntx, nrx, faithful=False):
    """MIMOChannel.generate_channel_frequency_response, src/channel_simulator.py:274-311."""
    nfft, cp, nsym = cfg["fft_size"], cfg["cp_length"], cfg["num_symbols"]
    fs = nfft * cfg["subcarrier_spacing"]
    step = nfft + cp
    starts = np.arange(nsym) * step
    if faithful:
        h = jakes_cir_faithful(model, doppler_hz, fs, jakes_u, nsym * step, ntx, nrx)[starts]
    else:
        h = jakes_cir(model, doppler_hz, fs, jakes_u, starts, ntx, nrx)
    return cfr_from_cir(h, nfft, used_bins(nfft, cfg["useful_subcarriers"]))


def apply_channel(tx, H, snr_db, noise_re, noise_im):
    """y = H x per resource element + AWGN scaled by the slot's measured power.

    src/channel_simulator.py:326-345.  tx[nsym, ntx, nsc], H[nsym, nrx, ntx, nsc];
    noise_re / noise_im are the two randn blocks (:342), real block first.
    """
    y = np.einsum("srtk,stk->srk", H, tx)
    p_sig = np.mean(np.abs(y) ** 2)
    sigma = np.sqrt(p_sig / 10 ** (snr_db / 10) / 2)
    return y + (np.asarray(noise_re) + 1j * np.asarray(noise_im)) * sigma


def simulate(cfg, ntx, nrx, model, doppler_hz, snr_db, density, draws, faithful=False):
    """simulate_transmission, src/channel_simulator.py:348-421, on injected draws.

    cfg: dict with fft_size, cp_length, num_symbols, useful_subcarriers,
    subcarrier_spacing.  draws: perm, pilot_phase, data_phase (radians),
    jakes_u, noise_re, noise_im.
    """
    nfft, nsym = cfg["fft_size"], cfg["num_symbols"]
    used = used_bins(nfft, cfg["useful_subcarriers"])
    nsc = used.size
    pidx, ppos, mask = pilot_layout(draws["perm"], nsc, nsym, density)
    x_p = np.exp(1j * np.asarray(draws["pilot_phase"], dtype=np.float64))
    x_d = np.exp(1j * np.asarray(draws["data_phase"], dtype=np.float64))
    grid = build_grid(mask, x_p, x_d)
    tx = np.repeat(grid[:, None, :], ntx, axis=1)          # same grid on every TX (:402-404)
    H = channel_frequency_response(model, doppler_hz, cfg, draws["jakes_u"], ntx, nrx, faithful)
    rx = apply_channel(tx, H, snr_db, draws["noise_re"], draws["noise_im"])
    return {"tx_symbols": tx, "rx_symbols": rx, "channel": H, "pilot_indices": pidx,
            "pilot_positions": ppos, "pilot_mask": mask, "pilot_symbols": x_p}


# --------------------------------------------------------------------------
# LS + interpolation (src/baseline_estimators.py:23-117)
# --------------------------------------------------------------------------
LS_EPS = 1e-12  # added to the (complex) pilot before dividing, :40 / :110 / :171


def ls_at_pilots(rx_grid, x_p, mask):
    """h_p = y_p / (x_p + 1e-12), row-major gather (:109-110)."""
    return rx_grid[mask] / (x_p + LS_EPS)


def query_points(nsym, nsc):
    """Row-major (sym, sc) query list, equal to the meshgrid/transposed form at :68."""
    s, k = np.divmod(np.arange(nsym * nsc), nsc)
    return np.column_stack([s, k])


def linear_plan(pilot_positions, nsym, nsc):
    """Barycentric interpolation plan equal to griddata(method='linear', fill_value=0).

    griddata -> LinearNDInterpolator -> Qhull Delaunay of the pilot points,
    point location, barycentric weights; queries outside the hull get the fill
    value.  Returns (idx[nsym*nsc, 3] int64, w[nsym*nsc, 3] float64); rows
    outside the hull have idx 0 and w 0.  Call sites: src/baseline_estimators.py:65-79.
    """
    from scipy.spatial import Delaunay
    pts = np.column_stack([pilot_positions[0], pilot_positions[1]]).astype(np.float64)
    tri = Delaunay(pts)
    q = query_points(nsym, nsc).astype(np.float64)
    simplex = tri.find_simplex(q)
    inside = simplex >= 0
    sidx = np.where(inside, simplex, 0)
    T = tri.transform[sidx]                                  # [n, 3, 2]
    b2 = np.einsum("nij,nj->ni", T[:, :2, :], q - T[:, 2, :])
    bary = np.concatenate([b2, 1.0 - b2.sum(axis=1, keepdims=True)], axis=1)
    idx = tri.simplices[sidx].astype(np.int64)
    idx[~inside] = 0
    bary[~inside] = 0.0
    return idx, bary


def nearest_plan(pilot_positions, nsym, nsc):
    """1-tap plan equal to griddata(method='nearest').

    NearestNDInterpolator builds scipy.spatial.KDTree (leafsize 10, not cKDTree's 16); pilots sit
    on an integer lattice so equidistant ties are common and the tree shape decides them."""
    from scipy.spatial import KDTree
    pts = np.column_stack([pilot_positions[0], pilot_positions[1]]).astype(np.float64)
    _, i = KDTree(pts).query(query_points(nsym, nsc).astype(np.float64))
    return i.astype(np.int64)


def plan_apply(idx, w, h_p, nsym, nsc):
    return (w * h_p[idx]).sum(axis=1).reshape(nsym, nsc)


def griddata_grid(h_p, pilot_positions, nsym, nsc, method):
    """The reference's own route: two griddata calls, real then imag (:65-81)."""
    from scipy.interpolate import griddata
    pts = np.column_stack([pilot_positions[0], pilot_positions[1]])
    q = query_points(nsym, nsc)
    re = griddata(pts, h_p.real, q, method=method, fill_value=0.0).reshape(nsym, nsc)
    im = griddata(pts, h_p.imag, q, method=method, fill_value=0.0).reshape(nsym, nsc)
    return re + 1j * im


def ls_estimate(rx4d, x_p, mask, pilot_positions, method="linear", faithful=False):
    """LSEstimator.estimate, src/baseline_estimators.py:83-117.

    rx4d[nsym, nrx, ntx, nsc] (callers replicate rx over tx,
    src/dataset_generator.py:63-64)."""
    nsym, nrx, ntx, nsc = rx4d.shape
    out = np.zeros(rx4d.shape, dtype=complex)
    plan = None
    if not faithful and method == "linear":
        plan = linear_plan(pilot_positions, nsym, nsc)
    for r in range(nrx):
        for t in range(ntx):
            h_p = ls_at_pilots(rx4d[:, r, t, :], x_p, mask)
            if plan is not None:
                out[:, r, t, :] = plan_apply(plan[0], plan[1], h_p, nsym, nsc)
            else:
                out[:, r, t, :] = griddata_grid(h_p, pilot_positions, nsym, nsc, method)
    return out


# --------------------------------------------------------------------------
# MMSE (src/baseline_estimators.py:155-270)
# --------------------------------------------------------------------------
def mmse_at_pilots(h_ls, snr_db, cov=None, faithful=False):
    """estimate_at_pilots, :169-196.  cov=None is the default
    (estimate_statistics=True) branch whose R_h = P*I with P = mean|h_ls|^2."""
    sigma2 = 1.0 / 10 ** (snr_db / 10)
    n = h_ls.size
    if cov is None:
        p = np.mean(np.abs(h_ls) ** 2)
        if not faithful:
            return (p / (p + sigma2)) * h_ls           # W = alpha*I exactly
        R = np.eye(n) * p
    else:
        R = np.asarray(cov)
    Ry = R + sigma2 * np.eye(n)
    try:
        W = R @ np.linalg.inv(Ry)
    except np.linalg.LinAlgError:
        W = R @ np.linalg.inv(Ry + 1e-6 * np.eye(n))
    return W @ h_ls


def wiener_matrix(cov, snr_db):
    """W = R (R + sigma^2 I)^-1 for a user covariance (:182-189)."""
    sigma2 = 1.0 / 10 ** (snr_db / 10)
    R = np.asarray(cov)
    return R @ np.linalg.inv(R + sigma2 * np.eye(R.shape[0]))


def mmse_estimate(rx4d, x_p, mask, pilot_positions, snr_db, cov=None, faithful=False):
    """MMSEEstimator.estimate, :232-270 (interpolation is always 'linear', :200)."""
    nsym, nrx, ntx, nsc = rx4d.shape
    out = np.zeros(rx4d.shape, dtype=complex)
    plan = None if faithful else linear_plan(pilot_positions, nsym, nsc)
    for r in range(nrx):
        for t in range(ntx):
            h_ls = ls_at_pilots(rx4d[:, r, t, :], x_p, mask)
            h_m = mmse_at_pilots(h_ls, snr_db, cov, faithful)
            if plan is not None:
                out[:, r, t, :] = plan_apply(plan[0], plan[1], h_m, nsym, nsc)
            else:
                out[:, r, t, :] = griddata_grid(h_m, pilot_positions, nsym, nsc, "linear")
    return out


def evaluate(H_true, H_est):
    """evaluate_estimator, src/baseline_estimators.py:326-337."""
    mse = np.mean(np.abs(H_true - H_est) ** 2)
    nmse = mse / (np.mean(np.abs(H_true) ** 2) + 1e-12)
    return {"mse": mse, "nmse": nmse, "nmse_db": 10 * np.log10(nmse + 1e-12)}


def nmse_pair00(H_est, H_true):
    """compute_nmse on antenna pair (0,0), run_phase8_pilot_optimization.py:32-37,149-154."""
    a, b = H_est[:, 0, 0, :], H_true[:, 0, 0, :]
    return np.mean(np.abs(a - b) ** 2) / (np.mean(np.abs(b) ** 2) + 1e-10)


# --------------------------------------------------------------------------
# OFDM modulate / demodulate (src/channel_simulator.py:150-203)
# --------------------------------------------------------------------------
def ofdm_modulate(symbols, fft_size, cp, useful):
    """[nsym, nsc] -> [nsym, fft_size+cp]: map, ifftshift, IFFT*sqrt(N), CP prepend."""
    used = used_bins(fft_size, useful)
    f = np.zeros((symbols.shape[0], fft_size), dtype=complex)
    f[:, used] = symbols
    t = np.fft.ifft(np.fft.ifftshift(f, axes=-1), axis=-1) * np.sqrt(fft_size)
    return np.concatenate([t[:, fft_size - cp:], t], axis=1)


def ofdm_demodulate(signal, fft_size, cp, useful):
    """[nsym, fft_size+cp] -> [nsym, nsc]: CP strip, FFT/sqrt(N), fftshift, gather."""
    used = used_bins(fft_size, useful)
    f = np.fft.fftshift(np.fft.fft(signal[:, cp:], axis=-1), axes=-1) / np.sqrt(fft_size)
    return f[:, used]


# --------------------------------------------------------------------------
# One whole slot of the benchmark pipeline (simulate + LS + MMSE), used by the
# CPU baseline.  Mirrors what quick_start.py:47-94 does per slot.
# --------------------------------------------------------------------------
def slot_pipeline(cfg, ntx, nrx, model, doppler_hz, snr_db, density, draws, faithful=False):
    sim = simulate(cfg, ntx, nrx, model, doppler_hz, snr_db, density, draws, faithful)
    rx4d = np.repeat(sim["rx_symbols"][:, :, None, :], ntx, axis=2)
    H_ls = ls_estimate(rx4d, sim["pilot_symbols"], sim["pilot_mask"], sim["pilot_positions"],
                       "linear", faithful)
    H_mm = mmse_estimate(rx4d, sim["pilot_symbols"], sim["pilot_mask"], sim["pilot_positions"],
                         snr_db, None, faithful)
    sim["H_ls"], sim["H_mmse"] = H_ls, H_mm
    return sim


def random_draws(rng, cfg, ntx, nrx, model, density):
    """Draws from a numpy Generator in the shapes simulate() expects (tests, CPU baseline)."""
    nsym = cfg["num_symbols"]
    nsc = used_bins(cfg["fft_size"], cfg["useful_subcarriers"]).size
    total = nsym * nsc
    n_p = int(total * density)
    P = len(TDL_NS[model.upper()])
    return {
        "perm": rng.permutation(total),
        "pilot_phase": rng.uniform(0, 2 * np.pi, n_p),
        "data_phase": rng.uniform(0, 2 * np.pi, total - n_p),
        "jakes_u": rng.random((P, ntx, nrx, 2, N_OSC)),
        "noise_re": rng.standard_normal((nsym, nrx, nsc)),
        "noise_im": rng.standard_normal((nsym, nrx, nsc)),
    }


# --------------------------------------------------------------------------
# "next" rows either side of the path: equaliser,
# QAM mapping, BER, ML feature packing.  Pinned by tests/golden/link_level.npz.
# --------------------------------------------------------------------------
def equalize(rx, H, method="zf"):
    """src/baseline_estimators.py:273-312: x = inv(H^H H + lam I) H^H y per RE, lam = 1e-8 ('zf', :297)
    or 0.01 ('mmse', :305-306).  rx [nsym,nrx,nsc], H [nsym,nrx,ntx,nsc] -> [nsym,ntx,nsc]."""
    if method not in ("zf", "mmse"):
        raise ValueError(f"Unknown equalization method: {method}")
    lam = 1e-8 if method == "zf" else 0.01
    Hm = np.moveaxis(np.asarray(H, dtype=complex), 3, 1)             # [nsym, nsc, nrx, ntx]
    y = np.moveaxis(np.asarray(rx, dtype=complex), 2, 1)[..., None]  # [nsym, nsc, nrx, 1]
    Hh = np.conj(np.swapaxes(Hm, -1, -2))
    A = Hh @ Hm + lam * np.eye(Hm.shape[-1])
    x = np.linalg.inv(A) @ Hh @ y
    return np.moveaxis(x[..., 0], 1, 2)


_QAM_LEVELS = np.array([-3.0, -1.0, 3.0, 1.0])


def qam_tables(M):
    """(constellation, gray_map) as listed at src/utils.py:91-103 / :126-138."""
    if M == 4:
        return np.array([1 + 1j, -1 + 1j, 1 - 1j, -1 - 1j]) / np.sqrt(2), np.array([0, 1, 3, 2])
    if M == 16:
        const = (_QAM_LEVELS[:, None] + 1j * _QAM_LEVELS[None, :]).reshape(-1) / np.sqrt(10)
        return const, np.array([0, 1, 3, 2, 4, 5, 7, 6, 12, 13, 15, 14, 8, 9, 11, 10])
    raise NotImplementedError(f"Modulation order {M} not implemented")


def qam_modulate(bits, M=4):
    """src/utils.py:71-108 as intended: the reference's `gray_map[decimal_values]` (:106) indexes a list
    with an ndarray and raises TypeError under NumPy 2; the evident meaning is the array lookup below
    (it is the inverse of qam_demodulation, which is how the golden file pins it)."""
    const, gray = qam_tables(M)
    bps = int(np.log2(M))
    bits = np.asarray(bits).reshape(-1)
    n = len(bits) // bps
    dec = bits[:n * bps].reshape(-1, bps) @ (2 ** np.arange(bps)[::-1])
    return const[gray[dec]]


def qam_demodulate(symbols, M=4):
    """src/utils.py:111-152: argmin |s - c| (first minimum), argsort(gray) back to decimal, MSB-first bits."""
    const, gray = qam_tables(M)
    bps = int(np.log2(M))
    symbols = np.asarray(symbols).reshape(-1)
    det = np.argmin(np.abs(symbols[:, None] - const[None, :]), axis=1)
    dec = np.argsort(gray)[det]
    return ((dec[:, None] >> np.arange(bps)[::-1]) & 1).reshape(-1)


def bit_error_rate(a, b):
    """src/utils.py:155-157."""
    a, b = np.asarray(a), np.asarray(b)
    return np.sum(a != b) / len(a)


def ber_approximation(H_est, H_true, snr_db):
    """run_phase5_evaluation.py:57-68: QPSK BER proxy from the estimation NMSE (power eps 1e-10, :40-42)."""
    nmse = np.mean(np.abs(H_est - H_true) ** 2) / (np.mean(np.abs(H_true) ** 2) + 1e-10)
    snr = 10 ** (snr_db / 10)
    return float(np.clip(0.5 * np.exp(-(snr / (1 + snr * nmse)) / 2), 1e-10, 0.5))


def ml_inputs(rx, H_ls, H_true, mask, normalize=True):
    """prepare_ml_inputs, src/dataset_generator.py:183-227: pair (0,0), channel-last."""
    c2r = lambda a: np.stack([a.real, a.imag], axis=-1)
    x = np.concatenate([c2r(rx[:, 0, :]), c2r(H_ls[:, 0, 0, :]), mask.astype(float)[..., None]], axis=-1)
    t = c2r(H_true[:, 0, 0, :])
    if normalize:
        x[..., :4] = x[..., :4] / (np.std(x[..., :4]) + 1e-8)
        t = t / (np.std(t) + 1e-8)
    return x, t


def dataset_norm(rx, H_ls, H_true):
    """ChannelDataset._compute_normalization_stats, src/train.py:41-57, on stacked arrays [N, ...]:
    (mean, std) for rx, H_ls, H_true pair-(0,0) rows = mean of the re / im means and of the re / im stds."""
    out = []
    for a in (rx[:, :, 0, :], H_ls[:, :, 0, 0, :], H_true[:, :, 0, 0, :]):
        out.append((np.mean([a.real.mean(), a.imag.mean()]), np.mean([a.real.std(), a.imag.std()])))
    return out


def dataset_item(rx, H_ls, H_true, mask, norm=None):
    """ChannelDataset.__getitem__, src/train.py:62-94, for one sample: channel-first float32 (5, nsym, nsc) /
    (2, nsym, nsc); norm = dataset_norm(...) or None."""
    c2r = lambda a: np.stack([a.real, a.imag], axis=0)
    parts = [c2r(rx[:, 0, :]), c2r(H_ls[:, 0, 0, :]), c2r(H_true[:, 0, 0, :])]
    if norm is not None:
        parts = [(p - m) / (s + 1e-8) for p, (m, s) in zip(parts, norm)]
    x = np.concatenate([parts[0], parts[1], mask[None].astype(float)], axis=0)
    return x.astype(np.float32), parts[2].astype(np.float32)


def snr_sweep_baselines(H_true, H_ls, snr_db):
    """run_phase5_evaluation.py:264-312 without a model: per-sample NMSE (eps 1e-10) of H_ls and of
    alpha * H_ls, alpha = 1/(1 + 1/snr), pair (0,0); 10 log10(mean + 1e-12) per SNR value."""
    nm = lambda e, t: np.mean(np.abs(e - t) ** 2) / (np.mean(np.abs(t) ** 2) + 1e-10)
    values = sorted(set(float(s) for s in snr_db))
    res = {s: ([], []) for s in values}
    for i in range(len(snr_db)):
        t, l = H_true[i, :, 0, 0, :], H_ls[i, :, 0, 0, :]
        a = 1 / (1 + 1 / 10 ** (float(snr_db[i]) / 10))
        res[float(snr_db[i])][0].append(nm(l, t))
        res[float(snr_db[i])][1].append(nm(a * l, t))
    db = lambda v: 10 * np.log10(np.mean(v) + 1e-12)
    return values, [db(res[s][0]) for s in values], [db(res[s][1]) for s in values]
