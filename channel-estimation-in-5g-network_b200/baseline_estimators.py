"""Drop-in for the reference's src/baseline_estimators.py (LS and MMSE channel estimators),
computed by libb2c on a B200.

LSEstimator / MMSEEstimator / evaluate_estimator keep the reference's signatures
(src/baseline_estimators.py:10-337) and return complex128 NumPy arrays.  The 2-D interpolation
of `scipy.interpolate.griddata` is split into a host-built plan (Qhull Delaunay / KDTree, once per
pilot pattern, cached) and a GPU blend per resource element; see _tables.py.
"""

from __future__ import annotations

import hashlib
from typing import Optional, Tuple

import numpy as np
import torch

import _tables
from _b2c import Geom
from engine import PatternPool, SlotEngine

_ENGINE = {}


def _engine() -> SlotEngine:
    """Estimators only need a device and launchers; geometry is passed per call."""
    dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
    if dev not in _ENGINE:
        _ENGINE[dev] = SlotEngine({"ofdm": {"fft_size": 1024, "cp_length": 72, "num_symbols": 14,
                                            "useful_subcarriers": 600, "subcarrier_spacing": 15000.0},
                                   "mimo": {"num_tx_antennas": 1, "num_rx_antennas": 1}})
    return _ENGINE[dev]


def _c64(a, device):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.complex64))).to(device)


def _pilot_index(pilot_positions, nsc):
    s, k = pilot_positions
    return np.asarray(s, dtype=np.int64) * nsc + np.asarray(k, dtype=np.int64)


def _interp_pairs(h_pairs, pilot_positions, grid_shape, method):
    """h_pairs [npair, Np] complex -> [npair, nsym, nsc] by the pattern's interpolation plan."""
    nsym, nsc = grid_shape
    eng = _engine()
    idx = _pilot_index(pilot_positions, nsc)
    order = np.argsort(idx, kind="stable")      # plans are keyed on the sorted (row-major) pilot set
    npair = h_pairs.shape[0]
    g = Geom(nsym, nsc, 1, npair, 1024, 72, 0.0)
    hp = _c64(np.asarray(h_pairs)[:, order][None], eng.device)
    if method == 'cubic':
        W = torch.from_numpy(_tables.cubic_matrix(idx[order], nsym, nsc)).to(eng.device)
        grid = eng.dense_real_apply(W, hp[0].contiguous(), ld_out=nsym * nsc)
        return grid.reshape(npair, nsym, nsc).cpu().numpy().astype(np.complex128)
    pool = PatternPool([idx[order]], nsym, nsc, method, eng.device)
    out = eng.ls_interp(None, None, pool, hp_in=hp, want=("H_ls",), geom=g)
    return out["H_ls"][0, :, :, 0, :].permute(1, 0, 2).cpu().numpy().astype(np.complex128)


class LSEstimator:
    """Least-squares estimator: pilot division + 2-D interpolation (src/baseline_estimators.py:10-117)."""

    def __init__(self, interpolation_method: str = 'linear'):
        self.interpolation_method = interpolation_method

    def estimate_at_pilots(self, rx_symbols: np.ndarray, tx_pilots: np.ndarray, pilot_mask: np.ndarray) -> np.ndarray:
        """rx_symbols[pilot_mask] / (tx_pilots + 1e-12)  (:23-42)."""
        eng = _engine()
        y = _c64(np.asarray(rx_symbols)[pilot_mask].reshape(1, -1), eng.device)
        return eng.pilot_vectors(y, _c64(tx_pilots, eng.device))[0].cpu().numpy().astype(np.complex128)

    def interpolate_channel(self, h_pilots: np.ndarray, pilot_positions: Tuple, grid_shape: Tuple[int, int]) -> np.ndarray:
        """Pilot estimates -> full (num_symbols, num_subcarriers) grid, 0 outside the hull (:44-81)."""
        return _interp_pairs(np.asarray(h_pilots)[None], pilot_positions, grid_shape, self.interpolation_method)[0]

    def estimate(self, rx_symbols: np.ndarray, tx_pilots: np.ndarray, pilot_mask: np.ndarray,
                 pilot_positions: Tuple) -> np.ndarray:
        """rx_symbols (num_symbols, num_rx, num_tx, num_subcarriers) -> estimate of the same shape (:83-117).
        Every (rx, tx) slice is estimated independently, as in the reference."""
        return _estimate_4d(rx_symbols, tx_pilots, pilot_mask, self.interpolation_method, None)


def _estimate_4d(rx_symbols, tx_pilots, pilot_mask, method, snr_db):
    rx_symbols = np.asarray(rx_symbols)
    nsym, nrx, ntx, nsc = rx_symbols.shape
    eng = _engine()
    idx = np.flatnonzero(np.asarray(pilot_mask).reshape(-1))           # row-major = grid[mask] order
    pool = PatternPool([idx], nsym, nsc, method, eng.device) if method != 'cubic' else None
    # (rx, tx) pairs are independent "receive antennas" of a 1-TX problem
    g = Geom(nsym, nsc, 1, nrx * ntx, 1024, 72, 0.0)
    rx = _c64(rx_symbols.reshape(nsym, nrx * ntx, nsc)[None], eng.device)
    want = ("H_mmse",) if snr_db is not None else ("H_ls",)
    if method == 'cubic':    # dense Clough-Tocher map on the tensor cores (LS only, as in the reference)
        out = eng.ls_cubic(rx, _c64(np.asarray(tx_pilots).reshape(1, -1), eng.device), idx, want=want, geom=g)
    else:
        out = eng.ls_interp(rx, _c64(np.asarray(tx_pilots).reshape(1, -1), eng.device), pool,
                            snr_db=snr_db, mmse=snr_db is not None, want=want, geom=g)
    H = out[want[0]][0, :, :, 0, :]
    return H.reshape(nsym, nrx, ntx, nsc).cpu().numpy().astype(np.complex128)


class MMSEEstimator:
    """MMSE estimator (src/baseline_estimators.py:120-270).  Default branch
    (estimate_statistics=True or no covariance): R_h = P*I, i.e. H_mmse = P/(P+sigma^2) * H_ls.
    Known-covariance branch: W = R (R + sigma^2 I)^-1 built once per SNR on the host, applied to
    the whole batch of pilot vectors by the dense GEMM kernel."""

    def __init__(self, channel_covariance: Optional[np.ndarray] = None, noise_variance: float = 0.01,
                 estimate_statistics: bool = True):
        self.channel_covariance = channel_covariance
        self.noise_variance = noise_variance
        self.estimate_statistics = estimate_statistics
        self._w_cache = {}

    def estimate_covariance(self, h_ls: np.ndarray) -> np.ndarray:
        """Sample covariance over the last axis (:137-153).  Host-side statistic, unused by the pipeline."""
        return np.cov(h_ls.reshape(-1, h_ls.shape[-1]), rowvar=False)

    def _dense(self):
        return not (self.estimate_statistics or self.channel_covariance is None)

    def _wiener(self, device):
        """W = R (R + sigma^2 I)^-1 in float64 on the host, cached per noise variance (:182-194)."""
        # keyed on the covariance's CONTENT: a matrix mutated in place, or a new one at a recycled id(), is a new W
        R = np.asarray(self.channel_covariance)
        key = (hashlib.sha1(np.ascontiguousarray(R).view(np.uint8)).hexdigest(), R.shape, str(R.dtype), float(self.noise_variance))
        if key not in self._w_cache:
            Ry = R + self.noise_variance * np.eye(R.shape[0])
            try:
                W = R @ np.linalg.inv(Ry)
            except np.linalg.LinAlgError:
                W = R @ np.linalg.inv(Ry + 1e-6 * np.eye(R.shape[0]))
            # applied to every (rx, tx) pair of every call at this SNR: keep it as a prepared GEMM operand
            self._w_cache = {key: _engine().prepare_dense(_c64(W, device))}
        return self._w_cache[key]

    def estimate_at_pilots(self, rx_symbols: np.ndarray, tx_pilots: np.ndarray, pilot_mask: np.ndarray,
                           snr_db: float = 10) -> np.ndarray:
        """LS at the pilots followed by the Wiener filter (:155-196); updates noise_variance like the reference."""
        eng = _engine()
        self.noise_variance = 1.0 / (10 ** (snr_db / 10))
        y = _c64(np.asarray(rx_symbols)[pilot_mask].reshape(1, -1), eng.device)
        x = _c64(tx_pilots, eng.device)
        if not self._dense():
            h = eng.pilot_vectors(y, x, snr_db=snr_db, mmse=True)
        else:
            h = eng.mmse_dense(self._wiener(eng.device), eng.pilot_vectors(y, x))
        return h[0].cpu().numpy().astype(np.complex128)

    def interpolate_channel(self, h_pilots: np.ndarray, pilot_positions: Tuple, grid_shape: Tuple[int, int],
                            interpolation_method: str = 'linear') -> np.ndarray:
        return _interp_pairs(np.asarray(h_pilots)[None], pilot_positions, grid_shape, interpolation_method)[0]

    def estimate(self, rx_symbols: np.ndarray, tx_pilots: np.ndarray, pilot_mask: np.ndarray,
                 pilot_positions: Tuple, snr_db: float = 10) -> np.ndarray:
        """Full-grid MMSE estimate, interpolation always 'linear' (:232-270)."""
        self.noise_variance = 1.0 / (10 ** (snr_db / 10))
        if not self._dense():
            return _estimate_4d(rx_symbols, tx_pilots, pilot_mask, 'linear', float(snr_db))
        rx_symbols = np.asarray(rx_symbols)
        nsym, nrx, ntx, nsc = rx_symbols.shape
        eng = _engine()
        idx = np.flatnonzero(np.asarray(pilot_mask).reshape(-1))
        pool = PatternPool([idx], nsym, nsc, 'linear', eng.device)
        g = Geom(nsym, nsc, 1, nrx * ntx, 1024, 72, 0.0)
        rx = _c64(rx_symbols.reshape(nsym, nrx * ntx, nsc)[None], eng.device)
        hp = eng.ls_interp(rx, _c64(np.asarray(tx_pilots).reshape(1, -1), eng.device), pool, want=("hp",), geom=g)["hp"]
        hm = eng.mmse_dense(self._wiener(eng.device), hp.reshape(nrx * ntx, -1)).reshape(1, nrx * ntx, -1)
        H = eng.ls_interp(None, None, pool, hp_in=hm, want=("H_ls",), geom=g)["H_ls"][0, :, :, 0, :]
        return H.reshape(nsym, nrx, ntx, nsc).cpu().numpy().astype(np.complex128)


def equalize_channel(rx_symbols: np.ndarray, H_est: np.ndarray, method: str = 'zf') -> np.ndarray:
    """Per-RE ZF / MMSE equaliser (src/baseline_estimators.py:273-312):
    x = (H^H H + lambda I)^-1 H^H y with lambda = 1e-8 ('zf') or 0.01 ('mmse'), solved in fp64 on the
    GPU (b2c_equalize) on complex128 buffers, so the reference's dtype and conditioning are kept."""
    if method not in ('zf', 'mmse'):
        raise ValueError(f"Unknown equalization method: {method}")
    eng = _engine()
    H = torch.from_numpy(np.ascontiguousarray(np.asarray(H_est, dtype=np.complex128))[None]).to(eng.device)
    y = torch.from_numpy(np.ascontiguousarray(np.asarray(rx_symbols, dtype=np.complex128))[None]).to(eng.device)
    return eng.equalize(y, H, method)[0].cpu().numpy()


def evaluate_estimator(H_true: np.ndarray, H_est: np.ndarray) -> dict:
    """mse / nmse / nmse_db over the whole 4-D array (src/baseline_estimators.py:315-337), reduced on
    the GPU by the statistics kernels."""
    H_true = np.asarray(H_true)
    err, pw = squared_error_sums(H_true, H_est)
    n = H_true.size
    mse = err / n
    nmse = mse / (pw / n + 1e-12)
    return {'mse': mse, 'nmse': nmse, 'nmse_db': 10 * np.log10(nmse + 1e-12)}


def squared_error_sums(H_true, H_est, width: int = 599):
    """(sum |H_true - H_est|^2, sum |H_true|^2) over arrays of any shape, reduced on the GPU: the data
    are laid out as zero-padded rows of `width` and pushed through K3's statistics path."""
    eng = _engine()
    t = np.asarray(H_true, dtype=np.complex64).reshape(-1)
    e = np.asarray(H_est, dtype=np.complex64).reshape(-1)
    rows = max(1, -(-t.size // width))
    tp, ep = np.zeros(rows * width, np.complex64), np.zeros(rows * width, np.complex64)
    tp[:t.size], ep[:e.size] = t, e
    g = Geom(1, width, 1, 1, 1024, 72, 0.0)
    stats = _sq_error_stats(eng, _c64(tp.reshape(rows, 1, 1, 1, width), eng.device),
                            _c64(ep.reshape(rows, 1, 1, width), eng.device), g)
    tot = stats[:, :, 1].sum(dim=(0, 1)).cpu().numpy()
    return float(tot[0]), float(tot[2])


_IDENTITY_POOLS = {}


def _sq_error_stats(eng, H_true, H_est_as_rx, g):
    """sum|H - H_est|^2 and sum|H|^2 per row, via K3 with an identity plan (every RE is its own pilot,
    unit pilots): H_ls == H_est exactly, so the kernel's error statistics are the wanted sums."""
    nsc = g.nsc
    key = (nsc, eng.device)
    if key not in _IDENTITY_POOLS:
        pool = PatternPool.__new__(PatternPool)
        pool.device, pool.nsym, pool.nsc, pool.method = eng.device, 1, nsc, "identity"
        pool.pilot_indices = [np.arange(nsc)]
        pool.np_max = nsc
        plan = np.zeros(2 * (nsc + 1), dtype=_tables.PLAN_DTYPE)
        plan["i0"][:nsc + 1] = plan["i1"][:nsc + 1] = plan["i2"][:nsc + 1] = np.arange(nsc + 1)   # row nsc -> the zero slot
        plan["w0"], plan["flags"] = 1.0, 1
        pool.npilots = torch.tensor([nsc], dtype=torch.int32, device=eng.device)
        pool.pilot_re = torch.arange(nsc, dtype=torch.int32, device=eng.device).reshape(1, nsc)
        pool.plan = torch.from_numpy(plan.view(np.uint8).reshape(2, (nsc + 1) * 16)).to(eng.device)
        from _b2c import Patterns
        pool.struct = Patterns(1, nsc, pool.npilots.data_ptr(), pool.pilot_re.data_ptr(), pool.plan.data_ptr())
        ones = torch.ones((1, nsc), dtype=torch.complex64, device=eng.device)
        _IDENTITY_POOLS[key] = (pool, ones)
    pool, ones = _IDENTITY_POOLS[key]
    return eng.ls_interp(H_est_as_rx, ones, pool, H_true=H_true, want=("stats",), geom=g)["stats"]
