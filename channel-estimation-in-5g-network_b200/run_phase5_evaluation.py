"""Drop-in for the classical-baseline part of the reference's run_phase5_evaluation.py: the metric helpers
(:37-68) and ModelEvaluator's per-SNR LS / MMSE aggregation over a stored test set (:264-312).  The trained
models, their loading and the plots are the ML side and out of scope.

The per-sample error sums of antenna pair (0,0) are reduced on the GPU (b2c_pair00_errors); the host only
groups N scalars by SNR.
"""

from __future__ import annotations

import numpy as np
import torch

from utils import linear2db


def compute_nmse(H_est: np.ndarray, H_true: np.ndarray) -> float:
    """mean|H_est - H_true|^2 / (mean|H_true|^2 + 1e-10) (run_phase5_evaluation.py:37-42)."""
    from baseline_estimators import squared_error_sums
    n = np.asarray(H_true).size
    err, pw = squared_error_sums(H_true, H_est)
    return float((err / n) / (pw / n + 1e-10))


def compute_mse(H_est: np.ndarray, H_true: np.ndarray) -> float:
    """mean|H_est - H_true|^2 (run_phase5_evaluation.py:45-48)."""
    from baseline_estimators import squared_error_sums
    return float(squared_error_sums(H_true, H_est)[0] / np.asarray(H_true).size)


def compute_mae(H_est: np.ndarray, H_true: np.ndarray) -> float:
    """mean|H_est - H_true| (run_phase5_evaluation.py:51-54), reduced on the GPU."""
    from baseline_estimators import _c64, _engine
    eng = _engine()
    n = np.asarray(H_true).size
    return float(eng.abs_diff_sum(_c64(np.asarray(H_est).reshape(-1), eng.device), _c64(np.asarray(H_true).reshape(-1), eng.device)).item() / n)


def compute_ber_approximation(H_est: np.ndarray, H_true: np.ndarray, snr_db: float) -> float:
    """QPSK BER proxy from the estimation NMSE (run_phase5_evaluation.py:57-68)."""
    nmse = compute_nmse(H_est, H_true)
    snr_linear = 10 ** (snr_db / 10)
    effective_snr = snr_linear / (1 + snr_linear * nmse)
    return float(np.clip(0.5 * np.exp(-effective_snr / 2), 1e-10, 0.5))


def per_sample_baseline_nmse(test_data: dict, batch: int = 4096):
    """(nmse_ls [N], nmse_mmse [N]) of run_phase5_evaluation.py:283-296: pair-(0,0) rows, MMSE = alpha * H_ls with
    alpha = 1 / (1 + 1/snr_linear).  `test_data` holds the stacked dataset arrays (numpy or CUDA tensors)."""
    from baseline_estimators import _engine
    from _b2c import Geom
    eng = _engine()
    H_true, H_ls, snr = test_data['H_true'], test_data['H_ls'], np.asarray(test_data['snr_db'], dtype=np.float64)
    N, nsym, nrx, ntx, nsc = H_true.shape
    g = Geom(nsym, nsc, ntx, nrx, 1024, 72, 0.0)
    alpha = (1.0 / (1.0 + 1.0 / 10 ** (snr / 10))).astype(np.float32)
    out = np.empty((N, 3))
    dev = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(
        device=eng.device, dtype=torch.complex64).contiguous()
    for i in range(0, N, batch):
        j = min(N, i + batch)
        out[i:j] = eng.pair00_errors(dev(H_ls[i:j]), dev(H_true[i:j]), alpha[i:j], geom=g).cpu().numpy()
    n = nsym * nsc
    den = out[:, 2] / n + 1e-10
    return (out[:, 0] / n) / den, (out[:, 1] / n) / den


def snr_sweep_baselines(test_data: dict) -> dict:
    """ModelEvaluator.snr_sweep_analysis without a model (run_phase5_evaluation.py:264-312):
    {'snr_db': sorted values, 'methods': {'LS': {'nmse_db': [...]}, 'MMSE': {...}}}."""
    snr = np.asarray(test_data['snr_db'])
    values = sorted(set(float(s) for s in snr))
    ls, mm = per_sample_baseline_nmse(test_data)
    summary = {'snr_db': values, 'methods': {}}
    for name, v in (('LS', ls), ('MMSE', mm)):
        summary['methods'][name] = {'nmse_db': [linear2db(np.mean(v[snr == s])) for s in values]}
    return summary
