"""Drop-in for the hot-path part of the reference's run_phase8_pilot_optimization.py (PilotOptimizer
.generate_test_sample / .analyze_pilot_density, :71-208): pilot-density x SNR sweep of simulate + LS,
NMSE on antenna pair (0,0), mean / std / dB per cell.  The CNN branch and the plots are out of scope
(ML side); an 'MMSE' method is reported next to 'LS' because the fused kernel produces it for free.

rng='numpy' draws in the reference's loop order (density, snr, sample) with one pilot pattern per
sample; rng='philox' (default here: this is a statistics sweep) shares one pattern per density and
folds the per-slot errors into per-cell accumulators on the GPU (b2c_stats_bins).
"""

from __future__ import annotations

import numpy as np
import torch

from dataset_generator import ChannelEstimationDataset
from utils import linear2db, load_config


def compute_nmse(H_est: np.ndarray, H_true: np.ndarray) -> float:
    """mean|H_est - H_true|^2 / (mean|H_true|^2 + 1e-10) (run_phase8_pilot_optimization.py:32-37), on the GPU."""
    from baseline_estimators import squared_error_sums
    n = np.asarray(H_true).size
    err, pw = squared_error_sums(H_true, H_est)
    return float((err / n) / (pw / n + 1e-10))


class PilotOptimizer:
    def __init__(self, config_path: str = 'configs/experiment_config.yaml', rng: str = 'philox', seed: int = 42):
        self.config = load_config(config_path)
        self.rng, self.seed = rng, seed

    def generate_test_sample(self, pilot_density: float, snr_db: float = 15.0, channel_type: str = 'EVA',
                             doppler_hz: float = 50.0) -> dict:
        """One slot + LS('linear'), numpy draw order (run_phase8_pilot_optimization.py:71-108)."""
        ds = ChannelEstimationDataset(self.config, rng='numpy', lists=([channel_type], [doppler_hz], [snr_db], [pilot_density]))
        s = ds.generate_sample(channel_type, doppler_hz, snr_db, pilot_density)
        return {'rx_symbols': s['rx_symbols'], 'H_ls': s['H_ls'], 'H_true': s['H_true'], 'pilot_mask': s['pilot_mask'], 'snr_db': snr_db}

    def analyze_pilot_density(self, pilot_densities: list, snr_values: list = None, num_samples: int = 100,
                              channel_type: str = 'EVA', doppler_hz: float = 50.0) -> dict:
        """NMSE vs pilot density (run_phase8_pilot_optimization.py:110-208)."""
        if snr_values is None:
            snr_values = [5, 10, 15, 20]
        nd, ns = len(pilot_densities), len(snr_values)
        ds = ChannelEstimationDataset(self.config, rng=self.rng, seed=self.seed,
                                      lists=([channel_type], [doppler_hz], list(snr_values), list(pilot_densities)))
        eng = ds.engine
        bins = torch.zeros((nd * ns, 14), dtype=torch.float64, device=eng.device)
        if self.rng == 'philox':
            pool = ds.pattern_pool()
            B = nd * ns * num_samples
            cell = np.arange(B) // num_samples                     # cell = density * ns + snr
            dens_i, snr_i = cell // ns, cell % ns
            pos = 0
            while pos < B:
                n = min(4096, B - pos)
                sl = slice(pos, pos + n)
                out = eng.run(n, 0, float(doppler_hz), np.asarray(snr_values, np.float32)[snr_i[sl]], dens_i[sl].astype(np.int32),
                              pool, slot0=pos, seed=self.seed, want=("stats",))
                eng.stats_bins(out["stats"], cell[sl].astype(np.int32), nd * ns, bins, snr_db=np.asarray(snr_values, np.float32)[snr_i[sl]])
                pos += n
        else:
            for di, dens in enumerate(pilot_densities):
                for si, snr in enumerate(snr_values):
                    samples = ds._numpy_batch([(channel_type, doppler_hz, snr, dens)] * num_samples, draw_params=False,
                                              want_stats=True)
                    eng.stats_bins(samples, np.full(len(samples), di * ns + si, np.int32), nd * ns, bins, snr_db=float(snr))
        b = bins.cpu().numpy()
        summary = {'pilot_densities': pilot_densities, 'snr_values': snr_values, 'methods': {'LS': {}, 'MMSE': {}}}
        for name, (c1, c2) in (('LS', (8, 9)), ('MMSE', (10, 11))):
            for si, snr in enumerate(snr_values):
                summary['methods'][name][snr] = {}
                for di, dens in enumerate(pilot_densities):
                    r = b[di * ns + si]
                    if r[0] > 0:
                        mean = r[c1] / r[0]
                        summary['methods'][name][snr][dens] = {
                            'nmse_mean': float(mean), 'nmse_db': float(linear2db(mean)),
                            'nmse_std': float(np.sqrt(max(r[c2] / r[0] - mean * mean, 0.0))),
                            # the sweep's BER curve: mean over the cell of compute_ber_approximation (run_phase5_evaluation.py:57-68)
                            'ber_proxy': float(r[12 if name == 'LS' else 13] / r[0])}
        return summary
