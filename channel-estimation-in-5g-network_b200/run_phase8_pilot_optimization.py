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
                              channel_type: str = 'EVA', doppler_hz: float = 50.0, with_ber: bool = False,
                              ber_samples: int = None) -> dict:
        """NMSE vs pilot density (run_phase8_pilot_optimization.py:110-208).  Every cell also carries 'ber_proxy', the
        mean of compute_ber_approximation (run_phase5_evaluation.py:57-68) over its samples.  with_ber=True (Philox
        mode) adds 'ber': a measured bit-error rate per cell -- a QPSK grid through the same slot pipeline, ZF
        equalisation with the cell's LS / MMSE estimate (equalize_channel), demapping and bit counting on the data
        resource elements, `ber_samples` slots per cell (default min(num_samples, 64)); the method 'PERFECT' (the true
        channel as the estimate) is the curve's floor."""
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
        ber = None
        if with_ber:
            if self.rng != 'philox':
                raise ValueError("with_ber needs rng='philox' (the QPSK grid is a Philox-mode option)")
            nb = int(ber_samples) if ber_samples is not None else min(int(num_samples), 64)
            Bb = nd * ns * nb
            cellb = np.arange(Bb) // nb
            ber = {k: np.zeros(nd * ns, np.int64) for k in ('H_ls', 'H_mmse', 'H_true', 'bits')}
            pos = 0
            while pos < Bb:
                n = min(512, Bb - pos)
                sl = slice(pos, pos + n)
                r = eng.ber_batch(n, 0, float(doppler_hz), np.asarray(snr_values, np.float32)[cellb[sl] % ns],
                                  (cellb[sl] // ns).astype(np.int32), pool, slot0=(1 << 40) + pos, seed=self.seed)
                for k in ('H_ls', 'H_mmse', 'H_true'):
                    np.add.at(ber[k], cellb[sl], r['errors'][k].cpu().numpy().astype(np.int64))
                np.add.at(ber['bits'], cellb[sl], r['bits'])
                pos += n
        b = bins.cpu().numpy()
        summary = {'pilot_densities': pilot_densities, 'snr_values': snr_values, 'methods': {'LS': {}, 'MMSE': {}}}
        if ber is not None:
            summary['methods']['PERFECT'] = {snr: {dens: {'ber': float(ber['H_true'][di * ns + si] / max(ber['bits'][di * ns + si], 1))}
                                                   for di, dens in enumerate(pilot_densities)} for si, snr in enumerate(snr_values)}
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
                        if ber is not None:
                            c = di * ns + si
                            summary['methods'][name][snr][dens]['ber'] = float(ber['H_ls' if name == 'LS' else 'H_mmse'][c] / max(ber['bits'][c], 1))
        return summary
