"""Drop-in for the reference's run_phase3_dataset_generation.py (DatasetGenerator, :29-226): fixed
parameter lists, per-split seeds, complex64 / float32 / 'U3' stacked outputs -- executed as batched
libb2c launches.

rng='numpy' (default) follows the reference's global-RNG draw order exactly (set_seed, one throw-away
probe sample :122, then 4 x choice + the sample's draws per iteration); rng='philox' is the
throughput mode (pattern pool, counter-based draws, same dict layout).
"""

from __future__ import annotations

import os

import numpy as np

from dataset_generator import ChannelEstimationDataset
from utils import load_config, set_seed

CHANNEL_TYPES = ['EPA', 'EVA', 'ETU']
DOPPLER_VALUES = [10, 50, 100, 200]
SNR_VALUES = [-5, 0, 5, 10, 15, 20, 25, 30]
PILOT_DENSITIES = [0.05, 0.10]
SPLIT_SEEDS = {'train': 42, 'val': 123, 'test': 456}
ARRAY_KEYS = ('rx_symbols', 'tx_symbols', 'H_ls', 'H_true')


def _typed_sample(s: dict) -> dict:
    """The twins' per-sample dtypes (run_phase3_dataset_generation.py:72-82, run_phase3_robust.py:82-93)."""
    return {**{k: s[k].astype(np.complex64) for k in ARRAY_KEYS},
            'pilot_mask': s['pilot_mask'].astype(np.float32), 'snr_db': s['snr_db'],
            'channel_type': s['channel_type'], 'doppler_hz': s['doppler_hz'], 'pilot_density': s['pilot_density']}


def stack_samples(samples) -> dict:
    """Stacked dataset dict with the dtypes of run_phase3_dataset_generation.py:135-143."""
    n = len(samples)
    out = {k: (np.stack([s[k] for s in samples]).astype(np.complex64) if n else np.zeros((0,), np.complex64))
           for k in ARRAY_KEYS}
    out['pilot_mask'] = np.stack([s['pilot_mask'] for s in samples]).astype(np.float32) if n else np.zeros((0,), np.float32)
    out['snr_db'] = np.array([s['snr_db'] for s in samples], dtype=np.float32)
    out['channel_type'] = np.array([str(s['channel_type']) for s in samples], dtype='U3')
    out['doppler_hz'] = np.array([s['doppler_hz'] for s in samples], dtype=np.float32)
    out['pilot_density'] = np.array([s['pilot_density'] for s in samples], dtype=np.float32)
    return out


class DatasetGenerator:
    """Phase-3 dataset generator with the reference twin's interface."""

    def __init__(self, config_path: str = 'configs/experiment_config.yaml', rng: str = 'numpy', batch_size: int = 256):
        self.config = load_config(config_path)
        self.rng, self.batch_size = rng, batch_size
        self._impl = None

    def _dataset(self, seed=42):
        if self._impl is None:
            self._impl = ChannelEstimationDataset(self.config, rng=self.rng, seed=seed, batch_size=self.batch_size,
                                                  lists=(CHANNEL_TYPES, DOPPLER_VALUES, SNR_VALUES, PILOT_DENSITIES),
                                                  out_dtype=np.complex64)
        return self._impl

    def generate_sample(self, channel_type: str, doppler_hz: float, snr_db: float, pilot_density: float) -> dict:
        """simulate_transmission + LS('linear') (run_phase3_dataset_generation.py:37-82)."""
        return _typed_sample(self._dataset().generate_sample(channel_type, doppler_hz, snr_db, pilot_density))

    def generate_dataset(self, num_samples: int, split: str = 'train', seed: int = None) -> dict:
        """Dict of stacked arrays (run_phase3_dataset_generation.py:84-205)."""
        if seed is None:
            seed = SPLIT_SEEDS.get(split, 42)
        set_seed(seed)
        ds = self._dataset(seed)
        ds.seed, ds._next_slot = seed, 0
        if self.rng == 'numpy':
            self.generate_sample('EPA', 50.0, 10.0, 0.1)      # the reference's shape probe consumes RNG (:122)
        samples = ds.generate_dataset(num_samples, split)
        for s in samples:    # the twins store python floats for the chosen parameters (:153-155)
            s['doppler_hz'], s['snr_db'], s['pilot_density'] = float(s['doppler_hz']), float(s['snr_db']), float(s['pilot_density'])
        print(f"\n  Generated {len(samples)} samples successfully")
        return stack_samples([_typed_sample(s) for s in samples])

    def save_dataset(self, dataset: dict, filepath: str):
        np.savez_compressed(filepath, **dataset)
        print(f"Saved to {filepath} ({os.path.getsize(filepath) / 2 ** 20:.1f} MB)")

    def print_dataset_info(self, dataset: dict, name: str):
        print(f"\n{name} Dataset Info:\n  Samples: {dataset['rx_symbols'].shape[0]}\n  RX symbols shape: {dataset['rx_symbols'].shape}"
              f"\n  H_true shape: {dataset['H_true'].shape}\n  SNR range: [{dataset['snr_db'].min():.0f}, {dataset['snr_db'].max():.0f}] dB"
              f"\n  Channel types: {np.unique(dataset['channel_type'])}")


def main():
    import argparse
    ap = argparse.ArgumentParser(description='Phase 3: dataset generation (B200)')
    ap.add_argument('--config', type=str, default='configs/experiment_config.yaml')
    ap.add_argument('--output-dir', type=str, default='data')
    ap.add_argument('--train-samples', type=int, default=10000)
    ap.add_argument('--val-samples', type=int, default=2000)
    ap.add_argument('--test-samples', type=int, default=2000)
    ap.add_argument('--rng', type=str, default='numpy', choices=['numpy', 'philox'])
    args = ap.parse_args()
    os.makedirs(args.output_dir, exist_ok=True)
    gen = DatasetGenerator(args.config, rng=args.rng)
    for split, n in (('train', args.train_samples), ('val', args.val_samples), ('test', args.test_samples)):
        data = gen.generate_dataset(n, split)
        gen.save_dataset(data, os.path.join(args.output_dir, f'{split}.npz'))
        gen.print_dataset_info(data, split)


if __name__ == '__main__':
    main()
