"""Drop-in for the reference's run_phase3_robust.py (RobustDatasetGenerator, :35-310): chunked .npz
output with a JSON progress checkpoint and resume.

rng='numpy' keeps the reference's semantics, including its resume rule `set_seed(seed + start_idx)`
(:145-156), which is NOT bit-identical to an uninterrupted run.  rng='philox' keys every sample by
its global index, so a resumed run reproduces exactly the arrays an uninterrupted run would have
written (tests/test_gpu_dropin.py::test_robust_generator_resume_is_exact_in_philox_mode).
"""

from __future__ import annotations

import json
import time
from datetime import datetime
from pathlib import Path

import numpy as np

from run_phase3_dataset_generation import (CHANNEL_TYPES, DOPPLER_VALUES, PILOT_DENSITIES, SNR_VALUES, SPLIT_SEEDS,
                                           DatasetGenerator, _typed_sample, stack_samples)
from utils import set_seed

KEYS = ('rx_symbols', 'tx_symbols', 'H_ls', 'H_true', 'pilot_mask', 'snr_db', 'channel_type', 'doppler_hz', 'pilot_density')


class RobustDatasetGenerator(DatasetGenerator):
    def __init__(self, config_path: str = 'configs/experiment_config.yaml', output_dir: str = 'data',
                 rng: str = 'numpy', batch_size: int = 256):
        super().__init__(config_path, rng=rng, batch_size=batch_size)
        self.output_dir = Path(output_dir)
        self.checkpoint_dir = self.output_dir / 'checkpoints'
        self.checkpoint_dir.mkdir(parents=True, exist_ok=True)

    def get_checkpoint_path(self, split: str) -> Path:
        return self.checkpoint_dir / f'{split}_checkpoint.json'

    def get_temp_data_path(self, split: str, chunk_id: int) -> Path:
        return self.checkpoint_dir / f'{split}_chunk_{chunk_id:04d}.npz'

    def save_checkpoint(self, split: str, completed: int, total: int, chunk_id: int, start_time: float):
        elapsed = time.time() - start_time
        with open(self.get_checkpoint_path(split), 'w') as fh:
            json.dump({'split': split, 'completed': completed, 'total': total, 'chunk_id': chunk_id,
                       'start_time': start_time, 'timestamp': datetime.now().isoformat(),
                       'samples_per_second': completed / elapsed if completed > 0 and elapsed > 0 else 0}, fh, indent=2)

    def load_checkpoint(self, split: str):
        path = self.get_checkpoint_path(split)
        if path.exists():
            with open(path) as fh:
                return json.load(fh)
        return None

    def generate_dataset_chunked(self, num_samples: int, split: str = 'train', seed: int = None,
                                 chunk_size: int = 500, resume: bool = False, stop_after_chunks: int = None) -> dict:
        """Chunked generation (run_phase3_robust.py:126-259).  `stop_after_chunks` simulates an
        interruption (tests)."""
        if seed is None:
            seed = SPLIT_SEEDS.get(split, 42)
        ck = self.load_checkpoint(split) if resume else None
        start_idx, chunk_id = (ck['completed'], ck['chunk_id']) if ck and ck['completed'] < ck['total'] else (0, 0)
        set_seed(seed + start_idx)
        ds = self._dataset(seed)
        ds.seed, ds._next_slot = seed, start_idx           # philox: sample i is slot i, resumed or not
        if self.rng == 'numpy':
            self.generate_sample('EPA', 50.0, 10.0, 0.1)   # shape probe of the reference (:171)
        t0 = time.time()
        pos, written = start_idx, 0
        while pos < num_samples:
            n = min(chunk_size, num_samples - pos)
            samples = [_typed_sample(s) for s in ds.generate_dataset(n, split)]
            np.savez_compressed(self.get_temp_data_path(split, chunk_id), **stack_samples(samples))
            pos += n
            chunk_id += 1
            written += 1
            self.save_checkpoint(split, pos, num_samples, chunk_id, t0)
            if stop_after_chunks is not None and written >= stop_after_chunks and pos < num_samples:
                return {}
        merged = self._merge_chunks(split, chunk_id)
        np.savez_compressed(str(self.output_dir / f'{split}.npz'), **merged)
        self._cleanup_chunks(split, chunk_id)
        return merged

    def _merge_chunks(self, split: str, num_chunks: int) -> dict:
        parts = {k: [] for k in KEYS}
        for c in range(num_chunks):
            path = self.get_temp_data_path(split, c)
            if path.exists():
                with np.load(path, allow_pickle=True) as z:
                    for k in KEYS:
                        parts[k].append(z[k])
        return {k: np.concatenate(v, axis=0) for k, v in parts.items() if v}

    def _cleanup_chunks(self, split: str, num_chunks: int):
        for c in range(num_chunks):
            self.get_temp_data_path(split, c).unlink(missing_ok=True)
        self.get_checkpoint_path(split).unlink(missing_ok=True)
