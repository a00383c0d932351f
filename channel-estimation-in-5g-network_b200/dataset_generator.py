"""Drop-in for the reference's dataset batch loop (src/dataset_generator.py:23-180 and its script
twins run_phase3_dataset_generation.py:29-205), executed as batched libb2c launches.

Two random-number modes:
  rng='numpy'  (default) -- draws come from the global numpy.random stream in exactly the
      reference's per-sample order (4 x choice, shuffle, 2 x uniform, Jakes rand, 2 x randn), so a
      seeded run returns the reference's samples; every sample has its own pilot pattern, hence
      its own host-built Delaunay plan (about 5 ms each -- this mode is for parity, not speed).
  rng='philox' -- counter-based draws keyed by (seed, global sample index) on the device, pilot
      patterns from a fixed pool: results are independent of batch size and of how samples are
      sharded over GPUs; this is the throughput mode (see generate_batch / sharded_statistics).
"""

from __future__ import annotations

from pathlib import Path
from typing import Dict, List, Optional

import numpy as np
import torch

import _b2c
import _tables
from engine import SlotEngine

_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
STREAM_PARAMS = 3


def philox_param_choice(seed: int, slot0: int, count: int, sizes):
    """Per-slot (model, doppler, snr, density) indices from Philox stream 3 (include/b2c.h):
    word i of counter (0, 3, slot lo, slot hi) -> (word * n_i) >> 32.  Host-side twin of the
    reference's four np.random.choice calls (src/dataset_generator.py:114-117)."""
    slot = np.arange(slot0, slot0 + count, dtype=np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    c = [np.zeros(count, np.uint64), np.full(count, STREAM_PARAMS, np.uint64), slot & mask, slot >> np.uint64(32)]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = np.uint64(_M0) * c[0], np.uint64(_M1) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0), p1 & mask, (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1), p0 & mask]
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return [((c[i] * np.uint64(n)) >> np.uint64(32)).astype(np.int64) for i, n in enumerate(sizes)]


class ChannelEstimationDataset:
    """Dataset generator with the reference's interface (src/dataset_generator.py:23-180)."""

    def __init__(self, config: Dict, rng: str = 'numpy', seed: int = 42, batch_size: int = 256,
                 patterns_per_density: int = 1, lists=None, out_dtype=np.complex128):
        self.config = config
        self.ofdm_config = config['ofdm']
        self.mimo_config = config['mimo']
        self.channel_config = config['channel']
        self.dataset_config = config.get('dataset', {})
        if rng not in ('numpy', 'philox'):
            raise ValueError(f"Unknown rng mode: {rng}")
        self.rng, self.seed, self.batch_size = rng, seed, batch_size
        self.patterns_per_density = patterns_per_density
        self._list_override = lists          # (models, dopplers, snrs, densities) of the script twins
        self.out_dtype = out_dtype           # complex128 like src/dataset_generator.py; the twins use complex64
        self._engine: Optional[SlotEngine] = None
        self._pool = None
        self._next_slot = 0

    # The Philox pattern pool is a function of the seed: re-seeding (the script twins do it per split) drops the pool,
    # so a split's pilot patterns never depend on which split this object generated before it -- a 'val' split resumed
    # in a fresh process gets the patterns of the uninterrupted train + val run.
    @property
    def seed(self):
        return self._seed

    @seed.setter
    def seed(self, value):
        if getattr(self, "_seed", None) != value:
            self._pool = None
        self._seed = value

    # ---- parameter lists (src/dataset_generator.py:105-108) -------------------------------------------
    def _lists(self):
        if self._list_override is not None:
            return tuple(list(v) for v in self._list_override)
        return (list(self.channel_config['models']), list(self.channel_config['doppler_hz']),
                list(self.config['simulation']['snr_range']), list(self.config['pilots']['density']))

    @property
    def engine(self) -> SlotEngine:
        if self._engine is None:
            self._engine = SlotEngine(self.config, models=tuple(str(m).upper() for m in self._lists()[0]))
        return self._engine

    # ---- one sample, reference draw order -----------------------------------------------------------------
    def generate_sample(self, channel_type: str, doppler_hz: float, snr_db: float, pilot_density: float) -> Dict:
        """simulate_transmission + LS('linear') for one slot (src/dataset_generator.py:34-89)."""
        return self._numpy_batch([(channel_type, doppler_hz, snr_db, pilot_density)], draw_params=False)[0]

    def generate_dataset(self, num_samples: int, split: str = 'train') -> List[Dict]:
        """List of sample dicts with randomly chosen conditions (src/dataset_generator.py:91-127)."""
        print(f"Generating {split} dataset with {num_samples} samples...")
        out: List[Dict] = []
        done = 0
        while done < num_samples:
            n = min(self.batch_size, num_samples - done)
            if self.rng == 'numpy':
                out += self._numpy_batch([None] * n, draw_params=True, first_index=done)
            else:
                out += self._philox_samples(n)
            done += n
        return out

    def _numpy_batch(self, params, draw_params, first_index=0, want_stats=False):
        eng = self.engine
        nsym, nsc, ntx, nrx = eng.nsym, eng.nsc, eng.ntx, eng.nrx
        models, dopplers, snrs, dens = self._lists() if draw_params else (None, None, None, None)
        meta, pats, turns, jakes, noise = [], [], [], [], []
        for i, p in enumerate(params):
            if draw_params:     # four choice() draws per sample, in the reference's order (:114-117)
                p = (np.random.choice(models), np.random.choice(dopplers), np.random.choice(snrs), np.random.choice(dens))
            ch, fd, snr, density = p
            total = nsym * nsc
            order = np.arange(total)
            np.random.shuffle(order)
            idx = np.sort(order[:int(total * density)])
            mask = np.zeros(total, dtype=bool)
            mask[idx] = True
            pilot_phase = np.random.uniform(0, 2 * np.pi, len(idx))
            data_phase = np.random.uniform(0, 2 * np.pi, total - len(idx))
            npaths = len(_tables.PDP[str(ch).upper()][0])
            ju = np.zeros((eng.p_max, ntx, nrx, 2, _tables.N_OSC))
            ju[:npaths] = np.random.rand(npaths, ntx, nrx, 2, _tables.N_OSC)
            z = np.random.randn(2, nsym, nrx, nsc)
            t = np.empty(total)
            t[mask] = pilot_phase / (2 * np.pi)
            t[~mask] = data_phase / (2 * np.pi)
            try:
                _tables.cached_plan(idx, nsym, nsc, 'linear')
            except Exception as exc:   # degenerate pilot set: the reference skips the sample (:123-125)
                print(f"Warning: Failed to generate sample {first_index + i}: {exc}")
                continue
            meta.append((ch, fd, snr, density, mask.reshape(nsym, nsc)))
            pats.append(idx)
            turns.append(t.reshape(nsym, nsc))
            jakes.append(ju)
            noise.append(z[0] + 1j * z[1])
        if not meta:
            return []
        B = len(meta)
        dev = eng.device
        pool = eng.pool(pats, 'linear')
        inject = {"jakes_u": torch.from_numpy(np.stack(jakes)).to(dev, torch.float32),
                  "sym_turns": torch.from_numpy(np.stack(turns)).to(dev, torch.float32),
                  "noise": torch.from_numpy(np.stack(noise)).to(dev, torch.complex64)}
        out = eng.run(B, [eng.models.index(str(m[0]).upper()) for m in meta], [float(m[1]) for m in meta],
                      [float(m[2]) for m in meta], np.arange(B), pool, inject=inject,
                      want=("stats",) if want_stats else ("H_true", "rx", "tx", "H_ls"))
        if want_stats:
            return out["stats"]
        host = {k: out[k].cpu().numpy().astype(self.out_dtype) for k in ("H_true", "rx", "tx", "H_ls")}
        return [{'rx_symbols': host["rx"][i], 'tx_symbols': host["tx"][i], 'H_ls': host["H_ls"][i],
                 'H_true': host["H_true"][i], 'pilot_mask': meta[i][4], 'snr_db': meta[i][2],
                 'channel_type': meta[i][0], 'doppler_hz': meta[i][1], 'pilot_density': meta[i][3]}
                for i in range(B)]

    # ---- throughput mode ---------------------------------------------------------------------------------
    def pattern_pool(self):
        if self._pool is None:
            self._pool = self.engine.random_pool(self._lists()[3], self.patterns_per_density, seed=self.seed)
        return self._pool

    def generate_batch(self, count: int, slot0: int = 0, want=("H_true", "rx", "tx", "H_ls"), out=None, ws=None, pitch=None,
                       **run_kwargs):
        """`count` slots starting at global sample index slot0, Philox mode, all on the device.
        Returns (dict of CUDA tensors, dict of per-slot parameter index arrays).  pitch=600: row-padded
        throughput layout (tensors are 599-wide views), see SlotEngine.run.  Further keyword arguments go to
        SlotEngine.run (compact=True, mmse="dense" + wiener=WienerBank for the known-covariance MMSE, qpsk=True)."""
        eng = self.engine
        models, dopplers, snrs, dens = self._lists()
        mi, di, si, pi = philox_param_choice(self.seed, slot0, count, (len(models), len(dopplers), len(snrs), len(dens)))
        # the pattern within a density's pool rotates with the sample index
        pid = pi * self.patterns_per_density + (np.arange(slot0, slot0 + count) % self.patterns_per_density)
        model_id = np.array([eng.models.index(str(m).upper()) for m in models], dtype=np.int32)[mi]
        res = eng.run(count, model_id, np.asarray(dopplers, dtype=np.float32)[di], np.asarray(snrs, dtype=np.float32)[si],
                      pid.astype(np.int32), self.pattern_pool(), slot0=slot0, seed=self.seed, want=want, out=out, ws=ws, pitch=pitch,
                      **run_kwargs)
        return res, {"model": mi, "doppler": di, "snr": si, "density": pi, "pattern": pid}

    def generate_feature_batch(self, count: int, slot0: int = 0, layout: str = "last", normalize: bool = True,
                               norm=None):
        """Training features straight from GPU-resident slots, skipping disk:
        `count` Philox slots -> (inputs, targets, params) with inputs = (rx re, rx im, H_ls re, H_ls im,
        pilot mask) and targets = H_true (re, im) of antenna pair (0,0), float32 CUDA tensors.
        layout='last' + normalize is prepare_ml_inputs per slot (src/dataset_generator.py:183-227);
        layout='first' + norm (see SlotEngine.normalization_from_moments) is ChannelDataset.__getitem__
        (src/train.py:62-94)."""
        wide = self.engine.nsc == 599 and self.engine.nsym % 2 == 0 and self.engine.ntx in (1, 2, 4, 8)
        res, par = self.generate_batch(count, slot0, want=("H_true", "rx", "tx", "H_ls"), pitch=600 if wide else None)
        x, y = self.engine.ml_features(res["rx"], res["H_ls"], res["H_true"], self.pattern_pool(),
                                       par["pattern"].astype(np.int32), layout, normalize, norm)
        return x, y, par

    def _philox_samples(self, n):
        models, dopplers, snrs, dens = self._lists()
        res, par = self.generate_batch(n, self._next_slot)
        self._next_slot += n
        pool = self.pattern_pool()
        host = {k: res[k].cpu().numpy() for k in ("H_true", "rx", "tx", "H_ls")}
        return [{'rx_symbols': host["rx"][i], 'tx_symbols': host["tx"][i], 'H_ls': host["H_ls"][i],
                 'H_true': host["H_true"][i], 'pilot_mask': pool.mask(int(par["pattern"][i])),
                 'snr_db': snrs[par["snr"][i]], 'channel_type': models[par["model"][i]],
                 'doppler_hz': dopplers[par["doppler"][i]], 'pilot_density': dens[par["density"][i]]} for i in range(n)]

    # ---- writers (src/dataset_generator.py:129-180) ---------------------------------------------------------
    def save_dataset(self, dataset: List[Dict], filepath: str, format: str = 'npz'):
        if format not in ('npz', 'h5'):
            raise ValueError(f"Unknown format: {format}")
        keys = ('rx_symbols', 'tx_symbols', 'H_ls', 'H_true', 'pilot_mask')
        stacked = {k: np.stack([s[k] for s in dataset]) for k in keys}
        for k in ('snr_db', 'channel_type', 'doppler_hz', 'pilot_density'):
            stacked[k] = np.array([s[k] for s in dataset])
        if format == 'npz':
            np.savez_compressed(filepath, **stacked)
        else:
            import h5py   # optional dependency, as in the reference
            with h5py.File(filepath, 'w') as fh:
                for k, v in stacked.items():
                    fh.create_dataset(k, data=v.astype('S10') if k == 'channel_type' else v)
        print(f"Dataset saved to {filepath}")


def sharded_statistics(config: Dict, total_slots: int, rank: int = 0, world_size: int = 1, batch: int = 8192,
                       seed: int = 42, bin_by: str = "snr", want_arrays=(), on_batch=None, dataset: Optional[ChannelEstimationDataset] = None,
                       **run_kwargs):
    """Multi-GPU dataset statistics: rank r simulates + estimates global samples
    [r*N/R, (r+1)*N/R) in batches and folds MSE/NMSE into per-bin float64 accumulators on its
    GPU; the caller all-reduces the returned [nbins, 14] tensor (see reduce_bins).  Philox keyed
    by the global sample index makes the result independent of world_size.  run_kwargs go to SlotEngine.run
    (e.g. mmse="dense", wiener=WienerBank: statistics of the known-covariance MMSE estimator)."""
    ds = dataset if dataset is not None else ChannelEstimationDataset(config, rng='philox', seed=seed)
    eng = ds.engine
    models, dopplers, snrs, dens = ds._lists()
    sizes = {"snr": len(snrs), "model": len(models), "density": len(dens), "doppler": len(dopplers)}
    nbins = sizes[bin_by]
    lo, hi = shard_range(total_slots, rank, world_size)
    bins = torch.zeros((nbins, _b2c.N_BINSTAT), dtype=torch.float64, device=eng.device)
    want = tuple(set(("stats",) + tuple(want_arrays)))
    out = ws = None
    pos = lo
    while pos < hi:
        n = min(batch, hi - pos)
        if out is None or n != batch:
            # H_true + rx + tx + H_ls requested (the dataset arrays; H_mmse optional): the row-padded throughput layout
            # (wide-store kernel); callers see 599-wide views
            wide = all(k in want for k in ("H_true", "rx", "tx", "H_ls"))
            out, ws = eng.alloc_outputs(n, want, pitch=600 if (wide and eng.nsc == 599 and eng.nsym % 2 == 0 and eng.ntx in (1, 2, 4, 8)) else None), eng.workspace(n)
        res, par = ds.generate_batch(n, pos, want=want, out=out, ws=ws, **run_kwargs)
        eng.stats_bins(res["stats"], par[bin_by].astype(np.int32), nbins, bins, snr_db=np.asarray(snrs, np.float32)[par["snr"]])
        if on_batch is not None:
            on_batch(pos, res, par)
        pos += n
    return bins


def shard_range(total: int, rank: int, world_size: int):
    """Contiguous block of global sample indices owned by `rank`."""
    return (total * rank) // world_size, (total * (rank + 1)) // world_size


def reduce_bins(bins: torch.Tensor) -> torch.Tensor:
    """Sum per-bin accumulators over ranks: the path's only collective (NCCL on GPUs, gloo in CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(bins, op=dist.ReduceOp.SUM)
    return bins


def summarize_bins(bins) -> List[Dict]:
    """Per-bin means in the reference's units: mse / nmse / nmse_db (evaluate_estimator) and the
    pair-(0,0) NMSE mean / std / dB of run_phase8_pilot_optimization.py:186-206."""
    b = bins.detach().cpu().numpy() if isinstance(bins, torch.Tensor) else np.asarray(bins)
    rows = []
    for r in b:
        n = max(r[0], 1.0)
        m = {"count": int(r[0]), "mse_ls": r[1] / n, "mse_mmse": r[2] / n, "nmse_ls": r[3] / n, "nmse_mmse": r[4] / n,
             "nmse00_ls_mean": r[8] / n, "nmse00_mmse_mean": r[10] / n}
        m["nmse_ls_db"] = 10 * np.log10(m["nmse_ls"] + 1e-12)
        m["nmse_mmse_db"] = 10 * np.log10(m["nmse_mmse"] + 1e-12)
        m["mse_ls_db"] = 10 * np.log10(m["mse_ls"] + 1e-12)
        m["mse_mmse_db"] = 10 * np.log10(m["mse_mmse"] + 1e-12)
        m["nmse00_ls_db"] = 10 * np.log10(m["nmse00_ls_mean"] + 1e-12)
        m["nmse00_ls_std"] = float(np.sqrt(max(r[9] / n - (r[8] / n) ** 2, 0.0)))
        if len(r) > 13:      # QPSK BER proxy of run_phase5_evaluation.py:57-68 (folded when the SNRs were given)
            m["ber_proxy_ls"], m["ber_proxy_mmse"] = r[12] / n, r[13] / n
        rows.append(m)
    return rows


_FEATURE_POOLS = {}


def prepare_ml_inputs(sample: Dict, normalize: bool = True):
    """(inputs [nsym, nsc, 5], targets [nsym, nsc, 2]) real-valued arrays from one dataset sample
    (src/dataset_generator.py:183-227): first RX / TX antenna, channels (rx re, rx im, H_ls re, H_ls im,
    pilot mask); with `normalize` the four signal channels are divided by their joint std + 1e-8 and the
    targets by theirs.  Packed and reduced on the GPU (b2c_ml_features)."""
    from engine import PatternPool
    from _b2c import Geom
    rx, H_ls, H_true = np.asarray(sample['rx_symbols']), np.asarray(sample['H_ls']), np.asarray(sample['H_true'])
    nsym, nrx, nsc = rx.shape
    ntx = H_ls.shape[2]
    mask = np.asarray(sample['pilot_mask']).astype(bool)
    idx = np.flatnonzero(mask.reshape(-1))
    dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
    key = (dev, nsym, nsc)
    if key not in _FEATURE_POOLS:
        from baseline_estimators import _engine
        _FEATURE_POOLS[key] = _engine()
    eng = _FEATURE_POOLS[key]
    # only the pilot positions of the pool are read by the feature kernel; the identity plan keeps this cheap
    pool = PatternPool.__new__(PatternPool)
    pool.device, pool.nsym, pool.nsc, pool.np_max = eng.device, nsym, nsc, max(1, idx.size)
    pool.npilots = torch.tensor([idx.size], dtype=torch.int32, device=eng.device)
    pool.pilot_re = torch.zeros((1, pool.np_max), dtype=torch.int32, device=eng.device)
    pool.pilot_re[0, :idx.size] = torch.from_numpy(idx.astype(np.int32)).to(eng.device)
    from _b2c import Patterns
    pool.struct = Patterns(1, pool.np_max, pool.npilots.data_ptr(), pool.pilot_re.data_ptr(), None)
    c64 = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.complex64))[None]).to(eng.device)
    g = Geom(nsym, nsc, ntx, nrx, 1024, 72, 0.0)
    x, y = eng.ml_features(c64(rx), c64(H_ls), c64(H_true), pool, 0, "last", normalize, None, geom=g)
    return x[0].cpu().numpy().astype(np.float64), y[0].cpu().numpy().astype(np.float64)


def main():
    """CLI with the reference's flags (src/dataset_generator.py:230-315)."""
    import argparse
    from utils import load_config, set_seed
    ap = argparse.ArgumentParser(description='Generate channel estimation dataset (B200)')
    ap.add_argument('--config', type=str, default='configs/experiment_config.yaml')
    ap.add_argument('--output-dir', type=str, default='data')
    ap.add_argument('--train-samples', type=int, default=None)
    ap.add_argument('--val-samples', type=int, default=None)
    ap.add_argument('--test-samples', type=int, default=None)
    ap.add_argument('--format', type=str, default='npz', choices=['npz', 'h5'])
    ap.add_argument('--seed', type=int, default=42)
    ap.add_argument('--rng', type=str, default='numpy', choices=['numpy', 'philox'])
    args = ap.parse_args()
    set_seed(args.seed)
    config = load_config(args.config)
    for split in ('train', 'val', 'test'):
        v = getattr(args, f'{split}_samples')
        if v is not None:
            config['dataset'][f'{split}_samples'] = v
    out_dir = Path(args.output_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    gen = ChannelEstimationDataset(config, rng=args.rng, seed=args.seed)
    for split in ('train', 'val', 'test'):
        data = gen.generate_dataset(config['dataset'][f'{split}_samples'], split=split)
        gen.save_dataset(data, out_dir / f'{split}.{args.format}', format=args.format)


if __name__ == '__main__':
    main()
