"""Drop-in for the reference's verify_phase3_datasets.py (:23-188): integrity and statistics of a stacked
dataset file written by run_phase3_dataset_generation / run_phase3_robust (here or by the reference).

Same result dictionary and status values ('failed', 'incomplete', 'shape_mismatch', 'data_errors', 'success').
The array-sized work -- NaN/Inf counts over rx_symbols / H_ls / H_true and the per-sample LS NMSE of antenna
pair (0,0) -- runs on the GPU (b2c_count_nonfinite, b2c_pair00_errors); the parameter histograms are host-side
bookkeeping on N scalars.
"""

from __future__ import annotations

import argparse
from pathlib import Path

import numpy as np
import torch

from utils import linear2db

REQUIRED_KEYS = ['rx_symbols', 'tx_symbols', 'H_ls', 'H_true', 'pilot_mask', 'snr_db', 'channel_type', 'doppler_hz']


def verify_dataset(filepath: str, detailed: bool = False, antennas=(2, 2), verbose: bool = True) -> dict:
    """verify_phase3_datasets.py:23-188.  `antennas` = (ntx, nrx) the shape check expects; the reference
    hard-codes its 2x2 default (:68-74)."""
    from baseline_estimators import _engine
    from _b2c import Geom
    say = print if verbose else (lambda *a, **k: None)
    try:
        data = np.load(filepath, allow_pickle=True)
    except Exception as e:
        return {'status': 'failed', 'error': str(e)}
    results = {'status': 'success', 'filepath': filepath}
    missing = [k for k in REQUIRED_KEYS if k not in data.keys()]
    if missing:
        results.update(status='incomplete', missing_keys=missing)
        return results
    arrs = {k: data[k] for k in ('rx_symbols', 'tx_symbols', 'H_ls', 'H_true', 'pilot_mask')}
    n = arrs['rx_symbols'].shape[0]
    results['num_samples'] = n
    ntx, nrx = antennas
    expected = {'rx_symbols': (n, 14, nrx, 599), 'tx_symbols': (n, 14, ntx, 599), 'H_ls': (n, 14, nrx, ntx, 599),
                'H_true': (n, 14, nrx, ntx, 599), 'pilot_mask': (n, 14, 599)}
    bad = [k for k, shp in expected.items() if arrs[k].shape != shp]
    for k, shp in expected.items():
        say(f"    {'ok ' if arrs[k].shape == shp else 'BAD'} {k}: {arrs[k].shape} (expected: {shp})")
    if bad:
        results.update(status='shape_mismatch', shape_errors=bad)

    eng = _engine()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(np.complex64, copy=False))).to(eng.device)
    counts = torch.zeros((2,), dtype=torch.int64, device=eng.device)
    on_dev = {}
    for k in ('rx_symbols', 'H_ls', 'H_true'):
        on_dev[k] = dev(arrs[k])
        eng.count_nonfinite(on_dev[k], counts)
    nan_count, inf_count = (int(v) for v in counts.cpu())
    if nan_count or inf_count:
        results.update(status='data_errors', nan_count=nan_count, inf_count=inf_count)

    snr = np.asarray(data['snr_db'])
    results['snr_range'] = [float(snr.min()), float(snr.max())]
    results['channel_types'] = list(np.unique(data['channel_type']))
    dop = np.asarray(data['doppler_hz'])
    results['doppler_range'] = [float(dop.min()), float(dop.max())]
    if 'pilot_density' in data.keys():
        results['pilot_densities'] = list(np.unique(data['pilot_density']))
    if detailed:
        for v, c in zip(*np.unique(snr, return_counts=True)):
            say(f"      {v:5.0f} dB: {c:5d} samples ({100 * c / n:5.1f}%)")

    # LS quality on up to 10 random samples (:159-170), calculate_nmse's 1e-12 epsilon, mean of the dB values
    if arrs['H_true'].ndim == 5 and arrs['H_ls'].shape == arrs['H_true'].shape and n > 0:
        idx = np.random.choice(n, min(10, n), replace=False)
        _, nsym, r, t, nsc = arrs['H_true'].shape
        sel = torch.from_numpy(np.sort(idx)).to(eng.device)
        err = eng.pair00_errors(on_dev['H_ls'][sel].contiguous(), on_dev['H_true'][sel].contiguous(),
                                geom=Geom(nsym, nsc, t, r, 1024, 72, 0.0)).cpu().numpy()
        m = nsym * nsc
        nmse = (err[:, 0] / m) / (err[:, 2] / m + 1e-12)
        results['avg_ls_nmse_db'] = float(np.mean([linear2db(v) for v in nmse]))
    mask = np.asarray(arrs['pilot_mask'][:min(10, n)], dtype=np.float64)
    results['avg_pilot_density'] = float(np.mean(mask.reshape(mask.shape[0], -1).mean(axis=1))) if n else 0.0
    say(f"  {'VERIFICATION PASSED' if results['status'] == 'success' else 'VERIFICATION FAILED: ' + results['status']}")
    return results


def main():
    ap = argparse.ArgumentParser(description='Verify Phase 3 Datasets')
    ap.add_argument('--data-dir', type=str, default='data')
    ap.add_argument('--detailed', action='store_true')
    ap.add_argument('--file', type=str, default=None)
    ap.add_argument('--antennas', type=int, nargs=2, default=(2, 2), metavar=('NTX', 'NRX'))
    args = ap.parse_args()
    files = [Path(args.file)] if args.file else sorted(Path(args.data_dir).glob('*.npz'))
    if not files:
        print(f"No .npz files found in {args.data_dir}")
        return
    ok = True
    for f in files:
        r = verify_dataset(str(f), detailed=args.detailed, antennas=tuple(args.antennas))
        print(f"  {f.stem}: {r.get('num_samples', 0)} samples - {r['status']}")
        ok = ok and r['status'] == 'success'
    print("All datasets verified successfully!" if ok else "Some datasets failed verification")


if __name__ == '__main__':
    main()
