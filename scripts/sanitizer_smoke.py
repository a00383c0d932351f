"""Small run of every libb2c entry point for compute-sanitizer (memcheck): kept tiny on purpose."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "channel-estimation-in-5g-network_b200")):
    sys.path.insert(0, p)
from engine import SlotEngine  # noqa: E402


def cfg(ntx, nrx, nsym=14, useful=600):
    return {"ofdm": {"fft_size": 1024, "cp_length": 72, "num_symbols": nsym, "useful_subcarriers": useful, "subcarrier_spacing": 15000},
            "mimo": {"num_tx_antennas": ntx, "num_rx_antennas": nrx}}


for ntx, nrx, nsym, useful in ((4, 4, 14, 600), (3, 2, 7, 300), (1, 1, 14, 600)):
    eng = SlotEngine(cfg(ntx, nrx, nsym, useful))
    pool = eng.random_pool([0.05, 0.10], seed=1)
    B = 3
    out = eng.run(B, [0, 1, 2], [10.0, 50.0, 200.0], [0.0, 10.0, 30.0], [0, 1, 1], pool, slot0=5, seed=3)
    part = eng.run(B, [0, 1, 2], 50.0, 10.0, [0, 1, 1], pool, want=("H_true", "rx", "H_ls"), compact=True)
    sim = eng.run(B, 2, 200.0, 10.0, want=("H_true", "rx", "tx"))
    bins = eng.stats_bins(out["stats"], np.array([0, 1, -1], np.int32), 2)
    xp = out["tx"][:, :, 0].reshape(B, -1)[:, torch.from_numpy(pool.pilot_indices[1]).to(eng.device)].contiguous()
    k3 = eng.ls_interp(out["rx"], xp, pool, pattern_id=1, snr_db=10.0, mmse=True, H_true=out["H_true"],
                       want=("H_ls", "H_mmse", "hp", "stats"))
    rx2 = eng.apply_channel(out["tx"], out["H_true"], [0.0, 10.0, 20.0], seed=4)
    npil = int(pool.npilots_host[1])
    W = torch.randn(npil, npil, dtype=torch.complex64, device=eng.device)
    hm = eng.mmse_dense(W, k3["hp"].reshape(B * nrx, -1)[:, :npil].contiguous())
    pv = eng.pilot_vectors(k3["hp"][0, :, :npil].contiguous(), xp[0, :npil].contiguous(), snr_db=5.0, mmse=True)
    h = eng.tdl_full("EVA", 50.0, 700, ntx, nrx)
    torch.cuda.synchronize()
    print("ok", ntx, nrx, nsym, useful, float(bins[:, 0].sum()))
e1 = SlotEngine(cfg(1, 1))
x = torch.randn(20, 599, dtype=torch.complex64, device=e1.device)
y = e1.ofdm_demodulate(e1.ofdm_modulate(x))
torch.cuda.synchronize()
print("ofdm roundtrip", float((x - y).abs().max()))
