"""Per-kernel micro-benchmarks on one B200 (CUDA events, 3 warm-ups, inputs >> L2 where it matters).
Writes one JSON object; the numbers go to profiles/kernels_rNN.json.  bench.py remains the headline."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "channel-estimation-in-5g-network_b200")):
    sys.path.insert(0, p)
from engine import SlotEngine  # noqa: E402

PEAK_HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3


def cfg(ntx, nrx):
    return {"ofdm": {"fft_size": 1024, "cp_length": 72, "num_symbols": 14, "useful_subcarriers": 600, "subcarrier_spacing": 15000},
            "mimo": {"num_tx_antennas": ntx, "num_rx_antennas": nrx}}


res = {"peak_hbm_gbs": PEAK_HBM}
eng = SlotEngine(cfg(4, 4))
dev = eng.device
pool = eng.random_pool([0.10], seed=42)

# K4 dense Wiener GEMM: W 838x838 complex, columns = slots x rx
npil = 838
W = (torch.randn(npil, npil, dtype=torch.complex64, device=dev) / np.sqrt(npil)).contiguous()
for ncols in (4096, 16384, 65536):
    h = torch.randn(ncols, npil, dtype=torch.complex64, device=dev)
    t = timeit(lambda: eng.mmse_dense(W, h), n=5)
    flops = 8.0 * npil * npil * ncols
    ref = (h[:64].to(torch.complex128) @ W.to(torch.complex128).T)
    err = ((eng.mmse_dense(W, h)[:64].to(torch.complex128) - ref).abs().max() / ref.abs().max()).item()
    res[f"k4_mmse_dense_{ncols}cols"] = {"ms": t * 1e3, "useful_tflops": flops / t / 1e12, "issued_tf32_tflops": 3 * flops / t / 1e12,
                                          "rel_err_vs_fp64": err}
    Wp = eng.prepare_dense(W)
    t = timeit(lambda: eng.mmse_dense(Wp, h), n=5)
    res[f"k4_mmse_dense_prepared_{ncols}cols"] = {"ms": t * 1e3, "useful_tflops": flops / t / 1e12,
                                                   "issued_tf32_tflops": 3 * flops / t / 1e12,
                                                   "note": "W pre-split into hi/lo smem tiles once (b2c_dense_prepare), A stages by bulk copy"}

# K4b cubic interpolation map: real [8386 x 838] on (re, im) of 4096 pilot vectors (1024 4x4 slots)
Wc = (torch.randn(8386, npil, device=dev) / np.sqrt(npil)).contiguous()
h = torch.randn(4096, npil, dtype=torch.complex64, device=dev)
t = timeit(lambda: eng.dense_real_apply(Wc, h), n=5)
flops = 4.0 * 8386 * npil * 4096
res["k4b_cubic_map_4096cols"] = {"ms": t * 1e3, "useful_tflops": flops / t / 1e12, "issued_tf32_tflops": 3 * flops / t / 1e12}
Wcp = eng.prepare_dense(Wc)
t = timeit(lambda: eng.dense_real_apply(Wcp, h), n=5)
res["k4b_cubic_map_prepared_4096cols"] = {"ms": t * 1e3, "useful_tflops": flops / t / 1e12, "issued_tf32_tflops": 3 * flops / t / 1e12}
del Wc, Wcp, h

# K3 stand-alone LS + MMSE + stats on resident rx / H_true, padded rows (wide accesses) and contiguous rows
B = 2048
for tag, pitch in (("k3_ls_interp_mmse_stats", 600), ("k3_ls_interp_mmse_stats_contiguous", None)):
    out = eng.run(B, 2, 200.0, 10.0, 0, pool, slot0=0, seed=1, pitch=pitch)
    rx, Ht = out["rx"], out["H_true"]
    tx_p = out["tx"][:, :, 0].reshape(B, -1)[:, torch.from_numpy(pool.pilot_indices[0]).to(dev)].contiguous()
    t = timeit(lambda: eng.ls_interp(rx, tx_p, pool, snr_db=10.0, mmse=True, H_true=Ht, want=("H_ls", "H_mmse", "stats")), n=5)
    bytes_k3 = B * (rx[0].numel() * 8 + 3 * Ht[0].numel() * 8)      # read rx + H_true, write H_ls + H_mmse (599 per row)
    res[tag] = {"ms": t * 1e3, "gbs": bytes_k3 / t / 1e9, "frac_hbm": bytes_k3 / t / 1e9 / PEAK_HBM, "slots": B,
                "note": "includes the output allocation of engine.ls_interp; " + ("pitch 600" if pitch else "contiguous")}
    del out, rx, Ht
    torch.cuda.empty_cache()

# K2 OFDM modulate / demodulate
e1 = SlotEngine(cfg(1, 1))
rows = 14 * 4 * 16384
x = torch.randn(rows, 599, dtype=torch.complex64, device=dev)
t = timeit(lambda: e1.ofdm_modulate(x), n=5)
bm = rows * (599 + 1096) * 8
res["k2_ofdm_modulate"] = {"ms": t * 1e3, "gbs": bm / t / 1e9, "frac_hbm": bm / t / 1e9 / PEAK_HBM, "rows": rows}
y = e1.ofdm_modulate(x)
t = timeit(lambda: e1.ofdm_demodulate(y), n=5)
res["k2_ofdm_demodulate"] = {"ms": t * 1e3, "gbs": bm / t / 1e9, "frac_hbm": bm / t / 1e9 / PEAK_HBM, "rows": rows}
del x, y
torch.cuda.empty_cache()

# stand-alone TDL (ChannelModel.generate_time_varying_channel): 15344 samples of a 4x4 ETU realisation, L = 78 taps
t = timeit(lambda: eng.tdl_full("ETU", 200.0, 15344, 4, 4), n=5)
res["tdl_full_4x4_etu_15344_samples"] = {"ms": t * 1e3, "out_MB": 15344 * 16 * 78 * 8 / 1e6,
                                           "note": "the reference spends ~1.5 s here per 4x4 ETU slot (SURVEY 8a3)"}

# dataset mode: H_true + rx + tx + H_ls (what generate_sample returns), padded rows
B = 4096
o = eng.alloc_outputs(B, ("H_true", "rx", "tx", "H_ls"), pitch=600)
ws = eng.workspace(B)
t = timeit(lambda: eng.run(B, 2, 200.0, 10.0, 0, pool, want=("H_true", "rx", "tx", "H_ls"), out=o, ws=ws), n=5)
bs = B * sum(o[k][0].numel() for k in ("H_true", "rx", "tx", "H_ls")) * 8
res["k1_dataset_mode"] = {"ms": t * 1e3, "gbs": bs / t / 1e9, "frac_hbm": bs / t / 1e9 / PEAK_HBM, "slots_per_s": B / t,
                          "note": "simulate + LS, no H_mmse / statistics written; includes K1a; pitch 600"}
del o

# K1a tap gains alone, and simulate-only (no estimation outputs)
B = 4096
ws = eng.workspace(B)
for tag, pitch in (("k1_simulate_only", 600), ("k1_simulate_only_contiguous", None)):
    o = eng.alloc_outputs(B, ("H_true", "rx", "tx"), pitch=pitch)
    t = timeit(lambda: eng.run(B, 2, 200.0, 10.0, want=("H_true", "rx", "tx"), out=o, ws=ws), n=5)
    bs = B * (o["H_true"][0].numel() + o["rx"][0].numel() + o["tx"][0].numel()) * 8     # 599 per row: algorithmic bytes
    res[tag] = {"ms": t * 1e3, "gbs": bs / t / 1e9, "frac_hbm": bs / t / 1e9 / PEAK_HBM, "slots_per_s": B / t,
                "note": "includes K1a tap gains; " + ("rows at pitch 600, 16-byte stores" if pitch else "contiguous rows, 8-byte stores")}
    del o
# unique-bytes (compact) forms of the same two modes: the tx grid written once, H_ls [B,14,nrx,600]
for tag, want in (("k1_dataset_mode_compact", ("H_true", "rx", "tx", "H_ls")), ("k1_simulate_only_compact", ("H_true", "rx", "tx"))):
    o = eng.alloc_outputs(B, want, compact=True, pitch=600)
    kw = dict(pattern_id=0, pool=pool) if "H_ls" in want else {}
    t = timeit(lambda: eng.run(B, 2, 200.0, 10.0, want=want, out=o, ws=ws, compact=True, **kw), n=5)
    bs = B * sum(o[k][0].numel() for k in want) * 8
    res[tag] = {"ms": t * 1e3, "gbs": bs / t / 1e9, "frac_hbm": bs / t / 1e9 / PEAK_HBM, "slots_per_s": B / t,
                "note": "includes K1a; compact layout at pitch 600 (every unique value once)"}
    del o
torch.cuda.empty_cache()

# time-domain statement of the slot: modulate -> circular TDL convolution -> demodulate (K2 exercised in situ)
B = 1024
fq = eng.run(B, 2, 200.0, 300.0, want=("H_true", "rx", "tx"))
tx = fq["tx"].contiguous()
t = timeit(lambda: eng.time_domain_slot(B, 2, 200.0, tx=tx), n=5)
rows_tx, rows_rx = B * 14 * 4, B * 14 * 4
bt = 8 * (rows_tx * (599 + 1096) + rows_tx * 1096 + rows_rx * 1096 + rows_rx * (1096 + 599))     # modulate + convolution + demodulate
res["time_domain_slot_4x4_etu"] = {"ms": t * 1e3, "slots_per_s": B / t, "gbs": bt / t / 1e9, "frac_hbm": bt / t / 1e9 / PEAK_HBM,
                                    "note": "K1a + K2 modulate + b2c_tdl_circular + K2 demodulate (incl. their output allocations); bytes = "
                                            "reads + writes of the three kernels"}
x_time = eng.ofdm_modulate(tx.reshape(-1, 599))
ws2 = eng.workspace(B)
from _b2c import check, dptr, lib, ref, stream_ptr  # noqa: E402
y_time = torch.empty((rows_rx, 1096), dtype=torch.complex64, device=dev)
mid = torch.full((B,), 2, dtype=torch.int32, device=dev)
t = timeit(lambda: check(lib().b2c_tdl_circular(ref(eng.geom), ref(eng.prof), mid.data_ptr(), B, dptr(ws2["gains"], "c64"), dptr(x_time, "c64"),
                                                 dptr(y_time, "c64"), stream_ptr())), n=5)
bc = 8 * 1096 * (rows_tx + rows_rx)
res["tdl_circular_4x4_etu"] = {"ms": t * 1e3, "gbs": bc / t / 1e9, "frac_hbm": bc / t / 1e9 / PEAK_HBM,
                               "note": "reads x_time once per rx CTA from L2 (4 rx share it), writes y_time; 4x4x9 complex MACs per sample"}
print(json.dumps(res))
