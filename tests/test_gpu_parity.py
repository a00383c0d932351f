"""GPU parity tests: the CUDA path (called through the C ABI via engine.SlotEngine / the
reference-shaped shims) against the CPU oracle and the golden vectors recorded from the reference.

Tolerances (north_star): bit-exact for integer / index data; float results within relative error
1e-4 of the float64 reference, measured norm-wise as max|a-b| / max|b| over each array (RTOL);
per-slot MSE/NMSE within 0.01 dB.  The observed errors are ~1e-6 (fp32 arithmetic).
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import OFDM_CFG, ROOT, SLOT_CASES, assert_close_elementwise, full_config, golden_draws, load_golden, relerr
from oracle import chanest_oracle as orc
from oracle import philox as opx

pytestmark = pytest.mark.gpu

RTOL = 1e-4
DB_TOL = 0.01
DIAG = os.path.join(ROOT, "gpurun_out", "parity_diag.jsonl")


def diag(**kw):
    os.makedirs(os.path.dirname(DIAG), exist_ok=True)
    with open(DIAG, "a") as fh:
        fh.write(json.dumps({k: (float(v) if isinstance(v, (np.floating, float)) else v) for k, v in kw.items()}) + "\n")


@pytest.fixture(scope="module")
def engines():
    from engine import SlotEngine
    cache = {}

    def get(ntx, nrx):
        if (ntx, nrx) not in cache:
            cache[(ntx, nrx)] = SlotEngine(full_config(ntx, nrx))
        return cache[(ntx, nrx)]
    return get


def inject_from_golden(eng, g):
    """Recorded reference draws -> the injected-draw tensors of include/b2c.h."""
    nsym, nsc = 14, 599
    mask = g["pilot_mask"]
    turns = np.empty((nsym, nsc))
    turns[mask] = g["pilot_phase"] / (2 * np.pi)
    turns[~mask] = g["data_phase"] / (2 * np.pi)
    ju = np.zeros((eng.p_max,) + g["jakes_u"].shape[1:])
    ju[:g["jakes_u"].shape[0]] = g["jakes_u"]
    dev = eng.device
    return {"jakes_u": torch.from_numpy(ju[None]).to(dev, torch.float32),
            "sym_turns": torch.from_numpy(turns[None]).to(dev, torch.float32),
            "noise": torch.from_numpy((g["noise_re"] + 1j * g["noise_im"])[None]).to(dev, torch.complex64)}


def db(x):
    return 10 * np.log10(x + 1e-12)


@pytest.mark.parametrize("name", SLOT_CASES)
def test_fused_pipeline_on_reference_draws(name, engines):
    """simulate + LS + default MMSE + statistics, injected with the reference's own draws."""
    g = load_golden(name)
    ntx, nrx = int(g["ntx"]), int(g["nrx"])
    eng = engines(ntx, nrx)
    pool = eng.pool([g["pilot_indices"]])
    out = eng.run(1, eng.models.index(str(g["model"])), float(g["doppler_hz"]), float(g["snr_db"]), 0, pool,
                  inject=inject_from_golden(eng, g))
    torch.cuda.synchronize()
    H = out["H_true"][0].cpu().numpy()
    rx = out["rx"][0].cpu().numpy()
    tx = out["tx"][0].cpu().numpy()
    H_ls = out["H_ls"][0].cpu().numpy()
    H_mm = out["H_mmse"][0].cpu().numpy()
    errs = {"H": relerr(H, g["channel"]), "rx": relerr(rx, g["rx_symbols"]),
            "tx": max(relerr(tx[:, t], g["tx_grid"]) for t in range(ntx)),
            "H_ls": max(relerr(H_ls[:, :, t], g["H_ls_tx0"]) for t in range(ntx)),
            "H_mmse": max(relerr(H_mm[:, :, t], g["H_mmse_tx0"]) for t in range(ntx))}
    diag(test="fused_injected", case=name, **errs)
    for k, v in errs.items():
        assert v < RTOL, (k, v)
    # element-wise as well: |a - b| <= 1e-4 |b| + 1e-6 max|b| on every resource element (deep fades, hull edges)
    assert_close_elementwise(H, g["channel"], what="H_true")
    assert_close_elementwise(rx, g["rx_symbols"], what="rx")
    for t in range(ntx):
        assert_close_elementwise(tx[:, t], g["tx_grid"], what="tx")
        assert_close_elementwise(H_ls[:, :, t], g["H_ls_tx0"], what="H_ls")
        assert_close_elementwise(H_mm[:, :, t], g["H_mmse_tx0"], what="H_mmse")
    # exact zeros outside the pilots' convex hull (griddata fill_value = 0.0)
    assert np.array_equal(H_ls[:, :, 0] == 0, g["H_ls_tx0"] == 0)
    # statistics vs evaluate_estimator on the reference's arrays
    st = out["stats"][0].cpu().numpy()[:, 1].sum(axis=0)      # [nrx, {pair (rx,0), all tx}, 3]
    n = H.size
    mse_ls, mse_mm, pw = st[0] / n, st[1] / n, st[2] / n
    d_ls = abs(db(mse_ls / (pw + 1e-12)) - g["metrics_ls"][2])
    d_mm = abs(db(mse_mm / (pw + 1e-12)) - g["metrics_mmse"][2])
    diag(test="fused_injected_stats", case=name, d_ls_db=d_ls, d_mm_db=d_mm)
    assert d_ls < DB_TOL and d_mm < DB_TOL
    assert abs(db(mse_ls) - db(g["metrics_ls"][0])) < DB_TOL and abs(db(mse_mm) - db(g["metrics_mmse"][0])) < DB_TOL


def test_simulate_only_and_h_only_variants(engines):
    """Optional outputs: simulate_transmission alone, and the CFR alone, equal the fused run."""
    g = load_golden("slot_2x2_eva")
    eng = engines(2, 2)
    inj = inject_from_golden(eng, g)
    args = (1, eng.models.index("EVA"), float(g["doppler_hz"]), float(g["snr_db"]))
    full = eng.run(*args, 0, eng.pool([g["pilot_indices"]]), inject=inj)
    sim = eng.run(*args, inject=inj, want=("H_true", "rx", "tx"))
    h_only = eng.run(*args, inject={"jakes_u": inj["jakes_u"]}, want=("H_true",))
    torch.cuda.synchronize()
    for k in ("H_true", "rx", "tx"):
        assert torch.equal(full[k], sim[k])
    assert torch.equal(full["H_true"], h_only["H_true"])


def test_fast_and_generic_instantiations_agree(engines):
    """The throughput (FAST) kernels and the generic ones (partial outputs) are the same arithmetic:
    Philox-mode results of one slot set must agree whichever instantiation produced them."""
    eng = engines(4, 4)
    pool = eng.random_pool([0.10], seed=2)
    args = dict(model_id=2, doppler_hz=200.0, snr_db=7.0, slot0=11, seed=3)
    full = eng.run(5, pattern_id=0, pool=pool, **args)                                    # FAST, estimation
    sim = eng.run(5, want=("H_true", "rx", "tx"), **args)                                 # FAST, simulate only
    part = eng.run(5, pattern_id=0, pool=pool, want=("H_true", "rx", "H_ls"), **args)     # generic, estimation
    h_rx = eng.run(5, want=("H_true", "rx"), **args)                                      # generic, simulate only
    torch.cuda.synchronize()
    for other in (sim, part, h_rx):
        for k in ("H_true", "rx", "tx", "H_ls"):
            if k in other:
                assert (other[k] - full[k]).abs().max().item() < 2e-6, k


@pytest.mark.parametrize("ntx,nrx,model,fd,snr,dens", [(2, 2, "EVA", 50.0, 15.0, 0.10), (4, 4, "ETU", 200.0, 5.0, 0.10),
                                                       (1, 1, "EPA", 10.0, 30.0, 0.05), (4, 2, "EPA", 100.0, -5.0, 0.02)])
def test_philox_mode_matches_oracle_on_twin_draws(ntx, nrx, model, fd, snr, dens, engines):
    """Throughput (Philox) mode: the oracle fed with oracle/philox.py's bit-exact twin of the device
    draws must reproduce the GPU arrays."""
    eng = engines(ntx, nrx)
    pool = eng.random_pool([dens], seed=7)
    seed, slot0, B = 1234, 2 ** 33 + 5, 3            # slot index beyond 32 bits exercises the counter words
    out = eng.run(B, eng.models.index(model), fd, snr, 0, pool, slot0=slot0, seed=seed)
    torch.cuda.synchronize()
    mask = pool.mask(0)
    pos = np.unravel_index(pool.pilot_indices[0], (14, 599))
    P = len(orc.TDL_NS[model])
    worst = {}
    for i in range(B):
        d = opx.slot_draws(seed, slot0 + i, 14, 599, P, ntx, nrx, mask)
        d["perm"] = np.concatenate([pool.pilot_indices[0], np.setdiff1d(np.arange(14 * 599), pool.pilot_indices[0])])
        ref = orc.slot_pipeline(OFDM_CFG, ntx, nrx, model, fd, snr, dens, d)
        assert np.array_equal(ref["pilot_mask"], mask)
        got = {k: out[k][i].cpu().numpy() for k in ("H_true", "rx", "tx", "H_ls", "H_mmse")}
        e = {"H": relerr(got["H_true"], ref["channel"]), "rx": relerr(got["rx"], ref["rx_symbols"]),
             "tx": relerr(got["tx"], ref["tx_symbols"]), "H_ls": relerr(got["H_ls"], ref["H_ls"]),
             "H_mmse": relerr(got["H_mmse"], ref["H_mmse"])}
        for k, v in e.items():
            worst[k] = max(worst.get(k, 0), v)
        # element-wise too.  The floor is 4e-6 max|b| here (1e-6 on injected draws): the Philox path evaluates
        # Box-Muller / symbol phases on the SFU (lg2.approx, sin/cos.approx: ~4e-7 absolute per unit amplitude).
        for k, rk in (("H_true", "channel"), ("rx", "rx_symbols"), ("tx", "tx_symbols"), ("H_ls", "H_ls"), ("H_mmse", "H_mmse")):
            assert_close_elementwise(got[k], ref[rk], floor=4e-6, what=k)
        st = out["stats"][i].cpu().numpy()[:, 1].sum(axis=0) / ref["channel"].size
        m_ls, m_mm = orc.evaluate(ref["channel"], ref["H_ls"]), orc.evaluate(ref["channel"], ref["H_mmse"])
        assert abs(db(st[0] / (st[2] + 1e-12)) - m_ls["nmse_db"]) < DB_TOL
        assert abs(db(st[1] / (st[2] + 1e-12)) - m_mm["nmse_db"]) < DB_TOL
    diag(test="philox_twin", case=f"{ntx}x{nrx}_{model}", **worst)
    for k, v in worst.items():
        assert v < RTOL, (k, v)


def test_results_do_not_depend_on_batching(engines):
    """Counter-based draws: slot g is the same array whether it is generated alone or inside a batch,
    which is what makes sharding over GPUs and resume exact."""
    eng = engines(2, 2)
    pool = eng.random_pool([0.05, 0.10], seed=3)
    B = 8
    model = np.array([0, 1, 2, 0, 1, 2, 0, 1], dtype=np.int32)
    fd = np.array([10, 50, 100, 200, 10, 50, 100, 200], dtype=np.float32)
    snr = np.array([-5, 0, 5, 10, 15, 20, 25, 30], dtype=np.float32)
    pid = np.array([0, 1, 0, 1, 1, 0, 1, 0], dtype=np.int32)
    big = eng.run(B, model, fd, snr, pid, pool, slot0=100, seed=9)
    for i in (0, 3, 7):
        one = eng.run(1, model[i:i + 1], fd[i:i + 1], snr[i:i + 1], pid[i:i + 1], pool, slot0=100 + i, seed=9)
        torch.cuda.synchronize()
        for k in ("H_true", "rx", "tx", "H_ls", "H_mmse", "stats"):
            assert torch.equal(big[k][i], one[k][0]), (k, i)
    other = eng.run(1, model[:1], fd[:1], snr[:1], pid[:1], pool, slot0=100, seed=10)
    assert not torch.equal(other["H_true"][0], big["H_true"][0])


def test_mixed_batch_against_oracle(engines):
    """Mixed EPA/EVA/ETU, Doppler, SNR and density in one launch (dataset-generation shape)."""
    eng = engines(2, 2)
    dens = [0.05, 0.10]
    pool = eng.random_pool(dens, seed=11)
    models = ["ETU", "EPA", "EVA", "EPA"]
    fd = [200.0, 10.0, 50.0, 100.0]
    snr = [0.0, 30.0, 10.0, -5.0]
    pid = [1, 0, 1, 0]
    out = eng.run(4, [eng.models.index(m) for m in models], fd, snr, pid, pool, slot0=40, seed=5)
    torch.cuda.synchronize()
    for i in range(4):
        mask = pool.mask(pid[i])
        d = opx.slot_draws(5, 40 + i, 14, 599, len(orc.TDL_NS[models[i]]), 2, 2, mask)
        d["perm"] = np.concatenate([pool.pilot_indices[pid[i]], np.setdiff1d(np.arange(8386), pool.pilot_indices[pid[i]])])
        ref = orc.slot_pipeline(OFDM_CFG, 2, 2, models[i], fd[i], snr[i], dens[pid[i]], d)
        for k, rk in (("H_true", "channel"), ("rx", "rx_symbols"), ("H_ls", "H_ls"), ("H_mmse", "H_mmse")):
            assert relerr(out[k][i].cpu().numpy(), ref[rk]) < RTOL, (i, k)


@pytest.mark.parametrize("name", ["slot_2x2_eva", "slot_4x4_etu", "slot_2x1_epa_1pct"])
def test_standalone_ls_kernel(name, engines):
    """b2c_ls_interp on the reference's rx grids: LS, default MMSE, pilot estimates, statistics."""
    g = load_golden(name)
    ntx, nrx = int(g["ntx"]), int(g["nrx"])
    eng = engines(ntx, nrx)
    pool = eng.pool([g["pilot_indices"]])
    dev = eng.device
    rx = torch.from_numpy(g["rx_symbols"][None]).to(dev, torch.complex64)
    xp = torch.from_numpy(g["pilot_symbols"][None]).to(dev, torch.complex64)
    Ht = torch.from_numpy(g["channel"][None]).to(dev, torch.complex64)
    out = eng.ls_interp(rx, xp, pool, snr_db=float(g["snr_db"]), mmse=True, H_true=Ht,
                        want=("H_ls", "H_mmse", "hp", "stats"))
    torch.cuda.synchronize()
    H_ls, H_mm = out["H_ls"][0].cpu().numpy(), out["H_mmse"][0].cpu().numpy()
    for t in range(ntx):
        assert relerr(H_ls[:, :, t], g["H_ls_tx0"]) < RTOL
        assert relerr(H_mm[:, :, t], g["H_mmse_tx0"]) < RTOL
    for r in range(nrx):
        hp = orc.ls_at_pilots(g["rx_symbols"][:, r], g["pilot_symbols"], g["pilot_mask"])
        assert relerr(out["hp"][0, r, :len(hp)].cpu().numpy(), hp) < RTOL
    st = out["stats"][0].cpu().numpy()[:, 1].sum(axis=0) / g["channel"].size
    assert abs(db(st[0] / (st[2] + 1e-12)) - g["metrics_ls"][2]) < DB_TOL
    assert abs(db(st[1] / (st[2] + 1e-12)) - g["metrics_mmse"][2]) < DB_TOL


def test_stats_bins_against_numpy(engines):
    eng = engines(2, 2)
    pool = eng.random_pool([0.10], seed=1)
    B = 64
    snr_idx = np.arange(B) % 8
    snr = np.array([-5, 0, 5, 10, 15, 20, 25, 30], dtype=np.float32)[snr_idx]
    out = eng.run(B, 1, 50.0, snr, 0, pool, slot0=0, seed=77)
    bin_id = snr_idx.astype(np.int32).copy()
    bin_id[5] = -1                                        # skipped slot
    bins = eng.stats_bins(out["stats"], bin_id, 8, snr_db=snr).cpu().numpy()
    H, Hl, Hm = (out[k].cpu().numpy().astype(np.complex128) for k in ("H_true", "H_ls", "H_mmse"))
    want = np.zeros((8, 14))
    for b in range(B):
        if bin_id[b] < 0:
            continue
        ls, mm = orc.evaluate(H[b], Hl[b]), orc.evaluate(H[b], Hm[b])
        n00l, n00m = orc.nmse_pair00(Hl[b], H[b]), orc.nmse_pair00(Hm[b], H[b])
        # the BER proxy of run_phase5_evaluation.py:57-68 on the pair-(0,0) rows (oracle restatement: ber_approximation)
        bl = orc.ber_approximation(Hl[b][:, 0, 0], H[b][:, 0, 0], float(snr[b]))
        bm = orc.ber_approximation(Hm[b][:, 0, 0], H[b][:, 0, 0], float(snr[b]))
        want[bin_id[b]] += [1, ls["mse"], mm["mse"], ls["nmse"], mm["nmse"], ls["nmse"] ** 2, mm["nmse"] ** 2,
                            np.mean(np.abs(H[b]) ** 2), n00l, n00l ** 2, n00m, n00m ** 2, bl, bm]
    assert np.array_equal(bins[:, 0], want[:, 0])
    assert np.allclose(bins, want, rtol=2e-5), np.abs(bins / want - 1).max()
    no_snr = eng.stats_bins(out["stats"], bin_id, 8).cpu().numpy()          # without SNRs the proxy fields stay zero
    assert np.array_equal(no_snr[:, :12], bins[:, :12]) and not no_snr[:, 12:].any()
    # accumulate-into semantics and determinism
    again = eng.stats_bins(out["stats"], bin_id, 8, torch.from_numpy(bins).to(eng.device), snr_db=snr).cpu().numpy()
    assert np.array_equal(again, 2 * bins)


def test_ofdm_modem(engines):
    g = load_golden("ofdm_modem")
    eng = engines(1, 1)
    dev = eng.device
    mod = eng.ofdm_modulate(torch.from_numpy(g["symbols"]).to(dev, torch.complex64))
    dem = eng.ofdm_demodulate(torch.from_numpy(g["signal"]).to(dev, torch.complex64))
    rt = eng.ofdm_demodulate(mod)
    torch.cuda.synchronize()
    e = {"mod": relerr(mod.cpu().numpy(), g["modulated"]), "demod": relerr(dem.cpu().numpy(), g["demodulated"]),
         "roundtrip": relerr(rt.cpu().numpy(), g["symbols"])}
    diag(test="ofdm", **e)
    assert max(e.values()) < RTOL
    # many rows (grid-stride path) + linearity
    x = torch.randn(5000, 599, dtype=torch.complex64, device=dev)
    y = eng.ofdm_modulate(x)
    assert relerr(eng.ofdm_demodulate(y).cpu().numpy(), x.cpu().numpy()) < RTOL
    assert relerr((eng.ofdm_modulate(2 * x[:64]) - 2 * y[:64]).abs().cpu().numpy() + 1, np.ones((64, 1096))) < 1e-5
    assert torch.equal(y[:, :72], y[:, 1024:])            # cyclic prefix
    # another prefix length (odd: rows 8-byte aligned only); same transform
    from engine import SlotEngine
    cfg = full_config(1, 1)
    cfg["ofdm"]["cp_length"] = 71
    e71 = SlotEngine(cfg)
    y71 = e71.ofdm_modulate(x[:300])
    assert y71.shape == (300, 1095) and torch.equal(y71[:, 71:], y[:300, 72:]) and torch.equal(y71[:, :71], y71[:, 1024:])
    assert relerr(e71.ofdm_demodulate(y71).cpu().numpy(), x[:300].cpu().numpy()) < RTOL


@pytest.mark.parametrize("model", ["EPA", "EVA", "ETU"])
def test_tdl_standalone(model, engines):
    g = load_golden("tdl_standalone")
    fd, ntx, nrx, ns = (g[f"{model}_meta"][0], *(int(v) for v in g[f"{model}_meta"][1:]))
    eng = engines(ntx, nrx)
    ju = torch.from_numpy(g[f"{model}_jakes_u"]).to(eng.device, torch.float32)
    h = eng.tdl_full(model, float(fd), ns, ntx, nrx, jakes_u=ju)
    torch.cuda.synchronize()
    assert tuple(h.shape) == tuple(g[f"{model}_shape"])
    got = h.cpu().numpy()
    assert relerr(got[::97], g[f"{model}_h"]) < RTOL
    # untouched delays are exactly zero (the reference allocates zeros, :97)
    delays = set(int(d) for d in g[f"{model}_delay_samples"])
    for d in range(got.shape[-1]):
        if d not in delays:
            assert not got[..., d].any()


def test_apply_channel_generic(engines):
    eng = engines(2, 2)
    rng = np.random.default_rng(5)
    B = 3
    tx = rng.standard_normal((B, 14, 2, 599)) + 1j * rng.standard_normal((B, 14, 2, 599))
    H = rng.standard_normal((B, 14, 2, 2, 599)) + 1j * rng.standard_normal((B, 14, 2, 2, 599))
    nz = rng.standard_normal((B, 14, 2, 599)) + 1j * rng.standard_normal((B, 14, 2, 599))
    snr = np.array([0.0, 10.0, 25.0], dtype=np.float32)
    dev = eng.device
    rx = eng.apply_channel(torch.from_numpy(tx).to(dev, torch.complex64), torch.from_numpy(H).to(dev, torch.complex64),
                           snr, noise=torch.from_numpy(nz).to(dev, torch.complex64))
    torch.cuda.synchronize()
    for b in range(B):
        assert relerr(rx[b].cpu().numpy(), orc.apply_channel(tx[b], H[b], float(snr[b]), nz[b].real, nz[b].imag)) < RTOL
    # Philox noise: empirical SNR close to the request
    rx2 = eng.apply_channel(torch.from_numpy(tx).to(dev, torch.complex64), torch.from_numpy(H).to(dev, torch.complex64),
                            snr, seed=3, slot0=0).cpu().numpy()
    for b in range(B):
        y = np.einsum("srtk,stk->srk", H[b], tx[b])
        est = 10 * np.log10(np.mean(np.abs(y) ** 2) / np.mean(np.abs(rx2[b] - y) ** 2))
        assert abs(est - snr[b]) < 0.2


def test_dense_wiener_path(engines):
    """Known-covariance MMSE: GPU LS at pilots -> dense W GEMM -> plan interpolation vs the reference."""
    g = load_golden("mmse_dense_2x2")
    eng = engines(2, 2)
    pos = np.unravel_index(g["pilot_indices"], (14, 599))
    ds = pos[0][:, None] - pos[0][None, :]
    dk = pos[1][:, None] - pos[1][None, :]
    R = 0.4 * np.exp(-np.abs(ds) / 20.0 - np.abs(dk) / 60.0) * np.exp(1j * 2 * np.pi * dk * 3 / 1024)
    W = orc.wiener_matrix(R, float(g["snr_db"]))
    dev = eng.device
    pool = eng.pool([g["pilot_indices"]])
    rx = torch.from_numpy(g["rx_symbols"][None]).to(dev, torch.complex64)
    xp = torch.from_numpy(g["pilot_symbols"][None]).to(dev, torch.complex64)
    hp = eng.ls_interp(rx, xp, pool, want=("hp",))["hp"]
    hm = eng.mmse_dense(torch.from_numpy(W).to(dev, torch.complex64), hp.reshape(2, -1)).reshape(1, 2, -1)
    H = eng.ls_interp(None, None, pool, hp_in=hm, want=("H_ls",))["H_ls"][0].cpu().numpy()
    for t in range(2):
        assert relerr(H[:, :, t], g["H_mmse_tx0"]) < RTOL
    # GEMM alone at an awkward size against numpy
    rng = np.random.default_rng(2)
    n, c = 167, 77
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    X = rng.standard_normal((c, n)) + 1j * rng.standard_normal((c, n))
    Y = eng.mmse_dense(torch.from_numpy(A).to(dev, torch.complex64), torch.from_numpy(X).to(dev, torch.complex64))
    assert relerr(Y.cpu().numpy(), X @ A.T) < RTOL
    # full-size Wiener matrix (838 pilots), ragged column count, padded leading dimension: the 3xTF32
    # tensor-core path measures ~1.3e-5 (tensor-core accumulation), plain TF32 would sit near 5e-4
    n, c, ld = 838, 300, 840
    A = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(n)
    X = np.zeros((c, ld), complex)
    X[:, :n] = rng.standard_normal((c, n)) + 1j * rng.standard_normal((c, n))
    At, Xt = torch.from_numpy(A).to(dev, torch.complex64), torch.from_numpy(X).to(dev, torch.complex64)
    Y = eng.mmse_dense(At, Xt).cpu().numpy()
    assert relerr(Y[:, :n], X[:, :n] @ A.T) < 5e-5 and not Y[:, n:].any()
    # prepared operand (pre-split tiles fetched by bulk copy): same MMAs on the same bits
    prep = eng.prepare_dense(At)
    # (>= 128 columns with 16-byte aligned rows take the TS kernel -- data operand in tensor memory --, the one-shot call
    # the SS kernel: same three TF32 products per K step, accumulated by different instruction shapes)
    for _ in range(2):                      # second call: the prepared buffer is reusable
        Yp = eng.mmse_dense(prep, Xt).cpu().numpy()
        assert relerr(Yp, Y) < 2e-6 and not Yp[:, n:].any()
    os.environ["B2C_DENSE_SS"] = "1"        # force the shared-memory-operand kernels: bit-identical to the one-shot form
    try:
        assert torch.equal(eng.mmse_dense(prep, Xt), torch.from_numpy(Y).to(dev))
    finally:
        del os.environ["B2C_DENSE_SS"]
    # 4000 columns: the 128 x 256 tile form is chosen (fuller last wave), ragged last tile, padded leading dimension
    Xw = np.zeros((4000, ld), complex)
    Xw[:, :n] = rng.standard_normal((4000, n)) + 1j * rng.standard_normal((4000, n))
    Xwt = torch.from_numpy(Xw).to(dev, torch.complex64)
    Yw = eng.mmse_dense(prep, Xwt)
    assert relerr(Yw.cpu().numpy()[:, :n], Xw[:, :n] @ A.T) < 5e-5 and not Yw[:, n:].abs().max().item()
    assert relerr(Yw.cpu().numpy(), eng.mmse_dense(At, Xwt).cpu().numpy()) < 2e-6      # TS kernel vs one-shot SS kernel
    # ragged everything on the TS kernel: odd pilot count (K tail inside a 16-byte load), 131 columns, padded rows
    n2, c2, ld2 = 419, 131, 420
    A2 = (rng.standard_normal((n2, n2)) + 1j * rng.standard_normal((n2, n2))) / np.sqrt(n2)
    X2 = np.full((c2, ld2), np.nan + 0j)                       # the padding element of each row must never be read into a result
    X2[:, :n2] = rng.standard_normal((c2, n2)) + 1j * rng.standard_normal((c2, n2))
    Y2 = eng.mmse_dense(eng.prepare_dense(torch.from_numpy(A2).to(dev, torch.complex64)), torch.from_numpy(X2).to(dev, torch.complex64)).cpu().numpy()
    assert relerr(Y2[:, :n2], X2[:, :n2] @ A2.T) < 5e-5 and not Y2[:, n2:].any()
    small = torch.from_numpy(rng.standard_normal((167, 167)) + 1j * rng.standard_normal((167, 167))).to(dev, torch.complex64)
    xs = torch.from_numpy(rng.standard_normal((77, 167)) + 1j * rng.standard_normal((77, 167))).to(dev, torch.complex64)
    assert torch.equal(eng.mmse_dense(eng.prepare_dense(small), xs), eng.mmse_dense(small, xs))


def model_cov(idx, nsc=599):
    """The formula-defined pilot covariance of the dense-MMSE fixtures (oracle/make_golden.py model_covariance)."""
    ps, pk = idx // nsc, idx % nsc
    ds, dk = ps[:, None] - ps[None, :], pk[:, None] - pk[None, :]
    return 0.4 * np.exp(-np.abs(ds) / 20.0 - np.abs(dk) / 60.0) * np.exp(1j * 2 * np.pi * dk * 3 / 1024)


def test_batched_dense_wiener_pipeline_on_reference_draws(engines):
    """SlotEngine.run(mmse="dense"): slot kernel (pilot vectors out) -> one GEMM per (pattern, SNR) -> K3 mode 2,
    driven with the reference's recorded draws of the 4x4 ETU slot, against the reference's known-covariance
    MMSEEstimator (838 x 838 Wiener matrix, tests/golden/mmse_dense_4x4_etu.npz)."""
    from engine import WienerBank
    g, d = load_golden("slot_4x4_etu"), load_golden("mmse_dense_4x4_etu")
    eng = engines(4, 4)
    pool = eng.pool([g["pilot_indices"]])
    snr = float(g["snr_db"])
    bank = WienerBank(eng, pool, {0: model_cov(g["pilot_indices"])}, [snr])
    out = eng.run(1, eng.models.index("ETU"), float(g["doppler_hz"]), snr, 0, pool, inject=inject_from_golden(eng, g),
                  mmse="dense", wiener=bank)
    torch.cuda.synchronize()
    H_mm = out["H_mmse"][0].cpu().numpy()
    errs = [relerr(H_mm[:, :, t], d["H_mmse_tx0"]) for t in range(4)]
    diag(test="dense_pipeline_injected", worst=max(errs))
    assert max(errs) < RTOL
    assert_close_elementwise(H_mm[:, :, 0], d["H_mmse_tx0"])
    assert relerr(out["H_ls"][0, :, :, 0].cpu().numpy(), g["H_ls_tx0"]) < RTOL          # LS part untouched
    # the opt-in cluster form of the GEMM (W tiles multicast between two CTAs): same products, same accumulation order
    os.environ["B2C_DENSE_CLUSTER"] = "1"
    try:
        out_cl = eng.run(1, eng.models.index("ETU"), float(g["doppler_hz"]), snr, 0, pool, inject=inject_from_golden(eng, g),
                         mmse="dense", wiener=bank)
        torch.cuda.synchronize()
    finally:
        del os.environ["B2C_DENSE_CLUSTER"]
    assert torch.equal(out_cl["H_mmse"], out["H_mmse"])
    st = out["stats"][0].cpu().numpy()[:, 1].sum(axis=0)
    n = g["channel"].size
    assert abs(db(st[1] / n / (st[2] / n + 1e-12)) - d["metrics_mmse"][2]) < DB_TOL     # dense MMSE NMSE
    assert abs(db(st[0] / n / (st[2] / n + 1e-12)) - g["metrics_ls"][2]) < DB_TOL       # LS NMSE from the slot kernel


@pytest.mark.parametrize("pitch", [600, None])
def test_batched_dense_wiener_pipeline_groups(pitch, engines):
    """A mixed batch (2 patterns of different pilot counts x 3 SNRs, shuffled) through the dense pipeline: every slot
    must get the filter of ITS (pattern, SNR) group -- checked against the oracle's W @ h_ls + interpolation on the
    GPU's own LS pilot estimates, and H_true / rx / tx / H_ls must equal the default-MMSE run bit for bit."""
    from engine import WienerBank
    eng = engines(2, 2)
    pool = eng.random_pool([0.05, 0.02], seed=9)
    snrs = [0.0, 10.0, 20.0]
    covs = {i: model_cov(pool.pilot_indices[i]) for i in range(2)}
    bank = WienerBank(eng, pool, covs, snrs)
    B = 13
    rng = np.random.default_rng(5)
    pid = rng.integers(0, 2, B).astype(np.int32)
    snr = np.asarray(snrs, np.float32)[rng.integers(0, 3, B)]
    args = dict(model_id=1, doppler_hz=50.0, snr_db=snr, pattern_id=pid, pool=pool, slot0=4242, seed=6)
    base = eng.run(B, pitch=pitch, **args)
    out = eng.run(B, pitch=pitch, mmse="dense", wiener=bank, **args)
    torch.cuda.synchronize()
    for k in ("H_true", "rx", "tx", "H_ls"):
        assert torch.equal(out[k], base[k]), k
    assert torch.equal(out["stats"][..., 0], base["stats"][..., 0]) and torch.equal(out["stats"][..., 2], base["stats"][..., 2])
    H_ls = out["H_ls"].cpu().numpy()
    H_true = out["H_true"].cpu().numpy().astype(np.complex128)
    H_mm = out["H_mmse"].cpu().numpy()
    stats = out["stats"].cpu().numpy()
    plans = {}
    for b in range(B):
        idx = pool.pilot_indices[pid[b]]
        pos = (idx // 599, idx % 599)
        if pid[b] not in plans:
            plans[pid[b]] = orc.linear_plan(pos, 14, 599)
        W = orc.wiener_matrix(covs[pid[b]], float(snr[b]))
        for r in range(2):
            h_ls_p = H_ls[b, :, r, 0].reshape(-1)[idx].astype(np.complex128)      # LS interpolation is exact at the pilots
            ref = orc.plan_apply(*plans[pid[b]], W @ h_ls_p, 14, 599)
            for t in range(2):
                assert relerr(H_mm[b, :, r, t], ref) < RTOL, (b, r, t)
            e = sum((np.abs(H_true[b, :, r, t] - ref) ** 2).sum() for t in range(2))
            assert abs(stats[b, r, 1, 1] / e - 1) < 1e-3, (b, r)
    # a prebuilt plan (grouping done once on the host) with device-resident parameters: same launches, same bits
    plan = bank.plan_batch(eng, pid, snr, B)
    dev_args = dict(args, snr_db=torch.from_numpy(snr).to(eng.device), pattern_id=torch.from_numpy(pid).to(eng.device))
    again = eng.run(B, pitch=pitch, mmse="dense", wiener=bank, dense_plan=plan, **dev_args)
    torch.cuda.synchronize()
    assert torch.equal(again["H_mmse"], out["H_mmse"]) and torch.equal(again["stats"], out["stats"])
    with pytest.raises(ValueError):            # device parameters without a plan: the grouping needs host values
        eng.run(B, pitch=pitch, mmse="dense", wiener=bank, **dev_args)
    # the grouped launch equals one prepared product per group, bit for bit
    hp = out["_keepalive"][2]["hp"]
    hm_grouped = out["_keepalive"][2]["hm"].clone()
    for gq in plan.groups:
        key = [k for k in plan.keys if bank.prepared[k].buf.data_ptr() == gq.prepared][0]
        rows = slice(gq.col0, gq.col0 + gq.ncols)
        one = eng.mmse_dense(bank.prepared[key], hp[rows].contiguous())
        n = bank.prepared[key].m
        assert (one[:, :n] - hm_grouped[rows, :n]).abs().max().item() <= 2e-6 * one[:, :n].abs().max().item(), key


@pytest.mark.parametrize("ntx,nrx,model", [(1, 1, "EPA"), (2, 2, "EVA"), (4, 4, "ETU"), (4, 2, "EVA")])
def test_dense_wiener_statistics_without_arrays(ntx, nrx, model, engines):
    """run(mmse="dense", want=("stats",)): slot kernel (statistics + pilot vectors) -> grouped GEMM -> b2c_dense_score,
    no resource-grid array in HBM.  The per-slot sums must equal those of the array-writing dense pipeline (which is
    checked against the reference above): the LS / power fields come from the same arithmetic, the MMSE field from the
    same filtered pilots interpolated against the regenerated CFR."""
    from engine import WienerBank
    eng = engines(ntx, nrx)
    pool = eng.random_pool([0.05, 0.02, 0.1], seed=4)
    snrs = [-5.0, 10.0, 25.0]
    bank = WienerBank(eng, pool, {i: model_cov(pool.pilot_indices[i]) for i in range(3)}, snrs)
    B = 37
    rng = np.random.default_rng(8)
    pid = rng.integers(0, 3, B).astype(np.int32)
    snr = np.asarray(snrs, np.float32)[rng.integers(0, 3, B)]
    args = dict(model_id=eng.models.index(model), doppler_hz=80.0, snr_db=snr, pattern_id=pid, pool=pool, slot0=977, seed=3,
                mmse="dense", wiener=bank)
    full = eng.run(B, **args)
    lean = eng.run(B, want=("stats",), **args)
    torch.cuda.synchronize()
    assert set(k for k in lean if not k.startswith("_")) == {"stats"}
    a, b = full["stats"].cpu().numpy(), lean["stats"].cpu().numpy()
    worst = np.abs(b / a - 1).max(axis=(0, 1, 2))
    diag(test="dense_score", case=f"{ntx}x{nrx}_{model}", worst_ls=worst[0], worst_mmse=worst[1], worst_pow=worst[2])
    assert worst.max() < 2e-5
    # against float64 on the array run's own outputs: sum |H_mmse - H_true|^2 over all tx of one rx
    Hm, Ht = full["H_mmse"].cpu().numpy().astype(np.complex128), full["H_true"].cpu().numpy().astype(np.complex128)
    e = (np.abs(Hm - Ht) ** 2).sum(axis=(1, 3, 4))                                     # [B, nrx]
    assert np.abs(b[:, :, 1, 1] / e - 1).max() < 2e-5
    # a prebuilt plan with device-resident parameters takes the same route
    plan = bank.plan_batch(eng, pid, snr, B)
    dev = dict(args, snr_db=torch.from_numpy(snr).to(eng.device), pattern_id=torch.from_numpy(pid).to(eng.device))
    again = eng.run(B, want=("stats",), dense_plan=plan, **dev)
    torch.cuda.synchronize()
    assert torch.equal(again["stats"], lean["stats"])
    # the scoring pass with per-thread plan staging instead of the bulk-copied row ring: same bits
    os.environ["B2C_PLAN_BULK"] = "0"
    try:
        per_thread = eng.run(B, want=("stats",), dense_plan=plan, **dev)
        torch.cuda.synchronize()
    finally:
        del os.environ["B2C_PLAN_BULK"]
    assert torch.equal(per_thread["stats"], lean["stats"])


@pytest.mark.parametrize("model,fd,ntx,nrx", [("EPA", 10.0, 1, 1), ("EVA", 50.0, 2, 2), ("ETU", 200.0, 4, 4)])
def test_time_domain_path_lands_on_the_frequency_domain_grid(model, fd, ntx, nrx, engines):
    """modulate (K2) -> per-symbol circular TDL convolution with the K1a tap gains -> demodulate (K2) reproduces the
    received grid of the frequency-domain pipeline (noise off) within 1e-4 -- for ETU too, whose last tap (77 samples)
    exceeds the 72-sample prefix: the reference's per-bin CFR product IS a circular convolution per symbol
    (src/channel_simulator.py:274-345).  One symbol is also checked against a NumPy time-domain convolution."""
    eng = engines(ntx, nrx)
    B, m = 3, eng.models.index(model)
    fq = eng.run(B, m, fd, 300.0, slot0=31, seed=12, want=("H_true", "rx", "tx"))          # 300 dB: the noise term is ~1e-15
    td = eng.time_domain_slot(B, m, fd, tx=fq["tx"], slot0=31, seed=12)
    torch.cuda.synchronize()
    assert td["rx"].shape == fq["rx"].shape
    err = relerr(td["rx"].cpu().numpy(), fq["rx"].cpu().numpy())
    diag(test="time_domain", case=model, err=err)
    assert err < RTOL
    assert_close_elementwise(td["rx"].cpu().numpy(), fq["rx"].cpu().numpy(), floor=4e-6, what="rx (time domain)")
    # time-domain samples against NumPy: y[n] = sum_tx sum_t g x[(n - d) mod N], prefix = last 72 samples
    gains = td["_keepalive"][2]["gains"].cpu().numpy().astype(np.complex128)               # [B, nrx, nsym, ntx, 16]
    delays = eng.host_tables["tap_delay"][m][:int(eng.host_tables["ntaps"][m])]
    b, s = 1, 5
    xt = orc.ofdm_modulate(fq["tx"][b, s].cpu().numpy().astype(np.complex128), 1024, 72, 600)   # [ntx, 1096]
    assert relerr(td["x_time"][b, s].cpu().numpy(), xt) < RTOL
    for r in range(nrx):
        body = sum(gains[b, r, s, t, i] * np.roll(xt[t, 72:], int(d)) for t in range(ntx) for i, d in enumerate(delays))
        want = np.concatenate([body[-72:], body])
        assert relerr(td["y_time"][b, s, r].cpu().numpy(), want) < RTOL
    # default transmit grid: the slot pipeline's own Philox grid for these slots
    td2 = eng.time_domain_slot(B, m, fd, slot0=31, seed=12)
    assert torch.equal(td2["tx"], fq["tx"]) and torch.equal(td2["rx"], td["rx"])


@pytest.mark.parametrize("ntx,nrx,model,snr", [(1, 1, "EPA", 8.0), (2, 2, "EVA", 4.0), (4, 4, "ETU", 0.0)])
def test_qpsk_bit_errors_against_the_oracle(ntx, nrx, model, snr, engines):
    """BER path: QPSK grid -> slot pipeline -> equalize_channel('zf') with H_ls / H_mmse / H_true -> qam_demodulation ->
    bit errors on the data REs, per slot.  (1) Philox mode: b2c_slots.qpsk against the oracle fed with the twin draws
    (oracle/philox.py qpsk=True): same grid, same per-slot error counts (a bit whose soft value sits within fp32
    rounding of a decision boundary may flip: <= 2 per slot allowed).  (2) injected QPSK phases: same check through
    the generic kernels."""
    eng = engines(ntx, nrx)
    dens = 0.05
    pool = eng.random_pool([dens], seed=3)
    mask = pool.mask(0)
    m, fd, B, seed, slot0 = eng.models.index(model), 70.0, 3, 77, 900
    P = len(orc.TDL_NS[model])
    perm = np.concatenate([pool.pilot_indices[0], np.setdiff1d(np.arange(14 * 599), pool.pilot_indices[0])])

    def oracle_errors(d):
        ref = orc.slot_pipeline(OFDM_CFG, ntx, nrx, model, fd, snr, dens, d)
        tx0 = ref["tx_symbols"][:, 0]
        bits_tx = orc.qam_demodulate(tx0, 4).reshape(14 * 599, 2)
        out = {}
        for name, H in (("H_ls", ref["H_ls"]), ("H_mmse", ref["H_mmse"]), ("H_true", ref["channel"])):
            xh = orc.equalize(ref["rx_symbols"], H, "zf")[:, 0]
            bits = orc.qam_demodulate(xh, 4).reshape(14 * 599, 2)
            out[name] = int(((bits != bits_tx) & ~mask.reshape(-1, 1)).sum())
        return ref, out

    got = eng.ber_batch(B, m, fd, snr, 0, pool, slot0=slot0, seed=seed)
    torch.cuda.synchronize()
    assert np.array_equal(got["bits"], np.full(B, 2 * (14 * 599 - len(pool.pilot_indices[0]))))
    tx_gpu = got["_keepalive"]["tx"].cpu().numpy()
    total = 0
    for i in range(B):
        d = opx.slot_draws(seed, slot0 + i, 14, 599, P, ntx, nrx, mask, qpsk=True)
        d["perm"] = perm
        ref, want = oracle_errors(d)
        assert relerr(tx_gpu[i], ref["tx_symbols"]) < RTOL
        assert np.allclose(np.abs(np.angle(ref["tx_symbols"]) / (np.pi / 4)) % 2, 1.0, atol=1e-5)      # QPSK points
        for name in ("H_ls", "H_mmse", "H_true"):
            assert abs(int(got["errors"][name][i]) - want[name]) <= 2, (name, i, int(got["errors"][name][i]), want[name])
        total += want["H_ls"]
    assert total > 0                                                    # the case is noisy enough to exercise the counter
    # injected QPSK phases (numpy draws) through the generic instantiation
    rng = np.random.default_rng(5)
    turns = (2 * rng.integers(0, 4, (1, 14, 599)) + 1) / 8.0
    ju = np.zeros((1, eng.p_max, ntx, nrx, 2, 20))
    ju[:, :P] = rng.random((1, P, ntx, nrx, 2, 20))
    z = rng.standard_normal((2, 14, nrx, 599))
    inj = {"jakes_u": torch.from_numpy(ju).to(eng.device, torch.float32), "sym_turns": torch.from_numpy(turns).to(eng.device, torch.float32),
           "noise": torch.from_numpy((z[0] + 1j * z[1])[None]).to(eng.device, torch.complex64)}
    got2 = eng.ber_batch(1, m, fd, snr, 0, pool, inject=inj)
    ph = 2 * np.pi * turns[0]
    d = {"perm": perm, "pilot_phase": ph[mask], "data_phase": ph[~mask], "jakes_u": ju[0, :P], "noise_re": z[0], "noise_im": z[1]}
    _, want2 = oracle_errors(d)
    for name in ("H_ls", "H_mmse", "H_true"):
        assert abs(int(got2["errors"][name][0]) - want2[name]) <= 2, (name, int(got2["errors"][name][0]), want2[name])


def test_pilot_sweep_ber_curves():
    """PilotOptimizer.analyze_pilot_density(with_ber=True): the measured BER falls with SNR, the true channel is the
    floor, and the proxy curve (compute_ber_approximation) is reported beside it."""
    import run_phase8_pilot_optimization as p8
    opt = p8.PilotOptimizer(rng='philox', seed=3)
    opt.config['mimo'] = {'num_tx_antennas': 1, 'num_rx_antennas': 2}
    res = opt.analyze_pilot_density([0.02, 0.10], [0, 10, 20], num_samples=24, channel_type='EPA', doppler_hz=10.0, with_ber=True, ber_samples=12)
    for dens in (0.02, 0.10):
        ls = [res['methods']['LS'][s][dens]['ber'] for s in (0, 10, 20)]
        pf = [res['methods']['PERFECT'][s][dens]['ber'] for s in (0, 10, 20)]
        assert ls[0] > ls[1] > ls[2] >= 0 and pf[0] > pf[1] >= pf[2]
        assert all(p <= l + 1e-3 for p, l in zip(pf, ls))              # perfect CSI is the floor
        assert 0 < ls[0] < 0.5 and 0 <= res['methods']['LS'][0][dens]['ber_proxy'] <= 0.5
    assert res['methods']['LS'][10][0.10]['ber'] <= res['methods']['LS'][10][0.02]['ber'] + 5e-3     # more pilots do not hurt


@pytest.mark.parametrize("ntx,nrx,compact", [(4, 4, False), (4, 4, True), (2, 2, False), (1, 2, True)])
def test_wide_kernel_with_bulk_copied_plan_rows(ntx, nrx, compact, engines):
    """The wide-store slot kernel takes its interpolation plan rows from a shared-memory ring filled by cp.async.bulk
    (default; B2C_PLAN_BULK=0 selects the older per-thread cp.async staging): same entries, same arithmetic -- every output bit-identical, over
    several waves of CTAs and mixed profiles / patterns."""
    eng = engines(ntx, nrx)
    pool = eng.random_pool([0.10, 0.03], seed=21)
    B = 700
    rng = np.random.default_rng(2)
    kw = dict(model_id=rng.integers(0, 3, B).astype(np.int32), doppler_hz=rng.uniform(5, 200, B).astype(np.float32),
              snr_db=rng.uniform(-5, 30, B).astype(np.float32), pattern_id=rng.integers(0, 2, B).astype(np.int32), pool=pool,
              slot0=5150, seed=8, pitch=600, compact=compact)
    res = []
    for flag in ("0", "1"):
        os.environ["B2C_PLAN_BULK"] = flag
        try:
            res.append(eng.run(B, **kw))
            torch.cuda.synchronize()
        finally:
            del os.environ["B2C_PLAN_BULK"]
    dflt = eng.run(B, **kw)
    torch.cuda.synchronize()
    for k in ("H_true", "rx", "tx", "H_ls", "H_mmse", "stats"):
        assert torch.equal(res[0][k], res[1][k]) and torch.equal(dflt[k], res[1][k]), k


@pytest.mark.parametrize("nsym", [2, 6, 16])
def test_wide_kernels_on_short_and_long_slots(nsym):
    """599 bins with 2 / 6 / 16 symbols per slot: fewer symbols than the plan ring has rows, a partial second lap, the
    compiled maximum.  The wide-store kernel (pitch 600), the statistics kernel and the dense scoring pass must agree with
    the generic contiguous-row kernel (the one the oracle parity tests above pin) and with their per-thread staging forms."""
    from engine import SlotEngine, WienerBank
    cfg = {"ofdm": {"fft_size": 1024, "cp_length": 72, "num_symbols": nsym, "useful_subcarriers": 600, "subcarrier_spacing": 15000},
           "mimo": {"num_tx_antennas": 2, "num_rx_antennas": 2}}
    eng = SlotEngine(cfg)
    pool = eng.random_pool([0.08], seed=3)
    B = 300
    rng = np.random.default_rng(nsym)
    snrs = [0.0, 15.0]
    snr = np.asarray(snrs, np.float32)[rng.integers(0, 2, B)]
    kw = dict(model_id=rng.integers(0, 3, B).astype(np.int32), doppler_hz=rng.uniform(5, 200, B).astype(np.float32), snr_db=snr,
              pattern_id=0, pool=pool, slot0=12, seed=4)
    generic = eng.run(B, **kw)
    wide = eng.run(B, pitch=600, **kw)
    stats_only = eng.run(B, want=("stats",), **kw)
    idx = pool.pilot_indices[0]
    bank = WienerBank(eng, pool, {0: model_cov(idx)}, snrs)
    dense = eng.run(B, mmse="dense", wiener=bank, **kw)
    lean = eng.run(B, mmse="dense", wiener=bank, want=("stats",), **kw)
    torch.cuda.synchronize()
    for k in ("H_true", "rx", "tx", "H_ls", "H_mmse"):
        assert relerr(wide[k].cpu().numpy(), generic[k].cpu().numpy()) < 2e-6, k
    assert torch.allclose(wide["stats"], generic["stats"], rtol=2e-5) and torch.allclose(stats_only["stats"], generic["stats"], rtol=2e-5)
    assert torch.allclose(lean["stats"], dense["stats"], rtol=2e-5)
    os.environ["B2C_PLAN_BULK"] = "0"
    try:
        wide0 = eng.run(B, pitch=600, **kw)
        lean0 = eng.run(B, mmse="dense", wiener=bank, want=("stats",), **kw)
        torch.cuda.synchronize()
    finally:
        del os.environ["B2C_PLAN_BULK"]
    for k in ("H_true", "rx", "tx", "H_ls", "H_mmse", "stats"):
        assert torch.equal(wide0[k], wide[k]), k
    assert torch.equal(lean0["stats"], lean["stats"])


@pytest.mark.parametrize("ntx,nrx", [(4, 4), (2, 2), (1, 1), (8, 2)])
def test_register_blocked_statistics_kernel(ntx, nrx, engines):
    """Statistics-only sweeps run slot2_kernel (160 threads, two mirror pairs = four bins per thread): same Philox
    counters and per-bin arithmetic as the 320-thread wide kernel, so the per-slot sums agree to the rounding of the
    partial-sum grouping (four bins per thread instead of two) with both that kernel and the full pipeline."""
    eng = engines(ntx, nrx)
    pool = eng.random_pool([0.10, 0.03], seed=13)
    B = 9
    pid = np.arange(B, dtype=np.int32) % 2
    mid = np.array([0, 1, 2, 2, 1, 0, 2, 1, 0], np.int32)                         # mixed profiles: 5 / 8 / 9 taps
    fd = np.array([10, 50, 100, 200, 70, 30, 150, 5, 90], np.float32)
    snr = np.linspace(-5, 30, B).astype(np.float32)
    kw = dict(model_id=mid, doppler_hz=fd, snr_db=snr, pattern_id=pid, pool=pool, slot0=31337, seed=5)
    got = eng.run(B, want=("stats",), **kw)["stats"]
    os.environ["B2C_NO_SLOT2"] = "1"
    try:
        ref = eng.run(B, want=("stats",), **kw)["stats"]
    finally:
        del os.environ["B2C_NO_SLOT2"]
    full = eng.run(B, pitch=600, **kw)["stats"]
    # plan rows by bulk copy (B2C_PLAN_BULK=1; the scoring pass's default) against the per-thread cp.async staging: same entries, same arithmetic, same bits --
    # on a batch large enough for several waves of CTAs (ring reuse, barrier races would show as differing sums)
    Bb = 1200
    kwb = dict(model_id=np.resize(mid, Bb), doppler_hz=np.resize(fd, Bb), snr_db=np.resize(snr, Bb), pattern_id=np.resize(pid, Bb),
               pool=pool, slot0=99, seed=5)
    forms = []
    for flag in ("1", "0"):
        os.environ["B2C_PLAN_BULK"] = flag
        try:
            forms.append(eng.run(Bb, want=("stats",), **kwb)["stats"])
        finally:
            del os.environ["B2C_PLAN_BULK"]
    torch.cuda.synchronize()
    assert torch.equal(forms[0], forms[1])
    assert torch.allclose(got, ref, rtol=2e-6, atol=0)
    assert torch.allclose(got, full, rtol=2e-6 if ntx < 4 else 2e-5, atol=0)
    assert not torch.equal(got, torch.zeros_like(got))


def test_error_reporting(engines):
    import _b2c
    eng = engines(2, 2)
    with pytest.raises(ValueError):
        eng.run(1, 0, 10.0, 10.0)                          # estimation outputs without a pattern pool
    with pytest.raises(_b2c.B2CError):
        eng.ofdm_modulate(torch.zeros((2, 599), dtype=torch.complex64))   # CPU tensor
    out = eng.run(0, 0, 10.0, 10.0, want=("H_true",))       # empty batch is a no-op
    assert out["H_true"].shape[0] == 0


@pytest.mark.parametrize("pitch", [600, None])
def test_full_size_properties(pitch, engines):
    """BASELINE config 3 shape (4x4 ETU 200 Hz, 10 % pilots) at a batch too large for the oracle:
    size-independent properties of the reference's algorithm, in the throughput layout bench.py measures
    (rows at pitch 600, wide-store kernel) and in the contiguous one."""
    eng = engines(4, 4)
    pool = eng.random_pool([0.10], seed=42)
    B = 512
    snr = np.array([-5, 0, 5, 10, 15, 20, 25, 30], dtype=np.float32)[np.arange(B) % 8]
    out = eng.run(B, eng.models.index("ETU"), 200.0, snr, 0, pool, slot0=0, seed=42, pitch=pitch)
    torch.cuda.synchronize()
    H, rx, tx, Hl, Hm = (out[k] for k in ("H_true", "rx", "tx", "H_ls", "H_mmse"))
    for t in (H, rx, tx, Hl, Hm):
        assert torch.isfinite(torch.view_as_real(t)).all()
    # unit-modulus symbols, the same grid on every TX antenna (:402-404)
    assert (tx.abs() - 1).abs().max() < 1e-5
    assert torch.equal(tx[:, :, 0], tx[:, :, 3])
    # Jakes normalisation: E|H|^2 = sum of surviving path powers / 2 = 0.5 for ETU
    assert abs(H.abs().pow(2).mean().item() - 0.5) < 0.02
    # estimates are replicated over tx; MMSE is a positive real shrinkage of LS per (slot, rx)
    assert torch.equal(Hl[:, :, :, 0], Hl[:, :, :, 2]) and torch.equal(Hm[:, :, :, 1], Hm[:, :, :, 3])
    ratio = (Hm[:, :, :, 0] * Hl[:, :, :, 0].conj()).real.sum(dim=(1, 3)) / Hl[:, :, :, 0].abs().pow(2).sum(dim=(1, 3))
    assert (ratio > 0).all() and (ratio < 1).all()
    # zeros exactly where the plan says "outside the hull"
    mask_out = torch.from_numpy(~pool_inside(pool)).to(eng.device)
    assert (Hl[:, :, 0, 0][:, mask_out] == 0).all() and (Hl[:, :, 0, 0][:, ~mask_out] != 0).any()
    # empirical SNR of rx against the noiseless sum_tx H x
    y = (H.sum(dim=3) * tx[:, :, :1])
    emp = 10 * torch.log10(y.abs().pow(2).mean(dim=(1, 2, 3)) / (rx - y).abs().pow(2).mean(dim=(1, 2, 3)))
    assert (emp.cpu().numpy() - snr).__abs__().max() < 0.15
    # LS at a pilot RE equals rx / x there
    e = int(pool.pilot_indices[0][123])
    s, k = divmod(e, 599)
    assert relerr((rx[:, s, :, k] / tx[:, s, 0, k][:, None]).cpu().numpy(), Hl[:, s, :, 0, k].cpu().numpy()) < RTOL


def pool_inside(pool):
    import _tables
    plan = _tables.cached_plan(pool.pilot_indices[0], pool.nsym, pool.nsc, pool.method)
    return plan["flags"].astype(bool).reshape(pool.nsym, pool.nsc)


@pytest.mark.parametrize("ntx,nrx,nsym,useful,model", [(3, 2, 7, 300, "EVA"), (1, 3, 5, 72, "EPA"), (8, 8, 14, 600, "ETU"),
                                                       (2, 5, 16, 640, "ETU")])
def test_other_geometries_generic_path(ntx, nrx, nsym, useful, model):
    """Non-default grids (odd symbol counts, other subcarrier counts, non power-of-two antenna
    counts, the 8x8 / 16-symbol / 639-bin limits) run the generic instantiations; same parity bar."""
    from engine import SlotEngine
    cfg = {"ofdm": {"fft_size": 1024, "cp_length": 72, "num_symbols": nsym, "useful_subcarriers": useful,
                    "subcarrier_spacing": 15000}, "mimo": {"num_tx_antennas": ntx, "num_rx_antennas": nrx}}
    eng = SlotEngine(cfg)
    nsc = useful - 1
    assert eng.nsc == nsc
    dens = 0.08
    pool = eng.random_pool([dens], seed=5)
    B, seed, slot0, snr, fd = 2, 99, 7, 12.0, 70.0
    out = eng.run(B, eng.models.index(model), fd, snr, 0, pool, slot0=slot0, seed=seed)
    torch.cuda.synchronize()
    mask = pool.mask(0)
    for i in range(B):
        d = opx.slot_draws(seed, slot0 + i, nsym, nsc, len(orc.TDL_NS[model]), ntx, nrx, mask)
        d["perm"] = np.concatenate([pool.pilot_indices[0], np.setdiff1d(np.arange(nsym * nsc), pool.pilot_indices[0])])
        ref = orc.slot_pipeline(cfg["ofdm"], ntx, nrx, model, fd, snr, dens, d)
        assert np.array_equal(ref["pilot_mask"], mask)
        for k, rk in (("H_true", "channel"), ("rx", "rx_symbols"), ("tx", "tx_symbols"), ("H_ls", "H_ls"), ("H_mmse", "H_mmse")):
            assert relerr(out[k][i].cpu().numpy(), ref[rk]) < RTOL, (k, i)
        st = out["stats"][i].cpu().numpy()
        m_ls = orc.evaluate(ref["channel"], ref["H_ls"])
        tot = st[:, 1].sum(axis=0) / ref["channel"].size
        assert abs(db(tot[0] / (tot[2] + 1e-12)) - m_ls["nmse_db"]) < DB_TOL
        assert abs(st[0, 0, 0] / (nsym * nsc) / (st[0, 0, 2] / (nsym * nsc) + 1e-10) / orc.nmse_pair00(ref["H_ls"], ref["channel"]) - 1) < 1e-4
    # stand-alone K3 on the same grids
    rx, Ht = out["rx"], out["H_true"]
    xp = out["tx"][:, :, 0].reshape(B, -1)[:, torch.from_numpy(pool.pilot_indices[0]).to(eng.device)].contiguous()
    k3 = eng.ls_interp(rx, xp, pool, snr_db=snr, mmse=True, H_true=Ht, want=("H_ls", "H_mmse", "stats"))
    assert (k3["H_ls"] - out["H_ls"]).abs().max().item() < 2e-5 and (k3["H_mmse"] - out["H_mmse"]).abs().max().item() < 2e-5
    assert torch.allclose(k3["stats"], out["stats"], rtol=1e-4)


@pytest.mark.parametrize("ntx,nrx,nsym,useful,want", [(3, 2, 7, 300, ("H_true", "rx", "tx", "H_ls", "H_mmse", "stats")),
                                                        (4, 4, 14, 600, ("H_true", "rx", "tx", "H_ls", "stats")),
                                                        (2, 2, 14, 600, ("H_true", "rx", "tx"))])
def test_host_pipeline_other_geometries_and_output_sets(ntx, nrx, nsym, useful, want):
    """HostPipeline's slab layout on a grid the wide kernels do not take (7 x 299, 3 TX: contiguous rows, generic kernel),
    in dataset mode (no H_mmse: what generate_sample returns) and simulate-only: same arrays as SlotEngine.run, ragged tail."""
    from engine import SlotEngine
    from host_pipeline import HostPipeline
    cfg = {"ofdm": {"fft_size": 1024, "cp_length": 72, "num_symbols": nsym, "useful_subcarriers": useful,
                    "subcarrier_spacing": 15000}, "mimo": {"num_tx_antennas": ntx, "num_rx_antennas": nrx}}
    eng = SlotEngine(cfg)
    est = "H_ls" in want
    pool = eng.random_pool([0.08], seed=5) if est else None
    n = 21
    snr = np.linspace(0, 20, n).astype(np.float32)
    ref = eng.run(n, 1, 60.0, snr, 0, pool, slot0=40, seed=3, want=want)
    for compact in (False, True):
        hp = HostPipeline(eng, pool, chunk=8, want=want, compact=compact)
        assert hp.pitch == (600 if (useful == 600 and ntx in (1, 2, 4, 8)) else eng.nsc)
        got = {}

        def consume(first, cnt, host):
            for k, v in host.items():
                got.setdefault(k, []).append(np.array(v[:cnt]))
        assert hp.run(np.full(n, 1), np.full(n, 60.0), snr, np.zeros(n), slot0=40, seed=3, consume=consume) == n
        for k in want:
            host = np.concatenate(got[k])
            assert host.shape == tuple(ref[k].shape), (k, host.shape)
            assert np.abs(host - ref[k].cpu().numpy()).max() <= 2e-6 * max(1.0, float(ref[k].abs().max())), (k, compact)


def test_compact_layout_and_host_pipeline(engines):
    """compact=True writes the tx-replicated arrays once; expanded views equal the full layout.  The
    host-buffer pipeline (pinned memory, double-buffered D2H) hands back the same arrays."""
    from host_pipeline import HostPipeline
    eng = engines(4, 4)
    pool = eng.random_pool([0.10], seed=2)
    n = 37                                                # ragged: 3 chunks of 16 -> 16, 16, 5
    snr = np.array([-5, 0, 5, 10, 15, 20, 25, 30], np.float32)[np.arange(n) % 8]
    args = dict(model_id=2, doppler_hz=200.0, snr_db=snr, pattern_id=0, pool=pool, slot0=500, seed=8)
    full = eng.run(n, **args)
    comp = eng.run(n, compact=True, **args)
    assert comp["H_ls"].shape == (n, 14, 4, 599) and comp["tx"].shape == (n, 14, 599)
    exp = eng.expand_compact(comp)
    torch.cuda.synchronize()
    for k in ("H_true", "rx", "tx", "H_ls", "H_mmse"):
        assert exp[k].shape == full[k].shape and (exp[k] - full[k]).abs().max().item() < 2e-6, k
    assert torch.allclose(comp["stats"], full["stats"], rtol=1e-5)
    for compact in (False, True):
        got = {}

        def consume(first, cnt, host):
            for k, v in host.items():
                got.setdefault(k, {})[first] = np.array(v[:cnt])      # copy out: buffers are reused
        hp = HostPipeline(eng, pool, chunk=16, compact=compact)
        done = hp.run(np.full(n, 2), np.full(n, 200.0), snr, np.zeros(n), slot0=500, seed=8, consume=consume)
        assert done == n and sorted(got["rx"]) == [500, 516, 532]
        for k in ("H_true", "rx", "tx", "H_ls", "H_mmse"):
            host = np.concatenate([got[k][f] for f in sorted(got[k])])
            assert host.shape == tuple(full[k].shape)
            assert np.abs(host - full[k].cpu().numpy()).max() < 2e-6, (k, compact)
        if compact:
            # work-sharing front end: chunks claimed from a shared counter (here a local one, claimed out of order by two
            # interleaved "ranks") produce the same arrays as the static run -- results follow the global slot index
            nxt, got2 = [0], {}

            def claim(k):
                first = nxt[0]
                nxt[0] += k
                return first + 500 if first < n else None

            def params_of(first, k):
                j = np.arange(first - 500, min(first - 500 + k, n))
                return np.stack([np.full(len(j), 2.0), np.full(len(j), 200.0), snr[j], np.zeros(len(j))])

            def consume2(first, cnt, host):
                got2[first] = {k: np.array(v[:cnt]) for k, v in host.items()}
            assert HostPipeline(eng, pool, chunk=16, compact=True, depth=3).run_dynamic(claim, params_of, seed=8, consume=consume2) == n
            assert sorted(got2) == [500, 516, 532] and got2[532]["rx"].shape[0] == 5
            for f in got2:
                for k in ("H_true", "rx", "H_ls"):
                    assert np.array_equal(got2[f][k], got[k][f]), (f, k)
        # what crosses the link: the unique bytes in padded rows (600 / 599) + statistics + 256-byte array alignment
        unique = (1945552 if compact else 3756928) * 600 / 599 + 192
        assert unique <= hp.d2h_bytes_per_slot <= unique + 6 * 256 / 16
        assert hp.pitch == 600


@pytest.mark.parametrize("ntx,nrx,model", [(4, 4, 2), (2, 2, 1), (1, 1, 0), (8, 2, 2)])
def test_compact_padded_layout_matches_full(ntx, nrx, model, engines):
    """compact = 1 with pitch 600 (the throughput form of the unique-bytes layout: H_ls / H_mmse [B,14,nrx,600],
    tx [B,14,600], one 16-byte store per lane) holds exactly the values of the full padded layout, inside guards."""
    import _b2c
    eng = engines(ntx, nrx)
    pool = eng.random_pool([0.10, 0.05], seed=4)
    B = 6
    pid = np.arange(B, dtype=np.int32) % 2
    snr = np.linspace(-5, 30, B).astype(np.float32)
    args = dict(model_id=model, doppler_hz=120.0, snr_db=snr, pattern_id=pid, pool=pool, slot0=7654321, seed=17)
    full = eng.run(B, pitch=_b2c.WIDE_PITCH, **args)
    # guarded buffers: [guard | payload | guard] per array, sentinel must survive
    P, nsym = _b2c.WIDE_PITCH, 14
    shapes = {"H_true": (B, nsym, nrx, ntx, P), "H_ls": (B, nsym, nrx, P), "H_mmse": (B, nsym, nrx, P),
              "rx": (B, nsym, nrx, P), "tx": (B, nsym, P)}
    G, bufs, out = 4096, {}, {}
    for k, sh in shapes.items():
        n = int(np.prod(sh))
        bufs[k] = torch.full((n + 2 * G,), 7.5 + 7.5j, dtype=torch.complex64, device=eng.device)
        out[k] = bufs[k][G:G + n].view(sh)[..., :599]
    out["stats"] = torch.empty((B, nrx, 2, 3), dtype=torch.float64, device=eng.device)
    comp = eng.run(B, out=out, compact=True, **args)
    torch.cuda.synchronize()
    for k, sh in shapes.items():
        assert torch.all(bufs[k][:G] == 7.5 + 7.5j) and torch.all(bufs[k][-G:] == 7.5 + 7.5j), k
        assert comp[k].stride(-2) == P
    exp = eng.expand_compact(comp)
    for k in ("H_true", "rx", "tx", "H_ls", "H_mmse"):
        assert exp[k].shape == full[k].shape and torch.equal(exp[k], full[k]), k
    assert torch.equal(comp["stats"], full["stats"])
    # simulate-only and dataset-mode (no H_mmse, no stats) calls of the same form
    sim = eng.run(B, want=("H_true", "rx", "tx"), compact=True, pitch=P, **{k: v for k, v in args.items() if k not in ("pool", "pattern_id")})
    dm = eng.run(B, want=("H_true", "rx", "tx", "H_ls"), compact=True, pitch=P, **args)
    torch.cuda.synchronize()
    assert sim["tx"].shape == (B, nsym, 599) and torch.equal(eng.expand_compact(sim)["tx"], full["tx"])
    assert torch.equal(sim["H_true"], full["H_true"]) and torch.equal(sim["rx"], full["rx"])
    assert torch.equal(eng.expand_compact(dm)["H_ls"], full["H_ls"]) and torch.equal(dm["rx"], full["rx"])
    # compact-layout readers (ls_sym_stride = nrx * pitch)
    x_f, t_f = eng.ml_features(full["rx"], full["H_ls"], full["H_true"], pool, pid, "last", True)
    x_c, t_c = eng.ml_features(comp["rx"], comp["H_ls"], comp["H_true"], pool, pid, "last", True)
    assert torch.equal(x_f, x_c) and torch.equal(t_f, t_c)


def test_dense_real_map_and_cubic_ls(engines):
    """K4b: real [m x k] map on complex columns (tcgen05, 3xTF32), and the cubic LS estimate built on it."""
    eng = engines(2, 2)
    dev = eng.device
    rng = np.random.default_rng(4)
    for m, k, c, ld_in in ((300, 167, 77, 170), (8386, 838, 70, 838), (129, 33, 3, 33)):
        W = rng.standard_normal((m, k)) / np.sqrt(k)
        X = np.zeros((c, ld_in), complex)
        X[:, :k] = rng.standard_normal((c, k)) + 1j * rng.standard_normal((c, k))
        Y = eng.dense_real_apply(torch.from_numpy(W).to(dev, torch.float32), torch.from_numpy(X).to(dev, torch.complex64),
                                 ld_out=m + 5).cpu().numpy()
        assert Y.shape == (c, m + 5) and not Y[:, m:].any()
        assert relerr(Y[:, :m], X[:, :k] @ W.T) < 5e-5, (m, k, c)
        Yp = eng.dense_real_apply(eng.prepare_dense(torch.from_numpy(W).to(dev, torch.float32)),
                                  torch.from_numpy(X).to(dev, torch.complex64), ld_out=m + 5).cpu().numpy()
        assert np.array_equal(Yp, Y), (m, k, c)             # prepared operand: bit-identical
    g, cub = load_golden("slot_2x2_eva"), load_golden("ls_cubic")["slot_2x2_eva"]
    rx = torch.from_numpy(g["rx_symbols"][None]).to(dev, torch.complex64)
    xp = torch.from_numpy(g["pilot_symbols"][None]).to(dev, torch.complex64)
    Ht = torch.from_numpy(g["channel"][None]).to(dev, torch.complex64)
    out = eng.ls_cubic(rx, xp, g["pilot_indices"], H_true=Ht, want=("H_ls", "stats"))
    H = out["H_ls"][0].cpu().numpy()
    for t in range(2):
        assert relerr(H[:, :, t], cub) < RTOL
    assert np.array_equal(H[:, :, 0] == 0, cub == 0)
    st = out["stats"][0].cpu().numpy()[:, 1].sum(axis=0) / g["channel"].size
    want = orc.evaluate(g["channel"], np.repeat(cub[:, :, None, :], 2, axis=2))
    assert abs(db(st[0] / (st[2] + 1e-12)) - want["nmse_db"]) < DB_TOL


def test_link_level_chain_at_full_size(engines):
    """Size-independent properties of the 'next' rows on 4x4 slots through the C ABI:
    bits -> QAM -> H x (noise 60 dB down) -> ZF equaliser with the true channel -> demap recovers every bit;
    the equaliser in complex64 I/O matches the fp64 oracle on one slot; the feature batch is consistent."""
    eng = engines(4, 4)
    B, nsym, nsc = 64, 14, 599
    pool = eng.random_pool([0.10], 1)
    out = eng.run(B, 0, 10.0, 20.0, 0, pool, slot0=11, seed=7, want=("H_true", "rx", "H_ls"))     # EPA: well conditioned mostly
    H = out["H_true"]
    rng = np.random.default_rng(5)
    for M in (4, 16):
        bps = 2 if M == 4 else 4
        bits = torch.from_numpy(rng.integers(0, 2, B * nsym * 4 * nsc * bps, dtype=np.uint8)).to(eng.device)
        x = eng.qam_modulate(bits, M).reshape(B, nsym, 4, nsc)
        y = eng.apply_channel(x.contiguous(), H, 60.0, slot0=3, seed=9)
        xh = eng.equalize(y, H, "zf")
        back = eng.qam_demodulate(xh.reshape(-1), M)
        nerr = int(eng.count_bit_errors(bits, back).item())
        assert nerr <= bits.numel() * 1e-3, (M, nerr)        # only REs where the 4x4 channel is near-singular
        if M == 4:
            ref = orc.equalize(y[0].cpu().numpy(), H[0].cpu().numpy(), "mmse")
            got = eng.equalize(y[:1], H[:1], "mmse")[0].cpu().numpy()
            assert relerr(got, ref) < 1e-5
    # feature batch: mask channel equals the pool mask, normalised signal channels have unit joint std
    xin, tgt = eng.ml_features(out["rx"], out["H_ls"], H, pool, 0, "last", True)
    assert xin.shape == (B, nsym, nsc, 5) and tgt.shape == (B, nsym, nsc, 2)
    assert np.array_equal(xin[3, :, :, 4].cpu().numpy().astype(bool), pool.mask(0))
    assert abs(float(xin[..., :4][5].double().std(unbiased=False)) - 1.0) < 1e-4
    assert abs(float(tgt[5].double().std(unbiased=False)) - 1.0) < 1e-4
    ref_x, ref_t = orc.ml_inputs(out["rx"][2].cpu().numpy(), out["H_ls"][2].cpu().numpy(), H[2].cpu().numpy(), pool.mask(0))
    assert relerr(xin[2].cpu().numpy(), ref_x) < RTOL and relerr(tgt[2].cpu().numpy(), ref_t) < RTOL
    err = eng.pair00_errors(out["H_ls"], H).cpu().numpy()
    d = (out["H_ls"][:, :, 0, 0] - H[:, :, 0, 0]).cpu().numpy()
    assert np.allclose(err[:, 0], (np.abs(d) ** 2).sum(axis=(1, 2)), rtol=1e-5)
    assert np.allclose(err[:, 1], err[:, 0], rtol=1e-12)                       # alpha = 1


@pytest.mark.parametrize("ntx,nrx,model", [(4, 4, 2), (2, 2, 1), (1, 1, 0), (8, 2, 2)])
def test_padded_row_layout_matches_contiguous(ntx, nrx, model, engines):
    """The wide-store kernel (rows padded to pitch 600, one 16-byte store per lane) is the same arithmetic as the
    contiguous throughput kernel: identical arrays, identical statistics; and the pitch-aware readers accept it."""
    import _b2c
    eng = engines(ntx, nrx)
    pool = eng.random_pool([0.10, 0.05], seed=4)
    B = 7
    pid = np.arange(B, dtype=np.int32) % 2
    snr = np.linspace(-5, 30, B).astype(np.float32)
    args = dict(model_id=model, doppler_hz=120.0, snr_db=snr, pattern_id=pid, pool=pool, slot0=1234567, seed=99)
    ref = eng.run(B, **args)
    pad = eng.run(B, pitch=_b2c.WIDE_PITCH, **args)
    sim_ref = eng.run(B, want=("H_true", "rx", "tx"), **{k: v for k, v in args.items() if k not in ("pool", "pattern_id")})
    sim_pad = eng.run(B, want=("H_true", "rx", "tx"), pitch=_b2c.WIDE_PITCH,
                      **{k: v for k, v in args.items() if k not in ("pool", "pattern_id")})
    torch.cuda.synchronize()
    for k in ("H_true", "rx", "tx", "H_ls", "H_mmse"):
        assert pad[k].shape == ref[k].shape and pad[k].stride(-2) == _b2c.WIDE_PITCH
        assert torch.equal(pad[k], ref[k]), k                      # same operations in the same order: bit-identical
        if k in sim_pad:
            assert torch.equal(sim_pad[k], sim_ref[k]), k
    # ntx >= 4: the wide kernel folds the error sums over tx (sum|h|^2 - 2 Re(conj(l) sum h) + ntx |l|^2), same value
    # to fp32 rounding of O(power)-sized terms
    assert torch.allclose(pad["stats"], ref["stats"], rtol=1e-6 if ntx < 4 else 2e-5, atol=0)
    # dataset mode: what the reference's generate_sample returns (no H_mmse, no statistics), same wide-store kernel
    dm = eng.run(B, want=("H_true", "rx", "tx", "H_ls"), pitch=_b2c.WIDE_PITCH, **args)
    assert set(k for k in dm if not k.startswith("_")) == {"H_true", "rx", "tx", "H_ls"}
    for k in ("H_true", "rx", "tx", "H_ls"):
        assert dm[k].stride(-2) == _b2c.WIDE_PITCH and torch.equal(dm[k], ref[k]), k
    # statistics-only call (pilot-density / SNR sweeps): the store-free instantiation gives the same sums
    so = eng.run(B, want=("stats",), **args)
    assert set(k for k in so if not k.startswith("_")) == {"stats"}
    assert torch.allclose(so["stats"], ref["stats"], rtol=1e-6 if ntx < 4 else 2e-5, atol=0)
    # against the oracle on the Philox twin draws for one slot (the contiguous path is pinned the same way)
    assert relerr(pad["H_true"][3].cpu().numpy(), ref["H_true"][3].cpu().numpy()) == 0.0
    # pitch-aware readers
    x_ref, t_ref = eng.ml_features(ref["rx"], ref["H_ls"], ref["H_true"], pool, pid, "last", True)
    x_pad, t_pad = eng.ml_features(pad["rx"], pad["H_ls"], pad["H_true"], pool, pid, "last", True)
    assert torch.equal(x_ref, x_pad) and torch.equal(t_ref, t_pad)
    assert torch.equal(eng.pair00_errors(ref["H_ls"], ref["H_true"], 0.5), eng.pair00_errors(pad["H_ls"], pad["H_true"], 0.5))
    assert torch.allclose(eng.pair00_moments(ref["rx"], ref["H_ls"], ref["H_true"]),
                          eng.pair00_moments(pad["rx"], pad["H_ls"], pad["H_true"]), rtol=1e-12)
    # the padded layout is only offered by the throughput configuration
    with pytest.raises(_b2c.B2CError):
        eng.run(B, want=("H_true", "rx"), pitch=_b2c.WIDE_PITCH, **{k: v for k, v in args.items() if k not in ("pool", "pattern_id")})
    # stand-alone K3 on the padded arrays (wide-access kernel) == K3 on the contiguous ones == the fused estimates
    pil = torch.zeros((B, pool.np_max), dtype=torch.complex64, device=eng.device)
    for b in range(B):
        idx = torch.from_numpy(pool.pilot_indices[pid[b]]).to(eng.device)
        pil[b, :idx.numel()] = ref["tx"][b, :, 0, :].reshape(-1)[idx]
    kw = dict(pattern_id=pid, snr_db=snr, mmse=True, want=("H_ls", "H_mmse", "stats"))
    k3_ref = eng.ls_interp(ref["rx"], pil, pool, H_true=ref["H_true"], **kw)
    k3_pad = eng.ls_interp(pad["rx"], pil, pool, H_true=pad["H_true"], **kw)
    assert k3_pad["H_ls"].stride(-2) == _b2c.WIDE_PITCH
    for k in ("H_ls", "H_mmse"):
        assert torch.equal(k3_pad[k], k3_ref[k]), k
        assert (k3_pad[k] - ref[k]).abs().max().item() < 2e-6, k
    assert torch.allclose(k3_pad["stats"], k3_ref["stats"], rtol=1e-6, atol=0)
    assert torch.allclose(k3_pad["stats"], ref["stats"], rtol=1e-4, atol=0)
    with pytest.raises(_b2c.B2CError):          # mixed layouts are refused
        eng.ls_interp(pad["rx"], pil, pool, H_true=ref["H_true"], **kw)
    # caller-made padded buffers whose padding element holds NaN: never read into a result
    def nan_padded(t):
        buf = torch.full(t.shape[:-1] + (_b2c.WIDE_PITCH,), float("nan"), dtype=torch.complex64, device=t.device)
        buf[..., :599] = t
        return buf[..., :599]
    k3_nan = eng.ls_interp(nan_padded(ref["rx"]), pil, pool, H_true=nan_padded(ref["H_true"]), **kw)
    assert torch.equal(k3_nan["stats"], k3_pad["stats"]) and torch.equal(k3_nan["H_ls"], k3_pad["H_ls"])
    xn, tn = eng.ml_features(nan_padded(ref["rx"]), nan_padded(ref["H_ls"]), nan_padded(ref["H_true"]), pool, pid, "last", True)
    assert torch.equal(xn, x_ref) and torch.equal(tn, t_ref)


@pytest.mark.parametrize("pitch", [600, None])
def test_slot_pipeline_writes_stay_inside_their_arrays(pitch, engines):
    """No out-of-bounds store (compute-sanitizer is not available on this pool): every output array sits between two
    guard regions filled with a sentinel; after the run the guards are intact, every payload element was written and
    the padding element of each row (pitch 600) holds a finite value."""
    eng = engines(4, 4)
    pool = eng.random_pool([0.10], seed=6)
    B, nsym, nsc, P = 5, 14, 599, (pitch or 599)
    guard = 4096                                    # complex elements on each side
    sentinel = torch.tensor(complex(-7.25e33, 3.5e-33), dtype=torch.complex64, device=eng.device)
    shapes = {"H_true": (B, nsym, 4, 4, P), "H_ls": (B, nsym, 4, 4, P), "H_mmse": (B, nsym, 4, 4, P), "rx": (B, nsym, 4, P),
              "tx": (B, nsym, 4, P)}
    bufs, out = {}, {}
    for k, shp in shapes.items():
        n = int(np.prod(shp))
        bufs[k] = torch.full((n + 2 * guard,), sentinel.item(), dtype=torch.complex64, device=eng.device)
        out[k] = bufs[k][guard:guard + n].view(shp)[..., :nsc]
    out["stats"] = torch.zeros((B, 4, 2, 3), dtype=torch.float64, device=eng.device)
    eng.run(B, 2, 200.0, 10.0, 0, pool, slot0=77, seed=5, out=out)
    torch.cuda.synchronize()
    for k, shp in shapes.items():
        n = int(np.prod(shp))
        assert (bufs[k][:guard] == sentinel).all() and (bufs[k][guard + n:] == sentinel).all(), k
        assert not (out[k] == sentinel).any(), k                       # every payload element was written
        if pitch:
            pad = bufs[k][guard:guard + n].view(shp)[..., nsc:]
            assert torch.isfinite(torch.view_as_real(pad)).all(), k
    ref = eng.run(B, 2, 200.0, 10.0, 0, pool, slot0=77, seed=5, pitch=pitch)
    for k in shapes:
        assert torch.equal(out[k], ref[k]), k


def test_dense_products_write_only_their_outputs(engines):
    """Guard regions around the outputs of the tensor-core products (one-shot, prepared 128-/256-column and
    warp-specialised forms, ragged sizes): nothing outside [ncols x m] is touched."""
    eng = engines(2, 2)
    dev = eng.device
    rng = np.random.default_rng(11)
    sentinel = complex(-7.25e33, 3.5e-33)
    guard = 2048

    def guarded(rows, ld):
        buf = torch.full((rows * ld + 2 * guard,), sentinel, dtype=torch.complex64, device=dev)
        return buf, buf[guard:guard + rows * ld].view(rows, ld)

    def intact(buf, rows, ld, m):
        body = buf[guard:guard + rows * ld].view(rows, ld)
        return bool((buf[:guard] == sentinel).all() and (buf[guard + rows * ld:] == sentinel).all()
                    and (body[:, m:] == sentinel).all() and not (body[:, :m] == sentinel).any())

    # 4000 / 1000 columns: 256-column tiles, warp-specialised ring; np = 12 / 40: one- and three-stage rings
    for n, c in ((167, 77), (838, 300), (838, 1000), (838, 4000), (12, 600), (40, 777)):
        ld = n + 3
        A = torch.from_numpy((rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(n)).to(dev, torch.complex64)
        X = torch.zeros((c, ld), dtype=torch.complex64, device=dev)
        X[:, :n] = torch.from_numpy(rng.standard_normal((c, n)) + 1j * rng.standard_normal((c, n))).to(dev, torch.complex64)
        want = eng.mmse_dense(A, X)
        for W in (A, eng.prepare_dense(A)):
            buf, out = guarded(c, ld)
            eng.mmse_dense(W, X, out=out)
            torch.cuda.synchronize()
            assert intact(buf, c, ld, n), (n, c, type(W).__name__)
            assert torch.equal(out[:, :n], want[:, :n])
    for m, k, c in ((300, 167, 77), (8386, 838, 70), (8386, 838, 500), (129, 33, 3), (200, 20, 300), (130, 65, 129)):
        Wr = torch.from_numpy(rng.standard_normal((m, k)) / np.sqrt(k)).to(dev, torch.float32)
        X = torch.from_numpy(rng.standard_normal((c, k)) + 1j * rng.standard_normal((c, k))).to(dev, torch.complex64)
        want = eng.dense_real_apply(Wr, X)
        for W in (Wr, eng.prepare_dense(Wr)):
            buf, out = guarded(c, m + 5)
            eng.dense_real_apply(W, X, out=out)
            torch.cuda.synchronize()
            assert intact(buf, c, m + 5, m), (m, k, c, type(W).__name__)
            assert torch.equal(out[:, :m], want)


@pytest.mark.parametrize("pitch", [600, None])
def test_ls_interp_writes_stay_inside_their_arrays(pitch, engines):
    """Guard regions around the stand-alone K3 outputs (contiguous and padded-row kernels)."""
    eng = engines(4, 4)
    pool = eng.random_pool([0.10], seed=6)
    B, nsym, nsc, P = 3, 14, 599, (pitch or 599)
    src = eng.run(B, 1, 50.0, 12.0, 0, pool, slot0=5, seed=8, pitch=pitch)
    idx = torch.from_numpy(pool.pilot_indices[0]).to(eng.device)
    pil = src["tx"][:, :, 0, :].reshape(B, -1)[:, idx].contiguous()
    guard = 4096
    sentinel = complex(-7.25e33, 3.5e-33)
    n = B * nsym * 4 * 4 * P
    bufs = {k: torch.full((n + 2 * guard,), sentinel, dtype=torch.complex64, device=eng.device) for k in ("H_ls", "H_mmse")}
    out = {k: v[guard:guard + n].view(B, nsym, 4, 4, P)[..., :nsc] for k, v in bufs.items()}
    res = eng.ls_interp(src["rx"], pil, pool, snr_db=12.0, mmse=True, H_true=src["H_true"], want=("H_ls", "H_mmse", "stats"), out=out)
    torch.cuda.synchronize()
    for k, v in bufs.items():
        assert (v[:guard] == sentinel).all() and (v[guard + n:] == sentinel).all(), k
        assert not (res[k] == sentinel).any(), k
        assert (res[k] - src[k]).abs().max().item() < 2e-6, k


def test_other_grid_on_reference_draws():
    """The reference's own arrays on a 7 x 299 grid with 3 TX x 2 RX (generic instantiations: odd symbol count, ntx not a
    power of two), injected with its recorded draws; plus the seeded drop-in call."""
    from engine import SlotEngine
    g = load_golden("slot_3x2_eva_7x299")
    nsym, useful, ntx, nrx = int(g["nsym"]), int(g["useful"]), int(g["ntx"]), int(g["nrx"])
    cfg = full_config(ntx, nrx)
    cfg["ofdm"].update(num_symbols=nsym, useful_subcarriers=useful)
    eng = SlotEngine(cfg)
    nsc = useful - 1
    mask = g["pilot_mask"]
    turns = np.empty((nsym, nsc))
    turns[mask] = g["pilot_phase"] / (2 * np.pi)
    turns[~mask] = g["data_phase"] / (2 * np.pi)
    ju = np.zeros((eng.p_max,) + g["jakes_u"].shape[1:])
    ju[:g["jakes_u"].shape[0]] = g["jakes_u"]
    inj = {"jakes_u": torch.from_numpy(ju[None]).to(eng.device, torch.float32),
           "sym_turns": torch.from_numpy(turns[None]).to(eng.device, torch.float32),
           "noise": torch.from_numpy((g["noise_re"] + 1j * g["noise_im"])[None]).to(eng.device, torch.complex64)}
    out = eng.run(1, eng.models.index("EVA"), float(g["doppler_hz"]), float(g["snr_db"]), 0, eng.pool([g["pilot_indices"]]), inject=inj)
    torch.cuda.synchronize()
    assert relerr(out["H_true"][0].cpu().numpy(), g["channel"]) < RTOL and relerr(out["rx"][0].cpu().numpy(), g["rx_symbols"]) < RTOL
    for t in range(ntx):
        assert relerr(out["tx"][0, :, t].cpu().numpy(), g["tx_grid"]) < RTOL
        assert relerr(out["H_ls"][0, :, :, t].cpu().numpy(), g["H_ls_tx0"]) < RTOL
        assert relerr(out["H_mmse"][0, :, :, t].cpu().numpy(), g["H_mmse_tx0"]) < RTOL
    st = out["stats"][0].cpu().numpy()[:, 1].sum(axis=0) / g["channel"].size
    assert abs(db(st[0] / (st[2] + 1e-12)) - g["metrics_ls"][2]) < DB_TOL and abs(db(st[1] / (st[2] + 1e-12)) - g["metrics_mmse"][2]) < DB_TOL
    # drop-in: np.random.seed(606); simulate_transmission(...) on this config, then the two estimators
    import baseline_estimators as be
    import channel_simulator as cs
    np.random.seed(606)
    sim = cs.simulate_transmission(cfg, channel_type="EVA", doppler_hz=70, snr_db=8, pilot_density=0.08)
    assert np.array_equal(sim["pilot_pattern"].pilot_indices, g["pilot_indices"])
    assert relerr(sim["channel"], g["channel"]) < RTOL and relerr(sim["rx_symbols"], g["rx_symbols"]) < RTOL
    pp = sim["pilot_pattern"]
    rx4d = np.repeat(sim["rx_symbols"].reshape(nsym, nrx, 1, nsc), ntx, axis=2)
    H_n = be.LSEstimator('nearest').estimate(rx4d[:, :, :1], sim["pilot_symbols"], pp.pilot_mask, pp.pilot_positions)
    assert relerr(H_n[:, :, 0], g["H_ls_nearest_tx0"]) < RTOL
    H_m = be.MMSEEstimator().estimate(rx4d, sim["pilot_symbols"], pp.pilot_mask, pp.pilot_positions, snr_db=8)
    assert relerr(H_m[:, :, 2], g["H_mmse_tx0"]) < RTOL


def test_c_abi_status_codes_and_empty_inputs(engines):
    """The C ABI never throws: bad arguments give negative status codes and a message, empty inputs are no-ops."""
    import ctypes as C
    import _b2c
    from _b2c import Geom, ref
    L = _b2c.lib()
    eng = engines(2, 2)
    dev = eng.device
    z = torch.zeros((8, 599), dtype=torch.complex64, device=dev)
    bits = torch.zeros((64,), dtype=torch.uint8, device=dev)
    P = lambda t: C.c_void_p(t.data_ptr())
    msg = lambda: L.b2c_last_error_string().decode()
    E_ARG, E_UNSUPPORTED = -1, -3
    # unsupported modulation order (the reference raises NotImplementedError for anything but 4 / 16)
    assert L.b2c_qam_modulate(P(bits), 8, 64, P(z), None) == E_UNSUPPORTED and "64" in msg()
    assert L.b2c_qam_demodulate(P(z), 8, 8, 0, P(bits), None) == E_UNSUPPORTED
    # null pointers
    g = Geom(14, 599, 2, 2, 1024, 72, 7.1e-5)
    assert L.b2c_equalize(ref(g), 1, None, P(z), P(z), 1e-8, 0, None) == E_ARG and "null" in msg()
    assert L.b2c_mmse_dense(None, 4, P(z), P(z), 1, 4, None) == E_ARG
    assert L.b2c_count_bit_errors(P(bits), None, 8, P(bits), None) == E_ARG
    assert L.b2c_dense_apply_grouped(None, 1, P(z), P(z), 4, None) == E_ARG
    grp = (_b2c.DenseGroup * 1)(_b2c.DenseGroup(z.data_ptr(), 0, 4, 4))
    assert L.b2c_dense_apply_grouped(grp, 33, P(z), P(bits), 4, None) == E_UNSUPPORTED and "33" in msg()     # at most 32 groups per call
    assert L.b2c_dense_apply_grouped(grp, 1, P(z), P(z), 4, None) == E_ARG                                    # in-place
    assert L.b2c_dense_apply_grouped(grp, 0, P(z), P(bits), 4, None) == 0                                     # no groups: no-op
    assert L.b2c_tdl_circular(ref(g), ref(eng.prof), None, 1, P(z), P(z), P(z), None) == E_ARG
    assert L.b2c_bit_errors_per_slot(ref(g), ref(eng.random_pool([0.05], seed=1).struct), None, 1, P(bits), P(bits), 2, P(bits), None) == E_ARG
    pool1 = eng.random_pool([0.05], seed=1)
    sl = eng._slots(1, 0, 10.0, 10.0, 0, 0, 1)[0]
    assert L.b2c_dense_score(ref(g), ref(eng.prof), ref(pool1.struct), ref(sl), 1, P(z), None, None, 2048, P(z), None) == E_ARG
    assert L.b2c_dense_score(ref(g), ref(eng.prof), ref(pool1.struct), ref(sl), 1, P(z), P(z), None, 4, P(z), None) == E_ARG and "np_max" in msg()
    g3 = Geom(14, 599, 3, 2, 1024, 72, 7.1e-5)
    assert L.b2c_dense_score(ref(g3), ref(eng.prof), ref(pool1.struct), ref(sl), 1, P(z), P(z), None, 2048, P(z), None) == E_UNSUPPORTED
    assert L.b2c_dense_score(ref(g), ref(eng.prof), ref(pool1.struct), ref(sl), 0, P(z), P(z), None, 2048, P(z), None) == 0
    # geometry outside the compiled limits: even bin count, too many symbols, too many antennas
    for bad in (Geom(14, 600, 2, 2, 1024, 72, 7.1e-5), Geom(17, 599, 2, 2, 1024, 72, 7.1e-5), Geom(14, 599, 9, 2, 1024, 72, 7.1e-5)):
        assert L.b2c_tap_gains(ref(bad), ref(eng.prof), ref(eng._slots(1, 0, 10.0, 10.0, 0, 0, 1)[0]), None, 1, P(z), P(z), None) == E_UNSUPPORTED
        assert msg()
    bad_fft = Geom(14, 599, 1, 1, 2048, 72, 7.1e-5)
    assert L.b2c_ofdm_modulate(ref(bad_fft), P(z), P(z), 1, None) == E_UNSUPPORTED and "1024" in msg()
    # row pitch on an entry point that takes contiguous rows only
    pitched = Geom(14, 599, 2, 2, 1024, 72, 7.1e-5, 600)
    assert L.b2c_tap_gains(ref(pitched), ref(eng.prof), ref(eng._slots(1, 0, 10.0, 10.0, 0, 0, 1)[0]), None, 1, P(z), P(z), None) == E_UNSUPPORTED
    # empty inputs: status 0, nothing launched
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert L.b2c_equalize(ref(g), 0, P(z), P(z), P(z), 1e-8, 0, s) == 0
    assert L.b2c_ofdm_modulate(ref(Geom(14, 599, 1, 1, 1024, 72, 7.1e-5)), P(z), P(z), 0, s) == 0
    assert L.b2c_mmse_dense(P(z), 4, P(z), P(bits), 0, 4, s) == 0
    assert L.b2c_count_bit_errors(P(bits), P(bits), 0, P(bits), s) == 0
    assert L.b2c_qam_modulate(P(bits), 0, 4, P(z), s) == 0
    torch.cuda.synchronize()
    assert eng.qam_modulate(torch.zeros((0,), dtype=torch.uint8, device=dev)).numel() == 0
    assert eng.count_bit_errors(bits[:0], bits[:0]).item() == 0
