"""GPU tests of the reference-shaped Python surface (channel_simulator, baseline_estimators,
dataset_generator, utils): same calls as the reference's own test scripts
(test_phase1_transmission.py, test_phase2_ls.py, test_phase2_mmse.py), but with numeric parity
against the golden vectors because the shims consume numpy.random in the reference's order."""
import os

import numpy as np
import pytest

from conftest import OFDM_CFG, SLOT_CASES, full_config, load_golden, relerr
from oracle import chanest_oracle as orc

pytestmark = pytest.mark.gpu
RTOL = 1e-4
SEEDS = {"slot_siso_epa": 101, "slot_2x2_eva": 202, "slot_4x4_etu": 303, "slot_2x2_etu_5pct": 404, "slot_2x1_epa_1pct": 505}


@pytest.mark.parametrize("name", SLOT_CASES)
def test_simulate_transmission_is_a_seeded_drop_in(name):
    """np.random.seed(s); simulate_transmission(...) returns the reference's arrays."""
    import channel_simulator as cs
    g = load_golden(name)
    ntx, nrx = int(g["ntx"]), int(g["nrx"])
    np.random.seed(SEEDS[name])
    sim = cs.simulate_transmission(full_config(ntx, nrx), channel_type=str(g["model"]), doppler_hz=float(g["doppler_hz"]),
                                   snr_db=float(g["snr_db"]), pilot_density=float(g["density"]))
    assert set(sim) == {"tx_symbols", "rx_symbols", "channel", "pilot_pattern", "pilot_symbols", "ofdm_config",
                        "mimo_config", "snr_db"}
    pp = sim["pilot_pattern"]
    assert np.array_equal(pp.pilot_indices, g["pilot_indices"]) and np.array_equal(pp.pilot_mask, g["pilot_mask"])
    assert np.array_equal(np.ravel_multi_index(pp.pilot_positions, (14, 599)), g["pilot_indices"])
    assert sim["channel"].shape == (14, nrx, ntx, 599) and sim["channel"].dtype == np.complex128
    assert sim["rx_symbols"].shape == (14, nrx, 599) and sim["tx_symbols"].shape == (14, ntx, 599)
    assert relerr(sim["channel"], g["channel"]) < RTOL
    assert relerr(sim["rx_symbols"], g["rx_symbols"]) < RTOL
    assert relerr(sim["pilot_symbols"], g["pilot_symbols"]) < RTOL
    for t in range(ntx):
        assert relerr(sim["tx_symbols"][:, t], g["tx_grid"]) < RTOL
    assert abs(np.mean(pp.pilot_mask) - float(g["density"])) < 0.05       # test_phase1_transmission.py:92-101


@pytest.mark.parametrize("name", ["slot_siso_epa", "slot_2x2_eva", "slot_2x2_etu_5pct"])
def test_ls_and_mmse_estimators(name):
    """The calls of test_phase2_ls.py:67-98 / test_phase2_mmse.py on the reference's rx grids."""
    import baseline_estimators as be
    g = load_golden(name)
    ntx, nrx = int(g["ntx"]), int(g["nrx"])
    rx4d = np.repeat(g["rx_symbols"].reshape(14, nrx, 1, 599), ntx, axis=2)
    pos = np.unravel_index(g["pilot_indices"], (14, 599))
    H_ls = be.LSEstimator(interpolation_method='linear').estimate(rx4d, g["pilot_symbols"], g["pilot_mask"], pos)
    est = be.MMSEEstimator(estimate_statistics=True)
    H_mm = est.estimate(rx4d, g["pilot_symbols"], g["pilot_mask"], pos, snr_db=float(g["snr_db"]))
    assert H_ls.shape == rx4d.shape and H_ls.dtype == np.complex128
    assert abs(est.noise_variance - 10 ** (-float(g["snr_db"]) / 10)) < 1e-12      # state mutation (:175)
    for t in range(ntx):
        assert relerr(H_ls[:, :, t], g["H_ls_tx0"]) < RTOL
        assert relerr(H_mm[:, :, t], g["H_mmse_tx0"]) < RTOL
    if f"H_ls_nearest_tx0" in g:
        H_n = be.LSEstimator('nearest').estimate(rx4d[:, :, :1], g["pilot_symbols"], g["pilot_mask"], pos)
        assert relerr(H_n[:, :, 0], g["H_ls_nearest_tx0"]) < RTOL
    m = be.evaluate_estimator(g["channel"], H_ls)
    assert set(m) == {"mse", "nmse", "nmse_db"}
    assert abs(m["nmse_db"] - g["metrics_ls"][2]) < 0.01 and abs(m["mse"] / g["metrics_ls"][0] - 1) < 2e-3
    # independent (rx, tx) slices: a tx slice that differs from its replica is estimated on its own
    rx_mod = rx4d.copy()
    rx_mod[:, 0, -1] *= 2.0
    H2 = be.LSEstimator().estimate(rx_mod, g["pilot_symbols"], g["pilot_mask"], pos)
    assert relerr(H2[:, 0, -1], 2.0 * g["H_ls_tx0"][:, 0]) < RTOL
    if name in load_golden("ls_cubic"):       # Clough-Tocher as a dense map on the tensor cores (test_phase2_ls.py:28)
        H_c = be.LSEstimator('cubic').estimate(rx4d, g["pilot_symbols"], g["pilot_mask"], pos)
        cub = load_golden("ls_cubic")[name]
        for t in range(ntx):
            assert relerr(H_c[:, :, t], cub) < RTOL
        assert np.array_equal(H_c[:, :, 0] == 0, cub == 0)
        grid = be.LSEstimator('cubic').interpolate_channel(
            orc.ls_at_pilots(g["rx_symbols"][:, 0], g["pilot_symbols"], g["pilot_mask"]), pos, (14, 599))
        assert relerr(grid, cub[:, 0]) < RTOL
    with pytest.raises(ValueError):
        be.LSEstimator('quintic').estimate(rx4d, g["pilot_symbols"], g["pilot_mask"], pos)
    with pytest.raises(ValueError):
        be.equalize_channel(g["rx_symbols"], H_ls, method='bogus')


def test_estimator_helper_methods():
    import baseline_estimators as be
    g = load_golden("slot_2x2_eva")
    pos = np.unravel_index(g["pilot_indices"], (14, 599))
    ls = be.LSEstimator()
    h_p = ls.estimate_at_pilots(g["rx_symbols"][:, 1], g["pilot_symbols"], g["pilot_mask"])
    want = orc.ls_at_pilots(g["rx_symbols"][:, 1], g["pilot_symbols"], g["pilot_mask"])
    assert relerr(h_p, want) < RTOL
    grid = ls.interpolate_channel(want, pos, (14, 599))
    assert relerr(grid, g["H_ls_tx0"][:, 1]) < RTOL
    mm = be.MMSEEstimator()
    h_m = mm.estimate_at_pilots(g["rx_symbols"][:, 1][g["pilot_mask"]], g["pilot_symbols"],
                                np.ones_like(g["pilot_symbols"], dtype=bool), float(g["snr_db"]))
    assert relerr(h_m, orc.mmse_at_pilots(want, float(g["snr_db"]))) < RTOL
    assert relerr(mm.interpolate_channel(orc.mmse_at_pilots(want, float(g["snr_db"])), pos, (14, 599)), g["H_mmse_tx0"][:, 1]) < RTOL


def test_known_covariance_mmse_estimator():
    import baseline_estimators as be
    g = load_golden("mmse_dense_2x2")
    pos = np.unravel_index(g["pilot_indices"], (14, 599))
    ds = pos[0][:, None] - pos[0][None, :]
    dk = pos[1][:, None] - pos[1][None, :]
    R = 0.4 * np.exp(-np.abs(ds) / 20.0 - np.abs(dk) / 60.0) * np.exp(1j * 2 * np.pi * dk * 3 / 1024)
    rx4d = np.repeat(g["rx_symbols"].reshape(14, 2, 1, 599), 2, axis=2)
    est = be.MMSEEstimator(channel_covariance=R, estimate_statistics=False)
    H = est.estimate(rx4d, g["pilot_symbols"], g["pilot_mask"], pos, snr_db=float(g["snr_db"]))
    for t in range(2):
        assert relerr(H[:, :, t], g["H_mmse_tx0"]) < RTOL


def test_channel_model_and_ofdm_classes():
    """test_phase1_channels.py:60-88 shape / finiteness checks + numeric parity on recorded draws."""
    import channel_simulator as cs
    g = load_golden("tdl_standalone")
    for model in ("EPA", "EVA", "ETU"):
        fd, ntx, nrx, ns = (g[f"{model}_meta"][0], *(int(v) for v in g[f"{model}_meta"][1:]))
        cm = cs.ChannelModel(model, float(fd), 2.0e9, 15.36e6)
        assert np.array_equal(cm.delay_samples, g[f"{model}_delay_samples"]) and cm.num_paths == len(cm.delays)
        assert np.allclose(cm.powers_linear, g[f"{model}_powers_linear"], rtol=1e-15)
        np.random.seed(808)
        h = cm.generate_time_varying_channel(ns, ntx, nrx)
        assert h.shape == (ns, nrx, ntx, int(cm.delay_samples.max()) + 1) and np.isfinite(h).all() and np.abs(h).sum() > 0
        assert relerr(h[::97], g[f"{model}_h"]) < RTOL
    with pytest.raises(KeyError):
        cs.ChannelModel("XYZ", 10, 2e9, 15.36e6)
    o = load_golden("ofdm_modem")
    sys_ = cs.OFDMSystem(cs.OFDMConfig())
    assert np.array_equal(sys_.used_indices, o["used_indices"]) and sys_.dc_idx == 512 and sys_.sampling_rate == 15.36e6
    assert relerr(sys_.modulate(o["symbols"]), o["modulated"]) < RTOL
    assert relerr(sys_.demodulate(o["signal"]), o["demodulated"]) < RTOL


def test_mimo_channel_methods():
    """generate_channel_frequency_response / apply_channel called on their own, seeded like the reference."""
    import channel_simulator as cs
    g = load_golden("slot_2x2_eva")
    ofdm, mimo = cs.OFDMConfig(), cs.MIMOConfig(2, 2)
    cm = cs.ChannelModel("EVA", 50.0, 2e9, 15.36e6)
    ch = cs.MIMOChannel(ofdm, mimo, cm)
    np.random.seed(1)
    ju = np.random.rand(9, 2, 2, 2, 20)
    nz = np.random.randn(2, 14, 2, 599)
    np.random.seed(1)
    H = ch.generate_channel_frequency_response(14)
    assert relerr(H, orc.channel_frequency_response("EVA", 50.0, OFDM_CFG, ju, 2, 2)) < RTOL
    tx = np.repeat(g["tx_grid"][:, None, :], 2, axis=1)
    rx = ch.apply_channel(tx, H, 12.0)
    assert relerr(rx, orc.apply_channel(tx, H, 12.0, nz[0], nz[1])) < RTOL
    pp = cs.PilotPattern(599, 14, 0.1)
    grid = pp.insert_pilots(np.arange(8386 - 838) + 0j, np.full(838, -1 + 0j))
    assert np.array_equal(pp.extract_pilots(grid), np.full(838, -1 + 0j)) and pp.get_pilot_positions() is pp.pilot_positions


def test_dataset_generator_numpy_mode_follows_reference_draw_order(tmp_path):
    import dataset_generator as dg
    from utils import default_config, set_seed
    cfg = default_config(2, 2)
    set_seed(123)
    gen = dg.ChannelEstimationDataset(cfg, batch_size=4)
    data = gen.generate_dataset(5, split='val')
    assert len(data) == 5
    # replay the same global-RNG stream through the oracle (src/dataset_generator.py:112-121 order)
    np.random.seed(123)
    for s in data:
        ch = np.random.choice(cfg["channel"]["models"])
        fd = np.random.choice(cfg["channel"]["doppler_hz"])
        snr = np.random.choice(cfg["simulation"]["snr_range"])
        dens = np.random.choice(cfg["pilots"]["density"])
        assert (s["channel_type"], s["doppler_hz"], s["snr_db"], s["pilot_density"]) == (ch, fd, snr, dens)
        perm = np.arange(8386)
        np.random.shuffle(perm)
        n_p = int(8386 * dens)
        draws = {"perm": perm, "pilot_phase": np.random.uniform(0, 2 * np.pi, n_p),
                 "data_phase": np.random.uniform(0, 2 * np.pi, 8386 - n_p),
                 "jakes_u": np.random.rand(len(orc.TDL_NS[ch]), 2, 2, 2, 20)}
        z = np.random.randn(2, 14, 2, 599)
        draws["noise_re"], draws["noise_im"] = z[0], z[1]
        ref = orc.simulate(OFDM_CFG, 2, 2, ch, float(fd), float(snr), float(dens), draws)
        rx4d = np.repeat(ref["rx_symbols"][:, :, None, :], 2, axis=2)
        H_ls = orc.ls_estimate(rx4d, ref["pilot_symbols"], ref["pilot_mask"], ref["pilot_positions"])
        assert np.array_equal(s["pilot_mask"], ref["pilot_mask"])
        assert relerr(s["H_true"], ref["channel"]) < RTOL and relerr(s["rx_symbols"], ref["rx_symbols"]) < RTOL
        assert relerr(s["tx_symbols"], ref["tx_symbols"]) < RTOL and relerr(s["H_ls"], H_ls) < RTOL
    f = tmp_path / "val.npz"
    gen.save_dataset(data, str(f), format='npz')
    z = np.load(f)
    assert set(z.files) == {"rx_symbols", "tx_symbols", "H_ls", "H_true", "pilot_mask", "snr_db", "channel_type",
                            "doppler_hz", "pilot_density"}
    assert z["H_true"].shape == (5, 14, 2, 2, 599) and z["rx_symbols"].shape == (5, 14, 2, 599)   # verify_phase3_datasets.py:68-74
    with pytest.raises(ValueError):
        gen.save_dataset(data, str(f), format='csv')


def test_dataset_generator_philox_mode_and_sharding():
    """Throughput mode: per-sample arrays and per-SNR statistics do not depend on how the global
    sample range is split across ranks (emulated here as two half-range runs on one GPU)."""
    import torch
    import dataset_generator as dg
    from utils import default_config
    cfg = default_config(2, 2)
    cfg["pilots"]["density"] = [0.05, 0.10]
    N = 96
    whole = dg.sharded_statistics(cfg, N, 0, 1, batch=40, seed=42)
    halves = [dg.sharded_statistics(cfg, N, r, 2, batch=17, seed=42) for r in range(2)]
    merged = halves[0] + halves[1]
    assert torch.equal(whole[:, 0], merged[:, 0]) and whole[:, 0].sum().item() == N
    assert torch.allclose(whole, merged, rtol=1e-12, atol=0)
    rows = dg.summarize_bins(whole)
    assert len(rows) == 8 and all(np.isfinite(r["nmse_ls_db"]) for r in rows if r["count"])
    ds = dg.ChannelEstimationDataset(cfg, rng='philox', seed=42)
    a, pa = ds.generate_batch(8, slot0=10)
    b, pb = ds.generate_batch(3, slot0=13)
    assert torch.equal(a["H_ls"][3:6], b["H_ls"]) and np.array_equal(pa["snr"][3:6], pb["snr"])
    lst = ds.generate_dataset(6)
    assert len(lst) == 6 and lst[0]["H_true"].shape == (14, 2, 2, 599) and lst[0]["pilot_mask"].dtype == bool
    # arrays handed to on_batch by the sharded loop (row-padded layout inside) are the per-sample arrays
    seen = {}
    with_arrays = dg.sharded_statistics(cfg, 16, batch=8, seed=42, want_arrays=("H_true", "rx", "tx", "H_ls", "H_mmse"),
                                        on_batch=lambda pos, res, par: seen.update({pos: (res["H_ls"].clone(), res["rx"].clone())}))
    full, _ = ds.generate_batch(16, slot0=0, want=("H_true", "rx", "tx", "H_ls", "H_mmse", "stats"))
    assert sorted(seen) == [0, 8] and with_arrays[:, 0].sum().item() == 16
    for pos, (hl, rx) in seen.items():
        assert torch.equal(hl, full["H_ls"][pos:pos + 8]) and torch.equal(rx, full["rx"][pos:pos + 8])
    # training features straight from the GPU-resident slots == prepare_ml_inputs on the same samples
    x, y, par = ds.generate_feature_batch(4, slot0=10)
    assert x.shape == (4, 14, 599, 5) and y.shape == (4, 14, 599, 2) and np.array_equal(par["snr"], pa["snr"][:4])
    pool = ds.pattern_pool()
    for i in range(4):
        smp = {"rx_symbols": a["rx"][i].cpu().numpy(), "H_ls": a["H_ls"][i].cpu().numpy(), "H_true": a["H_true"][i].cpu().numpy(),
               "pilot_mask": pool.mask(int(pa["pattern"][i]))}
        rx_ref, t_ref = orc.ml_inputs(smp["rx_symbols"], smp["H_ls"], smp["H_true"], smp["pilot_mask"])
        assert relerr(x[i].cpu().numpy(), rx_ref) < RTOL and relerr(y[i].cpu().numpy(), t_ref) < RTOL


def _replay_sample(cfg, ch, fd, snr, dens, ntx=2, nrx=2):
    """Oracle replay of one sample from the global numpy stream, reference draw order."""
    perm = np.arange(8386)
    np.random.shuffle(perm)
    n_p = int(8386 * dens)
    draws = {"perm": perm, "pilot_phase": np.random.uniform(0, 2 * np.pi, n_p),
             "data_phase": np.random.uniform(0, 2 * np.pi, 8386 - n_p),
             "jakes_u": np.random.rand(len(orc.TDL_NS[str(ch)]), ntx, nrx, 2, 20)}
    z = np.random.randn(2, 14, nrx, 599)
    draws["noise_re"], draws["noise_im"] = z[0], z[1]
    ref = orc.simulate(OFDM_CFG, ntx, nrx, str(ch), float(fd), float(snr), float(dens), draws)
    rx4d = np.repeat(ref["rx_symbols"][:, :, None, :], ntx, axis=2)
    ref["H_ls"] = orc.ls_estimate(rx4d, ref["pilot_symbols"], ref["pilot_mask"], ref["pilot_positions"])
    return ref


def test_phase3_generator_twin_matches_reference_order_and_dtypes(tmp_path):
    """run_phase3_dataset_generation.DatasetGenerator: set_seed, probe sample, 4 x choice + draws per
    sample (reference :98-176); stacked complex64 / float32 / U3 arrays (:135-143)."""
    import run_phase3_dataset_generation as p3
    gen = p3.DatasetGenerator(batch_size=3)
    data = gen.generate_dataset(4, split='val')
    assert data["rx_symbols"].shape == (4, 14, 2, 599) and data["H_true"].shape == (4, 14, 2, 2, 599)
    assert data["pilot_mask"].shape == (4, 14, 599)                      # verify_phase3_datasets.py:68-74
    assert data["H_ls"].dtype == np.complex64 and data["pilot_mask"].dtype == np.float32
    assert data["snr_db"].dtype == np.float32 and data["channel_type"].dtype == np.dtype('U3')
    np.random.seed(123)                                                  # 'val' split seed (:98-101)
    _replay_sample(None, 'EPA', 50.0, 10.0, 0.1)                         # the shape probe (:122)
    for i in range(4):
        ch = np.random.choice(p3.CHANNEL_TYPES)
        fd = float(np.random.choice(p3.DOPPLER_VALUES))
        snr = float(np.random.choice(p3.SNR_VALUES))
        dens = float(np.random.choice(p3.PILOT_DENSITIES))
        ref = _replay_sample(None, ch, fd, snr, dens)
        assert data["channel_type"][i] == ch and data["snr_db"][i] == snr and data["doppler_hz"][i] == fd
        assert np.array_equal(data["pilot_mask"][i] > 0, ref["pilot_mask"])
        assert relerr(data["H_true"][i], ref["channel"]) < RTOL and relerr(data["H_ls"][i], ref["H_ls"]) < RTOL
        assert relerr(data["rx_symbols"][i], ref["rx_symbols"]) < RTOL
    f = tmp_path / "val.npz"
    gen.save_dataset(data, str(f))
    assert set(np.load(f).files) == set(data)


def test_robust_generator_resume_is_exact_in_philox_mode(tmp_path):
    """Chunk files + JSON checkpoint + resume (run_phase3_robust.py:95-301).  With Philox keyed by the
    global sample index an interrupted-and-resumed run equals an uninterrupted one bit for bit."""
    import run_phase3_robust as rb
    a = rb.RobustDatasetGenerator(output_dir=str(tmp_path / "a"), rng='philox')
    whole = a.generate_dataset_chunked(10, 'train', chunk_size=4)
    assert whole["H_true"].shape == (10, 14, 2, 2, 599) and (tmp_path / "a" / "train.npz").exists()
    assert not list((tmp_path / "a" / "checkpoints").glob("*"))            # chunks + checkpoint cleaned up
    b = rb.RobustDatasetGenerator(output_dir=str(tmp_path / "b"), rng='philox')
    assert b.generate_dataset_chunked(10, 'train', chunk_size=4, stop_after_chunks=1) == {}
    ck = b.load_checkpoint('train')
    assert ck["completed"] == 4 and ck["chunk_id"] == 1 and ck["total"] == 10
    b2 = rb.RobustDatasetGenerator(output_dir=str(tmp_path / "b"), rng='philox')
    resumed = b2.generate_dataset_chunked(10, 'train', chunk_size=4, resume=True)
    for k in whole:
        assert np.array_equal(whole[k], resumed[k]), k
    # numpy mode keeps the reference's (non bit-reproducible) resume rule but the same file protocol
    c = rb.RobustDatasetGenerator(output_dir=str(tmp_path / "c"), rng='numpy', batch_size=2)
    out = c.generate_dataset_chunked(3, 'test', chunk_size=2)
    assert out["rx_symbols"].shape == (3, 14, 2, 599) and out["channel_type"].dtype == np.dtype('U3')


def test_resumed_val_split_in_a_fresh_generator_matches_uninterrupted_run(tmp_path):
    """Philox mode: the pattern pool follows the split seed, so 'val' generated after 'train' by one object, 'val'
    generated alone, and 'val' interrupted and resumed by a fresh object are all the same arrays."""
    import run_phase3_robust as rb
    a = rb.RobustDatasetGenerator(output_dir=str(tmp_path / "a"), rng='philox')
    a.generate_dataset_chunked(5, 'train', chunk_size=4)
    val_after_train = a.generate_dataset_chunked(9, 'val', chunk_size=4)
    b = rb.RobustDatasetGenerator(output_dir=str(tmp_path / "b"), rng='philox')
    assert b.generate_dataset_chunked(9, 'val', chunk_size=4, stop_after_chunks=1) == {}
    b2 = rb.RobustDatasetGenerator(output_dir=str(tmp_path / "b"), rng='philox')      # "fresh process": nothing generated before
    resumed = b2.generate_dataset_chunked(9, 'val', chunk_size=4, resume=True)
    for k in val_after_train:
        assert np.array_equal(val_after_train[k], resumed[k]), k


def test_reference_test_scripts_run_against_the_drop_in(tmp_path):
    """The reference's OWN test scripts for this path (staged unmodified under oracle/_ref by oracle/stage_ref.py) run
    against the drop-in through the `src.` namespace: `from src.channel_simulator import simulate_transmission` etc.
    resolve to the B200 modules.  The scripts print a verdict and exit 0 when their checks pass."""
    import shutil
    import subprocess
    import sys
    from conftest import PKG, ROOT
    ref = os.path.join(ROOT, "oracle", "_ref")
    scripts = ["test_phase1_transmission.py", "test_phase2_ls.py", "test_phase2_mmse.py"]
    if not all(os.path.exists(os.path.join(ref, s)) for s in scripts):
        pytest.skip("oracle/_ref not staged (python oracle/stage_ref.py needs /root/reference)")
    os.makedirs(tmp_path / "configs")
    shutil.copy(os.path.join(ref, "configs", "experiment_config.yaml"), tmp_path / "configs" / "experiment_config.yaml")
    env = dict(os.environ, PYTHONPATH=PKG, MPLBACKEND="Agg")
    for s in scripts:
        shutil.copy(os.path.join(ref, s), tmp_path / s)          # run from a scratch cwd: the scripts may write result files
        r = subprocess.run([sys.executable, s], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (s, r.stdout[-2000:], r.stderr[-2000:])
        assert "Traceback" not in r.stderr, (s, r.stderr[-2000:])
        # the scripts catch their own exceptions: the verdict is in what they print
        assert "COMPLETED SUCCESSFULLY" in r.stdout and "TESTS PASSED" in r.stdout, (s, r.stdout[-3000:])
        assert "\u274c" not in r.stdout and "FAILED" not in r.stdout, (s, r.stdout[-3000:])
        probe = subprocess.run([sys.executable, "-c", "import src.channel_simulator as m; print(m.__file__)"], cwd=tmp_path, env=env,
                               capture_output=True, text=True, timeout=300)
        assert PKG in probe.stdout, probe.stdout


def test_sharded_statistics_with_the_dense_wiener_estimator():
    """sharded_statistics(..., mmse="dense", wiener=bank): the per-SNR MMSE columns of the bins come from the
    known-covariance estimator; sharding over two emulated ranks reproduces the single-rank bins."""
    import torch
    import dataset_generator as dg
    from engine import WienerBank
    cfg = full_config(2, 2)
    cfg.update({"channel": {"models": ["EVA"], "doppler_hz": [50], "carrier_freq": 2.0e9}, "pilots": {"density": [0.02]},
                "simulation": {"snr_range": [0, 10, 20]}})
    ds = dg.ChannelEstimationDataset(cfg, rng='philox', seed=9)
    pool = ds.pattern_pool()
    idx = pool.pilot_indices[0]
    ps, pk = idx // 599, idx % 599
    R = 0.4 * np.exp(-np.abs(ps[:, None] - ps[None, :]) / 20.0 - np.abs(pk[:, None] - pk[None, :]) / 60.0) * np.exp(1j * 2 * np.pi * (pk[:, None] - pk[None, :]) * 3 / 1024)
    bank = WienerBank(ds.engine, pool, {0: R}, [0, 10, 20])
    one = dg.sharded_statistics(cfg, 48, 0, 1, batch=20, seed=9, dataset=ds, want_arrays=("H_true", "H_ls", "H_mmse"), mmse="dense", wiener=bank)
    two = sum(dg.sharded_statistics(cfg, 48, r, 2, batch=20, seed=9, dataset=ds, want_arrays=("H_true", "H_ls", "H_mmse"), mmse="dense", wiener=bank)
              for r in range(2))
    base = dg.sharded_statistics(cfg, 48, 0, 1, batch=20, seed=9, dataset=ds)
    torch.cuda.synchronize()
    assert torch.allclose(one, two, rtol=1e-12) and one[:, 0].sum().item() == 48
    assert torch.allclose(one[:, [1, 3]], base[:, [1, 3]], rtol=1e-5)          # the LS columns do not depend on the MMSE mode
    assert not torch.allclose(one[:, 4], base[:, 4], rtol=1e-3)                # the MMSE columns do


def test_pilot_density_sweep():
    """run_phase8 PilotOptimizer.analyze_pilot_density: per-(density, SNR) pair-(0,0) NMSE mean/std/dB."""
    import torch
    import run_phase8_pilot_optimization as p8
    opt = p8.PilotOptimizer(rng='philox', seed=7)
    dens, snrs, n = [0.02, 0.10], [5, 15], 6
    res = opt.analyze_pilot_density(dens, snrs, num_samples=n)
    assert res["pilot_densities"] == dens and res["snr_values"] == snrs and set(res["methods"]) == {"LS", "MMSE"}
    # recompute from the full arrays of the same (Philox-keyed) slots
    from dataset_generator import ChannelEstimationDataset
    ds = ChannelEstimationDataset(opt.config, rng='philox', seed=7, lists=(['EVA'], [50.0], snrs, dens))
    eng, pool = ds.engine, ds.pattern_pool()
    cell = np.arange(len(dens) * len(snrs) * n) // n
    out = eng.run(len(cell), 0, 50.0, np.asarray(snrs, np.float32)[cell % 2], (cell // 2).astype(np.int32), pool, slot0=0, seed=7)
    H, Hl = out["H_true"].cpu().numpy().astype(np.complex128), out["H_ls"].cpu().numpy().astype(np.complex128)
    for di, d in enumerate(dens):
        for si, s in enumerate(snrs):
            vals = [orc.nmse_pair00(Hl[b], H[b]) for b in np.flatnonzero(cell == di * 2 + si)]
            got = res["methods"]["LS"][s][d]
            assert abs(got["nmse_mean"] / np.mean(vals) - 1) < 1e-4 and abs(got["nmse_std"] - np.std(vals)) < 1e-3 * np.mean(vals)
            assert abs(got["nmse_db"] - 10 * np.log10(np.mean(vals) + 1e-12)) < 0.01
    # denser pilots estimate better at high SNR
    assert res["methods"]["LS"][15][0.10]["nmse_mean"] < res["methods"]["LS"][15][0.02]["nmse_mean"]
    # numpy mode: reference loop order (density, snr, sample), one pattern per sample
    np.random.seed(5)
    r2 = p8.PilotOptimizer(rng='numpy').analyze_pilot_density([0.05], [10], num_samples=2)
    np.random.seed(5)
    vals = []
    for _ in range(2):
        ref = _replay_sample(None, 'EVA', 50.0, 10, 0.05)
        vals.append(orc.nmse_pair00(ref["H_ls"], ref["channel"]))
    assert abs(r2["methods"]["LS"][10][0.05]["nmse_mean"] / np.mean(vals) - 1) < 1e-4
    s = p8.PilotOptimizer().generate_test_sample(0.1, snr_db=12.0)
    assert set(s) == {"rx_symbols", "H_ls", "H_true", "pilot_mask", "snr_db"}
    assert abs(p8.compute_nmse(s["H_ls"][:, 0, 0], s["H_true"][:, 0, 0]) / orc.nmse_pair00(s["H_ls"], s["H_true"]) - 1) < 1e-4


# ---- "next" rows -------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["2x2", "4x4", "2x4", "3x3", "rank1"])
def test_equalize_channel_drop_in(tag):
    """equalize_channel (src/baseline_estimators.py:273-312) against the reference's outputs."""
    import baseline_estimators as be
    g = load_golden("link_level")
    for meth in ("zf", "mmse"):
        x = be.equalize_channel(g[f"eq_{tag}_y"], g[f"eq_{tag}_H"], meth)
        assert x.dtype == np.complex128 and x.shape == g[f"eq_{tag}_{meth}"].shape
        assert relerr(x, g[f"eq_{tag}_{meth}"]) < (1e-5 if (tag == "rank1" and meth == "zf") else 1e-9)
    with pytest.raises(ValueError):
        be.equalize_channel(g["eq_2x2_y"], g["eq_2x2_H"], "mrc")


@pytest.mark.parametrize("M", [4, 16])
def test_qam_and_ber_drop_in(M):
    import utils as u
    g = load_golden("link_level")
    bits = g[f"qam{M}_bits"]
    sym = u.qam_modulation(bits, M)
    assert sym.dtype == np.complex128 and relerr(sym, g[f"qam{M}_symbols"]) < 1e-6
    assert np.array_equal(u.qam_demodulation(sym, M), bits)
    back = u.qam_demodulation(g[f"qam{M}_noisy"], M)
    assert np.array_equal(back, g[f"qam{M}_noisy_bits"])                      # bit-exact decisions
    assert u.calculate_ber(bits, back) == float(g[f"qam{M}_ber"])
    assert u.calculate_ber(bits, bits) == 0.0
    assert u.qam_modulation(bits[:len(bits) - 1], M).size == (len(bits) - 1) // int(np.log2(M))   # ragged tail dropped
    with pytest.raises(NotImplementedError):
        u.qam_modulation(bits, 64)
    with pytest.raises(NotImplementedError):
        u.qam_demodulation(sym, 256)


def test_prepare_ml_inputs_drop_in():
    import dataset_generator as dg
    g = load_golden("link_level")
    for i in range(g["ml_rx"].shape[0]):
        smp = {"rx_symbols": g["ml_rx"][i], "H_ls": g["ml_H_ls"][i], "H_true": g["ml_H_true"][i], "pilot_mask": g["ml_mask"][i]}
        for nz in (0, 1):
            x, t = dg.prepare_ml_inputs(smp, normalize=bool(nz))
            assert x.shape == (14, 96, 5) and t.shape == (14, 96, 2)
            assert np.array_equal(x[..., 4], g["ml_mask"][i].astype(float))           # mask channel bit-exact
            assert relerr(x, g[f"ml_inputs_{i}_{nz}"]) < RTOL and relerr(t, g[f"ml_targets_{i}_{nz}"]) < RTOL


def test_channel_dataset_features_and_phase5_baselines():
    """ChannelDataset normalisation + items (src/train.py:41-94) and the per-SNR LS / MMSE aggregation of
    run_phase5_evaluation.py:264-312 from device-resident arrays."""
    import torch
    import run_phase5_evaluation as p5
    from baseline_estimators import _engine
    from engine import PatternPool
    from _b2c import Geom
    g = load_golden("link_level")
    eng = _engine()
    rx, Hls, Htr, mask = g["ml_rx"], g["ml_H_ls"], g["ml_H_true"], g["ml_mask"]
    N, nsym, nrx, ntx, nsc = Htr.shape
    geom = Geom(nsym, nsc, ntx, nrx, 1024, 72, 0.0)
    dev = lambda a: torch.from_numpy(a.astype(np.complex64)).to(eng.device)
    mom = eng.pair00_moments(dev(rx), dev(Hls), dev(Htr), geom=geom)
    norm = eng.normalization_from_moments(mom, N * nsym * nsc)
    ref = g["ds_norm"]
    want = np.array([ref[0], 1 / (ref[1] + 1e-8), ref[2], 1 / (ref[3] + 1e-8), ref[4], 1 / (ref[5] + 1e-8)])
    # the means are ~1e-2 of the signal scale: compare them on that scale
    assert np.max(np.abs(norm[0::2] - want[0::2])) < 1e-5 and relerr(norm[1::2], want[1::2]) < 1e-5
    pool = PatternPool([np.flatnonzero(m.reshape(-1)) for m in mask], nsym, nsc, "nearest", eng.device)
    pid = np.arange(N, dtype=np.int32)
    for nz in (0, 1):
        x, t = eng.ml_features(dev(rx), dev(Hls), dev(Htr), pool, pid, "first", False, norm if nz else None, geom=geom)
        for i in range(N):
            assert relerr(x[i].cpu().numpy(), g[f"ds_inputs_{i}_{nz}"]) < RTOL
            assert relerr(t[i].cpu().numpy(), g[f"ds_targets_{i}_{nz}"]) < RTOL
    # compact H_ls layout ([B][nsym][nrx][nsc], tx = 0 only) gives the same features
    xc, _ = eng.ml_features(dev(rx), dev(Hls[:, :, :, 0]).contiguous(), dev(Htr), pool, pid, "first", False, norm, geom=geom)
    assert torch.equal(xc, x)
    # phase-5 metric helpers and the per-SNR sweep
    if "ber_approx" in g:
        for j, s in enumerate((-5.0, 10.0, 30.0)):
            assert abs(p5.compute_ber_approximation(Hls[0], Htr[0], s) / g["ber_approx"][j] - 1) < 1e-4
            assert abs(p5.compute_ber_approximation(Htr[0] * 1.01, Htr[0], s) / g["ber_approx_small_err"][j] - 1) < 1e-4
    assert abs(p5.compute_mae(Hls[0], Htr[0]) / np.mean(np.abs(Hls[0] - Htr[0])) - 1) < 1e-5
    assert abs(p5.compute_mse(Hls[1], Htr[1]) / np.mean(np.abs(Hls[1] - Htr[1]) ** 2) - 1) < 1e-5
    snr = np.array([0.0, 10.0, 0.0])
    sweep = p5.snr_sweep_baselines({"H_true": Htr, "H_ls": Hls, "snr_db": snr})
    vals, ls_db, mm_db = orc.snr_sweep_baselines(Htr, Hls, snr)
    assert sweep["snr_db"] == vals
    assert np.allclose(sweep["methods"]["LS"]["nmse_db"], ls_db, atol=1e-4)
    assert np.allclose(sweep["methods"]["MMSE"]["nmse_db"], mm_db, atol=1e-4)


def test_verify_phase3_datasets_drop_in(tmp_path):
    """verify_phase3_datasets.verify_dataset against the reference's own verdicts on the same tiny file
    (tests/golden/link_level.npz: clean -> 'shape_mismatch' (grid 14 x 96), planted NaN + Inf -> 'data_errors')."""
    import verify_phase3_datasets as v3
    g = load_golden("link_level")
    rx, Hls, Htr, mask = g["ml_rx"], g["ml_H_ls"], g["ml_H_true"], g["ml_mask"]
    meta = dict(snr_db=np.array([0.0, 10.0, 0.0], np.float32), channel_type=np.array(["EPA", "EVA", "EPA"]),
                doppler_hz=np.array([10.0, 50.0, 10.0], np.float32), pilot_density=np.array([0.1, 0.1, 0.05], np.float32))
    stacked = dict(rx_symbols=rx.astype(np.complex64), tx_symbols=rx.astype(np.complex64), H_ls=Hls.astype(np.complex64),
                   H_true=Htr.astype(np.complex64), pilot_mask=mask.astype(np.float32), **meta)
    f = str(tmp_path / "tiny.npz")
    np.savez(f, **stacked)
    np.random.seed(77)
    r = v3.verify_dataset(f, verbose=False)
    assert r["status"] == str(g["verify_status"][0]) == "shape_mismatch" and r["num_samples"] == 3
    assert abs(r["avg_ls_nmse_db"] - float(g["verify_ls_nmse_db"][0])) < 1e-3
    assert abs(r["avg_pilot_density"] - float(g["verify_pilot_density"][0])) < 1e-6      # the reference sums the float32 mask in float32
    assert r["snr_range"] + r["doppler_range"] == list(g["verify_ranges"])
    assert r["channel_types"] == ["EPA", "EVA"]
    bad = dict(stacked)
    bad["H_ls"] = stacked["H_ls"].copy()
    bad["H_ls"][1, 2, 0, 1, 5] = np.nan
    bad["rx_symbols"] = stacked["rx_symbols"].copy()
    bad["rx_symbols"][0, 0, 0, 0] = np.inf
    np.savez(f, **bad)
    r = v3.verify_dataset(f, verbose=False)
    assert r["status"] == str(g["verify_status"][1]) == "data_errors" and r["nan_count"] == 1 and r["inf_count"] == 1
    np.savez(f, **{k: v for k, v in stacked.items() if k != "H_true"})
    assert v3.verify_dataset(f, verbose=False) == {"status": "incomplete", "filepath": f, "missing_keys": ["H_true"]}
    assert v3.verify_dataset(str(tmp_path / "absent.npz"), verbose=False)["status"] == "failed"
    # a file written by the phase-3 twin passes at the default 2x2 geometry
    import run_phase3_dataset_generation as p3
    gen = p3.DatasetGenerator(batch_size=2)
    good = str(tmp_path / "val.npz")
    gen.save_dataset(gen.generate_dataset(2, split='val'), good)
    r = v3.verify_dataset(good, verbose=False)
    assert r["status"] == "success" and r["num_samples"] == 2 and np.isfinite(r["avg_ls_nmse_db"])
