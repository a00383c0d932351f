"""CPU-only checks of the product's host logic: table construction against the oracle, plan
construction against scipy.griddata semantics, Philox parameter choice, sharding, and the C-ABI
library's exported symbols (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

from conftest import OFDM_CFG, PKG, ROOT, SLOT_CASES, load_golden, relerr
from oracle import chanest_oracle as orc
from oracle import philox as opx

import _tables


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = opx.philox4x32_10(np.array(ctr, dtype=np.uint64), key)
        assert tuple(int(x) for x in got) == want


def test_philox_draw_statistics():
    u = opx.symbol_u(42, 7, 14, 599)
    assert u.min() > 0 and u.max() < 1 and abs(u.mean() - 0.5) < 0.01
    re, im = opx.noise(42, 7, 14, 4, 599)
    z = np.concatenate([re.ravel(), im.ravel()])
    assert abs(z.mean()) < 0.01 and abs(z.var() - 1) < 0.02 and abs(np.mean(z ** 4) - 3) < 0.15
    ju = opx.jakes_u(42, 7, 9, 4, 4)
    assert ju.shape == (9, 4, 4, 2, 20) and len(np.unique(ju)) > 0.99 * ju.size


def test_param_choice_matches_oracle_twin():
    import dataset_generator as dg
    sizes = (3, 4, 8, 2)
    got = dg.philox_param_choice(42, 1000, 64, sizes)
    for j in range(64):
        assert tuple(int(g[j]) for g in got) == opx.param_choice(42, 1000 + j, *sizes)
    for g, n in zip(dg.philox_param_choice(1, 0, 20000, sizes), sizes):
        assert np.bincount(g, minlength=n).min() > 0.8 * 20000 / n


@pytest.mark.parametrize("model", ["EPA", "EVA", "ETU"])
def test_profile_tables_match_oracle(model):
    used = _tables.used_subcarriers(1024, 600)
    assert np.array_equal(used, orc.used_bins(1024, 600))
    t = _tables.profile_tables([model], 15.36e6, 1024, used)
    prof = orc.tdl_profile(model, 15.36e6)
    d, own = orc.surviving_taps(prof["delay_samples"])
    nt = int(t["ntaps"][0])
    assert nt == len(d) and t["npaths"][0] == len(prof["delay_samples"])
    assert t["tap_delay"][0, :nt].tolist() == d.tolist() and t["tap_path"][0, :nt].tolist() == own.tolist()
    assert np.allclose(t["tap_amp"][0, :nt], np.sqrt(prof["powers_linear"][own] / 40), rtol=1e-6)
    # twiddle table reproduces fftshift(fft(h, 1024))[used] for a random CIR on the surviving taps
    rng = np.random.default_rng(0)
    h = np.zeros(int(d.max()) + 1, complex)
    g = rng.standard_normal(nt) + 1j * rng.standard_normal(nt)
    h[d] = g
    want = orc.cfr_from_cir(h[None, None, None, :], 1024, used)[0, 0, 0]
    got = (g[:, None] * t["tap_tw"][0, :nt].astype(complex)).sum(0)
    assert relerr(got, want) < 1e-6
    # quadratic form equals the bin-summed power
    c = t["tap_corr"][0, :nt, :nt].astype(complex)
    assert abs((g @ c @ g.conj()).real / np.sum(np.abs(want) ** 2) - 1) < 1e-5


@pytest.mark.parametrize("name", SLOT_CASES)
def test_plan_reproduces_reference_interpolation(name):
    """The 16-byte plan entries (fp32 weights) reproduce the reference's griddata output."""
    g = load_golden(name)
    plan = _tables.cached_plan(g["pilot_indices"], 14, 599, "linear")
    assert plan.dtype.itemsize == 16 and plan.shape == (14 * 599,)
    w0, w1 = plan["w0"].astype(np.float64), plan["w1"].astype(np.float64)
    inside = plan["flags"].astype(bool)
    for r in range(int(g["nrx"])):
        h_p = orc.ls_at_pilots(g["rx_symbols"][:, r], g["pilot_symbols"], g["pilot_mask"])
        val = w0 * h_p[plan["i0"]] + w1 * h_p[plan["i1"]] + (1 - w0 - w1) * h_p[plan["i2"]]
        val = np.where(inside, val, 0).reshape(14, 599)
        assert relerr(val, g["H_ls_tx0"][:, r]) < 1e-6
        assert np.array_equal(val == 0, g["H_ls_tx0"][:, r] == 0)


def test_finalized_plan_device_format():
    g = load_golden("slot_2x1_epa_1pct")
    plan = _tables.cached_plan(g["pilot_indices"], 14, 599, "linear")
    dev = _tables.finalize_plan(plan, 200)
    assert dev.shape == (8387,) and dev.dtype.itemsize == 16
    out = plan["flags"] == 0
    assert out.sum() > 0 and (dev["i0"][:-1][out] == 200).all() and (dev["i2"][:-1][out] == 200).all()
    assert (dev["w0"][:-1][out] == 1).all() and (dev["w1"][:-1][out] == 0).all() and dev["i1"][-1] == 200
    assert np.array_equal(dev[:-1][~out], plan[~out])


def test_nearest_plan_and_unknown_method():
    g = load_golden("slot_2x2_eva")
    plan = _tables.cached_plan(g["pilot_indices"], 14, 599, "nearest")
    pos = np.unravel_index(g["pilot_indices"], (14, 599))
    assert np.array_equal(plan["i0"], orc.nearest_plan(pos, 14, 599))
    with pytest.raises(ValueError):
        _tables.interpolation_plan(pos, 14, 599, "quintic")
    with pytest.raises(KeyError):
        _tables.path_tables("XYZ", 15.36e6)


@pytest.mark.parametrize("name", ["slot_siso_epa", "slot_2x1_epa_1pct"])
def test_cubic_interpolation_is_a_linear_map(name):
    """The dense matrix the GPU applies for interpolation_method='cubic' reproduces the reference's
    Clough-Tocher griddata output (fixture minted from the reference) -- and so does the oracle."""
    g, c = load_golden(name), load_golden("ls_cubic")
    W = _tables.cubic_matrix(g["pilot_indices"], 14, 599)
    assert W.shape == (14 * 599, len(g["pilot_indices"])) and W.dtype == np.float32
    pos = np.unravel_index(g["pilot_indices"], (14, 599))
    for r in range(int(g["nrx"])):
        h_p = orc.ls_at_pilots(g["rx_symbols"][:, r], g["pilot_symbols"], g["pilot_mask"])
        assert relerr((W.astype(np.float64) @ h_p).reshape(14, 599), c[name][:, r]) < 5e-6
    rx4d = g["rx_symbols"][:, :, None, :]
    H = orc.ls_estimate(rx4d, g["pilot_symbols"], g["pilot_mask"], pos, method="cubic")
    assert relerr(H[:, :, 0], c[name]) < 1e-12
    assert np.array_equal(c[name] == 0, (W != 0).any(axis=1).reshape(14, 599)[None].repeat(int(g["nrx"]), 0).transpose(1, 0, 2) == 0)


def test_shard_ranges_partition_the_samples():
    import dataset_generator as dg
    for total in (0, 1, 7, 1000, 100003):
        for ws in (1, 2, 3, 8):
            r = [dg.shard_range(total, k, ws) for k in range(ws)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(ws - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_summarize_bins_units():
    import dataset_generator as dg
    bins = np.zeros((2, 14))
    bins[0] = [4, 4 * 0.5, 4 * 0.25, 4 * 1.0, 4 * 0.5, 0, 0, 4 * 0.5, 4 * 2.0, 4 * 5.0, 0, 0, 4 * 0.125, 4 * 0.0625]
    rows = dg.summarize_bins(bins)
    assert rows[0]["count"] == 4 and abs(rows[0]["mse_ls"] - 0.5) < 1e-15
    assert rows[0]["ber_proxy_ls"] == 0.125 and rows[0]["ber_proxy_mmse"] == 0.0625
    assert abs(rows[0]["nmse_ls_db"]) < 1e-9 and abs(rows[0]["nmse00_ls_std"] - 1.0) < 1e-12
    assert rows[1]["count"] == 0


def test_library_exports_every_declared_symbol():
    """include/b2c.h is the contract: every function it declares is exported by libb2c.so."""
    import _b2c
    hdr = open(os.path.join(ROOT, "include", "b2c.h")).read()
    declared = set(re.findall(r"\b(b2c_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_b2c.EXPORTS), declared ^ set(_b2c.EXPORTS)
    if not os.path.exists(_b2c.LIB_PATH):
        import importlib.util
        spec = importlib.util.spec_from_file_location("b2c_build", os.path.join(PKG, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    L = _b2c.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.b2c_abi_version() == _b2c.ABI_VERSION == 3
    # argument validation runs before any CUDA call: safe without a GPU
    assert L.b2c_tap_gains(None, None, None, None, 1, None, None, None) == -1
    assert b"null argument" in L.b2c_last_error_string()
    g = _b2c.Geom(14, 599, 4, 4, 2048, 72, 7.1e-5)
    buf = (ctypes.c_float * 8)()
    assert L.b2c_ofdm_modulate(ctypes.byref(g), buf, buf, 1, None) == -3      # only fft 1024 is built


def test_product_fails_loudly_without_cuda():
    """No CPU fallback: constructing the engine without a GPU raises."""
    import torch
    import _b2c
    import engine
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from utils import default_config
    with pytest.raises(_b2c.B2CError):
        engine.SlotEngine(default_config())
    import channel_simulator as cs
    with pytest.raises(_b2c.B2CError):
        cs.simulate_transmission(default_config())


def test_product_never_imports_the_oracle():
    for fn in os.listdir(PKG):
        if fn.endswith(".py"):
            src = open(os.path.join(PKG, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn


def test_row_pitch_and_numa_helpers():
    """Host-side helpers of the padded-row layout and of the e2e leg (no GPU needed)."""
    import torch
    import _b2c
    import host_pipeline as hp
    buf = torch.zeros((3, 14, 4, 600), dtype=torch.complex64)
    view = buf[..., :599]
    assert _b2c.row_pitch(buf) == 600 and _b2c.row_pitch(view) == 600 and not view.is_contiguous()
    assert _b2c.row_pitch(torch.zeros((3, 14, 4, 599), dtype=torch.complex64)) == 599
    assert _b2c.row_pitch(torch.zeros((7,), dtype=torch.complex64)) == 7
    with pytest.raises(_b2c.B2CError):            # the product path takes CUDA tensors only
        _b2c.rows_ptr(view, 600)
    assert _b2c.rows_ptr(None, 600, optional=True) is None
    assert hp._cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11} and hp._cpulist("") == set()
    if not torch.cuda.is_available():
        assert hp.bind_to_gpu_numa_node(0) is None    # no device: nothing is bound, nothing raises


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the staged reference, or the oracle port when oracle/_ref is absent, on the host cores) prints one JSON line with the contract's keys,
    honours its wall budget, and non-zero ranks stay silent."""
    import json
    import subprocess
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1_siso_epa",
           "--steps", "50", "--warmup", "1", "--reference-budget", "6"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, check=True).stdout.strip().splitlines()
    line = json.loads(out[-1])
    assert line["impl"] == "reference" and line["unit"] == "slots/s" and line["higher_is_better"] is True
    assert 1 <= line["steps"] < 50 and "wall budget" in line["arm"]["note"]
    # `config` names the workload only, so that it is identical in both arms (the driver compares the strings)
    import bench
    assert line["config"] == {"workload": "c1_siso_epa: " + bench.WORKLOADS["c1_siso_epa"]["desc"]}
    from oracle import cpu_bench
    kind = "reference" if cpu_bench.reference_available() else "port"      # oracle/_ref staged (this container) or not
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == kind and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "slots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["vs_baseline"] is None
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    silent = subprocess.run(cmd, capture_output=True, text=True, timeout=120, check=True, env=env)
    assert silent.stdout.strip() == ""


def test_src_namespace_resolves_to_the_drop_ins():
    """The reference's scripts import `from src.channel_simulator import ...`; with the package directory on sys.path
    those lines bind the B200 drop-ins (same objects as the flat modules)."""
    import baseline_estimators as be
    import channel_simulator as cs
    import utils as ut
    import src.baseline_estimators as sbe
    import src.channel_simulator as scs
    import src.dataset_generator as sdg
    import src.utils as sut
    assert scs.simulate_transmission is cs.simulate_transmission and scs.PilotPattern is cs.PilotPattern
    assert sbe.LSEstimator is be.LSEstimator and sbe.MMSEEstimator is be.MMSEEstimator and sbe.evaluate_estimator is be.evaluate_estimator
    assert sut.calculate_nmse is ut.calculate_nmse and sut.linear2db is ut.linear2db
    assert hasattr(sdg, "ChannelEstimationDataset") and hasattr(sdg, "prepare_ml_inputs")


def test_load_config_only_falls_back_for_the_default_path(tmp_path, monkeypatch):
    """A mistyped --config must not silently run the packaged default (the reference raises FileNotFoundError)."""
    import utils as ut
    monkeypatch.chdir(tmp_path)                       # no configs/ here: the literal default falls back to the packaged copy
    assert ut.load_config()["ofdm"]["fft_size"] == 1024
    with pytest.raises(FileNotFoundError):
        ut.load_config("configs/experiment_confg.yaml")
    with pytest.raises(FileNotFoundError):
        ut.load_config(str(tmp_path / "nope.yaml"))


def test_reseeding_a_dataset_drops_its_pattern_pool():
    """The Philox pattern pool is a function of the seed: the script twins re-seed per split, and a split's pilot
    patterns must not depend on which split the same object generated first."""
    import dataset_generator as dg
    import utils as ut
    ds = dg.ChannelEstimationDataset(ut.default_config(), rng='philox', seed=42)
    ds._pool = object()
    ds.seed = 42
    assert ds._pool is not None                       # unchanged seed: pool kept
    ds.seed = 123
    assert ds._pool is None and ds.seed == 123


def test_plan_pool_builder_parallel_and_cached(tmp_path, monkeypatch):
    """plans_for: worker interpreters + on-disk cache give exactly the plans of the in-process builder."""
    monkeypatch.setenv("B2C_PLAN_CACHE", str(tmp_path / "plans"))
    rs = np.random.RandomState(3)
    pats = []
    for i in range(18):
        perm = np.arange(7 * 299)
        rs.shuffle(perm)
        pats.append(np.sort(perm[:60 + 5 * i]))
    _tables._PLAN_CACHE.clear()
    a = _tables.plans_for(pats, 7, 299, workers=2)                    # 18 patterns / 2 workers: the subprocess path
    assert len(list((tmp_path / "plans").glob("*.npy"))) == 18
    _tables._PLAN_CACHE.clear()
    b = _tables.plans_for(pats, 7, 299)                               # all from disk
    for i, p in enumerate(pats):
        ref = _tables.interpolation_plan(np.unravel_index(p, (7, 299)), 7, 299)
        assert np.array_equal(a[i], ref) and np.array_equal(b[i], ref)
