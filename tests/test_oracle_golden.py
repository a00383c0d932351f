"""Pin the CPU oracle against outputs of the real reference (tests/golden, minted by
oracle/make_golden.py).  Float tolerance 1e-12 (both sides float64); integer/index data bit-exact."""
import numpy as np
import pytest

from conftest import OFDM_CFG, SLOT_CASES, golden_draws, load_golden, relerr
from oracle import chanest_oracle as orc

TOL = 1e-12


@pytest.mark.parametrize("name", SLOT_CASES)
def test_simulate_matches_reference(name):
    g = load_golden(name)
    ntx, nrx = int(g["ntx"]), int(g["nrx"])
    sim = orc.simulate(OFDM_CFG, ntx, nrx, str(g["model"]), float(g["doppler_hz"]), float(g["snr_db"]),
                       float(g["density"]), golden_draws(g))
    assert np.array_equal(sim["pilot_indices"], g["pilot_indices"])
    assert np.array_equal(sim["pilot_mask"], g["pilot_mask"])
    assert relerr(sim["pilot_symbols"], g["pilot_symbols"]) < TOL
    for t in range(ntx):
        assert relerr(sim["tx_symbols"][:, t], g["tx_grid"]) < TOL
    assert relerr(sim["channel"], g["channel"]) < TOL
    assert relerr(sim["rx_symbols"], g["rx_symbols"]) < TOL


@pytest.mark.parametrize("name", SLOT_CASES)
def test_ls_mmse_match_reference(name):
    g = load_golden(name)
    ntx, nrx = int(g["ntx"]), int(g["nrx"])
    rx4d = np.repeat(g["rx_symbols"][:, :, None, :], ntx, axis=2)
    pos = np.unravel_index(g["pilot_indices"], g["pilot_mask"].shape)
    H_ls = orc.ls_estimate(rx4d, g["pilot_symbols"], g["pilot_mask"], pos)
    H_mm = orc.mmse_estimate(rx4d, g["pilot_symbols"], g["pilot_mask"], pos, float(g["snr_db"]))
    for t in range(ntx):
        assert relerr(H_ls[:, :, t], g["H_ls_tx0"]) < TOL
        assert relerr(H_mm[:, :, t], g["H_mmse_tx0"]) < TOL
    # exact zeros outside the pilots' convex hull (griddata fill_value=0.0)
    assert np.array_equal(H_ls[:, :, 0] == 0, g["H_ls_tx0"] == 0)
    m = orc.evaluate(g["channel"], H_ls)
    assert np.allclose([m["mse"], m["nmse"], m["nmse_db"]], g["metrics_ls"], rtol=1e-10, atol=1e-12)
    m = orc.evaluate(g["channel"], H_mm)
    assert np.allclose([m["mse"], m["nmse"], m["nmse_db"]], g["metrics_mmse"], rtol=1e-10, atol=1e-12)


def test_faithful_profile_equals_fast_profile():
    """The cost-faithful route (full time vector, griddata per pair, Np x Np inverse) gives the
    same arrays as the fast one -- on the smallest case so it stays quick."""
    g = load_golden("slot_2x1_epa_1pct")
    ntx, nrx = int(g["ntx"]), int(g["nrx"])
    a = orc.slot_pipeline(OFDM_CFG, ntx, nrx, "EPA", float(g["doppler_hz"]), float(g["snr_db"]),
                          float(g["density"]), golden_draws(g), faithful=True)
    assert relerr(a["channel"], g["channel"]) < TOL
    assert relerr(a["rx_symbols"], g["rx_symbols"]) < TOL
    assert relerr(a["H_ls"][:, :, 0], g["H_ls_tx0"]) < TOL
    assert relerr(a["H_mmse"][:, :, 1], g["H_mmse_tx0"]) < 1e-10   # explicit inverse of (P+s2) I


@pytest.mark.parametrize("name", ["slot_siso_epa", "slot_2x2_eva"])
def test_nearest_plan_matches_reference(name):
    g = load_golden(name)
    pos = np.unravel_index(g["pilot_indices"], g["pilot_mask"].shape)
    idx = orc.nearest_plan(pos, 14, 599)
    for r in range(int(g["nrx"])):
        h_p = orc.ls_at_pilots(g["rx_symbols"][:, r], g["pilot_symbols"], g["pilot_mask"])
        assert relerr(h_p[idx].reshape(14, 599), g["H_ls_nearest_tx0"][:, r]) < TOL


def test_dense_covariance_mmse_matches_reference():
    g = load_golden("mmse_dense_2x2")
    pos = np.unravel_index(g["pilot_indices"], g["pilot_mask"].shape)
    ds = pos[0][:, None] - pos[0][None, :]
    dk = pos[1][:, None] - pos[1][None, :]
    R = 0.4 * np.exp(-np.abs(ds) / 20.0 - np.abs(dk) / 60.0) * np.exp(1j * 2 * np.pi * dk * 3 / 1024)
    rx4d = np.repeat(g["rx_symbols"][:, :, None, :], 2, axis=2)
    H = orc.mmse_estimate(rx4d, g["pilot_symbols"], g["pilot_mask"], pos, float(g["snr_db"]), cov=R)
    assert relerr(H[:, :, 0], g["H_mmse_tx0"]) < 1e-10
    # W @ h route used by the GPU path: same numbers
    W = orc.wiener_matrix(R, float(g["snr_db"]))
    idx, w = orc.linear_plan(pos, 14, 599)
    for r in range(2):
        h = W @ orc.ls_at_pilots(g["rx_symbols"][:, r], g["pilot_symbols"], g["pilot_mask"])
        assert relerr(orc.plan_apply(idx, w, h, 14, 599), g["H_mmse_tx0"][:, r]) < 1e-10


def test_dense_covariance_mmse_4x4_matches_reference():
    """838-pilot Wiener filter on the 4x4 ETU fixture's slot (reference run: oracle/make_golden.py
    dense_from_slot_case): the oracle's W @ h_ls + plan route against the reference's estimate."""
    g, d = load_golden("slot_4x4_etu"), load_golden("mmse_dense_4x4_etu")
    pos = np.unravel_index(g["pilot_indices"], g["pilot_mask"].shape)
    ds = pos[0][:, None] - pos[0][None, :]
    dk = pos[1][:, None] - pos[1][None, :]
    R = 0.4 * np.exp(-np.abs(ds) / 20.0 - np.abs(dk) / 60.0) * np.exp(1j * 2 * np.pi * dk * 3 / 1024)
    W = orc.wiener_matrix(R, float(g["snr_db"]))
    idx, w = orc.linear_plan(pos, 14, 599)
    for r in range(4):
        h = W @ orc.ls_at_pilots(g["rx_symbols"][:, r], g["pilot_symbols"], g["pilot_mask"])
        assert relerr(orc.plan_apply(idx, w, h, 14, 599), d["H_mmse_tx0"][:, r]) < 1e-9
    H = np.repeat(d["H_mmse_tx0"][:, :, None, :], 4, axis=2)
    m = orc.evaluate(g["channel"], H)
    assert abs(m["nmse_db"] - d["metrics_mmse"][2]) < 1e-9


def test_ofdm_modem_matches_reference():
    g = load_golden("ofdm_modem")
    assert np.array_equal(orc.used_bins(1024, 600), g["used_indices"])
    assert relerr(orc.ofdm_modulate(g["symbols"], 1024, 72, 600), g["modulated"]) < TOL
    assert relerr(orc.ofdm_demodulate(g["signal"], 1024, 72, 600), g["demodulated"]) < TOL


@pytest.mark.parametrize("model", ["EPA", "EVA", "ETU"])
def test_tdl_standalone_matches_reference(model):
    g = load_golden("tdl_standalone")
    fd, ntx, nrx, ns = g[f"{model}_meta"]
    prof = orc.tdl_profile(model, 15.36e6)
    assert np.array_equal(prof["delay_samples"], g[f"{model}_delay_samples"])
    assert np.allclose(prof["powers_linear"], g[f"{model}_powers_linear"], rtol=1e-15)
    idx = np.arange(0, int(ns), 97)
    h = orc.jakes_cir(model, float(fd), 15.36e6, g[f"{model}_jakes_u"], idx, int(ntx), int(nrx))
    assert h.shape[1:] == tuple(g[f"{model}_shape"][1:])
    assert relerr(h, g[f"{model}_h"]) < TOL


def test_documented_integer_facts():
    """delay tables, used bins, pilot counts -- bit-exact integers."""
    assert orc.tdl_profile("EPA", 15.36e6)["delay_samples"].tolist() == [0, 0, 1, 1, 2, 3, 6]
    assert orc.tdl_profile("EVA", 15.36e6)["delay_samples"].tolist() == [0, 0, 2, 5, 6, 11, 17, 27, 39]
    assert orc.tdl_profile("ETU", 15.36e6)["delay_samples"].tolist() == [0, 1, 2, 3, 4, 8, 25, 35, 77]
    d, own = orc.surviving_taps(orc.tdl_profile("EPA", 15.36e6)["delay_samples"])
    assert d.tolist() == [0, 1, 2, 3, 6] and own.tolist() == [1, 3, 4, 5, 6]
    u = orc.used_bins(1024, 600)
    assert u.size == 599 and u[0] == 212 and u[299] == 511 and u[300] == 513 and u[-1] == 811
    assert [int(8386 * d) for d in (0.01, 0.02, 0.05, 0.10)] == [83, 167, 419, 838]
    with pytest.raises(KeyError):
        orc.tdl_profile("XYZ", 15.36e6)


# ---- "next" rows: equaliser, QAM, BER, ML features ------------------------------
EQ_CASES = ["2x2", "4x4", "2x4", "3x3", "rank1"]


@pytest.mark.parametrize("tag", EQ_CASES)
def test_equalizer_matches_reference(tag):
    g = load_golden("link_level")
    for meth in ("zf", "mmse"):
        # rank-1 ZF has condition number ~1e9: inv() itself is only good to ~1e-6 there
        tol = 1e-5 if (tag == "rank1" and meth == "zf") else 1e-9
        assert relerr(orc.equalize(g[f"eq_{tag}_y"], g[f"eq_{tag}_H"], meth), g[f"eq_{tag}_{meth}"]) < tol
    with pytest.raises(ValueError):
        orc.equalize(g["eq_2x2_y"], g["eq_2x2_H"], "mrc")


@pytest.mark.parametrize("M", [4, 16])
def test_qam_and_ber_match_reference(M):
    g = load_golden("link_level")
    bits = g[f"qam{M}_bits"]
    # the reference's demodulator inverts the restated modulator: pins constellation order and gray map
    assert np.array_equal(g[f"qam{M}_bits_back"], bits)
    assert relerr(orc.qam_modulate(bits, M), g[f"qam{M}_symbols"]) < TOL
    assert np.array_equal(orc.qam_demodulate(g[f"qam{M}_noisy"], M), g[f"qam{M}_noisy_bits"])
    assert orc.bit_error_rate(bits, g[f"qam{M}_noisy_bits"]) == float(g[f"qam{M}_ber"])
    assert abs(np.mean(np.abs(g[f"qam{M}_symbols"]) ** 2) - 1.0) < 0.05       # unit average power
    with pytest.raises(NotImplementedError):
        orc.qam_modulate(bits, 64)


def test_ml_feature_packing_matches_reference():
    g = load_golden("link_level")
    rx, Hls, Htr, mask = g["ml_rx"], g["ml_H_ls"], g["ml_H_true"], g["ml_mask"]
    for i in range(rx.shape[0]):
        for nz in (0, 1):
            x, t = orc.ml_inputs(rx[i], Hls[i], Htr[i], mask[i], bool(nz))
            assert relerr(x, g[f"ml_inputs_{i}_{nz}"]) < TOL and relerr(t, g[f"ml_targets_{i}_{nz}"]) < TOL
    # ChannelDataset works on the complex64 copies it loads from disk
    c = lambda a: a.astype(np.complex64)
    norm = orc.dataset_norm(c(rx), c(Hls), c(Htr))
    assert relerr(np.array(norm).reshape(-1), g["ds_norm"]) < 1e-6
    for i in range(rx.shape[0]):
        for nz in (0, 1):
            x, t = orc.dataset_item(c(rx[i]), c(Hls[i]), c(Htr[i]), mask[i], norm if nz else None)
            assert relerr(x, g[f"ds_inputs_{i}_{nz}"]) < 1e-6 and relerr(t, g[f"ds_targets_{i}_{nz}"]) < 1e-6
    if "ber_approx" in g:
        for j, s in enumerate((-5.0, 10.0, 30.0)):
            assert abs(orc.ber_approximation(Hls[0], Htr[0], s) - g["ber_approx"][j]) < 1e-15
            assert abs(orc.ber_approximation(Htr[0] * 1.01, Htr[0], s) - g["ber_approx_small_err"][j]) < 1e-15


def test_other_grid_matches_reference():
    """7 symbols x 299 used bins, 3 TX x 2 RX, EVA: the oracle on a grid other than the default one."""
    g = load_golden("slot_3x2_eva_7x299")
    cfg = dict(OFDM_CFG, num_symbols=int(g["nsym"]), useful_subcarriers=int(g["useful"]))
    ntx, nrx = int(g["ntx"]), int(g["nrx"])
    sim = orc.simulate(cfg, ntx, nrx, str(g["model"]), float(g["doppler_hz"]), float(g["snr_db"]), float(g["density"]), golden_draws(g))
    assert np.array_equal(sim["pilot_indices"], g["pilot_indices"]) and np.array_equal(sim["pilot_mask"], g["pilot_mask"])
    assert sim["channel"].shape == (7, 2, 3, 299)
    assert relerr(sim["channel"], g["channel"]) < TOL and relerr(sim["rx_symbols"], g["rx_symbols"]) < TOL
    rx4d = np.repeat(g["rx_symbols"][:, :, None, :], ntx, axis=2)
    pos = np.unravel_index(g["pilot_indices"], g["pilot_mask"].shape)
    assert relerr(orc.ls_estimate(rx4d, g["pilot_symbols"], g["pilot_mask"], pos)[:, :, 0], g["H_ls_tx0"]) < TOL
    assert relerr(orc.ls_estimate(rx4d[:, :, :1], g["pilot_symbols"], g["pilot_mask"], pos, "nearest")[:, :, 0], g["H_ls_nearest_tx0"]) < TOL
    assert relerr(orc.mmse_estimate(rx4d, g["pilot_symbols"], g["pilot_mask"], pos, float(g["snr_db"]))[:, :, 0], g["H_mmse_tx0"]) < TOL
