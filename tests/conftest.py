"""pytest configuration: marker registration, import paths, shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "channel-estimation-in-5g-network_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

OFDM_CFG = {"fft_size": 1024, "cp_length": 72, "num_symbols": 14,
            "useful_subcarriers": 600, "subcarrier_spacing": 15000}

SLOT_CASES = ["slot_siso_epa", "slot_2x2_eva", "slot_4x4_etu", "slot_2x2_etu_5pct", "slot_2x1_epa_1pct"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def golden_draws(g):
    return {k: g[k] for k in ("perm", "pilot_phase", "data_phase", "jakes_u", "noise_re", "noise_im")}


def full_config(ntx, nrx):
    return {"ofdm": dict(OFDM_CFG), "mimo": {"num_tx_antennas": int(ntx), "num_rx_antennas": int(nrx)},
            "channel": {"carrier_freq": 2.0e9}}


@pytest.fixture(scope="session")
def ofdm_cfg():
    return dict(OFDM_CFG)


def relerr(a, b):
    """Norm-wise relative error max|a-b| / max|b| used by the float parity checks."""
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def assert_close_elementwise(a, b, rtol=1e-4, floor=1e-6, what=""):
    """Element-wise companion of relerr: |a - b| <= rtol * |b| + floor * max|b| for EVERY element, so that a
    systematic error on the small-magnitude resource elements (deep fades, hull edges) cannot hide behind the
    array's largest value.  The floor (1e-6 of the largest magnitude, ~10 fp32 ulps of it) is what fp32 storage of a
    sum of O(1) terms can resolve; exact zeros of `b` (outside the pilots' hull) must be within it."""
    a, b = np.asarray(a), np.asarray(b)
    bound = rtol * np.abs(b) + floor * max(float(np.max(np.abs(b))), 1e-300)
    bad = np.abs(a - b) > bound
    if bad.any():
        i = np.unravel_index(np.argmax(np.abs(a - b) - bound), a.shape)
        raise AssertionError(f"{what}: {int(bad.sum())} of {a.size} elements outside |a-b| <= {rtol}|b| + {floor} max|b|; "
                             f"worst at {i}: got {a[i]}, want {b[i]}")
