"""Times the CPU oracle on the host cores.  TEST / BENCH INFRASTRUCTURE (used only by bench.py's
`cpu_baseline` leg and `--impl reference` arm).

Two kinds of CPU arm.  "reference": the unmodified reference staged under oracle/_ref/ by oracle/stage_ref.py
(pure Python, NumPy/SciPy; it travels to the GPU box with the working tree).  "port": when that directory is
absent, the oracle *port* run in its cost-faithful profile (oracle/chanest_oracle.py, faithful=True): the full
15344-sample Jakes accumulation per (path, tx, rx), two griddata calls per antenna pair for LS and
for MMSE, and one Np x Np inverse per antenna pair -- the work the reference's
simulate_transmission + LSEstimator('linear').estimate + MMSEEstimator().estimate do per slot.
One worker process per host core, BLAS pinned to one thread per worker.
"""

from __future__ import annotations

import os
import time

WORKLOADS = {
    # name: (ntx, nrx, model, doppler_hz, density)   -- SURVEY.md 8d C1..C3
    "c1_siso_epa": (1, 1, "EPA", 10.0, 0.10),
    "c2_2x2_eva": (2, 2, "EVA", 50.0, 0.10),
    "c3_4x4_etu": (4, 4, "ETU", 200.0, 0.10),
}
OFDM_CFG = {"fft_size": 1024, "cp_length": 72, "num_symbols": 14, "useful_subcarriers": 600,
            "subcarrier_spacing": 15000}
SNRS = (-5, 0, 5, 10, 15, 20, 25, 30)


def _worker(args):
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    workload, seed, nslots, faithful = args
    import numpy as np
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(1)
    except Exception:          # pragma: no cover
        limiter = None
    from oracle import chanest_oracle as orc
    ntx, nrx, model, fd, dens = WORKLOADS[workload]
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    acc = 0.0
    for i in range(nslots):
        snr = SNRS[(seed + i) % len(SNRS)]
        draws = orc.random_draws(rng, OFDM_CFG, ntx, nrx, model, dens)
        out = orc.slot_pipeline(OFDM_CFG, ntx, nrx, model, fd, snr, dens, draws, faithful=faithful)
        acc += float(abs(out["H_mmse"]).sum())
    del limiter
    return time.perf_counter() - t0, acc


REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_available() -> bool:
    """True when oracle/stage_ref.py has staged the unmodified reference under oracle/_ref/."""
    return all(os.path.exists(os.path.join(REF_DIR, "src", f)) for f in ("channel_simulator.py", "baseline_estimators.py"))


def _worker_reference(args):
    """One worker of the `kind = "reference"` CPU arm: the reference's OWN code from oracle/_ref/src --
    simulate_transmission (src/channel_simulator.py:348-421) + LSEstimator('linear').estimate
    (src/baseline_estimators.py:83-117) + MMSEEstimator().estimate (:232-270) + evaluate_estimator (:315-337) per slot,
    exactly the calls the reference's dataset / evaluation loops make."""
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    workload, seed, nslots, _ = args
    import sys
    import numpy as np
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(1)
    except Exception:          # pragma: no cover
        limiter = None
    sys.path.insert(0, os.path.join(REF_DIR, "src"))
    import baseline_estimators as be
    import channel_simulator as cs
    assert os.path.dirname(os.path.abspath(cs.__file__)) == os.path.join(REF_DIR, "src"), cs.__file__
    ntx, nrx, model, fd, dens = WORKLOADS[workload]
    cfg = {"ofdm": dict(OFDM_CFG), "mimo": {"num_tx_antennas": ntx, "num_rx_antennas": nrx}, "channel": {"carrier_freq": 2.0e9}}
    np.random.seed(seed % (2 ** 32))
    t0 = time.perf_counter()
    acc = 0.0
    for i in range(nslots):
        snr = SNRS[(seed + i) % len(SNRS)]
        sim = cs.simulate_transmission(cfg, channel_type=model, doppler_hz=fd, snr_db=snr, pilot_density=dens)
        rx, pp = sim["rx_symbols"], sim["pilot_pattern"]
        rx4d = np.repeat(rx.reshape(rx.shape[0], nrx, 1, rx.shape[2]), ntx, axis=2)      # src/dataset_generator.py:63-64
        H_ls = be.LSEstimator("linear").estimate(rx4d, sim["pilot_symbols"], pp.pilot_mask, pp.pilot_positions)
        H_mm = be.MMSEEstimator().estimate(rx4d, sim["pilot_symbols"], pp.pilot_mask, pp.pilot_positions, snr_db=snr)
        acc += be.evaluate_estimator(sim["channel"], H_ls)["nmse"] + be.evaluate_estimator(sim["channel"], H_mm)["nmse"]
    del limiter
    return time.perf_counter() - t0, acc


def run_sample(workload: str, cores: int, slots_per_core: int = 1, faithful: bool = True, seed0: int = 0):
    """Run cores x slots_per_core slots in parallel; returns (slots, wall seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    jobs = [(workload, seed0 + 1000 * c, slots_per_core, faithful) for c in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        _worker(jobs[0])
    else:
        with ctx.Pool(cores) as pool:
            pool.map(_worker, jobs)
    return cores * slots_per_core, time.perf_counter() - t0


class Pool:
    """Persistent worker pool so repeated steps do not pay process start-up."""

    def __init__(self, cores: int, kind: str = "port"):
        """kind = "reference": the staged reference itself (oracle/_ref); "port": the oracle restatement."""
        import multiprocessing as mp
        self.cores, self.kind = cores, kind
        self.fn = _worker_reference if kind == "reference" else _worker
        self.pool = mp.get_context("spawn").Pool(cores) if cores > 1 else None

    def step(self, workload: str, slots_per_core: int, faithful: bool, seed0: int):
        jobs = [(workload, seed0 + 1000 * c, slots_per_core, faithful) for c in range(self.cores)]
        t0 = time.perf_counter()
        if self.pool is None:
            self.fn(jobs[0])
        else:
            self.pool.map(self.fn, jobs)
        return self.cores * slots_per_core, time.perf_counter() - t0

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
