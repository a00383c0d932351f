"""Times the CPU oracle on the host cores.  TEST / BENCH INFRASTRUCTURE (used only by bench.py's
`cpu_baseline` leg and `--impl reference` arm).

The reference is pure Python (NumPy/SciPy) and cannot travel to the GPU box, so the CPU arm is the
oracle *port* run in its cost-faithful profile (oracle/chanest_oracle.py, faithful=True): the full
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

    def __init__(self, cores: int):
        import multiprocessing as mp
        self.cores = cores
        self.pool = mp.get_context("spawn").Pool(cores) if cores > 1 else None

    def step(self, workload: str, slots_per_core: int, faithful: bool, seed0: int):
        jobs = [(workload, seed0 + 1000 * c, slots_per_core, faithful) for c in range(self.cores)]
        t0 = time.perf_counter()
        if self.pool is None:
            _worker(jobs[0])
        else:
            self.pool.map(_worker, jobs)
        return self.cores * slots_per_core, time.perf_counter() - t0

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
