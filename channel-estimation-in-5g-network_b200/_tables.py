"""Host-side tables the kernels read: TDL profile tables and pilot-pattern interpolation plans.

Built once per (profile set, geometry) / per pilot pattern in float64 and uploaded in the fp32
layouts of include/b2c.h.  Plans use SciPy's Qhull Delaunay / KDTree -- the same third-party
code scipy.interpolate.griddata runs inside the reference's LSEstimator.interpolate_channel
(src/baseline_estimators.py:65-79) -- because only Qhull's own tie-breaking on the integer
pilot lattice reproduces the reference's triangles.  The per-resource-element blend over the
batch is done on the GPU (b2c_ls_interp / b2c_slot_pipeline); this module only builds indices.
"""

from __future__ import annotations

import hashlib
import os
from collections import OrderedDict

import numpy as np

MAX_TAPS = 16
N_OSC = 20

# 3GPP TS 36.104 Annex B.2 power-delay profiles (delay ns, relative power dB); the reference keeps
# the same tables at src/channel_simulator.py:41-54.
PDP = {
    "EPA": ((0, 30, 70, 90, 110, 190, 410), (0.0, -1.0, -2.0, -3.0, -8.0, -17.2, -20.8)),
    "EVA": ((0, 30, 150, 310, 370, 710, 1090, 1730, 2510), (0.0, -1.5, -1.4, -3.6, -0.6, -9.1, -7.0, -12.0, -16.9)),
    "ETU": ((0, 50, 120, 200, 230, 500, 1600, 2300, 5000), (-1.0, -1.0, -1.0, 0.0, 0.0, 0.0, -3.0, -5.0, -7.0)),
}


def path_tables(model_type: str, sampling_rate: float):
    """delays [s], powers_db, normalised powers_linear, integer delay_samples
    (ChannelModel.__init__, src/channel_simulator.py:72-82)."""
    ns, db = PDP[model_type]          # KeyError for an unknown profile, like the reference (:72)
    delays = np.array(ns, dtype=np.float64) * 1e-9
    powers_db = np.array(db, dtype=np.float64)
    lin = 10 ** (powers_db / 10)
    lin = lin / np.sum(lin)
    return delays, powers_db, lin, np.round(delays * sampling_rate).astype(int)


def resolve_taps(delay_samples):
    """(tap_delay, owner_path): later paths overwrite earlier ones that share a sample delay
    (the tap write at src/channel_simulator.py:125 is an assignment)."""
    last = {}
    for p, d in enumerate(delay_samples):
        last[int(d)] = p
    d_sorted = sorted(last)
    return d_sorted, [last[d] for d in d_sorted]


def used_subcarriers(fft_size: int, useful: int):
    """OFDMSystem.used_indices (src/channel_simulator.py:141-148)."""
    dc = fft_size // 2
    idx = np.arange(dc - useful // 2, dc + useful // 2)
    return idx[idx != dc]


def profile_tables(models, sampling_rate, fft_size, used):
    """Numpy arrays for b2c_profiles."""
    M, nsc = len(models), len(used)
    ntaps = np.zeros(M, np.int32)
    npaths = np.zeros(M, np.int32)
    tap_path = np.zeros((M, MAX_TAPS), np.int32)
    tap_delay = np.zeros((M, MAX_TAPS), np.int32)
    tap_amp = np.zeros((M, MAX_TAPS), np.float32)
    tap_tw = np.zeros((M, MAX_TAPS, nsc), np.complex64)
    tap_corr = np.zeros((M, MAX_TAPS, MAX_TAPS), np.complex64)
    f = (np.asarray(used, dtype=np.int64) - fft_size // 2)
    for m, name in enumerate(models):
        _, _, lin, dsamp = path_tables(name, sampling_rate)
        delays, owners = resolve_taps(dsamp)
        if len(delays) > MAX_TAPS:
            raise ValueError(f"profile {name} has {len(delays)} taps > {MAX_TAPS}")
        ntaps[m], npaths[m] = len(delays), len(dsamp)
        tw = np.zeros((MAX_TAPS, nsc), np.complex128)
        for t, (d, p) in enumerate(zip(delays, owners)):
            tap_path[m, t], tap_delay[m, t] = p, d
            tap_amp[m, t] = np.sqrt(lin[p]) / np.sqrt(2 * N_OSC)
            # fftshift(fft(h, N))[i] = sum_d h[d] exp(-j 2 pi (i - N/2) d / N); exact integer phase index
            tw[t] = np.exp(-2j * np.pi * ((f * d) % fft_size) / fft_size)
        tap_tw[m] = tw
        tap_corr[m] = tw @ tw.conj().T
    return {"ntaps": ntaps, "npaths": npaths, "tap_path": tap_path, "tap_delay": tap_delay,
            "tap_amp": tap_amp, "tap_tw": tap_tw, "tap_corr": tap_corr}


PLAN_DTYPE = np.dtype([("i0", "<u2"), ("i1", "<u2"), ("i2", "<u2"), ("flags", "<u2"), ("w0", "<f4"), ("w1", "<f4")])
assert PLAN_DTYPE.itemsize == 16


def _queries(nsym, nsc):
    s, k = np.divmod(np.arange(nsym * nsc), nsc)
    return np.column_stack([s, k]).astype(np.float64)


def interpolation_plan(pilot_positions, nsym, nsc, method="linear"):
    """16-byte plan entries for every resource element (row-major), see b2c_patterns in b2c.h."""
    pts = np.column_stack([np.asarray(pilot_positions[0]), np.asarray(pilot_positions[1])]).astype(np.float64)
    if len(pts) > 65534:
        raise ValueError("more than 65534 pilots")
    plan = np.zeros(nsym * nsc, dtype=PLAN_DTYPE)
    if method == "linear":
        from scipy.spatial import Delaunay
        tri = Delaunay(pts)                        # Qhull, same defaults as LinearNDInterpolator
        q = _queries(nsym, nsc)
        simplex = tri.find_simplex(q)
        inside = simplex >= 0
        sx = np.where(inside, simplex, 0)
        T = tri.transform[sx]
        b = np.einsum("nij,nj->ni", T[:, :2, :], q - T[:, 2, :])
        v = tri.simplices[sx]
        plan["i0"], plan["i1"], plan["i2"] = (np.where(inside, v[:, i], 0) for i in range(3))
        plan["w0"], plan["w1"] = np.where(inside, b[:, 0], 0.0), np.where(inside, b[:, 1], 0.0)
        plan["flags"] = inside.astype(np.uint16)
    elif method == "nearest":
        from scipy.spatial import KDTree            # NearestNDInterpolator's tree (leafsize 10)
        _, i = KDTree(pts).query(_queries(nsym, nsc))
        plan["i0"] = plan["i1"] = plan["i2"] = i
        plan["w0"], plan["w1"], plan["flags"] = 1.0, 0.0, 1
    else:
        raise ValueError(f"Unknown interpolation method {method!r} (griddata knows 'linear', 'nearest', 'cubic'; "
                         "'cubic' is a dense map, see cubic_matrix)")
    return plan


_CUBIC_CACHE: "OrderedDict[tuple, np.ndarray]" = OrderedDict()


def cubic_matrix(pilot_indices, nsym, nsc, cache_size=8):
    """griddata(method='cubic') as a dense linear map W [nsym*nsc, Np] (float32), fill_value 0.

    SciPy's CloughTocher2DInterpolator estimates vertex gradients with an iterative global solver and
    then evaluates a piecewise-cubic Bezier patch: for a fixed triangulation both steps are linear in
    the pilot values (the solver's stopping rule perturbs this by ~1e-6), so interpolating the identity
    matrix yields the map.  The basis run uses a tight tolerance; measured agreement with the
    reference's own two-griddata-call route is 2e-7 (tests/test_host_logic.py)."""
    from scipy.interpolate import CloughTocher2DInterpolator
    idx = np.ascontiguousarray(np.asarray(pilot_indices, dtype=np.int64))
    key = (hashlib.sha1(idx.tobytes()).hexdigest(), nsym, nsc)
    hit = _CUBIC_CACHE.get(key)
    if hit is not None:
        _CUBIC_CACHE.move_to_end(key)
        return hit
    pos = np.unravel_index(idx, (nsym, nsc))
    pts = np.column_stack([pos[0], pos[1]]).astype(np.float64)
    ct = CloughTocher2DInterpolator(pts, np.eye(len(idx)), fill_value=0.0, tol=1e-10, maxiter=2000)
    W = np.ascontiguousarray(ct(_queries(nsym, nsc)), dtype=np.float32)
    _CUBIC_CACHE[key] = W
    while len(_CUBIC_CACHE) > cache_size:
        _CUBIC_CACHE.popitem(last=False)
    return W


def identity_plan(n):
    """Plan in which resource element e is its own 'pilot' e (used to broadcast / score a grid that
    already holds per-RE values: the cubic path, evaluate_estimator)."""
    plan = np.zeros(n, dtype=PLAN_DTYPE)
    plan["i0"] = plan["i1"] = plan["i2"] = np.arange(n)
    plan["w0"], plan["flags"] = 1.0, 1
    return plan


def finalize_plan(plan, zero_slot):
    """Device form of a plan (b2c_patterns in b2c.h): resource elements outside the hull point all
    three taps at `zero_slot` (the zero entry the kernels append to the pilot vector, index np_max)
    with weights (1, 0, 0), and one extra all-outside row is appended for the kernels' idle lanes."""
    out = np.zeros(len(plan) + 1, dtype=PLAN_DTYPE)
    out[:-1] = plan
    outside = out["flags"] == 0
    for f in ("i0", "i1", "i2"):
        out[f][outside] = zero_slot
    out["w0"][outside], out["w1"][outside] = 1.0, 0.0
    return out


_PLAN_CACHE: "OrderedDict[tuple, np.ndarray]" = OrderedDict()


def cached_plan(pilot_indices, nsym, nsc, method="linear", cache_size=256):
    """Plans are pure functions of the pilot set; cache them by content hash."""
    idx = np.ascontiguousarray(np.asarray(pilot_indices, dtype=np.int64))
    key = (hashlib.sha1(idx.tobytes()).hexdigest(), nsym, nsc, method)
    hit = _PLAN_CACHE.get(key)
    if hit is not None:
        _PLAN_CACHE.move_to_end(key)
        return hit
    plan = interpolation_plan(np.unravel_index(idx, (nsym, nsc)), nsym, nsc, method)
    _PLAN_CACHE[key] = plan
    while len(_PLAN_CACHE) > cache_size:
        _PLAN_CACHE.popitem(last=False)
    return plan


# ---- pools of plans: built in parallel, cached on disk -------------------------------------------------------------
def _plan_cache_dir():
    d = os.environ.get("B2C_PLAN_CACHE", os.path.join(os.path.expanduser("~"), ".cache", "b2c_plans"))
    if d in ("", "0", "off"):
        return None
    try:
        os.makedirs(d, exist_ok=True)
        return d
    except OSError:
        return None


def _plan_key(idx, nsym, nsc, method):
    return f"{hashlib.sha1(idx.tobytes()).hexdigest()}_{nsym}x{nsc}_{method}"


def _build_plans(job):
    """Worker of plans_for (also run in-process): [(pilot index array)] -> [plan]."""
    idx_list, nsym, nsc, method = job
    return [interpolation_plan(np.unravel_index(i, (nsym, nsc)), nsym, nsc, method) for i in idx_list]


def _build_plans_subprocess(idx_list, nsym, nsc, method, workers):
    """Fan the patterns out over `workers` fresh interpreters running this file as a script (no fork of a process
    that holds a CUDA context, no re-import of the caller's __main__): job and result travel as .npz / .npy files."""
    import subprocess
    import sys
    import tempfile
    with tempfile.TemporaryDirectory(prefix="b2c_plans_") as tmp:
        procs = []
        for w in range(workers):
            mine = idx_list[w::workers]
            job, res = os.path.join(tmp, f"job{w}.npz"), os.path.join(tmp, f"res{w}.npy")
            np.savez(job, nsym=nsym, nsc=nsc, method=method, **{f"p{j}": a for j, a in enumerate(mine)})
            procs.append((len(mine), res, subprocess.Popen([sys.executable, os.path.abspath(__file__), "--build-plans", job, res])))
        out = [None] * len(idx_list)
        for w, (n, res, pr) in enumerate(procs):
            if pr.wait() != 0:
                raise RuntimeError(f"plan worker {w} failed (exit {pr.returncode})")
            plans = np.load(res)
            for j in range(n):
                out[w + j * workers] = plans[j]
    return out


def plans_for(pilot_index_list, nsym, nsc, method="linear", workers=None):
    """Plans of a whole pattern pool.  The reference draws a fresh pilot pattern per sample
    (src/channel_simulator.py:391); a realistic pool therefore holds hundreds of patterns and a Qhull
    triangulation + point location per pattern (~20-90 ms) adds up: plans missing from the in-memory and on-disk
    caches (B2C_PLAN_CACHE, default ~/.cache/b2c_plans, keyed by the pilot set's SHA-1) are built by a pool of
    worker interpreters (fresh processes, not forks: the parent usually holds a CUDA context)."""
    idxs = [np.ascontiguousarray(np.asarray(p, dtype=np.int64)) for p in pilot_index_list]
    out, todo = [None] * len(idxs), []
    cdir = _plan_cache_dir()
    for n, idx in enumerate(idxs):
        key = (hashlib.sha1(idx.tobytes()).hexdigest(), nsym, nsc, method)
        hit = _PLAN_CACHE.get(key)
        if hit is None and cdir is not None:
            path = os.path.join(cdir, _plan_key(idx, nsym, nsc, method) + ".npy")
            if os.path.exists(path):
                try:
                    hit = np.load(path)
                    if hit.dtype != PLAN_DTYPE or hit.shape != (nsym * nsc,):
                        hit = None
                except Exception:
                    hit = None
        if hit is None:
            todo.append(n)
        else:
            out[n] = hit
    if todo:
        if workers is None:
            try:
                workers = len(os.sched_getaffinity(0))
            except Exception:
                workers = os.cpu_count() or 1
        workers = max(1, min(int(workers), 32, len(todo) // 8))
        if workers <= 1:
            built = _build_plans(([idxs[n] for n in todo], nsym, nsc, method))
        else:
            built = _build_plans_subprocess([idxs[n] for n in todo], nsym, nsc, method, workers)
        for n, plan in zip(todo, built):
            out[n] = plan
            if cdir is not None:
                try:
                    tmp = os.path.join(cdir, f".{os.getpid()}_{n}.npy")
                    np.save(tmp, plan)
                    os.replace(tmp, os.path.join(cdir, _plan_key(idxs[n], nsym, nsc, method) + ".npy"))
                except OSError:
                    pass
    for idx, plan in zip(idxs, out):     # keep the small pools of the drop-in shims hot in memory
        if len(idxs) <= 256:
            _PLAN_CACHE[(hashlib.sha1(idx.tobytes()).hexdigest(), nsym, nsc, method)] = plan
    while len(_PLAN_CACHE) > 256:
        _PLAN_CACHE.popitem(last=False)
    return out


if __name__ == "__main__":      # plan worker of _build_plans_subprocess
    import sys
    if len(sys.argv) == 4 and sys.argv[1] == "--build-plans":
        with np.load(sys.argv[2]) as z:
            n = len([k for k in z.files if k.startswith("p")])
            plans = _build_plans(([z[f"p{j}"] for j in range(n)], int(z["nsym"]), int(z["nsc"]), str(z["method"])))
        np.save(sys.argv[3], np.stack(plans) if plans else np.zeros((0,), PLAN_DTYPE))
    else:
        sys.exit("usage: _tables.py --build-plans job.npz result.npy")
