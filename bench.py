#!/usr/bin/env python
"""Benchmark of the hot path: OFDM slots/s simulated + LS-estimated + MMSE-estimated.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME] [--layout full|compact]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is `launches_per_step` passes of the hot path (K1a tap gains -> fused slot kernel [-> dense Wiener GEMM -> K3]
-> K5 statistics fold), each over one batch of synthetic slots, into rotating output buffers; launches_per_step is
sized in the warm-up so that the timed region lasts >= --min-seconds whatever --steps is (a 50 ms burst never
reaches the board's power-capped steady state).  Default workload = BASELINE.json configs[2]: 4x4 ETU, 200 Hz
Doppler, FFT 1024 / CP 72, 14 symbols, 599 used bins, 10 % pilots, SNR cycling over {-5,...,30} dB.
Prints ONE JSON line.

Workloads (BASELINE.json `configs`):
  c1_siso_epa        config 1   SISO EPA 10 Hz
  c2_2x2_eva         config 2   2x2 EVA 50 Hz, default (alpha) MMSE
  c2_2x2_eva_dense   config 2   2x2 EVA 50 Hz, LS + dense Wiener MMSE on the tensor cores (one 838 x 838 W per SNR)
  c3_4x4_etu         config 3   4x4 ETU 200 Hz (default)
  c4_sweep           config 4   4x4 EVA 50 Hz, pilot density 1..10 % x 8 SNRs, MSE / NMSE / BER-proxy curves (statistics only)
  c5_mixed           config 5   4x4 mixed EPA/EVA/ETU x 4 Dopplers x 8 SNRs x 2 densities, dataset arrays + statistics
"""

from __future__ import annotations

import argparse
import hashlib
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "channel-estimation-in-5g-network_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "OFDM slots/sec simulated+LS/MMSE-estimated"
UNIT = "slots/s"
SNRS = (-5, 0, 5, 10, 15, 20, 25, 30)
ALL = ("H_true", "rx", "tx", "H_ls", "H_mmse", "stats")
_SNR_TXT = "SNR cycling over [-5, 0, 5, 10, 15, 20, 25, 30] dB"
WORKLOADS = {
    "c1_siso_epa": dict(ntx=1, nrx=1, models=["EPA"], dopplers=[10.0], densities=[0.10], mmse="default", want=ALL,
                        desc=f"SISO EPA 10 Hz, FFT 1024/CP 72, 10% pilots, {_SNR_TXT}, simulate+LS(linear)+MMSE(default)"),
    "c2_2x2_eva": dict(ntx=2, nrx=2, models=["EVA"], dopplers=[50.0], densities=[0.10], mmse="default", want=ALL,
                       desc=f"2x2 EVA 50 Hz, FFT 1024/CP 72, 10% pilots, {_SNR_TXT}, simulate+LS(linear)+MMSE(default)"),
    "c2_2x2_eva_dense": dict(ntx=2, nrx=2, models=["EVA"], dopplers=[50.0], densities=[0.10], mmse="dense", want=ALL, batch=18944,
                             desc=f"2x2 EVA 50 Hz, FFT 1024/CP 72, 10% pilots, {_SNR_TXT}, simulate+LS(linear)+MMSE(known-covariance "
                                  "Wiener filter W = R (R + s2 I)^-1, one 838 x 838 W per SNR)"),
    "c2_2x2_eva_dense_stats": dict(ntx=2, nrx=2, models=["EVA"], dopplers=[50.0], densities=[0.10], mmse="dense", want=("stats",),
                                   batch=18944,
                                   desc=f"2x2 EVA 50 Hz, FFT 1024/CP 72, 10% pilots, {_SNR_TXT}, simulate+LS(linear)+MMSE(known-covariance "
                                        "Wiener filter, one 838 x 838 W per SNR), per-SNR MSE / NMSE curves only (no array leaves the SM: "
                                        "slot kernel -> grouped GEMM -> b2c_dense_score)"),
    "c3_4x4_etu": dict(ntx=4, nrx=4, models=["ETU"], dopplers=[200.0], densities=[0.10], mmse="default", want=ALL,
                       desc=f"4x4 ETU 200 Hz, FFT 1024/CP 72, 10% pilots, {_SNR_TXT}, simulate+LS(linear)+MMSE(default)"),
    "c4_sweep": dict(ntx=4, nrx=4, models=["EVA"], dopplers=[50.0], densities=[0.01 * d for d in range(1, 11)], mmse="default",
                     want=("stats",),
                     desc="4x4 EVA 50 Hz, pilot density 1..10 % x SNR [-5, 0, 5, 10, 15, 20, 25, 30] dB, simulate+LS(linear)+MMSE(default), "
                          "per-cell MSE / NMSE / BER-proxy curves (statistics only, no array leaves the SM)"),
    "c5_mixed": dict(ntx=4, nrx=4, models=["EPA", "EVA", "ETU"], dopplers=[10.0, 50.0, 100.0, 200.0], densities=[0.05, 0.10],
                     mmse="default", want=("H_true", "rx", "tx", "H_ls", "stats"),
                     desc="4x4 mixed EPA/EVA/ETU x Doppler [10, 50, 100, 200] Hz x SNR [-5..30] dB x pilot density [5, 10] %, every "
                          "parameter drawn per slot (Philox stream 3), dataset arrays of generate_sample (H_true, rx, tx, H_ls) + statistics"),
}
CPU_WORKLOAD = {"c2_2x2_eva_dense": "c2_2x2_eva", "c2_2x2_eva_dense_stats": "c2_2x2_eva", "c4_sweep": "c3_4x4_etu", "c5_mixed": "c3_4x4_etu"}   # cost-equivalent CPU samples


def slot_bytes(ntx, nrx, nsym=14, nsc=599, compact=False, want=ALL):
    """Algorithmic bytes one slot of the pipeline writes: the requested arrays among H_true, H_ls,
    H_mmse, rx, tx, complex64.  compact: the tx-replicated arrays (H_ls, H_mmse, tx) counted once."""
    full, row = nsym * nrx * ntx * nsc, nsym * nrx * nsc
    rep = row if compact else full
    n = {"H_true": full, "H_ls": rep, "H_mmse": rep, "rx": row, "tx": nsym * nsc * (1 if compact else ntx)}
    return 8 * sum(v for k, v in n.items() if k in want)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            j = json.load(fh)
        return {"hbm_gbs": float(j["hbm_gbs"]), "bf16_tflops": float(j.get("bf16_tflops", 0) or 0),
                "bf16_tflops_sustained": float(j.get("bf16_tflops_sustained", 0) or 0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1600.0, "bf16_tflops_sustained": 1350.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms while the GPU is under load.  It is
    started before the warm-up steps (nvidia-smi needs ~100 ms to come up); samples whose timestamps
    fall between mark_begin() (start of warm-up) and stop() (end of the timed region) are kept."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0 = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        t1 = time.time()
        time.sleep(0.05)                    # let the last sample arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if len(r) >= 7 and (self.t0 is None or self.t0 <= t <= t1 + 0.05)]
        if not rows:                        # region shorter than one sampling period: nearest samples
            rows = [r for _, r in self.rows[-3:] if len(r) >= 7]

        def num(s):
            return s.replace(".", "", 1).isdigit()
        sm = [float(r[1]) for r in rows if num(r[1])]
        mx = [float(r[2]) for r in rows if num(r[2])]
        pw = [float(r[7]) for r in rows if len(r) > 7 and num(r[7])]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for j, n in enumerate(names) if any(r[3 + j].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(pw) if pw else None,
                "window": "warm-up + timed steps"}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_kind():
    from oracle import cpu_bench
    return "reference" if cpu_bench.reference_available() else "port"


def cpu_sample_text(kind, workload):
    if kind == "reference":
        return (f"slots of {workload} through the UNMODIFIED reference staged in oracle/_ref (src/channel_simulator.py "
                "simulate_transmission + src/baseline_estimators.py LSEstimator('linear').estimate + MMSEEstimator().estimate + "
                "evaluate_estimator)")
    return (f"slots of {workload} through the oracle port (oracle/chanest_oracle.py, cost-faithful profile: simulate + LS + MMSE; "
            "oracle/_ref is not staged on this box)")


def cpu_baseline(workload, cores, budget_s=20.0):
    """The reference's CPU implementation of the path on the host cores, bounded sample."""
    from oracle import cpu_bench
    kind = cpu_kind()
    wl = CPU_WORKLOAD.get(workload, workload)
    pool = cpu_bench.Pool(cores, kind)
    slots, secs = pool.step(wl, 1, True, 7)       # one slot per core per round; rounds sized to the budget from this probe
    rounds = max(0, min(3, int(budget_s / max(secs, 1e-3)) - 1))
    for r in range(rounds):
        s2, t2 = pool.step(wl, 1, True, 100 + r)
        slots, secs = slots + s2, secs + t2
    pool.close()
    return {"value": slots / secs, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{slots} {cpu_sample_text(kind, wl)}, {cores} worker processes, BLAS 1 thread each, {secs:.1f} s wall"}


def workload_config(args):
    """`config` of the JSON line: identical in both arms (arm-specific details go under `arm`)."""
    return {"workload": f"{args.workload}: {WORKLOADS[args.workload]['desc']}"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref when staged, else the oracle
    port) on all host cores, same metric / workload as the b200 arm."""
    if rank != 0:
        return
    from oracle import cpu_bench
    cores = host_cores()
    kind = cpu_kind()
    wl = CPU_WORKLOAD.get(args.workload, args.workload)
    pool = cpu_bench.Pool(cores, kind)
    # One step = one slot per worker process (the smallest sample that keeps every host core busy), ~2-7 s each.
    # A step count that would overrun the wall budget is cut short and the line reports the steps actually timed.
    t_start = time.perf_counter()
    for w in range(args.warmup):
        pool.step(wl, 1, True, 10 + w)
    slots = secs = 0.0
    done = 0
    for k in range(args.steps):
        s, t = pool.step(wl, 1, True, 1000 + k)
        slots, secs, done = slots + s, secs + t, done + 1
        if time.perf_counter() - t_start + 1.5 * t > args.reference_budget:
            break
    pool.close()
    truncated = done < args.steps
    args.steps = done
    value = slots / secs
    sample = f"each step = {cores} {cpu_sample_text(kind, wl)}, one per worker process"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "arm": {"slots_per_step": cores, "rng": "numpy global RandomState (reference) / Generator (port)",
                "cpu_sample_workload": wl,
                "note": (f"stopped after {done} timed steps: wall budget {args.reference_budget:.0f} s" if truncated else "all steps timed")},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def d2h_probe(torch, dev, nbytes, seconds=0.25):
    """Bare pinned device->host copies of one chunk slab (one cudaMemcpyAsync each), back to back: the link's own
    rate for exactly the transfer size the host pipeline issues.  Returns GB/s."""
    src = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    dst = [torch.empty((nbytes,), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        for i in range(2):
            dst[i].copy_(src, non_blocking=True)
        st.synchronize()
        n = max(2, int(seconds * 50e9 / nbytes))
        t0 = time.perf_counter()
        for i in range(n):
            dst[i & 1].copy_(src, non_blocking=True)
        st.synchronize()
        dt = time.perf_counter() - t0
    return n * nbytes / dt / 1e9


def run_b200(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist

    # CPU baseline first (rank 0, N=1 only), before CUDA is initialised in this process
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.workload, host_cores(), args.cpu_budget)

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    import _b2c
    from dataset_generator import ChannelEstimationDataset, philox_param_choice, shard_range, sharded_statistics
    from engine import SlotEngine, WienerBank
    from host_pipeline import HostPipeline, bind_to_gpu_numa_node

    W = WORKLOADS[args.workload]
    ntx, nrx, want, dense = W["ntx"], W["nrx"], tuple(W["want"]), W["mmse"] == "dense"
    cfg = {"ofdm": {"fft_size": 1024, "cp_length": 72, "num_symbols": 14, "useful_subcarriers": 600,
                    "subcarrier_spacing": 15000}, "mimo": {"num_tx_antennas": ntx, "num_rx_antennas": nrx},
           "channel": {"models": W["models"], "doppler_hz": W["dopplers"], "carrier_freq": 2.0e9},
           "pilots": {"density": W["densities"]}, "simulation": {"snr_range": list(SNRS)}}
    eng = SlotEngine(cfg, models=tuple(W["models"]))
    dev = eng.device
    ppd = 1 if dense else args.patterns            # one Wiener matrix per (pattern, SNR): the dense workload keeps one pattern
    t0 = time.perf_counter()
    if world > 1 and rank != 0:
        dist.barrier()                     # rank 0 builds the plans first (they land in the on-disk cache), the others read them
    pool = eng.random_pool(W["densities"], per_density=ppd, seed=42)
    if world > 1 and rank == 0:
        dist.barrier()
    pool_build_s = time.perf_counter() - t0
    B = args.batch
    stats_only = want == ("stats",)
    compact = args.layout == "compact" and not stats_only and not dense
    pitch = None if stats_only else args.pitch
    NBUF = 1 if stats_only else 2                   # rotating output buffers (each >> the 126 MB L2)
    outs = [eng.alloc_outputs(B, want, compact=compact, pitch=pitch) for _ in range(NBUF)]
    wss = [eng.workspace(B) for _ in range(NBUF)]

    # per-slot parameters of one batch (reused by every launch; Philox keys still differ: they follow the global slot index)
    nm, nd, ns, nden = len(W["models"]), len(W["dopplers"]), len(SNRS), len(W["densities"])
    idx = np.arange(B)
    if args.workload == "c5_mixed":
        mi, di, si, pi = philox_param_choice(args.seed, 0, B, (nm, nd, ns, nden))
    elif args.workload == "c4_sweep":
        mi, di, si, pi = np.zeros(B, np.int64), np.zeros(B, np.int64), idx % ns, (idx // ns) % nden
    else:
        mi, di, si, pi = np.zeros(B, np.int64), np.zeros(B, np.int64), idx % ns, np.zeros(B, np.int64)
    pat = (pi * ppd + (idx // (ns * nden)) % ppd).astype(np.int32)
    snr_host = np.asarray(SNRS, np.float32)[si]
    model_id = torch.from_numpy(mi.astype(np.int32)).to(dev)
    doppler = torch.from_numpy(np.asarray(W["dopplers"], np.float32)[di]).to(dev)
    snr = torch.from_numpy(snr_host).to(dev)
    pattern = torch.from_numpy(pat).to(dev)
    nbins = ns * nden if args.workload == "c4_sweep" else ns
    bin_id = torch.from_numpy((pi * ns + si if args.workload == "c4_sweep" else si).astype(np.int32)).to(dev)
    bins = torch.zeros((nbins, _b2c.N_BINSTAT), dtype=torch.float64, device=dev)

    bank = None
    if dense:      # known covariance of the LS pilot estimates: exponential time / frequency correlation model
        pidx = pool.pilot_indices[0]
        ps, pk = pidx // eng.nsc, pidx % eng.nsc
        R = 0.4 * np.exp(-np.abs(ps[:, None] - ps[None, :]) / 20.0 - np.abs(pk[:, None] - pk[None, :]) / 60.0) \
            * np.exp(1j * 2 * np.pi * (pk[:, None] - pk[None, :]) * 3 / 1024)
        bank = WienerBank(eng, pool, {0: R}, SNRS)
        dplan = bank.plan_batch(eng, pat, snr_host, B)     # the batch's (pattern, SNR) grouping: host work, done once

    launches = [0]

    def one_pass(gslot, buf):
        """One batch: global slots [gslot, gslot + B) of this rank through the engine's own launcher."""
        if dense:
            eng.run(B, model_id, doppler, snr, pattern, pool, slot0=gslot, seed=args.seed, out=outs[buf], ws=wss[buf],
                    mmse="dense", wiener=bank, dense_plan=dplan)
            launches[0] += 4                         # K1a, slot kernel, one grouped GEMM (8 SNR groups), K3 / b2c_dense_score
        else:
            eng.run(B, model_id, doppler, snr, pattern, pool, slot0=gslot, seed=args.seed, out=outs[buf], ws=wss[buf],
                    compact=compact)
            launches[0] += 2
        eng.stats_bins(outs[buf]["stats"], bin_id, nbins, bins, snr_db=snr)
        launches[0] += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- size the step: launches_per_step such that the timed region lasts >= --min-seconds ---------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    one_pass(0, 0)                                   # first launch (module load) outside the sampled window
    barrier()
    sampler.mark_begin()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for i in range(3):
        one_pass((1 + i) * B, i % NBUF)
    ev[1].record()
    barrier()
    pass_ms = ev[0].elapsed_time(ev[1]) / 3
    lps = args.launches_per_step or max(1, math.ceil(args.min_seconds * 1e3 / (pass_ms * args.steps)))
    if world > 1:                                    # every rank runs the same number of launches
        t = torch.tensor([lps], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lps = int(t.item())
    per_rank_passes = (args.warmup + args.steps) * lps + 32
    base = rank * per_rank_passes * B                # this rank's block of global slot indices

    n = 4
    for w in range(args.warmup):
        for j in range(lps):
            one_pass(base + n * B, n % NBUF)
            n += 1
    barrier()
    bins.zero_()
    launches[0] = 0
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_beg.record()
    for s in range(args.steps):
        for j in range(lps):
            one_pass(base + n * B, n % NBUF)
            n += 1
    if world > 1:
        dist.all_reduce(bins)            # the path's only collective: per-bin statistics over NVLink
    t_end.record()
    barrier()
    clocks = sampler.stop()
    ms = t_beg.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    slots_total = world * args.steps * lps * B
    value = slots_total / (ms * 1e-3)
    timed_launches = launches[0]
    nb = bins.cpu().numpy()

    # ---- dominant kernel alone, CUDA events on its stream, over rotating buffers (right after the timed region) --------
    L = _b2c.lib()
    from _b2c import PilotIO, Slots, check, dptr, ref, rows_ptr, stream_ptr
    P = _b2c.row_pitch(outs[0]["H_true"]) if "H_true" in outs[0] else eng.nsc
    g_out = eng._with_pitch(eng.geom, P)
    kev = []
    for i in range(14):
        o, w = outs[i % NBUF], wss[i % NBUF]
        sl = Slots(base + (n + i) * B, args.seed, model_id.data_ptr(), doppler.data_ptr(), snr.data_ptr(), pattern.data_ptr())
        check(L.b2c_tap_gains(ref(eng.geom), ref(eng.prof), ref(sl), None, B, dptr(w["gains"], "c64"), dptr(w["noise_std"], "f32"), stream_ptr()))
        pio = PilotIO(w["hp"].data_ptr(), dplan.col.data_ptr(), w["hp"].shape[1]) if dense else None    # as the dense pass launches it
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(L.b2c_slot_pipeline(ref(g_out), ref(eng.prof), ref(pool.struct), ref(sl), None, B, dptr(w["gains"], "c64"),
                                  dptr(w["noise_std"], "f32"), rows_ptr(o.get("H_true"), P, True), rows_ptr(o.get("rx"), P, True),
                                  rows_ptr(o.get("tx"), P, True), rows_ptr(o.get("H_ls"), P, True),
                                  None if dense else rows_ptr(o.get("H_mmse"), P, True), dptr(o["stats"], "f64"),
                                  int(compact), ref(pio), stream_ptr()))
        b.record()
        if i >= 2:
            kev.append((a, b))
    torch.cuda.synchronize()
    k_ms = statistics.mean(a.elapsed_time(b) for a, b in kev)
    slot_want = tuple(k for k in want if not (dense and k == "H_mmse"))
    alg = slot_bytes(ntx, nrx, compact=compact, want=slot_want) * B

    tensor = None
    if dense:     # the grouped GEMM (all SNR groups, one launch), alone: useful flops = 8 np^2 per column; issued = 3x (3xTF32)
        import ctypes as C
        npil = int(pool.npilots_host[0])
        ncols = B * nrx
        hp, hm = wss[0]["hp"], wss[0]["hm"]
        evs = []
        for i in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            check(L.b2c_dense_apply_grouped(C.byref(dplan.groups), len(dplan.groups), dptr(hp, "c64"), dptr(hm, "c64"), hp.shape[1],
                                            stream_ptr()))
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        g_ms = statistics.mean(a.elapsed_time(b) for a, b in evs[2:])
        useful = 8.0 * npil * npil * ncols / (g_ms * 1e-3) / 1e12
        tensor = {"gemm_ms": g_ms, "columns": ncols, "np": npil, "useful_tflops": useful, "issued_tf32_tflops": 3 * useful}
        if stats_only:      # the scoring pass alone (CFR regenerated from the gains, filtered pilots interpolated, error sums)
            sl = Slots(base, args.seed, model_id.data_ptr(), doppler.data_ptr(), snr.data_ptr(), pattern.data_ptr())
            evs = []
            for i in range(10):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                check(L.b2c_dense_score(ref(eng.geom), ref(eng.prof), ref(pool.struct), ref(sl), B, dptr(wss[0]["gains"], "c64"),
                                        dptr(hm, "c64"), dptr(dplan.col, "i32"), hp.shape[1], dptr(outs[0]["stats"], "f64"), stream_ptr()))
                b.record()
                evs.append((a, b))
            torch.cuda.synchronize()
            tensor["score_ms"] = statistics.mean(a.elapsed_time(b) for a, b in evs[2:])
        # a measured TF32 peak of this box (library matmul, measurement only -- not on the product path): 8192^3, best of 8
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            ma = torch.randn((8192, 8192), device=dev)
            mb = torch.randn((8192, 8192), device=dev)
            best = 1e9
            for i in range(10):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                torch.matmul(ma, mb)
                b.record()
                torch.cuda.synchronize()
                if i >= 2:
                    best = min(best, a.elapsed_time(b))
            tensor["tf32_peak_measured_tflops"] = 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
            del ma, mb
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

    # ---- the same metric through the public API (dataset_generator.sharded_statistics), device resident ----------------
    value_api = None
    if not args.no_api:
        ds = ChannelEstimationDataset(cfg, rng='philox', seed=args.seed, patterns_per_density=ppd)
        ds._pool = pool
        api_slots = max(B, min(slots_total // world, 40 * B))
        arrays = tuple(k for k in want if k != "stats")
        kw = dict(mmse="dense", wiener=bank) if dense else {}      # (the grouping by SNR is then redone per batch: SNRs are drawn per slot)
        sharded_statistics(cfg, 2 * B, 0, 1, batch=B, seed=args.seed, want_arrays=arrays, dataset=ds, **kw)      # warm-up
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        api_bins = sharded_statistics(cfg, api_slots * world, rank, world, batch=B, seed=args.seed, want_arrays=arrays, dataset=ds, **kw)
        if world > 1:
            dist.all_reduce(api_bins)
        b.record()
        barrier()
        api_ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([api_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            api_ms = float(t.item())
        value_api = {"value": api_slots * world / (api_ms * 1e-3), "unit": UNIT, "slots": api_slots * world,
                     "call": "dataset_generator.sharded_statistics(config, total_slots, rank, world_size, batch, want_arrays=...) "
                             "(host loop: per-batch parameter draw + upload, engine launches, K5 fold; every slot's parameters "
                             "drawn from Philox stream 3)"}

    # ---- fixed-range statistics checksum: global slots 0..65535 whatever N is (outside the timed region) --------------
    CHK = 65536
    lo, hi = shard_range(CHK, rank, world)
    qsum = torch.zeros((nrx * 6,), dtype=torch.int64, device=dev)
    pos = lo
    while pos < hi:
        m = min(B, hi - pos)
        j = np.arange(pos, pos + m)
        o = eng.run(m, (j % nm).astype(np.int32), float(W["dopplers"][0]), np.asarray(SNRS, np.float32)[j % ns],
                    ((j // ns) % nden * ppd).astype(np.int32), pool, slot0=pos, seed=args.seed, want=("stats",))
        # per-slot sums are a pure function of the global slot index; rounded to 2^-20 they add exactly in int64, so the
        # checksum is bit-identical however the range is split over ranks
        qsum += torch.round(o["stats"].reshape(m, -1) * float(2 ** 20)).to(torch.int64).sum(dim=0)
        pos += m
    if world > 1:
        dist.all_reduce(qsum)
    chk = qsum.cpu().numpy()
    checksum = {"slots": CHK, "range": "global slots [0, 65536), seed %d" % args.seed,
                "sha1": hashlib.sha1(chk.tobytes()).hexdigest(), "int64_sums_first4": [int(v) for v in chk[:4]],
                "how": "per-slot statistics (sum|H-H_ls|^2, sum|H-H_mmse|^2, sum|H|^2 per rx and group) rounded to 2^-20 and "
                       "summed exactly in int64 over the shards, then all-reduced: equal across N iff every rank computes "
                       "the same slots"}

    # ---- end-to-end through the host-buffer API: params from pinned memory, arrays back to pinned memory ---------------
    e2e = None
    if not stats_only and not dense:
        e2e_B = min(args.e2e_batch, B)
        numa = None if args.no_numa_bind else bind_to_gpu_numa_node(local_rank)     # before the pinned buffers exist
        hp_ = HostPipeline(eng, pool, chunk=min(args.e2e_chunk, e2e_B), want=want, compact=not args.e2e_full, depth=args.e2e_depth)
        par = (mi[:e2e_B], np.asarray(W["dopplers"], np.float32)[di[:e2e_B]], snr_host[:e2e_B], pat[:e2e_B])
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        for i in range(2):
            hp_.run(*par, slot0=(rank * 100 + i) * e2e_B, seed=args.seed)
        barrier()
        e2e_total = world * e2e_steps * e2e_B
        mine = e2e_steps * e2e_B
        t0 = time.perf_counter()
        if world == 1:
            for i in range(e2e_steps):
                hp_.run(*par, slot0=(2 + i) * e2e_B, seed=args.seed)
        else:
            # the job's slots [0, e2e_total) are CLAIMED chunk by chunk (atomic add on the rendezvous store): the GPUs of a box
            # feed host memory at different rates, a fixed shard would leave the fast links idle while the slow ones finish
            from host_pipeline import store_claimer
            claim = store_claimer(dist.distributed_c10d._get_default_store(), e2e_total, "b2c_e2e_next")
            par_np = np.stack([np.asarray(v, np.float32) for v in par])

            def params_of(first, k):
                k = min(k, e2e_total - first)
                j = (first + np.arange(k)) % e2e_B
                return par_np[:, j]
            mine = hp_.run_dynamic(claim, params_of, seed=args.seed)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        barrier()
        link = d2h_probe(torch, dev, hp_.slab_bytes)           # all ranks at once: the link under the same contention
        barrier()
        link_min = link_sum = link
        if world > 1:
            t = torch.tensor([e2e_s, -link], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s, link_min = float(t[0].item()), -float(t[1].item())
            t = torch.tensor([link], dtype=torch.float64, device=dev)
            dist.all_reduce(t)
            link_sum = float(t.item())
        e2e_value = e2e_total / e2e_s
        if world > 1:
            t = torch.tensor([float(mine)], dtype=torch.float64, device=dev)
            g = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(g, t)
            claimed = [int(v.item()) for v in g]
            assert sum(claimed) == e2e_total, (claimed, e2e_total)
        gbs = e2e_value * hp_.d2h_bytes_per_slot / 1e9
        e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": hp_.h2d_bytes_per_slot * e2e_B,
               "d2h_bytes_per_step": hp_.d2h_bytes_per_slot * e2e_B, "slots_per_step": e2e_B, "steps": e2e_steps,
               "numa_node_rank0": numa,
               "slots_per_rank": (claimed if world > 1 else [mine]),
               "sharding": ("static" if world == 1 else "dynamic: ranks claim 128-slot chunks of the global index range from the "
                            "rendezvous store as their buffers free up (Philox keyed by global index: same arrays whoever makes them)"),
               "roofline": {"bound": "pcie d2h", "achieved": gbs, "peak": link_sum, "unit": "GB/s", "frac": gbs / link_sum,
                            "peak_source": f"bare pinned cudaMemcpyAsync D2H of one {hp_.slab_bytes / 1e6:.0f} MB chunk slab, back to back, "
                                           f"measured in this run on all {world} rank(s) at once (sum over ranks; slowest rank {link_min:.1f} GB/s)"},
               "note": ("HostPipeline: params from pinned host memory, arrays + stats back to pinned host memory, one cudaMemcpyAsync per "
                        f"{hp_.chunk}-slot chunk, {hp_.depth} chunks in flight (PCIe-bound); "
                        + ("full replicated arrays cross PCIe" if args.e2e_full else
                           "tx-replicated arrays (H_ls, H_mmse, tx) are written and cross PCIe once and are exposed as full-shape NumPy broadcast views"))}

    if rank == 0:
        peaks = measured_peaks()
        peak = peaks["hbm_gbs"]
        achieved = alg / (k_ms * 1e-3) / 1e9 if alg else 0.0
        traffic, traffic_note = None, "not captured for this workload / layout"
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                tj = json.load(fh)
            ent = tj.get(f"{args.workload}:{args.layout}" if args.layout != "full" else args.workload, {})
            if ent.get("dram_bytes_per_slot"):
                traffic = ent["dram_bytes_per_slot"] * B
                traffic_note = (f"static: dram__bytes_read + dram__bytes_write per slot from the ncu capture {ent.get('capture', '?')} "
                                f"(profiles/), x {B} slots; not re-measured in this run")
        kname = ("slot2_kernel<%d, EST, store-free> (register-blocked, statistics only)" % ntx if stats_only else
                 "slot_kernel<%d, ..., 599, FAST, pitch %s%s>" % (ntx, P, ", compact" if compact else ""))
        share = k_ms * lps * args.steps / ms
        if stats_only:
            roof = {"bound": "issue", "kernel": kname, "achieved": None, "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
                    "note": "statistics-only sweep: the kernel writes 192 B per slot; it is bound by instruction issue / the FMA and LSU "
                            "pipes (see profiles/), not by a memory roofline", "kernel_ms": k_ms, "kernel_share_of_step": share}
        else:
            roof = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_note, "peak_source": peaks["source"],
                    "algorithmic_bytes_per_launch": alg, "kernel_ms": k_ms,
                    "kernel_ms_source": "CUDA events on the launching stream around the kernel alone, mean of 12 launches over rotating "
                                        "buffers right after the timed region",
                    "kernel_share_of_step": share}
        if tensor is not None:
            tf32_peak = peaks["bf16_tflops"] / 2 if peaks["bf16_tflops"] else None
            tensor.update({"bound": "tensor", "kernel": "dense_tc_ta_kernel<grouped> (tcgen05.mma kind::tf32 with the data operand in tensor memory, "
                                                        "3xTF32 split; 8 SNR groups in one grid)",
                           "achieved": tensor["issued_tf32_tflops"], "peak": tf32_peak, "unit": "TFLOP/s",
                           "frac": tensor["issued_tf32_tflops"] / tf32_peak if tf32_peak else None,
                           "frac_of_measured_tf32_matmul": tensor["issued_tf32_tflops"] / tensor["tf32_peak_measured_tflops"],
                           "peak_source": "measured bf16 dense burst peak / 2 (kind::tf32 issues at half the bf16 rate; MEASURED_PEAKS.json holds "
                                          "no TF32 figure); cross-check: tf32_peak_measured_tflops = torch.matmul fp32 with TF32 allowed, 8192^3, "
                                          "best of 8, measured in this run",
                           "gemm_share_of_step": tensor["gemm_ms"] * lps * args.steps / ms})
            roof["tensor"] = tensor
        per_snr = {}
        for j, s in enumerate(SNRS):
            rows = nb[j::ns] if args.workload == "c4_sweep" else nb[j:j + 1]
            c = max(rows[:, 0].sum(), 1)
            per_snr[str(s)] = [float(10 * np.log10(rows[:, 3].sum() / c + 1e-12)), float(10 * np.log10(rows[:, 4].sum() / c + 1e-12))]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "arm": {"slots_per_launch": B, "launches_per_step": lps, "slots_per_step": B * lps, "slots_total": slots_total,
                    "timed_seconds": ms * 1e-3, "rng": "philox4x32-10 keyed by global slot index",
                    "pilot_patterns": f"pool of {len(pool)} scattered patterns ({ppd} per density, RandomState(42)) rotated over the slots, "
                                      f"up to {pool.np_max} pilots; plans built in {pool_build_s:.1f} s",
                    "parallelism": f"dp{world} (slots sharded, NCCL all-reduce of per-bin stats)",
                    "l2_policy": (f"outputs per launch = {alg / 1e9:.1f} GB >> 126 MB L2, {NBUF} rotating buffers; no flush needed" if alg else
                                  "no output arrays; tables (plans, twiddles) are meant to stay L2-resident"),
                    "layout": args.layout + (" (H_ls, H_mmse, tx written once; stride-0 views over tx)" if compact else
                                             " (reference shapes, tx-replicated arrays written ntx times)"),
                    "hbm_layout": (f"rows of 599 complex64 at pitch {P}" + (" (one padding element per row: 16-byte stores; "
                                   "algorithmic bytes count 599)" if P != 599 else " (contiguous)"))},
            "roofline": roof,
            "cpu_baseline": cpu,
            "e2e": e2e if e2e is not None else {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                                                "note": "no host-buffer leg for this workload (statistics-only sweep / dense pipeline)"},
            "value_api": value_api,
            "stats_checksum": checksum,
            "gpu_launches": timed_launches,
            "clocks": clocks,
            "per_snr_nmse_db": per_snr,
        }
        if args.workload == "c4_sweep":
            def cell(d, s, c):
                return float(nb[d * ns + s, c] / max(nb[d * ns + s, 0], 1))
            line["curves"] = {"density_pct": [round(100 * d) for d in W["densities"]], "snr_db": list(SNRS),
                              "nmse00_ls_db": [[float(10 * np.log10(cell(d, s, 8) + 1e-12)) for s in range(ns)] for d in range(nden)],
                              "ber_proxy_ls": [[cell(d, s, 12) for s in range(ns)] for d in range(nden)],
                              "ber_proxy_mmse": [[cell(d, s, 13) for s in range(ns)] for d in range(nden)]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 50; 6 for --impl reference)")
    ap.add_argument("--warmup", type=int, default=None, help="warm-up steps (default 3; 1 for --impl reference)")
    ap.add_argument("--min-seconds", type=float, default=0.6,
                    help="lower bound of the timed region: launches per step are sized in the warm-up to reach it")
    ap.add_argument("--launches-per-step", type=int, default=0, help="fix the launches per step instead of sizing them")
    ap.add_argument("--reference-budget", type=float, default=240.0,
                    help="--impl reference: wall-clock budget in seconds; the step loop stops early rather than overrun it")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3_4x4_etu", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None,
                    help="slots per launch per GPU.  Default 4144 (x 4 rx CTAs = 56 full waves of 296 resident CTAs on 148 SMs); "
                         "c2_2x2_eva_dense: 18944 (37888 slot CTAs = 128 full waves; 8 SNR groups x 37 column tiles x 7 W tile pairs = 14 full GEMM waves)")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--patterns", type=int, default=64, help="pilot patterns per density in the pool (the reference draws one per sample)")
    ap.add_argument("--pitch", type=int, default=600, choices=[599, 600],
                    help="row pitch of the output arrays in HBM: 600 = padded rows / 16-byte stores (default), 599 = contiguous")
    ap.add_argument("--layout", default="full", choices=["full", "compact"],
                    help="full: the five arrays in the reference's shapes; compact: each unique value written once")
    ap.add_argument("--e2e-batch", type=int, default=2048)
    ap.add_argument("--e2e-chunk", type=int, default=128)
    ap.add_argument("--e2e-depth", type=int, default=3)
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--e2e-full", action="store_true", help="copy the tx-replicated arrays in full instead of once")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin each rank to its GPU's NUMA node for the e2e leg")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-api", action="store_true", help="skip the value_api leg (sharded_statistics)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.batch is None:
        args.batch = WORKLOADS[args.workload].get("batch", 4144)
    if args.steps is None:
        args.steps = 50 if args.impl == "b200" else 6
    if args.warmup is None:
        args.warmup = 3 if args.impl == "b200" else 1
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3              # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
