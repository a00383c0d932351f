"""Batched host API over libb2c: device-resident tables, pattern pools and the slot pipeline.

This is the layer the reference-shaped shims (channel_simulator.py, baseline_estimators.py,
dataset_generator.py) and bench.py call.  Torch supplies device memory and streams; every
computation is a libb2c kernel.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

import _b2c
import _tables
from _b2c import DenseGroup, Geom, Inject, Patterns, PilotIO, Profiles, Slots, check, dptr, lib, ref, row_pitch, rows_ptr, stream_ptr

DEFAULT_MODELS = ("EPA", "EVA", "ETU")
BIN_FIELDS = ("count", "sum_mse_ls", "sum_mse_mmse", "sum_nmse_ls", "sum_nmse_mmse", "sum_nmse_ls_sq",
              "sum_nmse_mmse_sq", "sum_power", "sum_nmse00_ls", "sum_nmse00_ls_sq", "sum_nmse00_mmse",
              "sum_nmse00_mmse_sq", "sum_ber_proxy_ls", "sum_ber_proxy_mmse")


def _cuda_device(device=None):
    if not torch.cuda.is_available():
        raise _b2c.B2CError("no CUDA device: this package runs on sm_100a only and has no CPU fallback")
    return torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")


class PatternPool:
    """Device-resident pool of pilot patterns + interpolation plans (b2c_patterns)."""

    def __init__(self, pilot_index_list, nsym, nsc, method="linear", device=None):
        self.device = _cuda_device(device)
        self.nsym, self.nsc, self.method = nsym, nsc, method
        self.pilot_indices = [np.asarray(p, dtype=np.int64) for p in pilot_index_list]
        n = len(self.pilot_indices)
        self.npilots_host = np.array([len(p) for p in self.pilot_indices], dtype=np.int32)
        self.np_max = int(self.npilots_host.max())
        pre = np.zeros((n, self.np_max), dtype=np.int32)
        plans = np.zeros((n + 1, nsym * nsc + 1), dtype=_tables.PLAN_DTYPE)   # +1 pattern of padding: kernels prefetch one symbol ahead
        built = _tables.plans_for(self.pilot_indices, nsym, nsc, method)     # parallel + disk-cached for large pools
        for i, p in enumerate(self.pilot_indices):
            pre[i, :len(p)] = p
            plans[i] = _tables.finalize_plan(built[i], self.np_max)
        self.npilots = torch.from_numpy(self.npilots_host).to(self.device)
        self.pilot_re = torch.from_numpy(pre).to(self.device)
        self.plan = torch.from_numpy(plans.view(np.uint8).reshape(n + 1, (nsym * nsc + 1) * 16)).to(self.device)
        self.struct = Patterns(n, self.np_max, self.npilots.data_ptr(), self.pilot_re.data_ptr(), self.plan.data_ptr())

    def __len__(self):
        return len(self.pilot_indices)

    def mask(self, i):
        m = np.zeros(self.nsym * self.nsc, dtype=bool)
        m[self.pilot_indices[i]] = True
        return m.reshape(self.nsym, self.nsc)


class PreparedDense:
    """A GEMM operand in b2c_dense_prepare's form (see SlotEngine.prepare_dense)."""

    def __init__(self, buf, m, k, is_complex):
        self.buf, self.m, self.k, self.is_complex = buf, int(m), int(k), bool(is_complex)


class WienerBank:
    """Known-covariance MMSE filters of a batch: W = R (R + sigma^2 I)^-1 per (pilot pattern, SNR), built ONCE in
    float64 on the host (src/baseline_estimators.py:174-190, noise_variance = 1 / snr_linear) and kept on the device
    as prepared tensor-core operands (b2c_dense_prepare).  SlotEngine.run(mmse="dense", wiener=bank) then applies
    each W to all pilot vectors of its (pattern, SNR) group with one GEMM.

    covariances: {pattern_id: R [np, np] complex} (np = that pattern's pilot count); snrs: the SNR values (dB) the
    batch may contain."""

    def __init__(self, engine, pool, covariances, snrs):
        self.pool = pool
        self.prepared, self.W_host = {}, {}
        for pid, R in covariances.items():
            R = np.asarray(R)
            n = int(pool.npilots_host[pid])
            if R.shape != (n, n):
                raise ValueError(f"covariance of pattern {pid} has shape {R.shape}, its pilot set has {n} pilots")
            for snr in snrs:
                sigma2 = 1.0 / (10.0 ** (float(snr) / 10.0))
                Ry = R + sigma2 * np.eye(n)
                try:
                    W = R @ np.linalg.inv(Ry)
                except np.linalg.LinAlgError:                       # the reference's regularised fallback (:191-194)
                    W = R @ np.linalg.inv(Ry + 1e-6 * np.eye(n))
                key = (int(pid), float(snr))
                self.W_host[key] = W
                self.prepared[key] = engine.prepare_dense(torch.from_numpy(np.ascontiguousarray(W.astype(np.complex64))).to(engine.device))

    def groups(self, pattern_id, snr_db):
        """Host-side grouping of a batch: (order, starts, keys) with `order` the slot indices sorted by
        (pattern, SNR) group and group g = order[starts[g]:starts[g+1]] using the filter keys[g]."""
        pid = np.asarray(pattern_id).astype(np.int64)
        snr = np.asarray(snr_db, dtype=np.float64)
        snr_vals, snr_idx = np.unique(snr, return_inverse=True)               # 1-D uniques: cheap on the host path of every batch
        code, inv = np.unique(pid * len(snr_vals) + snr_idx.reshape(-1), return_inverse=True)
        inv = inv.reshape(-1)
        keys = [(int(c // len(snr_vals)), float(snr_vals[c % len(snr_vals)])) for c in code]
        for k in keys:
            if k not in self.prepared:
                raise KeyError(f"no Wiener matrix for pattern {k[0]} at {k[1]} dB in this bank")
        order = np.argsort(inv, kind="stable")
        starts = np.concatenate([[0], np.cumsum(np.bincount(inv, minlength=len(keys)))])
        return order, starts, keys

    def plan_batch(self, engine, pattern_id, snr_db, B):
        """Grouping of one batch as the kernels consume it: the device column map (slot b's pilot vectors live in rows
        col[b] .. col[b] + nrx - 1 of the grouped pilot matrix) and the per-group GEMM descriptors.  A plan depends only on
        the batch's (pattern, SNR) values: build it once and pass it to every SlotEngine.run(dense_plan=...) that
        repeats them (the grouping is host work: np.unique + argsort + one upload)."""
        order, starts, keys = self.groups(np.broadcast_to(np.asarray(pattern_id), (B,)), np.broadcast_to(np.asarray(snr_db), (B,)))
        rank = np.empty(B, dtype=np.int64)
        rank[order] = np.arange(B)
        nrx = engine.nrx
        plan = DenseBatchPlan()
        plan.B, plan.keys = B, keys
        plan.col = torch.from_numpy((rank * nrx).astype(np.int32)).pin_memory().to(engine.device, non_blocking=True)
        plan.groups = (DenseGroup * len(keys))()
        for gi, key in enumerate(keys):
            W = self.prepared[key]
            plan.groups[gi] = DenseGroup(W.buf.data_ptr(), int(starts[gi]) * nrx, int(starts[gi + 1] - starts[gi]) * nrx, W.m)
        plan.keepalive = [self.prepared[k] for k in keys]
        return plan


class DenseBatchPlan:
    """See WienerBank.plan_batch."""


class SlotEngine:
    """Geometry + TDL profile tables on one GPU, and launchers for every libb2c entry point."""

    def __init__(self, config, models=DEFAULT_MODELS, device=None):
        self.device = _cuda_device(device)
        o, m = config["ofdm"], config["mimo"]
        self.fft_size, self.cp = int(o["fft_size"]), int(o["cp_length"])
        self.nsym, self.useful = int(o["num_symbols"]), int(o["useful_subcarriers"])
        self.ntx, self.nrx = int(m["num_tx_antennas"]), int(m["num_rx_antennas"])
        self.sampling_rate = self.fft_size * float(o["subcarrier_spacing"])
        self.used = _tables.used_subcarriers(self.fft_size, self.useful)
        self.nsc = len(self.used)
        self.models = tuple(models)
        self.geom = Geom(self.nsym, self.nsc, self.ntx, self.nrx, self.fft_size, self.cp,
                         (self.fft_size + self.cp) / self.sampling_rate)
        t = _tables.profile_tables(self.models, self.sampling_rate, self.fft_size, self.used)
        self.host_tables = t
        self._dev = {k: torch.from_numpy(np.ascontiguousarray(v)).to(self.device) for k, v in t.items()}
        self.prof = Profiles(len(self.models), self._dev["ntaps"].data_ptr(), self._dev["npaths"].data_ptr(),
                             self._dev["tap_path"].data_ptr(), self._dev["tap_amp"].data_ptr(),
                             self._dev["tap_tw"].data_ptr(), self._dev["tap_corr"].data_ptr(), self._dev["tap_delay"].data_ptr())
        self.p_max = int(t["npaths"].max())
        self._cubic, self._ident = {}, {}

    # ---- pools -------------------------------------------------------------------------------
    def pool(self, pilot_index_list, method="linear"):
        return PatternPool(pilot_index_list, self.nsym, self.nsc, method, self.device)

    def random_pool(self, densities, per_density=1, seed=42, method="linear"):
        """Fixed pool of scattered patterns (PilotPattern's rule, src/channel_simulator.py:223-229)
        drawn from numpy RandomState(seed): pattern id = density_index * per_density + j."""
        rs = np.random.RandomState(seed)
        total = self.nsym * self.nsc
        pats = []
        for d in densities:
            for _ in range(per_density):
                perm = np.arange(total)
                rs.shuffle(perm)
                pats.append(np.sort(perm[:int(total * d)]))
        return self.pool(pats, method)

    # ---- helpers -------------------------------------------------------------------------------
    _NP = {torch.int32: np.int32, torch.float32: np.float32}

    def _vec(self, v, B, dtype):
        """Per-slot parameter vector on the device.  Host values go through a pinned staging tensor and an ASYNCHRONOUS
        copy: a pageable-memory copy would make the host wait for every kernel already queued on the stream, i.e.
        serialise the host's preparation of batch i+1 with the GPU's work on batch i (torch's pinned-memory cache
        recycles the staging block only after the copy has completed)."""
        if isinstance(v, torch.Tensor):
            return v.to(device=self.device, dtype=dtype).contiguous()
        a = np.array(np.broadcast_to(np.asarray(v), (B,)), dtype=self._NP.get(dtype))     # typed, writable copy
        t = torch.from_numpy(a)
        if t.numel() and t.dtype == dtype:
            return t.pin_memory().to(device=self.device, non_blocking=True)
        return t.to(device=self.device, dtype=dtype)

    def _slots(self, B, model_id, doppler_hz, snr_db, pattern_id, slot0, seed, qpsk=False):
        keep = (self._vec(model_id, B, torch.int32), self._vec(doppler_hz, B, torch.float32),
                self._vec(snr_db, B, torch.float32), self._vec(pattern_id, B, torch.int32))
        s = Slots(int(slot0), int(seed) & 0xFFFFFFFFFFFFFFFF, *(k.data_ptr() for k in keep), 1 if qpsk else 0)
        return s, keep

    def _inject(self, inject):
        if inject is None:
            return None, ()
        keep = []
        ij = Inject()
        if inject.get("jakes_u") is not None:
            ju = inject["jakes_u"]
            ij.jakes_u = dptr(ju, "f32").value
            ij.p_max = ju.shape[1]
            keep.append(ju)
        if inject.get("sym_turns") is not None:
            ij.sym_turns = dptr(inject["sym_turns"], "f32").value
            ij.noise = dptr(inject["noise"], "c64").value
            keep += [inject["sym_turns"], inject["noise"]]
        return ij, keep

    def workspace(self, B):
        return {"gains": torch.empty((B, self.nrx, self.nsym, self.ntx, _b2c.MAX_TAPS), dtype=torch.complex64, device=self.device),
                "noise_std": torch.empty((B,), dtype=torch.float32, device=self.device)}

    @staticmethod
    def _with_pitch(g, pitch):
        return Geom(g.nsym, g.nsc, g.ntx, g.nrx, g.fft_size, g.cp_length, g.symbol_period_s, 0 if pitch == g.nsc else int(pitch))

    def alloc_outputs(self, B, want, compact=False, pitch=None):
        """Output buffers.  compact=True: the tx-replicated arrays (H_ls, H_mmse, tx) hold one copy.
        pitch=_b2c.WIDE_PITCH (600): rows padded by one element -- the layout of the wide-store kernel; the
        tensors returned are [..., :nsc] views of the padded buffers."""
        P = self.nsc if pitch is None else int(pitch)
        full = (B, self.nsym, self.nrx, self.ntx, P)
        est = (B, self.nsym, self.nrx, P) if compact else full
        shapes = {"H_true": full, "H_ls": est, "H_mmse": est, "rx": (B, self.nsym, self.nrx, P),
                  "tx": (B, self.nsym, P) if compact else (B, self.nsym, self.ntx, P)}
        out = {k: torch.empty(shapes[k], dtype=torch.complex64, device=self.device)[..., :self.nsc]
               for k in want if k in shapes}
        if "stats" in want:
            out["stats"] = torch.empty((B, self.nrx, 2, _b2c.N_STAT), dtype=torch.float64, device=self.device)
        return out

    # ---- K1a + fused slot kernel -------------------------------------------------------------------
    def run(self, B, model_id, doppler_hz, snr_db, pattern_id=0, pool=None, slot0=0, seed=42, inject=None,
            want=("H_true", "rx", "tx", "H_ls", "H_mmse", "stats"), out=None, ws=None, compact=False, pitch=None,
            mmse="default", wiener=None, dense_plan=None, qpsk=False):
        """Simulate B slots and (if any of H_ls/H_mmse/stats is wanted) estimate them.
        Per-slot parameters are scalars or length-B arrays/tensors.  Returns dict of CUDA tensors.
        compact=True writes the tx-replicated arrays once (see alloc_outputs); expand_compact() turns
        them into full-shape stride-0 views.  pitch=600 (or `out` from alloc_outputs(pitch=600)) selects the
        padded-row layout of the wide-store kernel (throughput configuration only, see include/b2c.h).
        mmse="default": MMSEEstimator()'s alpha * LS (src/baseline_estimators.py:177-180), fused in the slot kernel.
        mmse="dense" + wiener=WienerBank: the known-covariance branch (:181-190) for the whole batch -- the slot
        kernel also hands out h_ls at the pilots grouped by (pattern, SNR), one tensor-core GEMM per group applies
        W, and K3 interpolates the filtered pilots into H_mmse and scores them.  The grouping is decided on the host:
        either pattern_id / snr_db are host values, or dense_plan = wiener.plan_batch(...) built once for batches that
        repeat the same (pattern, SNR) values.
        qpsk=True (Philox draws only): every resource element carries a QPSK point instead of a uniform phase, i.e. two
        payload bits (b2c_slots.qpsk) -- the grid of the BER sweeps (ber_batch)."""
        if mmse not in ("default", "dense"):
            raise ValueError(f"Unknown mmse mode: {mmse}")
        dense = mmse == "dense" and any(k in (out if out is not None else want) for k in ("H_mmse", "stats"))
        score_only = False
        if dense:
            if wiener is None or pool is None:
                raise ValueError('mmse="dense" needs a WienerBank and its PatternPool')
            if compact:
                raise ValueError('mmse="dense" writes H_mmse in the full layout')
            if dense_plan is None and (isinstance(pattern_id, torch.Tensor) or isinstance(snr_db, torch.Tensor)):
                raise ValueError('mmse="dense" groups slots by (pattern, SNR) on the host: pass host values or a dense_plan')
            req = out if out is not None else want
            # statistics only: the array-free form (slot kernel -> GEMM -> b2c_dense_score); nothing but the pilot vectors
            # and the per-slot sums is written to HBM
            score_only = not any(k in req for k in ("H_true", "rx", "tx", "H_ls", "H_mmse")) and self.nsc == 599 \
                and self.nsym % 2 == 0 and self.ntx in (1, 2, 4, 8) and inject is None
            if out is None and not score_only:
                want = tuple(want) + tuple(k for k in ("H_true", "H_ls") if k not in want)   # K3 scores against H_true
        if out is None:
            out = self.alloc_outputs(B, want, compact, pitch)
        if ws is None:
            ws = self.workspace(B)
        est = any(k in out for k in ("H_ls", "H_mmse", "stats"))
        if est and pool is None:
            raise ValueError("estimation outputs need a PatternPool")
        if B == 0:
            return out
        if qpsk and inject is not None and inject.get("sym_turns") is not None:
            raise ValueError("qpsk selects the Philox grid; with injected symbols inject the QPSK phases themselves")
        slots, keep = self._slots(B, model_id, doppler_hz, snr_db, pattern_id, slot0, seed, qpsk)
        ij, keep_inj = self._inject(inject)
        L = lib()
        arrays = [out[k] for k in ("H_true", "rx", "tx", "H_ls", "H_mmse") if k in out]
        P = row_pitch(arrays[0]) if arrays else self.nsc
        g = self._with_pitch(self.geom, P)
        pio = None
        if dense:
            if "H_true" not in out and not score_only:
                raise ValueError('mmse="dense" needs H_true among the outputs (the MMSE error sums are taken against it)')
            if dense_plan is None:
                dense_plan = wiener.plan_batch(self, pattern_id, snr_db, B)
            if dense_plan.B != B:
                raise ValueError(f"dense_plan was built for {dense_plan.B} slots, this batch has {B}")
            col = dense_plan.col
            ld = pool.np_max
            hp = ws.get("hp")
            if hp is None or hp.shape[0] < B * self.nrx or hp.shape[1] != ld:
                hp = ws["hp"] = torch.empty((B * self.nrx, ld), dtype=torch.complex64, device=self.device)
                ws["hm"] = torch.empty_like(hp)
            pio = PilotIO(hp.data_ptr(), col.data_ptr(), ld)
        check(L.b2c_tap_gains(ref(self.geom), ref(self.prof), ref(slots), ref(ij), B,
                              dptr(ws["gains"], "c64"), dptr(ws["noise_std"], "f32"), stream_ptr()), "b2c_tap_gains")
        check(L.b2c_slot_pipeline(ref(g), ref(self.prof), ref(pool.struct) if pool is not None else None,
                                  ref(slots), ref(ij), B, dptr(ws["gains"], "c64"), dptr(ws["noise_std"], "f32"),
                                  rows_ptr(out.get("H_true"), P, True), rows_ptr(out.get("rx"), P, True),
                                  rows_ptr(out.get("tx"), P, True), rows_ptr(out.get("H_ls"), P, True),
                                  None if dense else rows_ptr(out.get("H_mmse"), P, True), dptr(out.get("stats"), "f64", True),
                                  1 if compact else 0, ref(pio), stream_ptr()), "b2c_slot_pipeline")
        if dense:
            hm = ws["hm"]
            ng = len(dense_plan.groups)
            for g0 in range(0, ng, 32):                           # every (pattern, SNR) group's GEMM in one launch (32 groups per call)
                n = min(32, ng - g0)
                check(L.b2c_dense_apply_grouped(C.byref(dense_plan.groups, g0 * C.sizeof(DenseGroup)), n, dptr(hp, "c64"), dptr(hm, "c64"),
                                                ld, stream_ptr()), "b2c_dense_apply_grouped")
            if score_only:
                # second pass of the statistics-only form: true CFR regenerated from the gains, filtered pilots interpolated,
                # MMSE error sums into stats[..., 1]
                check(L.b2c_dense_score(ref(self.geom), ref(self.prof), ref(pool.struct), ref(slots), B, dptr(ws["gains"], "c64"),
                                        dptr(hm, "c64"), dptr(col, "i32"), ld, dptr(out["stats"], "f64"), stream_ptr()),
                      "b2c_dense_score")
            else:
                # K3, mode 2: interpolate the filtered pilots into H_mmse, MMSE error sums into stats[..., 1]
                check(L.b2c_ls_interp(ref(g), ref(pool.struct), keep[3].data_ptr(), None, B, None, None, 0, dptr(hm, "c64"), 2,
                                      rows_ptr(out["H_true"], P) if "stats" in out else None, None,
                                      rows_ptr(out.get("H_mmse"), P, True), None, dptr(out.get("stats"), "f64", True),
                                      dptr(col, "i32"), ld, stream_ptr()), "b2c_ls_interp")
            keep = keep + (dense_plan,)
        out["_keepalive"] = (keep, keep_inj, ws)
        return out

    def expand_compact(self, out):
        """Full-shape (stride-0 over tx) views of compact outputs; works for torch tensors and numpy arrays."""
        res = dict(out)
        for k in ("H_ls", "H_mmse"):
            if k in out and out[k].ndim == 4:
                v = out[k][:, :, :, None, :]
                shape = v.shape[:3] + (self.ntx,) + v.shape[4:]
                res[k] = v.expand(*shape) if isinstance(v, torch.Tensor) else np.broadcast_to(v, shape)
        if "tx" in out and out["tx"].ndim == 3:
            v = out["tx"][:, :, None, :]
            shape = v.shape[:2] + (self.ntx,) + v.shape[3:]
            res["tx"] = v.expand(*shape) if isinstance(v, torch.Tensor) else np.broadcast_to(v, shape)
        return res

    # ---- K3 ------------------------------------------------------------------------------------------
    def ls_interp(self, rx, pilots, pool, pattern_id=0, snr_db=None, mmse=False, H_true=None, hp_in=None,
                  want=("H_ls",), geom=None, out=None, pitch=None):
        """rx [B,nsym,nrx,nsc] c64, pilots [B or 1, np_max] c64.  want subset of H_ls, H_mmse, hp, stats.
        If rx (and H_true) are padded-row views (pitch 600, as run(pitch=600) returns them) the outputs are
        padded the same way and the wide-access kernel runs."""
        g = geom if geom is not None else self.geom
        B = rx.shape[0] if rx is not None else hp_in.shape[0]
        P = row_pitch(rx) if rx is not None else (g.nsc if pitch is None else int(pitch))
        g = self._with_pitch(g, P)
        pid = self._vec(pattern_id, B, torch.int32)
        snr = self._vec(snr_db, B, torch.float32) if snr_db is not None else None
        out = dict(out) if out is not None else {}          # caller-supplied H_ls / H_mmse buffers (same row pitch as rx)
        shape = (B, g.nsym, g.nrx, g.ntx, P)
        if "H_ls" in want and "H_ls" not in out:
            out["H_ls"] = torch.empty(shape, dtype=torch.complex64, device=self.device)[..., :g.nsc]
        if "H_mmse" in want and "H_mmse" not in out:
            out["H_mmse"] = torch.empty(shape, dtype=torch.complex64, device=self.device)[..., :g.nsc]
        if "hp" in want:
            out["hp"] = torch.zeros((B, g.nrx, pool.np_max), dtype=torch.complex64, device=self.device)
        if "stats" in want:
            out["stats"] = torch.empty((B, g.nrx, 2, _b2c.N_STAT), dtype=torch.float64, device=self.device)
        stride = pilots.shape[1] if (pilots is not None and pilots.shape[0] > 1) else 0
        check(lib().b2c_ls_interp(ref(g), ref(pool.struct), dptr(pid, "i32"), dptr(snr, "f32", True), B,
                                  rows_ptr(rx, P, True), dptr(pilots, "c64", True), stride, dptr(hp_in, "c64", True),
                                  1 if mmse else 0, rows_ptr(H_true, P, True), rows_ptr(out.get("H_ls"), P, True),
                                  rows_ptr(out.get("H_mmse"), P, True), dptr(out.get("hp"), "c64", True),
                                  dptr(out.get("stats"), "f64", True), None, 0, stream_ptr()), "b2c_ls_interp")
        return out

    def pilot_vectors(self, y, x, snr_db=0.0, mmse=False):
        """y [nvec, n], x [n] complex64 CUDA -> alpha * y / (x + 1e-12)."""
        out = torch.empty_like(y)
        check(lib().b2c_pilot_vectors(dptr(y, "c64"), dptr(x, "c64"), y.shape[0], y.shape[1], float(snr_db),
                                      1 if mmse else 0, dptr(out, "c64"), stream_ptr()), "b2c_pilot_vectors")
        return out

    def dense_real_apply(self, W, h, ld_out=None, out=None):
        """W [m,k] float32 (or a PreparedDense of one), h [ncols, ld_in>=k] c64 -> out [ncols, ld_out>=m] with
        out[c,:m] = W @ h[c,:k] (columns m.. of a caller-supplied `out` are left as they are)."""
        m, k = (W.m, W.k) if isinstance(W, PreparedDense) else W.shape
        ld_out = (m if ld_out is None else ld_out) if out is None else out.shape[1]
        if out is None:
            out = torch.zeros((h.shape[0], ld_out), dtype=torch.complex64, device=self.device)
        if isinstance(W, PreparedDense):
            if W.is_complex:
                raise ValueError("prepared operand is a complex Wiener matrix, not a real map")
            check(lib().b2c_dense_apply_prepared(dptr(W.buf, "u8"), m, k, 0, dptr(h, "c64"), dptr(out, "c64"), h.shape[0],
                                                 h.shape[1], ld_out, stream_ptr()), "b2c_dense_apply_prepared")
            return out
        check(lib().b2c_dense_real_apply(dptr(W, "f32"), m, k, dptr(h, "c64"), dptr(out, "c64"), h.shape[0], h.shape[1],
                                         ld_out, stream_ptr()), "b2c_dense_real_apply")
        return out

    def prepare_dense(self, W):
        """One-time preparation of a GEMM operand that is applied many times (a Wiener matrix W [np,np] complex64, or
        a real interpolation map [m,k] float32): TF32 hi / lo split tiles in the kernel's shared-memory layout, so
        that every K stage of it is one bulk copy (b2c_dense_prepare)."""
        cplx = W.is_complex()
        m, k = W.shape
        nbytes = lib().b2c_dense_prepared_bytes(m, k, int(cplx))
        buf = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
        check(lib().b2c_dense_prepare(dptr(W, "c64" if cplx else "f32"), m, k, int(cplx), dptr(buf, "u8"), stream_ptr()),
              "b2c_dense_prepare")
        return PreparedDense(buf, m, k, cplx)

    def ls_cubic(self, rx, pilots, pilot_indices, H_true=None, want=("H_ls",), geom=None):
        """LS estimate with griddata's 'cubic' interpolation: LS at the pilots (K3), the dense
        Clough-Tocher map on the tensor cores (K4b), then K3 again with an identity plan to lay the
        grid out per (rx, tx) and score it.  rx [B,nsym,nrx,nsc], pilots [B or 1, Np].  (The reference's
        MMSE estimator always interpolates linearly, src/baseline_estimators.py:200, so there is no
        cubic MMSE.)"""
        g = geom if geom is not None else self.geom
        idx = np.asarray(pilot_indices, dtype=np.int64)
        nre = g.nsym * g.nsc
        pool = PatternPool([idx], g.nsym, g.nsc, "nearest", self.device)          # only its pilot list is used here
        B = rx.shape[0]
        hp = self.ls_interp(rx, pilots, pool, want=("hp",), geom=g)["hp"].reshape(B * g.nrx, -1)
        key = (idx.tobytes(), g.nsym, g.nsc)
        if key not in self._cubic:     # the map is applied to every batch of this pattern: keep it in prepared form
            self._cubic = {key: self.prepare_dense(torch.from_numpy(_tables.cubic_matrix(idx, g.nsym, g.nsc)).to(self.device))}
        grid = self.dense_real_apply(self._cubic[key], hp, ld_out=nre)             # [B*nrx, nre]
        return self.ls_interp(None, None, self._identity_pool(g.nsym, g.nsc), hp_in=grid.reshape(B, g.nrx, nre),
                              H_true=H_true, want=tuple(w for w in want if w in ("H_ls", "stats")), geom=g)

    def _identity_pool(self, nsym, nsc):
        nre = nsym * nsc
        if nre not in self._ident:
            pool = PatternPool.__new__(PatternPool)
            pool.device, pool.nsym, pool.nsc, pool.method = self.device, nsym, nsc, "identity"
            pool.pilot_indices, pool.np_max = [np.arange(nre)], nre
            pool.npilots_host = np.array([nre], np.int32)
            plan = np.zeros((2, nre + 1), dtype=_tables.PLAN_DTYPE)
            plan[0] = _tables.finalize_plan(_tables.identity_plan(nre), nre)
            pool.npilots = torch.tensor([nre], dtype=torch.int32, device=self.device)
            pool.pilot_re = torch.arange(nre, dtype=torch.int32, device=self.device).reshape(1, nre)
            pool.plan = torch.from_numpy(plan.view(np.uint8).reshape(2, (nre + 1) * 16)).to(self.device)
            pool.struct = Patterns(1, nre, pool.npilots.data_ptr(), pool.pilot_re.data_ptr(), pool.plan.data_ptr())
            self._ident[nre] = pool
        return self._ident[nre]

    def mmse_dense(self, W, h, out=None):
        """W [np,np] c64 (or a PreparedDense of one), h [ncols, ld>=np] c64 -> W @ h[c, :np] per column set."""
        if out is None:
            out = torch.zeros_like(h)
        if isinstance(W, PreparedDense):
            if not W.is_complex:
                raise ValueError("prepared operand is a real map, not a complex Wiener matrix")
            check(lib().b2c_dense_apply_prepared(dptr(W.buf, "u8"), W.m, W.k, 1, dptr(h, "c64"), dptr(out, "c64"), h.shape[0],
                                                 h.shape[1], h.shape[1], stream_ptr()), "b2c_dense_apply_prepared")
            return out
        check(lib().b2c_mmse_dense(dptr(W, "c64"), W.shape[0], dptr(h, "c64"), dptr(out, "c64"), h.shape[0],
                                   h.shape[1], stream_ptr()), "b2c_mmse_dense")
        return out

    # ---- K5 ------------------------------------------------------------------------------------------
    def stats_bins(self, stats, bin_id, nbins, bins=None, geom=None, snr_db=None):
        """Fold per-slot statistics into [nbins, 14] float64 accumulators (BIN_FIELDS).  snr_db (per slot) also folds
        the QPSK BER proxy of run_phase5_evaluation.py:57-68 (fields 12, 13)."""
        g = geom if geom is not None else self.geom
        B = stats.shape[0]
        if bins is None:
            bins = torch.zeros((nbins, _b2c.N_BINSTAT), dtype=torch.float64, device=self.device)
        bid = self._vec(bin_id, B, torch.int32)
        snr = None if snr_db is None else self._vec(snr_db, B, torch.float32)
        check(lib().b2c_stats_bins(ref(g), dptr(stats, "f64"), dptr(bid, "i32"), dptr(snr, "f32", True), B, nbins, dptr(bins, "f64"),
                                   stream_ptr()), "b2c_stats_bins")
        return bins

    # ---- K2 ------------------------------------------------------------------------------------------
    def ofdm_modulate(self, sym):
        rows = sym.shape[0]
        out = torch.empty((rows, self.fft_size + self.cp), dtype=torch.complex64, device=self.device)
        check(lib().b2c_ofdm_modulate(ref(self.geom), dptr(sym, "c64"), dptr(out, "c64"), rows, stream_ptr()),
              "b2c_ofdm_modulate")
        return out

    def ofdm_demodulate(self, sig):
        rows = sig.shape[0]
        out = torch.empty((rows, self.nsc), dtype=torch.complex64, device=self.device)
        check(lib().b2c_ofdm_demodulate(ref(self.geom), dptr(sig, "c64"), dptr(out, "c64"), rows, stream_ptr()),
              "b2c_ofdm_demodulate")
        return out

    # ---- time-domain statement of the slot (north_star kernels 1 / 2 as written) -------------------------------------
    def time_domain_slot(self, B, model_id, doppler_hz, tx=None, slot0=0, seed=42, inject=None):
        """OFDM-modulate the transmit grids, convolve every symbol body circularly with that symbol's tap gains (K1a,
        the very gains the fused kernel uses) and demodulate: the received grid of the frequency-domain pipeline
        WITHOUT noise, obtained through src/channel_simulator.py:150-203's modem (K2) and a time-domain TDL.
        tx [B, nsym, ntx, nsc] complex64 CUDA (default: the slot pipeline's own Philox grid).  Returns dict with
        rx [B, nsym, nrx, nsc], x_time / y_time [B, nsym, ant, fft + cp] and the tx used."""
        ws = self.workspace(B)
        slots, keep = self._slots(B, model_id, doppler_hz, 300.0, 0, slot0, seed)
        ij, keep_inj = self._inject(inject)
        L = lib()
        check(L.b2c_tap_gains(ref(self.geom), ref(self.prof), ref(slots), ref(ij), B, dptr(ws["gains"], "c64"),
                              dptr(ws["noise_std"], "f32"), stream_ptr()), "b2c_tap_gains")
        if tx is None:
            tx = self.run(B, model_id, doppler_hz, 300.0, slot0=slot0, seed=seed, inject=inject, want=("H_true", "rx", "tx"))["tx"]
        tx = tx.contiguous()
        x_time = self.ofdm_modulate(tx.reshape(B * self.nsym * self.ntx, self.nsc))
        y_time = torch.empty((B * self.nsym * self.nrx, self.fft_size + self.cp), dtype=torch.complex64, device=self.device)
        check(L.b2c_tdl_circular(ref(self.geom), ref(self.prof), keep[0].data_ptr(), B, dptr(ws["gains"], "c64"), dptr(x_time, "c64"),
                                 dptr(y_time, "c64"), stream_ptr()), "b2c_tdl_circular")
        rx = self.ofdm_demodulate(y_time).reshape(B, self.nsym, self.nrx, self.nsc)
        T = self.fft_size + self.cp
        return {"rx": rx, "tx": tx, "x_time": x_time.reshape(B, self.nsym, self.ntx, T), "y_time": y_time.reshape(B, self.nsym, self.nrx, T),
                "_keepalive": (keep, keep_inj, ws)}

    # ---- a8 / a3 stand-alone -----------------------------------------------------------------------------
    def apply_channel(self, tx, H, snr_db, noise=None, slot0=0, seed=42, geom=None):
        g = geom if geom is not None else self.geom
        B = tx.shape[0]
        slots, keep = self._slots(B, 0, 0.0, snr_db, 0, slot0, seed)
        ij, keep_inj = (None, ())
        if noise is not None:
            ij = Inject()
            ij.noise = dptr(noise, "c64").value
        rx = torch.empty((B, g.nsym, g.nrx, g.nsc), dtype=torch.complex64, device=self.device)
        scratch = torch.empty((B,), dtype=torch.float64, device=self.device)
        check(lib().b2c_apply_channel(ref(g), ref(slots), ref(ij), B, dptr(tx, "c64"), dptr(H, "c64"),
                                      dptr(rx, "c64"), dptr(scratch, "f64"), stream_ptr()), "b2c_apply_channel")
        return rx

    def tdl_full(self, model, doppler_hz, num_samples, ntx, nrx, jakes_u=None, seed=42, slot=0):
        m = self.models.index(model)
        nt = int(self.host_tables["ntaps"][m])
        delays = np.ascontiguousarray(self.host_tables["tap_delay"][m, :nt].astype(np.int32))
        L = int(delays.max()) + 1
        g = Geom(self.nsym, self.nsc, ntx, nrx, self.fft_size, self.cp, self.geom.symbol_period_s)
        out = torch.empty((num_samples, nrx, ntx, L), dtype=torch.complex64, device=self.device)
        check(lib().b2c_tdl_full(ref(g), ref(self.prof), m, float(doppler_hz), 1.0 / self.sampling_rate, num_samples, L,
                                 delays.ctypes.data_as(C.c_void_p), nt, dptr(jakes_u, "f32", True), int(seed), int(slot),
                                 dptr(out, "c64"), stream_ptr()), "b2c_tdl_full")
        return out

    # ---- next rows: equaliser, QAM, BER, ML features (b2c_link.cu) -------------------------------------
    def equalize(self, rx, H, method="zf"):
        """rx [B][nsym][nrx][nsc], H [B][nsym][nrx][ntx][nsc] (complex64 or complex128) -> [B][nsym][ntx][nsc]."""
        if method not in ("zf", "mmse"):
            raise ValueError(f"Unknown equalization method: {method}")
        lam = 1e-8 if method == "zf" else 0.01      # src/baseline_estimators.py:297, 305
        B, nsym, nrx, ntx, nsc = H.shape
        kind = "c128" if H.dtype == torch.complex128 else "c64"
        g = Geom(nsym, nsc, ntx, nrx, self.fft_size, self.cp, 0.0)
        out = torch.empty((B, nsym, ntx, nsc), dtype=H.dtype, device=self.device)
        check(lib().b2c_equalize(ref(g), B, dptr(rx, kind), dptr(H, kind), dptr(out, kind), lam, int(kind == "c128"),
                                 stream_ptr()), "b2c_equalize")
        return out

    @staticmethod
    def _qam_bits(M, what):
        if M not in (4, 16):
            raise NotImplementedError(f"{what} order {M} not implemented")
        return 2 if M == 4 else 4

    def qam_modulate(self, bits, M=4):
        """bits uint8 [n * log2 M] (0/1, MSB first) -> complex64 [n]."""
        bps = self._qam_bits(M, "Modulation")
        n = bits.numel() // bps
        out = torch.empty((n,), dtype=torch.complex64, device=self.device)
        if n:
            check(lib().b2c_qam_modulate(dptr(bits, "u8"), n, M, dptr(out, "c64"), stream_ptr()), "b2c_qam_modulate")
        return out

    def qam_demodulate(self, symbols, M=4):
        bps = self._qam_bits(M, "Demodulation")
        n = symbols.numel()
        kind = "c128" if symbols.dtype == torch.complex128 else "c64"
        bits = torch.empty((n * bps,), dtype=torch.uint8, device=self.device)
        if n:
            check(lib().b2c_qam_demodulate(dptr(symbols, kind), n, M, int(kind == "c128"), dptr(bits, "u8"), stream_ptr()),
                  "b2c_qam_demodulate")
        return bits

    def count_bit_errors(self, a, b, count=None):
        """Accumulates #{a != b} into the int64 device scalar `count` (created if None)."""
        if a.numel() != b.numel():
            raise ValueError("bit arrays differ in length")
        if count is None:
            count = torch.zeros((1,), dtype=torch.int64, device=self.device)
        if a.numel():
            check(lib().b2c_count_bit_errors(dptr(a, "u8"), dptr(b, "u8"), a.numel(), dptr(count, "i64"), stream_ptr()),
                  "b2c_count_bit_errors")
        return count

    def bit_errors_per_slot(self, bits_a, bits_b, pool, pattern_id, B, bps=2, geom=None):
        """int32 [B]: differing bits per slot on the DATA resource elements (pilots of the slot's pattern excluded).
        bits_a / bits_b: uint8 [B * nsym * nsc * bps] as qam_demodulate writes them for [B, nsym, nsc] grids."""
        g = geom if geom is not None else self.geom
        if bits_a.numel() != bits_b.numel() or bits_a.numel() != B * g.nsym * g.nsc * bps:
            raise ValueError("bit arrays do not match B x nsym x nsc x bps")
        pid = self._vec(pattern_id, B, torch.int32)
        counts = torch.empty((B,), dtype=torch.int32, device=self.device)
        if B:
            check(lib().b2c_bit_errors_per_slot(ref(g), ref(pool.struct), dptr(pid, "i32"), B, dptr(bits_a, "u8"), dptr(bits_b, "u8"), bps,
                                                C.c_void_p(counts.data_ptr()), stream_ptr()), "b2c_bit_errors_per_slot")
        return counts

    def ber_batch(self, B, model_id, doppler_hz, snr_db, pattern_id, pool, slot0=0, seed=42, method="zf",
                  estimates=("H_ls", "H_mmse", "H_true"), inject=None):
        """Link-level bit errors of B slots: a QPSK grid (2 payload bits per RE) through the slot pipeline, then for each
        channel estimate  equalize_channel (src/baseline_estimators.py:273-312) -> qam_demodulation (src/utils.py:111-152)
        -> bit errors against the transmitted bits on the data REs (calculate_ber, :155-157), all on the GPU.
        Every TX antenna sends the same grid (src/channel_simulator.py:402-404), so the payload is read off TX 0.
        `inject` (parity tests): recorded draws whose sym_turns are QPSK phases (2k+1)/8.
        Returns {"errors": {estimate: int32 [B]}, "bits": int64 [B] data bits per slot, "stats": per-slot error sums}."""
        out = self.run(B, model_id, doppler_hz, snr_db, pattern_id, pool, slot0=slot0, seed=seed, qpsk=inject is None, inject=inject)
        tx0 = out["tx"][:, :, 0, :].contiguous()
        bits_tx = self.qam_demodulate(tx0.reshape(-1), 4)
        res = {}
        for name in estimates:
            xh = self.equalize(out["rx"].contiguous(), out[name].contiguous(), method)[:, :, 0, :].contiguous()
            res[name] = self.bit_errors_per_slot(self.qam_demodulate(xh.reshape(-1), 4), bits_tx, pool, pattern_id, B)
        pid = np.broadcast_to(np.asarray(pattern_id.cpu() if isinstance(pattern_id, torch.Tensor) else pattern_id), (B,))
        nbits = 2 * (self.nsym * self.nsc - pool.npilots_host[pid].astype(np.int64))
        return {"errors": res, "bits": nbits, "stats": out["stats"], "_keepalive": out}

    def count_nonfinite(self, x, counts=None):
        """(#NaN, #Inf) elements of a float32 / complex64 CUDA tensor (numpy.isnan / isinf semantics), accumulated
        into the int64 pair `counts`."""
        if counts is None:
            counts = torch.zeros((2,), dtype=torch.int64, device=self.device)
        x = x.contiguous()
        if x.numel():
            check(lib().b2c_count_nonfinite(dptr(x, "c64" if x.is_complex() else "f32"), x.numel(), int(x.is_complex()),
                                            dptr(counts, "i64"), stream_ptr()), "b2c_count_nonfinite")
        return counts

    def abs_diff_sum(self, a, b):
        """sum |a - b| over two complex64 CUDA tensors of equal size (float64 device scalar)."""
        if a.numel() != b.numel():
            raise ValueError("arrays differ in size")
        out = torch.zeros((1,), dtype=torch.float64, device=self.device)
        if a.numel():
            check(lib().b2c_abs_diff_sum(dptr(a.contiguous(), "c64"), dptr(b.contiguous(), "c64"), a.numel(), dptr(out, "f64"),
                                         stream_ptr()), "b2c_abs_diff_sum")
        return out

    def _ls_sym_stride(self, H_ls, g):
        P = g.pitch if g.pitch else g.nsc
        if H_ls.dim() == 5:
            return g.nrx * g.ntx * P
        return g.nrx * P              # compact [B][nsym][nrx][nsc]

    def pair00_moments(self, rx, H_ls, H_true, moments=None, geom=None):
        """moments[3][4] += (sum re, sum im, sum re^2, sum im^2) of the pair-(0,0) rows of rx, H_ls, H_true."""
        P = row_pitch(rx)
        g = self._with_pitch(geom if geom is not None else self.geom, P)
        if moments is None:
            moments = torch.zeros((3, 4), dtype=torch.float64, device=self.device)
        check(lib().b2c_pair00_moments(ref(g), rx.shape[0], rows_ptr(rx, P), rows_ptr(H_ls, P), rows_ptr(H_true, P),
                                       self._ls_sym_stride(H_ls, g), dptr(moments, "f64"), stream_ptr()),
              "b2c_pair00_moments")
        return moments

    def pair00_errors(self, H_ls, H_true, alpha=None, geom=None):
        """[B][3] float64: sum |L-H|^2, sum |alpha L - H|^2, sum |H|^2 over the pair-(0,0) rows of each slot."""
        P = row_pitch(H_true)
        g = self._with_pitch(geom if geom is not None else self.geom, P)
        B = H_true.shape[0]
        out = torch.empty((B, 3), dtype=torch.float64, device=self.device)
        a = None if alpha is None else self._vec(alpha, B, torch.float32)
        check(lib().b2c_pair00_errors(ref(g), B, rows_ptr(H_ls, P), rows_ptr(H_true, P), self._ls_sym_stride(H_ls, g),
                                      dptr(a, "f32", True), dptr(out, "f64"), stream_ptr()), "b2c_pair00_errors")
        return out

    @staticmethod
    def normalization_from_moments(moments, count):
        """{rx_mean, rx_scale, ls_mean, ls_scale, true_mean, true_scale} as ChannelDataset does it
        (src/train.py:41-57, 80-83): mean of the re / im means, 1 / (mean of the re / im stds + 1e-8)."""
        m = np.asarray(moments.cpu() if isinstance(moments, torch.Tensor) else moments, dtype=np.float64)
        out = []
        for q in range(3):
            mu = m[q, :2] / count
            sd = np.sqrt(np.maximum(m[q, 2:] / count - mu * mu, 0.0))
            out += [mu.mean(), 1.0 / (sd.mean() + 1e-8)]
        return np.array(out, dtype=np.float32)

    def ml_features(self, rx, H_ls, H_true, pool, pattern_id=0, layout="last", normalize=True, norm=None, geom=None):
        """5-channel inputs / 2-channel targets for the ML side from GPU-resident slots.
        layout 'last' + normalize -> prepare_ml_inputs; layout 'first' + norm (6 floats) -> ChannelDataset items."""
        P = row_pitch(rx)
        g = self._with_pitch(geom if geom is not None else self.geom, P)
        B = rx.shape[0]
        lay = {"last": 0, "first": 1}[layout]
        pid = self._vec(pattern_id, B, torch.int32)
        mode, nt = 0, None
        if norm is not None:
            mode, nt = 2, torch.as_tensor(np.asarray(norm, dtype=np.float32)).to(self.device)
        elif normalize:
            mode = 1
        nre = g.nsym * g.nsc
        shape_in = (B, g.nsym, g.nsc, 5) if lay == 0 else (B, 5, g.nsym, g.nsc)
        shape_t = (B, g.nsym, g.nsc, 2) if lay == 0 else (B, 2, g.nsym, g.nsc)
        inputs = torch.empty(shape_in, dtype=torch.float32, device=self.device)
        targets = torch.empty(shape_t, dtype=torch.float32, device=self.device)
        check(lib().b2c_ml_features(ref(g), ref(pool.struct), dptr(pid, "i32"), B, rows_ptr(rx, P), rows_ptr(H_ls, P),
                                    rows_ptr(H_true, P), self._ls_sym_stride(H_ls, g), lay, mode, dptr(nt, "f32", True),
                                    dptr(inputs, "f32"), dptr(targets, "f32"), stream_ptr()), "b2c_ml_features")
        return inputs, targets
