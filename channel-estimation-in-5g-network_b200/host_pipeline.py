"""Host-buffer front end of the slot pipeline: per-slot parameters come from pinned host memory,
results land in pinned host buffers (what the reference's NumPy API hands its caller).

Chunks are multi-buffered over two CUDA streams so the device->host copy of chunk i overlaps the
kernels of chunk i+1.  This is the end-to-end (`e2e`) path bench.py times; it is PCIe-bound:
a 4x4 slot is 3.76 MB of results (1.95 MB with the tx-replicated arrays sent once).
"""

from __future__ import annotations

import numpy as np
import torch

ARRAYS = ("H_true", "rx", "tx", "H_ls", "H_mmse")


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if part:
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index):
    """Restrict this process to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned buffers are
    allocated (first touch then places them on that node).  With one rank per GPU all writing results to host
    memory, buffers that all land on one socket cap the aggregate device->host rate well below the sum of the
    PCIe links.  Returns the node, or None when the topology is not visible (containers, single-node hosts)."""
    import os
    try:
        p = torch.cuda.get_device_properties(device_index)
        if all(hasattr(p, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        else:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(p.uuid)).encode() if not str(p.uuid).startswith("GPU-") else str(p.uuid).encode())
            bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
            bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()[-12:]
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = _cpulist(fh.read())
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def store_claimer(store, total, key="b2c_next_slot"):
    """claim(n) for HostPipeline.run_dynamic on top of a torch.distributed store (TCPStore / the default rendezvous
    store): an atomic add hands every caller, on any rank, a fresh range [first, first + n) of the job's `total` slots;
    None once the job is exhausted (the last range may be shorter: clip with `total`)."""
    def claim(n):
        first = store.add(key, int(n)) - int(n)
        return int(first) if first < total else None
    return claim


class HostPipeline:
    """Slots in, NumPy arrays out: per-slot parameters from pinned host memory, results in pinned host memory.

    Each in-flight chunk owns ONE device slab and ONE pinned host slab with identical layouts
    ([array][slot]..., arrays 256-byte aligned), so a full chunk leaves the device as a single cudaMemcpyAsync;
    `depth` chunks are in flight (kernels of chunk i+1 overlap the copy of chunk i).  On the default grid the slot
    kernel runs in its wide-store form (rows padded to pitch 600: one 16-byte store per lane); the host arrays are
    the [..., :599] views of those rows.  compact=True keeps one copy of the tx-replicated arrays (H_ls, H_mmse, tx)
    on the device AND on the link and hands the consumer full-shape broadcast views: 48 % fewer bytes for 4x4."""

    def __init__(self, engine, pool, chunk=512, want=ARRAYS + ("stats",), compact=False, pitch=None, depth=2):
        self.eng, self.pool, self.chunk, self.want, self.compact = engine, pool, int(chunk), tuple(want), bool(compact)
        e = engine
        est = any(k in self.want for k in ("H_ls", "H_mmse", "stats"))
        wide_ok = (e.nsc == 599 and e.nsym % 2 == 0 and e.ntx in (1, 2, 4, 8) and all(k in self.want for k in ("H_true", "rx", "tx"))
                   and (not est or "H_ls" in self.want))
        self.pitch = int(pitch) if pitch is not None else (600 if wide_ok else e.nsc)
        P, c = self.pitch, self.chunk
        full = (c, e.nsym, e.nrx, e.ntx, P)
        rep = (c, e.nsym, e.nrx, P) if compact else full
        shapes = {"H_true": full, "H_ls": rep, "H_mmse": rep, "rx": (c, e.nsym, e.nrx, P),
                  "tx": (c, e.nsym, P) if compact else (c, e.nsym, e.ntx, P)}
        self.layout, off = {}, 0                        # name -> (byte offset, shape, torch dtype)
        for k in ARRAYS:
            if k in self.want:
                self.layout[k] = (off, shapes[k], torch.complex64)
                off += -(-int(np.prod(shapes[k])) * 8 // 256) * 256
        if "stats" in self.want:
            self.layout["stats"] = (off, (c, e.nrx, 2, 3), torch.float64)
            off += -(-c * e.nrx * 48 // 256) * 256
        self.slab_bytes = off
        self.depth = max(2, int(depth))
        self.dev_slab = [torch.empty((off,), dtype=torch.uint8, device=e.device) for _ in range(self.depth)]
        self.host_slab = [torch.empty((off,), dtype=torch.uint8, pin_memory=True) for _ in range(self.depth)]
        self.dev = [self._carve(sl) for sl in self.dev_slab]           # padded tensors (rows of `pitch`)
        self.host = [self._carve(sl) for sl in self.host_slab]
        self.ws = [e.workspace(c) for _ in range(self.depth)]
        self.par_host = [torch.empty((4, c), dtype=torch.float32, pin_memory=True) for _ in range(self.depth)]
        self.par_dev = [torch.empty((4, c), dtype=torch.float32, device=e.device) for _ in range(self.depth)]
        self.compute = torch.cuda.Stream(device=e.device)
        self.copy = torch.cuda.Stream(device=e.device)
        self.ev_done = [torch.cuda.Event() for _ in range(self.depth)]
        self.ev_copied = [torch.cuda.Event() for _ in range(self.depth)]
        self.d2h_bytes_per_slot = self.slab_bytes / c                # what crosses the link (padding included)
        self.h2d_bytes_per_slot = 16

    def _carve(self, slab):
        out = {}
        for k, (off, shape, dt) in self.layout.items():
            n = int(np.prod(shape)) * (8 if dt == torch.complex64 else 8)
            out[k] = slab[off:off + n].view(dt).view(shape)
        return out

    def _payload(self, arrs, n):
        """[:n] slots, rows cut to the nsc payload elements."""
        return {k: (v[:n] if k == "stats" else v[:n][..., :self.eng.nsc]) for k, v in arrs.items()}

    def run(self, model_id, doppler_hz, snr_db, pattern_id, slot0=0, seed=42, consume=None):
        """Process len(model_id) slots; `consume(first_slot, n, host_buffers)` is called once each
        chunk's arrays are complete in pinned memory (buffers are reused `depth` chunks later)."""
        total = len(model_id)
        par = np.stack([np.asarray(model_id, np.float32), np.asarray(doppler_hz, np.float32),
                        np.asarray(snr_db, np.float32), np.asarray(pattern_id, np.float32)])
        chunks = ((slot0 + start, par[:, start:start + self.chunk]) for start in range(0, total, self.chunk))
        return self._run_chunks(chunks, seed, consume)

    def run_dynamic(self, claim, params_of, seed=42, consume=None):
        """Work-sharing front end for several ranks feeding host memory at different link rates (on an 8-GPU box the
        per-GPU device->host rate differs by 1.5x between PCIe domains): instead of a fixed shard, every rank CLAIMS
        the next chunk of global slot indices when one of its buffers frees up -- `claim(n)` returns the first global
        index of n fresh slots or None when the job is exhausted (e.g. an atomic add on the torch.distributed store),
        `params_of(first, n)` the [4, n] parameter block (model, doppler, snr, pattern) of those slots.  Philox draws are
        keyed by the global slot index, so the arrays do not depend on which rank produced them.  Returns the number
        of slots this rank processed."""
        def chunks():
            while True:
                first = claim(self.chunk)
                if first is None:
                    return
                yield first, np.asarray(params_of(first, self.chunk), np.float32)
        return self._run_chunks(chunks(), seed, consume)

    def _run_chunks(self, chunks, seed, consume):
        D = self.depth
        pending = [None] * D
        done = 0

        def retire(i):
            self.ev_copied[i].synchronize()
            if consume is not None:
                consume(*pending[i], self.views(i, pending[i][1]))
            pending[i] = None

        for c, (first, par) in enumerate(chunks):
            i = c % D
            n = par.shape[1]
            if pending[i] is not None:                 # buffer i still owned by an earlier chunk
                retire(i)
            self.par_host[i][:, :n] = torch.from_numpy(np.ascontiguousarray(par))
            with torch.cuda.stream(self.compute):
                self.compute.wait_event(self.ev_copied[i])
                self.par_dev[i].copy_(self.par_host[i], non_blocking=True)
                p = self.par_dev[i]
                ws = {k: v[:n] for k, v in self.ws[i].items()}
                self.eng.run(n, p[0, :n].to(torch.int32), p[1, :n], p[2, :n], p[3, :n].to(torch.int32), self.pool,
                             slot0=first, seed=seed, want=self.want, out=self._payload(self.dev[i], n), ws=ws,
                             compact=self.compact)
                self.ev_done[i].record(self.compute)
            with torch.cuda.stream(self.copy):
                self.copy.wait_event(self.ev_done[i])
                if n == self.chunk:
                    self.host_slab[i].copy_(self.dev_slab[i], non_blocking=True)       # one cudaMemcpyAsync per chunk
                else:                                                                  # ragged tail: one per array
                    for k, v in self.dev[i].items():
                        self.host[i][k][:n].copy_(v[:n], non_blocking=True)
                self.ev_copied[i].record(self.copy)
            pending[i] = (first, n)
            done += n
        for i in sorted((j for j in range(D) if pending[j] is not None), key=lambda j: pending[j][0]):
            retire(i)
        return done

    def views(self, i, n=None):
        """NumPy views of host buffer i (zero-copy; full reference shapes, stride 0 over tx if compact)."""
        arrs = {k: v.numpy() for k, v in self._payload(self.host[i], self.chunk if n is None else n).items()}
        return self.eng.expand_compact(arrs) if self.compact else arrs
