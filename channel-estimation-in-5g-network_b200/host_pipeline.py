"""Host-buffer front end of the slot pipeline: per-slot parameters come from pinned host memory,
results land in pinned host buffers (what the reference's NumPy API hands its caller).

Chunks are double-buffered over two CUDA streams so the device->host copy of chunk i overlaps the
kernels of chunk i+1.  This is the end-to-end (`e2e`) path bench.py times; it is PCIe-bound:
a 4x4 slot is 3.76 MB of results.
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


class HostPipeline:
    def __init__(self, engine, pool, chunk=512, want=ARRAYS + ("stats",), compact=False):
        """compact=True moves the tx-replicated arrays (H_ls, H_mmse, tx) over PCIe once and hands the
        consumer full-shape NumPy broadcast views (SlotEngine.expand_compact): same values, 48 % fewer
        bytes for a 4x4 slot."""
        self.eng, self.pool, self.chunk, self.want, self.compact = engine, pool, chunk, tuple(want), compact
        self.dev = [engine.alloc_outputs(chunk, self.want, compact) for _ in range(2)]
        self.ws = [engine.workspace(chunk) for _ in range(2)]
        self.host = [{k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in d.items()} for d in self.dev]
        self.par_host = [torch.empty((4, chunk), dtype=torch.float32, pin_memory=True) for _ in range(2)]
        self.par_dev = [torch.empty((4, chunk), dtype=torch.float32, device=engine.device) for _ in range(2)]
        self.compute = torch.cuda.Stream(device=engine.device)
        self.copy = torch.cuda.Stream(device=engine.device)
        self.ev_done = [torch.cuda.Event() for _ in range(2)]
        self.ev_copied = [torch.cuda.Event() for _ in range(2)]
        self.d2h_bytes_per_slot = sum(v[0].numel() * v.element_size() for v in self.dev[0].values())
        self.h2d_bytes_per_slot = 16

    def run(self, model_id, doppler_hz, snr_db, pattern_id, slot0=0, seed=42, consume=None):
        """Process len(model_id) slots; `consume(first_slot, n, host_buffers)` is called once each
        chunk's arrays are complete in pinned memory (buffers are reused two chunks later)."""
        total = len(model_id)
        pending = [None, None]
        par = np.stack([np.asarray(model_id, np.float32), np.asarray(doppler_hz, np.float32),
                        np.asarray(snr_db, np.float32), np.asarray(pattern_id, np.float32)])
        for c, start in enumerate(range(0, total, self.chunk)):
            i = c & 1
            n = min(self.chunk, total - start)
            if pending[i] is not None:                 # buffer i still owned by an earlier chunk
                self.ev_copied[i].synchronize()
                if consume is not None:
                    consume(*pending[i], self.views(i))
            self.par_host[i][:, :n] = torch.from_numpy(par[:, start:start + n])
            with torch.cuda.stream(self.compute):
                self.compute.wait_event(self.ev_copied[i])
                self.par_dev[i].copy_(self.par_host[i], non_blocking=True)
                p = self.par_dev[i]
                out = {k: v[:n] for k, v in self.dev[i].items()}
                ws = {k: v[:n] for k, v in self.ws[i].items()}
                self.eng.run(n, p[0, :n].to(torch.int32), p[1, :n], p[2, :n], p[3, :n].to(torch.int32), self.pool,
                             slot0=slot0 + start, seed=seed, want=self.want, out=out, ws=ws, compact=self.compact)
                self.ev_done[i].record(self.compute)
            with torch.cuda.stream(self.copy):
                self.copy.wait_event(self.ev_done[i])
                for k, v in self.dev[i].items():
                    self.host[i][k][:n].copy_(v[:n], non_blocking=True)
                self.ev_copied[i].record(self.copy)
            pending[i] = (slot0 + start, n)
        for i in sorted(range(2), key=lambda j: pending[j][0] if pending[j] is not None else -1):
            if pending[i] is not None:
                self.ev_copied[i].synchronize()
                if consume is not None:
                    consume(*pending[i], self.views(i))
        return total

    def views(self, i):
        """NumPy views of host buffer i (zero-copy; full reference shapes, stride 0 over tx if compact)."""
        arrs = {k: v.numpy() for k, v in self.host[i].items()}
        return self.eng.expand_compact(arrs) if self.compact else arrs
