"""Bare device->host copy rate from 1/2/4/8 processes at once: is the end-to-end ceiling the link or the pipeline?

    python scripts/pcie_d2h_probe.py                       # one process, GPU 0
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/pcie_d2h_probe.py

Every rank copies slabs of the sizes HostPipeline moves (one cudaMemcpyAsync per slab into pinned memory, back to back
on one stream), all ranks between the same two barriers; rank 0 prints one JSON object with the per-rank and aggregate
GB/s per slab size.  bench.py's e2e.roofline.peak is the same measurement taken inside the bench run."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    out = {"world": world, "sizes_MB": [], "per_rank_gbs": [], "aggregate_gbs": [], "h2d_aggregate_gbs": []}
    for mb in (16, 64, 250, 1000):
        n = mb * 1000 * 1000
        src = torch.empty((n,), dtype=torch.uint8, device=dev)
        dst = [torch.empty((n,), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        st = torch.cuda.Stream(device=dev)
        res = []
        for direction in ("d2h", "h2d"):
            with torch.cuda.stream(st):
                for i in range(2):
                    (dst[i].copy_(src, non_blocking=True) if direction == "d2h" else src.copy_(dst[i], non_blocking=True))
                st.synchronize()
                if world > 1:
                    dist.barrier()
                reps = max(4, int(0.5 * 50e9 / n))
                t0 = time.perf_counter()
                for i in range(reps):
                    (dst[i & 1].copy_(src, non_blocking=True) if direction == "d2h" else src.copy_(dst[i & 1], non_blocking=True))
                st.synchronize()
                dt = time.perf_counter() - t0
            res.append(reps * n / dt / 1e9)
        t = torch.tensor(res, dtype=torch.float64, device=dev)
        if world > 1:
            allr = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allr, t)
        else:
            allr = [t]
        if rank == 0:
            out["sizes_MB"].append(mb)
            out["per_rank_gbs"].append([round(float(a[0]), 2) for a in allr])
            out["aggregate_gbs"].append(round(sum(float(a[0]) for a in allr), 2))
            out["h2d_aggregate_gbs"].append(round(sum(float(a[1]) for a in allr), 2))
        del src, dst
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
