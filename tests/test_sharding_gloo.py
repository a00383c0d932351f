"""N>1 host logic on CPU: two gloo ranks shard the global sample range exactly like the NCCL path
(dataset_generator.shard_range + reduce_bins) and reduce their per-bin accumulators.  The GPU work
itself is replaced by a deterministic stand-in keyed by the global sample index -- what is under
test is the partition, the parameter draw per global index and the collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_bins(lo, hi, seed, sizes, nbins):
    """Stand-in for sharded_statistics: per-slot 'nmse' is a pure function of the global index."""
    import dataset_generator as dg
    mi, di, si, pi = dg.philox_param_choice(seed, lo, hi - lo, sizes)
    g = np.arange(lo, hi, dtype=np.float64)
    nmse = 0.1 + (g % 97) / 97.0
    bins = np.zeros((nbins, 14))
    np.add.at(bins[:, 0], si, 1.0)
    np.add.at(bins[:, 3], si, nmse)
    np.add.at(bins[:, 5], si, nmse ** 2)
    return torch.from_numpy(bins)


def _worker(rank, world, port, total, out_dir):
    import sys
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import dataset_generator as dg
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = dg.shard_range(total, rank, world)
    bins = _fake_bins(lo, hi, 42, (3, 4, 8, 2), 8)
    bins = dg.reduce_bins(bins)
    if rank == 0:
        np.save(os.path.join(out_dir, "bins.npy"), bins.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_rank_sharding_and_reduction(world, tmp_path):
    total = 1001
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "bins.npy")
    want = _fake_bins(0, total, 42, (3, 4, 8, 2), 8).numpy()
    assert np.array_equal(got[:, 0], want[:, 0]) and got[:, 0].sum() == total
    assert np.allclose(got, want, rtol=1e-13, atol=0)


def test_reduce_bins_is_identity_without_a_process_group():
    import dataset_generator as dg
    b = torch.arange(28, dtype=torch.float64).reshape(2, 14)
    assert torch.equal(dg.reduce_bins(b.clone()), b)


def _claim_worker(rank, world, port, total, chunk, out_dir):
    import sys
    import time
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    from host_pipeline import store_claimer
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    claim = store_claimer(dist.distributed_c10d._get_default_store(), total, "job")
    mine = []
    while True:
        first = claim(chunk)
        if first is None:
            break
        mine.append((first, min(chunk, total - first)))
        time.sleep(0.001 * (1 + 3 * rank))                # ranks drain at different rates, like links of different speed
    np.save(os.path.join(out_dir, f"claims{rank}.npy"), np.array(mine, dtype=np.int64).reshape(-1, 2))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_dynamic_claims_partition_the_job(world, tmp_path):
    """HostPipeline.run_dynamic's work sharing: ranks claim chunks of the global slot range from the rendezvous store
    (atomic add).  Whatever their relative speeds, the claimed ranges are disjoint and cover [0, total) exactly, the
    last one clipped -- so the union of the ranks' outputs is the job, and (Philox keyed by the global index) the same
    arrays as any other split."""
    total, chunk = 1000, 48
    port = _free_port()
    mp.spawn(_claim_worker, args=(world, port, total, chunk, str(tmp_path)), nprocs=world, join=True)
    claims = [np.load(tmp_path / f"claims{r}.npy") for r in range(world)]
    allc = np.concatenate(claims)
    order = np.argsort(allc[:, 0])
    firsts, counts = allc[order, 0], allc[order, 1]
    assert firsts[0] == 0 and np.array_equal(firsts[1:], (firsts + counts)[:-1]) and firsts[-1] + counts[-1] == total
    assert np.all(counts[:-1] == chunk) and 0 < counts[-1] <= chunk
    assert all(len(c) > 0 for c in claims)                                     # every rank took part
    assert len(claims[0]) > len(claims[-1])                                    # the faster rank took more
