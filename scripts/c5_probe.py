"""Probe: C5 mixed workload (3 profiles x 4 Dopplers x 8 SNRs x 2 densities, 4x4) through sharded_statistics."""
import sys, time, json, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'channel-estimation-in-5g-network_b200'))
import torch
import dataset_generator as dg
from utils import default_config
cfg = default_config(4, 4)
res = {}
for batch in (2048, 8192, 16384):
    for tag, want in (("stats_only", ()), ("with_arrays", ("H_true", "rx", "tx", "H_ls", "H_mmse"))):
        if want and batch > 8192:
            continue
        ds = dg.ChannelEstimationDataset(cfg, rng='philox', seed=42)
        dg.sharded_statistics(cfg, 2 * batch, batch=batch, dataset=ds, want_arrays=want)   # warm-up (plans, module load)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        N = 131072
        bins = dg.sharded_statistics(cfg, N, batch=batch, dataset=ds, want_arrays=want)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res[f"{tag}_b{batch}"] = {"slots_per_s": round(N / dt), "count": float(bins[:, 0].sum())}
print(json.dumps(res))
