"""Probe: C4 pilot-density x SNR sweep (4x4 EVA 50 Hz, densities 1..10 %, 8 SNRs) through PilotOptimizer (Philox mode)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'channel-estimation-in-5g-network_b200'))
import torch
import run_phase8_pilot_optimization as p8

opt = p8.PilotOptimizer(rng='philox', seed=42)
opt.config["mimo"] = {"num_tx_antennas": 4, "num_rx_antennas": 4}
dens = [round(0.01 * i, 2) for i in range(1, 11)]
snrs = [-5, 0, 5, 10, 15, 20, 25, 30]
opt.analyze_pilot_density(dens, snrs, num_samples=20)            # warm-up: plans (Qhull) for the ten patterns, module load
torch.cuda.synchronize()
n = 2000
t0 = time.perf_counter()
res = opt.analyze_pilot_density(dens, snrs, num_samples=n)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
slots = len(dens) * len(snrs) * n
print(json.dumps({"slots": slots, "seconds": dt, "slots_per_s": slots / dt,
                  "ls_nmse_db_at_10dB": {str(d): round(res['methods']['LS'][10][d]['nmse_db'], 2) for d in dens}}))
