#!/usr/bin/env python
"""Where does a bench step spend its time: host wall per enqueue, device time per step, allocator activity."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from loguru import logger
logger.remove()
import torch
import scenarios
from monte_carlo_retirement_b200.config import Config
from monte_carlo_retirement_b200.simulation import RetirementMonteCarloSimulator

sim = RetirementMonteCarloSimulator(Config(**scenarios.SYNTH_C3))
sim.use_final_seeds()
n, wm = 1_000_000, 240
for _ in range(5):
    agg = sim.aggregates_device(wm, n, bands=True)
torch.cuda.synchronize()
for sync in (True, False):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    host = []
    st0 = torch.cuda.memory_stats()
    ev[0].record()
    for i in range(10):
        t0 = time.perf_counter()
        agg = sim.aggregates_device(wm, n, bands=True)
        host.append((time.perf_counter() - t0) * 1e3)
        ev[i + 1].record()
        if sync:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    st1 = torch.cuda.memory_stats()
    dev = [ev[i].elapsed_time(ev[i + 1]) for i in range(10)]
    print(f"sync={sync}: host enqueue ms {[round(h, 2) for h in host]}")
    print(f"           device ms       {[round(d, 2) for d in dev]}")
    print("           cudaMalloc calls", st1["num_device_alloc"] - st0["num_device_alloc"], "frees", st1["num_device_free"] - st0["num_device_free"],
          "retries", st1["num_alloc_retries"] - st0["num_alloc_retries"], "reserved GB", st1["reserved_bytes.all.current"] / 1e9)
print("plan", sim.last_series_plan)
