#!/usr/bin/env python
"""Which host call of a step is slow when a step's enqueue takes > 8 ms?"""
import os, sys, time, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from loguru import logger
logger.remove()
import torch
import scenarios
from monte_carlo_retirement_b200.config import Config
from monte_carlo_retirement_b200 import native
from monte_carlo_retirement_b200.simulation import RetirementMonteCarloSimulator

sim = RetirementMonteCarloSimulator(Config(**scenarios.SYNTH_C3)); sim.use_final_seeds()
n, wm = 1_000_000, 240
ctx = sim.native_context
log = []
def wrap(obj, name):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        log.append((name, (time.perf_counter() - t0) * 1e3))
        return r
    setattr(obj, name, g)
for nm in ("run_batch_device", "_select", "_final_balance_histograms", "_series_plan"):
    wrap(sim, nm)
for nm in ("first_year_rates", "select_rows", "quantiles_rows", "histogram", "run_batch"):
    if hasattr(ctx, nm): wrap(ctx, nm)
orig_empty = torch.empty
for _ in range(5):
    agg = sim.aggregates_device(wm, n, bands=True)
torch.cuda.synchronize()
slow = 0
for rep in range(30):
    torch.cuda.synchronize()
    for i in range(20):
        log.clear()
        a0 = torch.cuda.memory_stats()["num_device_alloc"]
        t0 = time.perf_counter()
        agg = sim.aggregates_device(wm, n, bands=True)
        dt = (time.perf_counter() - t0) * 1e3
        if dt > 8 and slow < 12:
            slow += 1
            print(f"rep {rep} step {i}: {dt:.1f} ms; cudaMallocs {torch.cuda.memory_stats()['num_device_alloc'] - a0};", " ".join(f"{k}={v:.1f}" for k, v in log if v > 0.5))
print("done; slow calls shown:", slow)
