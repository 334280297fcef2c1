#!/usr/bin/env python
"""Wall time of run_monte_carlo_simulations (host 7-tuple) for several chunk counts of the timeline launch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from loguru import logger
logger.remove()
import torch
import scenarios
from monte_carlo_retirement_b200.config import Config
from monte_carlo_retirement_b200.simulation import RetirementMonteCarloSimulator

n, wm = 1_000_000, 240
for chunks in (1, 2, 3, 4, 6, 8):
    sim = RetirementMonteCarloSimulator(Config(**scenarios.SYNTH_C3)); sim.use_final_seeds()
    sim.e2e_chunks = chunks
    for _ in range(3):
        tup = sim.run_monte_carlo_simulations(wm, n)
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); tup = sim.run_monte_carlo_simulations(wm, n); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"chunks={chunks}: median {sorted(ts)[len(ts)//2]:.3f} ms  min {min(ts):.3f}  all {[round(t,2) for t in ts]}")
