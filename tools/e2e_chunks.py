#!/usr/bin/env python
"""Wall time of run_monte_carlo_simulations (host 7-tuple) for several splits of the timeline launch."""
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
for chunks, frac in ((1, None), (2, 0.5), (2, 0.6), (2, 0.7), (2, 0.8), (2, 0.9), (3, None)):
    sim = RetirementMonteCarloSimulator(Config(**scenarios.SYNTH_C3)); sim.use_final_seeds()
    sim.e2e_chunks = chunks
    sim.e2e_first_fraction = frac
    for _ in range(3):
        tup = sim.run_monte_carlo_simulations(wm, n)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); tup = sim.run_monte_carlo_simulations(wm, n); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"chunks={chunks} first={frac}: median {sorted(ts)[len(ts)//2]:.3f} ms  min {min(ts):.3f}")
