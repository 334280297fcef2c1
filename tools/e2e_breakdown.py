#!/usr/bin/env python
"""Where the end-to-end time of run_monte_carlo_simulations(240, 1e6) goes (host side)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from loguru import logger
logger.remove()
import cProfile, pstats
import torch
import scenarios
from monte_carlo_retirement_b200.config import Config
from monte_carlo_retirement_b200.simulation import RetirementMonteCarloSimulator

sim = RetirementMonteCarloSimulator(Config(**scenarios.SYNTH_C3))
sim.use_final_seeds()
n = 1_000_000
for _ in range(2):
    sim.run_monte_carlo_simulations(240, n)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    sim.run_monte_carlo_simulations(240, n)
torch.cuda.synchronize()
print(f"e2e {1e3 * (time.perf_counter() - t0) / 5:.2f} ms/call")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    sim.run_monte_carlo_simulations(240, n)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
