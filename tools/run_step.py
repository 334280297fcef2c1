#!/usr/bin/env python
"""A few whole C3 steps (timeline + every aggregation) and nothing else — the target of ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from loguru import logger
logger.remove()
import torch
import scenarios
from monte_carlo_retirement_b200.config import Config
from monte_carlo_retirement_b200.simulation import RetirementMonteCarloSimulator

sim = RetirementMonteCarloSimulator(Config(**scenarios.SYNTH_C3)); sim.use_final_seeds()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    agg = sim.aggregates_device(240, 1_000_000, bands=True)
torch.cuda.synchronize()
print(agg.to_host()["success_probability"])
