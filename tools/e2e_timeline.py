#!/usr/bin/env python
"""Host wall-clock marks inside run_monte_carlo_simulations(240, 1e6) against the device's own timeline."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from loguru import logger
logger.remove()
import torch
import scenarios
from monte_carlo_retirement_b200.config import Config
from monte_carlo_retirement_b200 import simulation as S

sim = S.RetirementMonteCarloSimulator(Config(**scenarios.SYNTH_C3)); sim.use_final_seeds()
n = 1_000_000
marks = []
def mark(name): marks.append((name, time.perf_counter()))
# wrap the pieces
for nm in ("_staging", "run_batch_device", "_band_quantiles"):
    f = getattr(sim, nm)
    def g(*a, _f=f, _n=nm, **k):
        mark(_n + " >"); r = _f(*a, **k); mark(_n + " <"); return r
    setattr(sim, nm, g)
ctx = sim.native_context
f2 = ctx.gather_columns
def g2(*a, **k):
    mark("gather >"); r = f2(*a, **k); mark("gather <"); return r
ctx.gather_columns = g2
for _ in range(3):
    sim.run_monte_carlo_simulations(240, n)
torch.cuda.synchronize()
for rep in range(3):
    marks.clear()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record()
    tup = sim.run_monte_carlo_simulations(240, n)
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
    print(f"rep {rep}: wall {1e3*(t1-t0):.3f} ms; device first-event to last {e0.elapsed_time(e1):.3f} ms")
    print("   " + "  ".join(f"{nm} {1e3*(t-t0):.3f}" for nm, t in marks))
