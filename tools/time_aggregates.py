#!/usr/bin/env python
"""CUDA-event time of the device aggregations alone (series already in HBM) for a scenario."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from loguru import logger
logger.remove()
import torch
import scenarios
from monte_carlo_retirement_b200.config import Config
from monte_carlo_retirement_b200.simulation import RetirementMonteCarloSimulator, TRAJECTORY_QUANTILES, WITHDRAWAL_RATE_QUANTILES

ap = argparse.ArgumentParser()
ap.add_argument("--scenario", default="SYNTH_C3"); ap.add_argument("--wm", type=int, default=240); ap.add_argument("--n", type=int, default=1_000_000)
a = ap.parse_args()
sim = RetirementMonteCarloSimulator(Config(**getattr(scenarios, a.scenario)))
sim.use_final_seeds()
b = sim.run_batch_device(a.wm, a.n)
ctx = sim.native_context
T, R, n = b.T, b.R, a.n
out = torch.empty((T, 7), dtype=torch.float64, device="cuda"); outw = torch.empty((R, 5), dtype=torch.float64, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print(f"{a.scenario} wm={a.wm} n={n}: success {int(b.counters[0])/n:.4f}")
print(f"  traj bands  {t(lambda: ctx.quantiles(b.traj, n, n, T, TRAJECTORY_QUANTILES, out)):.3f} ms  ({T} rows, {T*n*8/1e6:.0f} MB)")
print(f"  real bands  {t(lambda: ctx.quantiles(b.real, n, n, T, TRAJECTORY_QUANTILES, out)):.3f} ms")
print(f"  wr bands    {t(lambda: ctx.quantiles(b.wr, n, n, R, WITHDRAWAL_RATE_QUANTILES, outw)):.3f} ms  ({R} rows)")
m = torch.empty(1, dtype=torch.float64, device="cuda")
print(f"  1-row median {t(lambda: ctx.quantiles(b.cols[1], n, n, 1, [0.5], m, median=True)):.3f} ms")
agg = sim.aggregates_device(a.wm, n, bands=True)
plan = sim.last_series_plan
kw = dict(n_global=n, offset=0, working_months=a.wm, bands=True, plan=plan, samples=False, part_first=0)
l0 = ctx.launch_count
ms = t(lambda: sim._aggregate_batch(agg.batch, **kw), reps=10)
print(f"  ALL aggregations of a step on the resident batch: {ms:.3f} ms, {(ctx.launch_count - l0) // 11} library launches")
