#!/usr/bin/env python
"""Are the rare slow bench runs Python garbage collections? 40 x (20 un-synchronised steps) per mode;
device time per step from CUDA events, the longest host gap inside each loop, and every gc pass."""
import gc, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from loguru import logger
logger.remove()
import torch
import scenarios
from monte_carlo_retirement_b200.config import Config
from monte_carlo_retirement_b200.simulation import RetirementMonteCarloSimulator

sim = RetirementMonteCarloSimulator(Config(**scenarios.SYNTH_C3)); sim.use_final_seeds()
n, wm = 1_000_000, 240
for _ in range(5):
    agg = sim.aggregates_device(wm, n, bands=True)
torch.cuda.synchronize()
gc_log = []
t_gc = [0.0]
def cb(phase, info):
    if phase == "start": t_gc[0] = time.perf_counter()
    else: gc_log.append((info["generation"], (time.perf_counter() - t_gc[0]) * 1e3))
gc.callbacks.append(cb)
for mode in ("gc on", "gc off", "gc frozen"):
    gc.enable()
    if mode == "gc off": gc.disable()
    if mode == "gc frozen": gc.collect(); gc.freeze()
    res, gaps = [], []
    gc_log.clear()
    for rep in range(40):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); worst = 0.0
        for i in range(20):
            t0 = time.perf_counter()
            agg = sim.aggregates_device(wm, n, bands=True)
            worst = max(worst, (time.perf_counter() - t0) * 1e3)
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 20); gaps.append(worst)
    s = sorted(res)
    print(f"{mode}: ms/step median {s[20]:.3f} p90 {s[36]:.3f} max {s[-1]:.3f}; worst host enqueue per loop: median {sorted(gaps)[20]:.2f} max {max(gaps):.2f} ms")
    print("   gc passes:", len(gc_log), "gen2:", [round(d, 1) for g, d in gc_log if g == 2], "longest gen0/1:", round(max([d for g, d in gc_log if g < 2] or [0]), 2))
