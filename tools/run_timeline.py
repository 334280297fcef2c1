#!/usr/bin/env python
"""Runs only the timeline kernel (C3 shape by default) a few times and prints its CUDA-event
time — the short command the ncu captures in profiles/ are taken on.

    python tools/run_timeline.py [--n 1000000] [--wm 240] [--reps 5] [--strict] [--no-series] [--scenario SYNTH_C3]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--wm", type=int, default=240)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--strict", action="store_true")
    ap.add_argument("--no-series", action="store_true")
    ap.add_argument("--scenario", default="SYNTH_C3")
    ap.add_argument("--search", type=int, default=0, help="also run a batched search over this many candidates")
    a = ap.parse_args()
    try:
        from loguru import logger

        logger.remove()
    except Exception:
        pass
    import torch

    import scenarios
    from monte_carlo_retirement_b200.config import Config
    from monte_carlo_retirement_b200.simulation import RetirementMonteCarloSimulator

    cfg = getattr(scenarios, a.scenario)
    sim = RetirementMonteCarloSimulator(Config(**cfg), strict=a.strict)
    sim.use_final_seeds()
    months = a.wm + 12 * cfg["retirement_years"]
    times = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b = sim.run_batch_device(a.wm, a.n, series=not a.no_series)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
        ex = int(b.counters[1])
        del b
    best = min(times)
    print(f"timeline {a.scenario} n={a.n} wm={a.wm} strict={a.strict} series={not a.no_series}: "
          f"best {best:.3f} ms median {sorted(times)[len(times)//2]:.3f} ms -> {a.n * months / best / 1e-3:.4e} nominal pm/s, "
          f"executed {ex} ({ex * 220 / best / 1e-3 / 1e12:.3f} Tslot/s @W=220)")
    if a.search:
        sim.use_search_seeds()
        cands = list(range(0, a.search))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sim.batched_success_counts(cands[:2], 1024)
        e0.record()
        counts, executed = sim.batched_success_counts(cands, a.n, with_executed=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        ex = int(executed.sum())
        nominal = sum(c + 12 * cfg["retirement_years"] for c in cands) * a.n
        print(f"search {len(cands)} candidates x {a.n} paths: {ms:.2f} ms, nominal {nominal:.3e} pm ({nominal / ms / 1e-3:.3e}/s), "
              f"executed {ex:.3e} pm ({ex / ms / 1e-3:.3e}/s)")


if __name__ == "__main__":
    main()
