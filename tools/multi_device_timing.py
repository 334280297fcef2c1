#!/usr/bin/env python
"""e2e timing of the single-process multi-device engine (MultiDeviceSimulator): the reference's
7-tuple for 1e6 paths per device, all devices of the box, one process, no launcher.

    python tools/multi_device_timing.py [--per-device 1000000] [--reps 5]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-device", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    from loguru import logger

    logger.remove()
    import torch

    import scenarios
    from monte_carlo_retirement_b200.config import Config
    from monte_carlo_retirement_b200.multi_device import MultiDeviceSimulator

    g = torch.cuda.device_count()
    sim = MultiDeviceSimulator(Config(**scenarios.SYNTH_C3))
    n = a.per_device * g
    for _ in range(3):
        sim.run_monte_carlo_simulations(240, n)
    t0 = time.perf_counter()
    for _ in range(a.reps):
        out = sim.run_monte_carlo_simulations(240, n)
    dt = (time.perf_counter() - t0) / a.reps
    for _ in range(2):
        sim.run_aggregates(240, n)
    t0 = time.perf_counter()
    for _ in range(a.reps):
        agg = sim.run_aggregates(240, n)
    dta = (time.perf_counter() - t0) / a.reps
    print(json.dumps({"devices": g, "paths": n, "e2e_7tuple_ms": dt * 1e3, "path_months_per_s_7tuple": n * 720 / dt,
                      "aggregates_ms": dta * 1e3, "path_months_per_s_aggregates": n * 720 / dta,
                      "rows": len(out[0]), "success_probability": agg["success_probability"]}))
    sim.close()


if __name__ == "__main__":
    main()
