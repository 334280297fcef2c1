#!/usr/bin/env python
"""BASELINE.json configs[4]: the big final run (default 1e9 paths over the ranks of this torchrun
job; one series at a time when the three yearly series do not fit in HBM together) with
trajectory percentile bands + final-balance histograms, aggregate-only (nothing N-sized leaves
the GPUs).

    torchrun --nproc-per-node=8 tools/run_big.py --paths 1000000000 --wm 233
    python tools/run_big.py --paths 125000000            # one GPU's shard of the 8-GPU job
"""
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
    ap.add_argument("--paths", type=int, default=1_000_000_000)
    ap.add_argument("--wm", type=int, default=233)
    ap.add_argument("--scenario", default="CONFIG_JSON")
    ap.add_argument("--no-bands", action="store_true")
    ap.add_argument("--repeats", type=int, default=2, help="timed calls; wall_s is the last one (allocator warm)")
    a = ap.parse_args()
    from loguru import logger

    logger.remove()
    import torch
    import torch.distributed as dist

    import scenarios
    from monte_carlo_retirement_b200.config import Config
    from monte_carlo_retirement_b200.simulation import RetirementMonteCarloSimulator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cfg = Config(**getattr(scenarios, a.scenario))
    if world > 1:
        from monte_carlo_retirement_b200.parallel import ShardedSimulator

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        sim = ShardedSimulator(cfg, device=local)
    else:
        sim = RetirementMonteCarloSimulator(cfg, device=local)
    sim.use_final_seeds()
    sim.run_aggregates(a.wm, 100_000 * world, bands=not a.no_bands)  # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from bench import ClockSampler  # SM clocks / throttle reasons during the timed calls (NVML)

    sampler = ClockSampler(local)
    sampler.start()
    walls = []
    for _ in range(max(1, a.repeats)):  # the first call also pays the cudaMalloc of the 70-120 GB series blocks
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        out = sim.run_aggregates(a.wm, a.paths, bands=not a.no_bands)
        torch.cuda.synchronize()
        walls.append(time.perf_counter() - t0)
    dt = walls[-1]
    clocks = sampler.stop()
    if int(os.environ.get("RANK", "0")) == 0:
        months = a.wm + 12 * cfg.retirement_years
        summary = {k: v for k, v in out.items() if not hasattr(v, "to_numpy") and k not in ("ruin_month_hist",)}
        summary["final_balance_hist_musd_100"] = {"range": out["final_balance_hist_musd_100"]["range"]}
        summary["final_balance_hist_60"] = {"range": out["final_balance_hist_60"]["range"]}
        if "trajectory_bands" in out:
            summary["median_trajectory_last"] = float(out["trajectory_bands"][0.5].iloc[-1])
        print(json.dumps({"paths": a.paths, "n_gpus": world, "working_months": a.wm, "wall_s": dt,
                          "wall_s_each_call": walls, "clocks_rank0": clocks,
                          "nominal_path_months_per_s": a.paths * months / dt,
                          "executed_path_months": out["executed_path_months"],
                          "series_passes": [list(g) for g in sim.last_series_plan], "result": summary}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
