#!/usr/bin/env python
"""BASELINE.json configs[3]: the full working-months search at 1e6 paths per candidate.

Times find_minimum_working_months() entry-to-return for the batched policies and checks that the
selected month equals the reference's decision procedure (the oracle's restatement of
simulation.py:1158-1342) applied to the same device success table.

    python tools/bench_search.py [--paths 1000000] [--scenario CONFIG_JSON] [--policies waves,grid,sequential]
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
    ap.add_argument("--paths", type=int, default=1_000_000)
    ap.add_argument("--scenario", default="CONFIG_JSON")
    ap.add_argument("--policies", default="auto,waves,grid,sequential")
    a = ap.parse_args()
    from loguru import logger

    logger.remove()
    import torch

    import scenarios
    from monte_carlo_retirement_b200.config import Config
    from monte_carlo_retirement_b200.simulation import RetirementMonteCarloSimulator
    from oracle import oracle as orc

    cfg_d = dict(getattr(scenarios, a.scenario), num_simulations_search=a.paths)
    cfg = Config(**cfg_d)
    R = cfg.retirement_years
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        from monte_carlo_retirement_b200.parallel import ShardedSimulator

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = {"scenario": a.scenario, "paths_per_candidate": a.paths, "n_gpus": world, "policies": {}}
    for pol in a.policies.split(","):
        if world > 1:
            if pol == "sequential":
                continue
            sim = ShardedSimulator(cfg, device=local, search_policy=pol)
        else:
            sim = RetirementMonteCarloSimulator(cfg, search_policy=pol)
        sim.find_minimum_working_months(verbose=False)  # warm-up (context, allocator)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        months, prob, curve = sim.find_minimum_working_months(verbose=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = dict(sim.last_search_stats)
        probes = [p["working_months"] for p in curve]
        # the reference's decisions on the same table
        table = dict(zip(sorted(set(probes)),
                         sim._reduce_counts(sim.batched_success_counts(sorted(set(probes)), a.paths)).cpu().tolist()))
        m2, p2, c2, _ = orc.search_decisions(lambda m: table[m] / a.paths * 100.0, cfg.starting_working_months_search,
                                             cfg.target_probability, a.paths)
        assert (m2, p2, c2) == (months, prob, curve), "device search disagrees with the reference decision procedure"
        res["policies"][pol] = {"wall_s": dt, "months": months, "probability": prob, "probes": len(curve),
                                "launches": st.get("launches"), "candidates_evaluated": st.get("candidates_evaluated"),
                                "nominal_path_months_probed": sum((m + 12 * R) * a.paths for m in probes)}
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
