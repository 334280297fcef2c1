#!/usr/bin/env python
"""torchrun --nproc-per-node=W tools/sharded_check.py — the W-rank engine against the single-GPU
engine on the same global path set (every rank recomputes the single-GPU answer locally)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    from loguru import logger

    logger.remove()
    import scenarios
    from monte_carlo_retirement_b200.config import Config
    from monte_carlo_retirement_b200.parallel import ShardedSimulator
    from monte_carlo_retirement_b200.simulation import RetirementMonteCarloSimulator

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = Config(**dict(scenarios.STRESSED, target_probability=70.0, num_simulations_search=20_000))
    n, wm = 50_001, 150
    sh = ShardedSimulator(cfg, device=local)
    one = RetirementMonteCarloSimulator(cfg, device=local)
    for s in (sh, one):
        s.use_final_seeds()
    a, b = sh.run_aggregates(wm, n), one.run_aggregates(wm, n)
    for k in b:
        va, vb = a[k], b[k]
        if hasattr(vb, "to_numpy"):
            assert np.array_equal(va.to_numpy(), vb.to_numpy(), equal_nan=True), k
        else:
            assert va == vb or (va != va and vb != vb), (k, va, vb)
    from monte_carlo_retirement_b200.parallel import shard_range

    ta, tb = sh.run_monte_carlo_simulations(wm, n), one.run_monte_carlo_simulations(wm, n)
    lo, cnt = shard_range(n, dist.get_rank(), dist.get_world_size())
    # summary_df: all rows on rank 0 (a view of the shared host block), the own shard elsewhere
    want0 = tb[0] if dist.get_rank() == 0 else tb[0].iloc[lo:lo + cnt].reset_index(drop=True)
    assert ta[0].equals(want0)
    # a second call while the first result is alive must not overwrite it (block generations)
    keep = ta[0].copy(deep=True)
    t2 = sh.run_monte_carlo_simulations(wm + 12, n)
    assert ta[0].equals(keep) and not t2[0].equals(keep)
    del t2
    t3 = sh.run_monte_carlo_simulations(wm, n)      # the freed generation is reused
    assert t3[0].equals(keep) and len(sh._shared_blocks) == 2
    for i in (1, 3, 4):
        assert np.array_equal(ta[i].to_numpy(), tb[i].to_numpy(), equal_nan=True), i
    assert ta[2] == tb[2] and ta[5] == tb[5] and ta[6] == tb[6]
    ra = sh.find_minimum_working_months(verbose=False)
    rb = one.find_minimum_working_months(verbose=False)
    assert ra == rb, (ra, rb)
    dist.barrier()
    if dist.get_rank() == 0:
        print(f"SHARDED CHECK OK world={dist.get_world_size()} months={ra[0]} p={ra[1]:.3f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
