"""Randomised GPU parity over the configuration space.

For the random scenarios of tests/test_oracle_fuzz_vs_reference.py (where the CPU oracle is
shown to equal the unmodified reference bit for bit) the strict CUDA build, replaying the
reference's own numpy draws, must give the oracle's success flags and ruin months exactly and
its balances / series within 1e-9 relative: (1) one strict thread per path through
`_run_single_simulation_path`, (2) the batch kernel through `run_batch_device(shocks=...)`,
(3) the fast build on the same draws — with MCR_FLAG_SMALL_RETURNS (the lean-capable variant
bench.py times) whenever the draws satisfy the scenario's proven return bound.
"""
from __future__ import annotations

import numpy as np
import pytest

from gpu_util import assert_close, device_batch_to_host, make_sim, small_returns_hold
from oracle import oracle as orc
from test_oracle_fuzz_vs_reference import _random_config

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("block", range(4))
def test_cuda_equals_oracle_on_random_scenarios(block):
    import torch

    rng = np.random.default_rng(20261018 + block)
    for k in range(12):
        cfg = _random_config(rng, 100 * block + k)
        wm = int(rng.choice([0, 1, 11, 12, 13, 37, int(rng.integers(0, 201))]))
        rng.uniform(0, 1e6, 4)  # keep the stream aligned with the CPU fuzz test
        mine = orc.OracleSimulator(cfg)
        mine.use_final_seeds()
        seeds = mine.seeds.path_seeds(16)
        shocks = orc.shocks_for_seeds(mine.p, wm, seeds)                 # (n, rows, 3), the reference's draws
        recs, traj, real, wr = orc.run_batch(mine.p, wm, shocks, n_threads=2)
        sim = make_sim(cfg, rng="numpy")
        sim.use_final_seeds()
        # (1) single strict thread, first 3 paths
        for i in range(3):
            got = sim._run_path_on_shocks(wm, shocks[i])
            assert got["Success"] is bool(recs["success"][i]), (cfg, wm, i)
            assert_close(got["Final Balance"], recs["final_balance"][i])
            assert_close(np.array(got["Trajectory"]), traj[i])
        # (2) batch kernel, strict, and (3) fast build on the same draws
        dev = torch.from_numpy(np.ascontiguousarray(shocks.transpose(1, 2, 0))).to("cuda")   # [rows, 3, n]
        small = small_returns_hold(sim, shocks)
        for fast in (False, True, "small") if small else (False, True):
            h = device_batch_to_host(sim.run_batch_device(wm, len(seeds), shocks=dev, _fast_replay=bool(fast),
                                                          _small_returns=fast == "small"))
            assert np.array_equal(h["success"], recs["success"].astype(bool)), (cfg, wm, fast)
            want_ruin = np.where(np.isnan(recs["years_to_ruin"]), -1, np.rint(recs["years_to_ruin"] * 12)).astype(int)
            assert np.array_equal(h["ruin_month"], want_ruin), (cfg, wm, fast)
            for key, ref in (("start", recs["start_balance"]), ("final", recs["final_balance"]),
                             ("fy_gross", recs["first_year_gross"]), ("fy_real", recs["first_year_real"]),
                             ("infl", recs["inflation_at_ret"])):
                assert_close(h[key], ref)
            assert_close(h["traj"], traj)
            assert_close(h["real"], real)
            assert np.array_equal(np.isnan(h["wr"]), np.isnan(wr))
            assert_close(np.nan_to_num(h["wr"]), np.nan_to_num(wr))
