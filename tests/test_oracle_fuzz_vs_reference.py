"""Randomised pin of the CPU oracle against the UNMODIFIED Python reference, run live.

Only where the reference tree is mounted (the build container; /root/reference does not exist
on the GPU box, where this module skips). For random scenarios — tax systems on/off with random
rates (incl. 0 and high), allocations incl. the corners, volatilities incl. 0, correlation in
[-1, 1], 0-4 income streams (indexed or not, fractional start ages, finite durations, zero
amounts), 1-40 retirement years, 0-200 working months — every key of the per-path dict of
`_run_single_simulation_path` (simulation.py:476-950) must be reproduced bit for bit by
oracle/path_oracle.c on the same numpy draws. The committed golden vectors pin 12 hand-picked
scenarios; this widens the pin to the whole configuration space.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import pytest

from oracle import oracle as orc

REF = "/root/reference/backend"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is not mounted here")


def _random_config(rng: np.random.Generator, k: int) -> dict:
    use1, use2 = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    pick = lambda *v: v[int(rng.integers(0, len(v)))]  # noqa: E731
    streams = []
    for s in range(int(rng.integers(0, 5))):
        streams.append({
            "name": f"s{s}", "monthly_amount_today": float(pick(0.0, rng.uniform(100, 6000))),
            "start_at_age": float(pick(40.0, 65.0, rng.uniform(35, 80), 60.5)),
            "duration_years": pick(None, 0, int(rng.integers(1, 30))),
            "inflation_indexed": bool(rng.integers(0, 2)), "tax_rate": float(pick(0.0, rng.uniform(0, 0.5))),
        })
    return {
        "scenario": f"fuzz{k}", "initial_balance": float(pick(0.0, rng.uniform(1e3, 2e6))),
        "monthly_contribution": float(pick(0.0, rng.uniform(0, 1e4))),
        "contribution_growth_rate_annual": float(pick(0.0, rng.uniform(0, 0.08))),
        "monthly_expenses": float(pick(0.0, rng.uniform(500, 2e4))), "current_age": float(pick(40.0, rng.uniform(25, 60))),
        "retirement_years": int(rng.integers(1, 41)), "allocation_inv1_pct": float(pick(0.0, 1.0, rng.uniform(0, 1))),
        "inv1_returns_mean": float(rng.uniform(-0.02, 0.15)), "inv1_returns_volatility": float(pick(0.0, rng.uniform(0, 0.3))),
        "inv1_annual_tax_on_gains_rate": float(pick(0.0, rng.uniform(0, 0.4))),
        "inv1_realized_gains_tax_rate": float(pick(0.0, rng.uniform(0, 0.6))), "inv1_use_realized_gains_tax_system": use1,
        "inv2_premium_over_inflation_mean": float(rng.uniform(-0.01, 0.06)),
        "inv2_premium_over_inflation_volatility": float(pick(0.0, rng.uniform(0, 0.08))),
        "inv2_annual_tax_on_gains_rate": float(pick(0.0, rng.uniform(0, 0.4))),
        "inv2_realized_gains_tax_rate": float(pick(0.0, rng.uniform(0, 0.6))), "inv2_use_realized_gains_tax_system": use2,
        "inflation_rate_mean": float(rng.uniform(0.0, 0.09)), "inflation_rate_volatility": float(pick(0.0, rng.uniform(0, 0.04))),
        "equity_inflation_correlation": float(pick(0.0, -1.0, 1.0, rng.uniform(-1, 1))),
        "num_simulations_main": 10, "num_simulations_search": 10, "target_probability": 90.0,
        "starting_working_months_search": 0, "seed": int(rng.integers(0, 2 ** 31)), "num_processes": 1,
        "other_income_streams": streams,
    }


def _reference_modules():
    saved_path = list(sys.path)
    saved = {k: sys.modules.pop(k, None) for k in ("config", "constants", "simulation", "utils")}
    sys.path.insert(0, REF)
    try:
        from loguru import logger

        logger.remove()
        import config as ref_config
        import simulation as ref_simulation
        return ref_config, ref_simulation
    finally:
        sys.path[:] = saved_path
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


def _same(a, b) -> bool:
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


@pytest.mark.parametrize("block", range(4))
def test_oracle_equals_reference_on_random_scenarios(block):
    ref_config, ref_simulation = _reference_modules()
    rng = np.random.default_rng(20261018 + block)
    failures = 0
    for k in range(12):
        cfg = _random_config(rng, 100 * block + k)
        wm = int(rng.choice([0, 1, 11, 12, 13, 37, int(rng.integers(0, 201))]))
        ref = ref_simulation.RetirementMonteCarloSimulator(ref_config.Config(**cfg))
        mine = orc.OracleSimulator(cfg)
        assert mine.main_seed == ref.main_seed
        ref.use_final_seeds()
        mine.use_final_seeds()
        seeds = list(ref._path_seeds(4))
        assert mine.seeds.path_seeds(4) == [int(s) for s in seeds]
        for seed in seeds:
            want = ref._run_single_simulation_path(wm, int(seed))
            got = mine.run_single(wm, int(seed))
            assert got.keys() == want.keys()
            assert got["Success"] is bool(want["Success"]), (cfg, wm, seed)
            failures += not want["Success"]
            for key in want:
                if key == "Success":
                    continue
                assert _same(got[key], want[key]), (key, cfg, wm, seed, got[key], want[key])
        # the aggregated 7-tuple (seed spawning for a NEW n, pandas quantiles, sampled columns, WR counts)
        if k % 3 == 0:
            w = ref.run_monte_carlo_simulations(wm, 7)
            g = mine.run(wm, 7)
            assert list(g[0].columns) == list(w[0].columns)
            for col in w[0].columns:
                assert _same(g[0][col].to_numpy(dtype=float), w[0][col].to_numpy(dtype=float)), (col, cfg, wm)
            for i in (1, 3, 4):
                assert list(g[i].columns) == list(w[i].columns) and _same(g[i].to_numpy(), w[i].to_numpy()), (i, cfg, wm)
            assert _same(g[2], w[2]) and _same(g[5], w[5]) and g[6] == w[6]
        # the helpers on this scenario's parameters
        b = rng.uniform(0, 1e6, 4)
        assert _same(orc.rebalance(mine.p, *b), ref._rebalance_portfolio(*[float(v) for v in b]))
    assert failures >= 0 and not math.isnan(float(failures))
