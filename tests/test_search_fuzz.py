"""Randomised check of the search driver (simulation.py:1138-1342 of the reference) on the CPU.

For random success tables k(m) — monotone, noisy / non-monotone, never reaching the target,
reaching it at once — the drop-in's `find_minimum_working_months` must take the reference's
decisions whatever way its probabilities are obtained:
  * `sequential`: one (replaced) `run_monte_carlo_simulations` per probe, as the reference tests do;
  * `waves` / `probe` / `grid`: success COUNTS from the batched search kernel — replaced here by a
    table lookup, so the host-side speculation (which candidates each launch evaluates, how the
    decisions are replayed over the tables) is what is under test.
The oracle's restatement of the procedure arbitrates; where the reference tree is mounted the
unmodified reference runs on the same table too (months, probability, curve and progress events).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import pandas as pd
import pytest
import torch

import scenarios
from monte_carlo_retirement_b200.config import Config
from monte_carlo_retirement_b200.simulation import RetirementMonteCarloSimulator
from oracle import oracle as orc

REF = "/root/reference/backend"


def _table(rng: np.random.Generator, n: int, start: int):
    """successes out of n for every month up to the search limit."""
    months = np.arange(0, start + 70 * 12 + 1)
    kind = rng.integers(0, 5)
    centre = start + rng.uniform(-20, 500)
    width = rng.uniform(2, 80)
    p = 1 / (1 + np.exp(-(months - centre) / width))
    if kind == 1:
        p = np.clip(p + rng.normal(0, 0.03, len(p)), 0, 1)      # noisy: non-monotone probes
    elif kind == 2:
        p = p * rng.uniform(0.3, 0.9)                           # saturates below the target
    elif kind == 3:
        p = np.clip(p + 0.5, 0, 1)                              # succeeds early
    return np.rint(p * n).astype(int)


def _df(k: int, n: int) -> pd.DataFrame:
    flags = np.zeros(n, dtype=bool)
    flags[:k] = True
    return pd.DataFrame({"Start Balance": np.full(n, 100.0), "Final Balance": flags.astype(float), "Success": flags})


def _reference_search(cfg: dict, k, n: int):
    saved_path = list(sys.path)
    saved = {m: sys.modules.pop(m, None) for m in ("config", "constants", "simulation", "utils")}
    sys.path.insert(0, REF)
    try:
        from loguru import logger

        logger.remove()
        import config as ref_config
        import simulation as ref_simulation

        sim = ref_simulation.RetirementMonteCarloSimulator(ref_config.Config(**cfg))
        sim.run_monte_carlo_simulations = lambda wm, num: (_df(int(k[wm]), num), None, None, None, None, None, None)
        events = []
        months, prob, curve = sim.find_minimum_working_months(verbose=False, progress_callback=events.append)
        return months, prob, curve, events
    finally:
        sys.path[:] = saved_path
        for m, v in saved.items():
            sys.modules.pop(m, None)
            if v is not None:
                sys.modules[m] = v


@pytest.mark.parametrize("block", range(5))
def test_search_policies_take_the_reference_decisions_on_random_tables(block, monkeypatch):
    rng = np.random.default_rng(777 + block)
    for _ in range(8):
        n = int(rng.choice([10, 40, 300, 1000]))
        start = int(rng.choice([0, 0, 3, 24, 120]))
        target = float(rng.choice([50.0, 80.0, 90.0, 97.0, 99.5]))
        cfg = dict(scenarios.TEST_BASE, num_simulations_search=n, starting_working_months_search=start,
                   target_probability=target)
        k = _table(rng, n, start)
        want = orc.search_decisions(lambda m: float(int(k[m]) / n * 100.0), start, target, n)
        results = {}
        # sequential: through a replaced run_monte_carlo_simulations
        sim = RetirementMonteCarloSimulator(Config(**cfg))
        sim.run_monte_carlo_simulations = lambda wm, num: (_df(int(k[wm]), num), None, None, None, None, None, None)
        ev = []
        results["sequential"] = (*sim.find_minimum_working_months(verbose=False, progress_callback=ev.append), ev)
        assert sim.last_search_stats["policy"] == "sequential"
        # device policies: the batched kernel replaced by the table
        for policy in ("waves", "probe", "grid"):
            sim = RetirementMonteCarloSimulator(Config(**cfg), search_policy=policy)
            launches = []

            def counts(candidates, num_simulations, *, first_path=0, with_executed=False, _l=launches):
                _l.append(len(candidates))
                assert num_simulations == n and first_path == 0
                return torch.tensor([int(k[c]) for c in candidates], dtype=torch.int64)

            monkeypatch.setattr(sim, "batched_success_counts", counts)
            ev = []
            results[policy] = (*sim.find_minimum_working_months(verbose=False, progress_callback=ev.append), ev)
            assert sim.last_search_stats["policy"] == policy
            if policy == "grid":
                assert launches[0] == 601   # one launch over start..start+600; later probes (rare) one by one
        for name, (months, prob, curve, ev) in results.items():
            assert (months, prob, curve) == want[:3], (name, cfg, months, want[0])
            assert [e["working_months"] for e in ev if e["type"] == "search_iter"] == want[3], name
            assert ev == results["sequential"][3], name               # identical progress events for every policy
        if os.path.isdir(REF):
            r_months, r_prob, r_curve, r_events = _reference_search(cfg, k, n)
            assert (r_months, r_prob, r_curve) == want[:3]
            assert r_events == results["sequential"][3]
