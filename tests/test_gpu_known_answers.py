"""The reference's own zero-volatility known-answer tests (tests/test_simulation_correctness.py),
restated against the drop-in mirror running on the GPU: `_run_single_simulation_path`,
`run_monte_carlo_simulations`, the three private helpers, `_draw_shock_path`."""
from __future__ import annotations

import math

import numpy as np
import pytest

import scenarios
from gpu_util import make_sim
from monte_carlo_retirement_b200.constants import MONTHS_PER_YEAR, SMALL_EPSILON
from monte_carlo_retirement_b200.simulation import median_first_year_withdrawal_rate, trajectory_time_points

pytestmark = pytest.mark.gpu

ZERO = dict(inflation_rate_mean=0.0, inflation_rate_volatility=0.0, inv1_returns_mean=0.0,
            inv1_returns_volatility=0.0, inv2_premium_over_inflation_mean=0.0,
            inv2_premium_over_inflation_volatility=0.0)


def _cfg(**over):
    d = dict(scenarios.TEST_BASE)
    d.update(over)
    return d


def _pension(age, amount=1_000.0):
    return [{"name": "Pension", "monthly_amount_today": amount, "start_at_age": age, "duration_years": None,
             "inflation_indexed": True, "tax_rate": 0.0}]


def test_partial_year_inflation_accrual():  # :84-107
    sim = make_sim(_cfg(**{**ZERO, "inflation_rate_mean": 0.06}, monthly_expenses=0.0, retirement_years=1, seed=7))
    r = sim._run_single_simulation_path(13, path_seed=99)
    assert abs(r["Inflation At Retirement"] - 1.06 ** (13 / MONTHS_PER_YEAR)) < 1e-9
    assert len(trajectory_time_points(13, 1)) == len(r["Trajectory"]) == 4


def test_partial_year_trajectory_keeps_equal_retirement_balance():  # :110-134
    r = make_sim(_cfg(**ZERO, initial_balance=100_000.0, monthly_expenses=1_000.0,
                      retirement_years=1))._run_single_simulation_path(working_months=13, path_seed=1)
    assert r["Trajectory"] == pytest.approx([100_000.0, 100_000.0, 100_000.0, 88_000.0])
    assert r["RealTrajectory"] == pytest.approx(r["Trajectory"])


def test_allocation_weights_conserve_every_dollar():  # :198-217
    r = make_sim(_cfg(**ZERO, initial_balance=100_000.0, allocation_inv1_pct=0.333333, monthly_expenses=0.0,
                      retirement_years=1))._run_single_simulation_path(working_months=0, path_seed=1)
    assert r["Start Balance"] == pytest.approx(100_000.0) and r["Trajectory"][0] == pytest.approx(100_000.0)


def test_income_stream_starts_at_age_and_fractional_month():  # :335-441
    base = _cfg(**ZERO, current_age=40.0, initial_balance=80_000.0, monthly_expenses=1000.0, retirement_years=10,
                other_income_streams=_pension(65.0), seed=1)
    with_p = make_sim(base)._run_single_simulation_path(working_months=240, path_seed=1)
    without = make_sim(dict(base, other_income_streams=[]))._run_single_simulation_path(working_months=240, path_seed=1)
    assert with_p["Final Balance"] > 0 and with_p["Final Balance"] > without["Final Balance"]
    r = make_sim(_cfg(**ZERO, current_age=60.0, initial_balance=6_000.0, monthly_expenses=1_000.0, retirement_years=2,
                      other_income_streams=_pension(60.5), seed=3))._run_single_simulation_path(0, path_seed=4)
    assert r["Success"] is True
    assert r["Final Balance"] == pytest.approx(0.0, abs=1e-6)
    assert r["First Year Gross Withdrawal"] == pytest.approx(6_000.0)


def test_pension_covers_after_portfolio_depleted():  # :444-493
    cfg = _cfg(**ZERO, current_age=60.0, initial_balance=12_000.0, monthly_expenses=1_000.0, retirement_years=10,
               other_income_streams=_pension(61.0), seed=1)
    sim = make_sim(cfg)
    r = sim._run_single_simulation_path(working_months=0, path_seed=1)
    assert r["Success"] is True and r["Final Balance"] == pytest.approx(0.0, abs=1e-6)
    assert make_sim(dict(cfg, other_income_streams=[]))._run_single_simulation_path(0, 1)["Success"] is False
    sim.use_final_seeds()
    summary = sim.run_monte_carlo_simulations(0, 5)[0]
    assert sim._success_probability(summary) == pytest.approx(100.0)
    assert (summary["Final Balance"] <= SMALL_EPSILON).all()


def test_withdrawal_rates():  # :220-256, :496-564
    sim = make_sim(_cfg(**ZERO, initial_balance=200_000.0, monthly_expenses=1_000.0, retirement_years=5, seed=1))
    r = sim._run_single_simulation_path(working_months=0, path_seed=1)
    wr = r["WithdrawalRateTrajectory"]
    expected = r["First Year Gross Withdrawal"] / r["Start Balance"] * 100.0
    assert len(wr) == 5 and wr[0] == pytest.approx(expected, abs=1e-6) and wr[1] == pytest.approx(wr[0], abs=1e-6)
    sim.use_final_seeds()
    summary, _, _, wr_pct, _, _, wr_counts = sim.run_monte_carlo_simulations(working_months=0, num_simulations=10)
    assert wr_counts == [10] * 5
    assert abs(wr_pct.iloc[0][0.50] - expected) < 0.5
    assert abs(median_first_year_withdrawal_rate(summary) - 6.0) < 0.5
    for _, row in summary.iterrows():
        assert abs(row["First Year Gross Withdrawal"] - 12_000.0) < 1.0
    r = make_sim(_cfg(**{**ZERO, "inflation_rate_mean": 0.06, "inv1_returns_mean": 0.06}, initial_balance=240_000.0,
                      monthly_expenses=1_000.0, retirement_years=8, seed=2))._run_single_simulation_path(0, 3)
    assert r["Success"] is True
    for rate in r["WithdrawalRateTrajectory"]:
        assert rate == pytest.approx(r["WithdrawalRateTrajectory"][0], abs=1e-4)
    assert r["WithdrawalRateTrajectory"][0] == pytest.approx(5.0, abs=0.05)


def test_years_to_ruin_and_real_trajectory():  # :567-602
    sim = make_sim(_cfg(**ZERO, initial_balance=5_000.0, monthly_expenses=2_000.0, retirement_years=10, seed=9))
    r = sim._run_single_simulation_path(working_months=0, path_seed=1)
    assert r["Success"] is False and r["YearsToRuin"] == pytest.approx(3 / 12)
    for nom, real in zip(r["Trajectory"], r["RealTrajectory"]):
        assert real == pytest.approx(nom, abs=1e-6)
    summary, traj, _, _, real_traj, _, wr_counts = sim.run_monte_carlo_simulations(0, 20)
    assert (summary["Success"] == False).all() and summary["YearsToRuin"].notna().all()  # noqa: E712
    assert len(real_traj) == len(traj) and wr_counts == [0] * 10


def test_helpers_known_answers():  # :605-662
    sim = make_sim(_cfg(inv1_use_realized_gains_tax_system=True, inv1_realized_gains_tax_rate=0.20))
    assert sim._calculate_withdrawal_and_update(100.0, 0.0, 90.0, True, 0.20) == pytest.approx((0.0, 0.0, 100.0, 80.0))
    assert sim._calculate_withdrawal_and_update(80.0, 100.0, 40.0, True, 0.20) == pytest.approx((40.0, 50.0, 40.0, 40.0))
    sim = make_sim(_cfg(allocation_inv1_pct=0.60, inv1_use_realized_gains_tax_system=True,
                        inv1_realized_gains_tax_rate=0.10, inv2_use_realized_gains_tax_system=True,
                        inv2_realized_gains_tax_rate=0.10))
    b1, cb1, b2, cb2 = sim._rebalance_portfolio(bal_inv1=70.0, cb_inv1=50.0, bal_inv2=30.0, cb_inv2=30.0)
    total = b1 + b2
    assert b1 / total == pytest.approx(0.60, abs=1e-10) and b2 / total == pytest.approx(0.40, abs=1e-10) and total < 100.0
    sale = 70.0 - b1
    br = 50.0 * (sale / 70.0)
    assert cb1 == pytest.approx(50.0 - br) and cb2 == pytest.approx(30.0 + sale - (sale - br) * 0.10)


def test_annual_tax_periods():  # :665-734
    common = dict(initial_balance=100_000.0, monthly_expenses=0.0, retirement_years=1, allocation_inv1_pct=0.50,
                  inv1_returns_mean=0.0, inv1_returns_volatility=0.0, inv2_premium_over_inflation_mean=1.0,
                  inv2_premium_over_inflation_volatility=0.0, inv2_use_realized_gains_tax_system=True,
                  inflation_rate_mean=0.0, inflation_rate_volatility=0.0, seed=11)
    a = make_sim(_cfg(**common, inv1_annual_tax_on_gains_rate=0.0))._run_single_simulation_path(12, 1)
    b = make_sim(_cfg(**common, inv1_annual_tax_on_gains_rate=1.0))._run_single_simulation_path(12, 1)
    assert b["Start Balance"] == pytest.approx(a["Start Balance"], rel=1e-10)
    assert b["Final Balance"] == pytest.approx(a["Final Balance"], rel=1e-10)
    r = make_sim(_cfg(**{**ZERO, "inv1_returns_mean": 0.12}, initial_balance=100.0, monthly_expenses=0.0,
                      retirement_years=1, allocation_inv1_pct=1.0, inv1_annual_tax_on_gains_rate=0.50,
                      seed=12))._run_single_simulation_path(working_months=13, path_seed=1)
    assert r["Start Balance"] == pytest.approx((112.0 - 6.0) * 1.12 ** (1 / 12), rel=1e-10)


def test_native_rng_handles_the_same_deterministic_cases():
    """Zero-volatility scenarios do not depend on the draws: the native-Philox batch (fast and
    strict builds) must give the same known answers as the single-path replay."""
    cfg = _cfg(**ZERO, initial_balance=5_000.0, monthly_expenses=2_000.0, retirement_years=10, seed=9)
    for strict in (False, True):
        sim = make_sim(cfg, strict=strict)
        summary = sim.run_monte_carlo_simulations(0, 64)[0]
        assert (summary["YearsToRuin"] == 0.25).all() and not summary["Success"].any()
    cfg = _cfg(**ZERO, current_age=60.0, initial_balance=12_000.0, monthly_expenses=1_000.0, retirement_years=10,
               other_income_streams=_pension(61.0), seed=1)
    for strict in (False, True):
        summary, traj, *_ = make_sim(cfg, strict=strict).run_monte_carlo_simulations(0, 64)
        assert summary["Success"].all() and (summary["Final Balance"] <= SMALL_EPSILON).all()
        assert traj[0.5].iloc[0] == 12_000.0 and traj[0.5].iloc[1] == pytest.approx(0.0, abs=1e-6)
    assert math.isfinite(make_sim(cfg).run_aggregates(0, 1000)["median_first_year_withdrawal_rate"])
