"""Scenario dictionaries shared by the golden-vector generator and the tests.

`CONFIG_JSON` / `JORGE_JSON` hold the field VALUES of the reference's two shipped scenarios
(/root/reference/config.json, /root/reference/jorge.json — SURVEY §8d configs #1/#2) with the
seed fixed; the others are the variants SURVEY §8d asks for (jorge+, #3, #3-annual, #3-vol) and
a few edge shapes modelled on the reference's own test configs
(tests/test_simulation_correctness.py:20-52).
"""
from __future__ import annotations

import copy

CONFIG_JSON = {
    "scenario": "Macunaima ret plan",
    "initial_balance": 240000.0,
    "monthly_contribution": 5000.0,
    "contribution_growth_rate_annual": 0.04,
    "monthly_expenses": 10000.0,
    "current_age": 40.0,
    "retirement_years": 50,
    "allocation_inv1_pct": 0.60,
    "inv1_returns_mean": 0.12,
    "inv1_returns_volatility": 0.02,
    "inv1_annual_tax_on_gains_rate": 0.0,
    "inv1_realized_gains_tax_rate": 0.10,
    "inv1_use_realized_gains_tax_system": True,
    "inv2_premium_over_inflation_mean": 0.05,
    "inv2_premium_over_inflation_volatility": 0.02,
    "inv2_annual_tax_on_gains_rate": 0.0,
    "inv2_realized_gains_tax_rate": 0.10,
    "inv2_use_realized_gains_tax_system": True,
    "inflation_rate_mean": 0.062,
    "inflation_rate_volatility": 0.0235,
    "num_simulations_main": 1000,
    "num_simulations_search": 300,
    "target_probability": 97.0,
    "starting_working_months_search": 0,
    "seed": 12345,
    "num_processes": 1,
    "other_income_streams": [
        {"name": "State Pension", "monthly_amount_today": 4000.0, "start_at_age": 65.0,
         "duration_years": None, "inflation_indexed": True, "tax_rate": 0.275},
        {"name": "Rental Income (Apt)", "monthly_amount_today": 0.0, "start_at_age": 40.0,
         "duration_years": 35, "inflation_indexed": False, "tax_rate": 0.20},
    ],
}

JORGE_JSON = {
    "scenario": "jorge",
    "initial_balance": 60000,
    "monthly_contribution": 9000,
    "contribution_growth_rate_annual": 0.04,
    "monthly_expenses": 4000,
    "current_age": 35,
    "retirement_years": 40,
    "allocation_inv1_pct": 0.6,
    "inv1_returns_mean": 0.12,
    "inv1_returns_volatility": 0.02,
    "inv1_annual_tax_on_gains_rate": 0,
    "inv1_realized_gains_tax_rate": 0.1,
    "inv1_use_realized_gains_tax_system": True,
    "inv2_premium_over_inflation_mean": 0.05,
    "inv2_premium_over_inflation_volatility": 0.02,
    "inv2_annual_tax_on_gains_rate": 0,
    "inv2_realized_gains_tax_rate": 0.1,
    "inv2_use_realized_gains_tax_system": True,
    "inflation_rate_mean": 0.062,
    "inflation_rate_volatility": 0.0235,
    "num_simulations_main": 1000,
    "num_simulations_search": 100,
    "target_probability": 98,
    "starting_working_months_search": 0,
    "seed": 12345,
    "num_processes": 1,
    "other_income_streams": [
        {"name": "State Pension", "monthly_amount_today": 8400, "start_at_age": 65,
         "duration_years": None, "inflation_indexed": True, "tax_rate": 0.275},
        {"name": "Rental Income (Apt)", "monthly_amount_today": 0, "start_at_age": 40,
         "duration_years": 35, "inflation_indexed": False, "tax_rate": 0.2},
    ],
}


def _derive(base, **over):
    d = copy.deepcopy(base)
    d.update(over)
    return d


# jorge+ : four streams that really exercise the stream logic (SURVEY §8d config #2 note)
JORGE_PLUS = _derive(
    JORGE_JSON,
    scenario="jorge-plus",
    inv1_returns_volatility=0.12,
    other_income_streams=[
        {"name": "State Pension", "monthly_amount_today": 2400.0, "start_at_age": 65.0,
         "duration_years": None, "inflation_indexed": True, "tax_rate": 0.275},
        {"name": "Annuity", "monthly_amount_today": 900.0, "start_at_age": 50.25,
         "duration_years": 20, "inflation_indexed": True, "tax_rate": 0.15},
        {"name": "Rental (fixed nominal)", "monthly_amount_today": 1500.0, "start_at_age": 40.0,
         "duration_years": 35, "inflation_indexed": False, "tax_rate": 0.20},
        {"name": "Royalty (fixed nominal)", "monthly_amount_today": 600.0, "start_at_age": 58.5,
         "duration_years": None, "inflation_indexed": False, "tax_rate": 0.0},
    ],
)

# config #3: the synthetic throughput shape (720 months at wm=240)
SYNTH_C3 = _derive(
    CONFIG_JSON,
    scenario="synthetic-c3",
    retirement_years=40,
    equity_inflation_correlation=-0.5,
    num_simulations_main=1_000_000,
    seed=20260101,
)
SYNTH_C3_ANNUAL = _derive(
    SYNTH_C3,
    scenario="synthetic-c3-annual",
    inv1_use_realized_gains_tax_system=False,
    inv1_annual_tax_on_gains_rate=0.15,
)
SYNTH_C3_VOL = _derive(SYNTH_C3, scenario="synthetic-c3-vol", inv1_returns_volatility=0.15)

# tax-heavy mix: asset 1 on the annual mark-to-market system, asset 2 on the realized system,
# high volatility so annual bills, losses above basis and failures all occur.
TAX_HEAVY = _derive(
    CONFIG_JSON,
    scenario="tax-heavy",
    retirement_years=30,
    monthly_expenses=14000.0,
    inv1_returns_volatility=0.22,
    inv1_use_realized_gains_tax_system=False,
    inv1_annual_tax_on_gains_rate=0.25,
    inv1_realized_gains_tax_rate=0.0,
    inv2_realized_gains_tax_rate=0.22,
    inv2_premium_over_inflation_volatility=0.06,
    equity_inflation_correlation=0.35,
    seed=777,
)

# both assets on the annual system, one with a realized-rate of zero but the flag on
ANNUAL_BOTH = _derive(
    CONFIG_JSON,
    scenario="annual-both",
    retirement_years=25,
    inv1_returns_volatility=0.18,
    inv1_use_realized_gains_tax_system=False,
    inv1_annual_tax_on_gains_rate=0.20,
    inv2_use_realized_gains_tax_system=False,
    inv2_annual_tax_on_gains_rate=0.30,
    monthly_expenses=9000.0,
    equity_inflation_correlation=-1.0,
    seed=4242,
)

# the reference tests' base config (tests/test_simulation_correctness.py:20-52): no tax at all
TEST_BASE = {
    "scenario": "test",
    "initial_balance": 500_000.0,
    "monthly_contribution": 0.0,
    "contribution_growth_rate_annual": 0.0,
    "monthly_expenses": 2_000.0,
    "current_age": 40.0,
    "retirement_years": 10,
    "allocation_inv1_pct": 0.6,
    "inv1_returns_mean": 0.08,
    "inv1_returns_volatility": 0.15,
    "inv1_annual_tax_on_gains_rate": 0.0,
    "inv1_realized_gains_tax_rate": 0.0,
    "inv1_use_realized_gains_tax_system": False,
    "inv2_premium_over_inflation_mean": 0.02,
    "inv2_premium_over_inflation_volatility": 0.01,
    "inv2_annual_tax_on_gains_rate": 0.0,
    "inv2_realized_gains_tax_rate": 0.0,
    "inv2_use_realized_gains_tax_system": False,
    "inflation_rate_mean": 0.03,
    "inflation_rate_volatility": 0.01,
    "equity_inflation_correlation": 0.0,
    "num_simulations_main": 50,
    "num_simulations_search": 40,
    "target_probability": 80.0,
    "starting_working_months_search": 0,
    "seed": 42,
    "num_processes": 1,
    "other_income_streams": [],
}

# a stressed no-tax scenario where a large share of paths fail (ruin months, NaN WR years)
STRESSED = _derive(
    TEST_BASE,
    scenario="stressed",
    initial_balance=300_000.0,
    monthly_contribution=1_500.0,
    contribution_growth_rate_annual=0.03,
    monthly_expenses=4_500.0,
    retirement_years=30,
    inv1_returns_volatility=0.20,
    equity_inflation_correlation=1.0,
    allocation_inv1_pct=0.85,
    seed=99,
)

# an all-in-asset-2 / zero-allocation corner and a broke start
CORNER_ALLOC0 = _derive(TEST_BASE, scenario="alloc0", allocation_inv1_pct=0.0,
                        inv2_use_realized_gains_tax_system=True, inv2_realized_gains_tax_rate=0.15,
                        monthly_contribution=800.0, seed=5)
CORNER_BROKE = _derive(TEST_BASE, scenario="broke", initial_balance=0.0, monthly_contribution=0.0,
                       monthly_expenses=1000.0, retirement_years=3, seed=6)

# name -> (config dict, [(stream, working_months, n_paths), ...])
GOLDEN_CASES = {
    "config_json": (CONFIG_JSON, [("final", 233, 24), ("search", 0, 16), ("search", 168, 24), ("final", 7, 8)]),
    "jorge_json": (JORGE_JSON, [("final", 75, 24), ("search", 48, 24), ("final", 13, 8)]),
    "jorge_plus": (JORGE_PLUS, [("final", 75, 24), ("search", 30, 24), ("final", 181, 16)]),
    "synth_c3": (SYNTH_C3, [("final", 240, 24), ("final", 100, 16)]),
    "synth_c3_annual": (SYNTH_C3_ANNUAL, [("final", 240, 24), ("search", 126, 16)]),
    "synth_c3_vol": (SYNTH_C3_VOL, [("final", 240, 24), ("search", 60, 24)]),
    "tax_heavy": (TAX_HEAVY, [("final", 200, 32), ("search", 90, 32), ("final", 0, 16), ("final", 125, 16)]),
    "annual_both": (ANNUAL_BOTH, [("final", 180, 32), ("search", 61, 24)]),
    "test_base": (TEST_BASE, [("final", 0, 16), ("final", 12, 16), ("search", 37, 16)]),
    "stressed": (STRESSED, [("final", 0, 32), ("final", 60, 32), ("search", 119, 32)]),
    "alloc0": (CORNER_ALLOC0, [("final", 24, 12), ("final", 5, 12)]),
    "broke": (CORNER_BROKE, [("final", 0, 4), ("final", 11, 4)]),
}

# search goldens: (config dict, num_simulations_search override or None)
SEARCH_CASES = {
    "config_json": (CONFIG_JSON, None),
    "jorge_json": (JORGE_JSON, None),
    "stressed": (_derive(STRESSED, target_probability=70.0, num_simulations_search=200), None),
    "tax_heavy": (_derive(TAX_HEAVY, target_probability=85.0, num_simulations_search=150), None),
    "unreachable": (_derive(CORNER_BROKE, target_probability=99.0, num_simulations_search=20,
                            retirement_years=40), None),
}
