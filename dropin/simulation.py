"""Flat-module shim: put this directory on sys.path AHEAD of the reference's backend/ (or copy
this file over backend/simulation.py) and `from simulation import ...` in backend/main.py,
backend/server.py, backend/plotting.py and tests/test_simulation_correctness.py resolves to the
B200 engine with the reference's names."""
from monte_carlo_retirement_b200.simulation import (  # noqa: F401
    RetirementMonteCarloSimulator,
    age_at_retirement_year,
    arithmetic_to_log_params,
    median_first_year_withdrawal_rate,
    retirement_age,
    stream_payment_start_age,
    stream_payment_start_month_index,
    trajectory_time_points,
    years_from_t0_to_age,
)
from monte_carlo_retirement_b200.constants import MONTHS_PER_YEAR, SMALL_EPSILON  # noqa: F401
