"""CLI outputs from device aggregates.

The reference's command line (`backend/main.py:112-171`) re-reduces the N-row `summary_df` on
the host three times: for the result log (`backend/utils.py:69-102`), for the text block of the
histogram figure (`backend/plotting.py:100-108`) and for the 100-bin histogram itself
(`plotting.py:46-59`). These helpers produce the same numbers and the same text from the dict
`RetirementMonteCarloSimulator.run_aggregates` returns — a few KB, whatever N is — so the CLI
can report a 1e8-path final run without materialising a frame.
"""
from __future__ import annotations

from typing import Any, Dict, List, Tuple

import numpy as np

from .constants import MONTHS_PER_YEAR
from .simulation import FINAL_BALANCE_QUANTILES


def analysis_summary(working_months: int, aggregates: Dict[str, Any]) -> Dict[str, Any]:
    """The `analysis_summary_for_plot` dict of main.py:139-145."""
    return {
        "required_working_months": working_months,
        "final_success_probability": aggregates["success_probability"],
        "median_start_retirement_balance": aggregates["median_start_balance"],
        "median_final_balance": aggregates["median_final_balance_successful"],
        "SWR": aggregates["median_first_year_withdrawal_rate"],
    }


def result_log_lines(config, working_months: int, aggregates: Dict[str, Any]) -> List[str]:
    """The messages `log_simulation_results` (utils.py:69-102) logs, in order."""
    a = aggregates
    lines = [
        f"--- Final Simulation Results for Scenario: '{config.Nickname}' ---",
        f"Determined Required Working Months: {working_months} ({working_months / MONTHS_PER_YEAR:.1f} years)",
        f"Probability of Not Running Out of Money (Final Sims): {a['success_probability']:.2f}% "
        f"(Target: {config.target_probability:.2f}%)",
        f"Median Balance at Start of Retirement (All Sims): ${a['median_start_balance']:,.2f}",
        f"Median Final Balance (Successful Sims Only): ${a['median_final_balance_successful']:,.2f}",
        "Est. First-year Real Withdrawal Rate (median, real gross / start bal): "
        f"{a['median_first_year_withdrawal_rate']:.2f}%",
        "Final Balance Percentiles (All Sims, $):",
    ]
    for q in FINAL_BALANCE_QUANTILES:
        lines.append(f"  {q * 100:.0f}th: {max(0, a['final_balance_quantiles'][q]):,.2f}")
    return lines


def final_balance_histogram(aggregates: Dict[str, Any]) -> Tuple[np.ndarray, np.ndarray]:
    """(counts, edges) of `plt.hist(successful final balances / 1e6, bins=100)` (plotting.py:46-59),
    i.e. numpy.histogram over the cohort's own [min, max]; draw it with
    `plt.stairs(counts, edges, fill=True)`. Empty cohort: (zeros(100), edges of [0, 1])."""
    h = aggregates["final_balance_hist_musd_100"]
    counts = np.asarray(h["counts"], dtype=np.int64)
    lo, hi = h["range"]
    if aggregates["success_count"] == 0 or not (lo == lo and hi == hi):
        return np.zeros(100, dtype=np.int64), np.linspace(0.0, 1.0, 101)
    if lo == hi:  # numpy widens a degenerate range by +-0.5
        lo, hi = lo - 0.5, hi + 0.5
    return counts, np.linspace(lo, hi, 101)


def histogram_label(aggregates: Dict[str, Any]) -> str:
    """Legend label of the histogram (plotting.py:39-43,57)."""
    n = aggregates["num_simulations"]
    rate = aggregates["success_count"] / n * 100 if n > 0 else 0.0
    return f"Successful Outcomes ({rate:.1f}%)"


def figure_result_lines(config, working_months: int, aggregates: Dict[str, Any]) -> List[str]:
    """The '--- Results ---' text block of the histogram figure (plotting.py:100-108)."""
    s = analysis_summary(working_months, aggregates)
    return [
        "--- Results ---",
        f"Req.Work: {s['required_working_months']}mo ({s['required_working_months'] / MONTHS_PER_YEAR:.1f}yr)",
        f"Success: {s['final_success_probability']:.1f}% (Target: {config.target_probability:.1f}%)",
        f"Med Start Bal: ${s['median_start_retirement_balance']:,.0f}",
        f"Med Final Bal (succ): ${s['median_final_balance']:,.0f}",
        f"SWR (1st-yr withdraw rate): {s['SWR']:.2f}%",
    ]
