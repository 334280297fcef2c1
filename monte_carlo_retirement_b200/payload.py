"""Response payload of one final run.

`build_result` assembles the dict the reference's `backend/server.py:_build_result` (416-565)
returns for `SimulationResponse` (server.py:100-109), in two modes:

  * "legacy"    — for the N the reference ships with. Calls the simulator's
                  `run_monte_carlo_simulations(working_months=..., num_simulations=...)` (keyword
                  call, as server.py:431-434 does, so the reference's FakeSimulator tests still
                  apply) and produces the reference's dict key for key, including the three O(N)
                  lists of `histogram` (server.py:553-563) and the O(failures) ruin list (:527-531).
  * "aggregate" — for large N (the lists are what breaks first above ~1e6 paths). Everything is
                  reduced on the device (`RetirementMonteCarloSimulator.aggregates_device`): the
                  summary block, the band tables, the 5 sample paths, and — instead of the lists —
                  the charts' own bins: `histogram.binned` holds what the dashboard's `binData`
                  (frontend/src/components/HistogramChart.jsx:13-60) computes from the lists and
                  `ruin_histogram.bins` what `binRuinYears` (RuinHistogramChart.jsx:12-29) computes.
                  The list fields stay present and empty, so the reference's response model
                  validates unchanged; a few KB leave the GPU regardless of N.

"auto" picks aggregate mode above `AGGREGATE_THRESHOLD` paths when the simulator offers it.
"""
from __future__ import annotations

import math
from decimal import ROUND_HALF_UP, Decimal
from typing import Any, Dict, List, Optional, Sequence

from .constants import MONTHS_PER_YEAR, SMALL_EPSILON
from .simulation import (FINAL_BALANCE_QUANTILES, median_first_year_withdrawal_rate, retirement_age,
                         stream_payment_start_month_index, trajectory_time_points)

AGGREGATE_THRESHOLD = 200_000   # paths; above this "auto" stops shipping per-path lists
HISTOGRAM_BINS = 60             # HistogramChart.jsx:13


# ---- small formatting helpers -------------------------------------------------------------------
def _cents(v: float) -> float:
    return round(float(v), 2)


def _finite_or_none(v: float, digits: int) -> Optional[float]:
    v = float(v)
    return None if (math.isnan(v) or math.isinf(v)) else round(v, digits)


def _pkey(q: float) -> str:
    return f"p{int(q * 100)}"   # server.py:214,451,506 — int() truncates exactly like the reference


def _js_fixed1(x: float) -> str:
    """Number.prototype.toFixed(1): exact decimal expansion, ties away from zero."""
    return str(Decimal(x).quantize(Decimal("0.1"), rounding=ROUND_HALF_UP))


# ---- blocks shared by both modes ----------------------------------------------------------------
def _band_block(table, samples, years: Sequence[float]) -> Optional[dict]:
    """server.py:205-228 — percentile series keyed p5..p95 plus the sampled paths, to cents."""
    if table is None or table.empty:
        return None
    if len(years) != len(table):
        raise ValueError("Trajectory time-point count does not match trajectory data "
                         f"({len(years)} != {len(table)}).")
    series = {_pkey(q): [_cents(v) for v in table[q]] for q in table.columns}
    paths = [[_cents(v) for v in path] for path in samples] if samples else []
    return {"years": list(years), "percentiles": series, "sample_paths": paths}


def _withdrawal_rate_block(table, counts, first_year: float, total_paths: int) -> Optional[dict]:
    """server.py:496-515 — NaN rows (no surviving path that year) become null, 3 decimals."""
    if table is None or table.empty:
        return None
    series = {_pkey(q): [_finite_or_none(v, 3) if v is not None else None for v in table[q]]
              for q in table.columns}
    return {"years": [first_year + i for i in range(len(table))], "percentiles": series,
            "observation_counts": counts or [], "total_paths": int(total_paths)}


def _reference_lines(config, working_months: int) -> List[dict]:
    """server.py:474-494 — retirement start plus one marker per paying income stream."""
    t_ret = working_months / MONTHS_PER_YEAR
    lines = [{"name": "Retirement Starts", "year": t_ret}]
    for s in (config.other_income_streams or []):
        if s.monthly_amount_today <= SMALL_EPSILON or s.duration_years == 0:
            continue
        first = stream_payment_start_month_index(config.current_age, working_months, s.start_at_age)
        lines.append({"name": s.name, "year": round(t_ret + first / MONTHS_PER_YEAR, 3)})
    return lines


def _search_curve_block(config, working_months: int, curve) -> Optional[dict]:
    """server.py:197-202,517-523 — last probability seen per candidate, ascending."""
    if not curve:
        return None
    latest: Dict[int, dict] = {}
    for point in curve:
        latest[int(point["working_months"])] = point
    return {"points": [latest[m] for m in sorted(latest)], "target_probability": config.target_probability,
            "selected_working_months": working_months}


def _summary_block(config, working_months: int, estimated: bool, success_pct: float, median_start: float,
                   median_final_ok: float, swr: float, final_quantiles: Dict[float, float]) -> dict:
    return {
        "required_working_months": working_months,
        "required_working_years": round(working_months / MONTHS_PER_YEAR, 1),
        "working_period_is_estimate": bool(estimated),
        "retirement_age": round(retirement_age(config.current_age, working_months), 1),
        "success_probability": round(float(success_pct), 2),
        "target_probability": config.target_probability,
        "median_start_balance": _cents(median_start),
        "median_final_balance_successful": _cents(median_final_ok),
        "swr": _finite_or_none(swr, 2),
        "final_balance_percentiles": {_pkey(q): _cents(max(0.0, float(v))) for q, v in final_quantiles.items()},
    }


# ---- the charts' own binning, from device histograms (aggregate mode) -----------------------------
def balance_bins(lo: float, hi: float, counts: Sequence[int], median: float, successful: int, total: int) -> dict:
    """What HistogramChart.jsx:binData returns, from the 60-bin floor histogram of the successful
    cohort's final balances over [lo, hi]."""
    rate = f"{_js_fixed1(successful / total * 100.0)}" if total else "0.0"
    if successful == 0:
        return {"bins": [], "median": 0, "successRate": rate}
    if hi <= lo:
        return {"bins": [{"label": f"${_js_fixed1(lo / 1e6)}M", "count": int(successful), "mid": lo / 1e6}],
                "median": median / 1e6, "successRate": rate}
    nb = len(counts)
    width = (hi - lo) / nb
    bins = []
    for i, c in enumerate(counts):
        mid = ((lo + i * width) + (lo + (i + 1) * width)) / 2 / 1e6
        bins.append({"label": f"${_js_fixed1(mid)}M", "count": int(c), "mid": mid})
    return {"bins": bins, "median": median / 1e6, "successRate": rate}


def ruin_year_bins(ruin_month_counts: Sequence[int]) -> List[dict]:
    """What RuinHistogramChart.jsx:binRuinYears returns, from the ruin-month histogram: a failure
    in retirement month m (YearsToRuin = m/12, shipped rounded to 0.1) lands in year
    max(ceil(m/12), 1); the last year shown is the last one with a failure."""
    months = [m for m, c in enumerate(ruin_month_counts) if c]
    if not months:
        return []
    year_of = lambda m: max(math.ceil(round(m / MONTHS_PER_YEAR, 1)), 1)  # noqa: E731
    per_year = [0] * year_of(months[-1])
    for m in months:
        per_year[year_of(m) - 1] += int(ruin_month_counts[m])
    return [{"year": i + 1, "label": str(i + 1), "count": c} for i, c in enumerate(per_year)]


# ---- the two modes ------------------------------------------------------------------------------
def _legacy(config, simulator, working_months: int, search_curve) -> dict:
    (summary, bands, samples, wr_bands, real_bands, real_samples, wr_counts) = simulator.run_monte_carlo_simulations(
        working_months=working_months, num_simulations=config.num_simulations_main)
    if summary.empty:
        raise ValueError(f"Simulation for '{config.Nickname}' yielded no results.")
    n = len(summary)
    final = summary["Final Balance"]
    ok = summary["Success"].astype(bool) if "Success" in summary.columns else final > SMALL_EPSILON
    final_ok = final[ok]
    quantiles = final.quantile(list(FINAL_BALANCE_QUANTILES))
    years = trajectory_time_points(working_months, config.retirement_years)
    ruin = None
    if "YearsToRuin" in summary.columns:
        failed_years = summary["YearsToRuin"][~ok].dropna()
        ruin = {"years_to_ruin": [round(float(v), 1) for v in failed_years],
                "failure_count": int(len(failed_years)), "total_paths": int(n)}
    return {
        "scenario": config.Nickname,
        "summary": _summary_block(config, working_months, bool(search_curve), ok.mean() * 100.0,
                                  float(summary["Start Balance"].median()),
                                  float(final_ok.median()) if len(final_ok) else 0.0,
                                  median_first_year_withdrawal_rate(summary),
                                  {float(q): float(v) for q, v in quantiles.items()}),
        "trajectory": _band_block(bands, samples, years),
        "trajectory_real": _band_block(real_bands, real_samples, years),
        "withdrawal_rate": _withdrawal_rate_block(wr_bands, wr_counts, working_months / MONTHS_PER_YEAR, n),
        "search_curve": _search_curve_block(config, working_months, search_curve),
        "ruin_histogram": ruin,
        "histogram": {"final_balances": [_cents(v) for v in final],
                      "start_balances": [_cents(v) for v in summary["Start Balance"]],
                      "success_flags": [bool(v) for v in summary["Success"]]},
        "reference_lines": _reference_lines(config, working_months),
    }


def _aggregate(config, simulator, working_months: int, search_curve) -> dict:
    n = int(config.num_simulations_main)
    if n <= 0:
        raise ValueError(f"Simulation for '{config.Nickname}' yielded no results.")
    a = simulator.run_aggregates(working_months, n, bands=True, samples=True)
    years = trajectory_time_points(working_months, config.retirement_years)
    n_ok = int(a["success_count"])
    h60 = a["final_balance_hist_60"]
    ruin_counts = a["ruin_month_hist"]
    return {
        "scenario": config.Nickname,
        "summary": _summary_block(config, working_months, bool(search_curve), a["success_probability"],
                                  a["median_start_balance"], a["median_final_balance_successful"],
                                  a["median_first_year_withdrawal_rate"], a["final_balance_quantiles"]),
        "trajectory": _band_block(a["trajectory_bands"], a["sample_paths"], years),
        "trajectory_real": _band_block(a["real_trajectory_bands"], a["real_sample_paths"], years),
        "withdrawal_rate": _withdrawal_rate_block(a["withdrawal_rate_bands"], a["withdrawal_rate_counts"],
                                                  working_months / MONTHS_PER_YEAR, n),
        "search_curve": _search_curve_block(config, working_months, search_curve),
        "ruin_histogram": {"years_to_ruin": [], "failure_count": int(sum(ruin_counts)), "total_paths": n,
                           "bins": ruin_year_bins(ruin_counts)},
        "histogram": {"final_balances": [], "start_balances": [], "success_flags": [],
                      "binned": balance_bins(h60["range"][0], h60["range"][1], h60["counts"],
                                             a["median_final_balance_successful"], n_ok, n)},
        "reference_lines": _reference_lines(config, working_months),
    }


def build_result(config, simulator, required_w_months: int, search_curve: Optional[List[dict]] = None, *,
                 mode: str = "auto", aggregate_threshold: int = AGGREGATE_THRESHOLD) -> Dict[str, Any]:
    """Run the final simulation and assemble the response dict — server.py:416-565.

    Same positional signature as the reference's `_build_result`; a maintainer switches with
    `from monte_carlo_retirement_b200.payload import build_result as _build_result`.
    """
    if mode not in ("auto", "legacy", "aggregate"):
        raise ValueError("mode must be 'auto', 'legacy' or 'aggregate'")
    if mode == "auto":
        big = int(config.num_simulations_main) > int(aggregate_threshold)
        mode = "aggregate" if big and hasattr(simulator, "run_aggregates") else "legacy"
    if mode == "aggregate":
        return _aggregate(config, simulator, int(required_w_months), search_curve)
    return _legacy(config, simulator, required_w_months, search_curve)
