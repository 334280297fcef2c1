"""Response payload (monte_carlo_retirement_b200/payload.py) — CPU part.

The legacy mode is host code over the simulator's 7-tuple, so it can be pinned here without a
GPU: fed by the CPU oracle (itself pinned bit-exact to the reference), `build_result` must
reproduce, key for key and digit for digit, the dicts the reference's own
`server._build_result` produced (tests/golden/payload.json, made by
tests/golden/make_payload_golden.py). The chart-binning helpers of the aggregate-only mode are
checked against a restatement of the dashboard's JavaScript.
"""
from __future__ import annotations

import json
import math
import os

import numpy as np
import pytest

from monte_carlo_retirement_b200 import payload
from monte_carlo_retirement_b200.config import Config
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "payload.json")))


class OracleBackedSimulator:
    """The one method `_build_result` calls, answered by the CPU oracle."""

    def __init__(self, cfg_dict):
        self.sim = orc.OracleSimulator(cfg_dict, n_threads=2)
        self.sim.use_final_seeds()

    def run_monte_carlo_simulations(self, working_months, num_simulations):
        return self.sim.run(working_months, num_simulations)


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_legacy_payload_equals_reference_build_result(name):
    g = GOLDEN[name]
    cfg = Config(**g["cfg"])
    got = payload.build_result(cfg, OracleBackedSimulator(g["cfg"]), g["working_months"],
                               search_curve=g["search_curve"], mode="legacy")
    got = json.loads(json.dumps(got, allow_nan=False))
    want = g["result"]
    assert got.keys() == want.keys()
    for key in want:
        assert got[key] == want[key], key


def test_auto_mode_is_legacy_for_small_n_and_for_simulators_without_aggregates():
    g = GOLDEN["test_base_override"]
    cfg = Config(**g["cfg"])
    got = payload.build_result(cfg, OracleBackedSimulator(g["cfg"]), g["working_months"])
    assert got["histogram"]["final_balances"] == g["result"]["histogram"]["final_balances"]
    big = Config(**dict(g["cfg"], num_simulations_main=payload.AGGREGATE_THRESHOLD + 1))
    calls = []

    class Fake:
        def run_monte_carlo_simulations(self, **kw):
            calls.append(kw)
            raise RuntimeError("stop here")

    with pytest.raises(RuntimeError, match="stop here"):
        payload.build_result(big, Fake(), 13)
    assert calls == [{"working_months": 13, "num_simulations": payload.AGGREGATE_THRESHOLD + 1}]
    with pytest.raises(ValueError):
        payload.build_result(cfg, Fake(), 13, mode="fastest")


def test_empty_result_is_a_value_error():
    g = GOLDEN["test_base_override"]
    cfg = Config(**g["cfg"])

    class Empty:
        def run_monte_carlo_simulations(self, **kw):
            import pandas as pd

            return pd.DataFrame(), None, None, None, None, None, None

    with pytest.raises(ValueError, match="yielded no results"):
        payload.build_result(cfg, Empty(), 13, mode="legacy")


# ---- the dashboard's binning, restated from the JavaScript for the check ---------------------------
def js_bin_data(values, flags, num_bins=60):
    """HistogramChart.jsx:13-60."""
    ok = [v for v, f in zip(values, flags) if f]
    rate = payload._js_fixed1(len(ok) / len(values) * 100) if values else "0.0"
    if not ok:
        return {"bins": [], "median": 0, "successRate": rate}
    s = sorted(ok)
    mid = len(s) // 2
    median = s[mid] if len(s) % 2 else (s[mid - 1] + s[mid]) / 2
    lo, hi = s[0], s[-1]
    if hi <= lo:
        return {"bins": [{"label": f"${payload._js_fixed1(lo / 1e6)}M", "count": len(ok), "mid": lo / 1e6}],
                "median": median / 1e6, "successRate": rate}
    width = (hi - lo) / num_bins
    counts = [0] * num_bins
    for v in ok:
        counts[min(math.floor((v - lo) / width), num_bins - 1)] += 1
    bins = []
    for i, c in enumerate(counts):
        m = ((lo + i * width) + (lo + (i + 1) * width)) / 2 / 1e6
        bins.append({"label": f"${payload._js_fixed1(m)}M", "count": c, "mid": m})
    return {"bins": bins, "median": median / 1e6, "successRate": rate}


def js_bin_ruin_years(years):
    """RuinHistogramChart.jsx:12-29 (without the chart's crash on an all-zero list)."""
    if not years:
        return []
    n = max(math.ceil(max(years)), 1)
    counts = [0] * n
    for y in years:
        counts[min(max(math.ceil(y) - 1, 0), n - 1)] += 1
    last = n - 1
    while last > 0 and counts[last] == 0:
        last -= 1
    return [{"year": i + 1, "label": str(i + 1), "count": counts[i]} for i in range(last + 1)]


def test_ruin_year_bins_match_the_dashboard_binning():
    rng = np.random.default_rng(5)
    for horizon in (1, 12, 13, 40 * 12):
        for _ in range(6):
            months = rng.integers(0, horizon + 1, size=rng.integers(1, 400))
            hist = np.bincount(months, minlength=horizon + 1)
            years = [round(float(m) / 12, 1) for m in months]          # server.py:528
            assert payload.ruin_year_bins(hist.tolist()) == js_bin_ruin_years(years)
    assert payload.ruin_year_bins([0] * 30) == []
    assert payload.ruin_year_bins([7]) == [{"year": 1, "label": "1", "count": 7}]


def test_balance_bins_match_the_dashboard_binning():
    rng = np.random.default_rng(6)
    for n, p_ok in ((500, 0.9), (61, 1.0), (3, 0.5), (40, 0.0)):
        values = (rng.lognormal(14, 1.0, n)).tolist()
        flags = (rng.random(n) < p_ok).tolist()
        want = js_bin_data(values, flags)
        ok = sorted(v for v, f in zip(values, flags) if f)
        if ok:
            lo, hi = ok[0], ok[-1]
            width = (hi - lo) / 60
            counts = [0] * 60
            for v in ok:
                counts[min(math.floor((v - lo) / width), 59)] += 1
            median = float(np.median(ok))
        else:
            lo = hi = float("nan")
            counts, median = [0] * 60, 0.0
        got = payload.balance_bins(lo, hi, counts, median, len(ok), n)
        assert got == want
    one = payload.balance_bins(2.5e6, 2.5e6, [4] + [0] * 59, 2.5e6, 4, 5)
    assert one == {"bins": [{"label": "$2.5M", "count": 4, "mid": 2.5}], "median": 2.5, "successRate": "80.0"}
    assert payload._js_fixed1(0.25) == "0.3" and payload._js_fixed1(2.675) == "2.7"


# ---- CLI outputs from aggregates (report.py) -------------------------------------
def host_aggregates(summary):
    """The subset of `run_aggregates()` the report helpers read, reduced on the host with the
    pandas / numpy calls the reference's CLI makes (main.py:112-133, utils.py:97-99)."""
    from monte_carlo_retirement_b200.simulation import FINAL_BALANCE_QUANTILES, median_first_year_withdrawal_rate

    ok = summary["Success"].astype(bool)
    succ = summary.loc[ok, "Final Balance"]
    q = summary["Final Balance"].quantile(FINAL_BALANCE_QUANTILES)
    musd = succ.to_numpy() / 1e6
    if len(musd):
        counts, edges = np.histogram(musd, bins=100)
        hist = {"range": [float(musd.min()), float(musd.max())], "counts": counts.tolist()}
    else:
        hist = {"range": [float("nan"), float("nan")], "counts": [0] * 100}
    return {
        "num_simulations": len(summary), "success_count": int(ok.sum()),
        "success_probability": float(ok.mean() * 100.0),
        "median_start_balance": float(summary["Start Balance"].median()),
        "median_final_balance_successful": float(succ.median()) if len(succ) else 0.0,
        "median_first_year_withdrawal_rate": median_first_year_withdrawal_rate(summary),
        "final_balance_quantiles": {float(k): float(v) for k, v in q.items()},
        "final_balance_hist_musd_100": hist,
    }


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_cli_report_from_aggregates_equals_reference_log_and_histogram(name):
    from monte_carlo_retirement_b200 import report

    g = GOLDEN[name]
    cfg = Config(**g["cfg"])
    summary = OracleBackedSimulator(g["cfg"]).run_monte_carlo_simulations(g["working_months"],
                                                                          cfg.num_simulations_main)[0]
    agg = host_aggregates(summary)
    assert report.result_log_lines(cfg, g["working_months"], agg) == g["cli"]["log"]
    counts, edges = report.final_balance_histogram(agg)
    if g["cli"]["hist100"] is None:
        assert counts.sum() == 0 and len(edges) == 101
    else:
        assert counts.tolist() == g["cli"]["hist100"]["counts"]
        assert edges.tolist() == g["cli"]["hist100"]["edges"]
    s = report.analysis_summary(g["working_months"], agg)
    assert set(s) == {"required_working_months", "final_success_probability", "median_start_retirement_balance",
                      "median_final_balance", "SWR"}
    fig = report.figure_result_lines(cfg, g["working_months"], agg)
    assert fig[0] == "--- Results ---" and fig[1].startswith(f"Req.Work: {g['working_months']}mo")
    assert report.histogram_label(agg) == f"Successful Outcomes ({agg['success_probability']:.1f}%)"


# ---- the reference's own API tests, restated against payload.build_result -------------------------
def test_outcomes_keep_success_flags_and_zero_balance_median():
    """tests/test_simulation_correctness.py:737-778 of the reference: the histogram cohort is the
    backend's successful-path cohort (a successful path may finish at exactly 0)."""
    import pandas as pd

    cfg = Config(**dict(__import__("scenarios").TEST_BASE, num_simulations_main=3, retirement_years=1,
                        other_income_streams=[]))
    summary = pd.DataFrame({
        "Start Balance": [100.0, 100.0, 100.0], "Final Balance": [0.0, 50.0, 25.0], "Success": [True, True, False],
        "YearsToRuin": [float("nan"), float("nan"), 0.5], "First Year Gross Withdrawal": [0.0, 10.0, 10.0],
        "First Year Real Gross Withdrawal": [0.0, 10.0, 10.0], "Inflation At Retirement": [1.0, 1.0, 1.0]})

    class FakeSimulator:
        def run_monte_carlo_simulations(self, **_kwargs):
            return summary, None, None, None, None, None, None

    result = payload.build_result(cfg, FakeSimulator(), required_w_months=0, search_curve=[])
    _validate_with_reference_model(result)
    assert result["summary"]["success_probability"] == pytest.approx(66.67)
    assert result["summary"]["median_final_balance_successful"] == pytest.approx(25.0)
    assert result["histogram"]["final_balances"] == [0.0, 50.0, 25.0]
    assert result["histogram"]["success_flags"] == [True, True, False]
    assert result["ruin_histogram"]["failure_count"] == 1
    assert result["ruin_histogram"]["years_to_ruin"] == [0.5]
    assert result["trajectory"] is None and result["withdrawal_rate"] is None and result["search_curve"] is None


def test_payload_preserves_exact_fractional_timeline():
    """tests/test_simulation_correctness.py:781-817: a 13-month working period stays 13/12 years."""
    import scenarios

    d = dict(scenarios.TEST_BASE, num_simulations_main=2, retirement_years=1, monthly_expenses=0.0, seed=5)
    cfg = Config(**d)
    result = payload.build_result(cfg, OracleBackedSimulator(d), required_w_months=13,
                                  search_curve=[{"working_months": 13, "working_years": 1.1, "probability": 100.0}])
    _validate_with_reference_model(result)
    ret = 13 / 12
    assert result["trajectory"]["years"] == pytest.approx([0.0, 1.0, ret, ret + 1])
    assert result["withdrawal_rate"]["years"][0] == pytest.approx(ret)
    assert result["reference_lines"][0]["year"] == pytest.approx(ret)
    assert result["summary"]["working_period_is_estimate"] is True


def _validate_with_reference_model(result) -> None:
    """`SimulationResponse.model_validate` with the reference's own pydantic model — only where the
    reference tree is mounted (the build container); the golden payloads were validated with it
    when they were generated (tests/golden/make_payload_golden.py)."""
    import sys

    backend = "/root/reference/backend"
    if not os.path.isdir(backend):
        return
    saved_path, saved_mods = list(sys.path), {k: sys.modules.get(k) for k in ("server", "config", "constants", "simulation", "utils")}
    sys.path.insert(0, backend)
    try:
        for k in saved_mods:
            sys.modules.pop(k, None)
        import server  # the reference's FastAPI module (no app is started)

        server.SimulationResponse.model_validate(result)
    finally:
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


def test_aggregate_shaped_payload_validates_against_the_reference_response_model():
    """The aggregate-only payload keeps every field the reference model requires (the O(N) lists
    stay present and empty; `binned` / `bins` are extra keys)."""
    g = GOLDEN["stressed"]
    legacy = json.loads(json.dumps(g["result"]))
    legacy["histogram"] = {"final_balances": [], "start_balances": [], "success_flags": [],
                           "binned": payload.balance_bins(0.0, 1.0, [0] * 60, 0.0, 0, 10)}
    legacy["ruin_histogram"] = {"years_to_ruin": [], "failure_count": 3, "total_paths": 10,
                                "bins": payload.ruin_year_bins([0, 1, 0, 2])}
    _validate_with_reference_model(legacy)
    assert legacy["ruin_histogram"]["bins"] == [{"year": 1, "label": "1", "count": 3}]


# ---- aggregate-only assembly on the CPU: a stub that answers run_aggregates() from the oracle -------
class OracleAggregates(OracleBackedSimulator):
    """`run_aggregates` (the dict DeviceAggregates.to_host returns) reduced on the host from the
    oracle's 7-tuple, so that payload's aggregate mode can be checked without a GPU."""

    def run_aggregates(self, working_months, num_simulations, *, bands=True, samples=False, first_path=0):
        summary, traj, smp, wr, real, real_smp, wr_counts = self.sim.run(working_months, num_simulations)
        agg = host_aggregates(summary)
        ok = summary["Success"].astype(bool)
        succ = summary.loc[ok, "Final Balance"].to_numpy()
        if len(succ) and succ.max() > succ.min():
            lo, hi = float(succ.min()), float(succ.max())
            width = (hi - lo) / 60
            counts = np.bincount(np.minimum(np.floor((succ - lo) / width), 59).astype(int), minlength=60).tolist()
        elif len(succ):
            lo = hi = float(succ.min())
            counts = [len(succ)] + [0] * 59
        else:
            lo = hi = float("nan")
            counts = [0] * 60
        months = np.rint(summary.loc[~ok, "YearsToRuin"].dropna().to_numpy() * 12).astype(int)
        R = len(wr)
        agg.update({
            "final_balance_hist_60": {"range": [lo, hi], "counts": counts},
            "ruin_month_hist": np.bincount(months, minlength=12 * R + 1).tolist(),
            "trajectory_bands": traj, "real_trajectory_bands": real, "withdrawal_rate_bands": wr,
            "withdrawal_rate_counts": wr_counts, "sample_paths": smp, "real_sample_paths": real_smp,
        })
        return agg


@pytest.mark.parametrize("name", ["config_json", "jorge_plus", "stressed", "broke"])
def test_aggregate_mode_assembles_the_legacy_payload_without_the_lists(name):
    g = GOLDEN[name]
    cfg = Config(**g["cfg"])
    wm, curve = g["working_months"], g["search_curve"]
    legacy = payload.build_result(cfg, OracleAggregates(g["cfg"]), wm, search_curve=curve, mode="legacy")
    agg = payload.build_result(cfg, OracleAggregates(g["cfg"]), wm, search_curve=curve, mode="aggregate")
    assert agg.keys() == legacy.keys()
    for key in ("scenario", "summary", "trajectory", "trajectory_real", "withdrawal_rate", "search_curve",
                "reference_lines"):
        assert json.dumps(agg[key], sort_keys=True) == json.dumps(legacy[key], sort_keys=True), key
    assert agg["histogram"]["final_balances"] == [] and agg["ruin_histogram"]["years_to_ruin"] == []
    assert agg["ruin_histogram"]["failure_count"] == legacy["ruin_histogram"]["failure_count"]
    assert agg["ruin_histogram"]["bins"] == js_bin_ruin_years(legacy["ruin_histogram"]["years_to_ruin"])
    summary = OracleAggregates(g["cfg"]).run_monte_carlo_simulations(wm, cfg.num_simulations_main)[0]
    want = js_bin_data(summary["Final Balance"].tolist(), summary["Success"].tolist())
    got = agg["histogram"]["binned"]
    assert got["successRate"] == want["successRate"] and got["median"] == want["median"]
    assert [b["count"] for b in got["bins"]] == [b["count"] for b in want["bins"]]
    assert [b["label"] for b in got["bins"]] == [b["label"] for b in want["bins"]]
    _validate_with_reference_model(json.loads(json.dumps(agg, allow_nan=False)))
    # "auto" picks the aggregate path above the threshold when the simulator offers it
    auto = payload.build_result(cfg, OracleAggregates(g["cfg"]), wm, search_curve=curve, aggregate_threshold=1)
    assert auto["histogram"]["final_balances"] == []
