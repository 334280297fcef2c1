"""Response payload on the GPU engine.

(1) legacy mode through the drop-in simulator replaying the reference's numpy draws against the
    dicts the reference's own `server._build_result` produced (tests/golden/payload.json):
    identical structure, flags, counts and labels; money values to the cent (a value within
    1e-9 relative of a rounding boundary may land on the neighbouring cent);
(2) aggregate-only mode against legacy mode on the same native-RNG batch: every shared field
    identical, the chart bins equal to the dashboard's own binning of the legacy lists.
"""
from __future__ import annotations

import json
import os

import numpy as np
import pytest

import scenarios
from gpu_util import make_sim
from test_payload import js_bin_data, js_bin_ruin_years

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "payload.json")))


def _same_money(got, want, path=""):
    """Recursive comparison: numbers to the cent (+1e-9 relative), everything else exact."""
    if isinstance(want, dict):
        assert isinstance(got, dict) and got.keys() == want.keys(), path
        for k in want:
            _same_money(got[k], want[k], f"{path}.{k}")
    elif isinstance(want, list):
        assert isinstance(got, list) and len(got) == len(want), path
        for i, (g, w) in enumerate(zip(got, want)):
            _same_money(g, w, f"{path}[{i}]")
    elif isinstance(want, float) and not isinstance(want, bool):
        assert isinstance(got, (int, float)), path
        assert abs(got - want) <= 0.0100001 + 1e-9 * abs(want), (path, got, want)
    else:
        assert got == want, (path, got, want)


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_legacy_payload_on_reference_draws_matches_reference(name):
    from monte_carlo_retirement_b200 import payload
    from monte_carlo_retirement_b200.config import Config

    g = GOLDEN[name]
    cfg = Config(**g["cfg"])
    sim = make_sim(cfg, rng="numpy")
    sim.use_final_seeds()
    got = payload.build_result(cfg, sim, g["working_months"], search_curve=g["search_curve"], mode="legacy")
    got = json.loads(json.dumps(got, allow_nan=False))
    want = g["result"]
    _same_money(got, want)
    assert got["histogram"]["success_flags"] == want["histogram"]["success_flags"]
    assert got["summary"]["success_probability"] == want["summary"]["success_probability"]


@pytest.mark.parametrize("cfg_dict,wm,n", [
    (scenarios.CONFIG_JSON, 233, 6000),
    (scenarios.JORGE_PLUS, 75, 4001),
    (scenarios.STRESSED, 150, 5000),
    (scenarios.CORNER_BROKE, 0, 300),
], ids=["config_json", "jorge_plus", "stressed", "broke"])
def test_aggregate_payload_equals_legacy_payload_without_the_lists(cfg_dict, wm, n):
    from monte_carlo_retirement_b200 import payload
    from monte_carlo_retirement_b200.config import Config

    cfg = Config(**dict(cfg_dict, num_simulations_main=n))
    curve = [{"working_months": wm, "working_years": wm / 12, "probability": 50.0}]
    sim = make_sim(cfg)
    sim.use_final_seeds()
    legacy = payload.build_result(cfg, sim, wm, search_curve=curve, mode="legacy")
    agg = payload.build_result(cfg, sim, wm, search_curve=curve, mode="aggregate")
    assert agg.keys() == legacy.keys()
    for key in ("scenario", "summary", "trajectory", "trajectory_real", "withdrawal_rate", "search_curve",
                "reference_lines"):
        assert json.dumps(agg[key], sort_keys=True) == json.dumps(legacy[key], sort_keys=True), key
    # the lists are gone, the counts stay
    assert agg["histogram"]["final_balances"] == [] and agg["histogram"]["success_flags"] == []
    assert agg["ruin_histogram"]["years_to_ruin"] == []
    assert agg["ruin_histogram"]["failure_count"] == legacy["ruin_histogram"]["failure_count"]
    assert agg["ruin_histogram"]["total_paths"] == n
    # and the charts get exactly what they would have computed from the lists
    assert agg["ruin_histogram"]["bins"] == js_bin_ruin_years(legacy["ruin_histogram"]["years_to_ruin"])
    summary = sim.run_monte_carlo_simulations(wm, n)[0]
    want = js_bin_data(summary["Final Balance"].tolist(), summary["Success"].tolist())
    got = agg["histogram"]["binned"]
    assert got["successRate"] == want["successRate"]
    assert [b["count"] for b in got["bins"]] == [b["count"] for b in want["bins"]]
    assert [b["label"] for b in got["bins"]] == [b["label"] for b in want["bins"]]
    np.testing.assert_allclose([b["mid"] for b in got["bins"]], [b["mid"] for b in want["bins"]], rtol=1e-14)
    assert got["median"] == want["median"]
    json.dumps(agg, allow_nan=False)  # serialisable as the SSE / JSON response


def test_auto_mode_switches_to_aggregates_above_the_threshold():
    from monte_carlo_retirement_b200 import payload
    from monte_carlo_retirement_b200.config import Config

    cfg = Config(**dict(scenarios.JORGE_JSON, num_simulations_main=3000))
    sim = make_sim(cfg)
    sim.use_final_seeds()
    small = payload.build_result(cfg, sim, 75)
    assert len(small["histogram"]["final_balances"]) == 3000 and "binned" not in small["histogram"]
    big = payload.build_result(cfg, sim, 75, aggregate_threshold=1000)
    assert big["histogram"]["final_balances"] == [] and len(big["histogram"]["binned"]["bins"]) == 60
    assert big["summary"] == small["summary"]


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_cli_report_from_device_aggregates_matches_reference_log(name):
    """report.py on `run_aggregates()` of the reference's own draws: the CLI's result log
    (utils.py:69-102) line for line, and the 100-bin histogram of plotting.py:46-59."""
    from monte_carlo_retirement_b200 import report
    from monte_carlo_retirement_b200.config import Config

    g = GOLDEN[name]
    cfg = Config(**g["cfg"])
    sim = make_sim(cfg, rng="numpy")
    sim.use_final_seeds()
    agg = sim.run_aggregates(g["working_months"], cfg.num_simulations_main)
    assert report.result_log_lines(cfg, g["working_months"], agg) == g["cli"]["log"]
    counts, edges = report.final_balance_histogram(agg)
    want = g["cli"]["hist100"]
    if want is None:
        assert counts.sum() == 0
    else:
        assert counts.sum() == sum(want["counts"])
        # a balance within 1e-9 of a bin edge may land in the neighbouring bin
        assert np.abs(np.cumsum(counts) - np.cumsum(want["counts"])).max() <= 1
        np.testing.assert_allclose(edges, want["edges"], rtol=1e-9)
