"""CPU-only tests: the C-ABI library loads and exports every symbol include/mcr.h declares,
host-side helpers equal the reference's known answers, the search driver makes the reference's
decisions on fake engines (the two monkey-patch tests of the reference), config validation.
No compute call is made (there is no GPU here and the engine has no CPU fallback)."""
from __future__ import annotations

import ctypes as C
import math
import os
import re

import numpy as np
import pandas as pd
import pytest

import golden_io
import scenarios
from monte_carlo_retirement_b200 import native
from monte_carlo_retirement_b200.config import Config, ConfigurationError, load_config_from_json
from monte_carlo_retirement_b200.simulation import (RetirementMonteCarloSimulator, age_at_retirement_year,
                                                    arithmetic_to_log_params, median_first_year_withdrawal_rate,
                                                    params_from_model, retirement_age, stream_payment_start_age,
                                                    stream_payment_start_month_index, trajectory_time_points,
                                                    years_from_t0_to_age)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg(**over):
    d = dict(scenarios.TEST_BASE)
    d.update(over)
    return Config(**d)


def test_library_exports_every_symbol_the_header_declares():
    header = open(os.path.join(ROOT, "include", "mcr.h")).read()
    declared = set(re.findall(r"\b(mcr_[a-z0-9_]+)\s*\(", header))
    declared -= {"mcr_ctx"}
    lib = native.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/mcr.h but not exported"
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)
    assert lib.mcr_abi_version() == 1


def test_pod_layouts_match_the_header():
    # sizes implied by include/mcr.h (explicitly padded PODs)
    assert C.sizeof(native.IncomeStream) == 32
    assert C.sizeof(native.Params) == 17 * 8 + 4 * 4 + 16 * 32
    assert C.sizeof(native.Outputs) == 15 * 8
    assert C.sizeof(native.PathRecord) == 5 * 8 + 4 * 4


def test_python_constants_match_the_header_macros():
    header = open(os.path.join(ROOT, "include", "mcr.h")).read()
    macros = {m.group(1): int(m.group(2).rstrip("uU"), 0)
              for m in re.finditer(r"#define\s+(MCR_[A-Z0-9_]+)\s+(0x[0-9a-fA-F]+[uU]?|\d+[uU]?)\b", header)}
    pairs = {"SEL_MEDIAN": "MCR_SEL_MEDIAN", "SEL_MINMAX": "MCR_SEL_MINMAX", "HIST_NUMPY": "MCR_HIST_NUMPY",
             "HIST_FLOOR": "MCR_HIST_FLOOR", "HIST_RAW_RANGE": "MCR_HIST_RAW_RANGE"}
    for py, c in pairs.items():
        assert c in macros, c
        assert getattr(native, py) == macros[c], (py, c)


def test_host_side_abi_helpers_match_reference_known_answers():
    lib = native.load_library()
    z = golden_io.load_helpers()
    for age, wm, start, exp in z["stream_start"]:
        assert lib.mcr_stream_start_month(age, int(wm), start) == int(exp)
        assert stream_payment_start_month_index(age, int(wm), start) == int(exp)
    for wm in (0, 1, 11, 12, 13, 24, 233, 240, 599):
        for R in (1, 40, 50):
            assert lib.mcr_trajectory_len(wm, R) == len(trajectory_time_points(wm, R))
    assert stream_payment_start_month_index(60.0, 0, 60.51) == 7          # tests/...:358-361
    assert retirement_age(40.0, 240) == pytest.approx(60.0)
    assert stream_payment_start_age(40.0, 240, 55.0) == pytest.approx(60.0)
    assert age_at_retirement_year(40.0, 240, 5) == pytest.approx(65.0)
    assert years_from_t0_to_age(40.0, 30.0) == 0.0
    assert trajectory_time_points(13, 1) == pytest.approx([0.0, 1.0, 13 / 12, 25 / 12])


def test_no_gpu_means_loud_failure_not_a_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    sim = RetirementMonteCarloSimulator(_cfg())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sim.run_monte_carlo_simulations(0, 4)
    with pytest.raises(RuntimeError):
        sim._run_single_simulation_path(0, 1)
    # creating a context without a device fails inside the library too
    h = C.c_void_p()
    rc = native.load_library().mcr_create(C.byref(params_from_model(_cfg())), 1, 0, C.byref(h))
    assert rc == -2 and b"no CPU fallback" in native.load_library().mcr_last_error(None)


def test_constructor_contract():
    for bad in (dict(inv1_returns_mean=-1.0), dict(inflation_rate_mean=-1.0),
                dict(inv2_premium_over_inflation_mean=-1.0), dict(num_simulations_search=0), dict(seed=-1)):
        with pytest.raises(ValueError):
            _cfg(**bad)
    with pytest.raises(ValueError):
        RetirementMonteCarloSimulator(_cfg(seed=0), main_seed_override=-1)
    sim = RetirementMonteCarloSimulator(_cfg(seed=7), main_seed_override=99)
    assert sim.main_seed == 99 and sim.params_model.seed == 7
    cfg = _cfg()
    sim = RetirementMonteCarloSimulator(cfg)
    cfg.monthly_expenses = 1.0  # the simulator works on a deep copy
    assert sim.params_model.monthly_expenses == 2000.0
    with pytest.raises(ValueError, match="at most 16"):
        RetirementMonteCarloSimulator(_cfg(other_income_streams=[
            dict(name=f"s{i}", monthly_amount_today=1.0, start_at_age=60, duration_years=None,
                 inflation_indexed=True, tax_rate=0.0) for i in range(17)]))
    assert _cfg().allocation_inv2_pct == pytest.approx(0.4)
    assert Config(**scenarios.CONFIG_JSON).Nickname == "Macunaima ret plan"
    with pytest.raises(ConfigurationError):
        load_config_from_json("/nonexistent.json")


def test_series_plan_is_decided_once_per_call_shape(monkeypatch):
    """cudaMemGetInfo takes tens of milliseconds while kernels run; a step must not ask for it again
    (it starved the launch queue: one bench run in ten was host-bound). Env overrides and key=None re-plan."""
    import torch

    calls = []

    def fake_mem_get_info(device=None):
        calls.append(device)
        return (int(8 * 1000 * 61 * 2.5 / 0.7), 0)   # room for two trajectory series and a bit

    monkeypatch.setattr(torch.cuda, "mem_get_info", fake_mem_get_info)
    monkeypatch.delenv("MCR_SERIES_BUDGET_BYTES", raising=False)
    monkeypatch.delenv("MCR_SERIES_SWEEP", raising=False)
    sim = RetirementMonteCarloSimulator(_cfg())
    monkeypatch.setattr(sim, "_torch_device", lambda: "cuda:0")   # (no device here: only the planning logic runs)
    key = (120, 1000, True)
    plan = sim._series_plan(1000, 61, 40, True, key)
    assert plan == [("traj", "real"), ("wr",)] and len(calls) == 1
    for _ in range(5):
        assert sim._series_plan(1000, 61, 40, True, key) is plan
    assert len(calls) == 1
    assert sim._series_plan(1000, 61, 40, False, key) == [()] and len(calls) == 1      # no bands: nothing to plan
    monkeypatch.setenv("MCR_SERIES_SWEEP", "1")                                       # an override is part of the key
    assert sim._series_plan(1000, 61, 40, True, key) == [("traj",), ("real",), ("wr",)] and len(calls) == 2
    monkeypatch.delenv("MCR_SERIES_SWEEP")
    assert sim._series_plan(1000, 61, 40, True, key) is plan and len(calls) == 2
    sim._series_plan(1000, 61, 40, True, None)                                        # ad-hoc query: not cached
    assert len(calls) == 3
    assert sim._series_plan(2000, 61, 40, True, (120, 2000, True)) == [("traj",), ("real",), ("wr",)] and len(calls) == 4


def test_seed_streams_and_numpy_draws_match_reference():
    """`_path_seeds` / `_draw_shock_path` keep the reference's numpy semantics (golden seeds)."""
    for name in ("config_json", "tax_heavy"):
        cfg, seed, cases = golden_io.load_paths(name)
        sim = RetirementMonteCarloSimulator(Config(**cfg))
        assert sim.main_seed == seed
        for case in cases:
            (sim.use_search_seeds if case["stream"] == "search" else sim.use_final_seeds)()
            seeds = sim._path_seeds(int(case["n"]))
            assert seeds == [int(s) for s in case["seeds"]]
            n_rows = max(int(case["wm"]) + cfg["retirement_years"] * 12, 1)
            sh = sim._draw_shock_path(n_rows, seeds[0])
            assert np.array_equal(np.vstack([sh[:3], sh[-1:]]), case["shock_probe"])
    pos = RetirementMonteCarloSimulator(_cfg(equity_inflation_correlation=1.0))._draw_shock_path(100, 4)
    neg = RetirementMonteCarloSimulator(_cfg(equity_inflation_correlation=-1.0))._draw_shock_path(100, 4)
    assert pos[:, 1] == pytest.approx(pos[:, 0]) and neg[:, 1] == pytest.approx(-neg[:, 0])


def test_log_params_and_medians():
    mu, sg = arithmetic_to_log_params(0.12, 0.15)
    assert math.exp(mu + 0.5 * sg * sg) == pytest.approx(1.12)
    assert arithmetic_to_log_params(0.05, 0.0) == (math.log(1.05), 0.0)
    with pytest.raises(ValueError):
        arithmetic_to_log_params(-1.0, 0.1)
    with pytest.raises(ValueError):
        arithmetic_to_log_params(0.1, -0.1)
    df = pd.DataFrame({"Start Balance": [100.0, 0.0, 200.0], "First Year Real Gross Withdrawal": [5.0, 1.0, 8.0]})
    assert median_first_year_withdrawal_rate(df) == pytest.approx(4.5)
    assert math.isnan(median_first_year_withdrawal_rate(pd.DataFrame()))
    sim = RetirementMonteCarloSimulator(_cfg())
    assert sim._success_probability(pd.DataFrame()) == 0.0
    assert sim._success_probability(pd.DataFrame({"Success": [True, False, True, True]})) == 75.0
    assert sim._success_probability(pd.DataFrame({"Final Balance": [1.0, 0.0]})) == 50.0


def _fake(df_for):
    def run(working_months: int, num_simulations: int):
        return df_for(working_months, num_simulations), None, None, None, None, None, None
    return run


def test_bisection_finds_true_minimum_through_patched_engine():  # tests/...:259-293
    threshold = 37
    sim = RetirementMonteCarloSimulator(_cfg(target_probability=90.0, num_simulations_search=10, seed=0))

    def df_for(wm, n):
        ok = wm >= threshold
        return pd.DataFrame({"Start Balance": [100.0] * n, "Final Balance": [1.0 if ok else 0.0] * n,
                             "Success": [ok] * n, "First Year Gross Withdrawal": [1.0] * n,
                             "Inflation At Retirement": [1.0] * n})

    sim.run_monte_carlo_simulations = _fake(df_for)
    months, prob, curve = sim.find_minimum_working_months(verbose=False)
    assert months == threshold and prob >= 90.0
    assert all("working_months" in p and "probability" in p for p in curve)
    assert sim.last_search_stats["policy"] == "sequential"  # probes went through the patched attribute


def test_search_verification_handles_non_monotone_probabilities():  # tests/...:296-332
    sim = RetirementMonteCarloSimulator(_cfg(target_probability=50.0, num_simulations_search=400, seed=0))

    def df_for(wm, n):
        k = 201 if wm == 4 else (213 if wm >= 24 else 199)
        flags = [True] * k + [False] * (n - k)
        return pd.DataFrame({"Start Balance": [100.0] * n, "Final Balance": [1.0 if f else 0.0 for f in flags],
                             "Success": flags})

    sim.run_monte_carlo_simulations = _fake(df_for)
    months, probability, _ = sim.find_minimum_working_months(verbose=False)
    assert months == 4 and probability == pytest.approx(50.25)


@pytest.mark.parametrize("name", ["config_json", "jorge_json", "stressed", "tax_heavy", "unreachable"])
def test_search_driver_replays_reference_decisions_and_events(name):
    """Feed the reference's own per-probe probabilities (golden search curve is rounded, so
    rebuild the exact table with the pinned oracle) through the drop-in's search driver: same
    months, probability, curve and progress events as the reference."""
    from oracle import oracle as orc

    g = golden_io.load_search()[name]
    o = orc.OracleSimulator(g["cfg"], n_threads=4)
    o.use_search_seeds()
    n = g["cfg"]["num_simulations_search"]
    table = {}

    def df_for(wm, num):
        assert num == n
        if wm not in table:
            recs, _, _, _ = orc.run_batch(o.p, wm, orc.shocks_for_seeds(o.p, wm, o.seeds.path_seeds(n)), 4,
                                          want_series=False)
            table[wm] = recs["success"].astype(bool)
        return pd.DataFrame({"Success": table[wm], "Final Balance": np.zeros(n)})

    sim = RetirementMonteCarloSimulator(Config(**g["cfg"]))
    sim.run_monte_carlo_simulations = _fake(df_for)
    events = []
    months, prob, curve = sim.find_minimum_working_months(verbose=False, progress_callback=events.append)
    assert months == g["months"] and prob == g["prob"] and curve == g["curve"]
    assert events == g["events"]


def test_header_is_a_plain_c_header(tmp_path):
    """include/mcr.h is the drop-in boundary: it must be consumable from C (cgo / ctypes / FFI
    generators), not only from the C++ that implements it."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include "mcr.h"\nint main(void) { mcr_params p; mcr_outputs o; (void)p; (void)o; '
                   'return (int)sizeof(mcr_path_record) == 0; }\n')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                        "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
