"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle and the
golden vectors of the unmodified reference, on the same inputs.

Bars (BASELINE.json north_star): replay of the reference's numpy draws -> bit-exact success
flags / ruin months / NaN patterns, balances within 1e-9 relative (fp64). Helper arithmetic has
no transcendental, so it is required to be bit-exact.
"""
from __future__ import annotations

import numpy as np
import pytest

import golden_io
import scenarios
from gpu_util import REL, assert_close, device_batch_to_host, make_sim

pytestmark = pytest.mark.gpu

CASES = list(golden_io.iter_cases())
IDS = [f"{n}-{c['stream']}-wm{c['wm']}" for n, _, _, c in CASES]


def _shocks_device(sim, wm, seeds):
    import torch

    R = sim.params_model.retirement_years
    n_rows = max(wm + 12 * R, 1)
    host = np.empty((n_rows, 3, len(seeds)))
    for i, s in enumerate(seeds):
        host[:, :, i] = sim._draw_shock_path(n_rows, int(s))
    return torch.from_numpy(host).to("cuda"), host


@pytest.mark.parametrize("name,cfg,seed,case", CASES, ids=IDS)
def test_replay_matches_reference_golden(name, cfg, seed, case):
    """mcr_replay on the reference's own numpy draws vs the reference's recorded outputs."""
    sim = make_sim(cfg)
    wm, n = int(case["wm"]), int(case["n"])
    shocks, host_shocks = _shocks_device(sim, wm, case["seeds"])
    # guard: the numpy bit-stream here is the one the fixtures were made with
    assert np.array_equal(np.vstack([host_shocks[:3, :, 0], host_shocks[-1:, :, 0]]), case["shock_probe"])
    b = sim.run_batch_device(wm, n, shocks=shocks)
    h = device_batch_to_host(b)
    assert np.array_equal(h["success"], case["success"])                     # bit-exact flags
    assert np.array_equal(np.isnan(h["ruin_years"]), np.isnan(case["ruin"]))
    assert np.array_equal(np.nan_to_num(h["ruin_years"], nan=-1.0), np.nan_to_num(case["ruin"], nan=-1.0))
    assert_close(h["start"], case["start"])
    assert_close(h["final"], case["final"])
    assert_close(h["fy_gross"], case["fy_gross"])
    assert_close(h["fy_real"], case["fy_real"])
    assert_close(h["infl"], case["infl"])
    assert_close(h["traj"], case["traj"])
    assert_close(h["real"], case["real"])
    assert np.array_equal(np.isnan(h["wr"]), np.isnan(case["wr"]))
    assert_close(np.nan_to_num(h["wr"]), np.nan_to_num(case["wr"]))
    assert int(h["success_count"]) == int(case["success"].sum())


@pytest.mark.parametrize("name,cfg,seed,case", CASES[::3], ids=IDS[::3])
def test_replay_matches_oracle(name, cfg, seed, case):
    """Same inputs through the C oracle (which is pinned bit-exact to the reference)."""
    from oracle import oracle as orc

    sim = make_sim(cfg)
    wm, n = int(case["wm"]), int(case["n"])
    shocks, host = _shocks_device(sim, wm, case["seeds"])
    p = orc.params_from_config(cfg)
    recs, traj, real, wr = orc.run_batch(p, wm, np.ascontiguousarray(host.transpose(2, 0, 1)))
    h = device_batch_to_host(sim.run_batch_device(wm, n, shocks=shocks))
    assert np.array_equal(h["success"], recs["success"].astype(bool))
    assert_close(h["final"], recs["final_balance"])
    assert_close(h["traj"], traj)
    assert_close(h["real"], real)
    assert np.array_equal(np.isnan(h["wr"]), np.isnan(wr))


@pytest.mark.parametrize("name", ["config_json", "jorge_plus", "tax_heavy", "annual_both", "broke"])
def test_single_path_matches_golden(name):
    """`_run_single_simulation_path(wm, path_seed)` — the call 11 reference tests make."""
    cfg, seed, cases = golden_io.load_paths(name)
    sim = make_sim(cfg)
    for case in cases:
        for i in range(min(3, int(case["n"]))):
            r = sim._run_single_simulation_path(int(case["wm"]), int(case["seeds"][i]))
            assert r["Success"] is bool(case["success"][i])
            assert_close(r["Start Balance"], case["start"][i])
            assert_close(r["Final Balance"], case["final"][i])
            assert_close(r["First Year Gross Withdrawal"], case["fy_gross"][i])
            assert_close(r["First Year Real Gross Withdrawal"], case["fy_real"][i])
            assert_close(r["Inflation At Retirement"], case["infl"][i])
            assert_close(np.array(r["Trajectory"]), case["traj"][i])
            assert_close(np.array(r["RealTrajectory"]), case["real"][i])
            wr = np.array(r["WithdrawalRateTrajectory"])
            assert np.array_equal(np.isnan(wr), np.isnan(case["wr"][i]))
            assert_close(np.nan_to_num(wr), np.nan_to_num(case["wr"][i]))
            g = case["ruin"][i]
            assert (np.isnan(g) and np.isnan(r["YearsToRuin"])) or r["YearsToRuin"] == g


def test_helpers_bit_exact_vs_reference():
    """_calculate_withdrawal_and_update / _net_liquidation_value / _rebalance_portfolio: plain
    IEEE arithmetic, so the strict CUDA thread must reproduce the reference bit for bit."""
    z = golden_io.load_helpers()
    sim = make_sim(scenarios.CONFIG_JSON)
    for row, exp_wd, exp_nl in zip(z["wd_in"][::2], z["wd_out"][::2], z["nl_out"][::2]):
        got = sim._calculate_withdrawal_and_update(row[0], row[1], row[2], bool(row[3]), row[4])
        assert np.array_equal(np.array(got), exp_wd), (row, got, exp_wd)
        assert sim._net_liquidation_value(row[0], row[1], bool(row[3]), row[4]) == exp_nl
    cfgs = {"tax_heavy": scenarios.TAX_HEAVY, "config_json": scenarios.CONFIG_JSON,
            "annual_both": scenarios.ANNUAL_BOTH, "test_base": scenarios.TEST_BASE}
    for name in z["sim_names"]:
        s = make_sim(cfgs[str(name)])
        for row, exp in zip(z["rb_in"][::3], z[f"rb_{name}"][::3]):
            got = s._rebalance_portfolio(row[0], row[1], row[2], row[3])
            assert np.array_equal(np.array(got), exp), (name, row, got, exp)


def test_annual_tax_helper_bit_exact_vs_reference():
    """_apply_annual_gain_taxes (simulation.py:361-450) as one strict CUDA thread: balances, cost
    bases and the tax_failed flag of the reference on random inputs, for scenarios with annual
    tax on one asset, on both, and on none (where only the trailing rebalance acts)."""
    z = golden_io.load_helpers()
    cfgs = {"tax_heavy": scenarios.TAX_HEAVY, "config_json": scenarios.CONFIG_JSON,
            "annual_both": scenarios.ANNUAL_BOTH, "test_base": scenarios.TEST_BASE}
    failed_seen = 0
    for name in z["sim_names"]:
        s = make_sim(cfgs[str(name)])
        for row, exp in zip(z["rb_in"][::2], z[f"at_{name}"][::2]):
            got = s._apply_annual_gain_taxes(*[float(v) for v in row[:6]])
            assert isinstance(got[4], bool)
            assert np.array_equal(np.array([*got[:4], float(got[4])]), exp), (name, row, got, exp)
            failed_seen += int(got[4])
    assert failed_seen > 0  # the fixture exercises the tax_failed branch


@pytest.mark.parametrize("name,cfg,seed,case", CASES[1::4], ids=IDS[1::4])
def test_dropin_7tuple_numpy_rng_matches_reference(name, cfg, seed, case):
    """run_monte_carlo_simulations with rng='numpy' against the reference's own 7-tuple: device
    radix-select bands, sample paths, WR bands and observation counts."""
    sim = make_sim(cfg, rng="numpy")
    (sim.use_search_seeds if case["stream"] == "search" else sim.use_final_seeds)()
    # same seeds as the fixture (the reference spawned them in fixture order on one simulator)
    sim._path_seed_cache[(sim._stream_name, int(case["n"]))] = [int(s) for s in case["seeds"]]
    summary, traj_pct, samples, wr_pct, real_pct, real_samples, wr_counts = sim.run_monte_carlo_simulations(
        int(case["wm"]), int(case["n"]))
    assert list(summary.columns) == ["Start Balance", "Final Balance", "Success", "YearsToRuin",
                                     "First Year Gross Withdrawal", "First Year Real Gross Withdrawal",
                                     "Inflation At Retirement"]
    assert np.array_equal(summary["Success"].to_numpy(), case["success"])
    assert summary["Success"].dtype == bool
    assert_close(summary["Start Balance"].to_numpy(), case["start"])
    assert_close(summary["Final Balance"].to_numpy(), case["final"])
    assert list(traj_pct.columns) == list(case["pct_cols"]) and list(wr_pct.columns) == list(case["wr_cols"])
    assert_close(traj_pct.to_numpy(), case["traj_pct"])
    assert_close(real_pct.to_numpy(), case["real_pct"])
    assert np.array_equal(np.isnan(wr_pct.to_numpy()), np.isnan(case["wr_pct"]))
    assert_close(np.nan_to_num(wr_pct.to_numpy()), np.nan_to_num(case["wr_pct"]))
    assert wr_counts == [int(v) for v in case["wr_counts"]]
    assert_close(np.array(samples), case["samples"])
    assert_close(np.array(real_samples), case["real_samples"])
    assert traj_pct[0.5].shape[0] == len(case["traj_pct"])  # float column labels work like the reference's


def test_search_numpy_rng_reproduces_reference_search():
    """find_minimum_working_months on the reference's draws: same months, probability,
    search_curve and probe order as the reference (config.json and a stressed scenario)."""
    g = golden_io.load_search()
    for name in ("jorge_json", "stressed", "unreachable"):
        sim = make_sim(g[name]["cfg"], rng="numpy")
        events = []
        months, prob, curve = sim.find_minimum_working_months(verbose=False, progress_callback=events.append)
        assert months == g[name]["months"]
        assert prob == g[name]["prob"]
        assert curve == g[name]["curve"]
        assert events == g[name]["events"]


def test_fast_and_strict_builds_agree_on_replay_inputs():
    """The throughput build (FMA contraction, shared reciprocals, short-range exp) must stay
    within 1e-9 relative of the parity build and produce the same success flags."""
    import torch

    for cfg, wm in ((scenarios.SYNTH_C3, 240), (scenarios.TAX_HEAVY, 150), (scenarios.JORGE_PLUS, 75),
                    (scenarios.STRESSED, 60), (scenarios.ANNUAL_BOTH, 100)):
        n = 4096
        strict = make_sim(cfg, strict=True)
        fast = make_sim(cfg, strict=False)
        # identical draws for both builds: native strict draws, replayed by the strict kernel,
        # against the fast kernel fed through its own Philox (same counters, MUFU normals differ
        # by ~1e-6), so compare fast vs strict on the SAME shocks via a fast replay of them.
        R = cfg["retirement_years"]
        n_rows = wm + 12 * R
        shocks = torch.empty((n_rows, 3, n), dtype=torch.float64, device="cuda")
        strict.native_context.draw_shocks(1, 0, n, n_rows, shocks, n, strict=True)
        hs = device_batch_to_host(strict.run_batch_device(wm, n, shocks=shocks))
        hf = device_batch_to_host(fast.run_batch_device(wm, n, shocks=shocks, _fast_replay=True))
        assert np.array_equal(hs["success"], hf["success"])
        assert np.array_equal(hs["ruin_month"], hf["ruin_month"])
        assert_close(hs["final"], hf["final"])
        assert_close(hs["traj"], hf["traj"])
        assert_close(hs["fy_real"], hf["fy_real"])


LOWVOL_CASES = [
    ("synth_c3", scenarios.SYNTH_C3, 240),
    ("config_json_all_fail", scenarios.CONFIG_JSON, 120),
    ("config_json", scenarios.CONFIG_JSON, 233),
    ("jorge_plus_lowvol", dict(scenarios.JORGE_PLUS, inv1_returns_volatility=0.05), 75),
    ("jorge_plus_lowvol_late", dict(scenarios.JORGE_PLUS, inv1_returns_volatility=0.05, monthly_expenses=9000), 31),
    ("stressed_lowvol_no_tax", dict(scenarios.STRESSED, inv1_returns_volatility=0.045, monthly_expenses=6500.0), 60),
    ("high_rates", dict(scenarios.SYNTH_C3, inv1_realized_gains_tax_rate=0.45, inv2_realized_gains_tax_rate=0.3,
                        inv1_returns_mean=0.03, inv1_returns_volatility=0.045, monthly_expenses=7000.0), 100),
    ("tiny_balances", dict(scenarios.SYNTH_C3, initial_balance=10.0, monthly_contribution=1.0, monthly_expenses=2.0), 24),
]


@pytest.mark.parametrize("name,cfg,wm", LOWVOL_CASES, ids=[c[0] for c in LOWVOL_CASES])
def test_benchmarked_variant_meets_the_replay_gate(name, cfg, wm):
    """The kernel variant bench.py times — fast build, bounded-return specialisation with the short
    exp polynomial and the lean month steps — fed the REFERENCE's numpy draws (proven on the host to
    satisfy the bound, MCR_FLAG_SMALL_RETURNS): bit-identical success flags / ruin months and
    balances within 1e-9 relative, against the strict build AND against the CPU oracle."""
    import torch

    from gpu_util import small_returns_hold
    from oracle import oracle as orc

    n = 2048
    o = orc.OracleSimulator(cfg)
    o.use_final_seeds()
    shocks = orc.shocks_for_seeds(o.p, wm, o.seeds.path_seeds(n))                  # (n, rows, 3)
    recs, traj, real, wr = orc.run_batch(o.p, wm, shocks, n_threads=4)
    sim = make_sim(cfg)
    assert small_returns_hold(sim, shocks)
    dev = torch.from_numpy(np.ascontiguousarray(shocks.transpose(1, 2, 0))).to("cuda")
    hs = device_batch_to_host(sim.run_batch_device(wm, n, shocks=dev))
    assert sim.native_context.last_variant in (1, 2)
    hf = device_batch_to_host(sim.run_batch_device(wm, n, shocks=dev, _fast_replay=True, _small_returns=True))
    variant = sim.native_context.last_variant
    assert variant in (3, 4, 5, 6)
    # ... which is the variant a native-draw launch of this scenario uses (what bench.py times)
    sim.run_batch_device(wm, 256, series=False)
    assert sim.native_context.last_variant == variant
    want_ruin = np.where(np.isnan(recs["years_to_ruin"]), -1, np.rint(recs["years_to_ruin"] * 12)).astype(np.int32)
    for h in (hs, hf):
        assert np.array_equal(h["success"], recs["success"].astype(bool))
        assert np.array_equal(h["ruin_month"], want_ruin)
        assert np.array_equal(np.isnan(h["wr"]), np.isnan(wr))
    # the parity build against the oracle: plain 1e-9 (it is ~1e-13 in practice)
    for key, ref in (("start", recs["start_balance"]), ("final", recs["final_balance"]),
                     ("fy_gross", recs["first_year_gross"]), ("fy_real", recs["first_year_real"]),
                     ("infl", recs["inflation_at_ret"]), ("traj", traj), ("real", real)):
        assert_close(hs[key], ref)
    assert_close(np.nan_to_num(hs["wr"]), np.nan_to_num(wr))
    # the benchmarked variant against the oracle (see assert_close_fast for the per-path floor)
    _assert_fast_gate(hf, {"start": recs["start_balance"], "final": recs["final_balance"],
                           "fy_gross": recs["first_year_gross"], "fy_real": recs["first_year_real"],
                           "infl": recs["inflation_at_ret"], "traj": traj, "real": real, "wr": wr})
    _assert_fast_gate(hf, hs)


def test_benchmarked_variant_meets_the_gate_on_its_own_draws():
    """Same gate on the engine's own Philox draws (their tails reach further than the 2048 numpy
    paths above): strict draws replayed by the strict build vs the fast bounded-return variant."""
    import torch

    for cfg, wm in ((scenarios.SYNTH_C3, 240), (scenarios.CONFIG_JSON, 200)):
        n = 16384
        strict = make_sim(cfg, strict=True)
        fast = make_sim(cfg)
        n_rows = wm + 12 * cfg["retirement_years"]
        shocks = torch.empty((n_rows, 3, n), dtype=torch.float64, device="cuda")
        strict.native_context.draw_shocks(1, 0, n, n_rows, shocks, n, strict=True)
        hs = device_batch_to_host(strict.run_batch_device(wm, n, shocks=shocks))
        hf = device_batch_to_host(fast.run_batch_device(wm, n, shocks=shocks, _fast_replay=True, _small_returns=True))
        assert fast.native_context.last_variant in (3, 5)
        assert np.array_equal(hs["success"], hf["success"])
        assert np.array_equal(hs["ruin_month"], hf["ruin_month"])
        assert np.array_equal(np.isnan(hf["wr"]), np.isnan(hs["wr"]))
        _assert_fast_gate(hf, hs)
        # how much of it needed the per-path floor at all: balances at their own scale obey the plain gate
        big = np.abs(hs["traj"]) >= 1e-3 * hs["traj"].max(axis=1, keepdims=True)
        rel_err = np.abs(hf["traj"] - hs["traj"])[big] / np.abs(hs["traj"])[big]
        assert rel_err.max() < 1e-9, rel_err.max()


def _assert_fast_gate(hf, ref):
    """hf (fast build, dict of device_batch_to_host) against ref (same keys): 1e-9 relative with the
    per-path floor of gpu_util.assert_close_fast on balances; withdrawals and rates are sums, not
    differences, and get the plain relative gate."""
    from gpu_util import assert_close_fast

    peak = np.abs(np.asarray(ref["traj"])).max(axis=1)
    peak_real = np.abs(np.asarray(ref["real"])).max(axis=1)
    assert_close_fast(hf["traj"], ref["traj"], peak)
    assert_close_fast(hf["real"], ref["real"], peak_real)
    assert_close_fast(hf["start"], ref["start"], peak)
    assert_close_fast(hf["final"], ref["final"], peak)
    for key in ("fy_gross", "fy_real", "infl"):
        assert_close(hf[key], ref[key])
    assert_close(np.nan_to_num(hf["wr"]), np.nan_to_num(ref["wr"]))


def test_years_to_ruin_column_is_exactly_the_reference_division():
    """summary_df["YearsToRuin"] = ruin_month / 12 with an IEEE division for every month value
    (a reciprocal multiply is 1 ulp off for a third of them)."""
    import torch

    sim = make_sim(scenarios.STRESSED)
    b = sim.run_batch_device(0, 4096, series=False)
    months = torch.arange(-1, 1300, dtype=torch.int32, device="cuda")
    b.ruin = months
    b.n = months.numel()
    out = torch.empty(months.numel(), dtype=torch.float64, device="cuda")
    b.years_to_ruin_into(out)
    got = out.cpu().numpy()
    assert np.isnan(got[0])
    assert np.array_equal(got[1:], np.arange(0, 1300) / 12)           # bit for bit
    df = sim.run_monte_carlo_simulations(0, 3000)[0]
    ruin = sim._last_batch.ruin.cpu().numpy()
    want = np.where(ruin < 0, np.nan, ruin / 12)
    assert np.array_equal(df["YearsToRuin"].to_numpy(), want, equal_nan=True)
