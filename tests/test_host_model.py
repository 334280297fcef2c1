"""CPU check of the device arithmetic: the engine's path headers compiled for the host
(tests/host_model, test infrastructure) against the CPU oracle on the reference's own numpy draws.

  * strict build  == oracle bit for bit (same operation order, libm exp on both sides);
  * fast build and the fast build's lean month steps (MCR_FLAG_SMALL_RETURNS variant — the one
    bench.py times) : identical success flags and ruin months, balances / series within 1e-9
    relative (BASELINE.json north_star), including scenarios where many paths fail, so the
    hand-over between lean and general months near ruin is exercised.

The same comparisons run against the real CUDA kernels in tests/test_gpu_parity.py; this file
exists so that a change to csrc/mcr_path.cuh is checked before GPU time is spent."""
from __future__ import annotations

import math

import numpy as np
import pytest

import host_model_util as hm
import scenarios
from gpu_util import assert_close
from oracle import oracle as orc

CASES = [
    ("synth_c3", scenarios.SYNTH_C3, 240), ("synth_c3_short", scenarios.SYNTH_C3, 7),
    ("synth_c3_vol", scenarios.SYNTH_C3_VOL, 240), ("config_json", scenarios.CONFIG_JSON, 233),
    ("config_json_early", scenarios.CONFIG_JSON, 120), ("jorge_plus", scenarios.JORGE_PLUS, 75),
    ("stressed", scenarios.STRESSED, 60), ("stressed0", scenarios.STRESSED, 0), ("test_base", scenarios.TEST_BASE, 36),
    ("tax_heavy", scenarios.TAX_HEAVY, 150), ("annual_both", scenarios.ANNUAL_BOTH, 100),
    ("synth_c3_annual", scenarios.SYNTH_C3_ANNUAL, 240), ("alloc0", scenarios.CORNER_ALLOC0, 50),
    ("broke", scenarios.CORNER_BROKE, 12),
    # low-volatility variants: every monthly log-return of the reference's draws stays below 0.1,
    # so the lean-capable kernel variant (MCR_FLAG_SMALL_RETURNS) is compared too
    ("jorge_plus_lowvol", dict(scenarios.JORGE_PLUS, inv1_returns_volatility=0.05), 75),
    ("jorge_plus_lowvol_late", dict(scenarios.JORGE_PLUS, inv1_returns_volatility=0.05, monthly_expenses=9000), 31),
    ("stressed_lowvol", dict(scenarios.STRESSED, inv1_returns_volatility=0.045, monthly_expenses=6500.0), 60),
    ("test_base_lowvol", dict(scenarios.TEST_BASE, inv1_returns_volatility=0.05), 36),
    ("high_rates", dict(scenarios.SYNTH_C3, inv1_realized_gains_tax_rate=0.45, inv2_realized_gains_tax_rate=0.3,
                        inv1_returns_mean=0.03, inv1_returns_volatility=0.045, monthly_expenses=7000.0), 100),
    ("rate_above_lean_limit", dict(scenarios.SYNTH_C3, inv1_realized_gains_tax_rate=0.95), 60),
    ("tiny_balances", dict(scenarios.SYNTH_C3, initial_balance=10.0, monthly_contribution=1.0, monthly_expenses=2.0), 24),
]
LEAN_EXPECTED = {"synth_c3", "config_json", "jorge_plus_lowvol", "stressed_lowvol", "test_base_lowvol", "high_rates"}


def _inputs(cfg, wm, n):
    sim = orc.OracleSimulator(cfg)
    sim.use_final_seeds()
    shocks = orc.shocks_for_seeds(sim.p, wm, sim.seeds.path_seeds(n))
    recs, traj, real, wr = orc.run_batch(sim.p, wm, shocks, n_threads=4)
    return sim.p, shocks, recs, traj, real, wr


def _compare(h, recs, traj, real, wr, rel):
    assert np.array_equal(h["success"], recs["success"].astype(bool))
    want_ruin = np.where(np.isnan(recs["years_to_ruin"]), -1, np.rint(recs["years_to_ruin"] * 12)).astype(np.int32)
    assert np.array_equal(h["ruin_month"], want_ruin)
    for key, ref in (("start", recs["start_balance"]), ("final", recs["final_balance"]), ("fy_gross", recs["first_year_gross"]),
                     ("fy_real", recs["first_year_real"]), ("infl", recs["inflation_at_ret"])):
        assert_close(h[key], ref, rel=rel)
    assert_close(h["traj"], traj, rel=rel)
    assert_close(h["real"], real, rel=rel)
    assert np.array_equal(np.isnan(h["wr"]), np.isnan(wr))
    assert_close(np.nan_to_num(h["wr"]), np.nan_to_num(wr), rel=rel)


@pytest.mark.parametrize("name,cfg,wm", CASES, ids=[c[0] for c in CASES])
def test_strict_header_equals_oracle(name, cfg, wm):
    p, shocks, recs, traj, real, wr = _inputs(cfg, wm, 64)
    h = hm.replay(p, wm, shocks, hm.STRICT)
    _compare(h, recs, traj, real, wr, rel=1e-13)


@pytest.mark.parametrize("name,cfg,wm", CASES, ids=[c[0] for c in CASES])
def test_fast_and_lean_headers_within_1e9_of_oracle(name, cfg, wm):
    p, shocks, recs, traj, real, wr = _inputs(cfg, wm, 256)
    h = hm.replay(p, wm, shocks, hm.FAST)
    _compare(h, recs, traj, real, wr, rel=1e-9)
    if hm.small_returns(p, shocks):
        h2 = hm.replay(p, wm, shocks, hm.FAST_SMALL)
        _compare(h2, recs, traj, real, wr, rel=1e-9)
        if name in LEAN_EXPECTED:
            assert h2["cfg"] in (3, 4, 5, 6)  # the lean-capable specialisations really ran ...
            assert h2["lean_months"] > 0.5 * int(h2["executed"].sum())  # ... and took the lean step
    else:
        assert name not in LEAN_EXPECTED


def test_philox_known_answer_vectors():
    """Random123 KAT for the numpy restatement the draw-layout tests (CPU and GPU) are built on."""
    import philox_ref

    z = np.zeros(1, dtype=np.uint32)
    out = philox_ref.philox4x32_10((z, z, z, z), (0, 0))
    assert [int(x[0]) for x in out] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = np.full(1, 0xFFFFFFFF, dtype=np.uint32)
    out = philox_ref.philox4x32_10((f, f, f, f), (0xFFFFFFFF, 0xFFFFFFFF))
    assert [int(x[0]) for x in out] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]


def test_native_draw_layout_is_philox_with_three_box_muller_pairs_per_call():
    """csrc/mcr_rng.cuh compiled for the host == the numpy restatement: counter layout (path,
    month / 2, stream), bit fields of the three pairs, month parity, odd month counts, 64-bit paths."""
    import philox_ref

    p = orc.params_from_config(dict(scenarios.SYNTH_C3, equity_inflation_correlation=-0.5))
    for first, n, months, stream in ((0, 33, 41, 1), (1_000_000_007, 17, 6, 0), ((1 << 33) + 5, 9, 13, 1)):
        got = hm.draw(p, 20260101, stream, first, n, months)
        want = philox_ref.shocks(20260101, stream, first, n, months, -0.5)
        assert got.shape == want.shape
        assert np.max(np.abs(got - want)) < 2e-5   # fp32 transform vs float64


def test_native_draws_are_standard_normal_and_uncorrelated():
    """Distribution of the layout itself (26-bit radius, 16- / 12-bit angles): moments, a
    Kolmogorov-Smirnov distance per component and per month parity, tails, and the correlations
    between the normals that share a Philox call."""
    from scipy import stats

    import philox_ref

    n, months = 120_000, 8
    x = philox_ref.shocks(777, 1, 0, n, months, 0.0)            # rho = 0: inflation == the independent normal
    for comp in range(3):
        for parity in (0, 1):
            v = x[parity::2, comp].ravel()
            assert abs(v.mean()) < 5 / math.sqrt(v.size)
            assert abs(v.var() - 1.0) < 5 * math.sqrt(2.0 / v.size)
            assert abs(stats.kurtosis(v)) < 0.03
            assert stats.kstest(v, "norm").statistic < 1.63 / math.sqrt(v.size)      # alpha = 1 %
            assert abs((np.abs(v) > 3.0).mean() / 0.0026998 - 1.0) < 0.1
    flat = x.transpose(0, 1, 2).reshape(months * 3, n)        # every (month, component) row against every other
    c = np.corrcoef(flat)
    off = c[~np.eye(len(c), dtype=bool)]
    assert np.abs(off).max() < 5 / math.sqrt(n)
    assert np.abs(x).max() < 6.2
