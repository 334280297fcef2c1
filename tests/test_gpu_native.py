"""GPU tests of the native-RNG path, the device aggregations and the batched search kernel."""
from __future__ import annotations

import math

import numpy as np
import pandas as pd
import pytest

import scenarios
from gpu_util import assert_close, device_batch_to_host, make_sim

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------------
# Philox + normals (the numpy restatement lives in tests/philox_ref.py; its Random123 KAT and the
# distribution of the draw layout are checked on the CPU in tests/test_host_model.py)
# ---------------------------------------------------------------------------------------------
def test_device_shocks_are_philox_box_muller():
    """mcr_draw_shocks == Philox4x32-10(key(main_seed); path, month / 2, stream) + three Box-Muller
    pairs per call (csrc/mcr_rng.cuh), for both arithmetic builds."""
    import torch

    import philox_ref

    cfg = dict(scenarios.SYNTH_C3, equity_inflation_correlation=-0.5)
    sim = make_sim(cfg, strict=True)
    for n, months, first in ((257, 41, 1_000_000_007), (64, 6, (1 << 33) + 5)):
        sh = torch.empty((months, 3, n), dtype=torch.float64, device="cuda")
        sim.native_context.draw_shocks(1, first, n, months, sh, n, strict=True)
        got = sh.cpu().numpy()
        want = philox_ref.shocks(sim.main_seed, 1, first, n, months, -0.5)
        assert np.max(np.abs(got - want)) < 5e-5  # fp32 transform on the device
        # the fast build (MUFU lg2/sin/cos) draws the same normals to ~1e-5
        sim.native_context.draw_shocks(1, first, n, months, sh, n, strict=False)
        assert np.max(np.abs(sh.cpu().numpy() - want)) < 2e-4


def test_native_shock_distribution():
    import torch

    sim = make_sim(dict(scenarios.SYNTH_C3, equity_inflation_correlation=-0.5))
    n, months = 200_000, 24
    sh = torch.empty((months, 3, n), dtype=torch.float64, device="cuda")
    sim.native_context.draw_shocks(0, 0, n, months, sh, n)
    x = sh.cpu().numpy()
    eq, inf, pr = x[:, 0].ravel(), x[:, 1].ravel(), x[:, 2].ravel()
    N = eq.size
    for v in (eq, inf, pr):
        assert abs(v.mean()) < 5 / math.sqrt(N)
        assert abs(v.var() - 1.0) < 5 * math.sqrt(2.0 / N)
        assert abs((v**4).mean() - 3.0) < 0.05
        assert abs(v).max() < 7.0
    assert abs(np.corrcoef(eq, inf)[0, 1] + 0.5) < 0.005
    assert abs(np.corrcoef(eq, pr)[0, 1]) < 0.005
    # months of one path and neighbouring paths are uncorrelated
    assert abs(np.corrcoef(x[0, 0], x[1, 0])[0, 1]) < 0.01
    assert abs(np.corrcoef(x[0, 0, :-1], x[0, 0, 1:])[0, 1]) < 0.01
    # KS distance to the normal CDF
    from scipy import stats

    assert stats.kstest(eq[:100_000], "norm").statistic < 0.006
    # seed streams differ
    sh2 = torch.empty_like(sh)
    sim.native_context.draw_shocks(1, 0, n, months, sh2, n)
    assert abs(np.corrcoef(x[0, 0], sh2[0, 0].cpu().numpy())[0, 1]) < 0.01


# ---------------------------------------------------------------------------------------------
# native-RNG timeline == replay of its own draws == CPU oracle on those draws
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg,wm", [(scenarios.SYNTH_C3, 240), (scenarios.TAX_HEAVY, 137), (scenarios.JORGE_PLUS, 75),
                                    (scenarios.STRESSED, 0), (scenarios.ANNUAL_BOTH, 61)],
                         ids=["c3", "tax_heavy", "jorge_plus", "stressed", "annual_both"])
def test_native_strict_equals_oracle_on_device_draws(cfg, wm):
    import torch

    from oracle import oracle as orc

    n = 2048
    sim = make_sim(cfg, strict=True)
    sim.use_final_seeds()
    R = cfg["retirement_years"]
    n_rows = max(wm + 12 * R, 1)
    sh = torch.empty((n_rows, 3, n), dtype=torch.float64, device="cuda")
    sim.native_context.draw_shocks(1, 0, n, n_rows, sh, n, strict=True)
    native = device_batch_to_host(sim.run_batch_device(wm, n))
    replay = device_batch_to_host(sim.run_batch_device(wm, n, shocks=sh))
    for k in ("start", "final", "fy_gross", "fy_real", "infl", "traj", "real"):
        assert np.array_equal(native[k], replay[k]), k  # same kernel body, same draws: bit-equal
    assert np.array_equal(native["success"], replay["success"])
    recs, traj, real, wr = orc.run_batch(orc.params_from_config(cfg), wm,
                                         np.ascontiguousarray(sh.cpu().numpy().transpose(2, 0, 1)), n_threads=4)
    assert np.array_equal(native["success"], recs["success"].astype(bool))
    assert_close(native["final"], recs["final_balance"])
    assert_close(native["start"], recs["start_balance"])
    assert_close(native["traj"], traj)
    assert_close(native["real"], real)
    assert np.array_equal(np.isnan(native["wr"]), np.isnan(wr))
    assert_close(np.nan_to_num(native["wr"]), np.nan_to_num(wr))
    ruin = np.where(np.isnan(recs["years_to_ruin"]), -1, np.round(recs["years_to_ruin"] * 12)).astype(int)
    assert np.array_equal(native["ruin_month"], ruin)
    # executed months: the oracle-side count is the number of shock rows consumed
    assert native["executed"] > 0 and native["executed"] <= n * n_rows
    assert native["ruin_hist"].sum() == (~native["success"]).sum()


def test_native_results_do_not_depend_on_sharding_or_candidate():
    """Path i's draws depend on (seed, stream, global index, month) only: splitting a batch in
    two launches (two GPUs' shards) gives the same per-path results bit for bit."""
    sim = make_sim(scenarios.SYNTH_C3_VOL)
    n, wm = 5000, 120
    whole = device_batch_to_host(sim.run_batch_device(wm, n))
    a = device_batch_to_host(sim.run_batch_device(wm, 1777, first_path=0))
    b = device_batch_to_host(sim.run_batch_device(wm, n - 1777, first_path=1777))
    for k in ("start", "final", "fy_real", "traj", "wr"):
        assert np.array_equal(np.concatenate([a[k], b[k]]), whole[k], equal_nan=True), k
    assert a["success_count"] + b["success_count"] == whole["success_count"]
    # common random numbers: the accumulation prefix of a longer candidate is the shorter one
    longer = device_batch_to_host(sim.run_batch_device(wm + 12, n))
    assert np.array_equal(longer["traj"][:, : wm // 12 + 1], whole["traj"][:, : wm // 12 + 1])


def test_native_success_probability_within_binomial_3sigma_of_reference_draws():
    """Native Philox runs vs the reference's numpy draws (through the pinned oracle)."""
    from oracle import oracle as orc

    cfg = dict(scenarios.SYNTH_C3_VOL)
    wm = 200
    n_ref = 6000
    o = orc.OracleSimulator(cfg, n_threads=8)
    recs, _, _, _ = o.run_raw(wm, n_ref)
    p_ref = recs["success"].mean()
    n = 400_000
    h = make_sim(cfg).run_aggregates(wm, n, bands=False)
    p = h["success_probability"] / 100.0
    sigma = math.sqrt(p_ref * (1 - p_ref) / n_ref + p * (1 - p) / n)
    assert 0.02 < p_ref < 0.98, p_ref  # a scenario where the test has power
    assert abs(p - p_ref) < 3 * sigma, (p, p_ref, sigma)
    # median start balance agrees statistically as well (1% is ~5 sigma of the median here)
    assert abs(h["median_start_balance"] / np.median(recs["start_balance"]) - 1) < 0.01


# ---------------------------------------------------------------------------------------------
# device aggregations
# ---------------------------------------------------------------------------------------------
def _dev(a, dtype=None):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda") if dtype is None else torch.from_numpy(
        np.ascontiguousarray(a).astype(dtype)).to("cuda")


@pytest.mark.parametrize("n", [1, 2, 3, 10, 257, 1000, 4097, 100_003])
def test_quantiles_match_pandas_bit_for_bit(n):
    import torch

    rng = np.random.default_rng(n)
    rows = 9
    x = np.exp(rng.normal(10, 2, (rows, n)))
    x[1] = np.round(x[1], -4)            # heavy duplicates
    x[2, rng.random(n) < 0.4] = 0.0      # zero-padded failed paths
    x[3, rng.random(n) < 0.5] = np.nan   # NaN-skipping (WR bands)
    x[4] = -x[4]                         # negative keys
    x[5, :] = np.nan                     # empty row
    x[6] = 7.25                          # constant row
    x[7, : n // 2] *= -1
    q = [0.05, 0.10, 0.25, 0.50, 0.75, 0.90, 0.95]
    sim = make_sim(scenarios.TEST_BASE)
    out = torch.empty((rows, len(q)), dtype=torch.float64, device="cuda")
    cnt = torch.empty(rows, dtype=torch.int64, device="cuda")
    sim.native_context.quantiles(_dev(x), n, n, rows, q, out, counts=cnt)
    want = pd.DataFrame(x.T).quantile(q, axis=0).T.to_numpy()
    got = out.cpu().numpy()
    assert np.array_equal(got, want, equal_nan=True), np.argwhere(~((got == want) | (np.isnan(got) & np.isnan(want))))
    assert cnt.cpu().tolist() == pd.DataFrame(x.T).count().tolist()
    # wider quantile set (server.py:453-455) and the median rule, with a cohort mask
    q9 = [0.01, 0.05, 0.10, 0.25, 0.50, 0.75, 0.90, 0.95, 0.99]
    mask = rng.random(n) < 0.7
    mask[0] = True
    out9 = torch.empty(len(q9), dtype=torch.float64, device="cuda")
    sim.native_context.quantiles(_dev(x[0]), n, n, 1, q9, out9, mask=_dev(mask, np.uint8))
    assert np.array_equal(out9.cpu().numpy(), pd.Series(x[0][mask]).quantile(q9).to_numpy())
    med = torch.empty(1, dtype=torch.float64, device="cuda")
    sim.native_context.quantiles(_dev(x[3]), n, n, 1, [0.5], med, median=True)
    want_med = pd.Series(x[3]).median()
    assert (np.isnan(want_med) and np.isnan(med.item())) or med.item() == want_med


def _zero_padded_rows(n, rng):
    """Band-row shapes of config #5: a bulk of positive balances next to a mass of exact zeros
    (failed paths pad +0.0, simulation.py:905-912)."""
    base = np.exp(rng.normal(17, 0.8, n))
    rows = []
    for frac in (0.001, 0.022, 0.5, 0.99):
        r = base.copy(); r[rng.random(n) < frac] = 0.0
        rows.append(r)                                         # 0-3: targets fall on both sides of the zero mass
    rows.append(np.zeros(n))                                   # 4: every path failed
    r = base.copy(); r[rng.random(n) < 0.3] = 0.0; r[rng.random(n) < 0.02] = 1e-9
    rows.append(r)                                             # 5: the mass below the bulk is NOT only zeros -> restart
    r = base.copy(); r[rng.random(n) < 0.2] = 0.0; r[rng.random(n) < 0.1] *= -1.0
    rows.append(r)                                             # 6: zeros between negative and positive keys
    r = base.copy(); r[rng.random(n) < 0.3] = 0.0; r[rng.random(n) < 0.3] = np.nan
    rows.append(r)                                             # 7: with NaN (skipped)
    r = np.zeros(n); r[: 3] = base[: 3]
    rows.append(r)                                             # 8: three survivors
    r = base.copy(); r[rng.random(n) < 0.03] = 0.0
    far = rng.choice(n, 5, replace=False); r[far[:3]] = 1e-3; r[far[3:]] = -2.0
    rows.append(r)                                             # 9: zeros + five keys far from the bulk (below AND
    return np.stack(rows)                                      #    above zero) that a sample is unlikely to see


@pytest.mark.parametrize("n", [300_000, 16384 * 1024 + 12_345])
def test_quantiles_of_rows_with_a_mass_of_zeros(n):
    """The adaptive start takes the extremes over the non-zero keys and resolves targets inside the
    zero mass directly; rows of more than 16 M elements are scanned in 1024 long chunks."""
    import torch

    rng = np.random.default_rng(5)
    x = _zero_padded_rows(n, rng)
    if n > 1_000_000:
        x = x[[1, 3, 9]]
    q = [0.01, 0.05, 0.10, 0.25, 0.50, 0.75, 0.90, 0.95, 0.99]
    sim = make_sim(scenarios.TEST_BASE)
    rows = x.shape[0]
    out = torch.empty((rows, len(q)), dtype=torch.float64, device="cuda")
    cnt = torch.empty(rows, dtype=torch.int64, device="cuda")
    xd = _dev(x)
    sim.native_context.quantiles(xd, n, n, rows, q, out, counts=cnt)
    want = np.stack([pd.Series(r).quantile(q).to_numpy() for r in x])
    got = out.cpu().numpy()
    assert np.array_equal(got, want, equal_nan=True), np.argwhere(~((got == want) | (np.isnan(got) & np.isnan(want))))
    assert cnt.cpu().tolist() == [int(np.count_nonzero(~np.isnan(r))) for r in x]
    mask = rng.random(n) < 0.6
    med = torch.empty((rows, 1), dtype=torch.float64, device="cuda")
    sim.native_context.quantiles(xd, n, n, rows, [0.5], med, mask=_dev(mask, np.uint8), median=True)
    want_med = np.array([pd.Series(r[mask]).median() for r in x])
    assert np.array_equal(med.cpu().numpy()[:, 0], want_med, equal_nan=True)


def test_quantiles_of_long_rows_are_exact_whatever_the_sampled_extremes_miss():
    """Rows longer than 8 chunks take their adaptive start from a SAMPLE (every 16th chunk of
    16384 elements); the first digit pass accounts for everything outside the sampled prefix
    exactly and a target that falls outside it restarts the row. Adversarial layouts, bit for
    bit against pandas."""
    import torch

    n, chunk = 400_000, 16384
    rng = np.random.default_rng(99)
    sampled = np.zeros(n, dtype=bool)
    for c0 in range(0, n, 16 * chunk):
        sampled[c0:c0 + chunk] = True
    base = np.exp(rng.normal(13, 0.4, n))
    rows = []
    rows.append(base.copy())                                   # 0: ordinary i.i.d. row
    rows.append(np.sort(base))                                 # 1: ascending
    rows.append(np.sort(base)[::-1].copy())                    # 2: descending
    r = base.copy(); r[~sampled] = np.where(rng.random((~sampled).sum()) < 0.45, 1e-3 * base[~sampled], base[~sampled])
    rows.append(r)                                             # 3: 40 % of the mass far BELOW anything the sample sees
    r = base.copy(); r[~sampled][:0] = 0; idx = np.flatnonzero(~sampled)[:10]; r[idx] = 0.0
    rows.append(r)                                             # 4: ten zeros the sample misses (stay 'below', no restart)
    r = base.copy(); r[sampled] = np.nan
    rows.append(r)                                             # 5: the sample sees no valid element at all
    r = base.copy(); r[sampled] = 5.0e5
    rows.append(r)                                             # 6: the sample sees ONE value, the row varies
    r = base.copy(); r[~sampled] = -base[~sampled]
    rows.append(r)                                             # 7: everything negative outside the sample
    x = np.stack(rows)
    q = [0.0, 0.05, 0.10, 0.25, 0.50, 0.75, 0.90, 0.95, 1.0]
    sim = make_sim(scenarios.TEST_BASE)
    out = torch.empty((len(rows), len(q)), dtype=torch.float64, device="cuda")
    cnt = torch.empty(len(rows), dtype=torch.int64, device="cuda")
    sim.native_context.quantiles(_dev(x), n, n, len(rows), q, out, counts=cnt)
    want = pd.DataFrame(x.T).quantile(q, axis=0).T.to_numpy()
    got = out.cpu().numpy()
    assert np.array_equal(got, want, equal_nan=True), np.argwhere(got != want)
    assert cnt.cpu().tolist() == pd.DataFrame(x.T).count().tolist()
    med = torch.empty(1, dtype=torch.float64, device="cuda")
    mask = np.ones(n, dtype=np.uint8); mask[sampled] = 0
    sim.native_context.quantiles(_dev(x[0]), n, n, 1, [0.5], med, median=True, mask=_dev(mask))   # masked-out sample
    assert med.item() == np.median(x[0][~sampled])

@pytest.mark.gpu
@pytest.mark.parametrize("n", [5_000, 1_000_000])
def test_quantiles_whatever_the_first_digit_window_looks_like(n):
    """The first digit pass lays ~8 K equal bins over the (sampled) key range of a row: ranges
    narrower than the low 32 key bits, ranges across a power of two, ranges stretched by outliers so
    far that the bulk stays in one bucket through every full pass (the tail then scans the row),
    buckets of a few distinct values, sign changes. Bit for bit against pandas."""
    import torch

    rng = np.random.default_rng(n + 17)
    u = rng.random(n)
    rows = [
        1.0 + u * 1e-12,                                   # 0: ~4500 distinct keys, all in the low key bits
        1.0 + u * 1e-7,                                    # 1: bin shift < 32
        1048576.0 * (0.999 + 0.002 * u),                   # 2: across 2^20 (no common leading bits to speak of)
        np.where(u < 0.5, -1e-300, 1e-300) * (1 + rng.random(n)),   # 3: across zero, nothing near it
        np.concatenate([[1e-200, 1e200], 1.0 + rng.random(n - 2) * 1e-9]),   # 4: bulk in one bucket of distinct values
        np.round(rng.normal(0, 3, n)),                     # 5: ~20 distinct values, negative to positive, with -0/+0
        rng.standard_cauchy(n) * 1e6,                      # 6: heavy tails on both sides
        np.exp(rng.normal(14, 1.0, n)),                    # 7: the ordinary band row
        np.full(n, -3.5),                                  # 8: constant, negative
        np.where(u < 0.9, 2.0, 2.0 + 4.4e-16),             # 9: two adjacent keys
    ]
    x = np.stack(rows)
    q = [0.0, 0.05, 0.10, 0.25, 0.50, 0.75, 0.90, 0.95, 1.0]
    sim = make_sim(scenarios.TEST_BASE)
    out = torch.empty((len(rows), len(q)), dtype=torch.float64, device="cuda")
    cnt = torch.empty(len(rows), dtype=torch.int64, device="cuda")
    sim.native_context.quantiles(_dev(x), n, n, len(rows), q, out, counts=cnt)
    want = pd.DataFrame(x.T).quantile(q, axis=0).T.to_numpy()
    got = out.cpu().numpy()
    assert np.array_equal(got, want, equal_nan=True), np.argwhere(got != want)
    assert cnt.cpu().tolist() == [n] * len(rows)
    mask = rng.random(n) < 0.5
    med = torch.empty((len(rows), 1), dtype=torch.float64, device="cuda")
    sim.native_context.quantiles(_dev(x), n, n, len(rows), [0.5], med, mask=_dev(mask, np.uint8), median=True)
    want_med = np.array([np.median(r[mask]) for r in x])
    assert np.array_equal(med.cpu().numpy()[:, 0], want_med)
    # MCR_SEL_MINMAX rows (the histogram ranges of a step): the exact extreme elements of the cohort
    ctx = sim.native_context
    xd, md = _dev(x), _dev(mask, np.uint8)
    mm = torch.empty((len(rows), 16), dtype=torch.float64, device="cuda")
    ctx.quantiles_rows(ctx.select_rows([(xd, n, md, [0.0, 1.0], "minmax")]), mm)
    got_mm = mm.cpu().numpy()[:, :2]
    assert np.array_equal(got_mm[:, 0], x[:, mask].min(axis=1)) and np.array_equal(got_mm[:, 1], x[:, mask].max(axis=1))


def test_multi_row_select_with_empty_and_tiny_rows_next_to_long_ones():
    """One launch sequence over rows of very different lengths — none, one element, a few, a long sampled
    row, 18 targets in one row: every grid line of the fused scan kernels must count its CTAs correctly,
    whether they had work or not."""
    import torch

    rng = np.random.default_rng(21)
    long_row = np.exp(rng.normal(11, 1.2, 700_003))
    rows = [np.empty(0), np.array([3.25]), np.array([2.0, -1.0, 5.5]), long_row, rng.normal(0, 1, 40_000)]
    q9 = [0.01, 0.05, 0.10, 0.25, 0.50, 0.75, 0.90, 0.95, 0.99]
    sim = make_sim(scenarios.TEST_BASE)
    ctx = sim.native_context
    dev = [None if len(r) == 0 else _dev(r) for r in rows]
    specs = [(d, len(r), None, q9, False) for d, r in zip(dev, rows)]
    specs += [(dev[3], len(long_row), None, [0.5], True), (dev[3], len(long_row), None, [0.0, 1.0], "minmax"),
              (None, 0, None, [0.0, 1.0], "minmax")]
    out = torch.empty((len(specs), 16), dtype=torch.float64, device="cuda")
    cnt = torch.empty(len(specs), dtype=torch.int64, device="cuda")
    for _ in range(2):   # twice: the per-row arrival counters must be back at zero
        out.fill_(-7.0)
        ctx.quantiles_rows(ctx.select_rows(specs), out, counts=cnt)
        got = out.cpu().numpy()
        for i, r in enumerate(rows):
            want = pd.Series(r).quantile(q9).to_numpy() if len(r) else np.full(9, np.nan)
            assert np.array_equal(got[i, :9], want, equal_nan=True), i
        assert got[5, 0] == np.median(long_row)
        assert got[6, 0] == long_row.min() and got[6, 1] == long_row.max()
        assert np.isnan(got[7, 0]) and np.isnan(got[7, 1])
        assert cnt.cpu().tolist() == [0, 1, 3, len(long_row), 40_000, len(long_row), len(long_row), 0]


def test_histograms_match_numpy_and_frontend_rule():
    import torch

    rng = np.random.default_rng(5)
    sim = make_sim(scenarios.TEST_BASE)
    for n in (1, 7, 5000, 200_001):
        v = np.exp(rng.normal(15, 1, n))
        mask = rng.random(n) < 0.8
        mask[0] = True
        d_v, d_m = _dev(v), _dev(mask, np.uint8)
        rng2 = torch.empty(2, dtype=torch.float64, device="cuda")
        hist = torch.zeros(100, dtype=torch.int64, device="cuda")
        sim.native_context.minmax(d_v, n, rng2, mask=d_m, divisor=1e6)
        sim.native_context.histogram(d_v, n, 100, rng2, hist, mask=d_m, divisor=1e6, mode=0)
        sel = v[mask] / 1e6
        assert rng2.cpu().tolist() == [sel.min(), sel.max()]
        want, _ = np.histogram(sel, bins=100)
        assert hist.cpu().tolist() == want.tolist()
        # dashboard rule (HistogramChart.jsx:31-52)
        hist60 = torch.zeros(60, dtype=torch.int64, device="cuda")
        sim.native_context.minmax(d_v, n, rng2, mask=d_m)
        sim.native_context.histogram(d_v, n, 60, rng2, hist60, mask=d_m, mode=1)
        s = v[mask]
        lo, hi = s.min(), s.max()
        if hi <= lo:
            want60 = [len(s)] + [0] * 59
        else:
            w = (hi - lo) / 60
            idx = np.minimum(np.floor((s - lo) / w), 59).astype(int)
            want60 = np.bincount(idx, minlength=60).tolist()
        assert hist60.cpu().tolist() == want60
    # empty cohort
    none = _dev(np.zeros(10, dtype=np.uint8))
    sim.native_context.minmax(_dev(np.ones(10)), 10, rng2, mask=none)
    assert all(math.isnan(x) for x in rng2.cpu().tolist())


def test_run_aggregates_matches_host_reductions_of_summary_df():
    """Aggregate-only mode == the reductions server.py / main.py make over summary_df."""
    from monte_carlo_retirement_b200.simulation import median_first_year_withdrawal_rate

    sim = make_sim(scenarios.STRESSED)
    n, wm = 20_000, 150
    sim.use_final_seeds()
    agg = sim.run_aggregates(wm, n)
    summary, traj_pct, _, wr_pct, real_pct, _, wr_counts = sim.run_monte_carlo_simulations(wm, n)
    ok = summary["Success"].astype(bool)
    assert agg["success_probability"] == sim._success_probability(summary)
    assert agg["median_start_balance"] == float(summary["Start Balance"].median())
    assert agg["median_final_balance_successful"] == float(summary.loc[ok, "Final Balance"].median())
    assert agg["median_first_year_withdrawal_rate"] == median_first_year_withdrawal_rate(summary)
    q = summary["Final Balance"].quantile(list(agg["final_balance_quantiles"]))
    assert list(agg["final_balance_quantiles"].values()) == q.tolist()
    want100, _ = np.histogram(summary.loc[ok, "Final Balance"] / 1e6, bins=100)
    assert agg["final_balance_hist_musd_100"]["counts"] == want100.tolist()
    failed = summary.loc[~ok, "YearsToRuin"].dropna()
    ruin_hist = np.bincount(np.round(failed.to_numpy() * 12).astype(int), minlength=12 * 30 + 1)
    assert agg["ruin_month_hist"] == ruin_hist.tolist()
    assert np.array_equal(agg["trajectory_bands"].to_numpy(), traj_pct.to_numpy())
    assert np.array_equal(agg["withdrawal_rate_bands"].to_numpy(), wr_pct.to_numpy(), equal_nan=True)
    assert agg["withdrawal_rate_counts"] == wr_counts
    # bands computed on the device == pandas on the device-produced series
    b = sim._last_batch
    want = pd.DataFrame(b.traj.cpu().numpy()).quantile([0.05, 0.10, 0.25, 0.50, 0.75, 0.90, 0.95], axis=1).T
    assert np.array_equal(traj_pct.to_numpy(), want.to_numpy())
    want_wr = pd.DataFrame(b.wr.cpu().numpy()).quantile([0.05, 0.25, 0.50, 0.75, 0.95], axis=1).T
    assert np.array_equal(wr_pct.to_numpy(), want_wr.to_numpy(), equal_nan=True)
    assert wr_counts == pd.DataFrame(b.wr.cpu().numpy()).count(axis=1).tolist()


# ---------------------------------------------------------------------------------------------
# batched search
# ---------------------------------------------------------------------------------------------
def test_search_batch_equals_one_launch_per_candidate():
    sim = make_sim(scenarios.STRESSED)
    sim.use_search_seeds()
    n = 3000
    cands = [0, 1, 5, 12, 13, 60, 119, 120, 121, 240, 37, 36]
    counts, executed = sim.batched_success_counts(cands, n, with_executed=True)
    counts = counts.cpu().tolist()
    executed = executed.cpu().tolist()
    for c, k, e in zip(cands, counts, executed):
        h = device_batch_to_host(sim.run_batch_device(c, n, series=False))
        assert k == h["success_count"], c
        assert e == h["executed"], c
    # sharded launch accumulates into the same buffer
    import torch

    acc = torch.zeros(len(cands), dtype=torch.int64, device="cuda")
    sim.native_context.search_batch(0, cands, 0, 1000, acc)
    sim.native_context.search_batch(0, cands, 1000, n - 1000, acc)
    assert acc.cpu().tolist() == counts


@pytest.mark.parametrize("policy", ["waves", "grid", "probe", "auto"])
def test_device_search_makes_the_reference_decisions(policy):
    """The batched search returns what the reference's decision procedure returns on the same
    success table (here: the table the device itself produces, probe by probe)."""
    from oracle import oracle as orc

    cfg = dict(scenarios.STRESSED, target_probability=70.0, num_simulations_search=2000)
    sim = make_sim(cfg, search_policy=policy)
    events = []
    months, prob, curve = sim.find_minimum_working_months(verbose=False, progress_callback=events.append)
    seq = make_sim(cfg, search_policy="sequential")
    m2, p2, c2 = seq.find_minimum_working_months(verbose=False)
    assert (months, prob, curve) == (m2, p2, c2)
    # and the oracle's restatement of the decision logic agrees on the device table
    table = {pt["working_months"]: None for pt in curve}
    counts = sim.batched_success_counts(sorted(table), 2000).cpu().tolist()
    table = dict(zip(sorted(table), counts))
    m3, p3, c3, order = orc.search_decisions(lambda m: table[m] / 2000 * 100.0, 0, 70.0, 2000)
    assert (m3, p3, c3) == (months, prob, curve)
    assert [e["working_months"] for e in events if e["type"] == "search_iter"] == order
    if policy in ("waves", "grid", "auto"):
        assert sim.last_search_stats["launches"] <= (1 if policy == "grid" else 3)
    else:
        assert sim.last_search_stats["launches"] == len({p["working_months"] for p in curve})
    assert months > 0 and prob >= 70.0


def test_device_search_target_met_at_start_and_unreachable():
    rich = dict(scenarios.TEST_BASE, initial_balance=5e7, target_probability=90.0, num_simulations_search=500)
    assert make_sim(rich).find_minimum_working_months(verbose=False)[0] == 0
    broke = dict(scenarios.CORNER_BROKE, target_probability=99.0, num_simulations_search=64, retirement_years=40)
    months, prob, curve = make_sim(broke).find_minimum_working_months(verbose=False)
    assert months == -1 and prob == 0.0 and len(curve) == 36


def test_crn_success_probability_is_monotone_in_working_months():
    """tests/test_simulation_correctness.py:55-81 on the native RNG."""
    cfg = dict(scenarios.TEST_BASE, initial_balance=100_000.0, monthly_contribution=3_000.0,
               monthly_expenses=5_000.0, retirement_years=30, inv1_returns_mean=0.10,
               inv1_returns_volatility=0.12, inflation_rate_mean=0.04, inflation_rate_volatility=0.015,
               num_simulations_main=80, seed=123)
    sim = make_sim(cfg)
    sim.use_search_seeds()
    probs = []
    for months in range(0, 61, 6):
        summary = sim.run_monte_carlo_simulations(months, 80)[0]
        probs.append(sim._success_probability(summary))
    assert all(b + 1e-9 >= a for a, b in zip(probs, probs[1:])), probs
    counts = sim.batched_success_counts(list(range(0, 61, 6)), 80).cpu().numpy()
    assert np.allclose(counts / 80 * 100.0, probs)


def test_series_sweep_mode_equals_all_at_once(monkeypatch):
    """Config #5 shape: when the three series do not fit in HBM together they are produced one
    at a time by recomputing the batch; the bands must be identical."""
    sim = make_sim(scenarios.SYNTH_C3_VOL)
    a = sim.run_aggregates(100, 30_000)
    monkeypatch.setenv("MCR_SERIES_SWEEP", "1")
    b = sim.run_aggregates(100, 30_000)
    for k in ("trajectory_bands", "real_trajectory_bands", "withdrawal_rate_bands"):
        assert np.array_equal(a[k].to_numpy(), b[k].to_numpy(), equal_nan=True), k
    assert a["withdrawal_rate_counts"] == b["withdrawal_rate_counts"]
    assert a["success_count"] == b["success_count"]
    # a budget that holds two series: [traj, real] with the summary pass, [wr] recomputed
    monkeypatch.delenv("MCR_SERIES_SWEEP")
    T = len(a["trajectory_bands"])
    monkeypatch.setenv("MCR_SERIES_BUDGET_BYTES", str(8 * 30_000 * 2 * T + 1))
    assert sim._series_plan(30_000, T, 40, True, None) == [("traj", "real"), ("wr",)]
    c = sim.run_aggregates(100, 30_000, samples=True)
    for k in ("trajectory_bands", "real_trajectory_bands", "withdrawal_rate_bands"):
        assert np.array_equal(a[k].to_numpy(), c[k].to_numpy(), equal_nan=True), k
    assert a["withdrawal_rate_counts"] == c["withdrawal_rate_counts"]
    monkeypatch.delenv("MCR_SERIES_BUDGET_BYTES")
    d = sim.run_aggregates(100, 30_000, samples=True)
    assert c["sample_paths"] == d["sample_paths"] and c["real_sample_paths"] == d["real_sample_paths"]


def test_large_batch_aggregate_only_mode():
    """8e6 paths in one launch without series: counts add up, medians are sane."""
    sim = make_sim(scenarios.SYNTH_C3)
    h = sim.run_aggregates(240, 8_000_000, bands=False)
    assert h["num_simulations"] == 8_000_000
    assert h["executed_path_months"] <= 8_000_000 * 720
    assert sum(h["ruin_month_hist"]) == 8_000_000 - h["success_count"]
    assert sum(h["final_balance_hist_musd_100"]["counts"]) == h["success_count"]
    assert 99.0 < h["success_probability"] < 100.0


@pytest.mark.parametrize("cfg,wm", [(scenarios.SYNTH_C3, 240), (scenarios.TEST_BASE, 36), (scenarios.TAX_HEAVY, 260)],
                         ids=["c3_small_exp", "no_tax", "generic"])
def test_fast_native_tracks_strict_native(cfg, wm):
    """Throughput build with its own draws (MUFU normals, short exp polynomial when the host
    proved |x| < 0.1, compile-time tax configuration) vs the parity build on the same Philox
    counters: the normals differ by ~1e-6 absolute, so balances agree to ~1e-5 relative and
    the success flags of all but knife-edge paths coincide."""
    n = 50_000
    fast = device_batch_to_host(make_sim(cfg, strict=False).run_batch_device(wm, n))
    strict = device_batch_to_host(make_sim(cfg, strict=True).run_batch_device(wm, n))
    assert (fast["success"] != strict["success"]).mean() < 2e-4
    both = fast["success"] & strict["success"]
    assert both.sum() > 100   # (TAX_HEAVY succeeds on ~6 % of the paths at 260 working months)
    rel = np.abs(fast["final"][both] - strict["final"][both]) / np.maximum(strict["final"][both], 1.0)
    # (the tail of `rel` belongs to paths that end almost dry: a final balance that is the small
    # difference of large numbers amplifies the 1e-6 difference of the two normal transforms)
    assert np.median(rel) < 1e-5 and np.quantile(rel, 0.99) < 1e-3, (np.median(rel), np.quantile(rel, 0.99), rel.max())
    rel0 = np.abs(fast["start"] - strict["start"]) / np.maximum(strict["start"], 1.0)
    assert rel0.max() < 1e-4


def test_concurrent_simulators_and_pipelined_reductions_do_not_share_scratch():
    """server.py runs simulations on executor worker threads (backend/server.py:309,405): two
    simulators at once, and one simulator whose reductions run on a side stream
    (aggregates_device(pipeline=True)) while a search and a single-path call are enqueued on the
    main one. Scratch is per (context, stream), so every result must equal the serial one."""
    import threading

    import torch

    cfg = dict(scenarios.SYNTH_C3_VOL, num_simulations_search=4000)
    ref = make_sim(cfg)
    want = ref.run_aggregates(100, 60_000)
    ref.use_search_seeds()
    want_counts = ref.batched_success_counts(list(range(0, 120, 6)), 20_000).cpu().tolist()
    ref.use_final_seeds()
    shocks = ref._draw_shock_path(100 + 12 * cfg["retirement_years"], 12345)
    want_path = ref._run_path_on_shocks(100, shocks)

    results, errors = {}, []

    def worker(tag):
        try:
            torch.cuda.set_device(0)
            with torch.cuda.stream(torch.cuda.Stream()):
                sim = make_sim(cfg)
                for rep in range(3):
                    results[(tag, rep)] = sim.run_aggregates(100, 60_000)
        except Exception as exc:  # pragma: no cover
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    # meanwhile, on this thread: pipelined reductions racing a search and a single path
    sim = make_sim(cfg)
    got_counts = got_path = None
    aggs = []
    for rep in range(3):
        aggs.append(sim.aggregates_device(100, 60_000, pipeline=True))
        sim.use_search_seeds()
        got_counts = sim.batched_success_counts(list(range(0, 120, 6)), 20_000).cpu().tolist()
        sim.use_final_seeds()
        got_path = sim._run_path_on_shocks(100, shocks)
    for t in threads:
        t.join()
    assert not errors, errors

    def same(a, b):
        for k in ("trajectory_bands", "real_trajectory_bands", "withdrawal_rate_bands"):
            assert np.array_equal(a[k].to_numpy(), b[k].to_numpy(), equal_nan=True), k
        for k in ("success_count", "median_first_year_withdrawal_rate", "median_start_balance",
                  "median_final_balance_successful", "final_balance_quantiles", "ruin_month_hist",
                  "final_balance_hist_musd_100", "final_balance_hist_60", "withdrawal_rate_counts"):
            assert a[k] == b[k], k

    for agg in aggs:
        same(agg.to_host(), want)
    for key, got in results.items():
        same(got, want)
    assert got_counts == want_counts
    assert got_path.keys() == want_path.keys()
    for k in want_path:   # the withdrawal-rate list of a failing path carries NaN
        assert np.array_equal(np.asarray(got_path[k], dtype=float), np.asarray(want_path[k], dtype=float), equal_nan=True), k


def test_scenario_sweep_equals_one_simulator_per_scenario():
    """Multi-scenario batching: many Configs in one launch per kernel variant, on the
    sweeping simulator's Philox streams == one simulator per scenario with the same seed. The grid
    mixes kernel variants (both taxed / no tax / annual tax), retirement lengths and income streams."""
    from monte_carlo_retirement_b200.config import Config

    base = dict(scenarios.SYNTH_C3, seed=4711)
    grid = []
    for expenses in (8000.0, 12000.0):
        for vol in (0.02, 0.12):
            grid.append(dict(base, monthly_expenses=expenses, inv1_returns_volatility=vol))
    grid.append(dict(scenarios.TEST_BASE, seed=4711, monthly_expenses=4000.0))                  # no tax
    grid.append(dict(scenarios.SYNTH_C3_ANNUAL, seed=4711))                                      # annual tax: generic variant
    grid.append(dict(scenarios.JORGE_PLUS, seed=4711))                                           # four income streams, R = 40
    grid.append(dict(base, retirement_years=25, allocation_inv1_pct=0.3))
    wms = [240, 240, 200, 200, 36, 240, 75, 180]
    n = 30_000
    sweeper = make_sim(base)
    sweeper.use_search_seeds()
    counts, executed = sweeper.sweep_success_counts([Config(**g) for g in grid], wms, n, with_executed=True)
    counts, executed = counts.cpu().tolist(), executed.cpu().tolist()
    for g, wm, c, ex in zip(grid, wms, counts, executed):
        one = make_sim(g)
        one.use_search_seeds()
        want, want_ex = one.batched_success_counts([wm], n, with_executed=True)
        assert (c, ex) == (int(want[0]), int(want_ex[0])), (g["scenario"], wm)
    probs = sweeper.sweep_success_probabilities([Config(**g) for g in grid[:4]], 240, n)
    assert probs[0] >= probs[1] and probs[0] > probs[2]      # same luck everywhere: cheaper living / lower vol never hurts here
    # sharded over two logical ranks: the shards' counts add up
    a = sweeper.sweep_success_counts([Config(**g) for g in grid], wms, 12_000).cpu()
    b = sweeper.sweep_success_counts([Config(**g) for g in grid], wms, n - 12_000, first_path=12_000).cpu()
    assert (a + b).tolist() == counts


def test_chunked_batch_equals_one_launch():
    """run_batch_device(chunks=k): k timeline launches over contiguous path ranges of the same output
    buffers (what lets the host-returning calls overlap the D2H of the summary columns with the
    simulation) == one launch, bit for bit, and on_chunk sees every range once."""
    import torch

    sim = make_sim(scenarios.SYNTH_C3_VOL)
    n, wm = 300_001, 120
    one = sim.run_batch_device(wm, n)
    seen = []
    many = sim.run_batch_device(wm, n, chunks=3, on_chunk=lambda b, lo, cnt: seen.append((lo, cnt)))
    assert len(seen) == 3 and seen[0][0] == 0 and sum(c for _, c in seen) == n
    assert all(a[0] + a[1] == b[0] for a, b in zip(seen, seen[1:])) and all(lo % 1024 == 0 for lo, _ in seen)
    for name in ("cols", "success", "ruin", "counters", "traj", "real"):
        assert torch.equal(getattr(one, name), getattr(many, name)), name
    assert torch.equal(torch.nan_to_num(one.wr, nan=-1.0), torch.nan_to_num(many.wr, nan=-1.0))
    # the 7-tuple through the chunked path equals the unchunked one
    sim.e2e_chunks = 1
    a = sim.run_monte_carlo_simulations(wm, n)
    sim.e2e_chunks = 3
    b = sim.run_monte_carlo_simulations(wm, n)
    assert a[0].equals(b[0]) and np.array_equal(a[1].to_numpy(), b[1].to_numpy()) and a[6] == b[6]
