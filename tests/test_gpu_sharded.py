"""Sharded (multi-GPU) path on the GPU. On a single-GPU box the ranks are emulated logically in
one process (one kernel launch per logical rank, the all-reduce replaced by a tensor sum) — the
profiling guide forbids several NCCL ranks on one GPU. With >= 2 GPUs the real torchrun/NCCL
path is exercised as well."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

import scenarios
from gpu_util import make_sim

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _logical_distributed_quantiles(ctx, shards, rows, q, median=False, masks=None):
    """Drive mcr_select_step for several logical ranks; `all-reduce` = sum of their histograms."""
    import torch

    lib = ctx.lib
    import ctypes as C

    from monte_carlo_retirement_b200 import native

    keep = [ctx.select_rows([(x[r], x.shape[-1], None if masks is None else masks[i], q, median) for r in range(rows)])
            for i, x in enumerate(shards)]
    descs = [C.cast(a.ctypes.data, C.POINTER(native.SelectRow)) for a in keep]
    st = [torch.empty(int(lib.mcr_select_state_bytes(rows)), dtype=torch.uint8, device="cuda") for _ in shards]
    hs = [torch.empty(int(lib.mcr_select_hist_bytes(rows)) // 4, dtype=torch.int32, device="cuda") for _ in shards]
    outs = [torch.empty((rows, 16), dtype=torch.float64, device="cuda") for _ in shards]
    cnts = [torch.empty(rows, dtype=torch.int64, device="cuda") for _ in shards]

    def step(r, kind, p=0):
        rc = lib.mcr_select_step(ctx.handle, kind, p, descs[r], rows, st[r].data_ptr(), hs[r].data_ptr(),
                                 outs[r].data_ptr(), cnts[r].data_ptr(), None)
        assert rc == 0, lib.mcr_last_error(ctx.handle)

    for r in range(len(shards)):
        step(r, 0)
    for p in range(8):
        if p == lib.mcr_select_full_passes():
            for r in range(len(shards)):
                step(r, 4)
        for r in range(len(shards)):
            step(r, 1, p)
        total = sum(hs[1:], hs[0].clone())
        for h in hs:
            h.copy_(total)
        for r in range(len(shards)):
            step(r, 2, p)
    for r in range(len(shards)):
        step(r, 3)
    return [o[:, : len(q)].contiguous() for o in outs], cnts


def _logical_pooled_quantiles(ctx, shards, rows, q, median=False):
    """The pooled-tail protocol of include/mcr.h (MCR_SELECT_POOL_*) with the ranks emulated on one
    GPU: every all-reduce is a torch reduction over the ranks' buffers. Returns the per-rank
    outputs, counts and the number of rows the pool could not finish."""
    import torch

    from monte_carlo_retirement_b200 import native

    lib = ctx.lib
    W = len(shards)
    keep = [ctx.select_rows([(x[r], x.shape[-1], None, q, median) for r in range(rows)]) for x in shards]
    descs = [C.cast(a.ctypes.data, C.POINTER(native.SelectRow)) for a in keep]
    st = [torch.empty(int(lib.mcr_select_state_bytes(rows)), dtype=torch.uint8, device="cuda") for _ in shards]
    hs = [torch.empty(int(lib.mcr_select_hist_bytes(rows)) // 4, dtype=torch.int32, device="cuda") for _ in shards]
    outs = [torch.empty((rows, 16), dtype=torch.float64, device="cuda") for _ in shards]
    cnts = [torch.empty(rows, dtype=torch.int64, device="cuda") for _ in shards]
    at4 = (C.c_int64 * 4)()
    lib.mcr_select_exchange_layout(rows, W, at4)
    at = [int(v) for v in at4]
    assert at[3] == lib.mcr_select_exchange_words(rows, W)
    xb = [torch.empty(at[3], dtype=torch.int64, device="cuda") for _ in shards]
    exts = [torch.empty((rows, 2), dtype=torch.int64, device="cuda") for _ in shards]

    def step(r, kind, p=0, buf=None, hist=None):
        rc = lib.mcr_select_step(ctx.handle, kind, p, descs[r], rows, st[r].data_ptr(),
                                 (hs[r] if hist is None else hist).data_ptr(),
                                 (outs[r] if buf is None else buf).data_ptr(), cnts[r].data_ptr(), None)
        assert rc == 0, lib.mcr_last_error(ctx.handle)

    def all_reduce(bufs, lo, hi, op):
        red = torch.stack([b[lo:hi] for b in bufs])
        red = red.sum(0) if op == "sum" else red.min(0).values
        for b in bufs:
            b[lo:hi] = red

    for r in range(W):
        step(r, 0, 3)
    for p in range(lib.mcr_select_full_passes()):
        for r in range(W):
            step(r, 1, p)
        if p == 0:
            for r in range(W):
                step(r, 5, 0, exts[r])
            low = torch.stack(exts).min(0).values
            for r in range(W):
                exts[r].copy_(low)
                step(r, 6, 0, exts[r])
        all_reduce(hs, 0, hs[0].numel(), "sum")
        for r in range(W):
            step(r, 2, p)
    for r in range(W):
        step(r, 4)
    for r in range(W):
        step(r, 7, r | (W << 8), hist=xb[r])
    all_reduce(xb, at[0], at[1], "sum")
    all_reduce(xb, at[1], at[2], "min")
    for r in range(W):
        step(r, 8, r | (W << 8), hist=xb[r])
    all_reduce(xb, at[2], at[3], "sum")
    for r in range(W):
        step(r, 9, r | (W << 8), hist=xb[r])
    unresolved = [int(b[0].item()) for b in xb]
    assert len(set(unresolved)) == 1  # every rank saw the same pool
    return [o[:, : len(q)].contiguous() for o in outs], cnts, unresolved[0]


@pytest.mark.parametrize("splits", [(1000, 3000), (1, 4095, 2), (2048, 0, 2048), (30000, 50000, 1, 40000)])
def test_pooled_tail_select_equals_global_quantiles(splits):
    rng = np.random.default_rng(sum(splits) + 1)
    rows, n = 6, sum(splits)
    x = np.exp(rng.normal(12, 1.5, (rows, n)))
    x[1, rng.random(n) < 0.3] = 0.0          # zero-padded failures: a big bucket of one value
    x[2, rng.random(n) < 0.5] = np.nan       # withdrawal-rate style rows
    x[3] = np.round(x[3], -3)                # heavy ties
    x[4] = 7.25                              # constant row: done after pass 0
    x[5, : n // 2] = -x[5, : n // 2]         # mixed signs: no common key prefix
    q = [0.05, 0.10, 0.25, 0.50, 0.75, 0.90, 0.95]
    import torch

    sim = make_sim(scenarios.TEST_BASE)
    ctx = sim.native_context
    bounds = np.cumsum((0,) + splits)
    shards = [torch.from_numpy(np.ascontiguousarray(x[:, a:b])).to("cuda") if b > a else
              torch.empty((rows, 0), dtype=torch.float64, device="cuda") for a, b in zip(bounds, bounds[1:])]
    outs, cnts, unresolved = _logical_pooled_quantiles(ctx, shards, rows, q)
    assert unresolved == 0
    want = pd.DataFrame(x.T).quantile(q, axis=0).T.to_numpy()
    for o, c in zip(outs, cnts):
        assert np.array_equal(o.cpu().numpy(), want, equal_nan=True)       # every rank: the exact global answer
        assert c.cpu().tolist() == pd.DataFrame(x.T).count().tolist()
    outs, _, unresolved = _logical_pooled_quantiles(ctx, [s[0:1].contiguous() for s in shards], 1, [0.5], median=True)
    assert unresolved == 0 and outs[0].item() == pd.Series(x[0]).median()


def test_pooled_tail_finishes_rows_with_a_mass_of_zeros():
    """Config #5's band rows (zero-padded failures next to a bulk of balances) over three logical
    ranks: the pooled tail must finish every row — the fallback is for pathological data only."""
    import torch

    from test_gpu_native import _zero_padded_rows

    splits = (70_000, 130_001, 99_999)
    n = sum(splits)
    x = _zero_padded_rows(n, np.random.default_rng(6))
    q = [0.01, 0.05, 0.10, 0.25, 0.50, 0.75, 0.90, 0.95, 0.99]
    sim = make_sim(scenarios.TEST_BASE)
    bounds = np.cumsum((0,) + splits)
    shards = [torch.from_numpy(np.ascontiguousarray(x[:, a:b])).to("cuda") for a, b in zip(bounds, bounds[1:])]
    outs, cnts, unresolved = _logical_pooled_quantiles(sim.native_context, shards, x.shape[0], q)
    assert unresolved == 0
    want = np.stack([pd.Series(r).quantile(q).to_numpy() for r in x])
    for o in outs:
        assert np.array_equal(o.cpu().numpy(), want, equal_nan=True)


def test_pooled_tail_resolves_requested_minima_and_maxima():
    """MCR_SEL_MINMAX rows (the histogram ranges of a step) over three logical ranks whose shards are long
    enough to be SAMPLED for the first digit window: the true extremes sit where no sample looks, the
    ranks exchange the extreme keys they saw outside the window with the candidate pool, and every rank
    ends with the exact global minimum and maximum — no fallback."""
    import torch

    splits = (200_000, 150_001, 180_000)
    n = sum(splits)
    rng = np.random.default_rng(12)
    x = np.exp(rng.normal(15, 0.7, (4, n)))
    x[0, 30_000], x[0, 250_000] = 3.5, 9.9e11            # far outside the bulk, in pieces the sample skips
    x[1, 20_000 + np.arange(5)] = -np.arange(1.0, 6.0)   # negative minimum next to a mass of zeros
    x[1, rng.random(n) < 0.2] = 0.0
    x[2] = 4.0                                           # constant row
    x[3, rng.random(n) < 0.4] = np.nan                   # NaN are skipped
    x[3, 401_000] = 1e-9
    sim = make_sim(scenarios.TEST_BASE)
    bounds = np.cumsum((0,) + splits)
    shards = [torch.from_numpy(np.ascontiguousarray(x[:, a:b])).to("cuda") for a, b in zip(bounds, bounds[1:])]
    outs, cnts, unresolved = _logical_pooled_quantiles(sim.native_context, shards, 4, [0.0, 1.0], median="minmax")
    assert unresolved == 0
    want = np.stack([np.nanmin(x, axis=1), np.nanmax(x, axis=1)], axis=1)
    for o in outs:
        assert np.array_equal(o.cpu().numpy(), want)


def test_pooled_tail_reports_rows_it_cannot_finish():
    """A dense cluster of distinct values inside a wide key range stays too big for the pool after
    all full passes: every rank reports the row, the caller falls back to the stepwise protocol."""
    import torch

    n = 24000
    x = np.empty((2, n))
    x[0] = 1.0 + np.arange(n) * 2.0 ** -50   # 24000 distinct values sharing 40+ leading key bits ...
    x[0, 0], x[0, 1] = 0.0, 1e300            # ... in a row whose extremes share none
    x[1] = np.linspace(1.0, 2.0, n)          # an ordinary row next to it
    q = [0.25, 0.5, 0.75]
    sim = make_sim(scenarios.TEST_BASE)
    ctx = sim.native_context
    shards = [torch.from_numpy(np.ascontiguousarray(x[:, a:b])).to("cuda") for a, b in ((0, 10000), (10000, n))]
    outs, _, unresolved = _logical_pooled_quantiles(ctx, shards, 2, q)
    assert unresolved == 1
    want = pd.DataFrame(x.T).quantile(q, axis=0).T.to_numpy()
    assert np.array_equal(outs[0][1].cpu().numpy(), want[1])               # the finished row is still exact
    legacy, _ = _logical_distributed_quantiles(ctx, shards, 2, q)          # and the fallback gets both
    assert np.array_equal(legacy[0].cpu().numpy(), want)
    single = torch.empty((2, 3), dtype=torch.float64, device="cuda")       # as does the single-GPU slow path
    whole = torch.from_numpy(x).to("cuda")
    ctx.quantiles(whole, n, n, 2, q, single)
    assert np.array_equal(single.cpu().numpy(), want)


@pytest.mark.parametrize("splits", [(1000, 3000), (1, 4095, 2), (2048, 0, 2048)])
def test_distributed_select_equals_global_quantiles(splits):
    import torch

    rng = np.random.default_rng(sum(splits))
    rows, n = 5, sum(splits)
    x = np.exp(rng.normal(12, 1.5, (rows, n)))
    x[1, rng.random(n) < 0.3] = 0.0
    x[2, rng.random(n) < 0.5] = np.nan
    x[3] = np.round(x[3], -3)
    q = [0.05, 0.10, 0.25, 0.50, 0.75, 0.90, 0.95]
    sim = make_sim(scenarios.TEST_BASE)
    ctx = sim.native_context
    bounds = np.cumsum((0,) + splits)
    shards = [torch.from_numpy(np.ascontiguousarray(x[:, a:b])).to("cuda") if b > a else
              torch.empty((rows, 0), dtype=torch.float64, device="cuda") for a, b in zip(bounds, bounds[1:])]
    outs, cnts = _logical_distributed_quantiles(ctx, shards, rows, q)
    want = pd.DataFrame(x.T).quantile(q, axis=0).T.to_numpy()
    for o, c in zip(outs, cnts):
        assert np.array_equal(o.cpu().numpy(), want, equal_nan=True)       # every rank: the exact global answer
        assert c.cpu().tolist() == pd.DataFrame(x.T).count().tolist()
    outs, _ = _logical_distributed_quantiles(ctx, [s[0:1].contiguous() for s in shards], 1, [0.5], median=True)
    assert outs[0].item() == pd.Series(x[0]).median()


def test_logical_shards_reproduce_single_gpu_aggregates():
    """Two logical ranks (disjoint global path ranges of one Philox stream) == one rank."""
    import torch

    sim = make_sim(scenarios.STRESSED)
    n, wm = 6000, 100
    whole = sim.run_batch_device(wm, n)
    a = sim.run_batch_device(wm, 2500, first_path=0)
    b = sim.run_batch_device(wm, 3500, first_path=2500)
    assert torch.equal(torch.cat([a.cols, b.cols], 1), whole.cols)
    assert torch.equal(a.counters + b.counters, whole.counters)
    q = [0.05, 0.10, 0.25, 0.50, 0.75, 0.90, 0.95]
    ref = torch.empty((whole.T, 7), dtype=torch.float64, device="cuda")
    sim.native_context.quantiles(whole.traj, n, n, whole.T, q, ref)
    outs, _ = _logical_distributed_quantiles(sim.native_context, [a.traj, b.traj], whole.T, q)
    assert torch.equal(outs[0], ref) and torch.equal(outs[1], ref)


def test_two_rank_nccl_run_matches_single_gpu():
    """Real torchrun x2 over NCCL (needs 2 GPUs): sharded aggregates, 7-tuple and search agree
    with the single-GPU engine."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(ROOT, "tools", "sharded_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SHARDED CHECK OK" in r.stdout
