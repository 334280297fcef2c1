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
