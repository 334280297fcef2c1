"""Path sharding over the GPUs of one box: one process per GPU (torchrun), torch.distributed
(NCCL over NVLink / NVSwitch) for the plumbing.

The hot path shards by construction — paths never interact (backend/simulation.py:987-990),
only aggregates couple them — so:

  * rank r of W owns the contiguous GLOBAL path range [r*N//W, (r+1)*N//W); Philox counters are
    global path indices, so any path's result is independent of W (tested bit-for-bit);
  * there is NO data-path collective. What crosses NVLink is tiny and latency-bound:
      - one all-reduce(sum) of the int64 counter block (success count, executed months,
        ruin-month histogram) and of the per-candidate success counts of the batched search;
      - min/max then histogram all-reduces for the final-balance histograms;
      - the <= 4 x [rows x 32 x 256] int32 digit histograms of the distributed radix select and one
        pooled block of the few hundred candidates per row that are left after them, which give
        EXACT global quantile bands without moving any per-path data;
      - (only when the caller asks for the reference's N-row summary_df) an all-gather of the
        7 summary columns.

The reference has no distributed layer at all (its only parallelism is multiprocessing.Pool
over paths, simulation.py:996-1001); this module is new functionality behind the same class
surface: `ShardedSimulator` IS a RetirementMonteCarloSimulator whose every rank returns the
same global answers.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import pandas as pd

from . import native
from .simulation import (FINAL_BALANCE_QUANTILES, TRAJECTORY_QUANTILES, WITHDRAWAL_RATE_QUANTILES,
                         RetirementMonteCarloSimulator)


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[first, first + count) of the global path range owned by `rank` (contiguous, balanced)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi - lo


class Collectives:
    """The few collectives the sharded engine needs, over the default process group."""

    def __init__(self, group=None):
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def sum_(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def min_(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        return t

    def max_(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return t

    def gather_cat(self, t, dim: int = -1, sizes: Optional[Sequence[int]] = None):
        """Concatenate the ranks' shards along `dim`. `sizes` (every rank's extent along dim) avoids
        an object all-gather when the caller can compute it (shard_range)."""
        import torch

        if sizes is None:
            sizes = [None] * self.world
            self.dist.all_gather_object(sizes, int(t.shape[dim]), group=self.group)
        m = max(sizes)
        if all(s == m for s in sizes):
            parts = [torch.empty_like(t) for _ in range(self.world)]
            self.dist.all_gather(parts, t.contiguous(), group=self.group)
            return torch.cat(parts, dim=dim)
        pad_shape = list(t.shape)
        pad_shape[dim] = m
        padded = torch.zeros(pad_shape, dtype=t.dtype, device=t.device)
        padded.narrow(dim, 0, t.shape[dim]).copy_(t)
        bufs = [torch.empty_like(padded) for _ in range(self.world)]
        self.dist.all_gather(bufs, padded, group=self.group)
        return torch.cat([b.narrow(dim, 0, s) for b, s in zip(bufs, sizes)], dim=dim)


class ShardedSimulator(RetirementMonteCarloSimulator):
    """RetirementMonteCarloSimulator over W ranks; every rank gets the global result."""

    _device_search_ok = True  # the class-level overrides below keep the device search path

    def __init__(self, params_model, main_seed_override: Optional[int] = None, *, collectives: Optional[Collectives] = None,
                 **kw):
        super().__init__(params_model, main_seed_override, **kw)
        self.coll = collectives or Collectives()
        self._agreed: Dict[Any, int] = {}
        self._select_flag = None
        self._stepwise_only = False
        self.select_fallbacks = 0
        self._sign4 = None
        if self.rng_mode != "philox":
            raise ValueError("sharding needs the counter-based Philox draws (rng='philox')")

    # ---- search: per-candidate counts are summed over the shards ------------------------------
    def batched_success_counts(self, candidates: Sequence[int], num_simulations: int, *, first_path: int = 0,
                               with_executed: bool = False):
        lo, cnt = shard_range(int(num_simulations), self.coll.rank, self.coll.world)
        return super().batched_success_counts(candidates, cnt, first_path=first_path + lo, with_executed=with_executed)

    def _reduce_counts(self, counts):
        return self.coll.sum_(counts)

    # ---- hooks: the base class's aggregate-only path runs unchanged on the local shard, with the
    # ---- few collectives injected here ------------------------------------------------------------
    def _shard(self, n_global: int) -> Tuple[int, int]:
        return shard_range(int(n_global), self.coll.rank, self.coll.world)

    def _agree_min(self, value: int, key=None) -> int:
        import torch

        if key is not None and key in self._agreed:
            return self._agreed[key]
        t = torch.tensor([int(value)], dtype=torch.int64, device=self._torch_device())
        out = int(self.coll.min_(t).item())
        if key is not None:
            self._agreed[key] = out
        return out

    def _select(self, specs, out16, counts=None) -> None:
        # exact GLOBAL order statistics: local digit histograms all-reduced per pass, then the few
        # candidates left are pooled across ranks (the "could not finish" count is checked at the
        # caller's next host sync: _selects_ok)
        ctx, coll = self.native_context, self.coll
        if self._stepwise_only:
            ctx.quantiles_rows(specs, out16, counts=counts, all_reduce=coll.sum_, all_reduce_min=coll.min_)
            return
        flag = ctx.quantiles_rows(specs, out16, counts=counts, all_reduce=coll.sum_, all_reduce_min=coll.min_,
                                  rank=coll.rank, world=coll.world, defer_check=True)
        if flag is not None:
            import torch

            self._select_flag = flag if self._select_flag is None else torch.maximum(self._select_flag, flag)

    def _selects_ok(self) -> bool:
        flag, self._select_flag = self._select_flag, None
        ok = flag is None or int(flag.item()) == 0
        if not ok:
            self.select_fallbacks += 1  # reported by bench.py: a fallback re-runs work outside a timed loop
        return ok

    def _stepwise_selects(self):
        import contextlib

        @contextlib.contextmanager
        def forced():
            before, self._stepwise_only = self._stepwise_only, True
            try:
                yield
            finally:
                self._stepwise_only = before

        return forced()

    def _reduce_samples(self, block):
        return self.coll.sum_(block)

    def _final_balance_histograms(self, b, rng_m, rng_1, hists) -> None:
        import torch

        ctx, coll, n = self.native_context, self.coll, b.n
        ctx.minmax(b.cols[1], n, rng_m, mask=b.success, divisor=1e6)
        ctx.minmax(b.cols[1], n, rng_1, mask=b.success, divisor=1.0)
        # global ranges in ONE all-reduce(MIN) of [lo_m, -hi_m, lo_1, -hi_1]; NaN (empty local
        # cohort) must not poison it
        if self._sign4 is None:
            self._sign4 = torch.tensor([1.0, -1.0, 1.0, -1.0], dtype=torch.float64, device=rng_m.device)
        sign = self._sign4
        packed = torch.nan_to_num(torch.cat([rng_m, rng_1]) * sign, nan=float("inf"))
        coll.min_(packed)
        packed = torch.where(torch.isinf(packed), float("nan"), packed * sign)
        rng_m.copy_(packed[0:2])
        rng_1.copy_(packed[2:4])
        ctx.histogram(b.cols[1], n, 100, rng_m, hists[0:], mask=b.success, divisor=1e6, mode=native.HIST_NUMPY)
        ctx.histogram(b.cols[1], n, 60, rng_1, hists[100:], mask=b.success, divisor=1.0, mode=native.HIST_FLOOR)
        coll.sum_(hists)

    # ---- the reference's 7-tuple, global, on every rank ------------------------------------------
    def run_monte_carlo_simulations(self, working_months: int, num_simulations: int):
        import torch

        coll = self.coll
        n_global = int(num_simulations)
        lo, n = shard_range(n_global, coll.rank, coll.world)
        b = self.run_batch_device(working_months, n, series=True, first_path=lo)
        dev = b.cols.device
        T, R = b.T, b.R
        nq, nw = len(TRAJECTORY_QUANTILES), len(WITHDRAWAL_RATE_QUANTILES)
        f64 = dict(dtype=torch.float64, device=dev)
        bands = torch.empty((T, nq), **f64)
        real_bands = torch.empty((T, nq), **f64)
        wr_bands = torch.empty((R, nw), **f64)
        wr_counts = torch.empty(R, dtype=torch.int64, device=dev)
        self._band_quantiles(b, bands, real_bands, wr_bands, wr_counts)  # one distributed multi-row select
        # sample paths: each rank contributes the columns it owns, summed into a zero block
        k = len(self._sample_columns(n_global))
        samples = torch.zeros((2, k, T), **f64)
        self._gather_samples(b.traj, n, T, lo, n_global, samples[0])
        self._gather_samples(b.real, n, T, lo, n_global, samples[1])
        self._reduce_samples(samples)
        # summary columns of all shards, in global path order: ONE all-gather of a packed
        # [5 f64 | success | ruin] block per rank over NVLink, one pinned D2H, zero-copy DataFrame
        sizes = [shard_range(n_global, r, coll.world)[1] for r in range(coll.world)]
        m = max(sizes)
        packed = torch.zeros((7, m), **f64)
        packed[0:5, :n] = b.cols
        packed[5, :n] = b.success.to(torch.float64)
        b.years_to_ruin_into(packed[6, :n])
        gathered = torch.empty((coll.world, 7, m), **f64)
        coll.dist.all_gather_into_tensor(gathered, packed, group=coll.group)
        host = torch.empty((7, n_global), dtype=torch.float64, pin_memory=True)
        if all(sz == m for sz in sizes):
            host.copy_(gathered.permute(1, 0, 2).reshape(7, -1), non_blocking=True)
        else:
            host.copy_(torch.cat([gathered[r, :, :sz] for r, sz in enumerate(sizes)], dim=1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if not self._selects_ok():  # rare: the pooled select gave up on a row -> stepwise protocol
            with self._stepwise_selects():
                self._band_quantiles(b, bands, real_bands, wr_bands, wr_counts)
        all_cols = host.numpy()
        self.last_d2h_bytes = n_global * 7 * 8 + (2 * T * nq + R * nw + 2 * k * T + R) * 8
        summary_df = pd.DataFrame({
            "Start Balance": all_cols[0],
            "Final Balance": all_cols[1],
            "Success": all_cols[5] != 0.0,
            "YearsToRuin": all_cols[6],
            "First Year Gross Withdrawal": all_cols[2],
            "First Year Real Gross Withdrawal": all_cols[3],
            "Inflation At Retirement": all_cols[4],
        }, copy=False)
        s = samples.cpu().numpy()
        self._last_batch = b
        return (summary_df,
                pd.DataFrame(bands.cpu().numpy(), columns=TRAJECTORY_QUANTILES),
                s[0].tolist(),
                pd.DataFrame(wr_bands.cpu().numpy(), columns=WITHDRAWAL_RATE_QUANTILES),
                pd.DataFrame(real_bands.cpu().numpy(), columns=TRAJECTORY_QUANTILES),
                s[1].tolist(),
                [int(v) for v in wr_counts.cpu().numpy()])
