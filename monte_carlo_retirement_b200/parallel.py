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
      - the 8 x [rows x 32 x 256] int32 digit histograms of the distributed radix select, which
        give EXACT global quantile bands without moving any per-path data;
      - (only when the caller asks for the reference's N-row summary_df) an all-gather of the
        7 summary columns.

The reference has no distributed layer at all (its only parallelism is multiprocessing.Pool
over paths, simulation.py:996-1001); this module is new functionality behind the same class
surface: `ShardedSimulator` IS a RetirementMonteCarloSimulator whose every rank returns the
same global answers.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

from . import native
from .constants import MONTHS_PER_YEAR
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
        if self.rng_mode != "philox":
            raise ValueError("sharding needs the counter-based Philox draws (rng='philox')")

    # ---- search: per-candidate counts are summed over the shards ------------------------------
    def batched_success_counts(self, candidates: Sequence[int], num_simulations: int, *, first_path: int = 0,
                               with_executed: bool = False):
        lo, cnt = shard_range(int(num_simulations), self.coll.rank, self.coll.world)
        return super().batched_success_counts(candidates, cnt, first_path=first_path + lo, with_executed=with_executed)

    def _reduce_counts(self, counts):
        return self.coll.sum_(counts)

    # ---- aggregate-only mode over the global path set ------------------------------------------
    def aggregates_device(self, working_months: int, num_simulations: int, *, bands: bool = True,
                          first_path: int = 0, timeline_events=None):
        import torch

        ctx = self.native_context
        coll = self.coll
        import os

        n_global = int(num_simulations)
        lo, n = shard_range(n_global, coll.rank, coll.world)
        series_bytes = 8 * n * (2 * self._trajectory_len(int(working_months)) + self.params_model.retirement_years)
        free_bytes, _ = torch.cuda.mem_get_info(self._torch_device())
        sweep_t = torch.tensor([int(bool(bands) and (series_bytes > 0.6 * free_bytes
                                                       or os.environ.get("MCR_SERIES_SWEEP") == "1"))],
                               dtype=torch.int32, device=self._torch_device())
        sweep = bool(coll.max_(sweep_t).item())  # every rank takes the same route
        if timeline_events is not None:
            timeline_events[0].record()
        b = self.run_batch_device(working_months, n, series=(bands and not sweep), first_path=first_path + lo)
        if timeline_events is not None:
            timeline_events[1].record()
        dev = b.cols.device
        T, R = b.T, b.R
        f64 = dict(dtype=torch.float64, device=dev)
        nq, nw, nf = len(TRAJECTORY_QUANTILES), len(WITHDRAWAL_RATE_QUANTILES), len(FINAL_BALANCE_QUANTILES)
        coll.sum_(b.counters)
        rates = torch.empty(n, **f64)
        ctx.first_year_rates(b.cols[0], b.cols[3], n, rates)
        small = torch.empty(3 + nf + 4, **f64)
        cnt = torch.empty(3, dtype=torch.int64, device=dev)
        ar = coll.sum_
        ctx.quantiles_distributed(rates, n, n, 1, [0.5], small[0:], ar, counts=cnt[0:], median=True)
        ctx.quantiles_distributed(b.cols[0], n, n, 1, [0.5], small[1:], ar, counts=cnt[1:], median=True)
        ctx.quantiles_distributed(b.cols[1], n, n, 1, [0.5], small[2:], ar, counts=cnt[2:], mask=b.success, median=True)
        ctx.quantiles_distributed(b.cols[1], n, n, 1, FINAL_BALANCE_QUANTILES, small[3:], ar)
        rng_m = small[3 + nf:3 + nf + 2]
        rng_1 = small[3 + nf + 2:3 + nf + 4]
        hists = torch.zeros(160, dtype=torch.int64, device=dev)
        for rng, divisor in ((rng_m, 1e6), (rng_1, 1.0)):
            ctx.minmax(b.cols[1], n, rng, mask=b.success, divisor=divisor)
            # global range: NaN (empty local cohort) must not poison min/max
            lo_v = torch.nan_to_num(rng[0:1], nan=float("inf"))
            hi_v = torch.nan_to_num(rng[1:2], nan=float("-inf"))
            coll.min_(lo_v)
            coll.max_(hi_v)
            empty = torch.isinf(lo_v)
            rng[0:1] = torch.where(empty, torch.full_like(lo_v, float("nan")), lo_v)
            rng[1:2] = torch.where(empty, torch.full_like(hi_v, float("nan")), hi_v)
        ctx.histogram(b.cols[1], n, 100, rng_m, hists[0:], mask=b.success, divisor=1e6, mode=native.HIST_NUMPY)
        ctx.histogram(b.cols[1], n, 60, rng_1, hists[100:], mask=b.success, divisor=1.0, mode=native.HIST_FLOOR)
        coll.sum_(hists)
        band_block = wr_counts = None
        if bands:
            band_block = torch.empty(2 * T * nq + R * nw, **f64)
            wr_counts = torch.empty(R, dtype=torch.int64, device=dev)
            if not sweep:
                ctx.quantiles_distributed(b.traj, n, n, T, TRAJECTORY_QUANTILES, band_block[0:], ar)
                ctx.quantiles_distributed(b.real, n, n, T, TRAJECTORY_QUANTILES, band_block[T * nq:], ar)
                ctx.quantiles_distributed(b.wr, n, n, R, WITHDRAWAL_RATE_QUANTILES, band_block[2 * T * nq:], ar,
                                          counts=wr_counts)
            else:  # one series at a time (recompute): 1e9-path jobs, see simulation.aggregates_device
                for which, rows, qs, off in (("traj", T, TRAJECTORY_QUANTILES, 0),
                                             ("real", T, TRAJECTORY_QUANTILES, T * nq),
                                             ("wr", R, WITHDRAWAL_RATE_QUANTILES, 2 * T * nq)):
                    part = self.run_batch_device(working_months, n, series=which, first_path=first_path + lo)
                    ctx.quantiles_distributed(getattr(part, which), n, n, rows, qs, band_block[off:], ar,
                                              counts=wr_counts if which == "wr" else None)
                    torch.cuda.current_stream().synchronize()
                    del part
        self._last_batch = b
        from .simulation import DeviceAggregates

        b.n_global = n_global
        agg = DeviceAggregates(batch=b, small=small, counts=cnt, hists=hists, band_block=band_block,
                               wr_counts=wr_counts, rates=rates)
        agg.n_override = n_global
        return agg

    def run_aggregates(self, working_months: int, num_simulations: int, *, bands: bool = True,
                       first_path: int = 0) -> Dict[str, Any]:
        agg = self.aggregates_device(working_months, num_simulations, bands=bands, first_path=first_path)
        out = agg.to_host()
        n = int(num_simulations)
        out["num_simulations"] = n
        out["success_probability"] = float(out["success_count"] / n * 100.0)
        return out

    # ---- the reference's 7-tuple, global, on every rank ------------------------------------------
    def run_monte_carlo_simulations(self, working_months: int, num_simulations: int):
        import torch

        ctx = self.native_context
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
        ar = coll.sum_
        ctx.quantiles_distributed(b.traj, n, n, T, TRAJECTORY_QUANTILES, bands, ar)
        ctx.quantiles_distributed(b.real, n, n, T, TRAJECTORY_QUANTILES, real_bands, ar)
        ctx.quantiles_distributed(b.wr, n, n, R, WITHDRAWAL_RATE_QUANTILES, wr_bands, ar, counts=wr_counts)
        # sample paths: each rank contributes the columns it owns, summed into a zero block
        cols = self._sample_columns(n_global)
        k = len(cols)
        samples = torch.zeros((2, k, T), **f64)
        mine = [(j, c - lo) for j, c in enumerate(cols) if lo <= c < lo + n]
        if mine:
            tmp = torch.empty((len(mine), T), **f64)
            ctx.gather_columns(b.traj, n, T, [c for _, c in mine], tmp)
            samples[0, [j for j, _ in mine]] = tmp
            ctx.gather_columns(b.real, n, T, [c for _, c in mine], tmp)
            samples[1, [j for j, _ in mine]] = tmp
        coll.sum_(samples)
        # summary columns of all shards, in global path order
        sizes = [shard_range(n_global, r, coll.world)[1] for r in range(coll.world)]
        all_cols = coll.gather_cat(b.cols, dim=1, sizes=sizes).cpu().numpy()
        all_succ = coll.gather_cat(b.success, dim=0, sizes=sizes).cpu().numpy().astype(bool)
        all_ruin = coll.gather_cat(b.ruin, dim=0, sizes=sizes).cpu().numpy()
        self.last_d2h_bytes = n_global * (5 * 8 + 1 + 4) + (2 * T * nq + R * nw + 2 * k * T + R) * 8
        summary_df = pd.DataFrame({
            "Start Balance": all_cols[0],
            "Final Balance": all_cols[1],
            "Success": all_succ,
            "YearsToRuin": np.where(all_ruin < 0, np.nan, all_ruin.astype(np.float64) / MONTHS_PER_YEAR),
            "First Year Gross Withdrawal": all_cols[2],
            "First Year Real Gross Withdrawal": all_cols[3],
            "Inflation At Retirement": all_cols[4],
        })
        s = samples.cpu().numpy()
        self._last_batch = b
        return (summary_df,
                pd.DataFrame(bands.cpu().numpy(), columns=TRAJECTORY_QUANTILES),
                s[0].tolist(),
                pd.DataFrame(wr_bands.cpu().numpy(), columns=WITHDRAWAL_RATE_QUANTILES),
                pd.DataFrame(real_bands.cpu().numpy(), columns=TRAJECTORY_QUANTILES),
                s[1].tolist(),
                [int(v) for v in wr_counts.cpu().numpy()])
