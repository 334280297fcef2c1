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

    def barrier(self) -> None:
        """All ranks have reached this point AND their previously enqueued device work is done."""
        import torch

        if torch.cuda.is_available() and self.dist.get_backend(self.group) == "nccl":
            torch.cuda.current_stream().synchronize()
        self.dist.barrier(group=self.group)

    def broadcast0_(self, t):
        self.dist.broadcast(t, src=0, group=self.group)
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
        self._shared_blocks: List["_SharedSummaryBlock"] = []   # rank-0-visible summary blocks (host, pinned)
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

    def _select(self, specs, out16, counts=None, stepwise: bool = False):
        # exact GLOBAL order statistics: local digit histograms all-reduced per pass, then the few
        # candidates left are pooled across ranks. The pooled tail reports rows it could not finish
        # in a device flag (identical on every rank) that the caller checks at its next host sync
        # and answers with `stepwise=True` on the still-resident rows.
        ctx, coll = self.native_context, self.coll
        if stepwise:
            ctx.quantiles_rows(specs, out16, counts=counts, all_reduce=coll.sum_, all_reduce_min=coll.min_)
            return None
        return ctx.quantiles_rows(specs, out16, counts=counts, all_reduce=coll.sum_, all_reduce_min=coll.min_,
                                  rank=coll.rank, world=coll.world, defer_check=True)

    def _reduce_samples(self, block):
        return self.coll.sum_(block)

    def _final_balance_histograms(self, b, rng_raw, hists) -> None:
        # (the range is GLOBAL already: min / max of the cohort come out of the distributed select)
        super()._final_balance_histograms(b, rng_raw, hists)
        self.coll.sum_(hists)

    # ---- the reference's 7-tuple ------------------------------------------------------------------
    def run_monte_carlo_simulations(self, working_months: int, num_simulations: int):
        """simulation.py:952-1128 over W ranks. Every rank returns the GLOBAL band frames, sample
        paths and observation counts. `summary_df` (N rows x 7 columns, the only N-sized output):
        rank 0 gets all N_global rows, the other ranks the rows of their own shard — all of them
        zero-copy views of one host block, so a shard crosses PCIe exactly once.

        No per-path data crosses NVLink: every GPU copies ITS OWN shard of the seven columns over
        ITS OWN PCIe link — underneath the select kernels — into one host block that all ranks of
        the box map (POSIX shared memory, page-locked in every process), at the offset of its global
        path range; rank 0's DataFrame is a zero-copy view of that block. (Round 1 all-gathered the
        columns to every GPU and copied all N rows to the host on every rank: 8 x 448 MB per call
        at 8 GPUs, 55.7 ms; this is 8 x 56 MB in parallel.)"""
        import torch

        coll = self.coll
        n_global = int(num_simulations)
        if n_global <= 0:
            return super().run_monte_carlo_simulations(working_months, n_global)
        lo, n = shard_range(n_global, coll.rank, coll.world)
        dev = self._torch_device()
        f64 = dict(dtype=torch.float64, device=dev)
        nq, nw = len(TRAJECTORY_QUANTILES), len(WITHDRAWAL_RATE_QUANTILES)

        # ---- the shard's columns start their trip to the host chunk by chunk, as soon as a chunk's
        # ---- timeline launch is done (underneath the next chunk, then underneath the selects)
        main = torch.cuda.current_stream()
        years = torch.empty(n, **f64)
        block = self._summary_block(n_global)          # collective: a generation no rank still views
        copy_stream = self._copy_stream

        def copy_out(b, at, cnt):
            b.years_to_ruin_slice(years, at, cnt)
            produced = torch.cuda.Event()
            produced.record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(produced)
                for c in range(5):
                    block.cols[c, lo + at:lo + at + cnt].copy_(b.cols[c, at:at + cnt], non_blocking=True)
                block.cols[5, lo + at:lo + at + cnt].copy_(years[at:at + cnt], non_blocking=True)
                block.succ[lo + at:lo + at + cnt].copy_(b.success[at:at + cnt], non_blocking=True)

        b = self.run_batch_device(working_months, n, series=True, first_path=lo, on_chunk=copy_out,
                                  chunks=self.e2e_chunks)
        T, R = b.T, b.R
        years.record_stream(copy_stream)

        # ---- global bands (one distributed multi-row select) and the sample paths, in ONE small block:
        # ---- [bands | real bands | WR bands | samples (2 x k x T) | WR counts | select flag]
        k = len(self._sample_columns(n_global))
        sizes = (T * nq, T * nq, R * nw, 2 * k * T, R, 1)
        small = torch.empty(sum(sizes), **f64)
        at, parts = 0, []
        for m in sizes:
            parts.append(small[at:at + m])
            at += m
        bands, real_bands, wr_bands = parts[0].view(T, nq), parts[1].view(T, nq), parts[2].view(R, nw)
        samples = parts[3].view(2, k, T)
        wr_counts = torch.empty(R, dtype=torch.int64, device=dev)
        flag = self._band_quantiles(b, bands, real_bands, wr_bands, wr_counts)
        parts[3].zero_()                                # each rank fills the columns it owns, summed into zeros
        self._gather_samples(b.traj, n, T, lo, n_global, samples[0])
        self._gather_samples(b.real, n, T, lo, n_global, samples[1])
        self._reduce_samples(parts[3])
        parts[4].copy_(wr_counts)                       # (exact in float64)
        if flag is None:
            parts[5].zero_()
        else:
            parts[5].copy_(flag)
        host_small = torch.empty(small.numel(), dtype=torch.float64, pin_memory=True)
        host_small.copy_(small, non_blocking=True)
        done = torch.cuda.Event()
        done.record(main)
        self.last_d2h_bytes = n * (6 * 8 + 1) + (small.numel() - 1) * 8
        # The frames wrap host memory the GPUs are still filling; they are built while the GPUs work and handed
        # out after every shard has landed. Views of the block keep its generation busy on this rank
        # (is_free): rank 0 all rows, the others the rows of their own shard.
        if coll.rank == 0:
            c, succ = block.np_cols, block.np_succ
        else:
            c, succ = block.np_cols[:, lo:lo + n], block.np_succ[lo:lo + n]
        summary_df = pd.DataFrame({
            "Start Balance": c[0],
            "Final Balance": c[1],
            "Success": succ.view("bool"),                # 0/1 bytes written by the kernel
            "YearsToRuin": c[5],
            "First Year Gross Withdrawal": c[2],
            "First Year Real Gross Withdrawal": c[3],
            "Inflation At Retirement": c[4],
        }, copy=False)
        h = host_small.numpy()
        o = [0]
        for m in sizes:
            o.append(o[-1] + m)
        traj_pct = pd.DataFrame(h[o[0]:o[1]].reshape(T, nq), columns=TRAJECTORY_QUANTILES, copy=False)
        real_pct = pd.DataFrame(h[o[1]:o[2]].reshape(T, nq), columns=TRAJECTORY_QUANTILES, copy=False)
        wr_pct = pd.DataFrame(h[o[2]:o[3]].reshape(R, nw), columns=WITHDRAWAL_RATE_QUANTILES, copy=False)
        done.synchronize()
        if h[o[5]] != 0.0:   # rare: the pooled select gave up on a row -> stepwise protocol, block copied again
            self.select_fallbacks += 1
            self._band_quantiles(b, bands, real_bands, wr_bands, wr_counts, stepwise=True)
            parts[4].copy_(wr_counts)
            host_small.copy_(small)
        copy_stream.synchronize()
        coll.barrier()                                  # every shard has landed in the shared block
        s2 = h[o[3]:o[4]].reshape(2, k, T)
        self._last_batch = b
        return (summary_df, traj_pct, s2[0].tolist(), wr_pct, real_pct, s2[1].tolist(), [int(v) for v in h[o[4]:o[5]]])

    def _staging(self, n: int, dev):
        if getattr(self, "_copy_stream", None) is None:
            import torch

            self._copy_stream = torch.cuda.Stream(device=dev)
        return super()._staging(n, dev)

    def _summary_block(self, n_global: int) -> "_SharedSummaryBlock":
        """The host block of this call: a generation of the right size that no live DataFrame of ANY
        rank still views — every rank reports which of its generations are free, one tiny
        all-reduce(MIN) of that mask picks the first common one; a new generation is created
        (shared memory + page-locking, ~0.1 s, once) when none is free."""
        import torch

        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self._torch_device())
        coll = self.coll
        slots = 16
        free = torch.zeros(slots, dtype=torch.int64)
        for i, blk in enumerate(self._shared_blocks[:slots]):
            free[i] = 1 if (blk.n == n_global and blk.is_free()) else 0
        free = coll.min_(free.to(self._torch_device())).cpu()
        hits = torch.nonzero(free).flatten().tolist()
        idx = hits[0] if hits else len(self._shared_blocks)
        if idx >= slots:   # every generation busy: results are being hoarded; recycle the oldest size class
            raise RuntimeError("more than 16 live results of the sharded run_monte_carlo_simulations")
        if idx == len(self._shared_blocks):
            self._shared_blocks.append(_SharedSummaryBlock(coll, self._block_tag(), idx, n_global))
        return self._shared_blocks[idx]

    def _block_tag(self) -> str:
        import os

        return f"mcr_b200_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}_{id(self) if self.coll.world == 1 else 0}"


class _SharedSummaryBlock:
    """[6][N] float64 + [N] uint8 in POSIX shared memory, mapped and page-locked by every rank of the
    box (cudaHostRegister), so each GPU's copy engine writes its shard straight into the block rank
    0 reads. Created by rank 0, opened by the others after a barrier, unlinked at once (the
    mapping keeps it alive)."""

    def __init__(self, coll: Collectives, tag: str, gen: int, n: int):
        import mmap
        import os
        import sys

        import numpy as np
        import torch

        self.n = n
        self.coll = coll
        if n == 0:          # placeholder generation of another size class
            self.cols = self.succ = None
            return
        path = f"/dev/shm/{tag}_{gen}_{n}"
        size = 6 * 8 * n + n
        size = (size + 4095) // 4096 * 4096
        if coll.rank == 0:
            fd = os.open(path, os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
            os.ftruncate(fd, size)
        coll.barrier()
        if coll.rank != 0:
            fd = os.open(path, os.O_RDWR)
        self._map = mmap.mmap(fd, size)
        os.close(fd)
        coll.barrier()
        if coll.rank == 0:
            os.unlink(path)
        # every numpy view of the block (the columns of rank 0's DataFrame) holds a reference to
        # `_raw`, the array that owns the mapping: its reference count tells whether a result still
        # lives on this generation
        self._raw = np.frombuffer(self._map, dtype=np.uint8)
        self.np_cols = self._raw[: 6 * 8 * n].view(np.float64).reshape(6, n)
        self.np_succ = self._raw[6 * 8 * n: 6 * 8 * n + n]
        self.cols = torch.from_numpy(self.np_cols)      # the D2H destinations
        self.succ = torch.from_numpy(self.np_succ)
        self._own_refs = sys.getrefcount(self._raw)
        rc = torch.cuda.cudart().cudaHostRegister(self.cols.data_ptr(), size, 0)
        if int(rc) != 0:
            raise RuntimeError(f"cudaHostRegister of the shared summary block failed ({rc})")
        self._registered = (self.cols.data_ptr(), size)

    def is_free(self) -> bool:
        """No numpy view of a previous result still references the block."""
        import sys

        return self.cols is not None and sys.getrefcount(self._raw) <= self._own_refs

    def __del__(self):  # pragma: no cover
        try:
            import torch

            if getattr(self, "_registered", None):
                torch.cuda.cudart().cudaHostUnregister(self._registered[0])
        except Exception:
            pass
