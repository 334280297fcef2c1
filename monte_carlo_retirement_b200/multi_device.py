"""All GPUs of the box from ONE process — the multi-GPU form the reference's callers can reach.

The reference is driven by a plain CLI and by FastAPI worker threads (backend/main.py:68-106,
backend/server.py:231-266,309,405): there is no torchrun in front of it. `MultiDeviceSimulator`
therefore runs one worker THREAD per device inside the caller's process; every worker owns a
`ShardedSimulator` bound to its GPU (contiguous global path ranges of one Philox stream, SURVEY
§8e) and the few collectives of the sharded engine go through `PeerCollectives`:
hand-written all-reduce kernels over NVLink / NVSwitch peer memory (csrc/mcr_comm.cu) — no NCCL,
no process group, no host synchronisation inside a collective.

    sim = MultiDeviceSimulator(cfg)                  # all visible devices
    sim = MultiDeviceSimulator(cfg, devices=[0, 1])
    sim.run_monte_carlo_simulations(240, 8_000_000)  # the reference's 7-tuple, global
    sim.find_minimum_working_months()

Results are bit-identical to the single-GPU engine for any number of devices (a path's draws
depend on its global index only; the distributed select is exact). Calls too small to be worth
sharding (`min_sharded_paths`) run on the first device alone — same numbers, less latency.
`dropin/simulation.py` returns this class when more than one device is visible (MCR_DEVICES).
"""
from __future__ import annotations

import ctypes as C
import queue
import threading
from typing import Any, Callable, Dict, List, Optional, Sequence

from . import native
from .parallel import ShardedSimulator
from .simulation import RetirementMonteCarloSimulator, _seed_from_timestamp

_OPS = {("sum", "int32"): 0, ("sum", "int64"): 1, ("min", "int64"): 2, ("max", "int64"): 3,
        ("sum", "float64"): 4, ("min", "float64"): 5, ("max", "float64"): 6}
COMM_BYTES = 32 << 20  # staging per half: the largest collective is the pooled candidate block (~11 MB)


class PeerGroup:
    """What the G ranks of one process share: a CPU barrier, a mailbox and the communicators."""

    def __init__(self, devices: Sequence[int]):
        self.devices = [int(d) for d in devices]
        self.world = len(self.devices)
        self.cpu_barrier = threading.Barrier(self.world, timeout=60)  # a rank that died breaks it instead of hanging the others
        self.mailbox: Dict[str, Any] = {}
        self.comms: List[Optional[int]] = [None] * self.world
        self.lock = threading.Lock()


class PeerCollectives:
    """The collectives interface of parallel.Collectives over csrc/mcr_comm.cu. One instance per
    worker thread; every method must be called by all ranks in the same order."""

    def __init__(self, group: PeerGroup, rank: int, ctx: native.Context):
        self.group = group
        self.rank = rank
        self.world = group.world
        self.lib = ctx.lib
        h = C.c_void_p()
        rc = self.lib.mcr_comm_create(ctx.handle, C.c_int64(COMM_BYTES), C.byref(h))
        if rc != 0:
            raise native.NativeError(f"mcr_comm_create failed ({rc})")
        self.handle = h
        group.comms[rank] = h.value
        group.cpu_barrier.wait()
        arr = (C.c_void_p * self.world)(*group.comms)
        rc = self.lib.mcr_comm_connect(self.handle, rank, self.world, arr)
        if rc != 0:
            raise native.NativeError(self.lib.mcr_comm_last_error(self.handle).decode())
        group.cpu_barrier.wait()

    def _all_reduce(self, t, what: str):
        import torch

        if t.numel() == 0:
            return t
        if not t.is_contiguous():
            raise ValueError("peer all-reduce needs a contiguous tensor")
        op = _OPS.get((what, str(t.dtype).replace("torch.", "")))
        if op is None:
            raise TypeError(f"peer all-reduce: unsupported {what} of {t.dtype}")
        width = t.element_size()
        per = COMM_BYTES // width
        flat = t.view(-1)
        stream = int(torch.cuda.current_stream().cuda_stream) or None
        for at in range(0, flat.numel(), per):
            part = flat[at:at + per]
            rc = self.lib.mcr_comm_all_reduce(self.handle, op, part.data_ptr(), part.numel(), stream)
            if rc != 0:
                raise native.NativeError(self.lib.mcr_comm_last_error(self.handle).decode())
        return t

    def sum_(self, t):
        return self._all_reduce(t, "sum")

    def min_(self, t):
        return self._all_reduce(t, "min")

    def max_(self, t):
        return self._all_reduce(t, "max")

    def barrier(self) -> None:
        """All ranks have reached this point and their enqueued device work is done."""
        import torch

        torch.cuda.current_stream().synchronize()
        self.check()
        self.group.cpu_barrier.wait()

    def broadcast0_(self, t):
        """Rank 0's (small) tensor to everyone, through the host."""
        g = self.group
        if self.rank == 0:
            g.mailbox["bcast"] = t.cpu()
        g.cpu_barrier.wait()
        if self.rank != 0:
            t.copy_(g.mailbox["bcast"])
        g.cpu_barrier.wait()
        return t

    def check(self) -> None:
        """A peer that never joined a collective makes the kernels give up after ~2 s of spinning."""
        st = int(self.lib.mcr_comm_status(self.handle))
        if st != 0:
            raise native.NativeError(f"peer all-reduce #{st} timed out waiting for another device of this process")

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.mcr_comm_destroy(self.handle)
            self.handle = None


class _LocalSummaryBlock:
    """The host block of the sharded 7-tuple path when all ranks live in one process: one pinned
    tensor, each GPU writes its shard over its own PCIe link."""

    def __init__(self, n: int):
        import sys

        import torch

        self.n = n
        self.cols = torch.empty((6, n), dtype=torch.float64, pin_memory=True)
        self.succ = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        self.np_cols = self.cols.numpy()
        self.np_succ = self.succ.numpy()
        self._own = (sys.getrefcount(self.np_cols), sys.getrefcount(self.np_succ))

    def is_free(self) -> bool:
        import sys

        return sys.getrefcount(self.np_cols) <= self._own[0] and sys.getrefcount(self.np_succ) <= self._own[1]


class _Worker(ShardedSimulator):
    """The ShardedSimulator of one device; only the host block of the 7-tuple path differs from the
    torchrun form (process-local pinned memory instead of POSIX shared memory)."""

    def _select(self, specs, out16, counts=None, stepwise: bool = False):
        # the whole pooled protocol in one library call (the ranks are threads under one interpreter lock)
        if stepwise:
            return super()._select(specs, out16, counts, stepwise=True)
        coll = self.coll
        return self.native_context.quantiles_rows_comm(coll.handle, coll.rank, coll.world, specs, out16, counts=counts)

    def _summary_block(self, n_global: int):
        import torch

        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self._torch_device())
        g = self.coll.group
        if self.coll.rank == 0:
            blocks = g.mailbox.setdefault("blocks", [])
            free = [b for b in blocks if b.n == n_global and b.is_free()]
            if not free:
                free = [_LocalSummaryBlock(n_global)]
                blocks.append(free[0])
                del blocks[:-4]          # generations still referenced by live results stay alive through them
            g.mailbox["block"] = free[0]
        g.cpu_barrier.wait()
        blk = g.mailbox["block"]
        g.cpu_barrier.wait()
        return blk


def _worker_main(pool: "_WorkerPool", rank: int, params_model, seed, kw) -> None:
    """Thread body of one device. It references the pool only (never the user-facing simulator),
    so dropping the simulator ends the threads (MultiDeviceSimulator.__del__ -> pool.close())."""
    import torch

    dev = pool.devices[rank]
    sim = None
    try:
        torch.cuda.set_device(dev)
        ctx_holder = RetirementMonteCarloSimulator(params_model, seed, device=dev, **kw)
        coll = PeerCollectives(pool.group, rank, ctx_holder.native_context)
        sim = _Worker(params_model, seed, collectives=coll, device=dev, **kw)
        sim._ctx = ctx_holder._ctx           # one native context per device
        pool.done.put((rank, "init", None, None))
    except BaseException as exc:  # noqa: BLE001
        pool.done.put((rank, "init", None, exc))
        return
    while True:
        job = pool.jobs[rank].get()
        if job is None:
            break
        tag, fn = job
        try:
            with torch.cuda.device(dev):
                out = fn(sim, rank)
                sim.coll.check()
            pool.done.put((rank, tag, out, None))
        except BaseException as exc:  # noqa: BLE001
            pool.done.put((rank, tag, None, exc))
    try:
        sim.coll.close()
    except Exception:  # pragma: no cover
        pass


class _WorkerPool:
    """One thread + one ShardedSimulator per device, created on the first call big enough to shard."""

    def __init__(self, devices: Sequence[int], params_model, seed, kw):
        self.devices = list(devices)
        self.group = PeerGroup(self.devices)
        self.jobs: List[queue.Queue] = [queue.Queue() for _ in self.devices]
        self.done: queue.Queue = queue.Queue()
        self.lock = threading.Lock()   # one collective call sequence at a time
        self.threads = [threading.Thread(target=_worker_main, args=(self, r, params_model, seed, kw), daemon=True,
                                         name=f"mcr-dev{d}") for r, d in enumerate(self.devices)]
        for t in self.threads:
            t.start()
        self.collect("init")

    def collect(self, tag: str):
        outs: Dict[int, Any] = {}
        first_exc = None
        for _ in self.devices:
            rank, got, out, exc = self.done.get()
            assert got == tag, (got, tag)
            outs[rank] = out
            if exc is not None and first_exc is None:
                first_exc = exc
        if first_exc is not None:
            raise first_exc
        return outs

    def on_all(self, tag: str, fn: Callable[[ShardedSimulator, int], Any]):
        """Run fn(worker_simulator, rank) on every device thread; rank 0's return value."""
        with self.lock:
            for q in self.jobs:
                q.put((tag, fn))
            return self.collect(tag)[0]

    def close(self) -> None:
        for q in self.jobs:
            q.put(None)
        for t in self.threads:
            t.join(timeout=10)


class MultiDeviceSimulator(RetirementMonteCarloSimulator):
    """RetirementMonteCarloSimulator over all (or the given) devices of this process."""

    _device_search_ok = True

    def __init__(self, params_model, main_seed_override: Optional[int] = None, *, devices: Optional[Sequence[int]] = None,
                 min_sharded_paths: int = 65536, **kw):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device is available: the B200 engine has no CPU fallback")
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        if main_seed_override is None and params_model.seed is None:
            main_seed_override = _seed_from_timestamp()   # one seed for every rank
        kw.pop("device", None)
        super().__init__(params_model, main_seed_override, device=int(devices[0]), **kw)
        if self.rng_mode != "philox":
            raise ValueError("several devices need the counter-based Philox draws (rng='philox')")
        self.devices = [int(d) for d in devices]
        self.min_sharded_paths = int(min_sharded_paths)
        self._pool: Optional[_WorkerPool] = None
        self._pool_args = (params_model.model_copy(deep=True), main_seed_override, dict(kw))

    def _on_all(self, tag: str, fn: Callable[[ShardedSimulator, int], Any]):
        if self._pool is None:
            self._pool = _WorkerPool(self.devices, *self._pool_args)
            if self._stream_name == "search":
                self._pool.on_all("seeds", lambda s, r: s.use_search_seeds())
        return self._pool.on_all(tag, fn)

    def close(self) -> None:
        if self._pool is not None:
            self._pool.close()
            self._pool = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ---- the reference surface ----------------------------------------------------------------------
    def use_search_seeds(self) -> None:
        super().use_search_seeds()
        if self._pool is not None:
            self._pool.on_all("seeds", lambda s, r: s.use_search_seeds())

    def use_final_seeds(self) -> None:
        super().use_final_seeds()
        if self._pool is not None:
            self._pool.on_all("seeds", lambda s, r: s.use_final_seeds())

    def _small(self, n: int) -> bool:
        return int(n) < self.min_sharded_paths or len(self.devices) == 1

    def run_monte_carlo_simulations(self, working_months: int, num_simulations: int):
        if self._small(num_simulations):
            return super().run_monte_carlo_simulations(working_months, num_simulations)
        return self._on_all("mc", lambda s, r: s.run_monte_carlo_simulations(working_months, num_simulations))

    def run_aggregates(self, working_months: int, num_simulations: int, **kw) -> Dict[str, Any]:
        if self._small(num_simulations):
            return super().run_aggregates(working_months, num_simulations, **kw)
        return self._on_all("agg", lambda s, r: s.run_aggregates(working_months, num_simulations, **kw))

    def batched_success_counts(self, candidates: Sequence[int], num_simulations: int, *, first_path: int = 0,
                               with_executed: bool = False):
        if self._small(num_simulations):
            return super().batched_success_counts(candidates, num_simulations, first_path=first_path,
                                                  with_executed=with_executed)

        def job(s, r):
            out = s.batched_success_counts(candidates, num_simulations, first_path=first_path, with_executed=with_executed)
            if with_executed:
                return s._reduce_counts(out[0]), s._reduce_counts(out[1])
            return s._reduce_counts(out)

        return self._on_all("counts", job)

    def find_minimum_working_months(self, verbose: bool = True, progress_callback=None):
        patched = "run_monte_carlo_simulations" in self.__dict__
        if patched or self._small(self.params_model.num_simulations_search):
            # small probes (or an instance-level replacement of run_monte_carlo_simulations, which the
            # reference's tests install): the single-device search of the base class
            return super().find_minimum_working_months(verbose=verbose, progress_callback=progress_callback)
        self.use_search_seeds()

        def job(s, r):
            return s.find_minimum_working_months(verbose=verbose and r == 0,
                                                 progress_callback=progress_callback if r == 0 else None)

        out = self._on_all("search", job)
        self.last_search_stats = self._on_all("stats", lambda s, r: s.last_search_stats)
        return out


def devices_from_env() -> List[int]:
    """MCR_DEVICES: "all" (default) | "0" | "0,1,2": the devices a drop-in simulator may use."""
    import os

    import torch

    if not torch.cuda.is_available():
        return []
    spec = os.environ.get("MCR_DEVICES", "all").strip().lower()
    n = torch.cuda.device_count()
    if spec in ("", "all"):
        return list(range(n))
    devs = [int(x) for x in spec.split(",") if x.strip() != ""]
    bad = [d for d in devs if not 0 <= d < n]
    if bad:
        raise ValueError(f"MCR_DEVICES names devices {bad} but {n} are visible")
    return devs
