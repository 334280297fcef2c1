"""World-size-2 tests of the N>1 host logic on CPU (gloo): shard arithmetic, the collectives
wrapper and the sharded search driver (per-shard success counts summed across ranks must lead
every rank to the single-process decisions). The CUDA kernels themselves are exercised by the
`-m gpu` tests; shard-invariance of the kernel results is tested there on one GPU by sharding
logically (tests/test_gpu_native.py::test_native_results_do_not_depend_on_sharding_or_candidate)."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import scenarios
from monte_carlo_retirement_b200.parallel import shard_range


def test_shard_range_partitions_the_path_set():
    for n in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (lo, c), (lo2, _) in zip(spans, spans[1:]):
                assert lo + c == lo2
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _path_succeeds(i: np.ndarray, wm: int) -> np.ndarray:
    """Deterministic stand-in for the kernel: path i succeeds iff wm >= its private threshold."""
    thr = (i * 2654435761 % 97) + 20 + (i % 5 == 0) * 40
    return wm >= thr


def _worker(rank, world, port, out_dir):  # noqa: C901
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from loguru import logger

        logger.remove()
        from monte_carlo_retirement_b200.config import Config
        from monte_carlo_retirement_b200.parallel import Collectives, ShardedSimulator
        from monte_carlo_retirement_b200.simulation import RetirementMonteCarloSimulator

        coll = Collectives()
        assert (coll.rank, coll.world) == (rank, world)
        # collectives wrapper
        t = torch.tensor([rank + 1, 10 * (rank + 1)], dtype=torch.int64)
        assert coll.sum_(t.clone()).tolist() == [3, 30]
        assert coll.min_(torch.tensor([float(rank)])).item() == 0.0
        assert coll.max_(torch.tensor([float(rank)])).item() == 1.0
        g = coll.gather_cat(torch.arange(3 + rank, dtype=torch.float64) + 100 * rank, dim=0)
        assert g.tolist() == [0.0, 1.0, 2.0, 100.0, 101.0, 102.0, 103.0]

        cfg = Config(**dict(scenarios.TEST_BASE, target_probability=75.0, num_simulations_search=1001,
                            starting_working_months_search=3))
        calls = []

        def fake_counts(self, candidates, num_simulations, *, first_path=0, with_executed=False):
            calls.append((list(candidates), num_simulations, first_path))
            idx = first_path + np.arange(num_simulations)
            return torch.tensor([int(_path_succeeds(idx, c).sum()) for c in candidates], dtype=torch.int64)

        RetirementMonteCarloSimulator.batched_success_counts = fake_counts  # stands in for the CUDA launch
        sim = ShardedSimulator(cfg, collectives=coll, search_policy="waves")
        sim._ctx = object()  # never touched: the fake replaces the only native call of the search
        events = []
        months, prob, curve = sim.find_minimum_working_months(verbose=False, progress_callback=events.append)
        lo, cnt = shard_range(1001, rank, world)
        assert all(c[1] == cnt and c[2] == lo for c in calls)   # each rank evaluated only its own shard
        assert sim.last_search_stats["launches"] <= 3
        # planning hooks of the sharded aggregate path: every rank must plan with the same numbers
        sim._torch_device = lambda: torch.device("cpu")
        assert sim._agree_min(1000 + 50 * rank, key="budget") == 1000       # minimum over ranks ...
        assert sim._agree_min(5, key="budget") == 1000                      # ... remembered per call shape
        free = (100 + 40 * rank) * 10 ** 9                                   # rank 1 has more free HBM
        torch.cuda.mem_get_info = lambda device=None: (free, 180 * 10 ** 9)
        n, T, R = 125_000_000, 71, 50                                        # config #5 shard: 71 + 71 + 50 GB
        plan = sim._series_plan(n, T, R, True, key=("c5",))
        assert plan == [("traj",), ("real",), ("wr",)]                      # 70 % of the SMALLER 100 GB holds one series
        assert sim._series_plan(1_000_000, 61, 40, True, key=("c3",)) == [("traj", "real", "wr")]
        assert sim._series_plan(n, T, R, False, key=("none",)) == [()]
        # the shared summary block of the sharded 7-tuple path (host side only here): rank 0 creates the
        # POSIX shared-memory generation, rank 1 maps the same pages; a generation is busy while a
        # numpy view of it (rank 0's DataFrame columns) is alive
        import monte_carlo_retirement_b200.parallel as par

        class _NoCudart:
            def cudaHostRegister(self, *a):
                return 0

            def cudaHostUnregister(self, *a):
                return 0

        torch.cuda.cudart = lambda: _NoCudart()
        blk = par._SharedSummaryBlock(coll, f"mcr_test_{port}", 0, 1000)
        lo, cnt = shard_range(1000, rank, world)
        blk.np_cols[:, lo:lo + cnt] = rank + 1.0
        blk.np_succ[lo:lo + cnt] = rank + 1
        coll.barrier()
        assert blk.np_cols[3, 0] == 1.0 and blk.np_cols[3, 999] == 2.0 and blk.np_succ[499] == 1 and blk.np_succ[500] == 2
        assert blk.is_free()
        view = blk.np_cols[2]
        assert not blk.is_free()
        del view
        assert blk.is_free()
        coll.barrier()
        torch.save({"months": months, "prob": prob, "curve": curve, "events": events, "plan": plan},
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_sharded_search_on_two_ranks_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "rank0.pt", weights_only=False)
    r1 = torch.load(tmp_path / "rank1.pt", weights_only=False)
    assert r0 == r1                                              # every rank holds the global answer
    # single-process reference decisions over all 1001 paths (oracle's restatement of the driver)
    from oracle import oracle as orc

    idx = np.arange(1001)
    months, prob, curve, order = orc.search_decisions(
        lambda m: float(int(_path_succeeds(idx, m).sum()) / 1001 * 100.0), 3, 75.0, 1001)
    assert (r0["months"], r0["prob"], r0["curve"]) == (months, prob, curve)
    assert [e["working_months"] for e in r0["events"] if e["type"] == "search_iter"] == order
