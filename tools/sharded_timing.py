#!/usr/bin/env python
"""torchrun --nproc-per-node=W tools/sharded_timing.py — phase times of one sharded C3 step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from loguru import logger
logger.remove()
import torch, torch.distributed as dist
import scenarios
from monte_carlo_retirement_b200.config import Config
from monte_carlo_retirement_b200.parallel import ShardedSimulator
from monte_carlo_retirement_b200.simulation import TRAJECTORY_QUANTILES, WITHDRAWAL_RATE_QUANTILES, FINAL_BALANCE_QUANTILES

local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W = dist.get_world_size()
sim = ShardedSimulator(Config(**scenarios.SYNTH_C3), device=local); sim.use_final_seeds()
n = 1_000_000
for _ in range(3): sim.aggregates_device(240, n * W)
torch.cuda.synchronize(); dist.barrier()
b = sim.run_batch_device(240, n, first_path=dist.get_rank() * n)
ctx = sim.native_context
specs = (ctx.series_rows(b.traj, n, b.T, TRAJECTORY_QUANTILES) + ctx.series_rows(b.real, n, b.T, TRAJECTORY_QUANTILES)
         + ctx.series_rows(b.wr, n, b.R, WITHDRAWAL_RATE_QUANTILES))
out16 = torch.empty((len(specs), 16), dtype=torch.float64, device="cuda")
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps
t_local = timed(lambda: ctx.quantiles_rows(specs, out16))
t_dist = timed(lambda: ctx.quantiles_rows(specs, out16, all_reduce=sim.coll.sum_))
t_dist_ad = timed(lambda: ctx.quantiles_rows(specs, out16, all_reduce=sim.coll.sum_, all_reduce_min=sim.coll.min_))
t_pool = timed(lambda: ctx.quantiles_rows(specs, out16, all_reduce=sim.coll.sum_, all_reduce_min=sim.coll.min_,
                                           rank=dist.get_rank(), world=W))
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
ctx.quantiles_rows(specs, out16, all_reduce=sim.coll.sum_, all_reduce_min=sim.coll.min_, rank=dist.get_rank(), world=W)
t_pool_cpu = 1e3 * (time.perf_counter() - t0)
torch.cuda.synchronize()
pool = torch.zeros(len(specs) * 8192, dtype=torch.int64, device="cuda")
t_ar_pool = timed(lambda: sim.coll.sum_(pool), reps=20)
small = torch.zeros(7, dtype=torch.float64, device="cuda"); hists = torch.zeros(160, dtype=torch.int64, device="cuda")
t_hist = timed(lambda: sim._final_balance_histograms(b, small[3:5], small[5:7], hists), reps=10)
t_agree = timed(lambda: sim._agree_min(1), reps=10)
t_tl = timed(lambda: sim.run_batch_device(240, n, first_path=dist.get_rank() * n))
h = torch.zeros(len(specs) * 32 * 256, dtype=torch.int32, device="cuda")
t_ar = timed(lambda: sim.coll.sum_(h), reps=20)
t_step = timed(lambda: sim.aggregates_device(240, n * W))
if dist.get_rank() == 0:
    print(f"W={W}: local select {t_local:.2f} ms | distributed fixed {t_dist:.2f} ms | distributed adaptive {t_dist_ad:.2f} ms | "
          f"one hist all-reduce ({h.numel()*4/1e6:.1f} MB) {t_ar:.3f} ms | full step {t_step:.2f} ms")
    print(f"      pooled select {t_pool:.2f} ms (CPU issue {t_pool_cpu:.2f} ms) | pool all-reduce ({pool.numel()*8/1e6:.1f} MB) {t_ar_pool:.3f} ms | "
          f"final-balance histograms {t_hist:.3f} ms | agree {t_agree:.3f} ms | timeline {t_tl:.2f} ms")
dist.destroy_process_group()
