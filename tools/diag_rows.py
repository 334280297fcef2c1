import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from loguru import logger
logger.remove()
import torch
import scenarios
from monte_carlo_retirement_b200.config import Config
from monte_carlo_retirement_b200.simulation import (RetirementMonteCarloSimulator, FINAL_BALANCE_QUANTILES,
                                                    TRAJECTORY_QUANTILES, WITHDRAWAL_RATE_QUANTILES)
sim = RetirementMonteCarloSimulator(Config(**scenarios.SYNTH_C3)); sim.use_final_seeds()
n = 1_000_000
b = sim.run_batch_device(240, n, series=True)
ctx = sim.native_context
rates = torch.empty(n, dtype=torch.float64, device="cuda"); ctx.first_year_rates(b.cols[0], b.cols[3], n, rates)
x, m = b.cols[1], b.success
T, R = b.T, b.R
def t(specs, reps=5):
    d = ctx.select_rows(specs)
    out = torch.empty((len(d), 16), dtype=torch.float64, device="cuda")
    ctx.quantiles_rows(d, out); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ctx.quantiles_rows(d, out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
print("rates median          %.0f us" % t([(rates, n, None, [0.5], True)]))
print("start median          %.0f us" % t([(b.cols[0], n, None, [0.5], True)]))
print("final succ median     %.0f us" % t([(x, n, m, [0.5], True)]))
print("final 9 quantiles     %.0f us" % t([(x, n, None, FINAL_BALANCE_QUANTILES, False)]))
print("final succ min/max    %.0f us" % t([(x, n, m, [0.0, 1.0], False)]))
for name, ser, rows, q in (("traj", b.traj, T, TRAJECTORY_QUANTILES), ("real", b.real, T, TRAJECTORY_QUANTILES), ("wr", b.wr, R, WITHDRAWAL_RATE_QUANTILES)):
    print(f"{name} all rows         %.0f us" % t(ctx.series_rows(ser, n, rows, q)))
    per = [t([(ser[r:r + 1], n, None, q, False)], reps=3) for r in range(rows)]
    print(f"{name} single rows (us):", " ".join("%.0f" % p for p in per))
