import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from loguru import logger
logger.remove()
import torch
import scenarios
from monte_carlo_retirement_b200.config import Config
from monte_carlo_retirement_b200.simulation import RetirementMonteCarloSimulator, FINAL_BALANCE_QUANTILES
sim = RetirementMonteCarloSimulator(Config(**scenarios.SYNTH_C3)); sim.use_final_seeds()
for n in (1_000_000, 8_000_000):
    b = sim.run_batch_device(240, n, series=False)
    ctx = sim.native_context
    x, m = b.cols[1], b.success
    sel = x[m.bool()]
    print(n, "torch min/max", sel.min().item(), sel.max().item(), "count", sel.numel())
    out = torch.empty((1, 16), dtype=torch.float64, device="cuda"); cnt = torch.empty(1, dtype=torch.int64, device="cuda")
    ctx.quantiles_rows(ctx.select_rows([(x, n, m, [0.0, 1.0], False)]), out, counts=cnt)
    print("   select alone ", out[0, 0].item(), out[0, 1].item(), cnt.item())
    rates = torch.empty(n, dtype=torch.float64, device="cuda"); ctx.first_year_rates(b.cols[0], b.cols[3], n, rates)
    specs = [(rates, n, None, [0.5], True), (b.cols[0], n, None, [0.5], True), (x, n, m, [0.5], True),
             (x, n, None, FINAL_BALANCE_QUANTILES, False), (x, n, m, [0.0, 1.0], False)]
    out = torch.empty((5, 16), dtype=torch.float64, device="cuda"); cnt = torch.empty(5, dtype=torch.int64, device="cuda")
    ctx.quantiles_rows(ctx.select_rows(specs), out, counts=cnt)
    print("   select 5 rows", out[4, 0].item(), out[4, 1].item(), cnt.tolist())
    print("   medians", out[0,0].item(), out[1,0].item(), out[2,0].item(), "torch", rates.nanmedian().item(), sel.median().item())
    h = sim.run_aggregates(240, n, bands=False)
    print("   hist sum", sum(h["final_balance_hist_musd_100"]["counts"]), h["success_count"], h["final_balance_hist_musd_100"].get("range"))
