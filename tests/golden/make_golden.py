"""Generate tests/golden/*.npz from the UNMODIFIED Python reference.

Run in the build container only (the reference lives at /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

For every (scenario, seed stream, working_months, n) of tests/scenarios.py it records the
reference's own path seeds (`_path_seeds`), the per-path dicts of
`_run_single_simulation_path`, the 7-tuple of `run_monte_carlo_simulations`, a probe of the
numpy shock stream (to detect a numpy bit-stream change on the machine that replays the
seeds), helper known answers and `find_minimum_working_months` results. Nothing in here is
product code; the fixtures it writes are what pins the oracle (and through it the CUDA path).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

REF = "/root/reference/backend"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(HERE))

from loguru import logger  # noqa: E402

logger.remove()

from config import Config  # noqa: E402  (reference)
from simulation import RetirementMonteCarloSimulator  # noqa: E402  (reference)

import scenarios  # noqa: E402


def path_cases():
    for name, (cfg_dict, cases) in scenarios.GOLDEN_CASES.items():
        cfg = Config(**cfg_dict)
        sim = RetirementMonteCarloSimulator(cfg)
        out = {"cfg_json": json.dumps(cfg_dict), "main_seed": sim.main_seed, "n_cases": len(cases)}
        for k, (stream, wm, n) in enumerate(cases):
            (sim.use_search_seeds if stream == "search" else sim.use_final_seeds)()
            seeds = list(sim._path_seeds(n))
            res = [sim._run_single_simulation_path(wm, s) for s in seeds]
            tup = sim.run_monte_carlo_simulations(wm, n)
            summary, traj_pct, samples, wr_pct, real_pct, real_samples, wr_counts = tup
            pre = f"case{k}_"
            out[pre + "stream"] = stream
            out[pre + "wm"] = wm
            out[pre + "n"] = n
            out[pre + "seeds"] = np.array(seeds, dtype=np.uint64)
            out[pre + "start"] = np.array([r["Start Balance"] for r in res])
            out[pre + "final"] = np.array([float(r["Final Balance"]) for r in res])
            out[pre + "success"] = np.array([r["Success"] for r in res], dtype=bool)
            out[pre + "ruin"] = np.array([r["YearsToRuin"] for r in res])
            out[pre + "fy_gross"] = np.array([r["First Year Gross Withdrawal"] for r in res])
            out[pre + "fy_real"] = np.array([r["First Year Real Gross Withdrawal"] for r in res])
            out[pre + "infl"] = np.array([r["Inflation At Retirement"] for r in res])
            out[pre + "traj"] = np.array([r["Trajectory"] for r in res])
            out[pre + "real"] = np.array([r["RealTrajectory"] for r in res])
            out[pre + "wr"] = np.array([r["WithdrawalRateTrajectory"] for r in res])
            # aggregated 7-tuple
            assert np.array_equal(summary["Start Balance"].to_numpy(), out[pre + "start"])
            out[pre + "traj_pct"] = traj_pct.to_numpy()
            out[pre + "pct_cols"] = np.array(list(traj_pct.columns), dtype=float)
            out[pre + "real_pct"] = real_pct.to_numpy()
            out[pre + "wr_pct"] = wr_pct.to_numpy()
            out[pre + "wr_cols"] = np.array(list(wr_pct.columns), dtype=float)
            out[pre + "wr_counts"] = np.array(wr_counts, dtype=np.int64)
            out[pre + "samples"] = np.array(samples)
            out[pre + "real_samples"] = np.array(real_samples)
            # shock-stream probe for the first path: first 3 rows and the last row
            n_rows = max(wm + cfg.retirement_years * 12, 1)
            sh = sim._draw_shock_path(n_rows, seeds[0])
            out[pre + "shock_probe"] = np.vstack([sh[:3], sh[-1:]])
            out[pre + "shock_sum"] = np.array([sh.sum()])
        yield name, out


def helper_cases():
    """Random known answers for the three private helpers + the annual-tax routine."""
    rng = np.random.default_rng(2026)
    cfg = Config(**scenarios.TAX_HEAVY)
    sims = {
        "tax_heavy": RetirementMonteCarloSimulator(Config(**scenarios.TAX_HEAVY)),
        "config_json": RetirementMonteCarloSimulator(Config(**scenarios.CONFIG_JSON)),
        "annual_both": RetirementMonteCarloSimulator(Config(**scenarios.ANNUAL_BOTH)),
        "test_base": RetirementMonteCarloSimulator(Config(**scenarios.TEST_BASE)),
    }
    del cfg
    n = 160
    bal = np.concatenate([rng.uniform(0, 2e6, n - 20), rng.uniform(0, 2e-6, 10), np.zeros(10)])
    cb = bal * rng.uniform(0.0, 1.6, n)
    tgt = np.concatenate([rng.uniform(-10, 3e5, n - 10), np.zeros(10)])
    use = rng.integers(0, 2, n).astype(bool)
    rate = np.where(rng.random(n) < 0.15, 0.0, rng.uniform(0, 0.6, n))
    any_sim = sims["config_json"]
    wd = np.array([any_sim._calculate_withdrawal_and_update(float(b), float(c), float(t), bool(u), float(r))
                   for b, c, t, u, r in zip(bal, cb, tgt, use, rate)])
    nl = np.array([any_sim._net_liquidation_value(float(b), float(c), bool(u), float(r))
                   for b, c, u, r in zip(bal, cb, use, rate)])
    out = {"wd_in": np.stack([bal, cb, tgt, use.astype(float), rate], 1), "wd_out": wd, "nl_out": nl}
    b2 = np.concatenate([rng.uniform(0, 2e6, n - 10), np.zeros(10)])
    cb2 = b2 * rng.uniform(0.0, 1.6, n)
    g1 = rng.normal(0, 5e4, n)
    g2 = rng.normal(0, 5e4, n)
    for name, sim in sims.items():
        rb = np.array([sim._rebalance_portfolio(float(a), float(b), float(c), float(d))
                       for a, b, c, d in zip(bal, cb, b2, cb2)])
        at = np.array([[float(x) for x in sim._apply_annual_gain_taxes(float(a), float(b), float(c), float(d),
                                                                      float(e), float(f))]
                       for a, b, c, d, e, f in zip(bal, cb, b2, cb2, g1, g2)])
        out[f"rb_{name}"] = rb
        out[f"at_{name}"] = at
    out["rb_in"] = np.stack([bal, cb, b2, cb2, g1, g2], 1)
    out["sim_names"] = np.array(list(sims.keys()))
    # stream start months over a sweep (simulation.py:47-63)
    from simulation import stream_payment_start_month_index

    ages = [40.0, 35.0, 60.0, 59.999999, 33.3]
    starts = [65.0, 40.0, 60.51, 60.5, 58.25, 0.0, 120.0, 62.0 + 1e-7]
    rows = []
    for a in ages:
        for s in starts:
            for wm in list(range(0, 40)) + [75, 233, 240, 241, 599, 600, 840]:
                rows.append((a, wm, s, stream_payment_start_month_index(a, wm, s)))
    out["stream_start"] = np.array(rows)
    return out


def search_cases():
    out = {}
    for name, (cfg_dict, _) in scenarios.SEARCH_CASES.items():
        d = dict(cfg_dict)
        d["num_processes"] = 8
        sim = RetirementMonteCarloSimulator(Config(**d))
        events = []
        months, prob, curve = sim.find_minimum_working_months(verbose=False, progress_callback=events.append)
        out[name] = {
            "cfg": cfg_dict,
            "months": months,
            "prob": prob,
            "curve": curve,
            "events": events,
        }
        print(f"search {name}: months={months} prob={prob} probes={len(curve)}", flush=True)
    return out


def main():
    for name, out in path_cases():
        path = os.path.join(HERE, f"paths_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)", flush=True)
    path = os.path.join(HERE, "helpers.npz")
    np.savez_compressed(path, **helper_cases())
    print(f"wrote {path}")
    with open(os.path.join(HERE, "search.json"), "w") as f:
        json.dump(search_cases(), f, indent=1)
    print("wrote search.json")
    with open(os.path.join(HERE, "PROVENANCE.txt"), "w") as f:
        import pandas as pd

        f.write("generated by tests/golden/make_golden.py from the unmodified reference at /root/reference\n"
                f"python {sys.version.split()[0]} numpy {np.__version__} pandas {pd.__version__}\n")


if __name__ == "__main__":
    main()
