"""Generate tests/golden/payload.json from the UNMODIFIED reference server code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_payload_golden.py

For a handful of (scenario, working_months, N, search_curve) cases it calls the reference's
own `server._build_result` (backend/server.py:416-565) on the reference simulator and stores the
response dict. These pin monte_carlo_retirement_b200/payload.py: the
legacy mode must reproduce them, the aggregate-only mode must agree on everything that is not
an O(N) list. Nothing in here is product code.
"""
from __future__ import annotations

import json
import os
import sys

REF = "/root/reference/backend"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(HERE))

from loguru import logger  # noqa: E402

logger.remove()

from config import Config  # noqa: E402  (reference)
from simulation import RetirementMonteCarloSimulator  # noqa: E402  (reference)
import server  # noqa: E402  (reference)

import scenarios  # noqa: E402

# (case name, scenario dict, working_months, N_main, search curve source in search.json or None)
CASES = [
    ("config_json", scenarios.CONFIG_JSON, 233, 200, "config_json"),
    ("jorge_json", scenarios.JORGE_JSON, 75, 200, "jorge_json"),
    ("jorge_plus", scenarios.JORGE_PLUS, 75, 120, None),
    ("stressed", scenarios.STRESSED, 150, 160, "stressed"),
    ("broke", scenarios.CORNER_BROKE, 0, 24, None),
    ("test_base_override", scenarios.TEST_BASE, 13, 40, None),
]


def cli_outputs(cfg, wm):
    """What backend/main.py:95-133 logs for the final run (utils.log_simulation_results) and the
    histogram plotting.py:46-59 draws (plt.hist(x, bins=100) == numpy.histogram(x, bins=100);
    matplotlib itself is not installed here)."""
    import numpy as np
    import utils  # reference
    from simulation import median_first_year_withdrawal_rate

    sim = RetirementMonteCarloSimulator(cfg)
    sim.use_final_seeds()
    df = sim.run_monte_carlo_simulations(working_months=wm, num_simulations=cfg.num_simulations_main)[0]
    ok = df["Success"].astype(bool)
    succ = df.loc[ok, "Final Balance"]
    messages = []
    sink = logger.add(lambda m: messages.append(m.record["message"]), level="INFO", format="{message}")
    try:
        utils.log_simulation_results(cfg, wm, ok.mean() * 100.0, df["Start Balance"].median(),
                                     succ.median() if not succ.empty else 0.0,
                                     median_first_year_withdrawal_rate(df), df)
    finally:
        logger.remove(sink)
    hist = None
    if not succ.empty:
        counts, edges = np.histogram(succ.to_numpy() / 1e6, bins=100)
        hist = {"counts": counts.tolist(), "edges": edges.tolist()}
    return {"log": messages, "hist100": hist}


def main() -> None:
    search = json.load(open(os.path.join(HERE, "search.json")))
    out = {}
    for name, cfg_dict, wm, n, curve_src in CASES:
        d = dict(cfg_dict)
        d["num_simulations_main"] = n
        cfg = Config(**d)
        sim = RetirementMonteCarloSimulator(cfg)
        sim.use_final_seeds()
        curve = search[curve_src]["curve"] if curve_src else None
        res = server._build_result(cfg, sim, wm, search_curve=curve)
        server.SimulationResponse(**res)  # schema-valid by the reference's own model
        out[name] = {"cfg": d, "working_months": wm, "search_curve": curve, "result": res,
                     "cli": cli_outputs(cfg, wm)}
    path = os.path.join(HERE, "payload.json")
    with open(path, "w") as f:
        json.dump(out, f, allow_nan=False)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
