"""Helpers shared by the GPU tests."""
from __future__ import annotations

import numpy as np

REL = 1e-9  # north_star: per-path balances within 1e-9 relative in fp64


def make_sim(cfg, **kw):
    from monte_carlo_retirement_b200.config import Config
    from monte_carlo_retirement_b200.simulation import RetirementMonteCarloSimulator

    if isinstance(cfg, dict):
        cfg = Config(**cfg)
    return RetirementMonteCarloSimulator(cfg, **kw)


def assert_close(got, want, rel=REL, abs_tol=1e-6):
    """|got - want| <= rel * |want| + abs_tol * rel-scale. The absolute term only covers values
    the engine itself snaps to zero (balances <= 1e-6 are zeroed by the reference)."""
    got = np.asarray(got, dtype=float)
    want = np.asarray(want, dtype=float)
    assert got.shape == want.shape, (got.shape, want.shape)
    err = np.abs(got - want)
    tol = rel * np.abs(want) + abs_tol * 1e-3
    bad = ~(err <= tol)
    if bad.any():
        i = np.argwhere(bad)[0]
        raise AssertionError(f"{bad.sum()} mismatches; first at {tuple(i)}: got {got[tuple(i)]!r} want "
                             f"{want[tuple(i)]!r} (rel err {err[tuple(i)] / max(abs(want[tuple(i)]), 1e-300):.3e})")


def device_batch_to_host(b):
    cols = b.cols.cpu().numpy()
    ruin = b.ruin.cpu().numpy()
    counters = b.counters.cpu().numpy()
    out = {
        "start": cols[0], "final": cols[1], "fy_gross": cols[2], "fy_real": cols[3], "infl": cols[4],
        "success": b.success.cpu().numpy().astype(bool),
        "ruin_month": ruin,
        "ruin_years": np.where(ruin < 0, np.nan, ruin / 12.0),
        "success_count": counters[0], "executed": counters[1], "ruin_hist": counters[2:],
    }
    if b.traj is not None:
        out["traj"] = b.traj.cpu().numpy().T.copy()  # -> (n, T) like the reference's lists
        out["real"] = b.real.cpu().numpy().T.copy()
        out["wr"] = b.wr.cpu().numpy().T.copy()
    return out


def small_returns_hold(sim, shocks_nrc) -> bool:
    """Do these draws (n, rows, 3) satisfy the bound on |monthly log-return| that the engine proved
    for the scenario's own Philox draws — the guarantee MCR_FLAG_SMALL_RETURNS asks the caller for?"""
    bound = sim.native_context.small_returns_bound
    if not bound > 0:
        return False
    r12 = np.sqrt(12.0)
    worst = 0.0
    for c, (mu, sg) in enumerate(((sim._inv1_mu_log, sim._inv1_sigma_log), (sim._inf_mu_log, sim._inf_sigma_log),
                                  (sim._inv2_prem_mu_log, sim._inv2_prem_sigma_log))):
        worst = max(worst, float(np.max(np.abs(mu / 12.0 + sg / r12 * np.asarray(shocks_nrc)[:, :, c]))))
    return worst < bound


def assert_close_fast(got, want, peak, rel=REL):
    """Contract of the FAST build (FMA contraction, shared reciprocals, short exp polynomial, lean
    month steps) against the parity build: 1e-9 relative, where a balance is measured against
    max(|balance|, 0.1 % of the path's peak balance). The floor is what any non-bit-exact
    arithmetic needs: the last yearly samples of a path that runs dry are the small difference of
    large numbers (a $14 k residue of a $50 M peak amplifies the ~1e-13 rounding-level difference
    of the two builds 3500 times), so an unfloored relative error is unbounded for every
    implementation that is not bit-identical. Balances at their own scale — every successful
    path, and failing paths until shortly before ruin — are held to the plain 1e-9.
    `peak`: per-path scale, shape (n,) (broadcast over a (n, T) series)."""
    got = np.asarray(got, dtype=float)
    want = np.asarray(want, dtype=float)
    assert got.shape == want.shape, (got.shape, want.shape)
    scale = np.asarray(peak, dtype=float)
    if want.ndim == 2:
        scale = scale[:, None]
    err = np.abs(got - want)
    tol = rel * np.maximum(np.abs(want), 1e-3 * scale) + 1e-9
    bad = ~(err <= tol)
    if bad.any():
        i = tuple(np.argwhere(bad)[0])
        raise AssertionError(f"{bad.sum()} mismatches; first at {i}: got {got[i]!r} want {want[i]!r} "
                             f"(err {err[i]:.3e}, tol {np.broadcast_to(tol, got.shape)[i]:.3e})")
