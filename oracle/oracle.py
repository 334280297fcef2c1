"""CPU ORACLE for the Monte Carlo path engine — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module. It is the checker, never the thing shipped or
measured as the product; the product (``monte_carlo_retirement_b200``) never imports it and
has no CPU fallback.

What it is: a restatement of the reference algorithm in
``/root/reference/backend/simulation.py`` — the per-path arithmetic in plain C
(``path_oracle.c``, loaded here through ctypes), the random draws through the very same numpy
calls the reference makes (numpy is the reference's third-party RNG dependency:
``SeedSequence.spawn`` / ``default_rng(seed).standard_normal((n, 3))``; numpy 2.4.2 pinned in
the reference's ``uv.lock:420-421``, 2.3.5 installed in this image), and the batch
aggregations through the same pandas calls. Each function cites the reference lines it
follows.

Parity status: PINNED — ``tests/test_oracle_golden.py`` checks it bit-for-bit against
``tests/golden/*.npz`` (outputs of the unmodified Python reference generated in the build
container by ``tests/golden/make_golden.py``) and against the reference's own known-answer
tests.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

MONTHS_PER_YEAR = 12  # backend/constants.py:1
SMALL_EPSILON = 1e-6  # backend/constants.py:3
MAX_STREAMS = 16

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


# --------------------------------------------------------------------------------------------
# ctypes mirror of include/mcr.h (params) and path_oracle.c (record)
# --------------------------------------------------------------------------------------------
class _Stream(C.Structure):
    _fields_ = [
        ("monthly_amount_today", C.c_double),
        ("start_at_age", C.c_double),
        ("tax_rate", C.c_double),
        ("duration_years", C.c_int32),
        ("inflation_indexed", C.c_int32),
    ]


class Params(C.Structure):
    _fields_ = [
        ("initial_balance", C.c_double),
        ("monthly_contribution", C.c_double),
        ("contribution_growth_rate_annual", C.c_double),
        ("monthly_expenses", C.c_double),
        ("current_age", C.c_double),
        ("allocation_inv1_pct", C.c_double),
        ("inv1_mu_log", C.c_double),
        ("inv1_sigma_log", C.c_double),
        ("inf_mu_log", C.c_double),
        ("inf_sigma_log", C.c_double),
        ("prem_mu_log", C.c_double),
        ("prem_sigma_log", C.c_double),
        ("equity_inflation_rho", C.c_double),
        ("inv1_annual_tax_on_gains_rate", C.c_double),
        ("inv1_realized_gains_tax_rate", C.c_double),
        ("inv2_annual_tax_on_gains_rate", C.c_double),
        ("inv2_realized_gains_tax_rate", C.c_double),
        ("inv1_use_realized_gains_tax_system", C.c_int32),
        ("inv2_use_realized_gains_tax_system", C.c_int32),
        ("retirement_years", C.c_int32),
        ("n_streams", C.c_int32),
        ("streams", _Stream * MAX_STREAMS),
    ]


class PathRecord(C.Structure):
    _fields_ = [
        ("start_balance", C.c_double),
        ("final_balance", C.c_double),
        ("years_to_ruin", C.c_double),
        ("first_year_gross", C.c_double),
        ("first_year_real", C.c_double),
        ("inflation_at_ret", C.c_double),
        ("success", C.c_int32),
        ("trajectory_len", C.c_int32),
    ]


RECORD_DTYPE = np.dtype(
    [
        ("start_balance", "<f8"),
        ("final_balance", "<f8"),
        ("years_to_ruin", "<f8"),
        ("first_year_gross", "<f8"),
        ("first_year_real", "<f8"),
        ("inflation_at_ret", "<f8"),
        ("success", "<i4"),
        ("trajectory_len", "<i4"),
    ]
)
assert RECORD_DTYPE.itemsize == C.sizeof(PathRecord)

_lib = None


def build(force: bool = False) -> str:
    """Compile path_oracle.c with gcc (see oracle/Makefile). Returns the .so path."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "path_oracle.c")
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.oracle_stream_start_month.restype = C.c_int32
        L.oracle_stream_start_month.argtypes = [C.c_double, C.c_int32, C.c_double]
        L.oracle_trajectory_len.restype = C.c_int32
        L.oracle_trajectory_len.argtypes = [C.c_int32, C.c_int32]
        L.oracle_net_liquidation.restype = C.c_double
        L.oracle_net_liquidation.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double]
        L.oracle_withdraw.restype = None
        L.oracle_withdraw.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_double, dp]
        L.oracle_rebalance.restype = None
        L.oracle_rebalance.argtypes = [C.POINTER(Params), dp]
        L.oracle_annual_tax.restype = C.c_int
        L.oracle_annual_tax.argtypes = [C.POINTER(Params), dp, C.c_double, C.c_double]
        L.oracle_run_path.restype = C.c_int
        L.oracle_run_path.argtypes = [C.POINTER(Params), C.c_int32, C.c_void_p, C.c_int32,
                                      C.POINTER(PathRecord), C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_run_batch.restype = C.c_int
        L.oracle_run_batch.argtypes = [C.POINTER(Params), C.c_int32, C.c_void_p, C.c_int32,
                                       C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int]
        _lib = L
    return _lib


# --------------------------------------------------------------------------------------------
# pure helpers
# --------------------------------------------------------------------------------------------
def log_params(mean: float, vol: float) -> Tuple[float, float]:
    """arithmetic_to_log_params — simulation.py:14-29."""
    if mean <= -1.0:
        raise ValueError("Arithmetic mean must be greater than -100%.")
    if vol < 0:
        raise ValueError("Volatility cannot be negative.")
    if vol == 0:
        return math.log(1.0 + mean), 0.0
    opm = 1.0 + mean
    sigma = math.sqrt(math.log(1.0 + (vol**2) / (opm**2)))
    mu = math.log(opm) - 0.5 * sigma**2
    return mu, sigma


def _get(cfg: Any, name: str, default=None):
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    return getattr(cfg, name, default)


def params_from_config(cfg: Any) -> Params:
    """Flatten a Config-like object / dict (backend/config.py:48-126) the way
    RetirementMonteCarloSimulator.__init__ consumes it (simulation.py:156-170)."""
    p = Params()
    for name in ("initial_balance", "monthly_contribution", "contribution_growth_rate_annual",
                 "monthly_expenses", "current_age", "allocation_inv1_pct",
                 "inv1_annual_tax_on_gains_rate", "inv1_realized_gains_tax_rate",
                 "inv2_annual_tax_on_gains_rate", "inv2_realized_gains_tax_rate"):
        default = 0.0 if name in ("contribution_growth_rate_annual", "inv1_realized_gains_tax_rate",
                                  "inv2_realized_gains_tax_rate") else None
        setattr(p, name, float(_get(cfg, name, default)))
    p.inv1_mu_log, p.inv1_sigma_log = log_params(_get(cfg, "inv1_returns_mean"),
                                                 _get(cfg, "inv1_returns_volatility"))
    p.inf_mu_log, p.inf_sigma_log = log_params(_get(cfg, "inflation_rate_mean"),
                                               _get(cfg, "inflation_rate_volatility"))
    p.prem_mu_log, p.prem_sigma_log = log_params(_get(cfg, "inv2_premium_over_inflation_mean"),
                                                 _get(cfg, "inv2_premium_over_inflation_volatility"))
    p.equity_inflation_rho = float(_get(cfg, "equity_inflation_correlation", 0.0))
    p.inv1_use_realized_gains_tax_system = int(bool(_get(cfg, "inv1_use_realized_gains_tax_system", False)))
    p.inv2_use_realized_gains_tax_system = int(bool(_get(cfg, "inv2_use_realized_gains_tax_system", True)))
    p.retirement_years = int(_get(cfg, "retirement_years"))
    streams = list(_get(cfg, "other_income_streams", []) or [])
    if len(streams) > MAX_STREAMS:
        raise ValueError(f"at most {MAX_STREAMS} other_income_streams are supported")
    p.n_streams = len(streams)
    for i, s in enumerate(streams):
        p.streams[i].monthly_amount_today = float(_get(s, "monthly_amount_today"))
        p.streams[i].start_at_age = float(_get(s, "start_at_age"))
        p.streams[i].tax_rate = float(_get(s, "tax_rate"))
        d = _get(s, "duration_years", None)
        p.streams[i].duration_years = -1 if d is None else int(d)
        p.streams[i].inflation_indexed = int(bool(_get(s, "inflation_indexed", True)))
    return p


def stream_start_month(current_age: float, working_months: int, start_at_age: float) -> int:
    return int(lib().oracle_stream_start_month(current_age, working_months, start_at_age))


def trajectory_len(working_months: int, retirement_years: int) -> int:
    return int(lib().oracle_trajectory_len(working_months, retirement_years))


def trajectory_time_points(working_months: int, retirement_years: int) -> List[float]:
    """simulation.py:99-123."""
    full, rem = divmod(working_months, MONTHS_PER_YEAR)
    pts = [0.0] + [float(y) for y in range(1, full + 1)]
    t_ret = working_months / MONTHS_PER_YEAR
    if rem:
        pts.append(t_ret)
    pts.extend(t_ret + y for y in range(1, retirement_years + 1))
    return pts


def withdraw(bal, cb, net_target, use_real_tax, rate) -> Tuple[float, float, float, float]:
    out = (C.c_double * 4)()
    lib().oracle_withdraw(bal, cb, net_target, int(bool(use_real_tax)), rate, out)
    return tuple(out)


def net_liquidation(bal, cb, use_real_tax, rate) -> float:
    return float(lib().oracle_net_liquidation(bal, cb, int(bool(use_real_tax)), rate))


def rebalance(p: Params, b1, cb1, b2, cb2) -> Tuple[float, float, float, float]:
    s = (C.c_double * 4)(b1, cb1, b2, cb2)
    lib().oracle_rebalance(C.byref(p), s)
    return tuple(s)


# --------------------------------------------------------------------------------------------
# random draws: the reference's own numpy calls
# --------------------------------------------------------------------------------------------
def draw_shock_path(n_months: int, path_seed: int, rho: float) -> np.ndarray:
    """_draw_shock_path — simulation.py:452-466."""
    rng = np.random.default_rng(path_seed)
    ind = rng.standard_normal((n_months, 3))
    eq = ind[:, 0]
    infl = rho * eq + math.sqrt(max(0.0, 1.0 - rho * rho)) * ind[:, 1]
    return np.column_stack((eq, infl, ind[:, 2]))


class SeedStreams:
    """Seed bookkeeping of the simulator — simulation.py:147-154,177-199 (search/final
    SeedSequence children; path seeds spawned once per (stream, n) and cached, so the spawn
    order matters exactly as in the reference)."""

    def __init__(self, main_seed: int):
        self.main_seed = main_seed
        seq = np.random.SeedSequence(main_seed)
        self._search, self._final = seq.spawn(2)
        self.stream = "final"
        self._cache: Dict[Tuple[str, int], List[int]] = {}

    def use(self, stream: str) -> None:
        assert stream in ("search", "final")
        self.stream = stream

    def path_seeds(self, n: int) -> List[int]:
        key = (self.stream, n)
        if key not in self._cache:
            seq = self._search if self.stream == "search" else self._final
            self._cache[key] = [int(c.generate_state(1)[0]) for c in seq.spawn(n)]
        return self._cache[key]


# --------------------------------------------------------------------------------------------
# path runs
# --------------------------------------------------------------------------------------------
def run_path(p: Params, working_months: int, shocks: np.ndarray) -> Dict[str, Any]:
    """_run_single_simulation_path (simulation.py:476-950) on a given shock matrix; returns the
    same 10-key dict."""
    shocks = np.ascontiguousarray(shocks, dtype=np.float64)
    T = trajectory_len(working_months, p.retirement_years)
    R = p.retirement_years
    traj = np.empty(T)
    real = np.empty(T)
    wr = np.empty(R)
    rec = PathRecord()
    rc = lib().oracle_run_path(C.byref(p), working_months, shocks.ctypes.data, shocks.shape[0],
                               C.byref(rec), traj.ctypes.data, real.ctypes.data, wr.ctypes.data)
    if rc != 0:
        raise ValueError("oracle_run_path: bad arguments")
    return {
        "Start Balance": rec.start_balance,
        "Final Balance": rec.final_balance,
        "Success": bool(rec.success),
        "YearsToRuin": rec.years_to_ruin,
        "First Year Gross Withdrawal": rec.first_year_gross,
        "First Year Real Gross Withdrawal": rec.first_year_real,
        "Trajectory": traj.tolist(),
        "RealTrajectory": real.tolist(),
        "WithdrawalRateTrajectory": wr.tolist(),
        "Inflation At Retirement": rec.inflation_at_ret,
    }


def run_batch(p: Params, working_months: int, shocks: np.ndarray, n_threads: int = 1,
              want_series: bool = True):
    """n paths at once. shocks: (n, n_rows, 3). Returns (records[n] structured array,
    traj[n, T], real[n, T], wr[n, R]) — series are None when want_series is False."""
    shocks = np.ascontiguousarray(shocks, dtype=np.float64)
    n, n_rows, _ = shocks.shape
    T = trajectory_len(working_months, p.retirement_years)
    R = p.retirement_years
    recs = np.zeros(n, dtype=RECORD_DTYPE)
    traj = np.empty((n, T)) if want_series else None
    real = np.empty((n, T)) if want_series else None
    wr = np.empty((n, R)) if want_series else None
    rc = lib().oracle_run_batch(
        C.byref(p), working_months, shocks.ctypes.data, n_rows, n, recs.ctypes.data,
        traj.ctypes.data if want_series else None, real.ctypes.data if want_series else None,
        wr.ctypes.data if want_series else None, n_threads)
    if rc != 0:
        raise ValueError("oracle_run_batch: bad arguments")
    return recs, traj, real, wr


def shocks_for_seeds(p: Params, working_months: int, seeds: Sequence[int]) -> np.ndarray:
    """(n, n_rows, 3) shock tensor the reference would draw for these path seeds
    (simulation.py:487-488)."""
    n_rows = max(working_months + p.retirement_years * MONTHS_PER_YEAR, 1)
    out = np.empty((len(seeds), n_rows, 3))
    for i, s in enumerate(seeds):
        out[i] = draw_shock_path(n_rows, s, p.equity_inflation_rho)
    return out


# --------------------------------------------------------------------------------------------
# batch aggregation — run_monte_carlo_simulations, simulation.py:1012-1128
# --------------------------------------------------------------------------------------------
TRAJ_Q = [0.05, 0.10, 0.25, 0.50, 0.75, 0.90, 0.95]  # :1045
WR_Q = [0.05, 0.25, 0.50, 0.75, 0.95]  # :1109


def aggregate(recs: np.ndarray, traj: np.ndarray, real: np.ndarray, wr: np.ndarray, main_seed: int):
    """The 7-tuple of run_monte_carlo_simulations from per-path results, through the same
    pandas calls the reference makes."""
    import pandas as pd

    summary = pd.DataFrame(
        {
            "Start Balance": recs["start_balance"],
            "Final Balance": recs["final_balance"],
            "Success": recs["success"].astype(bool),
            "YearsToRuin": recs["years_to_ruin"],
            "First Year Gross Withdrawal": recs["first_year_gross"],
            "First Year Real Gross Withdrawal": recs["first_year_real"],
            "Inflation At Retirement": recs["inflation_at_ret"],
        }
    )
    traj_df = pd.DataFrame(traj).transpose()  # (T, n) — :1056
    traj_pct = traj_df.quantile(TRAJ_Q, axis=1).transpose()  # :1059-1061
    k = min(traj_df.shape[1], 5)
    sampled = traj_df.sample(n=k, axis=1, random_state=main_seed)  # :1068-1072
    samples = sampled.values.T.tolist()
    real_df = pd.DataFrame(real).transpose()
    real_samples = real_df.loc[:, sampled.columns].values.T.tolist()  # :1075-1078
    real_pct = real_df.quantile(TRAJ_Q, axis=1).transpose()  # :1091-1093
    wr_df = pd.DataFrame(wr).transpose()
    wr_pct = wr_df.quantile(WR_Q, axis=1).transpose()  # :1108-1110
    wr_counts = [int(v) for v in wr_df.count(axis=1).tolist()]  # :1111-1113
    return summary, traj_pct, samples, wr_pct, real_pct, real_samples, wr_counts


def success_probability(summary) -> float:
    """_success_probability — simulation.py:1130-1136."""
    if summary.empty:
        return 0.0
    return float(summary["Success"].astype(bool).mean() * 100.0)


def median_first_year_withdrawal_rate(summary) -> float:
    """simulation.py:78-96."""
    if summary.empty:
        return float("nan")
    start = summary["Start Balance"]
    valid = start > SMALL_EPSILON
    if not valid.any():
        return float("nan")
    rates = (summary["First Year Real Gross Withdrawal"][valid] / start[valid]) * 100.0
    return float(rates.median())


class OracleSimulator:
    """The reference simulator surface needed by the parity tests, on top of the C oracle."""

    def __init__(self, cfg: Any, main_seed: Optional[int] = None, n_threads: int = 1):
        self.cfg = cfg
        self.p = params_from_config(cfg)
        seed = main_seed if main_seed is not None else _get(cfg, "seed")
        if seed is None:
            raise ValueError("the oracle needs an explicit seed")
        self.main_seed = int(seed)
        self.seeds = SeedStreams(self.main_seed)
        self.n_threads = n_threads

    def use_search_seeds(self):
        self.seeds.use("search")

    def use_final_seeds(self):
        self.seeds.use("final")

    def run_single(self, working_months: int, path_seed: int) -> Dict[str, Any]:
        n_rows = max(working_months + self.p.retirement_years * MONTHS_PER_YEAR, 1)
        return run_path(self.p, working_months, draw_shock_path(n_rows, path_seed, self.p.equity_inflation_rho))

    def run_raw(self, working_months: int, n: int):
        seeds = self.seeds.path_seeds(n)
        shocks = shocks_for_seeds(self.p, working_months, seeds)
        return run_batch(self.p, working_months, shocks, self.n_threads)

    def run(self, working_months: int, n: int):
        recs, traj, real, wr = self.run_raw(working_months, n)
        return aggregate(recs, traj, real, wr, self.main_seed)

    # find_minimum_working_months — simulation.py:1138-1342, restated as a decision procedure
    # over an abstract probability provider so the same logic can be checked on any table.
    def find_minimum_working_months(self, prob_fn: Optional[Callable[[int], float]] = None):
        self.use_search_seeds()
        n = int(_get(self.cfg, "num_simulations_search"))
        if prob_fn is None:
            def prob_fn(m: int) -> float:
                recs, _, _, _ = run_batch(self.p, m, shocks_for_seeds(self.p, m, self.seeds.path_seeds(n)),
                                          self.n_threads, want_series=False)
                return float(recs["success"].astype(bool).mean() * 100.0)
        return search_decisions(prob_fn, int(_get(self.cfg, "starting_working_months_search")),
                                float(_get(self.cfg, "target_probability")), n)


def search_decisions(prob_fn: Callable[[int], float], start: int, target: float, sim_count: int):
    """Bracket -> bisect -> verify — simulation.py:1158-1342. Returns (months, prob, curve,
    probe_order)."""
    limit = start + 70 * MONTHS_PER_YEAR  # :1161
    cache: Dict[int, float] = {}
    curve: List[Dict[str, float]] = []
    order: List[int] = []
    best_seen = -1.0

    def test(m: int) -> float:
        nonlocal best_seen
        if m in cache:
            return cache[m]
        pr = prob_fn(m)
        cache[m] = pr
        order.append(m)
        curve.append({"working_months": m, "working_years": round(m / MONTHS_PER_YEAR, 1),
                      "probability": round(pr, 2)})
        if pr > best_seen:
            best_seen = pr
        return pr

    step = 12
    cur = start
    p_lo = test(cur)
    if p_lo >= target:  # :1229-1232
        return cur, p_lo, curve, order
    lo, hi = start, None
    best_prob = None
    while cur < limit:  # :1234-1268
        gap = target - p_lo
        step = max(step, 24) if gap > 20 else (max(step, 12) if gap > 10 else max(step, 6))
        nxt = min(cur + step, limit)
        if nxt <= cur:
            break
        pr = test(nxt)
        if pr >= target:
            lo, hi, best_prob = cur, nxt, pr
            break
        lo = nxt
        p_lo = pr
        cur = nxt
    if hi is None:  # :1270-1279
        return -1, best_seen, curve, order
    best = hi
    while hi - lo > 1:  # :1282-1291
        mid = (lo + hi) // 2
        pr = test(mid)
        if pr >= target:
            best, best_prob, hi = mid, pr, mid
        else:
            lo = mid
    margin = min(100.0, 150.0 / math.sqrt(sim_count))  # :1296-1299
    tested = sorted(m for m in cache if m <= best)
    near = next((i for i, m in enumerate(tested) if cache[m] >= target - margin), len(tested) - 1)
    v_start = max(start, tested[max(0, near - 1)])  # :1312-1316
    for m in range(v_start, best + 1):  # :1322-1323
        test(m)
    ok = [m for m, pr in cache.items() if start <= m <= best and pr >= target]  # :1325-1332
    if ok:
        best = min(ok)
        best_prob = cache[best]
    return best, best_prob, curve, order
