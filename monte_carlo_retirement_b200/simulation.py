"""Drop-in mirror of the reference's `simulation` module (backend/simulation.py) whose Monte
Carlo path engine runs as hand-written sm_100a CUDA behind the C-ABI of include/mcr.h.

Same names, argument meaning, return shapes and error behaviour as the reference
(`RetirementMonteCarloSimulator`, the 8 module-level helpers), so `backend/main.py`,
`backend/server.py` and the reference's own tests run against it unchanged (put `dropin/` on
`sys.path` instead of the reference's `simulation.py`). What differs is where the work
happens:

  * `_run_single_simulation_path` / the three private helpers -> one strict CUDA thread each;
  * `run_monte_carlo_simulations` -> one launch of the fused timeline kernel for all paths plus
    device radix-select quantiles; only the N x 7 summary columns and the small band tables
    come back to the host. Random draws are counter-based Philox on the device
    (`rng="philox"`, default) or the reference's own numpy PCG64 draws uploaded and replayed
    (`rng="numpy"`, bit-compatible success flags, balances within 1e-9 relative);
  * `find_minimum_working_months` -> the reference's bracket / bisect / verify decisions,
    replayed verbatim on the host over success tables that the batched search kernel fills
    for many candidate `working_months` per launch (common random numbers across candidates).

There is no CPU fallback: without a CUDA device every compute call raises.
"""
from __future__ import annotations

import math
import os
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import pandas as pd

try:
    from loguru import logger
except Exception:  # pragma: no cover
    import logging

    logger = logging.getLogger("mcr_b200")

from . import native
from .constants import MONTHS_PER_YEAR, SMALL_EPSILON

TRAJECTORY_QUANTILES = [0.05, 0.10, 0.25, 0.50, 0.75, 0.90, 0.95]  # simulation.py:1045
WITHDRAWAL_RATE_QUANTILES = [0.05, 0.25, 0.50, 0.75, 0.95]  # simulation.py:1109
FINAL_BALANCE_QUANTILES = [0.01, 0.05, 0.10, 0.25, 0.50, 0.75, 0.90, 0.95, 0.99]  # server.py:453-455
SUMMARY_COLUMNS = [
    "Start Balance", "Final Balance", "Success", "YearsToRuin", "First Year Gross Withdrawal",
    "First Year Real Gross Withdrawal", "Inflation At Retirement",
]  # simulation.py:1013-1024


# ---------------------------------------------------------------------------------------------
# module-level helpers (simulation.py:14-123) — host-side scalar math, kept in Python so the
# constants handed to the kernel equal the reference's bit for bit
# ---------------------------------------------------------------------------------------------
def arithmetic_to_log_params(mean: float, vol: float) -> Tuple[float, float]:
    """Arithmetic annual (mean, vol) -> lognormal (mu_log, sigma_log) with E[gross] = 1 + mean
    (simulation.py:14-29)."""
    if mean <= -1.0:
        raise ValueError("Arithmetic mean must be greater than -100%.")
    if vol < 0:
        raise ValueError("Volatility cannot be negative.")
    if vol == 0:
        return math.log(1.0 + mean), 0.0
    one_plus_mean = 1.0 + mean
    sigma_log = math.sqrt(math.log(1.0 + (vol**2) / (one_plus_mean**2)))
    return math.log(one_plus_mean) - 0.5 * sigma_log**2, sigma_log


def retirement_age(current_age: float, working_months: int) -> float:
    """simulation.py:32-34."""
    return current_age + working_months / MONTHS_PER_YEAR


def stream_payment_start_age(current_age: float, working_months: int, start_at_age: float) -> float:
    """simulation.py:37-44."""
    return max(retirement_age(current_age, working_months), float(start_at_age))


def stream_payment_start_month_index(current_age: float, working_months: int, start_at_age: float) -> int:
    """First retirement month whose payment date is at/after eligibility (simulation.py:47-63)."""
    start = retirement_age(current_age, working_months)
    eligible = stream_payment_start_age(current_age, working_months, start_at_age)
    return max(0, int(math.ceil((eligible - start) * MONTHS_PER_YEAR - SMALL_EPSILON)))


def age_at_retirement_year(current_age: float, working_months: int, year_num: int) -> float:
    """simulation.py:66-70."""
    return retirement_age(current_age, working_months) + year_num


def years_from_t0_to_age(current_age: float, target_age: float) -> float:
    """simulation.py:73-75."""
    return max(0.0, float(target_age) - float(current_age))


def median_first_year_withdrawal_rate(summary_df: pd.DataFrame) -> float:
    """Median first-year real gross withdrawal / start balance, in % (simulation.py:78-96)."""
    if summary_df.empty:
        return float("nan")
    start = summary_df["Start Balance"]
    column = ("First Year Real Gross Withdrawal" if "First Year Real Gross Withdrawal" in summary_df.columns
              else "First Year Gross Withdrawal")
    valid = start > SMALL_EPSILON
    if not valid.any():
        return float("nan")
    return float(((summary_df[column][valid] / start[valid]) * 100.0).median())


def trajectory_time_points(working_months: int, retirement_years: int) -> List[float]:
    """Year value of every yearly trajectory sample (simulation.py:99-123)."""
    full_years, partial = divmod(working_months, MONTHS_PER_YEAR)
    points = [0.0]
    points.extend(float(y) for y in range(1, full_years + 1))
    retirement_time = working_months / MONTHS_PER_YEAR
    if partial:
        points.append(retirement_time)
    points.extend(retirement_time + y for y in range(1, retirement_years + 1))
    return points


def _seed_from_timestamp() -> int:
    """utils.py:9-11."""
    import datetime as _dt
    import hashlib

    ts = _dt.datetime.now(_dt.timezone.utc).isoformat()
    return int.from_bytes(hashlib.sha256(ts.encode()).digest()[:8], "big") % (2**32 - 1)


def params_from_model(p: Any) -> native.Params:
    """Flatten a Config (backend/config.py:48-126) into the POD the C-ABI takes, deriving the
    lognormal parameters exactly as the reference's constructor does (simulation.py:156-166)."""
    streams = list(getattr(p, "other_income_streams", None) or [])
    if len(streams) > native.MAX_STREAMS:
        raise ValueError(f"the CUDA engine supports at most {native.MAX_STREAMS} other_income_streams "
                         f"(got {len(streams)})")
    q = native.Params()
    q.initial_balance = p.initial_balance
    q.monthly_contribution = p.monthly_contribution
    q.contribution_growth_rate_annual = p.contribution_growth_rate_annual
    q.monthly_expenses = p.monthly_expenses
    q.current_age = p.current_age
    q.allocation_inv1_pct = p.allocation_inv1_pct
    q.inv1_mu_log, q.inv1_sigma_log = arithmetic_to_log_params(p.inv1_returns_mean, p.inv1_returns_volatility)
    q.inf_mu_log, q.inf_sigma_log = arithmetic_to_log_params(p.inflation_rate_mean, p.inflation_rate_volatility)
    q.prem_mu_log, q.prem_sigma_log = arithmetic_to_log_params(p.inv2_premium_over_inflation_mean,
                                                               p.inv2_premium_over_inflation_volatility)
    q.equity_inflation_rho = p.equity_inflation_correlation
    q.inv1_annual_tax_on_gains_rate = p.inv1_annual_tax_on_gains_rate
    q.inv1_realized_gains_tax_rate = p.inv1_realized_gains_tax_rate
    q.inv2_annual_tax_on_gains_rate = p.inv2_annual_tax_on_gains_rate
    q.inv2_realized_gains_tax_rate = p.inv2_realized_gains_tax_rate
    q.inv1_use_realized_gains_tax_system = int(bool(p.inv1_use_realized_gains_tax_system))
    q.inv2_use_realized_gains_tax_system = int(bool(p.inv2_use_realized_gains_tax_system))
    q.retirement_years = int(p.retirement_years)
    q.n_streams = len(streams)
    for i, s in enumerate(streams):
        q.streams[i].monthly_amount_today = s.monthly_amount_today
        q.streams[i].start_at_age = s.start_at_age
        q.streams[i].tax_rate = s.tax_rate
        q.streams[i].duration_years = -1 if s.duration_years is None else int(s.duration_years)
        q.streams[i].inflation_indexed = int(bool(s.inflation_indexed))
    return q


class DeviceBatch:
    """Device-resident result of one batch (torch tensors owned by the caller)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def years_to_ruin_into(self, dst) -> None:
        """dst[i] = ruin month / 12, NaN for paths that never failed — the reference's
        `YearsToRuin` column (simulation.py:825-828,943), filled on the device by one kernel with
        an IEEE division (a torch `div_` by a Python scalar multiplies by 1/12 instead, which is
        1 ulp off CPython's `m / 12` for a third of the months)."""
        assert dst.is_contiguous() and dst.numel() == self.n
        self.ctx.years_to_ruin(self.ruin, self.n, dst)

    def years_to_ruin_slice(self, dst, lo: int, count: int) -> None:
        """The same for paths [lo, lo + count) only (dst: the whole n-element column)."""
        self.ctx.years_to_ruin(self.ruin[lo:lo + count], count, dst[lo:lo + count])


class DeviceAggregates:
    """Device-resident aggregates of one batch (see RetirementMonteCarloSimulator.aggregates_device)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def wait(self) -> None:
        """Make the current stream wait for a pipelined call's reductions."""
        if getattr(self, "ready", None) is not None:
            import torch

            torch.cuda.current_stream().wait_event(self.ready)

    def to_host(self) -> Dict[str, Any]:
        """Copy the aggregates (a few KB) to the host. On several GPUs this is a COLLECTIVE call in
        the rare case handled below (every rank must call it, as every rank must have called
        aggregates_device): the flag it looks at is identical on all ranks."""
        if getattr(self, "ready", None) is not None:
            self.ready.synchronize()  # pipelined call: the reductions ran on their own stream
        flag = getattr(self, "select_flag", None)
        if flag is not None and int(flag.item()) != 0:
            # multi-GPU only, rare: the pooled select could not finish a row from the pooled
            # candidates. THIS aggregate's batch is still resident (self.batch owns the series), so
            # the rows are selected again with the plain stepwise protocol — nothing is re-simulated.
            sim, kw = self.reselect
            sim.select_fallbacks += 1
            again = sim._aggregate_batch(self.batch, stepwise=True, **kw)
            return again.to_host()
        b = self.batch
        n, T, R = int(getattr(self, "n_override", b.n)), b.T, b.R
        nq, nw, nf = len(TRAJECTORY_QUANTILES), len(WITHDRAWAL_RATE_QUANTILES), len(FINAL_BALANCE_QUANTILES)
        o16 = self.out16.cpu().numpy()          # [rows][16]: 3 medians, final quantiles, min / max, then the band rows
        cnt = self.cnt_all.cpu().numpy()
        hists = self.hists.cpu().numpy()
        counters = b.counters.cpu().numpy()
        rng_1 = o16[4, 0:2]
        rng_m = rng_1 / 1e6                     # (IEEE division, as the histogram kernel's)
        out: Dict[str, Any] = {}
        B0 = getattr(self, "first_band_row", None)
        if B0 is not None:
            out["trajectory_bands"] = pd.DataFrame(o16[B0:B0 + T, :nq].copy(), columns=TRAJECTORY_QUANTILES)
            out["real_trajectory_bands"] = pd.DataFrame(o16[B0 + T:B0 + 2 * T, :nq].copy(), columns=TRAJECTORY_QUANTILES)
            out["withdrawal_rate_bands"] = pd.DataFrame(o16[B0 + 2 * T:B0 + 2 * T + R, :nw].copy(),
                                                        columns=WITHDRAWAL_RATE_QUANTILES)
            out["withdrawal_rate_counts"] = [int(v) for v in cnt[B0 + 2 * T:B0 + 2 * T + R]]
        elif self.band_block is not None:       # series swept in several passes: bands gathered per pass
            blk = self.band_block.cpu().numpy()
            out["trajectory_bands"] = pd.DataFrame(blk[:T * nq].reshape(T, nq).copy(), columns=TRAJECTORY_QUANTILES)
            out["real_trajectory_bands"] = pd.DataFrame(blk[T * nq:2 * T * nq].reshape(T, nq).copy(),
                                                        columns=TRAJECTORY_QUANTILES)
            out["withdrawal_rate_bands"] = pd.DataFrame(blk[2 * T * nq:].reshape(R, nw).copy(),
                                                        columns=WITHDRAWAL_RATE_QUANTILES)
            out["withdrawal_rate_counts"] = [int(v) for v in self.wr_counts.cpu().numpy()]
        if getattr(self, "sample_block", None) is not None:
            smp = self.sample_block.cpu().numpy()
            out["sample_paths"] = smp[0].tolist()
            out["real_sample_paths"] = smp[1].tolist()
        n_success = int(counters[0])
        out.update({
            "num_simulations": n,
            "working_months": int(b.working_months),
            "success_count": n_success,
            "success_probability": float(n_success / n * 100.0),
            "executed_path_months": int(counters[1]),
            "median_first_year_withdrawal_rate": float(o16[0, 0]),
            "median_start_balance": float(o16[1, 0]),
            "median_final_balance_successful": float(o16[2, 0]) if cnt[2] > 0 else 0.0,
            "final_balance_quantiles": dict(zip(FINAL_BALANCE_QUANTILES, (float(v) for v in o16[3, :nf]))),
            "final_balance_hist_musd_100": {"range": [float(v) for v in rng_m], "counts": hists[:100].tolist()},
            "final_balance_hist_60": {"range": [float(v) for v in rng_1], "counts": hists[100:].tolist()},
            "ruin_month_hist": counters[2:].tolist(),
        })
        return out


class RetirementMonteCarloSimulator:
    """Monte Carlo retirement simulator — same surface as the reference class
    (simulation.py:126-1342); the path engine is the CUDA library."""

    def __init__(self, params_model, main_seed_override: Optional[int] = None, *, device: Optional[int] = None,
                 rng: Optional[str] = None, strict: Optional[bool] = None, search_policy: Optional[str] = None):
        self.params_model = params_model.model_copy(deep=True)
        if main_seed_override is not None:
            if main_seed_override < 0:
                raise ValueError("main_seed_override must be nonnegative.")
            self.main_seed = main_seed_override
        elif self.params_model.seed is not None:
            self.main_seed = self.params_model.seed
        else:
            self.main_seed = _seed_from_timestamp()

        # numpy seed streams, kept for rng="numpy" and `_path_seeds` (simulation.py:147-154)
        seed_seq = np.random.SeedSequence(self.main_seed)
        self._search_seed_seq, self._final_seed_seq = seed_seq.spawn(2)
        self._stream_name = "final"
        self._active_seed_seq = self._final_seed_seq
        self._path_seed_cache: Dict[Tuple[str, int], List[int]] = {}
        self._sample_column_cache: Dict[int, List[int]] = {}

        p = self.params_model
        self._inv1_mu_log, self._inv1_sigma_log = arithmetic_to_log_params(p.inv1_returns_mean,
                                                                           p.inv1_returns_volatility)
        self._inf_mu_log, self._inf_sigma_log = arithmetic_to_log_params(p.inflation_rate_mean,
                                                                         p.inflation_rate_volatility)
        self._inv2_prem_mu_log, self._inv2_prem_sigma_log = arithmetic_to_log_params(
            p.inv2_premium_over_inflation_mean, p.inv2_premium_over_inflation_volatility)
        self._equity_inflation_rho = p.equity_inflation_correlation
        self._native_params = params_from_model(p)  # raises ValueError for > 16 streams

        # GPU knobs: keyword-only / environment so the reference signatures stay intact
        self.rng_mode = (rng or os.environ.get("MCR_RNG", "philox")).lower()
        if self.rng_mode not in ("philox", "numpy"):
            raise ValueError("rng must be 'philox' or 'numpy'")
        self.strict = bool(int(os.environ.get("MCR_STRICT", "0"))) if strict is None else bool(strict)
        self.search_policy = (search_policy or os.environ.get("MCR_SEARCH_POLICY", "auto")).lower()
        if self.search_policy not in ("auto", "waves", "probe", "grid", "sequential"):
            raise ValueError("search_policy must be 'auto', 'waves', 'probe', 'grid' or 'sequential'")
        self._device_index = device
        self.select_fallbacks = 0  # several GPUs: selects repeated with the stepwise protocol (see DeviceAggregates.to_host)
        # host-returning batch calls split the timeline launch so that the device-to-host copy of the
        # summary columns overlaps the simulation (run_batch_device(chunks=...))
        self.e2e_chunks = int(os.environ.get("MCR_E2E_CHUNKS", "2"))
        self.e2e_first_fraction = float(os.environ.get("MCR_E2E_FIRST_FRACTION", "0.7"))
        self._ctx: Optional[native.Context] = None
        self.last_search_stats: Dict[str, Any] = {}
        logger.info(f"Simulator initialized for scenario '{p.Nickname}' with main seed: {self.main_seed}")

    # ---- native context (created on first compute call so pure host logic needs no GPU) ------
    @property
    def native_context(self) -> native.Context:
        if self._ctx is None:
            import torch

            if not torch.cuda.is_available():
                raise RuntimeError("no CUDA device is available: the B200 engine has no CPU fallback")
            if self._device_index is None:
                self._device_index = torch.cuda.current_device()
            self._ctx = native.Context(self._native_params, int(self.main_seed), int(self._device_index))
        return self._ctx

    def _torch_device(self):
        import torch

        self.native_context  # noqa: B018  (binds _device_index)
        return torch.device("cuda", int(self._device_index))

    # ---- seed streams (simulation.py:177-199) -------------------------------------------------
    def use_search_seeds(self) -> None:
        self._stream_name = "search"
        self._active_seed_seq = self._search_seed_seq

    def use_final_seeds(self) -> None:
        self._stream_name = "final"
        self._active_seed_seq = self._final_seed_seq

    def _seed_stream_id(self) -> int:
        return native.STREAM_SEARCH if self._stream_name == "search" else native.STREAM_FINAL

    def _path_seeds(self, num_simulations: int) -> List[int]:
        """numpy path seeds, spawned once per (stream, n) (simulation.py:187-199)."""
        key = (self._stream_name, num_simulations)
        if key not in self._path_seed_cache:
            children = self._active_seed_seq.spawn(num_simulations)
            self._path_seed_cache[key] = [int(c.generate_state(1)[0]) for c in children]
        return self._path_seed_cache[key]

    def _draw_shock_path(self, n_months: int, path_seed: int) -> np.ndarray:
        """Correlated unit shocks (n_months, 3) from the reference's numpy stream
        (simulation.py:452-466)."""
        independent = np.random.default_rng(path_seed).standard_normal((n_months, 3))
        equity = independent[:, 0]
        rho = self._equity_inflation_rho
        inflation = rho * equity + math.sqrt(max(0.0, 1.0 - rho * rho)) * independent[:, 1]
        return np.column_stack((equity, inflation, independent[:, 2]))

    def _monthly_gross_from_shock(self, mu_log: float, sigma_log: float, z: float) -> float:
        """simulation.py:468-474 (host convenience; the kernel evaluates the same expression)."""
        return float(math.exp(mu_log / MONTHS_PER_YEAR + sigma_log / math.sqrt(MONTHS_PER_YEAR) * z))

    # ---- the three private helpers the reference tests call (one strict CUDA thread each) ----
    def _calculate_withdrawal_and_update(self, bal_inv: float, cb_inv: float, net_withdrawal_target_for_inv: float,
                                         use_real_tax: bool, real_tax_rate: float) -> Tuple[float, float, float, float]:
        """simulation.py:201-254."""
        return self.native_context.helper_withdraw(float(bal_inv), float(cb_inv), float(net_withdrawal_target_for_inv),
                                                   bool(use_real_tax), float(real_tax_rate))

    def _net_liquidation_value(self, balance: float, cost_basis: float, use_realized_gains_tax: bool,
                               realized_gains_tax_rate: float) -> float:
        """simulation.py:256-272 (a staticmethod in the reference; callers use keyword-free
        instance calls, which keep working)."""
        return self.native_context.helper_net_liquidation(float(balance), float(cost_basis),
                                                          bool(use_realized_gains_tax), float(realized_gains_tax_rate))

    def _rebalance_portfolio(self, bal_inv1: float, cb_inv1: float, bal_inv2: float,
                             cb_inv2: float) -> Tuple[float, float, float, float]:
        """simulation.py:274-359."""
        return self.native_context.helper_rebalance(float(bal_inv1), float(cb_inv1), float(bal_inv2), float(cb_inv2))

    def _apply_annual_gain_taxes(self, balance_inv1: float, cost_basis_inv1: float, balance_inv2: float,
                                 cost_basis_inv2: float, gain_inv1: float,
                                 gain_inv2: float) -> Tuple[float, float, float, float, bool]:
        """simulation.py:361-450."""
        return self.native_context.helper_annual_tax(float(balance_inv1), float(cost_basis_inv1), float(balance_inv2),
                                                     float(cost_basis_inv2), float(gain_inv1), float(gain_inv2))

    # ---- one path (simulation.py:476-950) ----------------------------------------------------
    def _run_single_simulation_path(self, working_months: int, path_seed: int) -> Dict[str, Union[float, List[float]]]:
        p = self.params_model
        total_months = working_months + p.retirement_years * MONTHS_PER_YEAR
        shocks = self._draw_shock_path(max(total_months, 1), path_seed)
        return self._run_path_on_shocks(working_months, shocks)

    def _run_path_on_shocks(self, working_months: int, shocks: np.ndarray) -> Dict[str, Union[float, List[float]]]:
        rec, traj, real, wr = self.native_context.single_path(int(working_months), shocks)
        return {
            "Start Balance": rec.start_balance,
            "Final Balance": rec.final_balance,
            "Success": bool(rec.success),
            "YearsToRuin": float("nan") if rec.ruin_month < 0 else rec.ruin_month / MONTHS_PER_YEAR,
            "First Year Gross Withdrawal": rec.first_year_gross,
            "First Year Real Gross Withdrawal": rec.first_year_real,
            "Trajectory": traj.tolist(),
            "RealTrajectory": real.tolist(),
            "WithdrawalRateTrajectory": wr.tolist(),
            "Inflation At Retirement": rec.inflation_at_ret,
        }

    # ---- batch on the device -----------------------------------------------------------------
    def _trajectory_len(self, working_months: int) -> int:
        full = (working_months + MONTHS_PER_YEAR - 1) // MONTHS_PER_YEAR if working_months > 0 else 0
        return 1 + full + self.params_model.retirement_years

    def _numpy_shocks_device(self, working_months: int, num_simulations: int):
        """rng='numpy': the reference's own draws, laid out [month][component][path] and uploaded."""
        import torch

        n_rows = max(working_months + self.params_model.retirement_years * MONTHS_PER_YEAR, 1)
        host = torch.empty((n_rows, 3, num_simulations), dtype=torch.float64, pin_memory=True)
        view = host.numpy()
        for i, seed in enumerate(self._path_seeds(num_simulations)):
            view[:, :, i] = self._draw_shock_path(n_rows, seed)
        return host.to(self._torch_device(), non_blocking=True), n_rows

    def run_batch_device(self, working_months: int, num_simulations: int, *, first_path: int = 0,
                         series: bool = True, shocks=None, _fast_replay: bool = False,
                         _small_returns: bool = False, on_chunk=None, chunks: int = 1,
                         first_fraction: Optional[float] = None) -> DeviceBatch:
        """One launch of the timeline kernel for `num_simulations` paths; everything stays in HBM.
        `shocks` (device tensor [n_months, 3, n]) forces replay of those draws (strict build;
        `_fast_replay` / `_small_returns` are the tests' handles on the fast build and on its
        MCR_FLAG_SMALL_RETURNS variant, see include/mcr.h).
        `chunks` > 1 (native draws only) splits the batch into that many launches over contiguous
        path ranges of the same output buffers and calls `on_chunk(batch, lo, count)` after each —
        the host-returning callers start a chunk's device-to-host copy while the next chunk is
        still being simulated. The results are identical (a path depends on its index only)."""
        import torch

        ctx = self.native_context
        dev = self._torch_device()
        n = int(num_simulations)
        wm = int(working_months)
        R = self.params_model.retirement_years
        T = self._trajectory_len(wm)
        f64 = dict(dtype=torch.float64, device=dev)
        cols = torch.empty((5, n), **f64)  # start, final, first-year gross, first-year real, inflation
        success = torch.empty(n, dtype=torch.uint8, device=dev)
        ruin = torch.empty(n, dtype=torch.int32, device=dev)
        counters = torch.zeros(2 + 12 * R + 1, dtype=torch.int64, device=dev)  # success, executed, ruin hist
        # series: True (all three), False (none), one of "traj" / "real" / "wr" or a tuple of them
        # (huge batches whose three series do not fit in HBM together are swept in several passes)
        if series is True:
            want = {"traj", "real", "wr"}
        elif not series:
            want = set()
        else:
            want = {series} if isinstance(series, str) else set(series)
        traj = torch.empty((T, n), **f64) if "traj" in want else None
        real = torch.empty((T, n), **f64) if "real" in want else None
        wr = torch.empty((R, n), **f64) if "wr" in want else None
        out = native.Outputs()
        out.start_balance = cols[0].data_ptr()
        out.final_balance = cols[1].data_ptr()
        out.first_year_gross = cols[2].data_ptr()
        out.first_year_real = cols[3].data_ptr()
        out.inflation_at_ret = cols[4].data_ptr()
        out.success = success.data_ptr()
        out.ruin_month = ruin.data_ptr()
        if traj is not None:
            out.trajectory = traj.data_ptr()
        if real is not None:
            out.real_trajectory = real.data_ptr()
        if wr is not None:
            out.wr_trajectory = wr.data_ptr()
        out.series_ld = n
        out.success_count = counters[0:].data_ptr()
        out.executed_months = counters[1:].data_ptr()
        out.ruin_month_hist = counters[2:].data_ptr()
        if shocks is None and self.rng_mode == "numpy":
            shocks, _ = self._numpy_shocks_device(wm, n)
        batch = DeviceBatch(n=n, working_months=wm, T=T, R=R, cols=cols, success=success, ruin=ruin,
                            counters=counters, traj=traj, real=real, wr=wr, shocks=shocks, ctx=ctx)
        if shocks is not None:
            ctx.replay(shocks, int(shocks.shape[2]), int(shocks.shape[0]), wm, n, out, strict=not _fast_replay,
                       small_returns=_small_returns)
            if on_chunk is not None:
                on_chunk(batch, 0, n)
            return batch
        chunks = max(1, min(int(chunks), n // 65536 or 1))
        if chunks == 1:
            ctx.simulate(self._seed_stream_id(), wm, int(first_path), n, out, strict=self.strict)
            if on_chunk is not None:
                on_chunk(batch, 0, n)
            return batch
        # contiguous path ranges, multiples of 1024 paths (whole CTAs, 128-byte-aligned rows)
        step = (-(-n // chunks) + 1023) // 1024 * 1024
        starts = list(range(0, n, step))
        if chunks == 2 and first_fraction is not None:   # uneven halves: the LAST chunk's copy is the exposed one
            starts = [0, min(n - 1024, max(1024, int(n * float(first_fraction)) // 1024 * 1024))]
        for i, lo in enumerate(starts):
            cnt = (starts[i + 1] if i + 1 < len(starts) else n) - lo
            part = native.Outputs()
            for name, t, width in (("start_balance", cols[0], 8), ("final_balance", cols[1], 8),
                                   ("first_year_gross", cols[2], 8), ("first_year_real", cols[3], 8),
                                   ("inflation_at_ret", cols[4], 8), ("success", success, 1), ("ruin_month", ruin, 4),
                                   ("trajectory", traj, 8), ("real_trajectory", real, 8), ("wr_trajectory", wr, 8)):
                if t is not None:
                    setattr(part, name, t.data_ptr() + lo * width)
            part.series_ld = n
            part.success_count, part.executed_months = out.success_count, out.executed_months
            part.ruin_month_hist = out.ruin_month_hist
            ctx.simulate(self._seed_stream_id(), wm, int(first_path) + lo, cnt, part, strict=self.strict)
            if on_chunk is not None:
                on_chunk(batch, lo, cnt)
        return batch

    def _staging(self, n: int, dev):
        """Pinned host destination buffers + the copy stream."""
        import torch

        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        # fresh pinned blocks per call (torch's caching host allocator recycles them once the
        # previous call's DataFrame is garbage): summary_df is built on these without a copy
        return {"n": n, "stream": self._copy_stream,
                "cols": torch.empty((6, n), dtype=torch.float64, pin_memory=True),  # 5 summary columns + YearsToRuin
                "succ": torch.empty(n, dtype=torch.uint8, pin_memory=True)}

    # ---- hooks a sharded (multi-GPU) subclass overrides; identity on one GPU --------------------
    def _shard(self, n_global: int) -> Tuple[int, int]:
        """(offset, count) of the global path range this process owns."""
        return 0, n_global

    def _agree_min(self, value: int, key=None) -> int:
        """A quantity every rank must plan with identically: the minimum over ranks (`key`:
        remember it for this call shape)."""
        return int(value)

    def _series_plan(self, n: int, T: int, R: int, bands: bool, key) -> List[Tuple[str, ...]]:
        """Which of the three yearly series each timeline pass materialises. One pass when they
        fit in HBM together (1.3 KB per path at C3); otherwise they are packed, in order, into as
        few recomputed passes as 70 % of the free memory (minimum over ranks) allows — e.g. 1.25e8
        paths x 71 points: [traj] then [real, wr]. Philox makes every (path, month) draw
        reproducible, so a recomputed pass costs kernel time, not memory or precision."""
        import torch

        if not bands:
            return [()]
        # decided once per call shape: cudaMemGetInfo takes 10-60 ms while kernels are running — on the host
        # thread that is supposed to keep the launch queue full (aggregates_device re-plans after an OOM)
        cache = self.__dict__.setdefault("_plan_cache", {})
        ckey = None if key is None else (key, os.environ.get("MCR_SERIES_BUDGET_BYTES"), os.environ.get("MCR_SERIES_SWEEP"))
        if ckey is not None and ckey in cache:
            return cache[ckey]
        sizes = {"traj": 8 * n * T, "real": 8 * n * T, "wr": 8 * n * R}
        free_bytes, _ = torch.cuda.mem_get_info(self._torch_device())
        budget = int(0.7 * free_bytes)
        if os.environ.get("MCR_SERIES_BUDGET_BYTES"):
            budget = int(os.environ["MCR_SERIES_BUDGET_BYTES"])
        if os.environ.get("MCR_SERIES_SWEEP") == "1":
            budget = 0  # one series per pass
        budget = self._agree_min(budget, key)
        plan: List[Tuple[str, ...]] = []
        cur: List[str] = []
        used = 0
        for name in ("traj", "real", "wr"):
            if cur and used + sizes[name] > budget:
                plan.append(tuple(cur))
                cur, used = [], 0
            cur.append(name)
            used += sizes[name]
        plan.append(tuple(cur))
        if ckey is not None:
            cache[ckey] = plan
        return plan

    def _select(self, specs, out16, counts=None, stepwise: bool = False):
        """Hook: all select rows of a step in one launch sequence. Returns None, or (several GPUs,
        ShardedSimulator) a 1-element device tensor that is non-zero when the pooled shortcut could
        not finish some row — identical on every rank; the caller then selects again with
        `stepwise=True` while the rows are still resident."""
        self.native_context.quantiles_rows(specs, out16, counts=counts)
        return None

    def _band_quantiles(self, b: DeviceBatch, bands, real_bands, wr_bands, wr_counts, stepwise: bool = False):
        """7-quantile nominal / real bands and 5-quantile NaN-skipping withdrawal-rate bands
        (simulation.py:1045-1118) of a batch, as ONE multi-row select."""
        import torch

        ctx = self.native_context
        T, R, n = b.T, b.R, b.n
        nq, nw = len(TRAJECTORY_QUANTILES), len(WITHDRAWAL_RATE_QUANTILES)
        specs = (ctx.series_rows(b.traj, n, T, TRAJECTORY_QUANTILES) + ctx.series_rows(b.real, n, T, TRAJECTORY_QUANTILES)
                 + ctx.series_rows(b.wr, n, R, WITHDRAWAL_RATE_QUANTILES))
        out16 = torch.empty((2 * T + R, 16), dtype=torch.float64, device=b.cols.device)
        cnt = torch.empty(2 * T + R, dtype=torch.int64, device=b.cols.device)
        flag = self._select(specs, out16, cnt, stepwise=stepwise)
        bands.view(T, nq).copy_(out16[:T, :nq])
        real_bands.view(T, nq).copy_(out16[T:2 * T, :nq])
        wr_bands.view(R, nw).copy_(out16[2 * T:, :nw])
        wr_counts.copy_(cnt[2 * T:])
        return flag

    def _sample_columns(self, n: int) -> List[int]:
        """Columns DataFrame.sample(n=5, axis=1, random_state=main_seed) picks
        (simulation.py:1063-1072): RandomState(seed).choice(n, 5, replace=False)."""
        if n not in self._sample_column_cache:
            k = min(n, 5)
            cols = np.random.RandomState(self.main_seed).choice(n, size=k, replace=False)
            self._sample_column_cache[n] = [int(c) for c in cols]
        return self._sample_column_cache[n]

    def _reduce_samples(self, block):
        """Hook for multi-GPU sharding: every rank fills the sample rows it owns, the rest stay zero."""
        return block

    def _gather_samples(self, series, n_local: int, T: int, first: int, n_global: int, out) -> None:
        """Copy the sampled global columns this process owns from a time-major [T][n_local] series
        into the rows of `out` ([k, T], zero-initialised)."""
        import torch

        mine = [(j, c - first) for j, c in enumerate(self._sample_columns(n_global)) if first <= c < first + n_local]
        if not mine:
            return
        rows = [j for j, _ in mine]
        if rows == list(range(out.shape[0])):
            self.native_context.gather_columns(series, n_local, T, [c for _, c in mine], out)
        else:
            tmp = torch.empty((len(mine), T), dtype=torch.float64, device=out.device)
            self.native_context.gather_columns(series, n_local, T, [c for _, c in mine], tmp)
            out[rows] = tmp

    def run_monte_carlo_simulations(self, working_months: int, num_simulations: int):
        """simulation.py:952-1128 — returns (summary_df, traj_pct_df, samples, wr_pct_df,
        real_pct_df, real_samples, wr_counts)."""
        import torch

        n = int(num_simulations)
        if n <= 0:
            return pd.DataFrame(columns=SUMMARY_COLUMNS), None, None, None, None, None, None
        ctx = self.native_context
        dev = self._torch_device()
        # the N x 7 summary columns travel to the host (copy engine, side stream) chunk by chunk:
        # a chunk's copy starts as soon as ITS timeline launch is done, underneath the next chunk's
        # launch and, for the last one, underneath the select kernels
        main = torch.cuda.current_stream()
        stage = self._staging(n, dev)
        years = torch.empty(n, dtype=torch.float64, device=dev)

        def copy_out(b, lo, cnt):
            b.years_to_ruin_slice(years, lo, cnt)
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(stage["stream"]):
                stage["stream"].wait_event(done)
                for col in range(5):   # contiguous row slices: plain async copies
                    stage["cols"][col, lo:lo + cnt].copy_(b.cols[col, lo:lo + cnt], non_blocking=True)
                stage["cols"][5, lo:lo + cnt].copy_(years[lo:lo + cnt], non_blocking=True)
                stage["succ"][lo:lo + cnt].copy_(b.success[lo:lo + cnt], non_blocking=True)

        b = self.run_batch_device(working_months, n, series=True, on_chunk=copy_out, chunks=self.e2e_chunks,
                                  first_fraction=self.e2e_first_fraction)
        T, R = b.T, b.R
        nq, nw = len(TRAJECTORY_QUANTILES), len(WITHDRAWAL_RATE_QUANTILES)
        sample_cols = self._sample_columns(n)
        k = len(sample_cols)
        years.record_stream(stage["stream"])

        # one small result block: bands (T*7 *2), WR bands (R*5), samples (k*T *2), WR observation counts (R)
        small = torch.empty(2 * T * nq + R * nw + 2 * k * T + R, dtype=torch.float64, device=dev)
        o = 0
        bands = small[o:o + T * nq]; o += T * nq
        real_bands = small[o:o + T * nq]; o += T * nq
        wr_bands = small[o:o + R * nw]; o += R * nw
        samples = small[o:o + k * T]; o += k * T
        real_samples = small[o:o + k * T]; o += k * T
        wr_counts = torch.empty(R, dtype=torch.int64, device=dev)
        self._band_quantiles(b, bands, real_bands, wr_bands, wr_counts)
        ctx.gather_columns(b.traj, n, T, sample_cols, samples)
        ctx.gather_columns(b.real, n, T, sample_cols, real_samples)
        small[o:o + R].copy_(wr_counts)          # (exact in float64) — one block, one copy, one wait
        host_small = torch.empty(small.numel(), dtype=torch.float64, pin_memory=True)
        host_small.copy_(small, non_blocking=True)
        done = torch.cuda.Event()
        done.record(main)
        self.last_d2h_bytes = n * (6 * 8 + 1) + host_small.numel() * 8

        # The frames below wrap page-locked host memory that the GPU is still filling: they are built while it
        # works (nothing here reads a value), and handed out only after both streams have drained.
        c = stage["cols"].numpy()  # pinned block owned by this call's result (no copy)
        summary_df = pd.DataFrame({
            "Start Balance": c[0],
            "Final Balance": c[1],
            "Success": stage["succ"].numpy().view(np.bool_),  # 0/1 bytes written by the kernel
            "YearsToRuin": c[5],
            "First Year Gross Withdrawal": c[2],
            "First Year Real Gross Withdrawal": c[3],
            "Inflation At Retirement": c[4],
        }, copy=False)
        s = host_small.numpy()      # (the frames keep this call's pinned block alive)
        o = 0
        traj_pct = pd.DataFrame(s[o:o + T * nq].reshape(T, nq), columns=TRAJECTORY_QUANTILES, copy=False); o += T * nq
        real_pct = pd.DataFrame(s[o:o + T * nq].reshape(T, nq), columns=TRAJECTORY_QUANTILES, copy=False); o += T * nq
        wr_pct = pd.DataFrame(s[o:o + R * nw].reshape(R, nw), columns=WITHDRAWAL_RATE_QUANTILES, copy=False); o += R * nw
        done.synchronize()
        stage["stream"].synchronize()
        sample_list = s[o:o + k * T].reshape(k, T).tolist(); o += k * T
        real_sample_list = s[o:o + k * T].reshape(k, T).tolist(); o += k * T
        wr_observation_counts = [int(v) for v in s[o:o + R]]
        self.last_executed_months = None  # read lazily: b.counters[1]
        self._last_batch = b
        return (summary_df, traj_pct, sample_list, wr_pct, real_pct, real_sample_list, wr_observation_counts)

    def _success_probability(self, summary_df: pd.DataFrame) -> float:
        """simulation.py:1130-1136."""
        if summary_df.empty:
            return 0.0
        if "Success" in summary_df.columns:
            return float(summary_df["Success"].astype(bool).mean() * 100.0)
        return float((summary_df["Final Balance"] > SMALL_EPSILON).mean() * 100.0)

    # ---- aggregate-only mode: nothing N-sized leaves the device ----------
    def aggregates_device(self, working_months: int, num_simulations: int, *, bands: bool = True,
                          first_path: int = 0, timeline_events=None, samples: bool = False,
                          pipeline: bool = False) -> "DeviceAggregates":
        """Enqueue one batch and every device-side reduction the callers make over summary_df
        (server.py:439-461,525-532; main.py:112-133; utils.py:97-102; plotting.py:46-59;
        HistogramChart.jsx:13-60). Nothing is copied to the host and nothing synchronises.
        `timeline_events=(start, end)` records CUDA events around the timeline kernel;
        `samples=True` (with bands) also keeps the 5 sampled nominal / real paths of
        simulation.py:1063-1078 — 2 x 5 x T values, the only per-path data the payload needs.
        `pipeline=True` enqueues the reductions on a second (high-priority) stream, so that a
        caller looping over batches gets the HBM-bound reductions of batch i underneath the
        issue-bound timeline kernel of batch i+1; the result then carries a `ready` event that
        `to_host()` / `wait()` honour."""
        import torch

        n_global = int(num_simulations)
        try:
            return self._aggregates_device(working_months, n_global, bands, first_path, timeline_events, samples, pipeline)
        except torch.cuda.OutOfMemoryError:
            # the cached series plan was made when more memory was free: plan again (single GPU only — ranks
            # of a sharded run must keep identical plans, they agreed on the minimum over ranks up front)
            if self._shard(n_global) != (0, n_global) or not self.__dict__.get("_plan_cache"):
                raise
            self._plan_cache.clear()
            self._last_batch = None
            torch.cuda.empty_cache()
            return self._aggregates_device(working_months, n_global, bands, first_path, timeline_events, samples, pipeline)

    def _aggregates_device(self, working_months, n_global, bands, first_path, timeline_events, samples, pipeline):
        import contextlib

        import torch

        offset, n = self._shard(n_global)  # single GPU: (0, n_global)
        plan = self._series_plan(n, self._trajectory_len(int(working_months)), self.params_model.retirement_years,
                                 bool(bands), key=(int(working_months), n_global, bool(bands)))
        self.last_series_plan = plan
        if timeline_events is not None:
            timeline_events[0].record()
        part_first = first_path + offset
        b = self.run_batch_device(working_months, n, series=(True if len(plan[0]) == 3 else (plan[0] or False)),
                                  first_path=part_first)
        if timeline_events is not None:
            timeline_events[1].record()
        side = None
        if pipeline:
            side = self._reduction_stream()
            produced = torch.cuda.Event()
            produced.record()
            side.wait_event(produced)
            for t in (b.cols, b.success, b.ruin, b.counters, b.traj, b.real, b.wr):
                if t is not None:
                    t.record_stream(side)
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            kw = dict(n_global=n_global, offset=offset, working_months=working_months, bands=bands, plan=plan,
                      samples=samples, part_first=part_first)
            agg = self._aggregate_batch(b, **kw)
            if side is not None:
                agg.ready = torch.cuda.Event()
                agg.ready.record(side)
        agg.reselect = (self, kw)
        return agg

    def _reduction_stream(self):
        import torch

        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream(device=self._torch_device(), priority=-1)
        return self._side_stream

    def _aggregate_batch(self, b: DeviceBatch, n_global: int, offset: int, working_months: int, bands: bool,
                         plan: List[Tuple[str, ...]], samples: bool, part_first: int,
                         stepwise: bool = False) -> "DeviceAggregates":
        """Every reduction of aggregates_device over one resident batch, on the current stream.
        `stepwise`: the re-select of DeviceAggregates.to_host() (several GPUs, rare) — the counters
        of `b` are global already and the pooled shortcut of the select is not taken."""
        import torch

        ctx = self.native_context
        n = b.n
        if not stepwise:
            self._reduce_counts(b.counters)
        dev = b.cols.device
        T, R = b.T, b.R
        f64 = dict(dtype=torch.float64, device=dev)
        nq, nw, nf = len(TRAJECTORY_QUANTILES), len(WITHDRAWAL_RATE_QUANTILES), len(FINAL_BALANCE_QUANTILES)
        rates = torch.empty(n, **f64)
        ctx.first_year_rates(b.cols[0], b.cols[3], n, rates)
        # every order statistic of the step in ONE multi-row select: 3 medians (different columns /
        # cohorts), the 9 final-balance quantiles and, when the series are resident, the bands
        # (row 4: min and max of the successful cohort's final balances — the ranges of the two histograms)
        specs = [(rates, n, None, [0.5], True), (b.cols[0], n, None, [0.5], True),
                 (b.cols[1], n, b.success, [0.5], True), (b.cols[1], n, None, FINAL_BALANCE_QUANTILES, False),
                 (b.cols[1], n, b.success, [0.0, 1.0], "minmax")]
        B0 = len(specs)  # first band row
        with_bands = bool(bands) and len(plan) == 1
        if with_bands:
            specs += (ctx.series_rows(b.traj, n, T, TRAJECTORY_QUANTILES) + ctx.series_rows(b.real, n, T, TRAJECTORY_QUANTILES)
                      + ctx.series_rows(b.wr, n, R, WITHDRAWAL_RATE_QUANTILES))
        desc = ctx.select_rows(specs)  # one descriptor per row (5 + 2T + R of them with the bands)
        out16 = torch.empty((len(desc), 16), **f64)
        cnt_all = torch.empty(len(desc), dtype=torch.int64, device=dev)
        flag = self._select(desc, out16, cnt_all, stepwise=stepwise)
        if flag is not None and len(plan) > 1:
            # series swept in several passes: later passes release this batch's series, so nothing is
            # deferred — every select of the call is verified (and repeated stepwise) on the spot
            if int(flag.item()) != 0:
                self.select_fallbacks += 1
                self._select(desc, out16, cnt_all, stepwise=True)
            flag = None
        # Nothing is rearranged on the device: to_host() reads the medians, quantiles and bands straight out of the
        # select's [rows][16] output. Histogram ranges: row 4 holds [min, max] of the successful final balances
        # in $; the $M histogram (plotting.py:46-59: final / 1e6) lets the kernel divide them by 1e6 itself — a
        # division by a positive constant is monotone, so min(x / 1e6) == min(x) / 1e6 bit for bit.
        hists = torch.zeros(160, dtype=torch.int64, device=dev)
        self._final_balance_histograms(b, out16[4, 0:2], hists)
        band_block = wr_counts = sample_block = None
        if bands and samples:
            sample_block = torch.zeros((2, len(self._sample_columns(n_global)), T), **f64)
        if bands:
            if with_bands:                      # (the bands are rows B0.. of out16)
                if sample_block is not None:
                    self._gather_samples(b.traj, n, T, offset, n_global, sample_block[0])
                    self._gather_samples(b.real, n, T, offset, n_global, sample_block[1])
            else:
                # the three series do not fit together: one multi-row select per pass of the plan;
                # the first pass's series came with the summary batch, the others are recomputed
                band_block = torch.empty(2 * T * nq + R * nw, **f64)
                wr_counts = torch.empty(R, dtype=torch.int64, device=dev)
                layout = {"traj": (T, TRAJECTORY_QUANTILES, 0), "real": (T, TRAJECTORY_QUANTILES, T * nq),
                          "wr": (R, WITHDRAWAL_RATE_QUANTILES, 2 * T * nq)}
                for k, group in enumerate(plan):
                    part = b if k == 0 else self.run_batch_device(working_months, n, series=group,
                                                                  first_path=part_first)
                    specs_g, n_rows = [], 0
                    for which in group:
                        rows, qs, _ = layout[which]
                        specs_g += ctx.series_rows(getattr(part, which), n, rows, qs)
                        n_rows += rows
                    o16 = torch.empty((n_rows, 16), **f64)
                    c16 = torch.empty(n_rows, dtype=torch.int64, device=dev)
                    desc_g = ctx.select_rows(specs_g)
                    flag_g = self._select(desc_g, o16, c16, stepwise=stepwise)
                    # these series are released before the next pass, so a select that could not
                    # finish is repeated NOW, while they are resident (the host sync below is needed
                    # anyway); the flag is identical on every rank
                    if flag_g is not None and int(flag_g.item()) != 0:
                        self.select_fallbacks += 1
                        self._select(desc_g, o16, c16, stepwise=True)
                    at = 0
                    for which in group:
                        rows, qs, off = layout[which]
                        band_block[off:off + rows * len(qs)].view(rows, len(qs)).copy_(o16[at:at + rows, :len(qs)])
                        if which == "wr":
                            wr_counts.copy_(c16[at:at + rows])
                        elif sample_block is not None:
                            self._gather_samples(getattr(part, which), n, T, offset, n_global,
                                                 sample_block[0 if which == "traj" else 1])
                        at += rows
                    torch.cuda.current_stream().synchronize()  # release these series before the next pass
                    for which in group:
                        setattr(part, which, None)
                    del part, specs_g
        self._last_batch = b
        if sample_block is not None:
            self._reduce_samples(sample_block)
        agg = DeviceAggregates(batch=b, out16=out16, cnt_all=cnt_all, first_band_row=(B0 if with_bands else None),
                               hists=hists, band_block=band_block, wr_counts=wr_counts, rates=rates,
                               sample_block=sample_block)
        agg.n_override = n_global
        agg.select_flag = flag
        return agg

    def _final_balance_histograms(self, b: DeviceBatch, rng_raw, hists) -> None:
        """100-bin numpy histogram in $M (plotting.py:46-59) and the dashboard's 60-bin floor
        histogram (HistogramChart.jsx:13-60) of the successful cohort's final balances."""
        ctx = self.native_context
        n = b.n
        ctx.histogram(b.cols[1], n, 100, rng_raw, hists[0:], mask=b.success, divisor=1e6,
                      mode=native.HIST_NUMPY | native.HIST_RAW_RANGE)
        ctx.histogram(b.cols[1], n, 60, rng_raw, hists[100:], mask=b.success, divisor=1.0, mode=native.HIST_FLOOR)

    def run_aggregates(self, working_months: int, num_simulations: int, *, bands: bool = True,
                       first_path: int = 0, samples: bool = False) -> Dict[str, Any]:
        """aggregates_device(...) copied to the host as a dict (a few KB)."""
        return self.aggregates_device(working_months, num_simulations, bands=bands, first_path=first_path,
                                      samples=samples).to_host()

    # ---- batched search ------------------------------------------------------------------------
    def batched_success_counts(self, candidates: Sequence[int], num_simulations: int, *, first_path: int = 0,
                               with_executed: bool = False):
        """Success counts for many working_months in ONE launch (mcr_search_batch)."""
        import torch

        ctx = self.native_context
        dev = self._torch_device()
        cand = [int(c) for c in candidates]
        counts = torch.zeros(len(cand), dtype=torch.int64, device=dev)
        executed = torch.zeros(len(cand), dtype=torch.int64, device=dev) if with_executed else None
        if cand:
            ctx.search_batch(self._seed_stream_id(), cand, int(first_path), int(num_simulations), counts,
                             executed=executed, strict=self.strict)
        if with_executed:
            return counts, executed
        return counts

    def _reduce_counts(self, counts):
        """Hook for multi-GPU sharding (parallel.ShardedSimulator overrides): identity here."""
        return counts

    # ---- multi-scenario batching --------------------------------------------
    def sweep_success_counts(self, scenarios: Sequence[Any], working_months: Union[int, Sequence[int]],
                             num_simulations: int, *, first_path: int = 0, with_executed: bool = False):
        """Success counts of MANY scenarios (Config objects: a parameter sweep / sensitivity grid)
        in one launch per kernel variant, every scenario on the SAME Philox streams as this
        simulator (its main seed and active seed stream): the points of the grid differ by their
        parameters, not by their luck, so differences between them are far less noisy than between
        independent runs. `working_months`: one value for all, or one per scenario. The reference
        has no counterpart (one scenario per process); a scenario's count equals what a simulator
        built for it with the same seed returns from `batched_success_counts`."""
        import torch

        scen = list(scenarios)
        wms = [int(working_months)] * len(scen) if isinstance(working_months, int) else [int(w) for w in working_months]
        if len(wms) != len(scen):
            raise ValueError("working_months must be one value or one per scenario")
        dev = self._torch_device()
        counts = torch.zeros(len(scen), dtype=torch.int64, device=dev)
        executed = torch.zeros(len(scen), dtype=torch.int64, device=dev) if with_executed else None
        if scen:
            lo, n = self._shard(int(num_simulations))
            self.native_context.sweep_batch(self._seed_stream_id(), [params_from_model(c) for c in scen], wms,
                                            int(first_path) + lo, n, counts, executed=executed, strict=self.strict)
            counts = self._reduce_counts(counts)
            if with_executed:
                executed = self._reduce_counts(executed)
        return (counts, executed) if with_executed else counts

    def sweep_success_probabilities(self, scenarios: Sequence[Any], working_months: Union[int, Sequence[int]],
                                    num_simulations: int) -> List[float]:
        """`_success_probability` (simulation.py:1130-1136) of every scenario of a sweep, in percent."""
        counts = self.sweep_success_counts(scenarios, working_months, num_simulations)
        return [float(c / int(num_simulations) * 100.0) for c in counts.cpu().tolist()]

    def find_minimum_working_months(self, verbose: bool = True,
                                    progress_callback: Optional[Callable[[dict], None]] = None
                                    ) -> Tuple[int, float, List[Dict[str, float]]]:
        """Bracket -> bisect -> month-by-month verification (simulation.py:1138-1342).

        The decisions, their order, the search_curve and the progress events are the
        reference's; what changes is how a probe's probability is obtained: unless
        `run_monte_carlo_simulations` has been replaced on the instance (the reference tests
        do that) probes are answered from success tables filled by the batched search kernel,
        which evaluates whole sets of candidates per launch.
        """
        self.use_search_seeds()
        p = self.params_model
        start = p.starting_working_months_search
        target = p.target_probability
        sim_count = p.num_simulations_search
        max_total_months = start + 70 * MONTHS_PER_YEAR
        search_curve: List[Dict[str, float]] = []
        cache: Dict[int, float] = {}
        state = {"iteration": 0, "best_seen": -1.0, "lo": start, "hi": None}

        patched = "run_monte_carlo_simulations" in self.__dict__ or (
            type(self).run_monte_carlo_simulations is not RetirementMonteCarloSimulator.run_monte_carlo_simulations
            and not getattr(self, "_device_search_ok", False))
        use_device_batches = not patched and self.search_policy != "sequential" and self.rng_mode == "philox"
        # auto: speculative waves while one candidate cannot fill the GPU (a launch per probe is
        # then latency-bound), one search launch per probe once a single candidate saturates it
        policy = self.search_policy
        if policy == "auto":
            policy = "waves" if sim_count < 131072 else "probe"
        speculate = use_device_batches and policy == "waves"
        table: Dict[int, int] = {}  # working_months -> success count (device-evaluated)
        stats = {"launches": 0, "candidates_evaluated": 0, "policy": policy if use_device_batches else "sequential"}

        def prefetch(months: Iterable[int]) -> None:
            if not use_device_batches:
                return
            todo = sorted({int(m) for m in months if start <= int(m) <= max_total_months and int(m) not in table})
            if not todo:
                return
            counts = self._reduce_counts(self.batched_success_counts(todo, sim_count))
            for m, c in zip(todo, counts.cpu().tolist()):
                table[m] = int(c)
            stats["launches"] += 1
            stats["candidates_evaluated"] += len(todo)

        def probability(months: int) -> float:
            if use_device_batches:
                if months not in table:
                    prefetch([months])
                return float(table[months] / sim_count * 100.0)
            summary_df, _, _, _, _, _, _ = self.run_monte_carlo_simulations(months, sim_count)
            return self._success_probability(summary_df)

        if verbose:
            logger.info(f"Estimating working months to achieve {target:.2f}% success for '{p.Nickname}'.")
            logger.info(f"Starting search from {start} months. Simulations per test: {sim_count}.")

        def _test(months: int) -> float:
            if months in cache:
                return cache[months]
            state["iteration"] += 1
            it = state["iteration"]
            if verbose:
                logger.info(f"Search iter {it}: Testing {months} m ({months / MONTHS_PER_YEAR:.1f} yrs) "
                            f"with {sim_count} sims.")
            prob = probability(months)
            cache[months] = prob
            if verbose:
                logger.info(f"  Search iter {it}: Prob for {months} m: {prob:.2f}% (Target: {target:.2f}%)")
            search_curve.append({"working_months": months, "working_years": round(months / MONTHS_PER_YEAR, 1),
                                 "probability": round(prob, 2)})
            if progress_callback:
                progress_callback({"type": "search_iter", "iteration": it, "working_months": months,
                                   "working_years": round(months / MONTHS_PER_YEAR, 1),
                                   "probability": round(prob, 2), "target": target, "sim_count": sim_count,
                                   "lo": state["lo"], "hi": state["hi"]})
            if prob > state["best_seen"]:
                state["best_seen"] = prob
            return prob

        def finish(result):
            self.last_search_stats = stats
            return result

        # ---- phase 1: bracket (simulation.py:1224-1279)
        if use_device_batches and policy == "grid":
            prefetch(range(start, min(start + 600, max_total_months) + 1))
        step = 12
        current = start
        prob_at_lo = _test(current)
        if prob_at_lo >= target:
            if verbose:
                logger.info(f"  Target met at starting point {current} months.")
            return finish((current, prob_at_lo, search_curve))

        # every bracket probe is start + a multiple of 12 (steps are 12 or 24 and never shrink):
        # evaluate the yearly grid in two waves, the cheap near half first.
        horizon = [start + 12 * k for k in range(1, 71)]
        waves = [horizon[:30], horizon[30:]]
        best_prob = None
        while current < max_total_months:
            gap = target - prob_at_lo
            if gap > 20:
                step = max(step, 24)
            elif gap > 10:
                step = max(step, 12)
            else:
                step = max(step, 6)
            next_months = min(current + step, max_total_months)
            if next_months <= current:
                break
            if speculate and next_months not in table:
                for w in waves:
                    if next_months in w:
                        prefetch(w)
            prob = _test(next_months)
            if prob >= target:
                state["lo"], state["hi"] = current, next_months
                best_prob = prob
                if verbose:
                    logger.info(f"  Bracketed: lo={current} m (miss), hi={next_months} m (hit). Bisecting…")
                if progress_callback:
                    progress_callback({"type": "search_refining", "working_months": next_months,
                                       "lo": current, "hi": next_months})
                break
            state["lo"] = next_months
            prob_at_lo = prob
            current = next_months

        if state["hi"] is None:
            if verbose:
                logger.warning(f"Search for '{p.Nickname}' reached max limit "
                               f"({max_total_months / MONTHS_PER_YEAR:.1f} yrs). Target NOT met.")
                logger.warning(f"Highest probability achieved: {state['best_seen']:.2f}%.")
            return finish((-1, state["best_seen"], search_curve))

        # ---- phases 2+3 need every month from a conservative verification start up to hi: the
        # start can only move later once bisection points are known, so one launch covers both.
        margin = min(100.0, 150.0 / math.sqrt(sim_count))
        if speculate:
            tested = sorted(m for m in cache if m <= state["hi"])
            near = next((i for i, m in enumerate(tested) if cache[m] >= target - margin), len(tested) - 1)
            prefetch(range(max(start, tested[max(0, near - 1)]), state["hi"] + 1))

        # ---- phase 2: bisect (simulation.py:1281-1291)
        lo, hi = state["lo"], state["hi"]
        best = hi
        while hi - lo > 1:
            mid = (lo + hi) // 2
            prob = _test(mid)
            if prob >= target:
                best, best_prob, hi = mid, prob, mid
            else:
                lo = mid
            state["lo"], state["hi"] = lo, hi

        # ---- phase 3: verify the statistically plausible transition region (simulation.py:1293-1335)
        tested_before_best = sorted(m for m in cache if m <= best)
        near_target_index = next((i for i, m in enumerate(tested_before_best) if cache[m] >= target - margin),
                                 len(tested_before_best) - 1)
        verification_start = max(start, tested_before_best[max(0, near_target_index - 1)])
        if verbose:
            logger.info(f"  Verifying each month from {verification_start} to {best} "
                        "to handle locally non-monotone Monte Carlo estimates.")
        for month in range(verification_start, best + 1):
            _test(month)
        qualifying = [m for m, pr in cache.items() if start <= m <= best and pr >= target]
        if qualifying:
            best = min(qualifying)
            best_prob = cache[best]
        if verbose:
            logger.info(f"  Search complete: estimated minimum {best} months ({best / MONTHS_PER_YEAR:.1f} yrs) "
                        f"with prob {best_prob:.2f}%.")
        return finish((best, best_prob, search_curve))
