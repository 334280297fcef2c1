"""ctypes binding of libmcr_b200.so — the thin layer between the Python mirror of the
reference's `simulation` module and the hand-written sm_100a CUDA engine (include/mcr.h).

There is NO CPU fallback: loading fails loudly when the library is missing and every compute
call raises when no CUDA device is usable. PyTorch is used only to own device buffers and
streams; pointers cross the ABI as plain integers.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Any, Optional, Sequence

MAX_STREAMS = 16
FLAG_STRICT = 0x1
FLAG_SMALL_RETURNS = 0x2
STREAM_SEARCH = 0
STREAM_FINAL = 1
SEL_MEDIAN = 0x1
SEL_MINMAX = 0x2  # specs: median="minmax"
HIST_NUMPY = 0
HIST_RAW_RANGE = 0x100  # range2 holds min / max of the undivided values
HIST_FLOOR = 1
E_INVAL = -1

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "_lib", "libmcr_b200.so")


class IncomeStream(C.Structure):
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
        ("streams", IncomeStream * MAX_STREAMS),
    ]


class Outputs(C.Structure):
    _fields_ = [
        ("start_balance", C.c_void_p),
        ("final_balance", C.c_void_p),
        ("success", C.c_void_p),
        ("ruin_month", C.c_void_p),
        ("first_year_gross", C.c_void_p),
        ("first_year_real", C.c_void_p),
        ("inflation_at_ret", C.c_void_p),
        ("trajectory", C.c_void_p),
        ("real_trajectory", C.c_void_p),
        ("wr_trajectory", C.c_void_p),
        ("series_ld", C.c_int64),
        ("success_count", C.c_void_p),
        ("wr_obs_count", C.c_void_p),
        ("ruin_month_hist", C.c_void_p),
        ("executed_months", C.c_void_p),
    ]


class SelectRow(C.Structure):
    _fields_ = [
        ("values_dev", C.c_void_p),
        ("mask_dev", C.c_void_p),
        ("n", C.c_int64),
        ("n_q", C.c_int32),
        ("flags", C.c_uint32),
        ("q", C.c_double * 16),
    ]


def _select_row_dtype():
    import numpy as np

    return np.dtype([("values_dev", "<u8"), ("mask_dev", "<u8"), ("n", "<i8"), ("n_q", "<i4"), ("flags", "<u4"),
                     ("q", "<f8", (16,))])


SELECT_ROW_DTYPE = _select_row_dtype()
assert SELECT_ROW_DTYPE.itemsize == C.sizeof(SelectRow)


class PathRecord(C.Structure):
    _fields_ = [
        ("start_balance", C.c_double),
        ("final_balance", C.c_double),
        ("first_year_gross", C.c_double),
        ("first_year_real", C.c_double),
        ("inflation_at_ret", C.c_double),
        ("success", C.c_int32),
        ("ruin_month", C.c_int32),
        ("trajectory_len", C.c_int32),
        ("wr_len", C.c_int32),
    ]


# every symbol include/mcr.h declares: (restype, argtypes)
_VP, _I32, _I64, _U32, _U64, _D = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double
SIGNATURES = {
    "mcr_abi_version": (C.c_int, []),
    "mcr_create": (C.c_int, [C.POINTER(Params), _U64, C.c_int, C.POINTER(_VP)]),
    "mcr_destroy": (C.c_int, [_VP]),
    "mcr_last_error": (C.c_char_p, [_VP]),
    "mcr_launch_count": (_I64, [_VP]),
    "mcr_small_returns_bound": (_D, [_VP]),
    "mcr_last_variant": (_I32, [_VP]),
    "mcr_stream_start_month": (_I32, [_D, _I32, _D]),
    "mcr_trajectory_len": (_I32, [_I32, _I32]),
    "mcr_simulate": (C.c_int, [_VP, C.c_int, _I32, _I64, _I64, _U32, C.POINTER(Outputs), _VP]),
    "mcr_replay": (C.c_int, [_VP, _VP, _I64, _I32, _I32, _I64, _U32, C.POINTER(Outputs), _VP]),
    "mcr_single_path": (C.c_int, [_VP, _I32, _VP, _I32, C.POINTER(PathRecord), _VP, _VP, _VP]),
    "mcr_helper_withdraw": (C.c_int, [_VP, _D, _D, _D, _I32, _D, C.POINTER(_D)]),
    "mcr_helper_net_liquidation": (C.c_int, [_VP, _D, _D, _I32, _D, C.POINTER(_D)]),
    "mcr_helper_rebalance": (C.c_int, [_VP, _D, _D, _D, _D, C.POINTER(_D)]),
    "mcr_helper_annual_tax": (C.c_int, [_VP, _D, _D, _D, _D, _D, _D, C.POINTER(_D)]),
    "mcr_draw_shocks": (C.c_int, [_VP, C.c_int, _I64, _I64, _I32, _U32, _VP, _I64, _VP]),
    "mcr_search_batch": (C.c_int, [_VP, C.c_int, C.POINTER(_I32), _I32, _I64, _I64, _U32, _VP, _VP, _VP]),
    "mcr_sweep_batch": (C.c_int, [_VP, C.c_int, C.POINTER(Params), C.POINTER(_I32), _I32, _I64, _I64, _U32, _VP, _VP, _VP]),
    "mcr_quantiles": (C.c_int, [_VP, _VP, _I64, _I64, _I32, _VP, C.POINTER(_D), _I32, _U32, _VP, _VP, _VP]),
    "mcr_select_state_bytes": (_I64, [_I32]),
    "mcr_select_hist_bytes": (_I64, [_I32]),
    "mcr_select_full_passes": (_I32, []),
    "mcr_select_full_passes_for": (_I32, [_I64]),
    "mcr_select_exchange_words": (_I64, [_I32, _I32]),
    "mcr_select_exchange_layout": (None, [_I32, _I32, C.POINTER(_I64)]),
    "mcr_quantiles_rows": (C.c_int, [_VP, C.POINTER(SelectRow), _I32, _VP, _VP, _VP]),
    "mcr_select_step": (C.c_int, [_VP, _I32, _I32, C.POINTER(SelectRow), _I32, _VP, _VP, _VP, _VP, _VP]),
    "mcr_first_year_rates": (C.c_int, [_VP, _VP, _VP, _I64, _VP, _VP]),
    "mcr_years_to_ruin": (C.c_int, [_VP, _VP, _I64, _VP, _VP]),
    "mcr_minmax": (C.c_int, [_VP, _VP, _VP, _I64, _D, _VP, _VP]),
    "mcr_histogram": (C.c_int, [_VP, _VP, _VP, _I64, _D, _I32, _I32, _VP, _VP, _VP]),
    "mcr_gather_columns": (C.c_int, [_VP, _VP, _I64, _I32, C.POINTER(_I64), _I32, _VP, _VP]),
    "mcr_fp64_peak_slots_per_s": (C.c_int, [_VP, C.POINTER(_D)]),
    "mcr_comm_create": (C.c_int, [_VP, _I64, C.POINTER(_VP)]),
    "mcr_comm_connect": (C.c_int, [_VP, _I32, _I32, C.POINTER(_VP)]),
    "mcr_comm_all_reduce": (C.c_int, [_VP, _I32, _VP, _I64, _VP]),
    "mcr_quantiles_rows_comm": (C.c_int, [_VP, _VP, _I32, _I32, C.POINTER(SelectRow), _I32, _VP, _VP, _VP, _VP]),
    "mcr_comm_status": (_I32, [_VP]),
    "mcr_comm_calls": (_I64, [_VP]),
    "mcr_comm_last_error": (C.c_char_p, [_VP]),
    "mcr_comm_destroy": (C.c_int, [_VP]),
}

_lib = None
_lib_lock = threading.Lock()


def load_library(path: Optional[str] = None) -> C.CDLL:
    """dlopen the engine and bind every exported symbol. Raises if it is not built."""
    global _lib
    with _lib_lock:
        if _lib is not None and path is None:
            return _lib
        p = path or os.environ.get("MCR_LIB") or LIB_PATH  # MCR_LIB: an alternative build (tuning experiments)
        if not os.path.exists(p):
            raise RuntimeError(
                f"{p} is missing: build it with `python -m monte_carlo_retirement_b200.build` "
                "(this engine is CUDA-only; there is no CPU fallback)")
        lib = C.CDLL(p)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        if lib.mcr_abi_version() != 1:
            raise RuntimeError("libmcr_b200.so ABI version mismatch")
        if path is None:
            _lib = lib
        return lib


def _ptr(t: Any) -> Optional[int]:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return int(t.data_ptr())


def _stream_handle() -> Optional[int]:
    import torch

    return int(torch.cuda.current_stream().cuda_stream) or None


class NativeError(RuntimeError):
    pass


class Context:
    """One mcr_ctx (one simulator instance, one device)."""

    def __init__(self, params: Params, main_seed: int, device: int = 0):
        self.lib = load_library()
        self.params = params
        self.device = device
        h = _VP()
        rc = self.lib.mcr_create(C.byref(params), C.c_uint64(main_seed & 0xFFFFFFFFFFFFFFFF), device, C.byref(h))
        if rc != 0:
            msg = self.lib.mcr_last_error(None).decode()
            raise (ValueError if rc == E_INVAL else NativeError)(msg)
        self.handle = h

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.mcr_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> None:
        if rc != 0:
            msg = self.lib.mcr_last_error(self.handle).decode()
            raise (ValueError if rc == E_INVAL else NativeError)(msg)

    @property
    def launch_count(self) -> int:
        return int(self.lib.mcr_launch_count(self.handle))

    @property
    def last_variant(self) -> int:
        return int(self.lib.mcr_last_variant(self.handle))

    @property
    def small_returns_bound(self) -> float:
        return float(self.lib.mcr_small_returns_bound(self.handle))

    # ---- timeline -------------------------------------------------------------------------
    def simulate(self, seed_stream: int, working_months: int, first_path: int, n_paths: int, out: Outputs,
                 strict: bool = False) -> None:
        self._check(self.lib.mcr_simulate(self.handle, seed_stream, working_months, first_path, n_paths,
                                          FLAG_STRICT if strict else 0, C.byref(out), _stream_handle()))

    def replay(self, shocks, shocks_ld: int, n_months: int, working_months: int, n_paths: int, out: Outputs,
               strict: bool = True, small_returns: bool = False) -> None:
        flags = (FLAG_STRICT if strict else 0) | (FLAG_SMALL_RETURNS if small_returns and not strict else 0)
        self._check(self.lib.mcr_replay(self.handle, _ptr(shocks), shocks_ld, n_months, working_months, n_paths,
                                        flags, C.byref(out), _stream_handle()))

    def single_path(self, working_months: int, shocks_np):
        import numpy as np

        sh = np.ascontiguousarray(shocks_np, dtype=np.float64)
        if sh.ndim != 2 or sh.shape[1] != 3:
            raise ValueError("shocks must have shape (n_months, 3)")
        R = int(self.params.retirement_years)
        T = int(self.lib.mcr_trajectory_len(working_months, R))
        traj = np.empty(T)
        real = np.empty(T)
        wr = np.empty(R)
        rec = PathRecord()
        self._check(self.lib.mcr_single_path(self.handle, working_months, sh.ctypes.data, sh.shape[0], C.byref(rec),
                                             traj.ctypes.data, real.ctypes.data, wr.ctypes.data))
        return rec, traj, real, wr

    def helper_withdraw(self, bal, cb, target, use_tax, rate):
        out = (_D * 4)()
        self._check(self.lib.mcr_helper_withdraw(self.handle, bal, cb, target, int(bool(use_tax)), rate, out))
        return tuple(out)

    def helper_net_liquidation(self, bal, cb, use_tax, rate) -> float:
        out = _D()
        self._check(self.lib.mcr_helper_net_liquidation(self.handle, bal, cb, int(bool(use_tax)), rate, C.byref(out)))
        return out.value

    def helper_rebalance(self, b1, cb1, b2, cb2):
        out = (_D * 4)()
        self._check(self.lib.mcr_helper_rebalance(self.handle, b1, cb1, b2, cb2, out))
        return tuple(out)

    def helper_annual_tax(self, b1, cb1, b2, cb2, gain1, gain2):
        out = (_D * 5)()
        self._check(self.lib.mcr_helper_annual_tax(self.handle, b1, cb1, b2, cb2, gain1, gain2, out))
        return out[0], out[1], out[2], out[3], bool(out[4])

    def draw_shocks(self, seed_stream: int, first_path: int, n_paths: int, n_months: int, shocks, shocks_ld: int,
                    strict: bool = False) -> None:
        self._check(self.lib.mcr_draw_shocks(self.handle, seed_stream, first_path, n_paths, n_months,
                                             FLAG_STRICT if strict else 0, _ptr(shocks), shocks_ld, _stream_handle()))

    # ---- search ---------------------------------------------------------------------------
    def search_batch(self, seed_stream: int, candidates: Sequence[int], first_path: int, n_paths: int, counts,
                     executed=None, strict: bool = False) -> None:
        arr = (_I32 * len(candidates))(*[int(c) for c in candidates])
        self._check(self.lib.mcr_search_batch(self.handle, seed_stream, arr, len(candidates), first_path, n_paths,
                                              FLAG_STRICT if strict else 0, _ptr(counts), _ptr(executed),
                                              _stream_handle()))

    def sweep_batch(self, seed_stream: int, scenarios: Sequence[Params], working_months: Sequence[int], first_path: int,
                    n_paths: int, counts, executed=None, strict: bool = False) -> None:
        n = len(scenarios)
        arr = (Params * n)(*scenarios)
        wm = (_I32 * n)(*[int(w) for w in working_months])
        self._check(self.lib.mcr_sweep_batch(self.handle, seed_stream, arr, wm, n, first_path, n_paths,
                                             FLAG_STRICT if strict else 0, _ptr(counts), _ptr(executed), _stream_handle()))

    # ---- aggregations -----------------------------------------------------------------------
    def quantiles(self, values, n: int, ld: int, rows: int, q: Sequence[float], out, counts=None, mask=None,
                  median: bool = False) -> None:
        qs = (_D * len(q))(*[float(v) for v in q])
        self._check(self.lib.mcr_quantiles(self.handle, _ptr(values), n, ld, rows, _ptr(mask), qs, len(q),
                                           SEL_MINMAX if median == "minmax" else (SEL_MEDIAN if median else 0), _ptr(out),
                                           _ptr(counts), _stream_handle()))

    @staticmethod
    def select_rows(specs):
        """Row descriptors (mcr_select_row[]) as one numpy structured array, filled group-wise.
        A spec is (tensor, n, mask_or_None, quantiles, median): a 1-D tensor is one row, a 2-D
        time-major [rows, ld] tensor contributes one row per tensor row."""
        import numpy as np

        groups = []
        total = 0
        for x, n, mask, q, median in specs:
            rows = int(x.shape[0]) if (x is not None and x.dim() == 2) else 1
            groups.append((x, int(n), mask, q, median, rows))
            total += rows
        arr = np.zeros(total, dtype=SELECT_ROW_DTYPE)
        at = 0
        for x, n, mask, q, median, rows in groups:
            sl = arr[at:at + rows]
            if n > 0:
                base = int(x.data_ptr())
                stride = int(x.stride(0)) * 8 if x.dim() == 2 else 0
                sl["values_dev"] = base + stride * np.arange(rows, dtype=np.uint64)
            sl["mask_dev"] = 0 if mask is None else int(mask.data_ptr())
            sl["n"] = n
            sl["n_q"] = len(q)
            sl["flags"] = SEL_MINMAX if median == "minmax" else (SEL_MEDIAN if median else 0)
            sl["q"][:, :len(q)] = np.asarray(q, dtype=np.float64)
            at += rows
        return arr

    @staticmethod
    def series_rows(series, n: int, rows: int, q, mask=None, median: bool = False):
        """Select spec for the rows of a time-major [rows, ld] series tensor."""
        return [(series[:rows], n, mask, q, median)]

    def quantiles_rows(self, specs, out, counts=None, all_reduce=None, all_reduce_min=None, rank=None,
                       world=None, defer_check: bool = False):
        """All rows of `specs` in one launch sequence; out is a [n_rows, 16] f64 device tensor.
        With `all_reduce` (sums an integer device tensor in place across ranks) the rows are this
        rank's shards and every rank obtains the exact GLOBAL quantiles; `all_reduce_min`
        (element-wise MIN of an int64 device tensor) additionally enables the adaptive start,
        and with `rank` / `world` the pooled tail (mcr.h: MCR_SELECT_POOL_*). The pooled tail
        reports rows it could not finish (rare: a big bucket of distinct values); by default
        that count is read back here (one host sync) and the stepwise protocol re-run. With
        `defer_check` the count is RETURNED as a 1-element device tensor instead and the caller
        checks it at its next natural sync point (non-zero: call again without rank/world)."""
        import torch

        arr_np = specs if hasattr(specs, "dtype") else self.select_rows(specs)
        n_rows = len(arr_np)
        if n_rows == 0:
            return None
        arr = C.cast(arr_np.ctypes.data, C.POINTER(SelectRow))
        if all_reduce is None:
            self._check(self.lib.mcr_quantiles_rows(self.handle, arr, n_rows, _ptr(out), _ptr(counts), _stream_handle()))
            return None
        dev = out.device
        state = torch.empty(int(self.lib.mcr_select_state_bytes(n_rows)), dtype=torch.uint8, device=dev)
        hist = torch.empty(int(self.lib.mcr_select_hist_bytes(n_rows)) // 4, dtype=torch.int32, device=dev)

        def step(kind, p=0, buf=None):
            self._check(self.lib.mcr_select_step(self.handle, kind, p, arr, n_rows, _ptr(state), _ptr(hist),
                                                 _ptr(out if buf is None else buf), _ptr(counts), _stream_handle()))

        # global row length: the shards are balanced (shard_range), so local max x world
        n_max = int(arr_np["n"].max()) * int(world or 1)
        full = int(self.lib.mcr_select_full_passes_for(n_max))
        adaptive = all_reduce_min is not None
        if adaptive and world is not None:
            # pooled tail: `full` all-reduced passes at most, then the ranks pool their candidates
            at4 = (_I64 * 4)()
            self.lib.mcr_select_exchange_layout(n_rows, int(world), at4)
            at = [int(v) for v in at4]
            xbuf = torch.empty(int(at[3]), dtype=torch.int64, device=dev)
            who = int(rank) | (int(world) << 8)

            def pool_step(kind):
                self._check(self.lib.mcr_select_step(self.handle, kind, who, arr, n_rows, _ptr(state), _ptr(xbuf),
                                                     _ptr(out), _ptr(counts), _stream_handle()))

            step(0, 3)
            for p in range(full):
                step(1, p | (0x100 if p == 0 else 0))   # pass 0: only the sample CTAs
                if p == 0:
                    ext = torch.empty((n_rows, 2), dtype=torch.int64, device=dev)
                    step(5, 0, ext)
                    all_reduce_min(ext)
                    step(6, 0, ext)
                all_reduce(hist)
                step(2, p)
            step(4)
            pool_step(7)
            all_reduce(xbuf[at[0]:at[1]])
            all_reduce_min(xbuf[at[1]:at[2]])
            pool_step(8)
            all_reduce(xbuf[at[2]:at[3]])
            pool_step(9)
            if defer_check:
                return xbuf[0:1].clone()
            if int(xbuf[0].item()) == 0:  # identical on every rank: all of them saw the same pool
                return None
        step(0, 1 if adaptive else 0)
        # 8 digits of 8 bits resolve any key; the adaptive start spends pass 0 on the row extremes
        # and may give up its prefix after pass 1 (mcr_reduce.cu: advance_row), hence 2 more
        for p in range(10 if adaptive else 8):
            if p == full:
                step(4)  # COLLECT: candidates of the resolved prefixes (local shard)
            step(1, p | (0x100 if (p == 0 and adaptive) else 0))
            if p == 0 and adaptive:  # global row extremes -> skip the key bits every element shares
                ext = torch.empty((n_rows, 2), dtype=torch.int64, device=dev)
                step(5, 0, ext)
                all_reduce_min(ext)
                step(6, 0, ext)
            all_reduce(hist)
            step(2, p)
        step(3)
        return None

    def quantiles_rows_comm(self, comm_handle, rank: int, world: int, specs, out, counts=None):
        """The pooled distributed select in one call (several GPUs of this process, csrc/mcr_comm.cu).
        Returns the 1-element device tensor "rows the pool could not finish"."""
        import torch

        arr_np = specs if hasattr(specs, "dtype") else self.select_rows(specs)
        arr = C.cast(arr_np.ctypes.data, C.POINTER(SelectRow))
        flag = torch.empty(1, dtype=torch.int64, device=out.device)
        self._check(self.lib.mcr_quantiles_rows_comm(self.handle, comm_handle, rank, world, arr, len(arr_np), _ptr(out),
                                                     _ptr(counts), _ptr(flag), _stream_handle()))
        return flag

    def quantiles_distributed(self, values, n: int, ld: int, rows: int, q: Sequence[float], out, all_reduce,
                              counts=None, mask=None, median: bool = False) -> None:
        """Uniform-rows convenience over quantiles_rows(all_reduce=...): out is [rows, len(q)]."""
        import torch

        specs = [(values[:rows] if values.dim() == 2 else values, n, mask, q, median)]
        tmp = torch.empty((rows, 16), dtype=torch.float64, device=out.device)
        self.quantiles_rows(specs, tmp, counts=counts, all_reduce=all_reduce)
        out.view(-1)[: rows * len(q)].view(rows, len(q)).copy_(tmp[:, : len(q)])

    def first_year_rates(self, start, fy_real, n: int, rates) -> None:
        self._check(self.lib.mcr_first_year_rates(self.handle, _ptr(start), _ptr(fy_real), n, _ptr(rates),
                                                  _stream_handle()))

    def years_to_ruin(self, ruin, n: int, years) -> None:
        self._check(self.lib.mcr_years_to_ruin(self.handle, _ptr(ruin), n, _ptr(years), _stream_handle()))

    def minmax(self, values, n: int, out2, mask=None, divisor: float = 1.0) -> None:
        self._check(self.lib.mcr_minmax(self.handle, _ptr(values), _ptr(mask), n, divisor, _ptr(out2),
                                        _stream_handle()))

    def histogram(self, values, n: int, n_bins: int, range2, hist, mask=None, divisor: float = 1.0,
                  mode: int = HIST_NUMPY) -> None:
        self._check(self.lib.mcr_histogram(self.handle, _ptr(values), _ptr(mask), n, divisor, n_bins, mode,
                                           _ptr(range2), _ptr(hist), _stream_handle()))

    def gather_columns(self, series, ld: int, rows: int, cols: Sequence[int], out) -> None:
        arr = (_I64 * len(cols))(*[int(c) for c in cols])
        self._check(self.lib.mcr_gather_columns(self.handle, _ptr(series), ld, rows, arr, len(cols), _ptr(out),
                                                _stream_handle()))

    def fp64_peak_slots_per_s(self) -> float:
        out = _D()
        self._check(self.lib.mcr_fp64_peak_slots_per_s(self.handle, C.byref(out)))
        return out.value
