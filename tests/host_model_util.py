"""Builds and binds tests/host_model (TEST INFRASTRUCTURE ONLY): the engine's device headers
compiled for the host with g++, so the CPU suite can check the strict / fast / lean month
arithmetic against the oracle without a GPU. The product never loads this library."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host_model", "path_host_model.cpp")
OUT = os.path.join(HERE, "host_model", "_build", "libpath_host_model.so")
CSRC = os.path.join(ROOT, "monte_carlo_retirement_b200", "csrc")
_lib = None


def build(force: bool = False) -> str:
    deps = [SRC, os.path.join(ROOT, "include", "mcr.h")] + [os.path.join(CSRC, f) for f in
                                                          ("mcr_path.cuh", "mcr_rng.cuh", "mcr_derive.h", "mcr_portable.h")]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # -ffp-contract=off: only the explicit fma() calls of the source fuse, as in the strict device build
    cmd = ["g++", "-O2", "-std=c++17", "-march=native", "-ffp-contract=off", "-fPIC", "-shared",
           "-I", os.path.join(ROOT, "include"), SRC, "-o", OUT + ".tmp"]
    subprocess.run(cmd, check=True)
    os.replace(OUT + ".tmp", OUT)
    return OUT


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


STRICT, FAST, FAST_SMALL = 0, 1, 2


def replay(params, wm: int, shocks_nrc: np.ndarray, mode: int):
    """shocks_nrc: (n, rows, 3) as the oracle takes them. Returns a dict like gpu_util.device_batch_to_host."""
    n, rows, _ = shocks_nrc.shape
    dev_layout = np.ascontiguousarray(shocks_nrc.transpose(1, 2, 0))  # [rows, 3, n] == the device replay layout
    R = int(params.retirement_years)
    T = 1 + ((wm + 11) // 12 if wm > 0 else 0) + R
    cols = np.empty((5, n))
    success = np.empty(n, dtype=np.uint8)
    ruin = np.empty(n, dtype=np.int32)
    executed = np.empty(n, dtype=np.uint32)
    traj = np.empty((T, n))
    real = np.empty((T, n))
    wr = np.empty((R, n))
    cfg = C.c_int32(-1)
    lib().hm_lean_months()  # reset
    rc = lib().hm_replay(C.byref(params), C.c_int32(wm), C.c_void_p(dev_layout.ctypes.data), C.c_int64(n),
                         C.c_int32(rows), C.c_int64(n), C.c_int(mode), C.c_void_p(cols.ctypes.data),
                         C.c_void_p(success.ctypes.data), C.c_void_p(ruin.ctypes.data),
                         C.c_void_p(executed.ctypes.data), C.c_void_p(traj.ctypes.data), C.c_void_p(real.ctypes.data),
                         C.c_void_p(wr.ctypes.data), C.c_int64(n), C.byref(cfg))
    if rc != 0:
        raise ValueError("hm_replay: bad parameters")
    lib().hm_lean_months.restype = C.c_longlong
    lean = int(lib().hm_lean_months())
    return {"lean_months": lean, "start": cols[0], "final": cols[1], "fy_gross": cols[2], "fy_real": cols[3], "infl": cols[4],
            "success": success.astype(bool), "ruin_month": ruin, "executed": executed,
            "traj": traj.T.copy(), "real": real.T.copy(), "wr": wr.T.copy(), "cfg": int(cfg.value)}


def draw(params, main_seed: int, stream: int, first_path: int, n: int, n_months: int) -> np.ndarray:
    """The engine's native draw stream, [n_months, 3, n] (strict Box-Muller transform)."""
    out = np.empty((n_months, 3, n))
    rc = lib().hm_draw(C.byref(params), C.c_uint64(main_seed), C.c_uint32(stream), C.c_int64(first_path), C.c_int64(n),
                       C.c_int32(n_months), C.c_void_p(out.ctypes.data), C.c_int64(n))
    if rc != 0:
        raise ValueError("hm_draw: bad parameters")
    return out


def small_bound(params) -> float:
    """Bound on |monthly log-return| the engine proves for this scenario's own Philox draws (0: none)."""
    lib().hm_small_bound.restype = C.c_double
    return float(lib().hm_small_bound(C.byref(params)))


def small_returns(params, shocks_nrc: np.ndarray, bound=None) -> bool:
    """|mu_log/12 + sigma_log/sqrt(12) * z| < bound for every supplied shock of the three factors
    (the guarantee MCR_FLAG_SMALL_RETURNS asks the caller for; bound = the scenario's own class)."""
    bound = small_bound(params) if bound is None else bound
    if not bound > 0:
        return False
    r12 = np.sqrt(12.0)
    worst = 0.0
    for c, (mu, sg) in enumerate(((params.inv1_mu_log, params.inv1_sigma_log), (params.inf_mu_log, params.inf_sigma_log),
                                  (params.prem_mu_log, params.prem_sigma_log))):
        worst = max(worst, float(np.max(np.abs(mu / 12.0 + sg / r12 * shocks_nrc[:, :, c]))))
    return worst < bound
