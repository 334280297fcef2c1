"""numpy restatement of the engine's native draw stream (csrc/mcr_rng.cuh): Philox4x32-10
(Salmon et al., SC'11; checked against the Random123 known-answer vectors in the tests) keyed by
splitmix64(main_seed), one call per TWO months, three Box-Muller pairs per call."""
from __future__ import annotations

import math

import numpy as np


def philox4x32_10(c, k):
    """Philox4x32-10 on arrays of counters c = (c0, c1, c2, c3), key k = (k0, k1)."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
    c0, c1, c2, c3 = [np.asarray(x).astype(np.uint32) for x in c]
    k0 = np.uint32(k[0])
    k1 = np.uint32(k[1])
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
            c1 = p1.astype(np.uint32)
            c3 = p0.astype(np.uint32)
            c0, c2 = n0, n2
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return x ^ (x >> 31)


def key_from_seed(main_seed: int):
    key = splitmix64(splitmix64(int(main_seed)) ^ 0x6D63725F62323030)  # "mcr_b200"
    return key & 0xFFFFFFFF, key >> 32


def _box_muller(rad26, ang, bits):
    u1 = (rad26.astype(np.float64) + 0.5) / 2.0 ** 26
    th = (ang.astype(np.float64) + 0.5) / 2.0 ** (bits - 1) - 1.0   # theta / pi
    r = np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(np.pi * th), r * np.sin(np.pi * th)


def shocks(main_seed: int, stream: int, first_path: int, n: int, n_months: int, rho: float) -> np.ndarray:
    """[n_months, 3, n] correlated unit shocks (equity, inflation, premium) of global paths
    first_path .. first_path + n - 1 in double precision (the device transform is fp32)."""
    calls = (n_months + 1) // 2
    paths = np.uint64(first_path) + np.arange(n, dtype=np.uint64)
    p_lo = np.broadcast_to((paths & np.uint64(0xFFFFFFFF)).astype(np.uint32)[None, :], (calls, n))
    p_hi = np.broadcast_to((paths >> np.uint64(32)).astype(np.uint32)[None, :], (calls, n))
    cc = np.broadcast_to(np.arange(calls, dtype=np.uint32)[:, None], (calls, n))
    w0, w1, w2, w3 = philox4x32_10((p_lo, p_hi, cc, np.full((calls, n), stream, dtype=np.uint32)), key_from_seed(main_seed))
    a0, a1 = _box_muller(w0 >> np.uint32(6), w3 & np.uint32(0xFFFF), 16)        # pair A: month 2c
    b0, b1 = _box_muller(w1 >> np.uint32(6), w3 >> np.uint32(16), 16)           # pair B: month 2c + 1
    c0, c1 = _box_muller(w2 >> np.uint32(6), ((w0 & np.uint32(63)) << np.uint32(6)) | (w1 & np.uint32(63)), 12)
    rc = math.sqrt(max(0.0, 1.0 - rho * rho))
    out = np.empty((2 * calls, 3, n))
    out[0::2, 0], out[0::2, 1], out[0::2, 2] = a0, rho * a0 + rc * a1, c0
    out[1::2, 0], out[1::2, 1], out[1::2, 2] = b0, rho * b0 + rc * b1, c1
    return out[:n_months]
