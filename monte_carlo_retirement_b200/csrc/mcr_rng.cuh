// mcr_rng.cuh — shock sources for the timeline kernel.
//   ReplayShock : reads precomputed correlated shocks (the reference's own numpy draws,
//                 /root/reference/backend/simulation.py:452-466) from HBM, layout
//                 [(month*3 + component) * ld + path] so a warp reads 256 contiguous bytes.
//   PhiloxShock : counter-based Philox4x32-10 (Salmon et al., SC'11). key = f(main_seed),
//                 counter = (path_lo, path_hi, absolute month, seed stream). One call per
//                 month -> 4 x u32 -> two Box-Muller pairs -> (equity, indep, premium) unit
//                 normals; inflation = rho*equity + sqrt(1-rho^2)*indep as in :459-465. The
//                 draw for (path, month) never depends on working_months, launch geometry or
//                 shard layout: common random numbers across search candidates
//                 (simulation.py:152-154,192-199) and G-GPU invariance by construction.
// Normals are generated off the FP64 pipe (INT + FP32 + MUFU issue slots are otherwise idle
// in this FP64-bound kernel) and widened with one F2F each.
#pragma once
#include <cstdint>

#include "mcr_portable.h"

namespace mcr {

// The key schedule (k + r*W per round) does not depend on the counter: the host expands it once
// into the launch arguments (PhiloxKeys), so each round is 2 IMAD.WIDE + 2 three-input LOP3 with
// the round keys as constant-bank operands.
struct PhiloxKeys {
  uint32_t rk[20];  // rk[2r], rk[2r+1] = (k0 + r*0x9E3779B9, k1 + r*0xBB67AE85)
};

MCR_DEV void philox_expand_keys(uint32_t k0, uint32_t k1, PhiloxKeys& K) {
  for (int r = 0; r < 10; ++r) {
    K.rk[2 * r] = k0 + (uint32_t)r * 0x9E3779B9u;
    K.rk[2 * r + 1] = k1 + (uint32_t)r * 0xBB67AE85u;
  }
}

MCR_DEV void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              const PhiloxKeys& K, uint32_t out[4]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.rk[2 * r];
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.rk[2 * r + 1];
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Box-Muller on two 32-bit words. u1 = (a + 0.5) / 2^32 in (0, 1]; theta = 2*pi*(b + 0.5)/2^32 - pi.
// FAST: MUFU lg2 / sqrt / sin / cos (abs error ~5e-7 on a unit normal, invisible next to the
// 1/sqrt(N) sampling error). !FAST: full-precision logf / sqrtf / sincospif.
template <bool FAST>
MCR_DEV void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float u1 = fmaf(u32_to_float(a), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
#ifdef __CUDA_ARCH__
  if constexpr (FAST) {
    float lg, r, s, c;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u1));
    const float t = lg * -1.3862943611198906f;  // -2 ln2 * log2(u1) = -2 ln(u1) >= 0
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    const float th = fmaf(u32_to_float(b), 1.4629180792671596e-09f, -3.1415926535897931f);
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(th));
    n0 = r * c;
    n1 = r * s;
    return;
  }
#endif
  const float r = sqrtf(-2.0f * logf(u1));
  const float turns = fmaf(u32_to_float(b), 4.6566128730773926e-10f, -1.0f);  // theta / pi in [-1, 1]
  float s, c;
  sincospi_f(turns, s, c);
  n0 = r * c;
  n1 = r * s;
}

struct ReplayShock {
  const double* p;  // shocks_dev + path
  int64_t ld;
  int32_t left;     // rows not yet consumed (the timeline prefetches one month ahead)
  MCR_DEV void next(double& ze, double& zi, double& zp) {
    if (left > 0) {
      ze = load_stream(p);
      zi = load_stream(p + ld);
      zp = load_stream(p + 2 * ld);
      p += 3 * ld;
      --left;
    } else {
      ze = zi = zp = 0.0;  // past the end of the matrix: the prefetched month is never stepped
    }
  }
};

template <bool FAST>
struct PhiloxShock {
  const PhiloxKeys& keys;
  uint32_t p_lo, p_hi, month, strm;
  float rho_f, rho_c_f;
  double rho, rho_c;
  MCR_DEV void next(double& ze, double& zi, double& zp) {
    uint32_t r[4];
    philox4x32_10(p_lo, p_hi, month, strm, keys, r);
    ++month;
    float n0, n1, n2, n3;
    box_muller<FAST>(r[0], r[1], n0, n1);
    box_muller<FAST>(r[2], r[3], n2, n3);
    ze = (double)n0;
    if constexpr (FAST) {
      zi = (double)fmaf(rho_f, n0, rho_c_f * n1);
    } else {
      zi = rho * ze + rho_c * (double)n1;  // simulation.py:461-464
    }
    zp = (double)n2;
  }
};

}  // namespace mcr
