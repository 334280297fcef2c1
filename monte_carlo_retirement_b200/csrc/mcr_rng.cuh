// mcr_rng.cuh — shock sources for the timeline kernel.
//   ReplayShock : reads precomputed correlated shocks (the reference's own numpy draws,
//                 /root/reference/backend/simulation.py:452-466) from HBM, layout
//                 [(month*3 + component) * ld + path] so a warp reads 256 contiguous bytes.
//   PhiloxShock : counter-based Philox4x32-10 (Salmon et al., SC'11). key = f(main_seed),
//                 counter = (path_lo, path_hi, absolute month / 2, seed stream). One call per
//                 TWO months -> 4 x u32 -> three Box-Muller pairs -> 2 x (equity, indep, premium)
//                 unit normals; inflation = rho*equity + sqrt(1-rho^2)*indep as in :459-465. The
//                 draw for (path, month) never depends on working_months, launch geometry or
//                 shard layout: common random numbers across search candidates
//                 (simulation.py:152-154,192-199) and G-GPU invariance by construction.
// Why two months per call: on B200 the integer multiplies and logic ops of Philox issue at half
// rate on the datapath they share with the FP32 work (tools/microbench/issue_model2.cu), and that
// datapath — not the FP64 pipe — bounds the timeline kernel; the draws were > 1/3 of its load.
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

// Box-Muller. One Philox call (128 bits) feeds THREE pairs = six unit normals = the draws of two
// months: per pair a 26-bit radius uniform u1 = (a + 0.5) / 2^26 in (0, 1) and an angle of 16 (12
// for the third pair) bits, theta = 2*pi*(b + 0.5)/2^k - pi.
//   * radius: |n| <= sqrt(-2 ln 2^-27) = 6.12; the mass beyond is 9e-10 per normal, the radius grid
//     is finer than fp32's own resolution everywhere but the last ~4 atoms of the tail;
//   * angle: for N equally spaced angles and a continuous radius the joint characteristic function
//     of (r cos t, r sin t) differs from the Gaussian one only by Bessel terms J_{mN}(r|t|), m >= 1
//     (Jacobi-Anger), i.e. by less than (r|t|/2)^N / N! with N = 4096 or 65536: the marginals and
//     the pair are normal far beyond double precision. tests: KS distance, moments, tail counts,
//     pair / cross-month correlations (test_native_shock_distribution).
// FAST: MUFU lg2 / sqrt / sin / cos (abs error ~5e-7 on a unit normal, invisible next to the
// 1/sqrt(N) sampling error). !FAST: full-precision logf / sqrtf / sincospif.
template <bool FAST, int ANGLE_BITS>
MCR_DEV void box_muller(uint32_t a26, uint32_t b, float& n0, float& n1) {
  constexpr float kTwoPi = 6.2831853071795865f;
  constexpr float kStep = kTwoPi / (float)(1u << ANGLE_BITS);             // 2*pi / 2^k
  constexpr float kOff = -3.1415926535897931f + 0.5f * kStep;             // -pi + half a step
  const float u1 = fmaf(u32_to_float(a26), 1.4901161193847656e-08f, 7.4505805969238281e-09f);  // 2^-26, 2^-27
#ifdef __CUDA_ARCH__
  if constexpr (FAST) {
    float lg, r, s, c;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u1));
    const float t = lg * -1.3862943611198906f;  // -2 ln2 * log2(u1) = -2 ln(u1) >= 0
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    const float th = fmaf(u32_to_float(b), kStep, kOff);
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(th));
    n0 = r * c;
    n1 = r * s;
    return;
  }
#endif
  const float r = sqrtf(-2.0f * logf(u1));
  constexpr float kTurn = 2.0f / (float)(1u << ANGLE_BITS);
  const float turns = fmaf(u32_to_float(b), kTurn, -1.0f + 0.5f * kTurn);  // theta / pi in (-1, 1)
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

// Draw layout (what `mcr_draw_shocks` exposes and the tests restate in numpy): call c = month >> 1,
// counter (path_lo, path_hi, c, stream) -> words w0..w3;
//   pair A: radius w0 >> 6, angle w3 & 0xffff          -> (equity, independent) of month 2c
//   pair B: radius w1 >> 6, angle w3 >> 16             -> (equity, independent) of month 2c + 1
//   pair C: radius w2 >> 6, angle ((w0 & 63) << 6) | (w1 & 63)  (12 bits)
//                                                       -> premium of month 2c (cos), 2c + 1 (sin)
// inflation = rho * equity + sqrt(1 - rho^2) * independent (simulation.py:459-465); equity and
// independent come from ONE pair, so |inflation shock| <= that pair's radius.
template <bool FAST>
struct PhiloxShock {
  const PhiloxKeys& keys;
  uint32_t p_lo, p_hi, month, strm;
  float rho_f, rho_c_f;
  double rho, rho_c;
  float s0, s1, s2;   // the odd month's normals, drawn together with the even month's
  MCR_DEV void next(double& ze, double& zi, double& zp) {
    float n0, n1, n2;
    if ((month & 1u) == 0u) {
      uint32_t r[4];
      philox4x32_10(p_lo, p_hi, month >> 1, strm, keys, r);
      float c0, c1;
      box_muller<FAST, 16>(r[0] >> 6, r[3] & 0xffffu, n0, n1);
      box_muller<FAST, 16>(r[1] >> 6, r[3] >> 16, s0, s1);
      box_muller<FAST, 12>(r[2] >> 6, ((r[0] & 63u) << 6) | (r[1] & 63u), c0, c1);
      n2 = c0;
      s2 = c1;
    } else {
      n0 = s0;
      n1 = s1;
      n2 = s2;
    }
    ++month;
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
