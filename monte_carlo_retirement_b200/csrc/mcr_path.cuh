// mcr_path.cuh — the per-path timeline state machine, one thread = one path, all state in
// registers. Behavioural spec: /root/reference/backend/simulation.py:476-950
// (_run_single_simulation_path) with its helpers :201-450; the step order is the reference's
// (SURVEY Appendix A). This header is compiled twice:
//   * mcr_kernels_strict.cu  (-fmad=false): every product/sum rounds separately, in the
//     reference's operation order — the parity build used by replay / single-path / helpers;
//   * mcr_kernels_fast.cu    (-fmad=true): same formulas, FMA contraction, shared reciprocals
//     and a short-range exp — the throughput build used by native-RNG runs.
// Bit-exact elisions used in BOTH builds: an asset whose realized-gains tax
// is inactive (flag off or rate 0) has effective tax fraction 0, so `target / max(eps, 1.0)`
// == target, its sale pays no tax and its cost basis never reaches an output.
#pragma once
#include <cstdint>

#include "../../include/mcr.h"
#include "mcr_portable.h"

namespace mcr {

constexpr double kEps = MCR_SMALL_EPSILON;
constexpr int kMPY = MCR_MONTHS_PER_YEAR;

// One income stream that can pay something. Streams with monthly_amount_today == 0 are dropped
// on the host: they add +0.0 to the income sum, which is exact. The activity window
// [first, end) of a stream depends on working_months and travels with the launch arguments.
struct DevStream {
  double amount;        // monthly_amount_today
  double net_factor;    // 1.0 - tax_rate
  int32_t duration;     // months, -1 == None (informational; the window already encodes it)
  int32_t indexed;
};

// Scenario constants, passed by value as a __grid_constant__ kernel parameter (uniform loads
// from the constant bank). Everything here is derived on the host by mcr_create.
struct DevParams {
  double B0, C0, growth1p, E, a1, a2;
  double mu1, sg1, muI, sgI, muP, sgP;  // monthly: mu_log/12, sigma_log/sqrt(12)
  double rho, rho_c;                    // equity-inflation correlation, sqrt(max(0,1-rho^2))
  double rate1, rate2;                  // realized-gains tax rates
  double ann1, ann2;                    // annual mark-to-market rates
  float rho_f, rho_c_f;
  int32_t use1, use2;      // inv{1,2}_use_realized_gains_tax_system
  int32_t taxed1, taxed2;  // use && rate > 0 : realized tax actually bites
  int32_t growth_on;       // contribution_growth_rate_annual > 0
  int32_t annual_any;      // some asset can owe annual tax (needs the P&L accumulators)
  int32_t R;               // retirement_years
  int32_t algebra_ok;      // some realized-gains tax bites and every such rate is < 1: closed forms allowed
  int32_t exp_small;       // bound on |monthly log-return| proven for Philox normals (|z| <= 6.8): 2: < 0.05, 1: < 0.1, 0: none
  int32_t n_streams;       // streams with a positive amount, original order
  int32_t lean_cfg_ok;     // allocation in [0.01, 0.99] and every biting realized-gains rate <= 0.9 (lean months)
  double lean_cg;          // 0.5 * (1 - max biting rate) * exp(-0.2): W * lean_cg > E * level  =>  need < cap / 2
  double lean_need_coef;   // E / lean_cg: the month is lean-safe iff W > lean_need_coef * level + kLeanMinW
  double lean_level_max;   // yearly check: E * level stays < 1e8 for the next 12 months
  double hrate1, hrate2;   // rate / 2: max(0, g) * rate == (g + |g|) * (rate / 2) exactly
  DevStream streams[MCR_MAX_STREAMS];
};

// ---- lean months (fast build) -------------------------------------------------------------------
// While a path is far from every threshold of the reference's helpers, a month is straight-line
// arithmetic: no eps snap, no early-out, no clamp can be active. Sufficient conditions, all
// checked ahead of the month (DevParams::lean_*, `bal` in run_timeline):
//   * the portfolio is exactly on target after the last rebalance (b1 = a1*W, b2 = W - b1) and
//     W > kLeanMinW dollars, allocation in [0.01, 0.99]  => every balance the month sees is > 0.1;
//   * every monthly log-return is in (-0.1, 0.1) (Cfg::kExpSmall), every biting rate <= 0.9, and
//     W * lean_cg > E * level  => need <= E * level < cap / 2: the full need is met (target ==
//     need, the fraction withdrawn f < 1/2), no failure test can fire;
//   * price level in (4e-3, lean_level_max), re-checked every retirement year.
// Then (reference: simulation.py:692-796 for a retirement month, :519-553 for an accumulation
// month) the month reduces to: growth; tax due on liquidation tx_i = max(0, b_i - cb_i) * rate_i;
// cap = V - tx1 - tx2; f = need / cap; everything scales by q = 1 - f; and the rebalance, which
// solves (b_s - x) = a_s * (V - tax(x)) exactly, leaves the portfolio ON TARGET with
// W' = q * (V - fs * tx_s), fs = |drift| / (b_s - a_s * tx_s). Same formulas in exact
// arithmetic; only the association of the products differs (<= 1e-15 relative per month).
#ifndef MCR_COUNT_LEAN_MONTH
#define MCR_COUNT_LEAN_MONTH()   // tests/host_model counts the months that took the lean step
#endif
constexpr double kLeanMinW = 16.0;
constexpr double kLeanLevelMin = 4e-3;

// income of one stream in retirement month r (simulation.py:650-677); `lock` holds the nominal
// amount a non-indexed stream froze at its first payment (NaN-free sentinel: locked flag).
MCR_DEV void stream_income(const DevStream& st, int first, int end, int r, double level0,
                                              double& lock, bool& is_locked, double& income) {
  if (r >= first && r < end) {
    double nominal;
    if (st.indexed) {
      nominal = st.amount * level0;
    } else {
      if (!is_locked) { lock = st.amount * level0; is_locked = true; }
      nominal = lock;
    }
    income += nominal * st.net_factor;
  }
}

struct PathOut {
  double start_balance, final_balance, fy_gross, fy_real, infl_ret;
  int32_t success;
  int32_t ruin_month;  // -1 == NaN
  uint32_t executed;   // months stepped (shock rows consumed)
};

// Compile-time view of the scenario's tax switches. The switches are warp-uniform, but as run-time
// branches they cut the month into many small basic blocks and keep the two withdrawals (and
// other independent chains) from being interleaved by the scheduler; the launcher therefore
// picks a specialisation when the scenario matches one (-1 = read the flag at run time).
template <int TAXED1, int TAXED2, int ANNUAL, int EXPSMALL = 0>
struct Cfg {
  // EXPSMALL: |mu/12 + sigma/sqrt(12) * z| is bounded for all three factors — proven by the host
  // for the Box-Muller normals of 32-bit uniforms (|z| <= 6.77), or guaranteed by the caller of a
  // replay (MCR_FLAG_SMALL_RETURNS) — so a short Taylor polynomial of the unreduced argument needs
  // no range test: level 2 (< 0.05) degree 6, level 1 (< 0.1) degree 8; truncation x^(d+1)/(d+1)!
  // <= 3e-15 at the respective bound's typical 5-sigma draw, i.e. at rounding level.
  static constexpr int kExpDegree = EXPSMALL == 2 ? 6 : 8;
  static constexpr bool kExpSmall = EXPSMALL != 0;
  // kAlgebra: both assets taxed on realized gains with rate < 1 (checked by the launcher): the
  // fast build may use the closed forms of the withdrawal pair and of the rebalance sale (see
  // withdraw_pair_closed / rebalance_main).
  static constexpr bool kAlgebra = TAXED1 == 1 && TAXED2 == 1 && ANNUAL == 0;
  // kLean: the tax switches are compile-time, no annual tax, bounded monthly returns: the fast
  // build may take the lean month steps (see "lean months" above)
  static constexpr bool kLean = TAXED1 >= 0 && TAXED2 >= 0 && ANNUAL == 0 && EXPSMALL != 0;
  static constexpr bool kAnyTaxed = TAXED1 != 0 || TAXED2 != 0;
  // generic configuration: the same closed forms behind a (warp-uniform) run-time test
  static MCR_DEV bool algebra(const DevParams& P) {
    if constexpr (TAXED1 < 0) return P.algebra_ok != 0; else return kAlgebra;
  }
  static MCR_DEV bool taxed1(const DevParams& P) {
    if constexpr (TAXED1 < 0) return P.taxed1 != 0; else return TAXED1 != 0;
  }
  static MCR_DEV bool taxed2(const DevParams& P) {
    if constexpr (TAXED2 < 0) return P.taxed2 != 0; else return TAXED2 != 0;
  }
  static MCR_DEV bool annual(const DevParams& P) {
    if constexpr (ANNUAL < 0) return P.annual_any != 0; else return ANNUAL != 0;
  }
};
using CfgGeneric = Cfg<-1, -1, -1>;
using CfgBothTaxed = Cfg<1, 1, 0>;   // both assets on the realized-gains system with a positive rate, no annual tax
using CfgNoTax = Cfg<0, 0, 0>;       // no realized-gains tax bites and no annual tax
using CfgBothTaxedSmall = Cfg<1, 1, 0, 1>;   // |monthly log-return| < 0.1
using CfgNoTaxSmall = Cfg<0, 0, 0, 1>;
using CfgBothTaxedTight = Cfg<1, 1, 0, 2>;   // |monthly log-return| < 0.05
using CfgNoTaxTight = Cfg<0, 0, 0, 2>;

// CPython max(a, b) / min(a, b): first argument wins ties and NaN compares. Written as
// setp + selp so the compiler cannot canonicalise them into fmax/fmin, whose NaN-correct SASS
// expansion is 6 instructions instead of DSETP + 2 SEL.
MCR_DEV double pmax(double a, double b) {
#ifdef __CUDA_ARCH__
  double d;
  asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %2, %1;\n\tselp.f64 %0, %2, %1, p;\n\t}" : "=d"(d) : "d"(a), "d"(b));
  return d;
#else
  return b > a ? b : a;
#endif
}
MCR_DEV double pmin(double a, double b) {
#ifdef __CUDA_ARCH__
  double d;
  asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %2, %1;\n\tselp.f64 %0, %2, %1, p;\n\t}" : "=d"(d) : "d"(a), "d"(b));
  return d;
#else
  return b < a ? b : a;
#endif
}

// (a <= c) || (|b| <= c) and (a < c) || (b < c) as two DSETPs + a predicate OR. Written in PTX
// because the compiler otherwise rewrites the pair into fmin(a, b) < c, whose NaN-correct
// expansion costs 6 instructions.
MCR_DEV bool either_le_abs(double a, double b, double c) {
#ifndef __CUDA_ARCH__
  return a <= c || std::fabs(b) <= c;
#else
  uint32_t r;
  asm("{\n\t.reg .pred p, q;\n\t.reg .f64 t;\n\tabs.f64 t, %2;\n\tsetp.le.f64 p, %1, %3;\n\t"
      "setp.le.f64 q, t, %3;\n\tor.pred p, p, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r) : "d"(a), "d"(b), "d"(c));
  return r != 0;
#endif
}
MCR_DEV bool either_le(double a, double b, double c) {
#ifndef __CUDA_ARCH__
  return a <= c || b <= c;
#else
  uint32_t r;
  asm("{\n\t.reg .pred p, q;\n\tsetp.le.f64 p, %1, %3;\n\tsetp.le.f64 q, %2, %3;\n\tor.pred p, p, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r) : "d"(a), "d"(b), "d"(c));
  return r != 0;
#endif
}
MCR_DEV bool either_lt(double a, double b, double c) {
#ifndef __CUDA_ARCH__
  return a < c || b < c;
#else
  uint32_t r;
  asm("{\n\t.reg .pred p, q;\n\tsetp.lt.f64 p, %1, %3;\n\tsetp.lt.f64 q, %2, %3;\n\tor.pred p, p, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r) : "d"(a), "d"(b), "d"(c));
  return r != 0;
#endif
}

// exp(x) Taylor coefficients 1/13! .. 1/2! — constant-bank operands of the DFMAs (no UMOV pairs)
#define MCR_EXP_COEFFS                                                                             \
  {1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07,    \
   2.7557319223985893e-06, 2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03,    \
   8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5}
#if defined(__CUDACC__)
static __constant__ double kExpC[12] = MCR_EXP_COEFFS;
#endif
static constexpr double kExpCHost[12] = MCR_EXP_COEFFS;  // tests/host_model only
MCR_DEV double expc(int i) {
#ifdef __CUDA_ARCH__
  return kExpC[i];
#else
  return kExpCHost[i];
#endif
}

// ---------------------------------------------------------------------------------------------
// math policy
// ---------------------------------------------------------------------------------------------
template <bool FAST>
struct Math {
  // reciprocal to ~1 ulp: MUFU.RCP64H seed + the 5-DFMA refinement (fast build only)
  static MCR_DEV double rcp(double b) {
    if constexpr (FAST) {
      const double x0 = rcp_seed(b);
      double e = fma(-b, x0, 1.0);
      double e2 = fma(e, e, e);
      double x1 = fma(x0, e2, x0);
      double e3 = fma(-b, x1, 1.0);
      return fma(x1, e3, x1);
    } else {
      return 1.0 / b;
    }
  }
  // lean months: seed (relative error e ~ 2^-20) + one cubic step, x0 * (1 + e + e^2): relative
  // error e^3 ~ 1e-18 plus two roundings (<= ~2 ulp). The two extra DFMAs of rcp() only buy the
  // last ulp, which the 1e-9 contract of this build does not need.
  static MCR_DEV double rcp3(double b) {
    const double x0 = rcp_seed(b);
    const double e = fma(-b, x0, 1.0);
    const double e2 = fma(e, e, e);
    return fma(x0, e2, x0);
  }
  // lean months, where the quotient does not feed the balance recurrence at first order (the
  // fraction of the over-weight asset sold — it only sets the tax on the sale —, the real value of
  // a withdrawal): seed + ONE quadratic step, relative error e^2 ~ 1e-12. The withdrawal fraction
  // f = need / cap keeps rcp3: its error (always the same sign) would add up to 2e-12 of the
  // balance over 720 months and show at 1e-9 in the last samples of a path that runs dry.
  static MCR_DEV double rcp2(double b) {
    const double x0 = rcp_seed(b);
    return fma(x0, fma(-b, x0, 1.0), x0);
  }
  // a / b: IEEE in the strict build; a * rcp(b) (<= 2 ulp, no slow-path branch) in the fast one
  static MCR_DEV double div(double a, double b) {
    if constexpr (FAST) return a * rcp(b);
    else return a / b;
  }
  static MCR_DEV double poly_exp(double x) {
    double p = expc(0);
#pragma unroll
    for (int i = 1; i < 12; ++i) p = fma(p, x, expc(i));
    p = fma(p, x, 1.0);
    return fma(p, x, 1.0);
  }
  // the three monthly gross factors exp(mu/12 + sigma/sqrt(12)*z) — simulation.py:468-474,522-530.
  // Fast build: ONE combined range test (integer compare on the high words), then three
  // interleaved Horner chains in a single basic block.
  // bounded argument (Cfg::kExpSmall): Taylor polynomial of degree DEG, no range test
  template <int DEG>
  static MCR_DEV double poly_exp_small(double x) {
    double p = expc(13 - DEG);          // 1 / DEG!
#pragma unroll
    for (int i = 14 - DEG; i < 12; ++i) p = fma(p, x, expc(i));
    p = fma(p, x, 1.0);
    return fma(p, x, 1.0);
  }
  template <class C>
  static MCR_DEV void factors(const DevParams& P, double ze, double zi, double zp, double& G1,
                                                 double& GI, double& GP) {
    const double x1 = P.mu1 + P.sg1 * ze;  // strict: two roundings (no FMA); fast: one DFMA
    const double xi = P.muI + P.sgI * zi;
    const double xp = P.muP + P.sgP * zp;
    if constexpr (FAST && C::kExpSmall) {
      G1 = poly_exp_small<C::kExpDegree>(x1);
      GI = poly_exp_small<C::kExpDegree>(xi);
      GP = poly_exp_small<C::kExpDegree>(xp);
      return;
    }
    if constexpr (FAST) {
      const int h1 = hi_word(x1) & 0x7fffffff, hi = hi_word(xi) & 0x7fffffff, hp = hi_word(xp) & 0x7fffffff;
      const int hmax = h1 > hi ? (h1 > hp ? h1 : hp) : (hi > hp ? hi : hp);
      if (__builtin_expect(hmax < 0x3FD5C28F, 1)) {  // all |x| < 0.34
        G1 = poly_exp(x1);
        GI = poly_exp(xi);
        GP = poly_exp(xp);
        return;
      }
    }
    G1 = exp(x1);
    GI = exp(xi);
    GP = exp(xp);
  }
  // monthly gross factor exp(mu + sg * z) — simulation.py:468-474
  static MCR_DEV double gross(double mu, double sg, double z) {
    const double x = mu + sg * z;  // strict: two roundings (no FMA); fast: one DFMA
    if constexpr (FAST) {
      // |x| is a monthly log-return, far inside [-0.5 ln2, 0.5 ln2] for any sane scenario: a
      // degree-13 Taylor polynomial on the unreduced argument (truncation < 3e-18 relative at
      // |x| = 0.34, 13 DFMA instead of the library's 17 FP64 ops + range reduction), with the
      // general routine as the (warp-rare) fallback.
      if (__builtin_expect(fabs(x) < 0.34, 1)) {
        double p = expc(0);
#pragma unroll
        for (int i = 1; i < 12; ++i) p = fma(p, x, expc(i));
        p = fma(p, x, 1.0);
        return fma(p, x, 1.0);
      }
      return exp(x);
    } else {
      return exp(x);
    }
  }
};

// ---------------------------------------------------------------------------------------------
// helpers — simulation.py:201-450. FULL = also maintain the cost basis of an untaxed asset
// (needed only by the helper entry points that RETURN the basis).
//
// Clamps the reference writes but that are identities under the engine's invariants
// (bal >= 0, cost basis >= 0, 0 <= rate <= 1, correctly rounded IEEE ops) are dropped; each is
// annotated "identity:" with the reason. They return their second argument unchanged, so the
// results stay bit-identical (the helper parity tests are bit-exact on random inputs, including
// basis-above-market cases).
// ---------------------------------------------------------------------------------------------

// _net_liquidation_value — simulation.py:256-272
MCR_DEV double net_liq(double bal, double cb, bool taxed, double rate) {
  if (!taxed) return bal > kEps ? bal : 0.0;   // tax == 0.0; identity: max(0, bal - 0.0) with bal > eps
  const double gain = pmax(0.0, bal - cb);
  // identity: max(0, bal - gain*rate): gain <= bal and rate <= 1, so the difference is >= +0.0
  return bal > kEps ? bal - gain * rate : 0.0;
}

// _calculate_withdrawal_and_update — simulation.py:201-254
// main path (the early-out of :218-219 is tested by the callers)
template <bool FAST, bool FULL>
MCR_DEV void withdraw_main(double& bal, double& cb, double target, bool taxed, double rate,
                                              double& gross, double& net) {
  if (!taxed && !FULL) {
    // effective tax fraction 0 -> net_fraction == 1.0, target / 1.0 == target, tax_paid == 0.
    gross = pmin(target, bal);
    net = gross;                       // identity: max(0, gross - 0.0), gross > 0
    double nb = bal - gross;           // identity: max(0, bal - gross), gross <= bal
    if (nb <= kEps) nb = 0.0;
    bal = nb;
    return;  // cost basis of an untaxed asset is dead state
  }
  double fs;
  if constexpr (FAST) {
    const double rb = Math<FAST>::rcp(bal);
    const double gf = pmax(0.0, bal - cb) * rb;
    const double etf = taxed ? gf * rate : 0.0;
    const double nf = pmax(kEps, 1.0 - etf);
    gross = pmin(target * Math<FAST>::rcp(nf), bal);
    fs = pmin(1.0, gross * rb);        // the approximate reciprocal may land 1 ulp above 1
  } else {
    const double gf = pmax(0.0, bal - cb) / bal;         // :221
    const double etf = taxed ? gf * rate : 0.0;          // :222-226
    const double nf = pmax(kEps, 1.0 - etf);             // :227
    gross = pmin(target / nf, bal);                      // :228-231
    fs = gross / bal;                  // :233 identity: min(1, gross/bal), gross <= bal
  }
  const double br = cb * fs;           // :234 identity: min(cb, cb*fs), fs <= 1
  const double tg = pmax(0.0, gross - br);               // :235
  const double tax = taxed ? tg * rate : 0.0;            // :236-240
  net = gross - tax;                   // :241 identity: max(0, .), tax <= tg <= gross
  double nb = bal - gross;             // :243 identity: max(0, .), gross <= bal
  double ncb = cb - br;                // :244 identity: max(0, .), br <= cb
  if (nb <= kEps) { nb = 0.0; ncb = 0.0; }               // :245-247
  bal = nb;
  cb = ncb;
}

MCR_DEV bool withdraw_skips(double bal, double target) { return bal <= kEps || target <= 0; }

template <bool FAST, bool FULL>
MCR_DEV void withdraw(double& bal, double& cb, double target, bool taxed, double rate,
                                         double& gross, double& net) {
  if (__builtin_expect(withdraw_skips(bal, target), 0)) {  // :218-219
    bal = pmax(0.0, bal);
    cb = pmax(0.0, cb);
    gross = 0.0;
    net = 0.0;
    return;
  }
  withdraw_main<FAST, FULL>(bal, cb, target, taxed, rate, gross, net);
}

// Both monthly withdrawals with ONE combined early-out test, so the two dependent chains sit in
// one basic block and interleave (simulation.py:757-777).
template <bool FAST, class C>
MCR_DEV void withdraw_pair(const DevParams& P, double& b1, double& cb1, double t1, double& b2,
                                              double& cb2, double t2, double& gw1, double& nw1, double& gw2,
                                              double& nw2) {
  if (__builtin_expect(!(withdraw_skips(b1, t1) || withdraw_skips(b2, t2)), 1)) {
    withdraw_main<FAST, false>(b1, cb1, t1, C::taxed1(P), P.rate1, gw1, nw1);
    withdraw_main<FAST, false>(b2, cb2, t2, C::taxed2(P), P.rate2, gw2, nw2);
  } else {
    withdraw<FAST, false>(b1, cb1, t1, C::taxed1(P), P.rate1, gw1, nw1);
    withdraw<FAST, false>(b2, cb2, t2, C::taxed2(P), P.rate2, gw2, nw2);
  }
}

// _rebalance_portfolio — simulation.py:274-359. The sell-asset-1 / sell-asset-2 branches are
// folded into one straight-line body by selecting the roles (the direction differs per lane,
// the flags do not), so a warp never executes both.
// :290-296 — nothing to do when the portfolio is empty or already on target
MCR_DEV bool rebalance_skips(const DevParams& P, double b1, double b2) {
  const double total = b1 + b2;
  const double drift1 = b1 - total * P.a1;
  return either_le_abs(total, drift1, kEps);
}

template <bool FAST, bool FULL, class C = CfgGeneric>
MCR_DEV void rebalance_main(const DevParams& P, double& b1, double& cb1, double& b2, double& cb2) {
  const double total = b1 + b2;
  const double drift1 = b1 - total * P.a1;                // :293-294
  const bool sell1 = drift1 > 0;
  const double drift2 = b2 - total * P.a2;                // :328 (recomputed, not -drift1)
  const double bs = sell1 ? b1 : b2;
  const double bo = sell1 ? b2 : b1;
  const double drift = sell1 ? drift1 : drift2;
  // the rebalance consults only the `use` flag (:302-306); use && rate == 0 gives tpd == 0.0
  // exactly, so `taxed` decides the arithmetic in both cases.
  if (!C::taxed1(P) && !C::taxed2(P) && !FULL) {
    const double sale = pmin(bs, drift);   // denominator == max(eps, 1.0 - a*0.0) == 1.0; tax_paid == 0.0
    double nbs = bs - sale;                // identity: max(0, .), sale <= bs
    double nbo = bo + sale;
    if (nbs <= kEps) nbs = 0.0;            // :355-358
    if (nbo <= kEps) nbo = 0.0;
    b1 = sell1 ? nbs : nbo;
    b2 = sell1 ? nbo : nbs;
    return;
  }
  const double cbs = sell1 ? cb1 : cb2;
  const double cbo = sell1 ? cb2 : cb1;
  // use && rate == 0 and !use both mean "no tax on this sale": taxed{1,2} (compile-time in the
  // specialised configurations) selects the same value as the reference's `use` test
  const double rate = sell1 ? (C::taxed1(P) ? P.rate1 : 0.0) : (C::taxed2(P) ? P.rate2 : 0.0);
  const double as = sell1 ? P.a1 : P.a2;
  double sale, fs;
  if (FAST && !FULL && C::algebra(P)) {
    // gf = gain/bs, den = 1 - as*gf*rate = (bs - as*gain*rate)/bs, so
    //   fraction sold fs = sale/bs = drift/(bs - as*gain*rate)   (one reciprocal instead of two)
    //   sale = fs*bs, basis removed = fs*cbs, taxable gain = fs*gain, purchase = sale - fs*gain*rate.
    // den >= 1 - as*rate > eps because rate < 1, so the reference's max(eps, .) is inactive.
    const double gr = pmax(0.0, bs - cbs) * rate;
    fs = pmin(1.0, drift * Math<FAST>::rcp(bs - as * gr));
    sale = fs * bs;
    const double br_c = fs * cbs;
    const double buy_c = sale - fs * gr;
    double nbs = bs - sale, ncbs = cbs - br_c, nbo = bo + buy_c, ncbo = cbo + buy_c;
    if (__builtin_expect(either_le(nbs, nbo, kEps), 0)) {   // :355-358, one rarely-taken branch for both
      if (nbs <= kEps) { nbs = 0.0; ncbs = 0.0; }
      if (nbo <= kEps) { nbo = 0.0; ncbo = 0.0; }
    }
    b1 = sell1 ? nbs : nbo;
    cb1 = sell1 ? ncbs : ncbo;
    b2 = sell1 ? nbo : nbs;
    cb2 = sell1 ? ncbo : ncbs;
    return;
  }
  if constexpr (FAST) {
    const double rb = Math<FAST>::rcp(bs);
    const double gf = pmax(0.0, bs - cbs) * rb;
    const double den = pmax(kEps, 1.0 - as * (gf * rate));
    sale = pmin(bs, drift * Math<FAST>::rcp(den));
    fs = pmin(1.0, sale * rb);
  } else {
    const double gf = pmax(0.0, bs - cbs) / bs;           // :301 / :329
    const double den = pmax(kEps, 1.0 - as * (gf * rate)); // :302-310
    sale = pmin(bs, drift / den);                         // :311
    fs = sale / bs;                                       // :312
  }
  const double br = cbs * fs;                             // :313 identity: min(cb, cb*fs), fs <= 1
  const double tg = pmax(0.0, sale - br);                 // :314
  const double buy = sale - tg * rate;                    // :315-320
  double nbs = bs - sale;                                 // :322 identity: max(0, .), sale <= bs
  double ncbs = cbs - br;                                 // :323 identity: max(0, .), br <= cbs
  double nbo = bo + buy;                                  // :324
  double ncbo = cbo + buy;                                // :325
  if (nbs <= kEps) { nbs = 0.0; ncbs = 0.0; }             // :355-358
  if (nbo <= kEps) { nbo = 0.0; ncbo = 0.0; }
  b1 = sell1 ? nbs : nbo;
  cb1 = sell1 ? ncbs : ncbo;
  b2 = sell1 ? nbo : nbs;
  cb2 = sell1 ? ncbo : ncbs;
}

template <bool FAST, bool FULL, class C = CfgGeneric>
MCR_DEV void rebalance(const DevParams& P, double& b1, double& cb1, double& b2, double& cb2) {
  if (__builtin_expect(rebalance_skips(P, b1, b2), 0)) return;
  rebalance_main<FAST, FULL, C>(P, b1, cb1, b2, cb2);
}

// _apply_annual_gain_taxes — simulation.py:361-450. Returns tax_failed.
template <bool FAST, bool FULL, class C = CfgGeneric>
MCR_DEV bool annual_tax(const DevParams& P, double& b1, double& cb1, double& b2, double& cb2,
                                        double g1, double g2) {
  bool failed = false;
  if (C::annual(P)) {
    const double due1 = !P.use1 ? pmax(0.0, g1) * P.ann1 : 0.0;   // :380-384
    const double due2 = !P.use2 ? pmax(0.0, g2) * P.ann2 : 0.0;   // :385-389
    const double due = due1 + due2;
    const double cap1 = net_liq(b1, cb1, C::taxed1(P), P.rate1);  // :392-403
    const double cap2 = net_liq(b2, cb2, C::taxed2(P), P.rate2);
    const double cap = cap1 + cap2;
    const double pay = pmin(due, cap);                            // :405
    failed = pay < due - kEps;                                    // :406
    if (cap > kEps && pay > 0) {                                  // :408-430
      const double share1 = cap1 / cap;
      const double share2 = 1.0 - share1;
      double gw, n1, n2;
      withdraw<FAST, FULL>(b1, cb1, pay * share1, C::taxed1(P), P.rate1, gw, n1);
      withdraw<FAST, FULL>(b2, cb2, pay * share2, C::taxed2(P), P.rate2, gw, n2);
      if (n1 + n2 < due - kEps) failed = true;
    }
  }
  // else: no asset can owe annual tax -> due == 0.0, pay == min(0.0, cap) == 0.0, nothing is
  // sold and tax_failed is False; only the trailing rebalance of :432-442 runs.
  rebalance<FAST, FULL, C>(P, b1, cb1, b2, cb2);                  // :432-442
  return failed;
}

// ---------------------------------------------------------------------------------------------
// sinks for the yearly series
// ---------------------------------------------------------------------------------------------
struct NullSink {
  static constexpr bool kActive = false;
  MCR_DEV void point(int, double, double) {}
  MCR_DEV void wr(int, double) {}
};

// time-major [t][ld]: a warp's 32 paths store 256 contiguous bytes per point
struct SeriesSink {
  static constexpr bool kActive = true;
  double* traj;  // may be NULL
  double* real;  // may be NULL
  double* wrp;   // may be NULL
  int64_t ld;
  MCR_DEV void point(int t, double nominal, double price) {
    if (traj) store_stream(traj + (int64_t)t * ld, nominal);
    if (real) store_stream(real + (int64_t)t * ld, price > kEps ? nominal / price : 0.0);  // :928-931
  }
  MCR_DEV void wr(int y, double v) {
    if (wrp) store_stream(wrp + (int64_t)y * ld, v);
  }
};


// ---------------------------------------------------------------------------------------------
// lean month steps (fast build only; see "lean months" next to DevParams)
// ---------------------------------------------------------------------------------------------
struct Factors {
  double G1, GI, GP;
};

// max(0, gain) * rate as (gain + |gain|) * (rate / 2): the same value bit for bit (the sum is
// exactly 2 * max(0, gain), the halved rate only shifts the exponent), but two FP64-pipe
// instructions instead of a compare and two 32-bit selects — the selects issue on the half-rate
// integer / FP32 datapath that bounds this kernel on B200, the FP64 pipe has slack.
MCR_DEV double latent_tax(double gain, double half_rate) { return (gain + fabs(gain)) * half_rate; }

// The rebalance of a portfolio (nb1, nb2) with cost bases (c1, c2) that is guaranteed to trade
// without clamps: returns the fraction-sold bookkeeping applied to the cost bases and the total
// after the sale's tax, V - fs * tx_s (simulation.py:298-353 solved in closed form).
template <class C>
MCR_DEV double lean_rebalance(const DevParams& P, double nb1, double nb2, double V, double tx1, double tx2,
                              double& c1, double& c2) {
  if constexpr (!C::kAnyTaxed) {
    return V;  // no tax on the sale: the total is unchanged, cost bases are dead state
  } else {
    const double d0 = fma(-P.a1, V, nb1);          // drift of asset 1 (:293-294); asset 2 drifts by -d0
    const bool sell1 = d0 > 0.0;
    const double bs = sell1 ? nb1 : nb2;
    const double gs = sell1 ? tx1 : tx2;           // tax due if the WHOLE selling asset were liquidated
    const double cbs = sell1 ? c1 : c2;
    const double as = sell1 ? P.a1 : P.a2;
    const double fs = fabs(d0) * Math<true>::rcp2(fma(-as, gs, bs));   // fraction of the asset sold
    const double tax = fs * gs;
    const double buy = fma(fs, bs, -tax);          // sale - tax: what reaches the other asset
    const double br = fs * cbs;                    // basis removed from the selling asset
    if (sell1) { c1 -= br; c2 += buy; } else { c1 += buy; c2 -= br; }
    return V - tax;
  }
}

// accumulation month — simulation.py:519-553 (k1, k2: this year's contribution split)
template <class C>
MCR_DEV void lean_accumulate(const DevParams& P, const Factors& cur, double k1, double k2, double& b1, double& cb1,
                             double& b2, double& cb2, double& level, bool& bal) {
  const double G2 = cur.GI * cur.GP;
  const double nb1 = fma(b1, cur.G1, k1);
  const double nb2 = fma(b2, G2, k2);
  level *= cur.GI;
  const double V = nb1 + nb2;
  double c1 = cb1 + k1, c2 = cb2 + k2;
  const double tx1 = C::taxed1(P) ? latent_tax(nb1 - c1, P.hrate1) : 0.0;
  const double tx2 = C::taxed2(P) ? latent_tax(nb2 - c2, P.hrate2) : 0.0;
  const double Wn = lean_rebalance<C>(P, nb1, nb2, V, tx1, tx2, c1, c2);
  cb1 = c1;
  cb2 = c2;
  b1 = P.a1 * Wn;
  b2 = Wn - b1;
  bal = Wn > kLeanMinW;
}

// retirement month — simulation.py:644-796 (cnet = E - sum(indexed amount * (1 - tax)), fixed = sum of the locked
// nominal non-indexed payments * (1 - tax): need = max(0, cnet * level - fixed); the caller passes both halved)
template <class C>
MCR_DEV void lean_decumulate(const DevParams& P, const Factors& cur, double cnet_h, double fixed_h, double level_ret,
                             double& b1, double& cb1, double& b2, double& cb2, double& level, double& yr_gross,
                             double& yr_real, bool& bal) {
  const double level0 = level;
  const double half_need = fma(cnet_h, level0, -fixed_h);   // (E*level - income) / 2, exactly
  const double need = half_need + fabs(half_need);          // max(0, E*level - income) (:679-682)
  const double G2 = cur.GI * cur.GP;
  const double nb1 = b1 * cur.G1;
  const double nb2 = b2 * G2;
  const double lv = level0 * cur.GI;
  const double V = nb1 + nb2;
  double c1 = cb1, c2 = cb2;
  const double tx1 = C::taxed1(P) ? latent_tax(nb1 - c1, P.hrate1) : 0.0;   // tax due on full liquidation
  const double tx2 = C::taxed2(P) ? latent_tax(nb2 - c2, P.hrate2) : 0.0;
  double gross, q;
  if constexpr (C::kAnyTaxed) {
    // both withdrawals at once (:750-777): net target split by w_i = cap_i / cap and grossed up by
    // b_i / cap_i, i.e. ONE fraction f = need / cap of every balance, basis and latent tax is sold
    const double cap = V - (tx1 + tx2);
    const double f = need * Math<true>::rcp3(cap);   // (feeds the balance every month: full accuracy)
    q = 1.0 - f;
    gross = f * V;
  } else {
    gross = need;  // no tax: gross == net == need (< V / 2)
    q = 1.0;       // unused
  }
  yr_gross += gross;
  yr_real = fma(gross * level_ret, Math<true>::rcp2(level0), yr_real);      // :778-782
  // The rebalance's fraction sold is invariant under the common scale q, so it is computed from
  // the pre-withdrawal values (its reciprocal then runs beside the one above) and q applied once.
  const double Wu = lean_rebalance<C>(P, nb1, nb2, V, tx1, tx2, c1, c2);
  double Wn;
  if constexpr (C::kAnyTaxed) {
    cb1 = q * c1;
    cb2 = q * c2;
    Wn = q * Wu;
  } else {
    Wn = Wu - need;
  }
  b1 = P.a1 * Wn;
  b2 = Wn - b1;
  level = lv;
  bal = Wn > fma(P.lean_need_coef, lv, kLeanMinW);   // W > kLeanMinW and next month's need < cap / 2
}

// (re-)entry test after a general month / at a phase change: on target, big enough, in range
MCR_DEV bool lean_ready(const DevParams& P, double b1, double b2, double level, bool retired) {
  const double W = b1 + b2;
  bool ok = P.lean_cfg_ok != 0 && W > kLeanMinW && fabs(fma(-P.a1, W, b1)) <= 1e-13 * W;
  if (retired)
    ok = ok && W > fma(P.lean_need_coef, level, kLeanMinW) && level > kLeanLevelMin && level < P.lean_level_max;
  return ok;
}

// ---------------------------------------------------------------------------------------------
// the timeline
// ---------------------------------------------------------------------------------------------
// Software pipeline: the draws and gross factors of month a+1 are computed while month a is
// being stepped. They depend only on (seed, path, a+1), never on the path state, so this is a
// pure scheduling change: inside month a's basic blocks the compiler now has independent INT
// (Philox), FP32/MUFU (Box-Muller) and FP64 (three exp polynomials) work to interleave with
// the long dependent chains of the withdrawals and the rebalance.
template <bool FAST, class C, class Shock, class Sink>
MCR_DEV void run_timeline(const DevParams& P, const int wm,
                                             const int32_t* __restrict__ window, Shock& shock,
                                             Sink& sink, PathOut& o, int& years_observed) {
  const int R = P.R;
  double b1 = P.B0 * P.a1;                                   // :499-502
  double b2 = P.B0 - b1;
  double cb1 = b1, cb2 = b2;
  double contrib = P.C0;
  double g1 = 0.0, g2 = 0.0;
  double level = 1.0;
  bool pre_fail = false;
  uint32_t executed = 0;
  int t = 0;
  sink.point(t++, P.B0, 1.0);                                // :490-492

  Factors nxt;  // factors of the next month to be stepped (absolute month 0 first)
  {
    double ze, zi, zp;
    shock.next(ze, zi, zp);
    Math<FAST>::template factors<C>(P, ze, zi, zp, nxt.G1, nxt.GI, nxt.GP);
  }

  constexpr bool kLean = FAST && C::kLean;
  // lean mode (per path): the portfolio is exactly on target and far from every threshold, so the
  // month can take the straight-line step; re-evaluated after every month
  [[maybe_unused]] bool bal = false;
  if constexpr (kLean) bal = P.lean_cfg_ok != 0 && P.B0 > kLeanMinW;
  double k1 = contrib * P.a1;                                        // :540-547
  double k2 = contrib - k1;

  // ---- accumulation — :513-579
  int moy = 0;  // (m-1) % 12
  for (int m = 1; m <= wm; ++m) {
    if (moy == 0 && m > 1 && P.growth_on) {                          // :514-517
      contrib *= P.growth1p;
      k1 = contrib * P.a1;
      k2 = contrib - k1;
    }
    const Factors cur = nxt;
    double ze, zi, zp;
    shock.next(ze, zi, zp);                                        // draws of the FOLLOWING month
    ++executed;
    bool lean_month = false;
    if constexpr (kLean) {
      if (__builtin_expect(bal, 1)) {
        lean_accumulate<C>(P, cur, k1, k2, b1, cb1, b2, cb2, level, bal);
        MCR_COUNT_LEAN_MONTH();
        Math<FAST>::template factors<C>(P, ze, zi, zp, nxt.G1, nxt.GI, nxt.GP);
        lean_month = true;
      }
    }
    if (!lean_month) {
      const double G2 = cur.GI * cur.GP;                             // :532
      if (C::annual(P)) {
        g1 += b1 * (cur.G1 - 1.0);                                   // :534-535
        g2 += b2 * (G2 - 1.0);
      }
      b1 *= cur.G1;
      b2 *= G2;
      level *= cur.GI;
      b1 += k1; cb1 += k1;
      b2 += k2; cb2 += k2;
      if (__builtin_expect(!rebalance_skips(P, b1, b2), 1)) {        // :549-553
        rebalance_main<FAST, false, C>(P, b1, cb1, b2, cb2);
        Math<FAST>::template factors<C>(P, ze, zi, zp, nxt.G1, nxt.GI, nxt.GP);
      } else {
        Math<FAST>::template factors<C>(P, ze, zi, zp, nxt.G1, nxt.GI, nxt.GP);
      }
      if constexpr (kLean) bal = lean_ready(P, b1, b2, level, false);
    }
    if (++moy == kMPY) {                                           // m % 12 == 0 — :557-579
      moy = 0;
      // lean mode: no annual tax and the portfolio is on target, so the trailing rebalance of the
      // (empty) annual-tax event early-outs in the reference (:432-442 with |drift| <= eps)
      if (!(kLean && lean_month)) {
        if (annual_tax<FAST, false, C>(P, b1, cb1, b2, cb2, g1, g2)) pre_fail = true;
      }
      sink.point(t++, b1 + b2, level);
      g1 = 0.0; g2 = 0.0;
    }
  }

  const double S0 = b1 + b2;                                       // :581-582
  const double level_ret = level;
  if (wm > 0 && moy != 0) sink.point(t++, S0, level_ret);          // :590-594

  // lock state of non-indexed streams: the first two live in registers, the rest in local memory
  double lock0 = 0.0, lock1 = 0.0;
  bool locked0 = false, locked1 = false;
  double lockn[MCR_MAX_STREAMS];
  uint32_t lockn_mask = 0;
  const int ns = P.n_streams;
  // fast build: cached stream sums between two window boundaries, kept HALVED (exact scaling):
  // need = max(0, E*level - income) = t + |t| with t = cnet_h * level - fixed_h
  [[maybe_unused]] double cnet_h = 0.5 * P.E;  // (E - sum of indexed amount * (1 - tax)) / 2
  [[maybe_unused]] double fixed_h = 0.0;       // (sum of locked nominal payments * (1 - tax)) / 2
  int next_event = 0;                          // next retirement month at which they change
  if constexpr (kLean) bal = !pre_fail && lean_ready(P, b1, b2, level, true);
  double fy_gross = 0.0, fy_real = 0.0;
  bool ok = !pre_fail;                                             // :627-629
  int ruin = pre_fail ? 0 : -1;
  int tax_moy = moy;  // (absolute month) % 12 of the last completed month
  int y = 0;
  int n_obs = 0;      // retirement years with a withdrawal-rate observation

  // ---- decumulation — :632-868
  for (; y < R && !pre_fail; ++y) {
    double yr_g1 = 0.0, yr_g2 = 0.0, yr_real = 0.0;   // fast build: yr_g1 carries both assets' gross
    bool failed = false;
    int r = y * kMPY;
    if constexpr (kLean) {  // price level range of the lean months, valid for the next 12 of them
      if (!(level > kLeanLevelMin && level < P.lean_level_max)) bal = false;
    }
    for (int j = 0; j < kMPY; ++j, ++r) {
      const double level0 = level;                                 // :644-647
      double income = 0.0;                                         // :649-677
      if constexpr (FAST) {
        // The set of paying streams changes only at window boundaries (warp-uniform months):
        // between two boundaries net income = level * sum(indexed amount*(1-tax)) + sum(locked
        // nominal*(1-tax)), one DFMA per month. The sums are rebuilt at a boundary month.
        if (__builtin_expect(r == next_event, 0)) {
          double idx_coeff = 0.0, fixed_income = 0.0;
          int nxt_ev = 0x7fffffff;
          for (int k = 0; k < ns; ++k) {
            const int first = window[2 * k], end = window[2 * k + 1];
            if (first > r && first < nxt_ev) nxt_ev = first;
            if (end > r && end < nxt_ev) nxt_ev = end;
            if (r >= first && r < end) {
              if (P.streams[k].indexed) {
                idx_coeff += P.streams[k].amount * P.streams[k].net_factor;
              } else {
                if (!((lockn_mask >> k) & 1u)) { lockn[k] = P.streams[k].amount * level0; lockn_mask |= 1u << k; }
                fixed_income += lockn[k] * P.streams[k].net_factor;
              }
            }
          }
          next_event = nxt_ev;
          cnet_h = 0.5 * (P.E - idx_coeff);
          fixed_h = 0.5 * fixed_income;
        }
        if constexpr (kLean) {
          if (__builtin_expect(bal, 1)) {
            const Factors cur = nxt;
            double ze, zi, zp;
            shock.next(ze, zi, zp);                                // draws of the FOLLOWING month
            ++executed;
            lean_decumulate<C>(P, cur, cnet_h, fixed_h, level_ret, b1, cb1, b2, cb2, level, yr_g1, yr_real, bal);
            MCR_COUNT_LEAN_MONTH();
            Math<FAST>::template factors<C>(P, ze, zi, zp, nxt.G1, nxt.GI, nxt.GP);
            // the month cannot fail, and the (empty) annual-tax event at a tax-year end reduces to
            // a rebalance that early-outs on an on-target portfolio (:798-822, :432-442)
            if (++tax_moy == kMPY) tax_moy = 0;
            continue;
          }
        }
      } else {
        if (ns > 0) stream_income(P.streams[0], window[0], window[1], r, level0, lock0, locked0, income);
        if (ns > 1) stream_income(P.streams[1], window[2], window[3], r, level0, lock1, locked1, income);
        for (int k = 2; k < ns; ++k) {
          bool lk = (lockn_mask >> k) & 1u;
          double lv = lk ? lockn[k] : 0.0;
          const bool was = lk;
          stream_income(P.streams[k], window[2 * k], window[2 * k + 1], r, level0, lv, lk, income);
          if (lk && !was) { lockn[k] = lv; lockn_mask |= 1u << k; }
        }
      }
      double need;                                                 // :679-682
      if constexpr (FAST) {
        const double t = fma(cnet_h, level0, -fixed_h);
        need = pmax(0.0, t + t);
      } else {
        const double need_nominal = P.E * level0;
        need = pmax(0.0, need_nominal - income);
      }
      const bool wants = need > kEps;
      if (__builtin_expect(b1 + b2 <= kEps && wants, 0)) { failed = true; break; }  // :684-690

      const Factors cur = nxt;
      double ze, zi, zp;
      shock.next(ze, zi, zp);                                      // draws of the FOLLOWING month
      ++executed;
      const double G2 = cur.GI * cur.GP;
      if (C::annual(P)) {
        g1 += b1 * (cur.G1 - 1.0);                                 // :706-711
        g2 += b2 * (G2 - 1.0);
      }
      b1 *= cur.G1;
      b2 *= G2;
      level *= cur.GI;
      if (__builtin_expect(b1 + b2 <= kEps && wants, 0)) {         // :715-724
        b1 = pmax(0.0, b1);
        b2 = pmax(0.0, b2);
        failed = true;
        break;
      }
      double cap1, cap2, tx1 = 0.0, tx2 = 0.0;
      const bool algebra = FAST && C::algebra(P);
      if (algebra) {
        tx1 = C::taxed1(P) ? pmax(0.0, b1 - cb1) * P.rate1 : 0.0;   // tax due on full liquidation of each asset
        tx2 = C::taxed2(P) ? pmax(0.0, b2 - cb2) * P.rate2 : 0.0;
        cap1 = b1 > kEps ? b1 - tx1 : 0.0;     // == net_liq()
        cap2 = b2 > kEps ? b2 - tx2 : 0.0;
      } else {
        cap1 = net_liq(b1, cb1, C::taxed1(P), P.rate1); // :726-737
        cap2 = net_liq(b2, cb2, C::taxed2(P), P.rate2);
      }
      const double cap = cap1 + cap2;
      const double target = pmin(need, cap);   // :739-742 identity: max(0, .), need >= +0.0 and cap >= +0.0
      const double need_lo = need - kEps;
      double gw1, nw1, gw2, nw2;
      bool closed = false;
      if (algebra) {
        // Closed form of :750-777 when both assets really sell (b_i > eps, target > 0, cap > eps):
        // the net target is split by w_i = cap_i/cap and grossed up by 1/(1 - gf_i*rate_i) =
        // b_i/cap_i, so gross_i = target*b_i/cap = f*b_i with ONE fraction f = target/cap <= 1
        // for both assets; then basis removed = f*cb_i, tax = f*tx_i, net cash = f*cap_i. One
        // reciprocal instead of five; max(eps, 1 - etf) is inactive because rate < 1.
        closed = b1 > kEps && b2 > kEps && cap > kEps && target > 0;
        if (__builtin_expect(closed, 1)) {
          const double f = pmin(1.0, target * Math<FAST>::rcp(cap));
          gw1 = f * b1;
          gw2 = f * b2;
          nw1 = f * cap1;
          nw2 = f * cap2;
          double nb1 = b1 - gw1, ncb1 = cb1 - f * cb1, nb2 = b2 - gw2, ncb2 = cb2 - f * cb2;
          if (__builtin_expect(either_le(nb1, nb2, kEps), 0)) {   // :245-247, rarely taken
            if (nb1 <= kEps) { nb1 = 0.0; ncb1 = 0.0; }
            if (nb2 <= kEps) { nb2 = 0.0; ncb2 = 0.0; }
          }
          b1 = nb1; cb1 = ncb1; b2 = nb2; cb2 = ncb2;
        }
      }
      if (!closed) {
        const double w1 = cap > kEps ? Math<FAST>::div(cap1, cap) : P.a1;   // :750-755
        const double w2 = 1.0 - w1;
        withdraw_pair<FAST, C>(P, b1, cb1, target * w1, b2, cb2, target * w2, gw1, nw1, gw2, nw2);  // :757-777
      }
      if constexpr (FAST) {
        yr_g1 += gw1 + gw2;
      } else {
        yr_g1 += gw1;
        yr_g2 += gw2;
      }
      yr_real += Math<FAST>::div((gw1 + gw2) * level_ret, pmax(level0, kEps));   // :778-782
      if (wants && either_lt(target, nw1 + nw2, need_lo)) failed = true;   // :743-748 and :784-790
      if (__builtin_expect(!rebalance_skips(P, b1, b2), 1)) {      // :792-796
        rebalance_main<FAST, false, C>(P, b1, cb1, b2, cb2);
        Math<FAST>::template factors<C>(P, ze, zi, zp, nxt.G1, nxt.GI, nxt.GP);
      } else {
        Math<FAST>::template factors<C>(P, ze, zi, zp, nxt.G1, nxt.GI, nxt.GP);
      }
      if (++tax_moy == kMPY) tax_moy = 0;
      if (!failed && tax_moy == 0) {                               // :798-822
        const bool tf = annual_tax<FAST, false, C>(P, b1, cb1, b2, cb2, g1, g2);
        g1 = 0.0; g2 = 0.0;
        if (tf) failed = true;
      }
      if (__builtin_expect(failed, 0)) { ruin = r + 1; break; }    // :824-828
      if constexpr (kLean) bal = lean_ready(P, b1, b2, level, true);
    }
    const double yr_gross = yr_g1 + yr_g2;
    if (failed) {                                                  // :842-857
      ok = false;
      if (ruin < 0) ruin = r + 1;
      sink.point(t++, pmax(0.0, b1 + b2), level);
      sink.wr(y, nan_value());
      if (y == 0) { fy_gross = yr_gross; fy_real = yr_real; }
      ++y;
      break;
    }
    sink.wr(y, S0 > kEps ? (yr_real / S0) * 100.0 : 0.0);          // :834-840,859
    n_obs = y + 1;
    if (y == 0) { fy_gross = yr_gross; fy_real = yr_real; }
    sink.point(t++, b1 + b2, level);
  }

  // ---- final partial tax period — :873-898
  if (ok && tax_moy != 0) {
    if (annual_tax<FAST, false, C>(P, b1, cb1, b2, cb2, g1, g2)) { ok = false; ruin = R * kMPY; }
    sink.point(t - 1, b1 + b2, level);   // overwrites the last yearly sample
  }

  // ---- padding — :902-937 (failed paths pad 0.0; WR pads NaN)
  if (Sink::kActive) {
    const int T = 1 + (wm + kMPY - 1) / kMPY + R;
    for (; t < T; ++t) sink.point(t, 0.0, 1.0);
    for (; y < R; ++y) sink.wr(y, nan_value());
  }

  o.start_balance = S0;
  o.final_balance = pmax(0.0, b1 + b2);
  o.fy_gross = fy_gross;
  o.fy_real = fy_real;
  o.infl_ret = level_ret;
  o.success = ok ? 1 : 0;
  o.ruin_month = ruin;
  o.executed = executed;
  years_observed = n_obs;
}

}  // namespace mcr
