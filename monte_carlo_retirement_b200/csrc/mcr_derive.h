// mcr_derive.h — host-side derivation of the kernel's scenario constants from the C-ABI POD
// (plain C++, no CUDA): validation, the monthly log parameters, tax switches, the lean-month
// guards, the live income streams and their activity windows, and the Philox key. Used by
// mcr_api.cu (the product) and by tests/host_model (the header compiled for the host so the CPU
// suite can check the fast arithmetic against the oracle).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>

#include "mcr_path.cuh"
#include "mcr_rng.cuh"

namespace mcr {

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

inline void philox_key_from_seed(uint64_t main_seed, uint32_t& k0, uint32_t& k1) {
  const uint64_t key = splitmix64(splitmix64(main_seed) ^ 0x6D63725F62323030ull /* "mcr_b200" */);
  k0 = (uint32_t)key;
  k1 = (uint32_t)(key >> 32);
}

// stream_payment_start_month_index — simulation.py:47-63, same double operations as CPython
inline int32_t start_month_index(double current_age, int32_t wm, double start_at_age) {
  const double ret = current_age + (double)wm / (double)MCR_MONTHS_PER_YEAR;
  const double elig = (start_at_age > ret) ? start_at_age : ret;
  const double k = std::ceil((elig - ret) * (double)MCR_MONTHS_PER_YEAR - MCR_SMALL_EPSILON);
  if (!(k > 0)) return 0;
  return k > 2147483647.0 ? 2147483647 : (int32_t)k;
}

inline int validate_and_derive(const mcr_params& p, DevParams& d, double* live_start_age, int32_t* live_duration,
                        std::string& why) {
  auto bad = [&](const char* m) { why = m; return MCR_EINVAL; };
  if (!(p.initial_balance >= 0) || !(p.monthly_contribution >= 0) || !(p.monthly_expenses >= 0))
    return bad("balances, contributions and expenses must be >= 0");
  if (!(p.contribution_growth_rate_annual >= 0)) return bad("contribution_growth_rate_annual must be >= 0");
  if (!(p.allocation_inv1_pct >= 0 && p.allocation_inv1_pct <= 1)) return bad("allocation_inv1_pct must be in [0,1]");
  if (p.retirement_years <= 0) return bad("retirement_years must be > 0");
  if (p.retirement_years > 1000) return bad("retirement_years above 1000 is not supported");
  if (p.n_streams < 0 || p.n_streams > MCR_MAX_STREAMS)
    return bad("at most 16 other_income_streams are supported by the CUDA engine");
  if (!(p.inv1_sigma_log >= 0) || !(p.inf_sigma_log >= 0) || !(p.prem_sigma_log >= 0)) return bad("negative sigma_log");
  if (!(p.equity_inflation_rho >= -1 && p.equity_inflation_rho <= 1)) return bad("correlation must be in [-1,1]");
  const double rates[4] = {p.inv1_annual_tax_on_gains_rate, p.inv1_realized_gains_tax_rate,
                           p.inv2_annual_tax_on_gains_rate, p.inv2_realized_gains_tax_rate};
  for (double r : rates)
    if (!(r >= 0 && r <= 1)) return bad("tax rates must be in [0,1]");
  std::memset(&d, 0, sizeof(d));
  d.B0 = p.initial_balance;
  d.C0 = p.monthly_contribution;
  d.growth1p = 1 + p.contribution_growth_rate_annual;            // simulation.py:517
  d.E = p.monthly_expenses;
  d.a1 = p.allocation_inv1_pct;
  d.a2 = 1.0 - p.allocation_inv1_pct;                            // config.py:124-126
  const double mpy = (double)MCR_MONTHS_PER_YEAR;
  const double root = std::sqrt(mpy);
  d.mu1 = p.inv1_mu_log / mpy;  d.sg1 = p.inv1_sigma_log / root; // simulation.py:472-474
  d.muI = p.inf_mu_log / mpy;   d.sgI = p.inf_sigma_log / root;
  d.muP = p.prem_mu_log / mpy;  d.sgP = p.prem_sigma_log / root;
  d.rho = p.equity_inflation_rho;
  const double c2 = 1.0 - d.rho * d.rho;
  d.rho_c = std::sqrt(c2 > 0.0 ? c2 : 0.0);                      // simulation.py:463
  d.rho_f = (float)d.rho;
  d.rho_c_f = (float)d.rho_c;
  d.rate1 = p.inv1_realized_gains_tax_rate;
  d.rate2 = p.inv2_realized_gains_tax_rate;
  d.ann1 = p.inv1_annual_tax_on_gains_rate;
  d.ann2 = p.inv2_annual_tax_on_gains_rate;
  d.use1 = p.inv1_use_realized_gains_tax_system != 0;
  d.use2 = p.inv2_use_realized_gains_tax_system != 0;
  d.taxed1 = d.use1 && d.rate1 > 0;
  d.taxed2 = d.use2 && d.rate2 > 0;
  d.growth_on = p.contribution_growth_rate_annual > 0;
  d.algebra_ok = (d.taxed1 || d.taxed2) && (!d.taxed1 || d.rate1 <= 0.999) && (!d.taxed2 || d.rate2 <= 0.999);
  d.annual_any = (!d.use1 && d.ann1 > 0) || (!d.use2 && d.ann2 > 0);
  {
    // |mu/12| + sigma/sqrt(12) * z_max for the three factors; Box-Muller on a 26-bit radius
    // uniform gives r <= sqrt(-2 ln 2^-27) = 6.12, and the inflation shock rho*n0 + rho_c*n1 of ONE
    // pair (n0, n1) = r (cos t, sin t) is r cos(t - phi) with rho^2 + rho_c^2 = 1: bounded by r too
    const double zmax = 6.2;
    const double b1 = std::fabs(d.mu1) + d.sg1 * zmax, bi = std::fabs(d.muI) + d.sgI * zmax,
                 bp = std::fabs(d.muP) + d.sgP * zmax;
    const double worst = std::max(b1, std::max(bi, bp));
    d.exp_small = worst < 0.05 ? 2 : (worst < 0.1 ? 1 : 0);
  }
  {
    // lean months of the fast build (mcr_path.cuh "lean months"): allocation away from the corners
    // and every biting realized-gains rate <= 0.9, so that no balance of an on-target portfolio of
    // more than kLeanMinW dollars can come near the reference's 1e-6 snaps
    const double rmax = std::max(d.taxed1 ? d.rate1 : 0.0, d.taxed2 ? d.rate2 : 0.0);
    d.lean_cfg_ok = (d.a1 >= 0.01 && d.a1 <= 0.99 && rmax <= 0.9) ? 1 : 0;
    d.lean_cg = 0.5 * (1.0 - rmax) * std::exp(-0.2);
    d.lean_need_coef = d.E / d.lean_cg;
    // need <= E * level must stay below 1e8 for the 12 months a yearly check covers (|monthly
    // log-return| < 0.1); huge needs could trip the reference's net-cash test by rounding alone
    d.lean_level_max = d.E > 0.0 ? 1e8 / (d.E * std::exp(1.2)) : 1e300;
    d.hrate1 = 0.5 * d.rate1;
    d.hrate2 = 0.5 * d.rate2;
  }
  d.R = p.retirement_years;
  int live = 0;
  for (int k = 0; k < p.n_streams; ++k) {
    const mcr_income_stream& s = p.streams[k];
    if (!(s.monthly_amount_today >= 0) || !(s.tax_rate >= 0 && s.tax_rate <= 1) || !(s.start_at_age >= 0))
      return bad("bad other_income_streams entry");
    // a zero amount pays nominal 0.0 and adds +0.0 to the income sum (exact); a zero duration
    // is never active (simulation.py:653-656): neither can change any result.
    if (s.monthly_amount_today == 0.0 || s.duration_years == 0) continue;
    d.streams[live].amount = s.monthly_amount_today;
    d.streams[live].net_factor = 1.0 - s.tax_rate;               // simulation.py:675-677
    d.streams[live].duration = s.duration_years < 0 ? -1 : s.duration_years * MCR_MONTHS_PER_YEAR;
    d.streams[live].indexed = s.inflation_indexed != 0;
    live_start_age[live] = s.start_at_age;
    live_duration[live] = d.streams[live].duration;
    ++live;
  }
  d.n_streams = live;
  return MCR_OK;
}

// per live stream: [first paying retirement month, end) — simulation.py:602-621,653-656
inline void stream_windows(const DevParams& dev, double current_age, const double* live_start_age,
                           const int32_t* live_duration, int32_t wm, int32_t* out) {
  for (int k = 0; k < MCR_MAX_STREAMS; ++k) {
    int32_t first = 0, end = 0;
    if (k < dev.n_streams) {
      first = start_month_index(current_age, wm, live_start_age[k]);
      const int64_t e = live_duration[k] < 0 ? 2147483647ll : (int64_t)first + live_duration[k];
      end = e > 2147483647ll ? 2147483647 : (int32_t)e;
    }
    out[2 * k] = first;
    out[2 * k + 1] = end;
  }
}

// which compile-time specialisation of the timeline matches this scenario (0 generic, 1 both
// taxed, 2 no tax; +2 / +4 when the monthly log-returns are bounded by 0.1 / 0.05 — fast build only)
inline int pick_cfg_index(const DevParams& P, bool fast, int small_level) {
  const int small = fast ? 2 * small_level : 0;   // 0: none, 1: < 0.1, 2: < 0.05
  // the both-taxed specialisation uses closed forms that need 1 - rate > eps
  if (P.taxed1 && P.taxed2 && !P.annual_any && P.rate1 <= 0.999 && P.rate2 <= 0.999) return 1 + small;
  if (!P.taxed1 && !P.taxed2 && !P.annual_any) return 2 + small;
  return 0;
}

}  // namespace mcr
