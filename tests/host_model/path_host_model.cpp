// path_host_model.cpp — TEST INFRASTRUCTURE ONLY. The device headers of the engine
// (csrc/mcr_path.cuh, mcr_rng.cuh, mcr_derive.h) compiled for the HOST with g++, one path at a
// time, so that the `-m "not gpu"` suite can check the arithmetic of the strict, fast and lean
// month steps against the CPU oracle (tests/test_host_model.py) before a GPU is available.
// It is not part of libmcr_b200.so, nothing under monte_carlo_retirement_b200/ loads it, and it
// is no fallback: the product fails loudly without a CUDA device.
#include <cstdint>
#include <cstring>
#include <string>

static long long g_lean_months = 0;
#define MCR_COUNT_LEAN_MONTH() (++g_lean_months)
#include "../../monte_carlo_retirement_b200/csrc/mcr_derive.h"

using namespace mcr;

namespace {

struct HostSink {
  static constexpr bool kActive = true;
  double* traj;
  double* real;
  double* wrp;
  int64_t ld;
  void point(int t, double nominal, double price) {
    if (traj) traj[(int64_t)t * ld] = nominal;
    if (real) real[(int64_t)t * ld] = price > kEps ? nominal / price : 0.0;
  }
  void wr(int y, double v) {
    if (wrp) wrp[(int64_t)y * ld] = v;
  }
};

template <bool FAST, class C, class Shock>
void one(const DevParams& P, int wm, const int32_t* window, Shock& sh, HostSink& sink, PathOut& o) {
  int years = 0;
  run_timeline<FAST, C>(P, wm, window, sh, sink, o, years);
}

template <bool FAST, class Shock>
void dispatch(int cfg, const DevParams& P, int wm, const int32_t* window, Shock& sh, HostSink& sink, PathOut& o) {
  switch (cfg) {
    case 1: one<FAST, CfgBothTaxed>(P, wm, window, sh, sink, o); break;
    case 2: one<FAST, CfgNoTax>(P, wm, window, sh, sink, o); break;
    case 3: one<FAST, CfgBothTaxedSmall>(P, wm, window, sh, sink, o); break;
    case 4: one<FAST, CfgNoTaxSmall>(P, wm, window, sh, sink, o); break;
    case 5: one<FAST, CfgBothTaxedTight>(P, wm, window, sh, sink, o); break;
    case 6: one<FAST, CfgNoTaxTight>(P, wm, window, sh, sink, o); break;
    default: one<FAST, CfgGeneric>(P, wm, window, sh, sink, o); break;
  }
}

}  // namespace

extern "C" {

// months that took the lean step since the last call (test statistic)
long long hm_lean_months(void) {
  const long long v = g_lean_months;
  g_lean_months = 0;
  return v;
}

// the bound on |monthly log-return| the host proves for this scenario's own draws (0: none)
double hm_small_bound(const mcr_params* params) {
  DevParams P;
  std::string why;
  double live_age[MCR_MAX_STREAMS] = {0};
  int32_t live_dur[MCR_MAX_STREAMS] = {0};
  if (validate_and_derive(*params, P, live_age, live_dur, why) != MCR_OK) return -1.0;
  return P.exp_small == 2 ? 0.05 : (P.exp_small == 1 ? 0.1 : 0.0);
}

// mode: 0 strict, 1 fast, 2 fast + MCR_FLAG_SMALL_RETURNS. shocks: the device replay layout
// [(m*3 + c) * ld + i]. Outputs: cols[5][n] (start, final, first-year gross, first-year real,
// inflation at retirement), success[n], ruin[n], executed[n], series [T or R][series_ld] (may be NULL).
int hm_replay(const mcr_params* params, int32_t wm, const double* shocks, int64_t ld, int32_t n_months, int64_t n,
              int mode, double* cols, uint8_t* success, int32_t* ruin, uint32_t* executed, double* traj,
              double* real, double* wr, int64_t series_ld, int32_t* cfg_used) {
  DevParams P;
  std::string why;
  double live_age[MCR_MAX_STREAMS] = {0};
  int32_t live_dur[MCR_MAX_STREAMS] = {0};
  if (validate_and_derive(*params, P, live_age, live_dur, why) != MCR_OK) return -1;
  int32_t window[2 * MCR_MAX_STREAMS];
  stream_windows(P, params->current_age, live_age, live_dur, wm, window);
  const bool fast = mode != 0;
  const int cfg = pick_cfg_index(P, fast, mode == 2 ? P.exp_small : 0);
  if (cfg_used) *cfg_used = cfg;
  for (int64_t i = 0; i < n; ++i) {
    ReplayShock sh{shocks + i, ld, n_months};
    HostSink sink{traj ? traj + i : nullptr, real ? real + i : nullptr, wr ? wr + i : nullptr, series_ld};
    PathOut o = {};
    if (fast) dispatch<true>(cfg, P, wm, window, sh, sink, o);
    else dispatch<false>(cfg, P, wm, window, sh, sink, o);
    cols[0 * n + i] = o.start_balance;
    cols[1 * n + i] = o.final_balance;
    cols[2 * n + i] = o.fy_gross;
    cols[3 * n + i] = o.fy_real;
    cols[4 * n + i] = o.infl_ret;
    success[i] = (uint8_t)o.success;
    ruin[i] = o.ruin_month;
    if (executed) executed[i] = o.executed;
  }
  return 0;
}

// the native draw stream in the replay layout (strict transform: logf / sqrtf / sincospi)
int hm_draw(const mcr_params* params, uint64_t main_seed, uint32_t seed_stream, int64_t first_path, int64_t n,
            int32_t n_months, double* shocks, int64_t ld) {
  DevParams P;
  std::string why;
  double live_age[MCR_MAX_STREAMS] = {0};
  int32_t live_dur[MCR_MAX_STREAMS] = {0};
  if (validate_and_derive(*params, P, live_age, live_dur, why) != MCR_OK) return -1;
  uint32_t k0, k1;
  philox_key_from_seed(main_seed, k0, k1);
  PhiloxKeys keys;
  philox_expand_keys(k0, k1, keys);
  for (int64_t i = 0; i < n; ++i) {
    const uint64_t gp = (uint64_t)(first_path + i);
    PhiloxShock<false> sh{keys, (uint32_t)gp, (uint32_t)(gp >> 32), 0u, seed_stream, P.rho_f, P.rho_c_f, P.rho, P.rho_c, 0.f, 0.f, 0.f};
    for (int m = 0; m < n_months; ++m) {
      double ze, zi, zp;
      sh.next(ze, zi, zp);
      shocks[(int64_t)(3 * m + 0) * ld + i] = ze;
      shocks[(int64_t)(3 * m + 1) * ld + i] = zi;
      shocks[(int64_t)(3 * m + 2) * ld + i] = zp;
    }
  }
  return 0;
}

}  // extern "C"
