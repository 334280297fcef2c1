/*
 * path_oracle.c — CPU ORACLE for the Monte Carlo path engine. TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it. The shipped engine is the CUDA
 * library in monte_carlo_retirement_b200/csrc and has no CPU fallback.
 *
 * It is a plain-C restatement (IEEE double, no FMA contraction: build with
 * -ffp-contract=off, libm exp == CPython math.exp) of the reference algorithm in
 * /root/reference/backend/simulation.py; each function cites the lines it follows. The shock
 * draws (numpy SeedSequence + PCG64 + ziggurat, third-party, numpy 2.4.2 pinned in uv.lock,
 * 2.3.5 installed) are NOT restated: the oracle consumes a precomputed (n_months, 3) shock
 * matrix exactly as `_draw_shock_path` returns it, so it is bit-for-bit comparable with the
 * reference on the reference's own draws.
 *
 * Parity status: PINNED. tests/test_oracle_golden.py checks this file against
 *   (1) the reference's own known-answer tests (tests/test_simulation_correctness.py), restated;
 *   (2) the .npz files under tests/golden — outputs of the unmodified Python reference produced in the build
 *       container by tests/golden/make_golden.py (bit-exact equality is required).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mcr.h"

#define EPS MCR_SMALL_EPSILON
#define MPY MCR_MONTHS_PER_YEAR

/* CPython's builtin max(a, b) / min(a, b): the FIRST argument wins ties and NaN comparisons. */
static inline double py_max(double a, double b) { return (b > a) ? b : a; }
static inline double py_min(double a, double b) { return (b < a) ? b : a; }

/* ---- pure helpers ---------------------------------------------------------------------- */

/* stream_payment_start_month_index — simulation.py:47-63 (with retirement_age :32-34 and
 * stream_payment_start_age :37-44 folded in). */
int32_t oracle_stream_start_month(double current_age, int32_t working_months, double start_at_age) {
  double retirement_start = current_age + (double)working_months / (double)MPY;
  double eligible = py_max(retirement_start, start_at_age);
  double k = ceil((eligible - retirement_start) * (double)MPY - EPS);
  int32_t ki = (int32_t)k;
  return ki > 0 ? ki : 0;
}

/* expected trajectory length — simulation.py:585-589,902 */
int32_t oracle_trajectory_len(int32_t working_months, int32_t retirement_years) {
  int32_t wy = working_months > 0 ? (working_months + MPY - 1) / MPY : 0;
  return 1 + wy + retirement_years;
}

/* _net_liquidation_value — simulation.py:256-272 */
double oracle_net_liquidation(double balance, double cost_basis, int use_realized, double rate) {
  if (balance <= EPS) return 0.0;
  double taxable_gain = py_max(0.0, balance - cost_basis);
  double tax = (use_realized && rate > 0) ? taxable_gain * rate : 0.0;
  return py_max(0.0, balance - tax);
}

/* _calculate_withdrawal_and_update — simulation.py:201-254.
 * out = {new_balance, new_cost_basis, gross_withdrawal, net_cash_delivered} */
void oracle_withdraw(double bal, double cb, double net_target, int use_real_tax, double rate,
                     double out[4]) {
  if (bal <= EPS || net_target <= 0) { /* :218-219 */
    out[0] = py_max(0.0, bal);
    out[1] = py_max(0.0, cb);
    out[2] = 0.0;
    out[3] = 0.0;
    return;
  }
  int taxed = use_real_tax && rate > 0;
  double gain_fraction = py_max(0.0, bal - cb) / bal;               /* :221 */
  double eff_tax_fraction = taxed ? gain_fraction * rate : 0.0;     /* :222-226 */
  double net_fraction = py_max(EPS, 1.0 - eff_tax_fraction);        /* :227 */
  double gross = py_min(net_target / net_fraction, bal);            /* :228-231 */
  double fraction_sold = py_min(1.0, gross / bal);                  /* :233 */
  double basis_removed = py_min(cb, cb * fraction_sold);            /* :234 */
  double taxable_gain = py_max(0.0, gross - basis_removed);         /* :235 */
  double tax_paid = taxed ? taxable_gain * rate : 0.0;              /* :236-240 */
  double net_cash = py_max(0.0, gross - tax_paid);                  /* :241 */
  double nb = py_max(0.0, bal - gross);                             /* :243 */
  double ncb = py_max(0.0, cb - basis_removed);                     /* :244 */
  if (nb <= EPS) { nb = 0.0; ncb = 0.0; }                           /* :245-247 */
  out[0] = nb; out[1] = ncb; out[2] = gross; out[3] = net_cash;
}

/* _rebalance_portfolio — simulation.py:274-359. s = {b1, cb1, b2, cb2}, updated in place. */
void oracle_rebalance(const mcr_params* p, double s[4]) {
  double b1 = s[0], cb1 = s[1], b2 = s[2], cb2 = s[3];
  double a1 = p->allocation_inv1_pct;
  double a2 = 1.0 - a1; /* config.py:124-126 */
  double total = b1 + b2;
  if (total <= EPS) return;                                         /* :290-291 */
  double drift1 = b1 - total * a1;                                  /* :293-294 */
  if (fabs(drift1) <= EPS) return;                                  /* :295-296 */
  double nb1, ncb1, nb2, ncb2;
  if (drift1 > 0) { /* sell asset 1 — :298-325 */
    double gf = py_max(0.0, b1 - cb1) / b1;
    double tpd = p->inv1_use_realized_gains_tax_system ? gf * p->inv1_realized_gains_tax_rate : 0.0;
    double den = py_max(EPS, 1.0 - a1 * tpd);
    double sale = py_min(b1, drift1 / den);
    double fs = sale / b1;
    double br = py_min(cb1, cb1 * fs);
    double tg = py_max(0.0, sale - br);
    double tax = p->inv1_use_realized_gains_tax_system ? tg * p->inv1_realized_gains_tax_rate : 0.0;
    double buy = sale - tax;
    nb1 = py_max(0.0, b1 - sale);
    ncb1 = py_max(0.0, cb1 - br);
    nb2 = b2 + buy;
    ncb2 = cb2 + buy;
  } else { /* sell asset 2 — :326-353; note drift2 is recomputed, not -drift1 (:328) */
    double drift2 = b2 - total * a2;
    double gf = py_max(0.0, b2 - cb2) / b2;
    double tpd = p->inv2_use_realized_gains_tax_system ? gf * p->inv2_realized_gains_tax_rate : 0.0;
    double den = py_max(EPS, 1.0 - a2 * tpd);
    double sale = py_min(b2, drift2 / den);
    double fs = sale / b2;
    double br = py_min(cb2, cb2 * fs);
    double tg = py_max(0.0, sale - br);
    double tax = p->inv2_use_realized_gains_tax_system ? tg * p->inv2_realized_gains_tax_rate : 0.0;
    double buy = sale - tax;
    nb2 = py_max(0.0, b2 - sale);
    ncb2 = py_max(0.0, cb2 - br);
    nb1 = b1 + buy;
    ncb1 = cb1 + buy;
  }
  if (nb1 <= EPS) { nb1 = 0.0; ncb1 = 0.0; }                        /* :355-358 */
  if (nb2 <= EPS) { nb2 = 0.0; ncb2 = 0.0; }
  s[0] = nb1; s[1] = ncb1; s[2] = nb2; s[3] = ncb2;
}

/* _apply_annual_gain_taxes — simulation.py:361-450. Returns tax_failed. */
int oracle_annual_tax(const mcr_params* p, double s[4], double gain1, double gain2) {
  int use1 = p->inv1_use_realized_gains_tax_system, use2 = p->inv2_use_realized_gains_tax_system;
  double r1 = p->inv1_realized_gains_tax_rate, r2 = p->inv2_realized_gains_tax_rate;
  double due1 = !use1 ? py_max(0.0, gain1) * p->inv1_annual_tax_on_gains_rate : 0.0; /* :380-384 */
  double due2 = !use2 ? py_max(0.0, gain2) * p->inv2_annual_tax_on_gains_rate : 0.0; /* :385-389 */
  double due = due1 + due2;
  double cap1 = oracle_net_liquidation(s[0], s[1], use1, r1);                        /* :392-403 */
  double cap2 = oracle_net_liquidation(s[2], s[3], use2, r2);
  double cap = cap1 + cap2;
  double pay = py_min(due, cap);                                                     /* :405 */
  int failed = pay < due - EPS;                                                      /* :406 */
  if (cap > EPS && pay > 0) {                                                        /* :408-430 */
    double share1 = cap1 / cap;
    double share2 = 1.0 - share1;
    double w[4];
    oracle_withdraw(s[0], s[1], pay * share1, use1, r1, w);
    s[0] = w[0]; s[1] = w[1];
    double net1 = w[3];
    oracle_withdraw(s[2], s[3], pay * share2, use2, r2, w);
    s[2] = w[0]; s[3] = w[1];
    double net2 = w[3];
    if (net1 + net2 < due - EPS) failed = 1;
  }
  oracle_rebalance(p, s);                                                            /* :432-442 */
  return failed;
}

/* _monthly_gross_from_shock — simulation.py:468-474: exp(mu/12 + sigma/sqrt(12)*z) evaluated as
 * (mu/12) + ((sigma/sqrt(12))*z). */
static inline double monthly_gross(double mu_log, double sigma_log, double z) {
  return exp(mu_log / (double)MPY + sigma_log / sqrt((double)MPY) * z);
}

/* ---- the per-path timeline — simulation.py:476-950 -------------------------------------- */

typedef struct oracle_path_record {
  double start_balance, final_balance, years_to_ruin /* NaN if none */;
  double first_year_gross, first_year_real, inflation_at_ret;
  int32_t success;
  int32_t trajectory_len;
} oracle_path_record;

/*
 * shocks: row-major (n_rows, 3) = (equity, inflation, premium) correlated unit shocks.
 * traj / real_traj: caller buffers of oracle_trajectory_len() doubles; wr: retirement_years
 * doubles. Returns 0, or -1 on bad arguments.
 */
int oracle_run_path(const mcr_params* p, int32_t wm, const double* shocks, int32_t n_rows,
                    oracle_path_record* rec, double* traj, double* real_traj, double* wr) {
  const int R = p->retirement_years;
  const int total_months = wm + R * MPY;                                             /* :487 */
  if (n_rows < (total_months > 1 ? total_months : 1) || p->n_streams > MCR_MAX_STREAMS) return -1;
  const int T = oracle_trajectory_len(wm, R);
  double* price = (double*)malloc(sizeof(double) * (size_t)(T + 2));
  int n_traj = 0, n_wr = 0;
  const double a1 = p->allocation_inv1_pct;

  traj[n_traj] = p->initial_balance; price[n_traj] = 1.0; n_traj++;                  /* :490-492 */
  double years_to_ruin = NAN;
  double s[4];
  s[0] = p->initial_balance * a1;                                                    /* :499-502 */
  s[2] = p->initial_balance - s[0];
  s[1] = s[0];
  s[3] = s[2];
  double contrib = p->monthly_contribution;
  double g1 = 0.0, g2 = 0.0;
  double level = 1.0;
  int shock_idx = 0;
  int pre_fail = 0;

  for (int m = 1; m <= wm; ++m) {                                                    /* :513-579 */
    if ((m - 1) % MPY == 0 && m > 1 && p->contribution_growth_rate_annual > 0)
      contrib *= 1 + p->contribution_growth_rate_annual;
    const double* z = shocks + 3 * (size_t)shock_idx++;
    double G1 = monthly_gross(p->inv1_mu_log, p->inv1_sigma_log, z[0]);
    double GI = monthly_gross(p->inf_mu_log, p->inf_sigma_log, z[1]);
    double GP = monthly_gross(p->prem_mu_log, p->prem_sigma_log, z[2]);
    double G2 = GI * GP;
    g1 += s[0] * (G1 - 1.0);
    g2 += s[2] * (G2 - 1.0);
    s[0] *= G1;
    s[2] *= G2;
    level *= GI;
    double k1 = contrib * a1;
    double k2 = contrib - k1;
    s[0] += k1; s[1] += k1; s[2] += k2; s[3] += k2;
    oracle_rebalance(p, s);
    if (m % MPY == 0) {
      if (oracle_annual_tax(p, s, g1, g2)) pre_fail = 1;
      traj[n_traj] = s[0] + s[2]; price[n_traj] = level; n_traj++;
      g1 = 0.0; g2 = 0.0;
    }
  }

  const double start_balance = s[0] + s[2];                                          /* :581-582 */
  const double level_ret = level;
  if (wm > 0 && wm % MPY != 0) {                                                     /* :590-594 */
    traj[n_traj] = start_balance; price[n_traj] = level_ret; n_traj++;
  }

  int32_t start_month[MCR_MAX_STREAMS];
  int32_t duration[MCR_MAX_STREAMS]; /* -1 == None */
  double locked[MCR_MAX_STREAMS];
  int is_locked[MCR_MAX_STREAMS];
  for (int k = 0; k < p->n_streams; ++k) {                                           /* :602-621 */
    start_month[k] = oracle_stream_start_month(p->current_age, wm, p->streams[k].start_at_age);
    duration[k] = p->streams[k].duration_years < 0 ? -1 : p->streams[k].duration_years * MPY;
    is_locked[k] = 0; locked[k] = 0.0;
  }

  double fy_gross = 0.0, fy_real = 0.0;
  int ok = !pre_fail;                                                                /* :627-629 */
  if (pre_fail) years_to_ruin = 0.0;

  for (int y = 0; y < R && !pre_fail; ++y) {                                         /* :632-868 */
    double yr_g1 = 0.0, yr_g2 = 0.0, yr_real = 0.0;
    int failed = 0;
    int r = 0;
    for (int j = 0; j < MPY; ++j) {
      r = y * MPY + j;
      double level0 = level;
      double need_nominal = p->monthly_expenses * level0;
      double income = 0.0;
      for (int k = 0; k < p->n_streams; ++k) {                                       /* :650-677 */
        int active = r >= start_month[k] && (duration[k] < 0 || r < start_month[k] + duration[k]);
        if (!active) continue;
        double nominal;
        if (p->streams[k].inflation_indexed) {
          nominal = p->streams[k].monthly_amount_today * level0;
        } else {
          if (!is_locked[k]) { locked[k] = p->streams[k].monthly_amount_today * level0; is_locked[k] = 1; }
          nominal = locked[k];
        }
        income += nominal * (1.0 - p->streams[k].tax_rate);
      }
      double need = py_max(0.0, need_nominal - income);                              /* :679-682 */
      if (s[0] + s[2] <= EPS && need > EPS) { failed = 1; break; }                   /* :684-690 */

      int idx = shock_idx < n_rows - 1 ? shock_idx : n_rows - 1;                     /* :692 */
      const double* z = shocks + 3 * (size_t)idx;
      shock_idx++;
      double G1 = monthly_gross(p->inv1_mu_log, p->inv1_sigma_log, z[0]);
      double GI = monthly_gross(p->inf_mu_log, p->inf_sigma_log, z[1]);
      double GP = monthly_gross(p->prem_mu_log, p->prem_sigma_log, z[2]);
      double G2 = GI * GP;
      g1 += s[0] * (G1 - 1.0);
      g2 += s[2] * (G2 - 1.0);
      s[0] *= G1;
      s[2] *= G2;
      level *= GI;
      if (s[0] + s[2] <= EPS && need > EPS) {                                        /* :715-724 */
        s[0] = py_max(0.0, s[0]);
        s[2] = py_max(0.0, s[2]);
        failed = 1;
        break;
      }
      double cap1 = oracle_net_liquidation(s[0], s[1], p->inv1_use_realized_gains_tax_system,
                                           p->inv1_realized_gains_tax_rate);
      double cap2 = oracle_net_liquidation(s[2], s[3], p->inv2_use_realized_gains_tax_system,
                                           p->inv2_realized_gains_tax_rate);
      double cap = cap1 + cap2;
      double target = py_max(0.0, py_min(need, cap));                                /* :739-742 */
      if (need > EPS && target < need - EPS) failed = 1;                             /* :743-748 */
      double w1 = cap > EPS ? cap1 / cap : a1;                                       /* :750-755 */
      double w2 = 1.0 - w1;
      double o[4];
      oracle_withdraw(s[0], s[1], target * w1, p->inv1_use_realized_gains_tax_system,
                      p->inv1_realized_gains_tax_rate, o);
      s[0] = o[0]; s[1] = o[1];
      double gw1 = o[2], nw1 = o[3];
      yr_g1 += gw1;
      oracle_withdraw(s[2], s[3], target * w2, p->inv2_use_realized_gains_tax_system,
                      p->inv2_realized_gains_tax_rate, o);
      s[2] = o[0]; s[3] = o[1];
      double gw2 = o[2], nw2 = o[3];
      yr_g2 += gw2;
      yr_real += (gw1 + gw2) * level_ret / py_max(level0, EPS);                      /* :778-782 */
      if (need > EPS && nw1 + nw2 < need - EPS) failed = 1;                          /* :784-790 */
      oracle_rebalance(p, s);                                                        /* :792-796 */
      int abs_month = wm + r + 1;
      if (!failed && abs_month % MPY == 0) {                                         /* :798-822 */
        int tf = oracle_annual_tax(p, s, g1, g2);
        g1 = 0.0; g2 = 0.0;
        if (tf) failed = 1;
      }
      if (failed) { years_to_ruin = (double)(r + 1) / (double)MPY; break; }          /* :824-828 */
    }
    double yr_gross = yr_g1 + yr_g2;
    double wr_y = start_balance > EPS ? (yr_real / start_balance) * 100.0 : 0.0;     /* :834-840 */
    if (failed) {                                                                    /* :842-857 */
      ok = 0;
      if (isnan(years_to_ruin)) years_to_ruin = (double)(r + 1) / (double)MPY;
      traj[n_traj] = py_max(0.0, s[0] + s[2]); price[n_traj] = level; n_traj++;
      wr[n_wr++] = NAN;
      if (y == 0) { fy_gross = yr_gross; fy_real = yr_real; }
      break;
    }
    wr[n_wr++] = wr_y;
    if (y == 0) { fy_gross = yr_gross; fy_real = yr_real; }
    traj[n_traj] = s[0] + s[2]; price[n_traj] = level; n_traj++;
  }

  if (ok && total_months % MPY != 0) {                                               /* :873-898 */
    if (oracle_annual_tax(p, s, g1, g2)) { ok = 0; years_to_ruin = (double)R; }
    if (n_traj > 0) traj[n_traj - 1] = s[0] + s[2];
  }
  double final_total = s[0] + s[2];

  if (n_traj < T) {                                                                  /* :902-916 */
    double pad = !ok ? 0.0 : traj[n_traj - 1];
    double last_px = price[n_traj - 1];
    while (n_traj < T) { traj[n_traj] = pad; price[n_traj] = last_px; n_traj++; }
  }
  for (int t = 0; t < T; ++t) real_traj[t] = price[t] > EPS ? traj[t] / price[t] : 0.0; /* :928-931 */
  while (n_wr < R) wr[n_wr++] = NAN;                                                 /* :934-935 */

  rec->start_balance = start_balance;
  rec->final_balance = py_max(0.0, final_total);
  rec->success = ok;
  rec->years_to_ruin = years_to_ruin;
  rec->first_year_gross = fy_gross;
  rec->first_year_real = fy_real;
  rec->inflation_at_ret = level_ret;
  rec->trajectory_len = T;
  free(price);
  return 0;
}

/*
 * Batch runner (OpenMP over paths) — the fan-out of run_monte_carlo_simulations
 * (simulation.py:987-990). shocks: (n_paths, n_rows, 3) C-contiguous. Series outputs are
 * path-major here: traj[i*T + t], wr[i*R + y] (any may be NULL -> scratch).
 */
int oracle_run_batch(const mcr_params* p, int32_t wm, const double* shocks, int32_t n_rows,
                     int64_t n_paths, oracle_path_record* recs, double* traj, double* real_traj,
                     double* wr, int n_threads) {
  const int T = oracle_trajectory_len(wm, p->retirement_years);
  const int R = p->retirement_years;
  int err = 0;
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
  {
    double* t0 = (double*)malloc(sizeof(double) * (size_t)(2 * T + R + 4));
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n_paths; ++i) {
      double* tr = traj ? traj + i * T : t0;
      double* rt = real_traj ? real_traj + i * T : t0 + T;
      double* w = wr ? wr + i * R : t0 + 2 * T;
      if (oracle_run_path(p, wm, shocks + (size_t)i * (size_t)n_rows * 3, n_rows, recs + i, tr, rt, w))
        err = -1;
    }
    free(t0);
  }
  return err;
}
