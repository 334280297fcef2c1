// mcr_api.cu — the extern "C" boundary declared in include/mcr.h. Host-side only: parameter
// validation/derivation, context + scratch management, launches. No compute happens on the
// host and there is no CPU fallback: without a usable CUDA device every entry point fails.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "mcr_internal.h"
#include "mcr_path.cuh"
#include "mcr_reduce.h"

using namespace mcr;

struct mcr_ctx {
  int device = 0;
  int sm_count = 0;
  mcr_params params;
  DevParams dev;
  uint64_t main_seed = 0;
  uint32_t k0 = 0, k1 = 0;
  PhiloxKeys keys;
  std::mutex mu;
  std::string err;
  int64_t launches = 0;
  // live (positive-amount) streams in original order: eligibility age and duration in months
  double live_start_age[MCR_MAX_STREAMS];
  int32_t live_duration[MCR_MAX_STREAMS];
  // scratch (device), grown on demand
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
};

static thread_local std::string g_tls_err;

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int fail(mcr_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  g_tls_err = msg;
  return code;
}

int cuda_fail(mcr_ctx* ctx, cudaError_t e, const char* what) {
  return fail(ctx, MCR_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define MCR_CUDA(ctx, call)                                   \
  do {                                                        \
    cudaError_t e__ = (call);                                 \
    if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
  } while (0)

uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

int ensure_scratch(mcr_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->scratch_bytes) return MCR_OK;
  if (ctx->scratch) cudaFree(ctx->scratch);
  ctx->scratch = nullptr;
  ctx->scratch_bytes = 0;
  size_t want = bytes < (1u << 16) ? (1u << 16) : bytes;
  if (cudaMalloc(&ctx->scratch, want) != cudaSuccess) {
    cudaGetLastError();
    return fail(ctx, MCR_ENOMEM, "scratch allocation of " + std::to_string(want) + " bytes failed");
  }
  ctx->scratch_bytes = want;
  return MCR_OK;
}

// stream_payment_start_month_index — simulation.py:47-63, same double operations as CPython
int32_t start_month_index(double current_age, int32_t wm, double start_at_age) {
  const double ret = current_age + (double)wm / (double)MCR_MONTHS_PER_YEAR;
  const double elig = (start_at_age > ret) ? start_at_age : ret;
  const double k = std::ceil((elig - ret) * (double)MCR_MONTHS_PER_YEAR - MCR_SMALL_EPSILON);
  if (!(k > 0)) return 0;
  return k > 2147483647.0 ? 2147483647 : (int32_t)k;
}

int validate_and_derive(const mcr_params& p, DevParams& d, double* live_start_age, int32_t* live_duration,
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
    // |mu/12| + sigma/sqrt(12) * z_max for the three factors; Box-Muller on 32-bit uniforms
    // gives |n| <= sqrt(-2 ln 2^-33) = 6.77, the inflation shock is rho*n0 + rho_c*n1
    const double zmax = 6.8, zinf = zmax * (std::fabs(d.rho) + d.rho_c);
    const double b1 = std::fabs(d.mu1) + d.sg1 * zmax, bi = std::fabs(d.muI) + d.sgI * zinf,
                 bp = std::fabs(d.muP) + d.sgP * zmax;
    d.exp_small = (b1 < 0.1 && bi < 0.1 && bp < 0.1) ? 1 : 0;
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

const Launchers& pick(uint32_t flags) { return (flags & MCR_FLAG_STRICT) ? strict_launchers() : fast_launchers(); }

int check_months(mcr_ctx* ctx, int32_t wm) {
  if (wm < 0) return fail(ctx, MCR_EINVAL, "working_months must be >= 0");
  if ((int64_t)wm + (int64_t)ctx->dev.R * 12 > (1 << 24)) return fail(ctx, MCR_EINVAL, "timeline too long");
  return MCR_OK;
}

// per live stream: [first paying retirement month, end) — simulation.py:602-621,653-656
void fill_windows(const mcr_ctx* ctx, int32_t wm, int32_t* out) {
  for (int k = 0; k < MCR_MAX_STREAMS; ++k) {
    int32_t first = 0, end = 0;
    if (k < ctx->dev.n_streams) {
      first = start_month_index(ctx->params.current_age, wm, ctx->live_start_age[k]);
      const int64_t e = ctx->live_duration[k] < 0 ? 2147483647ll : (int64_t)first + ctx->live_duration[k];
      end = e > 2147483647ll ? 2147483647 : (int32_t)e;
    }
    out[2 * k] = first;
    out[2 * k + 1] = end;
  }
}

}  // namespace

extern "C" {

int mcr_abi_version(void) { return MCR_ABI_VERSION; }

const char* mcr_last_error(const mcr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_tls_err.c_str(); }

int32_t mcr_stream_start_month(double current_age, int32_t working_months, double start_at_age) {
  return start_month_index(current_age, working_months, start_at_age);
}

int32_t mcr_trajectory_len(int32_t working_months, int32_t retirement_years) {
  const int32_t wy = working_months > 0 ? (working_months + MCR_MONTHS_PER_YEAR - 1) / MCR_MONTHS_PER_YEAR : 0;
  return 1 + wy + retirement_years;
}

int mcr_create(const mcr_params* params, uint64_t main_seed, int device, mcr_ctx** out_ctx) {
  if (!params || !out_ctx) return fail(nullptr, MCR_EINVAL, "null argument");
  *out_ctx = nullptr;
  DevParams d;
  std::string why;
  double live_age[MCR_MAX_STREAMS] = {0};
  int32_t live_dur[MCR_MAX_STREAMS] = {0};
  if (validate_and_derive(*params, d, live_age, live_dur, why) != MCR_OK) return fail(nullptr, MCR_EINVAL, why);
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return fail(nullptr, MCR_ECUDA, std::string("no CUDA device available (this engine has no CPU fallback): ") +
                                        (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  }
  if (device < 0 || device >= n_dev) return fail(nullptr, MCR_EINVAL, "device index out of range");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
  if (prop.major < 10)
    return fail(nullptr, MCR_ECUDA, "device is not sm_100-class (the library carries sm_100a code only)");
  mcr_ctx* ctx = new mcr_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->params = *params;
  ctx->dev = d;
  std::memcpy(ctx->live_start_age, live_age, sizeof(live_age));
  std::memcpy(ctx->live_duration, live_dur, sizeof(live_dur));
  ctx->main_seed = main_seed;
  const uint64_t key = splitmix64(splitmix64(main_seed) ^ 0x6D63725F62323030ull /* "mcr_b200" */);
  ctx->k0 = (uint32_t)key;
  ctx->k1 = (uint32_t)(key >> 32);
  philox_expand_keys(ctx->k0, ctx->k1, ctx->keys);
  *out_ctx = ctx;
  return MCR_OK;
}

int mcr_destroy(mcr_ctx* ctx) {
  if (!ctx) return MCR_OK;
  {
    DeviceGuard g(ctx->device);
    if (ctx->scratch) cudaFree(ctx->scratch);
  }
  delete ctx;
  return MCR_OK;
}

int64_t mcr_launch_count(const mcr_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mcr_simulate(mcr_ctx* ctx, int seed_stream, int32_t working_months, int64_t first_path, int64_t n_paths,
                 uint32_t flags, const mcr_outputs* out, void* stream) {
  if (!ctx || !out) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (int rc = check_months(ctx, working_months)) return rc;
  if (n_paths < 0 || first_path < 0) return fail(ctx, MCR_EINVAL, "negative path range");
  if (seed_stream != MCR_STREAM_SEARCH && seed_stream != MCR_STREAM_FINAL) return fail(ctx, MCR_EINVAL, "bad seed stream");
  if ((out->trajectory || out->real_trajectory || out->wr_trajectory) && out->series_ld < n_paths)
    return fail(ctx, MCR_EINVAL, "series_ld < n_paths");
  if (n_paths == 0) return MCR_OK;
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(ctx, MCR_ECUDA, "cudaSetDevice failed");
  TimelineArgs A;
  std::memset(&A, 0, sizeof(A));
  A.wm = working_months;
  A.keys = ctx->keys; A.seed_stream = (uint32_t)seed_stream;
  A.first_path = first_path; A.n_paths = n_paths;
  fill_windows(ctx, working_months, A.window);
  A.out = *out;
  MCR_CUDA(ctx, pick(flags).timeline(ctx->dev, A, false, (cudaStream_t)stream));
  ctx->launches += 1;
  return MCR_OK;
}

int mcr_replay(mcr_ctx* ctx, const double* shocks_dev, int64_t shocks_ld, int32_t n_months, int32_t working_months,
               int64_t n_paths, uint32_t flags, const mcr_outputs* out, void* stream) {
  if (!ctx || !out || !shocks_dev) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (int rc = check_months(ctx, working_months)) return rc;
  const int64_t need = (int64_t)working_months + (int64_t)ctx->dev.R * 12;
  if (n_months < (need > 1 ? need : 1)) return fail(ctx, MCR_EINVAL, "shock matrix has too few months");
  if (n_paths < 0 || shocks_ld < n_paths) return fail(ctx, MCR_EINVAL, "bad path count / shocks_ld");
  if ((out->trajectory || out->real_trajectory || out->wr_trajectory) && out->series_ld < n_paths)
    return fail(ctx, MCR_EINVAL, "series_ld < n_paths");
  if (n_paths == 0) return MCR_OK;
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(ctx, MCR_ECUDA, "cudaSetDevice failed");
  TimelineArgs A;
  std::memset(&A, 0, sizeof(A));
  A.wm = working_months;
  A.n_paths = n_paths;
  fill_windows(ctx, working_months, A.window);
  A.shocks = shocks_dev;
  A.shocks_ld = shocks_ld;
  A.n_months = n_months;
  A.out = *out;
  MCR_CUDA(ctx, pick(flags).timeline(ctx->dev, A, true, (cudaStream_t)stream));
  ctx->launches += 1;
  return MCR_OK;
}

int mcr_single_path(mcr_ctx* ctx, int32_t working_months, const double* shocks_host, int32_t n_months,
                    mcr_path_record* rec, double* traj_host, double* real_host, double* wr_host) {
  if (!ctx || !shocks_host || !rec) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (int rc = check_months(ctx, working_months)) return rc;
  const int R = ctx->dev.R;
  const int64_t need = (int64_t)working_months + (int64_t)R * 12;
  if (n_months < (need > 1 ? need : 1)) return fail(ctx, MCR_EINVAL, "shock matrix has too few months");
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(ctx, MCR_ECUDA, "cudaSetDevice failed");
  const int T = mcr_trajectory_len(working_months, R);
  // scratch layout: shocks[n*3] | traj[T] | real[T] | wr[R] | 5 doubles | ruin i32 | success u8
  const size_t n_sh = (size_t)n_months * 3;
  const size_t doubles = n_sh + 2 * (size_t)T + (size_t)R + 5;
  if (int rc = ensure_scratch(ctx, doubles * 8 + 16)) return rc;
  double* base = (double*)ctx->scratch;
  double* d_sh = base;
  double* d_traj = d_sh + n_sh;
  double* d_real = d_traj + T;
  double* d_wr = d_real + T;
  double* d_sc = d_wr + R;
  int32_t* d_ruin = (int32_t*)(d_sc + 5);
  uint8_t* d_succ = (uint8_t*)(d_ruin + 1);
  cudaStream_t st = 0;
  MCR_CUDA(ctx, cudaMemcpyAsync(d_sh, shocks_host, n_sh * 8, cudaMemcpyHostToDevice, st));
  TimelineArgs A;
  std::memset(&A, 0, sizeof(A));
  A.wm = working_months;
  A.n_paths = 1;
  fill_windows(ctx, working_months, A.window);
  A.shocks = d_sh;
  A.shocks_ld = 1;  // [(m*3 + c) * 1 + 0] == row-major (n_months, 3)
  A.n_months = n_months;
  A.out.start_balance = d_sc + 0;
  A.out.final_balance = d_sc + 1;
  A.out.first_year_gross = d_sc + 2;
  A.out.first_year_real = d_sc + 3;
  A.out.inflation_at_ret = d_sc + 4;
  A.out.ruin_month = d_ruin;
  A.out.success = d_succ;
  A.out.trajectory = d_traj;
  A.out.real_trajectory = d_real;
  A.out.wr_trajectory = d_wr;
  A.out.series_ld = 1;
  MCR_CUDA(ctx, strict_launchers().timeline(ctx->dev, A, true, st));
  ctx->launches += 1;
  std::vector<double> h(2 * (size_t)T + R + 5 + 2);
  MCR_CUDA(ctx, cudaMemcpyAsync(h.data(), d_traj, (2 * (size_t)T + R + 5) * 8 + 8, cudaMemcpyDeviceToHost, st));
  MCR_CUDA(ctx, cudaStreamSynchronize(st));
  if (traj_host) std::memcpy(traj_host, h.data(), (size_t)T * 8);
  if (real_host) std::memcpy(real_host, h.data() + T, (size_t)T * 8);
  if (wr_host) std::memcpy(wr_host, h.data() + 2 * T, (size_t)R * 8);
  const double* sc = h.data() + 2 * T + R;
  rec->start_balance = sc[0];
  rec->final_balance = sc[1];
  rec->first_year_gross = sc[2];
  rec->first_year_real = sc[3];
  rec->inflation_at_ret = sc[4];
  int32_t ruin;
  uint8_t succ;
  std::memcpy(&ruin, sc + 5, 4);
  std::memcpy(&succ, (const char*)(sc + 5) + 4, 1);
  rec->ruin_month = ruin;
  rec->success = succ;
  rec->trajectory_len = T;
  rec->wr_len = R;
  return MCR_OK;
}

static int run_helper(mcr_ctx* ctx, int which, double a, double b, double c, double d, int use_tax, double rate,
                      double* out_host, int n_out, double e = 0.0) {
  if (!ctx || !out_host) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(ctx, MCR_ECUDA, "cudaSetDevice failed");
  if (int rc = ensure_scratch(ctx, 64)) return rc;
  MCR_CUDA(ctx, strict_launchers().helper(ctx->dev, which, a, b, c, d, use_tax, rate, e, (double*)ctx->scratch, 0));
  ctx->launches += 1;
  MCR_CUDA(ctx, cudaMemcpy(out_host, ctx->scratch, sizeof(double) * (size_t)n_out, cudaMemcpyDeviceToHost));
  return MCR_OK;
}

int mcr_helper_withdraw(mcr_ctx* ctx, double bal, double cost_basis, double net_target, int32_t use_real_tax,
                        double real_tax_rate, double out4_host[4]) {
  return run_helper(ctx, 0, bal, cost_basis, net_target, 0.0, use_real_tax, real_tax_rate, out4_host, 4);
}

int mcr_helper_net_liquidation(mcr_ctx* ctx, double bal, double cost_basis, int32_t use_real_tax, double real_tax_rate,
                               double* out_host) {
  return run_helper(ctx, 1, bal, cost_basis, 0.0, 0.0, use_real_tax, real_tax_rate, out_host, 1);
}

int mcr_helper_rebalance(mcr_ctx* ctx, double bal1, double cb1, double bal2, double cb2, double out4_host[4]) {
  return run_helper(ctx, 2, bal1, cb1, bal2, cb2, 0, 0.0, out4_host, 4);
}

int mcr_helper_annual_tax(mcr_ctx* ctx, double bal1, double cb1, double bal2, double cb2, double gain1, double gain2,
                          double out5_host[5]) {
  return run_helper(ctx, 3, bal1, cb1, bal2, cb2, 0, gain1, out5_host, 5, gain2);
}

int mcr_draw_shocks(mcr_ctx* ctx, int seed_stream, int64_t first_path, int64_t n_paths, int32_t n_months,
                    uint32_t flags, double* shocks_dev, int64_t shocks_ld, void* stream) {
  if (!ctx || !shocks_dev) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (n_paths < 0 || first_path < 0 || n_months < 0 || shocks_ld < n_paths) return fail(ctx, MCR_EINVAL, "bad shape");
  if (seed_stream != MCR_STREAM_SEARCH && seed_stream != MCR_STREAM_FINAL) return fail(ctx, MCR_EINVAL, "bad seed stream");
  if (n_paths == 0 || n_months == 0) return MCR_OK;
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(ctx, MCR_ECUDA, "cudaSetDevice failed");
  MCR_CUDA(ctx, pick(flags).draw(ctx->dev, ctx->keys, (uint32_t)seed_stream, first_path, n_paths, n_months,
                                 shocks_dev, shocks_ld, (cudaStream_t)stream));
  ctx->launches += 1;
  return MCR_OK;
}

int mcr_search_batch(mcr_ctx* ctx, int seed_stream, const int32_t* candidates_host, int32_t n_candidates,
                     int64_t first_path, int64_t n_paths, uint32_t flags, int64_t* success_counts_dev,
                     uint64_t* executed_months_dev, void* stream) {
  if (!ctx || !candidates_host || !success_counts_dev) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (n_candidates < 0 || n_candidates > 65535) return fail(ctx, MCR_EINVAL, "n_candidates must be in [0, 65535]");
  if (n_paths < 0 || first_path < 0) return fail(ctx, MCR_EINVAL, "negative path range");
  if (seed_stream != MCR_STREAM_SEARCH && seed_stream != MCR_STREAM_FINAL) return fail(ctx, MCR_EINVAL, "bad seed stream");
  for (int c = 0; c < n_candidates; ++c)
    if (int rc = check_months(ctx, candidates_host[c])) return rc;
  if (n_candidates == 0 || n_paths == 0) return MCR_OK;
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(ctx, MCR_ECUDA, "cudaSetDevice failed");
  // longest candidate first: the hardware block scheduler then back-fills with short ones
  std::vector<int32_t> order(n_candidates);
  for (int c = 0; c < n_candidates; ++c) order[c] = c;
  std::stable_sort(order.begin(), order.end(),
                   [&](int a, int b) { return candidates_host[a] > candidates_host[b]; });
  const size_t per = (size_t)n_candidates;
  std::vector<int32_t> h(per * (2 + 2 * MCR_MAX_STREAMS));
  int32_t* h_wm = h.data();
  int32_t* h_slot = h_wm + per;
  int32_t* h_sm = h_slot + per;
  for (size_t k = 0; k < per; ++k) {
    const int c = order[k];
    h_wm[k] = candidates_host[c];
    h_slot[k] = c;
    fill_windows(ctx, candidates_host[c], h_sm + k * 2 * MCR_MAX_STREAMS);
  }
  if (int rc = ensure_scratch(ctx, h.size() * 4)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // pageable source: the copy is staged before the call returns, so `h` may go out of scope
  MCR_CUDA(ctx, cudaMemcpyAsync(ctx->scratch, h.data(), h.size() * 4, cudaMemcpyHostToDevice, st));
  SearchArgs A;
  std::memset(&A, 0, sizeof(A));
  A.keys = ctx->keys; A.seed_stream = (uint32_t)seed_stream;
  A.n_candidates = n_candidates;
  A.first_path = first_path; A.n_paths = n_paths;
  A.wm = (const int32_t*)ctx->scratch;
  A.slot = A.wm + per;
  A.window = A.slot + per;
  A.success_counts = success_counts_dev;
  A.executed_months = executed_months_dev;
  MCR_CUDA(ctx, pick(flags).search(ctx->dev, A, st));
  ctx->launches += 1;
  return MCR_OK;
}

static int make_spec(mcr_ctx* ctx, const double* q_host, int32_t n_q, uint32_t sel_flags, QuantileSpec& spec) {
  if (n_q <= 0 || n_q > kMaxQuantiles) return fail(ctx, MCR_EINVAL, "n_q must be in [1,16]");
  std::memset(&spec, 0, sizeof(spec));
  spec.n_q = n_q;
  spec.median = (sel_flags & MCR_SEL_MEDIAN) ? 1 : 0;
  for (int k = 0; k < n_q; ++k) {
    spec.q[k] = spec.median ? 0.5 : (q_host ? q_host[k] : -1.0);
    if (!(spec.q[k] >= 0.0 && spec.q[k] <= 1.0)) return fail(ctx, MCR_EINVAL, "quantiles must be in [0,1]");
    if (k > 0 && spec.q[k] < spec.q[k - 1]) return fail(ctx, MCR_EINVAL, "quantiles must be ascending");
  }
  return MCR_OK;
}

static int make_descs(mcr_ctx* ctx, const mcr_select_row* rows_host, int32_t n_rows, std::vector<RowDesc>& d) {
  if (!rows_host || n_rows <= 0 || n_rows > 65535) return fail(ctx, MCR_EINVAL, "n_rows must be in [1, 65535]");
  d.resize((size_t)n_rows);
  for (int r = 0; r < n_rows; ++r) {
    const mcr_select_row& in = rows_host[r];
    if (in.n < 0 || (!in.values_dev && in.n > 0)) return fail(ctx, MCR_EINVAL, "bad select row");
    d[r].x = in.values_dev;
    d[r].mask = in.mask_dev;
    d[r].n = in.n;
    if (int rc = make_spec(ctx, in.q, in.n_q, in.flags, d[r].spec)) return rc;
  }
  return MCR_OK;
}

int mcr_quantiles_rows(mcr_ctx* ctx, const mcr_select_row* rows_host, int32_t n_rows, double* out_dev,
                       int64_t* counts_dev, void* stream) {
  if (!ctx || !out_dev) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  std::vector<RowDesc> d;
  if (int rc = make_descs(ctx, rows_host, n_rows, d)) return rc;
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(ctx, MCR_ECUDA, "cudaSetDevice failed");
  const size_t hb = quantile_hist_bytes(n_rows);
  if (int rc = ensure_scratch(ctx, hb + quantile_state_bytes(n_rows))) return rc;
  int n_launches = 0;
  MCR_CUDA(ctx, launch_quantiles_rows(n_rows, d.data(), out_dev, MCR_MAX_QUANTILES, counts_dev,
                                      (char*)ctx->scratch + hb, ctx->scratch, (cudaStream_t)stream, &n_launches));
  ctx->launches += n_launches;
  return MCR_OK;
}

int mcr_quantiles(mcr_ctx* ctx, const double* values_dev, int64_t n, int64_t ld, int32_t rows, const uint8_t* mask_dev,
                  const double* q_host, int32_t n_q, uint32_t sel_flags, double* out_dev, int64_t* counts_dev,
                  void* stream) {
  if (!ctx || (!values_dev && n > 0) || !out_dev) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (n < 0 || rows < 0 || rows > 65535 || (rows > 1 && ld < n)) return fail(ctx, MCR_EINVAL, "bad shape");
  QuantileSpec spec;
  if (int rc = make_spec(ctx, q_host, n_q, sel_flags, spec)) return rc;
  if (rows == 0) return MCR_OK;
  std::vector<RowDesc> d((size_t)rows);
  for (int r = 0; r < rows; ++r) {
    d[r].x = values_dev + (int64_t)r * ld;
    d[r].mask = mask_dev;
    d[r].n = n;
    d[r].spec = spec;
  }
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(ctx, MCR_ECUDA, "cudaSetDevice failed");
  const size_t hb = quantile_hist_bytes(rows);
  if (int rc = ensure_scratch(ctx, hb + quantile_state_bytes(rows))) return rc;
  int n_launches = 0;
  MCR_CUDA(ctx, launch_quantiles_rows(rows, d.data(), out_dev, n_q, counts_dev, (char*)ctx->scratch + hb, ctx->scratch,
                                      (cudaStream_t)stream, &n_launches));
  ctx->launches += n_launches;
  return MCR_OK;
}

int64_t mcr_select_state_bytes(int32_t rows) { return (int64_t)quantile_state_bytes(rows); }
int64_t mcr_select_hist_bytes(int32_t rows) { return (int64_t)quantile_hist_bytes(rows); }
int32_t mcr_select_full_passes(void) { return select_full_passes(); }
int64_t mcr_select_exchange_words(int32_t rows, int32_t world) { return (int64_t)select_exchange_words(rows, world); }
void mcr_select_exchange_layout(int32_t rows, int32_t world, int64_t* at4) {
  if (at4) select_exchange_layout(rows > 0 ? rows : 1, world > 0 ? world : 1, at4);
}

int mcr_select_step(mcr_ctx* ctx, int32_t step, int32_t pass, const mcr_select_row* rows_host, int32_t n_rows,
                    void* state_dev, void* hist_dev, double* out_dev, int64_t* counts_dev, void* stream) {
  if (!ctx || !state_dev || !hist_dev) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  const bool pooled_step = step >= MCR_SELECT_POOL_EXPORT && step <= MCR_SELECT_POOL_TAIL;
  const int rank = pass & 0xff, world = (pass >> 8) & 0xff;  // POOL_* steps: pass = rank | world << 8
  if (pooled_step ? (world < 1 || rank >= world) : (pass < 0 || pass > 15)) return fail(ctx, MCR_EINVAL, "bad pass");
  std::vector<RowDesc> d;
  if (int rc = make_descs(ctx, rows_host, n_rows, d)) return rc;
  int64_t max_n = 0;
  int max_nq = 1;
  for (const RowDesc& r : d) {
    max_n = r.n > max_n ? r.n : max_n;
    max_nq = r.spec.n_q > max_nq ? r.spec.n_q : max_nq;
  }
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(ctx, MCR_ECUDA, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  switch (step) {
    case MCR_SELECT_BEGIN:  // pass != 0: the caller will exchange the row extremes after HIST 0
      MCR_CUDA(ctx, launch_sel_begin(n_rows, d.data(), state_dev, hist_dev, st, /*adaptive=*/pass & 1,
                                     /*fused=*/(pass >> 1) & 1));
      break;
    case MCR_SELECT_EXTREMES_GET:
    case MCR_SELECT_EXTREMES_SET:
      if (!out_dev) return fail(ctx, MCR_EINVAL, "null extremes buffer");
      MCR_CUDA(ctx, launch_sel_extremes(n_rows, state_dev, (long long*)out_dev, step == MCR_SELECT_EXTREMES_SET, st));
      break;
    case MCR_SELECT_HIST:
      MCR_CUDA(ctx, launch_sel_hist(n_rows, max_n, pass, state_dev, hist_dev, st));
      break;
    case MCR_SELECT_COLLECT:
      MCR_CUDA(ctx, launch_sel_collect(n_rows, max_n, state_dev, st));
      ctx->launches += 1;
      break;
    case MCR_SELECT_ADVANCE:
      MCR_CUDA(ctx, launch_sel_advance(n_rows, max_nq, pass, state_dev, hist_dev, st));
      break;
    case MCR_SELECT_POOL_EXPORT:  // hist_dev is the exchange buffer for the POOL_* steps
      MCR_CUDA(ctx, launch_sel_export(n_rows, state_dev, (long long*)hist_dev, rank, world, st));
      break;
    case MCR_SELECT_POOL_PLACE:
      MCR_CUDA(ctx, launch_sel_place(n_rows, state_dev, (long long*)hist_dev, rank, world, st));
      break;
    case MCR_SELECT_POOL_TAIL:
      if (!out_dev) return fail(ctx, MCR_EINVAL, "null output");
      MCR_CUDA(ctx, launch_sel_tail_pooled(n_rows, state_dev, (long long*)hist_dev, world, out_dev, MCR_MAX_QUANTILES,
                                           counts_dev, st));
      break;
    case MCR_SELECT_FINISH:
      if (!out_dev) return fail(ctx, MCR_EINVAL, "null output");
      MCR_CUDA(ctx, launch_sel_finish(n_rows, state_dev, out_dev, MCR_MAX_QUANTILES, counts_dev, st));
      break;
    default:
      return fail(ctx, MCR_EINVAL, "bad select step");
  }
  ctx->launches += 1;
  return MCR_OK;
}

int mcr_first_year_rates(mcr_ctx* ctx, const double* start_dev, const double* first_year_real_dev, int64_t n,
                         double* rates_dev, void* stream) {
  if (!ctx || !start_dev || !first_year_real_dev || !rates_dev) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (n < 0) return fail(ctx, MCR_EINVAL, "negative n");
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(ctx, MCR_ECUDA, "cudaSetDevice failed");
  MCR_CUDA(ctx, launch_rates(start_dev, first_year_real_dev, n, rates_dev, (cudaStream_t)stream));
  ctx->launches += n > 0;
  return MCR_OK;
}

int mcr_minmax(mcr_ctx* ctx, const double* values_dev, const uint8_t* mask_dev, int64_t n, double divisor,
               double* minmax_dev, void* stream) {
  if (!ctx || !values_dev || !minmax_dev) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (n < 0 || !(divisor > 0)) return fail(ctx, MCR_EINVAL, "bad n / divisor");
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(ctx, MCR_ECUDA, "cudaSetDevice failed");
  if (int rc = ensure_scratch(ctx, 64)) return rc;
  MCR_CUDA(ctx, launch_minmax(values_dev, mask_dev, n, divisor, (unsigned long long*)ctx->scratch, minmax_dev,
                              (cudaStream_t)stream));
  ctx->launches += 2 + (n > 0);
  return MCR_OK;
}

int mcr_histogram(mcr_ctx* ctx, const double* values_dev, const uint8_t* mask_dev, int64_t n, double divisor,
                  int32_t n_bins, int32_t mode, const double* range_dev, int64_t* hist_dev, void* stream) {
  if (!ctx || !values_dev || !range_dev || !hist_dev) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (n < 0 || !(divisor > 0) || n_bins <= 0 || n_bins > 8192) return fail(ctx, MCR_EINVAL, "bad n / divisor / n_bins");
  if (mode != MCR_HIST_NUMPY && mode != MCR_HIST_FLOOR) return fail(ctx, MCR_EINVAL, "bad histogram mode");
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(ctx, MCR_ECUDA, "cudaSetDevice failed");
  MCR_CUDA(ctx, launch_histogram(values_dev, mask_dev, n, divisor, n_bins, mode, range_dev, hist_dev,
                                 (cudaStream_t)stream));
  ctx->launches += 1;
  return MCR_OK;
}

int mcr_gather_columns(mcr_ctx* ctx, const double* series_dev, int64_t ld, int32_t rows, const int64_t* cols_host,
                       int32_t n_cols, double* out_dev, void* stream) {
  if (!ctx || !series_dev || !cols_host || !out_dev) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (n_cols < 0 || n_cols > 1024 || rows < 0) return fail(ctx, MCR_EINVAL, "bad shape");
  for (int k = 0; k < n_cols; ++k)
    if (cols_host[k] < 0 || cols_host[k] >= ld) return fail(ctx, MCR_EINVAL, "column index out of range");
  if (n_cols == 0 || rows == 0) return MCR_OK;
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(ctx, MCR_ECUDA, "cudaSetDevice failed");
  if (int rc = ensure_scratch(ctx, (size_t)n_cols * 8)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  MCR_CUDA(ctx, cudaMemcpyAsync(ctx->scratch, cols_host, (size_t)n_cols * 8, cudaMemcpyHostToDevice, st));
  MCR_CUDA(ctx, launch_gather(series_dev, ld, rows, (const int64_t*)ctx->scratch, n_cols, out_dev, st));
  ctx->launches += 1;
  return MCR_OK;
}

int mcr_fp64_peak_slots_per_s(mcr_ctx* ctx, double* slots_per_s_host) {
  if (!ctx || !slots_per_s_host) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(ctx, MCR_ECUDA, "cudaSetDevice failed");
  if (int rc = ensure_scratch(ctx, 64)) return rc;
  cudaEvent_t e0, e1;
  MCR_CUDA(ctx, cudaEventCreate(&e0));
  MCR_CUDA(ctx, cudaEventCreate(&e1));
  int threads = 0, per_thread = 0;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {  // rep 0 is the warm-up
    MCR_CUDA(ctx, cudaEventRecord(e0, 0));
    MCR_CUDA(ctx, launch_fp64_peak(ctx->sm_count, 4096, (double*)ctx->scratch, 0, &threads, &per_thread));
    MCR_CUDA(ctx, cudaEventRecord(e1, 0));
    MCR_CUDA(ctx, cudaEventSynchronize(e1));
    ctx->launches += 1;
    float ms = 0.f;
    MCR_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    const double rate = (double)threads * (double)per_thread / ((double)ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *slots_per_s_host = best;
  return MCR_OK;
}

}  // extern "C"
