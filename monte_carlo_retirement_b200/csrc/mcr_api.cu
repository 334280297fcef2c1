// mcr_api.cu — the extern "C" boundary declared in include/mcr.h. Host-side only: parameter
// validation/derivation, context + scratch management, launches. No compute happens on the
// host and there is no CPU fallback: without a usable CUDA device every entry point fails.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "mcr_derive.h"
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
  int last_variant = -1;   // pick_cfg_index() of the last timeline / search launch
  // live (positive-amount) streams in original order: eligibility age and duration in months
  double live_start_age[MCR_MAX_STREAMS];
  int32_t live_duration[MCR_MAX_STREAMS];
  // Scratch (device), grown on demand. One buffer PER STREAM: calls that run on different streams
  // (aggregates_device(pipeline=True) puts the selects on a side stream while the next search /
  // single-path call is enqueued on the main one) never share live scratch. A buffer is only
  // replaced after its own stream has drained.
  struct Scratch {
    void* p = nullptr;
    size_t bytes = 0;
  };
  std::map<cudaStream_t, Scratch> scratch;
  // Page-locked staging for the small host arrays a call hands to the device (row descriptors of a
  // select, candidate lists of a search, scenario blocks of a sweep, column lists). cudaMemcpyAsync
  // from PAGEABLE memory waits for the stream's earlier work before it stages the source — the
  // timeline kernel of the same step — so the host could never run more than one kernel ahead of the
  // device and every hiccup of the calling thread showed up as idle GPU time. A ring of slots, each
  // guarded by an event recorded behind its copy.
  static constexpr int kStageSlots = 16;
  struct Staging {
    char* p = nullptr;
    size_t slot_bytes = 0;
    int next = 0;
    cudaEvent_t done[kStageSlots] = {};
    bool in_flight[kStageSlots] = {};
  } stage;
  // private non-blocking stream of the synchronous host-output entry points (single path,
  // helpers, peak): they must not serialise against other contexts through the legacy stream
  cudaStream_t own_stream = nullptr;
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

// scratch of at least `bytes` for work enqueued on `st` (see mcr_ctx::scratch)
int ensure_scratch(mcr_ctx* ctx, cudaStream_t st, size_t bytes, void** out) {
  mcr_ctx::Scratch& s = ctx->scratch[st];
  if (bytes > s.bytes) {
    if (s.p) {
      // work already enqueued on this stream may still read the old buffer
      cudaError_t e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaStreamSynchronize(scratch growth)");
      cudaFree(s.p);
    }
    s.p = nullptr;
    s.bytes = 0;
    const size_t want = bytes < (1u << 16) ? (1u << 16) : bytes;
    if (cudaMalloc(&s.p, want) != cudaSuccess) {
      cudaGetLastError();
      return fail(ctx, MCR_ENOMEM, "scratch allocation of " + std::to_string(want) + " bytes failed");
    }
    s.bytes = want;
  }
  *out = s.p;
  return MCR_OK;
}

// dst_dev[0, bytes) = src_host[0, bytes) behind the work already on `st`, without waiting for it; src_host
// may be reused as soon as this returns (see mcr_ctx::stage)
int stage_h2d(mcr_ctx* ctx, cudaStream_t st, void* dst_dev, const void* src_host, size_t bytes) {
  if (bytes == 0) return MCR_OK;
  mcr_ctx::Staging& S = ctx->stage;
  if (bytes > S.slot_bytes) {
    for (int k = 0; k < mcr_ctx::kStageSlots; ++k)
      if (S.in_flight[k]) {
        cudaEventSynchronize(S.done[k]);
        S.in_flight[k] = false;
      }
    if (S.p) cudaFreeHost(S.p);
    S.p = nullptr;
    S.slot_bytes = 0;
    size_t want = 1u << 16;
    while (want < bytes) want <<= 1;
    if (cudaHostAlloc(&S.p, want * mcr_ctx::kStageSlots, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      S.p = nullptr;
      return fail(ctx, MCR_ENOMEM, "page-locked staging allocation of " + std::to_string(want * mcr_ctx::kStageSlots) + " bytes failed");
    }
    S.slot_bytes = want;
    for (int k = 0; k < mcr_ctx::kStageSlots; ++k)
      if (!S.done[k] && cudaEventCreateWithFlags(&S.done[k], cudaEventDisableTiming) != cudaSuccess)
        return cuda_fail(ctx, cudaGetLastError(), "cudaEventCreate(staging)");
  }
  const int k = S.next;
  S.next = (S.next + 1) % mcr_ctx::kStageSlots;
  if (S.in_flight[k]) {
    cudaError_t e = cudaEventSynchronize(S.done[k]);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaEventSynchronize(staging)");
  }
  char* slot = S.p + (size_t)k * S.slot_bytes;
  std::memcpy(slot, src_host, bytes);
  cudaError_t e = cudaMemcpyAsync(dst_dev, slot, bytes, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMemcpyAsync(staging)");
  e = cudaEventRecord(S.done[k], st);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaEventRecord(staging)");
  S.in_flight[k] = true;
  return MCR_OK;
}

int own_stream(mcr_ctx* ctx, cudaStream_t* st) {
  if (!ctx->own_stream) {
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaStreamCreateWithFlags");
  }
  *st = ctx->own_stream;
  return MCR_OK;
}

const Launchers& pick(uint32_t flags) { return (flags & MCR_FLAG_STRICT) ? strict_launchers() : fast_launchers(); }

int check_months(mcr_ctx* ctx, int32_t wm) {
  if (wm < 0) return fail(ctx, MCR_EINVAL, "working_months must be >= 0");
  if ((int64_t)wm + (int64_t)ctx->dev.R * 12 > (1 << 24)) return fail(ctx, MCR_EINVAL, "timeline too long");
  return MCR_OK;
}

void fill_windows(const mcr_ctx* ctx, int32_t wm, int32_t* out) {
  stream_windows(ctx->dev, ctx->params.current_age, ctx->live_start_age, ctx->live_duration, wm, out);
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
  philox_key_from_seed(main_seed, ctx->k0, ctx->k1);
  philox_expand_keys(ctx->k0, ctx->k1, ctx->keys);
  *out_ctx = ctx;
  return MCR_OK;
}

int mcr_destroy(mcr_ctx* ctx) {
  if (!ctx) return MCR_OK;
  {
    DeviceGuard g(ctx->device);
    for (auto& kv : ctx->scratch)
      if (kv.second.p) {
        cudaStreamSynchronize(kv.first);
        cudaFree(kv.second.p);
      }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    for (int k = 0; k < mcr_ctx::kStageSlots; ++k)
      if (ctx->stage.done[k]) {
        if (ctx->stage.in_flight[k]) cudaEventSynchronize(ctx->stage.done[k]);
        cudaEventDestroy(ctx->stage.done[k]);
      }
    if (ctx->stage.p) cudaFreeHost(ctx->stage.p);
  }
  delete ctx;
  return MCR_OK;
}

double mcr_small_returns_bound(const mcr_ctx* ctx) {
  if (!ctx) return 0.0;
  return ctx->dev.exp_small == 2 ? 0.05 : (ctx->dev.exp_small == 1 ? 0.1 : 0.0);
}

int mcr_comm_device_of(const mcr_ctx* ctx) { return ctx ? ctx->device : 0; }  // mcr_comm.cu

int32_t mcr_last_variant(const mcr_ctx* ctx) { return ctx ? ctx->last_variant : -1; }

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
  // fast build: the variant for the bound on the monthly log-returns proven at mcr_create
  const bool fast = !(flags & MCR_FLAG_STRICT);
  const int cfg = pick_cfg_index(ctx->dev, fast, fast ? ctx->dev.exp_small : 0);
  ctx->last_variant = cfg;
  MCR_CUDA(ctx, pick(flags).timeline(ctx->dev, A, false, cfg, (cudaStream_t)stream));
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
  // supplied draws: bounded only if the caller vouches for it (MCR_FLAG_SMALL_RETURNS)
  const bool fast = !(flags & MCR_FLAG_STRICT);
  const int cfg = pick_cfg_index(ctx->dev, fast, (fast && (flags & MCR_FLAG_SMALL_RETURNS)) ? ctx->dev.exp_small : 0);
  ctx->last_variant = cfg;
  MCR_CUDA(ctx, pick(flags).timeline(ctx->dev, A, true, cfg, (cudaStream_t)stream));
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
  cudaStream_t st;
  if (int rc = own_stream(ctx, &st)) return rc;
  void* scratch;
  if (int rc = ensure_scratch(ctx, st, doubles * 8 + 16, &scratch)) return rc;
  double* base = (double*)scratch;
  double* d_sh = base;
  double* d_traj = d_sh + n_sh;
  double* d_real = d_traj + T;
  double* d_wr = d_real + T;
  double* d_sc = d_wr + R;
  int32_t* d_ruin = (int32_t*)(d_sc + 5);
  uint8_t* d_succ = (uint8_t*)(d_ruin + 1);
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
  MCR_CUDA(ctx, strict_launchers().timeline(ctx->dev, A, true, pick_cfg_index(ctx->dev, false, 0), st));
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
  cudaStream_t st;
  if (int rc = own_stream(ctx, &st)) return rc;
  void* scratch;
  if (int rc = ensure_scratch(ctx, st, 64, &scratch)) return rc;
  MCR_CUDA(ctx, strict_launchers().helper(ctx->dev, which, a, b, c, d, use_tax, rate, e, (double*)scratch, st));
  ctx->launches += 1;
  MCR_CUDA(ctx, cudaMemcpyAsync(out_host, scratch, sizeof(double) * (size_t)n_out, cudaMemcpyDeviceToHost, st));
  MCR_CUDA(ctx, cudaStreamSynchronize(st));
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
  cudaStream_t st = (cudaStream_t)stream;
  void* scratch;
  if (int rc = ensure_scratch(ctx, st, h.size() * 4, &scratch)) return rc;
  if (int rc = stage_h2d(ctx, st, scratch, h.data(), h.size() * 4)) return rc;
  SearchArgs A;
  std::memset(&A, 0, sizeof(A));
  A.keys = ctx->keys; A.seed_stream = (uint32_t)seed_stream;
  A.n_candidates = n_candidates;
  A.first_path = first_path; A.n_paths = n_paths;
  A.wm = (const int32_t*)scratch;
  A.slot = A.wm + per;
  A.window = A.slot + per;
  A.success_counts = success_counts_dev;
  A.executed_months = executed_months_dev;
  const bool fast = !(flags & MCR_FLAG_STRICT);
  const int cfg = pick_cfg_index(ctx->dev, fast, fast ? ctx->dev.exp_small : 0);
  ctx->last_variant = cfg;
  MCR_CUDA(ctx, pick(flags).search(ctx->dev, A, cfg, st));
  ctx->launches += 1;
  return MCR_OK;
}

int mcr_sweep_batch(mcr_ctx* ctx, int seed_stream, const mcr_params* scenarios_host, const int32_t* working_months_host,
                    int32_t n_items, int64_t first_path, int64_t n_paths, uint32_t flags, int64_t* success_counts_dev,
                    uint64_t* executed_months_dev, void* stream) {
  if (!ctx || !scenarios_host || !working_months_host || !success_counts_dev) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (n_items < 0 || n_items > 65535) return fail(ctx, MCR_EINVAL, "n_items must be in [0, 65535]");
  if (n_paths < 0 || first_path < 0) return fail(ctx, MCR_EINVAL, "negative path range");
  if (seed_stream != MCR_STREAM_SEARCH && seed_stream != MCR_STREAM_FINAL) return fail(ctx, MCR_EINVAL, "bad seed stream");
  if (n_items == 0 || n_paths == 0) return MCR_OK;
  const bool fast = !(flags & MCR_FLAG_STRICT);
  // derive every scenario like mcr_create does, and its variant
  std::vector<DevParams> dev((size_t)n_items);
  std::vector<int32_t> variant((size_t)n_items);
  std::vector<int32_t> windows((size_t)n_items * 2 * MCR_MAX_STREAMS);
  for (int k = 0; k < n_items; ++k) {
    std::string why;
    double live_age[MCR_MAX_STREAMS] = {0};
    int32_t live_dur[MCR_MAX_STREAMS] = {0};
    if (validate_and_derive(scenarios_host[k], dev[k], live_age, live_dur, why) != MCR_OK)
      return fail(ctx, MCR_EINVAL, "scenario " + std::to_string(k) + ": " + why);
    const int32_t wm = working_months_host[k];
    if (wm < 0 || (int64_t)wm + (int64_t)dev[k].R * 12 > (1 << 24)) return fail(ctx, MCR_EINVAL, "bad working_months");
    stream_windows(dev[k], scenarios_host[k].current_age, live_age, live_dur, wm, windows.data() + (size_t)k * 2 * MCR_MAX_STREAMS);
    variant[k] = pick_cfg_index(dev[k], fast, fast ? dev[k].exp_small : 0);
  }
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(ctx, MCR_ECUDA, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  // scratch: DevParams[n] | per group: scen | wm | slot | window
  const size_t per_item = 3 + 2 * MCR_MAX_STREAMS;
  const size_t dev_bytes = (sizeof(DevParams) * (size_t)n_items + 255) / 256 * 256;
  void* scratch;
  if (int rc = ensure_scratch(ctx, st, dev_bytes + (size_t)n_items * per_item * 4, &scratch)) return rc;
  std::vector<int32_t> h((size_t)n_items * per_item);
  // one launch per variant present; inside a launch the longest timelines first
  std::vector<int32_t> order((size_t)n_items);
  for (int k = 0; k < n_items; ++k) order[k] = k;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    if (variant[a] != variant[b]) return variant[a] < variant[b];
    return working_months_host[a] + 12 * dev[a].R > working_months_host[b] + 12 * dev[b].R;
  });
  if (int rc = stage_h2d(ctx, st, scratch, dev.data(), sizeof(DevParams) * (size_t)n_items)) return rc;
  int32_t* base_dev = (int32_t*)((char*)scratch + dev_bytes);
  size_t at = 0;
  struct Group { int cfg; size_t at, count; };
  std::vector<Group> groups;
  for (size_t lo = 0; lo < order.size();) {
    size_t hi = lo;
    while (hi < order.size() && variant[order[hi]] == variant[order[lo]]) ++hi;
    const size_t cnt = hi - lo;
    int32_t* scen = h.data() + at;
    int32_t* wm = scen + cnt;
    int32_t* slot = wm + cnt;
    int32_t* win = slot + cnt;
    for (size_t j = 0; j < cnt; ++j) {
      const int k = order[lo + j];
      scen[j] = k;
      wm[j] = working_months_host[k];
      slot[j] = k;
      std::memcpy(win + j * 2 * MCR_MAX_STREAMS, windows.data() + (size_t)k * 2 * MCR_MAX_STREAMS, sizeof(int32_t) * 2 * MCR_MAX_STREAMS);
    }
    groups.push_back({variant[order[lo]], at, cnt});
    at += cnt * per_item;
    lo = hi;
  }
  if (int rc = stage_h2d(ctx, st, base_dev, h.data(), h.size() * 4)) return rc;
  for (const Group& gr : groups) {
    SweepArgs A;
    std::memset(&A, 0, sizeof(A));
    A.keys = ctx->keys;
    A.seed_stream = (uint32_t)seed_stream;
    A.n_items = (int32_t)gr.count;
    A.first_path = first_path;
    A.n_paths = n_paths;
    A.scenarios = (const DevParams*)scratch;
    A.scen = base_dev + gr.at;
    A.wm = A.scen + gr.count;
    A.slot = A.wm + gr.count;
    A.window = A.slot + gr.count;
    A.success_counts = success_counts_dev;
    A.executed_months = executed_months_dev;
    MCR_CUDA(ctx, pick(flags).sweep(A, gr.cfg, st));
    ctx->launches += 1;
    ctx->last_variant = gr.cfg;
  }
  return MCR_OK;
}

static int make_spec(mcr_ctx* ctx, const double* q_host, int32_t n_q, uint32_t sel_flags, QuantileSpec& spec) {
  if (n_q <= 0 || n_q > kMaxQuantiles) return fail(ctx, MCR_EINVAL, "n_q must be in [1,16]");
  std::memset(&spec, 0, sizeof(spec));
  spec.n_q = n_q;
  spec.median = (sel_flags & MCR_SEL_MEDIAN) ? 1 : ((sel_flags & MCR_SEL_MINMAX) ? 2 : 0);
  if (spec.median == 2 && n_q != 2) return fail(ctx, MCR_EINVAL, "MCR_SEL_MINMAX needs n_q == 2");
  for (int k = 0; k < n_q; ++k) {
    spec.q[k] = spec.median == 1 ? 0.5 : (spec.median == 2 ? (double)k : (q_host ? q_host[k] : -1.0));
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
  void* scratch;
  if (int rc = ensure_scratch(ctx, (cudaStream_t)stream, hb + quantile_state_bytes(n_rows), &scratch)) return rc;
  int n_launches = 0;
  if (int rc = stage_h2d(ctx, (cudaStream_t)stream, select_desc_area((char*)scratch + hb, n_rows), d.data(),
                         sizeof(RowDesc) * d.size()))
    return rc;
  MCR_CUDA(ctx, launch_quantiles_rows(n_rows, d.data(), out_dev, MCR_MAX_QUANTILES, counts_dev,
                                      (char*)scratch + hb, scratch, (cudaStream_t)stream, &n_launches));
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
  void* scratch;
  if (int rc = ensure_scratch(ctx, (cudaStream_t)stream, hb + quantile_state_bytes(rows), &scratch)) return rc;
  int n_launches = 0;
  if (int rc = stage_h2d(ctx, (cudaStream_t)stream, select_desc_area((char*)scratch + hb, rows), d.data(),
                         sizeof(RowDesc) * d.size()))
    return rc;
  MCR_CUDA(ctx, launch_quantiles_rows(rows, d.data(), out_dev, n_q, counts_dev, (char*)scratch + hb, scratch,
                                      (cudaStream_t)stream, &n_launches));
  ctx->launches += n_launches;
  return MCR_OK;
}

int64_t mcr_select_state_bytes(int32_t rows) { return (int64_t)quantile_state_bytes(rows); }
int64_t mcr_select_hist_bytes(int32_t rows) { return (int64_t)quantile_hist_bytes(rows); }
int32_t mcr_select_full_passes(void) { return select_full_passes(); }
int32_t mcr_select_full_passes_for(int64_t n_global_max) { return select_full_passes_for(n_global_max); }
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
  const int sampled = (step == MCR_SELECT_HIST) ? (pass >> 8) & 1 : 0;   // HIST: MCR_SELECT_HIST_SAMPLED
  if (step == MCR_SELECT_HIST) pass &= 0xff;
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
      if (int rc = stage_h2d(ctx, st, select_desc_area(state_dev, n_rows), d.data(), sizeof(RowDesc) * d.size())) return rc;
      MCR_CUDA(ctx, launch_sel_begin(n_rows, state_dev, hist_dev, st, /*adaptive=*/pass & 1,
                                     /*fused (3: lists are pooled across ranks)=*/((pass >> 1) & 1) ? 3 : 0));
      break;
    case MCR_SELECT_EXTREMES_GET:
    case MCR_SELECT_EXTREMES_SET:
      if (!out_dev) return fail(ctx, MCR_EINVAL, "null extremes buffer");
      MCR_CUDA(ctx, launch_sel_extremes(n_rows, state_dev, (long long*)out_dev, step == MCR_SELECT_EXTREMES_SET, st));
      break;
    case MCR_SELECT_HIST:
      MCR_CUDA(ctx, launch_sel_hist(n_rows, max_n, pass, state_dev, hist_dev, st, sampled));
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

// The pooled distributed select (mcr.h: MCR_SELECT_POOL_*) driven entirely from here, with the
// all-reduces of csrc/mcr_comm.cu between its steps: ONE call per rank instead of ~45 (the ranks of
// a single-process multi-GPU run are Python threads that share one interpreter lock, so every
// host call saved is serial time saved).
int mcr_quantiles_rows_comm(mcr_ctx* ctx, mcr_comm* comm, int32_t rank, int32_t world, const mcr_select_row* rows_host,
                            int32_t n_rows, double* out_dev, int64_t* counts_dev, int64_t* unresolved_dev, void* stream) {
  if (!ctx || !comm || !out_dev || !unresolved_dev) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (world < 1 || rank < 0 || rank >= world) return fail(ctx, MCR_EINVAL, "bad rank / world");
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
  // scratch: hist | state | exchange buffer | extremes
  const size_t hb = (quantile_hist_bytes(n_rows) + 255) / 256 * 256;
  const size_t sb = (quantile_state_bytes(n_rows) + 255) / 256 * 256;
  const size_t xw = select_exchange_words(n_rows, world);
  const size_t xb = (xw * 8 + 255) / 256 * 256;
  void* scratch;
  if (int rc = ensure_scratch(ctx, st, hb + sb + xb + (size_t)n_rows * 16, &scratch)) return rc;
  char* base = (char*)scratch;
  void* hist = base;
  void* state = base + hb;
  long long* xbuf = (long long*)(base + hb + sb);
  long long* ext = (long long*)(base + hb + sb + xb);
  int64_t at[4];
  select_exchange_layout(n_rows, world, at);
  auto reduce = [&](int op, void* buf, int64_t n) -> int {
    if (mcr_comm_all_reduce(comm, op, buf, n, stream) != MCR_OK)
      return fail(ctx, MCR_ECUDA, std::string("peer all-reduce: ") + mcr_comm_last_error(comm));
    return MCR_OK;
  };
  if (int rc = stage_h2d(ctx, st, select_desc_area(state, n_rows), d.data(), sizeof(RowDesc) * d.size())) return rc;
  MCR_CUDA(ctx, launch_sel_begin(n_rows, state, hist, st, /*adaptive=*/1, /*fused, pooled lists=*/3));
  const int full = select_full_passes_for(max_n * world);   // shards are balanced: global length ~ local x world
  for (int p = 0; p < full; ++p) {
    MCR_CUDA(ctx, launch_sel_hist(n_rows, max_n, p, state, hist, st, /*sampled=*/p == 0));
    if (p == 0) {
      MCR_CUDA(ctx, launch_sel_extremes(n_rows, state, ext, /*store=*/0, st));
      if (int rc = reduce(MCR_COMM_MIN_I64, ext, (int64_t)n_rows * 2)) return rc;
      MCR_CUDA(ctx, launch_sel_extremes(n_rows, state, ext, /*store=*/1, st));
    }
    if (int rc = reduce(MCR_COMM_SUM_I32, hist, (int64_t)(quantile_hist_bytes(n_rows) / 4))) return rc;
    MCR_CUDA(ctx, launch_sel_advance(n_rows, max_nq, p, state, hist, st));
  }
  MCR_CUDA(ctx, launch_sel_collect(n_rows, max_n, state, st));
  MCR_CUDA(ctx, launch_sel_export(n_rows, state, xbuf, rank, world, st));
  if (int rc = reduce(MCR_COMM_SUM_I64, xbuf + at[0], at[1] - at[0])) return rc;
  if (int rc = reduce(MCR_COMM_MIN_I64, xbuf + at[1], at[2] - at[1])) return rc;
  MCR_CUDA(ctx, launch_sel_place(n_rows, state, xbuf, rank, world, st));
  if (int rc = reduce(MCR_COMM_SUM_I64, xbuf + at[2], at[3] - at[2])) return rc;
  MCR_CUDA(ctx, launch_sel_tail_pooled(n_rows, state, xbuf, world, out_dev, MCR_MAX_QUANTILES, counts_dev, st));
  MCR_CUDA(ctx, cudaMemcpyAsync(unresolved_dev, xbuf, 8, cudaMemcpyDeviceToDevice, st));
  ctx->launches += 9 + 2 * full;
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

int mcr_years_to_ruin(mcr_ctx* ctx, const int32_t* ruin_month_dev, int64_t n, double* years_dev, void* stream) {
  if (!ctx || !ruin_month_dev || !years_dev) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (n < 0) return fail(ctx, MCR_EINVAL, "negative n");
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(ctx, MCR_ECUDA, "cudaSetDevice failed");
  MCR_CUDA(ctx, launch_years_to_ruin(ruin_month_dev, n, years_dev, (cudaStream_t)stream));
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
  // the key pair lives behind the caller's output (minmax_dev[2..3] would be caller memory we do
  // not own): a per-stream scratch slot, consumed by k_minmax_finish before the call's last launch
  void* scratch;
  if (int rc = ensure_scratch(ctx, (cudaStream_t)stream, 64, &scratch)) return rc;
  MCR_CUDA(ctx, launch_minmax(values_dev, mask_dev, n, divisor, (unsigned long long*)scratch, minmax_dev,
                              (cudaStream_t)stream));
  ctx->launches += 2 + (n > 0);
  return MCR_OK;
}

int mcr_histogram(mcr_ctx* ctx, const double* values_dev, const uint8_t* mask_dev, int64_t n, double divisor,
                  int32_t n_bins, int32_t mode, const double* range_dev, int64_t* hist_dev, void* stream) {
  if (!ctx || !values_dev || !range_dev || !hist_dev) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (n < 0 || !(divisor > 0) || n_bins <= 0 || n_bins > 8192) return fail(ctx, MCR_EINVAL, "bad n / divisor / n_bins");
  if ((mode & ~MCR_HIST_RAW_RANGE) != MCR_HIST_NUMPY && (mode & ~MCR_HIST_RAW_RANGE) != MCR_HIST_FLOOR)
    return fail(ctx, MCR_EINVAL, "bad histogram mode");
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
  cudaStream_t st = (cudaStream_t)stream;
  void* scratch;
  if (int rc = ensure_scratch(ctx, st, (size_t)n_cols * 8, &scratch)) return rc;
  if (int rc = stage_h2d(ctx, st, scratch, cols_host, (size_t)n_cols * 8)) return rc;
  MCR_CUDA(ctx, launch_gather(series_dev, ld, rows, (const int64_t*)scratch, n_cols, out_dev, st));
  ctx->launches += 1;
  return MCR_OK;
}

int mcr_fp64_peak_slots_per_s(mcr_ctx* ctx, double* slots_per_s_host) {
  if (!ctx || !slots_per_s_host) return fail(ctx, MCR_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  if (!g.ok) return fail(ctx, MCR_ECUDA, "cudaSetDevice failed");
  cudaStream_t st;
  if (int rc = own_stream(ctx, &st)) return rc;
  void* scratch;
  if (int rc = ensure_scratch(ctx, st, 64, &scratch)) return rc;
  cudaEvent_t e0, e1;
  MCR_CUDA(ctx, cudaEventCreate(&e0));
  MCR_CUDA(ctx, cudaEventCreate(&e1));
  int threads = 0, per_thread = 0;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {  // rep 0 is the warm-up
    MCR_CUDA(ctx, cudaEventRecord(e0, st));
    MCR_CUDA(ctx, launch_fp64_peak(ctx->sm_count, 4096, (double*)scratch, st, &threads, &per_thread));
    MCR_CUDA(ctx, cudaEventRecord(e1, st));
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
