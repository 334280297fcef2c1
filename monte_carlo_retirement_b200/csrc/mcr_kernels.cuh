// mcr_kernels.cuh — __global__ entry points of the timeline / search kernels. Included by
// mcr_kernels_strict.cu (MCR_FAST=0, -fmad=false) and mcr_kernels_fast.cu (MCR_FAST=1), which
// export the same launcher table under two names (see mcr_internal.h).
#pragma once
#include "mcr_derive.h"
#include "mcr_internal.h"
#include "mcr_path.cuh"
#include "mcr_rng.cuh"

namespace mcr {

#ifndef MCR_BLOCK
#define MCR_BLOCK 128
#endif
constexpr int kBlock = MCR_BLOCK;  // threads per CTA
#ifndef MCR_MIN_BLOCKS
#define MCR_MIN_BLOCKS 6     // resident CTAs per SM the register allocator must leave room for (80 regs)
#endif

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Per-path outputs + the block/warp reductions that replace the host-side pandas passes
// (success mean: simulation.py:1130-1136; WR observation counts: :1111-1113; ruin list:
// server.py:525-532).
__device__ __forceinline__ void emit_outputs(const mcr_outputs& out, int64_t i, bool valid, const PathOut& o,
                                             int years_observed, int R, uint32_t* s_obs) {
  if (valid) {
    if (out.start_balance) __stcs(out.start_balance + i, o.start_balance);
    if (out.final_balance) __stcs(out.final_balance + i, o.final_balance);
    if (out.success) out.success[i] = (uint8_t)o.success;
    if (out.ruin_month) __stcs(out.ruin_month + i, o.ruin_month);
    if (out.first_year_gross) __stcs(out.first_year_gross + i, o.fy_gross);
    if (out.first_year_real) __stcs(out.first_year_real + i, o.fy_real);
    if (out.inflation_at_ret) __stcs(out.inflation_at_ret + i, o.infl_ret);
    if (out.ruin_month_hist && !o.success) atomicAdd((unsigned long long*)out.ruin_month_hist + o.ruin_month, 1ull);
  }
  const uint32_t lane = threadIdx.x & 31u;
  if (out.success_count) {
    const uint32_t ok = __ballot_sync(0xffffffffu, valid && o.success);
    if (lane == 0 && ok) atomicAdd((unsigned long long*)out.success_count, (unsigned long long)__popc(ok));
  }
  if (out.executed_months) {
    const uint32_t ex = warp_sum(valid ? o.executed : 0u);
    if (lane == 0 && ex) atomicAdd((unsigned long long*)out.executed_months, (unsigned long long)ex);
  }
  if (out.wr_obs_count) {
    // s_obs[k] = paths of this block with exactly k observed years; obs[y] = #paths with k > y
    for (int k = threadIdx.x; k <= R; k += blockDim.x) s_obs[k] = 0;
    __syncthreads();
    if (valid) atomicAdd(&s_obs[years_observed], 1u);
    __syncthreads();
    for (int y = threadIdx.x; y < R; y += blockDim.x) {
      uint32_t c = 0;
      for (int k = y + 1; k <= R; ++k) c += s_obs[k];
      if (c) atomicAdd((unsigned long long*)out.wr_obs_count + y, (unsigned long long)c);
    }
  }
}

template <bool FAST, bool REPLAY, class C>
__global__ void __launch_bounds__(kBlock, MCR_MIN_BLOCKS) k_timeline(const __grid_constant__ DevParams P,
                                                     const __grid_constant__ TimelineArgs A) {
  extern __shared__ uint32_t s_obs[];
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const bool valid = i < A.n_paths;
  PathOut o = {};
  int years_observed = 0;
  if (valid) {
    SeriesSink sink;
    sink.ld = A.out.series_ld;
    sink.traj = A.out.trajectory ? A.out.trajectory + i : nullptr;
    sink.real = A.out.real_trajectory ? A.out.real_trajectory + i : nullptr;
    sink.wrp = A.out.wr_trajectory ? A.out.wr_trajectory + i : nullptr;
    if constexpr (REPLAY) {
      ReplayShock sh{A.shocks + i, A.shocks_ld, A.n_months};
      run_timeline<FAST, C>(P, A.wm, A.window, sh, sink, o, years_observed);
    } else {
      const uint64_t gp = (uint64_t)(A.first_path + i);
      PhiloxShock<FAST> sh{A.keys, (uint32_t)gp, (uint32_t)(gp >> 32), 0u, A.seed_stream,
                           P.rho_f, P.rho_c_f, P.rho, P.rho_c, 0.f, 0.f, 0.f};
      run_timeline<FAST, C>(P, A.wm, A.window, sh, sink, o, years_observed);
    }
  }
  emit_outputs(A.out, i, valid, o, years_observed, P.R, s_obs);
}

// Batched search: blockIdx.y = candidate (host orders them longest first), blockIdx.x = tile
// of 128 paths. All 32 lanes of a warp share the candidate, so they fail at similar months;
// no series, no per-path outputs: only the success count (and executed months) per candidate.
template <bool FAST, class C>
__global__ void __launch_bounds__(kBlock, MCR_MIN_BLOCKS) k_search(const __grid_constant__ DevParams P,
                                                   const __grid_constant__ SearchArgs A) {
  const int c = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const bool valid = i < A.n_paths;
  PathOut o = {};
  int years_observed = 0;
  if (valid) {
    NullSink sink;
    const uint64_t gp = (uint64_t)(A.first_path + i);
    PhiloxShock<FAST> sh{A.keys, (uint32_t)gp, (uint32_t)(gp >> 32), 0u, A.seed_stream,
                         P.rho_f, P.rho_c_f, P.rho, P.rho_c, 0.f, 0.f, 0.f};
    run_timeline<FAST, C>(P, A.wm[c], A.window + (size_t)c * 2 * MCR_MAX_STREAMS, sh, sink, o, years_observed);
  }
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t ok = __ballot_sync(0xffffffffu, valid && o.success);
  if (lane == 0 && ok) atomicAdd((unsigned long long*)A.success_counts + A.slot[c], (unsigned long long)__popc(ok));
  if (A.executed_months) {
    const uint32_t ex = warp_sum(valid ? o.executed : 0u);
    if (lane == 0 && ex) atomicAdd((unsigned long long*)A.executed_months + A.slot[c], (unsigned long long)ex);
  }
}

// Multi-scenario sweep: blockIdx.y = item = (scenario, working_months), blockIdx.x = tile of 128
// paths of the SAME Philox streams for every item (common random numbers across scenarios:
// sensitivity grids differ by their parameters, not by their luck). The scenario constants come
// from global memory (one DevParams per scenario) instead of the launch's constant bank; the
// host groups the items by compile-time variant, so the lean month steps apply here as well.
template <bool FAST, class C>
__global__ void __launch_bounds__(kBlock, MCR_MIN_BLOCKS) k_sweep(const __grid_constant__ SweepArgs A) {
  const int c = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const bool valid = i < A.n_paths;
  const DevParams& P = A.scenarios[A.scen[c]];
  PathOut o = {};
  int years_observed = 0;
  if (valid) {
    NullSink sink;
    const uint64_t gp = (uint64_t)(A.first_path + i);
    PhiloxShock<FAST> sh{A.keys, (uint32_t)gp, (uint32_t)(gp >> 32), 0u, A.seed_stream,
                         P.rho_f, P.rho_c_f, P.rho, P.rho_c, 0.f, 0.f, 0.f};
    run_timeline<FAST, C>(P, A.wm[c], A.window + (size_t)c * 2 * MCR_MAX_STREAMS, sh, sink, o, years_observed);
  }
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t ok = __ballot_sync(0xffffffffu, valid && o.success);
  if (lane == 0 && ok) atomicAdd((unsigned long long*)A.success_counts + A.slot[c], (unsigned long long)__popc(ok));
  if (A.executed_months) {
    const uint32_t ex = warp_sum(valid ? o.executed : 0u);
    if (lane == 0 && ex) atomicAdd((unsigned long long*)A.executed_months + A.slot[c], (unsigned long long)ex);
  }
}

// Native shocks written out in the replay layout — the device analogue of `_draw_shock_path`
// (simulation.py:452-466); lets the tests replay the Philox draws through the CPU oracle.
template <bool FAST>
__global__ void k_draw_shocks(const __grid_constant__ DevParams P, const __grid_constant__ PhiloxKeys keys,
                              uint32_t seed_stream, int64_t first_path, int64_t n_paths, int32_t n_months,
                              double* shocks, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_paths) return;
  const uint64_t gp = (uint64_t)(first_path + i);
  PhiloxShock<FAST> sh{keys, (uint32_t)gp, (uint32_t)(gp >> 32), 0u, seed_stream, P.rho_f, P.rho_c_f, P.rho, P.rho_c, 0.f, 0.f, 0.f};
  for (int m = 0; m < n_months; ++m) {
    double ze, zi, zp;
    sh.next(ze, zi, zp);
    shocks[(int64_t)(3 * m + 0) * ld + i] = ze;
    shocks[(int64_t)(3 * m + 1) * ld + i] = zi;
    shocks[(int64_t)(3 * m + 2) * ld + i] = zp;
  }
}

// One-thread helper evaluations (strict build only is exported)
template <bool FAST>
__global__ void k_helper(const __grid_constant__ DevParams P, int which, double a, double b, double c, double d,
                         int use_tax, double rate, double e, double* out) {
  if (threadIdx.x || blockIdx.x) return;
  if (which == 0) {  // withdraw(bal=a, cb=b, target=c)
    double gross, net;
    const bool taxed = use_tax && rate > 0;
    withdraw<FAST, true>(a, b, c, taxed, rate, gross, net);
    out[0] = a; out[1] = b; out[2] = gross; out[3] = net;
  } else if (which == 1) {  // net liquidation
    out[0] = net_liq(a, b, use_tax && rate > 0, rate);
  } else if (which == 2) {  // rebalance(b1=a, cb1=b, b2=c, cb2=d)
    rebalance<FAST, true>(P, a, b, c, d);
    out[0] = a; out[1] = b; out[2] = c; out[3] = d;
  } else {  // annual tax(b1=a, cb1=b, b2=c, cb2=d, gain1=rate, gain2=e)
    const bool failed = annual_tax<FAST, true>(P, a, b, c, d, rate, e);
    out[0] = a; out[1] = b; out[2] = c; out[3] = d; out[4] = failed ? 1.0 : 0.0;
  }
}

// ---- launchers ---------------------------------------------------------------------------
template <class C>
static void launch_timeline_cfg(const DevParams& P, const TimelineArgs& A, bool replay, unsigned grid, size_t smem,
                                cudaStream_t st) {
  if (replay)
    k_timeline<MCR_FAST != 0, true, C><<<grid, kBlock, smem, st>>>(P, A);
  else
    k_timeline<MCR_FAST != 0, false, C><<<grid, kBlock, smem, st>>>(P, A);
}

// cfg: pick_cfg_index(P, fast, level of the bound on the monthly log-returns) — chosen by the API layer
static cudaError_t launch_timeline(const DevParams& P, const TimelineArgs& A, bool replay, int cfg, cudaStream_t st) {
  const unsigned grid = (unsigned)((A.n_paths + kBlock - 1) / kBlock);
  const size_t smem = A.out.wr_obs_count ? sizeof(uint32_t) * (size_t)(P.R + 1) : 0;
  switch (cfg) {
    case 1: launch_timeline_cfg<CfgBothTaxed>(P, A, replay, grid, smem, st); break;
    case 2: launch_timeline_cfg<CfgNoTax>(P, A, replay, grid, smem, st); break;
#if MCR_FAST
    case 3: launch_timeline_cfg<CfgBothTaxedSmall>(P, A, replay, grid, smem, st); break;
    case 4: launch_timeline_cfg<CfgNoTaxSmall>(P, A, replay, grid, smem, st); break;
    case 5: launch_timeline_cfg<CfgBothTaxedTight>(P, A, replay, grid, smem, st); break;
    case 6: launch_timeline_cfg<CfgNoTaxTight>(P, A, replay, grid, smem, st); break;
#endif
    default: launch_timeline_cfg<CfgGeneric>(P, A, replay, grid, smem, st); break;
  }
  return cudaGetLastError();
}

static cudaError_t launch_search(const DevParams& P, const SearchArgs& A, int cfg, cudaStream_t st) {
  dim3 grid((unsigned)((A.n_paths + kBlock - 1) / kBlock), (unsigned)A.n_candidates);
  switch (cfg) {
    case 1: k_search<MCR_FAST != 0, CfgBothTaxed><<<grid, kBlock, 0, st>>>(P, A); break;
    case 2: k_search<MCR_FAST != 0, CfgNoTax><<<grid, kBlock, 0, st>>>(P, A); break;
#if MCR_FAST
    case 3: k_search<true, CfgBothTaxedSmall><<<grid, kBlock, 0, st>>>(P, A); break;
    case 4: k_search<true, CfgNoTaxSmall><<<grid, kBlock, 0, st>>>(P, A); break;
    case 5: k_search<true, CfgBothTaxedTight><<<grid, kBlock, 0, st>>>(P, A); break;
    case 6: k_search<true, CfgNoTaxTight><<<grid, kBlock, 0, st>>>(P, A); break;
#endif
    default: k_search<MCR_FAST != 0, CfgGeneric><<<grid, kBlock, 0, st>>>(P, A); break;
  }
  return cudaGetLastError();
}

static cudaError_t launch_sweep(const SweepArgs& A, int cfg, cudaStream_t st) {
  dim3 grid((unsigned)((A.n_paths + kBlock - 1) / kBlock), (unsigned)A.n_items);
  switch (cfg) {
    case 1: k_sweep<MCR_FAST != 0, CfgBothTaxed><<<grid, kBlock, 0, st>>>(A); break;
    case 2: k_sweep<MCR_FAST != 0, CfgNoTax><<<grid, kBlock, 0, st>>>(A); break;
#if MCR_FAST
    case 3: k_sweep<true, CfgBothTaxedSmall><<<grid, kBlock, 0, st>>>(A); break;
    case 4: k_sweep<true, CfgNoTaxSmall><<<grid, kBlock, 0, st>>>(A); break;
    case 5: k_sweep<true, CfgBothTaxedTight><<<grid, kBlock, 0, st>>>(A); break;
    case 6: k_sweep<true, CfgNoTaxTight><<<grid, kBlock, 0, st>>>(A); break;
#endif
    default: k_sweep<MCR_FAST != 0, CfgGeneric><<<grid, kBlock, 0, st>>>(A); break;
  }
  return cudaGetLastError();
}

static cudaError_t launch_draw(const DevParams& P, const PhiloxKeys& keys, uint32_t seed_stream, int64_t first_path,
                               int64_t n_paths, int32_t n_months, double* shocks, int64_t ld, cudaStream_t st) {
  const unsigned grid = (unsigned)((n_paths + 127) / 128);
  k_draw_shocks<MCR_FAST != 0><<<grid, 128, 0, st>>>(P, keys, seed_stream, first_path, n_paths, n_months, shocks, ld);
  return cudaGetLastError();
}

static cudaError_t launch_helper(const DevParams& P, int which, double a, double b, double c, double d, int use_tax,
                                 double rate, double e, double* out, cudaStream_t st) {
  k_helper<MCR_FAST != 0><<<1, 32, 0, st>>>(P, which, a, b, c, d, use_tax, rate, e, out);
  return cudaGetLastError();
}

}  // namespace mcr
