// mcr_internal.h — host/device shared argument blocks and the launcher tables that the two
// arithmetic builds (strict / fast) export to the C-ABI layer (mcr_api.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/mcr.h"
#include "mcr_rng.cuh"

namespace mcr {

struct DevParams;

struct TimelineArgs {
  int32_t wm;
  uint32_t seed_stream;           // counter word 3
  PhiloxKeys keys;                // expanded Philox key schedule
  int64_t first_path, n_paths;
  int32_t window[2 * MCR_MAX_STREAMS];   // per live stream: first paying retirement month, end (exclusive)
  const double* shocks;           // replay only
  int64_t shocks_ld;
  int32_t n_months;               // replay only: rows of the shock matrix
  mcr_outputs out;
};

struct SearchArgs {
  PhiloxKeys keys;
  uint32_t seed_stream;
  int32_t n_candidates;
  int64_t first_path, n_paths;
  const int32_t* wm;              // [n_candidates] device, sorted longest first
  const int32_t* slot;            // [n_candidates] device: index into the caller's arrays
  const int32_t* window;          // [n_candidates][2 * MCR_MAX_STREAMS] device
  int64_t* success_counts;        // caller buffer, accumulated
  uint64_t* executed_months;      // caller buffer or NULL, accumulated
};

// Multi-scenario sweep: like SearchArgs, but every item brings its own scenario.
struct SweepArgs {
  PhiloxKeys keys;
  uint32_t seed_stream;
  int32_t n_items;
  int64_t first_path, n_paths;
  const DevParams* scenarios;     // [n_scenarios] device
  const int32_t* scen;            // [n_items] device: scenario of the item (items of one launch share a variant)
  const int32_t* wm;              // [n_items] device, sorted longest first
  const int32_t* slot;            // [n_items] device: index into the caller's arrays
  const int32_t* window;          // [n_items][2 * MCR_MAX_STREAMS] device
  int64_t* success_counts;        // caller buffer, accumulated
  uint64_t* executed_months;      // caller buffer or NULL, accumulated
};

struct Launchers {
  cudaError_t (*timeline)(const DevParams&, const TimelineArgs&, bool replay, int cfg, cudaStream_t);  // cfg: pick_cfg_index()
  cudaError_t (*search)(const DevParams&, const SearchArgs&, int cfg, cudaStream_t);
  cudaError_t (*draw)(const DevParams&, const PhiloxKeys& keys, uint32_t seed_stream, int64_t first_path,
                      int64_t n_paths, int32_t n_months, double* shocks, int64_t ld, cudaStream_t);
  cudaError_t (*helper)(const DevParams&, int which, double a, double b, double c, double d, int use_tax,
                        double rate, double e, double* out, cudaStream_t);
  cudaError_t (*sweep)(const SweepArgs&, int cfg, cudaStream_t);
};

const Launchers& strict_launchers();
const Launchers& fast_launchers();

}  // namespace mcr
