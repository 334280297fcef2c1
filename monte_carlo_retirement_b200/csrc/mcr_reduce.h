// mcr_reduce.h — launch interface of the device aggregations (mcr_reduce.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/mcr.h"

namespace mcr {

constexpr int kMaxQuantiles = 16;

struct QuantileSpec {
  int32_t n_q;
  int32_t median;  // 1: np.median rule (mean of the two middle order statistics), q ignored; 2: {min, max}
  double q[kMaxQuantiles];
};

// one row of a select call: where the values are, the cohort mask, what is asked
struct RowDesc {
  const double* x;
  const uint8_t* mask;
  int64_t n;
  QuantileSpec spec;
};

// Where a select call expects its rows' descriptors: the caller copies RowDesc[rows] there (on the call's
// stream) before launch_quantiles_rows / launch_sel_begin.
RowDesc* select_desc_area(void* state, int rows);
// all rows in ONE launch sequence (rows may differ in source, length, mask and quantiles); desc_host only
// sizes the grids here
cudaError_t launch_quantiles_rows(int rows, const RowDesc* desc_host, double* out, int out_stride, int64_t* counts,
                                  void* state, void* hist, cudaStream_t st, int* n_launches);
// the same select, one pass at a time (a multi-GPU caller all-reduces `hist` between hist and advance)
cudaError_t launch_sel_begin(int rows, void* state, void* hist, cudaStream_t st, int adaptive = 0, int fused = 0);
// fuse: the last CTA of a row runs the advance / closes the list itself (single-GPU sequence: no separate
// ADVANCE launch, no all-reduce in between)
cudaError_t launch_sel_hist(int rows, int64_t max_n, int pass, void* state, void* hist, cudaStream_t st,
                            int sampled = 0, int fuse = 0);
cudaError_t launch_sel_collect(int rows, int64_t max_n, void* state, cudaStream_t st, int fuse = 0);
int select_full_passes();
int select_full_passes_for(int64_t n_global_max);
// pooled tail for path shards on several GPUs (see "candidate exchange" in mcr_reduce.cu)
size_t select_exchange_words(int rows, int world);
void select_exchange_layout(int rows, int world, int64_t at[4]);
cudaError_t launch_sel_export(int rows, void* state, long long* xbuf, int rank, int world, cudaStream_t st);
cudaError_t launch_sel_place(int rows, void* state, long long* xbuf, int rank, int world, cudaStream_t st);
cudaError_t launch_sel_tail_pooled(int rows, void* state, long long* xbuf, int world, double* out, int out_stride,
                                   int64_t* counts, cudaStream_t st);
cudaError_t launch_sel_extremes(int rows, void* state, long long* ext, int store, cudaStream_t st);
cudaError_t launch_sel_advance(int rows, int max_nq, int pass, void* state, void* hist, cudaStream_t st);
cudaError_t launch_sel_finish(int rows, const void* state, double* out, int out_stride, int64_t* counts,
                              cudaStream_t st);
size_t quantile_state_bytes(int rows);
size_t quantile_hist_bytes(int rows);
cudaError_t launch_rates(const double* start, const double* fy_real, int64_t n, double* rates, cudaStream_t st);
cudaError_t launch_years_to_ruin(const int32_t* ruin, int64_t n, double* years, cudaStream_t st);
cudaError_t launch_minmax(const double* x, const uint8_t* mask, int64_t n, double divisor, unsigned long long* keys2,
                          double* minmax, cudaStream_t st);
cudaError_t launch_histogram(const double* x, const uint8_t* mask, int64_t n, double divisor, int n_bins, int mode,
                             const double* range_dev, int64_t* hist, cudaStream_t st);
cudaError_t launch_gather(const double* series, int64_t ld, int rows, const int64_t* cols_dev, int n_cols, double* out,
                          cudaStream_t st);
cudaError_t launch_fp64_peak(int sm_count, int iters, double* sink, cudaStream_t st, int* total_threads,
                             int* dfma_per_thread);

}  // namespace mcr
