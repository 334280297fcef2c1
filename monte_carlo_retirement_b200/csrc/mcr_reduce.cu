// mcr_reduce.cu — device aggregations that replace the host-side pandas/numpy passes of
// run_monte_carlo_simulations and its consumers (/root/reference/backend/simulation.py:
// 1045-1118, :78-96; backend/server.py:439-461,525-532; backend/plotting.py:46-59;
// frontend/src/components/HistogramChart.jsx:13-60).
//
//   k_sel_*       exact order statistics by MSD radix select (8-bit digits on the
//                 order-preserving 64-bit key): per pass one grid-wide histogram kernel
//                 (rows x chunks CTAs, every requested quantile of a row resolved in the same
//                 scan) + one tiny advance kernel; followed by numpy's 'linear' lerp
//                 (numpy/lib/_function_base_impl.py: _QuantileMethods['linear'], _get_indexes,
//                 _lerp) or the even/odd median rule of np.median.
//   k_rates       first-year withdrawal rates (simulation.py:92-95)
//   k_minmax / k_histogram   cohort min/max and equal-width histograms with numpy.histogram
//                 or frontend floor binning
//   k_gather      sample-path columns (simulation.py:1068-1078)
// Compiled with -fmad=false: the interpolation arithmetic must round like numpy's.
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "mcr_reduce.h"

namespace mcr {

namespace {

constexpr int kMaxTargets = 2 * kMaxQuantiles;

__device__ __forceinline__ uint64_t key_of(double v) {
  const uint64_t b = (uint64_t)__double_as_longlong(v);
  return b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(uint64_t k) {
  const uint64_t b = k ^ ((k >> 63) ? 0x8000000000000000ull : ~0ull);
  return __longlong_as_double((long long)b);
}

// Per-row selection state, resident in the context's scratch between the passes of one call.
// The radix select is split in per-pass kernels (histogram over all CTAs of a row -> global
// merge -> one small "advance" step) so that (a) a row is scanned by many SMs at once and (b) a
// multi-GPU caller can all-reduce the merged histograms between the two halves of a pass and
// obtain exact GLOBAL order statistics without moving any data.
struct SelRow {
  uint32_t hist[kMaxTargets][256];  // merged histogram of the current pass, per prefix group
  uint64_t prefix[kMaxTargets];     // per target (rank-sorted): key bits resolved so far
  int64_t rank[kMaxTargets];        // per target: rank inside the current prefix bucket
  uint64_t uprefix[kMaxTargets];    // sorted unique prefixes (groups) of the current pass
  int32_t group[kMaxTargets];       // target -> group
  int32_t slot[kMaxTargets];        // (quantile, lo/hi) -> rank-sorted target slot
  int32_t n_groups;
  int32_t pad_;
  int64_t n_valid;
};

constexpr int kHistThreads = 256;
constexpr int kChunk = 16384;  // elements of one row handled by one CTA

__global__ void k_sel_init(SelRow* __restrict__ rows) {
  SelRow& R = rows[blockIdx.x];
  for (int k = threadIdx.x; k < kMaxTargets * 256; k += blockDim.x) (&R.hist[0][0])[k] = 0;
  if (threadIdx.x == 0) { R.n_groups = 1; R.n_valid = 0; }
}

// One pass: histogram of digit `pass` (MSB first) inside every live prefix bucket.
__global__ void __launch_bounds__(kHistThreads) k_sel_hist(const double* __restrict__ values, int64_t n, int64_t ld,
                                                           const uint8_t* __restrict__ mask, SelRow* __restrict__ rows,
                                                           int pass) {
  extern __shared__ uint32_t sh[];  // [n_groups][256] + uprefix copy
  SelRow& R = rows[blockIdx.y];
  const int ng = R.n_groups;
  uint64_t* s_up = (uint64_t*)(sh + ng * 256);
  for (int k = threadIdx.x; k < ng * 256; k += kHistThreads) sh[k] = 0;
  if (pass > 0 && threadIdx.x < ng) s_up[threadIdx.x] = R.uprefix[threadIdx.x];
  __syncthreads();
  const double* __restrict__ x = values + (int64_t)blockIdx.y * ld;
  const int shift = 56 - 8 * pass;
  const int64_t begin = (int64_t)blockIdx.x * kChunk;
  const int64_t end = begin + kChunk < n ? begin + kChunk : n;
  const uint64_t up_lo = pass > 0 ? s_up[0] : 0, up_hi = pass > 0 ? s_up[ng - 1] : 0;
  for (int64_t base = begin; base < end; base += kHistThreads * 4) {
    double v[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {  // 4 independent loads in flight per thread
      const int64_t e = base + u * kHistThreads + threadIdx.x;
      ok[u] = e < end && (!mask || mask[e]);
      v[u] = ok[u] ? __ldcs(x + e) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int id = -1;
      if (ok[u] && v[u] == v[u]) {  // NaN never takes part (pandas skips it)
        const uint64_t k = key_of(v[u]);
        int g = 0;
        bool hit = true;
        if (pass > 0) {
          const uint64_t hi = k >> (shift + 8);
          hit = false;
          if (hi >= up_lo && hi <= up_hi) {
            int lo_i = 0, hi_i = ng - 1;
            while (lo_i <= hi_i) {
              const int mid = (lo_i + hi_i) >> 1;
              const uint64_t uu = s_up[mid];
              if (uu == hi) { g = mid; hit = true; break; }
              if (uu < hi) lo_i = mid + 1; else hi_i = mid - 1;
            }
          }
        }
        if (hit) id = g * 256 + (int)((k >> shift) & 255u);
      }
      // concentrated data puts whole warps in one bin: one atomic for the warp in that case
      const int id0 = __shfl_sync(0xffffffffu, id, 0);
      if (__all_sync(0xffffffffu, id == id0)) {
        if (id0 >= 0 && (threadIdx.x & 31) == 0) atomicAdd(&sh[id0], 32u);
      } else if (id >= 0) {
        atomicAdd(&sh[id], 1u);
      }
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < ng * 256; k += kHistThreads)
    if (sh[k]) atomicAdd(&(&R.hist[0][0])[k], sh[k]);
}

// Second half of a pass (one warp per row): consume the merged histogram.
__global__ void k_sel_advance(SelRow* __restrict__ rows, QuantileSpec spec, int pass) {
  SelRow& R = rows[blockIdx.x];
  const int tid = threadIdx.x;
  const int nq = spec.n_q, nt = 2 * nq;
  if (pass == 0) {
    if (tid == 0) {
      int64_t nv = 0;
      for (int d = 0; d < 256; ++d) nv += R.hist[0][d];
      R.n_valid = nv;
      for (int q = 0; q < nq; ++q) {
        int64_t lo = 0, hi = 0;
        if (nv > 0) {
          if (spec.median) {  // np.median: mean of the two middle order statistics
            lo = (nv - 1) / 2;
            hi = nv / 2;
          } else {  // numpy 'linear': virtual index (n - 1) * q ; previous = floor(vi)
            const double vi = __dmul_rn((double)(nv - 1), spec.q[q]);
            if (vi >= (double)(nv - 1)) { lo = hi = nv - 1; }
            else if (vi < 0) { lo = hi = 0; }
            else { lo = (int64_t)floor(vi); hi = lo + 1; }
          }
        }
        R.rank[2 * q] = lo;
        R.rank[2 * q + 1] = hi;
      }
      // rank-sort the targets (tiny insertion sort) so that prefixes stay sorted in every pass;
      // for small n the (lo, hi) pairs of different quantiles interleave.
      int32_t ord[kMaxTargets];
      for (int t = 0; t < nt; ++t) ord[t] = t;
      for (int i = 1; i < nt; ++i) {
        const int32_t o = ord[i];
        const int64_t r = R.rank[o];
        int j = i - 1;
        while (j >= 0 && R.rank[ord[j]] > r) { ord[j + 1] = ord[j]; --j; }
        ord[j + 1] = o;
      }
      int64_t sorted[kMaxTargets];
      for (int t = 0; t < nt; ++t) sorted[t] = R.rank[ord[t]];
      for (int t = 0; t < nt; ++t) {
        R.rank[t] = sorted[t];
        R.slot[ord[t]] = t;
        R.prefix[t] = 0;
        R.group[t] = 0;
      }
    }
    __syncwarp();
  }
  // each target walks its bucket histogram to the digit holding its rank
  if (tid < nt && R.n_valid > 0) {
    const uint32_t* h = R.hist[R.group[tid]];
    int64_t r = R.rank[tid];
    int d = 0;
    for (; d < 255; ++d) {
      const int64_t c = h[d];
      if (r < c) break;
      r -= c;
    }
    R.rank[tid] = r;
    R.prefix[tid] = (R.prefix[tid] << 8) | (uint64_t)d;
  }
  __syncwarp();
  // groups of the next pass (targets are rank-sorted, so prefixes are sorted) + clear histograms
  if (tid == 0) {
    int g = 0;
    for (int t = 0; t < nt; ++t) {
      if (t == 0 || R.prefix[t] != R.uprefix[g - 1]) R.uprefix[g++] = R.prefix[t];
      R.group[t] = g - 1;
    }
    R.n_groups = g < 1 ? 1 : g;
  }
  __syncwarp();
  for (int k = tid; k < kMaxTargets * 256; k += blockDim.x) (&R.hist[0][0])[k] = 0;
}

__global__ void k_sel_finish(const SelRow* __restrict__ rows, QuantileSpec spec, double* __restrict__ out,
                             int64_t* __restrict__ counts) {
  const SelRow& R = rows[blockIdx.x];
  const int tid = threadIdx.x;
  const int nq = spec.n_q;
  if (tid < nq) {
    const int64_t nv = R.n_valid;
    double res = CUDART_NAN;
    if (nv > 0) {
      const double a = value_of(R.prefix[R.slot[2 * tid]]);
      const double b = value_of(R.prefix[R.slot[2 * tid + 1]]);
      if (spec.median) {
        res = (nv & 1) ? a : __ddiv_rn(__dadd_rn(a, b), 2.0);
      } else {
        const double q = spec.q[tid];
        const double vi = __dmul_rn((double)(nv - 1), q);
        double prev = floor(vi);
        if (vi >= (double)(nv - 1)) prev = -1.0;  // numpy _get_indexes: both indexes -> last
        else if (vi < 0) prev = 0.0;
        const double t = __dsub_rn(vi, prev);     // gamma
        const double diff = __dsub_rn(b, a);      // _lerp
        res = __dadd_rn(a, __dmul_rn(diff, t));
        if (t >= 0.5) res = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, t)));
      }
    }
    out[(int64_t)blockIdx.x * nq + tid] = res;
  }
  if (tid == 0 && counts) counts[blockIdx.x] = R.n_valid;
}

__global__ void k_rates(const double* __restrict__ start, const double* __restrict__ fy_real, int64_t n,
                        double* __restrict__ rates) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double s = start[i];
  rates[i] = s > MCR_SMALL_EPSILON ? __dmul_rn(__ddiv_rn(fy_real[i], s), 100.0) : CUDART_NAN;
}

// ---- cohort min / max -----------------------------------------------------------------------
__device__ __forceinline__ void atomic_min_key(unsigned long long* p, unsigned long long k) { atomicMin(p, k); }
__device__ __forceinline__ void atomic_max_key(unsigned long long* p, unsigned long long k) { atomicMax(p, k); }

// keys[0] = min key (init ~0), keys[1] = max key (init 0)
__global__ void k_minmax(const double* __restrict__ x, const uint8_t* __restrict__ mask, int64_t n, double divisor,
                         unsigned long long* __restrict__ keys) {
  unsigned long long lo = ~0ull, hi = 0ull;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (mask && !mask[i]) continue;
    const double v = __ddiv_rn(x[i], divisor);
    if (v != v) continue;
    const unsigned long long k = key_of(v);
    lo = k < lo ? k : lo;
    hi = k > hi ? k : hi;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo, o);
    const unsigned long long h2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = l2 < lo ? l2 : lo;
    hi = h2 > hi ? h2 : hi;
  }
  if ((threadIdx.x & 31) == 0) {
    if (lo != ~0ull) atomic_min_key(keys, lo);
    if (hi != 0ull) atomic_max_key(keys + 1, hi);
  }
}

__global__ void k_minmax_init(unsigned long long* keys) {
  if (threadIdx.x == 0) { keys[0] = ~0ull; keys[1] = 0ull; }
}

__global__ void k_minmax_finish(const unsigned long long* __restrict__ keys, double* __restrict__ minmax) {
  if (threadIdx.x == 0) {
    const bool empty = keys[0] == ~0ull && keys[1] == 0ull;
    minmax[0] = empty ? CUDART_NAN : value_of(keys[0]);
    minmax[1] = empty ? CUDART_NAN : value_of(keys[1]);
  }
}

// mode 0: numpy.histogram(x, bins=n_bins) over [min, max] (matplotlib's plt.hist) — fast-path
//         index, then the +-1 corrections against linspace edges, last bin closed
//         (numpy/lib/_histograms_impl.py:851-863).
// mode 1: frontend rule idx = min(floor((v - min) / width), n_bins - 1), width = (max-min)/n_bins;
//         everything in bin 0 when max <= min (HistogramChart.jsx:31-52).
__global__ void k_histogram(const double* __restrict__ x, const uint8_t* __restrict__ mask, int64_t n, double divisor,
                            int n_bins, int mode, const double* __restrict__ range, unsigned long long* __restrict__ hist) {
  extern __shared__ uint32_t sh[];
  for (int k = threadIdx.x; k < n_bins; k += blockDim.x) sh[k] = 0;
  __syncthreads();
  double first = range[0], last = range[1];
  const bool empty = !(first == first);
  if (!empty) {
    if (mode == 0 && first == last) { first = __dsub_rn(first, 0.5); last = __dadd_rn(last, 0.5); }
    const double delta = __dsub_rn(last, first);
    const double step = __ddiv_rn(delta, (double)n_bins);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      if (mask && !mask[i]) continue;
      const double v = __ddiv_rn(x[i], divisor);
      if (v != v) continue;
      int idx;
      if (mode == 0) {
        if (v < first || v > last) continue;
        const double f = __dmul_rn(__ddiv_rn(__dsub_rn(v, first), delta), (double)n_bins);
        idx = (int)f;
        if (idx == n_bins) idx -= 1;
        auto edge = [&](int k) { return k == n_bins ? last : __dadd_rn(__dmul_rn((double)k, step), first); };
        if (v < edge(idx)) idx -= 1;
        if (v >= edge(idx + 1) && idx != n_bins - 1) idx += 1;
      } else {
        if (last <= first) idx = 0;
        else {
          const double fl = floor(__ddiv_rn(__dsub_rn(v, first), step));
          idx = fl < (double)(n_bins - 1) ? (int)fl : n_bins - 1;
          if (idx < 0) idx = 0;
        }
      }
      atomicAdd(&sh[idx], 1u);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n_bins; k += blockDim.x)
    if (sh[k]) atomicAdd(hist + k, (unsigned long long)sh[k]);
}

__global__ void k_gather(const double* __restrict__ series, int64_t ld, int rows, const int64_t* __restrict__ cols,
                         int n_cols, double* __restrict__ out) {
  const int k = blockIdx.x;
  if (k >= n_cols) return;
  for (int t = threadIdx.x; t < rows; t += blockDim.x) out[(int64_t)k * rows + t] = series[(int64_t)t * ld + cols[k]];
}

// DFMA-chain microbenchmark: 8 independent chains per thread, kPeakUnroll DFMAs per chain per
// iteration; every instruction is an FP64-pipe issue slot. The roofline denominator of SURVEY §8d.
constexpr int kPeakChains = 8;
constexpr int kPeakUnroll = 64;
__global__ void __launch_bounds__(256) k_fp64_peak(int iters, double seed, double* sink) {
  double a[kPeakChains];
#pragma unroll
  for (int c = 0; c < kPeakChains; ++c) a[c] = seed + (double)(threadIdx.x + c) * 1e-9;
  const double m = 1.0000000001, b = 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < kPeakUnroll; ++u) {
#pragma unroll
      for (int c = 0; c < kPeakChains; ++c) a[c] = fma(a[c], m, b);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < kPeakChains; ++c) s += a[c];
  if (s == 12345.678) sink[0] = s;  // keep the chains live
}

}  // namespace

cudaError_t launch_fp64_peak(int sm_count, int iters, double* sink, cudaStream_t st, int* total_threads,
                             int* dfma_per_thread) {
  const int blocks = sm_count * 4;
  k_fp64_peak<<<blocks, 256, 0, st>>>(iters, 1.0, sink);
  *total_threads = blocks * 256;
  *dfma_per_thread = iters * kPeakUnroll * kPeakChains;
  return cudaGetLastError();
}

size_t quantile_workspace_bytes(int rows) { return sizeof(SelRow) * (size_t)(rows > 0 ? rows : 1); }

cudaError_t launch_quantiles(const double* values, int64_t n, int64_t ld, int rows, const uint8_t* mask,
                             const QuantileSpec& spec, double* out, int64_t* counts, void* workspace, cudaStream_t st,
                             int* n_launches) {
  *n_launches = 0;
  if (rows <= 0) return cudaSuccess;
  SelRow* W = (SelRow*)workspace;
  const unsigned chunks = (unsigned)((n + kChunk - 1) / kChunk);
  const size_t smem = (size_t)kMaxTargets * 256 * sizeof(uint32_t) + kMaxTargets * sizeof(uint64_t);
  k_sel_init<<<rows, 256, 0, st>>>(W);
  ++*n_launches;
  for (int pass = 0; pass < 8; ++pass) {
    if (chunks > 0) {
      k_sel_hist<<<dim3(chunks, (unsigned)rows), kHistThreads, smem, st>>>(values, n, ld, mask, W, pass);
      ++*n_launches;
    }
    k_sel_advance<<<rows, 32, 0, st>>>(W, spec, pass);
    ++*n_launches;
  }
  k_sel_finish<<<rows, 32, 0, st>>>(W, spec, out, counts);
  ++*n_launches;
  return cudaGetLastError();
}

cudaError_t launch_rates(const double* start, const double* fy_real, int64_t n, double* rates, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_rates<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(start, fy_real, n, rates);
  return cudaGetLastError();
}

static unsigned reduce_grid(int64_t n) {
  int64_t g = (n + 255) / 256;
  const int64_t cap = 148 * 8;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

cudaError_t launch_minmax(const double* x, const uint8_t* mask, int64_t n, double divisor, unsigned long long* keys2,
                          double* minmax, cudaStream_t st) {
  k_minmax_init<<<1, 32, 0, st>>>(keys2);
  if (n > 0) k_minmax<<<reduce_grid(n), 256, 0, st>>>(x, mask, n, divisor, keys2);
  k_minmax_finish<<<1, 32, 0, st>>>(keys2, minmax);
  return cudaGetLastError();
}

cudaError_t launch_histogram(const double* x, const uint8_t* mask, int64_t n, double divisor, int n_bins, int mode,
                             const double* range_dev, int64_t* hist, cudaStream_t st) {
  k_histogram<<<reduce_grid(n), 256, sizeof(uint32_t) * (size_t)n_bins, st>>>(
      x, mask, n, divisor, n_bins, mode, range_dev, (unsigned long long*)hist);
  return cudaGetLastError();
}

cudaError_t launch_gather(const double* series, int64_t ld, int rows, const int64_t* cols_dev, int n_cols, double* out,
                          cudaStream_t st) {
  if (n_cols <= 0) return cudaSuccess;
  k_gather<<<n_cols, 128, 0, st>>>(series, ld, rows, cols_dev, n_cols, out);
  return cudaGetLastError();
}

}  // namespace mcr
