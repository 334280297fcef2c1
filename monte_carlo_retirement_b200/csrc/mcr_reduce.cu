// mcr_reduce.cu — device aggregations that replace the host-side pandas/numpy passes of
// run_monte_carlo_simulations and its consumers (/root/reference/backend/simulation.py:
// 1045-1118, :78-96; backend/server.py:439-461,525-532; backend/plotting.py:46-59;
// frontend/src/components/HistogramChart.jsx:13-60).
//
//   k_sel_*       exact order statistics by MSD radix select on the order-preserving 64-bit key:
//                 a sampled look at every row, ONE scan that histograms ~8 K equal bins laid over
//                 the sampled key range (every requested quantile of every row in the same
//                 scan), 8-bit digit scans for the rows that need more, a gather scan and a tail
//                 that walks the remaining digits on the gathered list; followed by numpy's
//                 'linear' lerp (numpy/lib/_function_base_impl.py: _QuantileMethods['linear'],
//                 _get_indexes, _lerp), the even/odd median rule of np.median, or the plain
//                 minimum / maximum.
//   k_rates       first-year withdrawal rates (simulation.py:92-95)
//   k_minmax / k_histogram   cohort min/max and equal-width histograms with numpy.histogram
//                 or frontend floor binning
//   k_gather      sample-path columns (simulation.py:1068-1078)
// Compiled with -fmad=false: the interpolation arithmetic must round like numpy's.
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "mcr_reduce.h"

#ifndef MCR_SCAN_UNROLL
#define MCR_SCAN_UNROLL 4
#endif
#ifndef MCR_FIRST_CHUNK
#define MCR_FIRST_CHUNK 32768
#endif

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
//
// MSD radix select on the order-preserving 64-bit key. A pass = one grid-wide histogram kernel
// (rows x chunks CTAs, every requested quantile of a row resolved in the same scan, merged into
// hist[row]) + an "advance" step that walks each target's bucket to the bin holding its rank —
// a kernel of its own in the multi-GPU protocol, the last CTA of the row's scan otherwise. What
// keeps the number of FULL scans of a row at two (histogram + gather) for ordinary data:
//   * adaptive start: pass 0 only looks at a sample of the row for its extreme keys (zeros
//     aside); pass 1 lays kFirstBins equal bins over that range — ~13 bits resolved by one scan
//     however the range sits in key space — and counts what falls outside exactly. A constant
//     row is finished after pass 1; a tightly concentrated row (early accumulation years) is as
//     cheap as a wide one;
//   * collect + tail: as soon as the live buckets of a row fit its candidate list together
//     (after pass 1 for C3's 1e6-element rows; one or two 8-bit digit passes later for rows of
//     1e8..1e9 elements or outlier-stretched ranges) the row stops scanning; one more scan gathers
//     the elements that share a target's prefix into the list; the tail kernel walks the
//     remaining digits on that list and interpolates. A bucket that stays too big for the list
//     but holds ONE repeated value (zero-padded failed paths) is recognised by its min == max;
//     only a big bucket of distinct values after all full passes falls back to scanning the row.
// The per-pass split also lets a multi-GPU caller all-reduce hist[] between the two halves of
// a pass and obtain exact GLOBAL order statistics without moving data; the row
// extremes then take one all-reduce(MIN) of their own (k_sel_extremes), without which that
// stepwise protocol falls back to fixed 8-bit digits from bit 0 (adaptive = 0).
// (hist[row] holds [groups][256] bins in the digit passes and kFirstBins + 4 words in pass 1.)
// Layout: hist[rows][kMaxTargets][256] u32 is its own contiguous array (the all-reduced
// buffer); the rest of the state is SelRow[rows] followed by the candidate lists.
struct SelRow {
  RowDesc d;                        // what this row selects from / for (copied in by BEGIN)
  uint64_t prefix[kMaxTargets];     // per target (rank-sorted): the `rb` key bits resolved so far
  int64_t rank[kMaxTargets];        // per target: rank inside the current prefix bucket
  uint64_t uprefix[kMaxTargets];    // sorted unique prefixes (groups) of the current pass
  uint64_t gmin[kMaxTargets];       // per group at collect time: smallest / largest key seen in the bucket
  uint64_t gmax[kMaxTargets];       //   (equal => the bucket is one repeated value)
  int64_t bcount[kMaxTargets];      // per target: population of its current bucket
  int32_t gbig[kMaxTargets];        // per group: bucket too large for the candidate list (not gathered)
  int32_t group[kMaxTargets];       // target -> group
  int32_t slot[kMaxTargets];        // (quantile, lo/hi) -> rank-sorted target slot
  uint64_t kmin, kmax;              // row extremes over the keys != +0.0 (pass 0, adaptive mode)
  uint64_t fbase;                   // first digit pass of an adaptive row: bin = (key >> fshift) - fbase
  uint64_t omin, omax;              // extreme keys seen OUTSIDE the window by the first digit pass (this GPU's elements)
  uint64_t fixed_key[kMaxTargets];  // per target of fixed_mask: its key
  uint32_t fixed_mask;              // targets resolved by the first digit pass itself (see advance_row) ...
  uint32_t omin_mask, omax_mask;    // ... of which: the row's minimum / maximum, i.e. omin / omax once those are final
  int32_t fshift;                   //   (64: no window — nothing sampled, or the row restarted from bit 0)
  int32_t cap;                      // candidate list capacity of this call (kCandCap; kPoolCap when ranks pool lists)
  int32_t whole;                    // the row is complete on this GPU (single-GPU call): omin / omax are the row's
  uint32_t done;                    // CTAs of the current scan that have finished this row (fused launches, see k_sel_hist)
  uint32_t pad1, pad2, pad3;
  int64_t n_valid;
  int32_t n_groups;
  int32_t rb;                       // resolved bits (64 == done)
  int32_t n_cand;                   // elements in this row's candidate list
  int32_t overflow;                 // candidate list overflowed
  int32_t collected;                // candidate list is valid
  int32_t adaptive;                 // row extremes are available after pass 0
  int32_t fused;                    // single-GPU call: k_sel_tail finishes the row (ungathered buckets allowed)
  int32_t ready;                    // every live bucket fits the candidate list: no more full scans needed
};

constexpr int kHistThreads = 256;
constexpr int kChunk = 16384;       // elements of one row handled by one CTA (short rows; long rows: chunk_for)
constexpr int kMaxChunksPerRow = 1024;  // a long row is cut into at most this many chunks: fewer CTAs merging their
                                        // shared histograms into the row's global one (1.25e8-element rows: 7630 -> 954)
constexpr int kCandCap = 16384;     // candidate list capacity per row (doubles)
constexpr int kPoolCap = 8192;      // ... of the list the ranks of a multi-GPU call pool (all-reduced: kept small)
constexpr int kBigBucket = 4096;    // buckets above this are not gathered (resolved by min == max, else by scanning)
constexpr int kFullPasses = 3;      // passes that may scan the rows before the collect: the sample, the ~13-bit first digit
                                    // pass and one 8-bit digit (21 bits over the sampled range);
constexpr int kFullPassesLong = 4;  // one more digit for rows of more than 2^24 elements (GLOBAL length), so that their
constexpr int64_t kLongRow = (int64_t)1 << 24;   // buckets still fit the candidate lists
constexpr int kHistWords = kMaxTargets * 256;  // per row
constexpr int kSampleStride = 16;   // adaptive pass 0 reads every 16th chunk of a long row ...
constexpr int kSampleMinChunks = 8; // ... rows of up to 8 chunks are read whole
// First digit pass of an adaptive row: kFirstBins equal bins laid over the sampled key range [kmin, kmax] —
// bin = (key >> s) - (kmin >> s) with the smallest s that fits the range — instead of an 8-bit digit under
// the common prefix of the two: at least half of the bins are in use whatever the range straddles (a range
// across a power of two shares almost no leading bits), ~13 bits are resolved by ONE scan, and every bin is
// still the set of keys with one (64 - s)-bit prefix, so that the following 8-bit passes apply unchanged.
constexpr int kFirstBins = 8128;
constexpr int kFirstChunk = MCR_FIRST_CHUNK;  // elements per CTA of that pass: few CTAs merge 8 K-bin histograms into the row's
constexpr int kBelowAt = kFirstBins, kAboveAt = kFirstBins + 1;  // H slots of the keys outside the sampled range
constexpr int kBelowZeroAt = kFirstBins + 2, kBelowNegAt = kFirstBins + 3;  // ... of which: exactly +0.0 / negative
constexpr int kFirstWords = kFirstBins + 4;
constexpr int kFirstBlocks = (kFirstBins + 255) / 256;   // 256-bin blocks of the window histogram (<= 32)
static_assert(kFirstBlocks <= 32, "one lane per block in advance_row");
static_assert(kFirstWords <= kHistWords, "first-digit histogram must fit the row's histogram");
// The engine pads the yearly series of a failed path with +0.0 (simulation.py:905-912) and clamps final
// balances at 0, so rows hold a mass of exact zeros next to a bulk of positive balances. Zeros share
// only the sign bit with the bulk: a common prefix over all keys would be worthless and the zero
// bucket never shrinks. The adaptive start therefore takes the row extremes over the keys != +0.0,
// the first digit pass counts the keys below the prefix in three classes (negative, +0.0, small
// positive), and a target whose rank falls among the zeros IS +0.0 — resolved on the spot. (Keys
// the sample missed just below the bulk or below zero only matter if a target lands on them.)
constexpr uint64_t kZeroKey = 0x8000000000000000ull;

__global__ void k_sel_init(SelRow* __restrict__ rows, uint32_t* __restrict__ hist, const RowDesc* __restrict__ desc,
                           int adaptive, int fused) {
  SelRow& R = rows[blockIdx.x];
  if (threadIdx.x == 0) R.d = desc[blockIdx.x];
  uint32_t* H = hist + (size_t)blockIdx.x * kHistWords;
  for (int k = threadIdx.x; k < kHistWords; k += blockDim.x) H[k] = 0;
  if (threadIdx.x == 0) {
    R.n_groups = 1; R.n_valid = 0; R.n_cand = 0; R.overflow = 0; R.collected = 0; R.rb = 0;
    R.kmin = ~0ull; R.kmax = 0ull; R.adaptive = adaptive; R.fused = fused != 0; R.ready = 0; R.fixed_mask = 0;
    R.fbase = 0; R.fshift = 64; R.cap = (fused & 2) ? kPoolCap : kCandCap;
    R.omin = ~0ull; R.omax = 0ull; R.whole = fused == 1; R.done = 0; R.omin_mask = 0; R.omax_mask = 0;
  }
  if (threadIdx.x < kMaxTargets) { R.gmin[threadIdx.x] = ~0ull; R.gmax[threadIdx.x] = 0ull; }
}

// ---- scan machinery ---------------------------------------------------------------------------
// The scans are instruction-issue bound before they are HBM bound (ncu, r01h: 77 warp
// instructions per element in the first version), so everything per-element is kept to 32-bit
// integer work on the two halves of the key, the pass geometry lives in registers (Probe), the
// shared-memory tables are addressed through precomputed 32-bit shared-window addresses, and the
// loop over a chunk has a branch-free body for full tiles.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void lds_u64(uint32_t a, uint32_t& lo, uint32_t& hi) {
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(a) : "memory");
}
// a value the compiler must keep in a register (a shared-window address it would otherwise rebuild from
// SR_CgaCtaId + constants — 4 instructions — at every use inside an unrolled loop)
__device__ __forceinline__ uint32_t opaque(uint32_t v) {
  asm volatile("mov.u32 %0, %0;" : "+r"(v));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
// red.shared.add under a predicate: one predicated instruction, no branch / reconvergence pair around it
__device__ __forceinline__ void red_shared_add_if(bool p, uint32_t a, uint32_t v) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q red.shared.add.u32 [%0], %1;\n\t}" ::"r"(a), "r"(v), "r"((uint32_t)p) : "memory");
}
__device__ __forceinline__ void red_shared_add(uint32_t a, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// order-preserving key as two 32-bit halves (3 integer instructions)
__device__ __forceinline__ void key_halves(double v, uint32_t& hi, uint32_t& lo) {
  lo = (uint32_t)__double2loint(v);
  hi = (uint32_t)__double2hiint(v);
  const uint32_t s = (uint32_t)((int32_t)hi >> 31);
  lo ^= s;
  hi ^= s | 0x80000000u;
}

// membership of a key among the live prefixes: LEFT-ALIGNED prefixes in a 256-entry chained
// table keyed by the last resolved byte (the most discriminating one)
struct PrefixTable {
  uint64_t up[kMaxTargets];   // prefix << (64 - rb)
  uint8_t head[256];          // group + 1, 0 == none
  uint8_t next[kMaxTargets];  // chain
};

__device__ __forceinline__ void build_table(PrefixTable& T, const SelRow& R, int ng, int rb) {
  for (int k = threadIdx.x; k < 256; k += blockDim.x) T.head[k] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int g = 0; g < ng; ++g) {
      const uint64_t u = R.uprefix[g];
      T.up[g] = u << (64 - rb);
      T.next[g] = T.head[u & 255u];
      T.head[u & 255u] = (uint8_t)(g + 1);
    }
  }
  __syncthreads();
}

// geometry of the current pass over one row, uniform across the CTA
struct Probe {
  uint32_t m_hi, m_lo;   // mask of the rb resolved (leading) key bits
  uint32_t p_hi, p_lo;   // group 0's left-aligned prefix (single-group fast path)
  int pshift;            // 64 - rb        : key >> pshift == prefix
  int dshift;            // 64 - rb - w    : position of the digit being histogrammed
  uint32_t dmask;
  uint32_t a_head, a_up, a_next, a_hist;  // shared-window addresses
};

__device__ __forceinline__ Probe make_probe(const SelRow& R, int rb, int w, const PrefixTable& T, const uint32_t* sh) {
  Probe P;
  const uint64_t m = rb == 0 ? 0ull : (~0ull << (64 - rb));
  const uint64_t p0 = rb == 0 ? 0ull : (R.uprefix[0] << (64 - rb));
  P.m_hi = (uint32_t)(m >> 32); P.m_lo = (uint32_t)m;
  P.p_hi = (uint32_t)(p0 >> 32); P.p_lo = (uint32_t)p0;
  P.pshift = 64 - rb;
  P.dshift = 64 - rb - w;
  P.dmask = (1u << w) - 1u;
  P.a_head = smem_addr(T.head); P.a_up = smem_addr(T.up); P.a_next = smem_addr(T.next);
  P.a_hist = smem_addr(sh);
  return P;
}

enum ScanMode { kScanTop = 1, kScanOne = 2, kScanTable = 3 };  // rb == 0 | one live prefix | several

template <int MODE>
__device__ __forceinline__ int group_of(const Probe& P, uint32_t hi, uint32_t lo) {
  if (MODE == kScanTop) return 0;
  if (MODE == kScanOne) return (((hi ^ P.p_hi) & P.m_hi) | ((lo ^ P.p_lo) & P.m_lo)) == 0u ? 0 : -1;
  const uint64_t k = ((uint64_t)hi << 32) | lo;
  uint32_t g = lds_u8(P.a_head + ((uint32_t)(k >> P.pshift) & 255u));
  while (g) {
    uint32_t ulo, uhi;
    lds_u64(P.a_up + (g - 1u) * 8u, ulo, uhi);
    if ((((hi ^ uhi) & P.m_hi) | ((lo ^ ulo) & P.m_lo)) == 0u) return (int)g - 1;
    g = lds_u8(P.a_next + g - 1u);
  }
  return -1;
}

// body(ok, v, hi, lo) for every element of x[0, cnt); all 32 lanes of a warp call it together
// (the body may shuffle). Masked-out and NaN elements arrive with ok == false (pandas skips NaN).
template <int NT, bool MASKED, typename Body>
__device__ __forceinline__ void scan_elements(const double* __restrict__ x, const uint8_t* __restrict__ m, int cnt,
                                              Body body) {
  constexpr int U = MCR_SCAN_UNROLL;  // independent loads in flight per thread
  const int tid = threadIdx.x;
  int base = 0;
  for (; base + U * NT <= cnt; base += U * NT) {  // full tiles: no bounds checks
    double v[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = base + u * NT + tid;
      v[u] = __ldcs(x + e);
      ok[u] = MASKED ? (m[e] != 0) : true;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint32_t hi, lo;
      key_halves(v[u], hi, lo);
      body(ok[u] && v[u] == v[u], v[u], hi, lo);
    }
  }
  for (; base < cnt; base += NT) {  // ragged end
    const int e = base + tid;
    bool ok = e < cnt;
    if (MASKED) ok = ok && m[e] != 0;
    const double v = ok ? __ldcs(x + e) : 0.0;
    uint32_t hi, lo;
    key_halves(v, hi, lo);
    body(ok && v == v, v, hi, lo);
  }
}

// shared-memory histogram of the next digit inside every live prefix bucket
template <int NT, bool MASKED, int MODE>
__device__ __forceinline__ void hist_elements(const Probe& P, const double* __restrict__ x,
                                              const uint8_t* __restrict__ m, int cnt) {
  const bool lane0 = (threadIdx.x & 31) == 0;
  scan_elements<NT, MASKED>(x, m, cnt, [&](bool ok, double, uint32_t hi, uint32_t lo) {
    int id = -1;
    if (ok) {
      const int g = group_of<MODE>(P, hi, lo);
      const uint64_t k = ((uint64_t)hi << 32) | lo;
      if (g >= 0) id = g * 256 + (int)((uint32_t)(k >> P.dshift) & P.dmask);
    }
    // concentrated data puts whole warps in one bin (and most keys of a late pass in none): one
    // atomic — or none — for the warp in that case
    const int id0 = __shfl_sync(0xffffffffu, id, 0);
    if (__all_sync(0xffffffffu, id == id0)) {
      if (id0 >= 0 && lane0) red_shared_add(P.a_hist + (uint32_t)id0 * 4u, 32u);
    } else if (id >= 0) {
      red_shared_add(P.a_hist + (uint32_t)id * 4u, 1u);
    }
  });
}

template <int NT>
__device__ __forceinline__ void hist_dispatch(const Probe& P, const double* __restrict__ x, const uint8_t* __restrict__ m,
                                              int cnt, int rb, int ng) {
  if (rb == 0) {
    if (m) hist_elements<NT, true, kScanTop>(P, x, m, cnt); else hist_elements<NT, false, kScanTop>(P, x, m, cnt);
  } else if (ng == 1) {
    if (m) hist_elements<NT, true, kScanOne>(P, x, m, cnt); else hist_elements<NT, false, kScanOne>(P, x, m, cnt);
  } else {
    if (m) hist_elements<NT, true, kScanTable>(P, x, m, cnt); else hist_elements<NT, false, kScanTable>(P, x, m, cnt);
  }
}

// adaptive pass 0: no digit yet — the extreme keys of (a sample of) the row. The leading bits the
// two share are very likely shared by every key of the row; the first digit pass verifies that
// exactly (keys outside are counted as below / above) and counts the valid elements.
template <int NT, bool MASKED>
__device__ __forceinline__ void extremes_elements(const double* __restrict__ x, const uint8_t* __restrict__ m, int cnt,
                                                  SelRow& R) {
  uint64_t lo_k = ~0ull, hi_k = 0ull;
  scan_elements<NT, MASKED>(x, m, cnt, [&](bool ok, double, uint32_t hi, uint32_t lo) {
    const uint64_t k = ((uint64_t)hi << 32) | lo;
    if (ok && k != kZeroKey) {
      lo_k = k < lo_k ? k : lo_k;
      hi_k = k > hi_k ? k : hi_k;
    }
  });
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t l2 = __shfl_xor_sync(0xffffffffu, lo_k, o), h2 = __shfl_xor_sync(0xffffffffu, hi_k, o);
    lo_k = l2 < lo_k ? l2 : lo_k;
    hi_k = h2 > hi_k ? h2 : hi_k;
  }
  if ((threadIdx.x & 31) == 0 && lo_k <= hi_k) {
    atomicMin((unsigned long long*)&R.kmin, (unsigned long long)lo_k);
    atomicMax((unsigned long long*)&R.kmax, (unsigned long long)hi_k);
  }
}

// first digit pass of an adaptive row: histogram of the keys over the kFirstBins bins of the sampled
// range, and the number of keys below / above it. HI32: the bin shift is >= 32 (the usual case — a row
// whose values spread by more than ~1e-4 relative), so the bin index comes from the key's high word alone.
template <int NT, bool MASKED, bool HI32>
__device__ __forceinline__ void first_digit_elements(uint32_t a_hist, uint64_t fbase, int fshift,
                                                     const double* __restrict__ x, const uint8_t* __restrict__ m,
                                                     int cnt, unsigned long long* s_out) {
  const bool lane0 = (threadIdx.x & 31) == 0;
  const uint32_t base32 = (uint32_t)fbase;       // HI32: fbase = kmin >> fshift fits 32 bits
  const int sh32 = fshift - 32;
  // the hot path is branch-free: bin index, one predicated shared-memory atomic, one vote
  scan_elements<NT, MASKED>(x, m, cnt, [&](bool ok, double, uint32_t hi, uint32_t lo) {
    uint32_t bin;
    bool above;
    if (HI32) {
      const uint32_t q = hi >> sh32;
      bin = q - base32;
      above = q > base32;                          // only read when outside
    } else {
      const uint64_t q = (((uint64_t)hi << 32) | lo) >> fshift;
      const uint64_t d = q - fbase;
      bin = d < (uint64_t)kFirstBins ? (uint32_t)d : 0xffffffffu;
      above = q > fbase;
    }
    const bool inside = ok && bin < (uint32_t)kFirstBins;
    red_shared_add_if(inside, a_hist + bin * 4u, 1u);
    const bool outside = ok && !inside;
    if (__any_sync(0xffffffffu, outside)) {
      // keys outside the sampled range (zero-padded failures — possibly most of the row —, whatever the
      // sample missed): one atomic per class and warp
      const bool below = outside && !above;
      const unsigned n_above = __popc(__ballot_sync(0xffffffffu, outside && above));
      const unsigned n_below = __popc(__ballot_sync(0xffffffffu, below));
      const unsigned n_zero = __popc(__ballot_sync(0xffffffffu, below && hi == 0x80000000u && lo == 0u));
      const unsigned n_neg = __popc(__ballot_sync(0xffffffffu, below && hi < 0x80000000u));
      // ... and the extreme keys out there: a target of rank 0 / n - 1 (a requested minimum / maximum,
      // which a sample practically never sees) is then known without restarting the row
      const unsigned long long k = ((unsigned long long)hi << 32) | lo;
      unsigned long long k_lo = below ? k : ~0ull, k_hi = (outside && above) ? k : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, k_lo, o), h2 = __shfl_xor_sync(0xffffffffu, k_hi, o);
        k_lo = l2 < k_lo ? l2 : k_lo;
        k_hi = h2 > k_hi ? h2 : k_hi;
      }
      if (lane0) {
        if (n_above) red_shared_add(a_hist + (uint32_t)kAboveAt * 4u, n_above);
        if (n_below) red_shared_add(a_hist + (uint32_t)kBelowAt * 4u, n_below);
        if (n_zero) red_shared_add(a_hist + (uint32_t)kBelowZeroAt * 4u, n_zero);
        if (n_neg) red_shared_add(a_hist + (uint32_t)kBelowNegAt * 4u, n_neg);
        if (n_below) atomicMin(&s_out[0], k_lo);
        if (n_above) atomicMax(&s_out[1], k_hi);
      }
    }
  });
}

// One pass, first half: histogram of the next digit inside every live prefix bucket. Rows that
// are finished or ready for the collect skip; rows with a valid candidate list scan that list
// (one CTA).
// chunk: elements per CTA (a multiple of 4 * kHistThreads); stride: pass 0 of an adaptive call
// launches one CTA per `stride` chunks (the sample), every other pass has stride 1.
__device__ __forceinline__ void hist_cta(SelRow& R, uint32_t* __restrict__ H, const double* __restrict__ cand, int pass,
                                         int chunk, int stride, uint32_t* sh) {
  __shared__ PrefixTable T;
  __shared__ unsigned long long s_out[2];
  const int rb = R.rb;
  const bool from_cand = R.collected && !R.overflow;
  if (from_cand && blockIdx.x > 0) return;
  const int64_t n = R.d.n;
  const bool sampling = pass == 0 && R.adaptive;
  // pass 0 of an adaptive row: a long row is sampled, one 16 K-element piece out of `stride` (what the
  // sample misses is handled exactly by the below / above counts of the next pass); rows of up to
  // kSampleMinChunks pieces are read whole by their first CTAs
  int64_t begin = (int64_t)blockIdx.x * chunk;
  if (sampling) {
    const int64_t pieces = (n + kChunk - 1) / kChunk;
    begin = pieces > kSampleMinChunks ? (int64_t)blockIdx.x * stride * kChunk : (int64_t)blockIdx.x * kChunk;
    chunk = kChunk;
  }
  if (from_cand) begin = 0;
  if (!from_cand && begin >= n) return;
  const int cnt = from_cand ? R.n_cand : (int)(begin + chunk < n ? chunk : n - begin);
  const double* __restrict__ x = (from_cand ? cand + (size_t)blockIdx.y * kCandCap : R.d.x) + begin;
  const uint8_t* __restrict__ m = (from_cand || !R.d.mask) ? nullptr : R.d.mask + begin;
  if (sampling) {
    if (m) extremes_elements<kHistThreads, true>(x, m, cnt, R); else extremes_elements<kHistThreads, false>(x, m, cnt, R);
    return;
  }
  const int fshift = R.fshift;
  const bool first_digit = pass == 1 && R.adaptive && fshift < 64;  // counts keys outside the sampled range too
  const int ng = R.n_groups;
  const int used = first_digit ? kFirstWords : ng * 256;
  for (int k = threadIdx.x; k < used; k += kHistThreads) sh[k] = 0;
  if (threadIdx.x == 0) { s_out[0] = ~0ull; s_out[1] = 0ull; }
  if (!first_digit && rb > 0 && ng > 1) build_table(T, R, ng, rb); else __syncthreads();
  if (first_digit) {
    const uint32_t a_hist = opaque(smem_addr(sh));
    const uint64_t fbase = R.fbase;
    if (fshift >= 32) {
      if (m) first_digit_elements<kHistThreads, true, true>(a_hist, fbase, fshift, x, m, cnt, s_out);
      else first_digit_elements<kHistThreads, false, true>(a_hist, fbase, fshift, x, m, cnt, s_out);
    } else {
      if (m) first_digit_elements<kHistThreads, true, false>(a_hist, fbase, fshift, x, m, cnt, s_out);
      else first_digit_elements<kHistThreads, false, false>(a_hist, fbase, fshift, x, m, cnt, s_out);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (s_out[0] != ~0ull) atomicMin((unsigned long long*)&R.omin, s_out[0]);
      if (s_out[1] != 0ull) atomicMax((unsigned long long*)&R.omax, s_out[1]);
    }
  } else {
    const int w = 64 - rb < 8 ? 64 - rb : 8;
    const Probe P = make_probe(R, rb, w, T, sh);
    hist_dispatch<kHistThreads>(P, x, m, cnt, rb, ng);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < used; k += kHistThreads)
    if (sh[k]) atomicAdd(&H[k], sh[k]);
}

// After the full passes: gather the elements that share a live prefix into the row's list. With several
// live prefixes (the usual case) a 64 K-bit filter over the last 16 resolved key bits decides, for a whole
// warp at once, that none of its elements is a candidate — ~95 % of the warps; the exact membership test
// and the list append only run for the others.
constexpr int kFilterBits = 16;
constexpr int kFilterWords = (1 << kFilterBits) / 32;

template <bool MASKED, int MODE, bool HI32>
__device__ __forceinline__ void collect_elements(const Probe& P, uint32_t a_filter, uint32_t fmask,
                                                 const double* __restrict__ x, const uint8_t* __restrict__ m, int cnt,
                                                 SelRow& R, double* __restrict__ out, unsigned long long* s_min,
                                                 unsigned long long* s_max, const int* s_big) {
  const int cap = R.cap;
  const unsigned lane = threadIdx.x & 31u;
  const int psh32 = P.pshift - 32;
  scan_elements<kHistThreads, MASKED>(x, m, cnt, [&](bool ok, double v, uint32_t hi, uint32_t lo) {
    bool maybe = ok;
    if (MODE == kScanOne) maybe = ok && (((hi ^ P.p_hi) & P.m_hi) | ((lo ^ P.p_lo) & P.m_lo)) == 0u;
    if (MODE == kScanTable) {
      uint32_t idx = HI32 ? (hi >> psh32) : (uint32_t)((((uint64_t)hi << 32) | lo) >> P.pshift);
      idx &= fmask;
      const uint32_t bit = (lds_u32(a_filter + ((idx >> 5) << 2)) >> (idx & 31u)) & 1u;   // (NaN lanes too: no branch)
      maybe = ok & (bit != 0u);
    }
    if (!__any_sync(0xffffffffu, maybe)) return;
    const int g = maybe ? (MODE == kScanTable ? group_of<kScanTable>(P, hi, lo) : 0) : -1;
    bool take = false;
    if (g >= 0) {
      // bucket extremes (shared-memory atomics only — no plain read races them —, merged once per
      // CTA): a bucket of one repeated value is recognised later by min == max. Few elements get
      // here: the live buckets are small by now, except for repeated values (zero-padded failures).
      const unsigned long long k = ((unsigned long long)hi << 32) | lo;
      atomicMin(&s_min[g], k);
      atomicMax(&s_max[g], k);
      take = !s_big[g];
    }
    // one list-cursor atomic per warp; an overflowing list is detected from n_cand by k_sel_collect_finish
    const unsigned takers = __ballot_sync(0xffffffffu, take);
    if (takers) {
      const int leader = __ffs(takers) - 1;
      int at = 0;
      if ((int)lane == leader) at = atomicAdd(&R.n_cand, __popc(takers));
      at = __shfl_sync(0xffffffffu, at, leader) + __popc(takers & ((1u << lane) - 1u));
      if (take && at < cap) out[at] = v;
    }
  });
}

__device__ __forceinline__ void collect_cta(SelRow& R, double* __restrict__ cand, int chunk) {
  __shared__ PrefixTable T;
  __shared__ unsigned long long s_min[kMaxTargets], s_max[kMaxTargets];
  __shared__ int s_big[kMaxTargets];
  __shared__ uint32_t s_filter[kFilterWords];
  const int rb = R.rb;
  const int64_t n = R.d.n;
  if ((int64_t)blockIdx.x * chunk >= n) return;
  const int ng = R.n_groups;
  if (threadIdx.x < kMaxTargets) {
    s_min[threadIdx.x] = ~0ull; s_max[threadIdx.x] = 0ull;
    s_big[threadIdx.x] = threadIdx.x < ng ? R.gbig[threadIdx.x] : 0;
  }
  const bool table = rb > 0 && ng > 1;
  const uint32_t fmask = (1u << (rb < kFilterBits ? rb : kFilterBits)) - 1u;
  if (table)
    for (int k = threadIdx.x; k < kFilterWords; k += kHistThreads) s_filter[k] = 0;
  if (rb > 0) build_table(T, R, ng, rb); else __syncthreads();
  if (table && threadIdx.x < ng) {
    const uint32_t idx = (uint32_t)R.uprefix[threadIdx.x] & fmask;
    atomicOr(&s_filter[idx >> 5], 1u << (idx & 31u));
  }
  __syncthreads();
  const int64_t begin = (int64_t)blockIdx.x * chunk;
  const int cnt = (int)(begin + chunk < n ? chunk : n - begin);
  const double* __restrict__ x = R.d.x + begin;
  const uint8_t* __restrict__ m = R.d.mask ? R.d.mask + begin : nullptr;
  double* __restrict__ out = cand + (size_t)blockIdx.y * kCandCap;
  const Probe P = make_probe(R, rb, 0, T, nullptr);
  const uint32_t a_filter = opaque(smem_addr(s_filter));
#define MCR_COLLECT(MODE, HI32)                                                                            \
  do {                                                                                                     \
    if (m) collect_elements<true, MODE, HI32>(P, a_filter, fmask, x, m, cnt, R, out, s_min, s_max, s_big);  \
    else collect_elements<false, MODE, HI32>(P, a_filter, fmask, x, m, cnt, R, out, s_min, s_max, s_big);   \
  } while (0)
  if (rb == 0) MCR_COLLECT(kScanTop, false);
  else if (ng == 1) MCR_COLLECT(kScanOne, false);
  else if (P.pshift >= 32) MCR_COLLECT(kScanTable, true);
  else MCR_COLLECT(kScanTable, false);
#undef MCR_COLLECT
  __syncthreads();
  if (threadIdx.x < ng && s_min[threadIdx.x] != ~0ull) {
    atomicMin((unsigned long long*)&R.gmin[threadIdx.x], s_min[threadIdx.x]);
    atomicMax((unsigned long long*)&R.gmax[threadIdx.x], s_max[threadIdx.x]);
  }
}

// Row extremes <-> a caller buffer ext[rows][2] of int64 in an order-preserving signed encoding
// chosen so that ONE all-reduce(MIN) across ranks yields the global min and max:
//   ext[r][0] = kmin ^ 2^63,  ext[r][1] = ~kmax ^ 2^63   (min over ranks of ~kmax == ~max kmax)
__global__ void k_sel_extremes(SelRow* __restrict__ rows, long long* __restrict__ ext, int n_rows, int store) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  SelRow& R = rows[r];
  const unsigned long long top = 0x8000000000000000ull;
  if (!store) {
    ext[2 * r] = (long long)(R.kmin ^ top);
    ext[2 * r + 1] = (long long)((~R.kmax) ^ top);
  } else {
    R.kmin = (unsigned long long)ext[2 * r] ^ top;
    R.kmax = ~((unsigned long long)ext[2 * r + 1] ^ top);
  }
}

// one thread, after every CTA of the collect has finished the row
__device__ __forceinline__ void collect_finish_row(SelRow& R) {
  const int n_cand = atomicAdd(&R.n_cand, 0);   // (the cursor the scan CTAs advanced: read at L2)
  if (n_cand > R.cap) { R.n_cand = R.cap; R.overflow = 1; }
  // stepwise (multi-GPU) protocol: later passes histogram the list, so a row with a bucket
  // that was not gathered keeps scanning the full row instead
  if (!R.fused)
    for (int g = 0; g < R.n_groups; ++g)
      if (R.gbig[g]) R.overflow = 1;
  R.collected = 1;
  R.ready = 0;  // later (stepwise) passes histogram the candidate list
}

__global__ void k_sel_collect_finish(SelRow* __restrict__ rows) {
  SelRow& R = rows[blockIdx.x];
  if (threadIdx.x == 0 && R.rb < 64 && !R.collected) collect_finish_row(R);
}

// the 0-based order statistics a row asks for, given its valid count: (lo, hi) per quantile,
// then rank-sorted so that prefixes stay sorted in every pass (for small n the pairs of
// different quantiles interleave). One thread.
__device__ void set_target_ranks(SelRow& R, int64_t nv) {
  const QuantileSpec& spec = R.d.spec;
  const int nq = spec.n_q, nt = 2 * nq;
  R.n_valid = nv;
  for (int q = 0; q < nq; ++q) {
    int64_t lo = 0, hi = 0;
    if (nv > 0) {
      if (spec.median == 1) {  // np.median: mean of the two middle order statistics
        lo = (nv - 1) / 2;
        hi = nv / 2;
      } else if (spec.median == 2) {  // {min, max}: the order statistics themselves
        lo = hi = q == 0 ? 0 : nv - 1;
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
  int32_t ord[kMaxTargets];
  for (int t = 0; t < nt; ++t) ord[t] = t;
  for (int i = 1; i < nt; ++i) {  // tiny insertion sort
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
  }
}

// One pass, second half (one warp per target + housekeeping): consume the merged histogram H
// of row R. Needs >= 32 * (2 * n_q) threads; ends with H cleared and the next pass's groups set.
//   adaptive rows:  pass 0 took (sampled) extreme keys -> the bin window of the first digit pass, nothing
//                   counted yet; pass 1 histogrammed the kFirstBins bins of that window, plus the keys
//                   outside it -> valid count, target ranks, first walk (64 - fshift bits resolved). A
//                   target that falls outside the window (only possible if the sample missed that
//                   much of the row) restarts the row from bit 0 as a non-adaptive one;
//   other rows:     pass 0 histogrammed the top digit -> valid count, target ranks, first walk.
// HG: H is the row's histogram in GLOBAL memory, merged by the atomics of other CTAs — of the SAME launch when
// the advance is fused into the scan kernel: read through L2.
template <bool HG>
__device__ __forceinline__ uint32_t ld_hist(const uint32_t* p) {
  return HG ? __ldcg(p) : *p;
}

template <bool HG>
__device__ __forceinline__ void advance_row(SelRow& R, uint32_t* H, int pass) {
  __shared__ int s_restart;
  __shared__ uint32_t s_blocksum[kFirstBlocks];
  const QuantileSpec& spec = R.d.spec;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int nt = 2 * spec.n_q;
  const int rb = R.rb;
  const bool adaptive = R.adaptive != 0;
  const bool extremes_only = adaptive && pass == 0;
  const int fshift = R.fshift;
  const bool first_digit = adaptive && pass == 1 && fshift < 64;  // window bins; below / above slots are live
  const bool counts_now = adaptive ? pass == 1 : pass == 0;       // this histogram carries the valid count
  if (tid == 0) s_restart = 0;
  __syncthreads();
  if (rb < 64 && !R.ready) {
    if (counts_now) {
      if (first_digit) {  // sums of the 256-bin blocks of the window histogram, all warps
        for (int blk = warp; blk < kFirstBlocks; blk += (int)(blockDim.x >> 5)) {
          uint32_t local = 0;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int b = blk * 256 + lane * 8 + k;
            local += b < kFirstBins ? ld_hist<HG>(H + b) : 0u;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
          if (lane == 0) s_blocksum[blk] = local;
        }
        __syncthreads();
      }
      if (warp == 0) {
        int64_t part = 0;
        if (first_digit) part = lane < kFirstBlocks ? (int64_t)s_blocksum[lane] : 0;
        else for (int d = lane; d < 256; d += 32) part += ld_hist<HG>(H + d);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) {
          const int64_t below = first_digit ? (int64_t)ld_hist<HG>(H + kBelowAt) : 0, above = first_digit ? (int64_t)ld_hist<HG>(H + kAboveAt) : 0;
          const int64_t below_neg = first_digit ? (int64_t)ld_hist<HG>(H + kBelowNegAt) : 0;
          const int64_t below_zero = first_digit ? (int64_t)ld_hist<HG>(H + kBelowZeroAt) : 0;
          set_target_ranks(R, part + below + above);
          if (!adaptive) {
            for (int t = 0; t < nt; ++t) { R.prefix[t] = 0; R.group[t] = 0; }
          } else {
            // ranks inside the prefix bucket. A target below it is +0.0 when its rank falls among the
            // zeros of the `below` mass (zero-padded failures: negative keys first, then the zeros,
            // then small positives); any other target outside the bucket restarts the row.
            // ... except the row's own minimum / maximum: the first digit pass tracked the extreme keys
            // outside the window (omin / omax; with path shards on several GPUs they become global in the
            // candidate exchange, before anything reads them — not in the plain stepwise protocol)
            uint32_t fixed = 0, omin_m = 0, omax_m = 0;
            const int64_t nv = R.n_valid;
            for (int t = 0; t < nt; ++t) {
              if (nv > 0) {
                const int64_t r = R.rank[t];
                if (r < below) {
                  if (r >= below_neg && r < below_neg + below_zero) { fixed |= 1u << t; R.fixed_key[t] = kZeroKey; }
                  else if (R.fused && r == 0) { fixed |= 1u << t; omin_m |= 1u << t; }
                  else s_restart = 1;
                } else if (r >= below + part) {
                  if (R.fused && r == nv - 1) { fixed |= 1u << t; omax_m |= 1u << t; }
                  else s_restart = 1;
                }
              }
              R.rank[t] -= below;
            }
            if (s_restart) {
              fixed = omin_m = omax_m = 0;
              for (int t = 0; t < nt; ++t) { R.rank[t] += below; R.prefix[t] = 0; R.group[t] = 0; R.bcount[t] = R.n_valid; }
            } else if (fixed) {
              // a fixed target shadows a live one from here on (same bucket walk, no group of its own):
              // targets are rank-sorted, so the fixed ones sit at the two ends; all fixed: row done
              int first_live = -1, last_live = -1;
              for (int t = 0; t < nt; ++t)
                if (!((fixed >> t) & 1u)) { if (first_live < 0) first_live = t; last_live = t; }
              if (first_live < 0) {
                s_restart = 2;   // nothing left to select
              } else {
                for (int t = 0; t < nt; ++t)
                  if ((fixed >> t) & 1u) R.rank[t] = R.rank[t < first_live ? first_live : last_live];
              }
            }
            R.fixed_mask = fixed;
            R.omin_mask = omin_m;
            R.omax_mask = omax_m;
          }
        }
      }
      __syncthreads();
    }
    const bool all_fixed = s_restart == 2;
    const bool restart = s_restart == 1;
    // warp t walks target t's bucket histogram (8 bins per lane + warp scan, 256 bins at a time) to
    // the bin holding its rank
    const int w = 64 - rb < 8 ? 64 - rb : 8;
    const int n_warps = (int)(blockDim.x >> 5);
    for (int tw = warp; tw < nt && !extremes_only && !restart && !all_fixed && R.n_valid > 0; tw += n_warps) {
      const int warp = tw;   // (the target this warp walks in this round)
      const uint32_t* h = first_digit ? H : H + R.group[warp] * 256;
      const int nb = first_digit ? kFirstBins : 256;
      const int64_t r = R.rank[warp];
      int64_t running = 0;
      bool found = false;
      int blk0 = 0;
      if (first_digit) {  // the block holding the rank, from the block sums
        const int64_t mine_sum = lane < kFirstBlocks ? (int64_t)s_blocksum[lane] : 0;
        int64_t incl = mine_sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int64_t up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        const unsigned who = __ballot_sync(0xffffffffu, r >= incl - mine_sum && r < incl);
        if (who) {
          const int src = __ffs(who) - 1;
          blk0 = src * 256;
          running = __shfl_sync(0xffffffffu, incl - mine_sum, src);
        } else {
          blk0 = nb;  // beyond the histogram: falls through to the "not found" rule
        }
      }
      for (int blk = blk0; blk < nb && !found; blk += 256) {
        uint32_t c[8];
        uint32_t local = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int b = blk + lane * 8 + k;
          c[k] = b < nb ? ld_hist<HG>(h + b) : 0u;
          local += c[k];
        }
        uint32_t incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        const int64_t before = running + (int64_t)incl - local;
        const bool mine = r >= before && r < running + (int64_t)incl;
        const unsigned who = __ballot_sync(0xffffffffu, mine);
        running += (int64_t)__shfl_sync(0xffffffffu, incl, 31);
        if (who == 0) continue;
        found = true;
        if (lane == __ffs(who) - 1) {
          int64_t rr = r - before;
          int k = 0;
          for (; k < 7; ++k) {
            if (rr < (int64_t)c[k]) break;
            rr -= c[k];
          }
          R.rank[warp] = rr;
          R.bcount[warp] = c[k];
          const uint64_t bin = (uint64_t)(blk + lane * 8 + k);
          R.prefix[warp] = first_digit ? R.fbase + bin : ((R.prefix[warp] << w) | bin);
        }
      }
      if (!found && lane == 31) {  // rank beyond the bucket (cannot happen for consistent histograms): last bin
        R.rank[warp] = 0;
        R.bcount[warp] = 0;
        R.prefix[warp] = first_digit ? R.fbase + (uint64_t)(kFirstBins - 1) : ((R.prefix[warp] << w) | (uint64_t)((1u << w) - 1u));
      }
    }
    __syncthreads();
    if (tid == 0) {
      int nrb = first_digit ? 64 - fshift : rb + w;
      bool counted = true;
      if (extremes_only) {
        // the window of the first digit pass: the smallest bin width 2^s that lays at most kFirstBins
        // bins over [kmin, kmax]. Nothing seen (empty / all-NaN / all-zero sample): no window, the next
        // pass histograms the top digit from bit 0.
        nrb = 0;
        if (R.kmin <= R.kmax) {
          int sft = 0;
          while (((R.kmax >> sft) - (R.kmin >> sft)) >= (uint64_t)kFirstBins) ++sft;
          R.fshift = sft;
          R.fbase = R.kmin >> sft;
        } else {
          R.fshift = 64;
        }
        for (int t = 0; t < nt; ++t) {
          R.prefix[t] = 0ull;
          R.bcount[t] = 0;
        }
        counted = false;
      } else if (restart) {
        nrb = 0;
        R.adaptive = 0;
        R.fshift = 64;
      } else if (R.n_valid <= 0 || all_fixed) {
        nrb = 64;
      }
      R.rb = nrb;
      // groups of the next pass (targets are rank-sorted, so prefixes are sorted)
      int64_t gcount[kMaxTargets];
      int64_t total = 0;
      int g = 0;
      for (int t = 0; t < nt; ++t) {
        if (t == 0 || R.prefix[t] != R.uprefix[g - 1]) {
          gcount[g] = R.bcount[t];
          total += R.bcount[t];
          R.uprefix[g++] = R.prefix[t];
        }
        R.group[t] = g - 1;
      }
      R.n_groups = g < 1 ? 1 : g;
      // once the live buckets fit the candidate list together the row skips the remaining full
      // scans; otherwise buckets above kBigBucket are left out of the gather (a bucket of one
      // repeated value — zero-padded failed paths — never shrinks)
      // (only where a tail kernel finishes the row from the gathered list: the plain stepwise
      // protocol keeps histogramming, so it has no use for an early stop)
      const bool ready = counted && !restart && R.fused && nrb < 64 && total <= R.cap;
      for (int k = 0; k < g; ++k) R.gbig[k] = (!ready && (!counted || restart || gcount[k] > kBigBucket)) ? 1 : 0;
      R.ready = ready ? 1 : 0;
    }
  }
  for (int k = tid; k < kHistWords; k += blockDim.x) H[k] = 0;
  __syncthreads();
}

// (the per-row state is a few KB that one thread walks back and forth: staged in shared memory)
// CG: read the source through L2 (it was updated by atomics of other CTAs of the same launch)
template <bool CG>
__device__ __forceinline__ void stage_row(SelRow& dst, const SelRow& src) {
  static_assert(sizeof(SelRow) % 8 == 0, "SelRow is copied in 8-byte words");
  const unsigned long long* s = (const unsigned long long*)&src;
  unsigned long long* d = (unsigned long long*)&dst;
  for (int k = threadIdx.x; k < (int)(sizeof(SelRow) / 8); k += blockDim.x) d[k] = CG ? __ldcg(s + k) : s[k];
  __syncthreads();
}

__global__ void __launch_bounds__(1024) k_sel_advance(SelRow* __restrict__ rows, uint32_t* __restrict__ hist,
                                                      int pass) {
  __shared__ SelRow s_row;
  // finished rows and rows waiting for the collect: no CTA touched their histogram (still all zero)
  if (rows[blockIdx.x].rb >= 64 || rows[blockIdx.x].ready) return;
  stage_row<false>(s_row, rows[blockIdx.x]);
  advance_row<true>(s_row, hist + (size_t)blockIdx.x * kHistWords, pass);
  stage_row<false>(rows[blockIdx.x], s_row);
}

// "this CTA is the last one of its row's grid line to finish": every CTA of the line calls it once, after its
// global atomics; true for exactly one of them, which then sees everything the others published
__device__ __forceinline__ bool last_cta_of_row(SelRow& R) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&R.done, 1u) == gridDim.x - 1u;
  __syncthreads();
  const bool last = s_last != 0;
  if (last) __threadfence();
  return last;
}

// One pass, first half (hist_cta) — and, in the single-GPU launch sequence (`fuse`), the second half too: the
// last CTA to finish a row walks the row's merged histogram right away (advance_row), while the other rows
// are still being scanned; no separate advance launch, and its serial part stays off the critical path.
__global__ void __launch_bounds__(kHistThreads) k_sel_hist(SelRow* __restrict__ rows, uint32_t* __restrict__ hist,
                                                              const double* __restrict__ cand, int pass, int chunk,
                                                              int stride, int fuse) {
  extern __shared__ __align__(16) uint32_t sh[];  // [n_groups][256] or the kFirstWords of the first digit pass
  static_assert(sizeof(SelRow) <= sizeof(uint32_t) * kHistWords, "the staged row reuses the histogram's shared memory");
  SelRow& s_row = *reinterpret_cast<SelRow*>(sh);   // (after the scan: 6 CTAs of 32 KB + tables per SM, no room to spare)
  SelRow& R = rows[blockIdx.y];
  if (R.rb >= 64 || R.ready) return;   // the whole grid line of the row returns here: nothing to advance
  uint32_t* H = hist + (size_t)blockIdx.y * kHistWords;
  hist_cta(R, H, cand, pass, chunk, stride, sh);
  if (!fuse || !last_cta_of_row(R)) return;
  stage_row<true>(s_row, R);
  advance_row<true>(s_row, H, pass);
  if (threadIdx.x == 0) s_row.done = 0;
  __syncthreads();
  stage_row<false>(R, s_row);
}

// The gather scan (collect_cta); `fuse`: the last CTA of a row also closes its list (k_sel_collect_finish).
__global__ void __launch_bounds__(kHistThreads) k_sel_collect(SelRow* __restrict__ rows, double* __restrict__ cand,
                                                              int chunk, int fuse) {
  SelRow& R = rows[blockIdx.y];
  if (R.rb >= 64 || R.collected) return;   // the whole grid line of the row
  collect_cta(R, cand, chunk);
  if (!fuse || !last_cta_of_row(R)) return;
  if (threadIdx.x == 0) {
    collect_finish_row(R);
    R.done = 0;
  }
}

// numpy 'linear' interpolation / np.median rule for quantile `k` of a finished row
__device__ __forceinline__ double finish_value(const SelRow& R, int k) {
  const QuantileSpec& spec = R.d.spec;
  const int64_t nv = R.n_valid;
  if (nv <= 0) return CUDART_NAN;
  const int sa = R.slot[2 * k], sb = R.slot[2 * k + 1];
  auto key_at = [&](int t) -> uint64_t {
    if ((R.omin_mask >> t) & 1u) return R.omin;
    if ((R.omax_mask >> t) & 1u) return R.omax;
    return ((R.fixed_mask >> t) & 1u) ? R.fixed_key[t] : R.prefix[t];
  };
  const double a = value_of(key_at(sa));
  const double b = value_of(key_at(sb));
  if (spec.median == 2) return a;
  if (spec.median) return (nv & 1) ? a : __ddiv_rn(__dadd_rn(a, b), 2.0);
  const double q = spec.q[k];
  const double vi = __dmul_rn((double)(nv - 1), q);
  double prev = floor(vi);
  if (vi >= (double)(nv - 1)) prev = -1.0;  // numpy _get_indexes: both indexes -> last
  else if (vi < 0) prev = 0.0;
  const double t = __dsub_rn(vi, prev);     // gamma
  const double diff = __dsub_rn(b, a);      // _lerp
  double res = __dadd_rn(a, __dmul_rn(diff, t));
  if (t >= 0.5) res = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, t)));
  return res;
}

// Single-GPU tail: whatever is left of a row after the collect, in ONE launch, followed by the
// interpolation: the remaining 8-bit digits are walked by this CTA alone, histogram in shared memory —
// over the gathered candidate list (L2-resident, at most kCandCap elements: every live bucket is either in
// it or one repeated value known from its extremes), or, when a bucket of distinct values overflowed the
// list, over the full row (slow, rare).
constexpr int kTailThreads = 1024;
__global__ void __launch_bounds__(kTailThreads) k_sel_tail(SelRow* __restrict__ rows, const double* __restrict__ cand,
                                                           int cand_stride, double* __restrict__ out, int out_stride,
                                                           int64_t* __restrict__ counts,
                                                           long long* __restrict__ unresolved) {
  extern __shared__ uint32_t s_hist[];  // [kHistWords]
  __shared__ PrefixTable T;
  __shared__ int need_scan;
  __shared__ int s_one[kMaxTargets];         // target sits in an ungathered bucket of one repeated value ...
  __shared__ uint64_t s_one_key[kMaxTargets];  // ... this one
  __shared__ SelRow s_row;
  stage_row<false>(s_row, rows[blockIdx.x]);
  SelRow& R = s_row;   // (nothing reads the row state after the tail: no copy back)
  const QuantileSpec& spec = R.d.spec;
  const int tid = threadIdx.x;
  const int nt = 2 * spec.n_q;
  if (R.rb < 64 && R.n_valid > 0) {
    // which targets can be read off the gathered list / off a one-value bucket, and is a scan needed?
    if (tid == 0) need_scan = 0;
    __syncthreads();
    if (tid < nt) {
      const int g = R.group[tid];
      s_one[tid] = R.gbig[g];
      s_one_key[tid] = R.gmin[g];
      if (R.gbig[g] ? (R.gmin[g] != R.gmax[g]) : (R.overflow != 0)) atomicOr(&need_scan, 1);
    }
    __syncthreads();
    if (need_scan && unresolved) {
      // path shards on several GPUs: this rank only holds part of the row, so the row cannot be
      // finished by scanning here — report it (the caller re-runs the stepwise protocol)
      if (tid == 0) atomicAdd((unsigned long long*)unresolved, 1ull);
    } else {
      const bool from_list = !need_scan;
      const double* __restrict__ x = from_list ? cand + (size_t)blockIdx.x * cand_stride : R.d.x;
      const uint8_t* __restrict__ mask = from_list ? nullptr : R.d.mask;
      const int64_t n = from_list ? (int64_t)R.n_cand : R.d.n;
      if (tid == 0) { R.ready = 0; R.fused = 0; }  // no early stop from here on: every digit is walked
      __syncthreads();
      for (int pass = kFullPassesLong; R.rb < 64; ++pass) {
        const int cur = R.rb, ng = R.n_groups;
        const int w = 64 - cur < 8 ? 64 - cur : 8;
        for (int k = tid; k < ng * 256; k += kTailThreads) s_hist[k] = 0;
        if (cur > 0 && ng > 1) build_table(T, R, ng, cur); else __syncthreads();
        const Probe P = make_probe(R, cur, w, T, s_hist);
        for (int64_t b0 = 0; b0 < n; b0 += (1 << 30)) {
          const int cnt = (int)(n - b0 < (1 << 30) ? n - b0 : (1 << 30));
          hist_dispatch<kTailThreads>(P, x + b0, mask ? mask + b0 : nullptr, cnt, cur, ng);
        }
        __syncthreads();
        advance_row<false>(R, s_hist, pass);
      }
      // (a one-value bucket is not in the list: its targets walked empty histograms above)
      if (from_list && tid < nt && s_one[tid]) R.prefix[tid] = s_one_key[tid];
      __syncthreads();
    }
  }
  if (tid < spec.n_q) out[(int64_t)blockIdx.x * out_stride + tid] = finish_value(R, tid);
  if (tid == 0 && counts) counts[blockIdx.x] = R.n_valid;
}

// ---- candidate exchange (path shards on several GPUs) -----------------------------------------
// After the collect every rank holds the LOCAL elements of each live bucket; the buckets are
// small by then (the rows stopped scanning because the GLOBAL histogram said they fit the
// list), so instead of more all-reduced digit passes the ranks pool the candidates and each
// finishes every row with the single-GPU tail. The exchange buffer is one int64 array
//   xbuf = [ unresolved | counts[world][rows] | extremes[rows][kMaxTargets + 1][2] (groups, then the row's omin / omax) | pool[rows][kPoolCap] ]
// whose regions are combined by plain all-reduces: SUM of the counts (each rank fills only its
// own line), MIN of the encoded extremes (as k_sel_extremes), SUM of the pool (each rank writes
// its candidates, as bit patterns, at its own offset into zeros).
__host__ __device__ inline size_t xbuf_counts_at() { return 1; }
__host__ __device__ inline size_t xbuf_extremes_at(int rows, int world) { return 1 + (size_t)world * rows; }
__host__ __device__ inline size_t xbuf_pool_at(int rows, int world) {
  return xbuf_extremes_at(rows, world) + (size_t)rows * (kMaxTargets + 1) * 2;
}

__global__ void k_sel_export(SelRow* __restrict__ rows, long long* __restrict__ xbuf, int n_rows, int rank, int world) {
  const int r = blockIdx.x;
  const SelRow& R = rows[r];
  const unsigned long long top = 0x8000000000000000ull;
  if (r == 0 && threadIdx.x == 0) xbuf[0] = 0;
  for (int q = threadIdx.x; q < world; q += blockDim.x) {
    long long c = 0;
    if (q == rank && R.rb < 64) c = R.overflow ? (long long)kPoolCap + 1 : (long long)R.n_cand;
    xbuf[xbuf_counts_at() + (size_t)q * n_rows + r] = c;
  }
  long long* e = xbuf + xbuf_extremes_at(n_rows, world) + (size_t)r * (kMaxTargets + 1) * 2;
  for (int g = threadIdx.x; g < kMaxTargets; g += blockDim.x) {
    e[2 * g] = (long long)(R.gmin[g] ^ top);
    e[2 * g + 1] = (long long)((~R.gmax[g]) ^ top);
  }
  if (threadIdx.x == 0) {   // this rank's extreme keys outside the first digit window (finished rows too)
    e[2 * kMaxTargets] = (long long)(R.omin ^ top);
    e[2 * kMaxTargets + 1] = (long long)((~R.omax) ^ top);
  }
}

__global__ void k_sel_place(SelRow* __restrict__ rows, const double* __restrict__ cand, long long* __restrict__ xbuf,
                            int n_rows, int rank, int world) {
  __shared__ long long s_off, s_total;
  const int r = blockIdx.x;
  SelRow& R = rows[r];
  const unsigned long long top = 0x8000000000000000ull;
  if (threadIdx.x == 0) {   // the GLOBAL omin / omax (a requested minimum / maximum may be all a finished row waits for)
    const long long* eo = xbuf + xbuf_extremes_at(n_rows, world) + (size_t)r * (kMaxTargets + 1) * 2 + 2 * kMaxTargets;
    R.omin = (unsigned long long)eo[0] ^ top;
    R.omax = ~((unsigned long long)eo[1] ^ top);
  }
  if (R.rb >= 64) return;
  if (threadIdx.x == 0) {
    long long off = 0, total = 0;
    for (int q = 0; q < world; ++q) {
      const long long c = xbuf[xbuf_counts_at() + (size_t)q * n_rows + r];
      if (q < rank) off += c;
      total += c;
    }
    s_off = off; s_total = total;
  }
  __syncthreads();
  const long long total = s_total, off = s_off;
  const int mine = R.n_cand;
  long long* pool = xbuf + xbuf_pool_at(n_rows, world) + (size_t)r * kPoolCap;
  if (total <= kPoolCap) {
    const double* __restrict__ c = cand + (size_t)r * kCandCap;
    for (int i = threadIdx.x; i < mine; i += blockDim.x) pool[off + i] = __double_as_longlong(c[i]);
  }
  const long long* e = xbuf + xbuf_extremes_at(n_rows, world) + (size_t)r * (kMaxTargets + 1) * 2;
  __syncthreads();
  if (threadIdx.x < kMaxTargets) {
    R.gmin[threadIdx.x] = (unsigned long long)e[2 * threadIdx.x] ^ top;
    R.gmax[threadIdx.x] = ~((unsigned long long)e[2 * threadIdx.x + 1] ^ top);
  }
  if (threadIdx.x == 0) {
    R.overflow = total > kPoolCap ? 1 : 0;
    R.n_cand = total > kPoolCap ? kPoolCap : (int)total;
  }
}

__global__ void k_sel_finish(const SelRow* __restrict__ rows, double* __restrict__ out, int out_stride,
                             int64_t* __restrict__ counts) {
  const SelRow& R = rows[blockIdx.x];
  if (threadIdx.x < R.d.spec.n_q) out[(int64_t)blockIdx.x * out_stride + threadIdx.x] = finish_value(R, threadIdx.x);
  if (threadIdx.x == 0 && counts) counts[blockIdx.x] = R.n_valid;
}

__global__ void k_rates(const double* __restrict__ start, const double* __restrict__ fy_real, int64_t n,
                        double* __restrict__ rates) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double s = start[i];
  rates[i] = s > MCR_SMALL_EPSILON ? __dmul_rn(__ddiv_rn(fy_real[i], s), 100.0) : CUDART_NAN;
}

// YearsToRuin = ruin_month / 12 (IEEE division, as CPython's (r + 1) / 12 at simulation.py:825-828),
// NaN for paths that never failed
__global__ void k_years_to_ruin(const int32_t* __restrict__ ruin, int64_t n, double* __restrict__ years) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t m = ruin[i];
  years[i] = m < 0 ? CUDART_NAN : __ddiv_rn((double)m, (double)MCR_MONTHS_PER_YEAR);
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
  if (mode & MCR_HIST_RAW_RANGE) {   // extremes of the undivided values: divide like every element below
    first = __ddiv_rn(first, divisor);
    last = __ddiv_rn(last, divisor);
    mode &= ~MCR_HIST_RAW_RANGE;
  }
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

size_t quantile_state_bytes(int rows) {
  const size_t r = (size_t)(rows > 0 ? rows : 1);
  return sizeof(SelRow) * r + sizeof(double) * kCandCap * r + sizeof(RowDesc) * r;  // state, candidate lists, descriptors
}
size_t quantile_hist_bytes(int rows) { return sizeof(uint32_t) * kHistWords * (size_t)(rows > 0 ? rows : 1); }
static double* cand_of(void* state, int rows) { return (double*)((char*)state + sizeof(SelRow) * (size_t)rows); }
RowDesc* select_desc_area(void* state, int rows) {
  return (RowDesc*)((char*)state + (sizeof(SelRow) + sizeof(double) * kCandCap) * (size_t)rows);
}

cudaError_t launch_sel_begin(int rows, void* state, void* hist, cudaStream_t st, int adaptive, int fused) {
  if (rows <= 0) return cudaSuccess;
  k_sel_init<<<rows, 256, 0, st>>>((SelRow*)state, (uint32_t*)hist, select_desc_area(state, rows), adaptive, fused);
  return cudaGetLastError();
}

// elements per CTA of a full scan: 16 K for rows of up to 16 M elements, else the row cut into
// kMaxChunksPerRow pieces (a multiple of the tile of scan_elements)
static int chunk_for(int64_t max_n) {
  if (max_n <= (int64_t)kChunk * kMaxChunksPerRow) return kChunk;
  const int64_t c = (max_n + kMaxChunksPerRow - 1) / kMaxChunksPerRow;
  const int64_t tile = MCR_SCAN_UNROLL * kHistThreads;
  return (int)((c + tile - 1) / tile * tile);
}

cudaError_t launch_sel_hist(int rows, int64_t max_n, int pass, void* state, void* hist, cudaStream_t st, int sampled,
                            int fuse) {
  if (rows <= 0 || max_n <= 0) return cudaSuccess;
  int chunk = chunk_for(max_n), stride = 1;
  // the wide first digit pass of adaptive rows, and the later passes most rows skip: few, long CTAs
  if (pass >= 1 && chunk < kFirstChunk) chunk = kFirstChunk;
  unsigned grid = (unsigned)((max_n + chunk - 1) / chunk);
  if (pass == 0) {
    // adaptive calls only sample the rows for their extreme keys in pass 0: ~64 pieces of 16 K
    // elements per long row. `sampled` = the caller knows the call is adaptive (grid of sample CTAs
    // only); otherwise the piece grid is launched and the CTAs off the sample exit at once.
    const int64_t pieces = (max_n + kChunk - 1) / kChunk;
    chunk = kChunk;
    if (sampled) {
      stride = (int)(pieces / 64 > kSampleStride ? pieces / 64 : kSampleStride);
      const int64_t g = (pieces + stride - 1) / stride;
      grid = (unsigned)(g > kSampleMinChunks ? g : (pieces < kSampleMinChunks ? pieces : kSampleMinChunks));
    } else {
      stride = kSampleStride;   // non-adaptive rows ignore it (pass 0 is their first digit pass) ...
      grid = (unsigned)pieces;  // ... and need every piece
      if (max_n > (int64_t)kChunk * kMaxChunksPerRow) {  // long non-adaptive rows: plain chunk grid, stride unused
        chunk = chunk_for(max_n);
        grid = (unsigned)((max_n + chunk - 1) / chunk);
        stride = 1;
      }
    }
  }
  const size_t smem = (size_t)kHistWords * sizeof(uint32_t);
  k_sel_hist<<<dim3(grid, (unsigned)rows), kHistThreads, smem, st>>>((SelRow*)state, (uint32_t*)hist,
                                                                     cand_of(state, rows), pass, chunk, stride, fuse);
  return cudaGetLastError();
}

cudaError_t launch_sel_collect(int rows, int64_t max_n, void* state, cudaStream_t st, int fuse) {
  if (rows <= 0) return cudaSuccess;
  const int chunk = chunk_for(max_n);
  const unsigned chunks = (unsigned)((max_n + chunk - 1) / chunk);
  if (chunks > 0)
    k_sel_collect<<<dim3(chunks, (unsigned)rows), kHistThreads, 0, st>>>((SelRow*)state, cand_of(state, rows), chunk,
                                                                         fuse);
  if (!fuse || chunks == 0) k_sel_collect_finish<<<rows, 32, 0, st>>>((SelRow*)state);
  return cudaGetLastError();
}

cudaError_t launch_sel_advance(int rows, int max_nq, int pass, void* state, void* hist, cudaStream_t st) {
  if (rows > 0) k_sel_advance<<<rows, 32 * 2 * max_nq, 0, st>>>((SelRow*)state, (uint32_t*)hist, pass);
  return cudaGetLastError();
}

cudaError_t launch_sel_extremes(int rows, void* state, long long* ext, int store, cudaStream_t st) {
  if (rows > 0) k_sel_extremes<<<(rows + 127) / 128, 128, 0, st>>>((SelRow*)state, ext, rows, store);
  return cudaGetLastError();
}

cudaError_t launch_sel_finish(int rows, const void* state, double* out, int out_stride, int64_t* counts, cudaStream_t st) {
  if (rows > 0) k_sel_finish<<<rows, 32, 0, st>>>((const SelRow*)state, out, out_stride, counts);
  return cudaGetLastError();
}

cudaError_t launch_quantiles_rows(int rows, const RowDesc* desc_host, double* out, int out_stride, int64_t* counts,
                                  void* state, void* hist, cudaStream_t st, int* n_launches) {
  *n_launches = 0;
  if (rows <= 0) return cudaSuccess;
  int64_t max_n = 0;
  int max_nq = 1;
  for (int r = 0; r < rows; ++r) {
    max_n = desc_host[r].n > max_n ? desc_host[r].n : max_n;
    max_nq = desc_host[r].spec.n_q > max_nq ? desc_host[r].spec.n_q : max_nq;
  }
  cudaError_t e = launch_sel_begin(rows, state, hist, st, /*adaptive=*/1, /*fused=*/1);
  ++*n_launches;
  // hist passes with the advance, and the collect with its finish, fused into the scan kernels (their last
  // CTA per row): init + (sample, first digit, [digit ...]) + collect + tail launches in all
  const int full = select_full_passes_for(max_n);
  for (int pass = 0; pass < full && e == cudaSuccess; ++pass) {
    if (max_n > 0) {
      e = launch_sel_hist(rows, max_n, pass, state, hist, st, /*sampled=*/1, /*fuse=*/1);
      ++*n_launches;
    } else {
      e = launch_sel_advance(rows, max_nq, pass, state, hist, st);   // empty rows: nothing to scan
      ++*n_launches;
    }
  }
  if (e == cudaSuccess) e = launch_sel_collect(rows, max_n, state, st, /*fuse=*/1);
  ++*n_launches;
  if (e != cudaSuccess) return e;
  // one launch for the remaining digits + interpolation (see k_sel_tail)
  k_sel_tail<<<rows, kTailThreads, sizeof(uint32_t) * kHistWords, st>>>((SelRow*)state, cand_of(state, rows), kCandCap, out,
                                                                        out_stride, counts, nullptr);
  ++*n_launches;
  return cudaGetLastError();
}

int select_full_passes() { return kFullPasses; }
int select_full_passes_for(int64_t n_global_max) { return n_global_max > kLongRow ? kFullPassesLong : kFullPasses; }

size_t select_exchange_words(int rows, int world) {
  rows = rows > 0 ? rows : 1;
  world = world > 0 ? world : 1;
  return xbuf_pool_at(rows, world) + (size_t)rows * kPoolCap;
}
void select_exchange_layout(int rows, int world, int64_t at[4]) {
  at[0] = (int64_t)xbuf_counts_at();
  at[1] = (int64_t)xbuf_extremes_at(rows, world);
  at[2] = (int64_t)xbuf_pool_at(rows, world);
  at[3] = (int64_t)select_exchange_words(rows, world);
}

cudaError_t launch_sel_export(int rows, void* state, long long* xbuf, int rank, int world, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(xbuf + xbuf_pool_at(rows, world), 0, sizeof(long long) * (size_t)rows * kPoolCap, st);
  if (e != cudaSuccess) return e;
  k_sel_export<<<rows, 64, 0, st>>>((SelRow*)state, xbuf, rows, rank, world);
  return cudaGetLastError();
}

cudaError_t launch_sel_place(int rows, void* state, long long* xbuf, int rank, int world, cudaStream_t st) {
  if (rows > 0) k_sel_place<<<rows, 256, 0, st>>>((SelRow*)state, cand_of(state, rows), xbuf, rows, rank, world);
  return cudaGetLastError();
}

cudaError_t launch_sel_tail_pooled(int rows, void* state, long long* xbuf, int world, double* out, int out_stride,
                                   int64_t* counts, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  k_sel_tail<<<rows, kTailThreads, sizeof(uint32_t) * kHistWords, st>>>(
      (SelRow*)state, (const double*)(xbuf + xbuf_pool_at(rows, world)), kPoolCap, out, out_stride, counts, xbuf);
  return cudaGetLastError();
}

cudaError_t launch_rates(const double* start, const double* fy_real, int64_t n, double* rates, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_rates<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(start, fy_real, n, rates);
  return cudaGetLastError();
}

cudaError_t launch_years_to_ruin(const int32_t* ruin, int64_t n, double* years, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_years_to_ruin<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ruin, n, years);
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
