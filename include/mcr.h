/*
 * mcr.h — C ABI of the B200-native Monte Carlo retirement path engine (libmcr_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of rflamino/monte_carlo_retirement: the
 * per-path timeline engine, its batch aggregations and the working-months search of
 * `backend/simulation.py`. The reference has no FFI today (it is pure Python), so every entry
 * point below cites the reference interface it replaces (file:line relative to the reference
 * root). The host-side mirror that binds these with ctypes is
 * `monte_carlo_retirement_b200/simulation.py`; INTEGRATION.md shows the binding a reference
 * maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative MCR_E* code otherwise; the message is
 *     available through mcr_last_error(ctx) (thread-local when ctx == NULL);
 *   - no exceptions, no torch types: plain pointers and sizes. Pointers named *_dev are DEVICE
 *     pointers owned by the caller (the Python side gets them from torch.Tensor.data_ptr());
 *     pointers named *_host are host pointers. The library never frees caller memory and never
 *     returns memory it allocated, only fills caller buffers;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream); all work
 *     is enqueued on it and the call returns without synchronising unless it fills a *_host
 *     output;
 *   - one mcr_ctx per simulator instance; a context is bound to one CUDA device and may be used
 *     from any host thread, one call at a time (calls on one context are serialised
 *     internally). There is no global mutable state, so concurrent contexts are safe
 *     (reference callers: worker threads of `backend/server.py:309,405`).
 *   - there is NO CPU fallback: every compute entry point fails with MCR_ECUDA when no
 *     sm_100-class device is usable.
 */
#ifndef MCR_H_
#define MCR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCR_ABI_VERSION 1
#define MCR_MAX_STREAMS 16      /* other_income_streams is unbounded in the reference
                                   (backend/config.py:99); more than this -> MCR_EINVAL */
#define MCR_MONTHS_PER_YEAR 12  /* backend/constants.py:1 */
#define MCR_SMALL_EPSILON 1e-6  /* backend/constants.py:3 */

/* error codes */
#define MCR_OK 0
#define MCR_EINVAL (-1)   /* bad argument: Python shim raises ValueError   */
#define MCR_ECUDA (-2)    /* CUDA runtime / launch error: RuntimeError     */
#define MCR_ENOMEM (-3)   /* scratch allocation failed: RuntimeError       */

/* flags for mcr_simulate / mcr_replay / mcr_search_batch */
#define MCR_FLAG_STRICT 0x1u /* reference operation order, no FMA contraction, no reciprocal
                                sharing: the parity build. Replay and single-path runs always
                                use it. Without it the fast arithmetic is used (same
                                formulas, FMA contraction + shared reciprocals, <=1e-12 rel). */

#define MCR_FLAG_SMALL_RETURNS 0x2u /* mcr_replay without MCR_FLAG_STRICT only: the caller has
                                verified on the host that |mu_log/12 + sigma_log/sqrt(12) * z| <
                                mcr_small_returns_bound(ctx) for every supplied shock of all three
                                factors — the bound mcr_create proved for this scenario's OWN
                                Philox draws. The fast build then runs the SAME kernel variant as
                                native-draw launches (short exp polynomial, lean month steps) on
                                the supplied draws, which is how the tests hold the benchmarked
                                variant to the <= 1e-9 replay contract. Ignored when the bound is
                                0; behaviour is undefined if the guarantee does not hold. */

/* seed streams — backend/simulation.py:147-151,177-185 (search vs final SeedSequence children) */
#define MCR_STREAM_SEARCH 0
#define MCR_STREAM_FINAL 1

/* One `other_income_streams` entry — backend/config.py:12-47 (OtherIncomeStreamConfig). */
typedef struct mcr_income_stream {
  double monthly_amount_today;
  double start_at_age;
  double tax_rate;
  int32_t duration_years;    /* -1 == None (paid indefinitely)                          */
  int32_t inflation_indexed; /* 0/1                                                      */
} mcr_income_stream;

/*
 * Flattened scenario — backend/config.py:48-126 (Config) plus the three lognormal parameter
 * pairs the simulator derives at construction (backend/simulation.py:156-166,
 * arithmetic_to_log_params :14-29). The log parameters are computed by the HOST caller with the
 * reference's own formula so that the constants the kernel sees equal the reference's bit for
 * bit (log/sqrt are not re-evaluated on the device).
 */
typedef struct mcr_params {
  double initial_balance;
  double monthly_contribution;
  double contribution_growth_rate_annual;
  double monthly_expenses;
  double current_age;
  double allocation_inv1_pct;
  double inv1_mu_log, inv1_sigma_log; /* equity                      */
  double inf_mu_log, inf_sigma_log;   /* inflation                   */
  double prem_mu_log, prem_sigma_log; /* inv2 premium over inflation */
  double equity_inflation_rho;
  double inv1_annual_tax_on_gains_rate;
  double inv1_realized_gains_tax_rate;
  double inv2_annual_tax_on_gains_rate;
  double inv2_realized_gains_tax_rate;
  int32_t inv1_use_realized_gains_tax_system;
  int32_t inv2_use_realized_gains_tax_system;
  int32_t retirement_years;
  int32_t n_streams;
  mcr_income_stream streams[MCR_MAX_STREAMS];
} mcr_params;

/*
 * Per-path outputs of one batch — replaces the list of 10-key dicts that
 * `_run_single_simulation_path` returns (backend/simulation.py:939-950) and that
 * `run_monte_carlo_simulations` turns into `summary_df` and the three series frames
 * (:1012-1038,1099-1103). All pointers are DEVICE pointers, any of them may be NULL (that
 * output is then not produced). Column layout is structure-of-arrays; the series are
 * time-major so that a warp's 32 paths store 256 contiguous bytes:
 *   trajectory[t * ld + i]   t in [0, T),  T = 1 + ceil(working_months/12) + retirement_years
 *   wr[y * ld + i]           y in [0, retirement_years)
 * with i the path's index inside this batch and ld >= n_paths (`series_ld`).
 */
typedef struct mcr_outputs {
  double* start_balance;      /* "Start Balance"                         */
  double* final_balance;      /* "Final Balance" (clamped >= 0)          */
  uint8_t* success;           /* "Success" 0/1                           */
  int32_t* ruin_month;        /* "YearsToRuin" * 12, -1 == NaN (success) */
  double* first_year_gross;   /* "First Year Gross Withdrawal"           */
  double* first_year_real;    /* "First Year Real Gross Withdrawal"      */
  double* inflation_at_ret;   /* "Inflation At Retirement"               */
  double* trajectory;         /* "Trajectory"      [T][ld]               */
  double* real_trajectory;    /* "RealTrajectory"  [T][ld]               */
  double* wr_trajectory;      /* "WithdrawalRateTrajectory" [R][ld], NaN = no observation */
  int64_t series_ld;          /* leading dimension of the three series   */
  /* device-side reductions (each may be NULL) */
  int64_t* success_count;     /* [1]  += number of successful paths (a15, :1130-1136)      */
  int64_t* wr_obs_count;      /* [R]  += non-NaN WR observations per year (:1111-1113)     */
  int64_t* ruin_month_hist;   /* [12R+1] += failed paths per ruin month (server.py:525-532)*/
  uint64_t* executed_months;  /* [1]  += months actually stepped (roofline accounting)     */
} mcr_outputs;

/* Host record of one path — the scalar part of the dict at backend/simulation.py:939-950. */
typedef struct mcr_path_record {
  double start_balance;
  double final_balance;
  double first_year_gross;
  double first_year_real;
  double inflation_at_ret;
  int32_t success;
  int32_t ruin_month; /* -1 == NaN */
  int32_t trajectory_len;
  int32_t wr_len;
} mcr_path_record;

typedef struct mcr_ctx mcr_ctx;

/* ---- lifetime ------------------------------------------------------------------------- */

/* Replaces RetirementMonteCarloSimulator.__init__ (backend/simulation.py:135-175): validates
 * and freezes the scenario, derives the Philox key from main_seed, binds to `device`. */
int mcr_create(const mcr_params* params, uint64_t main_seed, int device, mcr_ctx** out_ctx);
int mcr_destroy(mcr_ctx* ctx);
const char* mcr_last_error(const mcr_ctx* ctx);
int mcr_abi_version(void);
/* 0.05 / 0.1: every monthly log-return of this scenario's native draws is provably below it
 * (Box-Muller normals of 32-bit uniforms are bounded by 6.77); 0: no such bound. */
double mcr_small_returns_bound(const mcr_ctx* ctx);

/* ---- host-side pure helpers (a2/a3 of SURVEY §8) -------------------------------------- */

/* stream_payment_start_month_index (backend/simulation.py:47-63). */
int32_t mcr_stream_start_month(double current_age, int32_t working_months, double start_at_age);
/* Number of yearly trajectory points, 1 + ceil(wm/12) + R (backend/simulation.py:585-589,902). */
int32_t mcr_trajectory_len(int32_t working_months, int32_t retirement_years);

/* ---- the timeline kernel -------------------------------------------------------------- */

/* Native-RNG batch: replaces the fan-out of run_monte_carlo_simulations
 * (backend/simulation.py:973-1010) — `_path_seeds` (:187-199) + n x
 * `_run_single_simulation_path` (:476-950) with `_draw_shock_path` (:452-466) replaced by
 * counter-based Philox4x32-10 keyed by (main_seed, seed_stream) and counted by
 * (first_path + i, absolute month). Path i of this call is GLOBAL path first_path + i, so
 * shards of one run on several GPUs draw disjoint subsequences and the result of a path does
 * not depend on the shard layout or on working_months (common random numbers). */
int mcr_simulate(mcr_ctx* ctx, int seed_stream, int32_t working_months, int64_t first_path,
                 int64_t n_paths, uint32_t flags, const mcr_outputs* out, void* stream);

/* Replay batch: same kernel body fed with precomputed correlated shocks (the reference's own
 * numpy draws, backend/simulation.py:452-466), device layout shocks_dev[(m*3 + c) * shocks_ld
 * + i], m in [0, n_months), c in {equity, inflation, premium}. n_months must be
 * >= max(working_months + 12R, 1). Pass MCR_FLAG_STRICT for the parity contract (the Python
 * mirror always does); flags == 0 runs the fast arithmetic on the same inputs, which is how the
 * tests bound the fast-vs-strict drift. */
int mcr_replay(mcr_ctx* ctx, const double* shocks_dev, int64_t shocks_ld, int32_t n_months,
               int32_t working_months, int64_t n_paths, uint32_t flags, const mcr_outputs* out,
               void* stream);

/* One path from HOST shocks[n_months][3] (row-major, as `_draw_shock_path` returns) to HOST
 * outputs: replaces a direct call of `_run_single_simulation_path`
 * (backend/simulation.py:476-950; 11 reference tests call it). Runs the strict kernel with one
 * thread on the device and synchronises. trajectory/real_trajectory need trajectory_len
 * doubles, wr needs retirement_years doubles (may be NULL). */
int mcr_single_path(mcr_ctx* ctx, int32_t working_months, const double* shocks_host,
                    int32_t n_months, mcr_path_record* record_host, double* trajectory_host,
                    double* real_trajectory_host, double* wr_host);

/* Device evaluation of the private helpers of the path (backend/simulation.py:201-254,
 * :256-272, :274-359 — the three the reference tests call directly — and :361-450). One strict
 * thread each. */
int mcr_helper_withdraw(mcr_ctx* ctx, double bal, double cost_basis, double net_target,
                        int32_t use_real_tax, double real_tax_rate, double out4_host[4]);
int mcr_helper_net_liquidation(mcr_ctx* ctx, double bal, double cost_basis, int32_t use_real_tax,
                               double real_tax_rate, double* out_host);
int mcr_helper_rebalance(mcr_ctx* ctx, double bal1, double cb1, double bal2, double cb2,
                         double out4_host[4]);
/* `_apply_annual_gain_taxes` (backend/simulation.py:361-450): out5 = {bal1, cb1, bal2, cb2,
 * tax_failed (0.0 / 1.0)} after the period's mark-to-market tax and the trailing rebalance. */
int mcr_helper_annual_tax(mcr_ctx* ctx, double bal1, double cb1, double bal2, double cb2,
                          double gain1, double gain2, double out5_host[5]);

/* Native shocks in the replay layout: shocks_dev[(m*3 + c) * shocks_ld + i] for global paths
 * first_path + i — the device analogue of `_draw_shock_path` (backend/simulation.py:452-466).
 * Feeding them to mcr_replay (or to the CPU oracle) reproduces mcr_simulate with
 * MCR_FLAG_STRICT path for path. */
int mcr_draw_shocks(mcr_ctx* ctx, int seed_stream, int64_t first_path, int64_t n_paths,
                    int32_t n_months, uint32_t flags, double* shocks_dev, int64_t shocks_ld,
                    void* stream);

/* ---- the batched search kernel --------------------------------------------------------- */

/* Evaluates n_candidates values of working_months in ONE launch on the same Philox streams
 * (common random numbers), replacing one sequential run_monte_carlo_simulations per probe of
 * find_minimum_working_months (backend/simulation.py:1180-1222). success_counts_dev[c] and
 * executed_months_dev[c] (may be NULL) are ACCUMULATED (+=), so shards can share a buffer;
 * zero them first. */
int mcr_search_batch(mcr_ctx* ctx, int seed_stream, const int32_t* candidates_host,
                     int32_t n_candidates, int64_t first_path, int64_t n_paths, uint32_t flags,
                     int64_t* success_counts_dev, uint64_t* executed_months_dev, void* stream);

/* Multi-scenario batching (no reference counterpart — the reference runs one
 * scenario per process): n_items (scenario, working_months) pairs evaluated on the SAME Philox
 * streams of this context (common random numbers across scenarios: a sensitivity grid's points
 * differ by their parameters, not by their luck) in one launch per kernel variant present.
 * scenarios_host[k] is flattened like the argument of mcr_create (the context's own scenario
 * plays no role; its seed and device do). success_counts_dev[k] / executed_months_dev[k] (may be
 * NULL) are ACCUMULATED (+=). */
int mcr_sweep_batch(mcr_ctx* ctx, int seed_stream, const mcr_params* scenarios_host,
                    const int32_t* working_months_host, int32_t n_items, int64_t first_path,
                    int64_t n_paths, uint32_t flags, int64_t* success_counts_dev,
                    uint64_t* executed_months_dev, void* stream);

/* ---- device aggregations (a14/a16/a19 of SURVEY §8) ------------------------------------ */

#define MCR_SEL_MEDIAN 0x1u /* np.median rule: mean of the two middle order statistics
                               (Series.median at simulation.py:96, server.py:449-450);
                               q_host is ignored and n_q results per row are all the median */
#define MCR_SEL_MINMAX 0x2u /* the row's minimum and maximum over the valid elements: n_q must be 2, q ignored;
                              * out[0] = min, out[1] = max (NaN for an empty row) — exact elements, no lerp */

/* Exact order-statistic quantiles of `rows` rows of `n` doubles each (values_dev + r * ld) with
 * numpy's 'linear' interpolation, replacing DataFrame.quantile(q, axis=1) at
 * backend/simulation.py:1059-1061,1091-1093,1108-1110 and Series.quantile/median at
 * backend/simulation.py:96, backend/server.py:449-455. NaN elements never take part (pandas
 * skips them: the WR bands of :1106-1110); when mask_dev != NULL only elements with
 * mask_dev[i] != 0 take part (one mask shared by all rows: the successful-path cohort).
 * out_dev[r * n_q + k] is the k-th quantile of row r (NaN when the row has no valid element);
 * counts_dev[r] (may be NULL) receives the number of valid elements (the WR observation
 * counts of :1111-1113). q_host holds n_q <= 16 fractions in [0,1]. */
int mcr_quantiles(mcr_ctx* ctx, const double* values_dev, int64_t n, int64_t ld, int32_t rows,
                  const uint8_t* mask_dev, const double* q_host, int32_t n_q, uint32_t sel_flags,
                  double* out_dev, int64_t* counts_dev, void* stream);

#define MCR_MAX_QUANTILES 16

/* One row of a multi-row select: n doubles at values_dev, optional cohort mask, n_q <= 16
 * ascending fractions (ignored with MCR_SEL_MEDIAN). Rows of one call may differ in all of
 * these, so a whole batch's aggregations (medians of summary columns over different cohorts,
 * final-balance quantiles, nominal / real / withdrawal-rate bands) run as ONE launch sequence. */
typedef struct mcr_select_row {
  const double* values_dev;
  const uint8_t* mask_dev;
  int64_t n;
  int32_t n_q;
  uint32_t flags; /* MCR_SEL_MEDIAN | MCR_SEL_MINMAX */
  double q[MCR_MAX_QUANTILES];
} mcr_select_row;

/* out_dev[r * 16 + k] = k-th quantile of row r; counts_dev[r] (may be NULL) = valid elements. */
int mcr_quantiles_rows(mcr_ctx* ctx, const mcr_select_row* rows_host, int32_t n_rows,
                       double* out_dev, int64_t* counts_dev, void* stream);

/* The same select one step at a time, for path shards spread over several GPUs:
 *   BEGIN; for pass in 0..7 (0..9 with the EXTREMES exchange below) {
 *   if pass == mcr_select_full_passes(): COLLECT ;
 *   HIST (local shard) ; all-reduce(sum) hist_dev across ranks ; ADVANCE } ; FINISH.
 * After the all-reduce every rank holds the GLOBAL digit histogram, so all ranks walk to the
 * same exact global order statistics without moving any path data (COLLECT gathers the local
 * elements that share a resolved prefix into a short list so the remaining digits do not
 * rescan the rows). rows_host describes this rank's shard of every row (same order, n_q and
 * flags on all ranks) and must be passed to every step; state_dev / hist_dev are caller-owned
 * device buffers of mcr_select_state_bytes(n_rows) / mcr_select_hist_bytes(n_rows) bytes;
 * hist_dev is an array of uint32 counts (all-reduce it as int32). Output layout as
 * mcr_quantiles_rows. */
#define MCR_SELECT_BEGIN 0
#define MCR_SELECT_HIST 1
#define MCR_SELECT_ADVANCE 2
#define MCR_SELECT_FINISH 3
#define MCR_SELECT_COLLECT 4
/* Optional (BEGIN with pass = 1): after HIST of pass 0, EXTREMES_GET writes every row's local
 * (min, max) key into out_dev as int64[n_rows][2] in an encoding whose element-wise MIN over
 * ranks is the global pair; all-reduce(MIN) it and hand it back with EXTREMES_SET before
 * ADVANCE 0. Pass 0 then only looks at (a sample of) every row for its extreme keys, and pass 1
 * histograms ~8 K equal bins laid over that key range (about 13 bits resolved by one scan,
 * however concentrated the row); keys outside the range are accounted for exactly by the same
 * pass, which also counts the row. */
#define MCR_SELECT_EXTREMES_GET 5
#define MCR_SELECT_EXTREMES_SET 6
/* Pooled tail (BEGIN with pass = 3: adaptive start + pooled tail). Rows stop scanning as soon as
 * the GLOBAL histogram says their live buckets fit one candidate list; after
 *   for pass in 0..mcr_select_full_passes()-1 { HIST ; [EXTREMES_*] ; all-reduce ; ADVANCE } ; COLLECT
 * the ranks pool their few local candidates and every rank finishes all rows at once:
 *   POOL_EXPORT ; all-reduce(SUM) counts region ; all-reduce(MIN) extremes region ;
 *   POOL_PLACE  ; all-reduce(SUM) pool region   ; POOL_TAIL.
 * For these three steps hist_dev is the exchange buffer — int64[mcr_select_exchange_words(n_rows,
 * world)], regions at the word offsets mcr_select_exchange_layout() returns in at4 =
 * {counts, extremes, pool, end} — and pass = rank | world << 8. POOL_TAIL writes the quantiles
 * (layout as mcr_quantiles_rows) and leaves in word 0 of the buffer the number of rows it could
 * NOT finish from the pool (a big bucket of distinct values; identical on all ranks): if that
 * is not 0 the caller runs the plain stepwise protocol above instead. */
#define MCR_SELECT_POOL_EXPORT 7
#define MCR_SELECT_POOL_PLACE 8
#define MCR_SELECT_POOL_TAIL 9
int64_t mcr_select_exchange_words(int32_t rows, int32_t world);
void mcr_select_exchange_layout(int32_t rows, int32_t world, int64_t* at4);
int32_t mcr_select_full_passes(void);
/* the same for rows of up to n_global_max elements over all ranks: rows above 2^24 elements get
 * one more digit pass before the collect, so that their buckets still fit the candidate lists */
int32_t mcr_select_full_passes_for(int64_t n_global_max);
/* HIST of pass 0 of a call begun adaptively: OR this into `pass` so that only the CTAs of the
 * sample are launched (without it the whole piece grid is launched and exits off the sample). */
#define MCR_SELECT_HIST_SAMPLED 0x100
int64_t mcr_select_state_bytes(int32_t rows);
int64_t mcr_select_hist_bytes(int32_t rows);
int mcr_select_step(mcr_ctx* ctx, int32_t step, int32_t pass, const mcr_select_row* rows_host,
                    int32_t n_rows, void* state_dev, void* hist_dev, double* out_dev,
                    int64_t* counts_dev, void* stream);

/* rates[i] = first_year_real[i] / start[i] * 100 where start[i] > 1e-6 else NaN
 * (median_first_year_withdrawal_rate, backend/simulation.py:78-96). */
int mcr_first_year_rates(mcr_ctx* ctx, const double* start_dev, const double* first_year_real_dev,
                         int64_t n, double* rates_dev, void* stream);

/* years[i] = ruin_month[i] / 12.0 (IEEE division), NaN where ruin_month[i] < 0: the "YearsToRuin"
 * column of summary_df (backend/simulation.py:825-828,943,1018) filled on the device. */
int mcr_years_to_ruin(mcr_ctx* ctx, const int32_t* ruin_month_dev, int64_t n, double* years_dev,
                      void* stream);

#define MCR_HIST_NUMPY 0 /* numpy.histogram / matplotlib plt.hist(bins=n) — plotting.py:46-59  */
#define MCR_HIST_FLOOR 1 /* idx = min(floor((v-min)/width), n-1) — HistogramChart.jsx:13-60   */
#define MCR_HIST_RAW_RANGE 0x100 /* OR into mode: range_dev holds min / max of the UNDIVIDED values (e.g. straight out
                                  * of a MCR_SEL_MINMAX select row); the kernel divides them by `divisor` itself — the
                                  * same IEEE division it applies to every element */

/* min/max of (values / divisor) over the cohort mask (all when NULL) into minmax_dev[2]
 * (NaN, NaN when the cohort is empty). */
int mcr_minmax(mcr_ctx* ctx, const double* values_dev, const uint8_t* mask_dev, int64_t n,
               double divisor, double* minmax_dev, void* stream);

/* Equal-width histogram of (values / divisor) over the cohort: n_bins bins over
 * range_dev[0..1] (from mcr_minmax, or agreed across GPUs first). plotting.py divides by 1e6
 * ($M) and uses 100 bins with numpy semantics; the dashboard uses 60 bins with floor binning.
 * hist_dev[n_bins] is ACCUMULATED (+=). */
int mcr_histogram(mcr_ctx* ctx, const double* values_dev, const uint8_t* mask_dev, int64_t n,
                  double divisor, int32_t n_bins, int32_t mode, const double* range_dev,
                  int64_t* hist_dev, void* stream);

/* Gather `n_cols` columns (path indices cols_host) of a time-major series into
 * out_dev[k * rows + t] — the 5 sample paths of backend/simulation.py:1063-1078. */
int mcr_gather_columns(mcr_ctx* ctx, const double* series_dev, int64_t ld, int32_t rows,
                       const int64_t* cols_host, int32_t n_cols, double* out_dev, void* stream);

/* ---- several GPUs of one process ------------------------------------------------ */

/* The reference is driven by a CLI and by FastAPI worker threads (backend/main.py:68-106,
 * backend/server.py:309,405), so the multi-GPU form its callers can reach is ONE process with one
 * mcr_ctx per device (one host thread each). Path shards never exchange per-path data; the few
 * small reductions (success counts, digit histograms and pooled candidates of the distributed
 * select, histograms) are all-reduced by hand-written kernels over NVLink / NVSwitch peer memory:
 * one mcr_comm per context, connected to its peers once. All ranks call mcr_comm_all_reduce with
 * the same op and element count in the same order; the call only enqueues a kernel on `stream`
 * and the result replaces buf_dev on every rank. n * width must not exceed max_bytes. */
#define MCR_COMM_SUM_I32 0
#define MCR_COMM_SUM_I64 1
#define MCR_COMM_MIN_I64 2
#define MCR_COMM_MAX_I64 3
#define MCR_COMM_SUM_F64 4
#define MCR_COMM_MIN_F64 5
#define MCR_COMM_MAX_F64 6
typedef struct mcr_comm mcr_comm;
int mcr_comm_create(mcr_ctx* ctx, int64_t max_bytes, mcr_comm** out_comm);
/* comms[world]: every rank's communicator (same process), comms[rank] == comm; enables peer access. */
int mcr_comm_connect(mcr_comm* comm, int32_t rank, int32_t world, mcr_comm* const* comms);
int mcr_comm_all_reduce(mcr_comm* comm, int32_t op, void* buf_dev, int64_t n, void* stream);
/* The pooled distributed select (BEGIN ... POOL_TAIL above) in ONE call per rank, with the
 * all-reduces of `comm` between its steps; rows_host describes this rank's shards.
 * unresolved_dev[0] receives the number of rows the pool could not finish (identical on all
 * ranks; non-zero: repeat with the stepwise protocol). */
int mcr_quantiles_rows_comm(mcr_ctx* ctx, mcr_comm* comm, int32_t rank, int32_t world,
                            const mcr_select_row* rows_host, int32_t n_rows, double* out_dev,
                            int64_t* counts_dev, int64_t* unresolved_dev, void* stream);
/* non-zero: a kernel of this rank gave up waiting for a peer (~2 s); the number of that call */
int32_t mcr_comm_status(const mcr_comm* comm);
int64_t mcr_comm_calls(const mcr_comm* comm);
const char* mcr_comm_last_error(const mcr_comm* comm);
int mcr_comm_destroy(mcr_comm* comm);

/* ---- measurement ----------------------------------------------------------------------- */

/* DFMA-chain microbenchmark: FP64-pipe issue slots (lane-instructions) per second on this
 * device — the roofline denominator of SURVEY §8d. Synchronises. */
int mcr_fp64_peak_slots_per_s(mcr_ctx* ctx, double* slots_per_s_host);

/* Which compile-time specialisation the last mcr_simulate / mcr_replay / mcr_search_batch of this
 * context launched: 0 generic, 1 both assets taxed on realized gains, 2 no tax bites; +2 / +4 for
 * the fast build's bounded-return variants (< 0.1 / < 0.05: short exp polynomial + lean month
 * steps). -1 before the first launch. Lets the tests assert that the variant bench.py times is the
 * one their parity checks ran. */
int32_t mcr_last_variant(const mcr_ctx* ctx);

/* Kernels launched by this context so far (the `gpu_launches` claim of bench.py). */
int64_t mcr_launch_count(const mcr_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* MCR_H_ */
