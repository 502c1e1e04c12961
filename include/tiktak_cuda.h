/* tiktak_cuda.h -- C ABI of libtiktak_cuda, the B200 (sm_100a) implementation of TikTak's
 * data-parallel hot path: the Sobol pre-test stage and the batch of warm-start local searches.
 *
 * Every entry point replaces the body of one stage routine of the reference driver
 * (/root/reference/TiktakGlobalSearch.f90, internal procedures of PROGRAM TiktakGlobalSearch).
 * The file:line each one stands in for is cited on the declaration.  A Fortran driver reaches
 * these through ISO_C_BINDING (see INTEGRATION.md for the interface block); the in-tree host
 * mirror is tiktak_b200/ (ctypes) and driver/TiktakGlobalSearch.cpp.
 *
 * Conventions
 *   - plain C types only: REAL(DP) -> double, INTEGER(I4B) -> int32_t (nrtype.f90:5-7);
 *     64-bit integers only for Sobol point indices and counts.
 *   - arrays that mirror Fortran arrays are COLUMN-MAJOR with the reference's shapes, so the
 *     Fortran side passes them without transposition: p_range(nx,2), p_bound(nx,2)
 *     (genericParams.f90:353-360), x_starts(p_maxpoints, nx+1) (TiktakGlobalSearch.f90:851).
 *   - every output buffer is caller-allocated.  A pointer may be host memory OR device memory
 *     of the handle's GPU (detected with cudaPointerGetAttributes) -- device pointers let a
 *     multi-GPU host exchange rows with NCCL without staging.
 *   - the handle owns all device memory; calls block the calling thread; a handle is not
 *     re-entrant.  The library never exits the process: the reference's `stop 0` / exitState
 *     (stateControl.f90:203-216) become negative return codes, see tiktak_cuda_strerror().
 *   - point / restart numbering is 1-based like the reference (whichPoint, k).
 */
#ifndef TIKTAK_CUDA_H
#define TIKTAK_CUDA_H

#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tiktak_handle tiktak_handle;

/* status codes */
enum {
    TIKTAK_OK = 0,
    TIKTAK_ERR_ARG = -1,       /* bad argument / inconsistent config (reference: `stop 0` in parseConfig) */
    TIKTAK_ERR_CUDA = -2,      /* CUDA runtime failure; message in tiktak_cuda_last_error() */
    TIKTAK_ERR_STATE = -3,     /* stage called out of order (e.g. choose before pretest) */
    TIKTAK_ERR_NOGPU = -4,     /* no CUDA device: the library has NO CPU fallback */
    TIKTAK_ERR_UNSUPPORTED = -5
};

/* built-in objective plugins (device code in tiktak_b200/csrc/objectives.cuh; contract in
 * include/tiktak_obj.cuh).  0 restates the shipped objective.f90:76-142 (sleep removed). */
enum {
    TIKTAK_OBJ_TEMPLATE = 0,   /* Griewank/200 template shipped in objective.f90:96-117 */
    TIKTAK_OBJ_GRIEWANK = 1,   /* 1 + sum x^2/4000 - prod cos(x_i/sqrt(i)) */
    TIKTAK_OBJ_RASTRIGIN = 2,  /* 10 n + sum (x^2 - 10 cos(2 pi x)) */
    TIKTAK_OBJ_ROSENBROCK = 3, /* sum 100 (x_{i+1}-x_i^2)^2 + (1-x_i)^2 */
    TIKTAK_OBJ_SMM = 4         /* synthetic simulated-method-of-moments panel, 1200 moments */
};

/* parameters of TIKTAK_OBJ_SMM (pass through tiktak_config.user): a synthetic simulated-method-of-
 * moments problem of the size the reference's README quotes its scaling on (README.md:41-43: "1200+
 * moments and 7 parameters").  Income process y_it = a_i + z_it + e_it, z_it = rho z_i,t-1 + eta_it,
 * eta a two-component normal mixture; theta = (rho, sig_alpha, sig_eps, sig_eta1, sig_eta2, p_mix, mu_eta1).
 * Panel of np persons x T ages simulated from common random numbers (Philox4x32-10, `seed`) on every
 * evaluation; moments = for each age t and each of the 6 series {y_t, y_{t+k}-y_t, k=1,2,3,5,10} the 5 raw
 * statistics E s, E s^2, E s^3, E s^4, E|s|  ->  nm = 30 T (1200 at T = 40); data moments simulated at
 * theta_star with seed+1.  nx must be 7 and nm must be 30*T. */
typedef struct tiktak_smm_params {
    int32_t np;            /* persons, multiple of 256 */
    int32_t T;             /* ages, <= 40 */
    int32_t seed;          /* p_SEED = 314159265 in the reference (genericParams.f90:24) */
    int32_t reserved;
    double theta_star[7];
} tiktak_smm_params;

/* local-search algorithm ids = the reference's `alg` (genericParams.f90:12-13) */
enum { TIKTAK_ALG_DFNLS = 0, TIKTAK_ALG_AMOEBA = 1, TIKTAK_ALG_DFPMIN = 2 };

/* Mirrors the module variables of genericParams.f90:14-21 after parseConfig (:265-450), i.e.
 * AFTER defaults were applied (maxeval=40(nx+1), legitimate=qr_ndraw+10000, fvalmax=1e8, ...);
 * tiktak_cuda_config_defaults() applies them for callers holding raw config.txt values. */
typedef struct tiktak_config {
    int32_t nx;              /* p_nx */
    int32_t nm;              /* p_nmom */
    int32_t maxeval;         /* p_maxeval (DFNLS MAXFUN) */
    int32_t qr_ndraw;        /* p_qr_ndraw, must be < 2^30 (utilities.f90:937) */
    int32_t legitimate;      /* p_legitimate */
    int32_t maxpoints_plus1; /* p_maxpoints = config maxpoints + 1 (genericParams.f90:369) */
    int32_t search_type;     /* p_searchType: 0 = best point so far, 1 = rank lottery (getLotteryPoint :722-766) */
    int32_t lottery_points;  /* p_lotteryPoints: rows the lottery draws from (search_type == 1) */
    double fvalmax;          /* p_fvalmax */
    const double* range;     /* p_range(nx,2) column-major: lo[0..nx), hi[0..nx) */
    const double* init;      /* p_init(nx) */
    const double* bound;     /* p_bound(nx,2) column-major */
    /* --- knobs that do not exist in the reference (all default to reference behaviour) --- */
    int32_t objective_id;    /* TIKTAK_OBJ_* */
    int32_t device;          /* CUDA device ordinal */
    int32_t rank, world;     /* shard of the Sobol index range / restart set owned by this handle */
    int32_t nm_itmax;        /* amoeba evaluation cap; reference value 5000 (minimize.f90:27-35) */
    int32_t nm_shrink_mode;  /* 0 = code behaviour; 1 = README.md:36 simplex bounds from best minima */
    int32_t sobol_quirk;     /* 1 = the reference's single-precision rightmost-zero search (utilities.f90:1090): identical to 0 up to
                              * Sobol index 33,554,435; beyond it (a degenerate 4-point cycle) create() returns TIKTAK_ERR_UNSUPPORTED */
    int32_t text_roundtrip;  /* 1 = values pass through the f40.20 records like the reference's .dat files */
    int32_t round_size;      /* restarts blended against the same best-so-far; 1 = sequential reference order */
    double nm_ftol;          /* p_tolf_amoeba = 1e-4 (genericParams.f90:25) */
    const void* user;        /* objective-specific parameters (e.g. tiktak_smm_params) or NULL */
} tiktak_config;

/* genericParams.f90:307-338 defaults for negative ("-1") entries, applied in place to the raw
 * config.txt scalars; `maxpoints` is the raw value (NOT +1).  Returns TIKTAK_ERR_ARG where the
 * reference stops (maxpoints > qr_ndraw). */
int tiktak_cuda_config_defaults(int32_t nx, int32_t* maxeval, int32_t* qr_ndraw, int32_t* legitimate,
                                double* fvalmax, int32_t* maxpoints, int32_t* lottery_points);

/* obj_initialize (objective.f90:31-37) + allocation of device state.  Replaces the in-memory
 * globals of genericParams / the .dat counters of stateControl.f90 on the hot path. */
int tiktak_cuda_create(tiktak_handle** out, const tiktak_config* cfg);
void tiktak_cuda_destroy(tiktak_handle* h);

/* setupSobol (TiktakGlobalSearch.f90:825-838): insobl + I4_SOBOL + affine map, for points
 * first_1based .. first_1based+count-1.  out is (count, nx) column-major like sobol_trial. */
int tiktak_cuda_sobol_points(tiktak_handle* h, int64_t first_1based, int64_t count, double* out);

/* objFun (objective.f90:125-142) at n points; x is (n, nx) column-major (host or device). */
int tiktak_cuda_objfun(tiktak_handle* h, const double* x, int64_t n, double* fval);

/* solveAtSobolPoints (TiktakGlobalSearch.f90:405-471), states 2-5 in one call: evaluate points
 * 1,2,... of this handle's shard until the `legitimate` cut or qr_ndraw.  With world > 1 the
 * caller all-gathers per-block valid counts (tiktak_cuda_pretest_counts) and sets the global cut
 * with tiktak_cuda_pretest_set_cut; with world == 1 the cut is final on return.
 * n_evaluated = Kcut (the evaluated prefix 1..Kcut), n_legit = #(fval < fvalmax) within it. */
int tiktak_cuda_pretest(tiktak_handle* h, int64_t* n_evaluated, int64_t* n_legit);
/* When legitimate > qr_ndraw (the default, genericParams.f90:317-319) the cut cannot fire: called with
 * n_legit == NULL, tiktak_cuda_pretest then only QUEUES the evaluation of all qr_ndraw points and returns;
 * the valid count (the final legitSobol.dat value) is read back when first asked for, here or by choose. */
int tiktak_cuda_pretest_result(tiktak_handle* h, int64_t* n_evaluated, int64_t* n_legit);
/* the shard's own fval store in one copy: count blocks rank, rank+world, ... (TIKTAK_SOBOL_BLOCK values each, point n of
 * block b at [(b/world)*TIKTAK_SOBOL_BLOCK + n%TIKTAK_SOBOL_BLOCK]) up to the block holding the cut; n_local = values copied. */
int tiktak_cuda_pretest_values_shard(tiktak_handle* h, double* fval_local, int64_t capacity, int64_t* n_local);
/* the same copy queued on a second stream, behind the evaluation and beside whatever the handle does next (selection,
 * restarts); pretest_values_wait returns when it has landed.  The host buffer should be pinned. */
int tiktak_cuda_pretest_values_shard_async(tiktak_handle* h, double* fval_local, int64_t capacity, int64_t* n_local);
int tiktak_cuda_pretest_values_wait(tiktak_handle* h);
/* fval of points first..first+count-1 (the rows of sobolFnVal.dat, :467); points this shard
 * does not own or beyond the cut come back as NaN. */
int tiktak_cuda_pretest_values(tiktak_handle* h, int64_t first_1based, int64_t count, double* fval);
/* multi-GPU helpers: valid counts per TIKTAK_SOBOL_BLOCK-point block of the GLOBAL index range
 * (zeros for blocks of other ranks; summing over ranks gives the global histogram). */
#define TIKTAK_SOBOL_BLOCK 16384
int tiktak_cuda_pretest_wave(tiktak_handle* h, int64_t first_block, int64_t n_blocks); /* evaluate owned blocks */
int tiktak_cuda_pretest_counts(tiktak_handle* h, int64_t* counts, int64_t nblocks);
/* exact Sobol index of the `need`-th valid point inside count block `block` (0 unless this rank owns it) */
int tiktak_cuda_pretest_find(tiktak_handle* h, int64_t block, int64_t need, int64_t* n_out);
int tiktak_cuda_pretest_set_cut(tiktak_handle* h, int64_t kcut, int64_t n_legit);

/* chooseSobol (TiktakGlobalSearch.f90:840-899): row 1 = [objFun(p_init), p_init]; rows 2..K = the
 * best K-1 evaluated Sobol points in ascending fval (ties: ascending Sobol index).
 * x_starts is (K, nx+1) column-major, sobol_idx[K] (0 for row 1).  With world > 1 this selects
 * the shard-local top K-1; merge with tiktak_cuda_choose_merge. */
int tiktak_cuda_choose(tiktak_handle* h, double* x_starts, int32_t* sobol_idx);
/* export this shard's candidates for the all-gather: (K, 2) row-major doubles = K-1 rows [fval, (double)idx]
 * in ascending (fval, idx), NaN-padded, plus a trailer row [#valid points of the shard, #rows filled].
 * A device pointer keeps the export on the handle's stream (no host round trip). */
int tiktak_cuda_choose_candidates(tiktak_handle* h, double* cand);
/* merge the `world` gathered lists (nrows = world*K rows as above, host or device memory) into the
 * replicated x_starts: the position of a row in the union is its index in its own list plus the number of
 * rows below it in every other list -- one kernel, identical result on every rank. */
int tiktak_cuda_choose_merge(tiktak_handle* h, const double* cand, int64_t nrows, double* x_starts,
                             int32_t* sobol_idx);

/* LocalMinimizations + completeSearch (TiktakGlobalSearch.f90:473-600) for restarts
 * first_k .. first_k+n_k-1 of ONE round: starts = getModifiedParam (:683-720) blended against
 * the best row known to the handle when the round begins; then runAmoeba (:602-633) /
 * bobyqa_h (:572); then the re-evaluation (:582).  Blocking convenience form (world == 1 only) =
 * local_enqueue + local_ingest + local_fetch of the same restarts.
 * starts  : (n_k, nx)   column-major, the searchStart.dat rows (:557)
 * results : (n_k, nx+4) column-major = seqNo, k, #improvements, fval, x -- searchResults.dat (:594)
 * evals   : objective evaluations spent per restart (may be NULL). */
int tiktak_cuda_local(tiktak_handle* h, int32_t alg, int32_t first_k, int32_t n_k, double* starts,
                      double* results, int32_t* evals);

/* The same round, device-resident and asynchronous: nothing below synchronises the host, so a whole
 * LocalMinimizations stage is queued on the handle's stream back to back and read once at the end.
 *   local_enqueue : best-so-far -> blend -> local searches of this rank's restarts of the round
 *                   (slot s = k - first_k is run by rank s % world).  With world > 1 the outcomes are
 *                   packed into `send` (DEVICE memory, ceil(n_k/world) rows of nx+2 doubles =
 *                   fval, x, #evaluations) for ONE all-gather per round, issued by the caller on the
 *                   handle's stream (tiktak_cuda_stream) -- e.g. torch.distributed/NCCL over NVLink.
 *   local_ingest  : the round's rows (world > 1: `recv` = the gathered world x ceil(n_k/world) rows,
 *                   device memory) become rows k of the device-resident searchResults table and update
 *                   the best row, identically on every rank.  world == 1: send/recv are NULL.
 *   local_fetch   : synchronise and copy out rows first_k.. (same layouts as tiktak_cuda_local). */
int tiktak_cuda_local_enqueue(tiktak_handle* h, int32_t alg, int32_t first_k, int32_t n_k, double* send);
int tiktak_cuda_local_ingest(tiktak_handle* h, int32_t first_k, int32_t n_k, const double* recv);
int tiktak_cuda_local_fetch(tiktak_handle* h, int32_t first_k, int32_t n_k, double* starts, double* results,
                            int32_t* evals);
/* Asynchronous instances -- the reference's own parallel mode (N processes that each take the next restart number when they
 * finish one and blend the new start toward the best row of searchResults.dat as it is then, TiktakGlobalSearch.f90:546-600,
 * 683-720), with `instances` blocks (or warps: tiktak_cuda_local_async_clock below) of one GPU as the processes and a virtual clock
 * in place of wall time, so that the run is reproducible: restart numbers are handed out in the order of the virtual finish times
 * (ties: lower instance number first) and a start sees exactly the rows that finished at or before its virtual start time.
 * One kernel launch for restarts first_k .. first_k + n_k - 1, rows committed on the device; *used = instances run
 * (min(instances, n_k, what is resident)).  Nelder-Mead with a scalar objective (nx <= 127) or DFNLS (nx <= 100), search_type 0,
 * nm_shrink_mode 0, world 1 (local_async_supported tells); tiktak_cuda_local_fetch then numbers the #improvements column in that file order. */
#define TIKTAK_ASYNC_PICKUP_TICKS 32 /* virtual ticks an instance spends between two restarts (reading the rows, blending the start) */
#define TIKTAK_ASYNC_PICKUP_TICKS_NM_EVALS 128 /* ... on the evaluation clock of the one-warp Nelder-Mead instances: an evaluation is a much
                                                 shorter tick than a pass of the four-warp block, and the pickup is the lookahead every other
                                                 instance gets after a finish event (C3, 4140 instances: 32 ticks 13.4 ms, 128: 8.4 ms, 512: 8.3 ms) */
int tiktak_cuda_local_async(tiktak_handle* h, int32_t alg, int32_t first_k, int32_t n_k, int32_t instances, int32_t* used);
int tiktak_cuda_local_async_supported(tiktak_handle* h, int32_t alg, int32_t* yes);
/* The tick of the virtual clock.  TRIPS: Nelder-Mead with a block of four warps as the instance (nx <= 31, the default there): a restart
 * costs PICKUP + c0 + trips + shrinks * c0 ticks, c0 = ceil((nx+1)/4).  EVALS: a restart costs PICKUP + its objective evaluations --
 * DFNLS, and Nelder-Mead with ONE WARP as the instance (the throughput form: 3.3 times as many instances per GPU, less than half the
 * instructions per evaluation) -- taken for nx 32..127 and whenever `instances` is at least twice what the GPU holds as four-warp
 * blocks (12 per SM: 1776 on a B200); TIKTAK_ASYNC_FORM=quad|w4 in the environment overrides.  Positive `instances` only (free-running
 * instances have no clock). */
#define TIKTAK_ASYNC_CLOCK_TRIPS 0
#define TIKTAK_ASYNC_CLOCK_EVALS 1
int tiktak_cuda_local_async_clock(tiktak_handle* h, int32_t alg, int32_t instances, int32_t* clock);
/* instances < 0 in tiktak_cuda_local_async: |instances| FREE-RUNNING instances (Nelder-Mead, four-warp blocks or one warp by the same rule as
 * the reproducible form) -- no virtual clock: an instance takes the
 * next restart number from a counter and blends toward the best row committed at that moment, exactly as the reference's processes do
 * with searchResults.dat.  Like a multi-process run of the reference it is not reproducible run to run; every restart is still the
 * reference's search from its recorded start.  tiktak_cuda_local_async_info tells, for either form, which row each start was blended
 * toward (-1: none) and the restart's place in the order of the file. */
int tiktak_cuda_local_async_info(tiktak_handle* h, int32_t first_k, int32_t n_k, int32_t* base, int32_t* order);
/* how each restart's local search ended: 0 = as the reference's would; 2 = DFNLS refused the start (bounds closer than
 * 2 rhobeg, minimize.f90:354-362).  (1 was "stopped where the reference calls RESCUE_H" while RESCUE_H was not on the device.) */
int tiktak_cuda_local_status(tiktak_handle* h, int32_t first_k, int32_t n_k, int32_t* status);
/* how often restart k's DFNLS search rebuilt its interpolation system with RESCUE_H (minimize.f90:643-650, :1627-2023: after a
 * denominator was damaged by cancellation, :746-747, :782-783); saturates at 31.  0 for Nelder-Mead. */
int tiktak_cuda_local_nrescue(tiktak_handle* h, int32_t first_k, int32_t n_k, int32_t* nrescue);
/* Rounds started early.  A round depends on the earlier ones only through the best row, which changes in few
 * rounds (the #improvements column).  local_enqueue may therefore be given SEVERAL consecutive rounds at once
 * (n_k = a multiple of round_size, or up to the last restart): all of them start from the best row known now.
 * local_ingest_rounds then ingests them round by round for as long as that row stayed the best one, stops after the
 * first round that changes it (or that brings the row count above 1, which switches the blend on), and returns
 * `accepted` = restarts ingested; the caller queues the rest again.  The accepted sequence is bit for bit the one
 * of strictly sequential rounds; the idle SMs of a small round do the speculative work.  Synchronises the host.
 * local_capacity: how many restarts fit on the machine at once (one resident wave x world) -- the natural size of
 * such a launch; 0 when the mode (lottery, nm_shrink_mode) makes the start depend on more than the best row. */
int tiktak_cuda_local_ingest_rounds(tiktak_handle* h, int32_t first_k, int32_t n_k, int32_t round_size, const double* recv,
                                    int32_t* accepted);
int tiktak_cuda_local_capacity(tiktak_handle* h, int32_t alg, int32_t* restarts);
/* world == 1: local_enqueue + local_ingest_rounds of restarts first_k .. first_k+n_k-1 as ONE call.  With the throughput
 * form of the Nelder-Mead kernel this is one launch: restarts are handed to sub-warp groups by a device-side queue, the
 * group that finishes the last restart of a round commits the complete rounds in order into the results table
 * (completeSearch :582-597), and the first round that changes the best row stops the launch (`accepted` < n_k); how far
 * restarts may be started ahead of the committed rounds is bounded inside the kernel, so n_k may be "all that is left". */
int tiktak_cuda_local_launch(tiktak_handle* h, int32_t alg, int32_t first_k, int32_t n_k, int32_t round_size, int32_t* accepted);
/* SolveMissingSearch (TiktakGlobalSearch.f90:1104-1146), states 8-9: one restart k that is missing from searchResults.dat (an
 * instance was killed before it reported).  Start = getModifiedParam(p_maxpoints, x_starts(k,2:)) (:1121: the blend weight of
 * the last restart number, toward the best row known now), then completeSearch(alg, k).  Push the rows that do exist with
 * tiktak_cuda_push_results first (in file order) and fetch the restarts this handle ran itself (local_fetch) before the sweep.
 * Outputs (host memory, may be NULL): start (nx) = the searchStart.dat row, row (nx+4) = seqNo, k, #improvements, fval, x =
 * the searchResults.dat row, evals.  #improvements counts against the best of ALL rows known, as completeSearch :585-589 does. */
int tiktak_cuda_local_missing(tiktak_handle* h, int32_t alg, int32_t k, double* start, double* row, int32_t* evals);
/* make rows computed elsewhere (other ranks, or a resumed searchResults.dat) known to the handle;
 * rows is (n, nx+4) column-major.  Rows whose fval is NaN are ignored.  Also how the k=0 row
 * [fval(p_init), p_init] (:496-507) gets in: tiktak_cuda_choose pushes it automatically. */
int tiktak_cuda_push_results(tiktak_handle* h, const double* rows, int32_t n);

/* Multi-GPU: the handle owns its NCCL communicator.  The reference's instances coordinate through lock files on a
 * shared directory (stateControl.f90:12-216: getNextNumber / setState / waitState); here rank r of `world` (tiktak_config)
 * owns the Sobol count blocks r, r+world, ... and the restart slots r, r+world, ... of every launch, and once a
 * communicator exists tiktak_cuda_pretest / _choose / _local_launch do their exchanges themselves, on the handle's stream
 * (per-wave all-reduce of the valid counts when the `legitimate` cut can fire; one all-gather of candidate rows; one
 * all-gather of result rows per launch) and return the same global results on every rank.  NCCL is loaded with dlopen.
 *   one process per GPU : rank 0 calls comm_unique_id, the launcher carries the 128 bytes to the other ranks (MPI,
 *                         torch.distributed, a file), every rank calls comm_init -- a collective.
 *   one process, n GPUs : create n handles (rank r on device r) and call comm_init_all once; then one host thread per
 *                         handle makes the stage calls (each is collective over the n handles). */
#define TIKTAK_COMM_ID_BYTES 128
int tiktak_cuda_comm_unique_id(void* id, int32_t bytes);
int tiktak_cuda_comm_init(tiktak_handle* h, const void* id, int32_t bytes);
int tiktak_cuda_comm_init_all(tiktak_handle** handles, int32_t n_handles);

/* getBestLine (TiktakGlobalSearch.f90:781-802): best row so far. */
int tiktak_cuda_best(tiktak_handle* h, double* fbest, double* xbest, int32_t* k);

/* lastSearch (TiktakGlobalSearch.f90:635-681): EST_dfpmin polish from the best row. */
int tiktak_cuda_last_search(tiktak_handle* h, double* fval, double* x, int32_t* iters);

/* introspection for bench.py / tests */
int tiktak_cuda_device_buffer(tiktak_handle* h, const char* name, void** ptr, int64_t* bytes);
int64_t tiktak_cuda_launch_count(tiktak_handle* h);       /* kernels launched so far by this handle */
int tiktak_cuda_stage_ms(tiktak_handle* h, const char* stage, double* ms); /* CUDA-event time of last run */
void* tiktak_cuda_stream(tiktak_handle* h);               /* the cudaStream_t all work is queued on */

const char* tiktak_cuda_strerror(int code);
const char* tiktak_cuda_last_error(void);
const char* tiktak_cuda_version(void);

/* FP64 DFMA-chain microbenchmark (roofline denominator; not part of the reference). Returns FLOP/s. */
int tiktak_cuda_fp64_peak(int device, double* flops_per_s, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* TIKTAK_CUDA_H */
