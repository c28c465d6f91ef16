/* gkquad_b200.h -- C ABI of the B200-native batched adaptive quadrature engine.
 *
 * This is the drop-in boundary for ONE hot path of the Rust crate gkquad 0.0.4
 * (Kogia-sima/gkquad-rs): QAGS / QAGP adaptive Gauss-Kronrod quadrature (qk17 + qk25
 * rules, bisection worklist, Wynn epsilon table, infinite-range transform) and the nested
 * 2-D drivers QAGS2 / QAGP2.  Each entry point below names the reference interface it
 * replaces (paths relative to /root/reference/gkquad/src/).  A host crate binds these
 * symbols through FFI (see INTEGRATION.md for the Rust `extern "C"` block and the
 * `impl Algorithm<DeviceIntegrand> for QagsB200` that sits on top of it).
 *
 * Conventions
 *   - plain C: pointers, sizes, POD structs; no C++/torch types cross the boundary;
 *   - the unit of work is a BATCH of independent integrals (the reference integrates one
 *     closure per call; closures cannot cross an FFI, so integrands are device functors
 *     selected by ID from a registry -- gkq_functor_info / gkq_functor_lookup);
 *   - every function returns 0 on success or a negative gkq_error; numerical failures are
 *     NOT errors: like the reference's ValueWithError (common.rs:60-63) the partial value
 *     is always written and the per-integral `status` carries the RuntimeError;
 *   - there is no CPU fallback: without a CUDA device gkq_create fails with
 *     GKQ_ERR_NO_DEVICE and nothing can be computed.
 */
#ifndef GKQUAD_B200_H
#define GKQUAD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GKQ_ABI_VERSION 2

typedef struct gkq_handle_s* gkq_handle_t;

/* replaces: the Algorithm implementors QAGS / QAGP / AUTO
 * (single/algorithm/{qags.rs:19-63, qagp.rs:20-64, auto.rs:7-35}) and, for the 2-D entry
 * points, QAGS2 / QAGP2 / AUTO2 (double/algorithm/{qags.rs,qagp.rs,auto.rs}). */
/* GKQ_ALGO_QAG: the reference's deprecated QAG / QAG2 (single/algorithm/qag.rs,
 * double/algorithm/qag.rs): available through the same entries, not tuned. */
typedef enum gkq_algo { GKQ_ALGO_AUTO = 0, GKQ_ALGO_QAGS = 1, GKQ_ALGO_QAGP = 2, GKQ_ALGO_QAG = 3 } gkq_algo;

/* replaces: enum Tolerance (common.rs:9-19) and Tolerance::to_abs (common.rs:34-41). */
typedef enum gkq_tol_kind {
    GKQ_TOL_ABSOLUTE = 0,   /* delta <= abs                      */
    GKQ_TOL_RELATIVE = 1,   /* delta <= rel * |I|                */
    GKQ_TOL_ABS_OR_REL = 2, /* delta <= max(abs, rel * |I|)      */
    GKQ_TOL_ABS_AND_REL = 3 /* delta <= min(abs, rel * |I|)      */
} gkq_tol_kind;

typedef struct gkq_tolerance {
    int32_t kind; /* gkq_tol_kind */
    int32_t _pad;
    double abs_tol;
    double rel_tol;
} gkq_tolerance;

/* replaces: enum RuntimeError (error.rs:140-178); numbering = declaration order, 0 = the
 * `None` of ValueWithError::error. */
typedef enum gkq_status {
    GKQ_STATUS_NONE = 0,
    GKQ_STATUS_INSUFFICIENT_ITERATION = 1,
    GKQ_STATUS_ROUNDOFF_ERROR = 2,
    GKQ_STATUS_SUBRANGE_TOO_SMALL = 3,
    GKQ_STATUS_DIVERGENT = 4,
    GKQ_STATUS_NAN_VALUE_ENCOUNTERED = 5
} gkq_status;

typedef enum gkq_error {
    GKQ_OK = 0,
    GKQ_ERR_INVALID_ARGUMENT = -1, /* NULL pointer, negative size, unknown id ...            */
    GKQ_ERR_NAN_INPUT = -2,        /* NaN in tolerance / shared points (the reference panics:
                                      single/integrator.rs:71,86-90, single/common.rs:94)     */
    GKQ_ERR_NO_DEVICE = -3,        /* no usable CUDA device: there is no CPU fallback         */
    GKQ_ERR_CUDA = -4,             /* a CUDA runtime call failed; see gkq_last_error          */
    GKQ_ERR_OUT_OF_MEMORY = -5,
    GKQ_ERR_UNSUPPORTED = -6
} gkq_error;

/* replaces: IntegrationConfig (single/common.rs:108-126; defaults tolerance =
 * AbsOrRel(1.49e-8, 1.49e-8), max_evals = 2000, no points) and IntegrationConfig2
 * (double/common.rs:14-32; default max_evals = 100000).  `points` are the singular points
 * shared by every integral of the batch: for 1-D `npoints` doubles, for 2-D `npoints`
 * interleaved (x, y) pairs.  HOST pointer, read during the call.
 * `workspace_capacity` models WorkSpace::with_capacity(n) (single/workspace.rs:62-70): the
 * Vec capacity is the `limit` seen by qpsrt / increase_nrmax (workspace.rs:165,236).
 * 0 = a fresh WorkSpace::new() per integral, as single::integral does (integral.rs:17-19). */
typedef struct gkq_config {
    int32_t algo; /* gkq_algo */
    int32_t flags; /* GKQ_FLAG_* (0 = none) */
    gkq_tolerance tol;
    uint64_t max_evals;
    uint64_t workspace_capacity;
    const double* points;
    int64_t npoints;
} gkq_config;

/* 2-D entries: integrate x inside and y outside, i.e. the reference's DynamicX range
 * (double/range.rs:38-75): the algorithms swap the arguments of the integrand, hand the x-range
 * functor the outer variable and swap the coordinates of the singular points
 * (double/algorithm/qags.rs:29-44, qagp.rs:34-53).  With this flag the engine evaluates
 * f(inner, outer); x0/x1 are the OUTER (y) range, the range functor yields the inner (x) range and
 * points are given as (outer, inner) pairs -- the host mirrors do the swapping. */
#define GKQ_FLAG_SWAP_XY 1
/* Device entries only see pointers, so they cannot refuse NaN inputs for free the way the reference
 * does when a Range / Tolerance is constructed (single/common.rs:94, single/integrator.rs:71).  With
 * this flag a device entry first checks the range bounds and the per-integral tolerances of
 * gkq_overrides on the device (one small kernel + ONE synchronisation of `stream`) and returns
 * GKQ_ERR_NAN_INPUT before anything is integrated.  Without it: a NaN bound gives that integral
 * status NanValueEncountered (QAGS: after its 17 samples; QAGP: estimate NaN, delta f64::MAX,
 * 0 evaluations); a NaN tolerance makes every `delta <= tolerance` test false for that integral.
 * The host entries (host pointers) always check and need no flag. */
#define GKQ_FLAG_VALIDATE_INPUTS 2

/* Fills *cfg with the reference's defaults: dim 1 -> IntegrationConfig::default(),
 * dim 2 -> IntegrationConfig2::default(). */
int gkq_config_default(int dim, gkq_config* cfg);

/* ---- lifetime ------------------------------------------------------------------------- */

/* Creates an engine bound to CUDA device `device` (one handle per device per host thread;
 * calls on one handle are stream-ordered).  Replaces the owned state of an Integrator
 * (single/integrator.rs:31-35: algorithm + workspace): the handle owns all device scratch
 * (survivor worklists, per-thread interval slabs, staging buffers), grown on demand and
 * reused, so a warm call allocates nothing. */
int gkq_create(int device, gkq_handle_t* out);
/* The same, with a CUDA stream priority for the streams the handle owns (0 = default, negative =
 * scheduled first, clamped to the device's range).  For a caller that runs several handles on several
 * streams and wants one lane's batches out first -- e.g. the batch whose results are gathered to another
 * GPU while the other lanes still compute; pass the same priority to the stream of the calls. */
int gkq_create_prio(int device, int stream_priority, gkq_handle_t* out);
int gkq_destroy(gkq_handle_t h);
/* Message of the last failure on this handle (or of the last failed gkq_create when h is
 * NULL).  Never NULL. */
const char* gkq_last_error(gkq_handle_t h);
int gkq_abi_version(void);

/* ---- integrand registry ----------------------------------------------------------------
 * replaces: trait Integrand (single/common.rs:129-145) / Integrand2 (double/common.rs:35-45)
 * and DynamicY's range closure (double/range.rs:77-112).
 * dim: 1 = 1-D integrands f(x; p), 2 = 2-D integrands f(x, y; p), 0 = y-range functors
 * y(x; q) -> (y0, y1).  `flops_per_eval` is the algorithmic FP64 flop count of one
 * evaluation used for the roofline figure. */
int gkq_functor_count(int dim);
int gkq_functor_info(int dim, int id, const char** name, int* nparams, double* flops_per_eval);
int gkq_functor_lookup(int dim, const char* name); /* id, or GKQ_ERR_INVALID_ARGUMENT */

/* ---- 1-D batches ------------------------------------------------------------------------
 * replaces: Algorithm::integrate(&mut self, f, range, config) -> IntegrationResult
 * (single/algorithm/mod.rs:16-23) as reached from Integrator::run (single/integrator.rs:
 * 103-106), single::integral (integral.rs:17-19) and integral_with_config (:25-31), for `n`
 * independent integrals at once.
 *
 *   a, b          range of integral i is [a[i*range_stride], b[i*range_stride]];
 *                 range_stride 1 = per-integral, 0 = one shared range.  +-inf allowed:
 *                 either end non-finite => the variable transform of single/util.rs:74-103,
 *                 162-176 is applied to that integral (qags.rs:48, qagp.rs:50).
 *   params        structure-of-arrays: parameter k of integral i is
 *                 params[k*param_stride + i]; param_stride 0 = params[k] shared by all.
 *                 May be NULL when the functor has no parameters.
 *   pts_offsets,  optional per-integral singular points in CSR form (n+1 offsets into pts);
 *   pts           NULL = every integral uses cfg->points.  With GKQ_ALGO_AUTO an integral
 *                 with no points runs QAGS, otherwise QAGP (auto.rs:29-33).
 *   out_*         estimate, delta (Solution, common.rs:213-231), nevals, status.
 *
 * gkq_integrate_batch: every array pointer is a DEVICE pointer; work is enqueued on
 * `stream` (a cudaStream_t, NULL = default stream) and the call returns without
 * synchronising.  gkq_integrate_batch_host: every array pointer is a HOST pointer
 * (pageable or pinned); the call stages inputs to the device, computes, copies the four
 * result arrays back (chunked and overlapped) and returns when they are complete. */
int gkq_integrate_batch(gkq_handle_t h, const gkq_config* cfg, int functor_id, int64_t n,
                        const double* a, const double* b, int64_t range_stride,
                        const double* params, int64_t param_stride,
                        const int64_t* pts_offsets, const double* pts,
                        double* out_estimate, double* out_delta, uint32_t* out_nevals,
                        uint32_t* out_status, void* stream);

int gkq_integrate_batch_host(gkq_handle_t h, const gkq_config* cfg, int functor_id, int64_t n,
                             const double* a, const double* b, int64_t range_stride,
                             const double* params, int64_t param_stride,
                             const int64_t* pts_offsets, const double* pts,
                             double* out_estimate, double* out_delta, uint32_t* out_nevals,
                             uint32_t* out_status);

/* ---- 2-D batches ------------------------------------------------------------------------
 * replaces: Algorithm2::integrate(&mut self, f, range, config) (double/algorithm/mod.rs:7-10)
 * for R = Rectangle / DynamicY (double/range.rs:11-34, 77-112) as reached from
 * Integrator2::run (double/integrator.rs:80-87) and integral2 (double/integral.rs:9-16).
 * Outer range of integral i: [x0[i*range_stride], x1[i*range_stride]]; inner range
 * y(x) = yrange functor `yrange_id` with parameters yparams (SoA like params).  A Rectangle
 * is yrange "const" with q = (y0, y1) (double/algorithm/qags.rs:18-27).  cfg->points holds
 * (x, y) pairs; per-integral CSR pts_offsets counts pairs, pts_xy is interleaved. */
int gkq_integrate2_batch(gkq_handle_t h, const gkq_config* cfg, int functor2_id, int yrange_id,
                         int64_t n, const double* x0, const double* x1, int64_t range_stride,
                         const double* fparams, int64_t fparam_stride, const double* yparams,
                         int64_t yparam_stride, const int64_t* pts_offsets, const double* pts_xy,
                         double* out_estimate, double* out_delta, uint32_t* out_nevals,
                         uint32_t* out_status, void* stream);

int gkq_integrate2_batch_host(gkq_handle_t h, const gkq_config* cfg, int functor2_id,
                              int yrange_id, int64_t n, const double* x0, const double* x1,
                              int64_t range_stride, const double* fparams, int64_t fparam_stride,
                              const double* yparams, int64_t yparam_stride,
                              const int64_t* pts_offsets, const double* pts_xy,
                              double* out_estimate, double* out_delta, uint32_t* out_nevals,
                              uint32_t* out_status);

/* ---- per-integral configuration -------------------------------------
 * replaces: an Integrator whose tolerance / max_evals differ from integral to integral
 * (single/integrator.rs:69-84 called once per integral in a parameter sweep).  Each array is
 * optional (NULL = the value of `cfg` for every integral) and has n entries; the tolerance KIND
 * stays cfg->tol.kind.  max_evals[i] is clamped to cfg->max_evals, which sizes the engine's
 * interval lists.  NaN tolerances: rejected with GKQ_ERR_NAN_INPUT by the host entries and,
 * with GKQ_FLAG_VALIDATE_INPUTS, by the device entries (the reference panics, single/integrator.rs:71).  For the 2-D entries max_evals[i] is the TOTAL budget of nest i
 * (IntegrationConfig2::max_evals, double/common.rs:18).
 * *_batch_ex: device pointers; *_batch_host_ex: host pointers. */
typedef struct gkq_overrides {
    const double* abs_tol;
    const double* rel_tol;
    const uint64_t* max_evals;
} gkq_overrides;
int gkq_integrate_batch_ex(gkq_handle_t h, const gkq_config* cfg, const gkq_overrides* ov,
                           int functor_id, int64_t n, const double* a, const double* b,
                           int64_t range_stride, const double* params, int64_t param_stride,
                           const int64_t* pts_offsets, const double* pts, double* out_estimate,
                           double* out_delta, uint32_t* out_nevals, uint32_t* out_status, void* stream);
int gkq_integrate_batch_host_ex(gkq_handle_t h, const gkq_config* cfg, const gkq_overrides* ov,
                                int functor_id, int64_t n, const double* a, const double* b,
                                int64_t range_stride, const double* params, int64_t param_stride,
                                const int64_t* pts_offsets, const double* pts, double* out_estimate,
                                double* out_delta, uint32_t* out_nevals, uint32_t* out_status);

int gkq_integrate2_batch_ex(gkq_handle_t h, const gkq_config* cfg, const gkq_overrides* ov,
                            int functor2_id, int yrange_id, int64_t n, const double* x0,
                            const double* x1, int64_t range_stride, const double* fparams,
                            int64_t fparam_stride, const double* yparams, int64_t yparam_stride,
                            const int64_t* pts_offsets, const double* pts_xy, double* out_estimate,
                            double* out_delta, uint32_t* out_nevals, uint32_t* out_status, void* stream);
int gkq_integrate2_batch_host_ex(gkq_handle_t h, const gkq_config* cfg, const gkq_overrides* ov,
                                 int functor2_id, int yrange_id, int64_t n, const double* x0,
                                 const double* x1, int64_t range_stride, const double* fparams,
                                 int64_t fparam_stride, const double* yparams, int64_t yparam_stride,
                                 const int64_t* pts_offsets, const double* pts_xy, double* out_estimate,
                                 double* out_delta, uint32_t* out_nevals, uint32_t* out_status);

/* ---- rule level -------------------------------------------------------------------------
 * replaces: single::qk17 / qk25 / qk33 / qk41 / qk49 / qk57 (single/qk/mod.rs:28-73) -> QKResult
 * (:16-25).  `rule` is the number of points: 17, 25, 33, 41, 49 or 57.  Device pointers; out4 is
 * [4][n] SoA: estimate, delta, absvalue, asc. */
int gkq_qk_batch(gkq_handle_t h, int rule, int functor_id, int64_t n, const double* a,
                 const double* b, int64_t range_stride, const double* params,
                 int64_t param_stride, double* out4, void* stream);

/* replaces: Integrand::apply_to_slice (single/common.rs:135-137): y[i] = f(x[i]; params_i).
 * Device pointers.  Also the micro-kernel on which the registry's flops_per_eval figures are
 * measured (ncu: DADD + DMUL + 2*DFMA thread instructions / n). */
int gkq_functor_eval_batch(gkq_handle_t h, int functor_id, int64_t n, const double* x,
                           const double* params, int64_t param_stride, double* y, void* stream);
/* The same with the evaluation path chosen: mode 0 = the always-valid path (what gkq_functor_eval_batch
 * runs), mode 1 = the sample exactly as the rule kernels take it: branch-free math first (the engine's own
 * exp / sin / cos / log / pow; IEEE division and square root without the compiler's slow-path branch), then
 * re-evaluated through mode 0's path when an operand left the fast range.  Both modes must agree bit for
 * bit on every input; the parity tests use this entry to check that against the CPU. */
int gkq_functor_eval_batch_ex(gkq_handle_t h, int functor_id, int64_t n, const double* x,
                              const double* params, int64_t param_stride, double* y, int mode, void* stream);

/* Deferred-join mode (throughput pipelining of independent batches).  By default a batch call is
 * fully ordered on `stream`.  With deferred join enabled, gkq_integrate_batch still enqueues every
 * bulk kernel on `stream`, but the latency-bound remainders of QAGS batches (the few integrals
 * that need many bisection steps) accumulate as self-contained records; when the record buffer is
 * full (8 batches) ONE launch of the persistent kernel finishes them on a handle-owned stream,
 * beside whatever is enqueued next (the following batches append to the twin buffer).  The outputs
 * of a batch are complete only after gkq_join(h, stream) (finishes the pending records and makes
 * `stream` wait for all outstanding remainders) or gkq_sync.
 * The reference has no counterpart (it is synchronous); an Integrator-level sweep over many
 * parameter batches is where this applies.  A handle serves ONE stream at a time; to keep two
 * batches in flight (the single-wave survivor kernels of one fill the ramps and tails of the
 * other: +20 % on B200) drive two handles on two streams, as the host mirrors' BatchPipeline does. */
int gkq_set_deferred_join(gkq_handle_t h, int enable);
int gkq_join(gkq_handle_t h, void* stream);

/* ---- multi-GPU: static shard + one gather of results -------------------
 * replaces: nothing inside gkquad (it is single-threaded); this is what a host that runs
 * Algorithm::integrate / Algorithm2::integrate (single/algorithm/mod.rs:16-23,
 * double/algorithm/mod.rs:7-10) over a batch larger than one GPU binds: one process (or thread) and
 * one handle per GPU, a STATIC block-cyclic partition of the batch and no exchange while it is
 * integrated -- only the results travel.
 *
 *   partition    integral i belongs to rank (i / block) mod nranks; position j of rank r's shard is
 *                integral gkq_shard_index(n_total, nranks, r, block, j); every rank integrates its
 *                gkq_shard_size(..) integrals with the ordinary batch entries (heavy-tailed work --
 *                QAGP, 2-D -- balances statically because consecutive blocks go to different ranks).
 *   communicator gkq_comm_get_unique_id on one rank, the 128 bytes carried to the others by the
 *                host's own means (MPI, a file, torch.distributed ...), gkq_comm_init on every rank
 *                (collective).  NCCL is loaded at run time (libnccl.so.2; env GKQ_NCCL_LIB overrides);
 *                nranks == 1 needs no NCCL at all.
 *   gkq_gather   collective, asynchronous on `stream`: packs the local shard's results to 20 bytes
 *                per integral (estimate f64 | delta f64 | nevals << 3 | status u32 -- cfg->max_evals
 *                must stay below 2^29), moves them over NVLink (grouped ncclSend / ncclRecv to
 *                `root`, or ncclAllGather when root < 0) and un-permutes them into GLOBAL integral
 *                order on the receiving rank(s): g_* are device arrays of n_total entries there and
 *                ignored elsewhere.  est / dlt / nev / st are the device arrays the shard's results
 *                were written to (shard order).  On a side stream the gather of one batch overlaps
 *                the kernels of the next.
 *   gkq_gather_range  the same for shard positions [local_offset, local_offset + local_count) only
 *                (every rank passes the same range; local_offset a multiple of 4), so that a batch
 *                computed in chunks can be gathered chunk by chunk behind the kernels and only the
 *                last chunk's transfer is exposed. */
#define GKQ_COMM_ID_BYTES 128
int gkq_comm_get_unique_id(void* id_out /* [GKQ_COMM_ID_BYTES] */);
int gkq_comm_init(gkq_handle_t h, int nranks, int rank, const void* unique_id);
int gkq_comm_destroy(gkq_handle_t h);
int gkq_comm_info(gkq_handle_t h, int* nranks, int* rank);
int64_t gkq_shard_size(int64_t n_total, int nranks, int rank, int64_t block);
int64_t gkq_shard_index(int64_t n_total, int nranks, int rank, int64_t block, int64_t j);
int gkq_gather(gkq_handle_t h, int64_t n_total, int64_t block, int root, const double* est,
               const double* dlt, const uint32_t* nev, const uint32_t* st, double* g_est,
               double* g_dlt, uint32_t* g_nev, uint32_t* g_st, void* stream);
int gkq_gather_range(gkq_handle_t h, int64_t n_total, int64_t block, int root, int64_t local_offset,
                     int64_t local_count, const double* est, const double* dlt, const uint32_t* nev,
                     const uint32_t* st, double* g_est, double* g_dlt, uint32_t* g_nev,
                     uint32_t* g_st, void* stream);

/* ---- stream helpers / accounting --------------------------------------------------------- */
int gkq_sync(gkq_handle_t h, void* stream);

typedef struct gkq_stats {
    uint64_t kernel_launches;   /* kernels of this library launched through the handle     */
    uint64_t batches;           /* integrate calls                                         */
    uint64_t integrals;         /* integrals submitted                                     */
    uint64_t scratch_bytes;     /* device scratch currently owned by the handle            */
    uint64_t last_survivors_qk25;  /* worklist sizes of the last QAGS batch run with profiling */
    uint64_t last_survivors_loop;  /* on (gkq_set_profiling), valid after gkq_sync: -> qk25 pass, -> loop */
    uint64_t last_survivors_step2; /* integrals still unfinished after the first bisection  */
    uint64_t last_chain_loop[2];   /* the same two counts per chain of the QAGS pipeline:   */
    uint64_t last_chain_step2[2];  /* [0] straight from qk17, [1] through the qk25 pass      */
} gkq_stats;
int gkq_get_stats(gkq_handle_t h, gkq_stats* out);
int gkq_reset_stats(gkq_handle_t h);

/* Per-kernel timing for the roofline figure: when enabled every kernel launch of a batch call is
 * bracketed by CUDA events on the launching stream (adds ~2 us per launch, so leave it off for
 * throughput runs).  gkq_get_profile waits for the last batch and returns the number of
 * entries written: kernel names (static strings) and durations in ms, in launch order. */
int gkq_set_profiling(gkq_handle_t h, int enable);
int gkq_get_profile(gkq_handle_t h, int max_entries, const char** names, float* ms);

/* FP64 pipe calibration for the roofline denominator: runs a dependent-chain DFMA kernel
 * that fills every SM and writes the achieved FLOP/s (2 flops per DFMA) to *flops_per_s.
 * (MEASURED_PEAKS.json has no FP64 figure) */
int gkq_measure_fp64_peak(gkq_handle_t h, int repeats, double* flops_per_s, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* GKQUAD_B200_H */
