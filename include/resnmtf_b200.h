/*
 * resnmtf_b200.h -- C ABI of the B200-native ResNMTF multiplicative-update loop.
 *
 * The reference (eso28599/resnmtf) is pure R and has no FFI; the seam this library replaces is the
 * loop body of res_nmtf_inner(), R/main.r:50-109, whose only callees are update_matrices()
 * (R/update_steps.r:272-319) and calculate_error() (R/utils.r:157-166).  A drop-in R package keeps
 * res_nmtf_inner()/apply_resnmtf() unchanged and turns R/main.r:50-109 into one .Call that lands on
 * the functions below (see INTEGRATION.md for the .Call shim and the ctypes binding).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  All host matrices are column-major doubles, exactly
 *     as R stores them; host buffers are borrowed for the duration of the call only.
 *   - Views, rows and columns are 0-based here (R is 1-based; the shim subtracts 1).
 *   - Every function returns 0 (RESNMTF_OK) or a negative RESNMTF_E_* code; the message is available
 *     from resnmtf_last_error() (thread-local).  No C++ exception crosses the ABI.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with
 *     RESNMTF_E_CUDA.
 *   - Entry points on one ctx are not re-entrant; different ctx are independent (a placed fit uses all of its contexts).
 */
#ifndef RESNMTF_B200_H
#define RESNMTF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RESNMTF_OK 0
#define RESNMTF_E_INVALID (-1)     /* bad argument (NULL, out of range, shape mismatch)              */
#define RESNMTF_E_CUDA (-2)        /* CUDA runtime error, or no CUDA device                          */
#define RESNMTF_E_NOMEM (-3)       /* host or device allocation failed                               */
#define RESNMTF_E_STATE (-4)       /* call sequence error (data / factors not set before run)        */
#define RESNMTF_E_NAN (-5)         /* mean error became NaN in convergence mode: the reference's     */
                                   /* while(NA) at R/main.r:55 throws "missing value where           */
                                   /* TRUE/FALSE needed"; the R shim turns this code into that error */
#define RESNMTF_E_UNSUPPORTED (-6) /* k outside 1..RESNMTF_MAX_K                                     */
#define RESNMTF_E_COMM (-7)        /* NCCL error on the row-sharded path                             */

#define RESNMTF_MAX_K 16

/* which index space a shared-name map refers to (replaces the per-view `hash` objects built by
 * produce_indices(), R/utils.r:560-601, and consumed by star_prod_relevant(), R/utils.r:63-78) */
#define RESNMTF_MAP_ROW 0 /* rows of F (phi coupling)    -- row_indices[[v]]    */
#define RESNMTF_MAP_COL 1 /* rows of G (psi coupling)    -- column_indices[[v]] */

/* how the per-iteration error of calculate_error() (R/utils.r:157-166) is evaluated */
#define RESNMTF_ERR_AUTO 0      /* algebraic (no pass over X); re-done by a direct pass when < 1e-4  */
#define RESNMTF_ERR_ALGEBRAIC 1 /* ||X||^2 - 2<A,S> + <(F'F) S (G'G), S>, always                     */
#define RESNMTF_ERR_DIRECT 2    /* sum (X - F S G')^2 streamed over X, always (one extra pass)       */

/* streaming-kernel implementation of the two tall-skinny products */
#define RESNMTF_IMPL_AUTO 0
#define RESNMTF_IMPL_DFMA 1 /* CUDA-core FP64 FMA                                    */
#define RESNMTF_IMPL_DMMA 2 /* FP64 tensor-core mma.sync m8n8k4 (k <= 8), X loaded straight into fragments */
#define RESNMTF_IMPL_TMA 3  /* same MMAs, X staged through a shared-memory ring by TMA bulk copies; k <= 16 (two 8-wide
                               tiles in the factor dimension for k = 9..16, the k-extension loop of R/main.r:306-320)   */
#define RESNMTF_IMPL_FUSED 4 /* one pass over X per update-iteration: 8-row groups resident in the shared memory of a
                                cluster of 1..8 CTAs do the F step and the G step (k <= 8, p <= 8064, one GPU per view,
                                at most 7 phi partners); views that do not qualify run RESNMTF_IMPL_TMA.  This is what
                                RESNMTF_IMPL_AUTO picks for views of matrix size */
#define RESNMTF_IMPL_SMALL 5 /* the whole loop of a tiny fit (BASELINE configs[0]: every view <= 1024 x 1024 padded and <= 1 MB,
                                k <= 8, one GPU) as ONE persistent launch of one CTA: all views, all sweeps of a batch, the stop
                                rule on the device.  RESNMTF_IMPL_AUTO picks it for fits of at most 16384 padded entries in
                                total (where it beats one launch per view and sweep); a fit that does not qualify runs the
                                streaming kernels */

typedef struct resnmtf_ctx resnmtf_ctx;
typedef struct resnmtf_fit resnmtf_fit;

/* Counters of the last resnmtf_fit_run / resnmtf_fit_step on a fit. */
typedef struct resnmtf_counters {
  int64_t iterations;       /* update-iterations (update_matrices sweeps) done since set_factors      */
  int64_t kernel_launches;  /* CUDA kernels launched by the last run (graph nodes counted per replay) */
  double device_ms;         /* CUDA-event time of the last run on the fit's stream                    */
  double alg_bytes_per_iter;/* algorithmic HBM bytes per update-iteration (SURVEY 8d B_alg)           */
  int64_t direct_error_passes; /* how many times the direct residual pass ran in the last run        */
  int32_t converged;        /* 1 when the last run stopped on |d err| <= tol                          */
  int32_t impl;             /* RESNMTF_IMPL_* actually used                                           */
} resnmtf_counters;

/* ---- context: one CUDA device, one stream ------------------------------------------------------- */

/* device < 0 selects the current CUDA device. */
int resnmtf_ctx_create(int device, resnmtf_ctx** out);
int resnmtf_ctx_destroy(resnmtf_ctx* ctx);
/* cudaStream_t the context launches on, as an opaque pointer (for CUDA-event timing by the caller). */
void* resnmtf_ctx_stream(resnmtf_ctx* ctx);
int resnmtf_ctx_synchronize(resnmtf_ctx* ctx);
/* CUDA device index of the context (-1 for NULL). */
int resnmtf_ctx_device(resnmtf_ctx* ctx);
/* Thread-local message of the last failure on the calling thread (never NULL). */
const char* resnmtf_last_error(void);
/* Library version string, and the number of CUDA devices visible (0 when none / no driver). */
const char* resnmtf_version(void);
int resnmtf_device_count(void);

/* ---- fit: device-resident X, F, S, G, lambda, mu of all views of one res_nmtf_inner() call -------- */

/* n[v] x p[v] is the shape of view v, k[v] its number of clusters (k_vec of R/main.r:35). */
int resnmtf_fit_create(resnmtf_ctx* ctx, int n_views, const int64_t* n, const int64_t* p,
                       const int32_t* k, resnmtf_fit** out);
/* The same fit with its views PLACED on several GPUs of this process: view v lives on view_ctx[v] (contexts may repeat;
 * all equal = resnmtf_fit_create).  What the phi / psi / xi terms of a view need from its partners -- the mapped rows of
 * F^(w) and G^(w), the k x k S^(w) (star_prod_relevant / star_prod, R/utils.r:39-78) -- is read by the update kernels
 * straight from the partner GPU's memory over NVLink (peer access is enabled between the contexts), and the Gauss-Seidel
 * order of update_matrices() (R/update_steps.r:282-314) is kept across GPUs by events: a view's kernels start when the
 * coupled views before it have finished this sweep; views that are not coupled to each other update concurrently.
 * Results are identical to the single-GPU fit.  Every other entry point takes such a fit unchanged (set_data /
 * attach_data / set_data_device act on the view's GPU; resnmtf_fit_profile is not available). */
int resnmtf_fit_create_placed(resnmtf_ctx* const* view_ctx, int n_views, const int64_t* n, const int64_t* p,
                              const int32_t* k, resnmtf_fit** out);
int resnmtf_fit_destroy(resnmtf_fit* fit);

/* Copies view v (column-major, leading dimension ld >= n[v]) to the device and computes
 * data_norms[v] = ||X||_F^2 (R/main.r:48) there.  x may be pageable or pinned host memory. */
int resnmtf_fit_set_data(resnmtf_fit* fit, int v, const double* x, int64_t ld);
/* Same, but x is a DEVICE pointer on the fit's device (no host round trip). */
int resnmtf_fit_set_data_device(resnmtf_fit* fit, int v, const double* x_dev, int64_t ld);

/* A view's data uploaded ONCE and shared by several fits: apply_resnmtf() fits the same `data` for every k
 * of its sweep (R/main.r:279-287) and again in the k-extension loop (:306-320).  The handle is reference
 * counted: it may be destroyed while fits are still attached.  resnmtf_fit_attach_data replaces set_data
 * for that view (data_norms comes with the handle). */
typedef struct resnmtf_data resnmtf_data;
int resnmtf_data_create(resnmtf_ctx* ctx, int64_t n, int64_t p, const double* x, int64_t ld,
                        resnmtf_data** out);
/* Same, but x is a DEVICE pointer on ctx's device: the sub-sampled views of the stability analysis
 * (R/stability_analysis.r:215-253) are gathered on the device from the resident data and never visit the host. */
int resnmtf_data_create_device(resnmtf_ctx* ctx, int64_t n, int64_t p, const double* x_dev, int64_t ld,
                               resnmtf_data** out);
int resnmtf_data_destroy(resnmtf_data* data);
int resnmtf_fit_attach_data(resnmtf_fit* fit, int v, resnmtf_data* data);

/* ---- data handles on the device: prep, shuffles, sub-samples ------------------------------- */

/* resnmtf_data_create followed by make_non_neg_inner() and matrix_normalisation() (R/utils.r:20-27, 86-88) on the device:
 * every column is shifted by |min(0, min(column))| and divided by its sum.  *was_negative (may be NULL) is set to 1 when
 * a negative entry was seen -- the caller then issues the reference's warning "Matrix is not non-negative. Has been made
 * non-negative." (R/utils.r:24). */
int resnmtf_data_create_prepped(resnmtf_ctx* ctx, int64_t n, int64_t p, const double* x, int64_t ld,
                                int32_t* was_negative, resnmtf_data** out);
int resnmtf_data_shape(resnmtf_data* data, int64_t* n, int64_t* p);
/* The view back on the host, column-major with leading dimension ld (e.g. the prepped data for R-side post-processing). */
int resnmtf_data_download(resnmtf_data* data, double* x, int64_t ld);
/* colSums (length p) and rowSums (length n) of the view; either pointer may be NULL. */
int resnmtf_data_sums(resnmtf_data* data, double* col_sums, double* row_sums);
/* shuffle_view() (R/obtain_bicl.r:11-22): a new handle holding matrix(sample(x), nrow, ncol) -- ALL entries permuted by
 * a keyed bijection of the index range (a function of `seed` only; the reference draws from R's global stream) --
 * reshuffled until no row and no column sums to zero.  renormalise != 0 applies the prep apply_resnmtf() gives the
 * shuffled views (R/obtain_bicl.r:35 -> check_inputs).  *attempts (may be NULL) receives the number of shuffles. */
int resnmtf_data_shuffle(resnmtf_data* src, uint64_t seed, int renormalise, int64_t* attempts, resnmtf_data** out);
/* x[rows, cols] (0-based indices) as a new handle: the sub-samples of stability_repeat() (R/stability_analysis.r:215-253)
 * are gathered from the resident view; no re-normalisation (the reference does none, quirk Q10). */
int resnmtf_data_subsample(resnmtf_data* src, const int32_t* rows, int64_t n_rows, const int32_t* cols, int64_t n_cols,
                           resnmtf_data** out);
/* A copy of the view (and of its cached SVD triplets) on another context's GPU, device to device. */
int resnmtf_data_copy(resnmtf_data* src, resnmtf_ctx* dst_ctx, resnmtf_data** out);

/* ---- SVD initialisation ------------------------------------------------------------ */

/* What init_mats_inner() (R/update_steps.r:92-95) takes from svd(x): |U[, 1:k]| (n x k), d[1:k], |V[, 1:k]| (p x k),
 * column-major, computed on the device from the Gram matrix of the smaller side (FP64 tensor-core product), its top
 * eigenpairs (Chebyshev-filtered subspace iteration with locking, residuals at rounding level of the matrix norm) and one
 * more pass over the view for the other side.  Computed once per handle to width min(16, n, p) and cached: the fits of a
 * k-sweep slice the same triplets.  Any output pointer may be NULL. */
int resnmtf_data_svd_topk(resnmtf_data* data, int k, double* u, double* d, double* v);

/* Initial factors of view v: F n x k, S k x k, G p x k, lambda k, mu k (what init_mats(),
 * R/update_steps.r:36-66, hands to the loop).  lambda / mu may be NULL: they are then set to
 * colSums(F) / colSums(G) as R/update_steps.r:53-54 does.  Resets the iteration counter. */
int resnmtf_fit_set_factors(resnmtf_fit* fit, int v, const double* f, const double* s,
                            const double* g, const double* lambda, const double* mu);

/* phi, xi, psi: n_views x n_views column-major, ALREADY symmetrised by init_rest_mats()
 * (R/update_steps.r:12-24).  NULL means all-zero. */
int resnmtf_fit_set_restrictions(resnmtf_fit* fit, const double* phi, const double* xi,
                                 const double* psi);

/* Shared names between views v and w as index pairs: row (kind ROW) / column (kind COL) idx_v[i] of
 * view v carries the same name as idx_w[i] of view w.  len == 0 stores R's NA (the pair shares
 * nothing and is skipped by star_prod_relevant, R/utils.r:70).  Must be set for both (v,w) and (w,v)
 * when both directions are coupled.  A pair that was never set behaves like R's `indices = NULL`
 * (quirk of R/main.r:312: nothing is overwritten, the view is pulled towards itself). */
int resnmtf_fit_set_shared_map(resnmtf_fit* fit, int kind, int v, int w, const int32_t* idx_v,
                               const int32_t* idx_w, int64_t len);

/* Options: err_mode RESNMTF_ERR_*, impl RESNMTF_IMPL_* (both default AUTO). */
int resnmtf_fit_set_options(resnmtf_fit* fit, int err_mode, int impl);

/* Runs the loop of R/main.r:50-109 on the device.
 *   n_iters >= 0 : exactly n_iters sweeps (R/main.r:83-108).
 *   n_iters <  0 : until |mean_err - previous| <= tol (R/main.r:55-81; previous starts at 0);
 *                  max_iters > 0 adds a cap the reference does not have (<= 0: none).
 * iters_done receives the number of sweeps of this call.  Returns RESNMTF_E_NAN when the mean error
 * became NaN in convergence mode. */
int resnmtf_fit_run(resnmtf_fit* fit, int64_t n_iters, double tol, int64_t max_iters,
                    int64_t* iters_done);
/* One update_matrices() sweep + calculate_error() (for per-iteration parity checks). */
int resnmtf_fit_step(resnmtf_fit* fit);

/* Current (un-normalised) factors of view v, column-major, any pointer may be NULL. */
int resnmtf_fit_get_factors(resnmtf_fit* fit, int v, double* f, double* s, double* g,
                            double* lambda, double* mu);
/* normalisation_check() (R/utils.r:176-195) applied on the device to all views. */
int resnmtf_fit_normalise(resnmtf_fit* fit);
/* Mean error of every sweep since set_factors (All_Error of R/main.r:134); writes min(cap, count)
 * values, *count receives the total. */
int resnmtf_fit_get_errors(resnmtf_fit* fit, double* out, int64_t cap, int64_t* count);
/* Per-view error of the last sweep (calculate_error's vector) and data_norms. */
int resnmtf_fit_get_view_errors(resnmtf_fit* fit, double* err, double* data_norms);
int resnmtf_fit_get_counters(resnmtf_fit* fit, resnmtf_counters* out);

/* Runs n_iters sweeps with CUDA events between the kernel launches (no graph) and returns the summed
 * device time per kernel class: ms[0] F step (X.G + F update), ms[1] G step (X'.F + G, S, lambda, mu update +
 * algebraic error), ms[2] one-pass fused step (views on RESNMTF_IMPL_FUSED: both of the above in one launch), ms[3]
 * direct residual pass, ms[4] iteration
 * bookkeeping; launches[i] is the number of timed intervals of that class. */
int resnmtf_fit_profile(resnmtf_fit* fit, int64_t n_iters, double ms[5], int64_t launches[5]);

/* ---- post-fit reductions -------------------------------------------------------- */

/* jsd_calc() of R/utils.r:95-106 for a batch of column pairs, one thread block per pair:
 *   out[i] = JSD(density(vecs[, pair_a[i]]), density(vecs[, pair_b[i]]))
 * with stats::density's defaults (gaussian kernel, 512 points, from = 0, to = the larger of the two column maxima,
 * estimates zeroed above the column's own maximum) and philentropy::JSD's base-2 logarithm on the normalised
 * estimates.  This is the inner operation of the spurious-bicluster test (calculate_f_shuffle_jsd(),
 * R/obtain_bicl.r:55-68, and check_biclusters(), :113-133): 15 k^2 pairs per fit.
 * vecs: n x m column-major HOST matrix (leading dimension ld) -- the factor columns; bw[j] = bw.nrd0(vecs[, j]) and
 * vmax[j] = max(vecs[, j]) are computed by the caller once per column (O(n), need order statistics).  0-based column
 * indices.  The sums are evaluated in a fixed order: two calls give bit-identical results. */
int resnmtf_jsd_pairs(resnmtf_ctx* ctx, const double* vecs, int64_t n, int32_t m, int64_t ld,
                      const double* bw, const double* vmax, const int32_t* pair_a, const int32_t* pair_b,
                      int64_t n_pairs, double* out);

/* ---- bisilhouette distance blocks ------------------------------------------------------------- */

#define RESNMTF_DIST_EUCLIDEAN 0
#define RESNMTF_DIST_MANHATTAN 1
#define RESNMTF_DIST_COSINE 2

/* bisilhouette::bisilhouette() as obtain_biclusters() calls it per view (R/obtain_bicl.r:190-199) on the resident view:
 * for every non-empty bicluster (R_j, C_j) the mean silhouette of the rows of R_j, distances taken on the columns C_j
 * only, against the other row clusters (rows of R_j excluded) or -- when there is none -- against all other rows.
 * row_cl (n x k) / col_cl (p x k): column-major, non-zero = member.  vals[j] (may be NULL) receives the per-bicluster
 * value (0 for empty ones), *bisil their mean over the non-empty biclusters.  The package is not in the reference tree:
 * restated from its published definition, parity unpinned. */
int resnmtf_data_bisil(resnmtf_data* data, const double* row_cl, const double* col_cl, int k, int method, double* vals,
                       double* bisil);
/* The same for the biclusters j with want[j] != 0 only (vals[j]; 0 for the others): the per-bicluster values do not
 * depend on each other, so the biclusters of one fit can be scored on different GPUs that each hold a copy of the view
 * and combined by the caller -- *n_live (may be NULL) receives the number of non-empty biclusters of the whole
 * clustering, the denominator of the mean. */
int resnmtf_data_bisil_part(resnmtf_data* data, const double* row_cl, const double* col_cl, int k, int method,
                            const int32_t* want, double* vals, int32_t* n_live);

/* ---- fan-out: independent fits over the GPUs of a pool ----------------------------------- */

/* One default apply_resnmtf() call is 66 convergence loops: per k of the sweep one fit and num_repeats shuffled refits
 * (R/main.r:270-299, R/obtain_bicl.r:31-42), then per stability resample the same again
 * (R/stability_analysis.r:302-338); the reference runs them nested and serially (its %dopar% over k is unreachable).
 * Each of them is a UNIT below.  A pool owns one context and -- during resnmtf_batch_run -- one native worker thread per
 * GPU; the views of a data set are uploaded once and copied GPU to GPU on first use. */
typedef struct resnmtf_pool resnmtf_pool;

/* devices == NULL: the first n_devices visible GPUs (n_devices <= 0: all of them). */
int resnmtf_pool_create(const int* devices, int n_devices, resnmtf_pool** out);
int resnmtf_pool_destroy(resnmtf_pool* pool);
int resnmtf_pool_size(resnmtf_pool* pool);
/* Context of GPU `gpu` of the pool (owned by the pool). */
resnmtf_ctx* resnmtf_pool_ctx(resnmtf_pool* pool, int gpu);
/* Registers the views of one data set (the `data` list of apply_resnmtf) under `key`.  _put: handles that already
 * live on one of the pool's contexts (the pool takes its own reference); _put_host: host matrices, uploaded to the first
 * GPU, with prep != 0 through make_non_neg_inner + matrix_normalisation (resnmtf_data_create_prepped). */
int resnmtf_pool_put(resnmtf_pool* pool, int key, int n_views, resnmtf_data* const* views);
int resnmtf_pool_put_host(resnmtf_pool* pool, int key, int n_views, const int64_t* n, const int64_t* p,
                          const double* const* x, const int64_t* ld, int prep, int32_t* was_negative);
/* View `view` of set `key` on GPU `gpu` of the pool (copied there on first use); borrowed, owned by the pool. */
int resnmtf_pool_get(resnmtf_pool* pool, int key, int gpu, int view, resnmtf_data** out);
int resnmtf_pool_drop(resnmtf_pool* pool, int key);

#define RESNMTF_DERIVE_NONE 0
#define RESNMTF_DERIVE_SUBSAMPLE 1 /* x[rows, cols] first (stability_repeat, R/stability_analysis.r:215-253)            */
#define RESNMTF_DERIVE_SHUFFLE 2   /* then shuffle_view on every view (obtain_shuffled_f, R/obtain_bicl.r:31-42)        */

typedef struct resnmtf_map {  /* one resnmtf_fit_set_shared_map call */
  int32_t kind, v, w;
  const int32_t* idx_v;
  const int32_t* idx_w;
  int64_t len;
} resnmtf_map;

/* One res_nmtf_inner() core: [derive the data] -> initial factors -> loop of R/main.r:50-109 -> normalisation_check.
 * Per-view arrays have one entry per view of the data set; host buffers are borrowed until resnmtf_batch_run returns. */
typedef struct resnmtf_unit {
  /* ---- inputs ---- */
  int32_t data_key;             /* data set registered with resnmtf_pool_put(_host)                                    */
  int32_t derive;               /* RESNMTF_DERIVE_* bits                                                                */
  const int32_t* const* rows;   /* SUBSAMPLE: per view, 0-based row / column indices and their counts                   */
  const int64_t* n_rows;
  const int32_t* const* cols;
  const int64_t* n_cols;
  uint64_t seed;                /* SHUFFLE: key of the permutation (view v uses a function of seed and v)               */
  int32_t renormalise;          /* SHUFFLE: re-prep the shuffled views as apply_resnmtf does (R/obtain_bicl.r:35)        */
  int32_t n_maps;
  const int32_t* k;             /* k_vec                                                                                */
  const double* const* init_f;  /* explicit initial factors per view (R/update_steps.r:49-60) -- all three or none:     */
  const double* const* init_s;  /*   none = the SVD initialisation on the device (resnmtf_data_svd_topk) with ...        */
  const double* const* init_g;
  const double* const* noise;   /* ... per view the k x k draw abs(mvrnorm(k, 0, 0.05 I)) of R/update_steps.r:96-99      */
                                /*   made by the caller in the reference's order (NULL: no noise term)                   */
  const double* phi;            /* n_views x n_views, symmetrised (init_rest_mats); NULL = zero                         */
  const double* xi;
  const double* psi;
  const resnmtf_map* maps;      /* shared-name index maps, n_maps of them                                               */
  int64_t n_iters;              /* as resnmtf_fit_run                                                                   */
  double tol;
  int64_t max_iters;
  int32_t err_mode, impl;       /* as resnmtf_fit_set_options                                                           */
  /* ---- outputs (host buffers of the caller; any pointer may be NULL) ---- */
  double* const* out_f;         /* normalised factors (normalisation_check, R/utils.r:176-195)                          */
  double* const* out_s;
  double* const* out_g;
  double* const* out_lambda;    /* lambda / mu as the loop left them (R/main.r:137-138)                                 */
  double* const* out_mu;
  double* errors;               /* All_Error (R/main.r:134): first min(errors_cap, n_errors) values                     */
  int64_t errors_cap;
  int64_t n_errors;
  int64_t iters;                /* sweeps done                                                                          */
  int32_t status;               /* RESNMTF_OK or the RESNMTF_E_* code of this unit                                      */
  int32_t gpu;                  /* GPU of the pool that ran it                                                          */
  double seconds;               /* wall time of the unit on its worker thread                                           */
  char message[200];            /* resnmtf_last_error() of the worker thread when status != 0                           */
} resnmtf_unit;

/* Runs the units on one native worker thread per GPU of the pool (no interpreter, no host language involved) -- the
 * fits of the resident data before the units that derive their own, each group longest first; SVD triplets that several
 * units share are computed once, by the home GPU's worker, while the other GPUs start on derived units -- and returns
 * when all are done: RESNMTF_OK, or the code of the first failed unit (every unit carries its
 * own status).  What a unit returns does not depend on the GPU that ran it or on the number of GPUs. */
int resnmtf_batch_run(resnmtf_pool* pool, resnmtf_unit* units, int n_units);
/* sizeof(resnmtf_unit) as this library was built: lets a binding verify its own declaration of the struct. */
int resnmtf_unit_size(void);

/* ---- row-sharded view across ranks (one process per GPU; NCCL all-reduce of the p x k partials) ---- */

/* Size of the opaque NCCL unique id and its creation on rank 0 (to be broadcast by the host). */
int resnmtf_comm_id_size(void);
int resnmtf_comm_id_create(void* id_out);
/* Joins the communicator (NCCL is resolved with dlopen at this point; RESNMTF_E_COMM when it is missing).
 * n_ranks == 1 is allowed and runs the complete sharded code path on one GPU.
 * After this, every view of every fit created on ctx is treated as ROW-SHARDED: n[v] passed to fit_create
 * is the local row count (whole 64-row panels except on the last rank), F rows are local, G/S/lambda/mu
 * are replicated; per sweep [X'F | F'F | colSums(F)] is all-reduced once (p*k + k*k + k doubles) and the
 * G/S/lambda/mu updates are computed redundantly on every rank; ||X||^2 and the direct residual are
 * all-reduced scalars.  phi/psi gather maps then refer to LOCAL rows of equally sharded views. */
int resnmtf_ctx_join(resnmtf_ctx* ctx, const void* id, int rank, int n_ranks);

#ifdef __cplusplus
}
#endif
#endif /* RESNMTF_B200_H */
