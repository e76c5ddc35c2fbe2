/*
 * .Call shim between the R package `resnmtf` and libresnmtf_b200.so (include/resnmtf_b200.h).
 * NOT compiled in this repository's image (no R, no Rinternals.h); shipped for an R install:
 *   R CMD SHLIB r_shim.c -I../../include -L.. -lresnmtf_b200
 * It replaces the loop body of res_nmtf_inner(), R/main.r:50-109 -- see INTEGRATION.md.
 *
 * C_resnmtf_fit(data, init_f, init_s, init_g, lam, mu, phi, xi, psi, row_maps, col_maps, n_iters)
 *   data, init_*  : lists of double matrices (column-major, exactly R's storage)
 *   lam, mu       : lists of double vectors
 *   phi, xi, psi  : n_v x n_v double matrices, already symmetrised by init_rest_mats()
 *   row_maps/col_maps : list over v of list over w of either NULL (never set), integer(0) (NA pair) or a
 *                   2-row integer matrix rbind(idx_v, idx_w), 1-based (built in R with match())
 *   n_iters       : integer scalar, NA -> run to convergence (tol 1e-6)
 * returns list(F = list, S = list, G = list, lam = list, mu = list, total_err = numeric,
 *              Fn = list, Sn = list, Gn = list)   (raw factors, and the normalisation_check()-ed ones)
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdlib.h>
#include <string.h>

#include "resnmtf_b200.h"

static resnmtf_ctx* g_ctx = NULL;
static resnmtf_pool* g_pool = NULL; /* one context + (during a batch) one native worker thread per GPU */

/* The fit handle of a call lives in an R external pointer with a C finaliser for the whole .Call: whatever leaves the
 * call by longjmp -- Rf_error() here, or R_CheckUserInterrupt() on Ctrl-C -- leaves the handle to the garbage collector,
 * which destroys the fit (and frees its device memory) at the next collection.  The normal exit destroys it at once. */
static void fit_finalizer(SEXP guard) {
  resnmtf_fit* fit = (resnmtf_fit*)R_ExternalPtrAddr(guard);
  if (fit) {
    resnmtf_fit_destroy(fit);
    R_ClearExternalPtr(guard);
  }
}

static void fail_after_cleanup(SEXP guard) {
  /* copy the message first: the destroy below may overwrite the thread-local text */
  char msg[512];
  snprintf(msg, sizeof msg, "%s", resnmtf_last_error());
  fit_finalizer(guard);
  Rf_error("%s", msg);
}

static void set_maps(resnmtf_fit* fit, SEXP guard, int kind, SEXP maps, int n_v) {
  if (maps == R_NilValue) return; /* R's NULL: nothing set (quirk of R/main.r:312) */
  for (int v = 0; v < n_v; ++v) {
    SEXP mv = VECTOR_ELT(maps, v);
    if (mv == R_NilValue) continue;
    for (int w = 0; w < n_v; ++w) {
      if (w == v) continue;
      SEXP m = VECTOR_ELT(mv, w);
      if (m == R_NilValue) continue;
      const R_xlen_t len = XLENGTH(m) / 2;
      int *iv = NULL, *iw = NULL;
      if (len > 0) { /* R_alloc: released by R when the .Call returns, on every exit path */
        iv = (int*)R_alloc((size_t)len, sizeof(int));
        iw = (int*)R_alloc((size_t)len, sizeof(int));
        const int* p = INTEGER(m);
        for (R_xlen_t i = 0; i < len; ++i) {
          iv[i] = p[2 * i] - 1; /* R is 1-based */
          iw[i] = p[2 * i + 1] - 1;
        }
      }
      const int rc = resnmtf_fit_set_shared_map(fit, kind, v, w, iv, iw, (int64_t)len);
      if (rc != RESNMTF_OK) fail_after_cleanup(guard);
    }
  }
}

SEXP C_resnmtf_fit(SEXP data, SEXP init_f, SEXP init_s, SEXP init_g, SEXP lam, SEXP mu, SEXP phi, SEXP xi,
                   SEXP psi, SEXP row_maps, SEXP col_maps, SEXP n_iters) {
  const int n_v = LENGTH(data);
  if (!g_ctx && resnmtf_ctx_create(-1, &g_ctx) != RESNMTF_OK) Rf_error("%s", resnmtf_last_error());
  int64_t* n = (int64_t*)R_alloc(n_v, sizeof(int64_t));
  int64_t* p = (int64_t*)R_alloc(n_v, sizeof(int64_t));
  int32_t* k = (int32_t*)R_alloc(n_v, sizeof(int32_t));
  for (int v = 0; v < n_v; ++v) {
    n[v] = Rf_nrows(VECTOR_ELT(data, v));
    p[v] = Rf_ncols(VECTOR_ELT(data, v));
    k[v] = Rf_ncols(VECTOR_ELT(init_f, v));
  }
  resnmtf_fit* fit = NULL;
  if (resnmtf_fit_create(g_ctx, n_v, n, p, k, &fit) != RESNMTF_OK) Rf_error("%s", resnmtf_last_error());
  SEXP guard = PROTECT(R_MakeExternalPtr(fit, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(guard, fit_finalizer, TRUE);
  for (int v = 0; v < n_v; ++v) {
    if (resnmtf_fit_set_data(fit, v, REAL(VECTOR_ELT(data, v)), n[v]) != RESNMTF_OK ||
        resnmtf_fit_set_factors(fit, v, REAL(VECTOR_ELT(init_f, v)), REAL(VECTOR_ELT(init_s, v)),
                                REAL(VECTOR_ELT(init_g, v)), REAL(VECTOR_ELT(lam, v)),
                                REAL(VECTOR_ELT(mu, v))) != RESNMTF_OK)
      fail_after_cleanup(guard);
  }
  if (resnmtf_fit_set_restrictions(fit, REAL(phi), REAL(xi), REAL(psi)) != RESNMTF_OK) fail_after_cleanup(guard);
  set_maps(fit, guard, RESNMTF_MAP_ROW, row_maps, n_v);
  set_maps(fit, guard, RESNMTF_MAP_COL, col_maps, n_v);

  const int ni = Rf_asInteger(n_iters);
  int rc;
  if (ni != NA_INTEGER) {
    rc = resnmtf_fit_run(fit, ni, 1.0e-6, 0, NULL);
  } else {
    /* chunks of sweeps so that Ctrl-C works; the device keeps the stop-rule state between calls */
    int64_t done = 0;
    do {
      rc = resnmtf_fit_run(fit, -1, 1.0e-6, 256, &done);
      resnmtf_counters c;
      resnmtf_fit_get_counters(fit, &c);
      if (rc != RESNMTF_OK || c.converged) break;
      R_CheckUserInterrupt(); /* may longjmp: the guard's finaliser then destroys the fit */
    } while (1);
  }
  if (rc == RESNMTF_E_NAN) {
    fit_finalizer(guard);
    Rf_error("missing value where TRUE/FALSE needed"); /* what while(NA) raises at R/main.r:55 */
  }
  if (rc != RESNMTF_OK) fail_after_cleanup(guard);

  int64_t n_err = 0;
  resnmtf_fit_get_errors(fit, NULL, 0, &n_err);
  const char* names[] = {"F", "S", "G", "lam", "mu", "total_err", "Fn", "Sn", "Gn", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
  SEXP lists[9];
  for (int i = 0; i < 9; ++i) {
    if (i == 5) continue;
    lists[i] = Rf_allocVector(VECSXP, n_v);
    SET_VECTOR_ELT(out, i, lists[i]);
  }
  SEXP errs = Rf_allocVector(REALSXP, n_err);
  SET_VECTOR_ELT(out, 5, errs);
  resnmtf_fit_get_errors(fit, REAL(errs), n_err, &n_err);
  for (int pass = 0; pass < 2; ++pass) { /* raw factors, then normalisation_check() on the device */
    if (pass == 1 && resnmtf_fit_normalise(fit) != RESNMTF_OK) fail_after_cleanup(guard);
    for (int v = 0; v < n_v; ++v) {
      SEXP f = Rf_allocMatrix(REALSXP, (int)n[v], k[v]);
      SET_VECTOR_ELT(lists[pass ? 6 : 0], v, f);
      SEXP s = Rf_allocMatrix(REALSXP, k[v], k[v]);
      SET_VECTOR_ELT(lists[pass ? 7 : 1], v, s);
      SEXP g = Rf_allocMatrix(REALSXP, (int)p[v], k[v]);
      SET_VECTOR_ELT(lists[pass ? 8 : 2], v, g);
      double *pl = NULL, *pm = NULL;
      if (pass == 0) {
        SEXP l = Rf_allocVector(REALSXP, k[v]);
        SET_VECTOR_ELT(lists[3], v, l);
        SEXP m = Rf_allocVector(REALSXP, k[v]);
        SET_VECTOR_ELT(lists[4], v, m);
        pl = REAL(l);
        pm = REAL(m);
      }
      if (resnmtf_fit_get_factors(fit, v, REAL(f), REAL(s), REAL(g), pl, pm) != RESNMTF_OK) fail_after_cleanup(guard);
    }
  }
  fit_finalizer(guard); /* destroys the fit now and clears the pointer */
  UNPROTECT(2);
  return out;
}

/* jsd_calc() (R/utils.r:95-106) for a batch of column pairs: cols is the n x m matrix of factor columns, bw / vmax
 * their bw.nrd0 and maxima, pair_a / pair_b 1-based column indices.  Returns the numeric vector of JSD values in pair
 * order (resnmtf_jsd_pairs). */
SEXP C_resnmtf_jsd_pairs(SEXP cols, SEXP bw, SEXP vmax, SEXP pair_a, SEXP pair_b) {
  if (!g_ctx && resnmtf_ctx_create(-1, &g_ctx) != RESNMTF_OK) Rf_error("%s", resnmtf_last_error());
  if (!Rf_isReal(cols) || !Rf_isMatrix(cols)) Rf_error("cols must be a numeric matrix");
  const int64_t n = Rf_nrows(cols);
  const int m = Rf_ncols(cols);
  const R_xlen_t np = XLENGTH(pair_a);
  if (XLENGTH(pair_b) != np || XLENGTH(bw) != m || XLENGTH(vmax) != m) Rf_error("inconsistent argument lengths");
  int32_t* a = (int32_t*)R_alloc((size_t)np + 1, sizeof(int32_t));
  int32_t* b = (int32_t*)R_alloc((size_t)np + 1, sizeof(int32_t));
  for (R_xlen_t i = 0; i < np; ++i) { /* R is 1-based */
    a[i] = INTEGER(pair_a)[i] - 1;
    b[i] = INTEGER(pair_b)[i] - 1;
  }
  SEXP out = PROTECT(Rf_allocVector(REALSXP, np));
  const int rc = resnmtf_jsd_pairs(g_ctx, REAL(cols), n, m, n, REAL(bw), REAL(vmax), a, b, (int64_t)np, REAL(out));
  UNPROTECT(1);
  if (rc != RESNMTF_OK) Rf_error("%s", resnmtf_last_error());
  return out;
}


/* ---- fan-out: the independent fits of one apply_resnmtf() call as ONE .Call ---------------------------------------
 * C_resnmtf_batch(data, prep, units, n_gpus)
 *   data   : list of double matrices (the views); uploaded once, with prep = TRUE through make_non_neg_inner() +
 *            matrix_normalisation() (R/utils.r:20-27, 86-88) on the device
 *   units  : list of lists, one per res_nmtf_inner() core -- the k-sweep fits (R/main.r:270-299), the shuffled refits
 *            (R/obtain_bicl.r:31-42) and the stability resamples (R/stability_analysis.r:302-338).  Elements by name:
 *              k (integer, one per view); noise (list of k x k matrices: the abs(mvrnorm()) draw of
 *              R/update_steps.r:96-99, made in R in the reference's order) or init_f / init_s / init_g (lists);
 *              rows / cols (lists of 1-based integer vectors: sub-sample); shuffle_seed (numeric); renormalise (logical);
 *              phi / xi / psi (symmetrised matrices); row_maps / col_maps (as for C_resnmtf_fit); n_iters (integer, NA:
 *              run to convergence)
 *   n_gpus : integer, <= 0: every visible GPU
 * returns a list with one element per unit: list(F, S, G (normalised, lists over views), lam, mu, total_err, iters,
 * gpu, seconds, status, message).  The library's worker threads never touch the R API. */
static SEXP elt(SEXP list, const char* name) {
  SEXP names = Rf_getAttrib(list, R_NamesSymbol);
  if (names == R_NilValue) return R_NilValue;
  for (R_xlen_t i = 0; i < XLENGTH(list); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
  return R_NilValue;
}

static const double* const* matrix_list(SEXP list, int n_v) { /* NULL list -> NULL; NULL entries stay NULL */
  if (list == R_NilValue) return NULL;
  const double** out = (const double**)R_alloc((size_t)n_v, sizeof(double*));
  for (int v = 0; v < n_v; ++v) {
    SEXP m = VECTOR_ELT(list, v);
    out[v] = (m == R_NilValue) ? NULL : REAL(m);
  }
  return (const double* const*)out;
}

static int count_maps(SEXP maps, int n_v) {
  int c = 0;
  if (maps == R_NilValue) return 0;
  for (int v = 0; v < n_v; ++v) {
    SEXP mv = VECTOR_ELT(maps, v);
    if (mv == R_NilValue) continue;
    for (int w = 0; w < n_v; ++w)
      if (w != v && VECTOR_ELT(mv, w) != R_NilValue) ++c;
  }
  return c;
}

static int fill_maps(resnmtf_map* dst, int kind, SEXP maps, int n_v) {
  int c = 0;
  if (maps == R_NilValue) return 0;
  for (int v = 0; v < n_v; ++v) {
    SEXP mv = VECTOR_ELT(maps, v);
    if (mv == R_NilValue) continue;
    for (int w = 0; w < n_v; ++w) {
      if (w == v) continue;
      SEXP m = VECTOR_ELT(mv, w);
      if (m == R_NilValue) continue;
      const R_xlen_t len = XLENGTH(m) / 2;
      int32_t* iv = (int32_t*)R_alloc((size_t)len + 1, sizeof(int32_t));
      int32_t* iw = (int32_t*)R_alloc((size_t)len + 1, sizeof(int32_t));
      for (R_xlen_t i = 0; i < len; ++i) {
        iv[i] = INTEGER(m)[2 * i] - 1;
        iw[i] = INTEGER(m)[2 * i + 1] - 1;
      }
      dst[c].kind = kind;
      dst[c].v = v;
      dst[c].w = w;
      dst[c].idx_v = iv;
      dst[c].idx_w = iw;
      dst[c].len = (int64_t)len;
      ++c;
    }
  }
  return c;
}

SEXP C_resnmtf_batch(SEXP data, SEXP prep, SEXP units, SEXP n_gpus) {
  const int n_v = LENGTH(data);
  const int n_u = LENGTH(units);
  const int64_t err_cap = 1 << 16;
  if ((size_t)resnmtf_unit_size() != sizeof(resnmtf_unit)) Rf_error("resnmtf_unit layout differs from the library's");
  if (!g_pool && resnmtf_pool_create(NULL, Rf_asInteger(n_gpus), &g_pool) != RESNMTF_OK)
    Rf_error("%s", resnmtf_last_error());
  int64_t* n = (int64_t*)R_alloc((size_t)n_v, sizeof(int64_t));
  int64_t* p = (int64_t*)R_alloc((size_t)n_v, sizeof(int64_t));
  const double** x = (const double**)R_alloc((size_t)n_v, sizeof(double*));
  for (int v = 0; v < n_v; ++v) {
    n[v] = Rf_nrows(VECTOR_ELT(data, v));
    p[v] = Rf_ncols(VECTOR_ELT(data, v));
    x[v] = REAL(VECTOR_ELT(data, v));
  }
  const int key = 1;
  int32_t was_negative = 0;
  if (resnmtf_pool_put_host(g_pool, key, n_v, n, p, x, NULL, Rf_asLogical(prep) == TRUE, &was_negative) != RESNMTF_OK)
    Rf_error("%s", resnmtf_last_error());
  if (was_negative) Rf_warning("Matrix is not non-negative. Has been made non-negative."); /* R/utils.r:24 */

  resnmtf_unit* cu = (resnmtf_unit*)R_alloc((size_t)n_u, sizeof(resnmtf_unit));
  memset(cu, 0, (size_t)n_u * sizeof(resnmtf_unit));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, n_u));
  const char* names[] = {"F", "S", "G", "lam", "mu", "total_err", "iters", "gpu", "seconds", "status", "message", ""};
  for (int u = 0; u < n_u; ++u) {
    SEXP ru = VECTOR_ELT(units, u);
    resnmtf_unit* c = cu + u;
    SEXP kk = elt(ru, "k");
    if (LENGTH(kk) != n_v) {
      resnmtf_pool_drop(g_pool, key);
      Rf_error("unit %d: k must have one entry per view", u + 1);
    }
    int32_t* k = (int32_t*)R_alloc((size_t)n_v, sizeof(int32_t));
    for (int v = 0; v < n_v; ++v) k[v] = INTEGER(kk)[v];
    c->data_key = key;
    c->k = k;
    int64_t* un = (int64_t*)R_alloc((size_t)n_v, sizeof(int64_t)); /* shape of the unit's (derived) views */
    int64_t* up = (int64_t*)R_alloc((size_t)n_v, sizeof(int64_t));
    for (int v = 0; v < n_v; ++v) {
      un[v] = n[v];
      up[v] = p[v];
    }
    SEXP rows = elt(ru, "rows"), cols = elt(ru, "cols");
    if (rows != R_NilValue && cols != R_NilValue) { /* stability_repeat: x[rows, cols], 1-based in R */
      const int32_t** pr = (const int32_t**)R_alloc((size_t)n_v, sizeof(int32_t*));
      const int32_t** pc = (const int32_t**)R_alloc((size_t)n_v, sizeof(int32_t*));
      for (int v = 0; v < n_v; ++v) {
        SEXP rv = VECTOR_ELT(rows, v), cv = VECTOR_ELT(cols, v);
        un[v] = XLENGTH(rv);
        up[v] = XLENGTH(cv);
        int32_t* ir = (int32_t*)R_alloc((size_t)un[v] + 1, sizeof(int32_t));
        int32_t* ic = (int32_t*)R_alloc((size_t)up[v] + 1, sizeof(int32_t));
        for (int64_t i = 0; i < un[v]; ++i) ir[i] = INTEGER(rv)[i] - 1;
        for (int64_t i = 0; i < up[v]; ++i) ic[i] = INTEGER(cv)[i] - 1;
        pr[v] = ir;
        pc[v] = ic;
      }
      c->rows = pr;
      c->cols = pc;
      c->n_rows = un;
      c->n_cols = up;
      c->derive |= RESNMTF_DERIVE_SUBSAMPLE;
    }
    SEXP seed = elt(ru, "shuffle_seed");
    if (seed != R_NilValue) {
      c->derive |= RESNMTF_DERIVE_SHUFFLE;
      c->seed = (uint64_t)Rf_asReal(seed);
      SEXP rn = elt(ru, "renormalise");
      c->renormalise = (rn == R_NilValue) ? 1 : (Rf_asLogical(rn) == TRUE);
    }
    c->init_f = matrix_list(elt(ru, "init_f"), n_v);
    c->init_s = matrix_list(elt(ru, "init_s"), n_v);
    c->init_g = matrix_list(elt(ru, "init_g"), n_v);
    c->noise = matrix_list(elt(ru, "noise"), n_v);
    SEXP ph = elt(ru, "phi"), xi = elt(ru, "xi"), ps = elt(ru, "psi");
    c->phi = (ph == R_NilValue) ? NULL : REAL(ph);
    c->xi = (xi == R_NilValue) ? NULL : REAL(xi);
    c->psi = (ps == R_NilValue) ? NULL : REAL(ps);
    SEXP rm = elt(ru, "row_maps"), cm = elt(ru, "col_maps");
    const int nm = count_maps(rm, n_v) + count_maps(cm, n_v);
    if (nm) {
      resnmtf_map* maps = (resnmtf_map*)R_alloc((size_t)nm, sizeof(resnmtf_map));
      int at = fill_maps(maps, RESNMTF_MAP_ROW, rm, n_v);
      at += fill_maps(maps + at, RESNMTF_MAP_COL, cm, n_v);
      c->maps = maps;
      c->n_maps = at;
    }
    SEXP ni = elt(ru, "n_iters");
    const int iters = (ni == R_NilValue) ? NA_INTEGER : Rf_asInteger(ni);
    c->n_iters = (iters == NA_INTEGER) ? -1 : iters;
    c->tol = 1.0e-6; /* R/main.r:55 */
    c->max_iters = 0;
    c->err_mode = RESNMTF_ERR_AUTO;
    c->impl = RESNMTF_IMPL_AUTO;
    /* outputs: R objects allocated here on the main thread, filled by the library */
    SEXP res = Rf_mkNamed(VECSXP, names);
    SET_VECTOR_ELT(out, u, res);
    double** of = (double**)R_alloc((size_t)n_v, sizeof(double*));
    double** os = (double**)R_alloc((size_t)n_v, sizeof(double*));
    double** og = (double**)R_alloc((size_t)n_v, sizeof(double*));
    double** ol = (double**)R_alloc((size_t)n_v, sizeof(double*));
    double** om = (double**)R_alloc((size_t)n_v, sizeof(double*));
    for (int i = 0; i < 5; ++i) SET_VECTOR_ELT(res, i, Rf_allocVector(VECSXP, n_v));
    for (int v = 0; v < n_v; ++v) {
      SEXP f = Rf_allocMatrix(REALSXP, (int)un[v], k[v]);
      SET_VECTOR_ELT(VECTOR_ELT(res, 0), v, f);
      SEXP s = Rf_allocMatrix(REALSXP, k[v], k[v]);
      SET_VECTOR_ELT(VECTOR_ELT(res, 1), v, s);
      SEXP g = Rf_allocMatrix(REALSXP, (int)up[v], k[v]);
      SET_VECTOR_ELT(VECTOR_ELT(res, 2), v, g);
      SEXP l = Rf_allocVector(REALSXP, k[v]);
      SET_VECTOR_ELT(VECTOR_ELT(res, 3), v, l);
      SEXP m = Rf_allocVector(REALSXP, k[v]);
      SET_VECTOR_ELT(VECTOR_ELT(res, 4), v, m);
      of[v] = REAL(f);
      os[v] = REAL(s);
      og[v] = REAL(g);
      ol[v] = REAL(l);
      om[v] = REAL(m);
    }
    c->out_f = of;
    c->out_s = os;
    c->out_g = og;
    c->out_lambda = ol;
    c->out_mu = om;
    c->errors_cap = (c->n_iters >= 0) ? c->n_iters : err_cap;
    c->errors = (double*)R_alloc((size_t)c->errors_cap + 1, sizeof(double));
  }

  const int rc = resnmtf_batch_run(g_pool, cu, n_u); /* blocks; one native worker thread per GPU */
  resnmtf_pool_drop(g_pool, key);
  int nan_unit = 0;
  for (int u = 0; u < n_u; ++u) {
    const resnmtf_unit* c = cu + u;
    SEXP res = VECTOR_ELT(out, u);
    const int64_t ne = c->n_errors < c->errors_cap ? c->n_errors : c->errors_cap;
    SEXP errs = Rf_allocVector(REALSXP, (R_xlen_t)ne);
    SET_VECTOR_ELT(res, 5, errs);
    if (ne > 0) memcpy(REAL(errs), c->errors, (size_t)ne * sizeof(double));
    SET_VECTOR_ELT(res, 6, Rf_ScalarReal((double)c->iters));
    SET_VECTOR_ELT(res, 7, Rf_ScalarInteger(c->gpu));
    SET_VECTOR_ELT(res, 8, Rf_ScalarReal(c->seconds));
    SET_VECTOR_ELT(res, 9, Rf_ScalarInteger(c->status));
    SET_VECTOR_ELT(res, 10, Rf_mkString(c->message));
    if (c->status == RESNMTF_E_NAN) nan_unit = u + 1;
  }
  UNPROTECT(1);
  if (nan_unit) Rf_error("missing value where TRUE/FALSE needed"); /* while(NA) of R/main.r:55, unit nan_unit */
  if (rc != RESNMTF_OK) Rf_error("%s", resnmtf_last_error());
  return out;
}

/* ---- SVD initialisation and bisilhouette on the device ------------------------------------
 * C_resnmtf_svd_topk(x, k): what init_mats_inner() takes from svd(x) (R/update_steps.r:92-95) -- list(u = |U[, 1:k]|,
 * d = d[1:k], v = |V[, 1:k]|) -- from resnmtf_data_svd_topk (Gram matrix of the smaller side + filtered subspace
 * iteration on the GPU instead of a full LAPACK svd).
 * C_resnmtf_bisil(x, row_clusters, col_clusters, method): bisilhouette::bisilhouette(x, row_clusters, col_clusters,
 * method)$bisil as obtain_biclusters() calls it (R/obtain_bicl.r:190-199); method 0 euclidean, 1 manhattan, 2 cosine.
 * Returns list(bisil = , vals = per-bicluster values). */
static void data_finalizer(SEXP guard) {
  resnmtf_data* d = (resnmtf_data*)R_ExternalPtrAddr(guard);
  if (d) {
    resnmtf_data_destroy(d);
    R_ClearExternalPtr(guard);
  }
}

static SEXP upload_view(SEXP x, resnmtf_data** out) { /* returns the PROTECTed guard of the handle */
  if (!Rf_isReal(x) || !Rf_isMatrix(x)) Rf_error("x must be a numeric matrix");
  if (!g_ctx && resnmtf_ctx_create(-1, &g_ctx) != RESNMTF_OK) Rf_error("%s", resnmtf_last_error());
  resnmtf_data* d = NULL;
  if (resnmtf_data_create(g_ctx, Rf_nrows(x), Rf_ncols(x), REAL(x), Rf_nrows(x), &d) != RESNMTF_OK)
    Rf_error("%s", resnmtf_last_error());
  SEXP guard = PROTECT(R_MakeExternalPtr(d, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(guard, data_finalizer, TRUE);
  *out = d;
  return guard;
}

SEXP C_resnmtf_svd_topk(SEXP x, SEXP k) {
  const int kk = Rf_asInteger(k);
  resnmtf_data* d = NULL;
  SEXP guard = upload_view(x, &d);
  const char* names[] = {"u", "d", "v", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
  SEXP u = Rf_allocMatrix(REALSXP, Rf_nrows(x), kk);
  SET_VECTOR_ELT(out, 0, u);
  SEXP dd = Rf_allocVector(REALSXP, kk);
  SET_VECTOR_ELT(out, 1, dd);
  SEXP v = Rf_allocMatrix(REALSXP, Rf_ncols(x), kk);
  SET_VECTOR_ELT(out, 2, v);
  const int rc = resnmtf_data_svd_topk(d, kk, REAL(u), REAL(dd), REAL(v));
  data_finalizer(guard);
  UNPROTECT(2);
  if (rc != RESNMTF_OK) Rf_error("%s", resnmtf_last_error());
  return out;
}

SEXP C_resnmtf_bisil(SEXP x, SEXP row_cl, SEXP col_cl, SEXP method) {
  if (!Rf_isReal(row_cl) || !Rf_isReal(col_cl) || Rf_nrows(row_cl) != Rf_nrows(x) || Rf_nrows(col_cl) != Rf_ncols(x) ||
      Rf_ncols(row_cl) != Rf_ncols(col_cl))
    Rf_error("row_clusters must be n x k and col_clusters p x k numeric matrices");
  const int kk = Rf_ncols(row_cl);
  resnmtf_data* d = NULL;
  SEXP guard = upload_view(x, &d);
  const char* names[] = {"bisil", "vals", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
  SEXP vals = Rf_allocVector(REALSXP, kk);
  SET_VECTOR_ELT(out, 1, vals);
  double bisil = 0.0;
  const int rc = resnmtf_data_bisil(d, REAL(row_cl), REAL(col_cl), kk, Rf_asInteger(method), REAL(vals), &bisil);
  SET_VECTOR_ELT(out, 0, Rf_ScalarReal(bisil));
  data_finalizer(guard);
  UNPROTECT(2);
  if (rc != RESNMTF_OK) Rf_error("%s", resnmtf_last_error());
  return out;
}

static const R_CallMethodDef call_methods[] = {{"C_resnmtf_fit", (DL_FUNC)&C_resnmtf_fit, 12},
                                               {"C_resnmtf_jsd_pairs", (DL_FUNC)&C_resnmtf_jsd_pairs, 5},
                                               {"C_resnmtf_batch", (DL_FUNC)&C_resnmtf_batch, 4},
                                               {"C_resnmtf_svd_topk", (DL_FUNC)&C_resnmtf_svd_topk, 2},
                                               {"C_resnmtf_bisil", (DL_FUNC)&C_resnmtf_bisil, 4},
                                               {NULL, NULL, 0}};

void R_init_resnmtf(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}

/* library.dynam.unload() / detach(unload = TRUE): the context (stream, memory pool, workspaces) goes with the DLL */
void R_unload_resnmtf(DllInfo* dll) {
  (void)dll;
  if (g_pool) {
    resnmtf_pool_destroy(g_pool);
    g_pool = NULL;
  }
  if (g_ctx) {
    resnmtf_ctx_destroy(g_ctx);
    g_ctx = NULL;
  }
}
