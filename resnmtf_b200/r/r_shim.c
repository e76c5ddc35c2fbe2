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

#include "resnmtf_b200.h"

static resnmtf_ctx* g_ctx = NULL;

static void fail_after_cleanup(resnmtf_fit* fit, int* iv, int* iw) {
  /* copy the message first: no longjmp through live C resources */
  char msg[512];
  snprintf(msg, sizeof msg, "%s", resnmtf_last_error());
  if (fit) resnmtf_fit_destroy(fit);
  free(iv);
  free(iw);
  Rf_error("%s", msg);
}

static void set_maps(resnmtf_fit* fit, int kind, SEXP maps, int n_v) {
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
      if (len > 0) {
        iv = (int*)malloc(sizeof(int) * len);
        iw = (int*)malloc(sizeof(int) * len);
        const int* p = INTEGER(m);
        for (R_xlen_t i = 0; i < len; ++i) {
          iv[i] = p[2 * i] - 1; /* R is 1-based */
          iw[i] = p[2 * i + 1] - 1;
        }
      }
      const int rc = resnmtf_fit_set_shared_map(fit, kind, v, w, iv, iw, (int64_t)len);
      if (rc != RESNMTF_OK) fail_after_cleanup(fit, iv, iw);
      free(iv);
      free(iw);
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
  for (int v = 0; v < n_v; ++v) {
    if (resnmtf_fit_set_data(fit, v, REAL(VECTOR_ELT(data, v)), n[v]) != RESNMTF_OK ||
        resnmtf_fit_set_factors(fit, v, REAL(VECTOR_ELT(init_f, v)), REAL(VECTOR_ELT(init_s, v)),
                                REAL(VECTOR_ELT(init_g, v)), REAL(VECTOR_ELT(lam, v)),
                                REAL(VECTOR_ELT(mu, v))) != RESNMTF_OK)
      fail_after_cleanup(fit, NULL, NULL);
  }
  if (resnmtf_fit_set_restrictions(fit, REAL(phi), REAL(xi), REAL(psi)) != RESNMTF_OK)
    fail_after_cleanup(fit, NULL, NULL);
  set_maps(fit, RESNMTF_MAP_ROW, row_maps, n_v);
  set_maps(fit, RESNMTF_MAP_COL, col_maps, n_v);

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
      R_CheckUserInterrupt();
    } while (1);
  }
  if (rc == RESNMTF_E_NAN) {
    resnmtf_fit_destroy(fit);
    Rf_error("missing value where TRUE/FALSE needed"); /* what while(NA) raises at R/main.r:55 */
  }
  if (rc != RESNMTF_OK) fail_after_cleanup(fit, NULL, NULL);

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
    if (pass == 1 && resnmtf_fit_normalise(fit) != RESNMTF_OK) fail_after_cleanup(fit, NULL, NULL);
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
      if (resnmtf_fit_get_factors(fit, v, REAL(f), REAL(s), REAL(g), pl, pm) != RESNMTF_OK) {
        UNPROTECT(1);
        fail_after_cleanup(fit, NULL, NULL);
      }
    }
  }
  resnmtf_fit_destroy(fit);
  UNPROTECT(1);
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

static const R_CallMethodDef call_methods[] = {{"C_resnmtf_fit", (DL_FUNC)&C_resnmtf_fit, 12},
                                               {"C_resnmtf_jsd_pairs", (DL_FUNC)&C_resnmtf_jsd_pairs, 5},
                                               {NULL, NULL, 0}};

void R_init_resnmtf(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
