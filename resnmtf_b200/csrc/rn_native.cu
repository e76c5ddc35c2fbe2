// Host side of the data-handle operations and of the SVD initialisation (include/resnmtf_b200.h, sections "data
// handles" and "SVD initialisation"): everything apply_resnmtf() does to a view around the update loop, on the device
// and behind the C ABI -- prep (R/utils.r:20-27, 86-88), shuffles (R/obtain_bicl.r:11-22), sub-samples
// (R/stability_analysis.r:215-253), the top singular triplets init_mats_inner() takes from svd(x)
// (R/update_steps.r:92-95).  Kernels: rn_data.cuh, rn_linalg.cuh.  No cuBLAS / cuSOLVER, no CPU fallback.
#include "rn_host.h"
#include "rn_linalg.cuh"
#include "rn_bisil.cuh"
#include "rn_dense.h"

namespace {

inline int rn_blocks(int64_t work, int threads = 256, int cap = 1 << 16) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((work + threads - 1) / threads, cap));
}

// RAII device scratch from the context's pool (stream-ordered)
struct DevBuf {
  resnmtf_ctx* ctx = nullptr;
  void* p = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { reset(); }
  void reset() {
    if (p) rn_dev_free(ctx, p);
    p = nullptr;
  }
  cudaError_t alloc(resnmtf_ctx* c, size_t bytes, bool zero) {
    reset();
    ctx = c;
    cudaError_t e = rn_dev_alloc(c, &p, std::max<size_t>(bytes, 16));
    if (e == cudaSuccess && zero) e = cudaMemsetAsync(p, 0, std::max<size_t>(bytes, 16), c->stream);
    return e;
  }
  double* d() const { return static_cast<double*>(p); }
};

// column statistic of a whole view: out[pp] (device) = sum / min / sum of squares of every column, fixed order
int col_stat(resnmtf_ctx* ctx, const double* X, int64_t ldx, int64_t pp, int op, double* out_dev) {
  const int tiles = (int)(ldx / RN_ROW_TILE);
  DevBuf part;
  RN_CUDA(part.alloc(ctx, (size_t)tiles * pp * sizeof(double), false));
  dim3 grid((unsigned)((pp + 127) / 128), (unsigned)tiles);
  rn_col_stats<<<grid, 128, 0, ctx->stream>>>(X, pp, part.d(), op);
  rn_col_combine<<<(unsigned)((pp + 127) / 128), 128, 0, ctx->stream>>>(part.d(), pp, tiles, out_dev, op);
  RN_CUDA(cudaGetLastError());
  return RESNMTF_OK;
}

// does the view have an all-zero row or column?  (the entries are non-negative where this is asked)
int has_zero_line(resnmtf_data* d, bool* out) {
  resnmtf_ctx* ctx = d->ctx;
  DevBuf cs, rs;
  RN_CUDA(cs.alloc(ctx, (size_t)d->pp * sizeof(double), false));
  RN_CUDA(rs.alloc(ctx, (size_t)d->ldx * sizeof(double), false));
  int rc = col_stat(ctx, d->X, d->ldx, d->pp, RN_STAT_SUM, cs.d());
  if (rc) return rc;
  rn_row_sums<<<(unsigned)(d->ldx / RN_ROW_TILE), 256, 0, ctx->stream>>>(d->X, d->p, d->pp, rs.d());
  RN_CUDA(cudaGetLastError());
  std::vector<double> hc((size_t)d->p), hr((size_t)d->n);
  RN_CUDA(cudaMemcpyAsync(hc.data(), cs.d(), hc.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  RN_CUDA(cudaMemcpyAsync(hr.data(), rs.d(), hr.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  RN_CUDA(cudaStreamSynchronize(ctx->stream));
  bool z = false;
  for (double v : hc) z = z || v == 0.0;
  for (double v : hr) z = z || v == 0.0;
  *out = z;
  return RESNMTF_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// data handles
// ------------------------------------------------------------------------------------------------------------------

extern "C" int resnmtf_data_shape(resnmtf_data* data, int64_t* n, int64_t* p) {
  RN_CHECK(data != nullptr, RESNMTF_E_INVALID, "resnmtf_data_shape: data is NULL");
  if (n) *n = data->n;
  if (p) *p = data->p;
  return RESNMTF_OK;
}

// make_non_neg_inner (column-wise shift by |min(0, min(col))|) and matrix_normalisation (L1 column normalisation)
// in place on a handle; was_negative reports whether the warning of R/utils.r:24 is due.
static int prep_in_place(resnmtf_data* d, int32_t* was_negative) {
  resnmtf_ctx* ctx = d->ctx;
  DevBuf cmin, csum;
  RN_CUDA(cmin.alloc(ctx, (size_t)d->pp * sizeof(double), false));
  RN_CUDA(csum.alloc(ctx, (size_t)d->pp * sizeof(double), false));
  int rc = col_stat(ctx, d->X, d->ldx, d->pp, RN_STAT_MIN, cmin.d());
  if (rc) return rc;
  const int blocks = rn_blocks(d->ldx * d->p);
  rn_shift_scale<<<blocks, 256, 0, ctx->stream>>>(d->X, d->n, d->p, d->pp, cmin.d(), nullptr);
  if ((rc = col_stat(ctx, d->X, d->ldx, d->pp, RN_STAT_SUM, csum.d()))) return rc;
  rn_shift_scale<<<blocks, 256, 0, ctx->stream>>>(d->X, d->n, d->p, d->pp, nullptr, csum.d());
  RN_CUDA(cudaGetLastError());
  if (was_negative) {
    std::vector<double> h((size_t)d->p);
    RN_CUDA(cudaMemcpyAsync(h.data(), cmin.d(), h.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    RN_CUDA(cudaStreamSynchronize(ctx->stream));
    int32_t neg = 0;
    for (double v : h) neg |= (v < 0.0) ? 1 : 0;
    *was_negative = neg;
  }
  return RESNMTF_OK;
}

extern "C" int resnmtf_data_create_prepped(resnmtf_ctx* ctx, int64_t n, int64_t p, const double* x, int64_t ld,
                                           int32_t* was_negative, resnmtf_data** out) {
  RN_CHECK(out != nullptr, RESNMTF_E_INVALID, "resnmtf_data_create_prepped: out is NULL");
  resnmtf_data* d = nullptr;
  int rc = resnmtf_data_create(ctx, n, p, x, ld, &d);
  if (rc) return rc;
  if ((rc = prep_in_place(d, was_negative)) || (rc = rn_data_seal(d))) {
    resnmtf_data_destroy(d);
    return rc;
  }
  *out = d;
  return RESNMTF_OK;
}

extern "C" int resnmtf_data_download(resnmtf_data* data, double* x, int64_t ld) {
  RN_CHECK(data && x, RESNMTF_E_INVALID, "resnmtf_data_download: NULL argument");
  RN_CHECK(ld >= data->n, RESNMTF_E_INVALID, "resnmtf_data_download: ld < n");
  resnmtf_ctx* ctx = data->ctx;
  RN_CUDA(cudaSetDevice(ctx->device));
  DevBuf tmp;
  RN_CUDA(tmp.alloc(ctx, (size_t)data->n * data->p * sizeof(double), false));
  rn_panels_to_colmajor<<<rn_blocks(data->n * data->p), 256, 0, ctx->stream>>>(data->X, data->n, data->p, data->pp,
                                                                              tmp.d(), data->n);
  RN_CUDA(cudaGetLastError());
  RN_CUDA(cudaMemcpy2DAsync(x, (size_t)ld * sizeof(double), tmp.d(), (size_t)data->n * sizeof(double),
                            (size_t)data->n * sizeof(double), (size_t)data->p, cudaMemcpyDeviceToHost, ctx->stream));
  RN_CUDA(cudaStreamSynchronize(ctx->stream));
  return RESNMTF_OK;
}

extern "C" int resnmtf_data_sums(resnmtf_data* data, double* col_sums, double* row_sums) {
  RN_CHECK(data != nullptr, RESNMTF_E_INVALID, "resnmtf_data_sums: data is NULL");
  resnmtf_ctx* ctx = data->ctx;
  RN_CUDA(cudaSetDevice(ctx->device));
  DevBuf cs, rs;
  if (col_sums) {
    RN_CUDA(cs.alloc(ctx, (size_t)data->pp * sizeof(double), false));
    int rc = col_stat(ctx, data->X, data->ldx, data->pp, RN_STAT_SUM, cs.d());
    if (rc) return rc;
    RN_CUDA(cudaMemcpyAsync(col_sums, cs.d(), (size_t)data->p * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (row_sums) {
    RN_CUDA(rs.alloc(ctx, (size_t)data->ldx * sizeof(double), false));
    rn_row_sums<<<(unsigned)(data->ldx / RN_ROW_TILE), 256, 0, ctx->stream>>>(data->X, data->p, data->pp, rs.d());
    RN_CUDA(cudaGetLastError());
    RN_CUDA(cudaMemcpyAsync(row_sums, rs.d(), (size_t)data->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  RN_CUDA(cudaStreamSynchronize(ctx->stream));
  return RESNMTF_OK;
}

extern "C" int resnmtf_data_subsample(resnmtf_data* src, const int32_t* rows, int64_t n_rows, const int32_t* cols,
                                      int64_t n_cols, resnmtf_data** out) {
  RN_CHECK(src && rows && cols && out, RESNMTF_E_INVALID, "resnmtf_data_subsample: NULL argument");
  RN_CHECK(n_rows >= 1 && n_cols >= 1, RESNMTF_E_INVALID, "resnmtf_data_subsample: empty sub-sample");
  for (int64_t i = 0; i < n_rows; ++i)
    RN_CHECK(rows[i] >= 0 && rows[i] < src->n, RESNMTF_E_INVALID, "resnmtf_data_subsample: row index out of range");
  for (int64_t i = 0; i < n_cols; ++i)
    RN_CHECK(cols[i] >= 0 && cols[i] < src->p, RESNMTF_E_INVALID, "resnmtf_data_subsample: column index out of range");
  resnmtf_ctx* ctx = src->ctx;
  RN_CUDA(cudaSetDevice(ctx->device));
  resnmtf_data* d = nullptr;
  int rc = rn_data_alloc(ctx, n_rows, n_cols, &d);
  if (rc) return rc;
  DevBuf idx;
  cudaError_t e = idx.alloc(ctx, (size_t)(n_rows + n_cols) * sizeof(int32_t), false);
  int32_t* di = static_cast<int32_t*>(idx.p);
  if (e == cudaSuccess) e = cudaMemcpyAsync(di, rows, (size_t)n_rows * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(di + n_rows, cols, (size_t)n_cols * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    rn_gather<<<rn_blocks(n_rows * n_cols), 256, 0, ctx->stream>>>(src->X, src->pp, d->X, n_rows, n_cols, d->pp, di,
                                                                   di + n_rows);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // rows / cols are only borrowed for the call
  if (e != cudaSuccess) rc = rn_fail(RESNMTF_E_CUDA, std::string("resnmtf_data_subsample: ") + cudaGetErrorString(e));
  if (!rc) rc = rn_data_seal(d);
  if (rc) {
    resnmtf_data_destroy(d);
    return rc;
  }
  *out = d;
  return RESNMTF_OK;
}

extern "C" int resnmtf_data_shuffle(resnmtf_data* src, uint64_t seed, int renormalise, int64_t* attempts,
                                    resnmtf_data** out) {
  RN_CHECK(src && out, RESNMTF_E_INVALID, "resnmtf_data_shuffle: NULL argument");
  resnmtf_ctx* ctx = src->ctx;
  RN_CUDA(cudaSetDevice(ctx->device));
  resnmtf_data* d = nullptr;
  int rc = rn_data_alloc(ctx, src->n, src->p, &d);
  if (rc) return rc;
  const uint64_t N = (uint64_t)src->n * (uint64_t)src->p;
  int h = 1;
  while (((uint64_t)1 << (2 * h)) < N) ++h;  // balanced Feistel on 2h bits, cycle walking back into [0, N)
  int64_t tries = 0;
  for (;;) {  // shuffle_view (R/obtain_bicl.r:13-18): reshuffle while a row or a column sums to zero
    ++tries;
    uint64_t key = seed + 0x632be59bd9b4e019ULL * (uint64_t)tries;
    rn_shuffle<<<rn_blocks((int64_t)N), 256, 0, ctx->stream>>>(src->X, d->X, src->n, src->p, src->pp, h, key);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      rc = rn_fail(RESNMTF_E_CUDA, std::string("resnmtf_data_shuffle: ") + cudaGetErrorString(e));
      break;
    }
    bool zero = false;
    if ((rc = has_zero_line(d, &zero))) break;
    if (!zero) break;
    if (tries >= 1000) {
      rc = rn_fail(RESNMTF_E_STATE, "resnmtf_data_shuffle: every shuffle has an all-zero row or column");
      break;
    }
  }
  if (!rc && renormalise) rc = prep_in_place(d, nullptr);  // apply_resnmtf re-preps the shuffled views (R/obtain_bicl.r:35)
  if (!rc) rc = rn_data_seal(d);
  if (rc) {
    resnmtf_data_destroy(d);
    return rc;
  }
  if (attempts) *attempts = tries;
  *out = d;
  return RESNMTF_OK;
}

extern "C" int resnmtf_data_copy(resnmtf_data* src, resnmtf_ctx* dst_ctx, resnmtf_data** out) {
  RN_CHECK(src && dst_ctx && out, RESNMTF_E_INVALID, "resnmtf_data_copy: NULL argument");
  resnmtf_data* d = nullptr;
  int rc = rn_data_alloc(dst_ctx, src->n, src->p, &d);
  if (rc) return rc;
  // The source is a sealed handle: whatever wrote its X was synchronised when the handle was handed out (rn_data_seal),
  // and X is never written afterwards.  The source context's stream must NOT be touched from here: in a pool another
  // worker thread owns it and may have it in CUDA-graph capture (found on 8 GPUs: "operation not permitted when stream
  // is capturing").
  cudaError_t e = cudaSetDevice(dst_ctx->device);
  if (e == cudaSuccess)
    e = cudaMemcpyPeerAsync(d->X, dst_ctx->device, src->X, src->ctx->device, (size_t)src->ldx * src->pp * sizeof(double),
                            dst_ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(dst_ctx->stream);
  if (e != cudaSuccess) {
    resnmtf_data_destroy(d);
    return rn_fail(RESNMTF_E_CUDA, std::string("resnmtf_data_copy: ") + cudaGetErrorString(e));
  }
  d->xnorm2 = src->xnorm2;
  std::lock_guard<std::mutex> lk(src->svd_mu);
  d->svd_u = src->svd_u;
  d->svd_d = src->svd_d;
  d->svd_v = src->svd_v;
  d->svd_kc = src->svd_kc;
  *out = d;
  return RESNMTF_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// SVD initialisation: top singular triplets through the Gram matrix of the smaller side
// ------------------------------------------------------------------------------------------------------------------

namespace {

// a dense matrix in the panel layout on the device (rows padded to 64, columns to 32; padding zero)
struct Panel {
  DevBuf buf;
  int64_t rows = 0, cols = 0, ldx = 0, pp = 0;
  double* p() const { return buf.d(); }
  size_t count() const { return (size_t)ldx * pp; }
  cudaError_t alloc(resnmtf_ctx* ctx, int64_t r, int64_t c) {
    rows = r;
    cols = c;
    ldx = rn_round_up(r, RN_ROW_TILE);
    pp = rn_round_up(c, 32);
    return buf.alloc(ctx, count() * sizeof(double), true);
  }
};

struct AtbOut {  // where the product goes: a panel matrix or a compact column-major device buffer
  double* C = nullptr;
  bool panel = true;
  int64_t ppc = 0, ldc = 0;
};

// C = alpha A[:, :pa]' B[:, :pb] + beta E1 + gamma E2
int launch_atb(resnmtf_ctx* ctx, const double* A, int64_t ppa, int64_t pa, const double* B, int64_t ppb, int64_t pb,
               int row_tiles, bool symmetric, double alpha, double beta, const double* E1, double gamma, const double* E2,
               const AtbOut& out) {
  static bool attr_set[64] = {false};
  if (ctx->device < 64 && !attr_set[ctx->device]) {
    RN_CUDA(cudaFuncSetAttribute(rn_atb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rn_atb_smem()));
    attr_set[ctx->device] = true;
  }
  RnAtb a;
  std::memset(&a, 0, sizeof(a));
  a.A = A;
  a.B = B;
  a.ppa = ppa;
  a.ppb = ppb;
  a.pa = pa;
  a.pb = pb;
  a.row_tiles = row_tiles;
  a.symmetric = symmetric ? 1 : 0;
  a.tiles_i = (int)((pa + 63) / 64);
  a.tiles_j = (int)((pb + 63) / 64);
  const int tiles = symmetric ? a.tiles_i * (a.tiles_i + 1) / 2 : a.tiles_i * a.tiles_j;
  int splits = 1;
  if (tiles < ctx->sm_count) {  // few output tiles: split the contraction so that the whole chip works on it
    splits = std::max(1, std::min(ctx->sm_count / tiles, row_tiles / 2));  // one wave of CTAs
    splits = std::min(splits, 64);
  }
  a.splits = splits;
  DevBuf part, ticket;
  if (splits > 1) {
    RN_CUDA(part.alloc(ctx, (size_t)tiles * splits * 4096 * sizeof(double), false));
    RN_CUDA(ticket.alloc(ctx, (size_t)tiles * sizeof(int), true));
    a.part = part.d();
    a.ticket = static_cast<int*>(ticket.p);
  }
  a.alpha = alpha;
  a.beta = beta;
  a.gamma = gamma;
  a.E1 = E1;
  a.E2 = E2;
  a.C = out.C;
  a.c_panel = out.panel ? 1 : 0;
  a.ppc = out.ppc;
  a.ldc = out.ldc;
  dim3 grid((unsigned)tiles, (unsigned)splits);
  rn_atb<<<grid, RN_ATB_THREADS, rn_atb_smem(), ctx->stream>>>(a);
  RN_CUDA(cudaGetLastError());
  return RESNMTF_OK;  // part / ticket are released in stream order behind the kernel
}

int panel_mul(resnmtf_ctx* ctx, Panel& out, int co0, const Panel* Z, const Panel& Y, const double* M_dev, int ldm, int bi,
              int bo, double alpha, double beta) {
  rn_panel_mul<<<(unsigned)(Y.ldx / RN_ROW_TILE), 256, 0, ctx->stream>>>(out.p(), out.pp, co0, Z ? Z->p() : nullptr,
                                                                        Z ? Z->pp : 0, Y.p(), Y.pp, M_dev, ldm, bi, bo,
                                                                        alpha, beta);
  RN_CUDA(cudaGetLastError());
  return RESNMTF_OK;
}

// Y[:, :b] <- an orthonormal basis of its column space: shifted Cholesky QR, three rounds (the first with the shift
// that keeps the factorisation alive for condition numbers up to 1/u; Fukaya et al.'s shifted CholeskyQR3)
int orthonormalise(resnmtf_ctx* ctx, Panel& Y, int b) {
  DevBuf gdev, mdev;
  RN_CUDA(gdev.alloc(ctx, (size_t)b * b * sizeof(double), false));
  RN_CUDA(mdev.alloc(ctx, (size_t)b * b * sizeof(double), false));
  std::vector<double> G((size_t)b * b), R((size_t)b * b), Ri((size_t)b * b);
  Panel tmp;
  RN_CUDA(tmp.alloc(ctx, Y.rows, b));
  const double u = 1.1102230246251565e-16;
  for (int round = 0; round < 3; ++round) {
    AtbOut o;
    o.C = gdev.d();
    o.panel = false;
    o.ldc = b;
    int rc = launch_atb(ctx, Y.p(), Y.pp, b, Y.p(), Y.pp, b, (int)(Y.ldx / RN_ROW_TILE), true, 1.0, 0.0, nullptr, 0.0,
                        nullptr, o);
    if (rc) return rc;
    RN_CUDA(cudaMemcpyAsync(G.data(), gdev.d(), G.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    RN_CUDA(cudaStreamSynchronize(ctx->stream));
    double tr = 0.0;
    for (int i = 0; i < b; ++i) tr += G[(size_t)i * b + i];
    double shift = (round == 0) ? 11.0 * ((double)Y.rows * b + (double)b * (b + 1)) * u * tr : 0.0;
    bool ok = false;
    for (int attempt = 0; attempt < 8 && !ok; ++attempt) {
      std::vector<double> Gs = G;
      for (int i = 0; i < b; ++i) Gs[(size_t)i * b + i] += shift;
      ok = rn_cholesky_upper(b, Gs.data(), R.data());
      if (!ok) shift = std::max(shift * 100.0, 1.0e-12 * tr);
    }
    RN_CHECK(ok, RESNMTF_E_STATE, "SVD initialisation: the block orthonormalisation broke down");
    rn_upper_inverse(b, R.data(), Ri.data());
    RN_CUDA(cudaMemcpyAsync(mdev.d(), Ri.data(), Ri.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = panel_mul(ctx, tmp, 0, nullptr, Y, mdev.d(), b, b, b, 1.0, 0.0))) return rc;
    rn_panel_copy_cols<<<rn_blocks(Y.rows * b), 256, 0, ctx->stream>>>(Y.p(), Y.pp, 0, tmp.p(), tmp.pp, 0, b, Y.rows);
    RN_CUDA(cudaGetLastError());
    RN_CUDA(cudaStreamSynchronize(ctx->stream));  // Ri is reused by the next round
  }
  return RESNMTF_OK;
}

// The kc largest eigenpairs of the symmetric positive semi-definite W (m x m, panel layout) by Chebyshev-filtered
// subspace iteration with locking: a block of kc + guard vectors; per outer iteration a Rayleigh-Ritz step, locking of
// the leading Ritz pairs whose residual is at rounding level of the matrix norm (what a backward-stable dense solver
// delivers), explicit deflation of the locked pairs, then a Chebyshev filter that damps [0, smallest Ritz value of the
// block] -- degree bounded so that the block's most amplified direction gains at most e^23 on the least.  Only the top
// k <= 16 pairs of the Gram matrix are ever used (R/update_steps.r:93-95 keeps the first k singular triplets).  The
// start block is a fixed function of the seed: same matrix, same answer.  lam: kc values descending; V: m x kc panel.
int topk_eig(resnmtf_ctx* ctx, Panel& W, int kc, std::vector<double>& lam, Panel& V) {
  const int64_t m = W.rows;
  const int row_tiles = (int)(W.ldx / RN_ROW_TILE);
  const int guard = 48, max_outer = 60;
  const double tol = 2.0e-15;
  double best = INFINITY;  // smallest residual of the leading unlocked pair so far, and how long it has not moved
  int stall = 0;
  int b = (int)std::min<int64_t>(m, kc + guard);
  lam.clear();
  RN_CUDA(V.alloc(ctx, m, kc));
  Panel Q, AQ, T0, T1, T2, Wd;
  RN_CUDA(Q.alloc(ctx, m, b));
  RN_CUDA(AQ.alloc(ctx, m, b));
  RN_CUDA(T0.alloc(ctx, m, b));
  RN_CUDA(T1.alloc(ctx, m, b));
  RN_CUDA(T2.alloc(ctx, m, b));
  DevBuf hdev, ydev, thdev, ndev, mlock;
  RN_CUDA(hdev.alloc(ctx, (size_t)64 * 64 * sizeof(double), false));
  RN_CUDA(ydev.alloc(ctx, (size_t)64 * 64 * sizeof(double), false));
  RN_CUDA(thdev.alloc(ctx, (size_t)64 * sizeof(double), false));
  RN_CUDA(ndev.alloc(ctx, (size_t)Q.pp * sizeof(double), false));
  RN_CUDA(mlock.alloc(ctx, (size_t)16 * 64 * sizeof(double), false));
  Panel* work = &W;
  int nlock = 0;
  int rc;
  rn_panel_random<<<rn_blocks(m * b), 256, 0, ctx->stream>>>(Q.p(), m, Q.pp, b, 20260000ULL);
  RN_CUDA(cudaGetLastError());
  if ((rc = orthonormalise(ctx, Q, b))) return rc;

  // y <- y - V_lock (V_lock' y): keeps the block orthogonal to the locked vectors
  auto off_locked = [&](Panel& Y, int bw) -> int {
    if (!nlock) return RESNMTF_OK;
    AtbOut o;
    o.C = mlock.d();
    o.panel = false;
    o.ldc = nlock;
    int r2 = launch_atb(ctx, V.p(), V.pp, nlock, Y.p(), Y.pp, bw, row_tiles, false, 1.0, 0.0, nullptr, 0.0, nullptr, o);
    if (r2) return r2;
    return panel_mul(ctx, Y, 0, &Y, V, mlock.d(), nlock, nlock, bw, -1.0, 1.0);
  };

  std::vector<double> H((size_t)64 * 64), Yh((size_t)64 * 64), th(64), Ys((size_t)64 * 64), rs(64);
  for (int outer = 0; outer < max_outer; ++outer) {
    // AQ = work Q (work is symmetric: work Q = work' Q)
    AtbOut oq;
    oq.C = AQ.p();
    oq.ppc = AQ.pp;
    if ((rc = launch_atb(ctx, work->p(), work->pp, m, Q.p(), Q.pp, b, row_tiles, false, 1.0, 0.0, nullptr, 0.0, nullptr, oq)))
      return rc;
    if ((rc = off_locked(AQ, b))) return rc;
    // Rayleigh-Ritz: H = Q' AQ (b x b) on the host
    AtbOut oh;
    oh.C = hdev.d();
    oh.panel = false;
    oh.ldc = b;
    if ((rc = launch_atb(ctx, Q.p(), Q.pp, b, AQ.p(), AQ.pp, b, row_tiles, false, 1.0, 0.0, nullptr, 0.0, nullptr, oh)))
      return rc;
    RN_CUDA(cudaMemcpyAsync(H.data(), hdev.d(), (size_t)b * b * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    RN_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < b; ++i)
      for (int j = i + 1; j < b; ++j) {
        const double s = 0.5 * (H[(size_t)i + (size_t)j * b] + H[(size_t)j + (size_t)i * b]);
        H[(size_t)i + (size_t)j * b] = H[(size_t)j + (size_t)i * b] = s;
      }
    RN_CHECK(rn_sym_eig(b, H.data(), th.data(), Yh.data()), RESNMTF_E_STATE,
             "SVD initialisation: the Rayleigh-Ritz eigenproblem did not converge");
    // descending order
    for (int j = 0; j < b; ++j)
      for (int i = 0; i < b; ++i) Ys[(size_t)i + (size_t)j * b] = Yh[(size_t)i + (size_t)(b - 1 - j) * b];
    std::reverse(th.begin(), th.begin() + b);
    RN_CUDA(cudaMemcpyAsync(ydev.d(), Ys.data(), (size_t)b * b * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    RN_CUDA(cudaMemcpyAsync(thdev.d(), th.data(), (size_t)b * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    // Q <- Q Y, AQ <- AQ Y
    if ((rc = panel_mul(ctx, T0, 0, nullptr, Q, ydev.d(), b, b, b, 1.0, 0.0))) return rc;
    if ((rc = panel_mul(ctx, T1, 0, nullptr, AQ, ydev.d(), b, b, b, 1.0, 0.0))) return rc;
    std::swap(Q.buf.p, T0.buf.p);
    std::swap(AQ.buf.p, T1.buf.p);
    // residual norms of the Ritz pairs
    rn_panel_resid<<<rn_blocks((int64_t)row_tiles * b * 64), 256, 0, ctx->stream>>>(T2.p(), AQ.p(), Q.p(), thdev.d(), Q.pp, b,
                                                                                   row_tiles);
    RN_CUDA(cudaGetLastError());
    if ((rc = col_stat(ctx, T2.p(), T2.ldx, T2.pp, RN_STAT_SUMSQ, ndev.d()))) return rc;
    RN_CUDA(cudaMemcpyAsync(rs.data(), ndev.d(), (size_t)b * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    RN_CUDA(cudaStreamSynchronize(ctx->stream));
    const double scale = std::max(std::fabs(th[0]), lam.empty() ? 0.0 : lam[0]);
    int n_new = 0;  // leading pairs only, in order: nothing above a locked pair is still moving
    double accept = tol * scale;
    {  // a residual that sits at its rounding floor (a few ulp above tol on a wide matrix) is as converged as it gets
      const double r0 = std::sqrt(rs[0]);
      stall = (r0 > 0.5 * best) ? stall + 1 : 0;
      best = std::min(best, r0);
      if (stall >= 3 && r0 <= 1.0e-13 * scale) accept = std::max(accept, 2.0 * r0);
    }
    while (n_new < b && nlock + n_new < kc && std::sqrt(rs[n_new]) <= accept) ++n_new;
    if (n_new) {
      best = INFINITY;
      stall = 0;
    }
    if (n_new) {
      rn_panel_copy_cols<<<rn_blocks(m * n_new), 256, 0, ctx->stream>>>(V.p(), V.pp, nlock, Q.p(), Q.pp, 0, n_new, m);
      RN_CUDA(cudaGetLastError());
      if (work == &W) {  // deflated copy once something is locked: locked pairs then sit at eigenvalue 0
        RN_CUDA(Wd.alloc(ctx, m, m));
        RN_CUDA(cudaMemcpyAsync(Wd.p(), W.p(), W.count() * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        work = &Wd;
      }
      // the n_new new vectors are columns nlock .. of V; deflate with their Ritz values
      DevBuf lnew;
      RN_CUDA(lnew.alloc(ctx, (size_t)16 * sizeof(double), false));
      RN_CUDA(cudaMemcpyAsync(lnew.d(), th.data(), (size_t)n_new * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      {
        // V columns nlock.. as a (column-offset) panel view: rn_deflate reads column c of its V argument at offset c,
        // so hand it a temporary compact copy
        Panel vn;
        RN_CUDA(vn.alloc(ctx, m, n_new));
        rn_panel_copy_cols<<<rn_blocks(m * n_new), 256, 0, ctx->stream>>>(vn.p(), vn.pp, 0, Q.p(), Q.pp, 0, n_new, m);
        dim3 grid((unsigned)row_tiles, (unsigned)row_tiles);
        rn_deflate<<<grid, 256, 0, ctx->stream>>>(Wd.p(), m, Wd.pp, vn.p(), vn.pp, lnew.d(), n_new);
        RN_CUDA(cudaGetLastError());
        RN_CUDA(cudaStreamSynchronize(ctx->stream));  // th / vn go out of scope
      }
      for (int i = 0; i < n_new; ++i) lam.push_back(th[i]);
      nlock += n_new;
      // Q <- Q[:, n_new:]
      const int nb2 = b - n_new;
      if (nb2 > 0) {
        RN_CUDA(cudaMemsetAsync(T0.p(), 0, T0.count() * sizeof(double), ctx->stream));
        rn_panel_copy_cols<<<rn_blocks(m * nb2), 256, 0, ctx->stream>>>(T0.p(), T0.pp, 0, Q.p(), Q.pp, n_new, nb2, m);
        RN_CUDA(cudaGetLastError());
        std::swap(Q.buf.p, T0.buf.p);
      }
      for (int i = 0; i + n_new < b; ++i) th[i] = th[i + n_new];
      b = nb2;
    }
    if (nlock >= kc) {
      RN_CUDA(cudaStreamSynchronize(ctx->stream));
      return RESNMTF_OK;
    }
    double c = -1.0;  // the unwanted part of the (deflated) spectrum lies in [0, c]
    for (int i = b - 1; i >= 0; --i)
      if (th[i] > 0.0) {
        c = th[i];
        break;
      }
    RN_CHECK(b >= 2 && c > 0.0, RESNMTF_E_STATE, "SVD initialisation: the subspace iteration ran out of directions");
    const double t_max = std::max((2.0 * th[0] - c) / c, 1.0 + 1.0e-12);
    const int deg = (int)std::min(40.0, std::max(2.0, std::floor(23.0 / std::acosh(t_max))));
    const double e = 0.5 * c;
    // T_j((W - e I) / e) Q by the three-term recurrence; the shift is folded into the product's epilogue
    Panel* y0 = &Q;
    Panel* y1 = &T0;
    Panel* y2 = &T1;
    AtbOut o1;
    o1.C = y1->p();
    o1.ppc = y1->pp;
    if ((rc = launch_atb(ctx, work->p(), work->pp, m, y0->p(), y0->pp, b, row_tiles, false, 1.0 / e, -1.0, y0->p(), 0.0, nullptr, o1)))
      return rc;
    for (int dgr = 2; dgr <= deg; ++dgr) {
      AtbOut o2;
      o2.C = y2->p();
      o2.ppc = y2->pp;
      if ((rc = launch_atb(ctx, work->p(), work->pp, m, y1->p(), y1->pp, b, row_tiles, false, 2.0 / e, -2.0, y1->p(), -1.0,
                           y0->p(), o2)))
        return rc;
      Panel* t = y0;
      y0 = y1;
      y1 = y2;
      y2 = t;
    }
    if ((rc = off_locked(*y1, b))) return rc;
    if (y1 != &Q) {  // the filtered block becomes the new Q (buffers are the same size)
      std::swap(Q.buf.p, y1->buf.p);
    }
    if ((rc = orthonormalise(ctx, Q, b))) return rc;
  }
  return rn_fail(RESNMTF_E_STATE, "SVD initialisation: the subspace iteration did not converge");
}

// dense route for tiny Gram matrices (order <= 64): host eigensolver
int dense_eig(resnmtf_ctx* ctx, Panel& W, int kc, std::vector<double>& lam, Panel& V) {
  const int m = (int)W.rows;
  DevBuf tmp;
  RN_CUDA(tmp.alloc(ctx, (size_t)m * m * sizeof(double), false));
  rn_panels_to_colmajor<<<rn_blocks((int64_t)m * m), 256, 0, ctx->stream>>>(W.p(), m, m, W.pp, tmp.d(), m);
  RN_CUDA(cudaGetLastError());
  std::vector<double> A((size_t)m * m), w((size_t)m), Z((size_t)m * m);
  RN_CUDA(cudaMemcpyAsync(A.data(), tmp.d(), A.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  RN_CUDA(cudaStreamSynchronize(ctx->stream));
  RN_CHECK(rn_sym_eig(m, A.data(), w.data(), Z.data()), RESNMTF_E_STATE, "SVD initialisation: dense eigensolver failed");
  lam.assign((size_t)kc, 0.0);
  std::vector<double> vt((size_t)m * kc);
  for (int c = 0; c < kc; ++c) {
    lam[c] = w[m - 1 - c];
    for (int i = 0; i < m; ++i) vt[(size_t)i + (size_t)c * m] = Z[(size_t)i + (size_t)(m - 1 - c) * m];
  }
  RN_CUDA(V.alloc(ctx, m, kc));
  // at most 64 x 16 entries: laid out in the panel order on the host (host twin of rn_xidx)
  std::vector<double> vp((size_t)V.ldx * V.pp, 0.0);
  for (int c = 0; c < kc; ++c)
    for (int i = 0; i < m; ++i) {
      const int sigma = ((c & 1) << 2) | (c & 2);
      vp[((size_t)(i >> 6) * V.pp + c) * 64 + 2 * (((i & 63) >> 1) ^ sigma) + (i & 1)] = vt[(size_t)i + (size_t)c * m];
    }
  RN_CUDA(cudaMemcpyAsync(V.p(), vp.data(), vp.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  RN_CUDA(cudaStreamSynchronize(ctx->stream));
  return RESNMTF_OK;
}

// |U| (n x kc), d (kc), |V| (p x kc) of the view into the handle's cache
int compute_svd(resnmtf_data* data) {
  resnmtf_ctx* ctx = data->ctx;
  const int64_t n = data->n, p = data->p;
  const int kc = (int)std::min<int64_t>(RESNMTF_MAX_K, std::min(n, p));
  const bool cols_side = p <= n;  // Gram matrix of the smaller side
  // the matrix whose columns are the Gram side: X itself, or its transpose
  Panel Xt;
  const double* G_src = data->X;
  int64_t g_ldx = data->ldx, g_pp = data->pp, g_rows = n, g_cols = p;
  if (!cols_side) {
    RN_CUDA(Xt.alloc(ctx, p, n));
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((p + 31) / 32));
    rn_transpose_panels<<<grid, 256, 0, ctx->stream>>>(data->X, n, p, data->pp, Xt.p(), Xt.pp);
    RN_CUDA(cudaGetLastError());
    G_src = Xt.p();
    g_ldx = Xt.ldx;
    g_pp = Xt.pp;
    g_rows = p;
    g_cols = n;
  }
  const int64_t m = g_cols;
  Panel W, V;
  RN_CUDA(W.alloc(ctx, m, m));
  AtbOut ow;
  ow.C = W.p();
  ow.ppc = W.pp;
  int rc = launch_atb(ctx, G_src, g_pp, m, G_src, g_pp, m, (int)(g_ldx / RN_ROW_TILE), true, 1.0, 0.0, nullptr, 0.0, nullptr, ow);
  if (rc) return rc;
  std::vector<double> lam;
  rc = (m <= 64) ? dense_eig(ctx, W, kc, lam, V) : topk_eig(ctx, W, kc, lam, V);
  if (rc) return rc;
  // singular values and the side that came out of the eigensolver (column-major on the host, absolute values)
  std::vector<double> dvals((size_t)kc), side((size_t)m * kc);
  for (int c = 0; c < kc; ++c) dvals[c] = std::sqrt(std::max(lam[c], 0.0));
  DevBuf tmp, v16, ddev, other;
  RN_CUDA(tmp.alloc(ctx, (size_t)m * kc * sizeof(double), false));
  rn_panels_to_colmajor<<<rn_blocks(m * kc), 256, 0, ctx->stream>>>(V.p(), m, kc, V.pp, tmp.d(), m);
  RN_CUDA(cudaGetLastError());
  RN_CUDA(cudaMemcpyAsync(side.data(), tmp.d(), side.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  RN_CUDA(cudaStreamSynchronize(ctx->stream));
  // the other side: (G_src) V / d, one pass over the view
  std::vector<double> vrow((size_t)g_pp * 16, 0.0);
  for (int c = 0; c < kc; ++c)
    for (int64_t j = 0; j < m; ++j) vrow[(size_t)j * 16 + c] = side[(size_t)j + (size_t)c * m];
  RN_CUDA(v16.alloc(ctx, vrow.size() * sizeof(double), false));
  RN_CUDA(ddev.alloc(ctx, 16 * sizeof(double), true));
  RN_CUDA(other.alloc(ctx, (size_t)g_rows * 16 * sizeof(double), true));
  RN_CUDA(cudaMemcpyAsync(v16.d(), vrow.data(), vrow.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  RN_CUDA(cudaMemcpyAsync(ddev.d(), dvals.data(), (size_t)kc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  rn_xv16<<<(unsigned)(g_ldx / RN_ROW_TILE), 256, 0, ctx->stream>>>(G_src, g_rows, m, g_pp, v16.d(), ddev.d(), kc, other.d(),
                                                                    g_rows);
  RN_CUDA(cudaGetLastError());
  std::vector<double> oth((size_t)g_rows * kc);
  RN_CUDA(cudaMemcpyAsync(oth.data(), other.d(), oth.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  RN_CUDA(cudaStreamSynchronize(ctx->stream));
  for (double& x : side) x = std::fabs(x);
  std::lock_guard<std::mutex> pub(data->svd_mu);
  data->svd_d = dvals;
  if (cols_side) {
    data->svd_v = side;  // p x kc
    data->svd_u = oth;   // n x kc
  } else {
    data->svd_u = side;
    data->svd_v = oth;
  }
  data->svd_kc = kc;
  return RESNMTF_OK;
}

}  // namespace

// |U[, 1:k]|, d[1:k], |V[, 1:k]| of the view: what init_mats_inner() (R/update_steps.r:92-95) takes from svd(x)
int rn_data_svd(resnmtf_data* data) {
  auto cached = [&]() {
    std::lock_guard<std::mutex> lk(data->svd_mu);
    return data->svd_kc > 0;
  };
  if (cached()) return RESNMTF_OK;
  std::lock_guard<std::mutex> one(data->svd_compute_mu);  // svd_mu itself stays free: resnmtf_data_copy must not wait ~50 ms
  if (cached()) return RESNMTF_OK;
  RN_CUDA(cudaSetDevice(data->ctx->device));
  return compute_svd(data);
}

void rn_data_svd_adopt(resnmtf_data* dst, resnmtf_data* src) {
  if (!dst || !src || dst == src) return;
  std::vector<double> u, d, v;
  int kc;
  {
    std::lock_guard<std::mutex> lk(src->svd_mu);
    u = src->svd_u;
    d = src->svd_d;
    v = src->svd_v;
    kc = src->svd_kc;
  }
  if (kc <= 0) return;
  std::lock_guard<std::mutex> lk(dst->svd_mu);
  if (dst->svd_kc > 0) return;
  dst->svd_u.swap(u);
  dst->svd_d.swap(d);
  dst->svd_v.swap(v);
  dst->svd_kc = kc;
}

extern "C" int resnmtf_data_svd_topk(resnmtf_data* data, int k, double* u, double* d, double* v) {
  RN_CHECK(data != nullptr, RESNMTF_E_INVALID, "resnmtf_data_svd_topk: data is NULL");
  RN_CHECK(k >= 1 && k <= RESNMTF_MAX_K && k <= data->n && k <= data->p, RESNMTF_E_INVALID,
           "resnmtf_data_svd_topk: k out of range");
  int rc = rn_data_svd(data);
  if (rc) return rc;
  if (u) std::memcpy(u, data->svd_u.data(), (size_t)data->n * k * sizeof(double));
  if (d) std::memcpy(d, data->svd_d.data(), (size_t)k * sizeof(double));
  if (v) std::memcpy(v, data->svd_v.data(), (size_t)data->p * k * sizeof(double));
  return RESNMTF_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// bisilhouette
// ------------------------------------------------------------------------------------------------------------------

// Bisilhouette score of one view's biclustering (as obtain_biclusters() asks for it, R/obtain_bicl.r:190-199): for every
// non-empty bicluster (R_j, C_j) the silhouette of the rows of R_j computed on the columns C_j only, against the other
// row clusters R_l \ R_j (or, when there is none, against the rows outside R_j); vals[j] = mean row silhouette (0 for
// empty biclusters), *bisil = mean over the non-empty ones.  row_cl n x k, col_cl p x k: column-major, non-zero = member.
static int bisil_impl(resnmtf_data* data, const double* row_cl, const double* col_cl, int k, int method,
                      const int32_t* want, double* vals, double* bisil, int32_t* n_live);

extern "C" int resnmtf_data_bisil(resnmtf_data* data, const double* row_cl, const double* col_cl, int k, int method,
                                  double* vals, double* bisil) {
  RN_CHECK(bisil != nullptr, RESNMTF_E_INVALID, "resnmtf_data_bisil: NULL argument");
  return bisil_impl(data, row_cl, col_cl, k, method, nullptr, vals, bisil, nullptr);
}

// The same for a SUBSET of the biclusters (want[j] != 0): the per-bicluster values are independent of each other, so the
// biclusters of one fit can be scored on different GPUs (each holding a copy of the view) and combined by the caller --
// vals[j] of the wanted ones, 0 elsewhere; *n_live = number of non-empty biclusters of the whole clustering, the
// denominator of the mean.
extern "C" int resnmtf_data_bisil_part(resnmtf_data* data, const double* row_cl, const double* col_cl, int k, int method,
                                       const int32_t* want, double* vals, int32_t* n_live) {
  RN_CHECK(want && vals, RESNMTF_E_INVALID, "resnmtf_data_bisil_part: NULL argument");
  double ignored = 0.0;
  return bisil_impl(data, row_cl, col_cl, k, method, want, vals, &ignored, n_live);
}

static int bisil_impl(resnmtf_data* data, const double* row_cl, const double* col_cl, int k, int method,
                      const int32_t* want, double* vals, double* bisil, int32_t* n_live) {
  RN_CHECK(data && row_cl && col_cl && bisil, RESNMTF_E_INVALID, "resnmtf_data_bisil: NULL argument");
  RN_CHECK(k >= 1 && k <= 64, RESNMTF_E_INVALID, "resnmtf_data_bisil: k out of range");
  RN_CHECK(method >= 0 && method <= 2, RESNMTF_E_INVALID,
           "distance must be one of 'euclidean', 'manhattan' or 'cosine'.");
  resnmtf_ctx* ctx = data->ctx;
  RN_CUDA(cudaSetDevice(ctx->device));
  const int64_t n = data->n, p = data->p;
  std::vector<std::vector<int32_t>> R((size_t)k), Cc((size_t)k);
  for (int j = 0; j < k; ++j) {
    for (int64_t i = 0; i < n; ++i)
      if (row_cl[(size_t)i + (size_t)j * n] > 0.0) R[j].push_back((int32_t)i);
    for (int64_t i = 0; i < p; ++i)
      if (col_cl[(size_t)i + (size_t)j * p] > 0.0) Cc[j].push_back((int32_t)i);
  }
  std::vector<int> live;
  for (int j = 0; j < k; ++j)
    if (!R[j].empty() && !Cc[j].empty()) live.push_back(j);
  if (vals)
    for (int j = 0; j < k; ++j) vals[j] = 0.0;
  double total = 0.0;
  if (n_live) *n_live = (int32_t)live.size();
  for (int j : live) {
    if (want && !want[j]) continue;
    // segments: R_j, then R_l \ R_j for every other live l (non-empty ones), else the complement of R_j
    std::vector<std::vector<int32_t>> seg;
    seg.push_back(R[j]);
    std::vector<char> in_j((size_t)n, 0);
    for (int32_t r : R[j]) in_j[r] = 1;
    for (int l : live) {
      if (l == j) continue;
      std::vector<int32_t> g;
      for (int32_t r : R[l])
        if (!in_j[r]) g.push_back(r);
      if (!g.empty()) seg.push_back(g);
    }
    if (seg.size() == 1) {
      std::vector<int32_t> g;
      for (int64_t r = 0; r < n; ++r)
        if (!in_j[r]) g.push_back((int32_t)r);
      if (!g.empty()) seg.push_back(g);
    }
    if (seg.size() == 1) continue;  // no other rows at all: the bicluster scores 0
    RN_CHECK(seg.size() <= 17, RESNMTF_E_UNSUPPORTED, "resnmtf_data_bisil: more than 16 other row clusters");
    const int n_seg = (int)seg.size();
    std::vector<int32_t> slots, seg_of_tile;
    for (int sg = 0; sg < n_seg; ++sg) {
      for (int32_t r : seg[sg]) slots.push_back(r);
      while (slots.size() % 64) slots.push_back(-1);
      while (seg_of_tile.size() < slots.size() / 64) seg_of_tile.push_back(sg);
    }
    const int a_tiles = (int)((seg[0].size() + 63) / 64), b_tiles = (int)(slots.size() / 64);
    const int m = (int)Cc[j].size(), ldy = (m + 15) / 16 * 16;
    int splits = std::max(1, std::min(b_tiles, (4 * ctx->sm_count + a_tiles - 1) / a_tiles));
    DevBuf d_slots, d_cols, d_segt, d_Y, d_norms, d_part;
    RN_CUDA(d_slots.alloc(ctx, slots.size() * sizeof(int32_t), false));
    RN_CUDA(d_cols.alloc(ctx, (size_t)m * sizeof(int32_t), false));
    RN_CUDA(d_segt.alloc(ctx, seg_of_tile.size() * sizeof(int32_t), false));
    RN_CUDA(d_Y.alloc(ctx, slots.size() * (size_t)ldy * sizeof(double), false));
    RN_CUDA(d_norms.alloc(ctx, slots.size() * sizeof(double), false));
    const size_t part_count = (size_t)splits * n_seg * a_tiles * 64;
    RN_CUDA(d_part.alloc(ctx, part_count * sizeof(double), false));
    RN_CUDA(cudaMemcpyAsync(d_slots.p, slots.data(), slots.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    RN_CUDA(cudaMemcpyAsync(d_cols.p, Cc[j].data(), (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    RN_CUDA(cudaMemcpyAsync(d_segt.p, seg_of_tile.data(), seg_of_tile.size() * sizeof(int32_t), cudaMemcpyHostToDevice,
                            ctx->stream));
    rn_bisil_gather<<<(unsigned)((slots.size() + 7) / 8), 256, 0, ctx->stream>>>(
        data->X, data->pp, static_cast<const int32_t*>(d_slots.p), (int64_t)slots.size(),
        static_cast<const int32_t*>(d_cols.p), m, ldy, d_Y.d(), d_norms.d());
    RnBisil a;
    a.Y = d_Y.d();
    a.norms = method == RN_DIST_COSINE ? d_norms.d() : nullptr;
    a.rows = static_cast<const int32_t*>(d_slots.p);
    a.ldy = ldy;
    a.a_tiles = a_tiles;
    a.b_tiles = b_tiles;
    a.seg_of_tile = static_cast<const int32_t*>(d_segt.p);
    a.n_seg = n_seg;
    a.splits = splits;
    a.part = d_part.d();
    dim3 grid((unsigned)a_tiles, (unsigned)splits);
    if (method == RN_DIST_EUCLIDEAN) rn_bisil_dist<RN_DIST_EUCLIDEAN><<<grid, 256, 0, ctx->stream>>>(a);
    else if (method == RN_DIST_MANHATTAN) rn_bisil_dist<RN_DIST_MANHATTAN><<<grid, 256, 0, ctx->stream>>>(a);
    else rn_bisil_dist<RN_DIST_COSINE><<<grid, 256, 0, ctx->stream>>>(a);
    RN_CUDA(cudaGetLastError());
    std::vector<double> part(part_count);
    RN_CUDA(cudaMemcpyAsync(part.data(), d_part.d(), part_count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    RN_CUDA(cudaStreamSynchronize(ctx->stream));
    const size_t na = seg[0].size(), stride = (size_t)a_tiles * 64;
    double sil_sum = 0.0;
    for (size_t i = 0; i < na; ++i) {
      double own = 0.0, best = INFINITY;
      for (int sg = 0; sg < n_seg; ++sg) {
        double s = 0.0;
        for (int q = 0; q < splits; ++q) s += part[((size_t)q * n_seg + sg) * stride + i];
        if (sg == 0) own = na > 1 ? s / (double)std::max<size_t>(na - 1, 1) : 0.0;
        else best = std::min(best, s / (double)seg[sg].size());
      }
      const double den = std::max(own, best);
      sil_sum += den > 0.0 ? (best - own) / den : 0.0;
    }
    const double val = sil_sum / (double)na;
    if (vals) vals[j] = val;
    total += val;
  }
  *bisil = live.empty() ? 0.0 : total / (double)live.size();
  return RESNMTF_OK;
}
