// Hand-written sm_100a kernels of the ResNMTF update sweep (FP64 throughout).
//
// One update-iteration of view v (reference: update_matrices, R/update_steps.r:272-319) is, on the
// tensor-core path (k <= 8):
//   rn_f_step_sk   P = X.G streamed over X once (FP64 mma.sync), fused with the F update (update_f,
//                  :141-165) incl. the phi coupling gather (star_prod_relevant, R/utils.r:63-78)
//   rn_g_step_sk   T = X'.F streamed over X once, F'F and colSums(F) from the same fragments, fused with
//                  the G update (update_g, :180-207) incl. psi coupling, G'G, A = T'G, and in the last CTA
//                  the S update (update_s, :220-240; its numerator crossprod(F,X) G equals A, so X is not
//                  read a third time), lambda/mu (update_lm, :249-251), the algebraic error and the
//                  iteration bookkeeping (R/main.r:74-80)
// Both are persistent "stream-K" kernels: the grid is the number of resident CTAs, every CTA streams an
// equal, contiguous share of the 64x32 (F) / 64x64 (G) element units of X, and tiles that straddle two
// CTAs are combined by the last CTA to arrive, in CTA order.
// CUDA-core fallbacks (k 9..16, or RESNMTF_IMPL_DFMA): rn_f_step_dfma, rn_gram_f, rn_g_stream_dfma,
// rn_g_epilogue.  Optional direct error pass: rn_residual (calculate_error, R/utils.r:157-166).
// All cross-CTA reductions have a fixed summation order (no floating-point atomics), so results are
// bit-reproducible run to run.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "rn_types.h"
#include "rn_dev.cuh"

// ------------------------------------------------------------------------------------------------
// shared per-row math of the two multiplicative updates
// ------------------------------------------------------------------------------------------------

// update_f for one row r (R/update_steps.r:141-165).  P = (X G)[r,], Ssm = S, Wsm = crossprod(G) t(S),
// lamh = lambda/2.  Writes |F_new| in place.
template <int K>
__device__ __forceinline__ void rn_update_f_row(const RnView& vw, const RnFit& ft, const int v, const int64_t r,
                                                const double* P, const double* Ssm, const double* Wsm,
                                                const double* lamh, double* fnew = nullptr) {
  const int kp = vw.kp;
  const int V = ft.n_views;
  double f[K], N[K], FS[K], D[K];
#pragma unroll
  for (int c = 0; c < K; ++c) f[c] = vw.F[rn_fidx(r, c, kp)];
#pragma unroll
  for (int c = 0; c < K; ++c) {  // (X G) t(S)
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < K; ++a) s = fma(P[a], Ssm[c + a * K], s);
    N[c] = s;
  }
#pragma unroll
  for (int a = 0; a < K; ++a) {  // F S
    double s = 0.0;
#pragma unroll
    for (int b = 0; b < K; ++b) s = fma(f[b], Ssm[b + a * K], s);
    FS[a] = s;
  }
#pragma unroll
  for (int c = 0; c < K; ++c) {  // (F S) W
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < K; ++a) s = fma(FS[a], Wsm[a + c * K], s);
    D[c] = s;
  }
  double phisum = 0.0;
  for (int w = 0; w < V; ++w) phisum += ft.phi[w + v * V];
  if (phisum == 0.0) {
#pragma unroll
    for (int c = 0; c < K; ++c) {
      double ratio = N[c] / (D[c] + lamh[c]);
      if (isnan(ratio)) ratio = 1.0;
      const double o = fabs(f[c] * ratio);
      vw.F[rn_fidx(r, c, kp)] = o;
      if (fnew) fnew[c] = o;
    }
  } else {
    double pc[K];
#pragma unroll
    for (int c = 0; c < K; ++c) pc[c] = 0.0;
    for (int w = 0; w < V; ++w) {
      const double ph = ft.phi[w + v * V];
      if (ph == 0.0) continue;
      const int mode = ft.rowmode[w + v * V];
      if (mode == RN_MODE_NA) continue;
      const RnView* ow = ft.views + w;
      const double nw = (double)ow->n_glob;
      int64_t src = -1;
      if (mode == RN_MODE_MAP) src = ft.rowmap[w + v * V][r];
      const double* fw = ow->F;
      const int kpw = ow->kp;
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const double m = (src >= 0) ? fw[rn_fidx(src, c, kpw)] : f[c];
        pc[c] += (ph * m) * nw;
      }
    }
    const double nv = (double)vw.n_glob;
#pragma unroll
    for (int c = 0; c < K; ++c) {
      const double num = N[c] + pc[c] / nv;
      const double den = (D[c] + phisum * f[c]) + lamh[c];
      const double o = fabs(f[c] * (num / den));
      vw.F[rn_fidx(r, c, kp)] = o;
      if (fnew) fnew[c] = o;
    }
  }
}

// update_g for one data column j (R/update_steps.r:180-207).  Tj = crossprod(X, F)[j,], Ssm = S,
// Vs = crossprod(F) S, muh = mu/2.  Returns |G_new[j,]| in gn and writes it in place.
template <int K>
__device__ __forceinline__ void rn_update_g_row(const RnView& vw, const RnFit& ft, const int v, const int64_t j,
                                                const double* Tj, const double* Ssm, const double* Vs,
                                                const double* muh, double* gn) {
  const int KP = vw.kp;
  const int V = ft.n_views;
  double gj[K], N[K], GS[K], D[K];
#pragma unroll
  for (int c = 0; c < K; ++c) gj[c] = vw.G[j * KP + c];
#pragma unroll
  for (int c = 0; c < K; ++c) {  // crossprod(X, F) %*% S
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < K; ++a) s = fma(Tj[a], Ssm[a + c * K], s);
    N[c] = s;
  }
#pragma unroll
  for (int a = 0; a < K; ++a) {  // G t(S)
    double s = 0.0;
#pragma unroll
    for (int b = 0; b < K; ++b) s = fma(gj[b], Ssm[a + b * K], s);
    GS[a] = s;
  }
#pragma unroll
  for (int c = 0; c < K; ++c) {
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < K; ++a) s = fma(GS[a], Vs[a + c * K], s);
    D[c] = s;
  }
  if (ft.psi_total == 0.0) {
#pragma unroll
    for (int c = 0; c < K; ++c) {
      double ratio = rn_fast_div(N[c], D[c] + muh[c]);
      if (isnan(ratio)) ratio = 1.0;
      gn[c] = fabs(gj[c] * ratio);
    }
  } else {
    double psisum = 0.0;
    for (int w = 0; w < V; ++w) psisum += ft.psi[w + v * V];
    double pc[K];
#pragma unroll
    for (int c = 0; c < K; ++c) pc[c] = 0.0;
    for (int w = 0; w < V; ++w) {
      const double ps = ft.psi[w + v * V];
      if (ps == 0.0) continue;
      const int mode = ft.colmode[w + v * V];
      if (mode == RN_MODE_NA) continue;
      const RnView* ow = ft.views + w;
      const double pw = (double)ow->p;
      int64_t src = -1;
      if (mode == RN_MODE_MAP) src = ft.colmap[w + v * V][j];
      const double* gw = ow->G;
      const int kpw = ow->kp;
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const double m = (src >= 0) ? gw[src * kpw + c] : gj[c];
        pc[c] += (ps * m) * pw;
      }
    }
    const double pv = (double)vw.p;
#pragma unroll
    for (int c = 0; c < K; ++c) {
      const double num = N[c] + rn_fast_div(pc[c], pv);
      const double den = (D[c] + psisum * gj[c]) + muh[c];
      gn[c] = fabs(gj[c] * rn_fast_div(num, den));
    }
  }
#pragma unroll
  for (int c = 0; c < K; ++c) vw.G[j * KP + c] = gn[c];
}

// Iteration bookkeeping (R/main.r:74-80): mean error over the views, history, stop rule.  One thread.
__device__ __forceinline__ void rn_finish_dev(const RnFit& ft) {
  RnCtrl* c = ft.ctrl;
  double s = 0.0;
  for (int v = 0; v < ft.n_views; ++v) s += ft.views[v].scal[1];
  const double mean = s / (double)ft.n_views;
  if (c->hist_count < ft.hist_cap) ft.hist[c->hist_count] = mean;
  c->hist_count += 1;
  c->iters += 1;
  const double diff = fabs(mean - c->prev_err);
  c->last_diff = diff;
  c->prev_err = mean;
  if (c->conv_mode) {
    if (isnan(mean)) c->done = 2;
    else if (!(diff > c->tol)) c->done = 1;
  }
}

// Finishes view v once G'G | A | colSums(G) (fin[0..2K^2+K)) and F'F | colSums(F) (FtFs) are complete:
// update_s (R/update_steps.r:220-240), update_lm (:249-251), algebraic error, and -- when fuse_finish --
// the iteration bookkeeping.  Called by all NT threads of the last CTA.  Us / Sn / red: K*K scratch.
template <int K, int NT, bool CONSUMER_BAR = false>
__device__ __forceinline__ void rn_view_finish(const RnView& vw, const RnFit& ft, const int v, const int tid,
                                               const double* fin, const double* FtFs, const double* Ssm,
                                               double* Us, double* Sn, double* red, const int fuse_finish) {
  constexpr int KK = K * K;
  const int V = ft.n_views;
  const double* GtGn = fin;
  const double* An = fin + KK;
  const double* csGn = fin + 2 * KK;
  // this runs in ONE CTA at the very end of the iteration: issue the global loads its serial part needs now
  double lam_old = 0.0, mu_old = 0.0, xn = 0.0, err_before = 0.0;
  RnCtrl cc;
  if (tid < K) {
    lam_old = vw.lam[tid];
    mu_old = vw.mu[tid];
  }
  if (tid == 0) {
    xn = vw.scal[0];
    if (fuse_finish) {  // the views before this one finished in earlier launches: their errors are final
      cc = *ft.ctrl;
      for (int w = 0; w < v; ++w) err_before += ft.views[w].scal[1];
    }
  }
  for (int o = tid; o < KK; o += NT) {
    vw.GtG[o] = GtGn[o];
    vw.A[o] = An[o];
    vw.FtF[o] = FtFs[o];
  }
  if (tid < K) {
    vw.csG[tid] = csGn[tid];
    vw.csF[tid] = FtFs[KK + tid];
  }
  for (int o = tid; o < KK; o += NT) {  // U = crossprod(F) %*% S
    const int a = o % K, c = o / K;
    double s = 0.0;
    for (int b = 0; b < K; ++b) s = fma(FtFs[a + b * K], Ssm[b + c * K], s);
    Us[a + c * K] = s;
  }
  if constexpr (CONSUMER_BAR) asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
  else __syncthreads();
  for (int o = tid; o < KK; o += NT) {  // update_s
    const int a = o % K, b = o / K;
    double D = 0.0;
    for (int c = 0; c < K; ++c) D = fma(Us[a + c * K], GtGn[c + b * K], D);
    const double N = An[o];
    const double sv = Ssm[o];
    double out;
    if (ft.xi_total == 0.0) {
      double ratio = N / D;
      if (isnan(ratio)) ratio = 1.0;
      out = fabs(sv * ratio);
    } else {
      double xisum = 0.0, xs = 0.0;
      for (int w = 0; w < V; ++w) xisum += ft.xi[w + v * V];
      for (int w = 0; w < V; ++w) {
        const double x = ft.xi[w + v * V];
        if (x != 0.0) xs += x * ft.views[w].S[o];
      }
      out = fabs(sv * ((N + xs) / (D + xisum * sv)));
    }
    Sn[o] = out;
  }
  if constexpr (CONSUMER_BAR) asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
  else __syncthreads();
  for (int o = tid; o < KK; o += NT) vw.S[o] = Sn[o];
  if (tid < K) {  // update_lm
    vw.lam[tid] = FtFs[KK + tid] * lam_old;
    vw.mu[tid] = csGn[tid] * mu_old;
  }
  // algebraic error: (||X||^2 - 2 <A,S'> + <(F'F S') G'G, S'>) / ||X||^2
  for (int o = tid; o < KK; o += NT) {
    const int a = o % K, c = o / K;
    double s = 0.0;
    for (int b = 0; b < K; ++b) s = fma(FtFs[a + b * K], Sn[b + c * K], s);
    Us[a + c * K] = s;
  }
  if constexpr (CONSUMER_BAR) asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
  else __syncthreads();
  for (int o = tid; o < KK; o += NT) {
    const int a = o % K, b = o / K;
    double q = 0.0;
    for (int c = 0; c < K; ++c) q = fma(Us[a + c * K], GtGn[c + b * K], q);
    red[o] = (q - 2.0 * An[o]) * Sn[o];
  }
  if constexpr (CONSUMER_BAR) asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
  else __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int i = 0; i < KK; ++i) s += red[i];
    const double e = (xn + s) / xn;
    vw.scal[1] = e;
    vw.scal[2] = e;
    vw.flags[0] = (ft.err_mode == 2) ? 1 : 0;  // DIRECT: the residual pass that follows overwrites scal[1]
    // AUTO: the algebraic form loses digits to cancellation -- measured relative deviation from the direct residual
    // ~1e-15 / e with a factor of 1..10 (<= 1e-10 for e in [1e-4, 1e-3), up to 8e-10 in [1e-5, 1e-4)) -- so the direct
    // residual pass takes over below RN_AUTO_DIRECT_BELOW, a factor of 10 inside the 1e-9 bar
    const bool want = ft.err_mode == 0 && e < RN_AUTO_DIRECT_BELOW;
    if (want) ft.ctrl->want_direct = 1;
    if (fuse_finish) {  // the other views' errors were written by earlier launches, this view's by this thread
      if (ft.err_mode == 0 && (want || cc.want_direct)) {
        ft.ctrl->done = 3;  // pause: host re-does the error
      } else {  // rn_finish_dev on the control block read at entry (same summation order over the views)
        double sum = err_before + e;
        for (int w = v + 1; w < V; ++w) sum += ft.views[w].scal[1];
        const double mean = sum / (double)V;
        RnCtrl* c = ft.ctrl;
        if (cc.hist_count < ft.hist_cap) ft.hist[cc.hist_count] = mean;
        c->hist_count = cc.hist_count + 1;
        c->iters = cc.iters + 1;
        const double diff = fabs(mean - cc.prev_err);
        c->last_diff = diff;
        c->prev_err = mean;
        if (cc.conv_mode) {
          if (isnan(mean)) c->done = 2;
          else if (!(diff > cc.tol)) c->done = 1;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// F step, tensor-core path (k <= 8), persistent stream-K.
//   grid = resident CTAs, 256 threads.  Unit = 64 rows x 32 data columns of X (16 KB, contiguous in
//   the panel layout).  In a unit warp w owns columns 4w..4w+3; lane (g,t) loads rows 16m+2g+{0,1}
//   (m = 0..3) of column 4w+t, so a warp reads 2 KB contiguous per unit and the CTA 16 KB.
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256, 2) rn_f_step_sk(const RnView vw, const RnFit ft, const int v) {
  constexpr int KP = 8;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  if (ft.ctrl->done) return;  // state is frozen once the stop rule fired (uniform over the grid)

  __shared__ double Ps[RN_ROW_TILE * KP];
  __shared__ double Ssm[K * K], Wsm[K * K], lamh[K];
  __shared__ int s_flag;

  if (tid < K * K) Ssm[tid] = vw.S[tid];
  if (tid < K) lamh[tid] = 0.5 * vw.lam[tid];
  __syncthreads();
  if (tid < K * K) {  // W = crossprod(G) %*% t(S)
    const int a = tid % K, b = tid / K;
    double s = 0.0;
    for (int c = 0; c < K; ++c) s = fma(vw.GtG[a + c * K], Ssm[b + c * K], s);
    Wsm[a + b * K] = s;
  }

  const int64_t pp = vw.pp;
  const int64_t UPT = pp >> 5;  // units per row tile
  const int64_t U = (int64_t)vw.row_tiles * UPT;
  const int64_t C = gridDim.x, cta = blockIdx.x;
  const RnSplit sp(U, C);
  const int64_t u0 = sp.begin(cta), u1 = sp.begin(cta + 1);
  const double* __restrict__ G = vw.G;

  for (int64_t u = u0; u < u1;) {
    const int64_t tile = u / UPT;
    const int64_t cb0 = u - tile * UPT;
    const int64_t cb1 = min(UPT, cb0 + (u1 - u));
    const bool first_seg = (u == u0);
    u += cb1 - cb0;

    double acc[4][2][2];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int h = 0; h < 2; ++h) acc[m][h][0] = acc[m][h][1] = 0.0;
    const double* xt = vw.X + (tile * pp + 4 * warp + t) * RN_ROW_TILE;
    const double* gt = G + (4 * warp + t) * KP + g;
    int xo[4];  // swizzled piece offsets of rows 16m + 2g + {0,1} in data column 4w + t (mod 4 == t)
#pragma unroll
    for (int m = 0; m < 4; ++m) xo[m] = 2 * ((8 * m + g) ^ rn_sigma(t));
    int64_t cb = cb0;
    for (; cb + 3 < cb1; cb += 4) {
      double2 x[4][4];
      double b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double* xp = xt + (cb + q) * (32 * RN_ROW_TILE);
#pragma unroll
        for (int m = 0; m < 4; ++m) x[q][m] = rn_ld_stream2(xp + xo[m]);
        b[q] = gt[(cb + q) * (32 * KP)];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          rn_dmma(acc[m][0][0], acc[m][0][1], x[q][m].x, b[q]);
          rn_dmma(acc[m][1][0], acc[m][1][1], x[q][m].y, b[q]);
        }
    }
    for (; cb < cb1; ++cb) {
      const double* xp = xt + cb * (32 * RN_ROW_TILE);
      double2 x[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) x[m] = rn_ld_stream2(xp + xo[m]);
      const double b = gt[cb * (32 * KP)];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        rn_dmma(acc[m][0][0], acc[m][0][1], x[m].x, b);
        rn_dmma(acc[m][1][0], acc[m][1][1], x[m].y, b);
      }
    }

    __syncthreads();  // previous segment's epilogue is done with Ps
    for (int i = tid; i < RN_ROW_TILE * KP; i += 256) Ps[i] = 0.0;
    __syncthreads();
    for (int w = 0; w < 8; ++w) {  // fixed warp order
      if (warp == w) {
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            // fragment-major layout: (register, lane) -> conflict-free; rn_ps_index maps (row, col) back
            Ps[((m * 2 + h) * 2 + 0) * 32 + lane] += acc[m][h][0];
            Ps[((m * 2 + h) * 2 + 1) * 32 + lane] += acc[m][h][1];
          }
      }
      __syncthreads();
    }

    if (!(cb0 == 0 && cb1 == UPT)) {  // the tile straddles CTAs: combine in CTA order
      double* mine = vw.Ppart + (cta * 2 + (first_seg ? 0 : 1)) * (RN_ROW_TILE * KP);
      for (int i = tid; i < RN_ROW_TILE * KP; i += 256) mine[i] = Ps[i];
      __threadfence();
      __syncthreads();
      const int64_t c_first = sp.owner(tile * UPT), c_last = sp.owner(tile * UPT + UPT - 1);
      if (tid == 0) s_flag = (atomicAdd(&vw.tile_ticket[tile], 1) == (int)(c_last - c_first));
      __syncthreads();
      if (!s_flag) continue;
      __threadfence();
      for (int i = tid; i < RN_ROW_TILE * KP; i += 256) {
        double s = 0.0;
        for (int64_t c2 = c_first; c2 <= c_last; ++c2) {
          const int slot = (sp.begin(c2) / UPT == tile) ? 0 : 1;
          s += __ldcg(vw.Ppart + (c2 * 2 + slot) * (RN_ROW_TILE * KP) + i);
        }
        Ps[i] = s;
      }
      if (tid == 0) vw.tile_ticket[tile] = 0;
      __syncthreads();
    }

    if (tid < RN_ROW_TILE) {
      const int64_t r = tile * RN_ROW_TILE + tid;
      if (r < vw.n) {
        double P[K];
#pragma unroll
        for (int c = 0; c < K; ++c) P[c] = Ps[rn_ps_index(tid, c)];
        rn_update_f_row<K>(vw, ft, v, r, P, Ssm, Wsm, lamh);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// G step, tensor-core path (k <= 8), persistent stream-K, fused epilogue.
//   grid = resident CTAs, 128 threads.  Unit = 64 data columns x 64 rows of X (32 KB contiguous in the
//   panel layout); units are ordered column-group-major, so a CTA streams consecutive row steps of one
//   column group.  Lane (g,t) feeds, for every 8-column block jb, column 8jb+g and rows 8i+2t+{0,1}
//   (i = 0..7).  The F fragment of a step doubles as both operands of F'F (column group 0 only).
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128, 3) rn_g_step_sk(const RnView vw, const RnFit ft, const int v,
                                                       const int fuse_finish) {
  constexpr int KP = 8, KK = K * K, NFF = KK + K, NOUT = 2 * KK + K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  if (ft.ctrl->done) return;

  __shared__ double Ts[RN_COL_GROUP * KP];
  __shared__ double Gs[RN_COL_GROUP * K];
  __shared__ double FFs[NFF], FtFs[NFF];
  __shared__ double Ssm[KK], Vs[KK], muh[K];
  __shared__ double fin[NOUT], Us[KK], Sn[KK], red[KK];
  __shared__ int s_flag;

  const int64_t pp = vw.pp;
  const int64_t NS = vw.row_tiles, NG = vw.col_groups;
  const int64_t U = NS * NG;
  const int64_t C = gridDim.x, cta = blockIdx.x;
  const RnSplit sp(U, C);
  const int64_t u0 = sp.begin(cta), u1 = sp.begin(cta + 1);
  const int nffc = (int)sp.owner(NS - 1) + 1;  // CTAs that stream part of column group 0
  const double* __restrict__ F = vw.F;
  bool ff_ready = false;

  for (int i = tid; i < KK; i += 128) Ssm[i] = vw.S[i];
  if (tid < K) muh[tid] = 0.5 * vw.mu[tid];

  for (int64_t u = u0; u < u1;) {
    const int64_t grp = u / NS;
    const int64_t s0 = u - grp * NS;
    const int64_t s1 = min(NS, s0 + (u1 - u));
    const bool first_seg = (u == u0);
    u += s1 - s0;
    const int64_t j0 = grp * RN_COL_GROUP;
    const int njb = (int)min((int64_t)8, (pp - j0) >> 3);
    const bool doFF = (grp == 0);

    double acc[8][2];
#pragma unroll
    for (int jb = 0; jb < 8; ++jb) acc[jb][0] = acc[jb][1] = 0.0;
    double aff0 = 0.0, aff1 = 0.0, acs0 = 0.0, acs1 = 0.0;

    int xo[8];  // swizzled piece offsets of rows 8i + 2t + {0,1} in data column j0 + 8jb + g (mod 4 == g mod 4)
#pragma unroll
    for (int i = 0; i < 8; ++i) xo[i] = 2 * ((4 * i + t) ^ rn_sigma(g));
    for (int64_t s = s0 + warp; s < s1; s += 4) {
      const double* fp = F + (s * KP + g) * RN_ROW_TILE;
      const double* xp = vw.X + (s * pp + j0 + g) * RN_ROW_TILE;
      double2 f2[8], xa[8], xb[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) xa[i] = rn_ld_stream2(xp + xo[i]);
#pragma unroll
      for (int i = 0; i < 8; ++i) f2[i] = *reinterpret_cast<const double2*>(fp + xo[i]);
      if (doFF) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          rn_dmma(aff0, aff1, f2[i].x, f2[i].x);
          rn_dmma(aff0, aff1, f2[i].y, f2[i].y);
          rn_dmma(acs0, acs1, 1.0, f2[i].x);
          rn_dmma(acs0, acs1, 1.0, f2[i].y);
        }
      }
#pragma unroll
      for (int jb = 0; jb < 8; jb += 2) {
        if (jb + 1 < njb) {
          const double* xq = xp + (8 * (jb + 1)) * RN_ROW_TILE;
#pragma unroll
          for (int i = 0; i < 8; ++i) xb[i] = rn_ld_stream2(xq + xo[i]);
        }
        if (jb < njb) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            rn_dmma(acc[jb][0], acc[jb][1], xa[i].x, f2[i].x);
            rn_dmma(acc[jb][0], acc[jb][1], xa[i].y, f2[i].y);
          }
        }
        if (jb + 2 < njb) {
          const double* xq = xp + (8 * (jb + 2)) * RN_ROW_TILE;
#pragma unroll
          for (int i = 0; i < 8; ++i) xa[i] = rn_ld_stream2(xq + xo[i]);
        }
        if (jb + 1 < njb) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            rn_dmma(acc[jb + 1][0], acc[jb + 1][1], xb[i].x, f2[i].x);
            rn_dmma(acc[jb + 1][0], acc[jb + 1][1], xb[i].y, f2[i].y);
          }
        }
      }
    }

    __syncthreads();  // previous segment's epilogue is done with Ts / Gs / FFs
    for (int i = tid; i < RN_COL_GROUP * KP; i += 128) Ts[i] = 0.0;
    if (tid < NFF) FFs[tid] = 0.0;
    __syncthreads();
    for (int w = 0; w < 4; ++w) {  // fixed warp order
      if (warp == w) {
#pragma unroll
        for (int jb = 0; jb < 8; ++jb) {
          Ts[(8 * jb + g) * KP + 2 * t] += acc[jb][0];
          Ts[(8 * jb + g) * KP + 2 * t + 1] += acc[jb][1];
        }
        if (doFF) {
          if (g < K) {
            if (2 * t < K) FFs[g + (2 * t) * K] += aff0;
            if (2 * t + 1 < K) FFs[g + (2 * t + 1) * K] += aff1;
          }
          if (g == 0) {
            if (2 * t < K) FFs[KK + 2 * t] += acs0;
            if (2 * t + 1 < K) FFs[KK + 2 * t + 1] += acs1;
          }
        }
      }
      __syncthreads();
    }
    if (doFF) {  // publish this CTA's F'F | colSums(F) partial
      if (tid < NFF) vw.FFpart[cta * NFF + tid] = FFs[tid];
      __threadfence();
      __syncthreads();
      if (tid == 0) atomicAdd(&vw.misc_ticket[2], 1);
    }

    if (!(s0 == 0 && s1 == NS)) {  // the column group straddles CTAs: combine in CTA order
      double* mine = vw.Tpart + (cta * 2 + (first_seg ? 0 : 1)) * (RN_COL_GROUP * KP);
      for (int i = tid; i < RN_COL_GROUP * KP; i += 128) mine[i] = Ts[i];
      __threadfence();
      __syncthreads();
      const int64_t c_first = sp.owner(grp * NS), c_last = sp.owner(grp * NS + NS - 1);
      if (tid == 0) s_flag = (atomicAdd(&vw.group_ticket[grp], 1) == (int)(c_last - c_first));
      __syncthreads();
      if (!s_flag) continue;
      __threadfence();
      for (int i = tid; i < RN_COL_GROUP * KP; i += 128) {
        double s = 0.0;
        for (int64_t c2 = c_first; c2 <= c_last; ++c2) {
          const int slot = (sp.begin(c2) / NS == grp) ? 0 : 1;
          s += __ldcg(vw.Tpart + (c2 * 2 + slot) * (RN_COL_GROUP * KP) + i);
        }
        Ts[i] = s;
      }
      if (tid == 0) vw.group_ticket[grp] = 0;
      __syncthreads();
    }

    if (fuse_finish < 0) {  // row-sharded view: T goes to HBM for the all-reduce, the epilogue is its own launch
      for (int i = tid; i < 8 * njb * KP; i += 128) vw.T[j0 * KP + i] = Ts[i];
      continue;
    }
    // ---- epilogue of this column group: update_g, then the group's G'G | A | colSums(G) partial ----
    if (!ff_ready) {
      if (tid == 0) {
        while (rn_ld_acquire(&vw.misc_ticket[2]) < nffc) __nanosleep(64);
      }
      __syncthreads();
      if (tid < NFF) {
        double s = 0.0;
        for (int i = 0; i < nffc; ++i) s += __ldcg(vw.FFpart + (int64_t)i * NFF + tid);
        FtFs[tid] = s;
      }
      __syncthreads();
      for (int o = tid; o < KK; o += 128) {  // V = crossprod(F) %*% S
        const int a = o % K, c = o / K;
        double s = 0.0;
        for (int b = 0; b < K; ++b) s = fma(FtFs[a + b * K], Ssm[b + c * K], s);
        Vs[a + c * K] = s;
      }
      __syncthreads();
      ff_ready = true;
    }
    if (tid < RN_COL_GROUP) {
      const int64_t j = j0 + tid;
      double gn[K];
      if (j < vw.p) {
        double Tj[K];
#pragma unroll
        for (int c = 0; c < K; ++c) Tj[c] = Ts[tid * KP + c];
        rn_update_g_row<K>(vw, ft, v, j, Tj, Ssm, Vs, muh, gn);
      } else {
#pragma unroll
        for (int c = 0; c < K; ++c) gn[c] = 0.0;
      }
#pragma unroll
      for (int c = 0; c < K; ++c) Gs[tid * K + c] = gn[c];
    }
    __syncthreads();
    for (int o = tid; o < NOUT; o += 128) {
      double s = 0.0;
      if (o < KK) {
        const int a = o % K, b = o / K;
        for (int i = 0; i < RN_COL_GROUP; ++i) s = fma(Gs[i * K + a], Gs[i * K + b], s);
      } else if (o < 2 * KK) {
        const int a = (o - KK) % K, b = (o - KK) / K;
        for (int i = 0; i < RN_COL_GROUP; ++i) s = fma(Ts[i * KP + a], Gs[i * K + b], s);
      } else {
        const int c = o - 2 * KK;
        for (int i = 0; i < RN_COL_GROUP; ++i) s += Gs[i * K + c];
      }
      vw.GGpart[grp * NOUT + o] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_flag = (atomicAdd(&vw.misc_ticket[0], 1) == (int)NG - 1);
    __syncthreads();
    if (!s_flag) continue;
    __threadfence();
    // ---- last column group done: finish the view --------------------------------------------------
    for (int o = tid; o < NOUT; o += 128) fin[o] = rn_sum_strided(vw.GGpart + o, NOUT, NG);
    if (tid == 0) {
      vw.misc_ticket[0] = 0;
      vw.misc_ticket[2] = 0;
    }
    __syncthreads();
    rn_view_finish<K, 128>(vw, ft, v, tid, fin, FtFs, Ssm, Us, Sn, red, fuse_finish);
  }
}

// barrier over the 256 consumer threads only (the producer warp never joins it)
__device__ __forceinline__ void rn_consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

#define RN_TMA_THREADS 288  // 8 consumer warps + 1 producer warp
// k <= 8: factor rows of 8 doubles (kp = 8), one 8-wide MMA tile in the factor dimension; k = 9..16: kp = 16, two tiles.
// The rings shrink by one stage for kp = 16 (a stage carries the factor rows of its unit, twice as many bytes).
#define RN_F_STAGES_KP(kp) ((kp) == 8 ? 11 : 10)
#define RN_F_STAGE_BYTES_KP(kp) (32 * 512 + 32 * (kp) * 8)  // one 64x32 unit of X (16 KB, verbatim) + its 32 rows of G
#define RN_G_STAGES_KP(kp) ((kp) == 8 ? 5 : 4)
#define RN_G_STAGE_BYTES_KP(kp) (64 * 512 + (kp) * 512)     // one 64x64 unit of X (32 KB, verbatim) + the step's 64 rows of F
// dynamic shared memory of the two TMA kernels (stages | 2*NST mbarriers | epilogue scratch)
static inline size_t rn_f_tma_smem(int k) {
  const int kp = k <= 8 ? 8 : 16, nst = RN_F_STAGES_KP(kp);
  return (size_t)nst * RN_F_STAGE_BYTES_KP(kp) + 2 * nst * 8 + (RN_ROW_TILE * kp + 2 * k * k + k) * 8 + 16;
}
static inline size_t rn_g_tma_smem(int k) {
  const int kp = k <= 8 ? 8 : 16, nst = RN_G_STAGES_KP(kp);
  const int nff = k * k + k;
  return (size_t)nst * RN_G_STAGE_BYTES_KP(kp) + 2 * nst * 8 +
         (RN_COL_GROUP * kp + RN_COL_GROUP * k + 8 * nff + nff + 2 * k * k + k) * 8 + 16;
}

// ------------------------------------------------------------------------------------------------
// F step, tensor-core path fed by a TMA ring (k <= 16), persistent stream-K, 1 CTA per SM.
//   Same work decomposition and epilogue as rn_f_step_sk; what differs is how X reaches the MMAs: per 64x32
//   unit one thread issues ONE 16 KB bulk copy (the unit is contiguous in the panel layout) plus a 2 KB copy
//   of the unit's 32 rows of G into an 11-stage shared-memory ring; the 8 consumer warps wait on the stage's
//   mbarrier, read their fragments with LDS.128 (conflict-free thanks to the rn_sigma swizzle baked into the
//   stored layout) and release the stage.  ~200 KB per SM are in flight at all times, independent of
//   register pressure.
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(RN_TMA_THREADS, 1) rn_f_step_tma(const RnView vw, const RnFit ft, const int v) {
  // k = 9..16: NT = 2 tiles of 8 factor columns -- the B operand (the unit's rows of G) has two fragments per lane and
  // every X fragment feeds two MMAs; accumulators, Ps and the partials carry kp = 16 columns
  constexpr int KP = K <= 8 ? 8 : 16, NT = KP / 8, NST = RN_F_STAGES_KP(KP), XB = 32 * 512, GB = 32 * KP * 8;
  constexpr int STAGE = RN_F_STAGE_BYTES_KP(KP);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  if (ft.ctrl->done) return;

  extern __shared__ __align__(128) unsigned char rn_smem[];
  unsigned char* stages = rn_smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(rn_smem + NST * STAGE);
  uint64_t* empty = full + NST;
  double* Ps = reinterpret_cast<double*>(empty + NST);  // [NT][64 x 8 fragment-major]
  double* Ssm = Ps + RN_ROW_TILE * KP;
  double* Wsm = Ssm + K * K;
  double* lamh = Wsm + K * K;
  int* s_flag = reinterpret_cast<int*>(lamh + K);

  const int64_t pp = vw.pp;
  const int64_t UPT = pp >> 5;
  const int64_t U = (int64_t)vw.row_tiles * UPT;
  const int64_t C = gridDim.x, cta = blockIdx.x;
  const RnSplit sp(U, C);
  const int64_t u0 = sp.begin(cta), u1 = sp.begin(cta + 1);

  if (tid == 0) {
    for (int i = 0; i < NST; ++i) {
      rn_mbar_init(&full[i], 1);
      rn_mbar_init(&empty[i], 8);
    }
    rn_mbar_init_fence();
  }
  if (tid < K * K) Ssm[tid] = vw.S[tid];
  if (tid < K) lamh[tid] = 0.5 * vw.lam[tid];
  __syncthreads();

  if (warp == 8) {  // ---- producer warp ------------------------------------------------------------
    int64_t it = 0;
    for (int64_t u = u0; u < u1; ++u, ++it) {
      const int64_t tile = u / UPT, cb = u - tile * UPT;
      const int st = (int)(it % NST);
      const uint32_t ph = (uint32_t)((it / NST) & 1);
      rn_mbar_wait(&empty[st], ph ^ 1u);
      unsigned char* sb = stages + st * STAGE;
      if (lane == 0) {
        rn_mbar_expect_tx(&full[st], XB + GB);
        rn_bulk_g2s(sb, vw.X + (tile * pp + 32 * cb) * RN_ROW_TILE, XB, &full[st]);
        rn_bulk_g2s(sb + XB, vw.G + (32 * cb) * KP, GB, &full[st]);
      }
      __syncwarp();
    }
    return;
  }

  // ---- consumer warps ----------------------------------------------------------------------------
  if (tid < K * K) {  // W = crossprod(G) %*% t(S)
    const int a = tid % K, b = tid / K;
    double s = 0.0;
    for (int c = 0; c < K; ++c) s = fma(vw.GtG[a + c * K], Ssm[b + c * K], s);
    Wsm[a + b * K] = s;
  }
  int64_t it = 0;
  for (int64_t u = u0; u < u1;) {
    const int64_t tile = u / UPT;
    const int64_t cb0 = u - tile * UPT;
    const int64_t cb1 = min(UPT, cb0 + (u1 - u));
    const bool first_seg = (u == u0);
    u += cb1 - cb0;

    double acc[NT][4][2][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h) acc[nt][m][h][0] = acc[nt][m][h][1] = 0.0;
    int xo[4];  // swizzled piece offsets (doubles) of rows 16m + 2g + {0,1} in column 4w + t of the unit
#pragma unroll
    for (int m = 0; m < 4; ++m) xo[m] = (4 * warp + t) * RN_ROW_TILE + 2 * ((8 * m + g) ^ rn_sigma(t));
    for (int64_t cb = cb0; cb < cb1; ++cb, ++it) {
      const int st = (int)(it % NST);
      const uint32_t ph = (uint32_t)((it / NST) & 1);
      rn_mbar_wait(&full[st], ph);
      const unsigned char* sb = stages + st * STAGE;
      const double* xs = reinterpret_cast<const double*>(sb);
      double b[NT];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) b[nt] = *(reinterpret_cast<const double*>(sb + XB) + (4 * warp + t) * KP + 8 * nt + g);
      double2 x[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) x[m] = *reinterpret_cast<const double2*>(xs + xo[m]);
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          rn_dmma(acc[nt][m][0][0], acc[nt][m][0][1], x[m].x, b[nt]);
          rn_dmma(acc[nt][m][1][0], acc[nt][m][1][1], x[m].y, b[nt]);
        }
      __syncwarp();
      if (lane == 0) rn_mbar_arrive(&empty[st]);
    }

    rn_consumer_sync();  // previous segment's epilogue is done with Ps
    for (int i = tid; i < RN_ROW_TILE * KP; i += 256) Ps[i] = 0.0;
    rn_consumer_sync();
    for (int w = 0; w < 8; ++w) {  // fixed warp order
      if (warp == w) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              // fragment-major layout: (register, lane) -> conflict-free; rn_ps_index maps (row, col) back
              Ps[nt * 512 + ((m * 2 + h) * 2 + 0) * 32 + lane] += acc[nt][m][h][0];
              Ps[nt * 512 + ((m * 2 + h) * 2 + 1) * 32 + lane] += acc[nt][m][h][1];
            }
      }
      rn_consumer_sync();
    }

    if (!(cb0 == 0 && cb1 == UPT)) {  // the tile straddles CTAs: combine in CTA order
      double* mine = vw.Ppart + (cta * 2 + (first_seg ? 0 : 1)) * (RN_ROW_TILE * KP);
      for (int i = tid; i < RN_ROW_TILE * KP; i += 256) mine[i] = Ps[i];
      __threadfence();
      rn_consumer_sync();
      const int64_t c_first = sp.owner(tile * UPT), c_last = sp.owner(tile * UPT + UPT - 1);
      if (tid == 0) *s_flag = (atomicAdd(&vw.tile_ticket[tile], 1) == (int)(c_last - c_first));
      rn_consumer_sync();
      if (!*s_flag) continue;
      __threadfence();
      for (int i = tid; i < RN_ROW_TILE * KP; i += 256) {
        double s = 0.0;
        for (int64_t c2 = c_first; c2 <= c_last; ++c2) {
          const int slot = (sp.begin(c2) / UPT == tile) ? 0 : 1;
          s += __ldcg(vw.Ppart + (c2 * 2 + slot) * (RN_ROW_TILE * KP) + i);
        }
        Ps[i] = s;
      }
      if (tid == 0) vw.tile_ticket[tile] = 0;
      rn_consumer_sync();
    }

    if (tid < RN_ROW_TILE) {
      const int64_t r = tile * RN_ROW_TILE + tid;
      if (r < vw.n) {
        double P[K];
#pragma unroll
        for (int c = 0; c < K; ++c) P[c] = Ps[(c >> 3) * 512 + rn_ps_index(tid, c & 7)];
        rn_update_f_row<K>(vw, ft, v, r, P, Ssm, Wsm, lamh);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// G step, tensor-core path fed by a TMA ring (k <= 16), persistent stream-K, fused epilogue, 1 CTA per SM.
//   Unit = 64 data columns x 64 rows of X = one contiguous 32 KB run: ONE bulk copy per unit into a 6-stage
//   ring (~190 KB in flight per SM).  Consumer warp w owns data columns 8w..8w+7 of the group for every row
//   step, so there is no cross-warp reduction of T; the F fragment of a step comes straight from global/L1
//   (all 8 warps read the same 4 KB) and is prefetched one step ahead; F'F / colSums(F) (group 0 only) are
//   split over the warps by row block.
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(RN_TMA_THREADS, 1) rn_g_step_tma(const RnView vw, const RnFit ft, const int v,
                                                                   const int fuse_finish) {
  // k = 9..16: NT = 2 tiles of 8 factor columns -- two F fragments per X fragment, two accumulator tiles per warp,
  // F'F as 2 x 2 tiles; the stage's F block is kp x 512 B
  constexpr int KP = K <= 8 ? 8 : 16, NT = KP / 8, KK = K * K, NFF = KK + K, NOUT = 2 * KK + K;
  constexpr int NST = RN_G_STAGES_KP(KP), FB = KP * 512, STAGE = RN_G_STAGE_BYTES_KP(KP);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  if (ft.ctrl->done) return;

  extern __shared__ __align__(128) unsigned char rn_smem[];
  unsigned char* stages = rn_smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(rn_smem + NST * STAGE);
  uint64_t* empty = full + NST;
  double* Ts = reinterpret_cast<double*>(empty + NST);  // [64][KP]
  double* Gs = Ts + RN_COL_GROUP * KP;                   // [64][K]
  double* FFw = Gs + RN_COL_GROUP * K;                   // [8][NFF] per-warp partials; later fin / scratch
  double* FtFs = FFw + 8 * NFF;                          // [NFF]
  double* Ssm = FtFs + NFF;
  double* Vs = Ssm + KK;
  double* muh = Vs + KK;
  int* s_flag = reinterpret_cast<int*>(muh + K);
  static_assert(NOUT + 3 * KK <= 8 * NFF, "last-CTA scratch is aliased onto the per-warp F'F partials");
  double* fin = FFw;  // used only after the F'F partials have been published
  double* Us = fin + NOUT;
  double* Sn = Us + KK;
  double* red = Sn + KK;

  const int64_t pp = vw.pp;
  const int64_t NS = vw.row_tiles, NG = vw.col_groups;
  const int64_t U = NS * NG;
  const int64_t C = gridDim.x, cta = blockIdx.x;
  const RnSplit sp(U, C);
  const int64_t u0 = sp.begin(cta), u1 = sp.begin(cta + 1);
  const int nffc = (int)sp.owner(NS - 1) + 1;  // CTAs that stream part of column group 0

  if (tid == 0) {
    for (int i = 0; i < NST; ++i) {
      rn_mbar_init(&full[i], 1);
      rn_mbar_init(&empty[i], 8);
    }
    rn_mbar_init_fence();
  }
  for (int i = tid; i < KK; i += RN_TMA_THREADS) Ssm[i] = vw.S[i];
  if (tid < K) muh[tid] = 0.5 * vw.mu[tid];
  __syncthreads();

  if (warp == 8) {  // ---- producer warp ------------------------------------------------------------
    int64_t it = 0;
    for (int64_t u = u0; u < u1; ++u, ++it) {
      const int64_t grp = u / NS, s = u - grp * NS;
      const int64_t j0 = grp * RN_COL_GROUP;
      const int ncol = (int)min((int64_t)RN_COL_GROUP, pp - j0);
      const int st = (int)(it % NST);
      const uint32_t ph = (uint32_t)((it / NST) & 1);
      rn_mbar_wait(&empty[st], ph ^ 1u);
      unsigned char* sb = stages + st * STAGE;
      if (lane == 0) {
        rn_mbar_expect_tx(&full[st], (uint32_t)ncol * 512u + (uint32_t)FB);
        rn_bulk_g2s(sb, vw.X + (s * pp + j0) * RN_ROW_TILE, (uint32_t)ncol * 512u, &full[st]);
        rn_bulk_g2s(sb + 32768, vw.F + s * (KP * RN_ROW_TILE), (uint32_t)FB, &full[st]);
      }
      __syncwarp();
    }
    return;
  }

  // ---- consumer warps ----------------------------------------------------------------------------
  bool ff_ready = false;
  int64_t it = 0;
  for (int64_t u = u0; u < u1;) {
    const int64_t grp = u / NS;
    const int64_t s0 = u - grp * NS;
    const int64_t s1 = min(NS, s0 + (u1 - u));
    const bool first_seg = (u == u0);
    u += s1 - s0;
    const int64_t j0 = grp * RN_COL_GROUP;
    const int njb = (int)min((int64_t)8, (pp - j0) >> 3);
    const bool doFF = (grp == 0);
    const bool have_cols = warp < njb;

    double a0[NT], a1[NT], b0[NT], b1[NT];  // two MMA chains (even / odd i) per tile of 8 factor columns
    double aff[NT][NT][2], acs[NT][2];       // F'F tile (ta, tb): rows 8 ta + g, columns 8 tb + 2t + {0,1}; colSums(F)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      a0[nt] = a1[nt] = b0[nt] = b1[nt] = 0.0;
      acs[nt][0] = acs[nt][1] = 0.0;
#pragma unroll
      for (int n2 = 0; n2 < NT; ++n2) aff[nt][n2][0] = aff[nt][n2][1] = 0.0;
    }
    int xo[8];  // swizzled piece offsets (doubles) of rows 8i + 2t + {0,1} in column 8w + g of the unit
#pragma unroll
    for (int i = 0; i < 8; ++i) xo[i] = (8 * warp + g) * RN_ROW_TILE + 2 * ((4 * i + t) ^ rn_sigma(g));
    int fo[8];  // F fragment (column c = g of tile 0, same rows; tile nt: + 512 nt) inside the stage's F block
#pragma unroll
    for (int i = 0; i < 8; ++i) fo[i] = 4096 + g * RN_ROW_TILE + 2 * ((4 * i + t) ^ rn_sigma(g));
    for (int64_t s = s0; s < s1; ++s, ++it) {
      const int st = (int)(it % NST);
      const uint32_t ph = (uint32_t)((it / NST) & 1);
      rn_mbar_wait(&full[st], ph);
      const double* xs = reinterpret_cast<const double*>(stages + st * STAGE);
      double2 fw[NT];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        fw[nt] = make_double2(0.0, 0.0);
        // sigma depends on the column modulo 4 only: column 8 nt + g has the swizzle of column g
        if (doFF) fw[nt] = *reinterpret_cast<const double2*>(xs + 4096 + (8 * nt + g) * RN_ROW_TILE + 2 * ((4 * warp + t) ^ rn_sigma(g)));
      }
      if (have_cols) {
        double2 x2[8], f2[NT][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          x2[i] = *reinterpret_cast<const double2*>(xs + xo[i]);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) f2[nt][i] = *reinterpret_cast<const double2*>(xs + fo[i] + 512 * nt);
        }
#pragma unroll
        for (int i = 0; i < 8; i += 2)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            rn_dmma(a0[nt], a1[nt], x2[i].x, f2[nt][i].x);
            rn_dmma(b0[nt], b1[nt], x2[i + 1].x, f2[nt][i + 1].x);
            rn_dmma(a0[nt], a1[nt], x2[i].y, f2[nt][i].y);
            rn_dmma(b0[nt], b1[nt], x2[i + 1].y, f2[nt][i + 1].y);
          }
      }
      __syncwarp();
      if (lane == 0) rn_mbar_arrive(&empty[st]);
      if (doFF) {  // warp w covers rows 8w + 2t + {0,1} of the step
#pragma unroll
        for (int ta = 0; ta < NT; ++ta) {
#pragma unroll
          for (int tb = 0; tb < NT; ++tb) {
            rn_dmma(aff[ta][tb][0], aff[ta][tb][1], fw[ta].x, fw[tb].x);
            rn_dmma(aff[ta][tb][0], aff[ta][tb][1], fw[ta].y, fw[tb].y);
          }
          rn_dmma(acs[ta][0], acs[ta][1], 1.0, fw[ta].x);
          rn_dmma(acs[ta][0], acs[ta][1], 1.0, fw[ta].y);
        }
      }
    }

    rn_consumer_sync();  // previous segment's epilogue is done with Ts / Gs / FFw
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      Ts[(8 * warp + g) * KP + 8 * nt + 2 * t] = a0[nt] + b0[nt];
      Ts[(8 * warp + g) * KP + 8 * nt + 2 * t + 1] = a1[nt] + b1[nt];
    }
    if (doFF) {
      double* mine = FFw + warp * NFF;
#pragma unroll
      for (int ta = 0; ta < NT; ++ta) {
        const int ra = 8 * ta + g;
#pragma unroll
        for (int tb = 0; tb < NT; ++tb) {
          const int cb = 8 * tb + 2 * t;
          if (ra < K) {
            if (cb < K) mine[ra + cb * K] = aff[ta][tb][0];
            if (cb + 1 < K) mine[ra + (cb + 1) * K] = aff[ta][tb][1];
          }
        }
        if (g == 0) {
          const int cb = 8 * ta + 2 * t;
          if (cb < K) mine[KK + cb] = acs[ta][0];
          if (cb + 1 < K) mine[KK + cb + 1] = acs[ta][1];
        }
      }
    }
    rn_consumer_sync();
    if (doFF) {  // publish this CTA's F'F | colSums(F) partial (warps summed in order)
      for (int o = tid; o < NFF; o += 256) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += FFw[w * NFF + o];
        vw.FFpart[cta * NFF + o] = s;
      }
      __threadfence();
      rn_consumer_sync();
      if (tid == 0) atomicAdd(&vw.misc_ticket[2], 1);
    }

    if (!(s0 == 0 && s1 == NS)) {  // the column group straddles CTAs: combine in CTA order
      double* mine = vw.Tpart + (cta * 2 + (first_seg ? 0 : 1)) * (RN_COL_GROUP * KP);
      for (int i = tid; i < RN_COL_GROUP * KP; i += 256) mine[i] = Ts[i];
      __threadfence();
      rn_consumer_sync();
      const int64_t c_first = sp.owner(grp * NS), c_last = sp.owner(grp * NS + NS - 1);
      if (tid == 0) *s_flag = (atomicAdd(&vw.group_ticket[grp], 1) == (int)(c_last - c_first));
      rn_consumer_sync();
      if (!*s_flag) continue;
      __threadfence();
      for (int i = tid; i < RN_COL_GROUP * KP; i += 256) {
        double s = 0.0;
        for (int64_t c2 = c_first; c2 <= c_last; ++c2) {
          const int slot = (sp.begin(c2) / NS == grp) ? 0 : 1;
          s += __ldcg(vw.Tpart + (c2 * 2 + slot) * (RN_COL_GROUP * KP) + i);
        }
        Ts[i] = s;
      }
      if (tid == 0) vw.group_ticket[grp] = 0;
      rn_consumer_sync();
    }

    if (fuse_finish < 0) {  // row-sharded view: T goes to HBM for the all-reduce, the epilogue is its own launch
      for (int i = tid; i < 8 * njb * KP; i += 256) vw.T[j0 * KP + i] = Ts[i];
      continue;
    }
    // ---- epilogue of this column group: update_g, then the group's G'G | A | colSums(G) partial ----
    if (!ff_ready) {
      if (tid == 0) {
        while (rn_ld_acquire(&vw.misc_ticket[2]) < nffc) __nanosleep(64);
      }
      rn_consumer_sync();
      for (int o = tid; o < NFF; o += 256) {
        double s = 0.0;
        for (int i = 0; i < nffc; ++i) s += __ldcg(vw.FFpart + (int64_t)i * NFF + o);
        FtFs[o] = s;
      }
      rn_consumer_sync();
      for (int o = tid; o < KK; o += 256) {  // V = crossprod(F) %*% S
        const int a = o % K, c = o / K;
        double s = 0.0;
        for (int b = 0; b < K; ++b) s = fma(FtFs[a + b * K], Ssm[b + c * K], s);
        Vs[a + c * K] = s;
      }
      rn_consumer_sync();
      ff_ready = true;
    }
    if (tid < RN_COL_GROUP) {
      const int64_t j = j0 + tid;
      double gn[K];
      if (j < vw.p) {
        double Tj[K];
#pragma unroll
        for (int c = 0; c < K; ++c) Tj[c] = Ts[tid * KP + c];
        rn_update_g_row<K>(vw, ft, v, j, Tj, Ssm, Vs, muh, gn);
      } else {
#pragma unroll
        for (int c = 0; c < K; ++c) gn[c] = 0.0;
      }
#pragma unroll
      for (int c = 0; c < K; ++c) Gs[tid * K + c] = gn[c];
    }
    rn_consumer_sync();
    for (int o = tid; o < NOUT; o += 256) {
      double s = 0.0;
      if (o < KK) {
        const int a = o % K, b = o / K;
        for (int i = 0; i < RN_COL_GROUP; ++i) s = fma(Gs[i * K + a], Gs[i * K + b], s);
      } else if (o < 2 * KK) {
        const int a = (o - KK) % K, b = (o - KK) / K;
        for (int i = 0; i < RN_COL_GROUP; ++i) s = fma(Ts[i * KP + a], Gs[i * K + b], s);
      } else {
        const int c = o - 2 * KK;
        for (int i = 0; i < RN_COL_GROUP; ++i) s += Gs[i * K + c];
      }
      vw.GGpart[grp * NOUT + o] = s;
    }
    __threadfence();
    rn_consumer_sync();
    if (tid == 0) *s_flag = (atomicAdd(&vw.misc_ticket[0], 1) == (int)NG - 1);
    rn_consumer_sync();
    if (!*s_flag) continue;
    __threadfence();
    // ---- last column group done: finish the view --------------------------------------------------
    for (int o = tid; o < NOUT; o += 256) fin[o] = rn_sum_strided(vw.GGpart + o, NOUT, NG);
    if (tid == 0) {
      vw.misc_ticket[0] = 0;
      vw.misc_ticket[2] = 0;
    }
    rn_consumer_sync();
    rn_view_finish<K, 256, true>(vw, ft, v, tid, fin, FtFs, Ssm, Us, Sn, red, fuse_finish);
  }
}

// ------------------------------------------------------------------------------------------------
// F step, CUDA-core path (any k <= 16).  grid (row_tiles, cs), 256 threads.  A CTA owns 64 rows and the
// data columns of split `cs`; lanes own rows 2*lane+{0,1}, warp w takes columns jbeg+w, +8, ...  With
// cs > 1 the last CTA to arrive for a row tile sums the partials in split order and runs the epilogue.
// ------------------------------------------------------------------------------------------------
template <int K, int KP>
__global__ void __launch_bounds__(256) rn_f_step_dfma(const RnView vw, const RnFit ft, const int v) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, cs = blockIdx.y, ncs = gridDim.y;
  if (ft.ctrl->done) return;

  __shared__ double Ps[RN_ROW_TILE * KP];
  __shared__ double Ssm[K * K], Wsm[K * K], lamh[K];
  __shared__ int s_last;

  const int64_t pp = vw.pp;
  const int nb = (int)(pp >> 3);
  const int64_t jbeg = 8 * ((int64_t)nb * cs / ncs);
  const int64_t jend = 8 * ((int64_t)nb * (cs + 1) / ncs);
  const double* __restrict__ G = vw.G;

  for (int i = tid; i < RN_ROW_TILE * KP; i += 256) Ps[i] = 0.0;

  double a0[K], a1[K];
#pragma unroll
  for (int c = 0; c < K; ++c) a0[c] = a1[c] = 0.0;
  const double* xr = vw.X + (int64_t)tile * pp * RN_ROW_TILE;
  constexpr int U = 4;
  int64_t j = jbeg + warp;
  const int xo = 2 * (lane ^ rn_sigma(j));  // j advances in steps of 8: j mod 4 (hence the swizzle) is fixed
  for (; j + 8 * (U - 1) < jend; j += 8 * U) {
    double2 x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) x[u] = rn_ld_stream2(xr + (j + 8 * u) * RN_ROW_TILE + xo);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double* gr = G + (j + 8 * u) * KP;
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const double gv = gr[c];
        a0[c] = fma(x[u].x, gv, a0[c]);
        a1[c] = fma(x[u].y, gv, a1[c]);
      }
    }
  }
  for (; j < jend; j += 8) {
    const double2 x = rn_ld_stream2(xr + j * RN_ROW_TILE + xo);
    const double* gr = G + j * KP;
#pragma unroll
    for (int c = 0; c < K; ++c) {
      const double gv = gr[c];
      a0[c] = fma(x.x, gv, a0[c]);
      a1[c] = fma(x.y, gv, a1[c]);
    }
  }
  __syncthreads();
  for (int w = 0; w < 8; ++w) {
    if (warp == w) {
#pragma unroll
      for (int c = 0; c < K; ++c) {
        Ps[(2 * lane) * KP + c] += a0[c];
        Ps[(2 * lane + 1) * KP + c] += a1[c];
      }
    }
    __syncthreads();
  }

  if (ncs > 1) {
    double* mine = vw.Ppart + ((int64_t)cs * vw.row_tiles + tile) * (RN_ROW_TILE * KP);
    for (int i = tid; i < RN_ROW_TILE * KP; i += 256) mine[i] = Ps[i];
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&vw.tile_ticket[tile], 1) == ncs - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int i = tid; i < RN_ROW_TILE * KP; i += 256) {
      double s = 0.0;
      for (int c2 = 0; c2 < ncs; ++c2)
        s += __ldcg(vw.Ppart + ((int64_t)c2 * vw.row_tiles + tile) * (RN_ROW_TILE * KP) + i);
      Ps[i] = s;
    }
    if (tid == 0) vw.tile_ticket[tile] = 0;
  }

  if (tid < K * K) Ssm[tid] = vw.S[tid];
  if (tid < K) lamh[tid] = 0.5 * vw.lam[tid];
  __syncthreads();
  if (tid < K * K) {  // W = crossprod(G) %*% t(S)
    const int a = tid % K, b = tid / K;
    double s = 0.0;
    for (int c = 0; c < K; ++c) s = fma(vw.GtG[a + c * K], Ssm[b + c * K], s);
    Wsm[a + b * K] = s;
  }
  __syncthreads();
  if (tid < RN_ROW_TILE) {
    const int64_t r = (int64_t)tile * RN_ROW_TILE + tid;
    if (r < vw.n) {
      double P[K];
#pragma unroll
      for (int c = 0; c < K; ++c) P[c] = Ps[tid * KP + c];
      rn_update_f_row<K>(vw, ft, v, r, P, Ssm, Wsm, lamh);
    }
  }
}

// F'F and colSums(F) partials for the CUDA-core path.  grid (nff), 256 threads; FFpart[cta][k*k+k].
template <int K>
__global__ void __launch_bounds__(256) rn_gram_f(const RnView vw, const RnFit ft) {
  const int tid = threadIdx.x;
  if (ft.ctrl->done) return;
  constexpr int NOUT = K * K + K;
  constexpr int PER = (NOUT + 255) / 256;  // outputs per thread (2 for k = 16)
  __shared__ double rows[256 * K];
  const int64_t ldx = vw.ldx;
  const int nchunks = (int)((ldx + 255) / 256);
  const int c0 = (int)((int64_t)nchunks * blockIdx.x / gridDim.x);
  const int c1 = (int)((int64_t)nchunks * (blockIdx.x + 1) / gridDim.x);
  double acc[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q) acc[q] = 0.0;
  for (int ch = c0; ch < c1; ++ch) {
    const int64_t r = (int64_t)ch * 256 + tid;
#pragma unroll
    for (int c = 0; c < K; ++c) rows[tid * K + c] = (r < vw.n) ? vw.F[rn_fidx(r, c, vw.kp)] : 0.0;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int o = tid + 256 * q;
      if (o < K * K) {
        const int a = o % K, b = o / K;
        for (int i = 0; i < 256; ++i) acc[q] = fma(rows[i * K + a], rows[i * K + b], acc[q]);
      } else if (o < NOUT) {
        const int b = o - K * K;
        for (int i = 0; i < 256; ++i) acc[q] += rows[i * K + b];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int o = tid + 256 * q;
    if (o < NOUT) vw.FFpart[(int64_t)blockIdx.x * NOUT + o] = acc[q];
  }
}

// ------------------------------------------------------------------------------------------------
// G stream, CUDA-core path:  T = X'.F.  grid (col_groups32, rs), 256 threads.  Warp w owns data
// columns j0 + 4w .. +3, lanes own rows 2*lane + {0,1} of every 64-row step; F rows come through L1.
// ------------------------------------------------------------------------------------------------
template <int K, int KP>
__global__ void __launch_bounds__(256) rn_g_stream_dfma(const RnView vw, const RnFit ft) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = blockIdx.x, rs = blockIdx.y, nrs = gridDim.y;
  if (ft.ctrl->done) return;
  __shared__ double Ts[RN_COL_GROUP_DFMA * KP];
  __shared__ int s_last;

  const int64_t pp = vw.pp;
  const int ns = vw.row_tiles;
  const int s0 = (int)((int64_t)ns * rs / nrs), s1 = (int)((int64_t)ns * (rs + 1) / nrs);
  const int64_t j0 = (int64_t)grp * RN_COL_GROUP_DFMA;
  const int64_t jc = j0 + 4 * warp;
  const int ncols = (int)min((int64_t)RN_COL_GROUP_DFMA, pp - j0);
  const bool active = jc < pp;
  const double* __restrict__ F = vw.F;

  double acc[4][K];
#pragma unroll
  for (int jj = 0; jj < 4; ++jj)
#pragma unroll
    for (int c = 0; c < K; ++c) acc[jj][c] = 0.0;

  if (active) {
    for (int s = s0; s < s1; ++s) {
      const double* xp = vw.X + ((int64_t)s * pp + jc) * RN_ROW_TILE;
      double2 x[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)  // jc is a multiple of 4: the swizzle of column jc + jj is sigma(jj)
        x[jj] = rn_ld_stream2(xp + jj * RN_ROW_TILE + 2 * (lane ^ rn_sigma(jj)));
      double2 f2[K];
#pragma unroll
      for (int c = 0; c < K; ++c)
        f2[c] = *reinterpret_cast<const double2*>(F + ((int64_t)s * KP + c) * RN_ROW_TILE + 2 * (lane ^ rn_sigma(c)));
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int c = 0; c < K; ++c) acc[jj][c] = fma(x[jj].y, f2[c].y, fma(x[jj].x, f2[c].x, acc[jj][c]));
    }
  }
  for (int i = tid; i < RN_COL_GROUP_DFMA * KP; i += 256) Ts[i] = 0.0;
  __syncthreads();
#pragma unroll
  for (int jj = 0; jj < 4; ++jj)
#pragma unroll
    for (int c = 0; c < K; ++c) {
      const double s = rn_warp_sum(acc[jj][c]);
      if (lane == 0) Ts[(4 * warp + jj) * KP + c] = s;
    }
  __syncthreads();

  if (nrs == 1) {
    for (int i = tid; i < ncols * KP; i += 256) vw.T[j0 * KP + i] = Ts[i];
    return;
  }
  double* mine = vw.Tpart + ((int64_t)rs * pp + j0) * KP;
  for (int i = tid; i < ncols * KP; i += 256) mine[i] = Ts[i];
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&vw.group_ticket[grp], 1) == nrs - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = tid; i < ncols * KP; i += 256) {
    double s = 0.0;
    for (int r2 = 0; r2 < nrs; ++r2) s += __ldcg(vw.Tpart + ((int64_t)r2 * pp + j0) * KP + i);
    vw.T[j0 * KP + i] = s;
  }
  if (tid == 0) vw.group_ticket[grp] = 0;
}

// ------------------------------------------------------------------------------------------------
// Stand-alone G epilogue (CUDA-core path, and the row-sharded path where T / F'F arrive from an
// all-reduce): update_g from T in HBM, partial G'G | A | colSums(G); the last CTA finishes the view.
//   grid (gepi_ctas), 256 threads (128 for k > 8), one data column (row of G) per thread.
// ------------------------------------------------------------------------------------------------
template <int K, int KP>
__global__ void __launch_bounds__(RN_GEPI_THREADS(K)) rn_g_epilogue(const RnView vw, const RnFit ft, const int v,
                                                                    const int fuse_finish) {
  constexpr int NT = RN_GEPI_THREADS(K);
  constexpr int KK = K * K;
  constexpr int NOUT = 2 * KK + K;
  const int tid = threadIdx.x;
  if (ft.ctrl->done) return;

  __shared__ double Ssm[KK], FtFs[KK + K], Vs[KK], muh[K];
  __shared__ double Tsm[NT * K], Gsm[NT * K];
  __shared__ double Us[KK], Sn[KK];
  __shared__ int s_last;
  static_assert(NOUT <= NT * K && KK <= NT * K, "last-CTA scratch is aliased onto the row buffers");
  double* const fin = Tsm;  // only used by the last CTA, after every thread is done with Tsm / Gsm
  double* const red = Gsm;

  for (int o = tid; o < KK; o += NT) Ssm[o] = vw.S[o];
  if (tid < K) muh[tid] = 0.5 * vw.mu[tid];
  for (int o = tid; o < KK + K; o += NT) {  // F'F | colSums(F): fixed-order sum of the partials
    double s = 0.0;
    for (int i = 0; i < vw.nff; ++i) s += vw.FFpart[(int64_t)i * (KK + K) + o];
    FtFs[o] = s;
  }
  __syncthreads();
  for (int o = tid; o < KK; o += NT) {  // V = crossprod(F) %*% S
    const int a = o % K, c = o / K;
    double s = 0.0;
    for (int b = 0; b < K; ++b) s = fma(FtFs[a + b * K], Ssm[b + c * K], s);
    Vs[a + c * K] = s;
  }
  __syncthreads();

  const int64_t j = (int64_t)blockIdx.x * NT + tid;
  {
    double Tj[K], gn[K];
    if (j < vw.p) {
#pragma unroll
      for (int c = 0; c < K; ++c) Tj[c] = vw.T[j * KP + c];
      rn_update_g_row<K>(vw, ft, v, j, Tj, Ssm, Vs, muh, gn);
    } else {
#pragma unroll
      for (int c = 0; c < K; ++c) Tj[c] = gn[c] = 0.0;
    }
#pragma unroll
    for (int c = 0; c < K; ++c) {
      Tsm[tid * K + c] = Tj[c];
      Gsm[tid * K + c] = gn[c];
    }
  }
  __syncthreads();
  for (int o = tid; o < NOUT; o += NT) {
    double s = 0.0;
    if (o < KK) {
      const int a = o % K, b = o / K;
      for (int i = 0; i < NT; ++i) s = fma(Gsm[i * K + a], Gsm[i * K + b], s);
    } else if (o < 2 * KK) {
      const int a = (o - KK) % K, b = (o - KK) / K;
      for (int i = 0; i < NT; ++i) s = fma(Tsm[i * K + a], Gsm[i * K + b], s);
    } else {
      const int c = o - 2 * KK;
      for (int i = 0; i < NT; ++i) s += Gsm[i * K + c];
    }
    vw.GGpart[(int64_t)blockIdx.x * NOUT + o] = s;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&vw.misc_ticket[0], 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid == 0) vw.misc_ticket[0] = 0;
  for (int o = tid; o < NOUT; o += NT) fin[o] = rn_sum_strided(vw.GGpart + o, NOUT, (int64_t)gridDim.x);
  __syncthreads();
  rn_view_finish<K, NT>(vw, ft, v, tid, fin, FtFs, Ssm, Us, Sn, red, fuse_finish);
}

// ------------------------------------------------------------------------------------------------
// Direct residual: sum (X - (F S) G')^2, one extra pass over X.  grid (row_tiles, resid_cs), 256 thr.
// Runs only when flags[0] is set (error mode DIRECT) or when forced (the host's AUTO hand-over).
// ------------------------------------------------------------------------------------------------
template <int K, int KP>
__global__ void __launch_bounds__(256) rn_residual(const RnView vw, const RnFit ft, const int force) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (!force && (ft.ctrl->done || !vw.flags[0])) return;
  __shared__ double Ssm[K * K];
  __shared__ double wsum[8];
  __shared__ double bs[256];
  __shared__ int s_last;
  const int tile = blockIdx.x, cs = blockIdx.y, ncs = gridDim.y;
  const int64_t pp = vw.pp;
  const int nb = (int)(pp >> 3);
  const int64_t jbeg = 8 * ((int64_t)nb * cs / ncs), jend = 8 * ((int64_t)nb * (cs + 1) / ncs);
  if (tid < K * K) Ssm[tid] = vw.S[tid];
  __syncthreads();
  double fs0[K], fs1[K];
  {
    double f0[K], f1[K];
#pragma unroll
    for (int c = 0; c < K; ++c) {
      const double2 f = *reinterpret_cast<const double2*>(vw.F + ((int64_t)tile * KP + c) * RN_ROW_TILE +
                                                          2 * (lane ^ rn_sigma(c)));
      f0[c] = f.x;
      f1[c] = f.y;
    }
#pragma unroll
    for (int a = 0; a < K; ++a) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int b = 0; b < K; ++b) {
        s0 = fma(f0[b], Ssm[b + a * K], s0);
        s1 = fma(f1[b], Ssm[b + a * K], s1);
      }
      fs0[a] = s0;
      fs1[a] = s1;
    }
  }
  double acc = 0.0;
  const double* xr = vw.X + (int64_t)tile * pp * RN_ROW_TILE;
  for (int64_t j = jbeg + warp; j < jend; j += 8) {
    const double2 x = rn_ld_stream2(xr + j * RN_ROW_TILE + 2 * (lane ^ rn_sigma(j)));
    const double* gr = vw.G + j * KP;
    double h0 = 0.0, h1 = 0.0;
#pragma unroll
    for (int c = 0; c < K; ++c) {
      const double gv = gr[c];
      h0 = fma(fs0[c], gv, h0);
      h1 = fma(fs1[c], gv, h1);
    }
    const double d0 = x.x - h0, d1 = x.y - h1;
    acc = fma(d0, d0, acc);
    acc = fma(d1, d1, acc);
  }
  acc = rn_warp_sum(acc);
  if (lane == 0) wsum[warp] = acc;
  __syncthreads();
  const int nblk = gridDim.x * gridDim.y;
  const int bid = blockIdx.y * gridDim.x + blockIdx.x;
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += wsum[w];
    vw.Rpart[bid] = s;
    __threadfence();
    s_last = (atomicAdd(&vw.misc_ticket[1], 1) == nblk - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double s = 0.0;
  for (int i = tid; i < nblk; i += 256) s += __ldcg(vw.Rpart + i);
  bs[tid] = s;  // fixed-order block sum
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) bs[tid] += bs[tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    vw.misc_ticket[1] = 0;
    if (!vw.sharded) {
      const double e = bs[0] / vw.scal[0];
      vw.scal[1] = e;
      vw.scal[3] = e;
    } else {
      vw.scal[3] = bs[0];  // local residual sum; the host all-reduces and divides
    }
    ft.ctrl->direct_passes += 1;
  }
}

// The kernels below are not templates: they belong to ONE translation unit (resnmtf_capi.cu); a unit that only
// instantiates the templates above (rn_tma_gt8.cu) defines RN_KERNEL_TEMPLATES_ONLY.
#ifndef RN_KERNEL_TEMPLATES_ONLY
// Row-sharded path: sums the `count` F'F | colSums(F) partials in order into the tail of the T buffer
// (T | F'F | colSums(F) is then one contiguous all-reduce) and re-arms the F'F publication counter.
__global__ void rn_pack_ff(const RnView vw, const RnFit ft, const int count) {
  if (ft.ctrl->done) return;
  const int nff = vw.k * vw.k + vw.k;
  double* tail = vw.T + vw.pp * vw.kp;
  for (int o = threadIdx.x; o < nff; o += blockDim.x) {
    double s = 0.0;
    for (int i = 0; i < count; ++i) s += vw.FFpart[(int64_t)i * nff + o];
    tail[o] = s;
  }
  if (threadIdx.x == 0) vw.misc_ticket[2] = 0;
}

// Row-sharded path: the residual kernel left the all-reduced sum of squares in scal[3]; turn it into the error.
__global__ void rn_residual_scale(const RnView vw, const RnFit ft, const int force) {
  if (!force && (ft.ctrl->done || !vw.flags[0])) return;
  const double e = vw.scal[3] / vw.scal[0];
  vw.scal[1] = e;
  vw.scal[3] = e;
}

// Iteration bookkeeping as its own launch (DIRECT error mode, and the AUTO hand-over).  <<<1,1>>>.
__global__ void rn_finish(const RnFit ft, const int force) {
  if (!force && ft.ctrl->done) return;
  rn_finish_dev(ft);
}

// ------------------------------------------------------------------------------------------------
// set-up / tear-down kernels (once per fit, not on the per-iteration path)
// ------------------------------------------------------------------------------------------------

// Re-tiles `ncols` columns of a column-major source (leading dimension lds, n valid rows) into the
// panel layout X[tile][col][64] (16-byte pieces swizzled by rn_sigma) starting at data column col0.  One
// thread per 16-byte piece; padding rows of the last tile are written as zero.  grid-stride.
__global__ void __launch_bounds__(256) rn_to_panels(const RnView vw, const double* __restrict__ src, int64_t lds,
                                                    int64_t col0, int64_t ncols) {
  const int64_t pieces = (int64_t)vw.row_tiles * ncols * 32;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < pieces; i += (int64_t)gridDim.x * 256) {
    const int64_t h = i & 31;             // double2 index inside the 64-row column piece
    const int64_t cj = (i >> 5) % ncols;  // column inside the chunk
    const int64_t tile = (i >> 5) / ncols;
    const int64_t r = tile * RN_ROW_TILE + 2 * h;
    double2 val;
    val.x = (r < vw.n) ? src[cj * lds + r] : 0.0;
    val.y = (r + 1 < vw.n) ? src[cj * lds + r + 1] : 0.0;
    *reinterpret_cast<double2*>(vw.X + (tile * vw.pp + col0 + cj) * RN_ROW_TILE +
                                2 * (h ^ rn_sigma(col0 + cj))) = val;
  }
}

// ||X||_F^2, deterministic two-stage.  grid (nblk), 256 threads.
__global__ void __launch_bounds__(256) rn_xnorm2(const RnView vw, double* part, int32_t* ticket) {
  const int tid = threadIdx.x;
  const int64_t total2 = vw.ldx * vw.pp / 2;  // padding is zero
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + tid; i < total2; i += (int64_t)gridDim.x * 256) {
    const double2 x = rn_ld_stream2(vw.X + 2 * i);
    acc = fma(x.x, x.x, acc);
    acc = fma(x.y, x.y, acc);
  }
  __shared__ double bs[256];
  __shared__ int s_last;
  bs[tid] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) bs[tid] += bs[tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    part[blockIdx.x] = bs[0];
    __threadfence();
    s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double s = 0.0;
  for (int i = tid; i < (int)gridDim.x; i += 256) s += __ldcg(part + i);
  bs[tid] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) bs[tid] += bs[tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    vw.scal[0] = bs[0];
    *ticket = 0;
  }
}

// Column sums of F and G plus G'G of the current factors: used after set_factors (lambda/mu defaults, the G'G the
// first F step needs) and by the final normalisation.  grid (2k + k*k), 1024 threads: one CTA per output value,
// fixed-order tree inside the CTA (bit-reproducible).
__global__ void __launch_bounds__(1024) rn_factor_sums(const RnView vw) {
  const int tid = threadIdx.x;
  const int K = vw.k, KP = vw.kp;
  const int o = blockIdx.x;
  __shared__ double bs[1024];
  double s = 0.0;
  if (o < K) {
    for (int64_t r = tid; r < vw.n; r += 1024) s += vw.F[rn_fidx(r, o, KP)];
  } else if (o < 2 * K) {
    const int c = o - K;
    for (int64_t j = tid; j < vw.p; j += 1024) s += vw.G[j * KP + c];
  } else {
    const int a = (o - 2 * K) % K, b = (o - 2 * K) / K;
    for (int64_t j = tid; j < vw.p; j += 1024) s = fma(vw.G[j * KP + a], vw.G[j * KP + b], s);
  }
  bs[tid] = s;
  __syncthreads();
  for (int w = 512; w > 0; w >>= 1) {
    if (tid < w) bs[tid] += bs[tid + w];
    __syncthreads();
  }
  if (tid == 0) {
    if (o < K) vw.csF[o] = bs[0];
    else if (o < 2 * K) vw.csG[o - K] = bs[0];
    else vw.GtG[o - 2 * K] = bs[0];
  }
}

// lambda <- colSums(F), mu <- colSums(G)  (R/update_steps.r:53-54) when the caller passed none.
__global__ void rn_default_lm(const RnView vw, int set_lam, int set_mu) {
  const int c = threadIdx.x;
  if (c < vw.k) {
    if (set_lam) vw.lam[c] = vw.csF[c];
    if (set_mu) vw.mu[c] = vw.csG[c];
  }
}

// normalisation_check (R/utils.r:176-195): S[,j] *= csF[j]*csG[j]; F[,j] /= csF[j]; G[,j] /= csG[j].
// Needs csF / csG from rn_factor_sums.  grid-stride over max(n, p).
__global__ void __launch_bounds__(256) rn_normalise(const RnView vw) {
  const int K = vw.k, KP = vw.kp;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < vw.n)
    for (int c = 0; c < K; ++c) vw.F[rn_fidx(i, c, KP)] /= vw.csF[c];
  if (i < vw.p)
    for (int c = 0; c < K; ++c) vw.G[i * KP + c] /= vw.csG[c];
  if (blockIdx.x == 0 && threadIdx.x < K * K) {
    const int b = threadIdx.x / K;
    vw.S[threadIdx.x] *= vw.csF[b] * vw.csG[b];
  }
}
#endif  // RN_KERNEL_TEMPLATES_ONLY
