// Whole fits that are small enough to live in one SM's caches -- BASELINE configs[0] (the README toy: two views 100 x 50
// and 100 x 30, 64 KB of X) and the reference's own test data (2 x 180 x 180) -- as ONE persistent launch: a single CTA
// runs `sweeps` update-iterations of ALL views back to back (update_matrices, R/update_steps.r:272-319, in its literal
// Gauss-Seidel order; calculate_error, R/utils.r:157-166; the stop rule of R/main.r:55-81 on the device), so a sweep
// costs a handful of block barriers instead of one or two kernel launches per view.  At these sizes the streaming
// kernels are pure launch / grid-synchronisation latency (35 us per sweep for the toy, measured); X (<= 1 MB) is read
// from L1 / L2 through the read-only path; F, G, S stay in global memory exactly where every other entry point expects
// them, with a shared-memory copy of the current view's F and G for the reductions (what a kernel has just written to
// global memory is an L2 round trip away from its own next load -- the first version, without the copies, took 51 us per
// sweep of the toy against 35 us for two launches per view); the per-row / per-column update math is the same device
// code the streaming kernels use (rn_update_f_row, rn_update_g_row, rn_view_finish), so the numbers obey the same 1e-9
// bar.  Every sum has a fixed order: thread partials over interleaved slices, slices combined in slice order.
#pragma once
#include "rn_kernels.cuh"

#define RN_SM_THREADS 1024
#define RN_SM_MAXDIM 1024            // ldx and pp of every view
#define RN_SM_MAXELEMS (1 << 17)     // ldx * pp of every view (1 MB of X)
#define RN_SM_AUTO_ELEMS (1 << 14)   // sum of ldx * pp over the views up to which RESNMTF_IMPL_AUTO picks this path
#define RN_SM_PART_DOUBLES (RN_SM_THREADS * 8)
// dynamic shared memory: part | small | S W V F'F fin U Sn red lambda/2 mu/2 | Fs [dim][8] | Gs [dim][8], dim = the largest
// padded row / column count of the fit's views
#define RN_SM_SMEM_DOUBLES (RN_SM_PART_DOUBLES + RN_SM_THREADS + 64 + 64 + 64 + 72 + 136 + 64 + 64 + 64 + 8 + 8)
static inline size_t rn_small_smem(int dim) { return ((size_t)RN_SM_SMEM_DOUBLES + (size_t)16 * dim) * sizeof(double); }

template <int K>
__device__ __noinline__ void rn_small_view(const RnView& vw, const RnFit& ft, const int v, const int fuse, double* part,
                                           const int dim) {
  constexpr int KK = K * K, NFF = KK + K, NOUT = 2 * KK + K, NT = RN_SM_THREADS;
  const int tid = threadIdx.x;
  const int KP = vw.kp;
  const int ldx = (int)vw.ldx, pp = (int)vw.pp, n = (int)vw.n, p = (int)vw.p;
  double* small = part + RN_SM_PART_DOUBLES;  // [RN_SM_THREADS]
  double* Ssm = small + RN_SM_THREADS;
  double* Wsm = Ssm + 64;
  double* Vs = Wsm + 64;
  double* FtFs = Vs + 64;
  double* fin = FtFs + 72;
  double* Us = fin + 136;
  double* Sn = Us + 64;
  double* red = Sn + 64;
  double* lamh = red + 64;
  double* muh = lamh + 8;
  double* Fs = muh + 8;        // [dim][K] new F of this view (rows >= n zero)
  double* Gs = Fs + 8 * dim;   // [dim][K] G of this view: old in the X G phase, new afterwards
  // ---- set-up: S, W = crossprod(G) t(S), lambda / 2, mu / 2 ----
  if (tid < KK) Ssm[tid] = vw.S[tid];
  if (tid < K) {
    lamh[tid] = 0.5 * vw.lam[tid];
    muh[tid] = 0.5 * vw.mu[tid];
  }
  for (int i = tid; i < pp * K; i += NT) Gs[i] = vw.G[(int64_t)(i / K) * KP + i % K];
  __syncthreads();
  if (tid < KK) {
    const int a = tid % K, b = tid / K;
    double s = 0.0;
    for (int c = 0; c < K; ++c) s = fma(vw.GtG[a + c * K], Ssm[b + c * K], s);
    Wsm[a + b * K] = s;
  }
  // ---- P = X G: thread (row, column slice) ----
  {
    const int nsl = NT / ldx, r = tid % ldx, s = tid / ldx;
    if (s < nsl) {
      double acc[K];
#pragma unroll
      for (int j = 0; j < K; ++j) acc[j] = 0.0;
#pragma unroll 4
      for (int c = s; c < p; c += nsl) {
        const double x = __ldg(vw.X + rn_fidx(r, c, pp));
        const double* g = Gs + c * K;
#pragma unroll
        for (int j = 0; j < K; ++j) acc[j] = fma(x, g[j], acc[j]);
      }
#pragma unroll
      for (int j = 0; j < K; ++j) part[(s * ldx + r) * K + j] = acc[j];
    }
    __syncthreads();
    if (tid < n) {  // update_f for row tid (R/update_steps.r:141-165), partner rows of earlier views are already new
      double P[K];
#pragma unroll
      for (int j = 0; j < K; ++j) P[j] = 0.0;
      for (int q = 0; q < nsl; ++q)
#pragma unroll
        for (int j = 0; j < K; ++j) P[j] += part[(q * ldx + tid) * K + j];
      double fn[K];
      rn_update_f_row<K>(vw, ft, v, tid, P, Ssm, Wsm, lamh, fn);
#pragma unroll
      for (int j = 0; j < K; ++j) Fs[tid * K + j] = fn[j];
    } else if (tid < ldx) {
#pragma unroll
      for (int j = 0; j < K; ++j) Fs[tid * K + j] = 0.0;
    }
    __syncthreads();
  }
  // ---- F'F | colSums(F): thread (output, row slice of 8) ----
  {
    const int o = tid & 127, s = tid >> 7;
    double acc = 0.0;
    if (o < NFF) {
      if (o < KK) {
        const int a = o % K, b = o / K;
#pragma unroll 4
        for (int r = s; r < n; r += 8) acc = fma(Fs[r * K + a], Fs[r * K + b], acc);
      } else {
        const int c = o - KK;
#pragma unroll 4
        for (int r = s; r < n; r += 8) acc += Fs[r * K + c];
      }
    }
    small[tid] = acc;
    __syncthreads();
    if (tid < NFF) {
      double t = 0.0;
      for (int q = 0; q < 8; ++q) t += small[q * 128 + tid];
      FtFs[tid] = t;
    }
    __syncthreads();
    if (tid < KK) {  // V = crossprod(F) %*% S
      const int a = tid % K, c = tid / K;
      double t = 0.0;
      for (int b = 0; b < K; ++b) t = fma(FtFs[a + b * K], Ssm[b + c * K], t);
      Vs[a + c * K] = t;
    }
  }
  // ---- T = X'F: thread (column, row slice) ----
  {
    const int nsl = NT / pp, c = tid % pp, s = tid / pp;
    if (s < nsl) {
      double acc[K];
#pragma unroll
      for (int j = 0; j < K; ++j) acc[j] = 0.0;
      if (c < p)
#pragma unroll 4
        for (int r = s; r < n; r += nsl) {
          const double x = __ldg(vw.X + rn_fidx(r, c, pp));
#pragma unroll
          for (int j = 0; j < K; ++j) acc[j] = fma(x, Fs[r * K + j], acc[j]);
        }
#pragma unroll
      for (int j = 0; j < K; ++j) part[(s * pp + c) * K + j] = acc[j];
    }
    __syncthreads();
    if (tid < p) {  // update_g for column tid (R/update_steps.r:180-207)
      double Tj[K], gn[K];
#pragma unroll
      for (int j = 0; j < K; ++j) Tj[j] = 0.0;
      for (int q = 0; q < nsl; ++q)
#pragma unroll
        for (int j = 0; j < K; ++j) Tj[j] += part[(q * pp + tid) * K + j];
#pragma unroll
      for (int j = 0; j < K; ++j) part[tid * K + j] = Tj[j];  // slice 0 of this column now holds T[tid,]
      rn_update_g_row<K>(vw, ft, v, tid, Tj, Ssm, Vs, muh, gn);
#pragma unroll
      for (int j = 0; j < K; ++j) Gs[tid * K + j] = gn[j];
    }
    __syncthreads();
  }
  // ---- G'G | A = T'G | colSums(G): thread (output, column slice of 4) ----
  {
    const int o = tid & 255, s = tid >> 8;
    double acc = 0.0;
    if (o < NOUT) {
      if (o < KK) {
        const int a = o % K, b = o / K;
#pragma unroll 4
        for (int j = s; j < p; j += 4) acc = fma(Gs[j * K + a], Gs[j * K + b], acc);
      } else if (o < 2 * KK) {
        const int a = (o - KK) % K, b = (o - KK) / K;
#pragma unroll 4
        for (int j = s; j < p; j += 4) acc = fma(part[j * K + a], Gs[j * K + b], acc);
      } else {
        const int c = o - 2 * KK;
#pragma unroll 4
        for (int j = s; j < p; j += 4) acc += Gs[j * K + c];
      }
    }
    small[tid] = acc;
    __syncthreads();
    if (tid < NOUT) fin[tid] = ((small[tid] + small[256 + tid]) + small[512 + tid]) + small[768 + tid];
    __syncthreads();
  }
  // ---- update_s, update_lm, algebraic error, and (last view, not DIRECT) the sweep's bookkeeping ----
  rn_view_finish<K, NT>(vw, ft, v, tid, fin, FtFs, Ssm, Us, Sn, red, fuse);
  __syncthreads();
  // ---- error mode DIRECT: sum (X - (F S) G')^2 with the new factors (calculate_error, R/utils.r:157-166) ----
  if (ft.err_mode == 2) {
    const int nsl = NT / ldx, r = tid % ldx, s = tid / ldx;
    double acc = 0.0;
    if (s < nsl && r < n) {
      double fs[K];
#pragma unroll
      for (int a = 0; a < K; ++a) {
        double t = 0.0;
#pragma unroll
        for (int b = 0; b < K; ++b) t = fma(Fs[r * K + b], Sn[b + a * K], t);  // Sn: the new S (rn_view_finish)
        fs[a] = t;
      }
#pragma unroll 4
      for (int c = s; c < p; c += nsl) {
        const double* g = Gs + c * K;
        double h = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) h = fma(fs[j], g[j], h);
        const double d = __ldg(vw.X + rn_fidx(r, c, pp)) - h;
        acc = fma(d, d, acc);
      }
    }
    small[tid] = acc;
    __syncthreads();
    for (int o = NT / 2; o > 0; o >>= 1) {
      if (tid < o) small[tid] += small[tid + o];
      __syncthreads();
    }
    if (tid == 0) {
      const double e = small[0] / vw.scal[0];
      vw.scal[1] = e;
      vw.scal[3] = e;
      ft.ctrl->direct_passes += 1;
    }
    __syncthreads();
  }
}

// grid 1, RN_SM_THREADS threads, rn_small_smem(dim) bytes of dynamic shared memory
__global__ void __launch_bounds__(RN_SM_THREADS, 1) rn_small_sweeps(const RnFit ft, const int64_t sweeps, const int dim) {
  extern __shared__ __align__(16) double rn_small_part[];
  __shared__ int s_done;
  __shared__ RnView s_vw;
  const bool direct = ft.err_mode == 2;
  for (int64_t it = 0; it < sweeps; ++it) {
    if (threadIdx.x == 0) s_done = ft.ctrl->done;
    __syncthreads();
    if (s_done) break;
    for (int v = 0; v < ft.n_views; ++v) {
      // the view's descriptor once per CTA into shared memory (a per-thread copy is 40 loads + 40 local stores for each
      // of the 1024 threads: 21 % of the stall samples of the first version, lg_throttle)
      static_assert(sizeof(RnView) % 8 == 0, "copied as 64-bit words");
      if (threadIdx.x < sizeof(RnView) / 8)
        reinterpret_cast<unsigned long long*>(&s_vw)[threadIdx.x] =
            reinterpret_cast<const unsigned long long*>(ft.views + v)[threadIdx.x];
      __syncthreads();
      const RnView& vw = s_vw;
      const int fuse = (!direct && v == ft.n_views - 1) ? 1 : 0;
      switch (vw.k) {
#define X(KC) case KC: rn_small_view<KC>(vw, ft, v, fuse, rn_small_part, dim); break;
        RN_K_CASES_LE8(X)
#undef X
      }
      __syncthreads();
    }
    if (direct && threadIdx.x == 0) rn_finish_dev(ft);
    __syncthreads();
  }
}
