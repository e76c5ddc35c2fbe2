// Hand-written sm_100a kernels of the ResNMTF update sweep (FP64 throughout).
//
// One update-iteration of view v (reference: update_matrices, R/update_steps.r:272-319) is
//   rn_f_step      P = X.G streamed over X once, fused with the F update (update_f, :141-165) incl. the
//                  phi coupling gather (star_prod_relevant, R/utils.r:63-78)
//   rn_g_stream_*  T = X'.F streamed over X once (+ F'F and colSums(F) on the tensor-core path)
//   rn_g_epilogue  G update (update_g, :180-207) incl. psi coupling, G'G, A = T'G, then in the last CTA
//                  the S update (update_s, :220-240; its numerator crossprod(F,X) G equals A, so X is not
//                  read a third time), lambda/mu (update_lm, :249-251) and the algebraic error
//   rn_residual    optional direct error pass (calculate_error, R/utils.r:157-166)
//   rn_finish      mean error, history, stop rule (R/main.r:74-80)
// All cross-CTA reductions are two-stage with a fixed summation order (no floating-point atomics), so
// results are bit-reproducible run to run.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "rn_types.h"

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------

// 128-bit streaming load of two consecutive doubles of X: read-only path, do not allocate in L1 (X has
// no reuse inside a pass; L1 is kept for the G / F rows every warp re-reads).
__device__ __forceinline__ double2 rn_ld_stream2(const double* p) {
  double2 v;
  asm("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// FP64 tensor-core MMA, D(8x8) += A(8x4, row) * B(4x8, col).  Lane l = 4*g + t holds
// a = A[g][t], b = B[t][g], c0/c1 = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void rn_dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__device__ __forceinline__ double rn_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// F step:  P = X.G  (streaming)  +  update_f epilogue
//   grid (row_tiles, cs), 256 threads.  A CTA owns 64 rows and the data columns of split `cs`; its 8
//   warps interleave over the columns.  With cs > 1 the last CTA to arrive for a row tile sums the
//   partials in split order and runs the epilogue.
// ------------------------------------------------------------------------------------------------
template <int K, int KP, bool MMA>
__global__ void __launch_bounds__(256) rn_f_step(const RnView vw, const RnFit ft, const int v) {
  static_assert(!MMA || KP == 8, "tensor-core path holds k in one 8-wide fragment");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, cs = blockIdx.y, ncs = gridDim.y;
  if (ft.ctrl->done) return;  // state is frozen once the stop rule fired (uniform over the grid)

  __shared__ double Ps[RN_ROW_TILE * KP];
  __shared__ double Ssm[K * K], Wsm[K * K], lamh[K];
  __shared__ int s_last;

  const int64_t ldx = vw.ldx;
  const int64_t r0 = (int64_t)tile * RN_ROW_TILE;
  const int nb = (int)(vw.pp >> 3);
  const int64_t jbeg = 8 * ((int64_t)nb * cs / ncs);
  const int64_t jend = 8 * ((int64_t)nb * (cs + 1) / ncs);
  const double* __restrict__ X = vw.X;
  const double* __restrict__ G = vw.G;

  for (int i = tid; i < RN_ROW_TILE * KP; i += 256) Ps[i] = 0.0;

  if constexpr (MMA) {
    // lane (g,t): rows r0 + 16m + 2g + {0,1} (m = 0..3), data column 4q + t.
    const int g = lane >> 2, t = lane & 3;
    double acc[4][2][2];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int h = 0; h < 2; ++h) acc[m][h][0] = acc[m][h][1] = 0.0;
    const double* xb = X + r0 + 2 * g;
    int64_t q = (jbeg >> 2) + warp;
    const int64_t qend = jend >> 2;
    for (; q + 8 < qend; q += 16) {
      const int64_t ja = q * 4 + t, jb = ja + 32;
      const double* pa = xb + ja * ldx;
      const double* pb = xb + jb * ldx;
      double2 xa[4], xbv[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) xa[m] = rn_ld_stream2(pa + 16 * m);
#pragma unroll
      for (int m = 0; m < 4; ++m) xbv[m] = rn_ld_stream2(pb + 16 * m);
      const double ba = G[ja * KP + g], bb = G[jb * KP + g];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        rn_dmma(acc[m][0][0], acc[m][0][1], xa[m].x, ba);
        rn_dmma(acc[m][1][0], acc[m][1][1], xa[m].y, ba);
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        rn_dmma(acc[m][0][0], acc[m][0][1], xbv[m].x, bb);
        rn_dmma(acc[m][1][0], acc[m][1][1], xbv[m].y, bb);
      }
    }
    for (; q < qend; q += 8) {
      const int64_t ja = q * 4 + t;
      const double* pa = xb + ja * ldx;
      double2 xa[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) xa[m] = rn_ld_stream2(pa + 16 * m);
      const double ba = G[ja * KP + g];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        rn_dmma(acc[m][0][0], acc[m][0][1], xa[m].x, ba);
        rn_dmma(acc[m][1][0], acc[m][1][1], xa[m].y, ba);
      }
    }
    __syncthreads();
    for (int w = 0; w < 8; ++w) {
      if (warp == w) {
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int row = 16 * m + 2 * g + h;
            Ps[row * KP + 2 * t] += acc[m][h][0];
            Ps[row * KP + 2 * t + 1] += acc[m][h][1];
          }
      }
      __syncthreads();
    }
  } else {
    // lane: rows r0 + 2*lane + {0,1}; warp w takes data columns jbeg + w, + 8, ...
    double a0[K], a1[K];
#pragma unroll
    for (int c = 0; c < K; ++c) a0[c] = a1[c] = 0.0;
    const double* xr = X + r0 + 2 * lane;
    constexpr int U = 4;
    int64_t j = jbeg + warp;
    for (; j + 8 * (U - 1) < jend; j += 8 * U) {
      double2 x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) x[u] = rn_ld_stream2(xr + (j + 8 * u) * ldx);
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
      const double2 x = rn_ld_stream2(xr + j * ldx);
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
  }

  if (ncs > 1) {
    double* mine = vw.Ppart + ((int64_t)cs * ldx + r0) * KP;
    for (int i = tid; i < RN_ROW_TILE * KP; i += 256) mine[i] = Ps[i];
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&vw.tile_ticket[tile], 1) == ncs - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int i = tid; i < RN_ROW_TILE * KP; i += 256) {
      double s = 0.0;
      for (int c2 = 0; c2 < ncs; ++c2) s += __ldcg(vw.Ppart + ((int64_t)c2 * ldx + r0) * KP + i);
      Ps[i] = s;
    }
    if (tid == 0) vw.tile_ticket[tile] = 0;
  }

  // ---- epilogue: update_f (R/update_steps.r:141-165) -------------------------------------------
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

  const int V = ft.n_views;
  if (tid < RN_ROW_TILE) {
    const int64_t r = r0 + tid;
    if (r < vw.n) {
      double P[K], f[K], N[K], FS[K], D[K];
#pragma unroll
      for (int c = 0; c < K; ++c) {
        P[c] = Ps[tid * KP + c];
        f[c] = vw.F[(int64_t)c * ldx + r];
      }
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
          vw.F[(int64_t)c * ldx + r] = fabs(f[c] * ratio);
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
          const double nw = (double)ow->n;
          int64_t src = -1;
          if (mode == RN_MODE_MAP) src = ft.rowmap[w + v * V][r];
          const double* fw = ow->F;
          const int64_t ldw = ow->ldx;
#pragma unroll
          for (int c = 0; c < K; ++c) {
            const double m = (src >= 0) ? fw[(int64_t)c * ldw + src] : f[c];
            pc[c] += (ph * m) * nw;
          }
        }
        const double nv = (double)vw.n;
#pragma unroll
        for (int c = 0; c < K; ++c) {
          const double num = N[c] + pc[c] / nv;
          const double den = (D[c] + phisum * f[c]) + lamh[c];
          vw.F[(int64_t)c * ldx + r] = fabs(f[c] * (num / den));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// G stream, tensor-core path:  T = X'.F  (+ F'F and colSums(F) from the CTAs of column group 0)
//   grid (col_groups, rs), 128 threads.  A CTA owns 64 data columns and the 64-row steps of split rs;
//   its 4 warps interleave over the steps.  Lane (g,t) of a warp feeds, for every 8-column block jb,
//   column 8*jb+g and rows 8i + 2t + {0,1} (i = 0..7) of the step, so that each LDG.128 of a warp covers
//   whole 32 B sectors and each column is read in 512 B contiguous pieces.
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128) rn_g_stream_mma(const RnView vw, const RnFit ft) {
  constexpr int KP = 8;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = blockIdx.x, rs = blockIdx.y, nrs = gridDim.y;
  const int g = lane >> 2, t = lane & 3;
  if (ft.ctrl->done) return;

  __shared__ double Ts[RN_COL_GROUP * KP];
  __shared__ double FFs[K * K + K];
  __shared__ int s_last;

  const int64_t ldx = vw.ldx;
  const int ns = (int)(ldx / RN_ROW_TILE);
  const int s0 = (int)((int64_t)ns * rs / nrs), s1 = (int)((int64_t)ns * (rs + 1) / nrs);
  const int64_t j0 = (int64_t)grp * RN_COL_GROUP;
  const int njb = (int)min((int64_t)8, (vw.pp - j0) >> 3);
  const bool doFF = (grp == 0);
  const double* __restrict__ X = vw.X;
  const double* __restrict__ F = vw.F;

  double acc[8][2];
#pragma unroll
  for (int jb = 0; jb < 8; ++jb) acc[jb][0] = acc[jb][1] = 0.0;
  double aff0 = 0.0, aff1 = 0.0, acs0 = 0.0, acs1 = 0.0;

  for (int i = tid; i < RN_COL_GROUP * KP; i += 128) Ts[i] = 0.0;
  if (tid < K * K + K) FFs[tid] = 0.0;

  for (int s = s0 + warp; s < s1; s += 4) {
    const int64_t rr = (int64_t)s * RN_ROW_TILE + 2 * t;
    const double* fp = F + (int64_t)g * ldx + rr;
    double2 f2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f2[i] = *reinterpret_cast<const double2*>(fp + 8 * i);
    if (doFF) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        rn_dmma(aff0, aff1, f2[i].x, f2[i].x);
        rn_dmma(aff0, aff1, f2[i].y, f2[i].y);
        rn_dmma(acs0, acs1, 1.0, f2[i].x);
        rn_dmma(acs0, acs1, 1.0, f2[i].y);
      }
    }
    const double* xp = X + (j0 + g) * ldx + rr;
    double2 xa[8], xb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xa[i] = rn_ld_stream2(xp + 8 * i);
#pragma unroll
    for (int jb = 0; jb < 8; jb += 2) {
      if (jb + 1 < njb) {
        const double* xq = xp + (int64_t)(8 * (jb + 1)) * ldx;
#pragma unroll
        for (int i = 0; i < 8; ++i) xb[i] = rn_ld_stream2(xq + 8 * i);
      }
      if (jb < njb) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          rn_dmma(acc[jb][0], acc[jb][1], xa[i].x, f2[i].x);
          rn_dmma(acc[jb][0], acc[jb][1], xa[i].y, f2[i].y);
        }
      }
      if (jb + 2 < njb) {
        const double* xq = xp + (int64_t)(8 * (jb + 2)) * ldx;
#pragma unroll
        for (int i = 0; i < 8; ++i) xa[i] = rn_ld_stream2(xq + 8 * i);
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

  __syncthreads();
  for (int w = 0; w < 4; ++w) {
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
          if (2 * t < K) FFs[K * K + 2 * t] += acs0;
          if (2 * t + 1 < K) FFs[K * K + 2 * t + 1] += acs1;
        }
      }
    }
    __syncthreads();
  }

  if (doFF && tid < K * K + K) vw.FFpart[(int64_t)rs * (K * K + K) + tid] = FFs[tid];

  if (nrs == 1) {
    double* out = vw.T + j0 * KP;
    for (int i = tid; i < 8 * njb * KP; i += 128) out[i] = Ts[i];
    return;
  }
  double* mine = vw.Tpart + ((int64_t)rs * vw.pp + j0) * KP;
  for (int i = tid; i < 8 * njb * KP; i += 128) mine[i] = Ts[i];
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&vw.group_ticket[grp], 1) == nrs - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = tid; i < 8 * njb * KP; i += 128) {
    double s = 0.0;
    for (int r2 = 0; r2 < nrs; ++r2) s += __ldcg(vw.Tpart + ((int64_t)r2 * vw.pp + j0) * KP + i);
    vw.T[j0 * KP + i] = s;
  }
  if (tid == 0) vw.group_ticket[grp] = 0;
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

  const int64_t ldx = vw.ldx;
  const int ns = (int)(ldx / RN_ROW_TILE);
  const int s0 = (int)((int64_t)ns * rs / nrs), s1 = (int)((int64_t)ns * (rs + 1) / nrs);
  const int64_t j0 = (int64_t)grp * RN_COL_GROUP_DFMA;
  const int64_t jc = j0 + 4 * warp;
  const int ncols = (int)min((int64_t)RN_COL_GROUP_DFMA, vw.pp - j0);
  const bool active = jc < vw.pp;
  const double* __restrict__ X = vw.X;
  const double* __restrict__ F = vw.F;

  double acc[4][K];
#pragma unroll
  for (int jj = 0; jj < 4; ++jj)
#pragma unroll
    for (int c = 0; c < K; ++c) acc[jj][c] = 0.0;

  if (active) {
    for (int s = s0; s < s1; ++s) {
      const int64_t rr = (int64_t)s * RN_ROW_TILE + 2 * lane;
      double2 x[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) x[jj] = rn_ld_stream2(X + (jc + jj) * ldx + rr);
      double2 f2[K];
#pragma unroll
      for (int c = 0; c < K; ++c) f2[c] = *reinterpret_cast<const double2*>(F + (int64_t)c * ldx + rr);
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
  double* mine = vw.Tpart + ((int64_t)rs * vw.pp + j0) * KP;
  for (int i = tid; i < ncols * KP; i += 256) mine[i] = Ts[i];
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&vw.group_ticket[grp], 1) == nrs - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = tid; i < ncols * KP; i += 256) {
    double s = 0.0;
    for (int r2 = 0; r2 < nrs; ++r2) s += __ldcg(vw.Tpart + ((int64_t)r2 * vw.pp + j0) * KP + i);
    vw.T[j0 * KP + i] = s;
  }
  if (tid == 0) vw.group_ticket[grp] = 0;
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
    for (int c = 0; c < K; ++c) rows[tid * K + c] = (r < vw.n) ? vw.F[(int64_t)c * ldx + r] : 0.0;
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
// G epilogue: update_g (R/update_steps.r:180-207), partial G'G | A | colSums(G); the last CTA then
// finishes the view: update_s (:220-240), update_lm (:249-251), algebraic error.
//   grid (gepi_ctas), 256 threads (128 for k > 8), one data column (row of G) per thread.
// ------------------------------------------------------------------------------------------------
template <int K, int KP>
__global__ void __launch_bounds__(RN_GEPI_THREADS(K)) rn_g_epilogue(const RnView vw, const RnFit ft, const int v) {
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

  const int V = ft.n_views;
  const int64_t j = (int64_t)blockIdx.x * NT + tid;
  {
    double Tj[K], gj[K], gn[K];
    if (j < vw.p) {
      double N[K], GS[K], D[K];
#pragma unroll
      for (int c = 0; c < K; ++c) {
        Tj[c] = vw.T[j * KP + c];
        gj[c] = vw.G[j * KP + c];
      }
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
          double ratio = N[c] / (D[c] + muh[c]);
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
          const double num = N[c] + pc[c] / pv;
          const double den = (D[c] + psisum * gj[c]) + muh[c];
          gn[c] = fabs(gj[c] * (num / den));
        }
      }
#pragma unroll
      for (int c = 0; c < K; ++c) vw.G[j * KP + c] = gn[c];
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
  // partial G'G | A = T'G | colSums(G) of this CTA's 256 data columns
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

  // ---- last CTA: finish the view ------------------------------------------------------------------
  for (int o = tid; o < NOUT; o += NT) {
    double s = 0.0;
    for (int i = 0; i < (int)gridDim.x; ++i) s += __ldcg(vw.GGpart + (int64_t)i * NOUT + o);
    fin[o] = s;
    if (o < KK) vw.GtG[o] = s;
    else if (o < 2 * KK) vw.A[o - KK] = s;
    else vw.csG[o - 2 * KK] = s;
  }
  for (int o = tid; o < KK; o += NT) vw.FtF[o] = FtFs[o];
  if (tid < K) vw.csF[tid] = FtFs[KK + tid];
  __syncthreads();
  const double* GtGn = fin;
  const double* An = fin + KK;
  const double* csGn = fin + 2 * KK;
  for (int o = tid; o < KK; o += NT) {  // U = crossprod(F) %*% S
    const int a = o % K, c = o / K;
    double s = 0.0;
    for (int b = 0; b < K; ++b) s = fma(FtFs[a + b * K], Ssm[b + c * K], s);
    Us[a + c * K] = s;
  }
  __syncthreads();
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
  __syncthreads();
  for (int o = tid; o < KK; o += NT) vw.S[o] = Sn[o];
  if (tid < K) {  // update_lm
    vw.lam[tid] = FtFs[KK + tid] * vw.lam[tid];
    vw.mu[tid] = csGn[tid] * vw.mu[tid];
  }
  // algebraic error: (||X||^2 - 2 <A,S'> + <(F'F S') G'G, S'>) / ||X||^2
  for (int o = tid; o < KK; o += NT) {
    const int a = o % K, c = o / K;
    double s = 0.0;
    for (int b = 0; b < K; ++b) s = fma(FtFs[a + b * K], Sn[b + c * K], s);
    Us[a + c * K] = s;
  }
  __syncthreads();
  for (int o = tid; o < KK; o += NT) {
    const int a = o % K, b = o / K;
    double q = 0.0;
    for (int c = 0; c < K; ++c) q = fma(Us[a + c * K], GtGn[c + b * K], q);
    red[o] = (q - 2.0 * An[o]) * Sn[o];
  }
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int i = 0; i < KK; ++i) s += red[i];
    const double xn = vw.scal[0];
    const double e = (xn + s) / xn;
    vw.scal[1] = e;
    vw.scal[2] = e;
    int need = 0;
    if (ft.err_mode == 2) need = 1;
    else if (ft.err_mode == 0 && e < 1.0e-3) need = 1;
    vw.flags[0] = need;
  }
}

// ------------------------------------------------------------------------------------------------
// Direct residual: sum (X - (F S) G')^2, one extra pass over X.  grid (row_tiles, resid_cs), 256 thr.
// Runs only when flags[0] is set (error mode DIRECT, or AUTO with a small algebraic error).
// ------------------------------------------------------------------------------------------------
template <int K, int KP>
__global__ void __launch_bounds__(256) rn_residual(const RnView vw, const RnFit ft) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (ft.ctrl->done || !vw.flags[0]) return;
  __shared__ double Ssm[K * K];
  __shared__ double wsum[8];
  __shared__ int s_last;
  const int tile = blockIdx.x, cs = blockIdx.y, ncs = gridDim.y;
  const int64_t ldx = vw.ldx;
  const int64_t r0 = (int64_t)tile * RN_ROW_TILE + 2 * lane;
  const int nb = (int)(vw.pp >> 3);
  const int64_t jbeg = 8 * ((int64_t)nb * cs / ncs), jend = 8 * ((int64_t)nb * (cs + 1) / ncs);
  if (tid < K * K) Ssm[tid] = vw.S[tid];
  __syncthreads();
  double fs0[K], fs1[K];
  {
    double f0[K], f1[K];
#pragma unroll
    for (int c = 0; c < K; ++c) {
      const double2 f = *reinterpret_cast<const double2*>(vw.F + (int64_t)c * ldx + r0);
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
  const double* xr = vw.X + r0;
  for (int64_t j = jbeg + warp; j < jend; j += 8) {
    const double2 x = rn_ld_stream2(xr + j * ldx);
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
  // fixed-order block sum
  __shared__ double bs[256];
  bs[tid] = s;
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

// Mean error over the views, history, stop rule (R/main.r:74-80).  <<<1,1>>>.
__global__ void rn_finish(const RnFit ft) {
  RnCtrl* c = ft.ctrl;
  if (c->done) return;
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

// ------------------------------------------------------------------------------------------------
// set-up / tear-down kernels (once per fit, not on the per-iteration path)
// ------------------------------------------------------------------------------------------------

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

// Column sums of F and G plus G'G of the current factors (single CTA, 1024 threads): used after
// set_factors (lambda/mu defaults, the G'G the first F step needs) and by the final normalisation.
__global__ void __launch_bounds__(1024) rn_factor_sums(const RnView vw) {
  const int tid = threadIdx.x;
  const int K = vw.k, KP = vw.kp;
  __shared__ double bs[1024];
  for (int c = 0; c < K; ++c) {
    double s = 0.0;
    for (int64_t r = tid; r < vw.n; r += 1024) s += vw.F[(int64_t)c * vw.ldx + r];
    bs[tid] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
      if (tid < o) bs[tid] += bs[tid + o];
      __syncthreads();
    }
    if (tid == 0) vw.csF[c] = bs[0];
    __syncthreads();
  }
  for (int c = 0; c < K; ++c) {
    double s = 0.0;
    for (int64_t j = tid; j < vw.p; j += 1024) s += vw.G[j * KP + c];
    bs[tid] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
      if (tid < o) bs[tid] += bs[tid + o];
      __syncthreads();
    }
    if (tid == 0) vw.csG[c] = bs[0];
    __syncthreads();
  }
  for (int o2 = 0; o2 < K * K; ++o2) {
    const int a = o2 % K, b = o2 / K;
    double s = 0.0;
    for (int64_t j = tid; j < vw.p; j += 1024) s = fma(vw.G[j * KP + a], vw.G[j * KP + b], s);
    bs[tid] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
      if (tid < o) bs[tid] += bs[tid + o];
      __syncthreads();
    }
    if (tid == 0) vw.GtG[o2] = bs[0];
    __syncthreads();
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
    for (int c = 0; c < K; ++c) vw.F[(int64_t)c * vw.ldx + i] /= vw.csF[c];
  if (i < vw.p)
    for (int c = 0; c < K; ++c) vw.G[i * KP + c] /= vw.csG[c];
  if (blockIdx.x == 0 && threadIdx.x < K * K) {
    const int b = threadIdx.x / K;
    vw.S[threadIdx.x] *= vw.csF[b] * vw.csG[b];
  }
}
