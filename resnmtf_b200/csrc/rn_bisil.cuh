// Distance blocks of the bisilhouette score (bisilhouette::bisilhouette as called by
// obtain_biclusters(), R/obtain_bicl.r:190-199 -- package source not in the reference tree, restated from its published
// definition in resnmtf_b200/bicluster.py: parity unpinned).  For bicluster j the rows R_j are compared, on the
// columns C_j only, with themselves and with every other row cluster:
//   rn_bisil_gather   Y = X[row list, C_j] as a compact row-major matrix (rows: R_j first, then the other groups, every
//                     segment padded to 64 rows; columns padded to 16) -- the distance kernel then streams
//                     contiguous rows instead of gathering 8-byte entries out of the panel layout
//   rn_bisil_dist     part[split][segment][i] = sum over the rows i' of the segment (of this split's share) of
//                     d(Y_i, Y_i'), i in R_j; 64 x 64 pair tiles, 4 x 4 pairs per thread, direct difference form
//                     (sum (a - b)^2, sum |a - b|) or dot products + norms (cosine); d(i, i) = 0 exactly
// All sums have a fixed order (per thread over column chunks, per row over the 16 threads of a tile in order, per
// segment over the tiles of a split in order, splits combined in order on the host).
#pragma once
#include "rn_data.cuh"

#define RN_DIST_EUCLIDEAN 0
#define RN_DIST_MANHATTAN 1
#define RN_DIST_COSINE 2

// Y[slot][c] = X[rows[slot]][cols[c]] (rows[slot] < 0: zero row; c >= m: zero), Y row-major with leading dimension ldy
__global__ void __launch_bounds__(256) rn_bisil_gather(const double* __restrict__ X, int64_t pp, const int32_t* __restrict__ rows,
                                                       int64_t n_slots, const int32_t* __restrict__ cols, int m, int ldy,
                                                       double* __restrict__ Y, double* __restrict__ norms) {
  // one warp per slot: lanes stride over the columns; the row norm (cosine) is reduced in lane order
  const int lane = threadIdx.x & 31;
  const int64_t slot = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (slot >= n_slots) return;
  const int r = rows[slot];
  double ss = 0.0;
  for (int c = lane; c < ldy; c += 32) {
    double v = 0.0;
    if (r >= 0 && c < m) v = X[rn_xidx(r, cols[c], pp)];
    Y[slot * ldy + c] = v;
    ss = fma(v, v, ss);
  }
  if (norms) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) norms[slot] = sqrt(ss);
  }
}

struct RnBisil {
  const double* Y;        // [slots][ldy]
  const double* norms;    // [slots] (cosine)
  const int32_t* rows;    // [slots] original row index (-1: padding), to recognise i' == i
  int ldy;
  int a_tiles;            // 64-row tiles of segment 0 (= R_j)
  int b_tiles;            // 64-row tiles of all segments (segment 0 included)
  const int32_t* seg_of_tile;  // [b_tiles]
  int n_seg;
  int splits;
  double* part;           // [splits][n_seg][a_tiles * 64]
};

template <int METHOD>
__global__ void __launch_bounds__(256) rn_bisil_dist(const RnBisil a) {
  __shared__ double As[16][68];
  __shared__ double Bs[16][68];
  __shared__ double red[16][64];
  __shared__ double segsum[17][64];
  // thread: A rows 4 ty .. 4 ty + 3 (one 32-byte run: a broadcast read per half warp), B rows tx, tx + 16, tx + 32,
  // tx + 48 (consecutive lanes read consecutive words: no bank conflicts) of the tile pair
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int at = blockIdx.x, split = blockIdx.y;
  for (int i = threadIdx.x; i < 17 * 64; i += 256) segsum[i / 64][i % 64] = 0.0;
  const RnSplit sp(a.b_tiles, a.splits);
  const int bt0 = (int)sp.begin(split), bt1 = (int)sp.begin(split + 1);
  const double* Ya = a.Y + (int64_t)at * 64 * a.ldy;
  int rowa[4];
  double na[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    rowa[u] = a.rows[at * 64 + 4 * ty + u];
    na[u] = a.norms ? a.norms[at * 64 + 4 * ty + u] : 0.0;
  }
  __syncthreads();
  for (int bt = bt0; bt < bt1; ++bt) {
    const double* Yb = a.Y + (int64_t)bt * 64 * a.ldy;
    double acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int w = 0; w < 4; ++w) acc[u][w] = 0.0;
    for (int c0 = 0; c0 < a.ldy; c0 += 16) {
      __syncthreads();
      for (int i = threadIdx.x; i < 64 * 16; i += 256) {  // 64 rows x 16 columns of each operand, transposed into smem
        const int r = i >> 4, c = i & 15;
        As[c][r] = Ya[(int64_t)r * a.ldy + c0 + c];
        Bs[c][r] = Yb[(int64_t)r * a.ldy + c0 + c];
      }
      __syncthreads();
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        double av[4], bv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) av[u] = As[c][4 * ty + u];
#pragma unroll
        for (int w = 0; w < 4; ++w) bv[w] = Bs[c][tx + 16 * w];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            if (METHOD == RN_DIST_EUCLIDEAN) {
              const double df = av[u] - bv[w];
              acc[u][w] = fma(df, df, acc[u][w]);
            } else if (METHOD == RN_DIST_MANHATTAN) {
              acc[u][w] += fabs(av[u] - bv[w]);
            } else {
              acc[u][w] = fma(av[u], bv[w], acc[u][w]);
            }
          }
      }
    }
    // distances of this tile pair, summed over the thread's 4 B rows (padding rows and i' == i contribute 0)
    double rs[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int slot_b = bt * 64 + tx + 16 * w;
      const int rowb = a.rows[slot_b];
      const double nb = a.norms ? a.norms[slot_b] : 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        double d;
        if (METHOD == RN_DIST_EUCLIDEAN) d = sqrt(acc[u][w]);
        else if (METHOD == RN_DIST_MANHATTAN) d = acc[u][w];
        else {
          const double den = na[u] * nb;
          double c = (den != 0.0) ? acc[u][w] / den : 0.0;  // nan_to_num of 0 / 0
          if (isnan(c) || isinf(c)) c = 0.0;
          d = 1.0 - c;
        }
        if (rowb < 0 || rowa[u] < 0 || rowb == rowa[u]) d = 0.0;
        rs[u] += d;
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u) red[tx][4 * ty + u] = rs[u];
    __syncthreads();
    if (threadIdx.x < 64) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < 16; ++q) s += red[q][threadIdx.x];
      segsum[a.seg_of_tile[bt]][threadIdx.x] += s;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.n_seg * 64; i += 256) {
    const int sg = i / 64, r = i % 64;
    a.part[((int64_t)split * a.n_seg + sg) * (a.a_tiles * 64) + at * 64 + r] = segsum[sg][r];
  }
}
