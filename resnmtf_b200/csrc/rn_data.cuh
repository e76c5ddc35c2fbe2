// Kernels on whole views in the library's panel layout (rn_types.h): the work apply_resnmtf() does on the DATA around
// the update loop -- SURVEY 8f row N2 -- so that a shuffled refit or a stability resample never sends a matrix through
// the host:
//   make_non_neg_inner + matrix_normalisation (R/utils.r:20-27, 86-88)      rn_col_stats / rn_col_combine / rn_shift_scale
//   shuffle_view (R/obtain_bicl.r:11-22): x_messed = matrix(sample(x), ...)    rn_shuffle (keyed bijection of the entries)
//   sub-samples of stability_repeat (R/stability_analysis.r:215-253)          rn_gather, row / column sums of a sub-sample
//   the transposed view for the Gram matrix of the smaller side (rn_linalg.cuh), U = X V / d of the SVD initialisation
// All reductions have a fixed order (per-tile partials combined in tile order): results are bit-reproducible.
#pragma once
#include "rn_dev.cuh"

// position of X[r][j] in the panel layout with pp (padded) columns
__device__ __forceinline__ int64_t rn_xidx(int64_t r, int64_t j, int64_t pp) {
  return ((r >> 6) * pp + j) * RN_ROW_TILE + 2 * ((int)((r & 63) >> 1) ^ rn_sigma(j)) + (r & 1);
}

#define RN_STAT_SUM 0
#define RN_STAT_MIN 1
#define RN_STAT_SUMSQ 2

// part[tile][j] = sum / min / sum of squares of column j over the 64 rows of row tile `tile` (padding rows are zero:
// neutral for the sums, and the only minimum ever asked for is min(0, min(col)), R/utils.r:22).  grid (ceil(pp/128), tiles)
__global__ void __launch_bounds__(128) rn_col_stats(const double* __restrict__ X, int64_t pp, double* __restrict__ part,
                                                    int op) {
  const int64_t j = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (j >= pp) return;
  const int64_t tile = blockIdx.y;
  const double2* src = reinterpret_cast<const double2*>(X + (tile * pp + j) * RN_ROW_TILE);
  const int sg = rn_sigma(j);
  double acc = 0.0;
  // rows in order: piece position of row pair rp is rp ^ sigma(j)
#pragma unroll 8
  for (int rp = 0; rp < 32; ++rp) {
    const double2 v = src[rp ^ sg];
    if (op == RN_STAT_SUM) acc = (acc + v.x) + v.y;
    else if (op == RN_STAT_MIN) acc = fmin(acc, fmin(v.x, v.y));
    else acc = fma(v.y, v.y, fma(v.x, v.x, acc));
  }
  part[tile * pp + j] = acc;
}

// out[j] = the per-tile partials of column j combined in tile order
__global__ void __launch_bounds__(128) rn_col_combine(const double* __restrict__ part, int64_t pp, int tiles,
                                                      double* __restrict__ out, int op) {
  const int64_t j = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (j >= pp) return;
  double acc = 0.0;
  for (int t = 0; t < tiles; ++t) {
    const double v = part[(int64_t)t * pp + j];
    acc = (op == RN_STAT_MIN) ? fmin(acc, v) : acc + v;
  }
  out[j] = acc;
}

// out[r] = sum over the p data columns of row r.  One CTA per row tile, 256 threads = 64 rows x 4 column quarters;
// quarter q adds the columns j = q (mod 4) in order, the four quarter sums are added in order.
__global__ void __launch_bounds__(256) rn_row_sums(const double* __restrict__ X, int64_t p, int64_t pp,
                                                   double* __restrict__ out) {
  __shared__ double sm[4][64];
  const int r = threadIdx.x & 63, q = threadIdx.x >> 6;
  const int64_t tile = blockIdx.x;
  const double* base = X + tile * pp * RN_ROW_TILE;
  double acc = 0.0;
  for (int64_t j = q; j < p; j += 4) acc += base[j * RN_ROW_TILE + 2 * ((r >> 1) ^ rn_sigma(j)) + (r & 1)];
  sm[q][r] = acc;
  __syncthreads();
  if (q == 0) out[tile * 64 + r] = ((sm[0][r] + sm[1][r]) + sm[2][r]) + sm[3][r];
}

// x <- (x + shift[j]) / den[j] on the n x p data entries (shift / den may be NULL); padding stays zero
__global__ void __launch_bounds__(256) rn_shift_scale(double* __restrict__ X, int64_t n, int64_t p, int64_t pp,
                                                      const double* __restrict__ colmin, const double* __restrict__ den) {
  const int64_t tiles = (n + 63) >> 6;
  const int64_t total = tiles * p * 64;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int pos = (int)(i & 63);
    const int64_t j = (i >> 6) % p, tile = (i >> 6) / p;
    const int rr = 2 * ((pos >> 1) ^ rn_sigma(j)) + (pos & 1);  // row inside the tile stored at this position
    if (tile * 64 + rr >= n) continue;
    double* q = X + (tile * pp + j) * RN_ROW_TILE + pos;
    double v = *q;
    if (colmin) v += fabs(fmin(0.0, colmin[j]));  // make_non_neg_inner: abs(min(0, min(col)))
    if (den) v /= den[j];
    *q = v;
  }
}

// ---- keyed bijection of [0, N): a balanced Feistel network on 2h bits with cycle walking -----------------------
__device__ __forceinline__ uint32_t rn_mix32(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return (uint32_t)x;
}
__device__ __forceinline__ uint64_t rn_permute(uint64_t i, uint64_t N, int h, uint64_t key) {
  const uint64_t mask = ((uint64_t)1 << h) - 1;
  do {
    uint64_t l = i >> h, r = i & mask;
#pragma unroll
    for (int round = 0; round < 6; ++round) {
      const uint64_t f = rn_mix32(r ^ (key + 0x9e3779b97f4a7c15ULL * (uint64_t)(round + 1))) & mask;
      const uint64_t nl = r;
      r = l ^ f;
      l = nl;
    }
    i = (l << h) | r;
  } while (i >= N);
  return i;
}
// dst = matrix(x[perm], n, p): entry i of the column-major vector of dst is entry perm(i) of src's
// (shuffle_view, R/obtain_bicl.r:13: `sample(x_i)` permutes ALL entries)
__global__ void __launch_bounds__(256) rn_shuffle(const double* __restrict__ src, double* __restrict__ dst, int64_t n,
                                                  int64_t p, int64_t pp, int h, uint64_t key) {
  const uint64_t N = (uint64_t)n * (uint64_t)p;
  for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < N; i += (uint64_t)gridDim.x * 256) {
    const uint64_t s = rn_permute(i, N, h, key);
    dst[rn_xidx((int64_t)(i % n), (int64_t)(i / n), pp)] = src[rn_xidx((int64_t)(s % n), (int64_t)(s / n), pp)];
  }
}

// dst (n_out x p_out) = src[rows, cols]
__global__ void __launch_bounds__(256) rn_gather(const double* __restrict__ src, int64_t pp_src, double* __restrict__ dst,
                                                 int64_t n_out, int64_t p_out, int64_t pp_dst,
                                                 const int32_t* __restrict__ rows, const int32_t* __restrict__ cols) {
  const int64_t total = n_out * p_out;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i % n_out, j = i / n_out;
    dst[rn_xidx(r, j, pp_dst)] = src[rn_xidx(rows[r], cols[j], pp_src)];
  }
}

// panel layout -> plain column-major (leading dimension ld)
__global__ void __launch_bounds__(256) rn_panels_to_colmajor(const double* __restrict__ X, int64_t n, int64_t p,
                                                             int64_t pp, double* __restrict__ out, int64_t ld) {
  const int64_t total = n * p;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i % n, j = i / n;
    out[r + j * ld] = X[rn_xidx(r, j, pp)];
  }
}

// dst (p x n, panel layout with ppd padded columns) = t(src) (n x p); 32 x 32 element blocks through shared memory
__global__ void __launch_bounds__(256) rn_transpose_panels(const double* __restrict__ src, int64_t n, int64_t p,
                                                           int64_t pps, double* __restrict__ dst, int64_t ppd) {
  __shared__ double tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int jj = ty; jj < 32; jj += 8) {
    const int64_t r = r0 + tx, j = j0 + jj;
    tile[jj][tx] = (r < n && j < p) ? src[rn_xidx(r, j, pps)] : 0.0;
  }
  __syncthreads();
  for (int rr = ty; rr < 32; rr += 8) {
    const int64_t j = j0 + tx, r = r0 + rr;  // dst row = j, dst column = r
    if (j < p && r < n) dst[rn_xidx(j, r, ppd)] = tile[tx][rr];
  }
}

// U[r][c] = |sum_j X[r][j] V[j][c]| / d[c]  (c < kc <= 16): the left factors of the SVD initialisation from the right
// ones, U = X V diag(1/d) (R/update_steps.r:93: only |U[, 1:k]| is used).  V: [pp][16] row-major, zero padded.
// One CTA per row tile, 256 threads = 64 rows x 4 column quarters (fixed-order combination); out: column-major n x 16.
__global__ void __launch_bounds__(256) rn_xv16(const double* __restrict__ X, int64_t n, int64_t p, int64_t pp,
                                               const double* __restrict__ V, const double* __restrict__ d, int kc,
                                               double* __restrict__ out, int64_t ldo) {
  __shared__ double vs[64][16];
  __shared__ double red[4][64][17];
  const int r = threadIdx.x & 63, q = threadIdx.x >> 6;
  const int64_t tile = blockIdx.x;
  const double* base = X + tile * pp * RN_ROW_TILE;
  double acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.0;
  for (int64_t j0 = 0; j0 < p; j0 += 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int64_t j = j0 + (i >> 4);
      vs[i >> 4][i & 15] = (j < p) ? V[j * 16 + (i & 15)] : 0.0;
    }
    __syncthreads();
    const int jn = (int)min((int64_t)64, p - j0);
    for (int jj = q; jj < jn; jj += 4) {
      const int64_t j = j0 + jj;
      const double x = base[j * RN_ROW_TILE + 2 * ((r >> 1) ^ rn_sigma(j)) + (r & 1)];
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] = fma(x, vs[jj][c], acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < 16; ++c) red[q][r][c] = acc[c];
  __syncthreads();
  const int64_t row = tile * 64 + r;
  if (row < n)
    for (int c = q; c < kc; c += 4) {
      const double s = ((red[0][r][c] + red[1][r][c]) + red[2][r][c]) + red[3][r][c];
      out[row + (int64_t)c * ldo] = fabs(s / d[c]);
    }
}
