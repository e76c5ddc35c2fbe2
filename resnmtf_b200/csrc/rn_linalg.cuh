// FP64 linear algebra for the SVD initialisation of a fit (init_mats_inner, R/update_steps.r:78-125; SURVEY 8a row a11,
// 8f row N3) on views in the panel layout -- hand-written for sm_100a, no cuBLAS / cuSOLVER:
//   rn_atb          C = alpha A'B (+ beta E1 + gamma E2): FP64 tensor-core MMAs (mma.sync m8n8k4), operands staged
//                   through a TMA ring; A, B are n x pa / n x pb panel matrices, the contraction runs down the rows.
//                   Gram matrix of a view (A = B = X, symmetric: upper tiles computed, mirrored), W Q for a symmetric W
//                   held in the panel layout (W Q = W'Q), Q'(W Q), Q'Q -- every product of the subspace iteration.
//   rn_panel_mul    out = beta Z + alpha Y M for a small dense M (<= 64 x 64): Ritz rotations, R^-1 of the
//                   Cholesky QR, deflation against locked vectors
//   rn_deflate      W -= V diag(lambda) V'
// A 64-row tile of 64 columns is one contiguous 32 KB run in the panel layout, and its verbatim copy in shared memory
// is bank-conflict-free for the LDS.128 fragment reads (rn_sigma swizzle): one LDS.128 yields rows (2q, 2q+1) of a
// column, i.e. the operand of two consecutive k-steps.
#pragma once
#include "rn_data.cuh"

#define RN_ATB_THREADS 288  // 8 consumer warps (2 x 4 over a 64 x 64 output tile) + 1 producer warp
#define RN_ATB_STAGES 3
#define RN_ATB_BLOCK_BYTES (64 * 512)
#define RN_ATB_STAGE_BYTES (2 * RN_ATB_BLOCK_BYTES)
static inline size_t rn_atb_smem() { return (size_t)RN_ATB_STAGES * RN_ATB_STAGE_BYTES + 2 * RN_ATB_STAGES * 8 + 64; }

struct RnAtb {
  const double* A;
  const double* B;
  int64_t ppa, ppb;  // padded column counts of A and B (multiples of 32)
  int64_t pa, pb;    // columns of A / B that produce output rows / columns
  int row_tiles;     // 64-row tiles of A and B (same number of rows)
  int symmetric;     // 1: A == B; only tiles ti <= tj are computed, the result is mirrored
  int tiles_i, tiles_j;
  int splits;        // split of the row tiles over gridDim.y CTAs per output tile (partials summed in order)
  double* part;      // [tiles][splits][64 * 64] when splits > 1
  int* ticket;       // [tiles], zero on entry and on exit
  double alpha, beta, gamma;
  const double* E1;  // optional, same layout as C
  const double* E2;
  double* C;
  int c_panel;       // 1: C in the panel layout with ppc padded columns; 0: column-major, leading dimension ldc
  int64_t ppc, ldc;
};

__device__ __forceinline__ int64_t rn_atb_cidx(const RnAtb& a, int64_t i, int64_t j) {
  return a.c_panel ? rn_xidx(i, j, a.ppc) : i + j * a.ldc;
}

__global__ void __launch_bounds__(RN_ATB_THREADS, 1) rn_atb(const RnAtb a) {
  extern __shared__ __align__(128) unsigned char rn_smem[];
  unsigned char* ring = rn_smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(rn_smem + (size_t)RN_ATB_STAGES * RN_ATB_STAGE_BYTES);
  uint64_t* empty = full + RN_ATB_STAGES;
  int* s_last = reinterpret_cast<int*>(empty + RN_ATB_STAGES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;

  // output tile of this CTA
  int ti, tj;
  if (a.symmetric) {  // upper triangle, row by row: tile index -> (ti, tj >= ti)
    int rem = blockIdx.x;
    ti = 0;
    while (rem >= a.tiles_j - ti) {
      rem -= a.tiles_j - ti;
      ++ti;
    }
    tj = ti + rem;
  } else {
    ti = blockIdx.x / a.tiles_j;
    tj = blockIdx.x % a.tiles_j;
  }
  const bool diag = a.symmetric && ti == tj;
  const int64_t i0 = (int64_t)ti * 64, j0 = (int64_t)tj * 64;
  const uint32_t bytes_a = (uint32_t)min((int64_t)64, a.ppa - i0) * 512u;
  const uint32_t bytes_b = diag ? 0u : (uint32_t)min((int64_t)64, a.ppb - j0) * 512u;
  const RnSplit sp(a.row_tiles, a.splits);
  const int t0 = (int)sp.begin(blockIdx.y), t1 = (int)sp.begin(blockIdx.y + 1);

  // columns past the padded width are never copied: they must read as zero
  for (int i = tid; i < RN_ATB_STAGES * RN_ATB_STAGE_BYTES / 16; i += RN_ATB_THREADS)
    reinterpret_cast<double2*>(ring)[i] = make_double2(0.0, 0.0);
  if (tid == 0) {
    for (int s = 0; s < RN_ATB_STAGES; ++s) {
      rn_mbar_init(&full[s], 1);
      rn_mbar_init(&empty[s], 8);
    }
    rn_mbar_init_fence();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the zero fill is ordered before the bulk copies
  __syncthreads();

  double acc[4][2][2];
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 2; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;

  if (warp == 8) {
    if (lane == 0) {
      for (int tt = t0; tt < t1; ++tt) {
        const int st = (tt - t0) % RN_ATB_STAGES;
        const uint32_t ph = (uint32_t)(((tt - t0) / RN_ATB_STAGES) & 1);
        rn_mbar_wait(&empty[st], ph ^ 1u);
        rn_mbar_expect_tx(&full[st], bytes_a + bytes_b);
        unsigned char* dst = ring + (size_t)st * RN_ATB_STAGE_BYTES;
        rn_bulk_g2s(dst, a.A + ((int64_t)tt * a.ppa + i0) * RN_ROW_TILE, bytes_a, &full[st]);
        if (!diag) rn_bulk_g2s(dst + RN_ATB_BLOCK_BYTES, a.B + ((int64_t)tt * a.ppb + j0) * RN_ROW_TILE, bytes_b, &full[st]);
      }
    }
  } else {
    const int wi = warp >> 2, wj = warp & 3;  // warp tile: rows 32 wi .. +31, columns 16 wj .. +15 of the output tile
    // byte offset of this lane's piece inside a 64-column block: column c, row pair 4 step + t
    uint32_t offa[4], offb[2];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int c = 32 * wi + 8 * x + g;
      offa[x] = (uint32_t)(c * 512 + 16 * (t ^ rn_sigma(c)));
    }
#pragma unroll
    for (int y = 0; y < 2; ++y) {
      const int c = 16 * wj + 8 * y + g;
      offb[y] = (uint32_t)(c * 512 + 16 * (t ^ rn_sigma(c))) + (diag ? 0u : (uint32_t)RN_ATB_BLOCK_BYTES);
    }
    for (int tt = t0; tt < t1; ++tt) {
      const int st = (tt - t0) % RN_ATB_STAGES;
      rn_mbar_wait(&full[st], (uint32_t)(((tt - t0) / RN_ATB_STAGES) & 1));
      const unsigned char* sb = ring + (size_t)st * RN_ATB_STAGE_BYTES;
#pragma unroll
      for (int step = 0; step < 8; ++step) {
        // (4 step + t) ^ sigma == (t ^ sigma) + 4 step when step is even ... not in general: sigma has bit 2 set for
        // odd columns, so the piece index is recomputed by XOR on the byte offset instead
        double2 fa[4], fb[2];
#pragma unroll
        for (int x = 0; x < 4; ++x) fa[x] = *reinterpret_cast<const double2*>(sb + (offa[x] ^ (uint32_t)(64 * step)));
#pragma unroll
        for (int y = 0; y < 2; ++y) fb[y] = *reinterpret_cast<const double2*>(sb + (offb[y] ^ (uint32_t)(64 * step)));
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 2; ++y) {
            rn_dmma(acc[x][y][0], acc[x][y][1], fa[x].x, fb[y].x);
            rn_dmma(acc[x][y][0], acc[x][y][1], fa[x].y, fb[y].y);
          }
      }
      __syncwarp();
      if (lane == 0) rn_mbar_arrive(&empty[st]);
    }
  }
  __syncthreads();
  if (warp == 8) return;

  // ---- epilogue: 256 consumer threads ----------------------------------------------------------------------
  const int wi = warp >> 2, wj = warp & 3;
  const int tile_id = blockIdx.x;
  if (a.splits > 1) {
    double* mine = a.part + ((int64_t)tile_id * a.splits + blockIdx.y) * 4096;
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 2; ++y) {
        const int li = 32 * wi + 8 * x + g, lj = 16 * wj + 8 * y + 2 * t;
        *reinterpret_cast<double2*>(mine + li * 64 + lj) = make_double2(acc[x][y][0], acc[x][y][1]);
      }
    __threadfence();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (tid == 0) *s_last = (atomicAdd(&a.ticket[tile_id], 1) == a.splits - 1);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (!*s_last) return;
    __threadfence();
    if (tid == 0) a.ticket[tile_id] = 0;
    const double* base = a.part + (int64_t)tile_id * a.splits * 4096;
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 2; ++y) {
        const int li = 32 * wi + 8 * x + g, lj = 16 * wj + 8 * y + 2 * t;
        double s0 = 0.0, s1 = 0.0;
        for (int s = 0; s < a.splits; ++s) {
          const double2 v = __ldcg(reinterpret_cast<const double2*>(base + (int64_t)s * 4096 + li * 64 + lj));
          s0 += v.x;
          s1 += v.y;
        }
        acc[x][y][0] = s0;
        acc[x][y][1] = s1;
      }
  }
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 2; ++y)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int64_t i = i0 + 32 * wi + 8 * x + g, j = j0 + 16 * wj + 8 * y + 2 * t + e;
        if (i >= a.pa || j >= a.pb) continue;
        const int64_t idx = rn_atb_cidx(a, i, j);
        double v = a.alpha * acc[x][y][e];
        if (a.E1) v = fma(a.beta, a.E1[idx], v);
        if (a.E2) v = fma(a.gamma, a.E2[idx], v);
        a.C[idx] = v;
        if (a.symmetric && ti != tj) a.C[rn_atb_cidx(a, j, i)] = v;
      }
}

// out[r][co0 + c] = beta Z[r][c] + alpha sum_b Y[r][b] M[b][c]   (c < bo, b < bi <= 64, bo <= 64) on the m rows of
// panel matrices; M: bi x bo column-major (leading dimension ldm) in global memory.  out may alias Z (not Y).
// One CTA per row tile, 256 threads = 64 rows x 4 column quarters.
__global__ void __launch_bounds__(256) rn_panel_mul(double* __restrict__ out, int64_t ppo, int co0, const double* Z,
                                                    int64_t ppz, const double* __restrict__ Y, int64_t ppy,
                                                    const double* __restrict__ M, int ldm, int bi, int bo, double alpha,
                                                    double beta) {
  __shared__ double ms[64][65];
  const int r = threadIdx.x & 63, q = threadIdx.x >> 6;
  for (int i = threadIdx.x; i < bi * bo; i += 256) ms[i % bi][i / bi] = M[(i % bi) + (int64_t)(i / bi) * ldm];
  __syncthreads();
  const int64_t tile = blockIdx.x;
  const double* yb = Y + tile * ppy * RN_ROW_TILE;
  double acc[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) acc[u] = 0.0;
  for (int b = 0; b < bi; ++b) {
    const double y = yb[(int64_t)b * RN_ROW_TILE + 2 * ((r >> 1) ^ rn_sigma(b)) + (r & 1)];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int c = q + 4 * u;
      if (c < bo) acc[u] = fma(y, ms[b][c], acc[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int c = q + 4 * u;
    if (c >= bo) continue;
    double v = alpha * acc[u];
    if (Z) v = fma(beta, Z[(tile * ppz + c) * RN_ROW_TILE + 2 * ((r >> 1) ^ rn_sigma(c)) + (r & 1)], v);
    const int co = co0 + c;
    out[(tile * ppo + co) * RN_ROW_TILE + 2 * ((r >> 1) ^ rn_sigma(co)) + (r & 1)] = v;
  }
}

// W[i][j] -= sum_c lam[c] V[i][c] V[j][c]  (i, j < m; c < kc <= 16): the locked eigenpairs are moved to eigenvalue 0.
// grid (tiles, tiles), 256 threads; W m x m and V m x kc in the panel layout.
__global__ void __launch_bounds__(256) rn_deflate(double* __restrict__ W, int64_t m, int64_t ppw, const double* __restrict__ V,
                                                  int64_t ppv, const double* __restrict__ lam, int kc) {
  __shared__ double vi[64][17], vj[64][17];
  const int64_t i0 = (int64_t)blockIdx.x * 64, j0 = (int64_t)blockIdx.y * 64;
  for (int e = threadIdx.x; e < 64 * kc; e += 256) {
    const int r = e & 63, c = e >> 6;
    vi[r][c] = (i0 + r < m) ? V[rn_xidx(i0 + r, c, ppv)] * lam[c] : 0.0;
    vj[r][c] = (j0 + r < m) ? V[rn_xidx(j0 + r, c, ppv)] : 0.0;
  }
  __syncthreads();
  const int r = threadIdx.x & 63;
  for (int jj = threadIdx.x >> 6; jj < 64; jj += 4) {
    const int64_t i = i0 + r, j = j0 + jj;
    if (i >= m || j >= m) continue;
    double s = 0.0;
    for (int c = 0; c < kc; ++c) s = fma(vi[r][c], vj[jj][c], s);
    W[rn_xidx(i, j, ppw)] -= s;
  }
}

// out = a X + b Y + c Z over whole panel buffers (padding stays zero); Y / Z may be NULL
__global__ void __launch_bounds__(256) rn_panel_axpbypcz(double* __restrict__ out, double a, const double* X, double b,
                                                         const double* Y, double c, const double* Z, int64_t count) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < count; i += (int64_t)gridDim.x * 256) {
    double v = a * X[i];
    if (Y) v = fma(b, Y[i], v);
    if (Z) v = fma(c, Z[i], v);
    out[i] = v;
  }
}

// out[:, j] = AQ[:, j] - theta[j] Q[:, j]  (residual block of the Rayleigh-Ritz pairs), j < b
__global__ void __launch_bounds__(256) rn_panel_resid(double* __restrict__ out, const double* __restrict__ AQ,
                                                      const double* __restrict__ Q, const double* __restrict__ theta,
                                                      int64_t pp, int b, int64_t tiles) {
  const int64_t total = tiles * b * 64;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int pos = (int)(i & 63);
    const int64_t j = (i >> 6) % b, tile = (i >> 6) / b;
    const int64_t idx = (tile * pp + j) * RN_ROW_TILE + pos;
    out[idx] = AQ[idx] - theta[j] * Q[idx];
  }
}

// dst[:, dc0 + c] = src[:, sc0 + c] for c < nc (both m-row panel matrices)
__global__ void __launch_bounds__(256) rn_panel_copy_cols(double* __restrict__ dst, int64_t ppd, int dc0,
                                                          const double* __restrict__ src, int64_t pps, int sc0, int nc,
                                                          int64_t m) {
  const int64_t total = m * nc;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i % m;
    const int c = (int)(i / m);
    dst[rn_xidx(r, dc0 + c, ppd)] = src[rn_xidx(r, sc0 + c, pps)];
  }
}

// deterministic start block of the subspace iteration: uniform(-1, 1) from a counter hash of (seed, row, column)
__global__ void __launch_bounds__(256) rn_panel_random(double* __restrict__ out, int64_t m, int64_t pp, int b,
                                                       uint64_t seed) {
  const int64_t total = m * b;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i % m, c = i / m;
    uint64_t x = seed + 0x9e3779b97f4a7c15ULL * (uint64_t)(i + 1);
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    out[rn_xidx(r, c, pp)] = (double)(x >> 11) * (2.0 / 9007199254740992.0) - 1.0;
  }
}
