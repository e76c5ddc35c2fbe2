// Post-fit reductions on the device: the Jensen-Shannon divergence between kernel density
// estimates of two factor columns, jsd_calc() of R/utils.r:95-106, for a BATCH of column pairs.
//
// The spurious-bicluster test of the reference (R/obtain_bicl.r:55-68, 113-133) calls jsd_calc for every pair of
// columns of the shuffled-refit factors (10 k^2 pairs) and for every (fitted column, shuffled column) pair (5 k^2):
// 960 pairs = 1920 density estimates of 20000 values at k = 8, per fit, 66 fits per apply_resnmtf call.  Once the
// update loop runs on the device this is a quarter of the wall time of the call; one CTA per pair does it here.
//
// jsd_calc(x1, x2):  max_val = max(max x1, max x2);  d_i = density(x_i, from = 0, to = max_val)$y on 512 points,
// zeroed where the grid exceeds max(x_i);  JSD(d1, d2) with base-2 logarithms on the normalised estimates
// (philentropy::JSD, est.prob = "empirical").  density() is stats::density.default with its defaults (gaussian
// kernel, bw.nrd0, n = 512, cut = 3): linear binning of the sample on [from - 4 bw, to + 4 bw] (BinDist), circular
// convolution with the gaussian evaluated on seq(0, 2 (up - lo), length = 2 n) -- the reference does it with fft();
// the same sums are evaluated directly here, in index order, so the result does not depend on the launch -- then
// approx() onto seq(from, to, length = 512).  The bandwidth bw.nrd0(x) and max(x) of every column are inputs (they
// are O(n) per column and need order statistics; the caller computes them once per column, the pairs reuse them).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define RN_KDE_N 512       // grid points of density.default (n = 512) == threads per CTA
#define RN_KDE_WARPS 16    // RN_KDE_N / 32: every warp bins its own contiguous sixteenth of the sample

struct RnKdeSmem {
  double hist[RN_KDE_WARPS][RN_KDE_N];  // per-warp private bin sums (no atomics, no inter-warp conflicts)
  double y[RN_KDE_N];
  double kern[RN_KDE_N];
  double dens[RN_KDE_N];
  double red[RN_KDE_N];
};
static_assert(sizeof(RnKdeSmem) <= 100 * 1024, "two CTAs per SM");

// Sum over the 512 threads in a fixed tree order (every thread receives the result).
__device__ __forceinline__ double rn_kde_block_sum(double v, double* red) {
  const int t = threadIdx.x;
  red[t] = v;
  __syncthreads();
#pragma unroll
  for (int s = RN_KDE_N / 2; s > 0; s >>= 1) {
    if (t < s) red[t] += red[t + s];
    __syncthreads();
  }
  const double r = red[0];
  __syncthreads();
  return r;
}

// density(x, from = 0, to = to)$y at grid point threadIdx.x, zeroed where the grid point exceeds xmax (R/utils.r:99-103).
__device__ double rn_kde_density(const double* __restrict__ x, int64_t n, double bw, double to, double xmax,
                                 RnKdeSmem& sm) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int N = RN_KDE_N;
  const double from = 0.0;
  const double lo = from - 4.0 * bw, up = to + 4.0 * bw;
  const double delta = (up - lo) / (double)(N - 1);
  const double w = 1.0 / (double)n;
  // ---- BinDist: y[ix] += w (1 - fx), y[ix + 1] += w fx ------------------------------------------------------------------
  // Factor columns pile most of their values into a few bins, so the bins cannot be the unit of work.  The SAMPLE is
  // split instead: warp q bins samples [q n/16, (q+1) n/16) into its own histogram, 32 samples at a time; the lanes of
  // a batch that hit the same bin are summed in lane order by the lowest of them, which alone touches the bin (first all
  // left-neighbour weights of the batch, then all right-neighbour weights); at the end the 16 histograms are added in
  // warp order.  Every sum has a fixed order: the result does not depend on the launch.
#pragma unroll
  for (int q = 0; q < RN_KDE_WARPS; ++q) sm.hist[q][t] = 0.0;
  __syncthreads();
  {
    double* hist = sm.hist[warp];
    const int64_t s0 = (n * warp) / RN_KDE_WARPS, s1 = (n * (warp + 1)) / RN_KDE_WARPS;
    for (int64_t i0 = s0; i0 < s1; i0 += 32) {
      const int64_t i = i0 + lane;
      int ixv = -8 - lane;  // no sample / non-finite / far outside: a bin of its own, contributes nowhere
      double wa = 0.0, wb = 0.0;
      if (i < s1) {
        const double xpos = (x[i] - lo) / delta;
        const double fl = floor(xpos);
        const double fx = xpos - fl;
        if (fl >= -1.0 && fl <= (double)(N - 1)) {
          ixv = (int)fl;
          wa = w * (1.0 - fx);
          wb = w * fx;
        }
      }
      const unsigned same = __match_any_sync(0xffffffffu, ixv);
      const bool leader = (__ffs(same) - 1) == lane;
      double sa = 0.0, sb = 0.0;
#pragma unroll
      for (int src = 0; src < 32; ++src) {  // lane order; all lanes run the shuffles, the members of my group count
        const double va = __shfl_sync(0xffffffffu, wa, src);
        const double vb = __shfl_sync(0xffffffffu, wb, src);
        if ((same >> src) & 1u) {
          sa += va;
          sb += vb;
        }
      }
      if (leader && ixv >= 0 && ixv < N) hist[ixv] += sa;
      __syncwarp();
      if (leader && ixv + 1 >= 0 && ixv + 1 < N) hist[ixv + 1] += sb;
      __syncwarp();
    }
  }
  __syncthreads();
  {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < RN_KDE_WARPS; ++q) acc += sm.hist[q][t];
    sm.y[t] = acc;
  }
  // ---- gaussian on the lag grid seq(0, 2 (up - lo), length = 2 n): lag m sits at m * h ------------------------------
  {
    const double h = 2.0 * (up - lo) / (double)(2 * N - 1);
    const double z = ((double)t * h) / bw;
    sm.kern[t] = exp(-0.5 * (z * z)) / (bw * sqrt(2.0 * 3.141592653589793));
  }
  __syncthreads();
  // ---- convolution (what fft(y) * Conj(fft(kords)) evaluates), clipped at 0 ------------------------------------------
  double s = 0.0;
  for (int j = 0; j < N; ++j) {
    const int lag = t >= j ? t - j : j - t;
    s = fma(sm.y[j], sm.kern[lag], s);
  }
  sm.dens[t] = s > 0.0 ? s : 0.0;
  __syncthreads();
  // ---- approx(xords, dens, xout): xords = seq(lo, up, length = n), xout = seq(from, to, length = n) -------------------
  const double xo = (t == N - 1) ? to : (double)t * ((to - from) / (double)(N - 1)) + from;
  auto xord = [&](int j) { return (j == N - 1) ? up : (double)j * delta + lo; };
  int j = (int)floor((xo - lo) / delta);
  j = j < 0 ? 0 : (j > N - 2 ? N - 2 : j);
  while (j > 0 && xord(j) > xo) --j;
  while (j < N - 2 && xord(j + 1) <= xo) ++j;
  double val;
  if (xo >= up) {
    val = sm.dens[N - 1];
  } else {
    const double slope = (sm.dens[j + 1] - sm.dens[j]) / (xord(j + 1) - xord(j));
    val = slope * (xo - xord(j)) + sm.dens[j];
  }
  if (xo > xmax) val = 0.0;
  __syncthreads();  // dens / y / kern are reused by the next density
  return val;
}

// One CTA (512 threads) per pair: out[pair] = jsd_calc(vecs[, a], vecs[, b]).
__global__ void __launch_bounds__(RN_KDE_N)
rn_jsd_pairs(const double* __restrict__ vecs, int64_t n, int64_t ld, const double* __restrict__ bw,
             const double* __restrict__ vmax, const int32_t* __restrict__ pair_a, const int32_t* __restrict__ pair_b,
             int64_t n_pairs, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char rn_kde_raw[];
  RnKdeSmem& sm = *reinterpret_cast<RnKdeSmem*>(rn_kde_raw);
  for (int64_t pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
    const int a = pair_a[pr], b = pair_b[pr];
    const double ma = vmax[a], mb = vmax[b];
    const double to = ma > mb ? ma : mb;  // max_val, R/utils.r:96
    const double d1 = rn_kde_density(vecs + (int64_t)a * ld, n, bw[a], to, ma, sm);
    const double d2 = rn_kde_density(vecs + (int64_t)b * ld, n, bw[b], to, mb, sm);
    const double s1 = rn_kde_block_sum(d1, sm.red);
    const double s2 = rn_kde_block_sum(d2, sm.red);
    const double p = d1 / s1, q = d2 / s2;
    const double m = 0.5 * (p + q);
    // m > 0: the direct sums keep the true far tail of a gaussian (down to denormals, where (p + q) / 2 can round to
    // zero under a non-zero p); the reference's fft() floors the tails at its rounding noise and never gets there
    const double ta = (p > 0.0 && m > 0.0) ? p * log2(p / m) : 0.0;
    const double tb = (q > 0.0 && m > 0.0) ? q * log2(q / m) : 0.0;
    const double ja = rn_kde_block_sum(ta, sm.red);
    const double jb = rn_kde_block_sum(tb, sm.red);
    if (threadIdx.x == 0) out[pr] = 0.5 * ja + 0.5 * jb;
  }
}
