// Small device helpers shared by every kernel file of the library: FP64 MMA wrapper, the swizzled panel layout,
// fixed-order reductions, fast division, the stream-K split, mbarrier / bulk-copy (TMA) primitives.
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

// In-column swizzle of the X layout: inside a panel, the 16-byte piece (row pair) rp of data column c is
// stored at piece position rp ^ rn_sigma(c).  With it the MMA fragment reads of BOTH passes hit 8 distinct
// 16 B bank groups per quarter-warp when a unit is copied verbatim into shared memory (TMA path), and
// global accesses stay permutations inside 128 B lines.
__device__ __forceinline__ int rn_sigma(int64_t c) { return (int)(((c & 1) << 2) | (c & 2)); }

// F is stored in 64-row panels like X (F[tile][c][64], pieces swizzled by rn_sigma(c)), so that the 64 rows
// of F a row step needs are one contiguous kp*512 B run (one TMA copy) with conflict-free fragment reads.
__device__ __forceinline__ int64_t rn_fidx(int64_t r, int c, int kp) {
  return ((r >> 6) * kp + c) * RN_ROW_TILE + 2 * ((int)((r & 63) >> 1) ^ rn_sigma(c)) + (r & 1);
}

// Position of (row, col) of a 64 x 8 tile in the fragment-major order the MMA F-step kernels accumulate in:
// lane (g,t) of a warp holds rows 16m + 2g + h, columns 2t + j in register (m, h, j).
__device__ __forceinline__ int rn_ps_index(int row, int col) {
  const int m = row >> 4, g = (row & 15) >> 1, h = row & 1, t = col >> 1, j = col & 1;
  return ((m * 2 + h) * 2 + j) * 32 + 4 * g + t;
}

__device__ __forceinline__ double rn_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum of p[0], p[stride], ..., p[(n-1)*stride] with 8 interleaved partial sums: the loads of a batch are
// independent (pipelined L2 reads instead of a serial chain) and the summation order is fixed.
__device__ __forceinline__ double rn_sum_strided(const double* p, int64_t stride, int64_t n) {
  double s[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  int64_t i = 0;
  for (; i + 8 <= n; i += 8) {
    double v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = __ldcg(p + (i + q) * stride);
#pragma unroll
    for (int q = 0; q < 8; ++q) s[q] += v[q];
  }
  double tail = 0.0;
  for (; i < n; ++i) tail += __ldcg(p + i * stride);
  return (((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]))) + tail;
}

__device__ __forceinline__ int rn_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// num / den without the ~30-instruction IEEE division sequence: hardware reciprocal seed, one Newton step and one
// residual correction (result within 1 ulp; the parity bar is 1e-9).  Operands outside the safe range (zero, huge,
// tiny, Inf, NaN) take the exact division so that Inf / NaN behave as in R.
__device__ __forceinline__ double rn_fast_div(double num, double den) {
  const double ad = fabs(den), an = fabs(num);
  if (!(ad > 1.0e-280 && ad < 1.0e280 && an < 1.0e280)) return num / den;
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
  const double e = fma(-den, r, 1.0);  // seed error <= 2^-20: one Newton step leaves 2^-40, the correction below
  r = fma(r, e, r);                    // squares that again
  double q = num * r;
  const double rem = fma(-den, q, num);
  return fma(rem, r, q);
}

// stream-K partition of U units over C CTAs: CTA c owns [begin(c), begin(c+1)); the first U % C CTAs
// get one unit more.
struct RnSplit {
  int64_t q, rem;
  __device__ __forceinline__ RnSplit(int64_t U, int64_t C) : q(U / C), rem(U % C) {}
  __device__ __forceinline__ int64_t begin(int64_t c) const { return c * q + (c < rem ? c : rem); }
  __device__ __forceinline__ int64_t owner(int64_t u) const {
    const int64_t cut = rem * (q + 1);
    return u < cut ? u / (q + 1) : rem + (u - cut) / q;
  }
};

// ------------------------------------------------------------------------------------------------
// TMA pipeline primitives (sm_90+): mbarrier + 1-D bulk async copy global -> shared.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rn_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void rn_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rn_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void rn_mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void rn_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rn_smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void rn_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rn_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void rn_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "RN_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra RN_DONE_%=;\n"
      "bra RN_WAIT_%=;\n"
      "RN_DONE_%=:\n"
      "}\n" ::"r"(rn_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// one contiguous run global -> shared, completion counted in bytes on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void rn_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   rn_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(rn_smem_u32(bar))
               : "memory");
}
