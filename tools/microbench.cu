// Micro-benchmarks that size the design decisions of DESIGN.md on the actual B200:
//   (1) FP64 FMA issue rate on the CUDA cores (DFMA), (2) FP64 tensor-core rate (mma.sync m8n8k4),
//   (3) streaming read bandwidth of 128-bit loads in the access patterns the two passes use.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void dfma_peak(double* out, int iters, double a, double b) {
  double c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i];
  if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void dmma_peak(double* out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

// latency of dependent DMMA chains: one warp per SM, NCH independent accumulator chains
template <int NCH>
__global__ void dmma_chain(double* out, int iters, double a, double b, long long* cyc) {
  double c[NCH][2];
#pragma unroll
  for (int i = 0; i < NCH; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__device__ __forceinline__ double2 lds2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
// linear streaming read, grid-stride, U loads in flight per thread
template <int U>
__global__ void read_linear(const double* __restrict__ x, size_t n2, double* out) {
  double acc = 0;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n2; i += U * stride) {
    double2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = lds2(x + 2 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y;
  }
  if (acc == 12345.678) out[0] = acc;
}
// column-walk pattern of the F step: a CTA owns 64 rows; warp w walks columns w, w+8, ...; a warp load is
// 512 B contiguous, successive loads of a warp are ld*8 bytes apart.
template <int U>
__global__ void read_colwalk(const double* __restrict__ x, int64_t ld, int64_t p, double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* xr = x + (int64_t)blockIdx.x * 64 + 2 * lane;
  double acc = 0;
  int64_t j = warp;
  for (; j + 8 * (U - 1) < p; j += 8 * U) {
    double2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = lds2(xr + (j + 8 * u) * ld);
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y;
  }
  if (acc == 12345.678) out[0] = acc;
}

// chunked access of the G step: a CTA reads `chunk` contiguous bytes, then jumps `stride` bytes; CTA b starts
// at chunk index b * chunks_per_cta (contiguous chunk ranges per CTA, like the stream-K unit ranges).
__global__ void read_chunks(const double* __restrict__ x, size_t chunk_bytes, size_t stride_bytes, int chunks_per_cta,
                            size_t total_bytes, double* out) {
  double acc = 0;
  const size_t per16 = chunk_bytes / 16;
  for (int c = 0; c < chunks_per_cta; ++c) {
    size_t base = ((size_t)blockIdx.x * chunks_per_cta + c) * stride_bytes;
    if (base + chunk_bytes > total_bytes) base = base % (total_bytes - chunk_bytes) / 512 * 512;
    const double* p = x + base / 8;
    for (size_t i = threadIdx.x; i < per16; i += blockDim.x * 4) {
      double2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        size_t j = i + (size_t)u * blockDim.x;
        v[u] = j < per16 ? lds2(p + 2 * j) : make_double2(0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += v[u].x + v[u].y;
    }
  }
  if (acc == 12345.678) out[0] = acc;
}

int main(int argc, char** argv) {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  double* out;
  CK(cudaMalloc(&out, 64));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float ms;
  const int sms = prop.multiProcessorCount;
  // ---- DFMA / DMMA peak
  for (int warps = 4; warps <= 32; warps *= 2) {
    int iters = 20000;
    dfma_peak<<<sms * 2, warps * 16>>>(out, 100, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    dfma_peak<<<sms * 2, warps * 16>>>(out, iters, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double fma = (double)sms * 2 * warps * 16 * 8.0 * iters;
    printf("dfma: %2d warps/SM  %.3f ms  %.2f TFMA/s (%.2f TFLOP/s)\n", warps, ms, fma / ms * 1e-9, 2 * fma / ms * 1e-9);
    dmma_peak<<<sms * 2, warps * 16>>>(out, 100, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    dmma_peak<<<sms * 2, warps * 16>>>(out, iters / 4, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double mfma = (double)sms * 2 * (warps * 16 / 32) * 8.0 * (iters / 4) * 256.0;
    printf("dmma: %2d warps/SM  %.3f ms  %.2f TFMA/s (%.2f TFLOP/s)\n", warps, ms, mfma / ms * 1e-9, 2 * mfma / ms * 1e-9);
  }
  {
    long long* cyc;
    CK(cudaMalloc(&cyc, 8));
    long long h;
    const int iters = 4096;
#define CHAIN(N)                                                                                   \
    dmma_chain<N><<<1, 32>>>(out, iters, 1.0000001, 1e-9, cyc);                                     \
    CK(cudaDeviceSynchronize());                                                                    \
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));                                             \
    printf("dmma chain: %d independent chains, 1 warp: %.1f cycles per MMA issue slot, %.1f per chain step\n", N, \
           (double)h / (iters * N), (double)h / iters);
    CHAIN(1) CHAIN(2) CHAIN(4) CHAIN(8)
  }
  // ---- streaming reads over 20000 x 4000 doubles (640 MB, > L2)
  const int64_t n = 20032, p = 4000;
  size_t bytes = (size_t)n * p * 8;
  double* x;
  CK(cudaMalloc(&x, bytes));
  CK(cudaMemset(x, 0, bytes));
  double* y;
  CK(cudaMalloc(&y, bytes));
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    CK(cudaMemcpyAsync(y, x, bytes, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  printf("memcpy d2d: %.3f ms  %.1f GB/s (read+write)\n", ms, 2.0 * bytes / ms * 1e-6);
#define RUN(name, launch)                                                    \
  for (int rep = 0; rep < 3; ++rep) {                                         \
    CK(cudaEventRecord(e0));                                                  \
    launch;                                                                   \
    CK(cudaEventRecord(e1));                                                  \
    CK(cudaDeviceSynchronize());                                              \
    CK(cudaEventElapsedTime(&ms, e0, e1));                                    \
  }                                                                           \
  printf("%-34s %.3f ms  %.1f GB/s\n", name, ms, bytes / ms * 1e-6);
  RUN("read_linear U=4 grid=sms*8 x256", (read_linear<4><<<sms * 8, 256>>>(x, bytes / 16, out)));
  RUN("read_linear U=8 grid=sms*8 x256", (read_linear<8><<<sms * 8, 256>>>(x, bytes / 16, out)));
  RUN("read_linear U=8 grid=sms*4 x512", (read_linear<8><<<sms * 4, 512>>>(x, bytes / 16, out)));
  RUN("read_linear U=16 grid=sms*4 x256", (read_linear<16><<<sms * 4, 256>>>(x, bytes / 16, out)));
  RUN("read_colwalk U=4 (313 CTAs x256)", (read_colwalk<4><<<(int)(n / 64), 256>>>(x, n, p, out)));
  RUN("read_colwalk U=8 (313 CTAs x256)", (read_colwalk<8><<<(int)(n / 64), 256>>>(x, n, p, out)));
  RUN("read_colwalk U=16 (313 CTAs x256)", (read_colwalk<16><<<(int)(n / 64), 256>>>(x, n, p, out)));
  {
    // 20000 chunks of 32 KB = 640 MB; 296 CTAs x 68 chunks
    const size_t chunk = 32768;
    const int ctas = sms * 2, per = 68;
    const size_t moved = (size_t)ctas * per * chunk;
    size_t strides[] = {32768, 65536, 262144, 2048000, 2097152};
    for (size_t st : strides) {
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        read_chunks<<<ctas, 256>>>(x, chunk, st, per, bytes, out);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&ms, e0, e1));
      }
      printf("read_chunks 32KB stride %8zu B      %.3f ms  %.1f GB/s\n", st, ms, moved / ms * 1e-6);
    }
  }
  printf("done\n");
  return 0;
}
