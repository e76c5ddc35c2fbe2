// How many thread-block clusters of each size stay resident on this GPU with the shared-memory footprint of the
// one-pass fused kernel (one CTA per SM)?  Decides which cluster shapes can occupy all 148 SMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/cluster_probe tools/cluster_probe.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>

__global__ void probe_kernel(int* out) {
  extern __shared__ unsigned char smem[];
  if (out && threadIdx.x == 0) out[blockIdx.x] = smem[0];
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  std::printf("%s: %d SMs, smem/block optin %zu\n", prop.name, prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
  const int smems[] = {64 << 10, 112 << 10, 200 << 10, 220 << 10};
  const int threads[] = {384, 512, 640};
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int smem : smems) {
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int th : threads) {
      std::printf("smem %3d KB, %d threads:", smem >> 10, th);
      for (int cs = 1; cs <= 16; ++cs) {
        if (cs > 8 && cs != 16 && cs != 12 && cs != 10) continue;
        cudaLaunchConfig_t cfg;
        std::memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)(cs * prop.multiProcessorCount));
        cfg.blockDim = dim3((unsigned)th);
        cfg.dynamicSmemBytes = (size_t)smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cs;
        attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
        if (e != cudaSuccess) {
          cudaGetLastError();
          std::printf("  c%d: err", cs);
        } else {
          std::printf("  c%d: %d (%d SMs)", cs, n, n * cs);
        }
      }
      std::printf("\n");
    }
  }
  return 0;
}
