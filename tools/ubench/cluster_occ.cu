// How many thread-block clusters of a given size (and shared-memory footprint) can be co-resident on this GPU?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int smem[] = {207 * 1024, 100 * 1024, 48 * 1024};
  for (int si = 0; si < 3; ++si) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem[si]);
    for (int cs : {16, 8, 4, 2}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 32); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem[si];
      cudaLaunchAttribute a[1];
      a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
      cfg.attrs = a; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
      printf("smem %3d KB cluster %2d: max active clusters %d (%d CTAs) %s\n", smem[si] / 1024, cs, n, n * cs, cudaGetErrorString(e));
    }
  }
  return 0;
}
