// How many clusters of a given size / shared-memory footprint can be resident at once on this GPU?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dummy(float* p) { extern __shared__ float sm[]; if (p) p[0] = sm[0]; }
int main() {
  int dev = 0; cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
  printf("%s SMs=%d\n", prop.name, prop.multiProcessorCount);
  int smems[] = {215 * 1024, 178 * 1024, 110 * 1024, 100 * 1024, 64 * 1024};
  int sizes[] = {2, 4, 8, 16};
  for (int cs : sizes) for (int sm : smems) {
    cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    if (cs > 8) cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = sm;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster %2d smem %3d KB: max active clusters %d (%d CTAs) %s\n", cs, sm / 1024, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
