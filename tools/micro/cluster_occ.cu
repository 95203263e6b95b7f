// How many thread-block clusters of size C with ~220 KB of dynamic shared memory can be co-resident on this GPU?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/micro/cluster_occ.cu -o tools/micro/_bin/cluster_occ
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int c : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148 / c * c);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = 220 * 1024;
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeClusterDimension;
    a[0].val.clusterDim.x = c; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    cfg.attrs = a; cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster size %2d: max active clusters %3d -> %3d SMs (%s)\n", c, n, n * c, cudaGetErrorString(e));
  }
  return 0;
}
