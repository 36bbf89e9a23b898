// probe: which cluster shapes does the launcher accept for a 192-thread kernel with ~98 KB dynamic smem
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(192) k(int *out) { extern __shared__ unsigned char s[]; if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) out[0] = s[0]; }
int main() {
    int *d; cudaMalloc(&d, 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    struct { dim3 grid, cl; size_t smem; } cases[] = {
        {dim3(4, 50, 1), dim3(1, 2, 1), 100000}, {dim3(50, 4, 1), dim3(2, 1, 1), 100000}, {dim3(4, 50, 1), dim3(1, 2, 1), 1000},
        {dim3(4, 50, 1), dim3(1, 1, 1), 100000}, {dim3(6, 7, 3), dim3(1, 1, 3), 100000}, {dim3(4, 50, 1), dim3(1, 2, 1), 150000},
    };
    for (auto &c : cases) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = c.grid; cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = c.smem;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = c.cl.x; at[0].val.clusterDim.y = c.cl.y; at[0].val.clusterDim.z = c.cl.z;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nc = -1; cudaError_t oe = cudaOccupancyMaxActiveClusters(&nc, k, &cfg);
        cudaError_t e = cudaLaunchKernelEx(&cfg, k, d);
        cudaError_t s = cudaDeviceSynchronize();
        printf("grid (%d,%d,%d) cluster (%d,%d,%d) smem %zu: occ %s (%d clusters) launch %s sync %s\n", c.grid.x, c.grid.y, c.grid.z, c.cl.x, c.cl.y, c.cl.z, c.smem,
               cudaGetErrorString(oe), nc, cudaGetErrorString(e), cudaGetErrorString(s));
        cudaGetLastError();
    }
    return 0;
}
