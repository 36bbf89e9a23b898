// GPU-box probe: pinned-memory DMA rates with every GPU copying at once — cudaHostAlloc pages (4 KB) vs transparent-huge-page
// backed memory registered with cudaHostRegister.  usage: dma_test <gpu> <mode 0|1> <seconds>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <chrono>
int main(int argc, char **argv) {
    int gpu = atoi(argv[1]), mode = atoi(argv[2]); double secs = atof(argv[3]);
    size_t n = 256u << 20;
    cudaSetDevice(gpu);
    void *h = nullptr, *d = nullptr;
    if (mode == 0) cudaHostAlloc(&h, n, cudaHostAllocDefault);
    else {
        h = mmap(nullptr, n + (2u << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        h = (void *)(((uintptr_t)h + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1));
        int rc = madvise(h, n, MADV_HUGEPAGE);
        memset(h, 1, n);
        cudaError_t e = cudaHostRegister(h, n, cudaHostRegisterDefault);
        if (rc || e != cudaSuccess) printf("gpu %d: madvise rc %d, register %s\n", gpu, rc, cudaGetErrorString(e));
    }
    cudaMalloc(&d, n);
    cudaStream_t s; cudaStreamCreate(&s);
    for (int dir = 0; dir < 2; ++dir) {
        cudaMemcpyAsync(dir ? d : h, dir ? h : d, n, dir ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s);
        auto t0 = std::chrono::steady_clock::now(); int it = 0; double el = 0;
        while (el < secs) {
            cudaMemcpyAsync(dir ? d : h, dir ? h : d, n, dir ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s);
            ++it; el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        printf("gpu %d mode %s %s %.1f GB/s\n", gpu, mode ? "thp+register" : "cudaHostAlloc", dir ? "H2D" : "D2H", it * (double)n / el / 1e9);
    }
    return 0;
}
