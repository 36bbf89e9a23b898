// Shared host/device helpers for libd2s_b200 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/d2s_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libd2s_b200 is written for sm_100a (B200) only"
#endif

namespace d2s {

extern thread_local std::string g_last_error;
extern std::atomic<long long> g_launch_count;

int set_error(int code, const char *fmt, ...);

#define D2S_CHECK_CUDA(expr)                                                                            \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return ::d2s::set_error(D2S_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                    __FILE__, __LINE__);                                                \
    } while (0)

#define D2S_REQUIRE(cond, ...)                                                 \
    do {                                                                       \
        if (!(cond)) return ::d2s::set_error(D2S_ERR_INVALID, __VA_ARGS__);    \
    } while (0)

// Programmatic dependent launch (PDL).  While the engine records or runs a plan, g_pdl is set and every kernel is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: kernel N+1 may be scheduled as soon as every CTA of kernel N has executed
// griddepcontrol.launch_dependents, runs its prologue (barrier init, TMEM allocation, tensor-map prefetch, index maths) under
// kernel N's tail, and blocks in griddepcontrol.wait until kernel N has completed and flushed its writes.  Every kernel of the
// plan therefore starts with pdl_sync() before it touches data a predecessor produced (a no-op for ordinary launches).
extern thread_local bool g_pdl;
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_wait(); pdl_launch_dependents(); }
#endif

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Every kernel launch of the library goes through this so gpu_launches can be reported.
#define D2S_LAUNCH(kernel, grid, block, smem, stream, ...)                                                                  \
    do {                                                                                                                    \
        (void)::d2s::launch_kernel(kernel, dim3(grid), dim3(block), (size_t)(smem), (cudaStream_t)(stream), __VA_ARGS__);   \
        ::d2s::g_launch_count.fetch_add(1, std::memory_order_relaxed);                                                      \
    } while (0)

#define D2S_POST_LAUNCH() D2S_CHECK_CUDA(cudaPeekAtLastError())

constexpr int kNumSMs = 148;  // B200

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// 2^x on the SFU (MUFU.EX2): 2 ulp, flushes denormals, exp2(-inf) = 0 — what the softmax kernels need, without exp2f()'s range fix-ups
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- dtype load/store (device) ----
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<uint8_t>(uint8_t v) { return (float)v; }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
// u8: round-to-nearest-even of the (already clamped) value, as cv2's saturate_cast does downstream
// of the reference (streamer.py:250-256).
template <> __device__ __forceinline__ uint8_t from_f32<uint8_t>(float v) {
    return (uint8_t)__float2int_rn(fminf(fmaxf(v, 0.f), 255.f));
}

// Round an fp32 value to the value set of T (what ATen's opmath kernels do on store).
template <typename T> __device__ __forceinline__ float round_to(float v) { return to_f32<T>(from_f32<T>(v)); }
template <> __device__ __forceinline__ float round_to<float>(float v) { return v; }

}  // namespace d2s
