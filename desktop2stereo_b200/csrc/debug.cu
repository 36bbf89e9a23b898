// Kernel-level entry points used by the parity tests to exercise the tcgen05 GEMM, the implicit-GEMM 3x3 conv and the
// fused attention in isolation (tests compare each against a plain PyTorch fp32 reference of the same op).
#include "gemm.cuh"
#include "layers.cuh"

using namespace d2s;

extern "C" int d2s_debug_gemm(const void *A, const void *Bw, const float *bias, void *C, int M, int N, int K, int act,
                              float *x32_accumulate, d2s_stream_t stream) {
    D2S_REQUIRE(A && Bw && (C || x32_accumulate), "d2s_debug_gemm: null argument");
    GemmEpi e; e.bias = bias; e.act = act; e.c16 = (__half *)C; e.x32 = x32_accumulate; e.ldc = N;
    GemmPlan p;
    int rc = gemm_plan_linear(&p, (const __half *)A, K, (const __half *)Bw, K, M, N, K, e);
    if (rc) return rc;
    return gemm_launch(&p, (cudaStream_t)stream);
}

extern "C" int d2s_debug_conv3x3(const void *A, const void *Wt, const float *bias, void *C, int B, int H, int W, int Cp, int N, int act,
                                 const void *res1, void *c_relu, d2s_stream_t stream) {
    D2S_REQUIRE(A && Wt && C, "d2s_debug_conv3x3: null argument");
    GemmEpi e; e.bias = bias; e.act = act; e.c16 = (__half *)C; e.res1 = (const __half *)res1; e.c16_relu = (__half *)c_relu; e.ldc = N;
    ConvGeom g; g.B = B; g.H = H; g.W = W; g.Cp = Cp;
    GemmPlan p;
    int rc = gemm_plan_conv3x3(&p, (const __half *)A, g, (const __half *)Wt, N, e);
    if (rc) return rc;
    return gemm_launch(&p, (cudaStream_t)stream);
}

extern "C" int d2s_debug_attention(const void *qkv, void *out, int B, int N, int D, int heads, d2s_stream_t stream) {
    D2S_REQUIRE(qkv && out, "d2s_debug_attention: null argument");
    return attention_launch((const __half *)qkv, (__half *)out, B, N, D, heads, (cudaStream_t)stream);
}
