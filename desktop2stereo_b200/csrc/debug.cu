// Kernel-level entry points used by the parity tests to exercise the tcgen05 GEMM, the implicit-GEMM 3x3 conv and the
// fused attention in isolation (tests compare each against a plain PyTorch fp32 reference of the same op).
#include <cstdlib>

#include "gemm.cuh"
#include "layers.cuh"

using namespace d2s;

static int launch_with_scratch(GemmPlan &p, cudaStream_t st) {
    long long *tr = nullptr;
    const char *tv = getenv("D2S_GEMM_TRACE");
    if (tv && tv[0] == '1') { D2S_CHECK_CUDA(cudaMalloc(&tr, 16 * sizeof(long long))); D2S_CHECK_CUDA(cudaMemsetAsync(tr, 0, 16 * sizeof(long long), st)); p.trace = tr; }
    int rc = gemm_launch(&p, st);
    if (tr) {
        long long h[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, tr, sizeof(h), cudaMemcpyDeviceToHost);
        cudaFree(tr);
        fprintf(stderr, "[d2s gemm trace] cycles since entry: setup %lld | first TMA issued %lld | all TMA issued %lld | first full %lld | last MMA committed %lld | "
                        "epilogue start %lld | epilogue end %lld | after final sync %lld | after dealloc %lld\n",
                h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0], h[5] - h[0], h[6] - h[0], h[7] - h[0], h[8] - h[0], h[9] - h[0]);
    }
    return rc;
}

// tile policy of the plans d2s_debug_gemm builds on this thread (enum d2s_policy): lets tests / bench time the GEMM kernel with the
// throughput-policy tiles the pipeline's plans use when several frames are in flight
extern "C" int d2s_debug_set_gemm_policy(int policy) {
    D2S_REQUIRE(policy == D2S_POLICY_LATENCY || policy == D2S_POLICY_THROUGHPUT, "d2s_debug_set_gemm_policy: policy %d", policy);
    gemm_set_plan_policy(policy);
    return D2S_OK;
}

extern "C" int d2s_debug_gemm(const void *A, const void *Bw, const float *bias, void *C, int M, int N, int K, int act,
                              float *x32_accumulate, d2s_stream_t stream) {
    D2S_REQUIRE(A && Bw && (C || x32_accumulate), "d2s_debug_gemm: null argument");
    GemmEpi e; e.bias = bias; e.act = act; e.c16 = (__half *)C; e.x32 = x32_accumulate; e.ldc = N;
    GemmPlan p;
    int rc = gemm_plan_linear(&p, (const __half *)A, K, (const __half *)Bw, K, M, N, K, e);
    if (rc) return rc;
    return launch_with_scratch(p, (cudaStream_t)stream);
}

extern "C" int d2s_debug_conv3x3(const void *A, const void *Wt, const float *bias, void *C, int B, int H, int W, int Cp, int N, int act,
                                 const void *res1, void *c_relu, d2s_stream_t stream) {
    D2S_REQUIRE(A && Wt && C, "d2s_debug_conv3x3: null argument");
    GemmEpi e; e.bias = bias; e.act = act; e.c16 = (__half *)C; e.res1 = (const __half *)res1; e.c16_relu = (__half *)c_relu; e.ldc = N;
    ConvGeom g; g.B = B; g.H = H; g.W = W; g.Cp = Cp;
    GemmPlan p;
    int rc = gemm_plan_conv3x3(&p, (const __half *)A, g, (const __half *)Wt, N, e);
    if (rc) return rc;
    return launch_with_scratch(p, (cudaStream_t)stream);
}

extern "C" int d2s_debug_attention(const void *qkv, void *out, int B, int N, int D, int heads, d2s_stream_t stream) {
    D2S_REQUIRE(qkv && out, "d2s_debug_attention: null argument");
    const char *v = getenv("D2S_ATTN");
    if (!(v && v[0] == 't')) return attention_launch((const __half *)qkv, (__half *)out, B, N, D, heads, (cudaStream_t)stream);   // mma.sync version (default); D2S_ATTN=tcgen05 runs the tcgen05 kernel below
    // tcgen05 version: throw-away V^T scratch (allocated, zeroed, used, synchronised, freed)
    __half *vt = nullptr;
    const size_t n = attention_tc_vt_elems(B, N, heads);
    D2S_CHECK_CUDA(cudaMalloc(&vt, n * sizeof(__half)));
    D2S_CHECK_CUDA(cudaMemsetAsync(vt, 0, n * sizeof(__half), (cudaStream_t)stream));
    AttnTcPlan p;
    int rc = attention_tc_plan(&p, (const __half *)qkv, vt, (__half *)out, B, N, D, heads);
    if (rc == D2S_OK) rc = attention_tc_launch(&p, (cudaStream_t)stream);
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(vt);
    return rc;
}
