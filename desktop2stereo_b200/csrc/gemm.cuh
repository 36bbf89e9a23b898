// tcgen05 GEMM of libd2s_b200:  C[M,N] = A[M,K] * B[N,K]^T  (+ fused epilogue), fp16 operands, fp32 accumulate in TMEM.
//
// A is either a row-major matrix (2-D TMA) or an NHWC activation read as an implicit 3x3/pad-1 im2col
// (4-D TMA boxes, out-of-bounds = zero = the conv padding).  B is always the [N, K] row-major weight.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace d2s {

enum { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_SIGMOID = 3 };

struct GemmEpi {
    const float *bias = nullptr;     // [N]
    int act = ACT_NONE;              // applied to acc + bias (before residuals)
    const __half *res1 = nullptr;    // fp16 residuals, addressed like c16
    const __half *res2 = nullptr;
    __half *c16 = nullptr;           // fp16 output
    __half *c16_relu = nullptr;      // optional relu(v) copy (input of the next pre-activation conv)
    // optional transposed copy of the columns >= vt_col0 (the V third of a fused qkv projection) for the tcgen05 attention:
    // vt[((row / vt_tokens) * vt_heads + head) * 64 + d][row % vt_tokens], row pitch vt_npad, head = (col - vt_col0) / 64
    __half *vt = nullptr;
    int vt_col0 = 0, vt_tokens = 1, vt_heads = 1, vt_npad = 0;
    float *x32 = nullptr;            // fp32 residual stream: x32[row, col] += v   (x32_assign: = v)
    int x32_assign = 0;
    float *c32 = nullptr;            // fp32 output (plain store)
    int ldc = 0;                     // leading dimension (elements) of every output / residual above
    // fused 1x1 head (requires N <= 32): depth[row] = final_act(sum_n v[n] * w3[n] + b3) * max_depth
    const float *w3 = nullptr;
    float b3 = 0.f;
    int final_act = ACT_RELU;
    float max_depth = 1.f;
    void *depth_out = nullptr;
    int depth_dtype = D2S_F32;
};

struct ConvGeom {  // NHWC activation [B, H, W, Cp] for the implicit-GEMM A operand
    int B = 0, H = 0, W = 0, Cp = 0;
};

struct GemmPlan {  // everything a launch needs; built once per shape, replayed every frame
    CUtensorMap tmA, tmB;
    int M, N, K;          // conv: M = B*tiles*128 (padded), K = 9*Cp
    int BN;               // 32 / 64 / 128 / 256
    int pair;             // 1: cta_group::2 pairs along M (256 x BN tiles)
    int persist;          // 1: persistent tile loop with a double-buffered TMEM accumulator (grid = one CTA or pair per SM)
    int mtiles;           // real number of 128-row M tiles (the grid of a pair plan is padded to an even count)
    int stages;
    int conv;             // 0 linear, 1 implicit 3x3
    int H, W, Cp, TH, TW, tiles_x, tiles_y, B;
    int kblocks;
    int splits, kb_per_split;      // split-K: grid.z = cluster size; partial sums meet through distributed shared memory
    long long *trace;              // debug: clock64 stamps of CTA (0,0,0)
    GemmEpi epi;
    dim3 grid;
    size_t smem;
};

// A: [M,K] row-major fp16 (lda elements), Bw: [N,K] row-major fp16 (ldb elements)
int gemm_plan_linear(GemmPlan *p, const __half *A, int lda, const __half *Bw, int ldb, int M, int N, int K, const GemmEpi &epi);
// A: NHWC fp16 [B,H,W,Cp]; Bw: [N, 9*Cp] with k = (ky*3+kx)*Cp + c.  Output rows are pixels (b,y,x) -> (b*H+y)*W+x.
int gemm_plan_conv3x3(GemmPlan *p, const __half *A, const ConvGeom &g, const __half *Bw, int N, const GemmEpi &epi);
int gemm_launch(const GemmPlan *p, cudaStream_t stream);
// fp16 [rows, cols] row-major (pitch ld_elems) -> tensor map with 64-column x box_rows boxes, SWIZZLE_128B (after gemm_init())
int tma_encode_2d(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows);
void gemm_set_plan_policy(int policy);   // 0 latency, 1 throughput: tile-shape heuristics of the plans built next on this thread
int gemm_init();  // resolves cuTensorMapEncodeTiled, sets kernel attributes (idempotent)

}  // namespace d2s
