// Warp-specialised tcgen05 GEMM for sm_100a:  TMA -> 128B-swizzled smem ring -> tcgen05.mma (fp32 accum in TMEM)
// -> tcgen05.ld epilogue with fused bias / GELU / ReLU / residual adds / fp32 residual-stream update / 1x1 head.
//
// Roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..5 = epilogue (each owns the 32 TMEM lanes its warp-id % 4 selects).
// One 128 x BN output tile per CTA; the ring depth is chosen so two CTAs fit per SM, which overlaps one CTA's
// epilogue with the other's main loop.
#include <mutex>

#include "gemm.cuh"
#include "ptx.cuh"

namespace d2s {

constexpr int BM = 128;
constexpr int BK = 64;                    // 64 fp16 = one 128-byte swizzle row
constexpr int kABytes = BM * BK * 2;      // 16 KiB
constexpr int kGemmThreads = 192;

struct GemmArgs {
    CUtensorMap tmA, tmB;
    int M, N, kblocks, stages;
    int conv, H, W, Cp, TW, TW_shift, tiles_x, tiles_y;
    GemmEpi epi;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == ACT_GELU) return gelu_erf(v);
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
    return v;
}

template <int BN>
__global__ void __launch_bounds__(kGemmThreads) gemm_tc_kernel(const __grid_constant__ GemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int kBBytes = BN * BK * 2;
    constexpr int kStageBytes = kABytes + kBBytes;
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B needs 1024-B alignment
    const int stages = g.stages;
    uint64_t *full = (uint64_t *)(smem + (size_t)stages * kStageBytes);
    uint64_t *empty = full + stages;
    uint64_t *tmem_full = empty + stages;
    uint32_t *tmem_slot = (uint32_t *)(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_n = blockIdx.x, tile_m = blockIdx.y;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&g.tmA);
        ptx::prefetch_tensormap(&g.tmB);
        for (int s = 0; s < stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
        ptx::mbar_init(tmem_full, 1);
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, BN);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // conv tile -> (image, y0, x0)
    int img = 0, y0 = 0, x0 = 0;
    if (g.conv) {
        int per_img = g.tiles_x * g.tiles_y;
        img = tile_m / per_img;
        int t = tile_m - img * per_img;
        y0 = (t / g.tiles_x) * (BM >> g.TW_shift);
        x0 = (t % g.tiles_x) * g.TW;
    }

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            const int cchunks = g.conv ? (g.Cp / BK) : 1;
            for (int kb = 0; kb < g.kblocks; ++kb) {
                const int s = kb % stages;
                const uint32_t ph = (kb / stages) & 1;
                ptx::mbar_wait(&empty[s], ph ^ 1);
                uint8_t *a = smem + (size_t)s * kStageBytes, *b = a + kABytes;
                ptx::mbar_arrive_expect_tx(&full[s], kStageBytes);
                if (g.conv) {
                    int tap = kb / cchunks, cc = kb - tap * cchunks;
                    int dy = tap / 3 - 1, dx = tap % 3 - 1;
                    ptx::tma_load_4d(a, &g.tmA, &full[s], cc * BK, x0 + dx, y0 + dy, img);
                } else {
                    ptx::tma_load_2d(a, &g.tmA, &full[s], kb * BK, tile_m * BM);
                }
                ptx::tma_load_2d(b, &g.tmB, &full[s], kb * BK, tile_n * BN);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc = ptx::make_idesc_f16(BM, BN, 0);
            for (int kb = 0; kb < g.kblocks; ++kb) {
                const int s = kb % stages;
                const uint32_t ph = (kb / stages) & 1;
                ptx::mbar_wait(&full[s], ph);
                ptx::tc_fence_after();
                const uint32_t a_addr = ptx::smem_u32(smem + (size_t)s * kStageBytes);
                const uint64_t da = ptx::make_sw128_kmajor_desc(a_addr);
                const uint64_t db = ptx::make_sw128_kmajor_desc(a_addr + kABytes);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)  // UMMA_K = 16 fp16 = 32 bytes: advance the start address by 2 (>>4)
                    ptx::umma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                ptx::umma_commit(&empty[s]);  // smem slot is free once these MMAs have read it
            }
            ptx::umma_commit(tmem_full);      // accumulator complete
        }
    } else {
        // ===== epilogue: TMEM -> registers -> global =====
        const GemmEpi &e = g.epi;
        const int q = warp & 3;               // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;          // row inside the tile
        long long orow;                        // output row (pixel / token) or -1
        if (g.conv) {
            int y = y0 + (r >> g.TW_shift), x = x0 + (r & (g.TW - 1));
            orow = (y < g.H && x < g.W) ? ((long long)img * g.H + y) * g.W + x : -1;
        } else {
            int m = tile_m * BM + r;
            orow = m < g.M ? m : -1;
        }
        ptx::mbar_wait(tmem_full, 0);
        ptx::tc_fence_after();
        float head_acc = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t raw[32];
            ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), raw);
            ptx::tmem_ld_wait();
            const int n0 = tile_n * BN + c * 32;
            if (orow < 0 || n0 >= g.N) continue;
            const long long off = orow * e.ldc + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                if (n0 + j >= g.N) break;      // N is a multiple of 8
                float v[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(raw[j + t]);
                if (e.bias) {
                    float4 b0 = __ldg((const float4 *)(e.bias + n0 + j)), b1 = __ldg((const float4 *)(e.bias + n0 + j + 4));
                    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                }
                if (e.act != ACT_NONE) {
#pragma unroll
                    for (int t = 0; t < 8; ++t) v[t] = apply_act(v[t], e.act);
                }
                if (e.res1) {
                    uint4 u = __ldg((const uint4 *)(e.res1 + off + j));
                    const __half2 *h = (const __half2 *)&u;
#pragma unroll
                    for (int t = 0; t < 4; ++t) { float2 f = __half22float2(h[t]); v[2 * t] += f.x; v[2 * t + 1] += f.y; }
                }
                if (e.res2) {
                    uint4 u = __ldg((const uint4 *)(e.res2 + off + j));
                    const __half2 *h = (const __half2 *)&u;
#pragma unroll
                    for (int t = 0; t < 4; ++t) { float2 f = __half22float2(h[t]); v[2 * t] += f.x; v[2 * t + 1] += f.y; }
                }
                if (e.w3) {
#pragma unroll
                    for (int t = 0; t < 8; ++t) head_acc = fmaf(v[t], __ldg(e.w3 + n0 + j + t), head_acc);
                }
                if (e.x32) {
                    float4 *p = (float4 *)(e.x32 + off + j);
                    float4 a = p[0], b = p[1];
                    a.x += v[0]; a.y += v[1]; a.z += v[2]; a.w += v[3]; b.x += v[4]; b.y += v[5]; b.z += v[6]; b.w += v[7];
                    p[0] = a; p[1] = b;
                }
                if (e.c32) {
                    float4 *p = (float4 *)(e.c32 + off + j);
                    p[0] = make_float4(v[0], v[1], v[2], v[3]);
                    p[1] = make_float4(v[4], v[5], v[6], v[7]);
                }
                if (e.c16) {
                    __half2 h[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) h[t] = __floats2half2_rn(v[2 * t], v[2 * t + 1]);
                    *(uint4 *)(e.c16 + off + j) = *(const uint4 *)h;
                }
                if (e.c16_relu) {
                    __half2 h[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) h[t] = __floats2half2_rn(fmaxf(v[2 * t], 0.f), fmaxf(v[2 * t + 1], 0.f));
                    *(uint4 *)(e.c16_relu + off + j) = *(const uint4 *)h;
                }
            }
        }
        if (e.w3 && orow >= 0 && tile_n == 0) {
            float d = apply_act(head_acc + e.b3, e.final_act) * e.max_depth;
            if (e.depth_dtype == D2S_F16) ((__half *)e.depth_out)[orow] = __float2half_rn(d);
            else ((float *)e.depth_out)[orow] = d;
        }
        ptx::tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, BN);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_once;
static int g_init_rc = D2S_OK;
constexpr size_t kMaxSmem = 227 * 1024;

template <int BN> static size_t smem_for(int stages) { return (size_t)stages * (kABytes + BN * BK * 2) + (2 * stages + 1) * 8 + 16 + 1024; }

int gemm_init() {
    std::call_once(g_once, [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            g_init_rc = set_error(D2S_ERR_CUDA, "cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
            return;
        }
        g_encode = (EncodeTiledFn)fn;
        cudaError_t a = cudaFuncSetAttribute(gemm_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
        if (a == cudaSuccess) a = cudaFuncSetAttribute(gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
        if (a == cudaSuccess) a = cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
        if (a != cudaSuccess) g_init_rc = set_error(D2S_ERR_CUDA, "cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(a));
    });
    return g_init_rc;
}

static int encode_2d(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld_elems * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t es[2] = {1, 1};
    if (((uintptr_t)base & 15) || (strides[0] & 15)) return set_error(D2S_ERR_INVALID, "gemm: operand base/pitch must be 16-byte aligned");
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(D2S_ERR_CUDA, "cuTensorMapEncodeTiled(2d %llux%llu ld %llu) failed: %d", (unsigned long long)rows,
                                            (unsigned long long)cols, (unsigned long long)ld_elems, (int)r);
    return D2S_OK;
}

static int encode_nhwc(CUtensorMap *m, const void *base, const ConvGeom &g, int TW, int TH) {
    cuuint64_t dims[4] = {(cuuint64_t)g.Cp, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.B};
    cuuint64_t strides[3] = {(cuuint64_t)g.Cp * 2, (cuuint64_t)g.W * g.Cp * 2, (cuuint64_t)g.H * g.W * g.Cp * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)TW, (cuuint32_t)TH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (((uintptr_t)base & 15) || (g.Cp % BK)) return set_error(D2S_ERR_INVALID, "conv: NHWC base must be 16-B aligned and Cp a multiple of %d", BK);
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(D2S_ERR_CUDA, "cuTensorMapEncodeTiled(nhwc %dx%dx%dx%d) failed: %d", g.B, g.H, g.W, g.Cp, (int)r);
    return D2S_OK;
}

static int pick_bn(int N) { return N <= 32 ? 32 : (N <= 64 ? 64 : 128); }

static void finish_plan(GemmPlan *p) {
    // ring depth: as deep as useful while leaving room for two CTAs per SM
    size_t stage = kABytes + (size_t)p->BN * BK * 2;
    int stages = (int)((kMaxSmem / 2 - 2048) / stage);
    if (stages > p->kblocks) stages = p->kblocks;
    if (stages < 2) stages = 2;
    if (stages > 8) stages = 8;
    p->stages = stages;
    p->smem = (size_t)stages * stage + (2 * stages + 1) * 8 + 16 + 1024;
}

static int check_epi(const GemmEpi &e, int N) {
    D2S_REQUIRE(N % 8 == 0, "gemm: N=%d must be a multiple of 8", N);
    D2S_REQUIRE(e.ldc % 8 == 0 || !(e.c16 || e.c16_relu || e.res1 || e.res2 || e.x32 || e.c32), "gemm: ldc=%d must be a multiple of 8", e.ldc);
    D2S_REQUIRE(!e.w3 || N <= 32, "gemm: fused 1x1 head needs N <= 32 (got %d)", N);
    return D2S_OK;
}

int gemm_plan_linear(GemmPlan *p, const __half *A, int lda, const __half *Bw, int ldb, int M, int N, int K, const GemmEpi &epi) {
    int rc = gemm_init();
    if (rc) return rc;
    if ((rc = check_epi(epi, N))) return rc;
    D2S_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
    *p = GemmPlan{};
    p->M = M; p->N = N; p->K = K; p->BN = pick_bn(N); p->conv = 0; p->epi = epi;
    p->kblocks = ceil_div(K, BK);
    if ((rc = encode_2d(&p->tmA, A, M, K, lda, BM))) return rc;
    if ((rc = encode_2d(&p->tmB, Bw, N, K, ldb, p->BN))) return rc;
    p->grid = dim3(ceil_div(N, p->BN), ceil_div(M, BM));
    finish_plan(p);
    return D2S_OK;
}

int gemm_plan_conv3x3(GemmPlan *p, const __half *A, const ConvGeom &g, const __half *Bw, int N, const GemmEpi &epi) {
    int rc = gemm_init();
    if (rc) return rc;
    if ((rc = check_epi(epi, N))) return rc;
    D2S_REQUIRE(g.B > 0 && g.H > 0 && g.W > 0 && g.Cp > 0, "conv: bad geometry");
    *p = GemmPlan{};
    p->conv = 1; p->B = g.B; p->H = g.H; p->W = g.W; p->Cp = g.Cp;
    p->TW = g.W <= 8 ? 8 : 16; p->TH = BM / p->TW;
    p->tiles_x = ceil_div(g.W, p->TW); p->tiles_y = ceil_div(g.H, p->TH);
    p->N = N; p->K = 9 * g.Cp; p->BN = pick_bn(N); p->epi = epi;
    p->M = g.B * p->tiles_x * p->tiles_y * BM;
    p->kblocks = 9 * (g.Cp / BK);
    if ((rc = encode_nhwc(&p->tmA, A, g, p->TW, p->TH))) return rc;
    if ((rc = encode_2d(&p->tmB, Bw, N, p->K, p->K, p->BN))) return rc;
    p->grid = dim3(ceil_div(N, p->BN), g.B * p->tiles_x * p->tiles_y);
    finish_plan(p);
    return D2S_OK;
}

int gemm_launch(const GemmPlan *p, cudaStream_t stream) {
    GemmArgs a;
    a.tmA = p->tmA; a.tmB = p->tmB;
    a.M = p->M; a.N = p->N; a.kblocks = p->kblocks; a.stages = p->stages;
    a.conv = p->conv; a.H = p->H; a.W = p->W; a.Cp = p->Cp; a.TW = p->TW; a.TW_shift = p->TW == 8 ? 3 : 4;
    a.tiles_x = p->tiles_x; a.tiles_y = p->tiles_y;
    a.epi = p->epi;
    switch (p->BN) {
        case 32: D2S_LAUNCH(gemm_tc_kernel<32>, p->grid, kGemmThreads, p->smem, stream, a); break;
        case 64: D2S_LAUNCH(gemm_tc_kernel<64>, p->grid, kGemmThreads, p->smem, stream, a); break;
        default: D2S_LAUNCH(gemm_tc_kernel<128>, p->grid, kGemmThreads, p->smem, stream, a); break;
    }
    D2S_POST_LAUNCH();
    return D2S_OK;
}

}  // namespace d2s
