// Warp-specialised tcgen05 GEMM for sm_100a:  TMA -> 128B-swizzled smem ring -> tcgen05.mma (fp32 accum in TMEM)
// -> tcgen05.ld epilogue with fused bias / GELU / ReLU / residual adds / fp32 residual-stream update / 1x1 head.
//
// Roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..5 = epilogue (each owns the 32 TMEM lanes its warp-id % 4 selects).
// One 128 x BN output tile per CTA; the ring depth is chosen so two CTAs fit per SM, which overlaps one CTA's
// epilogue with the other's main loop.
#include <cstdlib>
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
    int conv, H, W, Cp, TW, TW_shift, tiles_x, tiles_y, mtiles, ntiles;
    int splits, kb_per_split;   // split-K: blockIdx.z owns k-blocks [z*kb_per_split, ...)
    long long *trace;           // optional: clock64 stamps of CTA (0,0,0) phases (debug)
    GemmEpi epi;
};

// erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far below the fp16 rounding of the value it feeds): one MUFU.RCP, one
// MUFU.EX2 and six FMAs instead of erff()'s ~30 branchy instructions — the GELU epilogue of fc1 was epilogue-bound
// (profiles/r1_launches_large4k.txt: 100 us with GELU vs 47 us without at M = 6224, N = 4096).
__device__ __forceinline__ float erf_fast(float x) {
    const float ax = fabsf(x);
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    return copysignf(fmaf(-p * t, __expf(-ax * ax), 1.f), x);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erf_fast(x * 0.70710678118654752440f)); }

template <int ACT>
__device__ __forceinline__ float apply_act(float v) {
    if (ACT == ACT_GELU) return gelu_erf(v);
    if (ACT == ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
    return v;
}

// Epilogue flavours (compile-time, so that each instantiation's code stays small enough for the instruction cache: the
// first version carried every flavour behind runtime flags, 88 KB of SASS, and spent ~14k cycles per tile stalled on
// instruction fetch — profiles/r1_gemm_trace.txt).
enum { MODE_C16 = 0,    // fp16 store of act(acc + bias) (+ fp16 residuals, + optional relu copy)
       MODE_X32 = 1,    // fp32 residual stream += acc + bias
       MODE_HEAD = 2 }; // relu(acc + bias) . w3 + b3 -> final activation -> depth (N <= 32)

// The fused epilogue on 8 consecutive output columns [n, n+8) of one output row (off = row * ldc + n).
// sb: this tile's bias staged in shared memory (zeros if the GEMM has none); r1/r2: the fp16 residual operands, already
// loaded by the caller (so that their global-memory round trips overlap instead of forming a chain).
template <int ACT, int MODE>
__device__ __forceinline__ void epilogue8(const GemmEpi &e, float (&v)[8], const float *sb, uint4 r1, uint4 r2, long long off, int n,
                                          float &head_acc, long long orow = 0) {
    {
        float4 b0 = *(const float4 *)sb, b1 = *(const float4 *)(sb + 4);
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (ACT != ACT_NONE) {
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = apply_act<ACT>(v[t]);
    }
    if (MODE == MODE_HEAD) {
#pragma unroll
        for (int t = 0; t < 8; ++t) head_acc = fmaf(v[t], __ldg(e.w3 + n + t), head_acc);
    } else if (MODE == MODE_X32) {
        // fire-and-forget reductions into the fp32 stream (no load round trip; one add per element and GEMM when splits == 1)
        float4 *p = (float4 *)(e.x32 + off);
        if (e.x32_assign) {
            p[0] = make_float4(v[0], v[1], v[2], v[3]);
            p[1] = make_float4(v[4], v[5], v[6], v[7]);
        } else {
            atomicAdd(p, make_float4(v[0], v[1], v[2], v[3]));
            atomicAdd(p + 1, make_float4(v[4], v[5], v[6], v[7]));
        }
    } else {
        if (e.res1) {
            const __half2 *h = (const __half2 *)&r1;
#pragma unroll
            for (int t = 0; t < 4; ++t) { float2 f = __half22float2(h[t]); v[2 * t] += f.x; v[2 * t + 1] += f.y; }
        }
        if (e.res2) {
            const __half2 *h = (const __half2 *)&r2;
#pragma unroll
            for (int t = 0; t < 4; ++t) { float2 f = __half22float2(h[t]); v[2 * t] += f.x; v[2 * t + 1] += f.y; }
        }
        __half2 h[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) h[t] = __floats2half2_rn(v[2 * t], v[2 * t + 1]);
        *(uint4 *)(e.c16 + off) = *(const uint4 *)h;
        if (e.vt && n >= e.vt_col0) {   // V^T for the tcgen05 attention: adjacent lanes are adjacent tokens, so each store instruction is contiguous
            const int img = (int)(orow / e.vt_tokens), tok = (int)(orow - (long long)img * e.vt_tokens), cv = n - e.vt_col0;
            __half *dst = e.vt + ((size_t)(img * e.vt_heads + (cv >> 6)) * 64 + (cv & 63)) * e.vt_npad + tok;
            const __half *hv = (const __half *)h;
#pragma unroll
            for (int t = 0; t < 8; ++t) dst[(size_t)t * e.vt_npad] = hv[t];
        }
        if (e.c16_relu) {
#pragma unroll
            for (int t = 0; t < 4; ++t) h[t] = __hmax2(h[t], __float2half2_rn(0.f));
            *(uint4 *)(e.c16_relu + off) = *(const uint4 *)h;
        }
    }
}

// PAIR = 1: two CTAs adjacent in M (grid x = M tiles, cluster dims (2,1,1)) compute a 256 x BN tile with cta_group::2 MMAs.  Each CTA keeps its own
// 128 rows of A and HALF of the B tile (BN/2 rows) in shared memory and its own 128 x BN accumulator in TMEM, so a k-block costs a
// CTA 16 KB (A) + BN/2 x 128 B (B) of L2->SM traffic instead of 16 KB + BN x 128 B: 128 FLOP per ingested byte at BN = 256.
// Both CTAs run a producer (their loads complete on the leader's mbarrier), the leader (even CTA) issues the MMAs and multicasts
// the "slot free" / "accumulator ready" commits to both, both run the epilogue on their own rows.
template <int BN, int ACT, int MODE, int PAIR = 0>
__global__ void __launch_bounds__(kGemmThreads) gemm_tc_kernel(const __grid_constant__ GemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int kBRows = PAIR ? BN / 2 : BN;        // rows of the B tile held by this CTA
    constexpr int kBBytes = kBRows * BK * 2;
    constexpr int kStageBytes = kABytes + kBBytes;
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B needs 1024-B alignment
    const int stages = g.stages;
    uint64_t *full = (uint64_t *)(smem + (size_t)stages * kStageBytes);
    uint64_t *empty = full + stages;
    uint64_t *tmem_full = empty + stages;
    uint32_t *tmem_slot = (uint32_t *)(tmem_full + 1);
    float *s_bias = (float *)(((uintptr_t)(tmem_slot + 2) + 15) & ~(uintptr_t)15);  // this tile's bias (BN floats), staged by the epilogue warps during the main loop

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_n = PAIR ? blockIdx.y : blockIdx.x, tile_m = PAIR ? blockIdx.x : blockIdx.y;   // a pair = two CTAs adjacent in x
    const bool tr = g.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#define D2S_STAMP(i) do { if (tr) g.trace[i] = clock64(); } while (0)
    if (threadIdx.x == 0) D2S_STAMP(0);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&g.tmA);
        ptx::prefetch_tensormap(&g.tmB);
        for (int s = 0; s < stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
        ptx::mbar_init(tmem_full, 1);
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    if (warp == 1) {
        if (PAIR) { ptx::tmem_alloc_2sm(tmem_slot, BN); ptx::tmem_relinquish_2sm(); }
        else { ptx::tmem_alloc(tmem_slot, BN); ptx::tmem_relinquish(); }
    }
    ptx::tc_fence_before();
    if (PAIR) ptx::cluster_sync();   // the peer's barriers must be initialised before anything of ours can arrive on them
    else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;   // 0 = leader of the pair
    pdl_wait();   // everything above (barriers, TMEM, tensor-map prefetch) overlapped the previous kernel; operands are read below
    if (threadIdx.x == 0) D2S_STAMP(1);

    // conv tile -> (image, y0, x0)
    int img = 0, y0 = 0, x0 = 0;
    if (g.conv) {
        int per_img = g.tiles_x * g.tiles_y;
        img = tile_m / per_img;
        int t = tile_m - img * per_img;
        y0 = (t / g.tiles_x) * (BM >> g.TW_shift);
        x0 = (t % g.tiles_x) * g.TW;
    }

    // split-K range of this CTA
    const int kb0 = blockIdx.z * g.kb_per_split;
    const int nkb = min(kb0 + g.kb_per_split, g.kblocks) - kb0;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            const int cchunks = g.conv ? (g.Cp / BK) : 1;
            int tap = kb0 / cchunks, cc = kb0 - tap * cchunks;   // conv: k-block -> (filter tap, channel chunk)
            int s = 0;
            uint32_t ph = 0;
            uint8_t *a = smem;
            for (int i = 0; i < nkb; ++i) {
                if (i >= stages) ptx::mbar_wait(&empty[s], ph ^ 1);   // the first pass over the ring finds every slot free
                if (PAIR) {
                    if (rank == 0) ptx::mbar_arrive_expect_tx(&full[s], 2 * kStageBytes);   // both CTAs' loads land on the leader's barrier
                    if (g.conv) {
                        int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
                        ptx::tma_load_4d_2sm(a, &g.tmA, &full[s], cc * BK, x0 + dx, y0 + dy, img);
                        if (++cc == cchunks) { cc = 0; ++tap; }
                    } else {
                        ptx::tma_load_2d_2sm(a, &g.tmA, &full[s], (kb0 + i) * BK, tile_m * BM);
                    }
                    ptx::tma_load_2d_2sm(a + kABytes, &g.tmB, &full[s], (kb0 + i) * BK, tile_n * BN + (int)rank * kBRows);
                } else {
                    ptx::mbar_arrive_expect_tx(&full[s], kStageBytes);
                    if (g.conv) {
                        int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
                        ptx::tma_load_4d(a, &g.tmA, &full[s], cc * BK, x0 + dx, y0 + dy, img);
                        if (++cc == cchunks) { cc = 0; ++tap; }
                    } else {
                        ptx::tma_load_2d(a, &g.tmA, &full[s], (kb0 + i) * BK, tile_m * BM);
                    }
                    ptx::tma_load_2d(a + kABytes, &g.tmB, &full[s], (kb0 + i) * BK, tile_n * BN);
                }
                if (i == 0) D2S_STAMP(2);
                a += kStageBytes;
                if (++s == stages) { s = 0; ph ^= 1; a = smem; }
            }
            D2S_STAMP(3);
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // ===== MMA issuer (the leader CTA of a pair issues for both) =====
            const uint32_t idesc = ptx::make_idesc_f16(PAIR ? 2 * BM : BM, BN, 0);
            int s = 0;
            uint32_t ph = 0;
            uint32_t a_addr = ptx::smem_u32(smem);
            for (int i = 0; i < nkb; ++i) {
                ptx::mbar_wait(&full[s], ph);
                ptx::tc_fence_after();
                if (i == 0) D2S_STAMP(4);
                const uint64_t da = ptx::make_sw128_kmajor_desc(a_addr);
                const uint64_t db = ptx::make_sw128_kmajor_desc(a_addr + kABytes);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {  // UMMA_K = 16 fp16 = 32 bytes: advance the start address by 2 (>>4)
                    if (PAIR) ptx::umma_f16_2sm(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (i | k) != 0);
                    else ptx::umma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (i | k) != 0);
                }
                if (PAIR) ptx::umma_commit_2sm(&empty[s], 3);   // slot s is free in BOTH CTAs once these MMAs have read it
                else ptx::umma_commit(&empty[s]);             // smem slot is free once these MMAs have read it
                a_addr += kStageBytes;
                if (++s == stages) { s = 0; ph ^= 1; a_addr = ptx::smem_u32(smem); }
            }
            if (PAIR) ptx::umma_commit_2sm(tmem_full, 3);   // accumulators complete in both CTAs
            else ptx::umma_commit(tmem_full);             // accumulator complete
            D2S_STAMP(5);
        }
    } else {
        // ===== epilogue: TMEM -> registers -> global =====
        const GemmEpi &e = g.epi;
        const int q = warp & 3;               // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;          // row inside the tile
        long long orow;                        // output row (pixel / token) or -1
        if (g.conv) {
            int y = y0 + (r >> g.TW_shift), x = x0 + (r & (g.TW - 1));
            orow = (y < g.H && x < g.W && tile_m < g.mtiles) ? ((long long)img * g.H + y) * g.W + x : -1;   // (pairs pad the grid to an even number of M tiles)
        } else {
            int m = tile_m * BM + r;
            orow = m < g.M ? m : -1;
        }
        // stage the bias while the main loop runs (the epilogue warps have nothing else to do yet)
        {
            for (int t = threadIdx.x - 64; t < BN; t += 128) {
                const int n = tile_n * BN + t;
                s_bias[t] = (e.bias && n < g.N) ? __ldg(e.bias + n) : 0.f;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        const bool fixup = g.splits > 1;   // split-K: this CTA only parks its partial sums; see step 2 below
        const bool has_res = MODE == MODE_C16 && (e.res1 || e.res2);
        const uint4 z4 = make_uint4(0, 0, 0, 0);
        ptx::mbar_wait(tmem_full, 0);
        ptx::tc_fence_after();
        pdl_launch_dependents();   // the main loop is over: the next kernel's CTAs may move in and run their prologue under this epilogue
        if (threadIdx.x == 64) D2S_STAMP(6);
        float head_acc = 0.f;
        if (!fixup) {
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t raw[32];
                ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), raw);
                const int n0 = tile_n * BN + c * 32;
                const bool live = orow >= 0 && n0 < g.N;
                const long long off0 = orow * e.ldc + n0;
                uint4 r1[4] = {z4, z4, z4, z4}, r2[4] = {z4, z4, z4, z4};
                if (has_res && live) {   // all residual loads of this chunk in flight together, under the TMEM load
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n0 + j * 8 < g.N) {
                            if (e.res1) r1[j] = __ldg((const uint4 *)(e.res1 + off0 + j * 8));
                            if (e.res2) r2[j] = __ldg((const uint4 *)(e.res2 + off0 + j * 8));
                        }
                }
                ptx::tmem_ld_wait();
                if (!live) continue;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (n0 + j * 8 >= g.N) break;      // N is a multiple of 8
                    float v[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(raw[j * 8 + t]);
                    epilogue8<ACT, MODE>(e, v, s_bias + c * 32 + j * 8, r1[j], r2[j], off0 + j * 8, n0 + j * 8, head_acc, orow);
                }
            }
        } else {
            // split-K, step 1: park this split's fp32 partial tile in this CTA's shared memory (the operand ring is idle now:
            // every TMA load has landed and every MMA has retired).  Layout [8-column unit][half][row] of float4, so a warp's
            // 16-byte accesses are contiguous both here and when a peer CTA reads them through DSMEM.
            float4 *part = (float4 *)smem;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t raw[32];
                ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), raw);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    part[(c * 8 + j) * BM + r] = make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                                                             __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]));
            }
        }
        if (MODE == MODE_HEAD && orow >= 0 && tile_n == 0) {   // (never split: the plan keeps the fused head in one CTA)
            float d = (e.final_act == ACT_SIGMOID ? apply_act<ACT_SIGMOID>(head_acc + e.b3) : fmaxf(head_acc + e.b3, 0.f)) * e.max_depth;
            if (e.depth_dtype == D2S_F16) ((__half *)e.depth_out)[orow] = __float2half_rn(d);
            else ((float *)e.depth_out)[orow] = d;
        }
        if (threadIdx.x == 64) D2S_STAMP(7);
        ptx::tc_fence_before();
    }
    if (g.splits > 1) {
        // split-K, step 2: the CTAs of one output tile form a thread-block cluster (cluster dims (1,1,splits)).  After the
        // cluster barrier every partial tile is visible through distributed shared memory; CTA z reduces the 8-column units
        // u = z, z + splits, ... by adding the partials in split order 0..splits-1 (the same order whichever CTA does it, so
        // replays are bit-identical) and runs the fused epilogue on them.  No global scratch, no atomics on partial sums.
        ptx::cluster_sync();
        if (warp >= 2) {
            const GemmEpi &e = g.epi;
            const int r = (warp & 3) * 32 + lane;
            long long orow;
            if (g.conv) {
                int y = y0 + (r >> g.TW_shift), x = x0 + (r & (g.TW - 1));
                orow = (y < g.H && x < g.W) ? ((long long)img * g.H + y) * g.W + x : -1;
            } else {
                int m = tile_m * BM + r;
                orow = m < g.M ? m : -1;
            }
            const bool has_res = MODE == MODE_C16 && (e.res1 || e.res2);
            const uint32_t part0 = ptx::smem_u32(smem) + (uint32_t)r * 16u;
            float head_acc = 0.f;
#pragma unroll 1
            for (int u = blockIdx.z; u < BN / 8; u += g.splits) {
                const int n = tile_n * BN + u * 8;
                const bool live = orow >= 0 && n < g.N;
                const long long off = orow * e.ldc + n;
                uint4 r1 = make_uint4(0, 0, 0, 0), r2 = r1;
                if (has_res && live) {
                    if (e.res1) r1 = __ldg((const uint4 *)(e.res1 + off));
                    if (e.res2) r2 = __ldg((const uint4 *)(e.res2 + off));
                }
                float v[8];
#pragma unroll 1
                for (int z = 0; z < g.splits; ++z) {
                    const uint32_t ra = ptx::mapa(part0 + (uint32_t)(u * 2) * (BM * 16u), (uint32_t)z);
                    const float4 a = ptx::ld_dsmem_f4(ra), b = ptx::ld_dsmem_f4(ra + BM * 16u);
                    if (z == 0) { v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; }
                    else { v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w; }
                }
                if (live) epilogue8<ACT, MODE>(e, v, s_bias + u * 8, r1, r2, off, n, head_acc, orow);
            }
        }
        ptx::cluster_sync();   // nobody leaves (and frees its shared memory) while a peer may still be reading it
    }
    if (PAIR) ptx::cluster_sync();   // the leader's MMAs read the peer's shared memory: nobody leaves before both are done
    else __syncthreads();
    if (threadIdx.x == 0) D2S_STAMP(8);
    if (warp == 1) {
        ptx::tc_fence_after();
        if (PAIR) ptx::tmem_dealloc_2sm(tmem_base, BN);
        else ptx::tmem_dealloc(tmem_base, BN);
    }
    if (threadIdx.x == 32) D2S_STAMP(9);
#undef D2S_STAMP
}

// ------------------------------------------------------------------------------------------------
// Persistent variant for problems with more tiles than CTA slots: one CTA (or cta_group::2 pair) per SM loops over output
// tiles; the operand ring keeps streaming across tile boundaries and the accumulator is DOUBLE-BUFFERED in TMEM (2 x BN
// columns), so the epilogue of tile i (TMEM -> registers -> fused epilogue -> global) overlaps the main loop of tile i+1.
//   smem full/empty[stages]   TMA  <-> MMA
//   tmem_full/tmem_empty[2]   MMA  <-> epilogue warps   (pair: the peer's epilogue warps arrive on the leader's tmem_empty)
// Tiles are walked M-fastest, so the CTAs running at the same time share one B (weight) tile in L2.
// ------------------------------------------------------------------------------------------------
constexpr int kPersistThreads = 320;   // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter, each takes half of the columns)
template <int BN, int ACT, int MODE, int PAIR>
__global__ void __launch_bounds__(kPersistThreads) gemm_tc_persistent_kernel(const __grid_constant__ GemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int kBRows = PAIR ? BN / 2 : BN;
    constexpr int kBBytes = kBRows * BK * 2;
    constexpr int kStageBytes = kABytes + kBBytes;
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int stages = g.stages;
    uint64_t *full = (uint64_t *)(smem + (size_t)stages * kStageBytes);
    uint64_t *empty = full + stages;
    uint64_t *tmem_full = empty + stages;      // [2]
    uint64_t *tmem_empty = tmem_full + 2;      // [2]
    uint32_t *tmem_slot = (uint32_t *)(tmem_empty + 2);
    float *s_bias = (float *)(((uintptr_t)(tmem_slot + 2) + 15) & ~(uintptr_t)15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&g.tmA);
        ptx::prefetch_tensormap(&g.tmB);
        for (int s = 0; s < stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], PAIR ? 512 : 256); }
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    if (warp == 1) {
        if (PAIR) { ptx::tmem_alloc_2sm(tmem_slot, 2 * BN); ptx::tmem_relinquish_2sm(); }
        else { ptx::tmem_alloc(tmem_slot, 2 * BN); ptx::tmem_relinquish(); }
    }
    ptx::tc_fence_before();
    if (PAIR) ptx::cluster_sync(); else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;
    pdl_sync();

    // tile walk: unit = one CTA tile, or one pair tile (two M tiles); M-fastest
    const int units_m = PAIR ? (g.mtiles + 1) / 2 : g.mtiles;
    const int n_units = units_m * g.ntiles;
    const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int unit_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int nkb = g.kblocks;
    const int cchunks = g.conv ? (g.Cp / BK) : 1;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int s = 0; uint32_t ph = 0; long long issued = 0;
            uint8_t *a = smem;
            for (int u = unit0; u < n_units; u += unit_step) {
                const int tile_n = u / units_m, tile_m = (u - tile_n * units_m) * (PAIR ? 2 : 1) + (int)rank;
                int img = 0, y0 = 0, x0 = 0;
                if (g.conv) {
                    int per_img = g.tiles_x * g.tiles_y;
                    img = tile_m / per_img;
                    int t = tile_m - img * per_img;
                    y0 = (t / g.tiles_x) * (BM >> g.TW_shift);
                    x0 = (t % g.tiles_x) * g.TW;
                }
                int tap = 0, cc = 0;
                for (int i = 0; i < nkb; ++i, ++issued) {
                    if (issued >= stages) ptx::mbar_wait(&empty[s], ph ^ 1);
                    if (PAIR) {
                        if (rank == 0) ptx::mbar_arrive_expect_tx(&full[s], 2 * kStageBytes);
                        if (g.conv) {
                            int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
                            ptx::tma_load_4d_2sm(a, &g.tmA, &full[s], cc * BK, x0 + dx, y0 + dy, img);
                            if (++cc == cchunks) { cc = 0; ++tap; }
                        } else {
                            ptx::tma_load_2d_2sm(a, &g.tmA, &full[s], i * BK, tile_m * BM);
                        }
                        ptx::tma_load_2d_2sm(a + kABytes, &g.tmB, &full[s], i * BK, tile_n * BN + (int)rank * kBRows);
                    } else {
                        ptx::mbar_arrive_expect_tx(&full[s], kStageBytes);
                        if (g.conv) {
                            int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
                            ptx::tma_load_4d(a, &g.tmA, &full[s], cc * BK, x0 + dx, y0 + dy, img);
                            if (++cc == cchunks) { cc = 0; ++tap; }
                        } else {
                            ptx::tma_load_2d(a, &g.tmA, &full[s], i * BK, tile_m * BM);
                        }
                        ptx::tma_load_2d(a + kABytes, &g.tmB, &full[s], i * BK, tile_n * BN);
                    }
                    a += kStageBytes;
                    if (++s == stages) { s = 0; ph ^= 1; a = smem; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc = ptx::make_idesc_f16(PAIR ? 2 * BM : BM, BN, 0);
            int s = 0; uint32_t ph = 0;
            uint32_t a_addr = ptx::smem_u32(smem);
            int it = 0;
            for (int u = unit0; u < n_units; u += unit_step, ++it) {
                const int acc = it & 1;
                if (it >= 2) {                                        // the epilogue must have drained this accumulator (tile it-2)
                    ptx::mbar_wait(&tmem_empty[acc], (uint32_t)(((it >> 1) - 1) & 1));
                    ptx::tc_fence_after();
                }
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int i = 0; i < nkb; ++i) {
                    ptx::mbar_wait(&full[s], ph);
                    ptx::tc_fence_after();
                    const uint64_t da = ptx::make_sw128_kmajor_desc(a_addr);
                    const uint64_t db = ptx::make_sw128_kmajor_desc(a_addr + kABytes);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        if (PAIR) ptx::umma_f16_2sm(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (i | k) != 0);
                        else ptx::umma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (i | k) != 0);
                    }
                    if (PAIR) ptx::umma_commit_2sm(&empty[s], 3); else ptx::umma_commit(&empty[s]);
                    a_addr += kStageBytes;
                    if (++s == stages) { s = 0; ph ^= 1; a_addr = ptx::smem_u32(smem); }
                }
                if (PAIR) ptx::umma_commit_2sm(&tmem_full[acc], 3); else ptx::umma_commit(&tmem_full[acc]);
            }
        }
    } else {
        // ===== epilogue warps =====
        const GemmEpi &e = g.epi;
        const int q = warp & 3, r = q * 32 + lane;   // TMEM lane quarter this warp may access = warp id % 4
        const int chalf = (warp - 2) >> 2;           // which half of the tile's columns this warp handles
        constexpr int kChunks = BN / 64;             // 32-column chunks per warp
        const bool has_res = MODE == MODE_C16 && (e.res1 || e.res2);
        const uint4 z4 = make_uint4(0, 0, 0, 0);
        int it = 0;
        for (int u = unit0; u < n_units; u += unit_step, ++it) {
            const int acc = it & 1;
            const int tile_n = u / units_m, tile_m = (u - tile_n * units_m) * (PAIR ? 2 : 1) + (int)rank;
            long long orow;
            if (g.conv) {
                int per_img = g.tiles_x * g.tiles_y;
                int img = tile_m / per_img;
                int t = tile_m - img * per_img;
                int y = (t / g.tiles_x) * (BM >> g.TW_shift) + (r >> g.TW_shift), x = (t % g.tiles_x) * g.TW + (r & (g.TW - 1));
                orow = (y < g.H && x < g.W && tile_m < g.mtiles) ? ((long long)img * g.H + y) * g.W + x : -1;
            } else {
                int m = tile_m * BM + r;
                orow = m < g.M ? m : -1;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");          // everyone is done with the previous tile's bias
            for (int t = threadIdx.x - 64; t < BN; t += 256) {
                const int n = tile_n * BN + t;
                s_bias[t] = (e.bias && n < g.N) ? __ldg(e.bias + n) : 0.f;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            ptx::mbar_wait(&tmem_full[acc], (uint32_t)((it >> 1) & 1));
            ptx::tc_fence_after();
            float head_acc = 0.f;
#pragma unroll 1
            for (int c = chalf * kChunks; c < (chalf + 1) * kChunks; ++c) {
                uint32_t raw[32];
                ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), raw);
                const int n0 = tile_n * BN + c * 32;
                const bool live = orow >= 0 && n0 < g.N;
                const long long off0 = orow * e.ldc + n0;
                uint4 r1[4] = {z4, z4, z4, z4}, r2[4] = {z4, z4, z4, z4};
                if (has_res && live) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n0 + j * 8 < g.N) {
                            if (e.res1) r1[j] = __ldg((const uint4 *)(e.res1 + off0 + j * 8));
                            if (e.res2) r2[j] = __ldg((const uint4 *)(e.res2 + off0 + j * 8));
                        }
                }
                ptx::tmem_ld_wait();
                if (c == (chalf + 1) * kChunks - 1) {                  // this warp's share of the accumulator is in registers: hand it back
                    ptx::tc_fence_before();
                    if (PAIR) ptx::mbar_arrive_leader(&tmem_empty[acc]); else ptx::mbar_arrive(&tmem_empty[acc]);
                }
                if (!live) continue;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (n0 + j * 8 >= g.N) break;
                    float v[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(raw[j * 8 + t]);
                    epilogue8<ACT, MODE>(e, v, s_bias + c * 32 + j * 8, r1[j], r2[j], off0 + j * 8, n0 + j * 8, head_acc, orow);
                }
            }
        }
        ptx::tc_fence_before();
    }
    if (PAIR) ptx::cluster_sync(); else __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        if (PAIR) ptx::tmem_dealloc_2sm(tmem_base, 2 * BN); else ptx::tmem_dealloc(tmem_base, 2 * BN);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

// the instantiations the network needs: (BN, activation, epilogue mode)
typedef void (*GemmKernel)(const GemmArgs);
struct Variant { int bn, act, mode, pair, persist; GemmKernel fn; };
#define D2S_V(bn, act, mode) {bn, act, mode, 0, 0, gemm_tc_kernel<bn, act, mode, 0>}
#define D2S_P(bn, act, mode) {bn, act, mode, 1, 0, gemm_tc_kernel<bn, act, mode, 1>}
#define D2S_S(bn, act, mode, pair) {bn, act, mode, pair, 1, gemm_tc_persistent_kernel<bn, act, mode, pair>}
static const Variant kVariants[] = {
    D2S_S(256, ACT_NONE, MODE_C16, 0), D2S_S(256, ACT_RELU, MODE_C16, 0), D2S_S(256, ACT_GELU, MODE_C16, 0), D2S_S(256, ACT_NONE, MODE_X32, 0),
    D2S_S(256, ACT_NONE, MODE_C16, 1), D2S_S(256, ACT_RELU, MODE_C16, 1), D2S_S(256, ACT_GELU, MODE_C16, 1), D2S_S(256, ACT_NONE, MODE_X32, 1),

    D2S_P(256, ACT_NONE, MODE_C16), D2S_P(256, ACT_RELU, MODE_C16), D2S_P(256, ACT_GELU, MODE_C16), D2S_P(256, ACT_NONE, MODE_X32),
    D2S_V(256, ACT_NONE, MODE_C16), D2S_V(256, ACT_RELU, MODE_C16), D2S_V(256, ACT_GELU, MODE_C16), D2S_V(256, ACT_NONE, MODE_X32),
    D2S_V(128, ACT_NONE, MODE_C16), D2S_V(128, ACT_RELU, MODE_C16), D2S_V(128, ACT_GELU, MODE_C16), D2S_V(128, ACT_NONE, MODE_X32),
    D2S_V(64, ACT_NONE, MODE_C16),  D2S_V(64, ACT_RELU, MODE_C16),  D2S_V(64, ACT_GELU, MODE_C16),  D2S_V(64, ACT_NONE, MODE_X32),
    D2S_V(32, ACT_NONE, MODE_C16),  D2S_V(32, ACT_RELU, MODE_C16),  D2S_V(32, ACT_RELU, MODE_HEAD), D2S_V(32, ACT_NONE, MODE_X32),
};
#undef D2S_V
#undef D2S_P
#undef D2S_S
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);
static int epi_mode(const GemmEpi &e) { return e.w3 ? MODE_HEAD : (e.x32 ? MODE_X32 : MODE_C16); }
static std::once_flag g_once;
static int g_init_rc = D2S_OK;
constexpr size_t kMaxSmem = 227 * 1024;

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
        cudaError_t a = cudaSuccess;
        for (int i = 0; i < kNumVariants && a == cudaSuccess; ++i)
            a = cudaFuncSetAttribute((const void *)kVariants[i].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
        if (a != cudaSuccess) g_init_rc = set_error(D2S_ERR_CUDA, "cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(a));
    });
    return g_init_rc;
}

int tma_encode_2d(CUtensorMap *m, const void *base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows) {
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

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// Tile width.  The main loop is bound by what one SM can pull from L2 (the whole chip moves ~12 TB/s from L2, i.e. ~43 B/clk per
// SM when every SM streams): a 128 x 128 tile ingests 32 KB per 2.1 MFLOP k-block = 64 FLOP/B -> ~0.77 PFLOP/s chip-wide, which is
// what the kernel measures (0.80-0.86).  A 128 x 256 tile ingests 48 KB per 4.2 MFLOP = 85 FLOP/B.  It is used when the problem has
// enough 256-wide tiles to fill every SM (large batches; +16 % at M = 6224, profiles/r1_gemm_microbench.txt); small problems keep 128-wide tiles for parallelism (and split-K).
// Plan policy (set by the engine around plan construction): 0 = latency (one frame alone on the GPU: many small tiles, split-K),
// 1 = throughput (several frames in flight share the GPU: 256-wide tiles whenever N allows — fewer, fatter CTAs, 33 % less L2
// ingest per FLOP and half the per-CTA fixed cost; measured +14 % frames/s at 8 frames in flight for +0.6 ms single-frame latency).
static thread_local int g_plan_policy = 0;
void gemm_set_plan_policy(int policy) { g_plan_policy = policy; }

static int pick_bn(int N, int mtiles) {
    if (N <= 32) return 32;
    if (N <= 64) return 64;
    const int mode = env_int("D2S_GEMM_BN256", g_plan_policy == 1 ? 2 : 1);    // 0 never, 1 heuristic, 2 whenever N allows
    if (mode && N % 256 == 0 && (mode == 2 || (long long)mtiles * (N / 256) >= kNumSMs)) return 256;
    return 128;
}

// 2-CTA pairs (cta_group::2) for the 256-wide tiles.  D2S_GEMM_PAIR = 0 never, 2 whenever BN == 256, 1 (default): only for
// problems that run on the persistent kernel (more than two waves of tiles), and there only if the pair grid does not quantise
// worse: rounds = ceil(units / slots) with 148 single-CTA slots or 74 pair slots; a pair unit is two tiles in ~1.86x the time
// (profiles/r1_gemm_microbench.txt: M = 6224, N = 4096: 1.13 vs 1.05 PFLOP/s; N = 3072, where pairs need 5 rounds instead of 4:
// 1.00 vs 1.08; 8192^3: 1.43 vs 1.29).
static int use_pair(int BN, int mtiles, int ntiles) {
    const int mode = env_int("D2S_GEMM_PAIR", 1);
    if (BN != 256 || mtiles < 2 || mode == 0) return 0;
    if (mode == 2) return 1;
    const long long tiles = (long long)mtiles * ntiles;
    if (tiles <= 2 * kNumSMs || !env_int("D2S_GEMM_PERSIST", 1)) return 0;
    const long long r1 = (tiles + kNumSMs - 1) / kNumSMs, rp = ((long long)((mtiles + 1) / 2) * ntiles + kNumSMs / 2 - 1) / (kNumSMs / 2);
    return 93 * rp <= 100 * r1;
}

// Choose the split-K factor and the ring depth.  Most of this network's GEMMs are small (M = 778 tokens, or a few
// hundred pixels at the coarse DPT levels): with one 128 x BN tile per CTA the grid covers a fraction of the 148 SMs and
// the kernel time is a serial chain of k-blocks, each bounded by TMA latency / ring depth.  So: split K until the grid
// fills one wave, and make the ring as deep as the chain (one CTA per SM) unless the grid needs two CTAs per SM.
static void finish_plan(GemmPlan *p) {
    const int base = p->grid.x * p->grid.y;
    const GemmEpi &e = p->epi;
    const bool x32 = epi_mode(e) == MODE_X32;
    int splits = 1;
    const int max_splits = min(env_int("D2S_GEMM_MAX_SPLITS", 8), 8);   // the splits of a tile form one cluster: portable maximum 8
    if (epi_mode(e) == MODE_HEAD || p->BN == 256) {
        splits = 1;                                    // the fused 1x1 head reduces over all N columns inside one thread; 256-wide
                                                       // tiles are only chosen for problems that fill the GPU without splitting
    } else if (x32) {
        splits = kNumSMs / base;                       // bias-only epilogue into the fp32 stream: cheap fix-up
        if (splits > p->kblocks / 3) splits = p->kblocks / 3;
    } else if (base * 2 <= kNumSMs && p->kblocks >= 16) {
        splits = kNumSMs / base;                       // fix-up path: only when the chain is long and the grid small
        if (splits > p->kblocks / 4) splits = p->kblocks / 4;
    }
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p->kb_per_split = ceil_div(p->kblocks, splits);
    p->splits = ceil_div(p->kblocks, p->kb_per_split);   // no empty split
    p->grid.z = p->splits;

    const size_t stage = kABytes + (size_t)(p->pair ? p->BN / 2 : p->BN) * BK * 2;
    const size_t budget = (base * p->splits > kNumSMs) ? kMaxSmem / 2 : kMaxSmem;   // two CTAs per SM only if needed
    int stages = (int)((budget - 2048) / stage);
    if (stages > p->kb_per_split) stages = p->kb_per_split;
    if (stages < 2) stages = 2;
    if (stages > 12) stages = 12;
    // Ring depth cap (default 3): per-CTA TMA ingest, not ring depth, bounds the main loop (one SM takes ~64 B/clk from L2), and a
    // 96 KB ring lets two CTAs share an SM so that one CTA's prologue/epilogue overlaps another's main loop — measured +35 %
    // frames/s with several frames in flight, no change in single-frame latency (profiles/r1_sweep_ring_depth.txt).
    { int ms = env_int("D2S_GEMM_MAX_STAGES", 3); if (ms >= 2 && stages > ms) stages = ms; }
    if (p->BN == 256) {   // 2 x 48 KB (3 x 32 KB for a pair): two CTAs (2 x 256 TMEM columns) per SM
        stages = env_int("D2S_GEMM_BN256_STAGES", p->pair ? 3 : 2); if (stages < 2) stages = 2; if (stages > 6) stages = 6;
    }
    // Persistent tile loop (one CTA or pair per SM, double-buffered TMEM accumulator) when there are more tiles than one wave of
    // CTA slots.  D2S_GEMM_PERSIST = 0 never, 1 (default) heuristic.
    p->persist = 0;
    if (p->splits == 1 && p->BN == 256 && epi_mode(e) != MODE_HEAD && env_int("D2S_GEMM_PERSIST", 1)) {   // (128-wide tiles are ingest-bound: they need two CTAs per SM, not one deep ring)
        const long long tiles = (long long)p->mtiles * ceil_div(p->N, p->BN);
        if (tiles > 2 * kNumSMs) {
            p->persist = 1;
            stages = (int)((kMaxSmem - 4096 - (size_t)p->BN * 4) / stage);            // one CTA per SM: as deep a ring as fits
            { int ms = env_int("D2S_GEMM_PERSIST_STAGES", 0); if (ms >= 2 && stages > ms) stages = ms; }
            if (stages > 8) stages = 8;
            const int units = p->pair ? (p->mtiles + 1) / 2 * ceil_div(p->N, p->BN) : (int)tiles;
            const int slots = p->pair ? kNumSMs / 2 : kNumSMs;
            p->grid = dim3((unsigned)((units < slots ? units : slots) * (p->pair ? 2 : 1)), 1, 1);
        }
    }
    p->stages = stages;
    p->smem = (size_t)stages * stage + (2 * stages + 4) * 8 + 32 + (size_t)p->BN * 4 + 16 + 1024;
    if (env_int("D2S_VERBOSE", 0))
        fprintf(stderr, "[d2s gemm] %s M=%d N=%d K=%d BN=%d pair=%d persist=%d grid=(%u,%u,%u) kblocks=%d splits=%d stages=%d smem=%zu act=%d mode=%d\n", p->conv ? "conv" : "lin ",
                p->M, p->N, p->K, p->BN, p->pair, p->persist, p->grid.x, p->grid.y, p->grid.z, p->kblocks, p->splits, p->stages, p->smem, e.act, epi_mode(e));
}

static int check_epi(const GemmEpi &e, int N) {
    D2S_REQUIRE(N % 8 == 0, "gemm: N=%d must be a multiple of 8", N);
    D2S_REQUIRE(e.ldc % 8 == 0 || !(e.c16 || e.c16_relu || e.res1 || e.res2 || e.x32 || e.c32), "gemm: ldc=%d must be a multiple of 8", e.ldc);
    D2S_REQUIRE(!e.w3 || N <= 32, "gemm: fused 1x1 head needs N <= 32 (got %d)", N);
    D2S_REQUIRE(!e.x32 || !(e.c16 || e.c16_relu || e.res1 || e.res2 || e.w3 || e.act != ACT_NONE), "gemm: the fp32-stream epilogue takes bias only");
    D2S_REQUIRE(e.x32 || e.w3 || e.c16, "gemm: no output");
    D2S_REQUIRE(!e.c32, "gemm: fp32 plain-store epilogue was removed");
    return D2S_OK;
}

int gemm_plan_linear(GemmPlan *p, const __half *A, int lda, const __half *Bw, int ldb, int M, int N, int K, const GemmEpi &epi) {
    int rc = gemm_init();
    if (rc) return rc;
    if ((rc = check_epi(epi, N))) return rc;
    D2S_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
    *p = GemmPlan{};
    p->M = M; p->N = N; p->K = K; p->BN = pick_bn(N, ceil_div(M, BM)); p->conv = 0; p->epi = epi;
    p->kblocks = ceil_div(K, BK);
    p->mtiles = ceil_div(M, BM);
    p->pair = use_pair(p->BN, p->mtiles, ceil_div(N, p->BN));
    if ((rc = tma_encode_2d(&p->tmA, A, M, K, lda, BM))) return rc;
    if ((rc = tma_encode_2d(&p->tmB, Bw, N, K, ldb, p->pair ? p->BN / 2 : p->BN))) return rc;
    p->grid = p->pair ? dim3((p->mtiles + 1) / 2 * 2, ceil_div(N, p->BN)) : dim3(ceil_div(N, p->BN), p->mtiles);
    D2S_REQUIRE(p->grid.y <= 65535, "gemm: %d row tiles exceed the grid limit (M=%d): lower the batch", p->mtiles, M);
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
    p->N = N; p->K = 9 * g.Cp; p->BN = pick_bn(N, g.B * p->tiles_x * p->tiles_y); p->epi = epi;
    p->mtiles = g.B * p->tiles_x * p->tiles_y;
    p->M = p->mtiles * BM;
    p->kblocks = 9 * (g.Cp / BK);
    p->pair = use_pair(p->BN, p->mtiles, ceil_div(N, p->BN));
    if ((rc = encode_nhwc(&p->tmA, A, g, p->TW, p->TH))) return rc;
    if ((rc = tma_encode_2d(&p->tmB, Bw, N, p->K, p->K, p->pair ? p->BN / 2 : p->BN))) return rc;
    p->grid = p->pair ? dim3((p->mtiles + 1) / 2 * 2, ceil_div(N, p->BN)) : dim3(ceil_div(N, p->BN), p->mtiles);
    D2S_REQUIRE(p->grid.y <= 65535, "conv: %d pixel tiles exceed the grid limit (%dx%dx%d): lower the batch", p->mtiles, g.B, g.H, g.W);
    finish_plan(p);
    return D2S_OK;
}

int gemm_launch(const GemmPlan *p, cudaStream_t stream) {
    GemmArgs a;
    a.tmA = p->tmA; a.tmB = p->tmB;
    a.M = p->M; a.N = p->N; a.kblocks = p->kblocks; a.stages = p->stages;
    a.conv = p->conv; a.H = p->H; a.W = p->W; a.Cp = p->Cp; a.TW = p->TW; a.TW_shift = p->TW == 8 ? 3 : 4;
    a.tiles_x = p->tiles_x; a.tiles_y = p->tiles_y; a.mtiles = p->mtiles; a.ntiles = ceil_div(p->N, p->BN);
    a.trace = p->trace;
    a.splits = p->splits; a.kb_per_split = p->kb_per_split;
    a.epi = p->epi;
    const int mode = epi_mode(p->epi);
    GemmKernel fn = nullptr;
    for (int i = 0; i < kNumVariants; ++i)
        if (kVariants[i].bn == p->BN && kVariants[i].act == p->epi.act && kVariants[i].mode == mode && kVariants[i].pair == p->pair && kVariants[i].persist == p->persist) fn = kVariants[i].fn;
    if (!fn) return set_error(D2S_ERR_UNSUPPORTED, "gemm: no kernel variant for BN=%d act=%d mode=%d", p->BN, p->epi.act, mode);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = p->grid; cfg.blockDim = dim3(p->persist ? kPersistThreads : kGemmThreads); cfg.dynamicSmemBytes = p->smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL (common.cuh), appended below when enabled
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    attr[0].id = cudaLaunchAttributeClusterDimension;   // split-K: the CTAs of one output tile are one cluster
    attr[0].val.clusterDim.x = p->pair ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = (unsigned)p->splits;   // (a pair never splits K)
    const bool clustered = p->splits > 1 || p->pair;
    cfg.attrs = clustered ? attr : attr + 1; cfg.numAttrs = (clustered ? 1 : 0) + (g_pdl ? 1 : 0);
    cudaError_t le = cudaLaunchKernelEx(&cfg, fn, a);
    if (le != cudaSuccess) {
        int nc = -1;
        cudaError_t oe = cudaOccupancyMaxActiveClusters(&nc, fn, &cfg);
        cudaFuncAttributes fa = {};
        cudaFuncGetAttributes(&fa, (const void *)fn);
        cudaGetLastError();
        return set_error(D2S_ERR_CUDA, "gemm launch failed: %s (grid %u,%u,%u cluster %u,%u,%u smem %zu BN %d pair %d regs %d static smem %zu; occupancy query: %s, %d clusters)",
                         cudaGetErrorString(le), p->grid.x, p->grid.y, p->grid.z, attr[0].val.clusterDim.x, attr[0].val.clusterDim.y, attr[0].val.clusterDim.z,
                         p->smem, p->BN, p->pair, fa.numRegs, fa.sharedSizeBytes, cudaGetErrorString(oe), nc);
    }
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return D2S_OK;
}

}  // namespace d2s
