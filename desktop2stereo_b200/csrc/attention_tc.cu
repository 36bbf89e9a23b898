// Fused multi-head self-attention on tcgen05 tensor cores (head dim 64, non-causal): softmax(Q K^T / 8) V.
//
// One CTA = 128 query rows of one (image, head); it walks the keys in blocks of 128.
//   warp 0      TMA producer: Q once, then K_j [128 keys x 64] and V^T_j [64 x 128 keys] per block (single-buffered: K_j is free
//               as soon as S_j = Q K_j^T has been issued, V_j as soon as P_j V_j has, so the next loads run under the softmax)
//   warp 1      MMA issuer + TMEM owner:  S_j = Q K_j^T  (128x128x64, fp32 in TMEM),  PV_j = P_j V_j  (128x64x128, fp32 in TMEM)
//   warps 2..9  softmax: TWO threads per query row (keys 0..63 | 64..127 of the block, output dims 0..31 | 32..63), so that every
//               scheduler has two softmax warps per CTA to interleave (v1 had one thread per row and was bound by single-warp
//               issue latency: 3.5 us per block).  Two passes over S_j in TMEM (row max — exchanged between the two threads
//               through shared memory —, then exp2 / row sum), P_j written as fp16 into shared memory in the 128B-swizzled
//               K-major layout the MMA reads; the running output O stays in REGISTERS (32 fp32 per thread):
//               O = O * exp2(m_old - m_new) + PV_j, read back from TMEM one block later.
// TMEM: 128 columns S + 64 columns PV (256 allocated) and 80 KB of shared memory per CTA, so two CTAs share an SM and one CTA's
// softmax overlaps the other's MMAs.  V^T (keys contiguous) is produced by a small transpose kernel so that both MMAs read plain
// K-major operands.  Replaces HF Dinov2SelfAttention's SDPA call (HF modeling_dinov2.py:203-234).
#include <type_traits>

#include "gemm.cuh"
#include "layers.cuh"
#include "ptx.cuh"

namespace d2s {

constexpr int kTcThreads = 320;

struct AttnArgs {
    CUtensorMap tmQK;     // qkv16 [B*N, 3D], boxes 128 rows x 64 cols
    CUtensorMap tmV;      // vt16 [B*heads*64, Npad], boxes 64 rows x 64 cols
    __half *out;          // [B*N, D]
    int N, D, heads, nkb;
    float scale_log2e;
};

// V third of qkv16 [B*N, 3D] -> vt16 [B, heads, 64, Npad] (keys contiguous); columns N..Npad-1 stay zero
__global__ void v_transpose_kernel(const __half *__restrict__ qkv, __half *__restrict__ vt, int N, int D, int heads, int Npad) {
    pdl_sync();
    __shared__ __half tile[64][66];
    const int b = blockIdx.z, h = blockIdx.y, t0 = blockIdx.x * 64;
    const __half *src = qkv + (size_t)b * N * 3 * D + 2 * D + h * 64;
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
        const int t = i >> 6, d = i & 63;
        tile[t][d] = (t0 + t < N) ? src[(size_t)(t0 + t) * 3 * D + d] : __float2half(0.f);
    }
    __syncthreads();
    __half *dst = vt + ((size_t)(b * heads + h) * 64) * Npad;
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
        const int d = i >> 6, t = i & 63;
        if (t0 + t < N) dst[(size_t)d * Npad + t0 + t] = tile[t][d];
    }
}

__global__ void __launch_bounds__(kTcThreads, 2) attention_tc_kernel(const __grid_constant__ AttnArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sQ = smem, *sK = smem + 16384, *sV = smem + 32768, *sP = smem + 49152;   // 16 + 16 + 16 + 32 KB
    uint64_t *bars = (uint64_t *)(smem + 81920);
    uint64_t *q_full = bars, *k_full = bars + 1, *k_free = bars + 2, *v_full = bars + 3, *s_full = bars + 4, *s_free = bars + 5,
             *p_full = bars + 6, *pv_full = bars + 7, *pv_free = bars + 8;
    uint32_t *tmem_slot = (uint32_t *)(bars + 9);
    float *s_mx = (float *)(bars + 10);         // [2 block parities][2 halves][128 rows] partial row maxima
    float *s_l = s_mx + 512;                    // [2 halves][128 rows] partial row sums (end of kernel)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
    const int nkb = g.nkb;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&g.tmQK);
        ptx::prefetch_tensormap(&g.tmV);
        ptx::mbar_init(q_full, 1); ptx::mbar_init(k_full, 1); ptx::mbar_init(k_free, 1); ptx::mbar_init(v_full, 1);
        ptx::mbar_init(s_full, 1); ptx::mbar_init(s_free, 256); ptx::mbar_init(p_full, 256); ptx::mbar_init(pv_full, 1);
        ptx::mbar_init(pv_free, 256);
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
    }
    if (warp == 1) { ptx::tmem_alloc(tmem_slot, 256); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tS = tmem_base, tPV = tmem_base + 128;
    pdl_sync();   // barrier init / TMEM allocation above ran under the previous kernel's tail; qkv and V^T are read below

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            const int row_q = b * g.N + q0, row_k0 = b * g.N, row_v = (b * g.heads + h) * 64;
            ptx::mbar_arrive_expect_tx(q_full, 16384);
            ptx::tma_load_2d(sQ, &g.tmQK, q_full, h * 64, row_q);
            for (int j = 0; j < nkb; ++j) {
                if (j > 0) ptx::mbar_wait_backoff(k_free, (uint32_t)((j - 1) & 1));
                ptx::mbar_arrive_expect_tx(k_full, 16384);
                ptx::tma_load_2d(sK, &g.tmQK, k_full, g.D + h * 64, row_k0 + j * 128);
                if (j > 0) ptx::mbar_wait_backoff(pv_full, (uint32_t)((j - 1) & 1));
                ptx::mbar_arrive_expect_tx(v_full, 16384);
                ptx::tma_load_2d(sV, &g.tmV, v_full, j * 128, row_v);
                ptx::tma_load_2d(sV + 8192, &g.tmV, v_full, j * 128 + 64, row_v);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idS = ptx::make_idesc_f16(128, 128, 0), idPV = ptx::make_idesc_f16(128, 64, 0);
            const uint32_t aQ = ptx::smem_u32(sQ), aK = ptx::smem_u32(sK), aV = ptx::smem_u32(sV), aP = ptx::smem_u32(sP);
            ptx::mbar_wait_backoff(q_full, 0);
            for (int j = 0; j < nkb; ++j) {
                const uint32_t ph = (uint32_t)(j & 1), pph = (uint32_t)((j - 1) & 1);
                ptx::mbar_wait_backoff(k_full, ph);
                if (j > 0) ptx::mbar_wait_backoff(s_free, pph);
                ptx::tc_fence_after();
                const uint64_t dq = ptx::make_sw128_kmajor_desc(aQ), dk = ptx::make_sw128_kmajor_desc(aK);
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::umma_f16(tS, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idS, k != 0);
                ptx::umma_commit(s_full);
                ptx::umma_commit(k_free);
                ptx::mbar_wait_backoff(p_full, ph);
                ptx::mbar_wait_backoff(v_full, ph);
                if (j > 0) ptx::mbar_wait_backoff(pv_free, pph);
                ptx::tc_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint64_t dp = ptx::make_sw128_kmajor_desc(aP + c * 16384), dv = ptx::make_sw128_kmajor_desc(aV + c * 8192);
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::umma_f16(tPV, dp + (uint64_t)(k * 2), dv + (uint64_t)(k * 2), idPV, (c | k) != 0);
                }
                ptx::umma_commit(pv_full);
            }
        }
    } else {
        // ===== softmax warps: two threads per query row =====
        const int qd = warp & 3, r = qd * 32 + lane;     // TMEM lane quarter = warp id % 4
        const int half = (warp - 2) >> 2;                // 0: keys 0..63 / dims 0..31,  1: keys 64..127 / dims 32..63
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const uint32_t pRow = ptx::smem_u32(sP) + (uint32_t)half * 16384u + (uint32_t)r * 128u;
        const float sl2e = g.scale_log2e;
        float o[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0.f;
        float m_old = -INFINITY, l = 0.f, c_prev = 0.f;
        // one key block; RAGGED (keys beyond N to mask) is a compile-time flag so that the common blocks carry no per-element tests
        auto block = [&](int j, auto ragged_c) {
            constexpr bool RAGGED = decltype(ragged_c)::value;
            const uint32_t ph = (uint32_t)(j & 1), pph = (uint32_t)((j - 1) & 1);
            const int key0 = j * 128 + half * 64;
            ptx::mbar_wait(s_full, ph);
            ptx::tc_fence_after();
            // pass 1: row max over this thread's 64 keys, then over the row through shared memory
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t raw[32];
                ptx::tmem_ld_32x32(tS + lane_off + (uint32_t)(half * 64 + c * 32), raw);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int t = 0; t < 32; ++t)
                    if (!RAGGED || key0 + c * 32 + t < g.N) mx = fmaxf(mx, __uint_as_float(raw[t]));
            }
            s_mx[(j & 1) * 256 + half * 128 + r] = mx;
            asm volatile("bar.sync 2, 256;" ::: "memory");
            const float m_new = fmaxf(m_old, fmaxf(mx, s_mx[(j & 1) * 256 + (half ^ 1) * 128 + r]));
            const float cs = fast_exp2((m_old - m_new) * sl2e);   // first block: exp2(-inf) = 0
            const float ms = m_new * sl2e;
            // fold the previous block's P V into the running output (it was computed against m_old)
            if (j > 0) {
                ptx::mbar_wait(pv_full, pph);
                ptx::tc_fence_after();
                uint32_t raw[32];
                ptx::tmem_ld_32x32(tPV + lane_off + (uint32_t)(half * 32), raw);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int t = 0; t < 32; ++t) o[t] = fmaf(o[t], c_prev, __uint_as_float(raw[t]));
                ptx::tc_fence_before();
                ptx::mbar_arrive(pv_free);
            }
            // pass 2: P = exp2(s * scale - m), row sum, fp16 P into the swizzled K-major tile
            float rs = 0.f;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t raw[32];
                ptx::tmem_ld_32x32(tS + lane_off + (uint32_t)(half * 64 + c * 32), raw);
                ptx::tmem_ld_wait();
                if (c == 1) { ptx::tc_fence_before(); ptx::mbar_arrive(s_free); }   // S_j is in registers: the next Q K^T may overwrite it
                uint32_t pk[16];
#pragma unroll
                for (int t = 0; t < 32; t += 2) {
                    float p0 = fast_exp2(fmaf(__uint_as_float(raw[t]), sl2e, -ms)), p1 = fast_exp2(fmaf(__uint_as_float(raw[t + 1]), sl2e, -ms));
                    if (RAGGED) { if (key0 + c * 32 + t >= g.N) p0 = 0.f; if (key0 + c * 32 + t + 1 >= g.N) p1 = 0.f; }
                    rs += p0 + p1;
                    __half2 hh = __floats2half2_rn(p0, p1);
                    pk[t >> 1] = *(uint32_t *)&hh;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t u = (uint32_t)(c * 4 + i);                    // 16-byte unit inside the 128-byte row
                    ptx::st_shared_v4(pRow + ((u ^ (uint32_t)(r & 7)) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
                }
            }
            l = fmaf(l, cs, rs);
            m_old = m_new;
            c_prev = cs;
            ptx::fence_proxy_async();      // generic-proxy stores of P -> visible to the tensor core's async proxy
            ptx::mbar_arrive(p_full);
        };
        for (int j = 0; j < nkb - 1; ++j) block(j, std::false_type{});
        if (nkb * 128 == g.N) block(nkb - 1, std::false_type{});
        else block(nkb - 1, std::true_type{});
        // last block's P V, the other half's row sum, then normalise and store
        ptx::mbar_wait(pv_full, (uint32_t)((nkb - 1) & 1));
        ptx::tc_fence_after();
        {
            uint32_t raw[32];
            ptx::tmem_ld_32x32(tPV + lane_off + (uint32_t)(half * 32), raw);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 32; ++t) o[t] = fmaf(o[t], c_prev, __uint_as_float(raw[t]));
        }
        ptx::tc_fence_before();
        s_l[half * 128 + r] = l;
        asm volatile("bar.sync 2, 256;" ::: "memory");
        l += s_l[(half ^ 1) * 128 + r];
        const int q = q0 + r;
        if (q < g.N) {
            const float inv = 1.f / l;
            __half *dst = g.out + ((size_t)b * g.N + q) * g.D + h * 64 + half * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                __half2 h0 = __floats2half2_rn(o[i] * inv, o[i + 1] * inv), h1 = __floats2half2_rn(o[i + 2] * inv, o[i + 3] * inv);
                __half2 h2 = __floats2half2_rn(o[i + 4] * inv, o[i + 5] * inv), h3 = __floats2half2_rn(o[i + 6] * inv, o[i + 7] * inv);
                *(uint4 *)(dst + i) = make_uint4(*(uint32_t *)&h0, *(uint32_t *)&h1, *(uint32_t *)&h2, *(uint32_t *)&h3);
            }
        }
    }
    __syncthreads();
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 256); }
}

size_t attention_tc_vt_elems(int B, int N, int heads) { return (size_t)B * heads * 64 * (size_t)((N + 7) / 8 * 8); }

int attention_tc_plan(AttnTcPlan *p, const __half *qkv, __half *vt, __half *out, int B, int N, int D, int heads) {
    int rc = gemm_init();
    if (rc) return rc;
    D2S_REQUIRE(D == heads * 64, "attention: head dim must be 64 (D=%d heads=%d)", D, heads);
    const int Npad = (N + 7) / 8 * 8;
    p->qkv = qkv; p->vt = vt; p->out = out; p->B = B; p->N = N; p->D = D; p->heads = heads; p->Npad = Npad;
    if ((rc = tma_encode_2d(&p->tmQK, qkv, (uint64_t)B * N, (uint64_t)3 * D, (uint64_t)3 * D, 128))) return rc;
    if ((rc = tma_encode_2d(&p->tmV, vt, (uint64_t)B * heads * 64, (uint64_t)Npad, (uint64_t)Npad, 64))) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        D2S_CHECK_CUDA(cudaFuncSetAttribute((const void *)attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 88 * 1024));
        attr_set = true;
    }
    return D2S_OK;
}

int attention_tc_launch(const AttnTcPlan *p, cudaStream_t stream, bool vt_ready) {
    if (!vt_ready)   // (the engine lets the qkv GEMM epilogue write V^T directly)
        D2S_LAUNCH(v_transpose_kernel, dim3(ceil_div(p->N, 64), p->heads, p->B), 256, 0, stream, p->qkv, p->vt, p->N, p->D, p->heads, p->Npad);
    AttnArgs a;
    a.tmQK = p->tmQK; a.tmV = p->tmV; a.out = p->out; a.N = p->N; a.D = p->D; a.heads = p->heads; a.nkb = ceil_div(p->N, 128);
    a.scale_log2e = 0.125f * 1.4426950408889634f;
    D2S_LAUNCH(attention_tc_kernel, dim3(ceil_div(p->N, 128), p->heads, p->B), kTcThreads, 81920 + 128 + 3072 + 1024, stream, a);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

}  // namespace d2s
