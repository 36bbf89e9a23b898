// Fused multi-head self-attention for the DINOv2 encoder: softmax(Q K^T / sqrt(64)) V, head dim 64, non-causal,
// N <= a few thousand tokens.  Flash-style single pass: S and P never touch HBM.
//
// Round-1 implementation: warp-level mma.sync (m16n8k16, fp16 in / fp32 accumulate) with online softmax in
// registers (quad shuffles).  Attention is ~14 % of the encoder FLOPs at N = 778; the tcgen05 version is listed
// under "next" in DESIGN.md.  Replaces HF Dinov2SelfAttention's SDPA call (HF modeling_dinov2.py:203-234).
#include "common.cuh"
#include "layers.cuh"

namespace d2s {

constexpr int kAttBQ = 64, kAttBK = 64, kAttD = 64, kAttPitch = 72;  // pitch 72 halves = 144 B: conflict-free ldmatrix

__device__ __forceinline__ void cp_async16(void *dst, const void *src, bool valid) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    int sz = valid ? 16 : 0;  // src-size 0 -> 16 zero bytes
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void *p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void *p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *(uint32_t *)&h;
}

// load a 64 x 64 fp16 tile (rows row0.., columns col0..col0+63 of a [rows, ld] matrix) into padded smem
__device__ __forceinline__ void load_tile(__half *s, const __half *g, int ld, int row0, int nrows, int col0) {
    for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
        int r = i >> 3, c = (i & 7) * 8;
        bool ok = row0 + r < nrows;
        const __half *src = g + (size_t)(ok ? row0 + r : 0) * ld + col0 + c;
        cp_async16(s + r * kAttPitch + c, src, ok);
    }
}

__global__ void __launch_bounds__(128) attention_kernel(const __half *__restrict__ qkv, __half *__restrict__ out, int N, int D,
                                                        float scale_log2e) {
    pdl_wait();
    __shared__ __align__(16) __half sQ[kAttBQ * kAttPitch];
    __shared__ __align__(16) __half sK[2][kAttBK * kAttPitch];
    __shared__ __align__(16) __half sV[2][kAttBK * kAttPitch];
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kAttBQ;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
    const int ld = 3 * D;
    const __half *base = qkv + (size_t)b * N * ld;
    const int nkb = (N + kAttBK - 1) / kAttBK;

    load_tile(sQ, base, ld, q0, N, h * kAttD);
    load_tile(sK[0], base, ld, 0, N, D + h * kAttD);
    load_tile(sV[0], base, ld, 0, N, 2 * D + h * kAttD);
    cp_async_commit();

    uint32_t qf[4][4];
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int kb = 0; kb < nkb; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nkb) {
            load_tile(sK[buf ^ 1], base, ld, (kb + 1) * kAttBK, N, D + h * kAttD);
            load_tile(sV[buf ^ 1], base, ld, (kb + 1) * kAttBK, N, 2 * D + h * kAttD);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (kb == 0) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
                ldmatrix_x4(qf[ks], sQ + (warp * 16 + (lane & 15)) * kAttPitch + ks * 16 + (lane >> 4) * 8);
        }
        // S = Q K^T  (16 query rows x 64 keys per warp)
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
            for (int kp = 0; kp < 2; ++kp) {  // two k-steps (32 dims) per ldmatrix.x4
                uint32_t kf[4];
                ldmatrix_x4(kf, sK[buf] + (nt * 8 + (lane & 7)) * kAttPitch + kp * 32 + (lane >> 3) * 8);
                mma_16816(s[nt], qf[kp * 2], kf[0], kf[1]);
                mma_16816(s[nt], qf[kp * 2 + 1], kf[2], kf[3]);
            }
        }
        // mask the ragged tail of keys
        const int kbase = kb * kAttBK;
        if (kbase + kAttBK > N) {
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                int key = kbase + nt * 8 + tg * 2;
                if (key >= N) s[nt][0] = s[nt][2] = -INFINITY;
                if (key + 1 >= N) s[nt][1] = s[nt][3] = -INFINITY;
            }
        }
        // online softmax (rows g and g+8 of this warp's 16)
        float mx0 = m0, mx1 = m1;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float c0 = fast_exp2((m0 - mx0) * scale_log2e), c1 = fast_exp2((m1 - mx1) * scale_log2e);  // first block: exp2(-inf) = 0
        m0 = mx0; m1 = mx1;
        const float ms0 = mx0 * scale_log2e, ms1 = mx1 * scale_log2e;
        float rs0 = 0.f, rs1 = 0.f;
        uint32_t pf[4][4];  // P as the A operand of P V: 4 k-steps of 16 keys
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float p0 = fast_exp2(s[nt][0] * scale_log2e - ms0), p1 = fast_exp2(s[nt][1] * scale_log2e - ms0);
            float p2 = fast_exp2(s[nt][2] * scale_log2e - ms1), p3 = fast_exp2(s[nt][3] * scale_log2e - ms1);
            rs0 += p0 + p1; rs1 += p2 + p3;
            pf[nt >> 1][(nt & 1) * 2 + 0] = pack_half2(p0, p1);
            pf[nt >> 1][(nt & 1) * 2 + 1] = pack_half2(p2, p3);
        }
        l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) { o[dt][0] *= c0; o[dt][1] *= c0; o[dt][2] *= c1; o[dt][3] *= c1; }
        // O += P V
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {  // two 8-wide d tiles per ldmatrix.x4.trans
                uint32_t vf[4];
                ldmatrix_x4_trans(vf, sV[buf] + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kAttPitch + dp * 16 + (lane >> 4) * 8);
                mma_16816(o[dp * 2], pf[ks], vf[0], vf[1]);
                mma_16816(o[dp * 2 + 1], pf[ks], vf[2], vf[3]);
            }
        }
        __syncthreads();
    }
    pdl_launch_dependents();
    // finalise: row sums live in quads
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    __half *ob = out + (size_t)b * N * D + h * kAttD;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
        int d = dt * 8 + tg * 2;
        if (r0 < N) *(__half2 *)(ob + (size_t)r0 * D + d) = __floats2half2_rn(o[dt][0] * i0, o[dt][1] * i0);
        if (r1 < N) *(__half2 *)(ob + (size_t)r1 * D + d) = __floats2half2_rn(o[dt][2] * i1, o[dt][3] * i1);
    }
}

int attention_launch(const __half *qkv, __half *out, int B, int N, int D, int heads, cudaStream_t stream) {
    D2S_REQUIRE(D == heads * kAttD, "attention: head dim must be 64 (D=%d heads=%d)", D, heads);
    dim3 grid(ceil_div(N, kAttBQ), heads, B);
    const float scale_log2e = 0.125f * 1.4426950408889634f;  // head_dim^-0.5 * log2(e)
    D2S_LAUNCH(attention_kernel, grid, 128, 0, stream, qkv, out, N, D, scale_log2e);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

}  // namespace d2s
