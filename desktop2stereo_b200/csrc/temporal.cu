// Temporal (video) layers of streaming Video-Depth-Anything — reference models/video_depth_anything/motion_module/
// motion_module.py:68-134 (GroupNorm -> proj_in -> block -> proj_out -> +res), :137-187 (2 x [LN -> temporal attention -> +res],
// LN -> GEGLU FF -> +res), :212-321 (attention over the frame axis, sinusoidal APE), attention.py:182-211, :363-384,
// and the stream state machine of vda2_s.py:177-224.
//
// B200-first restatement of the streaming step.  The reference keeps, per attention block, the last 31 frames' LayerNorm'd
// hidden states [(h*w), 31, C], shifts that whole cache left by one frame every step and re-projects all 32 positions to K and
// V.  Because to_k / to_v are linear and bias-free,  to_k(h_f + pe[p]) = to_k(h_f) + to_k(pe[p]):  each frame's K' = to_k(h_f),
// V' = to_v(h_f) is computed ONCE (one [d,C] x [C,3C] GEMM on the newest frame only), stored in a ring indexed by frame number
// (no shifting copy), and the position terms PQ/PK/PV = pe @ W^T are [32,3C] tables packed with the weights.  The attention
// kernel adds the table row of the position a ring slot currently occupies.  Per frame this reads each cached value once
// (4 bytes per (location, frame, channel)) instead of re-projecting 32 frames.
#include "layers.cuh"

namespace d2s {

// ---- GroupNorm(32 groups) over an NHWC fp16 map [d, Cp] (C real channels) -> fp16 [d, C] --------------------------------
// pass 1: per (pixel slice, channel) sums -> per (slice, group) partial sums (fixed order => deterministic)
constexpr int kGnSlice = 64;   // pixels per block
__global__ void __launch_bounds__(256) gn_stats_kernel(const __half *__restrict__ x, float *__restrict__ part, int d, int C, int Cp) {
    pdl_sync();
    __shared__ float s_sum[1024], s_sq[1024];
    const int p0 = blockIdx.x * kGnSlice, p1 = min(p0 + kGnSlice, d);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f, q = 0.f;
        for (int p = p0; p < p1; ++p) { float v = __half2float(x[(size_t)p * Cp + c]); s += v; q += v * v; }
        s_sum[c] = s; s_sq[c] = q;
    }
    __syncthreads();
    const int cg = C / 32;
    if (threadIdx.x < 32) {
        float s = 0.f, q = 0.f;
        for (int c = threadIdx.x * cg; c < (threadIdx.x + 1) * cg; ++c) { s += s_sum[c]; q += s_sq[c]; }
        part[((size_t)blockIdx.x * 32 + threadIdx.x) * 2] = s;
        part[((size_t)blockIdx.x * 32 + threadIdx.x) * 2 + 1] = q;
    }
}
// pass 2: finish the statistics (double, fixed order) and normalise
__global__ void __launch_bounds__(256) gn_apply_kernel(const __half *__restrict__ x, const float *__restrict__ part, const float *__restrict__ w,
                                                       const float *__restrict__ b, __half *__restrict__ y, int d, int C, int Cp, int nslices, float eps) {
    pdl_sync();
    __shared__ float s_mean[32], s_rstd[32];
    if (threadIdx.x < 32) {
        double s = 0.0, q = 0.0;
        for (int i = 0; i < nslices; ++i) { s += (double)part[((size_t)i * 32 + threadIdx.x) * 2]; q += (double)part[((size_t)i * 32 + threadIdx.x) * 2 + 1]; }
        const double n = (double)d * (double)(C / 32), mean = s / n;
        double var = q / n - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[threadIdx.x] = (float)mean;
        s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    const int cg = C / 32;
    const long long total = (long long)d * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long p = i / C;
        const int g = c / cg;
        const float v = (__half2float(x[p * Cp + c]) - s_mean[g]) * s_rstd[g] * __ldg(w + c) + __ldg(b + c);
        y[i] = __float2half_rn(v);
    }
}

int groupnorm32_launch(const __half *x, float *partials, const float *w, const float *b, __half *y, int d, int C, int Cp, float eps, cudaStream_t stream) {
    D2S_REQUIRE(C % 32 == 0 && C <= 1024, "groupnorm: C=%d must be a multiple of 32, <= 1024", C);
    const int ns = ceil_div(d, kGnSlice);
    D2S_LAUNCH(gn_stats_kernel, ns, 256, 0, stream, x, partials, d, C, Cp);
    D2S_POST_LAUNCH();
    const int blocks = min(ceil_div((long long)d * C, 256 * 4), 4 * kNumSMs);
    D2S_LAUNCH(gn_apply_kernel, blocks, 256, 0, stream, x, partials, w, b, y, d, C, Cp, ns, eps);
    D2S_POST_LAUNCH();
    return D2S_OK;
}
size_t groupnorm32_partial_floats(int d) { return (size_t)ceil_div(d, kGnSlice) * 64; }

// ---- GEGLU: [d, 8C] (value | gate) -> value * gelu(gate) [d, 4C]  (attention.py:363-384) ---------------------------------
__global__ void geglu_kernel(const __half *__restrict__ in, __half *__restrict__ out, long long rows, int inner) {
    pdl_sync();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread = 8 channels
    const int n8 = inner / 8;
    if (i >= rows * n8) return;
    const long long r = i / n8;
    const int c = (int)(i % n8) * 8;
    const uint4 a = *(const uint4 *)(in + r * 2 * inner + c), g = *(const uint4 *)(in + r * 2 * inner + inner + c);
    const __half2 *ah = (const __half2 *)&a, *gh = (const __half2 *)&g;
    uint4 o;
    __half2 *oh = (__half2 *)&o;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 av = __half22float2(ah[j]), gv = __half22float2(gh[j]);
        oh[j] = __floats2half2_rn(av.x * (0.5f * gv.x * (1.f + erff(gv.x * 0.70710678118654752440f))),
                                  av.y * (0.5f * gv.y * (1.f + erff(gv.y * 0.70710678118654752440f))));
    }
    *(uint4 *)(out + r * inner + c) = o;
}
int geglu_launch(const __half *in, __half *out, long long rows, int inner, cudaStream_t stream) {
    D2S_REQUIRE(inner % 8 == 0, "geglu: inner=%d must be a multiple of 8", inner);
    D2S_LAUNCH(geglu_kernel, ceil_div(rows * (inner / 8), 256), 256, 0, stream, in, out, rows, inner);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

// ---- fp32 -> fp16 (GEMM A operand of proj_out) -----------------------------------------------------------------------------
__global__ void cast_f16_kernel(const float *__restrict__ in, __half *__restrict__ out, long long n4) {
    pdl_sync();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = ((const float4 *)in)[i];
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    ((uint2 *)out)[i] = make_uint2(*(const uint32_t *)&a, *(const uint32_t *)&b);
}
int cast_f16_launch(const float *in, __half *out, long long n, cudaStream_t stream) {
    D2S_REQUIRE(n % 4 == 0, "cast: n must be a multiple of 4");
    D2S_LAUNCH(cast_f16_kernel, ceil_div(n / 4, 256), 256, 0, stream, in, out, n / 4);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

// ---- streaming temporal attention ---------------------------------------------------------------------------------------------
// One block per spatial location, one warp per head (8 heads), lane = key position p (32 frame positions).
//   qkv   [d, 3C] fp16   Q' | K' | V' of the newest frame (position-free projections of its LayerNorm'd hidden state)
//   ring  [d, 32, 2C] fp16   K' | V' of the last 32 frames; frame f lives in slot f % 32
//   pe    [32, 3C] fp32  PQ | PK | PV = pe @ {to_q,to_k,to_v}^T
//   t     frames seen so far on this stream (device counter; 0 = first frame)
// First frame (vda2_s.py:196-209, motion_module.py:252-254): the sequence is that one frame at position 0.
// Later frames: positions 0..30 hold frames t-31..t-1 (frames before the first are the first frame, because the reference
// seeds its cache with 31 copies of it), position 31 is the newest frame and the only query.
__global__ void __launch_bounds__(256) temporal_attention_kernel(const __half *__restrict__ qkv, __half *__restrict__ ring, const float *__restrict__ pe,
                                                                 const long long *__restrict__ t_ptr, __half *__restrict__ out, int C, float scale) {
    pdl_sync();
    extern __shared__ float s_mem[];                 // [8 warps][hd] query  +  [8][32] probabilities
    const int hd = C >> 3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long d = blockIdx.x;
    const long long t = *t_ptr;
    const bool first = t == 0;
    const int pq = first ? 0 : 31;                   // position of the query
    float *qs = s_mem + warp * hd, *probs = s_mem + 8 * hd + warp * 32;
    const __half *cur = qkv + d * 3 * C;
    __half *rg = ring + d * 32 * 2 * C;
    const int hc = warp * hd;                        // first channel of this head
    for (int c = lane; c < hd; c += 32) qs[c] = __half2float(cur[hc + c]) + __ldg(pe + (size_t)pq * 3 * C + hc + c);
    __syncwarp();
    // scores: lane p
    const int p = lane;
    float sc = -INFINITY;
    if (!first || p == 0) {
        const long long f = t - 31 + p;                                   // frame at position p (later frames)
        const __half *kp = (first || p == 31) ? cur + C + hc : rg + (size_t)((f < 0 ? 0 : f) & 31) * 2 * C + hc;
        const float *pk = pe + (size_t)p * 3 * C + C + hc;
        float acc = 0.f;
        for (int c = 0; c < hd; c += 8) {
            const uint4 kv = *(const uint4 *)(kp + c);
            const __half2 *kh = (const __half2 *)&kv;
            const float4 p0 = __ldg((const float4 *)(pk + c)), p1 = __ldg((const float4 *)(pk + c + 4));
            const float2 k0 = __half22float2(kh[0]), k1 = __half22float2(kh[1]), k2 = __half22float2(kh[2]), k3 = __half22float2(kh[3]);
            acc += qs[c] * (k0.x + p0.x) + qs[c + 1] * (k0.y + p0.y) + qs[c + 2] * (k1.x + p0.z) + qs[c + 3] * (k1.y + p0.w) +
                   qs[c + 4] * (k2.x + p1.x) + qs[c + 5] * (k2.y + p1.y) + qs[c + 6] * (k3.x + p1.z) + qs[c + 7] * (k3.y + p1.w);
        }
        sc = acc * scale;
    }
    float m = sc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float ex = (sc == -INFINITY) ? 0.f : __expf(sc - m);
    float sum = ex;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    probs[lane] = ex / sum;
    __syncwarp();
    // output: lane c
    const int np = first ? 1 : 32;
    for (int c = lane; c < hd; c += 32) {
        float acc = 0.f;
        for (int pp = 0; pp < np; ++pp) {
            const long long f = t - 31 + pp;
            const __half *vp = (first || pp == 31) ? cur + 2 * C + hc : rg + (size_t)((f < 0 ? 0 : f) & 31) * 2 * C + C + hc;
            acc += probs[pp] * (__half2float(vp[c]) + __ldg(pe + (size_t)pp * 3 * C + 2 * C + hc + c));
        }
        out[d * C + hc + c] = __float2half_rn(acc);
    }
    // the newest frame enters the ring (slot t % 32 held frame t-32, which no position refers to any more)
    __half *slot = rg + (size_t)(t & 31) * 2 * C;
    for (int c = lane; c < hd; c += 32) { slot[hc + c] = cur[C + hc + c]; slot[C + hc + c] = cur[2 * C + hc + c]; }
}

int temporal_attention_launch(const __half *qkv, __half *ring, const float *pe, const long long *t_ptr, __half *out, int d, int C, cudaStream_t stream) {
    D2S_REQUIRE(C % 64 == 0 && C <= 1024, "temporal attention: C=%d must be a multiple of 64, <= 1024 (8 heads, head dim multiple of 8)", C);
    const int hd = C / 8;
    const size_t smem = (size_t)(8 * hd + 8 * 32) * sizeof(float);
    D2S_LAUNCH(temporal_attention_kernel, d, 256, smem, stream, qkv, ring, pe, t_ptr, out, C, 1.0f / sqrtf((float)hd));
    D2S_POST_LAUNCH();
    return D2S_OK;
}

__global__ void frame_counter_kernel(long long *t) {
    pdl_sync();
    *t += 1;
}
int frame_counter_inc_launch(long long *t, cudaStream_t stream) {
    D2S_LAUNCH(frame_counter_kernel, 1, 1, 0, stream, t);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

}  // namespace d2s
