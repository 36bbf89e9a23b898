// Stereo warp + pad + SBS/TAB pack + Half-mode 2:1 mean + clamp — ONE kernel, HBM-bound, no tensor cores.
//
// Replaces make_sbs_core (reference depth.py:2122-2184), pad_to_aspect_tensor (:2106-2119), the
// HWC/float conversion of chw_tensor_to_numpy (:767-773) and, when the depth map is passed at model
// resolution, the final bilinear upsample of predict_depth (:1998-2004).
//
// Arithmetic contract (bit-exact against oracle/warp_oracle.c, which is pinned bit-exactly on the
// reference's CPU output): strict fp32, compiled with -fmad=false; every FMA below is explicit and sits
// exactly where ATen's linspace / grid_sampler_2d kernels contract one.
//
// Work decomposition: one thread produces 4 consecutive OUTPUT pixels of one output row.  The reference
// materialises shifts, two [h,w,2] fp32 grids, two eye images, the concatenation and the pooled copy
// (~10x the algorithmic bytes); here each source byte is read from HBM once and each
// output byte written once.
#include <cstdlib>

#include "common.cuh"

namespace d2s {

struct WarpK {
    const void *rgb; long long rsc, rsy, rsx;
    void *out; long long osc, osy, osx;
    const void *depth; int dh, dw; float dscale_h, dscale_w;
    int h, w, ph, pw, top, left, oh, ow;
    int tab, half, gather, lowres, rgb_round;
    float conv, ratio, max_px, strength, two_over_wm1;
    float xstep; int xhalf; float ystep; int yhalf;
    int *idx_l, *idx_r;
};

// torch.linspace(-1, 1, n)[i]  (RangeFactories: two-sided, one FMA per element)
__device__ __forceinline__ float linspace_pm1(int i, int n, float step, int half) {
    if (n == 1) return -1.0f;
    return (i < half) ? __fmaf_rn(step, (float)i, -1.0f) : __fmaf_rn(-step, (float)(n - i - 1), 1.0f);
}

// reflect_coordinates for |coord| >= span (rare: only within |shift| of the right border); fmod/floor/div are exact here.
__device__ __noinline__ float reflect_slow(float in, float span) {
    float extra = fmodf(in, span);
    int flips = (int)floorf(__fdiv_rn(in, span));
    return (flips % 2 == 0) ? extra : __fsub_rn(span, extra);
}

// grid_sampler_compute_source_index, padding_mode=reflection, align_corners=True (GridSampler.cuh)
__device__ __forceinline__ float source_index(float coord, int size) {
    coord = __fmul_rn(__fmul_rn(__fadd_rn(coord, 1.f), 0.5f), (float)(size - 1));
    if (size == 1) return 0.f;
    const float span = (float)(size - 1);
    float in = fabsf(coord);
    if (in >= span) in = reflect_slow(in, span);
    // in < span: fmod(in,span)==in and floor(in/span)==0 exactly, so reflect_coordinates returns `in`.
    return fminf(span, fmaxf(in, 0.f));
}

struct RowCtx {
    int iy0;
    float wy0, wy1;  // (iy_se - iy), (iy - iy_nw)
    bool ok0, ok1;
};

__device__ __forceinline__ RowCtx make_row(const WarpK &k, int y) {
    RowCtx r;
    float iy = source_index(linspace_pm1(y, k.h, k.ystep, k.yhalf), k.h);
    r.iy0 = (int)floorf(iy);
    r.wy0 = __fsub_rn((float)(r.iy0 + 1), iy);
    r.wy1 = __fsub_rn(iy, (float)r.iy0);
    r.ok0 = r.iy0 >= 0 && r.iy0 < k.h;
    r.ok1 = r.iy0 + 1 < k.h;
    return r;
}

// depth at full-res pixel (y,x) in the value set of DT.  lowres: upsample_bilinear2d (align_corners=False),
// restated from ATen's CUDA kernel (UpSampleBilinear2d.cu): accumulate in fp32, round to DT.
// (noinline helpers take their parameters by value: a reference to the kernel's parameter struct would force a per-thread
// copy of the whole struct into local memory)
template <typename DT>
__device__ __noinline__ float load_depth_lowres(const DT *d, int dh, int dw, float dscale_h, float dscale_w, int y, int x);

template <typename DT>
__device__ __forceinline__ float load_depth(const WarpK &k, int y, int x) {
    const DT *d = (const DT *)k.depth;
    if (!k.lowres) return to_f32<DT>(__ldg(d + (size_t)y * k.w + x));
    return load_depth_lowres<DT>(d, k.dh, k.dw, k.dscale_h, k.dscale_w, y, x);
}

template <typename DT>
__device__ __noinline__ float load_depth_lowres(const DT *d, int dh, int dw, float dscale_h, float dscale_w, int y, int x) {
    float h1r = fmaxf(__fmaf_rn(dscale_h, (float)y + 0.5f, -0.5f), 0.f);
    float w1r = fmaxf(__fmaf_rn(dscale_w, (float)x + 0.5f, -0.5f), 0.f);
    int h1 = (int)h1r, w1 = (int)w1r;
    int h1p = (h1 < dh - 1) ? 1 : 0, w1p = (w1 < dw - 1) ? 1 : 0;
    float h1l = __fsub_rn(h1r, (float)h1), h0l = __fsub_rn(1.f, h1l);
    float w1l = __fsub_rn(w1r, (float)w1), w0l = __fsub_rn(1.f, w1l);
    const DT *r0 = d + (size_t)h1 * dw, *r1 = d + (size_t)(h1 + h1p) * dw;
    float a = to_f32<DT>(__ldg(r0 + w1)), b = to_f32<DT>(__ldg(r0 + w1 + w1p));
    float c = to_f32<DT>(__ldg(r1 + w1)), e = to_f32<DT>(__ldg(r1 + w1 + w1p));
    float top = __fmaf_rn(w0l, a, __fmul_rn(w1l, b));
    float bot = __fmaf_rn(w0l, c, __fmul_rn(w1l, e));
    return round_to<DT>(__fmaf_rn(h0l, top, __fmul_rn(h1l, bot)));
}

template <typename RT, typename DT>
__device__ __forceinline__ float load_rgb(const WarpK &k, int c, int y, int x) {
    float v = to_f32<RT>(__ldg((const RT *)k.rgb + c * k.rsc + (long long)y * k.rsy + (long long)x * k.rsx));
    if (sizeof(RT) != 1 && k.rgb_round) v = round_to<DT>(v);  // rgb.to(depth.dtype), depth.py:2209-2215
    if (sizeof(RT) != 1) v = fminf(fmaxf(v, 0.f), 255.f);  // img.clamp(0,255), depth.py:2142 (no-op for u8)
    return v;
}

// One warped eye pixel.  e: 0 left, 1 right.
template <typename RT, typename DT>
__device__ __forceinline__ void eye_pixel(const WarpK &k, const RowCtx &row, int e, int y, int x, float rgb[3]) {
    // shift chain, depth.py:2143-2147 (+ :2154): each tensor-scalar op is fp32 math rounded to DT
    float d = round_to<DT>(__fsub_rn(load_depth<DT>(k, y, x), k.conv));
    float inv = round_to<DT>(__fmul_rn(-d, k.ratio));
    float s = round_to<DT>(__fmul_rn(inv, k.max_px));
    s = round_to<DT>(__fmul_rn(s, k.strength));
    int *idx = e ? k.idx_r : k.idx_l;
    if (k.gather) {
        // depth.py:2163-2172
        float c = e ? __fsub_rn((float)x, s) : __fadd_rn((float)x, s);
        int ci = (int)fminf(fmaxf(c, 0.f), (float)(k.w - 1));
        if (idx) idx[(size_t)y * k.w + x] = ci;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) rgb[ch] = load_rgb<RT, DT>(k, ch, y, ci);
        return;
    }
    // depth.py:2152-2160
    float sn = round_to<DT>(__fmul_rn(s, k.two_over_wm1));
    float xs = linspace_pm1(x, k.w, k.xstep, k.xhalf);
    float gx = e ? __fsub_rn(xs, sn) : __fadd_rn(xs, sn);
    float ix = source_index(gx, k.w);
    int ix0 = (int)floorf(ix);
    if (idx) idx[(size_t)y * k.w + x] = ix0;
    float wx0 = __fsub_rn((float)(ix0 + 1), ix), wx1 = __fsub_rn(ix, (float)ix0);
    float nw = __fmul_rn(wx0, row.wy0), ne = __fmul_rn(wx1, row.wy0);
    float sw = __fmul_rn(wx0, row.wy1), se = __fmul_rn(wx1, row.wy1);
    bool okx1 = ix0 + 1 < k.w;  // ix0 in [0,w-1] after the clip
    // A tap whose weight is exactly 0 contributes fma(v,0,acc)==acc: skipping its load is bit-identical.
    bool t_nw = row.ok0, t_ne = row.ok0 && okx1 && ne != 0.f;
    bool t_sw = row.ok1 && sw != 0.f, t_se = row.ok1 && okx1 && se != 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float acc = 0.f;
        if (t_nw) acc = __fmaf_rn(load_rgb<RT, DT>(k, ch, row.iy0, ix0), nw, acc);
        if (t_ne) acc = __fmaf_rn(load_rgb<RT, DT>(k, ch, row.iy0, ix0 + 1), ne, acc);
        if (t_sw) acc = __fmaf_rn(load_rgb<RT, DT>(k, ch, row.iy0 + 1, ix0), sw, acc);
        if (t_se) acc = __fmaf_rn(load_rgb<RT, DT>(k, ch, row.iy0 + 1, ix0 + 1), se, acc);
        rgb[ch] = acc;
    }
}

// element (cy,cx) of cat([pad(left), pad(right)]) — depth.py:2175-2181
template <typename RT, typename DT>
__device__ __forceinline__ void cat_pixel(const WarpK &k, const RowCtx *rows, int cy, int cx, float rgb[3]) {
    int e, ey, ex;
    if (k.tab) { e = cy >= k.ph; ey = cy - e * k.ph; ex = cx; }
    else       { e = cx >= k.pw; ex = cx - e * k.pw; ey = cy; }
    ey -= k.top; ex -= k.left;
    if (ey < 0 || ey >= k.h || ex < 0 || ex >= k.w) { rgb[0] = rgb[1] = rgb[2] = 0.f; return; }
    // rows[] caches the per-row y interpolation context (warp-uniform): slot = which of the thread's rows
    const RowCtx &row = rows[(k.half && k.tab) ? (cy & 1) : 0];
    eye_pixel<RT, DT>(k, row, e, ey, ex, rgb);
}

// 4 consecutive elements of OT packed into one register vector (16 B f32, 8 B f16/bf16, 4 B u8)
template <typename OT> struct Vec4;
template <> struct Vec4<float> { typedef float4 type; static __device__ __forceinline__ type pack(float a, float b, float c, float d) { return make_float4(a, b, c, d); } };
template <> struct Vec4<__half> { typedef uint2 type; static __device__ __forceinline__ type pack(float a, float b, float c, float d) {
    __half2 lo = __halves2half2(__float2half_rn(a), __float2half_rn(b)), hi = __halves2half2(__float2half_rn(c), __float2half_rn(d));
    return make_uint2(*(uint32_t *)&lo, *(uint32_t *)&hi); } };
template <> struct Vec4<__nv_bfloat16> { typedef uint2 type; static __device__ __forceinline__ type pack(float a, float b, float c, float d) {
    __nv_bfloat162 lo = __halves2bfloat162(__float2bfloat16_rn(a), __float2bfloat16_rn(b)), hi = __halves2bfloat162(__float2bfloat16_rn(c), __float2bfloat16_rn(d));
    return make_uint2(*(uint32_t *)&lo, *(uint32_t *)&hi); } };
template <> struct Vec4<uint8_t> { typedef uint32_t type; static __device__ __forceinline__ type pack(float a, float b, float c, float d) {
    return (uint32_t)from_f32<uint8_t>(a) | ((uint32_t)from_f32<uint8_t>(b) << 8) | ((uint32_t)from_f32<uint8_t>(c) << 16) | ((uint32_t)from_f32<uint8_t>(d) << 24); } };

// 4 pixels x 3 channels -> memory (vectorised when the layout allows)
template <typename OT>
__device__ __forceinline__ void store_px4(const WarpK &k, int oy, int ox, int n, const float (*v)[3]) {
    typedef typename Vec4<OT>::type V;
    OT *base = (OT *)k.out + (long long)oy * k.osy + (long long)ox * k.osx;
    if (n == 4 && k.osx == 3 && k.osc == 1 && ((uintptr_t)base % sizeof(V)) == 0) {
        // HWC: 12 contiguous elements
        V *d = (V *)base;
        d[0] = Vec4<OT>::pack(v[0][0], v[0][1], v[0][2], v[1][0]);
        d[1] = Vec4<OT>::pack(v[1][1], v[1][2], v[2][0], v[2][1]);
        d[2] = Vec4<OT>::pack(v[2][2], v[3][0], v[3][1], v[3][2]);
        return;
    }
    if (n == 4 && k.osx == 1 && ((uintptr_t)base % sizeof(V)) == 0 && ((k.osc * (long long)sizeof(OT)) % sizeof(V)) == 0) {
        // CHW: 4 contiguous elements per plane
#pragma unroll
        for (int c = 0; c < 3; ++c) *(V *)(base + c * k.osc) = Vec4<OT>::pack(v[0][c], v[1][c], v[2][c], v[3][c]);
        return;
    }
    for (int p = 0; p < n; ++p)
#pragma unroll
        for (int c = 0; c < 3; ++c) base[c * k.osc + p * k.osx] = from_f32<OT>(v[p][c]);
}

template <typename RT, typename DT, typename OT>
__global__ void __launch_bounds__(256) warp_sbs_kernel(const WarpK k) {
    const int ox0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int oy = blockIdx.y * blockDim.y + threadIdx.y;
    if (oy >= k.oh || ox0 >= k.ow) return;
    const int n = min(4, k.ow - ox0);

    // y-interpolation context for the (one or two) eye rows this thread touches
    RowCtx rows[2];
    {
        int cy0 = (k.half && k.tab) ? 2 * oy : oy;
        int ey0 = (k.tab ? (cy0 >= k.ph ? cy0 - k.ph : cy0) : cy0) - k.top;
        // Half-TAB pairs (2oy, 2oy+1) never straddle the eye seam unless ph is odd; rows[] is indexed by
        // (cy & 1), so slot 0 holds the even cat row and slot 1 the odd one.
        int cy1 = cy0 + 1;
        int ey1 = (k.tab ? (cy1 >= k.ph ? cy1 - k.ph : cy1) : cy1) - k.top;
        rows[0] = make_row(k, min(max(ey0, 0), k.h - 1));
        rows[1] = (k.half && k.tab) ? make_row(k, min(max(ey1, 0), k.h - 1)) : rows[0];
    }

    float v[4][3];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        if (p >= n) { v[p][0] = v[p][1] = v[p][2] = 0.f; continue; }
        const int ox = ox0 + p;
        float a[3];
        if (!k.half) {
            cat_pixel<RT, DT>(k, rows, oy, ox, a);
        } else {
            // F.interpolate(mode="area") with an exact 2:1 ratio == adaptive_avg_pool: (a + b) / 2
            float b[3];
            if (k.tab) { cat_pixel<RT, DT>(k, rows, 2 * oy, ox, a); cat_pixel<RT, DT>(k, rows, 2 * oy + 1, ox, b); }
            else       { cat_pixel<RT, DT>(k, rows, oy, 2 * ox, a); cat_pixel<RT, DT>(k, rows, oy, 2 * ox + 1, b); }
#pragma unroll
            for (int c = 0; c < 3; ++c) a[c] = __fmul_rn(__fadd_rn(a[c], b[c]), 0.5f);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) v[p][c] = fminf(fmaxf(a[c], 0.f), 255.f);  // final clamp, depth.py:2184
    }
    store_px4<OT>(k, oy, ox0, n, v);
}

// ------------------------------------------------------------------------------------------------
// Fast path (bilinear branch, no 16:9 padding, Full-SBS / Full-TAB / Half-SBS): source-centric.
// One block = one source row segment.  The rgb row(s) the segment can reach (|shift| is bounded by the parameters for
// depth in [0,1]) are staged once in shared memory as fp32 planes with coalesced loads; each thread then owns NP
// consecutive source pixels, reads their depth once, runs the shift chain once and produces BOTH eyes from smem taps.
// Same arithmetic, same FMA order as the generic kernel => bit-identical output (tests compare both against the oracle).
// ------------------------------------------------------------------------------------------------
// the taps of one eye pixel straight from global memory (fast kernel: the tap left the staged window, i.e. depth outside [0,1])
template <typename RT, typename DT>
__device__ __forceinline__ float stage_value(float v, bool rgb_round) {
    if (sizeof(RT) != 1) {
        if (rgb_round) v = round_to<DT>(v);
        v = fminf(fmaxf(v, 0.f), 255.f);
    }
    return v;
}
template <typename RT, typename DT>
__device__ __noinline__ float3 taps_global(const RT *rgb, long long rsc, long long rsy, long long rsx, int w, int rgb_round, int iy0, bool two,
                                           int ix0, float nw, float ne, float sw, float se) {
    const bool okx1 = ix0 + 1 < w;
    float o[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const RT *p0 = rgb + ch * rsc + (long long)iy0 * rsy + (long long)ix0 * rsx;
        float acc = __fmaf_rn(stage_value<RT, DT>(to_f32<RT>(__ldg(p0)), rgb_round), nw, 0.f);
        if (okx1) acc = __fmaf_rn(stage_value<RT, DT>(to_f32<RT>(__ldg(p0 + rsx)), rgb_round), ne, acc);
        if (two) {
            acc = __fmaf_rn(stage_value<RT, DT>(to_f32<RT>(__ldg(p0 + rsy)), rgb_round), sw, acc);
            if (okx1) acc = __fmaf_rn(stage_value<RT, DT>(to_f32<RT>(__ldg(p0 + rsy + rsx)), rgb_round), se, acc);
        }
        o[ch] = acc;
    }
    return make_float3(o[0], o[1], o[2]);
}

// ---- staging: one source row segment -> fp32 plane in shared memory (clamped / rounded exactly like load_rgb) ----
// 16 bytes of RT -> 16/sizeof(RT) staged floats
template <typename RT, typename DT> struct StageVec;
template <typename DT> struct StageVec<__half, DT> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void unpack(uint4 raw, bool rgb_round, float *f) {
        const __half2 lo = __float2half2_rn(0.f), hi = __float2half2_rn(255.f);
        const __half2 *h = (const __half2 *)&raw;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // clamp in fp16 (exact: both bounds are fp16 values; rounding to DT afterwards commutes with the clamp)
            float2 t = __half22float2(__hmin2(__hmax2(h[j], lo), hi));
            f[2 * j] = rgb_round ? round_to<DT>(t.x) : t.x;
            f[2 * j + 1] = rgb_round ? round_to<DT>(t.y) : t.y;
        }
    }
};
template <typename DT> struct StageVec<uint8_t, DT> {
    static constexpr int N = 16;
    static __device__ __forceinline__ void unpack(uint4 raw, bool, float *f) {
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            f[4 * j] = (float)(w[j] & 0xffu); f[4 * j + 1] = (float)((w[j] >> 8) & 0xffu);
            f[4 * j + 2] = (float)((w[j] >> 16) & 0xffu); f[4 * j + 3] = (float)(w[j] >> 24);
        }
    }
};
template <typename DT> struct StageVec<float, DT> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void unpack(uint4 raw, bool rgb_round, float *f) {
        const float w[4] = {__uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z), __uint_as_float(raw.w)};
#pragma unroll
        for (int j = 0; j < 4; ++j) f[j] = stage_value<float, DT>(w[j], rgb_round);
    }
};
template <typename DT> struct StageVec<__nv_bfloat16, DT> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void unpack(uint4 raw, bool rgb_round, float *f) {
        const __nv_bfloat16 *h = (const __nv_bfloat16 *)&raw;
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = stage_value<__nv_bfloat16, DT>(__bfloat162float(h[j]), rgb_round);
    }
};

// NP consecutive depth values with one vector load (pointer aligned to NP * sizeof(DT))
template <typename DT, int NP> struct alignas(sizeof(DT) * NP <= 16 ? sizeof(DT) * NP : 16) DepthPack { DT v[NP]; };

// One warped eye pixel straight from global memory: the rare cases the staged window cannot serve (a tap outside the window =
// depth outside [0,1], or a coordinate that needs grid_sample's reflection).  Same arithmetic as eye_pixel().
template <typename RT, typename DT>
__device__ __noinline__ float3 slow_eye(const RT *rgb, long long rsc, long long rsy, long long rsx, int w, int rgb_round, int iy0, float wy0, float wy1,
                                        bool two, float gx) {
    const float ix = source_index(gx, w);
    const int ix0 = (int)floorf(ix);
    const float wx0 = __fsub_rn((float)(ix0 + 1), ix), wx1 = __fsub_rn(ix, (float)ix0);
    return taps_global<RT, DT>(rgb, rsc, rsy, rsx, w, rgb_round, iy0, two, ix0, __fmul_rn(wx0, wy0), __fmul_rn(wx1, wy0), __fmul_rn(wx0, wy1),
                               __fmul_rn(wx1, wy1));
}

__device__ __forceinline__ float clamp255(float v) { return fminf(fmaxf(v, 0.f), 255.f); }
// staged pixel p lives at float4 index p + p/8: one pad slot per 8 pixels keeps both the 8-pixels-per-thread staging stores and
// the strided tap loads free of shared-memory bank conflicts
__device__ __forceinline__ int px_slot(int p) { return p + (p >> 3); }

// OL: output layout known at compile time — 0: HWC contiguous (sx = 3, sc = 1), 1: planar CHW (sx = 1)
// flags: bit 0 = rgb rows may be staged with 16-byte loads, bit 1 = depth rows may be read with vector loads
template <typename RT, typename DT, typename OT, int HALF, int OL, int THREADS>
__global__ void __launch_bounds__(THREADS) warp_sbs_fast_kernel(const WarpK k, int margin, int flags) {
    typedef typename Vec4<OT>::type V;
    constexpr int NP = HALF ? 8 : 4;                 // source pixels per thread (4 output pixels per eye)
    constexpr int SEG = THREADS * NP;                // source pixels per block
    extern __shared__ __align__(16) float4 s_px[];   // [rows(1|2)][pitch4]: staged pixels, (r, g, b, -) as fp32
    const int y = blockIdx.y;
    const int seg0 = blockIdx.x * SEG;
    const int pitch4 = px_slot(SEG + 2 * margin) + 1;
    const int lo = max(seg0 - margin, 0), hi = min(seg0 + SEG + margin, k.w);   // staged source columns [lo, hi); margin % 16 == 0
    const int tw = hi - lo;
    const RowCtx row = make_row(k, y);
    const bool two = row.ok1 && row.wy1 != 0.f;      // second source row contributes (block-uniform)
    const int nrows = two ? 2 : 1;
    // the depth values of this thread's pixels: issued first, so that their DRAM round trip overlaps the staging loads
    const int x0 = seg0 + threadIdx.x * NP;
    const int npx = min(NP, k.w - x0);               // <= 0: thread beyond the row; < NP only for the last thread of a row
    DepthPack<DT, NP> pk;
    const bool dvec = (flags & 2) && npx == NP;
    if (dvec) pk = *(const DepthPack<DT, NP> *)((const DT *)k.depth + (size_t)y * k.w + x0);
    {
        typedef StageVec<RT, DT> SV;
        const bool rr = k.rgb_round;
        const int nv = (flags & 1) ? tw / SV::N : 0;
        const RT *src = (const RT *)k.rgb + (long long)row.iy0 * k.rsy + (long long)lo * k.rsx;
        for (int i = threadIdx.x; i < nv; i += THREADS) {
            uint4 raw[2][3];                         // every load of both rows in flight before the first use
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                raw[0][c] = __ldg((const uint4 *)(src + c * k.rsc) + i);
                if (two) raw[1][c] = __ldg((const uint4 *)(src + k.rsy + c * k.rsc) + i);
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (r == 1 && !two) break;
                float f0[SV::N], f1[SV::N], f2[SV::N];
                SV::unpack(raw[r][0], rr, f0); SV::unpack(raw[r][1], rr, f1); SV::unpack(raw[r][2], rr, f2);
                float4 *dst = s_px + r * pitch4;
#pragma unroll
                for (int j = 0; j < SV::N; ++j) dst[px_slot(i * SV::N + j)] = make_float4(f0[j], f1[j], f2[j], 0.f);
            }
        }
        for (int r = 0; r < nrows; ++r) {
            float4 *dst = s_px + r * pitch4;
            for (int x = nv * SV::N + threadIdx.x; x < tw; x += THREADS) {
                const RT *p = src + r * k.rsy + (long long)x * k.rsx;
                dst[px_slot(x)] = make_float4(stage_value<RT, DT>(to_f32<RT>(__ldg(p)), rr), stage_value<RT, DT>(to_f32<RT>(__ldg(p + k.rsc)), rr),
                                              stage_value<RT, DT>(to_f32<RT>(__ldg(p + 2 * k.rsc)), rr), 0.f);
            }
        }
    }
    __syncthreads();
    if (x0 >= k.w) return;
    const float4 *s0 = s_px, *s1 = s_px + pitch4;

    float dv[NP];
    if (dvec) {
#pragma unroll
        for (int p = 0; p < NP; ++p) dv[p] = to_f32<DT>(pk.v[p]);
    } else {
#pragma unroll
        for (int p = 0; p < NP; ++p) dv[p] = load_depth<DT>(k, y, min(x0 + p, k.w - 1));   // clamped duplicates are never stored
    }

    const float span = (float)(k.w - 1);
    float v[2][4][3];                                // [eye][output pixel][channel]
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int x = min(x0 + p, k.w - 1);
        // shift chain, depth.py:2143-2147, :2154 (identical to eye_pixel)
        float d = round_to<DT>(__fsub_rn(dv[p], k.conv));
        float inv = round_to<DT>(__fmul_rn(-d, k.ratio));
        float s = round_to<DT>(__fmul_rn(inv, k.max_px));
        s = round_to<DT>(__fmul_rn(s, k.strength));
        const float sn = round_to<DT>(__fmul_rn(s, k.two_over_wm1));
        const float xs = linspace_pm1(x, k.w, k.xstep, k.xhalf);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            // grid_sampler source index (same operations as source_index(); the reflection case is left to slow_eye)
            const float gx = e ? __fsub_rn(xs, sn) : __fadd_rn(xs, sn);
            const float in = fabsf(__fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), span));
            const float ix = fminf(span, fmaxf(in, 0.f));
            const int ix0 = __float2int_rz(ix);      // ix >= 0 after the clip: truncation == floor
            const float fx0 = (float)ix0;
            const float wx0 = __fsub_rn(__fadd_rn(fx0, 1.f), ix), wx1 = __fsub_rn(ix, fx0);   // (float)(ix0 + 1) == fx0 + 1 exactly (ix0 < 2^24)
            const int a = ix0 - lo, b = a + (ix0 + 1 < k.w ? 1 : 0);
            // the east taps are taken unconditionally: when ix0 + 1 == w their weights are exactly 0 (ix == ix0 after the clip)
            const int ia = px_slot(min(max(a, 0), tw - 1)), ib = px_slot(min(max(b, 0), tw - 1));
            const float4 p00 = s0[ia], p01 = s0[ib];
            const float nw = __fmul_rn(wx0, row.wy0), ne = __fmul_rn(wx1, row.wy0);
            float3 c;
            c.x = __fmaf_rn(p01.x, ne, __fmaf_rn(p00.x, nw, 0.f));
            c.y = __fmaf_rn(p01.y, ne, __fmaf_rn(p00.y, nw, 0.f));
            c.z = __fmaf_rn(p01.z, ne, __fmaf_rn(p00.z, nw, 0.f));
            if (two) {
                const float4 p10 = s1[ia], p11 = s1[ib];
                const float sw = __fmul_rn(wx0, row.wy1), se = __fmul_rn(wx1, row.wy1);
                c.x = __fmaf_rn(p11.x, se, __fmaf_rn(p10.x, sw, c.x));
                c.y = __fmaf_rn(p11.y, se, __fmaf_rn(p10.y, sw, c.y));
                c.z = __fmaf_rn(p11.z, se, __fmaf_rn(p10.z, sw, c.z));
            }
            if (in >= span || a < 0 || b >= tw)      // rare: outside the staged window, or a coordinate grid_sample would reflect
                c = slow_eye<RT, DT>((const RT *)k.rgb, k.rsc, k.rsy, k.rsx, k.w, k.rgb_round, row.iy0, row.wy0, row.wy1, two, gx);
            if (!HALF) {
                v[e][p & 3][0] = clamp255(c.x); v[e][p & 3][1] = clamp255(c.y); v[e][p & 3][2] = clamp255(c.z);   // final clamp, depth.py:2184
            } else if ((p & 1) == 0) {
                v[e][p >> 1][0] = c.x; v[e][p >> 1][1] = c.y; v[e][p >> 1][2] = c.z;
            } else {   // F.interpolate(mode="area") at an exact 2:1 ratio: (a + b) / 2, then the final clamp
                v[e][p >> 1][0] = clamp255(__fmul_rn(__fadd_rn(v[e][p >> 1][0], c.x), 0.5f));
                v[e][p >> 1][1] = clamp255(__fmul_rn(__fadd_rn(v[e][p >> 1][1], c.y), 0.5f));
                v[e][p >> 1][2] = clamp255(__fmul_rn(__fadd_rn(v[e][p >> 1][2], c.z), 0.5f));
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int oy = k.tab ? e * k.h + y : y;
        const int ox = (k.tab ? 0 : e * (HALF ? k.w / 2 : k.w)) + (HALF ? x0 / 2 : x0);
        const int valid = HALF ? npx / 2 : npx;
        OT *base = (OT *)k.out + (long long)oy * k.osy + (long long)ox * (OL == 0 ? 3 : 1);
        if (valid == 4 && ((uintptr_t)base % sizeof(V)) == 0 && (OL == 0 || (k.osc * (long long)sizeof(OT)) % sizeof(V) == 0)) {
            if (OL == 0) {
                V *dvp = (V *)base;
                dvp[0] = Vec4<OT>::pack(v[e][0][0], v[e][0][1], v[e][0][2], v[e][1][0]);
                dvp[1] = Vec4<OT>::pack(v[e][1][1], v[e][1][2], v[e][2][0], v[e][2][1]);
                dvp[2] = Vec4<OT>::pack(v[e][2][2], v[e][3][0], v[e][3][1], v[e][3][2]);
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) *(V *)(base + c * k.osc) = Vec4<OT>::pack(v[e][0][c], v[e][1][c], v[e][2][c], v[e][3][c]);
            }
        } else {
#pragma unroll
            for (int pp = 0; pp < 4; ++pp)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    if (pp < valid) base[OL == 0 ? pp * 3 + c : c * k.osc + pp] = from_f32<OT>(v[e][pp][c]);
        }
    }
}

static void pad_geometry(int h, int w, int fill, int *ph, int *pw, int *top, int *left) {
    // depth.py:2106-2119 (python float == double)
    *ph = h; *pw = w; *top = 0; *left = 0;
    if (!fill) return;
    double r_img = (double)w / (double)h, r_t = 16.0 / 9.0;
    if (fabs(r_img - r_t) < 1e-3) return;
    if (r_img > r_t) { int nh = (int)nearbyint((double)w / r_t); *ph = nh; *top = (nh - h) / 2; }
    else             { int nw = (int)nearbyint((double)h * r_t); *pw = nw; *left = (nw - w) / 2; }
}

template <typename RT, typename DT>
static int launch_out(const WarpK &k, int out_dtype, dim3 grid, dim3 block, d2s_stream_t st) {
    switch (out_dtype) {
        case D2S_F32: D2S_LAUNCH((warp_sbs_kernel<RT, DT, float>), grid, block, 0, st, k); break;
        case D2S_F16: D2S_LAUNCH((warp_sbs_kernel<RT, DT, __half>), grid, block, 0, st, k); break;
        case D2S_U8:  D2S_LAUNCH((warp_sbs_kernel<RT, DT, uint8_t>), grid, block, 0, st, k); break;
        case D2S_BF16: D2S_LAUNCH((warp_sbs_kernel<RT, DT, __nv_bfloat16>), grid, block, 0, st, k); break;
        default: return set_error(D2S_ERR_UNSUPPORTED, "d2s_make_sbs: out dtype %d unsupported", out_dtype);
    }
    return D2S_OK;
}
template <typename RT>
static int launch_depth(const WarpK &k, int depth_dtype, int out_dtype, dim3 grid, dim3 block, d2s_stream_t st) {
    switch (depth_dtype) {
        case D2S_F32: return launch_out<RT, float>(k, out_dtype, grid, block, st);
        case D2S_F16: return launch_out<RT, __half>(k, out_dtype, grid, block, st);
        case D2S_BF16: return launch_out<RT, __nv_bfloat16>(k, out_dtype, grid, block, st);
        default: return set_error(D2S_ERR_UNSUPPORTED, "d2s_make_sbs: depth dtype %d unsupported", depth_dtype);
    }
}

template <typename RT, typename DT, typename OT>
static int launch_fast_t(const WarpK &k, bool half, int margin, d2s_stream_t st) {
    constexpr int kFull = 128, kHalf = 128;          // threads per block: 512 (Full) / 1024 (Half-SBS) source pixels per block
    const int seg = half ? 1024 : 512;               // small blocks: while some blocks of an SM wait on their staging loads, others compute
    dim3 grid(ceil_div(k.w, seg), k.h);
    const int padded = seg + 2 * margin;
    size_t smem = (size_t)2 * (padded + (padded >> 3) + 1) * sizeof(float4);
    const bool hwc = k.osx == 3 && k.osc == 1, chw = k.osx == 1;
    if (!hwc && !chw) return -1;
    int flags = 0;
    // 16-byte staging loads: contiguous pixels, every row/plane start 16-byte aligned (lo is a multiple of 8... of 16 for u8)
    const size_t es = sizeof(RT);
    if (k.rsx == 1 && ((uintptr_t)k.rgb % 16) == 0 && (k.rsy * (long long)es) % 16 == 0 && (k.rsc * (long long)es) % 16 == 0 && (margin * es) % 16 == 0) flags |= 1;
    const int np = half ? 8 : 4;
    if (!k.lowres && k.w % np == 0 && ((uintptr_t)k.depth % 16) == 0) flags |= 2;
    if (half) { if (hwc) D2S_LAUNCH((warp_sbs_fast_kernel<RT, DT, OT, 1, 0, kHalf>), grid, kHalf, smem, st, k, margin, flags);
                else     D2S_LAUNCH((warp_sbs_fast_kernel<RT, DT, OT, 1, 1, kHalf>), grid, kHalf, smem, st, k, margin, flags); }
    else      { if (hwc) D2S_LAUNCH((warp_sbs_fast_kernel<RT, DT, OT, 0, 0, kFull>), grid, kFull, smem, st, k, margin, flags);
                else     D2S_LAUNCH((warp_sbs_fast_kernel<RT, DT, OT, 0, 1, kFull>), grid, kFull, smem, st, k, margin, flags); }
    return D2S_OK;
}
// Instantiated for the dtype combinations the pipeline produces; anything else (-1) takes the generic kernel.
static int launch_fast(const WarpK &k, int rgb_dt, int depth_dt, int out_dt, bool half, int margin, d2s_stream_t st) {
#define D2S_FAST(RD, RT, DD, DT, OD, OT) if (rgb_dt == RD && depth_dt == DD && out_dt == OD) return launch_fast_t<RT, DT, OT>(k, half, margin, st)
    D2S_FAST(D2S_F16, __half, D2S_F16, __half, D2S_F32, float);
    D2S_FAST(D2S_F16, __half, D2S_F16, __half, D2S_U8, uint8_t);
    D2S_FAST(D2S_F16, __half, D2S_F16, __half, D2S_F16, __half);
    D2S_FAST(D2S_U8, uint8_t, D2S_F16, __half, D2S_F32, float);
    D2S_FAST(D2S_U8, uint8_t, D2S_F16, __half, D2S_U8, uint8_t);
    D2S_FAST(D2S_U8, uint8_t, D2S_F32, float, D2S_F32, float);
    D2S_FAST(D2S_U8, uint8_t, D2S_F32, float, D2S_U8, uint8_t);
    D2S_FAST(D2S_F32, float, D2S_F32, float, D2S_F32, float);
#undef D2S_FAST
    return -1;
}

}  // namespace d2s

using namespace d2s;

extern "C" int d2s_sbs_out_shape(int h, int w, int display_mode, int fill_16_9, int *out_h, int *out_w) {
    D2S_REQUIRE(h > 0 && w > 0 && out_h && out_w, "d2s_sbs_out_shape: bad arguments");
    D2S_REQUIRE(display_mode >= 0 && display_mode <= 3, "d2s_sbs_out_shape: display_mode %d", display_mode);
    int ph, pw, t, l;
    pad_geometry(h, w, fill_16_9, &ph, &pw, &t, &l);
    bool tab = display_mode == D2S_FULL_TAB || display_mode == D2S_HALF_TAB;
    bool half = display_mode == D2S_HALF_SBS || display_mode == D2S_HALF_TAB;
    *out_h = half ? ph : (tab ? 2 * ph : ph);
    *out_w = half ? pw : (tab ? pw : 2 * pw);
    return D2S_OK;
}

extern "C" int d2s_make_sbs(const d2s_warp_params *p, d2s_stream_t stream) {
    D2S_REQUIRE(p != nullptr, "d2s_make_sbs: null params");
    D2S_REQUIRE(p->rgb.base && p->out.base && p->depth, "d2s_make_sbs: null buffer");
    D2S_REQUIRE(p->h >= 1 && p->w >= 2, "d2s_make_sbs: eye size %dx%d unsupported (need w >= 2)", p->h, p->w);
    D2S_REQUIRE(p->display_mode >= 0 && p->display_mode <= 3, "d2s_make_sbs: display_mode %d", p->display_mode);
    D2S_REQUIRE(p->warp_mode == D2S_WARP_BILINEAR || p->warp_mode == D2S_WARP_GATHER, "d2s_make_sbs: warp_mode %d", p->warp_mode);
    D2S_REQUIRE(p->depth_h >= 1 && p->depth_w >= 1, "d2s_make_sbs: depth size %dx%d", p->depth_h, p->depth_w);
    WarpK k{};
    k.rgb = p->rgb.base; k.rsc = p->rgb.sc; k.rsy = p->rgb.sy; k.rsx = p->rgb.sx;
    k.out = p->out.base; k.osc = p->out.sc; k.osy = p->out.sy; k.osx = p->out.sx;
    k.depth = p->depth; k.dh = p->depth_h; k.dw = p->depth_w;
    k.lowres = !(p->depth_h == p->h && p->depth_w == p->w);
    k.dscale_h = (float)p->depth_h / (float)p->h;   // area_pixel_compute_scale, align_corners=False
    k.dscale_w = (float)p->depth_w / (float)p->w;
    k.h = p->h; k.w = p->w;
    pad_geometry(p->h, p->w, p->fill_16_9, &k.ph, &k.pw, &k.top, &k.left);
    k.tab = p->display_mode == D2S_FULL_TAB || p->display_mode == D2S_HALF_TAB;
    k.half = p->display_mode == D2S_HALF_SBS || p->display_mode == D2S_HALF_TAB;
    k.oh = k.half ? k.ph : (k.tab ? 2 * k.ph : k.ph);
    k.ow = k.half ? k.pw : (k.tab ? k.pw : 2 * k.pw);
    k.gather = p->warp_mode == D2S_WARP_GATHER;
    k.rgb_round = p->rgb_round_to_depth_dtype != 0;
    k.conv = (float)p->convergence; k.ratio = (float)p->depth_ratio;
    k.max_px = (float)(p->ipd_uv * (double)p->w);  // python: ipd_uv * W (double), then fp32 scalar
    k.strength = (float)0.05;
    k.two_over_wm1 = (float)(2.0 / (double)(p->w - 1));
    k.xstep = 2.0f / (float)(p->w - 1); k.xhalf = p->w / 2;
    k.ystep = p->h > 1 ? 2.0f / (float)(p->h - 1) : 0.f; k.yhalf = p->h / 2;
    k.idx_l = p->idx_left; k.idx_r = p->idx_right;

    // fast path: bilinear, no pad, no index taps, Full-SBS / Full-TAB / Half-SBS with even width
    {
        const bool half_sbs = k.half && !k.tab;
        const double smax = fmax(fabs(0.0 - p->convergence), fabs(1.0 - p->convergence)) * fabs(p->depth_ratio) * fabs(p->ipd_uv * p->w) * 0.05;
        const int margin = ((int)ceil(smax) + 2 + 15) / 16 * 16;   // multiple of 16: staged windows start 16-byte aligned for every dtype
        const char *nf = getenv("D2S_WARP_GENERIC");
        if (!(nf && nf[0] == '1') && !k.gather && !p->fill_16_9 && !k.idx_l && !k.idx_r && (!k.half || (half_sbs && k.w % 2 == 0)) && margin <= 128) {
            int rc = launch_fast(k, p->rgb.dtype, p->depth_dtype, p->out.dtype, half_sbs, margin, stream);
            if (rc != -1) { if (rc == D2S_OK) D2S_POST_LAUNCH(); return rc; }
        }
    }
    dim3 block(128, 2);
    dim3 grid(ceil_div(ceil_div(k.ow, 4), block.x), ceil_div(k.oh, block.y));
    int rc;
    switch (p->rgb.dtype) {
        case D2S_U8:  rc = launch_depth<uint8_t>(k, p->depth_dtype, p->out.dtype, grid, block, stream); break;
        case D2S_F16: rc = launch_depth<__half>(k, p->depth_dtype, p->out.dtype, grid, block, stream); break;
        case D2S_F32: rc = launch_depth<float>(k, p->depth_dtype, p->out.dtype, grid, block, stream); break;
        case D2S_BF16: rc = launch_depth<__nv_bfloat16>(k, p->depth_dtype, p->out.dtype, grid, block, stream); break;
        default: return set_error(D2S_ERR_UNSUPPORTED, "d2s_make_sbs: rgb dtype %d unsupported", p->rgb.dtype);
    }
    if (rc != D2S_OK) return rc;
    D2S_POST_LAUNCH();
    return D2S_OK;
}
