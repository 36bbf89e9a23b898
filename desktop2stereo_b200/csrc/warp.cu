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
#include <mutex>

#include "common.cuh"

namespace d2s {

static std::atomic<int> g_force_generic{0};   // d2s_debug_force_generic_warp: tests compare the fast path with the generic kernel

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
__device__ __forceinline__ float clamp255(float v) { return fminf(fmaxf(v, 0.f), 255.f); }

// One packed output pixel of the generic path (what warp_sbs_kernel computes per pixel): used by the fast kernel to redo the rare
// pixels its staged window cannot serve.  k is a __grid_constant__ kernel parameter, so taking its address costs no local copy.
template <typename RT, typename DT>
__device__ __noinline__ void generic_out_pixel(const WarpK *kp, int oy, int ox, float *v) {
    const WarpK &k = *kp;
    RowCtx rows[2];
    int cy0 = (k.half && k.tab) ? 2 * oy : oy;
    int ey0 = (k.tab ? (cy0 >= k.ph ? cy0 - k.ph : cy0) : cy0) - k.top;
    int cy1 = cy0 + 1;
    int ey1 = (k.tab ? (cy1 >= k.ph ? cy1 - k.ph : cy1) : cy1) - k.top;
    rows[0] = make_row(k, min(max(ey0, 0), k.h - 1));
    rows[1] = (k.half && k.tab) ? make_row(k, min(max(ey1, 0), k.h - 1)) : rows[0];
    float a[3];
    if (!k.half) cat_pixel<RT, DT>(k, rows, oy, ox, a);
    else {
        float b[3];
        if (k.tab) { cat_pixel<RT, DT>(k, rows, 2 * oy, ox, a); cat_pixel<RT, DT>(k, rows, 2 * oy + 1, ox, b); }
        else       { cat_pixel<RT, DT>(k, rows, oy, 2 * ox, a); cat_pixel<RT, DT>(k, rows, oy, 2 * ox + 1, b); }
#pragma unroll
        for (int c = 0; c < 3; ++c) a[c] = __fmul_rn(__fadd_rn(a[c], b[c]), 0.5f);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = clamp255(a[c]);
}

// ---- bulk (TMA, non-tensor) store of a contiguous shared-memory run to global memory ----
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group.read 0;" ::: "memory");
}

// Fast path v5.  One block = one source row segment of SEG pixels:
//   1. the rgb window the segment can reach (SEG + 2*margin pixels of 1-2 source rows) is staged in shared memory as
//      pixel-interleaved fp32 (r,g,b,-), so that a bilinear tap is one LDS.128;
//   2. thread t produces the output pixels t, t+128, t+256, t+384 of the segment for BOTH eyes (consecutive lanes = consecutive
//      pixels: tap loads and depth loads are contiguous across a warp, no bank conflicts, no padding) and parks them in a
//      shared-memory image of the output run;
//   3. the two output runs (left eye, right eye) of the segment are contiguous in the packed frame: each leaves with bulk
//      (TMA) stores, i.e. without per-thread address arithmetic and store instructions.
// Pixels whose taps leave the staged window (depth outside [0,1]) or need grid_sample's reflection are redone through the
// generic path (generic_out_pixel) before the run is stored.  Same arithmetic, same FMA order as the generic kernel =>
// bit-identical output (tests compare both against the oracle).
// OL: 0 = HWC contiguous (sx = 3, sc = 1), 1 = planar CHW (sx = 1).  flags bit 0: rgb rows may be staged with 16-byte loads,
// bit 2: the output runs may leave with bulk stores.
constexpr int kFastThreads = 128;

// two neighbouring pixels of one colour plane (pointer aligned to 2 elements), clamped / rounded exactly like load_rgb
template <typename RT, typename DT> __device__ __forceinline__ float2 load_px2(const RT *p, bool rgb_round);
template <> __device__ __forceinline__ float2 load_px2<__half, __half>(const __half *p, bool) {
    const uint32_t raw = __ldg((const uint32_t *)p);
    const __half2 v = __hmin2(__hmax2(*(const __half2 *)&raw, __float2half2_rn(0.f)), __float2half2_rn(255.f));   // clamp in fp16: exact
    return __half22float2(v);
}
template <> __device__ __forceinline__ float2 load_px2<uint8_t, __half>(const uint8_t *p, bool) {
    const uint16_t raw = __ldg((const uint16_t *)p);
    return make_float2((float)(raw & 0xffu), (float)(raw >> 8));
}
template <> __device__ __forceinline__ float2 load_px2<uint8_t, float>(const uint8_t *p, bool) {
    const uint16_t raw = __ldg((const uint16_t *)p);
    return make_float2((float)(raw & 0xffu), (float)(raw >> 8));
}
template <> __device__ __forceinline__ float2 load_px2<float, float>(const float *p, bool) {
    const float2 v = __ldg((const float2 *)p);
    return make_float2(stage_value<float, float>(v.x, false), stage_value<float, float>(v.y, false));
}

template <typename OT> __device__ __forceinline__ OT out_value(float v) { return from_f32<OT>(clamp255(v)); }   // final clamp, depth.py:2184
template <> __device__ __forceinline__ uint8_t out_value<uint8_t>(float v) { return (uint8_t)__float2int_rn(clamp255(v)); }

// step 2 for one thread; TWO: the second source row contributes (block-uniform, so the kernel branches once, not per pixel)
template <typename DT, typename OT, int HALF, int OL, bool TWO>
__device__ __forceinline__ unsigned fast_pixels(const WarpK &k, const float4 *__restrict__ s0, const float4 *__restrict__ s1, OT *__restrict__ s_out,
                                                const float *dv, const RowCtx &row, int seg0, int lo, int tw, int tid) {
    constexpr int THREADS = kFastThreads, OPX = 4, OSEG = THREADS * OPX;
    const float span = (float)(k.w - 1);
    const unsigned last_ok = (unsigned)(tw - 1);     // window index a is usable iff 0 <= a and a + 1 <= tw - 1
    unsigned bad = 0;                                // bit (2*j + e): output pixel j of eye e must be redone by the generic path
#pragma unroll
    for (int j = 0; j < OPX; ++j) {
        const int q = j * THREADS + tid;             // output pixel inside the segment
        float acc[2][3];                             // Half-SBS: the first pixel of the pair, per eye
#pragma unroll
        for (int hp = 0; hp < (HALF ? 2 : 1); ++hp) {
            const int xr = seg0 + (HALF ? 2 * q + hp : q);
            const int x = min(xr, k.w - 1);
            // shift chain, depth.py:2143-2147, :2154 (identical to eye_pixel)
            const float d = round_to<DT>(__fsub_rn(dv[HALF ? 2 * j + hp : j], k.conv));
            const float inv = round_to<DT>(__fmul_rn(-d, k.ratio));
            float s = round_to<DT>(__fmul_rn(inv, k.max_px));
            s = round_to<DT>(__fmul_rn(s, k.strength));
            const float sn = round_to<DT>(__fmul_rn(s, k.two_over_wm1));
            const float xs = linspace_pm1(x, k.w, k.xstep, k.xhalf);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                // grid_sampler source index (same operations as source_index(); the reflection case is left to the generic path)
                const float gx = e ? __fsub_rn(xs, sn) : __fadd_rn(xs, sn);
                const float in = fabsf(__fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), span));
                const float ix = fminf(span, in);    // in >= 0
                const int ix0 = __float2int_rz(ix);  // ix >= 0 after the clip: truncation == floor
                const float fx0 = (float)ix0;
                const float wx0 = __fsub_rn(__fadd_rn(fx0, 1.f), ix), wx1 = __fsub_rn(ix, fx0);   // (float)(ix0 + 1) == fx0 + 1 exactly (ix0 < 2^24)
                // Both taps ix0, ix0 + 1 must lie inside the staged window [lo, lo + tw).  One unsigned compare covers a < 0,
                // a + 1 >= tw AND the reflection case: in >= span gives ix0 = w - 1, i.e. a >= tw - 1 (the window ends at or before w).
                const unsigned a = (unsigned)(ix0 - lo);
                const bool oob = a >= last_ok;
                const unsigned ac = oob ? 0u : a;
                const float4 p00 = s0[ac], p01 = s0[ac + 1];
                const float nw = __fmul_rn(wx0, row.wy0), ne = __fmul_rn(wx1, row.wy0);
                float cx = __fmaf_rn(p01.x, ne, __fmaf_rn(p00.x, nw, 0.f));
                float cy = __fmaf_rn(p01.y, ne, __fmaf_rn(p00.y, nw, 0.f));
                float cz = __fmaf_rn(p01.z, ne, __fmaf_rn(p00.z, nw, 0.f));
                if (TWO) {
                    const float4 p10 = s1[ac], p11 = s1[ac + 1];
                    const float sw = __fmul_rn(wx0, row.wy1), se = __fmul_rn(wx1, row.wy1);
                    cx = __fmaf_rn(p11.x, se, __fmaf_rn(p10.x, sw, cx));
                    cy = __fmaf_rn(p11.y, se, __fmaf_rn(p10.y, sw, cy));
                    cz = __fmaf_rn(p11.z, se, __fmaf_rn(p10.z, sw, cz));
                }
                if (oob && xr < k.w) bad |= 1u << (2 * j + e);
                if (HALF && hp == 0) { acc[e][0] = cx; acc[e][1] = cy; acc[e][2] = cz; continue; }
                if (HALF) {   // F.interpolate(mode="area") at an exact 2:1 ratio: (a + b) / 2
                    cx = __fmul_rn(__fadd_rn(acc[e][0], cx), 0.5f); cy = __fmul_rn(__fadd_rn(acc[e][1], cy), 0.5f); cz = __fmul_rn(__fadd_rn(acc[e][2], cz), 0.5f);
                }
                OT *o = s_out + e * 3 * OSEG;
                if (OL == 0) { o[q * 3] = out_value<OT>(cx); o[q * 3 + 1] = out_value<OT>(cy); o[q * 3 + 2] = out_value<OT>(cz); }
                else { o[q] = out_value<OT>(cx); o[OSEG + q] = out_value<OT>(cy); o[2 * OSEG + q] = out_value<OT>(cz); }
            }
        }
    }
    return bad;
}

template <typename RT, typename DT, typename OT, int HALF, int OL>
__global__ void __launch_bounds__(kFastThreads) warp_sbs_fast_kernel(const __grid_constant__ WarpK k, int margin, int flags) {
    constexpr int THREADS = kFastThreads;
    constexpr int OPX = 4;                           // output pixels per thread and eye
    constexpr int OSEG = THREADS * OPX;              // output pixels per block and eye (512)
    constexpr int SEG = HALF ? 2 * OSEG : OSEG;      // source pixels per block
    extern __shared__ __align__(16) uint8_t s_raw[];
    const int pitch4 = SEG + 2 * margin;
    float4 *s_px = (float4 *)s_raw;                                  // [rows(2)][pitch4]
    OT *s_out = (OT *)(s_raw + (size_t)2 * pitch4 * sizeof(float4));  // [eye][3 * OSEG] (HWC: q*3+c, planar: c*OSEG+q)
    const int y = blockIdx.y;
    const int seg0 = blockIdx.x * SEG;
    const int lo = max(seg0 - margin, 0), hi = min(seg0 + SEG + margin, k.w);   // staged source columns [lo, hi); margin % 16 == 0
    const int tw = hi - lo;
    const RowCtx row = make_row(k, y);
    const bool two = row.ok1 && row.wy1 != 0.f;      // second source row contributes (block-uniform)
    const int tid = threadIdx.x;
    // ---- 0. this thread's depth values: issued first, so that their DRAM round trip overlaps the staging loads
    float dv[HALF ? 2 * OPX : OPX];
    {
        const DT *drow = (const DT *)k.depth + (size_t)y * k.w;
#pragma unroll
        for (int j = 0; j < OPX; ++j) {
            const int q = j * THREADS + tid;
            if (!HALF) {
                const int x = min(seg0 + q, k.w - 1);
                dv[j] = k.lowres ? load_depth<DT>(k, y, x) : to_f32<DT>(__ldg(drow + x));
            } else {
                const int x = min(seg0 + 2 * q, k.w - 2);              // w is even on this path
                dv[2 * j] = k.lowres ? load_depth<DT>(k, y, x) : to_f32<DT>(__ldg(drow + x));
                dv[2 * j + 1] = k.lowres ? load_depth<DT>(k, y, x + 1) : to_f32<DT>(__ldg(drow + x + 1));
            }
        }
    }
    // ---- 1. staging: a thread converts TWO neighbouring pixels per step (one 2-pixel load per colour plane), so that a warp's
    //         16-byte shared-memory stores are (almost) contiguous — wider per-thread loads make those stores collide 8-way, and
    //         every warp of the block waits for them at the barrier below
    {
        const bool rr = k.rgb_round;
        const int np2 = (flags & 1) ? tw / 2 : 0;
        const RT *src = (const RT *)k.rgb + (long long)row.iy0 * k.rsy + (long long)lo * k.rsx;
        for (int r = 0; r < (two ? 2 : 1); ++r) {
            float4 *dst = s_px + r * pitch4;
            const RT *srow = src + r * k.rsy;
            for (int i = tid; i < np2; i += THREADS) {
                float2 c0 = load_px2<RT, DT>(srow + 2 * i, rr), c1 = load_px2<RT, DT>(srow + k.rsc + 2 * i, rr), c2 = load_px2<RT, DT>(srow + 2 * k.rsc + 2 * i, rr);
                dst[2 * i] = make_float4(c0.x, c1.x, c2.x, 0.f);
                dst[2 * i + 1] = make_float4(c0.y, c1.y, c2.y, 0.f);
            }
            for (int x = 2 * np2 + tid; x < tw; x += THREADS) {
                const RT *p = srow + (long long)x * k.rsx;
                dst[x] = make_float4(stage_value<RT, DT>(to_f32<RT>(__ldg(p)), rr), stage_value<RT, DT>(to_f32<RT>(__ldg(p + k.rsc)), rr),
                                     stage_value<RT, DT>(to_f32<RT>(__ldg(p + 2 * k.rsc)), rr), 0.f);
            }
        }
    }
    __syncthreads();
    const int nsrc = min(SEG, k.w - seg0);           // valid source pixels of this segment
    const int nout = HALF ? nsrc / 2 : nsrc;         // valid output pixels per eye
    // ---- 2. both eyes of this thread's pixels -> s_out
    const unsigned bad = two ? fast_pixels<DT, OT, HALF, OL, true>(k, s_px, s_px + pitch4, s_out, dv, row, seg0, lo, tw, tid)
                             : fast_pixels<DT, OT, HALF, OL, false>(k, s_px, s_px + pitch4, s_out, dv, row, seg0, lo, tw, tid);
    // ---- 2b. the rare pixels the staged window could not serve: generic path
    if (bad) {
#pragma unroll 1
        for (int b = 0; b < 2 * OPX; ++b) {
            if (!((bad >> b) & 1u)) continue;
            const int j = b >> 1, e = b & 1, q = j * THREADS + tid;
            if (q >= nout) continue;
            const int oy = k.tab ? e * k.h + y : y;
            const int ox = (k.tab ? 0 : e * (HALF ? k.w / 2 : k.w)) + (HALF ? seg0 / 2 : seg0) + q;
            float v[3];
            generic_out_pixel<RT, DT>(&k, oy, ox, v);
            OT *o = s_out + e * 3 * OSEG;
            if (OL == 0) { o[q * 3] = from_f32<OT>(v[0]); o[q * 3 + 1] = from_f32<OT>(v[1]); o[q * 3 + 2] = from_f32<OT>(v[2]); }
            else { o[q] = from_f32<OT>(v[0]); o[OSEG + q] = from_f32<OT>(v[1]); o[2 * OSEG + q] = from_f32<OT>(v[2]); }
        }
    }
    // ---- 3. the output runs leave
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes of s_out -> visible to the bulk-copy engine
    __syncthreads();
    const int ox0 = HALF ? seg0 / 2 : seg0;
    const bool bulk = (flags & 4) != 0 && ((nout * (int)sizeof(OT)) % 16) == 0;   // flags bit 2: every run start is 16-byte aligned
    if (bulk) {
        if (tid < 2) {
            const int e = tid;
            const int oy = k.tab ? e * k.h + y : y;
            const int ox = (k.tab ? 0 : e * (HALF ? k.w / 2 : k.w)) + ox0;
            OT *g = (OT *)k.out + (long long)oy * k.osy + (long long)ox * (OL == 0 ? 3 : 1);
            const OT *sm = s_out + e * 3 * OSEG;
            if (OL == 0) bulk_store(g, sm, (uint32_t)(nout * 3 * sizeof(OT)));
            else {
#pragma unroll
                for (int c = 0; c < 3; ++c) bulk_store(g + c * k.osc, sm + c * OSEG, (uint32_t)(nout * sizeof(OT)));
            }
            bulk_commit_wait();                      // the run has been READ out of shared memory before the block may exit
        }
    } else {
#pragma unroll 1
        for (int e = 0; e < 2; ++e) {
            const int oy = k.tab ? e * k.h + y : y;
            const int ox = (k.tab ? 0 : e * (HALF ? k.w / 2 : k.w)) + ox0;
            OT *g = (OT *)k.out + (long long)oy * k.osy + (long long)ox * (OL == 0 ? 3 : 1);
            const OT *sm = s_out + e * 3 * OSEG;
            if (OL == 0) { for (int i = tid; i < nout * 3; i += THREADS) g[i] = sm[i]; }
            else {
                for (int c = 0; c < 3; ++c)
                    for (int i = tid; i < nout; i += THREADS) g[c * k.osc + i] = sm[c * OSEG + i];
            }
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
    const int seg = half ? 1024 : 512;               // source pixels per block (512 output pixels per eye)
    dim3 grid(ceil_div(k.w, seg), k.h);
    const int padded = seg + 2 * margin;
    const size_t smem = (size_t)2 * padded * sizeof(float4) + (size_t)2 * 3 * 512 * sizeof(OT);
    const bool hwc = k.osx == 3 && k.osc == 1, chw = k.osx == 1;
    if (!hwc && !chw) return -1;
    int flags = 0;
    // 16-byte staging loads: contiguous pixels, every row/plane start 16-byte aligned (lo is a multiple of 8... of 16 for u8)
    const size_t es = sizeof(RT), oes = sizeof(OT);
    if (k.rsx == 1 && ((uintptr_t)k.rgb % 16) == 0 && (k.rsy * (long long)es) % 16 == 0 && (k.rsc * (long long)es) % 16 == 0 && (margin * es) % 16 == 0) flags |= 1;
    // bulk stores: every output run starts 16-byte aligned (base, row pitch, the right eye's offset, the 512-pixel segment pitch;
    // planar: the plane pitch too)
    {
        const long long ew = half ? k.w / 2 : k.w;
        const long long px = hwc ? 3 : 1;
        bool ok = ((uintptr_t)k.out % 16) == 0 && (k.osy * (long long)oes) % 16 == 0 && (512 * px * (long long)oes) % 16 == 0 &&
                  (k.tab || (ew * px * (long long)oes) % 16 == 0) && (hwc || (k.osc * (long long)oes) % 16 == 0);
        if (ok) flags |= 4;
    }
    static std::once_flag once;                      // (one per <RT, DT, OT> instantiation) Half-SBS windows exceed the 48 KB default
    std::call_once(once, [] {
        cudaFuncSetAttribute((const void *)warp_sbs_fast_kernel<RT, DT, OT, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute((const void *)warp_sbs_fast_kernel<RT, DT, OT, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute((const void *)warp_sbs_fast_kernel<RT, DT, OT, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute((const void *)warp_sbs_fast_kernel<RT, DT, OT, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    });
    if (half) { if (hwc) D2S_LAUNCH((warp_sbs_fast_kernel<RT, DT, OT, 1, 0>), grid, kFastThreads, smem, st, k, margin, flags);
                else     D2S_LAUNCH((warp_sbs_fast_kernel<RT, DT, OT, 1, 1>), grid, kFastThreads, smem, st, k, margin, flags); }
    else      { if (hwc) D2S_LAUNCH((warp_sbs_fast_kernel<RT, DT, OT, 0, 0>), grid, kFastThreads, smem, st, k, margin, flags);
                else     D2S_LAUNCH((warp_sbs_fast_kernel<RT, DT, OT, 0, 1>), grid, kFastThreads, smem, st, k, margin, flags); }
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

extern "C" int d2s_debug_force_generic_warp(int on) {
    g_force_generic.store(on ? 1 : 0, std::memory_order_relaxed);
    return D2S_OK;
}

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
        if (!g_force_generic.load(std::memory_order_relaxed) && !k.gather && !p->fill_16_9 && !k.idx_l && !k.idx_r && (!k.half || (half_sbs && k.w % 2 == 0)) && margin <= 128) {
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
