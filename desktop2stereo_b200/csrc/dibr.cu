// Occlusion-aware stereo rendering (DIBR with disocclusion confidence + push-pull inpaint) — SURVEY §8f N2.
//
// Replaces, as a tensor-in / tensor-out CUDA kernel, what the reference does in its OpenGL viewer: the fragment shader
// FRAGMENT_SHADER of reference viewer.py:386-631 run once per eye view (viewer.py:2680-2760: u_eye_offset = -/+ ipd/2,
// u_depth_strength = 0.1 * depth_ratio, viewer.py:1334, 2686).  Still a BACKWARD warp — every output pixel gathers — but with the
// reference's occlusion handling: 3-tap depth smoothing along the parallax direction (:545-549), non-linear depth shaping (:554),
// edge falloff of the parallax (:559-563), a 2-tap depth-jump "disocclusion confidence" (:421-435), a directional push-pull
// inpaint that only accepts background samples (:437-506) blended in by that confidence (:567-576), border alpha (:582-583),
// optional feathering (:586-613) and rounded corners (:617-626).  texture() = GL_LINEAR + GL_REPEAT (moderngl's defaults; the
// reference sets neither), evaluated with exact fp32 weights.
//
// One thread = one output pixel of one eye; HBM/L2-gather bound, no tensor cores.  Strict fp32 in the shader's operation order,
// compiled with -fmad=false, so the result is bit-identical to oracle/dibr_oracle.c (tests/test_dibr_gpu.py).  Texel coordinates
// inside the image skip the GL_REPEAT wrap (fmodf) — same result, ~3x fewer instructions per fetch.  exp() weights and
// cos/sin(roll) arrive as fp32 values computed on the host in double.  u_resolution is a parameter: the reference never sets it
// (see the oracle's header), the default used by the host wrapper is the eye view's size, as the shader's comment says.
#include "common.cuh"

namespace d2s {

struct DibrK {
    const void *rgb; long long rsc, rsy, rsx; int rgb_dtype;
    const void *depth; int depth_dtype;
    void *out; long long osc, osy, osx; int out_dtype;
    int w, h, vw, vh, tab;
    float res_x, res_y, ipd_half, depth_strength, convergence, c, s;
    int search;
    float tol, blur_radius;
    float w1[33], w2[33];
    int feather; float feather_width, corner_radius;
};

__device__ __noinline__ int wrap_texel_slow(float f, int n) {   // GL_REPEAT for a texel coordinate outside [0, n)
    float r = fmodf(f, (float)n);
    if (r < 0.f) r = __fadd_rn(r, (float)n);
    int i = (int)r;
    return i >= n ? n - 1 : i;
}
__device__ __forceinline__ int wrap_texel(float f, int n) {     // f is integral (a floorf result): inside the image nothing wraps
    return (f >= 0.f && f < (float)n) ? (int)f : wrap_texel_slow(f, n);
}

struct Taps { int x0, x1, y0, y1; float fx, fy; };
__device__ __forceinline__ Taps make_taps(float u, float v, int w, int h) {
    Taps t;
    const float x = __fsub_rn(__fmul_rn(u, (float)w), 0.5f), y = __fsub_rn(__fmul_rn(v, (float)h), 0.5f);
    const float x0 = floorf(x), y0 = floorf(y);
    t.fx = __fsub_rn(x, x0); t.fy = __fsub_rn(y, y0);
    t.x0 = wrap_texel(x0, w); t.x1 = wrap_texel(__fadd_rn(x0, 1.f), w);
    t.y0 = wrap_texel(y0, h); t.y1 = wrap_texel(__fadd_rn(y0, 1.f), h);
    return t;
}
__device__ __forceinline__ float bilerp(float t00, float t10, float t01, float t11, float fx, float fy) {
    const float gx = __fsub_rn(1.f, fx), gy = __fsub_rn(1.f, fy);
    const float top = __fadd_rn(__fmul_rn(t00, gx), __fmul_rn(t10, fx));
    const float bot = __fadd_rn(__fmul_rn(t01, gx), __fmul_rn(t11, fx));
    return __fadd_rn(__fmul_rn(top, gy), __fmul_rn(bot, fy));
}

template <typename DT>
__device__ __forceinline__ float tex_depth_t(const DibrK &k, const Taps &t) {
    const DT *d = (const DT *)k.depth;
    const DT *r0 = d + (size_t)t.y0 * k.w, *r1 = d + (size_t)t.y1 * k.w;
    return bilerp(to_f32<DT>(__ldg(r0 + t.x0)), to_f32<DT>(__ldg(r0 + t.x1)), to_f32<DT>(__ldg(r1 + t.x0)), to_f32<DT>(__ldg(r1 + t.x1)), t.fx, t.fy);
}
template <typename DT>
__device__ __forceinline__ float tex_depth(const DibrK &k, float u, float v) { return tex_depth_t<DT>(k, make_taps(u, v, k.w, k.h)); }

// colour texel as the normalised texture holds it: value / 255 (the reference uploads u8 RGB, viewer.py:2385)
template <typename RT>
__device__ __forceinline__ float texel(const DibrK &k, int c, int y, int x) {
    return __fdiv_rn(to_f32<RT>(__ldg((const RT *)k.rgb + c * k.rsc + (long long)y * k.rsy + (long long)x * k.rsx)), 255.f);
}
template <typename RT>
__device__ __forceinline__ float3 tex_color_t(const DibrK &k, const Taps &t) {
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
        o[c] = bilerp(texel<RT>(k, c, t.y0, t.x0), texel<RT>(k, c, t.y0, t.x1), texel<RT>(k, c, t.y1, t.x0), texel<RT>(k, c, t.y1, t.x1), t.fx, t.fy);
    return make_float3(o[0], o[1], o[2]);
}
template <typename RT>
__device__ __forceinline__ float3 tex_color(const DibrK &k, float u, float v) { return tex_color_t<RT>(k, make_taps(u, v, k.w, k.h)); }

__device__ __forceinline__ float sstep(float e0, float e1, float x) {
    const float t = fminf(fmaxf(__fdiv_rn(__fsub_rn(x, e0), __fsub_rn(e1, e0)), 0.f), 1.f);
    return __fmul_rn(__fmul_rn(t, t), __fsub_rn(3.f, __fmul_rn(2.f, t)));
}

template <typename RT, typename DT>
__device__ __noinline__ float3 push_pull(const DibrK &k, float ux, float uy, float center, float pdx, float pdy, float psx, float psy, float sweep_sign) {
    float bx = 0.f, by = 0.f, bz = 0.f, bw = 0.f;
    const float swx = __fmul_rn(__fmul_rn(pdx, psx), sweep_sign), swy = __fmul_rn(__fmul_rn(pdy, psx), sweep_sign);
    const float thr = __fadd_rn(center, k.tol);
    for (int i = 1; i <= k.search; ++i) {            // phase 1: sweep towards the side the background is revealed from
        const float qx = __fadd_rn(ux, __fmul_rn(swx, (float)i)), qy = __fadd_rn(uy, __fmul_rn(swy, (float)i));
        if (qx < 0.f || qy < 0.f || qx > 1.f || qy > 1.f) continue;
        const Taps tq = make_taps(qx, qy, k.w, k.h);       // depth and colour are fetched at the same coordinate
        const float sdi = __fsub_rn(1.f, tex_depth_t<DT>(k, tq));
        if (sdi > thr) {
            const float3 sc = tex_color_t<RT>(k, tq);
            const float dw = __fadd_rn(1.f, __fmul_rn(__fsub_rn(sdi, center), 10.f));
            const float wgt = __fmul_rn(k.w1[i], dw);
            bx = __fadd_rn(bx, __fmul_rn(sc.x, wgt)); by = __fadd_rn(by, __fmul_rn(sc.y, wgt)); bz = __fadd_rn(bz, __fmul_rn(sc.z, wgt));
            bw = __fadd_rn(bw, wgt);
            if (bw > 5.f) break;
        }
    }
    if (bw < 2.f) {                                  // phase 2: opposite sweep
        for (int i = 1; i <= k.search; ++i) {
            const float qx = __fsub_rn(ux, __fmul_rn(swx, (float)i)), qy = __fsub_rn(uy, __fmul_rn(swy, (float)i));
            if (qx < 0.f || qy < 0.f || qx > 1.f || qy > 1.f) continue;
            const Taps tq = make_taps(qx, qy, k.w, k.h);
            const float sdi = __fsub_rn(1.f, tex_depth_t<DT>(k, tq));
            if (sdi > thr) {
                const float3 sc = tex_color_t<RT>(k, tq);
                const float wgt = k.w2[i];
                bx = __fadd_rn(bx, __fmul_rn(sc.x, wgt)); by = __fadd_rn(by, __fmul_rn(sc.y, wgt)); bz = __fadd_rn(bz, __fmul_rn(sc.z, wgt));
                bw = __fadd_rn(bw, wgt);
            }
        }
    }
    if (bw > 0.01f) {                                // phase 3: 3-tap vertical blur
        float ax = __fmul_rn(__fdiv_rn(bx, bw), 0.5f), ay = __fmul_rn(__fdiv_rn(by, bw), 0.5f), az = __fmul_rn(__fdiv_rn(bz, bw), 0.5f);
        float vwgt = 0.5f;
        const float thr2 = __fadd_rn(center, __fmul_rn(k.tol, 0.5f));
#pragma unroll
        for (int dy = -1; dy <= 1; dy += 2) {
            const float vy = __fadd_rn(uy, __fmul_rn(__fmul_rn((float)dy, psy), k.blur_radius));
            if (vy >= 0.f && vy <= 1.f) {
                const float vdi = __fsub_rn(1.f, tex_depth<DT>(k, ux, vy));
                if (vdi > thr2) {
                    const float3 sc = tex_color<RT>(k, ux, vy);
                    ax = __fadd_rn(ax, __fmul_rn(sc.x, 0.25f)); ay = __fadd_rn(ay, __fmul_rn(sc.y, 0.25f)); az = __fadd_rn(az, __fmul_rn(sc.z, 0.25f));
                    vwgt = __fadd_rn(vwgt, 0.25f);
                }
            }
        }
        return make_float3(__fdiv_rn(ax, vwgt), __fdiv_rn(ay, vwgt), __fdiv_rn(az, vwgt));
    }
    return tex_color<RT>(k, ux, uy);
}

// One output pixel (view column j, view row i from the top, eye e).  DEFER: a pixel whose disocclusion confidence calls for the
// inpaint sweep is not finished here — it returns true and the caller queues it for the dense second pass.
template <typename RT, typename DT, typename OT, bool DEFER>
__device__ __forceinline__ bool dibr_pixel(const DibrK &k, int j, int i, int e) {
    const float eye = e ? k.ipd_half : -k.ipd_half;
    const float uvx = __fdiv_rn(__fadd_rn((float)j, 0.5f), (float)k.vw);
    const float uvy = __fdiv_rn(__fadd_rn((float)(k.vh - 1 - i), 0.5f), (float)k.vh);     // gl_FragCoord.y counts from the bottom
    const float fx = uvx, fy = __fsub_rn(1.f, uvy);                                         // flipped_uv
    const float psx = __fdiv_rn(1.f, k.res_x), psy = __fdiv_rn(1.f, k.res_y);
    const float sg = eye > 0.f ? 1.f : (eye < 0.f ? -1.f : 0.f);
    const float pdx = __fmul_rn(k.c, sg), pdy = __fmul_rn(k.s, sg);
    const float sweep_sign = eye > 0.f ? -1.f : 1.f;
    // 3-tap depth smoothing along the parallax direction (viewer.py:545-549)
    const float dsx = __fmul_rn(__fmul_rn(pdx, psx), 1.5f), dsy = __fmul_rn(__fmul_rn(pdy, psy), 1.5f);
    const float d0 = tex_depth<DT>(k, fx, fy), dm = tex_depth<DT>(k, __fsub_rn(fx, dsx), __fsub_rn(fy, dsy)), dp = tex_depth<DT>(k, __fadd_rn(fx, dsx), __fadd_rn(fy, dsy));
    const float depth = __fadd_rn(__fadd_rn(__fmul_rn(d0, 0.7f), __fmul_rn(dm, 0.15f)), __fmul_rn(dp, 0.15f));
    const float depth_inv = -depth;
    const float shaped = __fmul_rn(depth_inv, __fadd_rn(1.f, __fmul_rn(0.35f, __fsub_rn(1.f, depth))));   // viewer.py:554
    const float shift = __fadd_rn(shaped, k.convergence);
    const float margin = 0.05f;
    const float falloff = __fmul_rn(sstep(0.f, margin, fx), sstep(1.f, __fsub_rn(1.f, margin), fx));      // viewer.py:560-562
    const float px = __fmul_rn(__fmul_rn(__fmul_rn(eye, shift), k.depth_strength), falloff);
    const float sx = __fsub_rn(fx, __fmul_rn(px, k.c)), sy = __fsub_rn(fy, __fmul_rn(px, k.s));
    // disocclusion confidence (viewer.py:421-435)
    float conf;
    if (sx < 0.f || sx > 1.f || sy < 0.f || sy > 1.f) conf = 1.f;
    else {
        const float s2x = __fmul_rn(__fmul_rn(pdx, psx), 2.f), s2y = __fmul_rn(__fmul_rn(pdy, psy), 2.f);
        const float dl = tex_depth<DT>(k, __fsub_rn(fx, s2x), __fsub_rn(fy, s2y)), dr = tex_depth<DT>(k, __fadd_rn(fx, s2x), __fadd_rn(fy, s2y));
        conf = sstep(0.04f, 0.10f, fabsf(__fsub_rn(dl, dr)));
    }
    if (DEFER && conf > 0.001f) return true;
    float3 col = tex_color<RT>(k, sx, sy);
    if (conf > 0.001f) {
        const float3 f = push_pull<RT, DT>(k, fx, fy, depth_inv, pdx, pdy, psx, psy, sweep_sign);
        const float ic = __fsub_rn(1.f, conf);
        col.x = __fadd_rn(__fmul_rn(col.x, ic), __fmul_rn(f.x, conf));
        col.y = __fadd_rn(__fmul_rn(col.y, ic), __fmul_rn(f.y, conf));
        col.z = __fadd_rn(__fmul_rn(col.z, ic), __fmul_rn(f.z, conf));
    }
    const float bxa = __fmul_rn(sstep(-0.001f, 0.001f, sx), sstep(1.001f, 0.999f, sx));
    const float bya = __fmul_rn(sstep(-0.001f, 0.001f, sy), sstep(1.001f, 0.999f, sy));
    float alpha = fminf(bxa, bya);
    if (k.feather) {
        const float f = k.feather_width;
        float fo = __fmul_rn(__fmul_rn(__fmul_rn(sstep(0.f, f, uvx), sstep(0.f, f, __fsub_rn(1.f, uvx))), sstep(0.f, f, uvy)), sstep(0.f, f, __fsub_rn(1.f, uvy)));
        fo = powf(fo, 0.7f);
        col.x = __fmul_rn(col.x, fo); col.y = __fmul_rn(col.y, fo); col.z = __fmul_rn(col.z, fo);
    }
    {   // rounded corners: Inigo Quilez' rounded-box SDF (viewer.py:617-626)
        const float r = k.corner_radius;
        const float dx = __fadd_rn(__fsub_rn(fabsf(__fsub_rn(uvx, 0.5f)), 0.5f), r), dy = __fadd_rn(__fsub_rn(fabsf(__fsub_rn(uvy, 0.5f)), 0.5f), r);
        const float mx = fmaxf(dx, 0.f), my = fmaxf(dy, 0.f);
        const float sdf = __fsub_rn(__fadd_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(mx, mx), __fmul_rn(my, my))), fminf(fmaxf(dx, dy), 0.f)), r);
        alpha = fminf(alpha, __fsub_rn(1.f, sstep(0.f, 0.01f, sdf)));
    }
    // the frame the viewer presents: colour over a black background, as 0..255 like make_sbs
    const int oy = k.tab ? e * k.vh + i : i, ox = k.tab ? j : e * k.vw + j;
    OT *o = (OT *)k.out + (long long)oy * k.osy + (long long)ox * k.osx;
    const float v[3] = {col.x, col.y, col.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c * k.osc] = from_f32<OT>(__fmul_rn(fminf(fmaxf(__fmul_rn(v[c], alpha), 0.f), 1.f), 255.f));
    return false;
}

// single pass: every pixel finished by its own thread (no workspace)
template <typename RT, typename DT, typename OT>
__global__ void __launch_bounds__(256) dibr_kernel(const __grid_constant__ DibrK k) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k.vw) return;
    dibr_pixel<RT, DT, OT, false>(k, j, blockIdx.y, blockIdx.z);
}

// Two passes (with a caller workspace).  The inpaint sweep runs up to 2 x search_radius (depth + colour) fetches for the few
// per cent of pixels that sit on a depth edge; in the single-pass kernel every warp that contains ONE such pixel walks the whole sweep
// with 31 idle lanes.  Pass A finishes the ordinary pixels and queues the others (one warp-aggregated atomic per warp); pass B
// runs the queued pixels densely, one per thread, persistent blocks striding over the queue.  Same per-pixel function, so the
// frame is bit-identical to the single-pass kernel's.
template <typename RT, typename DT, typename OT>
__global__ void __launch_bounds__(256) dibr_pass_a_kernel(const __grid_constant__ DibrK k, uint32_t *__restrict__ queue, uint32_t *__restrict__ count) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, e = blockIdx.z;
    const bool defer = j < k.vw && dibr_pixel<RT, DT, OT, true>(k, j, i, e);
    const unsigned m = __ballot_sync(0xffffffffu, defer);
    if (m) {
        const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(count, (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (defer) queue[base + __popc(m & ((1u << lane) - 1u))] = ((uint32_t)e << 31) | ((uint32_t)i << 16) | (uint32_t)j;
    }
}
template <typename RT, typename DT, typename OT>
__global__ void __launch_bounds__(128) dibr_pass_b_kernel(const __grid_constant__ DibrK k, const uint32_t *__restrict__ queue, const uint32_t *__restrict__ count) {
    const uint32_t n = *count;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const uint32_t v = queue[q];
        dibr_pixel<RT, DT, OT, false>(k, (int)(v & 0xffffu), (int)((v >> 16) & 0x7fffu), (int)(v >> 31));
    }
}

template <typename RT, typename DT, typename OT>
static int launch_dibr_t(const DibrK &k, dim3 grid, uint32_t *ws, d2s_stream_t st) {
    if (!ws) { D2S_LAUNCH((dibr_kernel<RT, DT, OT>), grid, 256, 0, st, k); return D2S_OK; }
    D2S_CHECK_CUDA(cudaMemsetAsync(ws, 0, sizeof(uint32_t), (cudaStream_t)st));     // ws[0] = queue length, ws[4..] = queue
    D2S_LAUNCH((dibr_pass_a_kernel<RT, DT, OT>), grid, 256, 0, st, k, ws + 4, ws);
    D2S_LAUNCH((dibr_pass_b_kernel<RT, DT, OT>), 8 * kNumSMs, 128, 0, st, k, ws + 4, ws);
    return D2S_OK;
}
template <typename RT, typename DT>
static int launch_dibr_out(const DibrK &k, dim3 grid, uint32_t *ws, d2s_stream_t st) {
    switch (k.out_dtype) {
        case D2S_F32: return launch_dibr_t<RT, DT, float>(k, grid, ws, st);
        case D2S_U8: return launch_dibr_t<RT, DT, uint8_t>(k, grid, ws, st);
        case D2S_F16: return launch_dibr_t<RT, DT, __half>(k, grid, ws, st);
        default: return set_error(D2S_ERR_UNSUPPORTED, "d2s_make_sbs_dibr: out dtype %d", k.out_dtype);
    }
}
template <typename RT>
static int launch_dibr_depth(const DibrK &k, dim3 grid, uint32_t *ws, d2s_stream_t st) {
    switch (k.depth_dtype) {
        case D2S_F32: return launch_dibr_out<RT, float>(k, grid, ws, st);
        case D2S_F16: return launch_dibr_out<RT, __half>(k, grid, ws, st);
        default: return set_error(D2S_ERR_UNSUPPORTED, "d2s_make_sbs_dibr: depth dtype %d", k.depth_dtype);
    }
}

}  // namespace d2s

using namespace d2s;

extern "C" int d2s_dibr_out_shape(int h, int w, int display_mode, int *view_h, int *view_w, int *out_h, int *out_w) {
    D2S_REQUIRE(h > 0 && w > 0 && display_mode >= 0 && display_mode <= 3, "d2s_dibr_out_shape: bad arguments");
    // viewer.py:2689-2760: Full modes render each eye at the texture size, Half modes into half-width (SBS) / half-height (TAB) viewports
    int vh = display_mode == D2S_HALF_TAB ? h / 2 : h, vw = display_mode == D2S_HALF_SBS ? w / 2 : w;
    const bool tab = display_mode == D2S_FULL_TAB || display_mode == D2S_HALF_TAB;
    if (view_h) *view_h = vh; if (view_w) *view_w = vw;
    if (out_h) *out_h = tab ? 2 * vh : vh; if (out_w) *out_w = tab ? vw : 2 * vw;
    return D2S_OK;
}

extern "C" size_t d2s_dibr_workspace_bytes(int h, int w, int display_mode) {
    int vh, vw, oh, ow;
    if (d2s_dibr_out_shape(h, w, display_mode, &vh, &vw, &oh, &ow)) return 0;
    return sizeof(uint32_t) * (4 + (size_t)2 * vh * vw);      // a counter + one queue slot per output pixel (worst case: every pixel deferred)
}

extern "C" int d2s_make_sbs_dibr(const d2s_dibr_params *p, d2s_stream_t stream) {
    D2S_REQUIRE(p && p->rgb.base && p->out.base && p->depth, "d2s_make_sbs_dibr: null argument");
    D2S_REQUIRE(p->h >= 2 && p->w >= 2, "d2s_make_sbs_dibr: frame %dx%d", p->h, p->w);
    D2S_REQUIRE(p->search_radius >= 0 && p->search_radius <= 32, "d2s_make_sbs_dibr: search_radius %d (0..32)", p->search_radius);
    DibrK k{};
    int oh, ow;
    int rc = d2s_dibr_out_shape(p->h, p->w, p->display_mode, &k.vh, &k.vw, &oh, &ow);
    if (rc) return rc;
    D2S_REQUIRE(k.vh >= 1 && k.vw >= 1, "d2s_make_sbs_dibr: empty eye view");
    k.rgb = p->rgb.base; k.rsc = p->rgb.sc; k.rsy = p->rgb.sy; k.rsx = p->rgb.sx; k.rgb_dtype = p->rgb.dtype;
    k.depth = p->depth; k.depth_dtype = p->depth_dtype;
    k.out = p->out.base; k.osc = p->out.sc; k.osy = p->out.sy; k.osx = p->out.sx; k.out_dtype = p->out.dtype;
    k.w = p->w; k.h = p->h;
    k.tab = p->display_mode == D2S_FULL_TAB || p->display_mode == D2S_HALF_TAB;
    k.res_x = p->resolution_x > 0.f ? p->resolution_x : (float)k.vw;      // u_resolution: "viewport resolution" (viewer.py:395)
    k.res_y = p->resolution_y > 0.f ? p->resolution_y : (float)k.vh;
    k.ipd_half = (float)(p->ipd_uv / 2.0);
    k.depth_strength = (float)(0.1 * p->depth_ratio);                      // viewer.py:1334, 2686
    k.convergence = (float)p->convergence;
    k.c = (float)cos(p->roll); k.s = (float)sin(p->roll);
    k.search = p->search_radius;
    k.tol = p->depth_tolerance; k.blur_radius = p->blur_radius;
    for (int i = 0; i <= 32; ++i) { k.w1[i] = (float)exp(-(double)i * 0.15); k.w2[i] = (float)exp(-(double)i * 0.2); }
    k.feather = p->feather_enabled; k.feather_width = p->feather_width; k.corner_radius = p->corner_radius;
    dim3 grid(ceil_div(k.vw, 256), k.vh, 2);
    D2S_REQUIRE(k.vh <= 32767 && k.vw <= 65535, "d2s_make_sbs_dibr: eye view %dx%d too large", k.vh, k.vw);
    uint32_t *ws = nullptr;
    if (p->workspace) {
        D2S_REQUIRE(p->workspace_bytes >= d2s_dibr_workspace_bytes(p->h, p->w, p->display_mode) && ((uintptr_t)p->workspace & 15) == 0,
                    "d2s_make_sbs_dibr: workspace too small or misaligned (%zu bytes, need %zu)", p->workspace_bytes, d2s_dibr_workspace_bytes(p->h, p->w, p->display_mode));
        ws = (uint32_t *)p->workspace;
    }
    switch (p->rgb.dtype) {
        case D2S_U8: rc = launch_dibr_depth<uint8_t>(k, grid, ws, stream); break;
        case D2S_F16: rc = launch_dibr_depth<__half>(k, grid, ws, stream); break;
        case D2S_F32: rc = launch_dibr_depth<float>(k, grid, ws, stream); break;
        default: return set_error(D2S_ERR_UNSUPPORTED, "d2s_make_sbs_dibr: rgb dtype %d", p->rgb.dtype);
    }
    if (rc) return rc;
    D2S_POST_LAUNCH();
    return D2S_OK;
}
