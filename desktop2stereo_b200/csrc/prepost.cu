// Pre- and post-processing around the depth network (all HBM/latency-bound, no tensor cores).
//
//   d2s_process          depth.py:542-566   BGRA/BGR u8 HWC -> RGB CHW (+ optional bilinear-antialias downscale)
//   d2s_preprocess       depth.py:676-706 + :1931 + :1946-1948   bicubic-antialias resize, /255, (x-mean)/std
//   d2s_postprocess      depth.py:806-867 (+:775, :709-736, :740-765), :1865-1887 EMA, :1998-2004 upsample
//
// The separable antialias resampler follows ATen's upsample_gen2d_aa_out_frame (UpSampleBicubic2d.cu /
// UpSample.cuh, torch 2.11): per-output span [xmin, xmin+xsize), filter evaluated at
// (j + (xmin - center) + 0.5) * invscale, weights normalised by their sum, horizontal dot products first,
// then the vertical one — evaluated in fp32.  Compiled with -fmad=false; FMAs are explicit.
#include <math.h>

#include <mutex>

#include "prepost.cuh"

namespace d2s {

// ------------------------------------------------------------------------------------------------
// antialias resampler
// ------------------------------------------------------------------------------------------------
struct AATable {  // per output index: span + normalised weights (K floats)
    int *xmin;
    int *xsize;
    float *w;
    int K;
};

__device__ __forceinline__ float bicubic_aa(float x) {  // BicubicFilterFunctor, a = -0.5
    const float a = -0.5f;
    if (x < 0) x = -x;
    if (x < 1.f) return __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(a + 2.f, x), a + 3.f), x), x), 1.f);
    if (x < 2.f) return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(x, 5.f), x), 8.f), x), 4.f), a);
    return 0.f;
}
__device__ __forceinline__ float bilinear_aa(float x) {  // BilinearFilterFunctor
    if (x < 0) x = -x;
    return x < 1.f ? __fsub_rn(1.f, x) : 0.f;
}

template <bool CUBIC>
__global__ void aa_table_kernel(AATable t, int in_size, int out_size, float scale, float support) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= out_size) return;
    float center = __fmul_rn(scale, (float)i + 0.5f);
    int xmin = max((int)(__fadd_rn(__fsub_rn(center, support), 0.5f)), 0);
    int xsize = min((int)(__fadd_rn(__fadd_rn(center, support), 0.5f)), in_size) - xmin;
    xsize = min(max(xsize, 0), t.K);
    float invscale = scale >= 1.f ? __fdiv_rn(1.f, scale) : 1.f;
    float xmc = __fsub_rn((float)xmin, center);
    float *w = t.w + (size_t)i * t.K;
    float total = 0.f;
    for (int j = 0; j < xsize; ++j) {
        float x = __fmul_rn(__fadd_rn(__fadd_rn((float)j, xmc), 0.5f), invscale);
        float v = CUBIC ? bicubic_aa(x) : bilinear_aa(x);
        w[j] = v;
        total = __fadd_rn(total, v);
    }
    for (int j = 0; j < xsize; ++j)
        if (total != 0.f) w[j] = __fdiv_rn(w[j], total);
    for (int j = xsize; j < t.K; ++j) w[j] = 0.f;
    t.xmin[i] = xmin;
    t.xsize[i] = xsize;
}

struct ImgView { const void *base; long long sc, sy, sx; };

// horizontal pass: src (strided, ST) [3,h,w] -> tmp [3,h,ow] fp32
template <typename ST>
__global__ void aa_resize_h_kernel(ImgView src, int h, int ow, AATable t, float *__restrict__ tmp) {
    int ox = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (ox >= ow) return;
    int xmin = t.xmin[ox], xsize = t.xsize[ox];
    const float *w = t.w + (size_t)ox * t.K;
    const ST *row = (const ST *)src.base + (long long)y * src.sy + (long long)xmin * src.sx;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int j = 0; j < xsize; ++j) {
        float wj = w[j];
        const ST *p = row + (long long)j * src.sx;
        float v0 = to_f32<ST>(__ldg(p)), v1 = to_f32<ST>(__ldg(p + src.sc)), v2 = to_f32<ST>(__ldg(p + 2 * src.sc));
        if (j == 0) { a0 = __fmul_rn(v0, wj); a1 = __fmul_rn(v1, wj); a2 = __fmul_rn(v2, wj); }
        else { a0 = __fmaf_rn(v0, wj, a0); a1 = __fmaf_rn(v1, wj, a1); a2 = __fmaf_rn(v2, wj, a2); }
    }
    size_t plane = (size_t)h * ow, o = (size_t)y * ow + ox;
    tmp[o] = a0; tmp[plane + o] = a1; tmp[2 * plane + o] = a2;
}

// The same horizontal pass for planar sources with contiguous rows (what process() produces): one block = one source row, whose three
// colour planes are first staged in shared memory with coalesced (16-byte where aligned) loads — at a 7.4x downscale every output
// reads ~30 source pixels that overlap its neighbours', which the direct kernel fetches as strided 2-byte loads (60 us per 4K frame,
// 0.8 TB/s).  Same products in the same order: bit-identical to aa_resize_h_kernel.
template <typename ST>
__global__ void __launch_bounds__(256) aa_resize_h_staged_kernel(ImgView src, int h, int w, int ow, AATable t, float *__restrict__ tmp) {
    extern __shared__ __align__(16) uint8_t s_row_raw[];
    ST *s_row = (ST *)s_row_raw;                       // [3][w]
    const int y = blockIdx.x;
    const ST *base = (const ST *)src.base + (long long)y * src.sy;
    constexpr int V = 16 / sizeof(ST);
    const bool vec = (w % V) == 0 && (((uintptr_t)base) & 15) == 0 && ((src.sc * (long long)sizeof(ST)) & 15) == 0;
    for (int c = 0; c < 3; ++c) {
        const ST *p = base + c * src.sc;
        ST *d = s_row + (size_t)c * w;
        if (vec) { for (int i = threadIdx.x; i < w / V; i += blockDim.x) ((uint4 *)d)[i] = __ldg((const uint4 *)p + i); }
        else { for (int i = threadIdx.x; i < w; i += blockDim.x) d[i] = __ldg(p + i); }
    }
    __syncthreads();
    const size_t plane = (size_t)h * ow;
    for (int ox = threadIdx.x; ox < ow; ox += blockDim.x) {
        const int xmin = t.xmin[ox], xsize = t.xsize[ox];
        const float *wt = t.w + (size_t)ox * t.K;
        const ST *r0 = s_row + xmin, *r1 = r0 + w, *r2 = r1 + w;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int j = 0; j < xsize; ++j) {
            const float wj = wt[j];
            const float v0 = to_f32<ST>(r0[j]), v1 = to_f32<ST>(r1[j]), v2 = to_f32<ST>(r2[j]);
            if (j == 0) { a0 = __fmul_rn(v0, wj); a1 = __fmul_rn(v1, wj); a2 = __fmul_rn(v2, wj); }
            else { a0 = __fmaf_rn(v0, wj, a0); a1 = __fmaf_rn(v1, wj, a1); a2 = __fmaf_rn(v2, wj, a2); }
        }
        const size_t o = (size_t)y * ow + ox;
        tmp[o] = a0; tmp[plane + o] = a1; tmp[2 * plane + o] = a2;
    }
}

// vertical pass + epilogue.  NORM: ((v/255) - mean)/std  (depth.py:1931, 1946-1948)
template <typename OT, bool NORM>
__global__ void aa_resize_v_kernel(const float *__restrict__ tmp, int h, int ow, int oh, AATable t, OT *__restrict__ dst,
                                   float m0, float m1, float m2, float s0, float s1, float s2) {
    int ox = blockIdx.x * blockDim.x + threadIdx.x;
    int oy = blockIdx.y, c = blockIdx.z;
    if (ox >= ow) return;
    int ymin = t.xmin[oy], ysize = t.xsize[oy];
    const float *w = t.w + (size_t)oy * t.K;
    const float *col = tmp + (size_t)c * h * ow + (size_t)ymin * ow + ox;
    float acc = 0.f;
    for (int j = 0; j < ysize; ++j) {
        float v = col[(size_t)j * ow];
        acc = (j == 0) ? __fmul_rn(v, w[0]) : __fmaf_rn(v, w[j], acc);
    }
    if (NORM) {
        float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
        acc = __fdiv_rn(__fsub_rn(__fdiv_rn(acc, 255.0f), mean), sd);
    }
    dst[(size_t)c * oh * ow + (size_t)oy * ow + ox] = from_f32<OT>(acc);
}

// no-resize fast path of process(): BGR(A) u8 HWC -> RGB CHW, 4 pixels per thread
template <typename OT>
__global__ void swizzle_chw_kernel(const uint8_t *__restrict__ src, int h, int w, int ch, OT *__restrict__ dst) {
    long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    long long n = (long long)h * w;
    if (i4 >= n) return;
    float r[4], g[4], b[4];
    int cnt = (int)min(4LL, n - i4);
    if (ch == 4 && cnt == 4) {
        uint4 q = __ldg((const uint4 *)(src + i4 * 4));
        uint32_t px[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int p = 0; p < 4; ++p) { b[p] = (float)(px[p] & 0xff); g[p] = (float)((px[p] >> 8) & 0xff); r[p] = (float)((px[p] >> 16) & 0xff); }
    } else {
        for (int p = 0; p < cnt; ++p) {
            const uint8_t *q = src + (i4 + p) * ch;
            b[p] = (float)q[0]; g[p] = (float)q[1]; r[p] = (float)q[2];
        }
    }
    for (int p = 0; p < cnt; ++p) {
        dst[i4 + p] = from_f32<OT>(r[p]);
        dst[n + i4 + p] = from_f32<OT>(g[p]);
        dst[2 * n + i4 + p] = from_f32<OT>(b[p]);
    }
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct ResizePlan {
    float scale_h, scale_w, support_h, support_w;
    int Kh, Kw;
    size_t off_tmp, off_xmin_w, off_xsize_w, off_ww, off_xmin_h, off_xsize_h, off_wh, total;
};

static ResizePlan plan_resize(int h, int w, int oh, int ow, bool cubic) {
    ResizePlan p;
    const float interp = cubic ? 4.f : 2.f;
    p.scale_h = (float)h / (float)oh;  // area_pixel_compute_scale, align_corners=False, no scale_factor
    p.scale_w = (float)w / (float)ow;
    p.support_h = p.scale_h >= 1.f ? (interp * 0.5f) * p.scale_h : interp * 0.5f;
    p.support_w = p.scale_w >= 1.f ? (interp * 0.5f) * p.scale_w : interp * 0.5f;
    p.Kh = (int)ceilf(p.support_h) * 2 + 1;
    p.Kw = (int)ceilf(p.support_w) * 2 + 1;
    size_t o = 0;
    p.off_tmp = o; o = align_up(o + sizeof(float) * 3 * (size_t)h * ow, 256);
    p.off_xmin_w = o; o = align_up(o + sizeof(int) * ow, 256);
    p.off_xsize_w = o; o = align_up(o + sizeof(int) * ow, 256);
    p.off_ww = o; o = align_up(o + sizeof(float) * (size_t)ow * p.Kw, 256);
    p.off_xmin_h = o; o = align_up(o + sizeof(int) * oh, 256);
    p.off_xsize_h = o; o = align_up(o + sizeof(int) * oh, 256);
    p.off_wh = o; o = align_up(o + sizeof(float) * (size_t)oh * p.Kh, 256);
    p.total = o;
    return p;
}

// what: bit 0 = build the two weight tables (input-independent: a caller that keeps its workspace may do it once per shape),
//       bit 1 = run the two filter passes
template <bool CUBIC, bool NORM>
static int run_resize(const d2s_image *src, int h, int w, void *dst, int dst_dtype, int oh, int ow, const float *mean,
                      const float *std, void *ws, size_t ws_bytes, d2s_stream_t st, int what = 3) {
    ResizePlan p = plan_resize(h, w, oh, ow, CUBIC);
    D2S_REQUIRE(ws && ws_bytes >= p.total, "resize: workspace too small (%zu < %zu)", ws_bytes, p.total);
    char *b = (char *)ws;
    AATable tw{(int *)(b + p.off_xmin_w), (int *)(b + p.off_xsize_w), (float *)(b + p.off_ww), p.Kw};
    AATable th{(int *)(b + p.off_xmin_h), (int *)(b + p.off_xsize_h), (float *)(b + p.off_wh), p.Kh};
    float *tmp = (float *)(b + p.off_tmp);
    if (what & 1) {
        D2S_LAUNCH((aa_table_kernel<CUBIC>), ceil_div(ow, 128), 128, 0, st, tw, w, ow, p.scale_w, p.support_w);
        D2S_LAUNCH((aa_table_kernel<CUBIC>), ceil_div(oh, 128), 128, 0, st, th, h, oh, p.scale_h, p.support_h);
    }
    if (!(what & 2)) { D2S_POST_LAUNCH(); return D2S_OK; }
    ImgView v{src->base, src->sc, src->sy, src->sx};
    dim3 gh(ceil_div(ow, 128), h);
    // planar source, contiguous rows, a downscale (several source pixels per output), row fits in shared memory: staged kernel
    const size_t es = src->dtype == D2S_F32 ? 4 : (src->dtype == D2S_U8 ? 1 : 2);
    if (src->sx == 1 && src->sc > 0 && w >= 2 * ow && (size_t)3 * w * es <= 96 * 1024 && src->dtype != D2S_BF16) {
        const size_t smem = (size_t)3 * w * es;
        static std::once_flag once;
        std::call_once(once, [] {
            cudaFuncSetAttribute((const void *)aa_resize_h_staged_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            cudaFuncSetAttribute((const void *)aa_resize_h_staged_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            cudaFuncSetAttribute((const void *)aa_resize_h_staged_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        });
        switch (src->dtype) {
            case D2S_U8:  D2S_LAUNCH((aa_resize_h_staged_kernel<uint8_t>), h, 256, smem, st, v, h, w, ow, tw, tmp); break;
            case D2S_F16: D2S_LAUNCH((aa_resize_h_staged_kernel<__half>), h, 256, smem, st, v, h, w, ow, tw, tmp); break;
            default:      D2S_LAUNCH((aa_resize_h_staged_kernel<float>), h, 256, smem, st, v, h, w, ow, tw, tmp); break;
        }
    } else
    switch (src->dtype) {
        case D2S_U8:  D2S_LAUNCH((aa_resize_h_kernel<uint8_t>), gh, 128, 0, st, v, h, ow, tw, tmp); break;
        case D2S_F16: D2S_LAUNCH((aa_resize_h_kernel<__half>), gh, 128, 0, st, v, h, ow, tw, tmp); break;
        case D2S_F32: D2S_LAUNCH((aa_resize_h_kernel<float>), gh, 128, 0, st, v, h, ow, tw, tmp); break;
        case D2S_BF16: D2S_LAUNCH((aa_resize_h_kernel<__nv_bfloat16>), gh, 128, 0, st, v, h, ow, tw, tmp); break;
        default: return set_error(D2S_ERR_UNSUPPORTED, "resize: source dtype %d", src->dtype);
    }
    dim3 gv(ceil_div(ow, 128), oh, 3);
    float m0 = mean ? mean[0] : 0, m1 = mean ? mean[1] : 0, m2 = mean ? mean[2] : 0;
    float s0 = std ? std[0] : 1, s1 = std ? std[1] : 1, s2 = std ? std[2] : 1;
    switch (dst_dtype) {
        case D2S_F32: D2S_LAUNCH((aa_resize_v_kernel<float, NORM>), gv, 128, 0, st, tmp, h, ow, oh, th, (float *)dst, m0, m1, m2, s0, s1, s2); break;
        case D2S_F16: D2S_LAUNCH((aa_resize_v_kernel<__half, NORM>), gv, 128, 0, st, tmp, h, ow, oh, th, (__half *)dst, m0, m1, m2, s0, s1, s2); break;
        case D2S_BF16: D2S_LAUNCH((aa_resize_v_kernel<__nv_bfloat16, NORM>), gv, 128, 0, st, tmp, h, ow, oh, th, (__nv_bfloat16 *)dst, m0, m1, m2, s0, s1, s2); break;
        default: return set_error(D2S_ERR_UNSUPPORTED, "resize: destination dtype %d", dst_dtype);
    }
    D2S_POST_LAUNCH();
    return D2S_OK;
}

// ------------------------------------------------------------------------------------------------
// post-process
// ------------------------------------------------------------------------------------------------
constexpr int kSortCap = 8192;

// metric models (depth.py:837-841): inv = where(d > 0, 1 / clamp(d, 1e-12), d), each op rounded to CT
template <typename IT, typename CT>
__device__ __forceinline__ float metric_inv(IT raw, bool &valid) {
    const float x = round_to<CT>(to_f32<IT>(raw));
    valid = x > 0.f;
    return valid ? round_to<CT>(__fdiv_rn(1.0f, fmaxf(x, round_to<CT>(1e-12f)))) : x;
}

// Q1: strided subsample -> bitonic sort in shared memory -> k-th smallest / k-th largest  (depth.py:784-794, 849-858).
// metric: the sample is taken from the COMPACTED list of valid (d > 0) inverted values, v = inv[valid]; vv = v[::step] with
// step = ceil(len(v) / cap) — a block-wide scan gives every valid element its rank in v, so the strided pick is exact.
template <typename IT, typename CT>
__global__ void __launch_bounds__(1024) post_bounds_kernel(const IT *__restrict__ d, int n, int step, int ns, int k, int metric, int cap,
                                                           double lo_q, float *__restrict__ bounds) {
    __shared__ float s[kSortCap];
    __shared__ int s_scan[1024];
    __shared__ int s_meta[3];   // n valid, ns, k
    int n_eff = n;
    if (!metric) {
        for (int i = threadIdx.x; i < kSortCap; i += blockDim.x)
            s[i] = i < ns ? round_to<CT>(to_f32<IT>(d[(size_t)i * step])) : INFINITY;
    } else {
        const int chunk = (n + 1023) / 1024, i0 = min((int)threadIdx.x * chunk, n), i1 = min(i0 + chunk, n);
        int cnt = 0;
        for (int i = i0; i < i1; ++i) { bool v; metric_inv<IT, CT>(d[i], v); cnt += v ? 1 : 0; }
        s_scan[threadIdx.x] = cnt;
        for (int i = threadIdx.x; i < kSortCap; i += blockDim.x) s[i] = INFINITY;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {            // inclusive Hillis-Steele scan of the per-thread counts
            int v = threadIdx.x >= off ? s_scan[threadIdx.x - off] : 0;
            __syncthreads();
            s_scan[threadIdx.x] += v;
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const int nv = s_scan[1023];
            int st = 1, m = nv;
            if (nv > cap) { st = (nv + cap - 1) / cap; m = (nv + st - 1) / st; }
            int kk = (int)nearbyint(lo_q * (double)(m - 1)) + 1;
            kk = kk < 1 ? 1 : kk; kk = kk > m ? m : kk;
            s_meta[0] = nv; s_meta[1] = m; s_meta[2] = kk;
            s_scan[1023] = st;   // (the last inclusive value is no longer needed: reuse the slot for the step)
        }
        __syncthreads();
        const int st = s_scan[1023];
        int rank = threadIdx.x == 0 ? 0 : (threadIdx.x == 1023 ? s_meta[0] - cnt : s_scan[threadIdx.x - 1]);
        for (int i = i0; i < i1; ++i) {
            bool v;
            const float x = metric_inv<IT, CT>(d[i], v);
            if (v) { if (rank % st == 0) s[rank / st] = x; ++rank; }
        }
        n_eff = s_meta[0]; ns = s_meta[1]; k = s_meta[2];
    }
    __syncthreads();
    for (int size = 2; size <= kSortCap; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < kSortCap / 2; t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool up = ((lo & size) == 0);
                float a = s[lo], b = s[hi];
                if ((a > b) == up) { s[lo] = b; s[hi] = a; }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        float lo, hi;
        if (n_eff <= 10) { lo = 0.f; hi = 0.f; }             // depth.py:845-847
        else if (k >= ns) { lo = s[0]; hi = s[ns - 1]; }       // tail_count == n -> min/max
        else { lo = s[k - 1]; hi = s[ns - k]; }
        bounds[0] = lo; bounds[1] = hi;
    }
}

// Q1 (normalise) + Q2 (gamma, foreground scale), per element, each op rounded to CT like ATen's opmath kernels
template <typename IT, typename CT>
__global__ void post_point_kernel(const IT *__restrict__ d, int n, const float *__restrict__ bounds, float gamma,
                                  int fg_on, float fg_exp, int metric, float *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float lo = bounds[0], hi = bounds[1];
    float denom = round_to<CT>(fmaxf(round_to<CT>(__fsub_rn(hi, lo)), 1e-6f));
    bool valid;
    float x = metric ? metric_inv<IT, CT>(d[i], valid) : round_to<CT>(to_f32<IT>(d[i]));
    float t = round_to<CT>(__fdiv_rn(round_to<CT>(__fsub_rn(x, lo)), denom));
    t = fminf(fmaxf(t, 0.f), 1.f);
    t = round_to<CT>(powf(t, gamma));
    t = fminf(fmaxf(t, 0.f), 1.f);  // apply_foreground_scale: depth.clamp(0,1)
    if (fg_on) {
        float dist = round_to<CT>(__fsub_rn(t, 0.5f));
        float p = round_to<CT>(powf(fabsf(dist), fg_exp));
        float sg = dist > 0.f ? 1.f : (dist < 0.f ? -1.f : 0.f);
        t = round_to<CT>(__fadd_rn(0.5f, round_to<CT>(__fmul_rn(sg, p))));
        t = fminf(fmaxf(t, 0.f), 1.f);
    }
    out[i] = t;
}

struct BlurW { float w[64]; int k; };

// Q2 anti_alias: separable Gaussian, zero padding (F.conv2d padding=k//2).  AXIS 0: along x, 1: along y.
// The vertical pass also applies the EMA (Q3) and writes the low-res result.
template <typename CT, int AXIS>
__global__ void post_blur_kernel(const float *__restrict__ in, int H, int W, BlurW bw, float *__restrict__ out,
                                 CT *__restrict__ ema, int ema_valid, float ema_w, CT *__restrict__ out_low) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    int r = bw.k / 2;
    float acc = 0.f;
    for (int j = 0; j < bw.k; ++j) {
        int xx = AXIS == 0 ? x + j - r : x, yy = AXIS == 0 ? y : y + j - r;
        if (xx >= 0 && xx < W && yy >= 0 && yy < H) acc = __fmaf_rn(in[(size_t)yy * W + xx], bw.w[j], acc);
    }
    acc = round_to<CT>(acc);
    size_t o = (size_t)y * W + x;
    if (AXIS == 1) {
        if (ema) {  // DepthStabilizer: prev.lerp_(depth, 1-alpha)  (Lerp.h: weight < 0.5 -> self + w*(end-self))
            const float prev = to_f32<CT>(ema[o]);
            if (ema_valid == 1 || (ema_valid == 2 && prev == prev)) acc = round_to<CT>(__fmaf_rn(ema_w, __fsub_rn(acc, prev), prev));
            ema[o] = from_f32<CT>(acc);
        }
        if (out_low) out_low[o] = from_f32<CT>(acc);
    }
    out[o] = acc;
}

// pass-through used when the blur is disabled (k < 3): EMA + low-res store only
template <typename CT>
__global__ void post_ema_kernel(float *__restrict__ buf, int n, CT *__restrict__ ema, int ema_valid, float ema_w,
                                CT *__restrict__ out_low) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = buf[i];
    if (ema) {
        const float prev = to_f32<CT>(ema[i]);
        if (ema_valid == 1 || (ema_valid == 2 && prev == prev)) acc = round_to<CT>(__fmaf_rn(ema_w, __fsub_rn(acc, prev), prev));
        ema[i] = from_f32<CT>(acc);
    }
    if (out_low) out_low[i] = from_f32<CT>(acc);
    buf[i] = acc;
}

// Q4: upsample_bilinear2d, align_corners=False (UpSampleBilinear2d.cu), fp32 accumulate, rounded to CT then stored as OT
template <typename CT, typename OT>
__global__ void post_upsample_kernel(const float *__restrict__ in, int H, int W, OT *__restrict__ out, int oh, int ow,
                                     float sh, float sw) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= ow) return;
    float v;
    if (oh == H && ow == W) v = in[(size_t)y * W + x];
    else {
        float h1r = fmaxf(__fmaf_rn(sh, (float)y + 0.5f, -0.5f), 0.f);
        float w1r = fmaxf(__fmaf_rn(sw, (float)x + 0.5f, -0.5f), 0.f);
        int h1 = (int)h1r, w1 = (int)w1r;
        int h1p = (h1 < H - 1) ? 1 : 0, w1p = (w1 < W - 1) ? 1 : 0;
        float h1l = __fsub_rn(h1r, (float)h1), h0l = __fsub_rn(1.f, h1l);
        float w1l = __fsub_rn(w1r, (float)w1), w0l = __fsub_rn(1.f, w1l);
        const float *r0 = in + (size_t)h1 * W, *r1 = in + (size_t)(h1 + h1p) * W;
        float top = __fmaf_rn(w0l, r0[w1], __fmul_rn(w1l, r0[w1 + w1p]));
        float bot = __fmaf_rn(w0l, r1[w1], __fmul_rn(w1l, r1[w1 + w1p]));
        v = round_to<CT>(__fmaf_rn(h0l, top, __fmul_rn(h1l, bot)));
    }
    out[(size_t)y * ow + x] = from_f32<OT>(v);
}

// the same upsample, 4 consecutive output pixels per thread (row terms computed once, one vector store): ow % 4 == 0
template <typename CT, typename OT>
__global__ void post_upsample4_kernel(const float *__restrict__ in, int H, int W, OT *__restrict__ out, int oh, int ow, float sh, float sw) {
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x0 >= ow) return;
    const float h1r = fmaxf(__fmaf_rn(sh, (float)y + 0.5f, -0.5f), 0.f);
    const int h1 = (int)h1r, h1p = (h1 < H - 1) ? 1 : 0;
    const float h1l = __fsub_rn(h1r, (float)h1), h0l = __fsub_rn(1.f, h1l);
    const float *r0 = in + (size_t)h1 * W, *r1 = in + (size_t)(h1 + h1p) * W;
    OT v[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const float w1r = fmaxf(__fmaf_rn(sw, (float)(x0 + p) + 0.5f, -0.5f), 0.f);
        const int w1 = (int)w1r, w1p = (w1 < W - 1) ? 1 : 0;
        const float w1l = __fsub_rn(w1r, (float)w1), w0l = __fsub_rn(1.f, w1l);
        const float top = __fmaf_rn(w0l, r0[w1], __fmul_rn(w1l, r0[w1 + w1p]));
        const float bot = __fmaf_rn(w0l, r1[w1], __fmul_rn(w1l, r1[w1 + w1p]));
        v[p] = from_f32<OT>(round_to<CT>(__fmaf_rn(h0l, top, __fmul_rn(h1l, bot))));
    }
    OT *o = out + (size_t)y * ow + x0;
    if (sizeof(OT) == 2) *(uint2 *)o = *(const uint2 *)v;
    else if (sizeof(OT) == 4) *(uint4 *)o = *(const uint4 *)v;
    else { o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = v[3]; }
}

template <typename CT> static float host_round(float v);
template <> float host_round<float>(float v) { return v; }
template <> float host_round<__half>(float v) { return __half2float(__float2half_rn(v)); }
template <> float host_round<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// anti_alias's kernel, built the way depth.py:752-757 builds it, each op rounded to CT
template <typename CT>
static BlurW make_gauss(float strength) {
    BlurW bw{};
    int k = (int)(3 * strength) | 1;
    bw.k = k;
    if (k < 3 || k > 63) return bw;
    double sigma = 0.5 * (double)strength;
    float denom = (float)(2 * sigma * sigma);
    float sum = 0.f;
    for (int j = 0; j < k; ++j) {
        float c = host_round<CT>((float)(j - k / 2));
        float sq = host_round<CT>(c * c);
        float e = host_round<CT>(-sq / denom);
        bw.w[j] = host_round<CT>(expf(e));
        sum += bw.w[j];
    }
    sum = host_round<CT>(sum);
    for (int j = 0; j < k; ++j) bw.w[j] = host_round<CT>(bw.w[j] / sum);
    return bw;
}

// phases: POST_PHASE_HEAD = percentile bounds + point ops + horizontal blur (no cross-frame state),
//         POST_PHASE_EMA  = vertical blur + DepthStabilizer EMA + low-res store (the one cross-frame dependency of the path),
//         POST_PHASE_UP   = final bilinear upsample
template <typename IT, typename CT>
static int run_post(const d2s_post_params *p, d2s_stream_t st, int phases) {
    const int H = p->H, W = p->W, n = H * W;
    size_t need = d2s_postprocess_workspace_bytes(H, W);
    D2S_REQUIRE(p->workspace && p->workspace_bytes >= need, "d2s_postprocess: workspace too small (%zu < %zu)", p->workspace_bytes, need);
    float *bounds = (float *)p->workspace;
    float *bufA = bounds + 64, *bufB = bufA + align_up(n, 64);
    // depth.py:849-858 integer maths
    int step = 1, ns = n;
    if (n > p->subsample_cap) { step = (n + p->subsample_cap - 1) / p->subsample_cap; ns = (n + step - 1) / step; }
    D2S_REQUIRE(ns <= kSortCap, "d2s_postprocess: subsample of %d values exceeds %d", ns, kSortCap);
    double lo_q = fmax(0.0, fmin(1.0, (double)p->percentile / 100.0));
    int k = (int)nearbyint(lo_q * (double)(ns - 1)) + 1;
    k = k < 1 ? 1 : k; k = k > ns ? ns : k;
    int fg_on = fabs((double)p->foreground_scale) >= 1e-6;
    float fg_exp = (float)(1.0 / (1.0 + (double)p->foreground_scale));
    if (phases & POST_PHASE_HEAD) {
        D2S_LAUNCH((post_bounds_kernel<IT, CT>), 1, 1024, 0, st, (const IT *)p->depth_in, n, step, ns, k, p->metric ? 1 : 0, p->subsample_cap, lo_q, bounds);
        D2S_LAUNCH((post_point_kernel<IT, CT>), ceil_div(n, 256), 256, 0, st, (const IT *)p->depth_in, n, bounds, p->gamma, fg_on, fg_exp, p->metric ? 1 : 0, bufA);
    }
    BlurW bw = make_gauss<CT>(p->aa_strength);
    D2S_REQUIRE(bw.k <= 63, "d2s_postprocess: anti-alias kernel size %d too large", bw.k);
    float ema_w = (float)(1.0 - (double)p->ema_alpha);
    CT *ema = (CT *)p->ema_state;
    CT *out_low = (CT *)p->out_lowres;
    dim3 g(ceil_div(W, 128), H);
    float *res = bufA;
    if (bw.k >= 3) {
        if (phases & POST_PHASE_HEAD) D2S_LAUNCH((post_blur_kernel<CT, 0>), g, 128, 0, st, bufA, H, W, bw, bufB, (CT *)nullptr, 0, 0.f, (CT *)nullptr);
        if (phases & POST_PHASE_EMA) D2S_LAUNCH((post_blur_kernel<CT, 1>), g, 128, 0, st, bufB, H, W, bw, bufA, ema, p->ema_valid, ema_w, out_low);
    } else if ((ema || out_low) && (phases & POST_PHASE_EMA)) {
        D2S_LAUNCH((post_ema_kernel<CT>), ceil_div(n, 256), 256, 0, st, bufA, n, ema, p->ema_valid, ema_w, out_low);
    }
    if (p->out && (phases & POST_PHASE_UP)) {
        dim3 gu(ceil_div(p->out_w, 128), p->out_h);
        float sh = (float)H / (float)p->out_h, sw = (float)W / (float)p->out_w;
        const bool vec4 = !(p->out_h == H && p->out_w == W) && p->out_w % 4 == 0 && ((uintptr_t)p->out & 15) == 0;
        if (vec4) {
            dim3 g4(ceil_div(p->out_w / 4, 128), p->out_h);
            switch (p->out_dtype) {
                case D2S_F32: D2S_LAUNCH((post_upsample4_kernel<CT, float>), g4, 128, 0, st, res, H, W, (float *)p->out, p->out_h, p->out_w, sh, sw); break;
                case D2S_F16: D2S_LAUNCH((post_upsample4_kernel<CT, __half>), g4, 128, 0, st, res, H, W, (__half *)p->out, p->out_h, p->out_w, sh, sw); break;
                case D2S_BF16: D2S_LAUNCH((post_upsample4_kernel<CT, __nv_bfloat16>), g4, 128, 0, st, res, H, W, (__nv_bfloat16 *)p->out, p->out_h, p->out_w, sh, sw); break;
                default: return set_error(D2S_ERR_UNSUPPORTED, "d2s_postprocess: out dtype %d", p->out_dtype);
            }
        } else
        switch (p->out_dtype) {
            case D2S_F32: D2S_LAUNCH((post_upsample_kernel<CT, float>), gu, 128, 0, st, res, H, W, (float *)p->out, p->out_h, p->out_w, sh, sw); break;
            case D2S_F16: D2S_LAUNCH((post_upsample_kernel<CT, __half>), gu, 128, 0, st, res, H, W, (__half *)p->out, p->out_h, p->out_w, sh, sw); break;
            case D2S_BF16: D2S_LAUNCH((post_upsample_kernel<CT, __nv_bfloat16>), gu, 128, 0, st, res, H, W, (__nv_bfloat16 *)p->out, p->out_h, p->out_w, sh, sw); break;
            default: return set_error(D2S_ERR_UNSUPPORTED, "d2s_postprocess: out dtype %d", p->out_dtype);
        }
    }
    D2S_POST_LAUNCH();
    return D2S_OK;
}

template <typename IT>
static int run_post_ct(const d2s_post_params *p, d2s_stream_t st, int phases) {
    switch (p->compute_dtype) {
        case D2S_F32: return run_post<IT, float>(p, st, phases);
        case D2S_F16: return run_post<IT, __half>(p, st, phases);
        case D2S_BF16: return run_post<IT, __nv_bfloat16>(p, st, phases);
        default: return set_error(D2S_ERR_UNSUPPORTED, "d2s_postprocess: compute dtype %d", p->compute_dtype);
    }
}

}  // namespace d2s

using namespace d2s;

extern "C" int d2s_model_input_shape(int h, int w, int target, int patch, int *new_h, int *new_w) {
    D2S_REQUIRE(h > 0 && w > 0 && target > 0 && patch > 0 && new_h && new_w, "d2s_model_input_shape: bad arguments");
    // depth.py:676-692 (python floats are doubles; round() is half-to-even)
    int longest = h > w ? h : w;
    double scale = longest != target ? (double)target / (double)longest : 1.0;
    long sh = (long)nearbyint((double)h * scale), sw = (long)nearbyint((double)w * scale);
    sh = sh < 1 ? 1 : sh; sw = sw < 1 ? 1 : sw;
    auto nearest = [patch](long x) { long down = (x / patch) * patch, up = down + patch; return (up - x) <= (x - down) ? up : down; };
    long nh = nearest(sh), nw = nearest(sw);
    *new_h = (int)(nh < 1 ? 1 : nh); *new_w = (int)(nw < 1 ? 1 : nw);
    return D2S_OK;
}

extern "C" size_t d2s_preprocess_workspace_bytes(int h, int w, int new_h, int new_w) {
    if (h <= 0 || w <= 0 || new_h <= 0 || new_w <= 0) return 0;
    size_t a = plan_resize(h, w, new_h, new_w, true).total, b = plan_resize(h, w, new_h, new_w, false).total;
    return a > b ? a : b;
}

extern "C" int d2s_preprocess(const d2s_image *src, int h, int w, void *dst, int dst_dtype, int new_h, int new_w,
                              const float mean[3], const float std[3], void *workspace, size_t workspace_bytes,
                              d2s_stream_t stream) {
    return d2s::preprocess_phases(src, h, w, dst, dst_dtype, new_h, new_w, mean, std, workspace, workspace_bytes, 3, stream);
}

int d2s::preprocess_phases(const d2s_image *src, int h, int w, void *dst, int dst_dtype, int new_h, int new_w, const float mean[3],
                           const float std[3], void *workspace, size_t workspace_bytes, int what, d2s_stream_t stream) {
    D2S_REQUIRE(src && (src->base || !(what & 2)) && (dst || !(what & 2)) && h > 0 && w > 0 && new_h > 0 && new_w > 0, "d2s_preprocess: bad arguments");
    D2S_REQUIRE(mean && std, "d2s_preprocess: mean/std required");
    return run_resize<true, true>(src, h, w, dst, dst_dtype, new_h, new_w, mean, std, workspace, workspace_bytes, stream, what);
}

size_t d2s::process_workspace_bytes(int h0, int w0, int h, int w) { return (h == h0 && w == w0) ? 0 : plan_resize(h0, w0, h, w, false).total; }

// process() with a caller-owned workspace (no stream-ordered allocation: capturable, and the tables can be built once)
int d2s::process_phases(const uint8_t *frame, int h0, int w0, int channels, void *out, int out_dtype, int h, int w, void *ws, size_t ws_bytes,
                        int what, d2s_stream_t stream) {
    if (h == h0 && w == w0) return (what & 2) ? d2s_process(frame, h0, w0, channels, out, out_dtype, h, w, stream) : D2S_OK;
    d2s_image src{};
    src.base = (void *)(frame + 2); src.dtype = D2S_U8; src.sc = -1; src.sy = (int64_t)w0 * channels; src.sx = channels;
    return run_resize<false, false>(&src, h0, w0, out, out_dtype, h, w, nullptr, nullptr, ws, ws_bytes, stream, what);
}

extern "C" int d2s_process(const uint8_t *frame, int h0, int w0, int channels, void *out, int out_dtype, int h, int w,
                           d2s_stream_t stream) {
    D2S_REQUIRE(frame && out && h0 > 0 && w0 > 0 && (channels == 3 || channels == 4), "d2s_process: bad arguments");
    D2S_REQUIRE(h > 0 && w > 0, "d2s_process: bad output size");
    if (h == h0 && w == w0) {
        int grid = ceil_div(ceil_div((long long)h * w, 4), 256);
        switch (out_dtype) {
            case D2S_F16: D2S_LAUNCH((swizzle_chw_kernel<__half>), grid, 256, 0, stream, frame, h, w, channels, (__half *)out); break;
            case D2S_F32: D2S_LAUNCH((swizzle_chw_kernel<float>), grid, 256, 0, stream, frame, h, w, channels, (float *)out); break;
            case D2S_U8:  D2S_LAUNCH((swizzle_chw_kernel<uint8_t>), grid, 256, 0, stream, frame, h, w, channels, (uint8_t *)out); break;
            case D2S_BF16: D2S_LAUNCH((swizzle_chw_kernel<__nv_bfloat16>), grid, 256, 0, stream, frame, h, w, channels, (__nv_bfloat16 *)out); break;
            default: return set_error(D2S_ERR_UNSUPPORTED, "d2s_process: out dtype %d", out_dtype);
        }
        D2S_POST_LAUNCH();
        return D2S_OK;
    }
    // bilinear + antialias downscale (depth.py:560-566).  Scratch is stream-ordered and library-owned.
    d2s_image src{};
    src.base = (void *)(frame + 2); src.dtype = D2S_U8; src.sc = -1; src.sy = (int64_t)w0 * channels; src.sx = channels;
    size_t bytes = plan_resize(h0, w0, h, w, false).total;
    void *ws = nullptr;
    D2S_CHECK_CUDA(cudaMallocAsync(&ws, bytes, (cudaStream_t)stream));
    int rc = run_resize<false, false>(&src, h0, w0, out, out_dtype, h, w, nullptr, nullptr, ws, bytes, stream);
    cudaFreeAsync(ws, (cudaStream_t)stream);
    return rc;
}

extern "C" size_t d2s_postprocess_workspace_bytes(int H, int W) {
    if (H <= 0 || W <= 0) return 0;
    return sizeof(float) * (64 + 2 * align_up((size_t)H * W, 64));
}

extern "C" int d2s_postprocess(const d2s_post_params *p, d2s_stream_t stream) { return d2s::postprocess_phases(p, POST_PHASE_ALL, stream); }

int d2s::postprocess_phases(const d2s_post_params *p, int phases, d2s_stream_t stream) {
    D2S_REQUIRE(p && p->depth_in && p->H > 0 && p->W > 0, "d2s_postprocess: bad arguments");
    D2S_REQUIRE(p->ema_valid >= 0 && p->ema_valid <= 2, "d2s_postprocess: ema_valid %d", p->ema_valid);
    D2S_REQUIRE(p->out == nullptr || (p->out_h > 0 && p->out_w > 0), "d2s_postprocess: bad output size");
    D2S_REQUIRE(p->subsample_cap >= 1, "d2s_postprocess: subsample_cap");
    D2S_REQUIRE(!p->metric || p->subsample_cap <= 8192, "d2s_postprocess: subsample_cap %d exceeds the sort capacity", p->subsample_cap);
    switch (p->in_dtype) {
        case D2S_F32: return run_post_ct<float>(p, stream, phases);
        case D2S_F16: return run_post_ct<__half>(p, stream, phases);
        case D2S_BF16: return run_post_ct<__nv_bfloat16>(p, stream, phases);
        default: return set_error(D2S_ERR_UNSUPPORTED, "d2s_postprocess: in dtype %d", p->in_dtype);
    }
}
