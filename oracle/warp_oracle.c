/* ORACLE — test infrastructure, never on the product path.
 *
 * CPU restatement of the reference's stereo warp + SBS pack:
 *   make_sbs_core                /root/reference/depth.py:2122-2184
 *   pad_to_aspect_tensor         /root/reference/depth.py:2106-2119
 * and of the two ATen ops it calls, restated from their published semantics
 * (torch 2.11, aten/src/ATen/native/cuda/GridSampler.cuh, RangeFactories.cu):
 *   torch.linspace (fp32)        two-sided formula, start + step*i | end - step*(n-1-i)
 *   F.grid_sample                bilinear, padding_mode="reflection", align_corners=True
 *
 * Strict IEEE fp32: compile with -ffp-contract=off.  Where the CUDA build of ATen contracts
 * a*b+c into one FMA (nvcc default), `fma_mode=1` uses fmaf(); `fma_mode=0` rounds the product
 * first (what an un-contracted CPU build does).  fma_mode=1 is the semantic the product kernel
 * implements; fma_mode=0 is used to pin this file against goldens produced by the reference on CPU.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { D2S_DT_F32 = 0, D2S_DT_F16 = 1, D2S_DT_BF16 = 2 };
enum { D2S_FULL_SBS = 0, D2S_HALF_SBS = 1, D2S_FULL_TAB = 2, D2S_HALF_TAB = 3 };
enum { D2S_WARP_BILINEAR = 0, D2S_WARP_GATHER = 1 };

static float round_bf16(float x) {
    uint32_t u; memcpy(&u, &x, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return x;           /* NaN */
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb; u &= 0xffff0000u;
    float r; memcpy(&r, &u, 4); return r;
}
static float round_f16(float x) { return (float)(_Float16)x; }
static float round_dt(float x, int dt) {
    return dt == D2S_DT_F16 ? round_f16(x) : dt == D2S_DT_BF16 ? round_bf16(x) : x;
}

/* depth.py:2143-2147 + :2154.  Every tensor-scalar op computes in fp32 and rounds to the
 * tensor dtype (ATen opmath for Half/BFloat16). Returns shifts (pixels) and shift_norm. */
static void shift_chain(float depth, int dt, float conv, float ratio, float max_px, float strength,
                        float two_over_wm1, float *shift_px, float *shift_norm) {
    float d = round_dt(depth - conv, dt);
    float inv = round_dt((-d) * ratio, dt);
    float s = round_dt(inv * max_px, dt);
    s = round_dt(s * strength, dt);
    *shift_px = s;
    *shift_norm = round_dt(s * two_over_wm1, dt);
}

void d2s_oracle_linspace(float *out, int n, int fma_mode) {
    /* RangeFactories.cu linspace: step=(end-start)/(steps-1); i<steps/2: start+step*i else end-step*(steps-i-1) */
    const float start = -1.0f, end = 1.0f;
    if (n == 1) { out[0] = start; return; }
    float step = (end - start) / (float)(n - 1);
    int half = n / 2;
    for (int i = 0; i < n; ++i) {
        if (i < half) out[i] = fma_mode ? fmaf(step, (float)i, start) : start + step * (float)i;
        else          out[i] = fma_mode ? fmaf(-step, (float)(n - i - 1), end) : end - step * (float)(n - i - 1);
    }
}

/* GridSampler.cuh: grid_sampler_compute_source_index for reflection + align_corners=True */
static float source_index(float coord, int size) {
    coord = ((coord + 1.f) / 2) * (float)(size - 1);            /* unnormalize */
    int twice_high = 2 * (size - 1);
    if (twice_high == 0) coord = 0.f;                            /* reflect_coordinates(…,0,0) */
    else {
        float span = (float)twice_high / 2;
        float in = fabsf(coord - 0.f);
        float extra = fmodf(in, span);
        int flips = (int)floorf(in / span);
        coord = (flips % 2 == 0) ? extra + 0.f : span - extra + 0.f;
    }
    coord = fminf((float)(size - 1), fmaxf(coord, 0.f));         /* clip_coordinates */
    return coord;
}

static void pad_geometry(int h, int w, int fill, int *ph, int *pw, int *top, int *left) {
    /* depth.py:2106-2119 (python float maths = double) */
    *ph = h; *pw = w; *top = 0; *left = 0;
    if (!fill) return;
    double r_img = (double)w / (double)h, r_t = 16.0 / 9.0;
    if (fabs(r_img - r_t) < 1e-3) return;
    if (r_img > r_t) { int nh = (int)nearbyint((double)w / r_t); *ph = nh; *top = (nh - h) / 2; }
    else             { int nw = (int)nearbyint((double)h * r_t); *pw = nw; *left = (nw - w) / 2; }
}

void d2s_oracle_out_shape(int h, int w, int display_mode, int fill_16_9, int *out_h, int *out_w) {
    int ph, pw, t, l; pad_geometry(h, w, fill_16_9, &ph, &pw, &t, &l);
    int tab = (display_mode == D2S_FULL_TAB || display_mode == D2S_HALF_TAB);
    int half = (display_mode == D2S_HALF_SBS || display_mode == D2S_HALF_TAB);
    int ch = tab ? 2 * ph : ph, cw = tab ? pw : 2 * pw;
    *out_h = half ? ph : ch; *out_w = half ? pw : cw;
}

/* rgb: [3,h,w] fp32 (already in the eye dtype's value set), depth: [h,w] fp32 holding values of
 * dtype `depth_dt`.  out: [3,out_h,out_w] fp32.  idx_left/idx_right (optional, [h,w] int32):
 * floor(ix) (bilinear) or the gather coordinate.  xs/ys optional overrides of the linspace grids.
 * out_dt: dtype of the reference's result tensor (fp32 for the grid_sample branch, the eye dtype for the
 * gather branch) — values are rounded to it.
 * Returns 0. */
int d2s_oracle_make_sbs(const float *rgb, const float *depth, int depth_dt, int h, int w,
                        float ipd_uv_w /* (float)(ipd_uv*W) */, float depth_ratio, float convergence,
                        int display_mode, int fill_16_9, int warp_mode, int fma_mode, int out_dt,
                        const float *xs_in, const float *ys_in,
                        float *out, int32_t *idx_left, int32_t *idx_right) {
    const size_t plane = (size_t)h * w;
    float *eyes = (float *)malloc(sizeof(float) * 6 * plane);   /* left[3,h,w], right[3,h,w] */
    float *xs = (float *)malloc(sizeof(float) * w), *ys = (float *)malloc(sizeof(float) * h);
    if (xs_in) memcpy(xs, xs_in, sizeof(float) * w); else d2s_oracle_linspace(xs, w, fma_mode);
    if (ys_in) memcpy(ys, ys_in, sizeof(float) * h); else d2s_oracle_linspace(ys, h, fma_mode);
    const float two_over_wm1 = (float)(2.0 / (double)(w - 1));
    const float strength = (float)0.05;

    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            float spx, snorm;
            shift_chain(depth[(size_t)y * w + x], depth_dt, convergence, depth_ratio, ipd_uv_w, strength,
                        two_over_wm1, &spx, &snorm);
            for (int eye = 0; eye < 2; ++eye) {
                float *dst = eyes + (size_t)eye * 3 * plane + (size_t)y * w + x;
                int32_t *idx = eye ? idx_right : idx_left;
                if (warp_mode == D2S_WARP_GATHER) {
                    /* depth.py:2163-2172 */
                    float c = eye ? (float)x - spx : (float)x + spx;
                    c = fminf(fmaxf(c, 0.f), (float)(w - 1));
                    int ci = (int)c;                                 /* .long() truncates */
                    if (idx) idx[(size_t)y * w + x] = ci;
                    for (int ch = 0; ch < 3; ++ch) {
                        float v = rgb[ch * plane + (size_t)y * w + ci];
                        dst[ch * plane] = fminf(fmaxf(v, 0.f), 255.f);   /* img clamp :2142 */
                    }
                } else {
                    /* depth.py:2152-2160 */
                    float gx = eye ? xs[x] - snorm : xs[x] + snorm;
                    float gy = ys[y];
                    float ix = source_index(gx, w), iy = source_index(gy, h);
                    int ix_nw = (int)floorf(ix), iy_nw = (int)floorf(iy);
                    int ix_ne = ix_nw + 1, iy_ne = iy_nw, ix_sw = ix_nw, iy_sw = iy_nw + 1;
                    int ix_se = ix_nw + 1, iy_se = iy_nw + 1;
                    float nw = ((float)ix_se - ix) * ((float)iy_se - iy);
                    float ne = (ix - (float)ix_sw) * ((float)iy_sw - iy);
                    float sw = ((float)ix_ne - ix) * (iy - (float)iy_ne);
                    float se = (ix - (float)ix_nw) * (iy - (float)iy_nw);
                    if (idx) idx[(size_t)y * w + x] = ix_nw;
                    for (int ch = 0; ch < 3; ++ch) {
                        const float *p = rgb + ch * plane;
                        float acc = 0.f;
#define TAP(yy, xx, wt)                                                                    \
    if ((yy) >= 0 && (yy) < h && (xx) >= 0 && (xx) < w) {                                  \
        float v = fminf(fmaxf(p[(size_t)(yy) * w + (xx)], 0.f), 255.f);                    \
        acc = fma_mode ? fmaf(v, (wt), acc) : acc + v * (wt);                              \
    }
                        TAP(iy_nw, ix_nw, nw) TAP(iy_ne, ix_ne, ne) TAP(iy_sw, ix_sw, sw) TAP(iy_se, ix_se, se)
#undef TAP
                        dst[ch * plane] = acc;
                    }
                }
            }
        }
    }

    /* depth.py:2175-2184: pad each eye, cat, area-downsample (exact 2:1 mean), clamp */
    int ph, pw, top, left; pad_geometry(h, w, fill_16_9, &ph, &pw, &top, &left);
    int tab = (display_mode == D2S_FULL_TAB || display_mode == D2S_HALF_TAB);
    int half = (display_mode == D2S_HALF_SBS || display_mode == D2S_HALF_TAB);
    int ch_ = tab ? 2 * ph : ph, cw = tab ? pw : 2 * pw;
    int oh = half ? ph : ch_, ow = half ? pw : cw;
    for (int c = 0; c < 3; ++c)
        for (int y = 0; y < oh; ++y)
            for (int x = 0; x < ow; ++x) {
                float v;
                int yy0 = y, xx0 = x, yy1 = y, xx1 = x;
                if (half) { if (tab) { yy0 = 2 * y; yy1 = 2 * y + 1; } else { xx0 = 2 * x; xx1 = 2 * x + 1; } }
                float a, b;
                {
                    int e, ey, ex;
                    /* element (yy0,xx0) of the concatenated padded image */
                    if (tab) { e = yy0 >= ph; ey = yy0 - e * ph; ex = xx0; } else { e = xx0 >= pw; ex = xx0 - e * pw; ey = yy0; }
                    ey -= top; ex -= left;
                    a = (ey >= 0 && ey < h && ex >= 0 && ex < w) ? eyes[(size_t)e * 3 * plane + c * plane + (size_t)ey * w + ex] : 0.f;
                    if (tab) { e = yy1 >= ph; ey = yy1 - e * ph; ex = xx1; } else { e = xx1 >= pw; ex = xx1 - e * pw; ey = yy1; }
                    ey -= top; ex -= left;
                    b = (ey >= 0 && ey < h && ex >= 0 && ex < w) ? eyes[(size_t)e * 3 * plane + c * plane + (size_t)ey * w + ex] : 0.f;
                }
                v = half ? (a + b) * 0.5f : a;      /* adaptive_avg_pool2d: sum / 2 (fp32 accumulate) */
                v = round_dt(v, out_dt);            /* result tensor dtype (gather branch: the eye dtype) */
                out[(size_t)c * oh * ow + (size_t)y * ow + x] = fminf(fmaxf(v, 0.f), 255.f);
            }
    free(eyes); free(xs); free(ys);
    return 0;
}
