/* ORACLE — test infrastructure, never on the product path.
 *
 * CPU restatement of the reference's occlusion-aware stereo renderer: the GLSL fragment shader FRAGMENT_SHADER of
 * /root/reference/viewer.py:386-631, evaluated per output pixel of each eye view exactly as the GPU would evaluate it per fragment:
 *   main()                         viewer.py:534-631   (3-tap depth smoothing :545-549, depth shaping :554, edge falloff :559-563,
 *                                                       parallax shift :563-564, confidence blend :567-576, border alpha :582-583,
 *                                                       feathering :586-613, rounded corners :617-626)
 *   disocclusion_confidence()      viewer.py:421-435
 *   push_pull_inpaint()            viewer.py:437-506
 * and of the fixed-function pieces the shader relies on, from the OpenGL 3.3 specification:
 *   texture()                      GL_LINEAR filtering, GL_REPEAT wrapping (moderngl's texture defaults; viewer.py:2385-2386 creates
 *                                  color_tex as 3 x u8 normalised and depth_tex as 1 x f32 and sets neither filter nor wrap),
 *                                  texel centres at (i + 0.5) / size, exact fp32 weights (real GPUs quantise them to 8 bits)
 *   smoothstep(), mix(), sign()    GLSL 3.30 built-ins
 *   the two eye passes             viewer.py:2680-2760: left view u_eye_offset = -ipd/2, right view +ipd/2, u_depth_strength =
 *                                  0.1 * depth_ratio (viewer.py:1334, 2686), each rendered into its own viewport of the packed frame
 *
 * PARITY UNPINNED.  There is no tensor oracle for this path: it needs an OpenGL context (none in the build container), and the
 * reference never sets `u_resolution` (nor `u_roll`): grep finds no assignment in the tree, so in the shipped application
 * pixel_size = 1.0 / vec2(0) and the smoothing / confidence / inpaint taps land at non-finite coordinates whose result is
 * driver-defined.  This restatement takes u_resolution as a parameter with the value the shader's own comment gives it
 * ("viewport resolution").  It is cross-checked against an independent numpy restatement (tests/test_oracle_dibr.py).
 *
 * Strict IEEE fp32, compile with -ffp-contract=off.  exp() terms and cos/sin(roll) are passed in as fp32 tables/values computed
 * by the caller in double precision, so that this file and the CUDA kernel evaluate identical arithmetic.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

typedef struct { float x, y, z, w; } vec4;

typedef struct {
    const float *color;   /* [h][w][3], values in [0,1] (the normalised u8 texture) */
    const float *depth;   /* [h][w] */
    int w, h;
    int vw, vh;           /* eye view (viewport) size */
    float res_x, res_y;   /* u_resolution */
    float eye_offset, depth_strength, convergence, c, s;
    int search_radius;
    float depth_tolerance, blur_radius;
    const float *w1, *w2; /* exp(-i*0.15), exp(-i*0.2), i = 0..search_radius */
    int feather_enabled;
    float feather_width, corner_radius;
} dibr_t;

static int wrapi(float f, int n) {          /* GL_REPEAT on an integer texel coordinate held in a float */
    float r = fmodf(f, (float)n);
    if (r < 0.f) r += (float)n;
    int i = (int)r;
    return i >= n ? n - 1 : i;
}

/* texture(sampler2D, uv) — bilinear, repeat.  ch = channels of the image, out = ch values */
static void tex(const float *img, int w, int h, int ch, float u, float v, float *out) {
    float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    float x0 = floorf(x), y0 = floorf(y);
    float fx = x - x0, fy = y - y0;
    int ix0 = wrapi(x0, w), ix1 = wrapi(x0 + 1.f, w), iy0 = wrapi(y0, h), iy1 = wrapi(y0 + 1.f, h);
    for (int c = 0; c < ch; ++c) {
        float t00 = img[((size_t)iy0 * w + ix0) * ch + c], t10 = img[((size_t)iy0 * w + ix1) * ch + c];
        float t01 = img[((size_t)iy1 * w + ix0) * ch + c], t11 = img[((size_t)iy1 * w + ix1) * ch + c];
        float top = t00 * (1.f - fx) + t10 * fx, bot = t01 * (1.f - fx) + t11 * fx;
        out[c] = top * (1.f - fy) + bot * fy;
    }
}
static float texd(const dibr_t *p, float u, float v) { float r; tex(p->depth, p->w, p->h, 1, u, v, &r); return r; }
static vec4 texc(const dibr_t *p, float u, float v) { float r[3]; tex(p->color, p->w, p->h, 3, u, v, r); vec4 o = {r[0], r[1], r[2], 1.f}; return o; }

static float clamp01(float t) { return fminf(fmaxf(t, 0.f), 1.f); }
static float smoothstep(float e0, float e1, float x) {
    float t = clamp01((x - e0) / (e1 - e0));
    return t * t * (3.f - 2.f * t);
}
static float signf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

/* viewer.py:421-435 */
static float disocclusion_confidence(const dibr_t *p, float bx, float by, float sx, float sy, float pdx, float pdy, float psx, float psy) {
    if (sx < 0.f || sx > 1.f || sy < 0.f || sy > 1.f) return 1.f;
    float s2x = pdx * psx * 2.f, s2y = pdy * psy * 2.f;
    float dl = texd(p, bx - s2x, by - s2y), dr = texd(p, bx + s2x, by + s2y);
    return smoothstep(0.04f, 0.10f, fabsf(dl - dr));
}

/* viewer.py:437-506 */
static vec4 push_pull_inpaint(const dibr_t *p, float ux, float uy, float center, float pdx, float pdy, float psx, float psy, float sweep_sign) {
    vec4 best = {0.f, 0.f, 0.f, 0.f};
    float bw = 0.f;
    float swx = pdx * psx * sweep_sign, swy = pdy * psx * sweep_sign;      /* g_par_dir * pixel_size.x * g_sweep_sign */
    for (int i = 1; i <= p->search_radius; ++i) {
        float qx = ux + swx * (float)i, qy = uy + swy * (float)i;
        if (qx < 0.f || qy < 0.f || qx > 1.f || qy > 1.f) continue;
        float sdi = 1.f - texd(p, qx, qy);
        if (sdi > center + p->depth_tolerance) {
            vec4 sc = texc(p, qx, qy);
            float dw = 1.f + (sdi - center) * 10.f;
            float wgt = p->w1[i] * dw;
            best.x += sc.x * wgt; best.y += sc.y * wgt; best.z += sc.z * wgt; best.w += sc.w * wgt;
            bw += wgt;
            if (bw > 5.f) break;
        }
    }
    if (bw < 2.f) {
        for (int i = 1; i <= p->search_radius; ++i) {
            float qx = ux - swx * (float)i, qy = uy - swy * (float)i;
            if (qx < 0.f || qy < 0.f || qx > 1.f || qy > 1.f) continue;
            float sdi = 1.f - texd(p, qx, qy);
            if (sdi > center + p->depth_tolerance) {
                vec4 sc = texc(p, qx, qy);
                float wgt = p->w2[i];
                best.x += sc.x * wgt; best.y += sc.y * wgt; best.z += sc.z * wgt; best.w += sc.w * wgt;
                bw += wgt;
            }
        }
    }
    if (bw > 0.01f) {
        vec4 acc = {best.x / bw * 0.5f, best.y / bw * 0.5f, best.z / bw * 0.5f, best.w / bw * 0.5f};
        float vwgt = 0.5f;
        for (int dy = -1; dy <= 1; dy += 2) {
            float vy = uy + (float)dy * psy * p->blur_radius;
            if (vy >= 0.f && vy <= 1.f) {
                float vdi = 1.f - texd(p, ux, vy);
                if (vdi > center + p->depth_tolerance * 0.5f) {
                    vec4 sc = texc(p, ux, vy);
                    acc.x += sc.x * 0.25f; acc.y += sc.y * 0.25f; acc.z += sc.z * 0.25f; acc.w += sc.w * 0.25f;
                    vwgt += 0.25f;
                }
            }
        }
        vec4 o = {acc.x / vwgt, acc.y / vwgt, acc.z / vwgt, acc.w / vwgt};
        return o;
    }
    return texc(p, ux, uy);
}

/* one fragment: view column j, view row i (from the top).  out = r,g,b,a (colour NOT multiplied by alpha), plus conf */
static void fragment(const dibr_t *p, int j, int i, float *out, float *conf_out) {
    float uvx = ((float)j + 0.5f) / (float)p->vw;
    float uvy = ((float)(p->vh - 1 - i) + 0.5f) / (float)p->vh;   /* gl_FragCoord.y counts from the bottom */
    float fx = uvx, fy = 1.f - uvy;                               /* flipped_uv */
    float psx = 1.f / p->res_x, psy = 1.f / p->res_y;
    float sg = signf(p->eye_offset);
    float pdx = p->c * sg, pdy = p->s * sg;
    float sweep_sign = p->eye_offset > 0.f ? -1.f : 1.f;
    float dsx = pdx * psx * 1.5f, dsy = pdy * psy * 1.5f;
    float d0 = texd(p, fx, fy), dm = texd(p, fx - dsx, fy - dsy), dp = texd(p, fx + dsx, fy + dsy);
    float depth = d0 * 0.7f + dm * 0.15f + dp * 0.15f;
    float depth_inv = -depth;
    float shaped = depth_inv * (1.f + 0.35f * (1.f - depth));
    float shift = shaped + p->convergence;
    float margin = 0.05f;
    float falloff = smoothstep(0.f, margin, fx) * smoothstep(1.f, 1.f - margin, fx);
    float px = p->eye_offset * shift * p->depth_strength * falloff;
    float sx = fx - px * p->c, sy = fy - px * p->s;
    float conf = disocclusion_confidence(p, fx, fy, sx, sy, pdx, pdy, psx, psy);
    vec4 col = texc(p, sx, sy);
    if (conf > 0.001f) {
        vec4 f = push_pull_inpaint(p, fx, fy, depth_inv, pdx, pdy, psx, psy, sweep_sign);
        col.x = col.x * (1.f - conf) + f.x * conf; col.y = col.y * (1.f - conf) + f.y * conf; col.z = col.z * (1.f - conf) + f.z * conf;
    }
    float bxa = smoothstep(-0.001f, 0.001f, sx) * smoothstep(1.001f, 0.999f, sx);
    float bya = smoothstep(-0.001f, 0.001f, sy) * smoothstep(1.001f, 0.999f, sy);
    float alpha = fminf(bxa, bya);
    if (p->feather_enabled) {
        float f = p->feather_width;
        float fo = smoothstep(0.f, f, uvx) * smoothstep(0.f, f, 1.f - uvx) * smoothstep(0.f, f, uvy) * smoothstep(0.f, f, 1.f - uvy);
        fo = powf(fo, 0.7f);
        col.x *= fo; col.y *= fo; col.z *= fo;
    }
    float r = p->corner_radius;
    float dx = fabsf(uvx - 0.5f) - 0.5f + r, dy = fabsf(uvy - 0.5f) - 0.5f + r;
    float mx = fmaxf(dx, 0.f), my = fmaxf(dy, 0.f);
    float sdf = sqrtf(mx * mx + my * my) + fminf(fmaxf(dx, dy), 0.f) - r;
    alpha = fminf(alpha, 1.f - smoothstep(0.f, 0.01f, sdf));
    out[0] = col.x; out[1] = col.y; out[2] = col.z; out[3] = alpha;
    if (conf_out) *conf_out = conf;
}

/* Render both eye views.  rgba_left / rgba_right: [vh][vw][4] fp32 (colour in [0,1], alpha), conf_*: optional [vh][vw]. */
int d2s_oracle_dibr(const float *color, const float *depth, int h, int w, int vw, int vh, float res_x, float res_y, float ipd_uv,
                    float depth_strength, float convergence, float c, float s, int search_radius, float depth_tolerance, float blur_radius,
                    const float *w1, const float *w2, int feather_enabled, float feather_width, float corner_radius,
                    float *rgba_left, float *rgba_right, float *conf_left, float *conf_right) {
    dibr_t p = {color, depth, w, h, vw, vh, res_x, res_y, 0.f, depth_strength, convergence, c, s, search_radius, depth_tolerance, blur_radius,
                w1, w2, feather_enabled, feather_width, corner_radius};
    for (int e = 0; e < 2; ++e) {
        p.eye_offset = e ? ipd_uv / 2.0f : -ipd_uv / 2.0f;
        float *o = e ? rgba_right : rgba_left, *cf = e ? conf_right : conf_left;
        for (int i = 0; i < vh; ++i)
            for (int j = 0; j < vw; ++j) fragment(&p, j, i, o + ((size_t)i * vw + j) * 4, cf ? cf + (size_t)i * vw + j : 0);
    }
    return 0;
}
