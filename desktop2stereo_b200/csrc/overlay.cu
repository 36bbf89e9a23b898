// FPS overlay (reference depth.py:2061-2103 overlay_fps, font depth.py:641-658, build_font :2029-2054).
//
// The reference builds an [H,W] mask of 0/1 from 5x3 bitmap glyphs scaled by max(1, min(8, H // 60)) and returns
// rgb * (1 - mask) + [0,255,0] * mask.  With a 0/1 mask that is "glyph pixels become (0,255,0), everything else is
// untouched" (x*1 + c*0 == x and x*0 + c*1 == c exactly for finite x), so the kernel only visits the text rectangle.
#include <cstring>

#include "common.cuh"

namespace d2s {

struct OverlayK {
    void *base; long long sc, sy, sx;
    int dtype, h, w, scale, n;
    unsigned short glyph[32];     // 15-bit bitmaps, bit (row*3 + col), row 0 = top, col 0 = left
};

template <typename T>
__device__ __forceinline__ void put_px(const OverlayK &k, int y, int x) {
    T *p = (T *)k.base + (long long)y * k.sy + (long long)x * k.sx;
    p[0] = from_f32<T>(0.f); p[k.sc] = from_f32<T>(255.f); p[2 * k.sc] = from_f32<T>(0.f);
}

__global__ void overlay_fps_kernel(const OverlayK k) {
    const int rx = blockIdx.x * blockDim.x + threadIdx.x, ry = blockIdx.y * blockDim.y + threadIdx.y;
    const int cw = 3 * k.scale, pitch = 4 * k.scale;            // char_w, char_w + spacing
    if (ry >= 5 * k.scale || rx >= k.n * pitch) return;
    const int y = 2 * k.scale + ry, x = 2 * k.scale + rx;         // margin_y, margin_x
    if (y >= k.h || x >= k.w) return;                             // x1 = min(W, ..), y1 = min(H, ..)
    const int i = rx / pitch, gx = rx - i * pitch;
    if (gx >= cw) return;
    if (!((k.glyph[i] >> ((ry / k.scale) * 3 + gx / k.scale)) & 1)) return;
    switch (k.dtype) {
        case D2S_F32: put_px<float>(k, y, x); break;
        case D2S_F16: put_px<__half>(k, y, x); break;
        case D2S_BF16: put_px<__nv_bfloat16>(k, y, x); break;
        default: put_px<uint8_t>(k, y, x); break;
    }
}

// depth.py:641-658
static const char *kFontChars = "0123456789FPS:. ";
static const char *kFontRows[16][5] = {
    {"111", "101", "101", "101", "111"}, {"010", "110", "010", "010", "111"}, {"111", "001", "111", "100", "111"},
    {"111", "001", "111", "001", "111"}, {"101", "101", "111", "001", "001"}, {"111", "100", "111", "001", "111"},
    {"111", "100", "111", "101", "111"}, {"111", "001", "010", "100", "100"}, {"111", "101", "111", "101", "111"},
    {"111", "101", "111", "001", "111"}, {"111", "100", "110", "100", "100"}, {"110", "101", "110", "100", "100"},
    {"111", "100", "111", "001", "111"}, {"000", "010", "000", "010", "000"}, {"000", "000", "000", "000", "010"},
    {"000", "000", "000", "000", "000"},
};

}  // namespace d2s

using namespace d2s;

extern "C" int d2s_overlay_fps(const d2s_image *rgb, int h, int w, const char *text, d2s_stream_t stream) {
    D2S_REQUIRE(rgb && rgb->base && text, "d2s_overlay_fps: null argument");
    D2S_REQUIRE(h >= 1 && w >= 1, "d2s_overlay_fps: image size %dx%d", h, w);
    D2S_REQUIRE(rgb->dtype == D2S_F32 || rgb->dtype == D2S_F16 || rgb->dtype == D2S_BF16 || rgb->dtype == D2S_U8, "d2s_overlay_fps: dtype %d", rgb->dtype);
    const size_t n = strlen(text);
    D2S_REQUIRE(n <= 32, "d2s_overlay_fps: text longer than 32 characters");
    if (n == 0) return D2S_OK;
    OverlayK k{};
    k.base = rgb->base; k.sc = rgb->sc; k.sy = rgb->sy; k.sx = rgb->sx; k.dtype = rgb->dtype;
    k.h = h; k.w = w; k.n = (int)n;
    int s = h / 60; if (s > 8) s = 8; if (s < 1) s = 1;           // scale = max(1, min(8, H // 60))
    k.scale = s;
    for (size_t i = 0; i < n; ++i) {
        const char *f = strchr(kFontChars, text[i]);
        const int ci = (f && text[i]) ? (int)(f - kFontChars) : 15;   // unknown characters render as " " (depth.py:2075)
        unsigned short bits = 0;
        for (int r = 0; r < 5; ++r)
            for (int c = 0; c < 3; ++c)
                if (kFontRows[ci][r][c] == '1') bits |= (unsigned short)(1u << (r * 3 + c));
        k.glyph[i] = bits;
    }
    dim3 block(32, 8), grid(ceil_div((long long)k.n * 4 * s, 32), ceil_div(5 * s, 8));
    D2S_LAUNCH(overlay_fps_kernel, grid, block, 0, stream, k);
    D2S_POST_LAUNCH();
    return D2S_OK;
}
