// Packed RGB u8 frame -> NV12 (planar Y, interleaved CbCr at half resolution), 1.5 bytes per pixel — SURVEY §8f N3.
//
// The reference hands the packed float32 frame to MJPEGStreamer.set_frame, whose encoder thread runs cv2.imencode(".jpg")
// (reference streamer.py:201-228, 250-256); make_sbs's device->host copy of that float32 frame (depth.py:767-773, 199 MB per 4K
// Full-SBS frame) is the largest cost of the Legacy-Streamer mode.  This kernel does the first two stages of that JPEG encoder on
// the device — libjpeg's RGB -> YCbCr colour conversion (JFIF full range, BT.601, 16-bit fixed point, jccolor.c) and its h2v2
// chroma downsample (2x2 mean with the alternating 1,2 rounding bias, jcsample.c) — so that 8x fewer bytes than the float32 frame
// cross PCIe, in the layout NVENC / nvJPEG hardware encoders take directly.  Integer arithmetic: bit-exact against oracle/nv12.py.
#include "common.cuh"

namespace d2s {

__device__ __forceinline__ int y_of(int r, int g, int b) { return (19595 * r + 38470 * g + 7471 * b + 32768) >> 16; }
__device__ __forceinline__ int cb_of(int r, int g, int b) { return (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16; }
__device__ __forceinline__ int cr_of(int r, int g, int b) { return (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16; }

// one thread: 2 rows x 4 columns of the frame (two chroma samples).  rgb: HWC u8 with row pitch `pitch` bytes.
__global__ void __launch_bounds__(256) rgb_to_nv12_kernel(const uint8_t *__restrict__ rgb, long long pitch, int h, int w, uint8_t *__restrict__ yp,
                                                          uint8_t *__restrict__ uvp) {
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y0 = blockIdx.y * 2;
    if (x0 >= w) return;
    const int n = min(4, w - x0);                    // w is even: n is 2 or 4
    int cb[2][4], cr[2][4];
    uint8_t yv[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const uint8_t *p = rgb + (long long)(y0 + r) * pitch + (long long)x0 * 3;
        uint8_t px[12];
        if (n == 4 && (((uintptr_t)p) & 3) == 0) {
            const uint32_t *q = (const uint32_t *)p;
            const uint32_t a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
            *(uint32_t *)&px[0] = a; *(uint32_t *)&px[4] = b; *(uint32_t *)&px[8] = c;
        } else {
            for (int i = 0; i < 3 * n; ++i) px[i] = __ldg(p + i);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int R = px[3 * i], G = px[3 * i + 1], B = px[3 * i + 2];
            yv[r][i] = (uint8_t)y_of(R, G, B); cb[r][i] = cb_of(R, G, B); cr[r][i] = cr_of(R, G, B);
        }
        uint8_t *yo = yp + (long long)(y0 + r) * w + x0;
        if (n == 4 && (((uintptr_t)yo) & 3) == 0) *(uint32_t *)yo = (uint32_t)yv[r][0] | ((uint32_t)yv[r][1] << 8) | ((uint32_t)yv[r][2] << 16) | ((uint32_t)yv[r][3] << 24);
        else for (int i = 0; i < n; ++i) yo[i] = yv[r][i];
    }
    uint8_t *uo = uvp + (long long)(y0 / 2) * w + x0;       // UV row pitch = w bytes (w/2 pairs)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (2 * c >= n) break;
        const int bias = (((x0 >> 1) + c) & 1) ? 2 : 1;     // jcsample.c h2v2_downsample: bias 1, 2, 1, 2, ... along the output row
        uo[2 * c] = (uint8_t)((cb[0][2 * c] + cb[0][2 * c + 1] + cb[1][2 * c] + cb[1][2 * c + 1] + bias) >> 2);
        uo[2 * c + 1] = (uint8_t)((cr[0][2 * c] + cr[0][2 * c + 1] + cr[1][2 * c] + cr[1][2 * c + 1] + bias) >> 2);
    }
}

int rgb_to_nv12_launch(const uint8_t *rgb, long long pitch, int h, int w, uint8_t *yp, uint8_t *uvp, cudaStream_t stream) {
    D2S_REQUIRE(rgb && yp && uvp && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0, "d2s_rgb_to_nv12: frame %dx%d must be even-sized", h, w);
    dim3 grid(ceil_div(ceil_div(w, 4), 256), h / 2);
    D2S_REQUIRE(grid.y <= 65535, "d2s_rgb_to_nv12: frame height %d", h);
    D2S_LAUNCH(rgb_to_nv12_kernel, grid, 256, 0, stream, rgb, pitch, h, w, yp, uvp);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

}  // namespace d2s

extern "C" int d2s_rgb_to_nv12(const uint8_t *rgb_hwc, int64_t row_pitch_bytes, int h, int w, uint8_t *nv12, d2s_stream_t stream) {
    D2S_REQUIRE(nv12 != nullptr, "d2s_rgb_to_nv12: null output");
    return d2s::rgb_to_nv12_launch(rgb_hwc, row_pitch_bytes > 0 ? row_pitch_bytes : (long long)w * 3, h, w, nv12, nv12 + (size_t)h * w, (cudaStream_t)stream);
}
