// Baseline JPEG encoder on the device — SURVEY §8f N3: the encode the reference's MJPEGStreamer runs on the host for every frame,
// cv2.imencode(".jpg", bgr, [IMWRITE_JPEG_QUALITY, q]) (reference streamer.py:250-256), moved in front of the device->host copy so
// that a compressed stream (typically 0.1-0.5 bytes per pixel) crosses PCIe instead of make_sbs's float32 frame (12 bytes per pixel,
// depth.py:767-773).  The stream is what OpenCV's bundled libjpeg-turbo writes for the same frame, quality and restart interval,
// byte for byte (tests/test_jpeg_gpu.py compares with cv2.imencode itself): JFIF 1.01, YCbCr 4:2:0, the Annex K quantisation tables
// scaled by the quality, the Annex K Huffman tables, the integer "islow" DCT.  The one difference from the reference's call is that
// the device encoder always uses restart intervals (DRI) — they are what makes entropy coding parallel — and restart markers do not
// change a single decoded pixel.
//
// Four kernels per frame, all integer arithmetic:
//   transform   CTA = 8 MCUs (128x16 pixels).  RGB -> YCbCr (jccolor.c fixed point) and the h2v2 chroma mean (jcsample.c, bias
//               1,2,1,2) into shared memory with libjpeg's edge rules (right edge replicated at full resolution, bottom edge
//               replicated after downsampling), then 8 threads per 8x8 block run jfdctint.c's row and column passes and
//               jcdctmgr.c's round-half-away quantisation; coefficients leave in zig-zag order, one coalesced 16-byte store a thread.
//   entropy     one warp per restart interval, the 64 coefficients of a block coded side by side (ballot -> zero runs -> codewords ->
//               warp scan -> OR into a shared-memory bit buffer): jchuff.c's encode_one_block, 0xFF stuffing and the 1-bit padding
//               included, into a private slot sized for the worst case (so it cannot overflow); writes the slot's length.
//   scan        exclusive prefix sum of (length + 2 marker bytes) over the intervals: one CTA.
//   gather      one warp per interval copies its slot behind the header and appends RSTn (or EOI after the last); writes the size.
// Luma blocks past the component's own block grid (1080 rows = 135 block rows, but 68 MCU rows hold 136) are libjpeg's "dummy
// blocks" (jccoefct.c): AC = 0 and DC = the previous block's DC; the entropy kernel synthesises them.
#include <string.h>

#include "common.cuh"

namespace d2s {

constexpr int JG = 8;                 // MCUs per CTA of the transform kernel
constexpr int JT = JG * 6 * 8;        // 8 threads per 8x8 block
constexpr int kBlockWorstBytes = 416; // 20 + 63 * 26 bits, every byte stuffed

struct JpegQuant { uint16_t div[2][64]; uint32_t rcp[2][64]; uint8_t izz[64]; };   // 8 * quantval (natural order), floor(2^32 / div) + 1, natural index -> zig-zag position
struct JpegHuff { uint32_t dc[2][16]; uint32_t ac[2][256]; };   // code << 5 | size
struct JpegHeader { uint8_t bytes[640]; int len; };

// ---------------------------------------------------------------------------------------------------------------- tables (host)
static const uint8_t kBaseQ[2][64] = {   // Annex K.1 / K.2 in zig-zag order
    {16, 11, 12, 14, 12, 10, 16, 14, 13, 14, 18, 17, 16, 19, 24, 40, 26, 24, 22, 22, 24, 49, 35, 37, 29, 40, 58, 51, 61, 60, 57, 51,
     56, 55, 64, 72, 92, 78, 64, 68, 87, 69, 55, 56, 80, 109, 81, 87, 95, 98, 103, 104, 103, 62, 77, 113, 121, 112, 100, 120, 92, 101, 103, 99},
    {17, 18, 18, 24, 21, 24, 47, 26, 26, 47, 99, 66, 56, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99}};
static const uint8_t kBits[4][16] = {    // Annex K.3: DC luma, AC luma, DC chroma, AC chroma
    {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0},
    {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 125},
    {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0},
    {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 119}};
static const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t kAcLuma[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1,
    0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56,
    0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85,
    0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa,
    0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
    0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
static const uint8_t kAcChroma[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42,
    0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19,
    0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55,
    0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8,
    0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4,
    0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

static void zigzag_positions(uint8_t *izz) {                     // izz[natural index] = position in the zig-zag scan
    int r = 0, c = 0;
    for (int k = 0; k < 64; ++k) {
        izz[r * 8 + c] = (uint8_t)k;
        if (((r + c) & 1) == 0) { if (c == 7) ++r; else if (r == 0) ++c; else { --r; ++c; } }
        else                    { if (r == 7) ++c; else if (c == 0) ++r; else { ++r; --c; } }
    }
}

static void derive_huffman(const uint8_t *bits, const uint8_t *vals, uint32_t *packed, int n_packed) {
    for (int i = 0; i < n_packed; ++i) packed[i] = 0;
    uint32_t code = 0;
    int k = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int i = 0; i < bits[len - 1]; ++i, ++k) packed[vals[k]] = (code++ << 5) | (uint32_t)len;
        code <<= 1;
    }
}

static uint8_t *put16(uint8_t *p, int v) { p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)v; return p + 2; }

// jcparam.c quality scaling + jcmarker.c's header sequence
static void build_tables(int h, int w, int quality, int restart_interval, JpegQuant &jq, JpegHuff &jh, JpegHeader &hd) {
    const int q = quality < 1 ? 1 : quality > 100 ? 100 : quality;
    const int scale = q < 50 ? 5000 / q : 200 - 2 * q;
    uint8_t qz[2][64];
    for (int t = 0; t < 2; ++t)
        for (int i = 0; i < 64; ++i) {
            const long v = ((long)kBaseQ[t][i] * scale + 50) / 100;
            qz[t][i] = (uint8_t)(v < 1 ? 1 : v > 255 ? 255 : v);
        }
    zigzag_positions(jq.izz);
    for (int t = 0; t < 2; ++t)
        for (int n = 0; n < 64; ++n) {
            jq.div[t][n] = (uint16_t)(qz[t][jq.izz[n]] << 3);
            jq.rcp[t][n] = (uint32_t)((1ull << 32) / jq.div[t][n]) + 1u;   // a / div == umulhi(a, rcp) while a * div < 2^32 (here a < 2^17, div < 2^11)
        }
    derive_huffman(kBits[0], kDcVals, jh.dc[0], 16);
    derive_huffman(kBits[2], kDcVals, jh.dc[1], 16);
    derive_huffman(kBits[1], kAcLuma, jh.ac[0], 256);
    derive_huffman(kBits[3], kAcChroma, jh.ac[1], 256);

    uint8_t *p = hd.bytes;
    static const uint8_t soi_app0[] = {0xFF, 0xD8, 0xFF, 0xE0, 0x00, 0x10, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
    memcpy(p, soi_app0, sizeof(soi_app0)); p += sizeof(soi_app0);
    for (int t = 0; t < 2; ++t) { *p++ = 0xFF; *p++ = 0xDB; p = put16(p, 67); *p++ = (uint8_t)t; memcpy(p, qz[t], 64); p += 64; }
    *p++ = 0xFF; *p++ = 0xC0; p = put16(p, 17); *p++ = 8; p = put16(p, h); p = put16(p, w); *p++ = 3;
    *p++ = 1; *p++ = 0x22; *p++ = 0; *p++ = 2; *p++ = 0x11; *p++ = 1; *p++ = 3; *p++ = 0x11; *p++ = 1;
    const uint8_t *vals[4] = {kDcVals, kAcLuma, kDcVals, kAcChroma};
    const int nvals[4] = {12, 162, 12, 162}, ids[4] = {0x00, 0x10, 0x01, 0x11};
    for (int t = 0; t < 4; ++t) {
        *p++ = 0xFF; *p++ = 0xC4; p = put16(p, 2 + 1 + 16 + nvals[t]); *p++ = (uint8_t)ids[t];
        memcpy(p, kBits[t], 16); p += 16; memcpy(p, vals[t], nvals[t]); p += nvals[t];
    }
    *p++ = 0xFF; *p++ = 0xDD; p = put16(p, 4); p = put16(p, restart_interval);
    static const uint8_t sos[] = {0xFF, 0xDA, 0x00, 0x0C, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 0x3F, 0};
    memcpy(p, sos, sizeof(sos)); p += sizeof(sos);
    hd.len = (int)(p - hd.bytes);
}

// ---------------------------------------------------------------------------------------------------------------- transform
__device__ __forceinline__ int jy_of(int r, int g, int b) { return (19595 * r + 38470 * g + 7471 * b + 32768) >> 16; }
__device__ __forceinline__ int jcb_of(int r, int g, int b) { return (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16; }
__device__ __forceinline__ int jcr_of(int r, int g, int b) { return (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16; }

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// jfdctint.c, one 8-point pass.  PASS 0: rows (results scaled up by PASS1_BITS = 2); PASS 1: columns (scaled back, net factor 8)
template <int PASS> __device__ __forceinline__ void fdct8(int *d) {
    const int t0 = d[0] + d[7], t7 = d[0] - d[7], t1 = d[1] + d[6], t6 = d[1] - d[6];
    const int t2 = d[2] + d[5], t5 = d[2] - d[5], t3 = d[3] + d[4], t4 = d[3] - d[4];
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    constexpr int SH = PASS ? 13 + 2 : 13 - 2;
    if (PASS == 0) { d[0] = (t10 + t11) << 2; d[4] = (t10 - t11) << 2; }
    else           { d[0] = descale(t10 + t11, 2); d[4] = descale(t10 - t11, 2); }
    int z1 = (t12 + t13) * 4433;
    d[2] = descale(z1 + t13 * 6270, SH);
    d[6] = descale(z1 + t12 * -15137, SH);
    z1 = t4 + t7;
    int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
    const int z5 = (z3 + z4) * 9633;
    const int a4 = t4 * 2446, a5 = t5 * 16819, a6 = t6 * 25172, a7 = t7 * 12299;
    z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
    z3 += z5; z4 += z5;
    d[7] = descale(a4 + z1 + z3, SH);
    d[5] = descale(a5 + z2 + z4, SH);
    d[3] = descale(a6 + z2 + z3, SH);
    d[1] = descale(a7 + z1 + z4, SH);
}

__global__ void __launch_bounds__(JT) jpeg_transform_kernel(const uint8_t *__restrict__ rgb, long long pitch, int h, int w, int mcus_x,
                                                            const __grid_constant__ JpegQuant q, int16_t *__restrict__ coef) {
    __shared__ int16_t sY[16][JG * 16 + 2];
    __shared__ int16_t sC[2][8][JG * 8 + 2];
    __shared__ int sW[JG * 6][8][9];
    __shared__ __align__(16) int16_t sO[JG * 6][64];
    __shared__ uint32_t sRcp[2][64];                                  // (lanes index these by their own column: not for the constant bank)
    __shared__ uint16_t sDiv[2][64];
    __shared__ uint8_t sIzz[64];
    if (threadIdx.x < 128) { sRcp[threadIdx.x >> 6][threadIdx.x & 63] = q.rcp[threadIdx.x >> 6][threadIdx.x & 63]; sDiv[threadIdx.x >> 6][threadIdx.x & 63] = q.div[threadIdx.x >> 6][threadIdx.x & 63]; }
    else if (threadIdx.x < 192) sIzz[threadIdx.x - 128] = q.izz[threadIdx.x - 128];
    const int my = blockIdx.y, mx0 = blockIdx.x * JG;
    const bool pairs_ok = ((pitch & 1) == 0) && ((((uintptr_t)rgb) & 1) == 0);

    // colour conversion + chroma mean, one 2x2 quad at a time
    for (int qd = threadIdx.x; qd < 8 * JG * 8; qd += JT) {
        const int qy = qd / (JG * 8), qx = qd % (JG * 8);
        const int x0 = mx0 * 16 + 2 * qx, cy = my * 8 + qy;
        const int cyc = min(cy, h / 2 - 1);                           // jcprepct.c: downsampled rows are replicated downwards
        int R[2][2], G[2][2], B[2][2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint8_t *row = rgb + (long long)(2 * cyc + r) * pitch;
            if (pairs_ok && x0 + 1 < w) {
                const uint16_t *p = (const uint16_t *)(row + (long long)x0 * 3);
                const uint32_t a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
                R[r][0] = a & 0xFF; G[r][0] = a >> 8; B[r][0] = b & 0xFF; R[r][1] = b >> 8; G[r][1] = c & 0xFF; B[r][1] = c >> 8;
            } else {
#pragma unroll
                for (int i = 0; i < 2; ++i) {                          // jcsample.c expand_right_edge: the last column is replicated
                    const uint8_t *p = row + (long long)min(x0 + i, w - 1) * 3;
                    R[r][i] = __ldg(p); G[r][i] = __ldg(p + 1); B[r][i] = __ldg(p + 2);
                }
            }
        }
        int cb = 0, cr = 0;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 2; ++i) { cb += jcb_of(R[r][i], G[r][i], B[r][i]); cr += jcr_of(R[r][i], G[r][i], B[r][i]); }
        const int bias = ((mx0 * 8 + qx) & 1) ? 2 : 1;                // h2v2_downsample: 1, 2, 1, 2, ... along the output row
        sC[0][qy][qx] = (int16_t)(((cb + bias) >> 2) - 128);
        sC[1][qy][qx] = (int16_t)(((cr + bias) >> 2) - 128);
        const bool below = cy != cyc;                                 // luma rows past the frame replicate the last row
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int s = below ? 1 : r;
#pragma unroll
            for (int i = 0; i < 2; ++i) sY[2 * qy + r][2 * qx + i] = (int16_t)(jy_of(R[s][i], G[s][i], B[s][i]) - 128);
        }
    }
    __syncthreads();

    const int blk = threadIdx.x >> 3, r = threadIdx.x & 7, m = blk / 6, b6 = blk % 6;
    int d[8];
    {
        const int16_t *s = b6 < 4 ? &sY[(b6 >> 1) * 8 + r][m * 16 + (b6 & 1) * 8] : &sC[b6 - 4][r][m * 8];
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = s[i];
    }
    fdct8<0>(d);
#pragma unroll
    for (int i = 0; i < 8; ++i) sW[blk][r][i] = d[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = sW[blk][i][r];
    fdct8<1>(d);
    const int tb = b6 < 4 ? 0 : 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {                                     // jcdctmgr.c quantize: round half away from zero
        const int nat = i * 8 + r, qv = sDiv[tb][nat];
        const int v = (int)__umulhi((uint32_t)(abs(d[i]) + (qv >> 1)), sRcp[tb][nat]);
        sO[blk][sIzz[nat]] = (int16_t)(d[i] < 0 ? -v : v);
    }
    __syncthreads();
    const int valid = min(JG, mcus_x - mx0) * 6 * 8;                  // 16-byte pieces of this CTA's MCUs
    if ((int)threadIdx.x < valid)
        ((uint4 *)(coef + ((size_t)my * mcus_x + mx0) * 6 * 64))[threadIdx.x] = ((const uint4 *)&sO[0][0])[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------------------------- entropy coding
// One WARP = one restart interval; the 64 coefficients of a block are coded side by side.  Lane l owns zig-zag positions l and
// l + 32: two ballots give the block's non-zero mask, from which every lane reads off the zero run in front of its coefficients
// (jchuff.c encode_one_block's run/size symbols, ZRLs in front when the run is 16+) and builds its <= 26-bit codeword; lane 0's first item is the DC
// difference, lane 31's second item is the EOB when position 63 is zero.  A warp scan of the lengths places the codewords, which
// are OR-ed into a big-endian bit buffer in shared memory; when the buffer might not hold another MCU (and at the end) the completed
// bytes go to the interval's slot with 0xFF stuffing (ballot + popc for the positions) and the last partial byte is carried over.
constexpr int EW = 8;                     // warps (restart intervals) per CTA
constexpr int kMcuBits = 6 * 1658;        // worst case of one MCU
constexpr int kBitWords = 640;            // bit buffer per warp: flushed when the next MCU might not fit (typically once per interval)

// OR `n` bits (right-aligned in v) into the buffer at bit offset o; stream bit i lives in word i / 32 at bit 31 - i % 32
__device__ __forceinline__ void put_item(uint32_t *buf, int o, uint32_t v, int n) {          // n <= 32: at most two words
    const uint32_t hi = v << (32 - n);
    const int w = o >> 5, s = o & 31;
    const uint32_t x0 = hi >> s, x1 = s ? hi << (32 - s) : 0u;
    if (x0) atomicOr(&buf[w], x0);
    if (x1) atomicOr(&buf[w + 1], x1);
}

// one AC coefficient t at zig-zag position k: v/n = Huffman code of (run % 16, size) followed by the value bits (<= 26 bits);
// nz = number of ZRL symbols (16 zeros each) that go in front of it
__device__ __forceinline__ void ac_item(int t, int k, unsigned long long mask, const uint32_t *ac, uint32_t &v, int &n, int &nz) {
    const unsigned long long below = mask & ((1ull << k) - 1ull);
    const int prev = below ? 63 - __clzll((long long)below) : 0;      // position 0 is the DC term
    const int run = k - 1 - prev;
    nz = run >> 4;
    const int nb = 32 - __clz(abs(t));
    const uint32_t sym = ac[((run & 15) << 4) | nb];
    v = ((sym >> 5) << nb) | ((uint32_t)(t < 0 ? t - 1 : t) & ((1u << nb) - 1u));
    n = (int)(sym & 31u) + nb;
}

__device__ __forceinline__ void put_ac(uint32_t *buf, int o, uint32_t v, int n, int nz, uint32_t zrl) {
    for (int i = 0; i < nz; ++i) { put_item(buf, o, zrl >> 5, (int)(zrl & 31u)); o += (int)(zrl & 31u); }   // rare: runs of 16+ zeros
    put_item(buf, o, v, n);
}

// completed bytes -> slot (with stuffing); the partial byte moves to the front of the cleared buffer
__device__ __forceinline__ void flush_bytes(uint32_t *buf, int &bitpos, uint8_t *out, uint32_t &outpos, int lane) {
    __syncwarp();
    const int nB = bitpos >> 3;
    uint32_t ffs = 0;
    for (int j0 = 0; j0 < nB; j0 += 32) {
        const int j = j0 + lane;
        const bool valid = j < nB;
        const uint32_t b = valid ? (buf[j >> 2] >> (24 - 8 * (j & 3))) & 0xFFu : 0u;
        const uint32_t ffm = __ballot_sync(0xffffffffu, valid && b == 0xFFu);
        if (valid) {
            const uint32_t p = outpos + (uint32_t)j + ffs + (uint32_t)__popc(ffm & ((1u << lane) - 1u));
            out[p] = (uint8_t)b;
            if (b == 0xFFu) out[p + 1] = 0;
        }
        ffs += (uint32_t)__popc(ffm);
    }
    outpos += (uint32_t)nB + ffs;
    const uint32_t carry = (bitpos & 7) ? (buf[nB >> 2] >> (24 - 8 * (nB & 3))) & 0xFFu : 0u;
    __syncwarp();
    for (int i = lane; i <= (bitpos >> 5) + 2 && i < kBitWords; i += 32) buf[i] = 0u;
    __syncwarp();
    if (lane == 0 && carry) buf[0] = carry << 24;
    bitpos &= 7;
    __syncwarp();
}

__global__ void __launch_bounds__(EW * 32) jpeg_entropy_kernel(const int16_t *__restrict__ coef, int n_mcus, int mcus_x, int yblk_w, int yblk_h,
                                                               int ri, int n_int, const __grid_constant__ JpegHuff hf,
                                                               uint8_t *__restrict__ slots, int slot_bytes, uint32_t *__restrict__ lens) {
    __shared__ uint32_t s_ac[2][256], s_dc[2][16];
    __shared__ uint32_t s_bits[EW][kBitWords];
    for (int i = threadIdx.x; i < 512; i += EW * 32) s_ac[i >> 8][i & 255] = hf.ac[i >> 8][i & 255];
    if (threadIdx.x < 32) s_dc[threadIdx.x >> 4][threadIdx.x & 15] = hf.dc[threadIdx.x >> 4][threadIdx.x & 15];
    for (int i = threadIdx.x; i < EW * kBitWords; i += EW * 32) (&s_bits[0][0])[i] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, it = blockIdx.x * EW + wid;
    if (it >= n_int) return;
    uint32_t *buf = s_bits[wid];
    uint8_t *out = slots + (size_t)it * slot_bytes;
    uint32_t outpos = 0;
    int bitpos = 0, last_y = 0, last_cb = 0, last_cr = 0;
    const int m0 = it * ri, m_end = min(n_mcus, m0 + ri);
    const int16_t *cp = coef + (size_t)m0 * 6 * 64;
    int nx0 = __ldg(cp + lane), nx1 = __ldg(cp + lane + 32);          // the next block's coefficients are always in flight
    for (int m = m0; m < m_end; ++m) {
        const int my = m / mcus_x, mx = m - my * mcus_x;
        int prev_dc = 0;
#pragma unroll 1
        for (int b = 0; b < 6; ++b) {
            int c0 = nx0, c1 = nx1;
            cp += 64;
            if (b < 5 || m + 1 < m_end) { nx0 = __ldg(cp + lane); nx1 = __ldg(cp + lane + 32); }
            const int tb = b < 4 ? 0 : 1;
            const bool dummy = b < 4 && (2 * my + (b >> 1) >= yblk_h || 2 * mx + (b & 1) >= yblk_w);   // jccoefct.c
            if (dummy) { c0 = 0; c1 = 0; }
            int dc = __shfl_sync(0xffffffffu, c0, 0);
            if (dummy) dc = prev_dc;
            prev_dc = dc;
            int diff;
            if (b < 4) { diff = dc - last_y; last_y = dc; } else if (b == 4) { diff = dc - last_cb; last_cb = dc; } else { diff = dc - last_cr; last_cr = dc; }
            const uint32_t mlo = __ballot_sync(0xffffffffu, c0 != 0 && lane != 0), mhi = __ballot_sync(0xffffffffu, c1 != 0);
            const unsigned long long mask = (unsigned long long)mlo | ((unsigned long long)mhi << 32);
            const int nbd = 32 - __clz(abs(diff));
            const uint32_t symd = s_dc[tb][nbd];
            const uint32_t vd = ((symd >> 5) << nbd) | ((uint32_t)(diff < 0 ? diff - 1 : diff) & ((1u << nbd) - 1u));
            const int nd = (int)(symd & 31u) + nbd;                   // <= 20 bits
            const uint32_t eob = s_ac[tb][0], zrl = s_ac[tb][0xF0];
            if (mask == 0ull) {                                       // no AC term at all (flat areas): DC difference + EOB, one store
                if (lane == 0) put_item(buf, bitpos, (vd << (eob & 31u)) | (eob >> 5), nd + (int)(eob & 31u));
                bitpos += nd + (int)(eob & 31u);
                continue;
            }
            uint32_t v0 = 0, v1 = 0;
            int n0 = 0, n1 = 0, z0 = 0, z1 = 0;
            if (lane == 0) { v0 = vd; n0 = nd; }
            else if (c0 != 0) ac_item(c0, lane, mask, s_ac[tb], v0, n0, z0);
            if (c1 != 0) ac_item(c1, lane + 32, mask, s_ac[tb], v1, n1, z1);
            else if (lane == 31) { v1 = eob >> 5; n1 = (int)(eob & 31u); }   // position 63 is zero: EOB
            const int l0 = n0 + z0 * (int)(zrl & 31u), l1 = n1 + z1 * (int)(zrl & 31u);
            uint32_t incl = (uint32_t)l0 | ((uint32_t)l1 << 16);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
            const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
            if (n0) put_ac(buf, bitpos + (int)(incl & 0xFFFFu) - l0, v0, n0, z0, zrl);
            if (n1) put_ac(buf, bitpos + (int)(tot & 0xFFFFu) + (int)(incl >> 16) - l1, v1, n1, z1, zrl);
            bitpos += (int)(tot & 0xFFFFu) + (int)(tot >> 16);
        }
        if (bitpos + kMcuBits > (kBitWords - 3) * 32) flush_bytes(buf, bitpos, out, outpos, lane);
    }
    flush_bytes(buf, bitpos, out, outpos, lane);
    if (bitpos & 7) {                                                 // jchuff.c flush_bits: pad the last byte with 1 bits
        const int pad = 8 - (bitpos & 7);
        if (lane == 0) put_item(buf, bitpos, (1u << pad) - 1u, pad);
        bitpos += pad;
        flush_bytes(buf, bitpos, out, outpos, lane);
    }
    if (lane == 0) lens[it] = outpos;
}

// exclusive scan of (len + 2) over the intervals; offs[n] = total stream length including the header
__global__ void __launch_bounds__(1024) jpeg_scan_kernel(const uint32_t *__restrict__ lens, int n, uint32_t header_len, uint32_t *__restrict__ offs) {
    __shared__ uint32_t warp_sum[32];
    const int chunk = (n + 1023) / 1024, lo = min(n, (int)threadIdx.x * chunk), hi = min(n, lo + chunk);
    uint32_t mine = 0;
    for (int i = lo; i < hi; i += 8) {                                // independent loads, eight at a time
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = i + k < hi ? lens[i + k] + 2u : 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) mine += v[k];
    }
    uint32_t incl = mine;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) warp_sum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = warp_sum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += u; }
        warp_sum[lane] = s;
    }
    __syncthreads();
    uint32_t base = header_len + (wid ? warp_sum[wid - 1] : 0u) + incl - mine;
    for (int i = lo; i < hi; i += 8) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = i + k < hi ? lens[i + k] + 2u : 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) { if (i + k < hi) offs[i + k] = base; base += v[k]; }
    }
    if (threadIdx.x == 1023) offs[n] = base;
}

// one warp per interval: slot -> its place in the stream, followed by RSTn (EOI after the last one)
__global__ void __launch_bounds__(256) jpeg_gather_kernel(const uint8_t *__restrict__ slots, int slot_bytes, const uint32_t *__restrict__ lens,
                                                          const uint32_t *__restrict__ offs, int n_int, const __grid_constant__ JpegHeader hd,
                                                          uint8_t *__restrict__ out, unsigned long long capacity, uint32_t *__restrict__ size_out) {
    const int lane = threadIdx.x & 31, it = blockIdx.x * 8 + (threadIdx.x >> 5);
    const uint32_t total = offs[n_int];
    const bool fits = (unsigned long long)total <= capacity;
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) *size_out = fits ? total : 0u;          // 0: the stream did not fit the caller's buffer
        if (fits) for (int i = threadIdx.x; i < hd.len; i += 256) out[i] = hd.bytes[i];
    }
    if (it >= n_int || !fits) return;
    const uint8_t *src = slots + (size_t)it * slot_bytes;
    const uint32_t len = lens[it];
    uint8_t *dst = out + offs[it];
    uint32_t i = 0;
    const uint32_t lead = min(len, (uint32_t)((4u - ((uintptr_t)dst & 3u)) & 3u));   // align the destination, then move words
    if (lane < (int)lead) dst[lane] = src[lane];
    i = lead;
    const uint32_t words = (len - i) / 4u;
    for (uint32_t k = lane; k < words; k += 32) {
        const uint8_t *s = src + i + 4u * k;
        *(uint32_t *)(dst + i + 4u * k) = (uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16) | ((uint32_t)s[3] << 24);
    }
    i += 4u * words;
    if (i + lane < len) dst[i + lane] = src[i + lane];
    if (lane == 0) { dst[len] = 0xFF; dst[len + 1] = (uint8_t)(it == n_int - 1 ? 0xD9 : 0xD0 + (it & 7)); }
}

struct JpegGeometry { int mcus_x, mcus_y, n_mcus, n_int, slot_bytes; size_t coef_bytes, slots_bytes, lens_off, offs_off, total; };

static JpegGeometry jpeg_geometry(int h, int w, int ri) {
    JpegGeometry g;
    g.mcus_x = ceil_div(w, 16); g.mcus_y = ceil_div(h, 16); g.n_mcus = g.mcus_x * g.mcus_y;
    g.n_int = ceil_div(g.n_mcus, ri);
    g.slot_bytes = ri * 6 * kBlockWorstBytes + 16;
    g.coef_bytes = ((size_t)g.n_mcus * 6 * 64 * sizeof(int16_t) + 255) & ~(size_t)255;
    g.slots_bytes = ((size_t)g.n_int * g.slot_bytes + 255) & ~(size_t)255;
    g.lens_off = g.coef_bytes + g.slots_bytes;
    g.offs_off = g.lens_off + (((size_t)g.n_int * 4 + 255) & ~(size_t)255);
    g.total = g.offs_off + (((size_t)(g.n_int + 1) * 4 + 255) & ~(size_t)255);
    return g;
}

static int jpeg_check(int h, int w, int ri) {
    D2S_REQUIRE(h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0 && h <= 65534 && w <= 65534, "d2s_jpeg: frame %dx%d must be even-sized and below 65535", h, w);
    D2S_REQUIRE(ri >= 1 && ri <= 65535, "d2s_jpeg: restart interval %d outside [1, 65535] MCUs (the device encoder is parallel over restart intervals)", ri);
    D2S_REQUIRE((long long)ceil_div(ceil_div(w, 16) * (long long)ceil_div(h, 16), ri) <= 1024LL * 256, "d2s_jpeg: too many restart intervals for %dx%d at %d MCUs each", h, w, ri);
    return D2S_OK;
}

int jpeg_encode_launch(const uint8_t *rgb, long long pitch, int h, int w, int quality, int ri, uint8_t *out, size_t capacity, uint32_t *size_out,
                       void *workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (int rc = jpeg_check(h, w, ri)) return rc;
    D2S_REQUIRE(rgb && out && size_out && workspace, "d2s_jpeg_encode: null pointer");
    D2S_REQUIRE((((uintptr_t)workspace) & 15) == 0, "d2s_jpeg_encode: workspace must be 16-byte aligned");
    const JpegGeometry g = jpeg_geometry(h, w, ri);
    D2S_REQUIRE(workspace_bytes >= g.total, "d2s_jpeg_encode: workspace %zu < %zu bytes", workspace_bytes, g.total);
    D2S_REQUIRE(capacity >= 1024, "d2s_jpeg_encode: output capacity %zu is below the header", capacity);
    JpegQuant jq; JpegHuff jh; JpegHeader hd;
    build_tables(h, w, quality, ri, jq, jh, hd);
    uint8_t *ws = (uint8_t *)workspace;
    int16_t *coef = (int16_t *)ws;
    uint8_t *slots = ws + g.coef_bytes;
    uint32_t *lens = (uint32_t *)(ws + g.lens_off), *offs = (uint32_t *)(ws + g.offs_off);
    D2S_REQUIRE(g.mcus_y <= 65535, "d2s_jpeg_encode: frame height %d", h);
    D2S_LAUNCH(jpeg_transform_kernel, dim3(ceil_div(g.mcus_x, JG), g.mcus_y), JT, 0, stream, rgb, pitch, h, w, g.mcus_x, jq, coef);
    D2S_LAUNCH(jpeg_entropy_kernel, ceil_div(g.n_int, EW), EW * 32, 0, stream, (const int16_t *)coef, g.n_mcus, g.mcus_x, ceil_div(w, 8), ceil_div(h, 8), ri,
               g.n_int, jh, slots, g.slot_bytes, lens);
    D2S_LAUNCH(jpeg_scan_kernel, 1, 1024, 0, stream, (const uint32_t *)lens, g.n_int, (uint32_t)hd.len, offs);
    D2S_LAUNCH(jpeg_gather_kernel, ceil_div(g.n_int, 8), 256, 0, stream, (const uint8_t *)slots, g.slot_bytes, (const uint32_t *)lens, (const uint32_t *)offs,
               g.n_int, hd, out, (unsigned long long)capacity, size_out);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

size_t jpeg_workspace(int h, int w, int ri) { return jpeg_geometry(h, w, ri).total; }
size_t jpeg_max_bytes(int h, int w, int ri) {
    const JpegGeometry g = jpeg_geometry(h, w, ri);
    return 1024 + (size_t)g.n_int * (g.slot_bytes + 2);
}

}  // namespace d2s

extern "C" size_t d2s_jpeg_workspace_bytes(int h, int w, int restart_interval) {
    if (d2s::jpeg_check(h, w, restart_interval)) return 0;
    return d2s::jpeg_workspace(h, w, restart_interval);
}

extern "C" size_t d2s_jpeg_max_bytes(int h, int w, int restart_interval) {
    if (d2s::jpeg_check(h, w, restart_interval)) return 0;
    return d2s::jpeg_max_bytes(h, w, restart_interval);
}

extern "C" int d2s_jpeg_encode(const uint8_t *rgb_hwc, int64_t row_pitch_bytes, int h, int w, int quality, int restart_interval, uint8_t *jpeg,
                               size_t capacity, uint32_t *size_out, void *workspace, size_t workspace_bytes, d2s_stream_t stream) {
    return d2s::jpeg_encode_launch(rgb_hwc, row_pitch_bytes > 0 ? row_pitch_bytes : (long long)w * 3, h, w, quality, restart_interval, jpeg, capacity,
                                   size_out, workspace, workspace_bytes, (cudaStream_t)stream);
}
