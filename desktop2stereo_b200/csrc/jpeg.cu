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
//   entropy     one thread per restart interval: jchuff.c's encode_one_block over the interval's MCUs into a private slot sized
//               for the worst case (so it cannot overflow), 0xFF stuffing and the 1-bit padding included; writes the slot's length.
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

struct JpegQuant { uint16_t div[2][64]; uint8_t izz[64]; };     // 8 * quantval in natural order; natural index -> zig-zag position
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
        for (int n = 0; n < 64; ++n) jq.div[t][n] = (uint16_t)(qz[t][jq.izz[n]] << 3);
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
        const int nat = i * 8 + r, qv = q.div[tb][nat];
        const int a = abs(d[i]) + (qv >> 1);
        const int v = a >= qv ? a / qv : 0;
        sO[blk][q.izz[nat]] = (int16_t)(d[i] < 0 ? -v : v);
    }
    __syncthreads();
    const int valid = min(JG, mcus_x - mx0) * 6 * 8;                  // 16-byte pieces of this CTA's MCUs
    if ((int)threadIdx.x < valid)
        ((uint4 *)(coef + ((size_t)my * mcus_x + mx0) * 6 * 64))[threadIdx.x] = ((const uint4 *)&sO[0][0])[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------------------------- entropy coding
// Bytes leave through a 32-bit staging word (slots are 4-byte aligned), so a thread issues one store per four stream bytes.
struct BitWriter {
    uint32_t *wp;          // next word of the slot
    uint32_t stage;        // bytes of the current word, little-endian
    uint32_t pos;          // bytes written so far
    uint64_t acc;
    int nbits;
    __device__ __forceinline__ void byte(uint32_t v) {
        stage |= v << ((pos & 3u) * 8u);
        if ((++pos & 3u) == 0u) { *wp++ = stage; stage = 0u; }
    }
    __device__ __forceinline__ void put(uint32_t bits, int size) {    // `bits` already masked to `size` bits
        acc = (acc << size) | bits;
        nbits += size;
        while (nbits >= 8) {
            const uint32_t v = (uint32_t)(acc >> (nbits - 8)) & 0xFFu;
            byte(v);
            if (v == 0xFFu) byte(0u);
            nbits -= 8;
        }
    }
    __device__ __forceinline__ uint32_t finish() {                    // pad with 1 bits (jchuff.c flush_bits), write the last partial word
        if (nbits) put((1u << (8 - nbits)) - 1u, 8 - nbits);
        if (pos & 3u) *wp = stage;
        return pos;
    }
};

__device__ __forceinline__ void put_value(BitWriter &bw, uint32_t sym, int t, int nb) {               // Huffman code, then the value bits
    const int t2 = t < 0 ? t - 1 : t;
    bw.put(((sym >> 5) << nb) | ((uint32_t)t2 & ((1u << nb) - 1u)), (int)(sym & 31u) + nb);
}

// four zig-zag coefficients packed in 64 bits: jump from non-zero to non-zero (jchuff.c encode_one_block's run/size loop)
__device__ __forceinline__ void code_four(BitWriter &bw, const uint32_t *s_ac, unsigned long long x, int &run) {
    int left = 4;
    while (x) {
        const int hz = (__ffsll((long long)x) - 1) >> 4;
        run += hz;
        const int t = (int)(int16_t)((x >> (16 * hz)) & 0xFFFFull);
        x = hz == 3 ? 0ull : x >> (16 * (hz + 1));
        left -= hz + 1;
        while (run > 15) { const uint32_t z = s_ac[0xF0]; bw.put(z >> 5, (int)(z & 31u)); run -= 16; }
        const int nb = 32 - __clz(abs(t));
        put_value(bw, s_ac[(run << 4) + nb], t, nb);
        run = 0;
    }
    run += left;
}

__device__ __forceinline__ void load_block(uint4 (&v)[8], const int16_t *coef, size_t block) {
    const uint4 *cp = (const uint4 *)(coef + block * 64);
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = __ldg(cp + c);
}

// One thread = one restart interval.  The thread's chain of dependent loads is what bounds it, so the 128 bytes of block b+1 are
// requested before block b is coded.
__global__ void __launch_bounds__(128) jpeg_entropy_kernel(const int16_t *__restrict__ coef, int n_mcus, int mcus_x, int yblk_w, int yblk_h,
                                                           int ri, int n_int, const __grid_constant__ JpegHuff hf,
                                                           uint8_t *__restrict__ slots, int slot_bytes, uint32_t *__restrict__ lens) {
    __shared__ uint32_t s_ac[2][256], s_dc[2][16];
    for (int i = threadIdx.x; i < 512; i += 128) s_ac[i >> 8][i & 255] = hf.ac[i >> 8][i & 255];
    if (threadIdx.x < 32) s_dc[threadIdx.x >> 4][threadIdx.x & 15] = hf.dc[threadIdx.x >> 4][threadIdx.x & 15];
    __syncthreads();
    const int it = blockIdx.x * 128 + threadIdx.x;
    if (it >= n_int) return;
    BitWriter bw{(uint32_t *)(slots + (size_t)it * slot_bytes), 0u, 0u, 0ull, 0};
    int last_y = 0, last_cb = 0, last_cr = 0;
    const int m0 = it * ri, m_end = min(n_mcus, m0 + ri);
    const size_t blk_end = (size_t)m_end * 6;
    uint4 cur[8], nxt[8];
    load_block(nxt, coef, (size_t)m0 * 6);
    for (int m = m0; m < m_end; ++m) {
        const int my = m / mcus_x, mx = m - my * mcus_x;
        int prev_dc = 0;
#pragma unroll 1
        for (int b = 0; b < 6; ++b) {
#pragma unroll
            for (int c = 0; c < 8; ++c) cur[c] = nxt[c];
            const size_t nb_idx = (size_t)m * 6 + b + 1;
            if (nb_idx < blk_end) load_block(nxt, coef, nb_idx);
            const int tb = b < 4 ? 0 : 1;
            const bool dummy = b < 4 && (2 * my + (b >> 1) >= yblk_h || 2 * mx + (b & 1) >= yblk_w);   // jccoefct.c
            const int dc = dummy ? prev_dc : (int)(int16_t)(cur[0].x & 0xFFFFu);
            prev_dc = dc;
            int diff;
            if (b < 4) { diff = dc - last_y; last_y = dc; } else if (b == 4) { diff = dc - last_cb; last_cb = dc; } else { diff = dc - last_cr; last_cr = dc; }
            const int nbd = 32 - __clz(abs(diff));
            const uint32_t symd = s_dc[tb][nbd];
            if (nbd) put_value(bw, symd, diff, nbd); else bw.put(symd >> 5, (int)(symd & 31u));
            int run = -1;                                             // the DC slot is masked to zero below and must not count
            if (!dummy) {
                cur[0].x &= 0xFFFF0000u;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    code_four(bw, s_ac[tb], (unsigned long long)cur[c].x | ((unsigned long long)cur[c].y << 32), run);
                    code_four(bw, s_ac[tb], (unsigned long long)cur[c].z | ((unsigned long long)cur[c].w << 32), run);
                }
            } else {
                run = 63;
            }
            if (run > 0) { const uint32_t e = s_ac[tb][0]; bw.put(e >> 5, (int)(e & 31u)); }
        }
    }
    lens[it] = bw.finish();
}

// exclusive scan of (len + 2) over the intervals; offs[n] = total stream length including the header
__global__ void __launch_bounds__(1024) jpeg_scan_kernel(const uint32_t *__restrict__ lens, int n, uint32_t header_len, uint32_t *__restrict__ offs) {
    __shared__ uint32_t warp_sum[32];
    const int chunk = (n + 1023) / 1024, lo = min(n, (int)threadIdx.x * chunk), hi = min(n, lo + chunk);
    uint32_t mine = 0;
    for (int i = lo; i < hi; ++i) mine += lens[i] + 2u;
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
    for (int i = lo; i < hi; ++i) { offs[i] = base; base += lens[i] + 2u; }
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
    D2S_LAUNCH(jpeg_entropy_kernel, ceil_div(g.n_int, 128), 128, 0, stream, (const int16_t *)coef, g.n_mcus, g.mcus_x, ceil_div(w, 8), ceil_div(h, 8), ri,
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
