/* ORACLE (test infrastructure only; never linked or called by the product) — baseline JPEG encoder as libjpeg-turbo runs it for
 * cv2.imencode(".jpg", bgr, [IMWRITE_JPEG_QUALITY, q]), the call the reference's MJPEGStreamer makes on every frame
 * (reference streamer.py:250-256).  libjpeg-turbo is OpenCV's bundled third-party dependency (3.1.2 in this image; not vendored in
 * /root/reference); its published algorithm is restated here stage by stage, sequentially, one MCU after another:
 *     jccolor.c   rgb_ycc_convert      JFIF BT.601 full range, 16-bit fixed point
 *     jcsample.c  h2v2_downsample      2x2 mean with the alternating 1,2 bias; right edge replicated at full resolution first
 *     jcprepct.c  pre_process_data     bottom edge: the DOWNSAMPLED rows are replicated to a whole iMCU
 *     jccoefct.c  compress_data        blocks past a component's own block grid are dummies: AC = 0, DC = the previous block's DC
 *     jfdctint.c  jpeg_fdct_islow      the default integer DCT (CONST_BITS 13, PASS1_BITS 2), output scaled by 8
 *     jcdctmgr.c  quantize             round-half-away division by 8 * quantval
 *     jcparam.c   jpeg_set_quality     quality -> scale (5000/q or 200-2q), (base * scale + 50) / 100 clamped to [1, 255]
 *     jchuff.c    encode_one_block     Annex K tables, DC differences, (run, size) AC symbols, ZRL / EOB, 0xFF stuffing, restart markers
 *     jcmarker.c  headers              SOI APP0(JFIF 1.01) DQT DQT SOF0 DHT x4 [DRI] SOS ... EOI
 * Pinned byte for byte on cv2.imencode itself (tests/test_oracle_jpeg.py), which runs in this container and on the GPU box. */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

static const uint8_t kBaseQ[2][64] = {   /* Annex K.1 / K.2 in zig-zag order (what a quality-50 DQT segment carries) */
    {16, 11, 12, 14, 12, 10, 16, 14, 13, 14, 18, 17, 16, 19, 24, 40, 26, 24, 22, 22, 24, 49, 35, 37, 29, 40, 58, 51, 61, 60, 57, 51,
     56, 55, 64, 72, 92, 78, 64, 68, 87, 69, 55, 56, 80, 109, 81, 87, 95, 98, 103, 104, 103, 62, 77, 113, 121, 112, 100, 120, 92, 101, 103, 99},
    {17, 18, 18, 24, 21, 24, 47, 26, 26, 47, 99, 66, 56, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99}};

static const uint8_t kBits[4][16] = {    /* DC lum, AC lum, DC chrom, AC chrom (Annex K.3) */
    {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0},
    {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 125},
    {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0},
    {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 119}};
static const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t kAcLum[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1,
    0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56,
    0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85,
    0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa,
    0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
    0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
static const uint8_t kAcChr[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42,
    0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19,
    0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55,
    0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8,
    0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4,
    0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

typedef struct { uint16_t code[256]; uint8_t size[256]; } huff_t;

static void derive(const uint8_t *bits, const uint8_t *vals, huff_t *t) {   /* jchuff.c jpeg_make_c_derived_tbl */
    memset(t, 0, sizeof(*t));
    int code = 0, k = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int i = 0; i < bits[len - 1]; ++i, ++k) { t->code[vals[k]] = (uint16_t)code++; t->size[vals[k]] = (uint8_t)len; }
        code <<= 1;
    }
}

static void zigzag_order(int *zz) {      /* zz[k] = natural (row-major) index of the k-th zig-zag coefficient */
    int r = 0, c = 0;
    for (int k = 0; k < 64; ++k) {
        zz[k] = r * 8 + c;
        if ((r + c) % 2 == 0) { if (c == 7) ++r; else if (r == 0) ++c; else { --r; ++c; } }
        else                  { if (r == 7) ++c; else if (c == 0) ++r; else { ++r; --c; } }
    }
}

typedef struct { uint8_t *p, *end; uint64_t acc; int nbits; int overflow; } bitw_t;

static void put_byte(bitw_t *b, int v) { if (b->p < b->end) *b->p++ = (uint8_t)v; else b->overflow = 1; }
static void put_bits(bitw_t *b, unsigned code, int size) {
    b->acc = (b->acc << size) | (code & ((1u << size) - 1)); b->nbits += size;
    while (b->nbits >= 8) {
        const int v = (int)((b->acc >> (b->nbits - 8)) & 0xFF);
        put_byte(b, v); if (v == 0xFF) put_byte(b, 0);
        b->nbits -= 8;
    }
}
static void flush_bits(bitw_t *b) { if (b->nbits) put_bits(b, 0x7F, 8 - b->nbits); b->acc = 0; b->nbits = 0; }   /* pad with 1 bits */

static int bit_length(int v) { int n = 0; while (v) { ++n; v >>= 1; } return n; }

static void encode_block(bitw_t *b, const int16_t *blk /* zig-zag order */, int *last_dc, const huff_t *dc, const huff_t *ac) {
    int temp = blk[0] - *last_dc, temp2 = temp;
    *last_dc = blk[0];
    if (temp < 0) { temp = -temp; --temp2; }
    int nb = bit_length(temp);
    put_bits(b, dc->code[nb], dc->size[nb]);
    if (nb) put_bits(b, (unsigned)temp2, nb);
    int run = 0;
    for (int k = 1; k < 64; ++k) {
        temp = blk[k];
        if (temp == 0) { ++run; continue; }
        while (run > 15) { put_bits(b, ac->code[0xF0], ac->size[0xF0]); run -= 16; }
        temp2 = temp;
        if (temp < 0) { temp = -temp; --temp2; }
        nb = bit_length(temp);
        put_bits(b, ac->code[(run << 4) + nb], ac->size[(run << 4) + nb]);
        put_bits(b, (unsigned)temp2, nb);
        run = 0;
    }
    if (run > 0) put_bits(b, ac->code[0], ac->size[0]);
}

#define DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))

static void fdct_islow(int *d) {         /* jfdctint.c: rows then columns, in place on 64 ints */
    for (int pass = 0; pass < 2; ++pass) {
        const int stride = pass ? 8 : 1, step = pass ? 1 : 8;
        for (int i = 0; i < 8; ++i) {
            int *p = d + i * step;
            const int t0 = p[0] + p[7 * stride], t7 = p[0] - p[7 * stride], t1 = p[stride] + p[6 * stride], t6 = p[stride] - p[6 * stride];
            const int t2 = p[2 * stride] + p[5 * stride], t5 = p[2 * stride] - p[5 * stride], t3 = p[3 * stride] + p[4 * stride], t4 = p[3 * stride] - p[4 * stride];
            const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
            const int sh = pass ? 13 + 2 : 13 - 2;
            if (!pass) { p[0] = (t10 + t11) << 2; p[4 * stride] = (t10 - t11) << 2; }
            else       { p[0] = DESCALE(t10 + t11, 2); p[4 * stride] = DESCALE(t10 - t11, 2); }
            int z1 = (t12 + t13) * 4433;
            p[2 * stride] = DESCALE(z1 + t13 * 6270, sh);
            p[6 * stride] = DESCALE(z1 + t12 * -15137, sh);
            z1 = t4 + t7; int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
            const int z5 = (z3 + z4) * 9633;
            const int a4 = t4 * 2446, a5 = t5 * 16819, a6 = t6 * 25172, a7 = t7 * 12299;
            z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
            z3 += z5; z4 += z5;
            p[7 * stride] = DESCALE(a4 + z1 + z3, sh);
            p[5 * stride] = DESCALE(a5 + z2 + z4, sh);
            p[3 * stride] = DESCALE(a6 + z2 + z3, sh);
            p[stride] = DESCALE(a7 + z1 + z4, sh);
        }
    }
}

static int y_of(int r, int g, int b) { return (19595 * r + 38470 * g + 7471 * b + 32768) >> 16; }
static int cb_of(int r, int g, int b) { return (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16; }
static int cr_of(int r, int g, int b) { return (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16; }

static int imin(int a, int b) { return a < b ? a : b; }

/* one component sample with libjpeg's edge rules.  comp 0: luma at (y, x); comp 1/2: chroma at (cy, cx) */
static int sample(const uint8_t *rgb, int h, int w, int comp, int y, int x) {
    if (comp == 0) {
        const uint8_t *p = rgb + ((size_t)imin(y, h - 1) * w + imin(x, w - 1)) * 3;
        return y_of(p[0], p[1], p[2]);
    }
    const int cy = imin(y, h / 2 - 1);                       /* downsampled rows are replicated */
    int sum = 0;
    for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx) {                     /* full-resolution columns are replicated before the mean */
            const uint8_t *p = rgb + ((size_t)(2 * cy + dy) * w + imin(2 * x + dx, w - 1)) * 3;
            sum += comp == 1 ? cb_of(p[0], p[1], p[2]) : cr_of(p[0], p[1], p[2]);
        }
    return (sum + ((x & 1) ? 2 : 1)) >> 2;
}

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

static void put16(uint8_t **p, int v) { *(*p)++ = (uint8_t)(v >> 8); *(*p)++ = (uint8_t)v; }

size_t d2s_oracle_jpeg_header(int h, int w, int quality, int restart_interval, uint8_t *out, uint8_t qtab[2][64]) {
    int q = quality < 1 ? 1 : quality > 100 ? 100 : quality;
    const int scale = q < 50 ? 5000 / q : 200 - 2 * q;
    for (int t = 0; t < 2; ++t)
        for (int i = 0; i < 64; ++i) {
            long v = ((long)kBaseQ[t][i] * scale + 50) / 100;
            qtab[t][i] = (uint8_t)(v < 1 ? 1 : v > 255 ? 255 : v);
        }
    uint8_t *p = out;
    static const uint8_t app0[] = {0xFF, 0xD8, 0xFF, 0xE0, 0x00, 0x10, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
    memcpy(p, app0, sizeof(app0)); p += sizeof(app0);
    for (int t = 0; t < 2; ++t) { *p++ = 0xFF; *p++ = 0xDB; put16(&p, 67); *p++ = (uint8_t)t; memcpy(p, qtab[t], 64); p += 64; }
    *p++ = 0xFF; *p++ = 0xC0; put16(&p, 17); *p++ = 8; put16(&p, h); put16(&p, w); *p++ = 3;
    *p++ = 1; *p++ = 0x22; *p++ = 0; *p++ = 2; *p++ = 0x11; *p++ = 1; *p++ = 3; *p++ = 0x11; *p++ = 1;
    const uint8_t *vals[4] = {kDcVals, kAcLum, kDcVals, kAcChr};
    const int nvals[4] = {12, 162, 12, 162}, ids[4] = {0x00, 0x10, 0x01, 0x11};
    for (int t = 0; t < 4; ++t) {
        *p++ = 0xFF; *p++ = 0xC4; put16(&p, 2 + 1 + 16 + nvals[t]); *p++ = (uint8_t)ids[t];
        memcpy(p, kBits[t], 16); p += 16; memcpy(p, vals[t], nvals[t]); p += nvals[t];
    }
    if (restart_interval > 0) { *p++ = 0xFF; *p++ = 0xDD; put16(&p, 4); put16(&p, restart_interval); }
    static const uint8_t sos[] = {0xFF, 0xDA, 0x00, 0x0C, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 0x3F, 0};
    memcpy(p, sos, sizeof(sos)); p += sizeof(sos);
    return (size_t)(p - out);
}

/* rgb: [h, w, 3] u8 (h, w even).  Returns the stream length, or 0 if `cap` is too small. */
size_t d2s_oracle_jpeg_encode(const uint8_t *rgb, int h, int w, int quality, int restart_interval, uint8_t *out, size_t cap) {
    if (cap < 1024 || h < 2 || w < 2 || (h & 1) || (w & 1)) return 0;
    uint8_t qtab[2][64];
    const size_t hdr = d2s_oracle_jpeg_header(h, w, quality, restart_interval, out, qtab);
    huff_t huff[4];
    derive(kBits[0], kDcVals, &huff[0]); derive(kBits[1], kAcLum, &huff[1]); derive(kBits[2], kDcVals, &huff[2]); derive(kBits[3], kAcChr, &huff[3]);
    int zz[64]; zigzag_order(zz);
    const int mcus_x = ceil_div(w, 16), mcus_y = ceil_div(h, 16);
    const int yblk_w = ceil_div(w, 8), yblk_h = ceil_div(h, 8);             /* luma block grid; chroma's is always whole MCUs */
    bitw_t bw = {out + hdr, out + cap - 2, 0, 0, 0};
    int last_dc[3] = {0, 0, 0}, to_go = restart_interval, rst = 0;
    for (int my = 0; my < mcus_y; ++my)
        for (int mx = 0; mx < mcus_x; ++mx) {
            if (restart_interval > 0 && to_go == 0) {
                flush_bits(&bw); put_byte(&bw, 0xFF); put_byte(&bw, 0xD0 + rst);
                rst = (rst + 1) & 7; last_dc[0] = last_dc[1] = last_dc[2] = 0; to_go = restart_interval;
            }
            int16_t prev_dc_coef = 0;                                    /* DC of the previous block in MCU order, for dummies */
            for (int blk = 0; blk < 6; ++blk) {
                const int comp = blk < 4 ? 0 : blk - 3;
                const int by = comp == 0 ? 2 * my + (blk >> 1) : my, bx = comp == 0 ? 2 * mx + (blk & 1) : mx;
                int16_t coef[64];
                if (comp == 0 && (by >= yblk_h || bx >= yblk_w)) {        /* jccoefct.c dummy block */
                    memset(coef, 0, sizeof(coef)); coef[0] = prev_dc_coef;
                } else {
                    int d[64];
                    for (int r = 0; r < 8; ++r)
                        for (int c = 0; c < 8; ++c) d[r * 8 + c] = sample(rgb, h, w, comp, by * 8 + r, bx * 8 + c) - 128;
                    fdct_islow(d);
                    const uint8_t *qt = qtab[comp ? 1 : 0];
                    for (int k = 0; k < 64; ++k) {
                        const int qv = qt[k] << 3;
                        int t = d[zz[k]];
                        if (t < 0) { t = -t; t += qv >> 1; t = t >= qv ? t / qv : 0; t = -t; }
                        else       { t += qv >> 1; t = t >= qv ? t / qv : 0; }
                        coef[k] = (int16_t)t;
                    }
                }
                prev_dc_coef = coef[0];
                encode_block(&bw, coef, &last_dc[comp], &huff[comp ? 2 : 0], &huff[comp ? 3 : 1]);
            }
            if (restart_interval > 0) --to_go;
        }
    flush_bits(&bw);
    put_byte(&bw, 0xFF); put_byte(&bw, 0xD9);
    return bw.overflow ? 0 : (size_t)(bw.p - out);
}
