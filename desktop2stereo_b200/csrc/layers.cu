// Non-GEMM layers of the depth network: LayerNorm, patch im2col, token assembly, position-table resampling,
// pixel shuffle (ConvTranspose as GEMM), stride-2 im2col, NHWC bilinear upsampling, weight packing.
#include "layers.cuh"

namespace d2s {

// ---- LayerNorm: one warp per row, two-pass statistics in registers (HF dinov2:348-386 norm1/norm2, :605-618) ----
__global__ void __launch_bounds__(256) layernorm_kernel(const float *__restrict__ x, const float *__restrict__ gamma,
                                                        const float *__restrict__ beta, __half *__restrict__ y, int rows, int D,
                                                        float eps, int skip_cls, int tokens_per_img) {
    pdl_sync();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    long long in_row = row;
    if (skip_cls) { int P = tokens_per_img - 1; in_row = (long long)(row / P) * tokens_per_img + 1 + row % P; }
    const float4 *xr = (const float4 *)(x + in_row * D);
    const int nv = (D + 127) >> 7;  // float4 slots per lane; slot i of a lane covers channels (lane + 32 i) * 4 .. + 3 when < D
    float4 v[8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (i < nv && (lane + i * 32) * 4 < D) { v[i] = xr[lane + i * 32]; sum += v[i].x + v[i].y + v[i].z + v[i].w; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)D;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (i < nv && (lane + i * 32) * 4 < D) {
            float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            sq += a * a + b * b + c * c + d * d;
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)D + eps);
    __half *yr = y + (long long)row * D;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (i < nv && (lane + i * 32) * 4 < D) {
            int c = (lane + i * 32) * 4;
            float4 g = __ldg((const float4 *)(gamma + c)), b = __ldg((const float4 *)(beta + c));
            __half2 h0 = __floats2half2_rn((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
            __half2 h1 = __floats2half2_rn((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
            uint2 u = make_uint2(*(uint32_t *)&h0, *(uint32_t *)&h1);
            *(uint2 *)(yr + c) = u;
        }
}

int layernorm_launch(const float *x, const float *gamma, const float *beta, __half *y, int rows, int D, float eps,
                     int skip_cls, int tokens_per_img, cudaStream_t stream) {
    D2S_REQUIRE(D % 4 == 0 && D <= 1024, "layernorm: D=%d must be a multiple of 4 and <= 1024", D);
    D2S_LAUNCH(layernorm_kernel, ceil_div(rows, 8), 256, 0, stream, x, gamma, beta, y, rows, D, eps, skip_cls, tokens_per_img);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

// ---- patch embedding as GEMM: im2col of non-overlapping 14x14 patches (HF dinov2:139-149) ----
template <typename T>
__global__ void patch_im2col_kernel(const T *__restrict__ pix, __half *__restrict__ out, int B, int H, int W, int patch, int Kp) {
    pdl_sync();
    const int ph = H / patch, pw = W / patch, K = 3 * patch * patch;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * ph * pw * K;
    if (i >= total) return;
    int k = (int)(i % K);
    long long m = i / K;
    int px = (int)(m % pw), py = (int)((m / pw) % ph), b = (int)(m / ((long long)pw * ph));
    int c = k / (patch * patch), r = k % (patch * patch), ky = r / patch, kx = r % patch;
    float v = to_f32<T>(pix[(((long long)b * 3 + c) * H + py * patch + ky) * W + px * patch + kx]);
    out[m * Kp + k] = __float2half_rn(v);
}

int patch_im2col_launch(const void *pix, int in_dtype, __half *out, int B, int H, int W, int patch, int Kp, cudaStream_t stream) {
    long long total = (long long)B * (H / patch) * (W / patch) * 3 * patch * patch;
    int grid = ceil_div(total, 256);
    if (in_dtype == D2S_F32) D2S_LAUNCH(patch_im2col_kernel<float>, grid, 256, 0, stream, (const float *)pix, out, B, H, W, patch, Kp);
    else if (in_dtype == D2S_F16) D2S_LAUNCH(patch_im2col_kernel<__half>, grid, 256, 0, stream, (const __half *)pix, out, B, H, W, patch, Kp);
    else return set_error(D2S_ERR_UNSUPPORTED, "d2s_infer: pixel_values dtype %d unsupported", in_dtype);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

// ---- cls token + position embeddings (HF dinov2:97-116) ----
__global__ void assemble_tokens_kernel(const __half *__restrict__ patches, const float *__restrict__ cls, const float *__restrict__ pos,
                                       float *__restrict__ x, int B, int P, int D) {
    pdl_sync();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * (P + 1) * D;
    if (i >= total) return;
    int d = (int)(i % D);
    long long t = i / D;
    int tok = (int)(t % (P + 1)), b = (int)(t / (P + 1));
    float v = tok == 0 ? cls[d] : __half2float(patches[((long long)b * P + tok - 1) * D + d]);
    x[i] = v + pos[(long long)tok * D + d];
}

int assemble_tokens_launch(const __half *patches, const float *cls, const float *pos, float *x, int B, int P, int D, cudaStream_t stream) {
    long long total = (long long)B * (P + 1) * D;
    D2S_LAUNCH(assemble_tokens_kernel, ceil_div(total, 256), 256, 0, stream, patches, cls, pos, x, B, P, D);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

// ---- bicubic resampling of the position table (ATen upsample_bicubic2d, A = -0.75, align_corners=False) ----
__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

__global__ void pos_embed_interp_kernel(const float *__restrict__ table, float *__restrict__ out, int g, int ph, int pw, int D, float sy, float sx) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)(1 + ph * pw) * D;
    if (i >= total) return;
    int d = (int)(i % D);
    int tok = (int)(i / D);
    if (tok == 0) { out[i] = table[d]; return; }   // class position embedding is kept as is
    int py = (tok - 1) / pw, px = (tok - 1) % pw;
    const float A = -0.75f;
    float ry = sy * ((float)py + 0.5f) - 0.5f, rx = sx * ((float)px + 0.5f) - 0.5f;
    int iy = (int)floorf(ry), ix = (int)floorf(rx);
    float ty = ry - (float)iy, tx = rx - (float)ix;
    float cy[4] = {cubic2(ty + 1.f, A), cubic1(ty, A), cubic1(1.f - ty, A), cubic2(2.f - ty, A)};
    float cx[4] = {cubic2(tx + 1.f, A), cubic1(tx, A), cubic1(1.f - tx, A), cubic2(2.f - tx, A)};
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        int yy = min(max(iy - 1 + a, 0), g - 1);
        float row = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int xx = min(max(ix - 1 + b, 0), g - 1);
            row += table[(long long)(1 + yy * g + xx) * D + d] * cx[b];
        }
        acc += row * cy[a];
    }
    out[i] = acc;
}

int pos_embed_interp_launch(const float *pos_table, float *pos_out, int grid, int ph, int pw, int D, float scale_y, float scale_x, cudaStream_t stream) {
    long long total = (long long)(1 + ph * pw) * D;
    D2S_LAUNCH(pos_embed_interp_kernel, ceil_div(total, 256), 256, 0, stream, pos_table, pos_out, grid, ph, pw, D, scale_y, scale_x);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

// ---- ConvTranspose (kernel == stride) scatter ----
__global__ void pixel_shuffle_kernel(const __half *__restrict__ in, __half *__restrict__ out, int B, int h, int w, int f, int C, int Cp) {
    pdl_sync();
    const int c8 = C / 8;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * h * w * f * f * c8;
    if (i >= total) return;
    int c = (int)(i % c8) * 8;
    long long t = i / c8;
    int ij = (int)(t % (f * f));
    long long m = t / (f * f);
    int x = (int)(m % w), y = (int)((m / w) % h), b = (int)(m / ((long long)w * h));
    int ii = ij / f, jj = ij % f;
    uint4 v = *(const uint4 *)(in + m * (long long)(f * f * C) + ij * C + c);
    *(uint4 *)(out + (((long long)b * h * f + y * f + ii) * (w * f) + x * f + jj) * Cp + c) = v;
}

int pixel_shuffle_launch(const __half *gemm_out, __half *out, int B, int h, int w, int f, int C, int Cp, cudaStream_t stream) {
    D2S_REQUIRE(C % 8 == 0, "pixel_shuffle: C=%d must be a multiple of 8", C);
    long long total = (long long)B * h * w * f * f * (C / 8);
    D2S_LAUNCH(pixel_shuffle_kernel, ceil_div(total, 256), 256, 0, stream, gemm_out, out, B, h, w, f, C, Cp);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

// ---- explicit im2col for the 3x3 / stride 2 / pad 1 conv of the coarsest reassemble level ----
__global__ void im2col_s2_kernel(const __half *__restrict__ in, __half *__restrict__ out, int B, int h, int w, int Cp, int oh, int ow) {
    pdl_sync();
    const int c8 = Cp / 8;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * oh * ow * 9 * c8;
    if (i >= total) return;
    int c = (int)(i % c8) * 8;
    long long t = i / c8;
    int tap = (int)(t % 9);
    long long m = t / 9;
    int ox = (int)(m % ow), oy = (int)((m / ow) % oh), b = (int)(m / ((long long)ow * oh));
    int y = 2 * oy + tap / 3 - 1, x = 2 * ox + tap % 3 - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < h && x >= 0 && x < w) v = *(const uint4 *)(in + (((long long)b * h + y) * w + x) * Cp + c);
    *(uint4 *)(out + m * (long long)(9 * Cp) + tap * Cp + c) = v;
}

int im2col_s2_launch(const __half *in, __half *out, int B, int h, int w, int Cp, int oh, int ow, cudaStream_t stream) {
    long long total = (long long)B * oh * ow * 9 * (Cp / 8);
    D2S_LAUNCH(im2col_s2_kernel, ceil_div(total, 256), 256, 0, stream, in, out, B, h, w, Cp, oh, ow);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

// ---- bilinear upsampling, align_corners=True, NHWC fp16 (ATen upsample_bilinear2d, fp32 accumulate) ----
__global__ void upsample_nhwc_kernel(const __half *__restrict__ in, __half *__restrict__ out, int B, int h, int w, int C, int oh, int ow,
                                     float rh, float rw) {
    pdl_sync();
    const int c8 = C / 8;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * oh * ow * c8;
    if (i >= total) return;
    int c = (int)(i % c8) * 8;
    long long m = i / c8;
    int x = (int)(m % ow), y = (int)((m / ow) % oh), b = (int)(m / ((long long)ow * oh));
    float h1r = rh * (float)y, w1r = rw * (float)x;
    int h1 = (int)h1r, w1 = (int)w1r;
    int h1p = (h1 < h - 1) ? 1 : 0, w1p = (w1 < w - 1) ? 1 : 0;
    float h1l = h1r - (float)h1, h0l = 1.f - h1l, w1l = w1r - (float)w1, w0l = 1.f - w1l;
    const __half *p00 = in + (((long long)b * h + h1) * w + w1) * C + c;
    const __half *p01 = p00 + (long long)w1p * C, *p10 = p00 + (long long)h1p * w * C, *p11 = p10 + (long long)w1p * C;
    uint4 u00 = *(const uint4 *)p00, u01 = *(const uint4 *)p01, u10 = *(const uint4 *)p10, u11 = *(const uint4 *)p11;
    const __half2 *a = (const __half2 *)&u00, *bq = (const __half2 *)&u01, *cq = (const __half2 *)&u10, *dq = (const __half2 *)&u11;
    __half2 r[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        float2 f00 = __half22float2(a[t]), f01 = __half22float2(bq[t]), f10 = __half22float2(cq[t]), f11 = __half22float2(dq[t]);
        float vx = h0l * (w0l * f00.x + w1l * f01.x) + h1l * (w0l * f10.x + w1l * f11.x);
        float vy = h0l * (w0l * f00.y + w1l * f01.y) + h1l * (w0l * f10.y + w1l * f11.y);
        r[t] = __floats2half2_rn(vx, vy);
    }
    *(uint4 *)(out + m * C + c) = *(const uint4 *)r;
}

int upsample_nhwc_launch(const __half *in, __half *out, int B, int h, int w, int C, int oh, int ow, cudaStream_t stream) {
    D2S_REQUIRE(C % 8 == 0, "upsample: C=%d must be a multiple of 8", C);
    float rh = oh > 1 ? (float)(h - 1) / (float)(oh - 1) : 0.f, rw = ow > 1 ? (float)(w - 1) / (float)(ow - 1) : 0.f;
    long long total = (long long)B * oh * ow * (C / 8);
    D2S_LAUNCH(upsample_nhwc_kernel, ceil_div(total, 256), 256, 0, stream, in, out, B, h, w, C, oh, ow, rh, rw);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

// ---- weight packing (engine creation) ----
__global__ void convert_pad_kernel(const float *__restrict__ src, __half *__restrict__ dst, int rows, int cols, int ld) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * ld) return;
    int c = (int)(i % ld);
    long long r = i / ld;
    dst[i] = __float2half_rn(c < cols ? src[r * cols + c] : 0.f);
}
int convert_pad_launch(const float *src, __half *dst, int rows, int cols, int ld, cudaStream_t stream) {
    D2S_LAUNCH(convert_pad_kernel, ceil_div((long long)rows * ld, 256), 256, 0, stream, src, dst, rows, cols, ld);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

__global__ void conv_weight_kernel(const float *__restrict__ src, __half *__restrict__ dst, int N, int Cin, int Cp) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * 9 * Cp) return;
    int c = (int)(i % Cp);
    int tap = (int)((i / Cp) % 9);
    long long n = i / (9LL * Cp);
    dst[i] = __float2half_rn(c < Cin ? src[(n * Cin + c) * 9 + tap] : 0.f);
}
int conv_weight_launch(const float *src, __half *dst, int N, int Cin, int Cp, cudaStream_t stream) {
    D2S_LAUNCH(conv_weight_kernel, ceil_div((long long)N * 9 * Cp, 256), 256, 0, stream, src, dst, N, Cin, Cp);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

__global__ void convt_weight_kernel(const float *__restrict__ src, __half *__restrict__ dst, int Cin, int Cout, int f, int Kp) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)f * f * Cout * Kp;
    if (i >= total) return;
    int ci = (int)(i % Kp);
    long long row = i / Kp;
    int co = (int)(row % Cout), ij = (int)(row / Cout);
    dst[i] = __float2half_rn(ci < Cin ? src[((long long)ci * Cout + co) * f * f + ij] : 0.f);
}
int convt_weight_launch(const float *src, __half *dst, int Cin, int Cout, int f, int Kp, cudaStream_t stream) {
    D2S_LAUNCH(convt_weight_kernel, ceil_div((long long)f * f * Cout * Kp, 256), 256, 0, stream, src, dst, Cin, Cout, f, Kp);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

__global__ void relu_copy_kernel(const __half2 *__restrict__ in, __half2 *__restrict__ out, size_t n2) {
    pdl_sync();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    out[i] = __hmax2(in[i], __float2half2_rn(0.f));
}
int relu_copy_launch(const __half *in, __half *out, size_t n, cudaStream_t stream) {
    D2S_LAUNCH(relu_copy_kernel, ceil_div((long long)(n / 2), 256), 256, 0, stream, (const __half2 *)in, (__half2 *)out, n / 2);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

int zero_launch(void *p, size_t bytes, cudaStream_t stream) {
    D2S_CHECK_CUDA(cudaMemsetAsync(p, 0, bytes, stream));
    return D2S_OK;
}

}  // namespace d2s
