// Non-GEMM layers of the depth network (all memory/latency bound, warp-shuffle reductions).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace d2s {

// attention.cu
int attention_launch(const __half *qkv, __half *out, int B, int N, int D, int heads, cudaStream_t stream);

// attention_tc.cu — the same attention on tcgen05 (plan = tensor maps over qkv / the V^T scratch; vt: attention_tc_vt_elems() halves, zeroed)
struct AttnTcPlan { CUtensorMap tmQK, tmV; const __half *qkv; __half *vt, *out; int B, N, D, heads, Npad; };
size_t attention_tc_vt_elems(int B, int N, int heads);
int attention_tc_plan(AttnTcPlan *p, const __half *qkv, __half *vt, __half *out, int B, int N, int D, int heads);
int attention_tc_launch(const AttnTcPlan *p, cudaStream_t stream, bool vt_ready = false);   // vt_ready: V^T was already written (fused qkv epilogue)

// layers.cu
// LayerNorm over the last dim of fp32 rows -> fp16 (GEMM A operand).  Row r of the output reads input row
// in_row(r) = r + r / rows_per_img * skip + skip  when skip_cls (drops each image's cls token), else r.
int layernorm_launch(const float *x, const float *gamma, const float *beta, __half *y, int rows, int D, float eps,
                     int skip_cls, int tokens_per_img, cudaStream_t stream);
// pixel_values [B,3,H,W] (fp32 or fp16) -> im2col'd patches [B*ph*pw, Kp] fp16, k = c*196 + ky*14 + kx, zero padded to Kp
int patch_im2col_launch(const void *pix, int in_dtype, __half *out, int B, int H, int W, int patch, int Kp, cudaStream_t stream);
// X[b, 0] = cls + pos[0];  X[b, 1+p] = patches[b*P + p] + pos[1+p]   (fp32 residual stream)
int assemble_tokens_launch(const __half *patches, const float *cls, const float *pos, float *x, int B, int P, int D, cudaStream_t stream);
// bicubic (A=-0.75, align_corners=False) resample of the [g,g,D] position table to [ph,pw,D], fp32 (HF dinov2:57-96)
// scale_y/scale_x: source step per output step (HF: grid/ph; VDA dinov2.py:179-210: grid/(ph+0.1), the scale_factor form)
int pos_embed_interp_launch(const float *pos_table, float *pos_out, int grid, int ph, int pw, int D, float scale_y, float scale_x, cudaStream_t stream);
// ConvTranspose2d with kernel == stride == f expressed as GEMM + this scatter:  gemm_out [B*h*w, f*f*C] -> NHWC [B, f*h, f*w, Cp]
int pixel_shuffle_launch(const __half *gemm_out, __half *out, int B, int h, int w, int f, int C, int Cp, cudaStream_t stream);
// explicit im2col for the stride-2 3x3/pad-1 conv: NHWC [B,h,w,Cp] -> [B*oh*ow, 9*Cp]
int im2col_s2_launch(const __half *in, __half *out, int B, int h, int w, int Cp, int oh, int ow, cudaStream_t stream);
// bilinear, align_corners=True, NHWC fp16 [B,h,w,C] -> [B,oh,ow,C]  (C multiple of 8)
int upsample_nhwc_launch(const __half *in, __half *out, int B, int h, int w, int C, int oh, int ow, cudaStream_t stream);
// fp32 [rows, cols] -> fp16 [rows, ld] with zero padding of columns cols..ld (weights at engine creation)
int convert_pad_launch(const float *src, __half *dst, int rows, int cols, int ld, cudaStream_t stream);
// 3x3 conv weight [N, Cin, 3, 3] fp32 -> [N, 9*Cp] fp16 with k = (ky*3+kx)*Cp + c
int conv_weight_launch(const float *src, __half *dst, int N, int Cin, int Cp, cudaStream_t stream);
// ConvTranspose weight [Cin, Cout, f, f] fp32 -> GEMM B matrix [f*f*Cout, Kp] fp16: row (i*f+j)*Cout + co, col ci
int convt_weight_launch(const float *src, __half *dst, int Cin, int Cout, int f, int Kp, cudaStream_t stream);
int relu_copy_launch(const __half *in, __half *out, size_t n, cudaStream_t stream);
int zero_launch(void *p, size_t bytes, cudaStream_t stream);


// temporal.cu — streaming Video-Depth-Anything layers
// GroupNorm(32) of an NHWC fp16 map [d, Cp] (C real channels) -> fp16 [d, C]; partials: groupnorm32_partial_floats(d) floats
int groupnorm32_launch(const __half *x, float *partials, const float *w, const float *b, __half *y, int d, int C, int Cp, float eps, cudaStream_t stream);
size_t groupnorm32_partial_floats(int d);
// [rows, 2*inner] (value | gate) -> value * gelu(gate)  [rows, inner]
int geglu_launch(const __half *in, __half *out, long long rows, int inner, cudaStream_t stream);
int cast_f16_launch(const float *in, __half *out, long long n, cudaStream_t stream);
// attention of the newest frame over the 32-frame ring (see temporal.cu); also stores the newest K'|V' into ring slot t % 32
int temporal_attention_launch(const __half *qkv, __half *ring, const float *pe, const long long *t_ptr, __half *out, int d, int C, cudaStream_t stream);
int frame_counter_inc_launch(long long *t, cudaStream_t stream);

}  // namespace d2s
