// The depth engine: DINOv2 ViT encoder + DPT neck/head of Depth Anything V2 as a replayable plan of sm_100a kernels.
//
// Replaces what DepthModelWrapper.__call__ dispatches to (reference depth.py:1763-1781), i.e. HF transformers'
// DepthAnythingForDepthEstimation.forward (modeling_depth_anything.py:330-413; dinov2/modeling_dinov2.py:38-624).
//
// Data layout in HBM
//   residual stream X        fp32 [B*N, D]                  (N = 1 + ph*pw tokens; never rounded to fp16)
//   GEMM operands            fp16, K-major: activations [rows, K], weights [N_out, K]  (nn.Linear layout as is)
//   feature maps             fp16 NHWC [B, h, w, Cp], Cp = channels rounded up to 64 (zero padded) so that a 64-channel
//                            slab is exactly one 128-byte TMA/UMMA swizzle row
//   accumulators             fp32 in TMEM
// LayerScale is folded into the proj / fc2 weights and biases when the blob is packed (weights.py), q/k/v are one
// [3D, D] matrix, ConvTranspose(k == stride) is a GEMM followed by a pixel shuffle, 3x3 convs are implicit GEMMs whose
// A operand is fetched by 4-D TMA boxes (out-of-bounds = zero = padding), the final 1x1 conv + ReLU lives in the epilogue
// of the last 3x3 conv.  One plan (buffers + tensor maps + CUDA graph) is built per (B, H, W) and replayed per frame.
#include "engine.cuh"

namespace d2s {

// ---- weight upload helpers -------------------------------------------------------------------------
struct BlobReader {
    const float *dev;  // blob already on the device
    size_t n, off = 0;
    bool ok = true;
    const float *take(size_t k) {
        if (off + k > n) { ok = false; return dev; }
        const float *p = dev + off;
        off += k;
        return p;
    }
};

template <typename T>
static int dev_alloc(d2s_engine *e, T **p, size_t count) {
    void *q = nullptr;
    D2S_CHECK_CUDA(cudaMalloc(&q, count * sizeof(T)));
    e->allocs.push_back(q);
    e->weight_bytes += count * sizeof(T);
    *p = (T *)q;
    return D2S_OK;
}
#define TRY(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

static int load_f32(d2s_engine *e, BlobReader &r, float **dst, size_t n, cudaStream_t st) {
    TRY(dev_alloc(e, dst, n));
    D2S_CHECK_CUDA(cudaMemcpyAsync(*dst, r.take(n), n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return D2S_OK;
}
// [rows, cols] fp32 -> fp16 [rows, ld]
static int load_mat(d2s_engine *e, BlobReader &r, __half **dst, int rows, int cols, int ld, cudaStream_t st) {
    TRY(dev_alloc(e, dst, (size_t)rows * ld));
    return convert_pad_launch(r.take((size_t)rows * cols), *dst, rows, cols, ld, st);
}
static int load_conv3(d2s_engine *e, BlobReader &r, __half **dst, int N, int Cin, int Cp, cudaStream_t st) {
    TRY(dev_alloc(e, dst, (size_t)N * 9 * Cp));
    return conv_weight_launch(r.take((size_t)N * Cin * 9), *dst, N, Cin, Cp, st);
}
__global__ void expand_bias_kernel(const float *b, float *out, int C, int reps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C * reps) out[i] = b[i % C];
}
static int load_convt(d2s_engine *e, BlobReader &r, __half **w, float **b, int C, int f, int Kp, cudaStream_t st) {
    TRY(dev_alloc(e, w, (size_t)f * f * C * Kp));
    TRY(convt_weight_launch(r.take((size_t)C * C * f * f), *w, C, C, f, Kp, st));
    TRY(dev_alloc(e, b, (size_t)f * f * C));
    D2S_LAUNCH(expand_bias_kernel, ceil_div(f * f * C, 256), 256, 0, st, r.take(C), *b, C, f * f);
    D2S_POST_LAUNCH();
    return D2S_OK;
}
static int load_rcu(d2s_engine *e, BlobReader &r, RcuW *w, int F, cudaStream_t st) {
    TRY(load_conv3(e, r, &w->c1_w, F, F, F, st)); TRY(load_f32(e, r, &w->c1_b, F, st));
    TRY(load_conv3(e, r, &w->c2_w, F, F, F, st)); TRY(load_f32(e, r, &w->c2_b, F, st));
    return D2S_OK;
}

static int upload_weights(d2s_engine *e, const void *blob, size_t nbytes) {
    const d2s_model_config &c = e->cfg;
    cudaStream_t st = 0;
    float *dblob = nullptr;
    D2S_CHECK_CUDA(cudaMalloc(&dblob, nbytes));
    D2S_CHECK_CUDA(cudaMemcpy(dblob, blob, nbytes, cudaMemcpyHostToDevice));
    BlobReader r{dblob, nbytes / sizeof(float)};
    const int D = e->D, K0 = 3 * c.patch * c.patch, F = e->F;
    int rc = [&]() -> int {
        TRY(load_mat(e, r, &e->patch_w, D, K0, e->Kpatch, st));
        TRY(load_f32(e, r, &e->patch_b, D, st));
        TRY(load_f32(e, r, &e->cls, D, st));
        TRY(load_f32(e, r, &e->pos_table, (size_t)(1 + c.pos_grid * c.pos_grid) * D, st));
        e->layers.resize(c.layers);
        for (auto &l : e->layers) {
            TRY(load_f32(e, r, &l.ln1_w, D, st)); TRY(load_f32(e, r, &l.ln1_b, D, st));
            TRY(load_mat(e, r, &l.qkv_w, 3 * D, D, D, st)); TRY(load_f32(e, r, &l.qkv_b, 3 * D, st));
            TRY(load_mat(e, r, &l.proj_w, D, D, D, st)); TRY(load_f32(e, r, &l.proj_b, D, st));
            TRY(load_f32(e, r, &l.ln2_w, D, st)); TRY(load_f32(e, r, &l.ln2_b, D, st));
            TRY(load_mat(e, r, &l.fc1_w, c.mlp_hidden, D, D, st)); TRY(load_f32(e, r, &l.fc1_b, c.mlp_hidden, st));
            TRY(load_mat(e, r, &l.fc2_w, D, c.mlp_hidden, c.mlp_hidden, st)); TRY(load_f32(e, r, &l.fc2_b, D, st));
        }
        TRY(load_f32(e, r, &e->norm_w, D, st)); TRY(load_f32(e, r, &e->norm_b, D, st));
        for (int i = 0; i < 4; ++i) {
            TRY(load_mat(e, r, &e->re_proj_w[i], e->c[i], D, D, st));
            TRY(load_f32(e, r, &e->re_proj_b[i], e->c[i], st));
        }
        TRY(load_convt(e, r, &e->up0_w, &e->up0_b, e->c[0], 4, e->cp[0], st));
        TRY(load_convt(e, r, &e->up1_w, &e->up1_b, e->c[1], 2, e->cp[1], st));
        TRY(load_conv3(e, r, &e->dn3_w, e->c[3], e->c[3], e->cp[3], st)); TRY(load_f32(e, r, &e->dn3_b, e->c[3], st));
        for (int i = 0; i < 4; ++i) TRY(load_conv3(e, r, &e->neck_w[i], F, e->c[i], e->cp[i], st));
        for (int j = 0; j < 4; ++j) {
            TRY(load_mat(e, r, &e->fus[j].proj_w, F, F, F, st)); TRY(load_f32(e, r, &e->fus[j].proj_b, F, st));
            TRY(load_rcu(e, r, &e->fus[j].rl1, F, st));
            TRY(load_rcu(e, r, &e->fus[j].rl2, F, st));
        }
        TRY(load_conv3(e, r, &e->head_c1_w, e->Fh, F, F, st)); TRY(load_f32(e, r, &e->head_c1_b, e->Fh, st));
        TRY(load_conv3(e, r, &e->head_c2_w, c.head_hidden, e->Fh, e->Fhp, st)); TRY(load_f32(e, r, &e->head_c2_b, c.head_hidden, st));
        TRY(load_f32(e, r, &e->head_c3_w, c.head_hidden, st));
        return D2S_OK;
    }();
    const float *b3 = rc == D2S_OK ? r.take(1) : nullptr;
    if (rc == D2S_OK && c.temporal) rc = [&]() -> int {
        const int Cs[4] = {e->c[2], e->c[3], F, F};
        for (int m = 0; m < 4; ++m) {
            TemporalW &t = e->tm[m];
            const int C = t.C = Cs[m];
            TRY(load_f32(e, r, &t.gn_w, C, st)); TRY(load_f32(e, r, &t.gn_b, C, st));
            TRY(load_mat(e, r, &t.in_w, C, C, C, st)); TRY(load_f32(e, r, &t.in_b, C, st));
            for (int a = 0; a < 2; ++a) {
                TAttnW &w = t.att[a];
                TRY(load_f32(e, r, &w.ln_w, C, st)); TRY(load_f32(e, r, &w.ln_b, C, st));
                TRY(load_mat(e, r, &w.qkv_w, 3 * C, C, C, st));
                TRY(load_f32(e, r, &w.pe, (size_t)32 * 3 * C, st));
                TRY(load_mat(e, r, &w.out_w, C, C, C, st)); TRY(load_f32(e, r, &w.out_b, C, st));
            }
            TRY(load_f32(e, r, &t.ffln_w, C, st)); TRY(load_f32(e, r, &t.ffln_b, C, st));
            TRY(load_mat(e, r, &t.ff1_w, 8 * C, C, C, st)); TRY(load_f32(e, r, &t.ff1_b, 8 * C, st));
            TRY(load_mat(e, r, &t.ff2_w, C, 4 * C, 4 * C, st)); TRY(load_f32(e, r, &t.ff2_b, C, st));
            TRY(load_mat(e, r, &t.out_w, C, C, C, st)); TRY(load_f32(e, r, &t.out_b, C, st));
        }
        return D2S_OK;
    }();
    if (rc == D2S_OK) {
        if (!r.ok || r.off != r.n) rc = set_error(D2S_ERR_INVALID, "d2s_create: weight blob has %zu floats, the config needs %zu", r.n, r.off);
        else if (cudaMemcpy(&e->head_c3_b, b3, sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) rc = set_error(D2S_ERR_CUDA, "d2s_create: blob readback failed");
    }
    cudaError_t se = cudaStreamSynchronize(st);
    cudaFree(dblob);
    if (rc == D2S_OK && se != cudaSuccess) rc = set_error(D2S_ERR_CUDA, "d2s_create: weight upload failed: %s", cudaGetErrorString(se));
    return rc;
}

// ---- plan construction -------------------------------------------------------------------------------
template <typename T>
static int plan_alloc(ShapePlan *sp, T **p, size_t count) {
    void *q = nullptr;
    size_t bytes = count * sizeof(T);
    D2S_CHECK_CUDA(cudaMalloc(&q, bytes));
    D2S_CHECK_CUDA(cudaMemset(q, 0, bytes));  // zero so that the channel padding of NHWC maps is (and stays) zero
    sp->allocs.push_back(q);
    sp->total_bytes += bytes;
    *p = (T *)q;
    return D2S_OK;
}

// the next K'|V' ring of the plan's stream state (allocated by the first plan built for this video, found again by the others)
static int state_ring(ShapePlan *sp, __half **p, size_t elems) {
    StreamState *ss = sp->state;
    if (!ss) return set_error(D2S_ERR_INVALID, "temporal plan without a stream state");
    const size_t i = sp->ring_cursor++;
    if (i < ss->rings.size()) {
        if (ss->ring_elems[i] != elems) return set_error(D2S_ERR_INVALID, "temporal stream state does not match the plan (ring %zu: %zu vs %zu elements)", i, ss->ring_elems[i], elems);
        *p = ss->rings[i];
        return D2S_OK;
    }
    void *q = nullptr;
    D2S_CHECK_CUDA(cudaMalloc(&q, elems * sizeof(__half)));
    D2S_CHECK_CUDA(cudaMemset(q, 0, elems * sizeof(__half)));
    ss->rings.push_back((__half *)q); ss->ring_elems.push_back(elems); ss->bytes += elems * sizeof(__half);
    *p = (__half *)q;
    return D2S_OK;
}

static int add_gemm(ShapePlan *sp, const GemmPlan &gp) {
    sp->gemms.push_back(gp);
    const GemmPlan *p = &sp->gemms.back();
    sp->ops.push_back([p](cudaStream_t st) { return gemm_launch(p, st); });
    return D2S_OK;
}
static int add_linear(ShapePlan *sp, const __half *A, int lda, const __half *Bw, int ldb, int M, int N, int K, const GemmEpi &epi) {
    GemmPlan gp;
    TRY(gemm_plan_linear(&gp, A, lda, Bw, ldb, M, N, K, epi));
    return add_gemm(sp, gp);
}
static int add_conv(ShapePlan *sp, const __half *A, int B, int H, int W, int Cp, const __half *Bw, int N, const GemmEpi &epi) {
    GemmPlan gp;
    ConvGeom g; g.B = B; g.H = H; g.W = W; g.Cp = Cp;
    TRY(gemm_plan_conv3x3(&gp, A, g, Bw, N, epi));
    return add_gemm(sp, gp);
}

// Pre-activation residual unit (HF depth_anything:96-136): out = conv2(relu(conv1(relu(x)))) + x (+ extra)
// x_relu is relu(x), produced by whoever produced x.  Writes out (and relu(out) if out_relu).
static int add_rcu(ShapePlan *sp, const RcuW &w, const __half *x, const __half *x_relu, const __half *extra, __half *tmp, __half *out,
                   __half *out_relu, int B, int h, int wd, int F) {
    GemmEpi e1; e1.bias = w.c1_b; e1.act = ACT_RELU; e1.c16 = tmp; e1.ldc = F;
    TRY(add_conv(sp, x_relu, B, h, wd, F, w.c1_w, F, e1));
    GemmEpi e2; e2.bias = w.c2_b; e2.res1 = x; e2.res2 = extra; e2.c16 = out; e2.c16_relu = out_relu; e2.ldc = F;
    TRY(add_conv(sp, tmp, B, h, wd, F, w.c2_w, F, e2));
    return D2S_OK;
}

// One TemporalModule on an NHWC fp16 map x [d, Cp] of one frame -> out [d, Cp] (motion_module.py:103-134; streaming form, see
// temporal.cu).  Kernels: GroupNorm (2) | proj_in GEMM | 2 x (LN, qkv GEMM, ring attention, out GEMM) | LN, ff1 GEMM, GEGLU,
// ff2 GEMM | cast, proj_out GEMM (+ residual x) = 18 launches, all GEMMs on the tcgen05 kernel.
static int add_temporal(ShapePlan *sp, const TemporalW &t, const __half *x, __half *out, int d, int Cp, int m) {
    const int C = t.C;
    float *part, *hs;
    __half *gn16, *ln16, *qkv16, *att16, *ff16, *gg16, *hs16;
    TRY(plan_alloc(sp, &part, groupnorm32_partial_floats(d)));
    TRY(plan_alloc(sp, &hs, (size_t)d * C));
    TRY(plan_alloc(sp, &gn16, (size_t)d * C)); TRY(plan_alloc(sp, &ln16, (size_t)d * C)); TRY(plan_alloc(sp, &qkv16, (size_t)d * 3 * C));
    TRY(plan_alloc(sp, &att16, (size_t)d * C)); TRY(plan_alloc(sp, &ff16, (size_t)d * 8 * C)); TRY(plan_alloc(sp, &gg16, (size_t)d * 4 * C));
    TRY(plan_alloc(sp, &hs16, (size_t)d * C));
    const float *gw = t.gn_w, *gb = t.gn_b;
    sp->ops.push_back([=](cudaStream_t st) { return groupnorm32_launch(x, part, gw, gb, gn16, d, C, Cp, 1e-6f, st); });
    GemmEpi ei; ei.bias = t.in_b; ei.x32 = hs; ei.x32_assign = 1; ei.ldc = C;
    TRY(add_linear(sp, gn16, C, t.in_w, C, d, C, C, ei));
    const long long *tc = sp->frame_counter;
    for (int a = 0; a < 2; ++a) {
        const TAttnW w = t.att[a];
        __half *ring;
        TRY(state_ring(sp, &ring, (size_t)d * 32 * 2 * C));
        sp->ops.push_back([=](cudaStream_t st) { return layernorm_launch(hs, w.ln_w, w.ln_b, ln16, d, C, 1e-5f, 0, 0, st); });
        GemmEpi eq; eq.c16 = qkv16; eq.ldc = 3 * C;
        TRY(add_linear(sp, ln16, C, w.qkv_w, C, d, 3 * C, C, eq));
        sp->ops.push_back([=](cudaStream_t st) { return temporal_attention_launch(qkv16, ring, w.pe, tc, att16, d, C, st); });
        GemmEpi eo; eo.bias = w.out_b; eo.x32 = hs; eo.ldc = C;
        TRY(add_linear(sp, att16, C, w.out_w, C, d, C, C, eo));
    }
    const float *fw = t.ffln_w, *fb = t.ffln_b;
    sp->ops.push_back([=](cudaStream_t st) { return layernorm_launch(hs, fw, fb, ln16, d, C, 1e-5f, 0, 0, st); });
    GemmEpi e1; e1.bias = t.ff1_b; e1.c16 = ff16; e1.ldc = 8 * C;
    TRY(add_linear(sp, ln16, C, t.ff1_w, C, d, 8 * C, C, e1));
    sp->ops.push_back([=](cudaStream_t st) { return geglu_launch(ff16, gg16, d, 4 * C, st); });
    GemmEpi e2; e2.bias = t.ff2_b; e2.x32 = hs; e2.ldc = C;
    TRY(add_linear(sp, gg16, 4 * C, t.ff2_w, 4 * C, d, C, 4 * C, e2));
    sp->ops.push_back([=](cudaStream_t st) { return cast_f16_launch(hs, hs16, (long long)d * C, st); });
    GemmEpi eo; eo.bias = t.out_b; eo.res1 = x; eo.c16 = out; eo.ldc = Cp;
    TRY(add_linear(sp, hs16, C, t.out_w, C, d, C, C, eo));
    sp->taps["temporal" + std::to_string(m)] = {out, (size_t)d * Cp, D2S_F16};
    return D2S_OK;
}

static int build_plan(d2s_engine *e, ShapePlan *sp) {
    const d2s_model_config &c = e->cfg;
    const int B = sp->B, H = sp->H, W = sp->W, D = e->D, F = e->F;
    const int ph = H / c.patch, pw = W / c.patch, P = ph * pw, N = P + 1, M = B * N, BP = B * P;
    const size_t in_es = sp->in_dtype == D2S_F32 ? 4 : 2, out_es = sp->out_dtype == D2S_F32 ? 4 : 2;
    sp->in_bytes = (size_t)B * 3 * H * W * in_es;
    sp->out_bytes = (size_t)B * H * W * out_es;
    uint8_t *in_stage, *out_stage;
    TRY(plan_alloc(sp, &in_stage, sp->in_bytes)); TRY(plan_alloc(sp, &out_stage, sp->out_bytes));
    if (c.temporal) {
        if (!sp->state->frame_counter) {
            D2S_CHECK_CUDA(cudaMalloc((void **)&sp->state->frame_counter, sizeof(long long)));
            D2S_CHECK_CUDA(cudaMemset(sp->state->frame_counter, 0, sizeof(long long)));
        }
        sp->frame_counter = sp->state->frame_counter;
    }
    sp->in_stage = in_stage; sp->out_stage = out_stage;

    // ---------------- embeddings ----------------
    __half *patches, *pe; float *X, *pos;
    TRY(plan_alloc(sp, &patches, (size_t)BP * e->Kpatch));
    TRY(plan_alloc(sp, &pe, (size_t)BP * D));
    TRY(plan_alloc(sp, &X, (size_t)M * D));
    TRY(plan_alloc(sp, &pos, (size_t)N * D));
    // position table for this grid: input independent, computed once here (HF dinov2:57-96)
    if (ph == c.pos_grid && pw == c.pos_grid) D2S_CHECK_CUDA(cudaMemcpy(pos, e->pos_table, (size_t)N * D * sizeof(float), cudaMemcpyDeviceToDevice));
    else if (c.pos_interp_offset != 0.f)   // VDA: F.interpolate(scale_factor=((ph+off)/g, (pw+off)/g)) -> source step = 1/scale_factor (dinov2.py:190-203)
        TRY(pos_embed_interp_launch(e->pos_table, pos, c.pos_grid, ph, pw, D, (float)(1.0 / ((double)(float)(ph + c.pos_interp_offset) / (double)c.pos_grid)),
                                    (float)(1.0 / ((double)(float)(pw + c.pos_interp_offset) / (double)c.pos_grid)), 0));
    else TRY(pos_embed_interp_launch(e->pos_table, pos, c.pos_grid, ph, pw, D, (float)c.pos_grid / (float)ph, (float)c.pos_grid / (float)pw, 0));
    D2S_CHECK_CUDA(cudaStreamSynchronize(0));
    {
        const int in_dtype = sp->in_dtype, patch = c.patch, Kp = e->Kpatch;
        sp->ops.push_back([=](cudaStream_t st) { return patch_im2col_launch(in_stage, in_dtype, patches, B, H, W, patch, Kp, st); });
        GemmEpi ep; ep.bias = e->patch_b; ep.c16 = pe; ep.ldc = D;
        TRY(add_linear(sp, patches, e->Kpatch, e->patch_w, e->Kpatch, BP, D, e->Kpatch, ep));
        const float *cls = e->cls;
        sp->ops.push_back([=](cudaStream_t st) { return assemble_tokens_launch(pe, cls, pos, X, B, P, D, st); });
    }

    // ---------------- encoder ----------------
    __half *ln16, *qkv16, *att16, *h16, *feat[4];
    TRY(plan_alloc(sp, &ln16, (size_t)M * D)); TRY(plan_alloc(sp, &qkv16, (size_t)M * 3 * D));
    TRY(plan_alloc(sp, &att16, (size_t)M * D)); TRY(plan_alloc(sp, &h16, (size_t)M * c.mlp_hidden));
    for (int i = 0; i < 4; ++i) TRY(plan_alloc(sp, &feat[i], (size_t)BP * D));
    const float eps = c.layer_norm_eps;
    const int heads = c.heads, mlp = c.mlp_hidden;
    // attention: the tcgen05 kernel (attention_tc.cu; V^T comes from the qkv GEMM's epilogue) when there are enough 128-query
    // tiles to fill the GPU (71 vs 101 us per layer at B = 8 x 16 heads), else the mma.sync flash kernel (equal at B = 1, and
    // its 64-query tiles spread over more SMs).  Throughput-policy plans (several frames in flight) always take the tcgen05 kernel:
    // per-SM efficiency, not the latency of one launch, is what counts there (measured: 1796 -> 1866 frames/s at 8 frames in flight,
    // base1080).  D2S_ATTN=tcgen05 | mma forces one.
    const char *attn_env = getenv("D2S_ATTN");
    const bool attn_tc = attn_env && attn_env[0] ? attn_env[0] == 't'
                                                 : (sp->policy == D2S_POLICY_THROUGHPUT || (long long)B * heads * ceil_div(N, 128) >= 2 * kNumSMs);
    AttnTcPlan attn{};
    if (attn_tc) {
        __half *vt16;
        TRY(plan_alloc(sp, &vt16, attention_tc_vt_elems(B, N, heads)));
        TRY(attention_tc_plan(&attn, qkv16, vt16, att16, B, N, D, heads));
    }
    for (int l = 0; l < c.layers; ++l) {
        const LayerW lw = e->layers[l];
        sp->ops.push_back([=](cudaStream_t st) { return layernorm_launch(X, lw.ln1_w, lw.ln1_b, ln16, M, D, eps, 0, N, st); });
        GemmEpi eq; eq.bias = lw.qkv_b; eq.c16 = qkv16; eq.ldc = 3 * D;
        if (attn_tc) { eq.vt = attn.vt; eq.vt_col0 = 2 * D; eq.vt_tokens = N; eq.vt_heads = heads; eq.vt_npad = attn.Npad; }
        TRY(add_linear(sp, ln16, D, lw.qkv_w, D, M, 3 * D, D, eq));
        if (attn_tc) sp->ops.push_back([=](cudaStream_t st) { return attention_tc_launch(&attn, st, true); });
        else sp->ops.push_back([=](cudaStream_t st) { return attention_launch(qkv16, att16, B, N, D, heads, st); });
        GemmEpi epj; epj.bias = lw.proj_b; epj.x32 = X; epj.ldc = D;   // LayerScale folded; X += ...
        TRY(add_linear(sp, att16, D, lw.proj_w, D, M, D, D, epj));
        sp->ops.push_back([=](cudaStream_t st) { return layernorm_launch(X, lw.ln2_w, lw.ln2_b, ln16, M, D, eps, 0, N, st); });
        GemmEpi e1; e1.bias = lw.fc1_b; e1.act = ACT_GELU; e1.c16 = h16; e1.ldc = mlp;
        TRY(add_linear(sp, ln16, D, lw.fc1_w, D, M, mlp, D, e1));
        GemmEpi e2; e2.bias = lw.fc2_b; e2.x32 = X; e2.ldc = D;
        TRY(add_linear(sp, h16, mlp, lw.fc2_w, mlp, M, D, mlp, e2));
        for (int i = 0; i < 4; ++i)
            if (c.out_indices[i] == l + 1) {  // tap: shared final LayerNorm, cls token dropped (HF dinov2:605-618, depth_anything:76-93)
                __half *f = feat[i];
                const float *nw = e->norm_w, *nb = e->norm_b;
                sp->ops.push_back([=](cudaStream_t st) { return layernorm_launch(X, nw, nb, f, BP, D, eps, 1, N, st); });
            }
    }
    sp->taps["hidden_last"] = {X, (size_t)M * D, D2S_F32};
    for (int i = 0; i < 4; ++i) sp->taps["feat" + std::to_string(i)] = {feat[i], (size_t)BP * D, D2S_F16};

    // ---------------- reassemble (HF depth_anything:31-93) ----------------
    const int hs[4] = {4 * ph, 2 * ph, ph, (ph - 1) / 2 + 1}, ws[4] = {4 * pw, 2 * pw, pw, (pw - 1) / 2 + 1};
    __half *R[4];
    {
        __half *r0, *r1, *r3, *ct0, *ct1, *col3;
        TRY(plan_alloc(sp, &r0, (size_t)BP * e->cp[0])); TRY(plan_alloc(sp, &r1, (size_t)BP * e->cp[1])); TRY(plan_alloc(sp, &r3, (size_t)BP * e->cp[3]));
        TRY(plan_alloc(sp, &R[2], (size_t)BP * e->cp[2]));
        TRY(plan_alloc(sp, &ct0, (size_t)BP * 16 * e->c[0])); TRY(plan_alloc(sp, &ct1, (size_t)BP * 4 * e->c[1]));
        TRY(plan_alloc(sp, &R[0], (size_t)B * hs[0] * ws[0] * e->cp[0])); TRY(plan_alloc(sp, &R[1], (size_t)B * hs[1] * ws[1] * e->cp[1]));
        TRY(plan_alloc(sp, &R[3], (size_t)B * hs[3] * ws[3] * e->cp[3]));
        TRY(plan_alloc(sp, &col3, (size_t)B * hs[3] * ws[3] * 9 * e->cp[3]));
        __half *rproj[4] = {r0, r1, R[2], r3};
        for (int i = 0; i < 4; ++i) {
            GemmEpi ep; ep.bias = e->re_proj_b[i]; ep.c16 = rproj[i]; ep.ldc = e->cp[i];
            TRY(add_linear(sp, feat[i], D, e->re_proj_w[i], D, BP, e->c[i], D, ep));
        }
        const int c0 = e->c[0], c1 = e->c[1], cp0 = e->cp[0], cp1 = e->cp[1], cp3 = e->cp[3];
        GemmEpi eu0; eu0.bias = e->up0_b; eu0.c16 = ct0; eu0.ldc = 16 * c0;
        TRY(add_linear(sp, r0, cp0, e->up0_w, cp0, BP, 16 * c0, cp0, eu0));
        __half *R0 = R[0], *R1 = R[1];
        sp->ops.push_back([=](cudaStream_t st) { return pixel_shuffle_launch(ct0, R0, B, ph, pw, 4, c0, cp0, st); });
        GemmEpi eu1; eu1.bias = e->up1_b; eu1.c16 = ct1; eu1.ldc = 4 * c1;
        TRY(add_linear(sp, r1, cp1, e->up1_w, cp1, BP, 4 * c1, cp1, eu1));
        sp->ops.push_back([=](cudaStream_t st) { return pixel_shuffle_launch(ct1, R1, B, ph, pw, 2, c1, cp1, st); });
        const int h3 = hs[3], w3 = ws[3];
        sp->ops.push_back([=](cudaStream_t st) { return im2col_s2_launch(r3, col3, B, ph, pw, cp3, h3, w3, st); });
        GemmEpi ed; ed.bias = e->dn3_b; ed.c16 = R[3]; ed.ldc = cp3;
        TRY(add_linear(sp, col3, 9 * cp3, e->dn3_w, 9 * cp3, B * h3 * w3, e->c[3], 9 * cp3, ed));
    }
    for (int i = 0; i < 4; ++i) sp->taps["reassemble" + std::to_string(i)] = {R[i], (size_t)B * hs[i] * ws[i] * e->cp[i], D2S_F16};
    if (c.temporal) {   // dpt_temporal.py:92-95: temporal modules on layer_3 and layer_4 before the neck convs
        for (int i = 2; i < 4; ++i) {
            __half *o;
            TRY(plan_alloc(sp, &o, (size_t)hs[i] * ws[i] * e->cp[i]));
            TRY(add_temporal(sp, e->tm[i - 2], R[i], o, hs[i] * ws[i], e->cp[i], i - 2));
            R[i] = o;
        }
    }

    // ---------------- neck convs (3x3, no bias) -> features + relu copies ----------------
    __half *Fm[4], *Fr[4];
    for (int i = 0; i < 4; ++i) {
        size_t n = (size_t)B * hs[i] * ws[i] * F;
        TRY(plan_alloc(sp, &Fm[i], n)); TRY(plan_alloc(sp, &Fr[i], n));
        GemmEpi ep; ep.c16 = Fm[i]; ep.c16_relu = Fr[i]; ep.ldc = F;
        TRY(add_conv(sp, R[i], B, hs[i], ws[i], e->cp[i], e->neck_w[i], F, ep));
        sp->taps["neck" + std::to_string(i)] = {Fm[i], n, D2S_F16};
    }

    // ---------------- fusion, coarse -> fine (HF depth_anything:139-203) ----------------
    __half *hidden = nullptr;  // projected output of the previous fusion layer, at the current level's size
    for (int j = 0; j < 4; ++j) {
        const int lv = 3 - j, h = hs[lv], w = ws[lv];
        const int oh = j < 3 ? hs[lv - 1] : 2 * h, ow = j < 3 ? ws[lv - 1] : 2 * w;
        size_t n = (size_t)B * h * w * F, no = (size_t)B * oh * ow * F;
        __half *tmp, *s = Fm[lv], *s_relu = Fr[lv], *u, *up, *proj;
        TRY(plan_alloc(sp, &tmp, n)); TRY(plan_alloc(sp, &u, n)); TRY(plan_alloc(sp, &up, no)); TRY(plan_alloc(sp, &proj, no));
        if (j > 0) {  // hidden + residual_layer1(features[lv])
            TRY(plan_alloc(sp, &s, n)); TRY(plan_alloc(sp, &s_relu, n));
            TRY(add_rcu(sp, e->fus[j].rl1, Fm[lv], Fr[lv], hidden, tmp, s, s_relu, B, h, w, F));
        }
        TRY(add_rcu(sp, e->fus[j].rl2, s, s_relu, nullptr, tmp, u, nullptr, B, h, w, F));
        sp->ops.push_back([=](cudaStream_t st) { return upsample_nhwc_launch(u, up, B, h, w, F, oh, ow, st); });
        GemmEpi ep; ep.bias = e->fus[j].proj_b; ep.c16 = proj; ep.ldc = F;
        TRY(add_linear(sp, up, F, e->fus[j].proj_w, F, B * oh * ow, F, F, ep));
        hidden = proj;
        sp->taps["fused" + std::to_string(j)] = {proj, no, D2S_F16};
        if (c.temporal && j < 2) {   // dpt_temporal.py:102-107: temporal modules on path_4 and path_3
            __half *o;
            TRY(plan_alloc(sp, &o, no));
            TRY(add_temporal(sp, e->tm[2 + j], proj, o, oh * ow, F, 2 + j));
            hidden = o;
        }
    }

    // ---------------- head (HF depth_anything:292-308) ----------------
    {
        const int h8 = 8 * ph, w8 = 8 * pw, Fh = e->Fh, Fhp = e->Fhp;
        __half *c1, *up2;
        TRY(plan_alloc(sp, &c1, (size_t)B * h8 * w8 * Fhp)); TRY(plan_alloc(sp, &up2, (size_t)B * H * W * Fhp));
        GemmEpi e1; e1.bias = e->head_c1_b; e1.c16 = c1; e1.ldc = Fhp;
        TRY(add_conv(sp, hidden, B, h8, w8, F, e->head_c1_w, Fh, e1));
        sp->ops.push_back([=](cudaStream_t st) { return upsample_nhwc_launch(c1, up2, B, h8, w8, Fhp, H, W, st); });
        GemmEpi e2; e2.bias = e->head_c2_b; e2.act = ACT_RELU; e2.w3 = e->head_c3_w; e2.b3 = e->head_c3_b;
        e2.final_act = c.metric ? ACT_SIGMOID : ACT_RELU; e2.max_depth = c.max_depth; e2.depth_out = out_stage; e2.depth_dtype = sp->out_dtype;
        TRY(add_conv(sp, up2, B, H, W, Fhp, e->head_c2_w, c.head_hidden, e2));
        sp->taps["head_conv1"] = {c1, (size_t)B * h8 * w8 * Fhp, D2S_F16};
    }
    sp->taps["depth"] = {out_stage, (size_t)B * H * W, sp->out_dtype};
    if (c.temporal) { long long *fc = sp->frame_counter; sp->ops.push_back([=](cudaStream_t st) { return frame_counter_inc_launch(fc, st); }); }
    // the buffers were zeroed on the legacy default stream; the plan may be replayed on a non-blocking stream
    D2S_CHECK_CUDA(cudaStreamSynchronize(0));
    return D2S_OK;
}

int engine_run_ops(ShapePlan *sp, cudaStream_t st) {
    // Programmatic dependent launch between the kernels of the plan (common.cuh; the GEMM and attention kernels release their
    // dependents when their main loop is over, so the next kernel's prologue runs under their epilogue).  Measured on B200:
    // +3.6 % frames/s with 8 frames in flight (1794 vs 1731), but +0.07 ms for one frame alone (1.48 vs 1.42 ms) — so it is on for
    // throughput-policy plans and off for latency-policy plans.  D2S_PDL=0 | 1 forces it.
    const char *pe = getenv("D2S_PDL");
    g_pdl = pe && pe[0] ? pe[0] == '1' : sp->policy == D2S_POLICY_THROUGHPUT;
    int rc = D2S_OK;
    for (auto &op : sp->ops)
        if ((rc = op(st))) break;
    g_pdl = false;
    return rc;
}

__global__ void tap_to_f32_kernel(const void *src, int dtype, float *dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[i] = dtype == D2S_F32 ? ((const float *)src)[i] : __half2float(((const __half *)src)[i]);
}

}  // namespace d2s

using namespace d2s;

extern "C" int d2s_create(const void *weight_blob, size_t nbytes, const d2s_model_config *cfg, int device, d2s_handle *out) {
    D2S_REQUIRE(weight_blob && cfg && out, "d2s_create: null argument");
    D2S_REQUIRE(cfg->hidden % 128 == 0 && cfg->hidden <= 1024, "d2s_create: hidden=%d must be a multiple of 128, <= 1024", cfg->hidden);
    D2S_REQUIRE(cfg->hidden == cfg->heads * 64, "d2s_create: head dim must be 64 (hidden=%d heads=%d)", cfg->hidden, cfg->heads);
    D2S_REQUIRE(cfg->patch == 14 && cfg->head_hidden == 32, "d2s_create: patch must be 14 and head_hidden 32");
    D2S_REQUIRE(cfg->fusion % 64 == 0 && cfg->mlp_hidden % 64 == 0, "d2s_create: fusion/mlp sizes must be multiples of 64");
    for (int i = 0; i < 4; ++i) D2S_REQUIRE(cfg->neck[i] % 8 == 0 && cfg->neck[i] > 0, "d2s_create: neck[%d]=%d must be a positive multiple of 8", i, cfg->neck[i]);
    if (cfg->temporal)
        D2S_REQUIRE(cfg->neck[2] % 64 == 0 && cfg->neck[3] % 64 == 0 && cfg->fusion % 64 == 0, "d2s_create: temporal modules need neck[2], neck[3] and fusion to be multiples of 64");
    D2S_CHECK_CUDA(cudaSetDevice(device));
    int rc = gemm_init();
    if (rc) return rc;
    d2s_engine *e = new d2s_engine();
    e->cfg = *cfg; e->device = device;
    e->D = cfg->hidden; e->L = cfg->layers; e->Kpatch = round_up(3 * cfg->patch * cfg->patch, 64);
    for (int i = 0; i < 4; ++i) { e->c[i] = cfg->neck[i]; e->cp[i] = round_up(cfg->neck[i], 64); }
    e->F = cfg->fusion; e->Fh = cfg->fusion / 2; e->Fhp = round_up(e->Fh, 64);
    const char *ng = getenv("D2S_NO_GRAPH");
    e->use_graph = !(ng && ng[0] == '1');
    if (const char *mp = getenv("D2S_MAX_PLANS")) { int v = atoi(mp); if (v >= 1) e->max_plans = (size_t)v; }
    rc = upload_weights(e, weight_blob, nbytes);
    if (rc) { d2s_destroy(e); return rc; }
    *out = e;
    return D2S_OK;
}

extern "C" int d2s_destroy(d2s_handle h) {
    if (!h) return D2S_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    h->plans.clear();
    h->states.clear();
    for (void *p : h->allocs) cudaFree(p);
    delete h;
    return D2S_OK;
}

// Find or build the plan for (shape, dtypes, stream, policy).  Building is the only host-synchronous path (like the reference's
// lazy engine build at depth.py:1842-1862): buffers, tensor maps, graph capture.  The cache is LRU-bounded (max_plans): the
// least recently used plan is destroyed after a device synchronise; temporal stream states outlive their plans.
int d2s::engine_plan(d2s_engine *h, int B, int H, int W, int in_dtype, int out_dtype, cudaStream_t st, ShapePlan **out) {
    const d2s_model_config &c = h->cfg;
    D2S_REQUIRE(B >= 1 && H >= 14 && W >= 14 && H % c.patch == 0 && W % c.patch == 0, "d2s_infer: input %dx%dx%d must be a multiple of the patch size", B, H, W);
    D2S_REQUIRE(in_dtype == D2S_F32 || in_dtype == D2S_F16, "d2s_infer: pixel_values dtype %d", in_dtype);
    D2S_REQUIRE(out_dtype == D2S_F32 || out_dtype == D2S_F16, "d2s_infer: output dtype %d", out_dtype);
    D2S_REQUIRE(!c.temporal || B == 1, "d2s_infer: a temporal (Video-Depth-Anything) engine takes one frame per call (B=%d)", B);
    D2S_REQUIRE(c.max_batch <= 0 || B <= c.max_batch, "d2s_infer: batch %d exceeds the engine's max_batch %d", B, c.max_batch);
    D2S_REQUIRE((c.max_h <= 0 || H <= c.max_h) && (c.max_w <= 0 || W <= c.max_w), "d2s_infer: input %dx%d exceeds the engine's max_h x max_w %dx%d", H, W, c.max_h, c.max_w);
    std::vector<long long> key = {B, H, W, in_dtype, out_dtype, (long long)(uintptr_t)st, h->policy};
    std::unique_lock<std::mutex> lock(h->mu);
    auto it = h->plans.find(key);
    if (it == h->plans.end()) {
        if (h->plans.size() >= h->max_plans) {
            auto victim = h->plans.begin();
            for (auto p = h->plans.begin(); p != h->plans.end(); ++p)
                if (p->second->last_use < victim->second->last_use) victim = p;
            D2S_CHECK_CUDA(cudaDeviceSynchronize());      // the victim may still be running on its stream
            if (h->last == victim->second.get()) h->last = nullptr;
            h->plans.erase(victim);
        }
        std::unique_ptr<ShapePlan> sp(new ShapePlan());
        sp->B = B; sp->H = H; sp->W = W; sp->in_dtype = in_dtype; sp->out_dtype = out_dtype; sp->stream = st; sp->policy = h->policy;
        if (c.temporal) {
            std::vector<long long> skey = {(long long)(uintptr_t)st, H, W};
            auto &slot = h->states[skey];
            if (!slot) slot.reset(new StreamState());
            sp->state = slot.get();
        }
        gemm_set_plan_policy(h->policy);
        int rc = build_plan(h, sp.get());
        gemm_set_plan_policy(0);
        if (rc) return rc;
        if (h->use_graph) {
            cudaStream_t cs;
            D2S_CHECK_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
            cudaError_t ce = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
            int orc = ce == cudaSuccess ? engine_run_ops(sp.get(), cs) : D2S_ERR_CUDA;
            cudaError_t ee = cudaStreamEndCapture(cs, &sp->graph);
            cudaStreamDestroy(cs);
            if (ce != cudaSuccess || ee != cudaSuccess || orc) return orc ? orc : set_error(D2S_ERR_CUDA, "d2s_infer: graph capture failed: %s", cudaGetErrorString(ee != cudaSuccess ? ee : ce));
            D2S_CHECK_CUDA(cudaGraphInstantiate(&sp->exec, sp->graph, 0));
        }
        it = h->plans.emplace(key, std::move(sp)).first;
    }
    ShapePlan *sp = it->second.get();
    sp->last_use = ++h->tick;
    h->last = sp;
    *out = sp;
    return D2S_OK;
}

int d2s::engine_run_plan(ShapePlan *sp, cudaStream_t st) {
    if (sp->exec) {
        D2S_CHECK_CUDA(cudaGraphLaunch(sp->exec, st));
        g_launch_count.fetch_add((long long)sp->ops.size(), std::memory_order_relaxed);
        return D2S_OK;
    }
    return engine_run_ops(sp, st);
}

extern "C" int d2s_infer(d2s_handle h, const void *pixel_values, int in_dtype, void *depth_out, int out_dtype, int B, int H, int W,
                         d2s_stream_t stream) {
    D2S_REQUIRE(h && pixel_values && depth_out, "d2s_infer: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    // One plan (activation buffers + tensor maps + graph) per shape AND per stream: frames submitted on different streams
    // run concurrently on disjoint buffers while sharing the (read-only) weights.
    ShapePlan *sp = nullptr;
    int rc = engine_plan(h, B, H, W, in_dtype, out_dtype, st, &sp);
    if (rc) return rc;
    D2S_CHECK_CUDA(cudaMemcpyAsync(sp->in_stage, pixel_values, sp->in_bytes, cudaMemcpyDeviceToDevice, st));
    if ((rc = engine_run_plan(sp, st))) return rc;
    D2S_CHECK_CUDA(cudaMemcpyAsync(depth_out, sp->out_stage, sp->out_bytes, cudaMemcpyDeviceToDevice, st));
    return D2S_OK;
}

extern "C" int d2s_set_policy(d2s_handle h, int policy) {
    D2S_REQUIRE(h != nullptr && (policy == D2S_POLICY_LATENCY || policy == D2S_POLICY_THROUGHPUT), "d2s_set_policy: bad arguments");
    std::unique_lock<std::mutex> lock(h->mu);
    h->policy = policy;
    return D2S_OK;
}

extern "C" int d2s_reset_stream(d2s_handle h, d2s_stream_t stream) {
    D2S_REQUIRE(h != nullptr, "d2s_reset_stream: null handle");
    std::unique_lock<std::mutex> lock(h->mu);
    for (auto &kv : h->states)
        if (kv.first[0] == (long long)(uintptr_t)stream && kv.second->frame_counter)
            D2S_CHECK_CUDA(cudaMemsetAsync(kv.second->frame_counter, 0, sizeof(long long), (cudaStream_t)stream));
    return D2S_OK;
}

extern "C" int d2s_release_stream(d2s_handle h, d2s_stream_t stream) {
    D2S_REQUIRE(h != nullptr, "d2s_release_stream: null handle");
    std::unique_lock<std::mutex> lock(h->mu);
    D2S_CHECK_CUDA(cudaDeviceSynchronize());     // (the stream handle itself may already be destroyed: do not synchronise on it)
    for (auto it = h->plans.begin(); it != h->plans.end();) {
        if (it->second->stream == (cudaStream_t)stream) {
            if (h->last == it->second.get()) h->last = nullptr;
            it = h->plans.erase(it);
        } else ++it;
    }
    for (auto it = h->states.begin(); it != h->states.end();)
        it = it->first[0] == (long long)(uintptr_t)stream ? h->states.erase(it) : std::next(it);
    return D2S_OK;
}

extern "C" int d2s_debug_tap(d2s_handle h, const char *name, float *dst, size_t max_elems, size_t *n_elems, d2s_stream_t stream) {
    D2S_REQUIRE(h && name && n_elems, "d2s_debug_tap: null argument");
    std::unique_lock<std::mutex> lock(h->mu);
    D2S_REQUIRE(h->last, "d2s_debug_tap: no inference has run yet");
    auto it = h->last->taps.find(name);
    if (it == h->last->taps.end()) return set_error(D2S_ERR_INVALID, "d2s_debug_tap: unknown tap '%s'", name);
    *n_elems = it->second.n;
    if (!dst) return D2S_OK;
    D2S_REQUIRE(max_elems >= it->second.n, "d2s_debug_tap: buffer too small (%zu < %zu)", max_elems, it->second.n);
    D2S_LAUNCH(tap_to_f32_kernel, ceil_div((long long)it->second.n, 256), 256, 0, stream, it->second.ptr, it->second.dtype, dst, it->second.n);
    D2S_POST_LAUNCH();
    return D2S_OK;
}

extern "C" size_t d2s_workspace_bytes(d2s_handle h) {
    if (!h) return 0;
    size_t t = h->weight_bytes;
    std::unique_lock<std::mutex> lock(h->mu);
    for (auto &kv : h->plans) t += kv.second->total_bytes;
    for (auto &kv : h->states) t += kv.second->bytes;
    return t;
}
