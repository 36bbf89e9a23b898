// The whole-frame pipeline: capture frame in -> packed stereo frame out, `slots` frames in flight, ONE C call per frame.
//
// This is the caller side of the hot path.  The reference runs capture -> process -> predict_depth -> make_sbs as
// three Python threads joined by two size-1 queues (reference main.py:67-68, 232-262, 1336-1341): a 3-deep software pipeline in
// which every frame costs ~15 framework calls.  Here each in-flight frame owns a slot: a CUDA stream, fixed device buffers (so
// nothing is allocated per frame), pinned host buffers, the engine plan of that stream, and two CUDA graphs that replay the
// frame's kernels:
//     [H2D copy] -> process kernel -> graph A { resize+normalise -> network (~140 kernels) -> percentile bounds, point ops, blur-x }
//                -> wait(previous frame's EMA) -> blur-y + DepthStabilizer EMA -> signal EMA
//                -> graph B { depth upsample + stereo warp/pack }  -> [D2H copy] -> done
// The EMA of DepthStabilizer (depth.py:1865-1887) is the only cross-frame dependency of the path; it is the single kernel between
// the two graphs, ordered frame-to-frame with an event, so the networks of consecutive frames overlap freely.  The antialias
// weight tables of both resizes are input-independent and are built once at creation (the per-call entry points rebuild them).
// Arithmetic is that of d2s_process / d2s_preprocess / d2s_infer / d2s_postprocess / d2s_make_sbs — the same kernels, launched
// through the same code — so a frame through the pipe equals the same frame through the five calls, bit for bit.
#include <string.h>

#include <algorithm>
#include <cstddef>
#include <vector>

#include "engine.cuh"
#include "prepost.cuh"

#define TRY_RC(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

namespace d2s {

struct PipeSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr, ema_ev = nullptr, in_ev = nullptr, t[4] = {nullptr, nullptr, nullptr, nullptr};
    uint8_t *d_frame = nullptr;      // captured frame [h0,w0,ch] u8
    void *d_rgb = nullptr;           // process() output [3,h,w] rgb_dtype
    void *ws_proc = nullptr, *ws_pre = nullptr, *ws_post = nullptr;
    void *d_depth = nullptr;         // [h,w] fp16: predict_depth's return value
    void *d_out = nullptr;           // packed stereo frame
    void *d_enc = nullptr;           // cfg.out_format != PACKED: what the caller receives (NV12 frame / d2s_pipe_jpeg_frame per stream)
    void *ws_jpeg = nullptr;         // D2S_OUT_JPEG: one d2s_jpeg_encode workspace per stream
    size_t copied = 0;               // D2S_OUT_JPEG with host_io: bytes per stream the submit copied to the host
    void *h_in = nullptr, *h_out = nullptr;
    ShapePlan *plan = nullptr;
    cudaGraphExec_t gA = nullptr, gB = nullptr;
    cudaGraph_t graphA = nullptr, graphB = nullptr;
    long long kernels_a = 0, kernels_b = 0;
    bool busy = false, traced = false;
    // streams > 1: the per-stream resize / post-process / warp chains of one step are independent: they are captured as parallel
    // branches of the graphs (fork / join through these), so e.g. the eight one-block percentile selections run side by side
    std::vector<cudaStream_t> aux;
    std::vector<cudaEvent_t> join_ev;
    cudaEvent_t fork_ev = nullptr;
};

}  // namespace d2s

struct d2s_pipe {
    d2s_pipe_config cfg;
    d2s_engine *engine;
    int device;
    int B;             // video streams per slot = frames per submit (cfg.streams)
    int h, w;          // size of process()'s output (== frame size unless target_height < frame_h)
    int Hm, Wm;        // model input
    int oh, ow;        // packed stereo frame
    size_t frame_bytes, out_bytes, res_bytes, rgb_es, out_es;   // out: the packed RGB frame, res: what the caller receives (per stream)
    size_t ws_proc_bytes, ws_pre_bytes, ws_post_bytes;
    size_t ws_jpeg_bytes = 0, jpeg_head = 0;  // D2S_OUT_JPEG: workspace per stream; bytes per stream the next submit copies to the host
    int jpeg_quality = 90, jpeg_ri = 4;
    char fps_text[40] = {0};                 // cfg.fps_overlay: drawn onto the RGB frame between the two graphs
    void *ema_state = nullptr;               // [Hm,Wm] fp16, NaN = unset (d2s_post_params.ema_valid == 2)
    cudaEvent_t last_ema = nullptr;          // EMA event of the most recently submitted frame
    bool trace = false;
    std::vector<d2s::PipeSlot> slots;
};

namespace d2s {

int rgb_to_nv12_launch(const uint8_t *rgb, long long pitch, int h, int w, uint8_t *yp, uint8_t *uvp, cudaStream_t stream);   // nv12.cu
int jpeg_encode_launch(const uint8_t *rgb, long long pitch, int h, int w, int quality, int ri, uint8_t *out, size_t capacity, uint32_t *size_out,
                       void *workspace, size_t workspace_bytes, cudaStream_t stream);                                          // jpeg.cu
size_t jpeg_workspace(int h, int w, int ri);

__global__ void pipe_join_kernel(int) {}      // a single kernel predecessor for whatever follows a join (programmatic launch edges need one)

static size_t dtype_size(int dt) { return dt == D2S_F32 ? 4 : (dt == D2S_U8 ? 1 : 2); }

// b: index of the video stream inside the slot's batch (cfg.streams frames go through the network per submit, one per stream)
static void fill_post(const d2s_pipe *p, const PipeSlot &s, int b, d2s_post_params *pp) {
    const d2s_pipe_config &c = p->cfg;
    *pp = d2s_post_params{};
    pp->depth_in = (const __half *)s.plan->out_stage + (size_t)b * p->Hm * p->Wm; pp->in_dtype = D2S_F16; pp->H = p->Hm; pp->W = p->Wm;
    pp->out = (__half *)s.d_depth + (size_t)b * p->h * p->w; pp->out_dtype = D2S_F16; pp->out_h = p->h; pp->out_w = p->w;
    pp->compute_dtype = D2S_F16;          // the reference's CUDA path post-processes the autocast (fp16) depth in fp16
    pp->metric = c.metric; pp->percentile = c.percentile; pp->subsample_cap = c.subsample_cap; pp->gamma = c.gamma;
    pp->foreground_scale = c.foreground_scale; pp->aa_strength = c.aa_strength;
    pp->ema_state = c.use_temporal_smooth ? (void *)((__half *)p->ema_state + (size_t)b * p->Hm * p->Wm) : nullptr; pp->ema_valid = 2; pp->ema_alpha = c.ema_alpha;
    pp->workspace = (char *)s.ws_post + (size_t)b * p->ws_post_bytes; pp->workspace_bytes = p->ws_post_bytes;
}

static void *rgb_of(const d2s_pipe *p, const PipeSlot &s, int b) { return (char *)s.d_rgb + (size_t)b * 3 * p->h * p->w * p->rgb_es; }
static const uint8_t *frame_of(const d2s_pipe *p, const uint8_t *base, int b) { return base + (size_t)b * p->frame_bytes; }

static void fill_warp(const d2s_pipe *p, const PipeSlot &s, int b, d2s_warp_params *wp) {
    const d2s_pipe_config &c = p->cfg;
    *wp = d2s_warp_params{};
    wp->rgb.base = rgb_of(p, s, b); wp->rgb.dtype = c.rgb_dtype; wp->rgb.sc = (int64_t)p->h * p->w; wp->rgb.sy = p->w; wp->rgb.sx = 1;
    wp->out.base = (char *)s.d_out + (size_t)b * p->out_bytes; wp->out.dtype = c.out_dtype; wp->out.sc = 1; wp->out.sy = (int64_t)3 * p->ow; wp->out.sx = 3;
    wp->depth = (__half *)s.d_depth + (size_t)b * p->h * p->w; wp->depth_dtype = D2S_F16; wp->depth_h = p->h; wp->depth_w = p->w; wp->h = p->h; wp->w = p->w;
    wp->ipd_uv = c.ipd_uv; wp->depth_ratio = c.depth_ratio; wp->convergence = c.convergence;
    wp->display_mode = c.display_mode; wp->fill_16_9 = c.fill_16_9; wp->warp_mode = D2S_WARP_BILINEAR;
    wp->rgb_round_to_depth_dtype = c.rgb_dtype != D2S_F16;   // make_sbs casts rgb to depth.dtype (depth.py:2209-2215)
}

static int capture_begin(cudaStream_t st) {
    D2S_CHECK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    return D2S_OK;
}
static int capture_end(cudaStream_t st, int rc, cudaGraph_t *g, cudaGraphExec_t *exec) {
    cudaError_t ee = cudaStreamEndCapture(st, g);
    if (rc) return rc;
    if (ee != cudaSuccess) return set_error(D2S_ERR_CUDA, "d2s_pipe_create: graph capture failed: %s", cudaGetErrorString(ee));
    D2S_CHECK_CUDA(cudaGraphInstantiate(exec, *g, 0));
    return D2S_OK;
}

static int build_slot(d2s_pipe *p, PipeSlot &s) {
    const d2s_pipe_config &c = p->cfg;
    D2S_CHECK_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    D2S_CHECK_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    D2S_CHECK_CUDA(cudaEventCreateWithFlags(&s.ema_ev, cudaEventDisableTiming));
    D2S_CHECK_CUDA(cudaEventCreateWithFlags(&s.in_ev, cudaEventDisableTiming));
    for (auto &e : s.t) D2S_CHECK_CUDA(cudaEventCreate(&e));
    const int B = p->B;
    D2S_CHECK_CUDA(cudaMalloc((void **)&s.d_frame, B * p->frame_bytes));
    D2S_CHECK_CUDA(cudaMalloc(&s.d_rgb, (size_t)B * 3 * p->h * p->w * p->rgb_es));
    if (p->ws_proc_bytes) D2S_CHECK_CUDA(cudaMalloc(&s.ws_proc, p->ws_proc_bytes));
    D2S_CHECK_CUDA(cudaMalloc(&s.ws_pre, B * p->ws_pre_bytes));       // (one per stream: the resizes of one step run concurrently)
    if (B > 1) {
        s.aux.resize(B); s.join_ev.resize(B);
        for (int b = 0; b < B; ++b) {
            D2S_CHECK_CUDA(cudaStreamCreateWithFlags(&s.aux[b], cudaStreamNonBlocking));
            D2S_CHECK_CUDA(cudaEventCreateWithFlags(&s.join_ev[b], cudaEventDisableTiming));
        }
        D2S_CHECK_CUDA(cudaEventCreateWithFlags(&s.fork_ev, cudaEventDisableTiming));
    }
    D2S_CHECK_CUDA(cudaMalloc(&s.ws_post, B * p->ws_post_bytes));     // (one per stream: its blur-x result waits there for the EMA step)
    D2S_CHECK_CUDA(cudaMalloc(&s.d_depth, (size_t)B * p->h * p->w * 2));
    D2S_CHECK_CUDA(cudaMalloc(&s.d_out, B * p->out_bytes));
    if (c.out_format != D2S_OUT_PACKED) D2S_CHECK_CUDA(cudaMalloc(&s.d_enc, B * p->res_bytes));
    if (c.out_format == D2S_OUT_JPEG) {
        D2S_CHECK_CUDA(cudaMalloc(&s.ws_jpeg, B * p->ws_jpeg_bytes));
        D2S_CHECK_CUDA(cudaMemsetAsync(s.d_enc, 0, B * p->res_bytes, s.stream));
    }
    if (c.host_io) {
        D2S_CHECK_CUDA(cudaHostAlloc(&s.h_in, B * p->frame_bytes, cudaHostAllocDefault));
        D2S_CHECK_CUDA(cudaHostAlloc(&s.h_out, B * p->res_bytes, cudaHostAllocDefault));
    }
    TRY_RC(engine_plan(p->engine, B, p->Hm, p->Wm, D2S_F16, D2S_F16, s.stream, &s.plan));

    // input-independent tables of the two antialias resizes: once, here (the workspaces are shared by the slot's streams: the
    // kernels of one slot run in stream order)
    auto rgb_view = [&](int b) {
        d2s_image v{};
        v.base = rgb_of(p, s, b); v.dtype = c.rgb_dtype; v.sc = (int64_t)p->h * p->w; v.sy = p->w; v.sx = 1;
        return v;
    };
    auto model_in = [&](int b) { return (void *)((__half *)s.plan->in_stage + (size_t)b * 3 * p->Hm * p->Wm); };
    auto ws_pre_of = [&](int b) { return (void *)((char *)s.ws_pre + (size_t)b * p->ws_pre_bytes); };
    TRY_RC(process_phases(s.d_frame, c.frame_h, c.frame_w, c.channels, s.d_rgb, c.rgb_dtype, p->h, p->w, s.ws_proc, p->ws_proc_bytes, 1, s.stream));
    for (int b = 0; b < B; ++b) {
        d2s_image src = rgb_view(b);
        TRY_RC(preprocess_phases(&src, p->h, p->w, model_in(b), D2S_F16, p->Hm, p->Wm, c.mean, c.std, ws_pre_of(b), p->ws_pre_bytes, 1, s.stream));
    }
    D2S_CHECK_CUDA(cudaStreamSynchronize(s.stream));

    const bool split = c.use_temporal_smooth != 0 || c.fps_overlay != 0;   // something happens between the network and the warp
    auto warp_and_pack = [&](int b, cudaStream_t st) -> int {      // stereo warp (+ the output encoder: NV12 stages, or the whole JPEG)
        d2s_warp_params wp; fill_warp(p, s, b, &wp);
        int r = d2s_make_sbs(&wp, st);
        const uint8_t *packed = (const uint8_t *)s.d_out + (size_t)b * p->out_bytes;
        uint8_t *enc = (uint8_t *)s.d_enc + (size_t)b * p->res_bytes;
        if (!r && c.out_format == D2S_OUT_NV12) r = rgb_to_nv12_launch(packed, (long long)p->ow * 3, p->oh, p->ow, enc, enc + (size_t)p->oh * p->ow, st);
        if (!r && c.out_format == D2S_OUT_JPEG)
            r = jpeg_encode_launch(packed, (long long)p->ow * 3, p->oh, p->ow, p->jpeg_quality, p->jpeg_ri, enc + offsetof(d2s_pipe_jpeg_frame, data),
                                   p->res_bytes - offsetof(d2s_pipe_jpeg_frame, data), (uint32_t *)enc, (char *)s.ws_jpeg + (size_t)b * p->ws_jpeg_bytes,
                                   p->ws_jpeg_bytes, st);
        return r;
    };
    // run body(b, stream) for every stream of the step: in line for one stream, as parallel branches of the capture otherwise
    auto for_each_stream = [&](auto body) -> int {
        if (B == 1) return body(0, s.stream);
        D2S_CHECK_CUDA(cudaEventRecord(s.fork_ev, s.stream));
        int r = D2S_OK;
        for (int b = 0; b < B; ++b) {
            D2S_CHECK_CUDA(cudaStreamWaitEvent(s.aux[b], s.fork_ev, 0));
            if (!r) r = body(b, s.aux[b]);
            D2S_CHECK_CUDA(cudaEventRecord(s.join_ev[b], s.aux[b]));
            D2S_CHECK_CUDA(cudaStreamWaitEvent(s.stream, s.join_ev[b], 0));
        }
        D2S_LAUNCH(pipe_join_kernel, 1, 1, 0, s.stream, 0);
        return r;
    };
    // graph A
    long long k0 = g_launch_count.load();
    TRY_RC(capture_begin(s.stream));
    int rc = for_each_stream([&](int b, cudaStream_t st) {
        d2s_image src = rgb_view(b);
        return preprocess_phases(&src, p->h, p->w, model_in(b), D2S_F16, p->Hm, p->Wm, c.mean, c.std, ws_pre_of(b), p->ws_pre_bytes, 2, st);
    });
    if (!rc) rc = engine_run_ops(s.plan, s.stream);
    if (!rc) rc = for_each_stream([&](int b, cudaStream_t st) {
        d2s_post_params pp; fill_post(p, s, b, &pp);
        int r = postprocess_phases(&pp, split ? POST_PHASE_HEAD : POST_PHASE_ALL, st);
        if (!r && !split) r = warp_and_pack(b, st);
        return r;
    });
    TRY_RC(capture_end(s.stream, rc, &s.graphA, &s.gA));
    s.kernels_a = g_launch_count.load() - k0;
    if (split) {
        k0 = g_launch_count.load();
        TRY_RC(capture_begin(s.stream));
        rc = for_each_stream([&](int b, cudaStream_t st) {
            d2s_post_params pp; fill_post(p, s, b, &pp);
            int r = postprocess_phases(&pp, POST_PHASE_UP, st);
            if (!r) r = warp_and_pack(b, st);
            return r;
        });
        TRY_RC(capture_end(s.stream, rc, &s.graphB, &s.gB));
        s.kernels_b = g_launch_count.load() - k0;
    }
    return D2S_OK;
}

static void free_slot(PipeSlot &s) {
    if (s.gA) cudaGraphExecDestroy(s.gA);
    if (s.gB) cudaGraphExecDestroy(s.gB);
    if (s.graphA) cudaGraphDestroy(s.graphA);
    if (s.graphB) cudaGraphDestroy(s.graphB);
    for (void *q : {(void *)s.d_frame, s.d_rgb, s.ws_proc, s.ws_pre, s.ws_post, s.d_depth, s.d_out, s.d_enc, s.ws_jpeg}) if (q) cudaFree(q);
    if (s.h_in) cudaFreeHost(s.h_in);
    if (s.h_out) cudaFreeHost(s.h_out);
    for (cudaEvent_t e : {s.done, s.ema_ev, s.in_ev, s.fork_ev, s.t[0], s.t[1], s.t[2], s.t[3]}) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : s.join_ev) if (e) cudaEventDestroy(e);
    for (cudaStream_t a : s.aux) if (a) cudaStreamDestroy(a);
    if (s.stream) cudaStreamDestroy(s.stream);
}

}  // namespace d2s

using namespace d2s;

extern "C" int d2s_pipe_create(d2s_handle engine, const d2s_pipe_config *cfg, d2s_pipe_handle *out) {
    D2S_REQUIRE(engine && cfg && out, "d2s_pipe_create: null argument");
    D2S_REQUIRE(cfg->frame_h > 0 && cfg->frame_w > 1 && (cfg->channels == 3 || cfg->channels == 4), "d2s_pipe_create: frame %dx%dx%d", cfg->frame_h, cfg->frame_w, cfg->channels);
    D2S_REQUIRE(cfg->slots >= 1 && cfg->slots <= 64, "d2s_pipe_create: slots=%d (1..64)", cfg->slots);
    D2S_REQUIRE(cfg->rgb_dtype == D2S_F16 || cfg->rgb_dtype == D2S_F32, "d2s_pipe_create: rgb_dtype %d (F16 or F32)", cfg->rgb_dtype);
    D2S_REQUIRE(cfg->out_dtype == D2S_F32 || cfg->out_dtype == D2S_U8 || cfg->out_dtype == D2S_F16, "d2s_pipe_create: out_dtype %d", cfg->out_dtype);
    D2S_REQUIRE(cfg->display_mode >= 0 && cfg->display_mode <= 3, "d2s_pipe_create: display_mode %d", cfg->display_mode);
    D2S_REQUIRE(cfg->depth_resolution > 0 && cfg->patch == engine->cfg.patch, "d2s_pipe_create: depth_resolution %d / patch %d", cfg->depth_resolution, cfg->patch);
    // a temporal engine keeps ONE video's window per stream: frames of that video cannot be spread over several slots
    D2S_REQUIRE(!engine->cfg.temporal || cfg->slots == 1, "d2s_pipe_create: a temporal (Video-Depth-Anything) engine needs slots == 1 (its frames are sequential; run one pipe per video)");
    D2S_REQUIRE(cfg->streams >= 0 && cfg->streams <= 64 && (!engine->cfg.temporal || cfg->streams <= 1), "d2s_pipe_create: streams=%d (1..64; 1 for a temporal engine)", cfg->streams);
    D2S_CHECK_CUDA(cudaSetDevice(engine->device));
    d2s_pipe *p = new d2s_pipe();
    p->cfg = *cfg; p->engine = engine; p->device = engine->device;
    p->B = cfg->streams > 0 ? cfg->streams : 1;
    if (cfg->target_height >= cfg->frame_h) { p->h = cfg->frame_h; p->w = cfg->frame_w; }
    else {   // depth.py:555-559
        p->h = (cfg->target_height / 2) * 2;
        p->w = ((int)((double)cfg->frame_w * (double)cfg->target_height / (double)cfg->frame_h) / 2) * 2;
    }
    int rc = d2s_model_input_shape(p->h, p->w, cfg->depth_resolution, cfg->patch, &p->Hm, &p->Wm);
    if (!rc) rc = d2s_sbs_out_shape(p->h, p->w, cfg->display_mode, cfg->fill_16_9, &p->oh, &p->ow);
    if (rc || p->h < 1 || p->w < 2) { delete p; return rc ? rc : set_error(D2S_ERR_INVALID, "d2s_pipe_create: processed size %dx%d", p->h, p->w); }
    p->rgb_es = dtype_size(cfg->rgb_dtype); p->out_es = dtype_size(cfg->out_dtype);
    p->frame_bytes = (size_t)cfg->frame_h * cfg->frame_w * cfg->channels;
    p->out_bytes = (size_t)p->oh * p->ow * 3 * p->out_es;
    p->res_bytes = cfg->out_format == D2S_OUT_PACKED ? p->out_bytes : (size_t)p->oh * p->ow * 3 / 2;
    if (cfg->out_format < D2S_OUT_PACKED || cfg->out_format > D2S_OUT_JPEG ||
        (cfg->out_format != D2S_OUT_PACKED && (cfg->out_dtype != D2S_U8 || p->oh % 2 || p->ow % 2))) {
        rc = set_error(D2S_ERR_INVALID, "d2s_pipe_create: out_format %d needs out_dtype U8 and an even-sized packed frame (%dx%d)", cfg->out_format, p->oh, p->ow);
        delete p;
        return rc;
    }
    if (cfg->out_format == D2S_OUT_JPEG) {
        p->jpeg_quality = cfg->jpeg_quality > 0 ? cfg->jpeg_quality : 90;
        p->jpeg_ri = cfg->jpeg_restart_interval > 0 ? cfg->jpeg_restart_interval : 4;
        p->res_bytes = ((size_t)p->oh * p->ow * 3 + offsetof(d2s_pipe_jpeg_frame, data) + 4096 + 255) & ~(size_t)255;   // the raw frame's size: above quality-100 noise (2 B/px)
        if (d2s_jpeg_workspace_bytes(p->oh, p->ow, p->jpeg_ri) == 0) { delete p; return D2S_ERR_INVALID; }
        p->ws_jpeg_bytes = (jpeg_workspace(p->oh, p->ow, p->jpeg_ri) + 255) & ~(size_t)255;
        p->jpeg_head = std::min(p->res_bytes, std::max((size_t)65536, ((size_t)p->oh * p->ow / 4 + 65535) & ~(size_t)65535));
    }
    p->ws_proc_bytes = process_workspace_bytes(cfg->frame_h, cfg->frame_w, p->h, p->w);
    p->ws_pre_bytes = d2s_preprocess_workspace_bytes(p->h, p->w, p->Hm, p->Wm);
    p->ws_post_bytes = d2s_postprocess_workspace_bytes(p->Hm, p->Wm);
    rc = [&]() -> int {
        D2S_CHECK_CUDA(cudaMalloc(&p->ema_state, (size_t)p->B * p->Hm * p->Wm * 2));      // one DepthStabilizer state per stream
        D2S_CHECK_CUDA(cudaMemset(p->ema_state, 0xFF, (size_t)p->B * p->Hm * p->Wm * 2));
        // several frames in flight share the GPU: the throughput tile policy (d2s_set_policy); one frame alone: latency
        const int old_policy = engine->policy;
        { std::unique_lock<std::mutex> lock(engine->mu); engine->policy = cfg->slots > 1 ? D2S_POLICY_THROUGHPUT : D2S_POLICY_LATENCY; }
        p->slots.resize(cfg->slots);
        int r = D2S_OK;
        for (auto &s : p->slots)
            if ((r = build_slot(p, s))) break;
        { std::unique_lock<std::mutex> lock(engine->mu); engine->policy = old_policy; }
        return r;
    }();
    if (rc) { d2s_pipe_destroy(p); return rc; }
    *out = p;
    return D2S_OK;
}

extern "C" int d2s_pipe_destroy(d2s_pipe_handle p) {
    if (!p) return D2S_OK;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    for (auto &s : p->slots) {
        if (s.stream) d2s_release_stream(p->engine, s.stream);
        free_slot(s);
    }
    if (p->ema_state) cudaFree(p->ema_state);
    delete p;
    return D2S_OK;
}

extern "C" int d2s_pipe_geometry(d2s_pipe_handle p, int *h, int *w, int *model_h, int *model_w, int *out_h, int *out_w, size_t *frame_bytes, size_t *out_bytes) {
    D2S_REQUIRE(p != nullptr, "d2s_pipe_geometry: null pipe");
    if (h) *h = p->h; if (w) *w = p->w; if (model_h) *model_h = p->Hm; if (model_w) *model_w = p->Wm;
    if (out_h) *out_h = p->oh; if (out_w) *out_w = p->ow; if (frame_bytes) *frame_bytes = p->frame_bytes; if (out_bytes) *out_bytes = p->res_bytes;
    return D2S_OK;
}

extern "C" int d2s_pipe_slot_buffers(d2s_pipe_handle p, int slot, void **host_in, void **host_out, void **dev_in, void **dev_out, void **dev_depth,
                                     d2s_stream_t *stream) {
    D2S_REQUIRE(p && slot >= 0 && slot < (int)p->slots.size(), "d2s_pipe_slot_buffers: bad slot");
    const PipeSlot &s = p->slots[slot];
    if (host_in) *host_in = s.h_in; if (host_out) *host_out = s.h_out; if (dev_in) *dev_in = s.d_frame; if (dev_out) *dev_out = s.d_enc ? s.d_enc : s.d_out;
    if (dev_depth) *dev_depth = s.d_depth; if (stream) *stream = s.stream;
    return D2S_OK;
}

extern "C" int d2s_pipe_submit(d2s_pipe_handle p, int slot, const void *frame, d2s_stream_t frame_ready_on) {
    D2S_REQUIRE(p && slot >= 0 && slot < (int)p->slots.size(), "d2s_pipe_submit: bad slot");
    PipeSlot &s = p->slots[slot];
    D2S_REQUIRE(!s.busy, "d2s_pipe_submit: slot %d still holds an uncollected frame (d2s_pipe_wait first)", slot);
    const d2s_pipe_config &c = p->cfg;
    cudaStream_t st = s.stream;
    if (frame && !c.host_io && frame_ready_on != (d2s_stream_t)st) {   // a device frame produced on another stream
        D2S_CHECK_CUDA(cudaEventRecord(s.in_ev, (cudaStream_t)frame_ready_on));
        D2S_CHECK_CUDA(cudaStreamWaitEvent(st, s.in_ev, 0));
    }
    s.traced = p->trace;
    if (s.traced) D2S_CHECK_CUDA(cudaEventRecord(s.t[0], st));
    const uint8_t *src;
    if (c.host_io) {
        D2S_CHECK_CUDA(cudaMemcpyAsync(s.d_frame, frame ? frame : s.h_in, p->B * p->frame_bytes, cudaMemcpyHostToDevice, st));
        src = s.d_frame;
    } else src = frame ? (const uint8_t *)frame : s.d_frame;
    int rc = D2S_OK;
    for (int b = 0; b < p->B; ++b)
        if ((rc = process_phases(frame_of(p, src, b), c.frame_h, c.frame_w, c.channels, rgb_of(p, s, b), c.rgb_dtype, p->h, p->w, s.ws_proc, p->ws_proc_bytes, 2, st))) return rc;
    if (s.traced) D2S_CHECK_CUDA(cudaEventRecord(s.t[1], st));
    D2S_CHECK_CUDA(cudaGraphLaunch(s.gA, st));
    long long kernels = s.kernels_a;
    if (s.gB) {
        if (c.use_temporal_smooth && p->last_ema && p->last_ema != s.ema_ev) D2S_CHECK_CUDA(cudaStreamWaitEvent(st, p->last_ema, 0));
        for (int b = 0; b < p->B; ++b) {
            d2s_post_params pp; fill_post(p, s, b, &pp);
            if ((rc = postprocess_phases(&pp, POST_PHASE_EMA, st))) return rc;
        }
        if (c.use_temporal_smooth) {
            D2S_CHECK_CUDA(cudaEventRecord(s.ema_ev, st));
            p->last_ema = s.ema_ev;
        }
        if (p->fps_text[0])                                   // make_sbs(fps=...): overlay_fps on the frame the warp reads (depth.py:2226-2227)
            for (int b = 0; b < p->B; ++b) {
                d2s_image img{};
                img.base = rgb_of(p, s, b); img.dtype = c.rgb_dtype; img.sc = (int64_t)p->h * p->w; img.sy = p->w; img.sx = 1;
                if ((rc = d2s_overlay_fps(&img, p->h, p->w, p->fps_text, (d2s_stream_t)st))) return rc;
            }
        if (s.traced) D2S_CHECK_CUDA(cudaEventRecord(s.t[2], st));
        D2S_CHECK_CUDA(cudaGraphLaunch(s.gB, st));
        kernels += s.kernels_b;
    }
    g_launch_count.fetch_add(kernels, std::memory_order_relaxed);
    if (s.traced) D2S_CHECK_CUDA(cudaEventRecord(s.t[3], st));
    if (c.host_io && c.out_format == D2S_OUT_JPEG) {      // only as much of every stream's buffer as recent frames needed
        s.copied = p->jpeg_head;
        D2S_CHECK_CUDA(cudaMemcpy2DAsync(s.h_out, p->res_bytes, s.d_enc, p->res_bytes, s.copied, p->B, cudaMemcpyDeviceToHost, st));
    } else if (c.host_io) D2S_CHECK_CUDA(cudaMemcpyAsync(s.h_out, s.d_enc ? s.d_enc : s.d_out, p->B * p->res_bytes, cudaMemcpyDeviceToHost, st));
    D2S_CHECK_CUDA(cudaEventRecord(s.done, st));
    s.busy = true;
    return D2S_OK;
}

extern "C" int d2s_pipe_wait(d2s_pipe_handle p, int slot) {
    D2S_REQUIRE(p && slot >= 0 && slot < (int)p->slots.size(), "d2s_pipe_wait: bad slot");
    PipeSlot &s = p->slots[slot];
    D2S_REQUIRE(s.busy, "d2s_pipe_wait: slot %d has no frame in flight", slot);
    D2S_CHECK_CUDA(cudaEventSynchronize(s.done));
    s.busy = false;
    if (p->cfg.host_io && p->cfg.out_format == D2S_OUT_JPEG) {
        size_t need = 0;
        for (int b = 0; b < p->B; ++b) {
            const d2s_pipe_jpeg_frame *f = (const d2s_pipe_jpeg_frame *)((const char *)s.h_out + (size_t)b * p->res_bytes);
            D2S_REQUIRE(f->size != 0, "d2s_pipe_wait: the JPEG stream of stream %d did not fit %zu bytes", b, p->res_bytes);
            need = std::max(need, offsetof(d2s_pipe_jpeg_frame, data) + (size_t)f->size);
        }
        if (need > s.copied) {        // this frame outgrew the estimate: fetch the tails now
            D2S_CHECK_CUDA(cudaMemcpy2DAsync((char *)s.h_out + s.copied, p->res_bytes, (const char *)s.d_enc + s.copied, p->res_bytes, need - s.copied, p->B,
                                             cudaMemcpyDeviceToHost, s.stream));
            D2S_CHECK_CUDA(cudaStreamSynchronize(s.stream));
        }
        p->jpeg_head = std::min(p->res_bytes, std::max((size_t)65536, (need + need / 4 + 65535) & ~(size_t)65535));
    }
    return D2S_OK;
}

extern "C" int d2s_pipe_set_fps_text(d2s_pipe_handle p, const char *text) {
    D2S_REQUIRE(p != nullptr, "d2s_pipe_set_fps_text: null pipe");
    D2S_REQUIRE(p->cfg.fps_overlay, "d2s_pipe_set_fps_text: the pipe was created without fps_overlay");
    D2S_REQUIRE(!text || strlen(text) <= 32, "d2s_pipe_set_fps_text: text longer than 32 characters");
    memset(p->fps_text, 0, sizeof(p->fps_text));
    if (text) strcpy(p->fps_text, text);
    return D2S_OK;
}

extern "C" int d2s_pipe_set_trace(d2s_pipe_handle p, int on) {
    D2S_REQUIRE(p != nullptr, "d2s_pipe_set_trace: null pipe");
    p->trace = on != 0;
    return D2S_OK;
}

extern "C" int d2s_pipe_slot_times(d2s_pipe_handle p, int slot, float ms[3]) {
    D2S_REQUIRE(p && ms && slot >= 0 && slot < (int)p->slots.size(), "d2s_pipe_slot_times: bad arguments");
    PipeSlot &s = p->slots[slot];
    D2S_REQUIRE(s.traced && !s.busy, "d2s_pipe_slot_times: slot %d was not traced, or is still in flight", slot);
    D2S_CHECK_CUDA(cudaEventElapsedTime(&ms[0], s.t[0], s.t[1]));
    if (s.gB) {
        D2S_CHECK_CUDA(cudaEventElapsedTime(&ms[1], s.t[1], s.t[2]));
        D2S_CHECK_CUDA(cudaEventElapsedTime(&ms[2], s.t[2], s.t[3]));
    } else {
        D2S_CHECK_CUDA(cudaEventElapsedTime(&ms[1], s.t[1], s.t[3]));
        ms[2] = 0.f;
    }
    return D2S_OK;
}

extern "C" int d2s_pipe_reset(d2s_pipe_handle p) {
    D2S_REQUIRE(p != nullptr, "d2s_pipe_reset: null pipe");
    for (auto &s : p->slots) D2S_CHECK_CUDA(cudaStreamSynchronize(s.stream));
    D2S_CHECK_CUDA(cudaMemset(p->ema_state, 0xFF, (size_t)p->B * p->Hm * p->Wm * 2));
    p->last_ema = nullptr;
    if (p->engine->cfg.temporal)
        for (auto &s : p->slots) { int rc = d2s_reset_stream(p->engine, s.stream); if (rc) return rc; }
    return D2S_OK;
}
