// Internals of the depth engine shared by engine.cu (plan construction, d2s_infer) and pipe.cu (the whole-frame pipeline):
// the per-shape plan, the temporal stream state and the engine object behind d2s_handle.
#pragma once
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "gemm.cuh"
#include "layers.cuh"

namespace d2s {

static int round_up(int x, int a) { return (x + a - 1) / a * a; }

struct Tap { const void *ptr; size_t n; int dtype; };

// Temporal engines: what VideoDepthAnything.forward keeps between calls (vda2_s.py:189-224) — the frame counter and one K'|V'
// ring per temporal attention block.  It belongs to a VIDEO, which the caller identifies by the CUDA stream it submits the
// video's frames on (and the input size): every plan bound to that stream — latency or throughput policy, any dtype — shares
// it, so switching policy mid-video does not restart the window.  d2s_reset_stream zeroes the counter, d2s_release_stream frees it.
struct StreamState {
    long long *frame_counter = nullptr;
    std::vector<__half *> rings;
    std::vector<size_t> ring_elems;
    size_t bytes = 0;
    ~StreamState() {
        if (frame_counter) cudaFree(frame_counter);
        for (__half *r : rings) cudaFree(r);
    }
};

struct ShapePlan {
    int B, H, W, in_dtype, out_dtype;
    cudaStream_t stream = nullptr;         // the stream this plan is bound to (temporal engines keep per-stream state)
    int policy = 0;                        // D2S_POLICY_* the plan was built under
    long long *frame_counter = nullptr;    // temporal: frames seen on this stream (device; owned by `state`)
    StreamState *state = nullptr;          // temporal: shared by every plan of this (stream, H, W)
    size_t ring_cursor = 0;                // next ring of `state` this plan's construction will bind
    unsigned long long last_use = 0;       // LRU stamp (engine tick)
    std::vector<void *> allocs;
    std::vector<std::function<int(cudaStream_t)>> ops;
    std::deque<GemmPlan> gemms;            // stable addresses: the op lambdas hold pointers into it
    std::map<std::string, Tap> taps;
    void *in_stage = nullptr, *out_stage = nullptr;
    size_t in_bytes = 0, out_bytes = 0, total_bytes = 0;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    ~ShapePlan() {
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        for (void *p : allocs) cudaFree(p);
    }
};

struct LayerW {
    float *ln1_w, *ln1_b, *qkv_b, *proj_b, *ln2_w, *ln2_b, *fc1_b, *fc2_b;
    __half *qkv_w, *proj_w, *fc1_w, *fc2_w;
};
struct RcuW { __half *c1_w, *c2_w; float *c1_b, *c2_b; };
struct FusionW { __half *proj_w; float *proj_b; RcuW rl1, rl2; };
struct TAttnW { float *ln_w, *ln_b, *pe, *out_b; __half *qkv_w, *out_w; };          // pe: [32, 3C] = pe @ {q,k,v}^T
struct TemporalW {                                                                   // one TemporalModule (motion_module.py:31-134)
    int C;
    float *gn_w, *gn_b, *in_b, *ffln_w, *ffln_b, *ff1_b, *ff2_b, *out_b;
    __half *in_w, *ff1_w, *ff2_w, *out_w;
    TAttnW att[2];
};

}  // namespace d2s

struct d2s_engine {
    d2s_model_config cfg;
    int device;
    int D, L, P14, Kpatch;            // Kpatch: 588 padded to 640
    int c[4], cp[4], F, Fh, Fhp;      // neck channels (+ padded), fusion width, head width (+ padded)
    std::vector<void *> allocs;
    size_t weight_bytes = 0;
    // weights
    __half *patch_w; float *patch_b, *cls, *pos_table;
    std::vector<d2s::LayerW> layers;
    float *norm_w, *norm_b;
    __half *re_proj_w[4]; float *re_proj_b[4];
    __half *up0_w, *up1_w, *dn3_w; float *up0_b, *up1_b, *dn3_b;   // up biases are expanded to f*f*C
    __half *neck_w[4];
    d2s::FusionW fus[4];
    __half *head_c1_w, *head_c2_w; float *head_c1_b, *head_c2_b, *head_c3_w; float head_c3_b;
    d2s::TemporalW tm[4];                                              // cfg.temporal only
    std::map<std::vector<long long>, std::unique_ptr<d2s::ShapePlan>> plans;   // keyed by shape, dtypes, stream AND policy; LRU-bounded
    std::map<std::vector<long long>, std::unique_ptr<d2s::StreamState>> states; // temporal: keyed by (stream, H, W)
    unsigned long long tick = 0;
    size_t max_plans = 64;               // D2S_MAX_PLANS; the least recently used plan is dropped beyond it
    std::mutex mu;
    d2s::ShapePlan *last = nullptr;
    bool use_graph = true;
    int policy = 0;                      // D2S_POLICY_LATENCY / D2S_POLICY_THROUGHPUT for the plans built next
};

namespace d2s {
// find or build (host-synchronous on first use) the plan for this shape / dtypes / stream under the engine's current policy
int engine_plan(d2s_engine *h, int B, int H, int W, int in_dtype, int out_dtype, cudaStream_t st, ShapePlan **out);
// enqueue the plan's kernels (its CUDA graph when built with one) on st; input is read from sp->in_stage, depth lands in sp->out_stage
int engine_run_plan(ShapePlan *sp, cudaStream_t st);
// enqueue the plan's kernels one by one (used while capturing a larger graph)
int engine_run_ops(ShapePlan *sp, cudaStream_t st);
}  // namespace d2s
