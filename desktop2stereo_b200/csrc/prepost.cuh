// Internal entry points of prepost.cu for the whole-frame pipeline (pipe.cu): the same kernels as d2s_process / d2s_preprocess /
// d2s_postprocess, split so that input-independent work is done once per shape and the one cross-frame step (the EMA) can be
// ordered between frames that are in flight on different streams.
#pragma once
#include "common.cuh"

namespace d2s {

enum { POST_PHASE_HEAD = 1, POST_PHASE_EMA = 2, POST_PHASE_UP = 4, POST_PHASE_ALL = 7 };
int postprocess_phases(const d2s_post_params *p, int phases, d2s_stream_t stream);
// what: bit 0 = build the antialias weight tables in the workspace, bit 1 = run the filter passes (tables must be there)
int preprocess_phases(const d2s_image *src, int h, int w, void *dst, int dst_dtype, int new_h, int new_w, const float mean[3],
                      const float std[3], void *workspace, size_t workspace_bytes, int what, d2s_stream_t stream);
size_t process_workspace_bytes(int h0, int w0, int h, int w);
int process_phases(const uint8_t *frame, int h0, int w0, int channels, void *out, int out_dtype, int h, int w, void *ws, size_t ws_bytes,
                   int what, d2s_stream_t stream);

}  // namespace d2s
