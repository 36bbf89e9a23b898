// Entry points not implemented yet return D2S_ERR_UNSUPPORTED (never a CPU fallback).
#include "common.cuh"
using namespace d2s;
#define NOT_YET(name) return set_error(D2S_ERR_UNSUPPORTED, name ": not implemented yet")
extern "C" int d2s_create(const void *, size_t, const d2s_model_config *, int, d2s_handle *) { NOT_YET("d2s_create"); }
extern "C" int d2s_destroy(d2s_handle) { NOT_YET("d2s_destroy"); }
extern "C" int d2s_infer(d2s_handle, const void *, int, void *, int, int, int, int, d2s_stream_t) { NOT_YET("d2s_infer"); }
extern "C" int d2s_debug_tap(d2s_handle, const char *, float *, size_t, size_t *, d2s_stream_t) { NOT_YET("d2s_debug_tap"); }
extern "C" size_t d2s_workspace_bytes(d2s_handle) { return 0; }
extern "C" int d2s_overlay_fps(const d2s_image *, int, int, const char *, d2s_stream_t) { NOT_YET("d2s_overlay_fps"); }
