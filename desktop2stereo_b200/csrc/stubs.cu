// Entry points not implemented yet return D2S_ERR_UNSUPPORTED (never a CPU fallback).
#include "common.cuh"
using namespace d2s;
#define NOT_YET(name) return set_error(D2S_ERR_UNSUPPORTED, name ": not implemented yet")
extern "C" int d2s_overlay_fps(const d2s_image *, int, int, const char *, d2s_stream_t) { NOT_YET("d2s_overlay_fps"); }
