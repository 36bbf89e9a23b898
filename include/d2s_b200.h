/* d2s_b200 — C ABI of the B200-native depth-inference + stereo-warp hot path.
 *
 * Drop-in boundary for lc700x/desktop2stereo's `depth.py` hot path.  Every entry
 * point takes plain pointers/sizes (device pointers unless stated), enqueues on the caller's
 * stream, never host-synchronises, never allocates caller-visible memory, and returns 0 on
 * success or a non-zero status with a message retrievable through d2s_last_error() — the same
 * convention as the reference's own ctypes→cudart wrapper (reference viewer.py:20-146, rc != 0
 * -> RuntimeError at viewer.py:104-105).
 *
 * Which reference interface each entry point replaces:
 *   d2s_make_sbs        depth.py:2122-2184 make_sbs_core (+ pad_to_aspect_tensor :2106-2119,
 *                       the HWC/float conversion of chw_tensor_to_numpy :767-773, and optionally
 *                       the depth upsample of predict_depth :1998-2004 fused into the gather)
 *   d2s_process         depth.py:542-566  process() CUDA branch (BGRA/BGR u8 HWC -> RGB CHW)
 *   d2s_preprocess      depth.py:676-706 _resize_patch_aligned_t (bicubic+antialias) fused with
 *                       the /255 and mean/std normalisation of predict_depth :1931,1946-1948
 *   d2s_create/_infer/_destroy
 *                       the engine object assigned to DepthModelWrapper.model
 *                       (depth.py:1763-1781; template: TensorRTEngine depth.py:1457-1536)
 *   d2s_postprocess     depth.py:806-867 post_process_depth/normalize, :775 apply_gamma,
 *                       :709-736 apply_foreground_scale, :740-765 anti_alias,
 *                       :1865-1887 DepthStabilizer, :1998-2004 final bilinear upsample
 *   d2s_overlay_fps     depth.py:2061-2103 overlay_fps
 */
#ifndef D2S_B200_H
#define D2S_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *d2s_stream_t;            /* cudaStream_t */
typedef struct d2s_engine *d2s_handle; /* opaque engine (packed weights + workspace), one per GPU */

enum d2s_status {
    D2S_OK = 0,
    D2S_ERR_INVALID = 1,   /* bad argument */
    D2S_ERR_CUDA = 2,      /* a CUDA runtime/driver call failed */
    D2S_ERR_UNSUPPORTED = 3,
    D2S_ERR_NOMEM = 4
};

enum d2s_dtype { D2S_F32 = 0, D2S_F16 = 1, D2S_BF16 = 2, D2S_U8 = 3 };
enum d2s_display_mode { D2S_FULL_SBS = 0, D2S_HALF_SBS = 1, D2S_FULL_TAB = 2, D2S_HALF_TAB = 3 };
enum d2s_warp_mode {
    D2S_WARP_BILINEAR = 0, /* depth.py:2152-2160 grid_sample branch (CUDA path of the reference) */
    D2S_WARP_GATHER = 1    /* depth.py:2163-2172 integer gather branch */
};

/* A strided 3-channel image view: element (c,y,x) lives at base[c*sc + y*sy + x*sx] (in elements).
 * CHW planar RGB: sc=h*w, sy=w, sx=1.   HWC RGB: sc=1, sy=3*w, sx=3.
 * HWC BGRA (capture format, depth.py:549): base=&px[2], sc=-1, sy=4*w, sx=4. */
typedef struct d2s_image {
    void *base;
    int32_t dtype; /* enum d2s_dtype */
    int32_t reserved;
    int64_t sc, sy, sx;
} d2s_image;

typedef struct d2s_warp_params {
    d2s_image rgb; /* source eye image, h x w, values 0..255 */
    d2s_image out; /* packed stereo frame, out_h x out_w (see d2s_sbs_out_shape) */
    const void *depth;   /* [depth_h, depth_w] contiguous, dtype depth_dtype, values in [0,1] */
    int32_t depth_dtype; /* D2S_F32 / D2S_F16 / D2S_BF16: the shift chain rounds to this dtype (depth.py:2143-2147) */
    int32_t depth_h, depth_w; /* == h,w: full-res depth.  Otherwise the bilinear align_corners=False
                                 upsample of depth.py:1998-2004 is evaluated inside the gather. */
    int32_t h, w;
    double ipd_uv, depth_ratio, convergence; /* python floats: scalars are rounded to fp32 only after ipd_uv*W (depth.py:2146) */
    int32_t display_mode; /* enum d2s_display_mode */
    int32_t fill_16_9;
    int32_t warp_mode;    /* enum d2s_warp_mode */
    int32_t rgb_round_to_depth_dtype; /* 1: round rgb values to depth_dtype on load (make_sbs casts
                                         rgb to depth.dtype, depth.py:2209-2215) */
    int32_t *idx_left;    /* optional [h,w] int32 debug taps: floor(ix) or the gather coordinate */
    int32_t *idx_right;
} d2s_warp_params;

/* Output geometry of make_sbs_core for an h x w eye (depth.py:2175-2183). */
int d2s_sbs_out_shape(int h, int w, int display_mode, int fill_16_9, int *out_h, int *out_w);

/* Stereo warp + pad + SBS/TAB pack + (Half modes) 2:1 area mean + clamp, one kernel. */
int d2s_make_sbs(const d2s_warp_params *p, d2s_stream_t stream);

/* process(): BGRA/BGR u8 HWC [h0,w0,ch] -> RGB CHW [3,h,w] of out_dtype (F16/F32).  If (h,w) ==
 * (h0,w0) it is a pure swizzle+cast, else the bilinear+antialias downscale of depth.py:560-566. */
int d2s_process(const uint8_t *frame, int h0, int w0, int channels, void *out, int out_dtype, int h, int w,
                d2s_stream_t stream);

/* Patch-aligned model-input size of depth.py:676-692 for an h x w frame. */
int d2s_model_input_shape(int h, int w, int target, int patch, int *new_h, int *new_w);

/* _resize_patch_aligned_t (bicubic, antialias, align_corners=False) + /255 + (x-mean)/std.
 * src: RGB image view (any layout/dtype of d2s_image), dst: [3,new_h,new_w] CHW of dst_dtype. */
int d2s_preprocess(const d2s_image *src, int h, int w, void *dst, int dst_dtype, int new_h, int new_w,
                   const float mean[3], const float std[3], void *workspace, size_t workspace_bytes,
                   d2s_stream_t stream);
size_t d2s_preprocess_workspace_bytes(int h, int w, int new_h, int new_w);

/* ---- depth network (DINOv2 ViT encoder + DPT head) ---- */
typedef struct d2s_model_config {
    int32_t hidden;      /* D: 384 / 768 / 1024 */
    int32_t layers;      /* L: 12 / 12 / 24 */
    int32_t heads;       /* 6 / 12 / 16 (head dim must be 64) */
    int32_t mlp_hidden;  /* 4*D */
    int32_t patch;       /* 14 */
    int32_t pos_grid;    /* 37 (pos-embed table is pos_grid^2 + 1 rows) */
    int32_t out_indices[4]; /* 1-based hidden-state indices, e.g. {3,6,9,12} */
    int32_t neck[4];     /* neck_hidden_sizes */
    int32_t fusion;      /* fusion_hidden_size */
    int32_t head_hidden; /* 32 */
    float layer_norm_eps;
    float max_depth;     /* 1.0 for relative models */
    int32_t metric;      /* 0: final ReLU, 1: final sigmoid*max_depth */
    int32_t max_batch;   /* > 0: d2s_infer rejects B > max_batch (bounds the per-plan workspace); 0: no limit */
    int32_t max_h, max_w;/* > 0: largest model input accepted; 0: no limit */
    /* Video-Depth-Anything (reference models/video_depth_anything/, depth.py:870-902): */
    int32_t temporal;    /* 0: Depth Anything V2 (per-frame).  1: streaming VDA: 4 temporal modules in the DPT head, one frame per
                            d2s_infer call (B must be 1), per-stream state = rings of the last 32 frames' K/V (vda2_s.py:177-224) */
    float pos_interp_offset; /* 0: HF size-based pos-embed interpolation; 0.1: VDA's scale_factor form (dinov2.py:179-210) */
    int32_t reserved[2];
} d2s_model_config;

/* weight_blob: host pointer to the packed fp32 parameter blob produced by
 * desktop2stereo_b200.weights.pack_state_dict (layout documented there and in DESIGN.md). */
int d2s_create(const void *weight_blob, size_t nbytes, const d2s_model_config *cfg, int device, d2s_handle *out);
int d2s_destroy(d2s_handle h);
/* pixel_values [B,3,H,W] (F16 or F32, normalised) -> predicted_depth [B,H,W] (F16 or F32). */
int d2s_infer(d2s_handle h, const void *pixel_values, int in_dtype, void *depth_out, int out_dtype,
              int B, int H, int W, d2s_stream_t stream);
/* Tile-shape policy of the plans d2s_infer builds from now on (plans of both policies coexist, keyed by stream and policy):
 * LATENCY (default): one frame alone on the GPU — many small tiles, split-K.  THROUGHPUT: several frames in flight on several
 * streams — fewer, wider tiles.  Results agree to fp16 rounding (different accumulation splits), each policy is deterministic. */
enum d2s_policy { D2S_POLICY_LATENCY = 0, D2S_POLICY_THROUGHPUT = 1 };
int d2s_set_policy(d2s_handle h, int policy);
/* temporal engines: forget the video state bound to `stream` (frame counter + K/V rings; ONE state per stream and input size,
 * shared by the plans of both policies and all dtypes) — the next frame is a first frame again (vda2_s.py:196
 * `if not self.transform`).  All frames of one video must be submitted on one stream, in order. */
int d2s_reset_stream(d2s_handle h, d2s_stream_t stream);
/* Free everything the engine keeps for `stream`: its plans (activation buffers, graphs) and, for temporal engines, the video's
 * state.  Call it before destroying a stream the engine has seen (a recycled stream handle would otherwise inherit the old
 * video's window).  Host-synchronous.  The plan cache is also LRU-bounded (64 plans, env D2S_MAX_PLANS). */
int d2s_release_stream(d2s_handle h, d2s_stream_t stream);
/* Debug/parity taps: copy an internal activation (by name) to a caller buffer as fp32. */
int d2s_debug_tap(d2s_handle h, const char *name, float *dst, size_t max_elems, size_t *n_elems, d2s_stream_t stream);
size_t d2s_workspace_bytes(d2s_handle h);
int64_t d2s_launch_count(void); /* kernels launched by this library since load (for gpu_launches) */

/* ---- depth post-process ---- */
typedef struct d2s_post_params {
    const void *depth_in; /* [H,W] raw predicted_depth */
    int32_t in_dtype;
    int32_t H, W;
    void *out;            /* [out_h,out_w], dtype out_dtype: post-processed (+EMA) depth, upsampled */
    int32_t out_dtype;
    int32_t out_h, out_w; /* == H,W: no upsample */
    int32_t compute_dtype; /* dtype the reference would compute in (F16 on CUDA, BF16 on CPU, F32) */
    int32_t metric;       /* 1: 1/d on d>0 before the percentile clip */
    float percentile;     /* 2.0 */
    int32_t subsample_cap;/* 6144 */
    float gamma;          /* 1.45 */
    float foreground_scale; /* settings/10 */
    float aa_strength;    /* settings*2 */
    void *ema_state;      /* optional [H,W] compute_dtype persistent buffer (DepthStabilizer.prev) */
    int32_t ema_valid;    /* 0: first frame (state := depth), 1: lerp, 2: per element — a NaN in the state means "unset" (state :=
                             depth there), so a state buffer filled with 0xFF bytes starts a stream without a host-side flag */
    float ema_alpha;      /* 0.9 */
    void *out_lowres;     /* optional [H,W] compute_dtype: the post-processed (+EMA) map before the upsample */
    void *workspace; size_t workspace_bytes;
} d2s_post_params;
size_t d2s_postprocess_workspace_bytes(int H, int W);
int d2s_postprocess(const d2s_post_params *p, d2s_stream_t stream);

/* overlay_fps: blends the "FPS: xx.x" glyph mask into an RGB image in place. */
int d2s_overlay_fps(const d2s_image *rgb, int h, int w, const char *text, d2s_stream_t stream);

/* ---- device-side output encode, first stages ----
 * Packed u8 HWC frame -> NV12 [h rows of Y | h/2 rows of interleaved CbCr], the colour conversion + 4:2:0 downsample of the JPEG
 * encoder the reference runs on the host (cv2.imencode in streamer.py:250-256; libjpeg jccolor.c / jcsample.c arithmetic), 1.5 B/px:
 * the layout NVENC / nvJPEG take.  h, w even.  row_pitch_bytes <= 0: tightly packed (3 * w). */
int d2s_rgb_to_nv12(const uint8_t *rgb_hwc, int64_t row_pitch_bytes, int h, int w, uint8_t *nv12, d2s_stream_t stream);

/* Packed u8 HWC frame -> a complete baseline JPEG stream, encoded on the device: replaces cv2.imencode(".jpg", bgr,
 * [IMWRITE_JPEG_QUALITY, quality]) in MJPEGStreamer._encoder_loop (reference streamer.py:250-256) AND make_sbs's float32
 * device->host copy (depth.py:767-773); MJPEGStreamer's `encoded_frame` (streamer.py:254) is the wire boundary that remains.
 * The stream is byte-identical to what cv2.imencode (OpenCV's libjpeg-turbo: JFIF, YCbCr 4:2:0, Annex K tables, islow DCT) writes
 * for the same frame with IMWRITE_JPEG_RST_INTERVAL = restart_interval; restart markers (>= 1 MCU of 16x16 pixels per interval,
 * required: the encoder is parallel over intervals) do not change decoded pixels.  h, w even.
 *   jpeg / capacity   device buffer for the stream; d2s_jpeg_max_bytes() can never overflow, smaller is allowed
 *   size_out          device uint32: stream length in bytes, or 0 if it did not fit `capacity`
 *   workspace         device scratch of d2s_jpeg_workspace_bytes(h, w, restart_interval)
 * Asynchronous on `stream`; never allocates. */
size_t d2s_jpeg_workspace_bytes(int h, int w, int restart_interval);
size_t d2s_jpeg_max_bytes(int h, int w, int restart_interval);
int d2s_jpeg_encode(const uint8_t *rgb_hwc, int64_t row_pitch_bytes, int h, int w, int quality, int restart_interval, uint8_t *jpeg,
                    size_t capacity, uint32_t *size_out, void *workspace, size_t workspace_bytes, d2s_stream_t stream);

/* ---- occlusion-aware stereo rendering ----
 * The reference's OpenGL viewer warps with a fragment shader that handles disocclusions (reference viewer.py:386-631:
 * 3-tap depth smoothing, depth shaping, edge falloff, 2-tap disocclusion confidence :421-435, push-pull inpaint :437-506, border
 * alpha, feathering, rounded corners); d2s_make_sbs_dibr evaluates that shader per output pixel of both eye views (left eye
 * u_eye_offset = -ipd_uv/2, right +ipd_uv/2, u_depth_strength = 0.1 * depth_ratio; viewer.py:1334, 2680-2760) and packs them like
 * the viewer's viewports: Full modes render each eye at frame size, Half-SBS / Half-TAB at half width / half height.
 * The output is colour x alpha over black, scaled to 0..255 like d2s_make_sbs.  Parity: bit-exact against oracle/dibr_oracle.c
 * (no OpenGL oracle exists: the reference never sets u_resolution, see that file's header). */
typedef struct d2s_dibr_params {
    d2s_image rgb;        /* source frame h x w, values 0..255 (u8 / f16 / f32); sampled as the normalised texture value / 255 */
    d2s_image out;        /* packed frame, out_h x out_w (d2s_dibr_out_shape), f32 / f16 / u8 */
    const void *depth;    /* [h, w] contiguous, F32 or F16: the texture the viewer uploads (predict_depth's map) */
    int32_t depth_dtype;
    int32_t h, w;
    int32_t display_mode; /* enum d2s_display_mode */
    double ipd_uv, depth_ratio, convergence, roll;   /* viewer.py:1326-1340; roll in radians (u_roll, 0 in the reference) */
    float resolution_x, resolution_y;  /* u_resolution; <= 0: the eye view's size ("viewport resolution", viewer.py:395) */
    int32_t search_radius;             /* 12   (u_search_radius, viewer.py:402) */
    float depth_tolerance;             /* 0.012 (u_depth_tolerance) */
    float blur_radius;                 /* 2.5  (u_blur_radius) */
    int32_t feather_enabled;           /* 0    (viewer.py:1326 feather_enabled=False) */
    float feather_width;
    float corner_radius;               /* 0 */
    void *workspace;                   /* optional device scratch (16-byte aligned, d2s_dibr_workspace_bytes): with it the inpaint sweeps of
                                          the pixels on depth edges run in a dense second pass (same frame, ~3x faster); NULL: one pass */
    size_t workspace_bytes;
} d2s_dibr_params;
size_t d2s_dibr_workspace_bytes(int h, int w, int display_mode);
int d2s_dibr_out_shape(int h, int w, int display_mode, int *view_h, int *view_w, int *out_h, int *out_w);
int d2s_make_sbs_dibr(const d2s_dibr_params *p, d2s_stream_t stream);

/* ---- whole-frame pipeline: the caller side of the hot path ----
 * Replaces the per-frame sequence main.py drives (reference main.py:232-262 process -> predict_depth, :1336-1341 make_sbs ->
 * streamer.set_frame) with ONE call per frame.  A pipe owns `slots` frame slots; each slot has a CUDA stream, fixed device
 * buffers, pinned host buffers (host_io) and CUDA graphs of the frame's ~150 kernels, so frames in flight overlap copy, network
 * and warp.  Results are bit-identical to d2s_process -> d2s_preprocess -> d2s_infer -> d2s_postprocess -> d2s_make_sbs on the
 * same frame (same kernels).  The DepthStabilizer EMA (depth.py:1865-1887) orders consecutive frames with an event between the
 * two graphs of a frame; submit frames of one video in order.  One caller thread per pipe. */
typedef struct d2s_pipe *d2s_pipe_handle;
enum d2s_out_format { D2S_OUT_PACKED = 0, D2S_OUT_NV12 = 1, D2S_OUT_JPEG = 2 };
/* D2S_OUT_JPEG result of one stream, at stride `out_bytes` (d2s_pipe_geometry) in the slot's output buffer.  With host_io the pipe
 * copies only as many bytes as recent frames needed (a frame that outgrows the estimate costs one extra copy inside d2s_pipe_wait);
 * d2s_pipe_wait fails with D2S_ERR_INVALID if a stream did not fit out_bytes (oh * ow * 3 + 4 KB: above uniform noise at quality 100). */
typedef struct d2s_pipe_jpeg_frame {
    uint32_t size;          /* bytes of `data` that hold the stream (SOI ... EOI) */
    uint32_t reserved[3];
    uint8_t data[1];
} d2s_pipe_jpeg_frame;
typedef struct d2s_pipe_config {
    int32_t frame_h, frame_w, channels; /* captured frame: u8 HWC, BGRA (4) or BGR (3) (depth.py:549) */
    int32_t target_height;        /* process(img, target_height): bilinear-antialias downscale when < frame_h (depth.py:555-566) */
    int32_t rgb_dtype;            /* dtype of process()'s tensor: D2S_F16 (FP16 setting on) or D2S_F32 */
    int32_t depth_resolution;     /* DEPTH_RESOLUTION (518) */
    int32_t patch;                /* 14 */
    float mean[3], std[3];        /* depth.py:1794-1799 */
    int32_t metric;               /* post_process_depth / DepthStabilizer: the fields of d2s_post_params */
    float percentile;
    int32_t subsample_cap;
    float gamma, foreground_scale, aa_strength;
    int32_t use_temporal_smooth;
    float ema_alpha;
    double ipd_uv, depth_ratio, convergence;   /* make_sbs (depth.py:2186) */
    int32_t display_mode, fill_16_9;
    int32_t out_dtype;            /* packed frame [oh, ow, 3] HWC: D2S_F32 = what make_sbs returns (depth.py:2231), D2S_U8, D2S_F16 */
    int32_t out_format;           /* what the caller receives per stream.  D2S_OUT_PACKED (0): the packed frame [oh, ow, 3] of out_dtype;
                                     D2S_OUT_NV12 (1; needs out_dtype U8, even oh / ow): [oh * 3 / 2, ow] u8 (d2s_rgb_to_nv12);
                                     D2S_OUT_JPEG (2; same needs): a d2s_pipe_jpeg_frame — the complete JPEG stream d2s_jpeg_encode
                                     writes for the packed frame (byte-identical to cv2.imencode, streamer.py:250-256) */
    int32_t slots;                /* frames in flight (1..64); > 1 builds throughput-policy plans */
    int32_t host_io;              /* 1: frames come from and results go to pinned HOST memory (H2D / D2H copies on the slot's stream) */
    int32_t streams;              /* concurrent video streams sharing the pipe (0/1: one).  One submit takes ONE frame of EVERY stream:
                                     `frame` is [streams][frame_h, frame_w, channels], the result [streams][oh, ow, 3]; the network runs
                                     them as one batch (BASELINE configs 3/5: 8 x 4K), each stream keeps its own DepthStabilizer state */
    int32_t jpeg_quality;         /* D2S_OUT_JPEG: IMWRITE_JPEG_QUALITY (MJPEGStreamer's `quality`, streamer.py:113); 0 -> 90 */
    int32_t jpeg_restart_interval;/* D2S_OUT_JPEG: MCUs (16x16 pixels) per restart interval; 0 -> 4 */
    int32_t fps_overlay;          /* 1: make_sbs's `fps=` argument (depth.py:2186, 2226-2227): d2s_pipe_set_fps_text() may give a text that
                                     overlay_fps (d2s_overlay_fps) draws onto the RGB frame before the warp — after the depth network
                                     has read the frame, as in the reference, where predict_depth sees the frame without it */
} d2s_pipe_config;
/* Host-synchronous (allocates the slots, builds their plans, captures their graphs).  The pipe borrows `engine`: destroy the pipe
 * before the engine. */
int d2s_pipe_create(d2s_handle engine, const d2s_pipe_config *cfg, d2s_pipe_handle *out);
int d2s_pipe_destroy(d2s_pipe_handle p);
/* frame_bytes / out_bytes are per stream (one frame); a slot's buffers hold `streams` of them back to back */
int d2s_pipe_geometry(d2s_pipe_handle p, int *h, int *w, int *model_h, int *model_w, int *out_h, int *out_w, size_t *frame_bytes, size_t *out_bytes);
/* The slot's own buffers (valid for the life of the pipe): pinned host frame / result (host_io), device frame / result / depth
 * [h,w] fp16 (what predict_depth returns), and its stream. */
int d2s_pipe_slot_buffers(d2s_pipe_handle p, int slot, void **host_in, void **host_out, void **dev_in, void **dev_out, void **dev_depth, d2s_stream_t *stream);
/* Enqueue one frame on `slot` and return.  frame: NULL = the slot's own input buffer, else a pinned host pointer (host_io) or a
 * device pointer whose contents are complete on stream `frame_ready_on`.  The result lands in the slot's host_out / dev_out. */
int d2s_pipe_submit(d2s_pipe_handle p, int slot, const void *frame, d2s_stream_t frame_ready_on);
int d2s_pipe_wait(d2s_pipe_handle p, int slot);   /* block until the slot's frame is complete; the slot may then be reused */
int d2s_pipe_reset(d2s_pipe_handle p);            /* new video: EMA state (and a temporal engine's window) forgotten; host-synchronous */
/* The text (<= 32 characters, e.g. "FPS: 59.9") drawn onto the frames of the following submits; NULL or "": none.  Needs
 * d2s_pipe_config.fps_overlay.  The every-10th-call refresh of the reference's text cache (depth.py:2061-2072) is the caller's:
 * desktop2stereo_b200/overlay.py keeps it. */
int d2s_pipe_set_fps_text(d2s_pipe_handle p, const char *text);
int d2s_pipe_set_trace(d2s_pipe_handle p, int on);               /* record per-stage CUDA events on the frames submitted from now on */
int d2s_pipe_slot_times(d2s_pipe_handle p, int slot, float ms[3]); /* after wait: process | resize+network+post | upsample+warp */

/* ---- kernel-level parity hooks (used by tests/ only; fp16 operands, row-major) ----
 * d2s_debug_gemm:      C[M,N] = act(A[M,K] * Bw[N,K]^T + bias)   and/or   x32[M,N] += A*Bw^T + bias   (tcgen05 GEMM)
 * d2s_debug_conv3x3:   NHWC [B,H,W,Cp] (*) Wt[N, 9*Cp] (k = (ky*3+kx)*Cp + c), pad 1, + bias, act, + res1; optional relu copy
 * d2s_debug_attention: qkv [B,N,3D] -> softmax(q k^T / 8) v, [B,N,D], head dim 64 */
int d2s_debug_gemm(const void *A, const void *Bw, const float *bias, void *C, int M, int N, int K, int act,
                   float *x32_accumulate, d2s_stream_t stream);
int d2s_debug_conv3x3(const void *A, const void *Wt, const float *bias, void *C, int B, int H, int W, int Cp, int N, int act,
                      const void *res1, void *c_relu, d2s_stream_t stream);
int d2s_debug_set_gemm_policy(int policy); /* enum d2s_policy for the plans d2s_debug_gemm builds on the calling thread */
int d2s_debug_attention(const void *qkv, void *out, int B, int N, int D, int heads, d2s_stream_t stream);
/* 1: d2s_make_sbs always takes the generic (output-centric) kernel — tests compare it with the shared-memory fast path */
int d2s_debug_force_generic_warp(int on);

const char *d2s_last_error(void);
const char *d2s_version(void);

#ifdef __cplusplus
}
#endif
#endif /* D2S_B200_H */
