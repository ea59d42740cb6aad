/*
 * tsvpp.h -- C ABI of the B200-native video-post-processing (VPP) path.
 *
 * This is the drop-in boundary for ONE hot path of osai-ai/tensor-stream: the per-frame chain
 *   crop -> NV12 resize (nearest/bilinear/bicubic/area) -> NV12->{RGB,BGR,Y800,HSV,...}
 *        -> [0,1] normalisation -> planar/merged layout
 * which the reference runs as 1-3 kernels + 3-7 cudaMalloc + 2-4 cudaFree per frame behind
 *   VideoProcessor::Convert            reference src/VideoProcessor.cpp:94-166
 *   cropHost / cropKernel              reference src/Crop.cu:4-48
 *   resizeKernel / resizeNV12*Kernel   reference src/Resize.cu:160-473
 *   colorConversionKernel<T> / NV12To* reference src/ColorConversion.cu:6-382
 * and which this library runs as ONE fused sm_100a kernel per (resize type, output format) pair
 * over a whole batch of frames.
 *
 * Plain pointers and sizes only (no torch / C++ types): this is what the reference's
 * VideoProcessor (C++), a cgo/JNI/ctypes binding, the PyTorch extension and bench.py all bind.
 * Results are bit-identical to the reference's for every format (see tests/).
 *
 * Numbering of the enums is the reference's (include/VideoProcessor.h:20-64) so the values can
 * be passed through unchanged.
 */
#ifndef TSVPP_H
#define TSVPP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* FourCC -- reference include/VideoProcessor.h:20-28 */
enum { TSVPP_Y800 = 0, TSVPP_RGB24 = 1, TSVPP_BGR24 = 2, TSVPP_NV12 = 3, TSVPP_UYVY = 4, TSVPP_YUV444 = 5, TSVPP_HSV = 6 };
/* Planes -- reference include/VideoProcessor.h:32-35 */
enum { TSVPP_PLANAR = 0, TSVPP_MERGED = 1 };
/* ResizeType -- reference include/VideoProcessor.h:59-64 */
enum { TSVPP_NEAREST = 0, TSVPP_BILINEAR = 1, TSVPP_BICUBIC = 2, TSVPP_AREA = 3 };

/* Status codes: 0, the reference's Internal enum (include/Common.h:19-24) or a positive
 * cudaError_t, exactly what VideoProcessor::Convert returns. */
enum { TSVPP_OK = 0, TSVPP_REPEAT = -1, TSVPP_UNSUPPORTED = -2, TSVPP_ERROR = -3 };

/* One decoded NV12 frame in DEVICE memory: what the reference reads from its input AVFrame
 * (data[0], data[1], linesize[0], linesize[1], width, height; src/VideoProcessor.cpp:94,
 * src/Resize.cu:418-421).  pitch 0 means "pitch = width" like the reference. */
typedef struct TsvppFrame {
    const uint8_t *y;  /* luma plane,  height rows of pitch_y bytes          */
    const uint8_t *uv; /* interleaved chroma plane, height/2 rows of pitch_uv */
    int pitch_y;
    int pitch_uv;
    int width;
    int height;
} TsvppFrame;

/* FrameParameters flattened (reference include/VideoProcessor.h:38-105). */
typedef struct TsvppParams {
    int crop_x0, crop_y0, crop_x1, crop_y1; /* CropOptions; applied iff 0<w<W && 0<h<H (VideoProcessor.cpp:109) */
    int dst_width, dst_height;              /* ResizeOptions; 0 => keep size                                   */
    int resize_type;                        /* TSVPP_NEAREST..TSVPP_AREA                                       */
    int fourcc;                             /* TSVPP_Y800..TSVPP_HSV                                           */
    int planes;                             /* TSVPP_PLANAR / TSVPP_MERGED (RGB24/BGR24 only)                  */
    int normalization;                      /* 1 => float32 output in [0,1]; HSV is always float32             */
} TsvppParams;

typedef struct TsvppContext TsvppContext;

/* Context bound to one CUDA device (plan cache, descriptor ring, staging buffers).
 * Replaces the per-VideoProcessor state of reference src/VideoProcessor.cpp:79-92; queries the
 * device it is created on (the reference always asked device 0, :83). */
int tsvpp_create(int device, TsvppContext **out);
void tsvpp_destroy(TsvppContext *ctx);

/* Output geometry after the dispatch rules of reference src/VideoProcessor.cpp:106-135.
 * *elem_size is 4 (float32) or 1 (uint8); *elems = channelsByFourCC * w * h
 * (src/VideoProcessor.cpp:4-14). */
int tsvpp_output_info(int src_width, int src_height, const TsvppParams *p, int *out_w, int *out_h,
                      size_t *elems, int *elem_size);

/* THE hot path: convert n frames (all of the same width/height/pitches) with one fused kernel
 * launch on `stream` (a cudaStream_t; NULL = default stream).  out[i] is a DEVICE buffer of
 * tsvpp_output_info() elems, tightly packed, layout as the reference's (MERGED (H,W,3), PLANAR
 * (3,H,W), Y800 (H,W), HSV (H,W,3) float).  Asynchronous w.r.t. the host; stream-ordered.
 * Replaces: VideoProcessor::Convert's crop/resize/colour stages, src/VideoProcessor.cpp:106-143. */
int tsvpp_convert_batch(TsvppContext *ctx, const TsvppFrame *frames, int n, const TsvppParams *p,
                        void *const *out, void *stream);

/* Same, but the frame descriptors / output pointers are uploaded once and the launch is replayed
 * (streaming pipelines that recycle a ring of decode surfaces; bench.py).  */
typedef struct TsvppBatch TsvppBatch;
int tsvpp_batch_create(TsvppContext *ctx, const TsvppFrame *frames, int n, const TsvppParams *p,
                       void *const *out, TsvppBatch **batch);
int tsvpp_batch_run(TsvppBatch *batch, void *stream);
void tsvpp_batch_destroy(TsvppBatch *batch);

/* Per-frame entry for pipelines that call once per decoded frame from several consumer threads (the reference's
 * TensorStream::getFrame -> VideoProcessor::Convert, src/Wrappers/WrapperPython.cpp:265-350): calls that are pending
 * at the same time and agree on geometry and parameters are converted by ONE fused launch on the context's
 * coalescing stream.  `out` must be usable on that stream (e.g. cudaMallocAsync'd on tsvpp_coalesce_stream()).
 * wait > 0: return when this frame's result is complete; wait == 0: return once the launch that covers it has been
 * issued (stream-ordered; the coalescing stream is a blocking stream, so work on the legacy default stream is
 * ordered after it like after the reference's per-consumer streams); wait < 0 ("posted"): return at once -- the
 * frame is picked up by the context's launcher thread; the input planes and `out` must stay untouched until
 * tsvpp_coalesce_wait() returns, which also reports the first error of a posted frame. */
int tsvpp_convert_coalesced(TsvppContext *ctx, const TsvppFrame *frame, const TsvppParams *p, void *out, int wait);
void *tsvpp_coalesce_stream(TsvppContext *ctx);
/* Returns when every coalesced frame issued so far is complete (wait == 0 callers use it before they reuse or read
 * their outputs on the host). */
int tsvpp_coalesce_wait(TsvppContext *ctx);
/* frames converted / launches issued by tsvpp_convert_coalesced so far (their ratio is the mean batch size) */
int tsvpp_coalesce_stats(const TsvppContext *ctx, unsigned long long *frames, unsigned long long *launches);
/* host nanoseconds spent issuing those launches (plan lookup + descriptor upload + launch + event record) */
unsigned long long tsvpp_coalesce_issue_ns(const TsvppContext *ctx);

/* End-to-end entry with HOST buffers: each frame is a host NV12 image (y: height*pitch_y bytes,
 * uv: height/2*pitch_uv bytes); outputs are written to host memory.  Copies are pipelined with
 * the kernel over internal streams; returns when all outputs are complete.  `frames[i].y/uv` and
 * out[i] are HOST pointers here (pinned memory gives the best rate). */
int tsvpp_convert_batch_host(TsvppContext *ctx, const TsvppFrame *frames, int n, const TsvppParams *p,
                             void *const *out);

/* Bit 0: the last tsvpp_convert_batch_host() call read its inputs in place from pinned host memory
 * (zero-copy, only the needed source rows cross PCIe) instead of staging whole frames with H2D
 * copies.  Bit 1: it wrote the outputs straight into pinned host memory (no D2H copy). */
int tsvpp_last_host_path(const TsvppContext *ctx);

/* Counters for evidence: number of fused-kernel launches issued by this context so far. */
unsigned long long tsvpp_launch_count(const TsvppContext *ctx);
/* Human-readable description of the plan chosen for (geometry, params): tile width, stages,
 * smem bytes, fill mode ("bulk" = cp.async.bulk/TMA, "ldg" = fallback for unaligned frames). */
int tsvpp_describe_plan(TsvppContext *ctx, const TsvppFrame *frame0, const TsvppParams *p, char *buf, size_t buflen);
/* number of plans currently cached by the context (LRU-bounded, TSVPP_MAX_PLANS, default 64) */
int tsvpp_plan_cache_size(TsvppContext *ctx);

/* Test hook: pow(w,2) and pow(w,3) as evaluated by the device's libdevice (the bicubic weights of
 * reference src/Resize.cu:45-48 depend on it bit for bit; the CPU oracle is pinned to these). */
int tsvpp_debug_pow23(const double *w, int n, double *w2, double *w3);

const char *tsvpp_version(void);
/* Text of the last failure on the calling thread (which runtime call, where). */
const char *tsvpp_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* TSVPP_H */
