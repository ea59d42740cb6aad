/* tsvpp_avframe.h -- the AVFrame the VPP boundary needs.
 * With FFmpeg headers available the real <libavutil/frame.h> is used (define TSVPP_HAVE_FFMPEG or let
 * __has_include find it).  Without them (this image ships none) a STAND-IN struct with exactly the fields
 * VideoProcessor::Convert touches is declared: data[], linesize[], width, height, opaque (reference
 * src/VideoProcessor.cpp:94-166, tests/src/VPPTests.cpp:566-590).  The stand-in is NOT layout-compatible with
 * libavutil's AVFrame (there `opaque` sits far behind `format`): a library built against it must only be used by
 * callers compiled against it too.  include/VideoProcessor.h checks that at start-up (AvFrameLayoutCheck). */
#ifndef TSVPP_AVFRAME_H
#define TSVPP_AVFRAME_H
#include <stdint.h>
#if defined(TSVPP_HAVE_FFMPEG) || (defined(__has_include) && __has_include(<libavutil/frame.h>))
#define TSVPP_AVFRAME_FLAVOUR "ffmpeg"
extern "C" {
#include <libavutil/frame.h>
}
#else
#define TSVPP_AVFRAME_FLAVOUR "shim"
extern "C" {
typedef struct AVFrame {
    uint8_t *data[8];
    int linesize[8];
    uint8_t **extended_data;
    int width, height;
    int nb_samples;
    int format;
    void *opaque;
} AVFrame;
static inline void av_frame_unref(AVFrame *f) { (void)f; }
/* reference counting is libavutil's business; the stand-in shares the planes by plain copy */
static inline int av_frame_ref(AVFrame *dst, const AVFrame *src) {
    *dst = *src;
    return 0;
}
}
#endif
#endif
