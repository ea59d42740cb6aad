/* tsvpp_avframe.h -- the AVFrame the VPP boundary needs.
 * With FFmpeg headers available the real <libavutil/frame.h> is used (define
 * TSVPP_HAVE_FFMPEG or let __has_include find it).  Without them (this image ships none) a
 * layout-compatible prefix with exactly the fields VideoProcessor::Convert touches is declared:
 * data[], linesize[], width, height, opaque (reference src/VideoProcessor.cpp:94-166,
 * tests/src/VPPTests.cpp:566-590). */
#ifndef TSVPP_AVFRAME_H
#define TSVPP_AVFRAME_H
#include <stdint.h>
#if defined(TSVPP_HAVE_FFMPEG) || (defined(__has_include) && __has_include(<libavutil/frame.h>))
extern "C" {
#include <libavutil/frame.h>
}
#else
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
}
#endif
#endif
