/* Minimal stand-in for FFmpeg's <libavutil/frame.h>, used ONLY to compile the reference's VPP
 * sources in place (oracle/Makefile, target _ref/libtsref_vpp.so).  The VPP touches nothing of
 * FFmpeg except these AVFrame fields and av_frame_unref (reference src/VideoProcessor.cpp:164,
 * tests/src/VPPTests.cpp:566-590). */
#ifndef TSVPP_SHIM_AVFRAME_H
#define TSVPP_SHIM_AVFRAME_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
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
#ifdef __cplusplus
}
#endif
#endif
