/* vpp_oracle.h -- C interface of the CPU oracle (test infrastructure; see vpp_oracle.c). */
#ifndef VPP_ORACLE_H
#define VPP_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* same numbering as the reference enums, include/VideoProcessor.h:20-64 */
enum { VPO_Y800 = 0, VPO_RGB24, VPO_BGR24, VPO_NV12, VPO_UYVY, VPO_YUV444, VPO_HSV };
enum { VPO_PLANAR = 0, VPO_MERGED };
enum { VPO_NEAREST = 0, VPO_BILINEAR, VPO_BICUBIC, VPO_AREA };

typedef struct {
    int crop_x0, crop_y0, crop_x1, crop_y1; /* CropOptions; zero area => no crop */
    int dst_width, dst_height;              /* ResizeOptions; 0 => no resize */
    int resize_type;
    int fourcc;
    int planes;
    int normalization;                       /* 1 => float32 output divided by 255 */
} vpo_params;

/* Whole chain on one frame.  `out` must hold channels*W*H elements (float if normalization
 * or HSV, else uint8).  Returns 0, or the reference's error code. */
int vpo_convert(const uint8_t *y, const uint8_t *uv, int pitch_y, int pitch_uv, int width, int height,
                const vpo_params *p, void *out, int *out_w, int *out_h);

float vpo_channels(int fourcc);
void vpo_yuv_to_rgb(int Y, int U, int V, int *rgb);
void vpo_rgb_to_hsv(float R, float G, float B, float *hsv);
void vpo_bicubic_weights(double w, double c[4]);
void vpo_set_bicubic_fma(int on);
/* pin pow(w,2)/pow(w,3) to externally computed values (see vpp_oracle.c) */
void vpo_pow_table_clear(void);
int vpo_pow_table_set(const double *w, const double *w2, const double *w3, int n);
void vpo_pow_record(int on);
int vpo_pow_recorded(double *out, int cap);
int vpo_area_pattern(float scale, float *out, int max_rows);
long vpo_oob_reads(void);
void vpo_reset_oob(void);
uint32_t vpo_av_crc32(uint32_t crc, const uint8_t *buf, size_t n);

#ifdef __cplusplus
}
#endif
#endif
