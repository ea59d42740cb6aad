/*
 * vpp_oracle.c -- CPU restatement of the reference VPP chain (crop -> NV12 resize ->
 * colour conversion -> normalisation -> layout) of osai-ai/tensor-stream.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (tensor_stream_b200/csrc) never links, imports or falls back to anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against the
 * reference's own golden CRC-32 values (tests/src/VPPTests.cpp, PythonTests.cpp) and its
 * seven float32 golden dumps (tests/resources/test_references), and tests/test_parity_gpu.py (test_reference_vpp_equals_oracle)
 * (GPU) checks it against the reference's unmodified .cu files compiled for sm_100a
 * (oracle/_ref).
 *
 * Every function cites the reference file:line it restates.  The arithmetic is the
 * reference's *as compiled by nvcc with default -fmad=true*: the positions of the fused
 * multiply-adds are part of the observable behaviour (bilinear/area results are truncated to
 * uint8), so they are written out with fmaf() and this file must be compiled with
 * -ffp-contract=off so that the host compiler adds none of its own.
 *
 * Out-of-bounds reads that the reference performs (undefined behaviour there; see
 * SURVEY.md A.7) are defined here as "read 0" and counted in vpo_oob_reads().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "vpp_oracle.h"

static _Thread_local long g_oob = 0;

long vpo_oob_reads(void) { return g_oob; }
void vpo_reset_oob(void) { g_oob = 0; }

/* ------------------------------------------------------------------------------------ */
/* A plane of bytes with a known extent; reads outside it return 0 and are counted.      */
typedef struct {
    const uint8_t *p;
    long size;
} plane_t;

static inline int rd(const plane_t *pl, long idx) {
    if (idx < 0 || idx >= pl->size) {
        g_oob++;
        return 0;
    }
    return pl->p[idx];
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* tight NV12 image: Y is w*h, UV is w*(h/2); both pitches == w (reference Resize.cu:411-412,
 * Crop.cu:29-30 allocate exactly this and the temp AVFrame's linesize stays 0 => pitch=width,
 * Resize.cu:420-421, ColorConversion.cu:301) */
typedef struct {
    uint8_t *y, *uv;
    int w, h;
} nv12_t;

static int nv12_alloc(nv12_t *im, int w, int h) {
    im->w = w;
    im->h = h;
    im->y = (uint8_t *)calloc((size_t)w * h + 64, 1);
    /* two extra zeroed chroma rows: for odd w/h the reference's colour conversion indexes one
     * row/byte past the chroma plane (undefined there); defined as 0 here */
    im->uv = (uint8_t *)calloc((size_t)w * (h / 2 + 2) + 64, 1);
    return (im->y && im->uv) ? 0 : -1;
}
static void nv12_free(nv12_t *im) {
    free(im->y);
    free(im->uv);
    im->y = im->uv = NULL;
}

/* ------------------------------------------------------------------------------------ */
/* Crop -- reference src/Crop.cu:4-21 (cropKernel)                                         */
static void crop_nv12(const plane_t *Y, const plane_t *UV, int pitchY, int pitchUV, int x0, int y0,
                      int x1, int y1, nv12_t *out) {
    int cw = x1 - x0, ch = y1 - y0;
    for (int i = 0; i < ch; i++) {
        for (int j = 0; j < cw; j++) {
            int UVRow = i / 2;
            int UVCol = (j % 2 == 0) ? j : j - 1;
            long uS = (long)(y0 / 2 + UVRow) * pitchUV + (UVCol + x0);
            long vS = uS + 1;
            long uD = (long)UVRow * cw + UVCol;
            out->y[j + (long)i * cw] = (uint8_t)rd(Y, (long)(x0 + j) + (long)(y0 + i) * pitchY);
            /* rows i >= 2*(ch/2) cannot occur for even ch; odd ch writes past ch/2 rows in the
             * reference (heap overflow) -- guarded here */
            if (UVRow < ch / 2) {
                out->uv[uD] = (uint8_t)rd(UV, uS);
                if (UVCol + 1 < cw) out->uv[uD + 1] = (uint8_t)rd(UV, vS);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* Bilinear tap -- reference src/Resize.cu:5-25 (calculateBillinearInterpolation).
 * x and y arrive as floats (the device function's signature), the start index is computed in
 * float arithmetic, contracted by nvcc to one FMA and truncated.                           */
static int bilinear_tap(const plane_t *d, float x, float y, int xDiff, int yDiff, int linesize,
                        int width, int height, float wx, float wy) {
    int s = (int)fmaf(y, (float)linesize, x);
    if (x + (float)xDiff >= (float)width) xDiff = 0;
    if (y + (float)yDiff >= (float)height) linesize = 0;
    float A = (float)rd(d, s);
    float B = (float)rd(d, (long)s + xDiff);
    float C = (float)rd(d, (long)s + (long)linesize * yDiff);
    float D = (float)rd(d, (long)s + (long)linesize * yDiff + xDiff);
    float omx = 1.0f - wx, omy = 1.0f - wy;
    float v = fmaf(A * omx, omy, (B * wx) * omy);
    v = fmaf(omx, C * wy, v);
    v = fmaf(D, wx * wy, v);
    return (int)v;
}

/* reference src/Resize.cu:269-312 (resizeNV12BilinearKernel) */
static void resize_bilinear(const plane_t *Y, const plane_t *UV, int sw, int sh, int lsY, int lsUV,
                            nv12_t *o) {
    int dw = o->w, dh = o->h;
    float xr = (float)sw / dw, yr = (float)sh / dh; /* Resize.cu:418-419 */
    for (int i = 0; i < dh; i++) {
        float yF = fmaf((float)(unsigned)i + 0.5f, yr, -0.5f);
        int y = (int)floorf(yF);
        float wy = yF - (float)y;
        if (y < 0) { y = 0; wy = 0; }
        if (y > sh - 1) { y = sh - 1; wy = 0; }
        for (int j = 0; j < dw; j++) {
            float xF = fmaf((float)(unsigned)j + 0.5f, xr, -0.5f);
            int x = (int)floorf(xF);
            float wx = xF - (float)x;
            if (x < 0) { x = 0; wx = 0; }
            if (x > sw - 1) { x = sw - 1; wx = 0; }
            o->y[(long)i * dw + j] =
                (uint8_t)bilinear_tap(Y, (float)x, (float)y, 1, 1, lsY, sw, sh, wx, wy);
            if (i < dh / 2 && j < dw / 2) {
                o->uv[(long)i * dw + 2 * j] =
                    (uint8_t)bilinear_tap(UV, (float)(2 * x), (float)y, 2, 1, lsUV, sw, sh / 2, wx, wy);
                o->uv[(long)i * dw + 2 * j + 1] =
                    (uint8_t)bilinear_tap(UV, (float)(2 * x + 1), (float)y, 2, 1, lsUV, sw, sh / 2, wx, wy);
            }
        }
    }
}

/* reference src/Resize.cu:242-267 (resizeNV12NearestKernel) */
static void resize_nearest(const plane_t *Y, const plane_t *UV, int sw, int sh, int lsY, int lsUV,
                           nv12_t *o) {
    int dw = o->w, dh = o->h;
    float xr = (float)sw / dw, yr = (float)sh / dh;
    (void)sh;
    for (int i = 0; i < dh; i++) {
        int y = (int)(yr * (float)(unsigned)i);
        for (int j = 0; j < dw; j++) {
            int x = (int)(xr * (float)(unsigned)j);
            o->y[(long)i * dw + j] = (uint8_t)rd(Y, (long)y * lsY + x);
            if (i < dh / 2 && j < dw / 2) {
                o->uv[(long)i * dw + 2 * j] = (uint8_t)rd(UV, (long)y * lsUV + 2 * x);
                o->uv[(long)i * dw + 2 * j + 1] = (uint8_t)rd(UV, (long)y * lsUV + 2 * x + 1);
            }
        }
    }
}

/* reference src/Resize.cu:214-240 (resizeNV12UpscaleAreaKernel) */
static void resize_area_up(const plane_t *Y, const plane_t *UV, int sw, int sh, int lsY, int lsUV,
                           nv12_t *o) {
    int dw = o->w, dh = o->h;
    float xr = (float)sw / dw, yr = (float)sh / dh;
    for (int i = 0; i < dh; i++) {
        int y = (int)floorf(yr * (float)(unsigned)i);
        float wy = (float)(unsigned)(i + 1) - (float)(y + 1) / yr;
        if (wy <= 0) wy = 0; else wy = wy - floorf(wy);
        for (int j = 0; j < dw; j++) {
            int x = (int)floorf(xr * (float)(unsigned)j);
            float wx = (float)(unsigned)(j + 1) - (float)(x + 1) / xr;
            if (wx <= 0) wx = 0; else wx = wx - floorf(wx);
            o->y[(long)i * dw + j] =
                (uint8_t)bilinear_tap(Y, (float)x, (float)y, 1, 1, lsY, sw, sh, wx, wy);
            if (i < dh / 2 && j < dw / 2) {
                o->uv[(long)i * dw + 2 * j] =
                    (uint8_t)bilinear_tap(UV, (float)(2 * x), (float)y, 2, 1, lsUV, sw, sh / 2, wx, wy);
                o->uv[(long)i * dw + 2 * j + 1] =
                    (uint8_t)bilinear_tap(UV, (float)(2 * x + 1), (float)y, 2, 1, lsUV, sw, sh / 2, wx, wy);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* Bicubic -- reference src/Resize.cu:27-91 (calculateBicubicSplineInterpolation), FP64.
 * Keys cubic a=-0.75, horizontal pass per row rounded+clamped to int, then vertical.
 * g_bicubic_fma selects the accumulate form: 0 = plain mul/add in source order,
 * 1 = the DFMA chain nvcc emits (mul, then three fma).  The two differ only when a sum lands
 * within 1 ulp of a .5 tie.                                                                */
static int g_bicubic_fma = 1;
void vpo_set_bicubic_fma(int on) { g_bicubic_fma = on; }

/* pow() is the one primitive of the path whose result is implementation-defined: the compiled
 * reference calls CUDA libdevice's pow (Resize.cu:45-48), which is NOT correctly rounded, and
 * with dyadic weights (ratios like 1.5, 2) sums land exactly on .5 ties, so a 1-ulp difference
 * in pow(w,3) flips pixels.  CPU-only runs use libm's pow.  GPU tests pin the oracle to the
 * device's values: record the distinct w arguments (vpo_pow_record), evaluate pow(w,2), pow(w,3)
 * on the GPU (tsvpp_debug_pow23), install them (vpo_pow_table_set); everything else of the
 * bicubic arithmetic is still evaluated here. */
#define POW_TAB_CAP 16384
static double g_pw[POW_TAB_CAP], g_pw2[POW_TAB_CAP], g_pw3[POW_TAB_CAP];
static int g_pow_n = 0;       /* installed entries (open addressing, cap POW_TAB_CAP) */
static uint8_t g_pow_used[POW_TAB_CAP];
static int g_pow_record = 0;
static double g_rec[POW_TAB_CAP], g_rec_key[POW_TAB_CAP];
static uint8_t g_rec_used[POW_TAB_CAP];
static int g_rec_n = 0;

static inline uint32_t pow_hash(double w) {
    uint64_t b;
    memcpy(&b, &w, 8);
    b ^= b >> 29;
    b *= 0x9E3779B97F4A7C15ull;
    return (uint32_t)(b >> 40) & (POW_TAB_CAP - 1);
}
void vpo_pow_table_clear(void) {
    memset(g_pow_used, 0, sizeof(g_pow_used));
    g_pow_n = 0;
}
int vpo_pow_table_set(const double *w, const double *w2, const double *w3, int n) {
    vpo_pow_table_clear();
    for (int i = 0; i < n; i++) {
        if (g_pow_n >= POW_TAB_CAP / 2) return -1;
        uint32_t h = pow_hash(w[i]);
        while (g_pow_used[h] && g_pw[h] != w[i]) h = (h + 1) & (POW_TAB_CAP - 1);
        if (!g_pow_used[h]) g_pow_n++;
        g_pow_used[h] = 1;
        g_pw[h] = w[i];
        g_pw2[h] = w2[i];
        g_pw3[h] = w3[i];
    }
    return 0;
}
void vpo_pow_record(int on) {
    g_pow_record = on;
    if (on) {
        g_rec_n = 0;
        memset(g_rec_used, 0, sizeof(g_rec_used));
    }
}
int vpo_pow_recorded(double *out, int cap) {
    int n = g_rec_n < cap ? g_rec_n : cap;
    memcpy(out, g_rec, sizeof(double) * (size_t)n);
    return g_rec_n;
}
static void pow23(double w, double *w2, double *w3) {
    if (g_pow_record) {
        uint32_t h = pow_hash(w);
        int seen = 0;
        while (g_rec_used[h]) {
            if (g_rec_key[h] == w) { seen = 1; break; }
            h = (h + 1) & (POW_TAB_CAP - 1);
        }
        if (!seen && g_rec_n < POW_TAB_CAP / 2) {
            g_rec_used[h] = 1;
            g_rec_key[h] = w;
            g_rec[g_rec_n++] = w;
        }
    }
    if (g_pow_n) {
        uint32_t h = pow_hash(w);
        while (g_pow_used[h]) {
            if (g_pw[h] == w) {
                *w2 = g_pw2[h];
                *w3 = g_pw3[h];
                return;
            }
            h = (h + 1) & (POW_TAB_CAP - 1);
        }
    }
    *w2 = pow(w, 2);
    *w3 = pow(w, 3);
}

void vpo_bicubic_weights(double w, double c[4]) {
    const double a = -0.75;
    double w2, w3;
    pow23(w, &w2, &w3);
    if (g_bicubic_fma) {
        /* as compiled for sm_100a (cuobjdump -sass of the unmodified reference, a folded to
         * -0.75): t=w2*1.5; q=w3*-0.75; c0=fma(w,-.75,t)+q; c1=fma(w3,1.25,fma(w2,-2.25,1));
         * c2=fma(w3,-1.25,fma(w,.75,t)); c3=fma(w2,-.75,-q) */
        double t = w2 * 1.5, q = w3 * -0.75;
        c[0] = fma(w, -0.75, t) + q;
        c[1] = fma(w3, 1.25, fma(w2, -2.25, 1.0));
        c[2] = fma(w3, -1.25, fma(w, 0.75, t));
        c[3] = fma(w2, -0.75, -q);
        return;
    }
    c[0] = a * w - 2 * a * w2 + a * w3;
    c[1] = 1 - (a + 3) * w2 + (a + 2) * w3;
    c[2] = -a * w + (2 * a + 3) * w2 - (a + 2) * w3;
    c[3] = a * w2 - a * w3;
}

static inline double dot4(const double c[4], double d0, double d1, double d2, double d3) {
    if (g_bicubic_fma) {
        /* SASS order: DMUL c1*d1; DFMA c0,d0; DFMA c2,d2; DFMA c3,d3 */
        double s = c[1] * d1;
        s = fma(c[0], d0, s);
        s = fma(c[2], d2, s);
        s = fma(c[3], d3, s);
        return s;
    }
    return ((c[0] * d0 + c[1] * d1) + c[2] * d2) + c[3] * d3;
}

static int bicubic_tap(const plane_t *d, int x, int y, int xDiff, int yDiff, int linesize, int width,
                       int height, double wx, double wy) {
    long s = (long)x + (long)y * linesize;
    int xDiffTop = xDiff, yDiffTop = yDiff;
    if (x + xDiff >= width) xDiff = 0;
    if (x + xDiff * 2 >= width) xDiff = 0;
    if (x - xDiffTop < 0) xDiffTop = 0;
    if (y + yDiff >= height) yDiff = 0;
    if (y + yDiff * 2 >= height) yDiff = 0;
    if (y - yDiffTop < 0) yDiffTop = 0;
    double cx[4], cy[4];
    vpo_bicubic_weights(wx, cx);
    vpo_bicubic_weights(wy, cy);
    long rowoff[4] = {-(long)linesize * yDiffTop, 0, (long)linesize * yDiff, 2L * linesize * yDiff};
    int b[4];
    for (int r = 0; r < 4; r++) {
        double v = dot4(cx, rd(d, s - xDiffTop + rowoff[r]), rd(d, s + rowoff[r]),
                        rd(d, s + xDiff + rowoff[r]), rd(d, s + 2 * xDiff + rowoff[r]));
        int bi = (int)round(v);
        bi = bi < 255 ? bi : 255;
        bi = bi > 0 ? bi : 0;
        b[r] = bi;
    }
    double v = dot4(cy, b[0], b[1], b[2], b[3]);
    int val = (int)round(v);
    val = val < 255 ? val : 255;
    val = val > 0 ? val : 0;
    return val;
}

/* reference src/Resize.cu:314-357 (resizeNV12BicubicKernel): coordinates in f32 (with the
 * FMA), widened to double afterwards */
static void resize_bicubic(const plane_t *Y, const plane_t *UV, int sw, int sh, int lsY, int lsUV,
                           nv12_t *o) {
    int dw = o->w, dh = o->h;
    float xr = (float)sw / dw, yr = (float)sh / dh;
    for (int i = 0; i < dh; i++) {
        double yF = (double)fmaf((float)(unsigned)i + 0.5f, yr, -0.5f);
        int y = (int)floor(yF);
        double wy = yF - y;
        if (y < 0) { y = 0; wy = 0; }
        if (y > sh - 1) { y = sh - 1; wy = 0; }
        for (int j = 0; j < dw; j++) {
            double xF = (double)fmaf((float)(unsigned)j + 0.5f, xr, -0.5f);
            int x = (int)floor(xF);
            double wx = xF - x;
            if (x < 0) { x = 0; wx = 0; }
            if (x > sw - 1) { x = sw - 1; wx = 0; }
            o->y[(long)i * dw + j] = (uint8_t)bicubic_tap(Y, x, y, 1, 1, lsY, sw, sh, wx, wy);
            if (i < dh / 2 && j < dw / 2) {
                o->uv[(long)i * dw + 2 * j] =
                    (uint8_t)bicubic_tap(UV, 2 * x, y, 2, 1, lsUV, sw, sh / 2, wx, wy);
                o->uv[(long)i * dw + 2 * j + 1] =
                    (uint8_t)bicubic_tap(UV, 2 * x + 1, y, 2, 1, lsUV, sw, sh / 2, wx, wy);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* Area (downscale) -- reference src/Resize.cu:359-386 (generateResizePattern), host f32.
 * Returns the number of pattern rows; rows are `stride`=ceil(scale) floats wide.          */
int vpo_area_pattern(float scale, float *out, int max_rows) {
    int stride = (int)ceilf(scale);
    int id = 0;
    float rest = 0;
    const float eps = 1.1920928955078125e-7f; /* std::numeric_limits<float>::epsilon() */
    while ((float)id * scale == 0 || ((float)id * scale - (float)(int)((float)id * scale) > eps)) {
        if (id >= max_rows) return -1;
        float dyn = scale;
        int n = 0;
        float *row = out + (long)id * stride;
        for (int k = 0; k < stride; k++) row[k] = 0;
        if (rest != 0) {
            if (n < stride) row[n] = rest;
            n++;
            dyn -= rest;
        }
        while (dyn - 1 > 0) {
            if (n < stride) row[n] = 1;
            n++;
            dyn -= 1.0f;
        }
        if (dyn > eps) {
            if (n < stride) row[n] = dyn;
            n++;
            rest = 1 - dyn;
        }
        /* n can exceed ceil(scale) (e.g. scale 2.4, third row): the reference's kernel reads
         * only ceil(scale) weights per row (Resize.cu:162-163), so the excess is dropped */
        id++;
    }
    return id;
}

/* reference src/Resize.cu:160-178 (calculateAreaInterpolation) */
static int area_tap(const plane_t *d, long start, float sx, float sy, int linesize, int stride,
                    const float *px, const float *py) {
    float sum = 0, div = 0;
    int rx = (int)ceilf(sx), ry = (int)ceilf(sy);
    for (int i = 0; i < ry; i++) {
        for (int j = 0; j < rx; j++) {
            long idx = start + (long)j * stride + (long)i * linesize;
            float w = px[j] * py[i];
            div += w;
            sum = fmaf((float)rd(d, idx), w, sum);
        }
    }
    sum /= div;
    return (int)sum;
}

/* reference src/Resize.cu:180-212 (resizeNV12DownscaleAreaKernel) */
static int resize_area_down(const plane_t *Y, const plane_t *UV, int sw, int sh, int lsY, int lsUV,
                            nv12_t *o) {
    int dw = o->w, dh = o->h;
    float xr = (float)sw / dw, yr = (float)sh / dh;
    int strideX = (int)ceilf(xr), strideY = (int)ceilf(yr);
    /* the kernel indexes the patterns with j % PX and i % PY (:187-188): a pattern at least as long as the
     * output never wraps, so generation may stop at dw / dh rows (the reference's loop only ends when
     * id*scale is integral in f32, which for scales like 2034/508 takes thousands of rows) */
    float *patX = (float *)malloc(sizeof(float) * (size_t)(dw + 1) * strideX);
    float *patY = (float *)malloc(sizeof(float) * (size_t)(dh + 1) * strideY);
    if (!patX || !patY) { free(patX); free(patY); return -1; }
    int PX = vpo_area_pattern(xr, patX, dw), PY = vpo_area_pattern(yr, patY, dh);
    if (PX < 0) PX = dw;
    if (PY < 0) PY = dh;
    if (PX <= 0 || PY <= 0) { free(patX); free(patY); return -1; }
    (void)sh;
    for (int i = 0; i < dh; i++) {
        float yF = (float)(int)(yr * (float)(unsigned)i);
        int y = (int)floorf(yF);
        const float *py = patY + (long)(i % PY) * strideY;
        for (int j = 0; j < dw; j++) {
            float xF = (float)(int)(xr * (float)(unsigned)j);
            int x = (int)floorf(xF);
            const float *px = patX + (long)(j % PX) * strideX;
            o->y[(long)i * dw + j] = (uint8_t)area_tap(Y, (long)y * lsY + x, xr, yr, lsY, 1, px, py);
            if (i < dh / 2 && j < dw / 2) {
                long idx = (long)y * lsUV + x * 2;
                o->uv[(long)i * dw + 2 * j] = (uint8_t)area_tap(UV, idx, xr, yr, lsUV, 2, px, py);
                o->uv[(long)i * dw + 2 * j + 1] = (uint8_t)area_tap(UV, idx + 1, xr, yr, lsUV, 2, px, py);
            }
        }
    }
    free(patX);
    free(patY);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* NV12 -> RGB -- reference src/ColorConversion.cu:6-39 (NV12toRGB24Kernel).  BT.601 limited
 * range; the FMA positions were read from sm_100a SASS of the unmodified reference
 * (SURVEY.md A.1).  Note both planes are indexed with the *same* pitch (:301).              */
static inline void yuv_to_rgb(int Yv, int U, int V, int *R, int *G, int *B) {
    float yp = fmaxf(0.f, (float)Yv - 16.f);
    float fu = (float)(U - 128), fv = (float)(V - 128);
    int r = (int)fmaf(yp, 1.163999557f, fmaf(fv, 1.5959997177f, 0.5f));
    int b = (int)fmaf(yp, 1.163999557f, fmaf(fu, 2.017999649f, 0.5f));
    float t = fu * -0.390999794f;
    t = fmaf(fv, -0.812999725f, t);
    t = t + 0.5f;
    int g = (int)fmaf(yp, 1.163999557f, t);
    *R = clampi(r, 0, 255);
    *G = clampi(g, 0, 255);
    *B = clampi(b, 0, 255);
}

void vpo_yuv_to_rgb(int Yv, int U, int V, int *rgb) { yuv_to_rgb(Yv, U, V, rgb, rgb + 1, rgb + 2); }

static inline void px_rgb(const plane_t *Y, const plane_t *UV, int i, int j, int pitch, int *R, int *G,
                          int *B) {
    int UVRow = i / 2;
    int UVCol = (j % 2 == 0) ? j : j - 1;
    int U = rd(UV, (long)UVRow * pitch + UVCol);
    int V = rd(UV, (long)UVRow * pitch + UVCol + 1);
    int Yv = rd(Y, (long)j + (long)i * pitch);
    yuv_to_rgb(Yv, U, V, R, G, B);
}

/* reference src/ColorConversion.cu:41-93 (NV12ToRGB24KernelPlanar / ...Merged).  For float
 * output the value is stored, then divided in place by 255 (IEEE division).                */
static void to_rgb(const plane_t *Y, const plane_t *UV, int w, int h, int pitch, int swapRB, int planar,
                   int norm, int as_float, void *dst) {
    uint8_t *d8 = (uint8_t *)dst;
    float *df = (float *)dst;
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            int c[3];
            px_rgb(Y, UV, i, j, pitch, &c[0], &c[1], &c[2]);
            if (swapRB) { int t = c[0]; c[0] = c[2]; c[2] = t; }
            for (int k = 0; k < 3; k++) {
                long idx = planar ? ((long)j + (long)i * w + (long)k * w * h) : ((long)j * 3 + (long)i * 3 * w + k);
                if (as_float) {
                    float v = (float)c[k];
                    if (norm) v /= 255;
                    df[idx] = v;
                } else {
                    /* uint8 with normalisation cannot happen: normalization selects T=float
                     * (VideoProcessor.cpp:139-142) */
                    d8[idx] = (uint8_t)c[k];
                }
            }
        }
}

/* reference src/ColorConversion.cu:95-105 (NV12ToY800) */
static void to_y800(const plane_t *Y, int w, int h, int pitch, int norm, int as_float, void *dst) {
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            int v = rd(Y, (long)j + (long)i * pitch);
            if (as_float) {
                float f = (float)v;
                if (norm) f /= 255;
                ((float *)dst)[(long)j + (long)i * w] = f;
            } else
                ((uint8_t *)dst)[(long)j + (long)i * w] = (uint8_t)v;
        }
}

/* reference src/ColorConversion.cu:107-127 (calculateUYVYChromaVertical) */
static int uyvy_chroma_vertical(const plane_t *UV, int i, int j, int width, int height) {
    int UVRow = i / 2, UVCol = j;
    int value = rd(UV, (long)UVCol + (long)UVRow * width);
    if (UVRow % 2 != 0) {
        int p1 = UVRow;
        int p2 = UVRow + 1; p2 = p2 < height / 2 - 1 ? p2 : height / 2 - 1;
        int p3 = UVRow - 1; p3 = p3 > 0 ? p3 : 0;
        int p4 = UVRow + 2; p4 = p4 < height / 2 - 1 ? p4 : height / 2 - 1;
        value = ((9 * (rd(UV, (long)p1 * width + UVCol) + rd(UV, (long)p2 * width + UVCol)) -
                  (rd(UV, (long)p3 * width + UVCol) + rd(UV, (long)p4 * width + UVCol)) + 8) >> 4);
        value = value < 255 ? value : 255;
        value = value > 0 ? value : 0;
    }
    return value;
}

/* reference src/ColorConversion.cu:177-209 (NV12ToUYVY) -- writes integer-valued elements into
 * dst (uint8 or float), 2*w*h of them */
static void to_uyvy(const plane_t *Y, const plane_t *UV, int w, int h, int pitch, int norm, int as_float,
                    void *dst) {
    uint8_t *d8 = (uint8_t *)dst;
    float *df = (float *)dst;
#define PUT(idx, v)                                                   \
    do {                                                              \
        if (as_float) { float f_ = (float)(v); if (norm) f_ /= 255; df[idx] = f_; } \
        else d8[idx] = (uint8_t)(v);                                  \
    } while (0)
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            long index = (long)j + (long)i * w;
            long isrc = (long)j + (long)i * pitch;
            if (index % 2 == 0) {
                long id = index * 2;
                int Uv = uyvy_chroma_vertical(UV, i, j, pitch, h) & 0xFF;
                PUT(id, Uv);
                PUT(id + 1, rd(Y, isrc));
                int Vv = uyvy_chroma_vertical(UV, i, j + 1, pitch, h) & 0xFF;
                PUT(id + 2, Vv);
            } else {
                PUT(index * 2 + 1, rd(Y, isrc));
            }
        }
#undef PUT
}

/* reference src/ColorConversion.cu:129-173 (calculateYUV444ChromaHorizontal / UYVYToYUV444).
 * `src` is the 2*w*h UYVY buffer living inside a 3*w*h allocation; indices >= 2*w*h read
 * uninitialised memory in the reference -- 0 here (the goldens reproduce with 0).          */
static void uyvy_to_yuv444_u8(const uint8_t *src, long nsrc, uint8_t *dst, int w, int h) {
    long wh = (long)w * h;
#define S(k) (((k) >= 0 && (k) < nsrc) ? (int)src[k] : 0)
    for (long index = 0; index < wh; index++) {
        long si = index * 2 + 1;
        dst[index] = (uint8_t)S(si);
        if (index % 2 == 0) {
            dst[wh + index] = (uint8_t)S(si - 1);
            dst[2 * wh + index] = (uint8_t)S(si + 1);
        } else {
            for (int sh = 0; sh <= 2; sh += 2) {
                long p1 = si - 3 + sh, p2 = si + 1 + sh, p3 = si - 7 + sh, p4 = si + 5 + sh;
                if (p3 < 0) p3 = p1;
                if (p4 > wh * 2 - 1) p4 = p2;
                /* int arithmetic, C division (toward zero), then narrowed to unsigned char
                 * *before* the (therefore ineffective) clamp */
                int v = (9 * (S(p1) + S(p2)) - (S(p3) + S(p4)) + 8) / 16;
                dst[(sh == 0 ? wh : 2 * wh) + index] = (uint8_t)v;
            }
        }
    }
#undef S
}

static void uyvy_to_yuv444_f32(const float *src, long nsrc, float *dst, int w, int h, int norm) {
    long wh = (long)w * h;
#define S(k) (((k) >= 0 && (k) < nsrc) ? src[k] : 0.0f)
    for (long index = 0; index < wh; index++) {
        long si = index * 2 + 1;
        float v = S(si);
        if (norm) v /= 255;
        dst[index] = v;
        if (index % 2 == 0) {
            float u = S(si - 1), vv = S(si + 1);
            if (norm) { u /= 255; vv /= 255; }
            dst[wh + index] = u;
            dst[2 * wh + index] = vv;
        } else {
            for (int sh = 0; sh <= 2; sh += 2) {
                long p1 = si - 3 + sh, p2 = si + 1 + sh, p3 = si - 7 + sh, p4 = si + 5 + sh;
                if (p3 < 0) p3 = p1;
                if (p4 > wh * 2 - 1) p4 = p2;
                float f = ((9 * (S(p1) + S(p2)) - (S(p3) + S(p4)) + 8) / 16);
                f = fminf(f, 255.f);
                f = fmaxf(f, 0.f);
                if (norm) f /= 255;
                dst[(sh == 0 ? wh : 2 * wh) + index] = f;
            }
        }
    }
#undef S
}

/* reference src/ColorConversion.cu:211-233 (NV12MergeBuffers) */
static void to_nv12_merged(const plane_t *Y, const plane_t *UV, int w, int h, int pitch, int norm,
                           int as_float, void *dst) {
    uint8_t *d8 = (uint8_t *)dst;
    float *df = (float *)dst;
    long wh = (long)w * h;
    /* odd w: the reference's last column writes its V one element past the row -- into the next chroma row
     * (rewritten by that row's first column; a race on the GPU) and, on the last row, past the
     * allocation.  Sequential order here; the out-of-bounds store is dropped. */
    long total = wh + (long)(h / 2) * w;
#define PUT(idx, v)                                                   \
    do {                                                              \
        if ((idx) >= total) break;                                    \
        if (as_float) { float f_ = (float)(v); if (norm) f_ /= 255; df[idx] = f_; } \
        else d8[idx] = (uint8_t)(v);                                  \
    } while (0)
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            PUT((long)j + (long)i * w, rd(Y, (long)j + (long)i * pitch));
            if (i % 2 == 0 && j % 2 == 0) {
                long iuv = (long)(i / 2) * w + j, isrc = (long)(i / 2) * pitch + j;
                PUT(wh + iuv, rd(UV, isrc));
                PUT(wh + iuv + 1, rd(UV, isrc + 1));
            }
        }
#undef PUT
}

/* reference src/ColorConversion.cu:235-278 (RGBMergedToHSVMerged<float>) on one pixel */
void vpo_rgb_to_hsv(float R, float G, float B, float *hsv) {
    float mn = fminf(fminf(R, G), B);
    float mx = fmaxf(fmaxf(R, G), B);
    float delta = mx - mn;
    float H = 0, S = 0;
    if (mx != 0) S = 1 - mn / mx;
    if (mx == mn) {
        H = 0;
    } else {
        if (R == mx && G >= B) H = 60 * (G - B) / delta;
        else if (R == mx && G < B) H = 60 * (G - B) / delta + 360;
        else if (G == mx) H = 60 * (B - R) / delta + 120;
        else if (B == mx) H = 60 * (R - G) / delta + 240;
        if (H < 0) H += 360;
        H /= 360;
    }
    hsv[0] = H;
    hsv[1] = S;
    hsv[2] = mx;
}

/* ------------------------------------------------------------------------------------ */
/* channelsByFourCC -- reference src/VideoProcessor.cpp:4-14 */
float vpo_channels(int fourcc) {
    if (fourcc == VPO_Y800) return 1;
    if (fourcc == VPO_UYVY) return 2;
    if (fourcc == VPO_NV12) return 1.5f;
    return 3;
}

/* Convert -- reference src/VideoProcessor.cpp:94-166 (dispatch rules) + Resize.cu:408-473 +
 * ColorConversion.cu:280-382 */
int vpo_convert(const uint8_t *y, const uint8_t *uv, int pitch_y, int pitch_uv, int width, int height,
                const vpo_params *p, void *out, int *out_w, int *out_h) {
    if (pitch_y == 0) pitch_y = width;   /* Resize.cu:420-421 / Crop.cu:37-38 */
    if (pitch_uv == 0) pitch_uv = width;
    plane_t Y = {y, (long)pitch_y * height};
    plane_t UV = {uv, (long)pitch_uv * (height / 2)};
    nv12_t cur = {0, 0, 0, 0}, nxt = {0, 0, 0, 0};
    int cw = p->crop_x1 - p->crop_x0, ch = p->crop_y1 - p->crop_y0;
    int crop = 0, resize = 0;
    int ow = 0, oh = 0;
    const plane_t *sY = &Y, *sUV = &UV;
    plane_t tY, tUV;
    int sw = width, sh = height, lsY = pitch_y, lsUV = pitch_uv;

    if (cw > 0 && ch > 0 && cw < width && ch < height) { /* VideoProcessor.cpp:109 */
        crop = 1;
        if (nv12_alloc(&cur, cw, ch)) return -1;
        crop_nv12(&Y, &UV, pitch_y, pitch_uv, p->crop_x0, p->crop_y0, p->crop_x1, p->crop_y1, &cur);
        ow = cw; oh = ch;
        tY.p = cur.y; tY.size = (long)cw * ch;
        tUV.p = cur.uv; tUV.size = (long)cw * (ch / 2);
        sY = &tY; sUV = &tUV; sw = cw; sh = ch; lsY = lsUV = cw;
    }
    if (p->dst_width && p->dst_height) { /* :118-130 */
        if (p->dst_width != sw || p->dst_height != sh) resize = 1;
        if (resize) {
            if (nv12_alloc(&nxt, p->dst_width, p->dst_height)) return -1;
            float xr = (float)sw / p->dst_width, yr = (float)sh / p->dst_height;
            int rc = 0;
            switch (p->resize_type) {
                case VPO_BILINEAR: resize_bilinear(sY, sUV, sw, sh, lsY, lsUV, &nxt); break;
                case VPO_NEAREST: resize_nearest(sY, sUV, sw, sh, lsY, lsUV, &nxt); break;
                case VPO_BICUBIC: resize_bicubic(sY, sUV, sw, sh, lsY, lsUV, &nxt); break;
                case VPO_AREA:
                    if (xr > 1 && yr > 1) rc = resize_area_down(sY, sUV, sw, sh, lsY, lsUV, &nxt);
                    else resize_area_up(sY, sUV, sw, sh, lsY, lsUV, &nxt);
                    break;
                default: rc = -2;
            }
            if (rc) { nv12_free(&cur); nv12_free(&nxt); return rc; }
            nv12_free(&cur);
            cur = nxt;
            ow = p->dst_width; oh = p->dst_height;
            tY.p = cur.y; tY.size = (long)ow * oh;
            tUV.p = cur.uv; tUV.size = (long)ow * (oh / 2);
            sY = &tY; sUV = &tUV; sw = ow; sh = oh; lsY = lsUV = ow;
        }
    }
    if (ow == 0 || oh == 0) { ow = width; oh = height; } /* :132-135 */

    /* colour conversion: pitchNV12 = src->linesize[0] ? : width, used for BOTH planes
     * (ColorConversion.cu:301) */
    int pitch = (resize || crop) ? ow : pitch_y;
    int as_float = p->normalization ? 1 : 0; /* VideoProcessor.cpp:139-142 */
    int norm = p->normalization ? 1 : 0;
    int w = ow, h = oh;
    int rc = 0;
    switch (p->fourcc) {
        case VPO_BGR24:
        case VPO_RGB24:
            to_rgb(sY, sUV, w, h, pitch, p->fourcc == VPO_BGR24, p->planes == VPO_PLANAR, norm, as_float, out);
            break;
        case VPO_Y800: to_y800(sY, w, h, pitch, norm, as_float, out); break;
        case VPO_UYVY: to_uyvy(sY, sUV, w, h, pitch, norm, as_float, out); break;
        case VPO_YUV444: {
            /* ColorConversion.cu:341-352: UYVY pass without normalisation into a 3*w*h buffer,
             * then the 4:4:4 pass */
            long n3 = 3L * w * h;
            if (as_float) {
                float *tmp = (float *)calloc((size_t)n3 + 16, sizeof(float));
                to_uyvy(sY, sUV, w, h, pitch, 0, 1, tmp);
                uyvy_to_yuv444_f32(tmp, n3, (float *)out, w, h, norm);
                free(tmp);
            } else {
                uint8_t *tmp = (uint8_t *)calloc((size_t)n3 + 16, 1);
                to_uyvy(sY, sUV, w, h, pitch, 0, 0, tmp);
                uyvy_to_yuv444_u8(tmp, n3, (uint8_t *)out, w, h);
                free(tmp);
            }
        } break;
        case VPO_NV12: to_nv12_merged(sY, sUV, w, h, pitch, norm, as_float, out); break;
        case VPO_HSV: {
            /* ColorConversion.cu:360-372: always float, always merged, from normalised RGB */
            float *o = (float *)out;
            for (int i = 0; i < h; i++)
                for (int j = 0; j < w; j++) {
                    int R, G, B;
                    px_rgb(sY, sUV, i, j, pitch, &R, &G, &B);
                    float r = (float)R / 255, g = (float)G / 255, b = (float)B / 255;
                    vpo_rgb_to_hsv(r, g, b, o + ((long)j * 3 + (long)i * w * 3));
                }
        } break;
        default: rc = 1; /* cudaErrorMissingConfiguration, ColorConversion.cu:376 */
    }
    nv12_free(&cur);
    if (out_w) *out_w = w;
    if (out_h) *out_h = h;
    return rc;
}

/* FFmpeg av_crc(AV_CRC_32_IEEE): MSB-first CRC-32, poly 0x04C11DB7, caller-supplied init,
 * no reflection, no final xor (reference tests/src/VPPTests.cpp:113) */
uint32_t vpo_av_crc32(uint32_t crc, const uint8_t *buf, size_t n) {
    static uint32_t tab[256];
    static int init = 0;
    if (!init) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i << 24;
            for (int k = 0; k < 8; k++) c = (c << 1) ^ ((c & 0x80000000u) ? 0x04C11DB7u : 0);
            /* FFmpeg stores the table byte-swapped for little-endian use */
            tab[i] = ((c >> 24) & 0xFF) | ((c >> 8) & 0xFF00) | ((c << 8) & 0xFF0000) | (c << 24);
        }
        init = 1;
    }
    while (n--) crc = tab[((uint8_t)crc) ^ *buf++] ^ (crc >> 8);
    return crc;
}
