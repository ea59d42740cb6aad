// vpp_march.h -- plan of the exact-integer separable ("column march") kernels, vpp_march_kernels.cu.
//
// When every resampling weight of a geometry is a dyadic rational -- 1080p -> 1280x720, 2^k x 2^k, 2x
// upscales ...: the weights of reference src/Resize.cu:45-48 (bicubic) and :359-386 (area patterns) are then
// small integers over a power of two -- the reference's FP64 / FP32 sums are EXACT (no rounding until its
// round() / division), so the same results come out of integer dot products (dp4a) bit for bit, at a
// fraction of the instruction count of the float paths in vpp_kernels.cu.  The plan builder
// (tsvpp_capi.cpp: build_march) checks the exactness conditions on the actual weight values (the bicubic
// ones come from the device's own pow()) and falls back to the float kernels otherwise.
//
// Work decomposition: item = (frame, strip of R = 2*rpg*G output rows, tile of tw output columns).  The
// producer warp stages every DISTINCT source row a strip touches once (slot s = s-th needed row of the strip),
// consumers are G row groups x tw/2 threads; a thread owns two adjacent output columns and walks down the
// rpg row pairs of its group.
//   bicubic: horizontal 4-tap sums (rounded and clamped to a byte, as the reference does between its two
//            passes, src/Resize.cu:59) are evaluated once per (source row, output column) and kept in a 4-byte
//            register window; the vertical 4-tap sum is one dot product of that window.
//   area:    per source row of the window a horizontal dot of KX bytes with the integer pattern row, weighted
//            into an integer accumulator; one exact division at the end (src/Resize.cu:160-178).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "vpp_plan.h"

namespace tsvpp {

enum MarchMode : int { MARCH_BICUBIC = 0, MARCH_AREA = 1, MARCH_BILINEAR = 2 };

// output columns per consumer thread: bilinear 4 (12-byte / float4 row stores), the others 2
#ifdef __CUDACC__
__host__ __device__
#endif
constexpr int march_bw(int mode) { return mode == MARCH_BILINEAR ? 4 : 2; }

#ifndef TSVPP_MARCH_MAXTHREADS
#define TSVPP_MARCH_MAXTHREADS 288
#endif
constexpr int kMarchMaxThreads = TSVPP_MARCH_MAXTHREADS;  // 1 producer warp + up to 256 consumers

// bicubic: one record per luma column, and one per chroma PAIR (U and V share position and weights)
struct MarchColB {
    int addr;       // byte offset (multiple of 4) of the first word of the tap window inside a staged row
    unsigned sel0;  // luma: PRMT selector picking the 4 taps out of {w0,w1}.
                    // chroma: selA of U (low 16 bits) and of V (high 16): picks from {w0,w1}
    unsigned sel1;  // chroma: selB of U / V: merges the first pick with w2
    int exact;      // 1: the FP64 weights ARE m / 2^q; 0: they are within 2^-44 of it (libdevice pow is not correctly
                    // rounded): sums that land exactly on a .5 tie are then re-evaluated with the FP64 chain
    int w[4];       // weight digits, base 128 (short: w[0] low, w[1] high; long: w[0..1] = (m & 16383), w[2..3] = m >> 14)
};
// area: one record per luma column and one per chroma pair
struct MarchColA {
    int addr;    // byte offset (multiple of 4) of the first word
    int shift;   // 8 * (first tap - addr): funnel-shift amount that aligns the window to a word boundary
    int prow;    // horizontal pattern row (j % PX), selects the weight words and, with the row's pattern row, the divisor
    int pad;
};
// bicubic: per output row (luma rows and chroma rows have separate tables).  Windows are kept newest-first:
// byte 0 = the slot the group has marched to when this row is evaluated.
struct MarchRowB {
    unsigned meta;  // bits 0-11: strip-relative slot the window must have been marched to; bits 12-27: PRMT selector
                    // that rebuilds the four taps (in tap order) from the window, for the FP64 tie path; bit 28: exact
    int pad0;
    int w[4];       // vertical weight digits folded onto the window bytes (w[0..1] sit in the first 16 bytes)
    int pad1[2];
};
// bilinear (also AREA when it is not a strict downscale, reference src/Resize.cu:214-240): the horizontal two-tap sum
// is kept UNROUNDED (A*(2^ex - mx) + B*mx < 2^15), the vertical one is shifted once at the end -- the reference
// truncates a single exact f32 expression (src/Resize.cu:5-25), so there is no intermediate rounding to mirror.
struct MarchColL {  // one luma column, or one chroma pair
    int addr;       // byte offset (multiple of 4) of the first of the two words the taps lie in
    unsigned sel0;  // luma: PRMT selector putting tap 0 / tap 1 into bytes 0 / 1.  chroma: the same for U
    unsigned sel1;  // chroma: for V
    unsigned w;     // byte 0 = weight of tap 0 (2^ex - mx, or 2^ex when the taps coincide), byte 1 = weight of tap 1
};
struct MarchRowL {
    unsigned packed;  // bits 0-15: strip-relative slot the two-row window must have been marched to;
                      // bits 16-23: weight of the older row of the window (0 .. 2^ey <= 128); bits 24-31: of the newest
};
// area: per output row
struct MarchRowA {
    int first;  // strip-relative slot of the first window row
    int prow;   // vertical pattern row (i % PY)
};

struct MarchPlanDev {
    PlanDev e;               // the fields emit_block reads: dw, dh, cw2, ch2, swap_rb (everything else zero)
    int mode;                // MarchMode
    int tw, ntiles, nstrips;
    int G, rpg;              // row groups per strip, row pairs per group; strip = 2*G*rpg luma rows
    int nsl, nsc;            // slots per stage: luma, chroma (maxima over strips)
    int slot_stride, stage_bytes, nstages;
    int pitch_y, pitch_uv;
    int need_uv;
    int vshift;              // bilinear: ex + ey
    // bicubic
    int xlong, ylong;        // weight digits: 0 = short (q <= 13), 1 = long
    int qx, qy;              // weights are m / 2^q
    // area
    int K, KX, NW;           // window rows, window columns, weight words per pattern row (ceil(KX/4) rounded to 1,2,4)
    int PX, PY;
    // tables
    const int *seg_a0_l, *seg_len_l, *seg_a0_c, *seg_len_c;  // [ntiles]
    const int *slotrow_l, *slotrow_c;                        // [nstrips][nsl] / [nstrips][nsc]: source row of each slot
    const int *nslots_l, *nslots_c;                          // [nstrips]
    const void *col_l, *col_c;                               // [ntiles*tw] / [ntiles*tw/2]
    const void *row_l, *row_c;                               // [nstrips*R] / [nstrips*R/2]
    const int *gfirst_l, *gfirst_c;                          // bicubic: [nstrips*G] slot of the oldest tap of a group's first row
    const double *cxd;                                       // bicubic: [dw][4] the reference's FP64 horizontal weights
    const double *cyd;                                       // bicubic: [dh][4] vertical ones (row i; chroma row ci uses row ci)
    const int *px_w;                                         // area: [PX][NW] packed u8 weights
    const int *py_w;                                         // area: [PY][K] integer weights
    const unsigned *div_mul;                                 // area: [PY][PX] magic multiplier of the divisor
    const int *div_shift;                                    // area: [PY][PX]
};

int march_block_threads(const MarchPlanDev &P);
size_t march_smem_bytes(const MarchPlanDev &P);
int configure_march(const MarchPlanDev &P, int ok, size_t smem_optin_max);
int march_max_ctas_per_sm(const MarchPlanDev &P, int ok, int threads, size_t smem_bytes);
int launch_march(const MarchPlanDev &P, int ok, const FrameDev *frames_dev, int nframes, int vec_ok, int grid,
                 size_t smem_bytes, cudaStream_t stream);

}  // namespace tsvpp
