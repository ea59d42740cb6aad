// vpp_plan.h -- geometry plan shared by the host plan builder and the fused sm_100a kernels.
//
// A plan is everything about a (source geometry, FrameParameters) pair that does not depend on
// pixel data: per-output-column tap offsets and weights, per-output-row source rows and weights,
// the tiling of the output into work items and the shared-memory stage layout.  The reference
// recomputes all of this per pixel per frame (src/Resize.cu:269-357) and re-uploads the AREA
// pattern tables per frame (src/Resize.cu:437-452); here it is built once and cached.
#pragma once
#include <cstdint>

namespace tsvpp {

// which resampling arithmetic Phase A runs
enum ResampleKind : int {
    RS_NEAREST = 0,   // also "no resize" (identity tables) -- reference Resize.cu:242-267
    RS_BILINEAR = 1,  // also AREA when not a strict downscale -- Resize.cu:269-312 / :214-240
    RS_BICUBIC = 2,   // Resize.cu:314-357, 27-91 (FP64)
    RS_AREA = 3,      // strict downscale, pattern tables -- Resize.cu:180-212, 160-178
    RS_COUNT = 4
};

// which colour conversion / layout Phase B runs
enum OutKind : int {
    OUT_RGB_PLANAR_U8 = 0,
    OUT_RGB_PLANAR_F32 = 1,
    OUT_RGB_MERGED_U8 = 2,
    OUT_RGB_MERGED_F32 = 3,
    OUT_Y800_U8 = 4,
    OUT_Y800_F32 = 5,
    OUT_HSV_F32 = 6,
    OUT_NV12_U8 = 7,
    OUT_NV12_F32 = 8,
    OUT_COUNT = 9
};

constexpr int kMaxThreads = 288;    // 1 producer warp + up to 256 consumer threads (one 2x2 block each)
constexpr int kLdgProducerWarps = 4;  // unaligned frames (shifted bulk fill): the slots are shared by 4 producer warps
constexpr int kMaxStages = 8;       // mbarrier ring depth upper bound

// column parameters, one per output sample column (luma: per output x; chroma: per chroma byte)
struct ColNearest { int o; };                                  // smem byte offset inside a slot
struct ColBilinear { int o0, o1; float w; int pad; };          // taps at o0 / o1, weight of o1
// Keys a=-0.75 taps; c = the reference's FP64 weights, cf = the same rounded to f32, exact = all four
// are multiples of 2^-12 (then the f32 evaluation of a row sum is exact, ties included)
struct ColBicubic { int o[4]; double c[4]; float cf[4]; int exact; int pad; };
struct ColArea { int o; int prow; };                           // first tap, pattern row index

struct FrameDev {
    const uint8_t *y;
    const uint8_t *uv;
    void *out;
};

// Device-side view of a plan (passed to the kernel by value).
struct PlanDev {
    int dw, dh;             // output size
    int cw2, ch2;           // number of chroma pairs per row (dw/2) and chroma rows (dh/2) that exist
    int tw;                 // tile width in output columns (32..256, multiple of 32)
    int ntiles, nstrips;    // work items per frame = nstrips * ntiles
    int rp;                 // row pairs per strip (1 or 2): a strip is 2*rp luma rows, rp chroma rows
    int nprod;              // producer warps: 1 (aligned frames) or kLdgProducerWarps (unaligned frames, shifted bulk fill)
    int bw;                 // output columns per consumer thread (2 or 4); a thread owns 2 rows x bw columns
    int K;                  // vertical taps = slots per output row
    int KX;                 // horizontal taps (AREA only)
    int slot_stride;        // bytes between slots (multiple of 16)
    int stage_bytes;        // 3*rp*K*slot_stride (2*rp*K for luma-only outputs)
    int nstages;            // ring depth
    int pitch_y, pitch_uv;  // source pitches actually used for addressing
    long plane_bytes_y, plane_bytes_uv;  // extents for the guarded (non-bulk) fill
    int swap_rb;            // BGR
    int need_uv;            // 0 for Y800
    int ident;              // nearest only: every aligned group of bw columns taps bw consecutive, bw-aligned staged
                            // bytes (no horizontal resize): a thread reads its row of a block with one shared load
    int PX, PY;             // AREA pattern periods
    // direct (unstaged) path: no resize, so source rows/columns are affine in the output coordinates
    int direct;             // 1: the plan may run vpp_direct_kernel
    int d_row_l0, d_row_c0; // source row of output luma row 0 / chroma row 0
    int d_x_l0, d_x_c0;     // source byte of output column 0 in a luma / chroma row (multiples of 4)
    unsigned d_div_mul;     // n / (dw/4) == __umulhi(n, d_div_mul) >> d_div_shift for every block index n of a frame
    int d_div_shift;
    // tables (device pointers)
    const int *seg_a0_l;    // [ntiles] byte offset of the staged segment inside a luma row
    const int *seg_len_l;   // [ntiles] bytes (multiple of 16)
    const int *seg_a0_c;    // [ntiles] same for the chroma plane
    const int *seg_len_c;
    const int *rowsrc_l;    // [nstrips*2*rp][K] source row (luma plane, crop included)
    const int *rowsrc_c;    // [nstrips*rp][K]   source row (chroma plane)
    const void *col_l;      // [ntiles*tw] Col* for luma columns
    const void *col_c;      // [ntiles*tw] Col* for chroma bytes
    const float *wy;        // [nstrips*2*rp] bilinear vertical weight per output row index
    const double *cy;       // [nstrips*2*rp][4] bicubic vertical weights per output row index
    const float *cyf;       // same, rounded to f32
    const int *cy_exact;    // [nstrips*2*rp] 1 if the row's four weights are multiples of 2^-12
    const int *row_shift;   // [nstrips*rp] bicubic: source-row window of luma row 2p+1 = that of row 2p shifted by
                            // this many rows (0..3), 4 = no reuse; lets row 2p+1 reuse row 2p's horizontal sums
    const float *px;        // [PX][KX] AREA horizontal pattern
    const float *py;        // [PY][K]  AREA vertical pattern
    const float *area_div;  // [PY][PX] sum of px[b]*py[a] in the kernel's order (rows outer, columns inner)
};

}  // namespace tsvpp
