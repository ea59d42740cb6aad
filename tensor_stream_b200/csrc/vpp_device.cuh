// vpp_device.cuh -- device-side helpers shared by the sm_100a VPP kernels (vpp_kernels.cu: row-slot
// staging, f32 resamplers; vpp_march_kernels.cu: distinct-row staging, exact-integer separable resamplers):
// mbarrier / bulk-copy PTX, exact conversions, the reference's colour arithmetic, block stores.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <stdint.h>

#include "vpp_launch.h"
#include "vpp_plan.h"

namespace tsvpp {

// ---------------------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk async copy (TMA unit, descriptor-free 1-D form)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; size and both addresses must be multiples of 16 bytes.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                         uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
        : "memory");
}
// The same copy issued by ONE elected lane of a converged warp (all 32 lanes must execute this with
// identical operands): an elect.sync predicate is known to be single-lane, so the copy goes straight
// to the uniform datapath instead of through a loop over the active lanes.
__device__ __forceinline__ void bulk_g2s_elected(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                                 uint64_t policy) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n\t}"
        ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// ---------------------------------------------------------------------------------------
// Exact small-integer <-> float conversions on the FMA/ALU pipes
// ---------------------------------------------------------------------------------------
constexpr uint32_t kMagicBits = 0x4B000000u;  // 2^23 as float bits
constexpr float kMagic = 8388608.0f;

// (float)b for 0 <= b < 2^23, exact
__device__ __forceinline__ float u2f(uint32_t b) {
#ifdef TSVPP_U2F_MAGIC
    return __fsub_rn(__uint_as_float(kMagicBits | b), kMagic);
#else
    // cvt on a 32-bit register compiles to I2FP.F32.U32 (one full-rate instruction); the u8/u16
    // forms the compiler picks for (float)uchar are the quarter-rate XU-pipe I2F the first version
    // was bound by.  Measured equal-or-faster than the 2-instruction magic form (A/B, round 1).
    float f;
    asm("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(b));
    return f;
#endif
}

// (int)v clamped to [0,255] in ONE instruction (SASS F2IP.U8.F32.TRUNC): a float->u8 conversion
// truncates toward zero and saturates to the destination range (PTX cvt semantics), i.e. exactly
// min(max((int)v,0),255) of reference src/ColorConversion.cu:26-38 -- and exactly (unsigned char)(int)v
// for the resamplers, whose results lie in [0,256).  Replaces FADD.RZ + 2 IMNMX + LOP.
__device__ __forceinline__ uint32_t f2u8_sat(float v) {
    uint32_t r;
    asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

// ---------------------------------------------------------------------------------------
// Arithmetic shared with the oracle (each cites the reference lines it reproduces)
// ---------------------------------------------------------------------------------------

// reference src/Resize.cu:5-25 (calculateBillinearInterpolation), FMA positions as compiled:
//   v = fma(A*(1-wx), (1-wy), (B*wx)*(1-wy)); v = fma((1-wx), C*wy, v); v = fma(D, wx*wy, v)
// and (int)v stored to unsigned char (v in [0, 256)).
__device__ __forceinline__ int bilinear_px(uint32_t a, uint32_t b, uint32_t c, uint32_t d, float wx, float wy) {
    const float A = u2f(a), B = u2f(b), C = u2f(c), D = u2f(d);
    const float omx = __fsub_rn(1.0f, wx), omy = __fsub_rn(1.0f, wy);
    float v = __fmaf_rn(__fmul_rn(A, omx), omy, __fmul_rn(__fmul_rn(B, wx), omy));
    v = __fmaf_rn(omx, __fmul_rn(C, wy), v);
    v = __fmaf_rn(D, __fmul_rn(wx, wy), v);
    return (int)f2u8_sat(v);
}

// reference src/Resize.cu:45-59: one 4-tap row of the bicubic, DFMA chain as compiled
// (DMUL c1*d1; DFMA c0,d0; DFMA c2,d2; DFMA c3,d3), then round-half-away and clamp.
__device__ __forceinline__ int bicubic_dot(const double c0, const double c1, const double c2, const double c3,
                                           int d0, int d1, int d2, int d3) {
    double s = __dmul_rn(c1, (double)d1);
    s = __fma_rn(c0, (double)d0, s);
    s = __fma_rn(c2, (double)d2, s);
    s = __fma_rn(c3, (double)d3, s);
    int b = (int)round(s);
    b = min(b, 255);
    b = max(b, 0);
    return b;
}

// The same row sum, evaluated in float32 first.  With t32 = fl(s32 + .5):
//   |t32 - (s_real + .5)| <= 2^-24 * 255 * sum|c_k|   (weights rounded to f32, sum|c_k| <= 1.3)   2.0e-5
//                          + 2^-17                      (fmul, |c1*d1| < 256)                        0.8e-5
//                          + 4 * 2^-16                  (three fma and the +.5, magnitudes < 512)    6.1e-5
//                         <  8.9e-5 < 2^-13,
// and the FP64 chain is within 1e-12 of s_real.  So whenever frac(t32) is farther than 2^-12 from an
// integer, floor(t32) IS round(s64).  Otherwise -- or never, when the weights are multiples of 2^-12
// and the f32 sum is exact, ties included -- fall back to the reference's FP64 chain.  Bit-exact by
// construction; the window only sets how often the slow path runs (a tie-prone 1.5x resize lands in
// it for 1/256 + 2^-11 of the dots instead of 1/256 + 2^-8 with the 2^-9 window of round 1).
__device__ __forceinline__ int bicubic_dot_fast(const float *cf, const double *cd, int exact, uint32_t d0, uint32_t d1,
                                                uint32_t d2, uint32_t d3) {
    float s = __fmul_rn(cf[1], u2f(d1));
    s = __fmaf_rn(cf[0], u2f(d0), s);
    s = __fmaf_rn(cf[2], u2f(d2), s);
    s = __fmaf_rn(cf[3], u2f(d3), s);
    const float t = __fadd_rn(s, 0.5f);
    const uint32_t r = f2u8_sat(t);  // floor(t) for 0 <= t < 256, saturated outside: clamp(round-half-up(s))
    if (!exact) {
        const float frac = __fsub_rn(t, u2f(r));  // in [0,1) wherever the rounding can matter
        const float eps = 0.000244140625f;        // 2^-12
        if (t >= 0.0f && (frac < eps || frac > 1.0f - eps))
            return bicubic_dot(cd[0], cd[1], cd[2], cd[3], (int)d0, (int)d1, (int)d2, (int)d3);
    }
    return (int)r;
}

// reference src/ColorConversion.cu:6-39 (NV12toRGB24Kernel); chroma terms are shared by the
// pixels of a 2x2 block.  R=(int)fma(Y',1.164,fma(V,1.596,.5)) etc..
struct ChromaTerms {
    float r, g, b;
};
// `Um`, `Vm`: the samples as magic-float bits (kMagicBits | value)
__device__ __forceinline__ ChromaTerms chroma_terms(uint32_t Um, uint32_t Vm) {
    // (float)(U-128): exact via the magic form with a folded bias (2^23 + 128)
    const float fu = __fsub_rn(__uint_as_float(Um), kMagic + 128.0f);
    const float fv = __fsub_rn(__uint_as_float(Vm), kMagic + 128.0f);
    ChromaTerms t;
    t.r = __fmaf_rn(fv, 1.5959997177f, 0.5f);
    t.b = __fmaf_rn(fu, 2.017999649f, 0.5f);
    float g = __fmul_rn(fu, -0.390999794f);
    g = __fmaf_rn(fv, -0.812999725f, g);
    t.g = __fadd_rn(g, 0.5f);
    return t;
}
// one pixel -> the three channels, each already clamped to 0..255.  `Ym` = kMagicBits | Y.
__device__ __forceinline__ void yuv_px_bits(uint32_t Ym, const ChromaTerms &t, int &R, int &G, int &B) {
    // max(0.f, (float)Y - 16.f): (2^23 + Y) - (2^23 + 16) is exact, one FADD on the FMA pipe
    const float yp = fmaxf(__fsub_rn(__uint_as_float(Ym), kMagic + 16.0f), 0.0f);
    R = (int)f2u8_sat(__fmaf_rn(yp, 1.163999557f, t.r));
    G = (int)f2u8_sat(__fmaf_rn(yp, 1.163999557f, t.g));
    B = (int)f2u8_sat(__fmaf_rn(yp, 1.163999557f, t.b));
}
__device__ __forceinline__ uint32_t bits_u8(int c) { return (uint32_t)c; }

// c / 255.0f with IEEE rounding for c in 0..255 given as magic bits: reciprocal multiply plus one
// residual correction; verified exhaustively against the division (tests/test_host_logic.py).
__device__ __forceinline__ float div255_x(float x) {
    const float r = 0x1.010102p-8f;  // RN(1/255)
    const float q = __fmul_rn(x, r);
    const float e = __fmaf_rn(-q, 255.0f, x);
    return __fmaf_rn(e, r, q);
}
__device__ __forceinline__ float div255_bits(int c) { return div255_x(u2f((uint32_t)c)); }
__device__ __forceinline__ float div255_u8(uint32_t c) { return div255_x(u2f(c)); }

// reference src/ColorConversion.cu:235-278 (RGBMergedToHSVMerged<float>)
__device__ __forceinline__ void rgb_to_hsv(float R, float G, float B, float &H, float &S, float &V) {
    const float mn = fminf(fminf(R, G), B);
    const float mx = fmaxf(fmaxf(R, G), B);
    const float delta = __fsub_rn(mx, mn);
    V = mx;
    S = 0.0f;
    if (mx != 0.0f) S = __fsub_rn(1.0f, __fdiv_rn(mn, mx));
    H = 0.0f;
    if (mx == mn) return;
    if (R == mx && G >= B)
        H = __fdiv_rn(__fmul_rn(60.0f, __fsub_rn(G, B)), delta);
    else if (R == mx && G < B)
        H = __fadd_rn(__fdiv_rn(__fmul_rn(60.0f, __fsub_rn(G, B)), delta), 360.0f);
    else if (G == mx)
        H = __fadd_rn(__fdiv_rn(__fmul_rn(60.0f, __fsub_rn(B, R)), delta), 120.0f);
    else if (B == mx)
        H = __fadd_rn(__fdiv_rn(__fmul_rn(60.0f, __fsub_rn(R, G)), delta), 240.0f);
    if (H < 0.0f) H = __fadd_rn(H, 360.0f);
    H = __fdiv_rn(H, 360.0f);
}

// ---------------------------------------------------------------------------------------
// Stage fill (producer warp)
// ---------------------------------------------------------------------------------------
struct ItemCoord {
    int frame, strip, tile;
};
__device__ __forceinline__ ItemCoord decode_item(const PlanDev &P, int item) {
    ItemCoord c;
    const int per_frame = P.nstrips * P.ntiles;
    c.frame = item / per_frame;
    const int r = item - c.frame * per_frame;
    c.strip = r / P.ntiles;
    c.tile = r - c.strip * P.ntiles;
    return c;
}

// Work items of one CTA advance by gridDim.x: keep (frame, strip, tile) as a mixed-radix counter
// instead of dividing every iteration.
struct ItemIter {
    int frame, strip, tile;
    int d_frame, d_strip, d_tile;
    __device__ __forceinline__ void init(const PlanDev &P, int first, int step) {
        const ItemCoord c = decode_item(P, first);
        frame = c.frame; strip = c.strip; tile = c.tile;
        d_tile = step % P.ntiles;
        const int q = step / P.ntiles;
        d_strip = q % P.nstrips;
        d_frame = q / P.nstrips;
    }
    __device__ __forceinline__ void next(const PlanDev &P) {
        tile += d_tile;
        int carry = tile >= P.ntiles ? 1 : 0;
        tile -= carry ? P.ntiles : 0;
        strip += d_strip + carry;
        carry = strip >= P.nstrips ? 1 : 0;
        strip -= carry ? P.nstrips : 0;
        frame += d_frame + carry;
    }
};

// One bulk copy per (output row of the strip, vertical tap).  Every operand of the copies is
// warp-uniform (item coordinates, read-only table entries at uniform indices), so the copies are
// issued from the uniform datapath by one lane without a per-lane serialisation loop.
template <bool NEED_UV>
__device__ __forceinline__ void issue_stage_bulk(const PlanDev &P, const FrameDev *__restrict__ frames,
                                                 const ItemIter &ic, unsigned char *stage, uint64_t *bar, int lane,
                                                 uint64_t policy, int pwarp, int npwarps) {
    const unsigned long long *fd = reinterpret_cast<const unsigned long long *>(frames + ic.frame);
    const unsigned char *fy = reinterpret_cast<const unsigned char *>(__ldg(fd));
    const unsigned char *fuv = reinterpret_cast<const unsigned char *>(__ldg(fd + 1));
    const int nl = 2 * P.rp * P.K;            // luma slots
    const int nc = NEED_UV ? P.rp * P.K : 0;  // chroma slots
    const int a0l = __ldg(P.seg_a0_l + ic.tile), lenl = __ldg(P.seg_len_l + ic.tile);
    const int a0c = NEED_UV ? __ldg(P.seg_a0_c + ic.tile) : 0, lenc = NEED_UV ? __ldg(P.seg_len_c + ic.tile) : 0;
    const int *rs_l = P.rowsrc_l + (size_t)ic.strip * nl;
    const int *rs_c = P.rowsrc_c + (size_t)ic.strip * nc;
    // slots are dealt round-robin to the producer warps; each announces the bytes of its own copies
    const int myl = (nl - pwarp + npwarps - 1) / npwarps, myc = (nc - pwarp + npwarps - 1) / npwarps;
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)(myl * lenl + myc * lenc));
    __syncwarp();
    fy += a0l;
    fuv += a0c;
#pragma unroll 4
    for (int c = pwarp; c < nl; c += npwarps) {
        const int row = __ldg(rs_l + c);
        bulk_g2s_elected(stage + (size_t)c * P.slot_stride, fy + (size_t)row * P.pitch_y, lenl, bar, policy);
    }
    if (NEED_UV) {
        unsigned char *sc = stage + (size_t)nl * P.slot_stride;
#pragma unroll 2
        for (int c = pwarp; c < nc; c += npwarps) {
            const int row = __ldg(rs_c + c);
            bulk_g2s_elected(sc + (size_t)c * P.slot_stride, fuv + (size_t)row * P.pitch_uv, lenc, bar, policy);
        }
    }
}

// Guarded row copy: the warp copies one row segment with aligned 32-bit global loads + funnel shifts (any source
// alignment; bytes outside the plane read as 0).  Used by fill_stage_shifted for the rows at the edges of a plane.
__device__ __forceinline__ uint32_t ld_word_guarded(const unsigned char *plane, long off, long plane_bytes) {
    // the 4 bytes plane[off .. off+3]; `plane + off` is 4-byte aligned
    if (off >= 0 && off + 4 <= plane_bytes) return __ldg(reinterpret_cast<const uint32_t *>(plane + off));
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long o = off + k;
        if (o >= 0 && o < plane_bytes) v |= (uint32_t)__ldg(plane + o) << (8 * k);
    }
    return v;
}
__device__ __forceinline__ void copy_row_ldg(unsigned char *dst, const unsigned char *plane, long base, int len,
                                             long plane_bytes, int lane) {
    const int mis = (int)((reinterpret_cast<uintptr_t>(plane) + (uintptr_t)base) & 3);
    const long abase = base - mis;  // 4-byte aligned address relative to the plane pointer
    uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);
    const int nwords = len >> 2;
    for (int k = lane; k < nwords; k += 32) {
        const uint32_t lo = ld_word_guarded(plane, abase + 4L * k, plane_bytes);
        const uint32_t hi = mis ? ld_word_guarded(plane, abase + 4L * k + 4, plane_bytes) : 0u;
        d32[k] = __funnelshift_r(lo, hi, mis * 8);
    }
}

// Fill for frames whose base / pitch are NOT 16-byte aligned (tight odd-width clips, sliced tensors): still the TMA
// unit.  A bulk copy needs 16-byte aligned source, destination and size, so every row segment is fetched as its
// 16-byte aligned SUPERSET [src & ~15, round_up(src + len, 16)) into the slot, and the consumers are told where the
// segment starts inside the slot (`shifts`, one byte per slot, written before the barrier arrive).  Only rows whose
// superset would leave the plane (the first bytes of the first row, the last bytes of the last one) are copied by the
// warp with guarded loads.  Slots are dealt round-robin to the producer warps; the "full" barrier counts one arrival
// per producer warp plus the bulk bytes.
struct ShiftedRow {  // where a slot's row segment lives and how its aligned superset is fetched
    const unsigned char *plane;
    long base, pbytes, s0;
    int len, mis, nbytes;
    bool bulk;
};
template <bool NEED_UV>
__device__ __forceinline__ ShiftedRow shifted_row(const PlanDev &P, const FrameDev &fr, const int *rs_l, const int *rs_c,
                                                  int nl, int a0l, int lenl, int a0c, int lenc, int c) {
    ShiftedRow r;
    const bool luma = !NEED_UV || c < nl;
    r.plane = luma ? fr.y : fr.uv;
    r.base = luma ? (long)rs_l[c] * P.pitch_y + a0l : (long)rs_c[c - nl] * P.pitch_uv + a0c;
    r.len = luma ? lenl : lenc;
    r.pbytes = luma ? P.plane_bytes_y : P.plane_bytes_uv;
    r.mis = (int)((reinterpret_cast<uintptr_t>(r.plane) + (uintptr_t)r.base) & 15);
    r.s0 = r.base - r.mis;
    r.nbytes = (r.mis + r.len + 15) & ~15;
    r.bulk = r.s0 >= 0 && r.s0 + r.nbytes <= r.pbytes;
    return r;
}
template <bool NEED_UV>
__device__ __forceinline__ void fill_stage_shifted(const PlanDev &P, const FrameDev *__restrict__ frames,
                                                   const ItemIter &ic, unsigned char *stage, unsigned char *shifts,
                                                   uint64_t *bar, int lane, uint64_t policy, int pwarp, int npwarps) {
    const FrameDev fr = frames[ic.frame];
    const int nl = 2 * P.rp * P.K;
    const int nc = NEED_UV ? P.rp * P.K : 0;
    const int a0l = P.seg_a0_l[ic.tile], lenl = P.seg_len_l[ic.tile];
    const int a0c = NEED_UV ? P.seg_a0_c[ic.tile] : 0, lenc = NEED_UV ? P.seg_len_c[ic.tile] : 0;
    const int *rs_l = P.rowsrc_l + (size_t)ic.strip * nl;
    const int *rs_c = P.rowsrc_c + (size_t)ic.strip * nc;
    const int nslots = nl + nc;
    // every lane takes one slot of this warp's share (slot = pwarp + npwarps * (32 * round + lane)); rows whose
    // superset would leave the plane are copied by the whole warp with guarded loads, the others are announced to the
    // barrier and fetched by one bulk copy per lane
    uint32_t bytes = 0;
    for (int c0 = pwarp; c0 < nslots; c0 += 32 * npwarps) {
        const int c = c0 + npwarps * lane;
        ShiftedRow r;
        r.bulk = false;
        if (c < nslots) {
            r = shifted_row<NEED_UV>(P, fr, rs_l, rs_c, nl, a0l, lenl, a0c, lenc, c);
            shifts[c] = r.bulk ? (unsigned char)r.mis : (unsigned char)0;
            if (r.bulk) bytes += (uint32_t)r.nbytes;
        }
        unsigned guarded = __ballot_sync(0xffffffffu, c < nslots && !r.bulk);
        while (guarded) {
            const int l = __ffs(guarded) - 1;
            guarded &= guarded - 1;
            const int cg = c0 + npwarps * l;
            const ShiftedRow g = shifted_row<NEED_UV>(P, fr, rs_l, rs_c, nl, a0l, lenl, a0c, lenc, cg);
            copy_row_ldg(stage + (size_t)cg * P.slot_stride, g.plane, g.base, g.len, g.pbytes, lane);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
    __syncwarp();
    if (lane == 0) {  // shift bytes and guarded rows are released to the consumers with this arrive
        if (bytes) mbar_expect_tx(bar, bytes);
        else mbar_arrive(bar);
    }
    __syncwarp();
    for (int c0 = pwarp; c0 < nslots; c0 += 32 * npwarps) {
        const int c = c0 + npwarps * lane;
        if (c < nslots) {
            const ShiftedRow r = shifted_row<NEED_UV>(P, fr, rs_l, rs_c, nl, a0l, lenl, a0c, lenc, c);
            if (r.bulk) bulk_g2s(stage + (size_t)c * P.slot_stride, r.plane + r.s0, (uint32_t)r.nbytes, bar, policy);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Resampling of one output sample (one byte of the reference's resized NV12 image)
// ---------------------------------------------------------------------------------------
// The dynamic shared memory (stage ring + mbarriers) of the fused kernel.
extern __shared__ __align__(128) unsigned char smem[];

// A position inside `smem` as a 32-bit shared-space address.  The consumers derive every tap address
// from one laundered base register and load with ld.shared: with `unsigned char *` (or `smem[i]`)
// arithmetic the compiler re-materialised the address of `smem` (S2UR SR_CgaCtaId / UMOV / ULEA /
// IMAD.U32) in every basic block of the sampling code -- four extra instructions per block, the ULEA
// among the top stall sites of the bilinear kernels.
struct SPtr {
    uint32_t addr;
    __device__ __forceinline__ SPtr operator+(int d) const { return SPtr{addr + (uint32_t)d}; }
    __device__ __forceinline__ uint32_t operator[](int i) const {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr + (uint32_t)i));
        return v;
    }
};
// shared-space address of smem[0], opaque to the optimiser so that it stays in one register
__device__ __forceinline__ uint32_t smem_base_laundered() {
    uint32_t b = smem_addr(smem);
    asm volatile("" : "+r"(b));
    return b;
}

// The slots of one output row: slot k (vertical tap k) starts at base + k * stride, plus -- for stages filled by
// fill_stage_shifted -- the byte the producer recorded for it.
template <bool SH>
struct RowSlots {
    SPtr base;
    uint32_t stride;
    uint32_t sh;  // shared address of the row's first shift byte (SH only)
    __device__ __forceinline__ SPtr row(int k) const {
        uint32_t a = base.addr + (uint32_t)k * stride;
        if constexpr (SH) {
            uint32_t v;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(sh + (uint32_t)k));
            a += v;
        }
        return SPtr{a};
    }
};

template <int RS>
struct ColOf;
template <>
struct ColOf<RS_NEAREST> { using T = ColNearest; };
template <>
struct ColOf<RS_BILINEAR> { using T = ColBilinear; };
template <>
struct ColOf<RS_BICUBIC> { using T = ColBicubic; };
template <>
struct ColOf<RS_AREA> { using T = ColArea; };

// AREA window sum with the horizontal extent known at compile time: the KXT horizontal weights stay
// in registers, the taps are immediate-offset shared loads.  Summation order and rounding are the
// reference's (rows outer, columns inner; w = px*py; sum = fma(d, w, sum)).
template <int KXT, int STRIDE, bool SH>
__device__ __noinline__ int area_sample(const RowSlots<SH> sl, int o, int K, const float *__restrict__ px,
                                        const float *__restrict__ py, float div) {
    float pxr[KXT];
#pragma unroll
    for (int b = 0; b < KXT; ++b) pxr[b] = __ldg(px + b);
    float sum = 0.0f;
    for (int a = 0; a < K; ++a) {
        const SPtr s = sl.row(a) + o;
        const float wyv = __ldg(py + a);
#pragma unroll
        for (int b = 0; b < KXT; ++b) sum = __fmaf_rn(u2f(s[b * STRIDE]), __fmul_rn(pxr[b], wyv), sum);
    }
    return (int)f2u8_sat(__fdiv_rn(sum, div));
}

// `slots` points at slot 0 of this output row; `orow` is the output row index used to look up
// the vertical weights (luma: i, chroma: ci -- the reference evaluates the *luma-grid* formula
// at the chroma index, src/Resize.cu:303-310), `stride` is 1 for luma and 2 for chroma bytes,
// `wy` is the preloaded bilinear vertical weight.
template <int RS, bool SH>
__device__ __forceinline__ int resample(const PlanDev &P, const RowSlots<SH> &slots,
                                        const typename ColOf<RS>::T &col, int orow, int stride, float wy) {
    if constexpr (RS == RS_NEAREST) {
        return slots.row(0)[col.o];
    } else if constexpr (RS == RS_BILINEAR) {
        const SPtr s0 = slots.row(0), s1 = slots.row(1);
        return bilinear_px(s0[col.o0], s0[col.o1], s1[col.o0], s1[col.o1], col.w, wy);
    } else if constexpr (RS == RS_BICUBIC) {
        const int o0 = col.o[0], o1 = col.o[1], o2 = col.o[2], o3 = col.o[3];
        const float cf[4] = {col.cf[0], col.cf[1], col.cf[2], col.cf[3]};
        const int exact_x = col.exact;
        uint32_t b[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const SPtr s = slots.row(k);
            b[k] = (uint32_t)bicubic_dot_fast(cf, col.c, exact_x, s[o0], s[o1], s[o2], s[o3]);
        }
        const float4 cyf4 = __ldg(reinterpret_cast<const float4 *>(P.cyf) + orow);
        const float cyf[4] = {cyf4.x, cyf4.y, cyf4.z, cyf4.w};
        return bicubic_dot_fast(cyf, P.cy + (size_t)orow * 4, __ldg(P.cy_exact + orow), b[0], b[1], b[2], b[3]);
    } else {
        // reference src/Resize.cu:160-178 (calculateAreaInterpolation): rows outer, columns
        // inner; w = px*py; sum = fma(d, w, sum); (int)(sum/div) -- div precomputed (same order)
        const float *px = P.px + (size_t)col.prow * P.KX;
        const int prow_y = orow % P.PY;
        const float *py = P.py + (size_t)prow_y * P.K;
        const float div = __ldg(P.area_div + (size_t)prow_y * P.PX + col.prow);
        if (stride == 1) {
            switch (P.KX) {
                case 2: return area_sample<2, 1, SH>(slots, col.o, P.K, px, py, div);
                case 3: return area_sample<3, 1, SH>(slots, col.o, P.K, px, py, div);
                case 4: return area_sample<4, 1, SH>(slots, col.o, P.K, px, py, div);
                case 5: return area_sample<5, 1, SH>(slots, col.o, P.K, px, py, div);
                case 6: return area_sample<6, 1, SH>(slots, col.o, P.K, px, py, div);
                case 7: return area_sample<7, 1, SH>(slots, col.o, P.K, px, py, div);
                case 8: return area_sample<8, 1, SH>(slots, col.o, P.K, px, py, div);
                default: break;
            }
        } else {
            switch (P.KX) {
                case 2: return area_sample<2, 2, SH>(slots, col.o, P.K, px, py, div);
                case 3: return area_sample<3, 2, SH>(slots, col.o, P.K, px, py, div);
                case 4: return area_sample<4, 2, SH>(slots, col.o, P.K, px, py, div);
                case 5: return area_sample<5, 2, SH>(slots, col.o, P.K, px, py, div);
                case 6: return area_sample<6, 2, SH>(slots, col.o, P.K, px, py, div);
                case 7: return area_sample<7, 2, SH>(slots, col.o, P.K, px, py, div);
                case 8: return area_sample<8, 2, SH>(slots, col.o, P.K, px, py, div);
                default: break;
            }
        }
        float sum = 0.0f;  // generic window width
        for (int a = 0; a < P.K; ++a) {
            const SPtr s = slots.row(a) + col.o;
            const float wyv = __ldg(py + a);
            for (int b = 0; b < P.KX; ++b) sum = __fmaf_rn(u2f(s[b * stride]), __fmul_rn(__ldg(px + b), wyv), sum);
        }
        return (int)f2u8_sat(__fdiv_rn(sum, div));
    }
}

// Bicubic, both luma rows of a block for one column.  Row 1 re-uses the horizontal sums of row 0
// wherever the two rows tap the same source row (`shift`, warp-uniform, from the plan): the sum
// depends on (source row, column) only, so this is the reference's arithmetic evaluated once
// instead of twice.
template <bool SH>
__device__ __forceinline__ void bicubic_luma_pair(const PlanDev &P, const RowSlots<SH> &sl0, const RowSlots<SH> &sl1,
                                                  const ColBicubic &col, int i0, int shift, bool two_rows, int &y0,
                                                  int &y1) {
    const int o0 = col.o[0], o1 = col.o[1], o2 = col.o[2], o3 = col.o[3];
    const float cf[4] = {col.cf[0], col.cf[1], col.cf[2], col.cf[3]};
    const int exact_x = col.exact;
    uint32_t b0[4], b1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const SPtr s = sl0.row(k);
        b0[k] = (uint32_t)bicubic_dot_fast(cf, col.c, exact_x, s[o0], s[o1], s[o2], s[o3]);
    }
    {
        const float4 c4 = __ldg(reinterpret_cast<const float4 *>(P.cyf) + i0);
        const float cyf[4] = {c4.x, c4.y, c4.z, c4.w};
        y0 = bicubic_dot_fast(cyf, P.cy + (size_t)i0 * 4, __ldg(P.cy_exact + i0), b0[0], b0[1], b0[2], b0[3]);
    }
    y1 = 0;
    if (!two_rows) return;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        // taps k <= 3-shift are row 0's k+shift; `shift` is uniform across the warp
        bool reuse = false;
        uint32_t v = 0;
        if (shift == 0) { reuse = true; v = b0[k]; }
        else if (shift == 1 && k <= 2) { reuse = true; v = b0[k + 1 > 3 ? 3 : k + 1]; }
        else if (shift == 2 && k <= 1) { reuse = true; v = b0[k + 2 > 3 ? 3 : k + 2]; }
        else if (shift == 3 && k == 0) { reuse = true; v = b0[3]; }
        if (!reuse) {
            const SPtr s = sl1.row(k);
            v = (uint32_t)bicubic_dot_fast(cf, col.c, exact_x, s[o0], s[o1], s[o2], s[o3]);
        }
        b1[k] = v;
    }
    const float4 c4 = __ldg(reinterpret_cast<const float4 *>(P.cyf) + i0 + 1);
    const float cyf[4] = {c4.x, c4.y, c4.z, c4.w};
    y1 = bicubic_dot_fast(cyf, P.cy + (size_t)(i0 + 1) * 4, __ldg(P.cy_exact + i0 + 1), b1[0], b1[1], b1[2], b1[3]);
}

// ---------------------------------------------------------------------------------------
// Colour conversion + layout for one block of 2 rows x BW columns (BW = 2 or 4) per thread
// ---------------------------------------------------------------------------------------
// NW consecutive 32-bit words to `p`; AW = guaranteed alignment of p in words (1, 2 or 4).
template <int NW, int AW>
__device__ __forceinline__ void st_words(void *p, const uint32_t (&w)[NW]) {
    if constexpr (AW >= 4 && NW % 4 == 0) {
#pragma unroll
        for (int k = 0; k < NW; k += 4)
            __stcs(reinterpret_cast<uint4 *>(p) + k / 4, make_uint4(w[k], w[k + 1], w[k + 2], w[k + 3]));
    } else if constexpr (AW >= 2 && NW % 2 == 0) {
#pragma unroll
        for (int k = 0; k < NW; k += 2) __stcs(reinterpret_cast<uint2 *>(p) + k / 2, make_uint2(w[k], w[k + 1]));
    } else {
#pragma unroll
        for (int k = 0; k < NW; ++k) __stcs(reinterpret_cast<unsigned int *>(p) + k, w[k]);
    }
}

// one row segment of NE float elements (NE = BW or 3*BW)
template <int NE, int BW>
__device__ __forceinline__ void st_row_f32(float *dst, const float (&e)[NE], bool full, int nvalid) {
    if (full) {
        uint32_t w[NE];
#pragma unroll
        for (int k = 0; k < NE; ++k) w[k] = __float_as_uint(e[k]);
        st_words<NE, BW>(dst, w);  // alignment: BW floats (8 or 16 bytes) by construction
    } else {
        for (int k = 0; k < nvalid; ++k) dst[k] = e[k];
    }
}
// b0 | b1<<8 | b2<<16 | b3<<24 for clean bytes, as integer multiply-adds
__device__ __forceinline__ uint32_t pack4(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
    uint32_t w;
    asm("{\n\t.reg .u32 t;\n\t"
        "mad.lo.u32 t, %2, 256, %1;\n\t"
        "mad.lo.u32 t, %3, 65536, t;\n\t"
        "mad.lo.u32 %0, %4, 16777216, t;\n\t}"
        : "=r"(w) : "r"(b0), "r"(b1), "r"(b2), "r"(b3));
    return w;
}
// one row segment of NE byte elements (values 0..255 in uint32)
template <int NE, int BW>
__device__ __forceinline__ void st_row_u8(unsigned char *dst, const uint32_t (&e)[NE], bool full, int nvalid) {
    if (full) {
        if constexpr (NE % 4 == 0) {
            uint32_t w[NE / 4];
#pragma unroll
            for (int k = 0; k < NE / 4; ++k)  // three multiply-adds per word: the FMA pipe has the headroom, the ALU pipe not
                w[k] = pack4(e[4 * k], e[4 * k + 1], e[4 * k + 2], e[4 * k + 3]);
            st_words<NE / 4, 1>(dst, w);
        } else {  // NE == 2 or 6: 16-bit stores
#pragma unroll
            for (int k = 0; k < NE / 2; ++k)
                *reinterpret_cast<unsigned short *>(dst + 2 * k) = (unsigned short)(e[2 * k] | (e[2 * k + 1] << 8));
        }
    } else {
        for (int k = 0; k < nvalid; ++k) dst[k] = (unsigned char)e[k];
    }
}

// Yv[a][b]: resized luma of pixel (i0+a, j0+b); UV[p] = {U,V} of column pair p (0 if it does not
// exist).  nx, ny: how many columns/rows of the block lie inside the image; full: nx==BW && ny==2
// and the vector-store alignment conditions hold.
// MAGIC: Yv / UV already hold kMagicBits | value (RGB-family outputs only).
// Element offsets are 32-bit (the plan refuses outputs of 2^32 elements or more): one IMAD.WIDE per store
// address instead of a 64-bit multiply chain.
template <int OK, int BW, bool MAGIC = false>
__device__ __forceinline__ void emit_block(const PlanDev &P, void *outp, int i0, int j0, const int (&Yv)[2][BW],
                                           const int (&UV)[BW / 2][2], int nx, int ny, bool full) {
    const int dw = P.dw, dh = P.dh;
    const int nrow = full ? 2 : ny;
    if constexpr (OK == OUT_Y800_U8 || OK == OUT_NV12_U8) {
        unsigned char *o = (unsigned char *)outp;
        for (int a = 0; a < nrow; ++a) {
            uint32_t e[BW];
#pragma unroll
            for (int b = 0; b < BW; ++b) e[b] = (uint32_t)Yv[a][b];
            st_row_u8<BW, BW>(o + ((uint32_t)(i0 + a) * (uint32_t)dw + (uint32_t)j0), e, full, nx);
        }
        if constexpr (OK == OUT_NV12_U8) {
            // reference src/ColorConversion.cu:211-233 (NV12MergeBuffers): Y plane then UV plane
            if ((i0 >> 1) < P.ch2) {
                uint32_t e[BW];
#pragma unroll
                for (int p = 0; p < BW / 2; ++p) { e[2 * p] = (uint32_t)UV[p][0]; e[2 * p + 1] = (uint32_t)UV[p][1]; }
                const int nuv = 2 * max(0, min(BW / 2, P.cw2 - (j0 >> 1)));
                st_row_u8<BW, BW>(o + ((uint32_t)dw * (uint32_t)dh + (uint32_t)(i0 >> 1) * (uint32_t)dw + (uint32_t)j0), e, full && nuv == BW, nuv);
            }
        }
    } else if constexpr (OK == OUT_Y800_F32 || OK == OUT_NV12_F32) {
        float *o = (float *)outp;
        for (int a = 0; a < nrow; ++a) {
            float e[BW];
#pragma unroll
            for (int b = 0; b < BW; ++b) e[b] = div255_u8((uint32_t)Yv[a][b]);
            st_row_f32<BW, BW>(o + ((uint32_t)(i0 + a) * (uint32_t)dw + (uint32_t)j0), e, full, nx);
        }
        if constexpr (OK == OUT_NV12_F32) {
            if ((i0 >> 1) < P.ch2) {
                float e[BW];
#pragma unroll
                for (int p = 0; p < BW / 2; ++p) { e[2 * p] = div255_u8((uint32_t)UV[p][0]); e[2 * p + 1] = div255_u8((uint32_t)UV[p][1]); }
                const int nuv = 2 * max(0, min(BW / 2, P.cw2 - (j0 >> 1)));
                st_row_f32<BW, BW>(o + ((uint32_t)dw * (uint32_t)dh + (uint32_t)(i0 >> 1) * (uint32_t)dw + (uint32_t)j0), e, full && nuv == BW, nuv);
            }
        }
    } else {
        // RGB family: channel values 0..255, [row][col][channel]
        int C[2][BW][3];
#pragma unroll
        for (int p = 0; p < BW / 2; ++p) {
            ChromaTerms t = chroma_terms(MAGIC ? (uint32_t)UV[p][0] : (kMagicBits | (uint32_t)UV[p][0]),
                                         MAGIC ? (uint32_t)UV[p][1] : (kMagicBits | (uint32_t)UV[p][1]));
            if (OK != OUT_HSV_F32 && P.swap_rb) {  // BGR: first channel is B (ColorConversion.cu:304-305);
                const float tmp = t.r;              // swapping the two chroma terms once per pair swaps the
                t.r = t.b;                          // R and B results of its four pixels
                t.b = tmp;
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b)
                    yuv_px_bits(MAGIC ? (uint32_t)Yv[a][2 * p + b] : (kMagicBits | (uint32_t)Yv[a][2 * p + b]), t,
                                C[a][2 * p + b][0], C[a][2 * p + b][1], C[a][2 * p + b][2]);
        }
        if constexpr (OK == OUT_HSV_F32) {
            // reference src/ColorConversion.cu:360-372: from normalised RGB (never BGR), merged
            for (int a = 0; a < nrow; ++a) {
                float e[3 * BW];
#pragma unroll
                for (int b = 0; b < BW; ++b)
                    rgb_to_hsv(div255_bits(C[a][b][0]), div255_bits(C[a][b][1]), div255_bits(C[a][b][2]), e[3 * b], e[3 * b + 1], e[3 * b + 2]);
                st_row_f32<3 * BW, BW>((float *)outp + ((uint32_t)(i0 + a) * (uint32_t)dw + (uint32_t)j0) * 3u, e, full, 3 * nx);
            }
            return;
        }
        if constexpr (OK == OUT_RGB_PLANAR_F32 || OK == OUT_RGB_PLANAR_U8) {
            const uint32_t plane = (uint32_t)dw * (uint32_t)dh;
            for (int a = 0; a < nrow; ++a) {
                const uint32_t off = (uint32_t)(i0 + a) * (uint32_t)dw + (uint32_t)j0;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    if constexpr (OK == OUT_RGB_PLANAR_F32) {
                        float e[BW];
#pragma unroll
                        for (int b = 0; b < BW; ++b) e[b] = div255_bits(C[a][b][ch]);
                        st_row_f32<BW, BW>((float *)outp + ((uint32_t)ch * plane + off), e, full, nx);
                    } else {
                        uint32_t e[BW];
#pragma unroll
                        for (int b = 0; b < BW; ++b) e[b] = bits_u8(C[a][b][ch]);
                        st_row_u8<BW, BW>((unsigned char *)outp + ((uint32_t)ch * plane + off), e, full, nx);
                    }
                }
            }
        } else {
            for (int a = 0; a < nrow; ++a) {
                const uint32_t off = ((uint32_t)(i0 + a) * (uint32_t)dw + (uint32_t)j0) * 3u;
                if constexpr (OK == OUT_RGB_MERGED_F32) {
                    float e[3 * BW];
#pragma unroll
                    for (int b = 0; b < BW; ++b) {
                        e[3 * b] = div255_bits(C[a][b][0]);
                        e[3 * b + 1] = div255_bits(C[a][b][1]);
                        e[3 * b + 2] = div255_bits(C[a][b][2]);
                    }
                    st_row_f32<3 * BW, BW>((float *)outp + off, e, full, 3 * nx);
                } else {
                    uint32_t e[3 * BW];
#pragma unroll
                    for (int b = 0; b < BW; ++b) {
                        e[3 * b] = bits_u8(C[a][b][0]);
                        e[3 * b + 1] = bits_u8(C[a][b][1]);
                        e[3 * b + 2] = bits_u8(C[a][b][2]);
                    }
                    st_row_u8<3 * BW, BW>((unsigned char *)outp + off, e, full, 3 * nx);
                }
            }
        }
    }
}

// BW consecutive staged bytes (BW-aligned) as one shared load
template <int BW>
__device__ __forceinline__ uint32_t ld_group(SPtr p) {
    uint32_t v;
    if constexpr (BW == 4) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(p.addr));
    else asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(p.addr));
    return v;
}

// one column-parameter record (4, 8 or 16 bytes, naturally aligned in the plan tables) as one load
template <class T>
__device__ __forceinline__ T ld_col(const T *p) {
    T t;
    if constexpr (sizeof(T) == 16) {
        const int4 v = __ldg(reinterpret_cast<const int4 *>(p));
        memcpy(&t, &v, sizeof(T));
    } else if constexpr (sizeof(T) == 8) {
        const int2 v = __ldg(reinterpret_cast<const int2 *>(p));
        memcpy(&t, &v, sizeof(T));
    } else if constexpr (sizeof(T) == 4) {
        const int v = __ldg(reinterpret_cast<const int *>(p));
        memcpy(&t, &v, sizeof(T));
    } else {
        t = *p;
    }
    return t;
}

}  // namespace tsvpp
