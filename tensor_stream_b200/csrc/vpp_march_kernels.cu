// vpp_march_kernels.cu -- exact-integer separable resamplers for sm_100a (B200): bicubic and area
// downscale for geometries whose weights are dyadic rationals (see vpp_march.h for the argument and the
// work decomposition).  Replaces, for those geometries, the float paths of vpp_kernels.cu that reproduce
// reference src/Resize.cu:27-91,314-357 (bicubic, FP64) and :160-212,359-386 (area, FP32): same bytes out,
// 4-6x fewer instructions per output sample.
//
//   producer warp : one cp.async.bulk (TMA unit, SASS UBLKCP) per DISTINCT source row a strip touches --
//                   a row that feeds several output rows is fetched and staged once (the slot-per-tap
//                   layout of vpp_kernels.cu stages it up to four times).
//   consumers     : thread = 2 adjacent output columns x the rpg row pairs of its row group.
//                   bicubic: per newly needed source row and column, 2 aligned shared loads + 1 PRMT give
//                   the four taps as packed bytes, 2 (or 4) dp4a give the exact horizontal sum, which is
//                   rounded / clamped like the reference's b0..b3 and shifted into a 4-byte register window;
//                   each output row is then one more packed dot product of the window.
//                   area: per window row, funnel-shifted words x packed integer pattern weights (dp4a),
//                   accumulated with the row's integer weight; one exact division by multiply-shift.
// Colour conversion, normalisation and the stores are vpp_device.cuh's emit_block, shared with the float path.
#include "vpp_device.cuh"
#include "vpp_march.h"

// CTAs per SM the register allocation is bounded for.  Bicubic: 3 (72 registers, a few spilled words) measured
// +12 % (C3) / +22 % (1080p -> 2160p) over 2 (96 registers): the dependent dp4a / PRMT chains want the warps more
// than the registers (gpurun r2c).  Area: 2.
#ifndef TSVPP_MARCH_MINBLOCKS_BICUBIC
#define TSVPP_MARCH_MINBLOCKS_BICUBIC 3
#endif
#ifndef TSVPP_MARCH_MINBLOCKS_BILINEAR
#define TSVPP_MARCH_MINBLOCKS_BILINEAR 2
#endif
#ifndef TSVPP_MARCH_MINBLOCKS_AREA
#define TSVPP_MARCH_MINBLOCKS_AREA 2
#endif

namespace tsvpp {

// u8 x s8 and u8 x u8 four-way dot products with 32-bit accumulate (SASS IDP.4A)
__device__ __forceinline__ int dp4a_us(uint32_t a_u8x4, int b_s8x4, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t dp4a_uu(uint32_t a_u8x4, uint32_t b_u8x4, uint32_t c) {
    uint32_t d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_u8x4), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// round_half_away(sum_k taps_k * m_k / 2^q) for integer weights m_k given as base-128 digits, NOT yet clamped
// -- reference src/Resize.cu:49-51 (`int b = round(a0+a1+a2+a3); min 255; max 0`).  For s >= 0 round-half-away
// is floor(s + 1/2); negative sums end up <= 0 and clamp to 0 either way.
//   short (q <= 13): m = 128*w[1] + w[0], w[0] in [0,127], w[1] signed
//   long  (15 <= q <= 27): m = 2^14 * (128*w[3] + w[2]) + (128*w[1] + w[0]), the low part in [0, 16383]:
//         floor((hi*2^14 + lo + 2^(q-1)) / 2^q) = floor((hi + 2^(q-15) + floor(lo / 2^14)) / 2^(q-14))  (lo >= 0)
// `tie` reports a sum whose fraction is exactly one half.  The reference's FP64 weights come out of libdevice's
// pow(), which is not correctly rounded: they can sit a few 1e-17 off the dyadic values m / 2^q.  That moves
// the FP64 sum by < 1e-12, which changes round() only where the exact sum is a tie -- every other multiple of
// 2^-q is at least 2^-27 away from one.  Ties of inexact weights are re-evaluated with the reference's own
// FP64 chain (bicubic_dot); everything else is exact by construction.
struct QConst {  // per axis, in registers: short {q, 2^(q-1), 2^q - 1}; long {q-14, 2^(q-15), 2^(q-14) - 1}
    int sh, half, mask;
    __device__ __forceinline__ void set(int q, bool lng) {
        sh = lng ? q - 14 : q;
        half = lng ? (1 << (q - 15)) : (1 << (q - 1));
        mask = (1 << sh) - 1;
        asm volatile("" : "+r"(sh), "+r"(half), "+r"(mask));  // keep them out of the constant bank re-loads
    }
};
template <bool LONG>
__device__ __forceinline__ int idot_raw(uint32_t taps, const int (&w)[4], const QConst &qc, bool &tie) {
    if constexpr (!LONG) {
        const int lo = dp4a_us(taps, w[0], qc.half);
        const int hi = dp4a_us(taps, w[1], 0);
        const int t = hi * 128 + lo;
        tie = (t & qc.mask) == 0;
        return t >> qc.sh;
    } else {
        const int lo = dp4a_us(taps, w[1], 0) * 128 + dp4a_us(taps, w[0], 0);
        const int hi = dp4a_us(taps, w[3], 0) * 128 + dp4a_us(taps, w[2], qc.half);
        const int u = hi + (lo >> 14);
        tie = ((lo & 16383) == 0) && ((u & qc.mask) == 0);
        return u >> qc.sh;
    }
}
__device__ __forceinline__ uint32_t clamp_u8(int r) { return (uint32_t)min(max(r, 0), 255); }
// (c << 16) | (sat_u8(a) << 8) | sat_u8(b): two saturating conversions and a two-byte shift of `c` in ONE
// instruction (SASS I2IP.U8.S32.SAT) -- marches a window by two slots: a = older, b = newest
__device__ __forceinline__ uint32_t pack_sat2(int a, int b, uint32_t c) {
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// The tie path: reference src/Resize.cu:45-51 with its FP64 weights (taps in tap order, one per byte), DFMA chain
// as compiled (vpp_device.cuh: bicubic_dot).  `r` is the integer path's result floor(s_exact + 1/2) = n + 1 with
// s_exact = n + 1/2: the FP64 sum is within 1e-10 of that, so round-half-away of it is n + 1 if it is >= n + 1/2
// and n otherwise (n + 1/2 is exact in FP64; for n < 0 both clamp to 0) -- no libdevice round() needed.
#ifndef TSVPP_TIE_INLINE
#define TSVPP_TIE_INLINE __forceinline__
#endif
__device__ TSVPP_TIE_INLINE int bicubic_tie_fp64(const double *__restrict__ c, uint32_t taps, int r) {
    const double2 c01 = __ldg(reinterpret_cast<const double2 *>(c));
    const double2 c23 = __ldg(reinterpret_cast<const double2 *>(c) + 1);
    double s = __dmul_rn(c01.y, (double)(int)((taps >> 8) & 0xffu));
    s = __fma_rn(c01.x, (double)(int)(taps & 0xffu), s);
    s = __fma_rn(c23.x, (double)(int)((taps >> 16) & 0xffu), s);
    s = __fma_rn(c23.y, (double)(int)(taps >> 24), s);
    return s >= (double)r - 0.5 ? r : r - 1;
}

// ---------------------------------------------------------------------------------------
// producer: one bulk copy per distinct source row of the strip
// ---------------------------------------------------------------------------------------
struct MarchIter {  // mixed-radix (frame, strip, tile) counter advancing by gridDim.x items
    int frame, strip, tile;
    int d_frame, d_strip, d_tile;
    __device__ __forceinline__ void init(const MarchPlanDev &P, int first, int step) {
        const int per_frame = P.nstrips * P.ntiles;
        frame = first / per_frame;
        const int r = first - frame * per_frame;
        strip = r / P.ntiles;
        tile = r - strip * P.ntiles;
        d_tile = step % P.ntiles;
        const int q = step / P.ntiles;
        d_strip = q % P.nstrips;
        d_frame = q / P.nstrips;
    }
    __device__ __forceinline__ void next(const MarchPlanDev &P) {
        tile += d_tile;
        int carry = tile >= P.ntiles ? 1 : 0;
        tile -= carry ? P.ntiles : 0;
        strip += d_strip + carry;
        carry = strip >= P.nstrips ? 1 : 0;
        strip -= carry ? P.nstrips : 0;
        frame += d_frame + carry;
    }
};

template <bool NEED_UV>
__device__ __forceinline__ void march_fill(const MarchPlanDev &P, const FrameDev *__restrict__ frames,
                                           const MarchIter &ic, unsigned char *stage, uint64_t *bar, int lane,
                                           uint64_t policy) {
    const unsigned long long *fd = reinterpret_cast<const unsigned long long *>(frames + ic.frame);
    const unsigned char *fy = reinterpret_cast<const unsigned char *>(__ldg(fd));
    const unsigned char *fuv = reinterpret_cast<const unsigned char *>(__ldg(fd + 1));
    const int nl = __ldg(P.nslots_l + ic.strip);
    const int nc = NEED_UV ? __ldg(P.nslots_c + ic.strip) : 0;
    const int a0l = __ldg(P.seg_a0_l + ic.tile), lenl = __ldg(P.seg_len_l + ic.tile);
    const int a0c = NEED_UV ? __ldg(P.seg_a0_c + ic.tile) : 0, lenc = NEED_UV ? __ldg(P.seg_len_c + ic.tile) : 0;
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)(nl * lenl + nc * lenc));
    __syncwarp();
    fy += a0l;
    fuv += a0c;
    const int *rl = P.slotrow_l + (size_t)ic.strip * P.nsl;
#pragma unroll 4
    for (int s = 0; s < nl; ++s)
        bulk_g2s_elected(stage + (size_t)s * P.slot_stride, fy + (size_t)__ldg(rl + s) * P.pitch_y, lenl, bar, policy);
    if (NEED_UV) {
        const int *rc = P.slotrow_c + (size_t)ic.strip * P.nsc;
        unsigned char *sc = stage + (size_t)P.nsl * P.slot_stride;
#pragma unroll 4
        for (int s = 0; s < nc; ++s)
            bulk_g2s_elected(sc + (size_t)s * P.slot_stride, fuv + (size_t)__ldg(rc + s) * P.pitch_uv, lenc, bar, policy);
    }
}

// ---------------------------------------------------------------------------------------
// per-thread column state
// ---------------------------------------------------------------------------------------
template <bool XLONG>
struct BicubicCols {
    // two luma columns and one chroma pair
    int addr0, addr1, addrc;
    uint32_t sel0, sel1, selcA, selcB;
    int inexact;  // bit 0 / 1 / 2: the weights of column 0 / column 1 / the chroma pair are only NEAR dyadic (FP64 tie path)
    int w0[4], w1[4], wc[4];
    __device__ __forceinline__ void load(const MarchColB *cl, const MarchColB *cc) {
        const int4 a0 = __ldg(reinterpret_cast<const int4 *>(cl));
        const int4 b0 = __ldg(reinterpret_cast<const int4 *>(cl) + 1);
        const int4 a1 = __ldg(reinterpret_cast<const int4 *>(cl + 1));
        const int4 b1 = __ldg(reinterpret_cast<const int4 *>(cl + 1) + 1);
        addr0 = a0.x; sel0 = (uint32_t)a0.y;
        addr1 = a1.x; sel1 = (uint32_t)a1.y;
        inexact = (a0.w ? 0 : 1) | (a1.w ? 0 : 2);
        w0[0] = b0.x; w0[1] = b0.y; w0[2] = b0.z; w0[3] = b0.w;
        w1[0] = b1.x; w1[1] = b1.y; w1[2] = b1.z; w1[3] = b1.w;
        if (cc) {
            const int4 ac = __ldg(reinterpret_cast<const int4 *>(cc));
            const int4 bc = __ldg(reinterpret_cast<const int4 *>(cc) + 1);
            addrc = ac.x; selcA = (uint32_t)ac.y; selcB = (uint32_t)ac.z;
            inexact |= ac.w ? 0 : 4;
            wc[0] = bc.x; wc[1] = bc.y; wc[2] = bc.z; wc[3] = bc.w;
        }
    }
};

// ---------------------------------------------------------------------------------------
// bicubic: one thread, two output columns, the rpg row pairs of its row group
// ---------------------------------------------------------------------------------------
// Windows hold the horizontal sums of the last four marched slots, byte 0 = newest.  A march of two slots is
// one I2IP (pack_sat2), of one slot a clamp + PRMT; the FP64 tie path patches the raw sums before they are packed.
template <bool XL, bool YL, int OK, bool NEED_UV>
__device__ __forceinline__ void march_bicubic_group(const MarchPlanDev &P, const BicubicCols<XL> &bc, uint32_t stage,
                                                    uint32_t stage_c, uint32_t sstride, int strip, int g, int j0,
                                                    bool have_c, bool vec, void *outp) {
    const int R2 = P.G * P.rpg;
    const int pr0 = strip * R2 + g * P.rpg;  // first row pair (global index) of this thread's group
    const MarchRowB *rowl = reinterpret_cast<const MarchRowB *>(P.row_l);
    const MarchRowB *rowc = reinterpret_cast<const MarchRowB *>(P.row_c);
    const double *cxd = P.cxd, *cyd = P.cyd;
    // the RGB family takes its samples as kMagicBits | value (the float 2^23 + value): one PRMT either way
    constexpr bool MAGIC = !(OK == OUT_Y800_U8 || OK == OUT_Y800_F32 || OK == OUT_NV12_U8 || OK == OUT_NV12_F32);
    constexpr uint32_t kHi = MAGIC ? kMagicBits : 0u;
    constexpr uint32_t kSelA = MAGIC ? 0x7541u : 0x4441u;  // {byte 1 of the pack, 0, 0, 0x4B | 0}
    QConst qx, qy;
    qx.set(P.qx, XL);
    qy.set(P.qy, YL);
    const int dh = P.e.dh, ch2 = P.e.ch2, rpg = P.rpg;

    // one march step of the two luma columns over `n` (1 or 2) slots starting at shared address sl
    auto luma_step2 = [&](uint32_t sl, uint32_t &W0, uint32_t &W1) {
        const uint32_t a0 = lds32(sl + bc.addr0), a1 = lds32(sl + bc.addr0 + 4);
        const uint32_t b0 = lds32(sl + bc.addr1), b1 = lds32(sl + bc.addr1 + 4);
        const uint32_t c0 = lds32(sl + sstride + bc.addr0), c1 = lds32(sl + sstride + bc.addr0 + 4);
        const uint32_t d0 = lds32(sl + sstride + bc.addr1), d1 = lds32(sl + sstride + bc.addr1 + 4);
        const uint32_t tA0 = __byte_perm(a0, a1, bc.sel0), tA1 = __byte_perm(b0, b1, bc.sel1);
        const uint32_t tB0 = __byte_perm(c0, c1, bc.sel0), tB1 = __byte_perm(d0, d1, bc.sel1);
        bool zA0, zA1, zB0, zB1;
        int rA0 = idot_raw<XL>(tA0, bc.w0, qx, zA0), rA1 = idot_raw<XL>(tA1, bc.w1, qx, zA1);
        int rB0 = idot_raw<XL>(tB0, bc.w0, qx, zB0), rB1 = idot_raw<XL>(tB1, bc.w1, qx, zB1);
        if ((((zA0 | zB0) ? 1 : 0) | ((zA1 | zB1) ? 2 : 0)) & bc.inexact) {
            if (zA0 && (bc.inexact & 1)) rA0 = bicubic_tie_fp64(cxd + 4 * (size_t)j0, tA0, rA0);
            if (zB0 && (bc.inexact & 1)) rB0 = bicubic_tie_fp64(cxd + 4 * (size_t)j0, tB0, rB0);
            if (zA1 && (bc.inexact & 2)) rA1 = bicubic_tie_fp64(cxd + 4 * (size_t)(j0 + 1), tA1, rA1);
            if (zB1 && (bc.inexact & 2)) rB1 = bicubic_tie_fp64(cxd + 4 * (size_t)(j0 + 1), tB1, rB1);
        }
        W0 = pack_sat2(rA0, rB0, W0);
        W1 = pack_sat2(rA1, rB1, W1);
    };
    auto luma_step1 = [&](uint32_t sl, uint32_t &W0, uint32_t &W1) {
        const uint32_t a0 = lds32(sl + bc.addr0), a1 = lds32(sl + bc.addr0 + 4);
        const uint32_t b0 = lds32(sl + bc.addr1), b1 = lds32(sl + bc.addr1 + 4);
        const uint32_t tA0 = __byte_perm(a0, a1, bc.sel0), tA1 = __byte_perm(b0, b1, bc.sel1);
        bool zA0, zA1;
        int rA0 = idot_raw<XL>(tA0, bc.w0, qx, zA0), rA1 = idot_raw<XL>(tA1, bc.w1, qx, zA1);
        if (((zA0 ? 1 : 0) | (zA1 ? 2 : 0)) & bc.inexact) {
            if (zA0 && (bc.inexact & 1)) rA0 = bicubic_tie_fp64(cxd + 4 * (size_t)j0, tA0, rA0);
            if (zA1 && (bc.inexact & 2)) rA1 = bicubic_tie_fp64(cxd + 4 * (size_t)(j0 + 1), tA1, rA1);
        }
        W0 = __byte_perm(W0, clamp_u8(rA0), 0x2104);  // (W << 8) | new
        W1 = __byte_perm(W1, clamp_u8(rA1), 0x2104);
    };
    // chroma pair cj = j0/2 uses the horizontal weights of luma column cj (src/Resize.cu:351-353)
    auto chroma_taps = [&](uint32_t sl, uint32_t &tu, uint32_t &tv) {
        const uint32_t x0 = lds32(sl), x1 = lds32(sl + 4), x2 = lds32(sl + 8);
        tu = __byte_perm(__byte_perm(x0, x1, bc.selcA & 0xffffu), x2, bc.selcB & 0xffffu);
        tv = __byte_perm(__byte_perm(x0, x1, bc.selcA >> 16), x2, bc.selcB >> 16);
    };
    auto chroma_step2 = [&](uint32_t sl, uint32_t &WU, uint32_t &WV) {
        uint32_t uA, vA, uB, vB;
        chroma_taps(sl, uA, vA);
        chroma_taps(sl + sstride, uB, vB);
        bool z0, z1, z2, z3;
        int rUA = idot_raw<XL>(uA, bc.wc, qx, z0), rVA = idot_raw<XL>(vA, bc.wc, qx, z1);
        int rUB = idot_raw<XL>(uB, bc.wc, qx, z2), rVB = idot_raw<XL>(vB, bc.wc, qx, z3);
        if ((z0 | z1 | z2 | z3) && (bc.inexact & 4)) {
            const double *cd = cxd + 4 * (size_t)(j0 >> 1);
            if (z0) rUA = bicubic_tie_fp64(cd, uA, rUA);
            if (z1) rVA = bicubic_tie_fp64(cd, vA, rVA);
            if (z2) rUB = bicubic_tie_fp64(cd, uB, rUB);
            if (z3) rVB = bicubic_tie_fp64(cd, vB, rVB);
        }
        WU = pack_sat2(rUA, rUB, WU);
        WV = pack_sat2(rVA, rVB, WV);
    };
    auto chroma_step1 = [&](uint32_t sl, uint32_t &WU, uint32_t &WV) {
        uint32_t uA, vA;
        chroma_taps(sl, uA, vA);
        bool z0, z1;
        int rUA = idot_raw<XL>(uA, bc.wc, qx, z0), rVA = idot_raw<XL>(vA, bc.wc, qx, z1);
        if ((z0 | z1) && (bc.inexact & 4)) {
            const double *cd = cxd + 4 * (size_t)(j0 >> 1);
            if (z0) rUA = bicubic_tie_fp64(cd, uA, rUA);
            if (z1) rVA = bicubic_tie_fp64(cd, vA, rVA);
        }
        WU = __byte_perm(WU, clamp_u8(rUA), 0x2104);
        WV = __byte_perm(WV, clamp_u8(rVA), 0x2104);
    };
    // row records: first 16 bytes {meta, -, w0, w1}, next 16 {w2, w3, -, -} (read for long vertical digits only)
    auto load_row = [&](const MarchRowB *r, int4 &lo, int4 &hi) {
        // volatile: the loads stay where they are written (requested a row ahead, consumed late) instead of being
        // sunk to their first use
        asm volatile("ld.global.nc.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "l"(r));
        if (YL)
            asm volatile("ld.global.nc.v4.s32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                         : "l"(reinterpret_cast<const int4 *>(r) + 1));
    };
    // vertical sums of two windows that share a row record; tie path with the row's FP64 weights
    auto vsum2 = [&](uint32_t Wa, uint32_t Wb, const int4 &lo, const int4 &hi, const double *cd, int &ra, int &rb) {
        const int wv[4] = {lo.z, lo.w, YL ? hi.x : 0, YL ? hi.y : 0};
        bool za, zb;
        ra = idot_raw<YL>(Wa, wv, qy, za);
        rb = idot_raw<YL>(Wb, wv, qy, zb);
        const uint32_t meta = (uint32_t)lo.x;
        if ((za | zb) && !(meta >> 28)) {
            const uint32_t tapsel = (meta >> 12) & 0xffffu;  // window bytes back into tap order
            if (za) ra = bicubic_tie_fp64(cd, __byte_perm(Wa, 0u, tapsel), ra);
            if (zb) rb = bicubic_tie_fp64(cd, __byte_perm(Wb, 0u, tapsel), rb);
        }
        const uint32_t pk = pack_sat2(ra, rb, 0u);  // byte 1 = clamp(ra), byte 0 = clamp(rb)
        ra = (int)__byte_perm(pk, kHi, kSelA);
        rb = (int)__byte_perm(pk, kHi, kSelA - 1u);
    };

    uint32_t WY0 = 0, WY1 = 0, WU = 0, WV = 0;
    int curl = __ldg(P.gfirst_l + strip * P.G + g) - 1;
    int curc = NEED_UV ? __ldg(P.gfirst_c + strip * P.G + g) - 1 : 0;
    uint32_t sl_l = stage + (uint32_t)(curl + 1) * sstride;    // shared address of the next luma slot to march
    uint32_t sl_c = stage_c + (uint32_t)(curc + 1) * sstride + (uint32_t)bc.addrc;
    int4 ra_lo, ra_hi = make_int4(0, 0, 0, 0);
    load_row(rowl + 2 * pr0, ra_lo, ra_hi);
    for (int p = 0; p < rpg; ++p) {
        const int pr = pr0 + p;
        const int i0 = 2 * pr;
        if (i0 >= dh) break;  // uniform: rows past the image
        // the records of this pair's second row and chroma row are requested before the first row's work
        int4 rb_lo, rb_hi = make_int4(0, 0, 0, 0), rc_lo = make_int4(0, 0, 0, 0), rc_hi = make_int4(0, 0, 0, 0);
        load_row(rowl + i0 + 1, rb_lo, rb_hi);
        if (NEED_UV) load_row(rowc + pr, rc_lo, rc_hi);
        int Yv[2][2];
        int UV[1][2] = {{(int)kHi, (int)kHi}};
        {
            int n = (int)((uint32_t)ra_lo.x & 0xfffu) - curl;  // uniform
            curl += n;
            for (; n >= 2; n -= 2, sl_l += 2 * sstride) luma_step2(sl_l, WY0, WY1);
            if (n > 0) { luma_step1(sl_l, WY0, WY1); sl_l += sstride; }
            vsum2(WY0, WY1, ra_lo, ra_hi, cyd + 4 * (size_t)i0, Yv[0][0], Yv[0][1]);
        }
        if (p + 1 < rpg) load_row(rowl + i0 + 2, ra_lo, ra_hi);  // next pair's first row
        {
            int n = (int)((uint32_t)rb_lo.x & 0xfffu) - curl;
            curl += n;
            for (; n >= 2; n -= 2, sl_l += 2 * sstride) luma_step2(sl_l, WY0, WY1);
            if (n > 0) { luma_step1(sl_l, WY0, WY1); sl_l += sstride; }
            vsum2(WY0, WY1, rb_lo, rb_hi, cyd + 4 * (size_t)(i0 + 1), Yv[1][0], Yv[1][1]);
        }
        if (NEED_UV && pr < ch2) {
            int n = (int)((uint32_t)rc_lo.x & 0xfffu) - curc;
            curc += n;
            for (; n >= 2; n -= 2, sl_c += 2 * sstride) chroma_step2(sl_c, WU, WV);
            if (n > 0) { chroma_step1(sl_c, WU, WV); sl_c += sstride; }
            int u, v;
            vsum2(WU, WV, rc_lo, rc_hi, cyd + 4 * (size_t)pr, u, v);
            if (have_c) { UV[0][0] = u; UV[0][1] = v; }
        }
        const int nx = min(2, P.e.dw - j0), ny = min(2, dh - i0);
        if (nx > 0) {
            if (nx == 2 && ny == 2 && vec)
                emit_block<OK, 2, MAGIC>(P.e, outp, i0, j0, Yv, UV, 2, 2, true);
            else
                emit_block<OK, 2, MAGIC>(P.e, outp, i0, j0, Yv, UV, nx, ny, false);
        }
    }
}

// ---------------------------------------------------------------------------------------
// bilinear: one thread, FOUR output columns (two chroma pairs), the rpg row pairs of its row group
// ---------------------------------------------------------------------------------------
struct BilinearCols {
    int addr[4], caddr[2];
    uint32_t sel[4], w[4], selu[2], selv[2], cw[2];
};

// reference src/Resize.cu:5-25 with exact dyadic weights: (int)v = (H_old * w_old + H_new * w_new) >> (ex + ey) with
// H = A * (2^ex - mx) + B * mx per source row.  Nothing rounds before the final truncation, values stay in [0, 255].
template <int OK, bool NEED_UV>
__device__ __forceinline__ void march_bilinear_group(const MarchPlanDev &P, const BilinearCols &lc, uint32_t stage,
                                                     uint32_t stage_c, uint32_t sstride, int strip, int g, int j0, bool vec,
                                                     void *outp) {
    const int R2 = P.G * P.rpg;
    const int pr0 = strip * R2 + g * P.rpg;
    const MarchRowL *rowl = reinterpret_cast<const MarchRowL *>(P.row_l);
    const MarchRowL *rowc = reinterpret_cast<const MarchRowL *>(P.row_c);
    const int dh = P.e.dh, ch2 = P.e.ch2, rpg = P.rpg, vsh = P.vshift;
    uint32_t H0[4] = {0, 0, 0, 0}, H1[4] = {0, 0, 0, 0};
    uint32_t U0[2] = {0, 0}, U1[2] = {0, 0}, V0[2] = {0, 0}, V1[2] = {0, 0};
    int curl = __ldg(P.gfirst_l + strip * P.G + g) - 1;
    int curc = NEED_UV ? __ldg(P.gfirst_c + strip * P.G + g) - 1 : 0;
    uint32_t sl_l = stage + (uint32_t)(curl + 1) * sstride;
    uint32_t sl_c = stage_c + (uint32_t)(curc + 1) * sstride;
    // row records are one word each; the next pair's three are requested a whole pair ahead (volatile: the loads stay
    // here instead of being sunk to their first use)
    auto ldrow = [](const MarchRowL *r) -> uint32_t {
        uint32_t v;
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(r));
        return v;
    };
    uint32_t nra = ldrow(rowl + 2 * pr0), nrb = ldrow(rowl + 2 * pr0 + 1), nrc = NEED_UV ? ldrow(rowc + pr0) : 0u;
    for (int p = 0; p < rpg; ++p) {
        const int pr = pr0 + p;
        const int i0 = 2 * pr;
        if (i0 >= dh) break;
        const uint32_t ra = nra, rb = nrb, rc = nrc;
        if (p + 1 < rpg) {  // tables are padded by two rows
            nra = ldrow(rowl + i0 + 2);
            nrb = ldrow(rowl + i0 + 3);
            if (NEED_UV) nrc = ldrow(rowc + pr + 1);
        }
        int Yv[2][4];
        int UV[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const uint32_t r = a ? rb : ra;
            for (int n = (int)(r & 0xffffu) - curl; n > 0; --n, ++curl, sl_l += sstride) {  // uniform
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t t = __byte_perm(lds32(sl_l + lc.addr[c]), lds32(sl_l + lc.addr[c] + 4), lc.sel[c]);
                    H0[c] = H1[c];
                    H1[c] = dp4a_uu(t, lc.w[c], 0u);
                }
            }
            const uint32_t w0 = (r >> 16) & 0xffu, w1 = r >> 24;
#pragma unroll
            for (int c = 0; c < 4; ++c) Yv[a][c] = (int)((H0[c] * w0 + H1[c] * w1) >> vsh);
        }
        if (NEED_UV && pr < ch2) {
            for (int n = (int)(rc & 0xffffu) - curc; n > 0; --n, ++curc, sl_c += sstride) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint32_t x0 = lds32(sl_c + lc.caddr[c]), x1 = lds32(sl_c + lc.caddr[c] + 4);
                    U0[c] = U1[c];
                    V0[c] = V1[c];
                    U1[c] = dp4a_uu(__byte_perm(x0, x1, lc.selu[c]), lc.cw[c], 0u);
                    V1[c] = dp4a_uu(__byte_perm(x0, x1, lc.selv[c]), lc.cw[c], 0u);
                }
            }
#pragma unroll
            for (int c = 0; c < 2; ++c)
                if ((j0 >> 1) + c < P.e.cw2) {
                    UV[c][0] = (int)((U0[c] * ((rc >> 16) & 0xffu) + U1[c] * (rc >> 24)) >> vsh);
                    UV[c][1] = (int)((V0[c] * ((rc >> 16) & 0xffu) + V1[c] * (rc >> 24)) >> vsh);
                }
        }
        const int nx = min(4, P.e.dw - j0), ny = min(2, dh - i0);
        if (nx > 0) {
            if (nx == 4 && ny == 2 && vec)
                emit_block<OK, 4>(P.e, outp, i0, j0, Yv, UV, 4, 2, true);
            else
                emit_block<OK, 4>(P.e, outp, i0, j0, Yv, UV, nx, ny, false);
        }
    }
}

// ---------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------
// MODE: MARCH_BICUBIC / MARCH_AREA.  VAR: bicubic: bit 0 = long horizontal digits, bit 1 = long vertical digits;
// area: NW (1, 2 or 4).
template <int MODE, int VAR, int OK>
__global__ void __launch_bounds__(kMarchMaxThreads, MODE == MARCH_BICUBIC ? TSVPP_MARCH_MINBLOCKS_BICUBIC : MODE == MARCH_BILINEAR ? TSVPP_MARCH_MINBLOCKS_BILINEAR : TSVPP_MARCH_MINBLOCKS_AREA)
    vpp_march_kernel(const MarchPlanDev P, const FrameDev *__restrict__ frames, int nframes, int vec_ok) {
    constexpr bool NEED_UV = !(OK == OUT_Y800_U8 || OK == OUT_Y800_F32);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NS = P.nstages;
    unsigned char *stages = smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)NS * P.stage_bytes);
    uint64_t *empty = full + kMaxStages;
    const int ncons_warps = ((int)blockDim.x >> 5) - 1;

    const int total = nframes * P.nstrips * P.ntiles;
    const int first = blockIdx.x;
    const int my_items = first < total ? (total - first + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(full + s, 1u);
            mbar_init(empty + s, (uint32_t)ncons_warps);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    if (warp == 0) {
        // ================= producer =================
        const uint64_t policy = l2_policy_evict_first();
        MarchIter ic;
        ic.init(P, first, (int)gridDim.x);
        int s = 0;
        uint32_t phase = 1;  // a fresh "empty" barrier passes a wait on parity 1
        for (int it = 0; it < my_items; ++it, ic.next(P)) {
            mbar_wait(empty + s, phase);
            march_fill<NEED_UV>(P, frames, ic, stages + (size_t)s * P.stage_bytes, full + s, lane, policy);
            if (++s == NS) {
                s = 0;
                phase ^= 1u;
            }
        }
        return;
    }

    // ================= consumers =================
    const int ctid = tid - 32;
    constexpr int BW = march_bw(MODE);
    const int blocks_x = P.tw / BW;
    const int tx = ctid % blocks_x;  // column block inside the tile
    const int g = ctid / blocks_x;   // row group inside the strip
    const bool active = g < P.G;
    const bool vec = vec_ok != 0;
    const int R2 = P.G * P.rpg;      // row pairs per strip
    const uint32_t sbase = smem_base_laundered();
    const uint32_t sstride = (uint32_t)P.slot_stride;

    int cur_tile = -1;
    // per-tile column state
    BicubicCols<(MODE == MARCH_BICUBIC) && ((VAR & 1) != 0)> bc;
    BilinearCols lc;
    int a_addr0 = 0, a_addr1 = 0, a_addrc = 0, a_sh0 = 0, a_sh1 = 0, a_shc = 0, a_pr0 = 0, a_pr1 = 0, a_prc = 0;
    constexpr int NW = (MODE == MARCH_AREA) ? VAR : 1;
    uint32_t a_w0[NW], a_w1[NW], a_wc[NW];

    MarchIter ic;
    ic.init(P, first, (int)gridDim.x);
    int s = 0;
    uint32_t phase = 0;
    for (int it = 0; it < my_items; ++it, ic.next(P)) {
        const uint32_t stage = sbase + (uint32_t)(s * P.stage_bytes);
        const uint32_t stage_c = stage + (uint32_t)P.nsl * sstride;
        const int j0 = ic.tile * P.tw + BW * tx;
        const bool have_c = NEED_UV && (j0 >> 1) < P.e.cw2;
        if (active && ic.tile != cur_tile) {
            cur_tile = ic.tile;
            const int cidx = ic.tile * P.tw + BW * tx;  // tables are padded to ntiles*tw columns
            if constexpr (MODE == MARCH_BILINEAR) {
                const MarchColL *cl = reinterpret_cast<const MarchColL *>(P.col_l) + cidx;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int4 v = __ldg(reinterpret_cast<const int4 *>(cl + c));
                    lc.addr[c] = v.x; lc.sel[c] = (uint32_t)v.y; lc.w[c] = (uint32_t)v.w;
                }
                if (NEED_UV) {
                    const MarchColL *cc = reinterpret_cast<const MarchColL *>(P.col_c) + (cidx >> 1);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int4 v = __ldg(reinterpret_cast<const int4 *>(cc + c));
                        lc.caddr[c] = v.x; lc.selu[c] = (uint32_t)v.y; lc.selv[c] = (uint32_t)v.z; lc.cw[c] = (uint32_t)v.w;
                    }
                }
            } else if constexpr (MODE == MARCH_BICUBIC) {
                const MarchColB *cl = reinterpret_cast<const MarchColB *>(P.col_l) + cidx;
                const MarchColB *cc = NEED_UV ? reinterpret_cast<const MarchColB *>(P.col_c) + (cidx >> 1) : nullptr;
                bc.load(cl, cc);
            } else {
                const MarchColA *cl = reinterpret_cast<const MarchColA *>(P.col_l) + cidx;
                const int4 c0 = __ldg(reinterpret_cast<const int4 *>(cl));
                const int4 c1 = __ldg(reinterpret_cast<const int4 *>(cl + 1));
                a_addr0 = c0.x; a_sh0 = c0.y; a_pr0 = c0.z;
                a_addr1 = c1.x; a_sh1 = c1.y; a_pr1 = c1.z;
#pragma unroll
                for (int k = 0; k < NW; ++k) {
                    a_w0[k] = (uint32_t)__ldg(P.px_w + a_pr0 * NW + k);
                    a_w1[k] = (uint32_t)__ldg(P.px_w + a_pr1 * NW + k);
                }
                if (NEED_UV) {
                    const int4 cc = __ldg(reinterpret_cast<const int4 *>(reinterpret_cast<const MarchColA *>(P.col_c) + (cidx >> 1)));
                    a_addrc = cc.x; a_shc = cc.y; a_prc = cc.z;
#pragma unroll
                    for (int k = 0; k < NW; ++k) a_wc[k] = (uint32_t)__ldg(P.px_w + a_prc * NW + k);
                }
            }
        }

        mbar_wait(full + s, phase);

        if (active) {
            void *outp = frames[ic.frame].out;
            if constexpr (MODE == MARCH_BILINEAR) {
                march_bilinear_group<OK, NEED_UV>(P, lc, stage, stage_c, sstride, ic.strip, g, j0, vec, outp);
            } else if constexpr (MODE == MARCH_BICUBIC) {
                march_bicubic_group<(VAR & 1) != 0, (VAR & 2) != 0, OK, NEED_UV>(P, bc, stage, stage_c, sstride, ic.strip, g, j0,
                                                                                  have_c, vec, outp);
            } else {
                // ---- area ----
                const MarchRowA *rowl = reinterpret_cast<const MarchRowA *>(P.row_l);
                const MarchRowA *rowc = reinterpret_cast<const MarchRowA *>(P.row_c);
                const int K = P.K;
                const int pr0 = ic.strip * R2 + g * P.rpg;  // first row pair (global index) of this thread's group
                for (int p = 0; p < P.rpg; ++p) {
                    const int pr = pr0 + p;
                    const int i0 = 2 * pr;
                    if (i0 >= P.e.dh) break;
                    int Yv[2][2];
                    int UV[1][2] = {{0, 0}};
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        const int2 ra = __ldg(reinterpret_cast<const int2 *>(rowl + i0 + a));
                        const int *pyw = P.py_w + ra.y * K;
                        uint32_t acc0 = 0, acc1 = 0;
                        uint32_t sl = stage + (uint32_t)ra.x * sstride;
                        for (int k = 0; k < K; ++k, sl += sstride) {
                            const uint32_t wy = (uint32_t)__ldg(pyw + k);
                            if (wy == 0) continue;  // zero padding of the pattern row (uniform)
                            uint32_t h0 = 0, h1 = 0;
                            uint32_t lo0 = lds32(sl + a_addr0), lo1 = lds32(sl + a_addr1);
#pragma unroll
                            for (int w = 0; w < NW; ++w) {
                                const uint32_t hi0 = lds32(sl + a_addr0 + 4 * (w + 1));
                                const uint32_t hi1 = lds32(sl + a_addr1 + 4 * (w + 1));
                                h0 = dp4a_uu(__funnelshift_r(lo0, hi0, a_sh0), a_w0[w], h0);
                                h1 = dp4a_uu(__funnelshift_r(lo1, hi1, a_sh1), a_w1[w], h1);
                                lo0 = hi0;
                                lo1 = hi1;
                            }
                            acc0 += wy * h0;
                            acc1 += wy * h1;
                        }
                        const int dj0 = ra.y * P.PX + a_pr0, dj1 = ra.y * P.PX + a_pr1;
                        Yv[a][0] = (int)(__umulhi(acc0, __ldg(P.div_mul + dj0)) >> __ldg(P.div_shift + dj0));
                        Yv[a][1] = (int)(__umulhi(acc1, __ldg(P.div_mul + dj1)) >> __ldg(P.div_shift + dj1));
                    }
                    if (NEED_UV && pr < P.e.ch2 && have_c) {
                        const int2 ra = __ldg(reinterpret_cast<const int2 *>(rowc + pr));
                        const int *pyw = P.py_w + ra.y * K;
                        uint32_t accu = 0, accv = 0;
                        uint32_t sl = stage_c + (uint32_t)ra.x * sstride + (uint32_t)a_addrc;
                        for (int k = 0; k < K; ++k, sl += sstride) {
                            const uint32_t wy = (uint32_t)__ldg(pyw + k);
                            if (wy == 0) continue;
                            uint32_t hu = 0, hv = 0;
                            uint32_t lo = lds32(sl);
#pragma unroll
                            for (int w = 0; w < NW; ++w) {
                                const uint32_t m1 = lds32(sl + 8 * w + 4), m2 = lds32(sl + 8 * w + 8);
                                const uint32_t x0 = __funnelshift_r(lo, m1, a_shc), x1 = __funnelshift_r(m1, m2, a_shc);
                                hu = dp4a_uu(__byte_perm(x0, x1, 0x6420), a_wc[w], hu);  // U bytes of 4 pairs
                                hv = dp4a_uu(__byte_perm(x0, x1, 0x7531), a_wc[w], hv);  // V bytes
                                lo = m2;
                            }
                            accu += wy * hu;
                            accv += wy * hv;
                        }
                        const int dj = ra.y * P.PX + a_prc;
                        const uint32_t mul = __ldg(P.div_mul + dj);
                        const int sh = __ldg(P.div_shift + dj);
                        UV[0][0] = (int)(__umulhi(accu, mul) >> sh);
                        UV[0][1] = (int)(__umulhi(accv, mul) >> sh);
                    }
                    const int nx = min(2, P.e.dw - j0), ny = min(2, P.e.dh - i0);
                    if (nx > 0) {
                        if (nx == 2 && ny == 2 && vec)
                            emit_block<OK, 2>(P.e, outp, i0, j0, Yv, UV, 2, 2, true);
                        else
                            emit_block<OK, 2>(P.e, outp, i0, j0, Yv, UV, nx, ny, false);
                    }
                }
            }
        }
        // all shared-memory reads of this stage are done: hand it back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
        if (++s == NS) {
            s = 0;
            phase ^= 1u;
        }
    }
}

// ---------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------
using MarchFn = void (*)(const MarchPlanDev, const FrameDev *, int, int);

template <int MODE, int VAR>
static MarchFn pick_ok(int ok) {
    switch (ok) {
        case OUT_RGB_PLANAR_U8: return vpp_march_kernel<MODE, VAR, OUT_RGB_PLANAR_U8>;
        case OUT_RGB_PLANAR_F32: return vpp_march_kernel<MODE, VAR, OUT_RGB_PLANAR_F32>;
        case OUT_RGB_MERGED_U8: return vpp_march_kernel<MODE, VAR, OUT_RGB_MERGED_U8>;
        case OUT_RGB_MERGED_F32: return vpp_march_kernel<MODE, VAR, OUT_RGB_MERGED_F32>;
        case OUT_Y800_U8: return vpp_march_kernel<MODE, VAR, OUT_Y800_U8>;
        case OUT_Y800_F32: return vpp_march_kernel<MODE, VAR, OUT_Y800_F32>;
        case OUT_HSV_F32: return vpp_march_kernel<MODE, VAR, OUT_HSV_F32>;
        case OUT_NV12_U8: return vpp_march_kernel<MODE, VAR, OUT_NV12_U8>;
        case OUT_NV12_F32: return vpp_march_kernel<MODE, VAR, OUT_NV12_F32>;
    }
    return nullptr;
}
static MarchFn pick_march(const MarchPlanDev &P, int ok) {
    if (P.mode == MARCH_BILINEAR) return pick_ok<MARCH_BILINEAR, 0>(ok);
    if (P.mode == MARCH_BICUBIC) {
        switch ((P.xlong ? 1 : 0) | (P.ylong ? 2 : 0)) {
            case 0: return pick_ok<MARCH_BICUBIC, 0>(ok);
            case 1: return pick_ok<MARCH_BICUBIC, 1>(ok);
            case 2: return pick_ok<MARCH_BICUBIC, 2>(ok);
            default: return pick_ok<MARCH_BICUBIC, 3>(ok);
        }
    }
    switch (P.NW) {
        case 1: return pick_ok<MARCH_AREA, 1>(ok);
        case 2: return pick_ok<MARCH_AREA, 2>(ok);
        case 4: return pick_ok<MARCH_AREA, 4>(ok);
    }
    return nullptr;
}

int march_block_threads(const MarchPlanDev &P) {
    const int cons = (P.tw / march_bw(P.mode)) * P.G;
    return 32 + (cons + 31) / 32 * 32;
}
size_t march_smem_bytes(const MarchPlanDev &P) { return (size_t)P.nstages * P.stage_bytes + 2 * 8 * kMaxStages + 128; }

int configure_march(const MarchPlanDev &P, int ok, size_t smem_optin_max) {
    MarchFn fn = pick_march(P, ok);
    if (!fn) return (int)cudaErrorMissingConfiguration;
    return (int)cudaFuncSetAttribute(reinterpret_cast<const void *>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_optin_max);
}
int march_max_ctas_per_sm(const MarchPlanDev &P, int ok, int threads, size_t smem_bytes) {
    MarchFn fn = pick_march(P, ok);
    if (!fn) return 0;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, reinterpret_cast<const void *>(fn), threads, smem_bytes) !=
        cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}
int launch_march(const MarchPlanDev &P, int ok, const FrameDev *frames_dev, int nframes, int vec_ok, int grid,
                 size_t smem_bytes, cudaStream_t stream) {
    MarchFn fn = pick_march(P, ok);
    if (!fn) return (int)cudaErrorMissingConfiguration;
    fn<<<grid, march_block_threads(P), smem_bytes, stream>>>(P, frames_dev, nframes, vec_ok);
    return (int)cudaGetLastError();
}

}  // namespace tsvpp
