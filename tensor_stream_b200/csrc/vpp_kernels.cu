// vpp_kernels.cu -- the fused VPP kernels for sm_100a (B200).
//
// One kernel instantiation per (resample kind, output kind) pair replaces the reference's
// crop kernel + resize kernel + colour-conversion kernel(s) and the uint8 NV12 intermediates
// they pass through HBM (reference src/Crop.cu:4-21, src/Resize.cu:180-357,
// src/ColorConversion.cu:6-278, dispatched by src/VideoProcessor.cpp:106-143).
//
// Structure (one persistent CTA loops over work items = (frame, strip of 2*rp output rows,
// tile of tw output columns)):
//   producer  : warp 0 issues one cp.async.bulk (TMA, SASS UBLKCP) per needed source-row
//               segment into a ring of shared-memory stages, completion tracked by mbarriers;
//               only the source rows / bytes the taps touch are ever fetched from HBM.
//   phase A   : every thread resamples its output column from the staged rows with the
//               reference's exact arithmetic (explicit __f*_rn intrinsics so the FMA positions
//               equal the reference's nvcc -fmad=true contraction) and rounds through uint8 --
//               the reference's NV12 intermediate lives in shared memory only.
//   phase B   : 4 consecutive pixels per thread: BT.601 conversion, optional /255, layout;
//               float4 / uchar4 stores.
// Bit-exactness contract: see oracle/vpp_oracle.c (same formulas, same citations).
#include <cuda_runtime.h>
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
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// ---------------------------------------------------------------------------------------
// Arithmetic shared with the oracle (each cites the reference lines it reproduces)
// ---------------------------------------------------------------------------------------

// reference src/Resize.cu:5-25 (calculateBillinearInterpolation), FMA positions as compiled:
//   v = fma(A*(1-wx), (1-wy), (B*wx)*(1-wy)); v = fma((1-wx), C*wy, v); v = fma(D, wx*wy, v)
__device__ __forceinline__ int bilinear_px(int a, int b, int c, int d, float wx, float wy) {
    const float A = (float)a, B = (float)b, C = (float)c, D = (float)d;
    const float omx = __fsub_rn(1.0f, wx), omy = __fsub_rn(1.0f, wy);
    float v = __fmaf_rn(__fmul_rn(A, omx), omy, __fmul_rn(__fmul_rn(B, wx), omy));
    v = __fmaf_rn(omx, __fmul_rn(C, wy), v);
    v = __fmaf_rn(D, __fmul_rn(wx, wy), v);
    return __float2int_rz(v) & 0xFF;
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

// reference src/ColorConversion.cu:6-39 (NV12toRGB24Kernel); chroma terms are shared by the two
// pixels of a pair.  R=(int)fma(Y',1.164,fma(V,1.596,.5)) etc.
struct ChromaTerms {
    float r, g, b;
};
__device__ __forceinline__ ChromaTerms chroma_terms(int U, int V) {
    const float fu = (float)(U - 128), fv = (float)(V - 128);
    ChromaTerms t;
    t.r = __fmaf_rn(fv, 1.5959997177f, 0.5f);
    t.b = __fmaf_rn(fu, 2.017999649f, 0.5f);
    float g = __fmul_rn(fu, -0.390999794f);
    g = __fmaf_rn(fv, -0.812999725f, g);
    t.g = __fadd_rn(g, 0.5f);
    return t;
}
__device__ __forceinline__ void yuv_px(int Y, const ChromaTerms &t, int &R, int &G, int &B) {
    const float yp = fmaxf(0.0f, __fsub_rn((float)Y, 16.0f));
    R = min(max(__float2int_rz(__fmaf_rn(yp, 1.163999557f, t.r)), 0), 255);
    G = min(max(__float2int_rz(__fmaf_rn(yp, 1.163999557f, t.g)), 0), 255);
    B = min(max(__float2int_rz(__fmaf_rn(yp, 1.163999557f, t.b)), 0), 255);
}

// (float)c / 255.0f with IEEE rounding for c in 0..255: reciprocal multiply plus one residual
// correction; verified exhaustively against the division (tests/test_host_logic.py).
__device__ __forceinline__ float div255(int c) {
    const float x = (float)c;
    const float r = 0x1.010102p-8f;  // RN(1/255)
    const float q = __fmul_rn(x, r);
    const float e = __fmaf_rn(-q, 255.0f, x);
    return __fmaf_rn(e, r, q);
}

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
// Stage fill
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

// Producer: executed by warp 0.  One bulk copy per (output row of the strip, vertical tap).
template <bool NEED_UV>
__device__ __forceinline__ void issue_stage_bulk(const PlanDev &P, const FrameDev *__restrict__ frames, int item,
                                                 unsigned char *stage, uint64_t *bar, int lane, uint64_t policy) {
    const ItemCoord ic = decode_item(P, item);
    const FrameDev fr = frames[ic.frame];
    const int nl = 2 * P.rp * P.K;                 // luma slots
    const int nc = NEED_UV ? P.rp * P.K : 0;       // chroma slots
    const int a0l = P.seg_a0_l[ic.tile], lenl = P.seg_len_l[ic.tile];
    const int a0c = NEED_UV ? P.seg_a0_c[ic.tile] : 0, lenc = NEED_UV ? P.seg_len_c[ic.tile] : 0;
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)(nl * lenl + nc * lenc));
    __syncwarp();
    const int *rs_l = P.rowsrc_l + (size_t)ic.strip * nl;
    for (int c = lane; c < nl; c += 32) {
        const int row = rs_l[c];
        bulk_g2s(stage + (size_t)c * P.slot_stride, fr.y + (size_t)row * P.pitch_y + a0l, lenl, bar, policy);
    }
    if (NEED_UV) {
        const int *rs_c = P.rowsrc_c + (size_t)ic.strip * nc;
        for (int c = lane; c < nc; c += 32) {
            const int row = rs_c[c];
            bulk_g2s(stage + (size_t)(nl + c) * P.slot_stride, fr.uv + (size_t)row * P.pitch_uv + a0c, lenc, bar,
                     policy);
        }
    }
}

// Fallback fill for frames whose base/pitch are not 16-byte aligned: guarded byte loads by the
// whole CTA (correct for any alignment; bytes past the plane end read as 0).
template <bool NEED_UV>
__device__ __forceinline__ void fill_stage_ldg(const PlanDev &P, const FrameDev *__restrict__ frames, int item,
                                               unsigned char *stage, int tid) {
    const ItemCoord ic = decode_item(P, item);
    const FrameDev fr = frames[ic.frame];
    const int nl = 2 * P.rp * P.K;
    const int nc = NEED_UV ? P.rp * P.K : 0;
    const int a0l = P.seg_a0_l[ic.tile], lenl = P.seg_len_l[ic.tile];
    const int *rs_l = P.rowsrc_l + (size_t)ic.strip * nl;
    for (int c = 0; c < nl; ++c) {
        const long base = (long)rs_l[c] * P.pitch_y + a0l;
        unsigned char *dst = stage + (size_t)c * P.slot_stride;
        for (int b = tid; b < lenl; b += kThreads) {
            const long g = base + b;
            dst[b] = (g < P.plane_bytes_y) ? __ldg(fr.y + g) : (unsigned char)0;
        }
    }
    if (NEED_UV) {
        const int a0c = P.seg_a0_c[ic.tile], lenc = P.seg_len_c[ic.tile];
        const int *rs_c = P.rowsrc_c + (size_t)ic.strip * nc;
        for (int c = 0; c < nc; ++c) {
            const long base = (long)rs_c[c] * P.pitch_uv + a0c;
            unsigned char *dst = stage + (size_t)(nl + c) * P.slot_stride;
            for (int b = tid; b < lenc; b += kThreads) {
                const long g = base + b;
                dst[b] = (g < P.plane_bytes_uv) ? __ldg(fr.uv + g) : (unsigned char)0;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Phase A: resample one output sample (one byte of the resized NV12 strip)
// ---------------------------------------------------------------------------------------
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

// `slots` points at slot 0 of this output row; `orow` is the output row index used to look up
// the vertical weights (luma: i, chroma: ci -- the reference evaluates the *luma-grid* formula
// at the chroma index, src/Resize.cu:303-310), `stride` is 1 for luma and 2 for chroma bytes.
template <int RS>
__device__ __forceinline__ int resample(const PlanDev &P, const unsigned char *slots,
                                        const typename ColOf<RS>::T &col, int orow, int stride) {
    if constexpr (RS == RS_NEAREST) {
        const ColNearest &c = col;
        return slots[c.o];
    } else if constexpr (RS == RS_BILINEAR) {
        const ColBilinear &c = col;
        const unsigned char *s0 = slots, *s1 = slots + P.slot_stride;
        const float wy = __ldg(P.wy + orow);
        return bilinear_px(s0[c.o0], s0[c.o1], s1[c.o0], s1[c.o1], c.w, wy);
    } else if constexpr (RS == RS_BICUBIC) {
        const ColBicubic &c = col;
        int b[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned char *s = slots + k * P.slot_stride;
            b[k] = bicubic_dot(c.c[0], c.c[1], c.c[2], c.c[3], s[c.o[0]], s[c.o[1]], s[c.o[2]], s[c.o[3]]);
        }
        const double *cy = P.cy + (size_t)orow * 4;
        return bicubic_dot(__ldg(cy + 0), __ldg(cy + 1), __ldg(cy + 2), __ldg(cy + 3), b[0], b[1], b[2], b[3]);
    } else {
        // reference src/Resize.cu:160-178 (calculateAreaInterpolation): rows outer, columns
        // inner; w = px*py; div += w; sum = fma(d, w, sum); (int)(sum/div)
        const ColArea &c = col;
        const float *px = P.px + (size_t)c.prow * P.KX;
        const float *py = P.py + (size_t)(orow % P.PY) * P.K;
        float sum = 0.0f, div = 0.0f;
        for (int a = 0; a < P.K; ++a) {
            const unsigned char *s = slots + a * P.slot_stride + c.o;
            const float wyv = __ldg(py + a);
            for (int b = 0; b < P.KX; ++b) {
                const float w = __fmul_rn(__ldg(px + b), wyv);
                div = __fadd_rn(div, w);
                sum = __fmaf_rn((float)s[b * stride], w, sum);
            }
        }
        return __float2int_rz(__fdiv_rn(sum, div)) & 0xFF;
    }
}

// ---------------------------------------------------------------------------------------
// Phase B: colour conversion + layout for 4 consecutive pixels of one output row
// ---------------------------------------------------------------------------------------
template <int OK>
__device__ __forceinline__ void emit4(const PlanDev &P, void *outp, int i, int j0, uint32_t y4, uint32_t uv4,
                                      bool vec_ok) {
    const int dw = P.dw, dh = P.dh;
    const int nvalid = min(4, dw - j0);
    int Yv[4] = {(int)(y4 & 0xFF), (int)((y4 >> 8) & 0xFF), (int)((y4 >> 16) & 0xFF), (int)(y4 >> 24)};

    if (OK == OUT_Y800_U8) {
        unsigned char *o = (unsigned char *)outp + (size_t)i * dw + j0;
        if (vec_ok && nvalid == 4) {
            *reinterpret_cast<uint32_t *>(o) = y4;
        } else {
            for (int k = 0; k < nvalid; ++k) o[k] = (unsigned char)Yv[k];
        }
        return;
    }
    if (OK == OUT_Y800_F32) {
        float *o = (float *)outp + (size_t)i * dw + j0;
        float f[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) f[k] = div255(Yv[k]);
        if (vec_ok && nvalid == 4) {
            __stcs(reinterpret_cast<float4 *>(o), make_float4(f[0], f[1], f[2], f[3]));
        } else {
            for (int k = 0; k < nvalid; ++k) o[k] = f[k];
        }
        return;
    }
    if (OK == OUT_NV12_U8 || OK == OUT_NV12_F32) {
        // reference src/ColorConversion.cu:211-233 (NV12MergeBuffers): Y plane then UV plane
        const size_t wh = (size_t)dw * dh;
        const bool do_uv = ((i & 1) == 0) && ((i >> 1) < P.ch2);
        if (OK == OUT_NV12_U8) {
            unsigned char *o = (unsigned char *)outp;
            for (int k = 0; k < nvalid; ++k) o[(size_t)i * dw + j0 + k] = (unsigned char)Yv[k];
            if (do_uv)
                for (int k = 0; k < nvalid; ++k)
                    if (((j0 + k) >> 1) < P.cw2) o[wh + (size_t)(i >> 1) * dw + j0 + k] = (unsigned char)((uv4 >> (8 * k)) & 0xFF);
        } else {
            float *o = (float *)outp;
            for (int k = 0; k < nvalid; ++k) o[(size_t)i * dw + j0 + k] = div255(Yv[k]);
            if (do_uv)
                for (int k = 0; k < nvalid; ++k)
                    if (((j0 + k) >> 1) < P.cw2) o[wh + (size_t)(i >> 1) * dw + j0 + k] = div255((int)((uv4 >> (8 * k)) & 0xFF));
        }
        return;
    }

    // RGB family
    const ChromaTerms t0 = chroma_terms((int)(uv4 & 0xFF), (int)((uv4 >> 8) & 0xFF));
    const ChromaTerms t1 = chroma_terms((int)((uv4 >> 16) & 0xFF), (int)(uv4 >> 24));
    int R[4], G[4], B[4];
    yuv_px(Yv[0], t0, R[0], G[0], B[0]);
    yuv_px(Yv[1], t0, R[1], G[1], B[1]);
    yuv_px(Yv[2], t1, R[2], G[2], B[2]);
    yuv_px(Yv[3], t1, R[3], G[3], B[3]);

    if (OK == OUT_HSV_F32) {
        // reference src/ColorConversion.cu:360-372: from normalised RGB (never BGR), merged
        float *o = (float *)outp + ((size_t)i * dw + j0) * 3;
        float v[12];
#pragma unroll
        for (int k = 0; k < 4; ++k) rgb_to_hsv(div255(R[k]), div255(G[k]), div255(B[k]), v[3 * k], v[3 * k + 1], v[3 * k + 2]);
        if (vec_ok && nvalid == 4) {
            float4 *o4 = reinterpret_cast<float4 *>(o);
            __stcs(o4 + 0, make_float4(v[0], v[1], v[2], v[3]));
            __stcs(o4 + 1, make_float4(v[4], v[5], v[6], v[7]));
            __stcs(o4 + 2, make_float4(v[8], v[9], v[10], v[11]));
        } else {
            for (int k = 0; k < 3 * nvalid; ++k) o[k] = v[k];
        }
        return;
    }

    if (P.swap_rb) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int t = R[k];
            R[k] = B[k];
            B[k] = t;
        }
    }
    if (OK == OUT_RGB_PLANAR_F32) {
        const size_t plane = (size_t)dw * dh;
        float *o = (float *)outp + (size_t)i * dw + j0;
        if (vec_ok && nvalid == 4) {
            __stcs(reinterpret_cast<float4 *>(o), make_float4(div255(R[0]), div255(R[1]), div255(R[2]), div255(R[3])));
            __stcs(reinterpret_cast<float4 *>(o + plane), make_float4(div255(G[0]), div255(G[1]), div255(G[2]), div255(G[3])));
            __stcs(reinterpret_cast<float4 *>(o + 2 * plane), make_float4(div255(B[0]), div255(B[1]), div255(B[2]), div255(B[3])));
        } else {
            for (int k = 0; k < nvalid; ++k) {
                o[k] = div255(R[k]);
                o[plane + k] = div255(G[k]);
                o[2 * plane + k] = div255(B[k]);
            }
        }
    } else if (OK == OUT_RGB_PLANAR_U8) {
        const size_t plane = (size_t)dw * dh;
        unsigned char *o = (unsigned char *)outp + (size_t)i * dw + j0;
        if (vec_ok && nvalid == 4) {
            *reinterpret_cast<uint32_t *>(o) = (uint32_t)R[0] | ((uint32_t)R[1] << 8) | ((uint32_t)R[2] << 16) | ((uint32_t)R[3] << 24);
            *reinterpret_cast<uint32_t *>(o + plane) = (uint32_t)G[0] | ((uint32_t)G[1] << 8) | ((uint32_t)G[2] << 16) | ((uint32_t)G[3] << 24);
            *reinterpret_cast<uint32_t *>(o + 2 * plane) = (uint32_t)B[0] | ((uint32_t)B[1] << 8) | ((uint32_t)B[2] << 16) | ((uint32_t)B[3] << 24);
        } else {
            for (int k = 0; k < nvalid; ++k) {
                o[k] = (unsigned char)R[k];
                o[plane + k] = (unsigned char)G[k];
                o[2 * plane + k] = (unsigned char)B[k];
            }
        }
    } else if (OK == OUT_RGB_MERGED_F32) {
        float *o = (float *)outp + ((size_t)i * dw + j0) * 3;
        if (vec_ok && nvalid == 4) {
            float4 *o4 = reinterpret_cast<float4 *>(o);
            __stcs(o4 + 0, make_float4(div255(R[0]), div255(G[0]), div255(B[0]), div255(R[1])));
            __stcs(o4 + 1, make_float4(div255(G[1]), div255(B[1]), div255(R[2]), div255(G[2])));
            __stcs(o4 + 2, make_float4(div255(B[2]), div255(R[3]), div255(G[3]), div255(B[3])));
        } else {
            for (int k = 0; k < nvalid; ++k) {
                o[3 * k] = div255(R[k]);
                o[3 * k + 1] = div255(G[k]);
                o[3 * k + 2] = div255(B[k]);
            }
        }
    } else if (OK == OUT_RGB_MERGED_U8) {
        unsigned char *o = (unsigned char *)outp + ((size_t)i * dw + j0) * 3;
        if (vec_ok && nvalid == 4) {
            uint32_t *o4 = reinterpret_cast<uint32_t *>(o);
            o4[0] = (uint32_t)R[0] | ((uint32_t)G[0] << 8) | ((uint32_t)B[0] << 16) | ((uint32_t)R[1] << 24);
            o4[1] = (uint32_t)G[1] | ((uint32_t)B[1] << 8) | ((uint32_t)R[2] << 16) | ((uint32_t)G[2] << 24);
            o4[2] = (uint32_t)B[2] | ((uint32_t)R[3] << 8) | ((uint32_t)G[3] << 16) | ((uint32_t)B[3] << 24);
        } else {
            for (int k = 0; k < nvalid; ++k) {
                o[3 * k] = (unsigned char)R[k];
                o[3 * k + 1] = (unsigned char)G[k];
                o[3 * k + 2] = (unsigned char)B[k];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// The fused kernel
// ---------------------------------------------------------------------------------------
template <int RS, int OK, bool BULK>
__global__ void __launch_bounds__(kThreads) vpp_fused_kernel(const PlanDev P, const FrameDev *__restrict__ frames,
                                                             int nframes, int vec_ok) {
    constexpr bool NEED_UV = !(OK == OUT_Y800_U8 || OK == OUT_Y800_F32);
    using ColT = typename ColOf<RS>::T;
    extern __shared__ __align__(128) unsigned char smem[];

    const int tid = threadIdx.x;
    const int NS = BULK ? P.nstages : 1;
    unsigned char *stages = smem;
    const int rbuf_bytes = 3 * P.rp * P.tw;
    unsigned char *rbuf = smem + (size_t)NS * P.stage_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(rbuf + 2 * rbuf_bytes);  // rbuf_bytes is a multiple of 32

    const int total = nframes * P.nstrips * P.ntiles;
    const int first = blockIdx.x;
    if (first >= total) return;
    const int my_items = (total - first + (int)gridDim.x - 1) / (int)gridDim.x;

    uint64_t policy = 0;
    if (BULK) {
        if (tid == 0) {
            for (int s = 0; s < NS; ++s) mbar_init(bars + s, 1);
            fence_barrier_init();
            fence_proxy_async();
        }
        __syncthreads();
        if (tid < 32) {
            policy = l2_policy_evict_first();
            for (int p = 0; p < NS - 1 && p < my_items; ++p)
                issue_stage_bulk<NEED_UV>(P, frames, first + p * gridDim.x, stages + (size_t)p * P.stage_bytes, bars + p,
                                          tid, policy);
        }
    }

    const int tw = P.tw;
    const int tx = tid & (tw - 1);
    const int ty = tid / tw;
    const int ny = kThreads / tw;
    const int nrg = (NEED_UV ? 3 : 2) * P.rp;  // row groups per strip: 2*rp luma (+ rp chroma)
    const int units_per_row = tw >> 2;

    for (int it = 0; it < my_items; ++it) {
        const int item = first + it * gridDim.x;
        const ItemCoord ic = decode_item(P, item);
        unsigned char *stage;
        if (BULK) {
            const int pf = it + NS - 1;
            if (tid < 32 && pf < my_items) {
                const int ps = pf % NS;
                issue_stage_bulk<NEED_UV>(P, frames, first + pf * gridDim.x, stages + (size_t)ps * P.stage_bytes,
                                          bars + ps, tid, policy);
            }
            stage = stages + (size_t)(it % NS) * P.stage_bytes;
        } else {
            stage = stages;
            fill_stage_ldg<NEED_UV>(P, frames, item, stage, tid);
        }

        // column parameters of this thread (issued before the barrier wait so the latency hides)
        const int col = ic.tile * tw + tx;
        ColT cl = reinterpret_cast<const ColT *>(P.col_l)[col];
        ColT cc;
        if (NEED_UV) cc = reinterpret_cast<const ColT *>(P.col_c)[col];

        if (BULK)
            mbar_wait(bars + (it % NS), (uint32_t)((it / NS) & 1));
        else
            __syncthreads();

        // ---- phase A: resized NV12 strip -> rbuf (bytes) ----
        unsigned char *rb = rbuf + (size_t)(it & 1) * rbuf_bytes;
        for (int rg = ty; rg < nrg; rg += ny) {
            int v = 0;
            if (rg < 2 * P.rp) {
                const int i = ic.strip * 2 * P.rp + rg;
                if (i < P.dh && col < P.dw)
                    v = resample<RS>(P, stage + (size_t)rg * P.K * P.slot_stride, cl, i, 1);
            } else if (NEED_UV) {
                const int ci = ic.strip * P.rp + (rg - 2 * P.rp);
                if (ci < P.ch2 && (col >> 1) < P.cw2)
                    v = resample<RS>(P, stage + (size_t)rg * P.K * P.slot_stride, cc, ci, 2);
            }
            rb[rg * tw + tx] = (unsigned char)v;
        }
        __syncthreads();

        // ---- phase B: colour conversion + layout ----
        void *outp = frames[ic.frame].out;
        const int nunits = 2 * P.rp * units_per_row;
        for (int u = tid; u < nunits; u += kThreads) {
            const int r = u / units_per_row;
            const int c4 = u - r * units_per_row;
            const int i = ic.strip * 2 * P.rp + r;
            const int j0 = ic.tile * tw + 4 * c4;
            if (i < P.dh && j0 < P.dw) {
                const uint32_t y4 = *reinterpret_cast<const uint32_t *>(rb + r * tw + 4 * c4);
                const uint32_t uv4 =
                    NEED_UV ? *reinterpret_cast<const uint32_t *>(rb + (2 * P.rp + (r >> 1)) * tw + 4 * c4) : 0u;
                emit4<OK>(P, outp, i, j0, y4, uv4, vec_ok != 0);
            }
        }
        if (!BULK) __syncthreads();  // stage is refilled by everyone next iteration
    }
}

// ---------------------------------------------------------------------------------------
// Bicubic weight tables are built on the device so that pow() is the very libdevice routine
// the compiled reference calls (reference src/Resize.cu:45-48, weights folded by nvcc to:
//   t=w2*1.5; q=w3*-.75; c0=fma(w,-.75,t)+q; c1=fma(w3,1.25,fma(w2,-2.25,1));
//   c2=fma(w3,-1.25,fma(w,.75,t)); c3=fma(w2,-.75,-q)   [cuobjdump -sass, sm_100a])
// ---------------------------------------------------------------------------------------
__global__ void bicubic_weights_kernel(const double *__restrict__ w, double *__restrict__ c, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = w[i];
    const double w2 = pow(x, 2.0), w3 = pow(x, 3.0);
    const double t = __dmul_rn(w2, 1.5), q = __dmul_rn(w3, -0.75);
    c[4 * i + 0] = __dadd_rn(__fma_rn(x, -0.75, t), q);
    c[4 * i + 1] = __fma_rn(w3, 1.25, __fma_rn(w2, -2.25, 1.0));
    c[4 * i + 2] = __fma_rn(w3, -1.25, __fma_rn(x, 0.75, t));
    c[4 * i + 3] = __fma_rn(w2, -0.75, -q);
}

// pow(w,2), pow(w,3) as the device evaluates them (exported for the tests: the CPU oracle is
// pinned to these so that the comparison isolates everything but libdevice's pow)
__global__ void pow23_kernel(const double *__restrict__ w, double *__restrict__ w2, double *__restrict__ w3, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    w2[i] = pow(w[i], 2.0);
    w3[i] = pow(w[i], 3.0);
}
void launch_pow23(const double *w_dev, double *w2_dev, double *w3_dev, int n, cudaStream_t stream) {
    pow23_kernel<<<(n + 127) / 128, 128, 0, stream>>>(w_dev, w2_dev, w3_dev, n);
}

void launch_bicubic_weights(const double *w_dev, double *c_dev, int n, cudaStream_t stream) {
    bicubic_weights_kernel<<<(n + 127) / 128, 128, 0, stream>>>(w_dev, c_dev, n);
}

// ---------------------------------------------------------------------------------------
// Host-side dispatch over the (resample, output, fill) instantiations
// ---------------------------------------------------------------------------------------
using KernelFn = void (*)(const PlanDev, const FrameDev *, int, int);

template <int RS, int OK>
static KernelFn pick_fill(bool bulk) {
    return bulk ? vpp_fused_kernel<RS, OK, true> : vpp_fused_kernel<RS, OK, false>;
}
template <int RS>
static KernelFn pick_out(int ok, bool bulk) {
    switch (ok) {
        case OUT_RGB_PLANAR_U8: return pick_fill<RS, OUT_RGB_PLANAR_U8>(bulk);
        case OUT_RGB_PLANAR_F32: return pick_fill<RS, OUT_RGB_PLANAR_F32>(bulk);
        case OUT_RGB_MERGED_U8: return pick_fill<RS, OUT_RGB_MERGED_U8>(bulk);
        case OUT_RGB_MERGED_F32: return pick_fill<RS, OUT_RGB_MERGED_F32>(bulk);
        case OUT_Y800_U8: return pick_fill<RS, OUT_Y800_U8>(bulk);
        case OUT_Y800_F32: return pick_fill<RS, OUT_Y800_F32>(bulk);
        case OUT_HSV_F32: return pick_fill<RS, OUT_HSV_F32>(bulk);
        case OUT_NV12_U8: return pick_fill<RS, OUT_NV12_U8>(bulk);
        case OUT_NV12_F32: return pick_fill<RS, OUT_NV12_F32>(bulk);
    }
    return nullptr;
}
static KernelFn pick_kernel(int rs, int ok, bool bulk) {
    switch (rs) {
        case RS_NEAREST: return pick_out<RS_NEAREST>(ok, bulk);
        case RS_BILINEAR: return pick_out<RS_BILINEAR>(ok, bulk);
        case RS_BICUBIC: return pick_out<RS_BICUBIC>(ok, bulk);
        case RS_AREA: return pick_out<RS_AREA>(ok, bulk);
    }
    return nullptr;
}

size_t fused_smem_bytes(const PlanDev &P, bool bulk) {
    const int NS = bulk ? P.nstages : 1;
    return (size_t)NS * P.stage_bytes + 2 * (size_t)(3 * P.rp * P.tw) + 8 * kMaxStages + 128;
}

int configure_fused(int rs, int ok, bool bulk, size_t smem_bytes) {
    KernelFn fn = pick_kernel(rs, ok, bulk);
    if (!fn) return (int)cudaErrorMissingConfiguration;  // reference ColorConversion.cu:376
    return (int)cudaFuncSetAttribute(reinterpret_cast<const void *>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes);
}

int max_ctas_per_sm(int rs, int ok, bool bulk, size_t smem_bytes) {
    KernelFn fn = pick_kernel(rs, ok, bulk);
    if (!fn) return 0;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, reinterpret_cast<const void *>(fn), kThreads, smem_bytes) !=
        cudaSuccess)
        return 0;
    return n;
}

int launch_fused(int rs, int ok, bool bulk, const PlanDev &P, const FrameDev *frames_dev, int nframes, int vec_ok,
                 int grid, size_t smem_bytes, cudaStream_t stream) {
    KernelFn fn = pick_kernel(rs, ok, bulk);
    if (!fn) return (int)cudaErrorMissingConfiguration;
    fn<<<grid, kThreads, smem_bytes, stream>>>(P, frames_dev, nframes, vec_ok);
    return (int)cudaGetLastError();
}

}  // namespace tsvpp
