// vpp_kernels.cu -- the fused VPP kernels for sm_100a (B200).
//
// One kernel instantiation per (resample kind, output kind) pair replaces the reference's
// crop kernel + resize kernel + colour-conversion kernel(s) and the uint8 NV12 intermediates
// they pass through HBM (reference src/Crop.cu:4-21, src/Resize.cu:180-357,
// src/ColorConversion.cu:6-278, dispatched by src/VideoProcessor.cpp:106-143).
//
// Structure: persistent, warp-specialised CTAs loop over work items = (frame, strip of 2*rp
// output rows, tile of tw output columns).
//   producer warp : for each item, one cp.async.bulk (TMA unit, SASS UBLKCP) per needed
//                   source-row segment into a ring of shared-memory stages; completion is
//                   tracked by "full" mbarriers, reuse by "empty" mbarriers.  Only the source
//                   rows / bytes the taps touch are ever fetched from HBM.
//                   Frames whose base / pitch are not 16-byte aligned are staged by the TMA unit too:
//                   four producer warps bulk-copy the 16-byte aligned SUPERSET of every row segment
//                   and record per slot where the segment starts; the consumers add that byte
//                   (RowSlots<SH>).  Guarded loads only for the rows at the edges of a plane.
//   consumer warps: each thread owns a block of 2 rows x BW columns (BW = 2, or 4 for wide outputs
//                   of the light resamplers): the luma samples and the (U,V) pairs they share
//                   are resampled from the staged rows with the
//                   reference's exact arithmetic (explicit __f*_rn intrinsics so the FMA
//                   positions equal the reference's nvcc -fmad=true contraction), rounded
//                   through uint8 exactly where the reference stores its NV12 intermediate,
//                   then converted (BT.601), normalised and stored with vector stores.
//                   No block-wide barrier anywhere in the steady state.  Taps are ld.shared from
//                   addresses derived from one laundered base register (SPtr below).
// No-resize conversions on device-resident frames skip the staging altogether (vpp_direct_kernel,
// further down): nothing is reused between output blocks there, so coalesced loads + occupancy win.
// uint8 -> float uses cvt.rn.f32.u32 on 32-bit registers (SASS I2FP) and float -> uint8 one
// saturating cvt.rzi.u8.f32 (SASS F2IP: truncation and the [0,255] clamp in one instruction)
// instead of the byte-typed I2F/F2I forms the compiler picks for (float)uchar / (int)float
// (quarter-rate XU pipe: ncu showed XU as the busiest pipe of the first version).
// Bit-exactness contract: see oracle/vpp_oracle.c (same formulas, same citations).
#include "vpp_device.cuh"

namespace tsvpp {

// ---------------------------------------------------------------------------------------
// The fused kernel.  blockDim.x = 32 (producer warp) + round_up(tw/BW * rp, 32) consumers;
// every consumer thread owns a block of 2 rows x BW columns of the output.
// ---------------------------------------------------------------------------------------
template <int RS, int OK, bool BULK, int BW>
__global__ void __launch_bounds__((BW == 4 ? 160 : kMaxThreads) + (BULK ? (BW == 4 ? 32 : 0) : 32 * (kLdgProducerWarps - 1)),
                                  BULK ? (BW == 4 ? 5 : ((RS == RS_BICUBIC || RS == RS_AREA) ? 3 : 4)) : 2)
    vpp_fused_kernel(const PlanDev P, const FrameDev *__restrict__ frames, int nframes, int vec_ok) {
    constexpr bool NEED_UV = !(OK == OUT_Y800_U8 || OK == OUT_Y800_F32);
    using ColT = typename ColOf<RS>::T;
    constexpr bool SMALL_COL = sizeof(ColT) <= 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NS = P.nstages;
    unsigned char *stages = smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)NS * P.stage_bytes);
    uint64_t *empty = full + kMaxStages;
    const int NPROD = P.nprod;  // producer warps (bulk: 1, or 2 when the row segments are short; unaligned frames: 4)
    // unaligned frames: one byte per slot = where the row segment starts inside its slot (fill_stage_shifted)
    constexpr bool SH = !BULK;
    const int nslots_pad = (((NEED_UV ? 3 : 2) * P.rp * P.K) + 15) & ~15;
    unsigned char *shift_tab = smem + (size_t)NS * P.stage_bytes + 2 * 8 * kMaxStages;
    const int ncons_warps = ((int)blockDim.x >> 5) - NPROD;

    const int total = nframes * P.nstrips * P.ntiles;
    const int first = blockIdx.x;
    const int my_items = first < total ? (total - first + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(full + s, (uint32_t)NPROD);
            mbar_init(empty + s, (uint32_t)ncons_warps);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    if (warp < NPROD) {
        // ================= producer(s) =================
        const uint64_t policy = l2_policy_evict_first();
        ItemIter ic;
        ic.init(P, first, (int)gridDim.x);
        int s = 0;
        uint32_t phase = 1;  // a fresh "empty" barrier passes a wait on parity 1: every stage starts free
        for (int it = 0; it < my_items; ++it, ic.next(P)) {
            mbar_wait(empty + s, phase);
            unsigned char *stage = stages + (size_t)s * P.stage_bytes;
            if (BULK)
                issue_stage_bulk<NEED_UV>(P, frames, ic, stage, full + s, lane, policy, warp, NPROD);
            else
                fill_stage_shifted<NEED_UV>(P, frames, ic, stage, shift_tab + (size_t)s * nslots_pad, full + s, lane, policy,
                                            warp, NPROD);
            if (++s == NS) {
                s = 0;
                phase ^= 1u;
            }
        }
        return;
    }

    // ================= consumers =================
    const int ctid = tid - 32 * NPROD;
    const int blocks_x = P.tw / BW;
    const int tx = ctid % blocks_x;  // column block inside the tile
    const int ty = ctid / blocks_x;  // row pair inside the strip
    const bool active = ty < P.rp;
    const bool vec = vec_ok != 0;
    const ColT *col_l = reinterpret_cast<const ColT *>(P.col_l);
    const ColT *col_c = reinterpret_cast<const ColT *>(P.col_c);

    const uint32_t sbase = smem_base_laundered();
    int cur_tile = -1;
    ColT lreg[BW], creg[BW];  // column parameters (kept in registers when small)

    ItemIter ic;
    ic.init(P, first, (int)gridDim.x);
    int s = 0;
    uint32_t phase = 0;
    for (int it = 0; it < my_items; ++it, ic.next(P)) {
        const SPtr stage{sbase + (uint32_t)(s * P.stage_bytes)};
        const int colbase = ic.tile * P.tw + BW * tx;
        const int i0 = (ic.strip * P.rp + ty) * 2;  // first output row of this thread's block
        const int ci = ic.strip * P.rp + ty;        // its chroma row
        const int j0 = colbase;                     // first output column
        float wy0 = 0.f, wy1 = 0.f, wyc = 0.f;
        int row_shift = 4;
        if (active) {
            if (SMALL_COL && ic.tile != cur_tile) {
#pragma unroll
                for (int b = 0; b < BW; ++b) {
                    lreg[b] = ld_col(col_l + colbase + b);
                    if (NEED_UV) creg[b] = ld_col(col_c + colbase + b);
                }
                cur_tile = ic.tile;
            }
            if (RS == RS_BILINEAR) {
                wy0 = __ldg(P.wy + i0);
                wy1 = __ldg(P.wy + i0 + 1);
                wyc = __ldg(P.wy + ci);
            }
            if (RS == RS_BICUBIC) row_shift = __ldg(P.row_shift + ci);
        }

        mbar_wait(full + s, phase);

        int Yv[2][BW];
        int UV[BW / 2][2];
#pragma unroll
        for (int b = 0; b < BW; ++b) Yv[0][b] = Yv[1][b] = 0;
#pragma unroll
        for (int p = 0; p < BW / 2; ++p) UV[p][0] = UV[p][1] = 0;
        const int nx = min(BW, P.dw - j0), ny = min(2, P.dh - i0);
        const bool work = active && nx > 0 && ny > 0;
        bool sampled = false;
        if constexpr (RS == RS_NEAREST && SMALL_COL) {
            // no horizontal resize: the block's bytes are consecutive and aligned in the staged rows
            if (!SH && work && P.ident && nx == BW && ny == 2 && (!NEED_UV || (ci < P.ch2 && (j0 >> 1) + BW / 2 <= P.cw2))) {
                const SPtr sl0 = stage + ((2 * ty) * P.slot_stride + lreg[0].o);
                const uint32_t w0 = ld_group<BW>(sl0), w1 = ld_group<BW>(sl0 + P.slot_stride);
#pragma unroll
                for (int b = 0; b < BW; ++b) {
                    Yv[0][b] = (int)__byte_perm(w0, 0u, 0x4440u + b);  // byte b, zero-extended: one PRMT
                    Yv[1][b] = (int)__byte_perm(w1, 0u, 0x4440u + b);
                }
                if (NEED_UV) {
                    const uint32_t wc = ld_group<BW>(stage + ((2 * P.rp + ty) * P.slot_stride + creg[0].o));
#pragma unroll
                    for (int p = 0; p < BW / 2; ++p) {
                        UV[p][0] = (int)__byte_perm(wc, 0u, 0x4440u + 2 * p);
                        UV[p][1] = (int)__byte_perm(wc, 0u, 0x4440u + 2 * p + 1);
                    }
                }
                sampled = true;
            }
        }
        if (work && !sampled) {
            const uint32_t shs = sbase + (uint32_t)(NS * P.stage_bytes + 2 * 8 * kMaxStages + s * nslots_pad);  // this stage's shift bytes
            const RowSlots<SH> sl0{stage + (2 * ty) * P.K * P.slot_stride, (uint32_t)P.slot_stride, shs + (uint32_t)((2 * ty) * P.K)};
            const RowSlots<SH> sl1{sl0.base + P.K * P.slot_stride, (uint32_t)P.slot_stride, sl0.sh + (uint32_t)P.K};
#pragma unroll
            for (int b = 0; b < BW; ++b) {
                if (b < nx) {
                    const ColT &L = SMALL_COL ? lreg[b] : col_l[colbase + b];
                    if constexpr (RS == RS_BICUBIC) {
                        bicubic_luma_pair(P, sl0, sl1, L, i0, row_shift, ny > 1, Yv[0][b], Yv[1][b]);
                    } else {
                        Yv[0][b] = resample<RS, SH>(P, sl0, L, i0, 1, wy0);
                        if (ny > 1) Yv[1][b] = resample<RS, SH>(P, sl1, L, i0 + 1, 1, wy1);
                    }
                }
            }
            if (NEED_UV && ci < P.ch2) {
                const RowSlots<SH> sc{stage + (2 * P.rp + ty) * P.K * P.slot_stride, (uint32_t)P.slot_stride,
                                      shs + (uint32_t)((2 * P.rp + ty) * P.K)};
#pragma unroll
                for (int p = 0; p < BW / 2; ++p) {
                    if ((j0 >> 1) + p < P.cw2) {
                        const ColT &C0 = SMALL_COL ? creg[2 * p] : col_c[colbase + 2 * p];
                        const ColT &C1 = SMALL_COL ? creg[2 * p + 1] : col_c[colbase + 2 * p + 1];
                        UV[p][0] = resample<RS, SH>(P, sc, C0, ci, 2, wyc);
                        UV[p][1] = resample<RS, SH>(P, sc, C1, ci, 2, wyc);
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

        if (work) {
            void *outp = frames[ic.frame].out;
            if (nx == BW && ny == 2 && vec)
                emit_block<OK, BW>(P, outp, i0, j0, Yv, UV, BW, 2, true);  // constants fold: straight-line vector stores
            else
                emit_block<OK, BW>(P, outp, i0, j0, Yv, UV, nx, ny, false);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Direct kernel: colour conversion / layout at the source size (the reference's default read():
// no resize, optional crop).  Nothing is re-used between output blocks, so there is nothing to
// stage: a warp takes 32*kDirectSub consecutive 2x4 blocks of a frame (row-major over blocks), every
// lane owns blocks lane, lane+32, ... -- its three 32-bit loads per block (two luma rows, one chroma
// row) and its stores are then contiguous across the warp.  All loads of a lane are issued before
// the first conversion.  Same arithmetic as the fused kernel (emit_block).
// ---------------------------------------------------------------------------------------
// blocks per lane: 2 (32 registers, 8 CTAs = 64 warps per SM) measured 530 k fps on 1080p RGB24, 4 (40
// registers, 48 warps) 486 k, 8 (48 registers, 40 warps) 469 k -- same box, profiles/r1b_direct_blocks_per_lane.md
#ifndef TSVPP_DIRECT_SUB
#define TSVPP_DIRECT_SUB 2
#endif
constexpr int kDirectSub = TSVPP_DIRECT_SUB;

// one warp unit in flight: where it goes and the words it loaded
struct DirectUnit {
    void *out;
    int pr0, gx0;  // block coordinates of this lane's first block
    uint32_t w0[kDirectSub], w1[kDirectSub], wc[kDirectSub];
};

// `f`: frame, `uu`: warp unit inside the frame
template <bool NEED_UV>
__device__ __forceinline__ void direct_load(const PlanDev &P, const FrameDev *__restrict__ frames, unsigned f,
                                            unsigned uu, int ngx, int npr, int lane, DirectUnit &d) {
    const int g0 = (int)uu * (32 * kDirectSub) + lane;
    const unsigned long long *fd = reinterpret_cast<const unsigned long long *>(frames + f);
    const unsigned char *fy = reinterpret_cast<const unsigned char *>(__ldg(fd));
    const unsigned char *fuv = reinterpret_cast<const unsigned char *>(__ldg(fd + 1));
    d.out = reinterpret_cast<void *>(__ldg(fd + 2));
    fy += (size_t)P.d_row_l0 * P.pitch_y + P.d_x_l0;
    fuv += (size_t)P.d_row_c0 * P.pitch_uv + P.d_x_c0;
    int pr = (int)(__umulhi((unsigned)g0, P.d_div_mul) >> P.d_div_shift), gx = g0 - pr * ngx;  // g0 / ngx (plan)
    d.pr0 = pr;
    d.gx0 = gx;
#pragma unroll
    for (int k = 0; k < kDirectSub; ++k) {
        d.w0[k] = d.w1[k] = d.wc[k] = 0;
        if (pr < npr) {
            const unsigned char *ly = fy + (size_t)(2 * pr) * P.pitch_y + 4 * gx;
            d.w0[k] = __ldcs(reinterpret_cast<const uint32_t *>(ly));
            d.w1[k] = __ldcs(reinterpret_cast<const uint32_t *>(ly + P.pitch_y));
            if (NEED_UV) d.wc[k] = __ldcs(reinterpret_cast<const uint32_t *>(fuv + (size_t)pr * P.pitch_uv + 4 * gx));
        }
        gx += 32;
        while (gx >= ngx) {
            gx -= ngx;
            ++pr;
        }
    }
}

template <int OK>
__device__ __forceinline__ void direct_emit(const PlanDev &P, int ngx, int npr, int vec_ok, const DirectUnit &d) {
    // RGB family: one PRMT per sample builds kMagicBits | byte (the float 2^23 + value) directly
    constexpr bool MAGIC = !(OK == OUT_Y800_U8 || OK == OUT_Y800_F32 || OK == OUT_NV12_U8 || OK == OUT_NV12_F32);
    constexpr uint32_t hi = MAGIC ? kMagicBits : 0u;
    constexpr uint32_t sel = MAGIC ? 0x7540u : 0x4440u;  // bytes {src b, 0, 0, 0x4B} / {src b, 0, 0, 0}
    int pr = d.pr0, gx = d.gx0;
#pragma unroll
    for (int k = 0; k < kDirectSub; ++k) {
        if (pr >= npr) break;
        int Yv[2][4], UV[2][2];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            Yv[0][b] = (int)__byte_perm(d.w0[k], hi, sel + b);
            Yv[1][b] = (int)__byte_perm(d.w1[k], hi, sel + b);
        }
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            UV[p][0] = (int)__byte_perm(d.wc[k], hi, sel + 2 * p);
            UV[p][1] = (int)__byte_perm(d.wc[k], hi, sel + 2 * p + 1);
        }
        if (vec_ok)
            emit_block<OK, 4, MAGIC>(P, d.out, 2 * pr, 4 * gx, Yv, UV, 4, 2, true);
        else
            emit_block<OK, 4, MAGIC>(P, d.out, 2 * pr, 4 * gx, Yv, UV, 4, 2, false);
        gx += 32;
        while (gx >= ngx) {
            gx -= ngx;
            ++pr;
        }
    }
}

// One unit per warp; blockIdx.y is the frame.  A persistent variant that prefetched the next unit's loads
// needed 64 instead of 40 registers (4 instead of 6 CTAs per SM) and was 15 % slower (478 k vs 409 k fps,
// same box): the load latency is better hidden by occupancy here.
template <int OK>
__global__ void __launch_bounds__(256) vpp_direct_kernel(const PlanDev P, const FrameDev *__restrict__ frames,
                                                         int nframes, int vec_ok) {
    constexpr bool NEED_UV = !(OK == OUT_Y800_U8 || OK == OUT_Y800_F32);
    const int lane = threadIdx.x & 31;
    const int ngx = P.dw >> 2, npr = P.dh >> 1;
    const unsigned upf = (unsigned)(ngx * npr + 32 * kDirectSub - 1) / (32 * kDirectSub);  // warp units per frame
    const unsigned uu = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (uu >= upf || (int)blockIdx.y >= nframes) return;
    DirectUnit cur;
    direct_load<NEED_UV>(P, frames, blockIdx.y, uu, ngx, npr, lane, cur);
    direct_emit<OK>(P, ngx, npr, vec_ok, cur);
}

using DirectFn = void (*)(const PlanDev, const FrameDev *, int, int);
static DirectFn pick_direct(int ok) {
    switch (ok) {
        case OUT_RGB_PLANAR_U8: return vpp_direct_kernel<OUT_RGB_PLANAR_U8>;
        case OUT_RGB_PLANAR_F32: return vpp_direct_kernel<OUT_RGB_PLANAR_F32>;
        case OUT_RGB_MERGED_U8: return vpp_direct_kernel<OUT_RGB_MERGED_U8>;
        case OUT_RGB_MERGED_F32: return vpp_direct_kernel<OUT_RGB_MERGED_F32>;
        case OUT_Y800_U8: return vpp_direct_kernel<OUT_Y800_U8>;
        case OUT_Y800_F32: return vpp_direct_kernel<OUT_Y800_F32>;
        case OUT_HSV_F32: return vpp_direct_kernel<OUT_HSV_F32>;
        case OUT_NV12_U8: return vpp_direct_kernel<OUT_NV12_U8>;
        case OUT_NV12_F32: return vpp_direct_kernel<OUT_NV12_F32>;
    }
    return nullptr;
}
int launch_direct(int ok, const PlanDev &P, const FrameDev *frames_dev, int nframes, int vec_ok, cudaStream_t stream) {
    DirectFn fn = pick_direct(ok);
    if (!fn) return (int)cudaErrorMissingConfiguration;
    const int ngx = P.dw >> 2, npr = P.dh >> 1;
    const long upf = ((long)ngx * npr + 32 * kDirectSub - 1) / (32 * kDirectSub);
    const long gx = (upf + 7) / 8;
    if (gx <= 0 || gx > 0x7fffffffL) return (int)cudaErrorInvalidConfiguration;
    for (int f0 = 0; f0 < nframes; f0 += 65535) {  // gridDim.y limit
        const int nf = std::min(65535, nframes - f0);
        fn<<<dim3((unsigned)gx, (unsigned)nf), 256, 0, stream>>>(P, frames_dev + f0, nf, vec_ok);
    }
    return (int)cudaGetLastError();
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
static KernelFn pick_fill(bool bulk, int bw) {
    // 4-column blocks only where they pay (wide outputs of the light resamplers); the bicubic and
    // area instantiations are tap-bound, not store-bound
    if constexpr (RS == RS_NEAREST || RS == RS_BILINEAR) {
        if (bw == 4) return bulk ? vpp_fused_kernel<RS, OK, true, 4> : vpp_fused_kernel<RS, OK, false, 4>;
    }
    return bulk ? vpp_fused_kernel<RS, OK, true, 2> : vpp_fused_kernel<RS, OK, false, 2>;
}
template <int RS>
static KernelFn pick_out(int ok, bool bulk, int bw) {
    switch (ok) {
        case OUT_RGB_PLANAR_U8: return pick_fill<RS, OUT_RGB_PLANAR_U8>(bulk, bw);
        case OUT_RGB_PLANAR_F32: return pick_fill<RS, OUT_RGB_PLANAR_F32>(bulk, bw);
        case OUT_RGB_MERGED_U8: return pick_fill<RS, OUT_RGB_MERGED_U8>(bulk, bw);
        case OUT_RGB_MERGED_F32: return pick_fill<RS, OUT_RGB_MERGED_F32>(bulk, bw);
        case OUT_Y800_U8: return pick_fill<RS, OUT_Y800_U8>(bulk, bw);
        case OUT_Y800_F32: return pick_fill<RS, OUT_Y800_F32>(bulk, bw);
        case OUT_HSV_F32: return pick_fill<RS, OUT_HSV_F32>(bulk, bw);
        case OUT_NV12_U8: return pick_fill<RS, OUT_NV12_U8>(bulk, bw);
        case OUT_NV12_F32: return pick_fill<RS, OUT_NV12_F32>(bulk, bw);
    }
    return nullptr;
}
static KernelFn pick_kernel(int rs, int ok, bool bulk, int bw) {
    switch (rs) {
        case RS_NEAREST: return pick_out<RS_NEAREST>(ok, bulk, bw);
        case RS_BILINEAR: return pick_out<RS_BILINEAR>(ok, bulk, bw);
        case RS_BICUBIC: return pick_out<RS_BICUBIC>(ok, bulk, bw);
        case RS_AREA: return pick_out<RS_AREA>(ok, bulk, bw);
    }
    return nullptr;
}

bool fused_supports_bw4(int rs) { return rs == RS_NEAREST || rs == RS_BILINEAR; }

int fused_block_threads(const PlanDev &P) {
    const int cons = (P.tw / P.bw) * P.rp;
    return 32 * P.nprod + (cons + 31) / 32 * 32;
}

size_t fused_smem_bytes(const PlanDev &P) {
    // stages, 2 x kMaxStages mbarriers, and (always reserved, used for unaligned frames) one shift byte per slot and stage
    const int nslots_pad = (((P.need_uv ? 3 : 2) * P.rp * P.K) + 15) & ~15;
    return (size_t)P.nstages * P.stage_bytes + 2 * 8 * kMaxStages + (size_t)P.nstages * nslots_pad + 128;
}

// the dynamic-shared-memory opt-in is a per-FUNCTION attribute shared by every plan that uses
// the instantiation: always raise it to the device maximum
int configure_fused(int rs, int ok, bool bulk, int bw, size_t smem_optin_max) {
    KernelFn fn = pick_kernel(rs, ok, bulk, bw);
    if (!fn) return (int)cudaErrorMissingConfiguration;  // reference ColorConversion.cu:376
    return (int)cudaFuncSetAttribute(reinterpret_cast<const void *>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_optin_max);
}

int max_ctas_per_sm(int rs, int ok, bool bulk, int bw, int threads, size_t smem_bytes) {
    KernelFn fn = pick_kernel(rs, ok, bulk, bw);
    if (!fn) return 0;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, reinterpret_cast<const void *>(fn), threads, smem_bytes) !=
        cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

int launch_fused(int rs, int ok, bool bulk, const PlanDev &P, const FrameDev *frames_dev, int nframes, int vec_ok,
                 int grid, size_t smem_bytes, cudaStream_t stream) {
    KernelFn fn = pick_kernel(rs, ok, bulk, P.bw);
    if (!fn) return (int)cudaErrorMissingConfiguration;
    fn<<<grid, fused_block_threads(P), smem_bytes, stream>>>(P, frames_dev, nframes, vec_ok);
    return (int)cudaGetLastError();
}

}  // namespace tsvpp
