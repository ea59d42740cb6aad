// vpp_yuv_kernels.cu -- the packed / planar YUV outputs (FourCC UYVY and YUV444).
//
// These two formats filter chroma across neighbouring chroma rows (UYVY) and across
// neighbouring pixel pairs in *linear* memory order (YUV444), so they do not fit the
// strip-local fused kernel; they run as "fused resize -> tight NV12 (u8)" followed by one of
// the kernels below (or directly on the input frame when there is neither crop nor resize).
// The reference needs NV12ToUYVY (+ UYVYToYUV444 and a 3*W*H temp) after its crop/resize
// kernels: src/ColorConversion.cu:107-209, 336-352.
//
// Every quirk the reference's goldens depend on is kept:
//   * the vertical (9,9,-1,-1)/16 filter triggers on ODD chroma rows, taps clamped to the plane;
//   * pixels are paired by LINEAR index (index % 2), not by column;
//   * the 4:4:4 horizontal filter runs over linear memory (no row-edge handling), the uint8
//     variant narrows to unsigned char before the clamp (wraps), C division truncates;
//   * entries past the 2*W*H UYVY payload read as 0 (uninitialised in the reference).
#include <cuda_runtime.h>
#include <stdint.h>

#include "vpp_launch.h"
#include "vpp_plan.h"

namespace tsvpp {

struct YuvGeom {
    int w, h;        // image size
    int pitch;       // bytes per row, used for BOTH planes (ColorConversion.cu:301)
    long y_bytes;    // readable extent of the luma plane
    long uv_bytes;   // readable extent of the chroma plane
    int norm;        // divide by 255 (float output)
};

__device__ __forceinline__ int ld_guard(const uint8_t *p, long idx, long extent) {
    return (idx >= 0 && idx < extent) ? (int)__ldg(p + idx) : 0;
}

// reference src/ColorConversion.cu:107-127 (calculateUYVYChromaVertical)
__device__ __forceinline__ int chroma_vertical(const uint8_t *UV, const YuvGeom &g, int i, int col) {
    const int r = i / 2;
    int value = ld_guard(UV, (long)r * g.pitch + col, g.uv_bytes);
    if (r & 1) {
        const int last = g.h / 2 - 1;
        const int p1 = r, p2 = min(r + 1, last), p3 = max(r - 1, 0), p4 = min(r + 2, last);
        value = (9 * (ld_guard(UV, (long)p1 * g.pitch + col, g.uv_bytes) + ld_guard(UV, (long)p2 * g.pitch + col, g.uv_bytes)) -
                 (ld_guard(UV, (long)p3 * g.pitch + col, g.uv_bytes) + ld_guard(UV, (long)p4 * g.pitch + col, g.uv_bytes)) + 8) >> 4;
        value = min(value, 255);
        value = max(value, 0);
    }
    return value;
}

// U (comp 0) / V (comp 1) entry of UYVY pair k (the pair that starts at linear pixel 2k)
__device__ __forceinline__ int pair_chroma(const uint8_t *UV, const YuvGeom &g, long k, int comp) {
    const long idx = 2 * k;
    if (k < 0 || idx >= (long)g.w * g.h) return 0;
    const int i = (int)(idx / g.w), j = (int)(idx - (long)i * g.w);
    return chroma_vertical(UV, g, i, j + comp);
}

__device__ __forceinline__ float div255f(int c) {
    const float x = (float)c, r = 0x1.010102p-8f;
    const float q = __fmul_rn(x, r);
    return __fmaf_rn(__fmaf_rn(-q, 255.0f, x), r, q);
}

// reference src/ColorConversion.cu:177-209 (NV12ToUYVY): one thread per linear pixel pair
template <typename T>
__global__ void nv12_to_uyvy_kernel(const FrameDev *__restrict__ frames, YuvGeom g) {
    const FrameDev fr = frames[blockIdx.y];
    const long npx = (long)g.w * g.h;
    const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long idx = 2 * k;
    if (idx >= npx) return;
    const int i = (int)(idx / g.w), j = (int)(idx - (long)i * g.w);
    const int U = chroma_vertical(fr.uv, g, i, j);
    const int V = chroma_vertical(fr.uv, g, i, j + 1);
    const int Y0 = ld_guard(fr.y, (long)i * g.pitch + j, g.y_bytes);
    const bool has1 = idx + 1 < npx;
    int Y1 = 0;
    if (has1) {
        const long idx1 = idx + 1;
        const int i1 = (int)(idx1 / g.w), j1 = (int)(idx1 - (long)i1 * g.w);
        Y1 = ld_guard(fr.y, (long)i1 * g.pitch + j1, g.y_bytes);
    }
    T *o = (T *)fr.out + 2 * idx;
    if (sizeof(T) == 1) {
        if (has1 && ((2 * idx) & 3) == 0) {
            *reinterpret_cast<uint32_t *>(o) = (uint32_t)U | ((uint32_t)Y0 << 8) | ((uint32_t)V << 16) | ((uint32_t)Y1 << 24);
        } else {
            o[0] = (T)U; o[1] = (T)Y0; o[2] = (T)V;
            if (has1) o[3] = (T)Y1;
        }
    } else {
        float f[4] = {(float)U, (float)Y0, (float)V, (float)Y1};
        if (g.norm) { f[0] = div255f(U); f[1] = div255f(Y0); f[2] = div255f(V); f[3] = div255f(Y1); }
        float *of = (float *)o;
        of[0] = f[0]; of[1] = f[1]; of[2] = f[2];
        if (has1) of[3] = f[3];
    }
}

// reference src/ColorConversion.cu:129-173 (calculateYUV444ChromaHorizontal / UYVYToYUV444):
// planar Y | U | V, one thread per linear pixel
template <typename T>
__global__ void nv12_to_yuv444_kernel(const FrameDev *__restrict__ frames, YuvGeom g) {
    const FrameDev fr = frames[blockIdx.y];
    const long npx = (long)g.w * g.h;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= npx) return;
    const int i = (int)(idx / g.w), j = (int)(idx - (long)i * g.w);
    const int Yv = ld_guard(fr.y, (long)i * g.pitch + j, g.y_bytes);
    int Ui, Vi;      // integer results (uint8 path / even pixels)
    float Uf, Vf;    // float results of the odd-pixel filter (float path keeps the fraction)
    bool filtered = false;
    if ((idx & 1) == 0) {
        Ui = pair_chroma(fr.uv, g, idx / 2, 0);
        Vi = pair_chroma(fr.uv, g, idx / 2, 1);
        Uf = (float)Ui;
        Vf = (float)Vi;
    } else {
        filtered = true;
        const long k = (idx - 1) / 2;
        const long si = 2 * idx + 1;  // position of this pixel's Y inside the UYVY buffer
        int t[2];
#pragma unroll
        for (int comp = 0; comp < 2; ++comp) {
            const int shift = 2 * comp;
            const long k3 = (si - 7 + shift < 0) ? k : k - 1;                // point3 < 0 -> point1
            const long k4 = (si + 5 + shift > 2 * npx - 1) ? k + 1 : k + 2;  // point4 past payload -> point2
            const int a = pair_chroma(fr.uv, g, k, comp), b = pair_chroma(fr.uv, g, k + 1, comp);
            const int c = pair_chroma(fr.uv, g, k3, comp), d = pair_chroma(fr.uv, g, k4, comp);
            t[comp] = 9 * (a + b) - (c + d) + 8;
        }
        Ui = (t[0] / 16) & 0xFF;  // C division (toward zero), narrowed to unsigned char: wraps
        Vi = (t[1] / 16) & 0xFF;
        Uf = fminf(fmaxf((float)t[0] / 16.0f, 0.0f), 255.0f);  // exact: small integers / 16
        Vf = fminf(fmaxf((float)t[1] / 16.0f, 0.0f), 255.0f);
    }
    if (sizeof(T) == 1) {
        unsigned char *o = (unsigned char *)fr.out;
        o[idx] = (unsigned char)Yv;
        o[npx + idx] = (unsigned char)Ui;
        o[2 * npx + idx] = (unsigned char)Vi;
    } else {
        float *o = (float *)fr.out;
        float y = (float)Yv, u = Uf, v = Vf;
        if (g.norm) {
            y = div255f(Yv);
            if (!filtered) { u = div255f(Ui); v = div255f(Vi); }
            else { u = __fdiv_rn(Uf, 255.0f); v = __fdiv_rn(Vf, 255.0f); }
        }
        o[idx] = y;
        o[npx + idx] = u;
        o[2 * npx + idx] = v;
    }
}


// ---------------------------------------------------------------------------------------
// Fast paths: width % 8 == 0, even height, 8-byte aligned planes / pitch, 16-byte aligned outputs.
// A thread owns 8 consecutive pixels of one row (4 linear pairs: with an even width pairs never straddle rows).
// Same arithmetic as the generic kernels above, on packed bytes.
// ---------------------------------------------------------------------------------------
// eight chroma bytes U0 V0 U1 V1 | U2 V2 U3 V3 of luma row i at byte column j, vertically filtered on odd chroma
// rows: (9 (p1 + p2) - (p3 + p4) + 8) >> 4 clamped -- reference src/ColorConversion.cu:107-127 -- evaluated on two
// 16-bit lanes per register with a bias of 512 per lane (the sum lies in [-502, 4598]): ((s + 512) >> 4) - 32.
__device__ __forceinline__ uint2 chroma8(const uint8_t *__restrict__ uv, int pitch, int h, int i, int j) {
    const int r = i >> 1;
    const uint2 p1 = __ldg(reinterpret_cast<const uint2 *>(uv + (size_t)r * pitch + j));
    if (!(r & 1)) return p1;
    const int last = (h >> 1) - 1;
    const uint2 p2 = __ldg(reinterpret_cast<const uint2 *>(uv + (size_t)min(r + 1, last) * pitch + j));
    const uint2 p3 = __ldg(reinterpret_cast<const uint2 *>(uv + (size_t)max(r - 1, 0) * pitch + j));
    const uint2 p4 = __ldg(reinterpret_cast<const uint2 *>(uv + (size_t)min(r + 2, last) * pitch + j));
    auto filt = [](uint32_t a, uint32_t b, uint32_t c, uint32_t d) -> uint32_t {
        uint32_t out = 0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {  // even bytes, then odd bytes, as two 16-bit lanes
            const int sh = 8 * half;
            const uint32_t ae = (a >> sh) & 0x00ff00ffu, be = (b >> sh) & 0x00ff00ffu;
            const uint32_t ce = (c >> sh) & 0x00ff00ffu, de = (d >> sh) & 0x00ff00ffu;
            const uint32_t s = 9u * (ae + be) + 0x02080208u - (ce + de);  // + 8 + 512 per lane: no borrow between lanes
            uint32_t w = (s >> 4) & 0x0fff0fffu;                          // floor(s / 16) + 32 per lane
            w = __vsub2(__vmaxu2(w, 0x00200020u), 0x00200020u);           // max(x, 0)
            w = __vminu2(w, 0x00ff00ffu);
            out |= w << sh;
        }
        return out;
    };
    return make_uint2(filt(p1.x, p2.x, p3.x, p4.x), filt(p1.y, p2.y, p3.y, p4.y));
}

template <typename T>
__global__ void __launch_bounds__(256) nv12_to_uyvy_fast_kernel(const FrameDev *__restrict__ frames, YuvGeom g) {
    const FrameDev fr = frames[blockIdx.y];
    const int gpr = g.w >> 3;  // 8-pixel groups per row
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (unsigned)(gpr * g.h)) return;
    const int i = (int)(t / (unsigned)gpr), j = 8 * (int)(t - (unsigned)i * (unsigned)gpr);
    const uint2 y = __ldg(reinterpret_cast<const uint2 *>(fr.y + (size_t)i * g.pitch + j));
    const uint2 c = chroma8(fr.uv, g.pitch, g.h, i, j);
    const size_t o = 2 * ((size_t)i * g.w + j);
    if (sizeof(T) == 1) {
        // U0 Y0 V0 Y1 | U1 Y2 V1 Y3 | U2 Y4 V2 Y5 | U3 Y6 V3 Y7
        const uint4 w = make_uint4(__byte_perm(c.x, y.x, 0x5140), __byte_perm(c.x, y.x, 0x7362),
                                   __byte_perm(c.y, y.y, 0x5140), __byte_perm(c.y, y.y, 0x7362));
        __stcs(reinterpret_cast<uint4 *>((uint8_t *)fr.out + o), w);
    } else {
        float *of = (float *)fr.out + o;
        const uint32_t cw[2] = {c.x, c.y}, yw[2] = {y.x, y.y};
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // pair q: U, Y, V, Y
            const int u = (int)((cw[q >> 1] >> (16 * (q & 1))) & 0xffu), v = (int)((cw[q >> 1] >> (16 * (q & 1) + 8)) & 0xffu);
            const int y0 = (int)((yw[q >> 1] >> (16 * (q & 1))) & 0xffu), y1 = (int)((yw[q >> 1] >> (16 * (q & 1) + 8)) & 0xffu);
            float4 f = make_float4((float)u, (float)y0, (float)v, (float)y1);
            if (g.norm) f = make_float4(div255f(u), div255f(y0), div255f(v), div255f(y1));
            __stcs(reinterpret_cast<float4 *>(of) + q, f);
        }
    }
}

// 4:4:4 planar.  The odd pixels' chroma is filtered over the pairs k-1, k, k+1, k+2 in LINEAR order (the
// reference filters its UYVY temp as one long array: src/ColorConversion.cu:129-143), so a block first builds the
// vertically filtered pair stream of its rows -- plus the last pair of the row before and the first two of the row
// after -- in shared memory, then filters from there.  kRows rows per block.
constexpr int kYuv444Rows = 4;
template <typename T>
__global__ void __launch_bounds__(256) nv12_to_yuv444_fast_kernel(const FrameDev *__restrict__ frames, YuvGeom g) {
    extern __shared__ __align__(16) unsigned char sm444[];  // [(kRows + 2) * w] chroma bytes, row-major, halo rows first/last
    const FrameDev fr = frames[blockIdx.y];
    const int w = g.w, h = g.h, gpr = w >> 3;
    const int i0 = blockIdx.x * kYuv444Rows;
    const int nrows = min(kYuv444Rows, h - i0);
    // phase 1: rows i0-1 .. i0+nrows (halo rows only where they exist; what the filter reads outside the image is 0)
    for (int t = threadIdx.x; t < (nrows + 2) * gpr; t += blockDim.x) {
        const int lr = t / gpr, j = 8 * (t - lr * gpr);
        const int i = i0 - 1 + lr;
        uint2 c = make_uint2(0u, 0u);
        const bool halo = lr == 0 || lr == nrows + 1;
        if (i >= 0 && i < h && (!halo || (lr == 0 ? j == w - 8 : j == 0))) c = chroma8(fr.uv, g.pitch, h, i, j);
        *reinterpret_cast<uint2 *>(sm444 + (size_t)lr * w + j) = c;
    }
    __syncthreads();
    const long npx = (long)w * h;
    for (int t = threadIdx.x; t < nrows * gpr; t += blockDim.x) {
        const int lr = t / gpr, j = 8 * (t - lr * gpr);
        const int i = i0 + lr;
        const uint2 y = __ldg(reinterpret_cast<const uint2 *>(fr.y + (size_t)i * g.pitch + j));
        const unsigned char *row = sm444 + (size_t)(lr + 1) * w;  // pair p of this row: bytes row[2p], row[2p+1]
        const long idx0 = (long)i * w + j;
        int U[8], V[8];
        float Uf[8], Vf[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int jp = j + 2 * q;  // byte column of pair k = (idx0 + 2q) / 2 inside this row
            // even pixel 2q: the pair's own (vertically filtered) chroma
            U[2 * q] = row[jp];
            V[2 * q] = row[jp + 1];
            Uf[2 * q] = (float)U[2 * q];
            Vf[2 * q] = (float)V[2 * q];
            // odd pixel 2q+1: pairs k-1, k, k+1, k+2; row[-2] / row[w] reach into the halo rows, which hold the
            // previous row's last pair / the next row's first pairs (or 0 outside the image)
            const long idx = idx0 + 2 * q + 1;
            const long si = 2 * idx + 1;
#pragma unroll
            for (int comp = 0; comp < 2; ++comp) {
                const int shift = 2 * comp;
                const int o3 = (si - 7 + shift < 0) ? 0 : -2;                // point3 < 0 -> point1
                const int o4 = (si + 5 + shift > 2 * npx - 1) ? 2 : 4;       // point4 past the payload -> point2
                const int a = row[jp + comp], b = row[jp + 2 + comp], c = row[jp + o3 + comp], d = row[jp + o4 + comp];
                const int tt = 9 * (a + b) - (c + d) + 8;
                if (comp == 0) {
                    U[2 * q + 1] = (tt / 16) & 0xFF;  // C division, narrowed to unsigned char before the clamp: wraps
                    Uf[2 * q + 1] = fminf(fmaxf((float)tt / 16.0f, 0.0f), 255.0f);
                } else {
                    V[2 * q + 1] = (tt / 16) & 0xFF;
                    Vf[2 * q + 1] = fminf(fmaxf((float)tt / 16.0f, 0.0f), 255.0f);
                }
            }
        }
        if (sizeof(T) == 1) {
            uint8_t *o = (uint8_t *)fr.out;
            auto pk = [](const int (&e)[8], int k) { return (uint32_t)e[k] | ((uint32_t)e[k + 1] << 8) | ((uint32_t)e[k + 2] << 16) | ((uint32_t)e[k + 3] << 24); };
            __stcs(reinterpret_cast<uint2 *>(o + idx0), y);
            __stcs(reinterpret_cast<uint2 *>(o + npx + idx0), make_uint2(pk(U, 0), pk(U, 4)));
            __stcs(reinterpret_cast<uint2 *>(o + 2 * npx + idx0), make_uint2(pk(V, 0), pk(V, 4)));
        } else {
            float *o = (float *)fr.out;
            float fy[8], fu[8], fv[8];
            const uint32_t yw[2] = {y.x, y.y};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int yv = (int)((yw[k >> 2] >> (8 * (k & 3))) & 0xffu);
                fy[k] = g.norm ? div255f(yv) : (float)yv;
                if (k & 1) {  // filtered: keeps the fraction, IEEE division like the reference's `/= 255`
                    fu[k] = g.norm ? __fdiv_rn(Uf[k], 255.0f) : Uf[k];
                    fv[k] = g.norm ? __fdiv_rn(Vf[k], 255.0f) : Vf[k];
                } else {
                    fu[k] = g.norm ? div255f(U[k]) : Uf[k];
                    fv[k] = g.norm ? div255f(V[k]) : Vf[k];
                }
            }
#pragma unroll
            for (int k = 0; k < 8; k += 4) {
                __stcs(reinterpret_cast<float4 *>(o + idx0 + k), make_float4(fy[k], fy[k + 1], fy[k + 2], fy[k + 3]));
                __stcs(reinterpret_cast<float4 *>(o + npx + idx0 + k), make_float4(fu[k], fu[k + 1], fu[k + 2], fu[k + 3]));
                __stcs(reinterpret_cast<float4 *>(o + 2 * npx + idx0 + k), make_float4(fv[k], fv[k + 1], fv[k + 2], fv[k + 3]));
            }
        }
    }
}

// fast != 0: the caller has checked width % 8 == 0, even height, 8-byte aligned planes and pitch, 16-byte aligned
// outputs, (kYuv444Rows + 2) * w bytes of shared memory available, and planes that cover pitch * rows
int launch_packed_yuv(int yuv444, int is_float, const FrameDev *frames_dev, int nframes, int w, int h, int pitch,
                      long y_bytes, long uv_bytes, int norm, int fast, cudaStream_t stream) {
    YuvGeom g;
    g.w = w; g.h = h; g.pitch = pitch; g.y_bytes = y_bytes; g.uv_bytes = uv_bytes; g.norm = norm;
    const long npx = (long)w * h;
    const int threads = 256;
    if (fast) {
        if (yuv444) {
            const dim3 grid((unsigned)((h + kYuv444Rows - 1) / kYuv444Rows), (unsigned)nframes);
            const size_t smem = (size_t)(kYuv444Rows + 2) * w;
            if (is_float) nv12_to_yuv444_fast_kernel<float><<<grid, threads, smem, stream>>>(frames_dev, g);
            else nv12_to_yuv444_fast_kernel<unsigned char><<<grid, threads, smem, stream>>>(frames_dev, g);
        } else {
            const dim3 grid((unsigned)(((long)(w / 8) * h + threads - 1) / threads), (unsigned)nframes);
            if (is_float) nv12_to_uyvy_fast_kernel<float><<<grid, threads, 0, stream>>>(frames_dev, g);
            else nv12_to_uyvy_fast_kernel<unsigned char><<<grid, threads, 0, stream>>>(frames_dev, g);
        }
        return (int)cudaGetLastError();
    }
    if (yuv444) {
        dim3 grid((unsigned)((npx + threads - 1) / threads), (unsigned)nframes);
        if (is_float) nv12_to_yuv444_kernel<float><<<grid, threads, 0, stream>>>(frames_dev, g);
        else nv12_to_yuv444_kernel<unsigned char><<<grid, threads, 0, stream>>>(frames_dev, g);
    } else {
        const long npairs = (npx + 1) / 2;
        dim3 grid((unsigned)((npairs + threads - 1) / threads), (unsigned)nframes);
        if (is_float) nv12_to_uyvy_kernel<float><<<grid, threads, 0, stream>>>(frames_dev, g);
        else nv12_to_uyvy_kernel<unsigned char><<<grid, threads, 0, stream>>>(frames_dev, g);
    }
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// AREA downscale with windows too large for the fused kernel's shared-memory stages (1080p -> 8x8: 240 x 135
// taps per sample).  reference src/Resize.cu:160-212 (calculateAreaInterpolation, resizeNV12DownscaleAreaKernel)
// after src/Crop.cu:4-21: one thread per output sample walks its window in the reference's order (rows outer,
// columns inner, one FMA chain -- the chain is the arithmetic, it cannot be split across threads) and writes
// the tight NV12 intermediate the reference's resize stage produces; the colour stage then runs on that
// intermediate like any other frame.  Outputs of such ratios are a few hundred samples: latency, not
// bandwidth, and irrelevant next to the source read.
__device__ __forceinline__ int area_wide_read(const uint8_t *plane, long idx, const AreaWideParams &P, bool chroma) {
    if (idx < 0) return 0;
    if (P.crop) {
        // the reference's window indexes the tight cropped copy linearly: undo that copy
        const long size = (long)P.sw * (chroma ? P.sh / 2 : P.sh);
        if (idx >= size) return 0;
        const long r = idx / P.sw, c = idx - r * P.sw;
        const long src = chroma ? ((long)(P.y0 / 2) + r) * P.pitch_uv + P.x0 + c : ((long)P.y0 + r) * P.pitch_y + P.x0 + c;
        if (src >= (chroma ? P.plane_uv : P.plane_y)) return 0;
        return plane[src];
    }
    if (idx >= (chroma ? P.plane_uv : P.plane_y)) return 0;
    return plane[idx];
}

__global__ void __launch_bounds__(128) area_wide_kernel(const FrameDev *frames, AreaWideParams P) {
    const FrameDev f = frames[blockIdx.y];
    const long nl = (long)P.dw * P.dh, nc = (long)(P.dw / 2) * (P.dh / 2) * 2;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nl + nc) return;
    const bool chroma = t >= nl;
    int i, j, comp = 0;
    if (!chroma) {
        i = (int)(t / P.dw);
        j = (int)(t - (long)i * P.dw);
    } else {
        const long u = t - nl;
        comp = (int)(u & 1);
        const long q = u >> 1;
        i = (int)(q / (P.dw / 2));
        j = (int)(q - (long)i * (P.dw / 2));
    }
    const int y = (int)floorf((float)(int)__fmul_rn(P.yr, (float)(unsigned)i));  // Resize.cu:186-190
    const int x = (int)floorf((float)(int)__fmul_rn(P.xr, (float)(unsigned)j));
    const float *px = P.patx + (long)(j % P.PX) * P.KX;
    const float *py = P.paty + (long)(i % P.PY) * P.K;
    const int ls = P.crop ? P.sw : (chroma ? P.pitch_uv : P.pitch_y);
    const int stride = chroma ? 2 : 1;
    const long start = (long)y * ls + (chroma ? 2 * x + comp : x);
    const uint8_t *plane = chroma ? f.uv : f.y;
    float sum = 0.f, div = 0.f;
    for (int a = 0; a < P.K; ++a) {
        const float wy = __ldg(py + a);
        for (int b = 0; b < P.KX; ++b) {
            const float w = __fmul_rn(__ldg(px + b), wy);
            div = __fadd_rn(div, w);
            sum = __fmaf_rn((float)area_wide_read(plane, start + (long)b * stride + (long)a * ls, P, chroma), w, sum);
        }
    }
    sum = __fdiv_rn(sum, div);
    uint8_t *o = static_cast<uint8_t *>(f.out);
    if (!chroma) o[(long)i * P.dw + j] = (uint8_t)(int)sum;
    else o[nl + (long)i * P.dw + 2 * j + comp] = (uint8_t)(int)sum;
}

int launch_area_wide(const AreaWideParams &P, const FrameDev *frames_dev, int nframes, cudaStream_t stream) {
    const long total = (long)P.dw * P.dh + (long)(P.dw / 2) * (P.dh / 2) * 2;
    const int threads = 128;
    const dim3 grid((unsigned)((total + threads - 1) / threads), (unsigned)nframes);
    area_wide_kernel<<<grid, threads, 0, stream>>>(frames_dev, P);
    return (int)cudaGetLastError();
}

}  // namespace tsvpp
