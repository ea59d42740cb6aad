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

int launch_packed_yuv(int yuv444, int is_float, const FrameDev *frames_dev, int nframes, int w, int h, int pitch,
                      long y_bytes, long uv_bytes, int norm, cudaStream_t stream) {
    YuvGeom g;
    g.w = w; g.h = h; g.pitch = pitch; g.y_bytes = y_bytes; g.uv_bytes = uv_bytes; g.norm = norm;
    const long npx = (long)w * h;
    const int threads = 256;
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

}  // namespace tsvpp
