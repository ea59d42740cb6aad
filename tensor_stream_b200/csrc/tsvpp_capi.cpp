// tsvpp_capi.cpp -- plan builder, context and the C ABI of include/tsvpp.h.
//
// Host-side logic of the hot path: the dispatch rules of the reference's
// VideoProcessor::Convert (src/VideoProcessor.cpp:106-143) and the per-pixel coordinate / weight
// formulas of its resize kernels (src/Resize.cu:180-357), evaluated ONCE per (geometry, params)
// into tables instead of per pixel per frame.  All float math here mirrors the device arithmetic
// of the reference bit for bit: fmaf() where nvcc contracted, plain IEEE ops elsewhere; this file
// is compiled with -ffp-contract=off.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tsvpp.h"
#include "vpp_launch.h"
#include "vpp_march.h"
#include "vpp_plan.h"

using namespace tsvpp;

namespace {

thread_local char g_last_error[256] = "";

inline int note_error(int code, const char *what, int line) {
    std::snprintf(g_last_error, sizeof(g_last_error), "%s -> %d (%s) at tsvpp_capi.cpp:%d", what, code,
                  code > 0 ? cudaGetErrorString((cudaError_t)code) : "tsvpp status", line);
    return code;
}

#define TSVPP_CUDA(expr)                                              \
    do {                                                              \
        cudaError_t e_ = (expr);                                      \
        if (e_ != cudaSuccess) return note_error((int)e_, #expr, __LINE__); \
    } while (0)

// Every entry point that allocates, builds a plan or launches runs on the CONTEXT's device and puts the
// caller's current device back on exit: two contexts on different GPUs may be driven from one thread, and a
// Python __del__ -> tsvpp_destroy must not change torch's current device behind its back.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) {
            err = cudaSetDevice(device);
            switched = err == cudaSuccess;
        }
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};
#define TSVPP_ON_DEVICE(ctx)                                                        \
    DeviceGuard device_guard_((ctx)->device);                                       \
    if (device_guard_.err != cudaSuccess) return note_error((int)device_guard_.err, "cudaSetDevice", __LINE__)

// no C++ exception may cross the C ABI (std::bad_alloc / length_error from a plan build would terminate a
// cgo / JNI / ctypes caller)
template <class F>
int abi_guard(F &&f, int line) {
    try {
        return f();
    } catch (const std::bad_alloc &) {
        return note_error((int)cudaErrorMemoryAllocation, "std::bad_alloc", line);
    } catch (const std::exception &e) {
        return note_error(TSVPP_ERROR, e.what(), line);
    } catch (...) {
        return note_error(TSVPP_ERROR, "unknown C++ exception", line);
    }
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline int env_int(const char *name, int dflt) {
    const char *s = std::getenv(name);
    return (s && *s) ? std::atoi(s) : dflt;
}

// ---------------------------------------------------------------------------------------
// Dispatch rules -- reference src/VideoProcessor.cpp:106-135
// ---------------------------------------------------------------------------------------
struct Geometry {
    bool crop = false, resize = false;
    int x0 = 0, y0 = 0;  // crop origin in the source
    int sw = 0, sh = 0;  // working (post-crop) size
    int dw = 0, dh = 0;  // final size
};

Geometry resolve_geometry(int W, int H, const TsvppParams &p) {
    Geometry g;
    g.sw = W;
    g.sh = H;
    const int cw = p.crop_x1 - p.crop_x0, ch = p.crop_y1 - p.crop_y0;
    if (cw > 0 && ch > 0 && cw < W && ch < H) {  // VideoProcessor.cpp:109 (both strictly smaller)
        g.crop = true;
        g.x0 = p.crop_x0;
        g.y0 = p.crop_y0;
        g.sw = cw;
        g.sh = ch;
    }
    g.dw = g.sw;
    g.dh = g.sh;
    if (p.dst_width && p.dst_height && (p.dst_width != g.sw || p.dst_height != g.sh)) {  // :118-124
        g.resize = true;
        g.dw = p.dst_width;
        g.dh = p.dst_height;
    }
    return g;
}

// What the reference never checks and then indexes with: sizes must be positive (a frame has at least one
// chroma row), the requested output size non-negative and small enough for 32-bit tables.  A crop box that
// leaves the frame stays accepted: its taps are clamped into the frame (the reference reads out of bounds).
int validate_geometry(int W, int H, const TsvppParams &p) {
    if (W <= 0 || H < 2 || W > (1 << 16) || H > (1 << 16)) return TSVPP_ERROR;
    if (p.dst_width < 0 || p.dst_height < 0 || p.dst_width > (1 << 16) || p.dst_height > (1 << 16)) return TSVPP_ERROR;
    const Geometry g = resolve_geometry(W, H, p);
    if (g.sw <= 0 || g.sh <= 0 || g.dw <= 0 || g.dh <= 0) return TSVPP_ERROR;
    return TSVPP_OK;
}

// channelsByFourCC -- reference src/VideoProcessor.cpp:4-14
// channels * w * h of the reference's (size_t)(channelsByFourCC * w * h) allocations, in exact integer
// arithmetic (a float product is only exact below 2^24 elements)
size_t elems_of(int fourcc, int w, int h) {
    const size_t px = (size_t)w * (size_t)h;
    if (fourcc == TSVPP_Y800) return px;
    if (fourcc == TSVPP_UYVY) return 2 * px;
    if (fourcc == TSVPP_NV12) return px + px / 2;  // floor(1.5 * w * h)
    return 3 * px;
}

int out_kind_of(const TsvppParams &p) {
    const bool f32 = p.normalization != 0;
    switch (p.fourcc) {
        case TSVPP_RGB24:
        case TSVPP_BGR24:
            if (p.planes == TSVPP_PLANAR) return f32 ? OUT_RGB_PLANAR_F32 : OUT_RGB_PLANAR_U8;
            return f32 ? OUT_RGB_MERGED_F32 : OUT_RGB_MERGED_U8;
        case TSVPP_Y800: return f32 ? OUT_Y800_F32 : OUT_Y800_U8;
        case TSVPP_NV12: return f32 ? OUT_NV12_F32 : OUT_NV12_U8;
        case TSVPP_HSV: return OUT_HSV_F32;  // always float, always merged (ColorConversion.cu:360-372)
        default: return -1;
    }
}

// reference src/Resize.cu:359-386 (generateResizePattern); rows are ceil(scale) floats wide,
// surplus entries dropped as the kernel never reads them (Resize.cu:162-163).
// The kernel indexes the pattern with (output index % rows): once `need` (= output columns / rows) pattern
// rows exist the period no longer matters, so generation stops there (some scales, e.g. 2034/508 in f32, only
// reach an integral id*scale after thousands of rows).
bool area_pattern(float scale, int need, std::vector<float> &out, int &rows) {
    const int stride = (int)std::ceil(scale);
    const float eps = std::numeric_limits<float>::epsilon();
    int id = 0;
    float rest = 0;
    out.clear();
    while ((float)id * scale == 0 || ((float)id * scale - (float)(int)((float)id * scale) > eps)) {
        if (id >= need) break;
        out.resize((size_t)(id + 1) * stride, 0.f);
        float *row = out.data() + (size_t)id * stride;
        float dyn = scale;
        int n = 0;
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
        id++;
    }
    rows = id;
    return true;
}

// One axis of the resampling: for output index k, the integer source coordinate, the fractional
// weight and (bicubic) the double weight, exactly as the reference kernels compute them.
struct AxisSample {
    int pos = 0;
    float w = 0.f;
    double wd = 0.0;
};

AxisSample axis_sample(int rs_api, bool area_down, bool identity, int k, float ratio, int src_len) {
    AxisSample a;
    if (identity) {
        a.pos = k;
        return a;
    }
    switch (rs_api) {
        case TSVPP_NEAREST:  // Resize.cu:250-251
            a.pos = (int)(ratio * (float)(unsigned)k);
            break;
        case TSVPP_BILINEAR: {  // Resize.cu:277-303
            const float f = fmaf((float)(unsigned)k + 0.5f, ratio, -0.5f);
            int x = (int)floorf(f);
            float w = f - (float)x;
            if (x < 0) { x = 0; w = 0; }
            if (x > src_len - 1) { x = src_len - 1; w = 0; }
            a.pos = x;
            a.w = w;
        } break;
        case TSVPP_BICUBIC: {  // Resize.cu:322-348: f32 coordinate (FMA), widened to double
            const double f = (double)fmaf((float)(unsigned)k + 0.5f, ratio, -0.5f);
            int x = (int)std::floor(f);
            double w = f - x;
            if (x < 0) { x = 0; w = 0; }
            if (x > src_len - 1) { x = src_len - 1; w = 0; }
            a.pos = x;
            a.wd = w;
        } break;
        case TSVPP_AREA:
            if (area_down) {  // Resize.cu:187-193
                a.pos = (int)floorf((float)(int)(ratio * (float)(unsigned)k));
            } else {  // Resize.cu:220-232
                const int x = (int)floorf(ratio * (float)(unsigned)k);
                float w = (float)(unsigned)(k + 1) - (float)(x + 1) / ratio;
                if (w <= 0) w = 0; else w = w - floorf(w);
                a.pos = x;
                a.w = w;
            }
            break;
    }
    return a;
}

// taps of one sample along one axis (positions in the working image; stride 1 or 2)
// bilinear -- Resize.cu:7-10; bicubic -- Resize.cu:32-43
void bilinear_taps(int X, int diff, int len, int t[2]) {
    int d = diff;
    if ((float)X + (float)d >= (float)len) d = 0;
    t[0] = X;
    t[1] = X + d;
}
void bicubic_taps(int X, int diff, int len, int t[4]) {
    int d = diff, dtop = diff;
    if (X + d >= len) d = 0;
    if (X + d * 2 >= len) d = 0;
    if (X - dtop < 0) dtop = 0;
    t[0] = X - dtop;
    t[1] = X;
    t[2] = X + d;
    t[3] = X + 2 * d;
}

// all four weights multiples of 2^-12 with |c| <= 2: products with a byte and their partial sums
// are then exact in float32 (<= 23 significant bits), so the f32 row sum equals the FP64 one
int weights_exact(const double *c) {
    for (int k = 0; k < 4; ++k) {
        const double v = c[k] * 4096.0;
        if (std::fabs(c[k]) > 2.0 || v != std::floor(v)) return 0;
    }
    return 1;
}

struct DeviceArena {
    std::vector<unsigned char> host;
    size_t add(const void *p, size_t bytes) {
        const size_t off = (host.size() + 255) / 256 * 256;
        host.resize(off + bytes);
        if (bytes) std::memcpy(host.data() + off, p, bytes);
        return off;
    }
};

struct PlanKey {
    int w, h, pitch_y, pitch_uv, bulk;
    TsvppParams p;
    bool operator<(const PlanKey &o) const { return std::memcmp(this, &o, sizeof(PlanKey)) < 0; }
    bool operator==(const PlanKey &o) const { return std::memcmp(this, &o, sizeof(PlanKey)) == 0; }
};

struct Plan {
    PlanDev dev{};
    int rs = 0, ok = 0;
    bool bulk = true;
    size_t smem = 0;
    int ctas_per_sm = 1;
    int out_w = 0, out_h = 0;
    size_t out_elems = 0;
    int elem_size = 1;
    int vec_ok_geom = 0;  // dw % 2 == 0
    long fetch_bytes_per_frame = 0;
    void *dev_tables = nullptr;
    std::string desc;
    // exact-integer separable kernel (vpp_march_kernels.cu), used instead of the float kernel when the
    // geometry's weights are dyadic (march_plan.inl)
    bool march = false;
    MarchPlanDev mdev{};
    void *march_tables = nullptr;
    size_t march_smem = 0;
    int march_ctas_per_sm = 1;
    ~Plan() {
        if (dev_tables) cudaFree(dev_tables);
        if (march_tables) cudaFree(march_tables);
    }
};

#include "march_plan.inl"

}  // namespace

struct TsvppContext {
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 227 * 1024;
    std::mutex mu;
    // plan cache: shared ownership (a launch in flight or a TsvppBatch keeps its plan alive), LRU-bounded
    std::map<PlanKey, std::shared_ptr<Plan>> plans;
    std::list<PlanKey> lru;  // most recently used first
    size_t max_plans = 64;
    bool no_march = false;  // TSVPP_NO_MARCH=1 at tsvpp_create: keep the float bicubic / area kernels (A/B, tests)
    std::atomic<unsigned long long> launches{0};
    // descriptor ring for tsvpp_convert_batch (pinned host + device), guarded by events
    static constexpr int kRing = 32;  // concurrent convert calls in flight (the reference allows maxConsumers = 5 by default)
    FrameDev *ring_host[kRing] = {};
    FrameDev *ring_dev[kRing] = {};
    cudaEvent_t ring_ev[kRing] = {};
    int ring_cap[kRing] = {};
    bool ring_busy[kRing] = {};    // claimed by a call that has not yet recorded its event
    bool ring_recorded[kRing] = {};
    int ring_next = 0;
    std::condition_variable ring_cv;
    std::mutex host_mu;  // tsvpp_convert_batch_host owns hstream / stage_in / stage_out for the whole call
    // coalescing of per-frame calls (tsvpp_convert_coalesced): requests pending at the same time share a launch
    struct CoReq {
        TsvppFrame f;
        TsvppParams p;
        void *out = nullptr;
        int rc = 0;
        unsigned seq = 0;            // completion ticket (stream memory-op path)
        cudaEvent_t ev = nullptr;    // completion event (fallback path)
        std::atomic<int> done{0};
    };
    static constexpr int kCoEvents = 64;
    static constexpr int kCoSlots = 32;
    std::mutex co_qmu;       // guards co_q
    std::vector<CoReq *> co_q;
    std::mutex co_launch;    // whoever holds it drains co_q and launches (flat combining)
    std::once_flag co_once;
    int co_init_rc = 0;
    cudaStream_t co_stream = nullptr;
    cudaEvent_t co_ev[kCoEvents] = {};
    unsigned co_evnext = 0;
    // Completion tickets: after every coalesced launch the stream writes the launch's sequence number into a pinned
    // host word (cuStreamWriteValue32).  Waiters and descriptor-slot reuse poll that word -- no CUDA call per frame,
    // which matters because driver calls from many threads serialise on the context lock (measured: 8 consumers each
    // calling cudaMallocAsync / cudaEventSynchronize stretch a 10 us launch path to 40 us).
    bool co_memops = false;
    // cuStreamWriteValue32, resolved through the runtime (cudaGetDriverEntryPoint): the library must load on machines
    // without a driver (CPU-only build / export checks), so it does not link libcuda
    CUresult (*co_write32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
    volatile unsigned *co_done_host = nullptr;
    CUdeviceptr co_done_dev = 0;
    unsigned co_seq = 0;  // last ticket issued (launch lock held)
    std::atomic<unsigned> co_seq_pub{0};  // the same, readable without the lock
    // Launcher thread (default; TSVPP_CO_THREAD=0 lets the callers combine among themselves instead): one thread
    // drains the queue and issues the launches, so the issue path stays on one core with warm caches and the callers
    // never touch the driver.  It spins while frames keep coming and sleeps on a condition variable when idle.
    bool co_use_thread = false;
    std::thread co_thread;
    std::atomic<bool> co_stop{false};
    std::atomic<int> co_pending{0};
    std::atomic<int> co_sleeping{0};
    std::mutex co_cv_mu;
    std::condition_variable co_cv;
    // posted requests (wait < 0): copied into the queue, the caller does not wait for the launch to be issued
    struct CoPosted {
        TsvppFrame f;
        TsvppParams p;
        void *out;
    };
    std::vector<CoPosted> co_qp;      // guarded by co_qmu
    std::atomic<int> co_posted_rc{0};  // first error of a posted frame, reported by tsvpp_coalesce_wait
    std::atomic<unsigned long long> co_posted_total{0}, co_posted_issued{0};  // frames posted / posted frames launched
    struct CoSlot {
        FrameDev *host = nullptr, *dev = nullptr;
        FrameDev *host_dev = nullptr;  // device address of the (mapped) pinned host copy
        int cap = 0;
        unsigned seq = 0;
        bool used = false;
    } co_slot[kCoSlots];
    unsigned co_slotnext = 0;
    std::atomic<unsigned long long> co_frames{0}, co_launches{0}, co_drain_ns{0};
    // host-buffer path
    cudaStream_t hstream[2] = {};
    unsigned char *stage_in[2] = {};
    unsigned char *stage_out[2] = {};
    size_t stage_in_cap[2] = {}, stage_out_cap[2] = {};
    int last_host_zero_copy = 0;
    cudaEvent_t hev[2] = {};
};

struct TsvppBatch {
    TsvppContext *ctx = nullptr;
    std::shared_ptr<Plan> plan;
    // UYVY / YUV444 batches replay the two-kernel path from a host copy of the descriptors
    bool packed_yuv = false;
    TsvppParams params{};
    std::vector<TsvppFrame> frames;
    std::vector<void *> outs;
    FrameDev *frames_dev = nullptr;
    int n = 0;
    int vec_ok = 0;
    int grid = 0;
};

namespace {

int build_plan(TsvppContext *ctx, int W, int H, int pitch_y_in, int pitch_uv_in, const TsvppParams &p, bool bulk,
               std::unique_ptr<Plan> &out) {
    if (int vrc = validate_geometry(W, H, p)) return vrc;
    std::unique_ptr<Plan> pl(new Plan());
    const Geometry g = resolve_geometry(W, H, p);
    const int ok = out_kind_of(p);
    if (ok < 0) {
        if (p.fourcc == TSVPP_UYVY || p.fourcc == TSVPP_YUV444) return TSVPP_UNSUPPORTED;
        return (int)cudaErrorMissingConfiguration;  // unknown FourCC, ColorConversion.cu:376
    }
    if (p.resize_type < TSVPP_NEAREST || p.resize_type > TSVPP_AREA) return TSVPP_UNSUPPORTED;

    int pitch_y = pitch_y_in ? pitch_y_in : W;    // Resize.cu:420-421
    int pitch_uv = pitch_uv_in ? pitch_uv_in : W;
    const long plane_y = (long)pitch_y * H, plane_uv = (long)pitch_uv * (H / 2);
    // colour conversion straight from the input frame indexes BOTH planes with linesize[0]
    // (ColorConversion.cu:301); kept, including when the two pitches differ.
    if (!g.crop && !g.resize) pitch_uv = pitch_y;

    const bool identity = !g.resize;
    const float xr = (float)g.sw / g.dw, yr = (float)g.sh / g.dh;  // Resize.cu:418-419
    const bool area_down = g.resize && p.resize_type == TSVPP_AREA && xr > 1 && yr > 1;  // :436
    int rs;
    if (identity || p.resize_type == TSVPP_NEAREST) rs = RS_NEAREST;
    else if (p.resize_type == TSVPP_BILINEAR) rs = RS_BILINEAR;
    else if (p.resize_type == TSVPP_BICUBIC) rs = RS_BICUBIC;
    else rs = area_down ? RS_AREA : RS_BILINEAR;

    const bool need_uv = !(ok == OUT_Y800_U8 || ok == OUT_Y800_F32);
    const int dw = g.dw, dh = g.dh;
    const int cw2 = dw / 2, ch2 = dh / 2;  // chroma samples that exist (Resize.cu:305, i<dstH/2 && j<dstW/2)

    // ---- pattern tables (AREA downscale) ----
    std::vector<float> patx, paty;
    int PX = 1, PY = 1;
    int KX = 1, K = 1;
    if (rs == RS_AREA) {
        if (!area_pattern(xr, dw, patx, PX) || !area_pattern(yr, dh, paty, PY)) return TSVPP_UNSUPPORTED;
        KX = (int)std::ceil(xr);
        K = (int)std::ceil(yr);
    } else if (rs == RS_BILINEAR) {
        K = 2;
    } else if (rs == RS_BICUBIC) {
        K = 4;
    }
    const int ntap_x = (rs == RS_NEAREST) ? 1 : (rs == RS_BILINEAR) ? 2 : (rs == RS_BICUBIC) ? 4 : 1;

    // ---- per-column samples (absolute byte positions inside a source row) ----
    std::vector<AxisSample> xs(dw), ys(std::max(dh, 1));
    for (int j = 0; j < dw; ++j) xs[j] = axis_sample(p.resize_type, area_down, identity, j, xr, g.sw);
    for (int i = 0; i < dh; ++i) ys[i] = axis_sample(p.resize_type, area_down, identity, i, yr, g.sh);

    // luma taps per column [dw][ntap_x], chroma taps per chroma byte [dw][ntap_x]; -1 = absent
    std::vector<int> tap_l((size_t)dw * 4, 0), tap_c((size_t)dw * 4, 0);
    std::vector<int> lo_l(dw), hi_l(dw), lo_c(dw, 0), hi_c(dw, 0);
    const int row_bytes_max_y = pitch_y, row_bytes_max_uv = pitch_uv;
    auto clamp_abs = [](int v, int hi) { return std::max(0, std::min(v, hi)); };
    for (int j = 0; j < dw; ++j) {
        int t[4] = {0, 0, 0, 0};
        const int X = xs[j].pos;
        if (rs == RS_NEAREST) t[0] = X;
        else if (rs == RS_BILINEAR) bilinear_taps(X, 1, g.sw, t);
        else if (rs == RS_BICUBIC) bicubic_taps(X, 1, g.sw, t);
        else t[0] = X;
        int lo = 1 << 30, hi = -1;
        for (int k = 0; k < ntap_x; ++k) {
            const int a = clamp_abs(g.x0 + t[k], row_bytes_max_y - 1);
            tap_l[(size_t)j * 4 + k] = a;
            lo = std::min(lo, a);
            hi = std::max(hi, a);
        }
        if (rs == RS_AREA) hi = std::min(tap_l[(size_t)j * 4] + (KX - 1), row_bytes_max_y - 1);
        lo_l[j] = lo;
        hi_l[j] = hi;
    }
    if (need_uv) {
        for (int b = 0; b < dw; ++b) {
            const int cj = b >> 1, comp = b & 1;
            if (cj >= cw2) {  // no chroma sample here: point at byte 0 of the segment
                lo_c[b] = -1;
                continue;
            }
            int t[4] = {0, 0, 0, 0};
            const int X = 2 * xs[cj].pos + comp;  // Resize.cu:306-307: (2*x, 2*x+1) with xDiff=2
            if (rs == RS_NEAREST) t[0] = X;
            else if (rs == RS_BILINEAR) bilinear_taps(X, 2, g.sw, t);
            else if (rs == RS_BICUBIC) bicubic_taps(X, 2, g.sw, t);
            else t[0] = X;
            int lo = 1 << 30, hi = -1;
            for (int k = 0; k < ntap_x; ++k) {
                const int a = clamp_abs(g.x0 + t[k], row_bytes_max_uv - 1);
                tap_c[(size_t)b * 4 + k] = a;
                lo = std::min(lo, a);
                hi = std::max(hi, a);
            }
            if (rs == RS_AREA) hi = std::min(tap_c[(size_t)b * 4] + 2 * (KX - 1), row_bytes_max_uv - 1);
            lo_c[b] = lo;
            hi_c[b] = hi;
        }
    }

    // ---- tiling and shared-memory budget ----
    const int want_tw = env_int("TSVPP_TW", 256);
    const int want_rp = env_int("TSVPP_RP", 2);
    const int want_ns = env_int("TSVPP_STAGES", 0);
    const size_t smem_cap = ctx->smem_optin;
    int tw = 32, rp = 1, nstages = 1, ntiles = 1, slot_stride = 16;
    std::vector<int> a0l, lenl, a0c, lenc;
    bool found = false;
    // Scored search over (row pairs per strip, tile width, ring depth): maximise the number of
    // consumer threads resident per SM (occupancy hides the stage latency -- profiles/r1_v*_sweep),
    // then prefer wider tiles (fewer, longer bulk copies) and a deeper ring.
    // columns per consumer thread: 4 (float4 / uchar4 stores, half the per-item overhead per pixel)
    // for wide outputs of the light resamplers, else 2
    // (measured, profiles/r1_v5_*: +20-30 % for nearest / colour-only at >= 512 columns, +10-15 % for
    // bilinear downscales, but -35 % for bilinear upscales, which are latency- not store-bound)
    int bw = (fused_supports_bw4(rs) && dw % 4 == 0 && dw >= 512 && (rs == RS_NEAREST || (xr >= 1.0f && yr >= 1.0f))) ? 4 : 2;
    if (env_int("TSVPP_BW", 0) == 2) bw = 2;
    if (env_int("TSVPP_BW", 0) == 4 && fused_supports_bw4(rs) && dw % 4 == 0) bw = 4;
    int nprod_bulk = 1;  // bulk-copy producer warps (2 only with 4-column blocks: __launch_bounds__)
    if (bw == 4 && env_int("TSVPP_NPROD", 1) == 2) nprod_bulk = 2;
    const int regs_assumed = bw == 4 ? 80 : ((rs == RS_NEAREST || rs == RS_BILINEAR) ? 56 : 72);  // caps implied by __launch_bounds__
    long best_score = -1;
    // pass 0: tiles of 32..256 columns.  pass 1, only when no such tile fits (AREA windows of an extreme
    // downscale: 3 * rp * ceil(yr) row slots of tw * xr bytes): narrower tiles down to one column block.
    for (int pass = 0; pass < 2 && !found; ++pass)
    for (int cand_rp = 2; cand_rp >= 1; --cand_rp) {
        if (cand_rp == 2 && dh < 4) continue;
        if (env_int("TSVPP_RP", 0) && cand_rp != want_rp) continue;
        for (int cand = pass ? 16 : 256; cand >= (pass ? bw : 32); cand >>= 1) {
            if (cand > want_tw) continue;
            if (cand > 32 && cand >= 2 * round_up(dw, 32)) continue;  // don't over-size tiles for narrow outputs
            const int nt = (dw + cand - 1) / cand;
            std::vector<int> A0l(nt), Ll(nt), A0c(nt, 0), Lc(nt, 16);
            int stride = 16;
            for (int t = 0; t < nt; ++t) {
                const int j0 = t * cand, j1 = std::min(dw, j0 + cand);
                int lo = 1 << 30, hi = -1;
                for (int j = j0; j < j1; ++j) {
                    lo = std::min(lo, lo_l[j]);
                    hi = std::max(hi, hi_l[j]);
                }
                A0l[t] = lo & ~15;
                int need = hi + 1 - A0l[t];
                Ll[t] = std::min(round_up(need, 16), round_up(pitch_y - A0l[t], 16));
                stride = std::max(stride, round_up(need, 16));
                if (need_uv) {
                    lo = 1 << 30;
                    hi = -1;
                    for (int b = j0; b < j1; ++b) {
                        if (lo_c[b] < 0) continue;
                        lo = std::min(lo, lo_c[b]);
                        hi = std::max(hi, hi_c[b]);
                    }
                    if (hi < 0) { lo = 0; hi = 0; }
                    A0c[t] = lo & ~15;
                    need = hi + 1 - A0c[t];
                    Lc[t] = std::min(round_up(need, 16), round_up(pitch_uv - A0c[t], 16));
                    stride = std::max(stride, round_up(need, 16));
                }
            }
            if (!bulk) stride += 16;  // unaligned frames: a row segment starts up to 15 bytes into its slot (fill_stage_shifted)
            const int nslots = (need_uv ? 3 : 2) * cand_rp * K;
            const size_t stage = (size_t)nslots * stride;
            const size_t fixed = 2 * 8 * kMaxStages + 128 + 4 * (size_t)((nslots + 15) & ~15);
            const int consumers = (cand / bw) * cand_rp;
            const int threads = 32 * (bulk ? nprod_bulk : kLdgProducerWarps) + round_up(consumers, 32);
            const int reg_ctas = std::max(1, 65536 / (threads * regs_assumed));
            for (int ns = (want_ns > 0 ? want_ns : 4); ns >= 1; --ns) {
                if (want_ns > 0 && ns != want_ns) continue;
                const size_t smem = fixed + stage * ns;
                if (smem > smem_cap) continue;
                // 1 KB per CTA is reserved by the driver; 228 KB per SM
                const int by_smem = (int)((size_t)(228 * 1024) / (smem + 1024));
                const int ctas = std::max(1, std::min(std::min(by_smem, reg_ctas), 2048 / threads));
                // one stage can only be refilled while another is consumed: a single stage serialises
                // resident consumer threads, discounted by the share of the CTA that is the producer warp
                // (small CTAs spend a third of their warps on issuing ever smaller copies)
                const long eff = (long)ctas * consumers * consumers / (consumers + 32);
                const long score = eff * (ns >= 2 ? 8 : 1) * 64 + (cand / 32) * 4 + std::min(ns, 3);
                if (score > best_score) {
                    best_score = score;
                    found = true;
                    tw = cand;
                    rp = cand_rp;
                    nstages = ns;
                    ntiles = nt;
                    slot_stride = stride;
                    a0l = A0l; lenl = Ll; a0c = A0c; lenc = Lc;
                }
            }
        }
    }
    if (!found) return TSVPP_UNSUPPORTED;  // taps of one strip do not fit in shared memory

    const int nstrips = (dh + 2 * rp - 1) / (2 * rp);
    const int nrows_l = nstrips * 2 * rp, nrows_c = nstrips * rp;

    // ---- row tables ----
    std::vector<int> rowsrc_l((size_t)nrows_l * K), rowsrc_c((size_t)std::max(1, nrows_c) * K, 0);
    std::vector<float> wy(nrows_l, 0.f);
    std::vector<double> wyd(nrows_l, 0.0);
    for (int i = 0; i < nrows_l; ++i) {
        const int ii = std::min(i, dh - 1);
        const AxisSample &s = ys[ii];
        wy[i] = s.w;
        wyd[i] = s.wd;
        int t[4] = {0, 0, 0, 0};
        if (rs == RS_NEAREST) t[0] = s.pos;
        else if (rs == RS_BILINEAR) bilinear_taps(s.pos, 1, g.sh, t);
        else if (rs == RS_BICUBIC) bicubic_taps(s.pos, 1, g.sh, t);
        for (int k = 0; k < K; ++k) {
            const int r = (rs == RS_AREA) ? s.pos + k : t[k];
            rowsrc_l[(size_t)i * K + k] = std::max(0, std::min(g.y0 + r, H - 1));
        }
    }
    if (need_uv) {
        for (int ci = 0; ci < nrows_c; ++ci) {
            const int cc = std::min(ci, std::max(ch2 - 1, 0));
            const AxisSample &s = ys[std::min(cc, dh - 1)];  // luma-grid formula at the chroma index
            int t[4] = {0, 0, 0, 0};
            if (rs == RS_NEAREST) t[0] = s.pos;
            else if (rs == RS_BILINEAR) bilinear_taps(s.pos, 1, g.sh / 2, t);
            else if (rs == RS_BICUBIC) bicubic_taps(s.pos, 1, g.sh / 2, t);
            for (int k = 0; k < K; ++k) {
                const int r = (rs == RS_AREA) ? s.pos + k : t[k];
                rowsrc_c[(size_t)ci * K + k] = std::max(0, std::min(g.y0 / 2 + r, H / 2 - 1));
            }
        }
    }

    // ---- bicubic weights, computed on the device with the reference's own pow() ----
    std::vector<double> cxw, cyw;  // [dw][4], [nrows_l][4]
    if (rs == RS_BICUBIC) {
        std::vector<double> wall((size_t)dw + nrows_l);
        for (int j = 0; j < dw; ++j) wall[j] = xs[j].wd;
        for (int i = 0; i < nrows_l; ++i) wall[(size_t)dw + i] = wyd[i];
        struct DevBuf {  // freed on every exit path
            double *p = nullptr;
            ~DevBuf() {
                if (p) cudaFree(p);
            }
        } dW, dC;
        const int n = (int)wall.size();
        TSVPP_CUDA(cudaMalloc(&dW.p, sizeof(double) * n));
        TSVPP_CUDA(cudaMalloc(&dC.p, sizeof(double) * 4 * n));
        TSVPP_CUDA(cudaMemcpy(dW.p, wall.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
        launch_bicubic_weights(dW.p, dC.p, n, 0);
        std::vector<double> call((size_t)4 * n);
        TSVPP_CUDA(cudaMemcpy(call.data(), dC.p, sizeof(double) * 4 * n, cudaMemcpyDeviceToHost));
        cxw.assign(call.begin(), call.begin() + (size_t)4 * dw);
        cyw.assign(call.begin() + (size_t)4 * dw, call.end());
    }

    // ---- column tables with tile-relative offsets ----
    const int ncol = ntiles * tw;
    DeviceArena ar;
    size_t off_col_l = 0, off_col_c = 0;
    auto rel = [&](int abs_pos, int a0) { return std::max(0, abs_pos - a0); };
    int ident = 0;
    if (rs == RS_NEAREST) {
        std::vector<ColNearest> cl(ncol, ColNearest{0}), cc(ncol, ColNearest{0});
        for (int j = 0; j < dw; ++j) {
            const int t = j / tw;
            cl[j].o = rel(tap_l[(size_t)j * 4], a0l[t]);
            if (need_uv && lo_c[j] >= 0) cc[j].o = rel(tap_c[(size_t)j * 4], a0c[t]);
        }
        ident = 1;
        for (int j0 = 0; j0 + bw <= dw && ident; j0 += bw) {
            for (int b = 0; b < bw; ++b) {
                if (cl[j0 + b].o != cl[j0].o + b) ident = 0;
                if (need_uv && (lo_c[j0 + b] < 0 || cc[j0 + b].o != cc[j0].o + b)) ident = 0;
            }
            if (cl[j0].o % bw || (need_uv && cc[j0].o % bw)) ident = 0;
        }
        off_col_l = ar.add(cl.data(), cl.size() * sizeof(ColNearest));
        off_col_c = ar.add(cc.data(), cc.size() * sizeof(ColNearest));
    } else if (rs == RS_BILINEAR) {
        std::vector<ColBilinear> cl(ncol, ColBilinear{0, 0, 0.f, 0}), cc(ncol, ColBilinear{0, 0, 0.f, 0});
        for (int j = 0; j < dw; ++j) {
            const int t = j / tw;
            cl[j].o0 = rel(tap_l[(size_t)j * 4 + 0], a0l[t]);
            cl[j].o1 = rel(tap_l[(size_t)j * 4 + 1], a0l[t]);
            cl[j].w = xs[j].w;
            if (need_uv && lo_c[j] >= 0) {
                cc[j].o0 = rel(tap_c[(size_t)j * 4 + 0], a0c[t]);
                cc[j].o1 = rel(tap_c[(size_t)j * 4 + 1], a0c[t]);
                cc[j].w = xs[j >> 1].w;
            }
        }
        off_col_l = ar.add(cl.data(), cl.size() * sizeof(ColBilinear));
        off_col_c = ar.add(cc.data(), cc.size() * sizeof(ColBilinear));
    } else if (rs == RS_BICUBIC) {
        ColBicubic z;
        std::memset(&z, 0, sizeof(z));
        std::vector<ColBicubic> cl(ncol, z), cc(ncol, z);
        for (int j = 0; j < dw; ++j) {
            const int t = j / tw;
            for (int k = 0; k < 4; ++k) {
                cl[j].o[k] = rel(tap_l[(size_t)j * 4 + k], a0l[t]);
                cl[j].c[k] = cxw[(size_t)4 * j + k];
                cl[j].cf[k] = (float)cl[j].c[k];
            }
            cl[j].exact = weights_exact(cl[j].c);
            if (need_uv && lo_c[j] >= 0) {
                for (int k = 0; k < 4; ++k) {
                    cc[j].o[k] = rel(tap_c[(size_t)j * 4 + k], a0c[t]);
                    cc[j].c[k] = cxw[(size_t)4 * (j >> 1) + k];
                    cc[j].cf[k] = (float)cc[j].c[k];
                }
                cc[j].exact = weights_exact(cc[j].c);
            }
        }
        off_col_l = ar.add(cl.data(), cl.size() * sizeof(ColBicubic));
        off_col_c = ar.add(cc.data(), cc.size() * sizeof(ColBicubic));
    } else {
        std::vector<ColArea> cl(ncol, ColArea{0, 0}), cc(ncol, ColArea{0, 0});
        for (int j = 0; j < dw; ++j) {
            const int t = j / tw;
            cl[j].o = rel(tap_l[(size_t)j * 4], a0l[t]);
            cl[j].prow = j % PX;  // Resize.cu:198
            if (need_uv && lo_c[j] >= 0) {
                cc[j].o = rel(tap_c[(size_t)j * 4], a0c[t]);
                cc[j].prow = (j >> 1) % PX;
            }
        }
        off_col_l = ar.add(cl.data(), cl.size() * sizeof(ColArea));
        off_col_c = ar.add(cc.data(), cc.size() * sizeof(ColArea));
    }
    const size_t off_a0l = ar.add(a0l.data(), a0l.size() * sizeof(int));
    const size_t off_lenl = ar.add(lenl.data(), lenl.size() * sizeof(int));
    const size_t off_a0c = ar.add(a0c.data(), a0c.size() * sizeof(int));
    const size_t off_lenc = ar.add(lenc.data(), lenc.size() * sizeof(int));
    const size_t off_rsl = ar.add(rowsrc_l.data(), rowsrc_l.size() * sizeof(int));
    const size_t off_rsc = ar.add(rowsrc_c.data(), rowsrc_c.size() * sizeof(int));
    const size_t off_wy = ar.add(wy.data(), wy.size() * sizeof(float));
    const size_t off_cy = ar.add(cyw.data(), cyw.size() * sizeof(double));
    std::vector<float> cyf(cyw.size());
    std::vector<int> cy_exact(cyw.size() / 4 + 1, 0);
    for (size_t i = 0; i < cyw.size(); ++i) cyf[i] = (float)cyw[i];
    for (size_t r = 0; r + 3 < cyw.size(); r += 4) cy_exact[r / 4] = weights_exact(&cyw[r]);
    // bicubic: luma rows 2p and 2p+1 usually tap overlapping source rows; when row 2p+1's four
    // source rows are row 2p's shifted by d (checked on the resolved row table, so the clamped
    // borders simply report "no reuse"), its horizontal sums for taps k <= 3-d are row 2p's k+d
    std::vector<int> row_shift((size_t)std::max(1, nrows_c), 4);
    if (rs == RS_BICUBIC) {
        for (int pr = 0; pr < nrows_c; ++pr) {
            const int *r0 = &rowsrc_l[(size_t)(2 * pr) * K], *r1 = &rowsrc_l[(size_t)(2 * pr + 1) * K];
            for (int d = 0; d <= 3; ++d) {
                bool ok = true;
                for (int k = 0; k + d <= 3; ++k) ok = ok && (r1[k] == r0[k + d]);
                if (ok) {
                    row_shift[pr] = d;
                    break;
                }
            }
        }
    }
    const size_t off_rshift = ar.add(row_shift.data(), row_shift.size() * sizeof(int));
    const size_t off_cyf = ar.add(cyf.data(), cyf.size() * sizeof(float));
    const size_t off_cye = ar.add(cy_exact.data(), cy_exact.size() * sizeof(int));
    const size_t off_px = ar.add(patx.data(), patx.size() * sizeof(float));
    const size_t off_py = ar.add(paty.data(), paty.size() * sizeof(float));
    // AREA: the divisor sum(px[b]*py[a]) does not depend on pixel data; evaluate it once per
    // (vertical pattern row, horizontal pattern row) in the reference's order and rounding
    // (Resize.cu:166-173: rows outer, columns inner, w = px*py, div += w), not 40 FADDs per sample
    std::vector<float> area_div;
    if (rs == RS_AREA) {
        area_div.resize((size_t)PX * PY);
        for (int r = 0; r < PY; ++r)
            for (int c = 0; c < PX; ++c) {
                float div = 0.f;
                for (int a = 0; a < K; ++a)
                    for (int b = 0; b < KX; ++b) {
                        const float w = patx[(size_t)c * KX + b] * paty[(size_t)r * K + a];
                        div = div + w;
                    }
                area_div[(size_t)r * PX + c] = div;
            }
    }
    const size_t off_adiv = ar.add(area_div.data(), area_div.size() * sizeof(float));

    TSVPP_CUDA(cudaMalloc(&pl->dev_tables, ar.host.size() + 256));
    TSVPP_CUDA(cudaMemcpy(pl->dev_tables, ar.host.data(), ar.host.size(), cudaMemcpyHostToDevice));
    const unsigned char *base = (const unsigned char *)pl->dev_tables;

    PlanDev &D = pl->dev;
    D.dw = dw; D.dh = dh; D.cw2 = cw2; D.ch2 = ch2;
    D.tw = tw; D.ntiles = ntiles; D.nstrips = nstrips; D.rp = rp; D.bw = bw;
    // unaligned frames: four producer warps share the slots (two passes of address arithmetic per slot and the guarded
    // edge rows make the shifted fill dearer per slot: 1.71 M vs 1.18 M frames/s with one warp on a tight 1080x608 clip)
    D.nprod = bulk ? nprod_bulk : kLdgProducerWarps;
    D.K = K; D.KX = KX;
    D.slot_stride = slot_stride;
    D.stage_bytes = (need_uv ? 3 : 2) * rp * K * slot_stride;
    D.nstages = nstages;
    D.pitch_y = pitch_y; D.pitch_uv = pitch_uv;
    D.plane_bytes_y = plane_y; D.plane_bytes_uv = plane_uv;
    D.swap_rb = (p.fourcc == TSVPP_BGR24) ? 1 : 0;  // ColorConversion.cu:305
    D.need_uv = need_uv ? 1 : 0;
    D.ident = ident;
    // direct path: no resize (rows and columns affine with unit step, nothing clamped), 4-column
    // blocks, 32-bit aligned loads
    {
        bool direct = rs == RS_NEAREST && bulk && dw % 4 == 0 && dh % 2 == 0 && dw >= 8 && dh >= 2 &&
                      !env_int("TSVPP_NO_DIRECT", 0);
        for (int j = 0; j < dw && direct; ++j) {
            if (tap_l[(size_t)j * 4] != tap_l[0] + j) direct = false;
            if (need_uv && (lo_c[j] < 0 || tap_c[(size_t)j * 4] != tap_c[0] + j)) direct = false;
        }
        for (int i = 0; i < dh && direct; ++i)
            if (rowsrc_l[(size_t)i * K] != rowsrc_l[0] + i) direct = false;
        if (need_uv)
            for (int ci = 0; ci < dh / 2 && direct; ++ci)
                if (rowsrc_c[(size_t)ci * K] != rowsrc_c[0] + ci) direct = false;
        if (direct && ((tap_l[0] & 3) || (need_uv && (tap_c[0] & 3)))) direct = false;
        // block index -> (row pair, column group) without an integer division in the kernel: with
        // L = ceil(log2(d)), M = floor(2^(31+L) / d) + 1, n / d == (n * M) >> (31 + L) for all n < 2^31
        D.d_div_mul = 0;
        D.d_div_shift = 0;
        if (direct) {
            const unsigned d = (unsigned)(dw / 4);
            int L = 0;
            while ((1u << L) < d) ++L;
            const unsigned long long M = ((1ull << (31 + L)) / d) + 1;
            D.d_div_mul = (unsigned)M;
            D.d_div_shift = L - 1;
            const unsigned long long last = (unsigned long long)d * (unsigned)(dh / 2) + 4096;  // a unit may overrun the frame
            for (unsigned long long n = d - 1; n <= last && direct; n += d) {  // every n = k*d - 1 and k*d
                if (((n * M) >> (31 + L)) != n / d || (((n + 1) * M) >> (31 + L)) != (n + 1) / d) direct = false;
            }
            if (M >> 32) direct = false;
        }
        D.direct = direct ? 1 : 0;
        D.d_row_l0 = direct ? rowsrc_l[0] : 0;
        D.d_row_c0 = (direct && need_uv) ? rowsrc_c[0] : 0;
        D.d_x_l0 = direct ? tap_l[0] : 0;
        D.d_x_c0 = (direct && need_uv) ? tap_c[0] : 0;
    }
    D.PX = PX; D.PY = PY;
    D.seg_a0_l = (const int *)(base + off_a0l);
    D.seg_len_l = (const int *)(base + off_lenl);
    D.seg_a0_c = (const int *)(base + off_a0c);
    D.seg_len_c = (const int *)(base + off_lenc);
    D.rowsrc_l = (const int *)(base + off_rsl);
    D.rowsrc_c = (const int *)(base + off_rsc);
    D.col_l = base + off_col_l;
    D.col_c = base + off_col_c;
    D.wy = (const float *)(base + off_wy);
    D.cy = (const double *)(base + off_cy);
    D.cyf = (const float *)(base + off_cyf);
    D.cy_exact = (const int *)(base + off_cye);
    D.row_shift = (const int *)(base + off_rshift);
    D.px = (const float *)(base + off_px);
    D.py = (const float *)(base + off_py);
    D.area_div = (const float *)(base + off_adiv);

    pl->rs = rs;
    pl->ok = ok;
    pl->bulk = bulk;
    pl->smem = fused_smem_bytes(D);
    if (pl->smem > smem_cap) return TSVPP_UNSUPPORTED;
    int rc = configure_fused(rs, ok, bulk, bw, smem_cap);
    if (rc) return rc;
    pl->ctas_per_sm = std::max(1, max_ctas_per_sm(rs, ok, bulk, bw, fused_block_threads(D), pl->smem));
    const int cap_ctas = env_int("TSVPP_CTAS_PER_SM", 0);
    if (cap_ctas > 0) pl->ctas_per_sm = std::min(pl->ctas_per_sm, cap_ctas);
    pl->out_w = dw;
    pl->out_h = dh;
    pl->out_elems = elems_of(p.fourcc, dw, dh);
    if (pl->out_elems >= 4294967296ull) return TSVPP_UNSUPPORTED;  // 32-bit element offsets
    pl->elem_size = (p.normalization || p.fourcc == TSVPP_HSV) ? 4 : 1;
    pl->vec_ok_geom = (dw % bw == 0) ? 1 : 0;
    // bytes one frame's work items fetch from the source planes (what crosses HBM / PCIe)
    long fetch = 0;
    for (int t = 0; t < ntiles; ++t)
        fetch += (long)nstrips * ((long)2 * rp * K * lenl[t] + (need_uv ? (long)rp * K * lenc[t] : 0));
    pl->fetch_bytes_per_frame = fetch;
    char buf[640];
    static const char *rs_names[] = {"nearest", "bilinear", "bicubic", "area"};
    std::snprintf(buf, sizeof(buf),
                  "resample=%s out_kind=%d fill=%s bw=%d tw=%d rp=%d K=%d KX=%d ntiles=%d nstrips=%d slot_stride=%d "
                  "stage_bytes=%d nstages=%d smem=%zu threads=%d ctas_per_sm=%d crop=%d resize=%d fetch_bytes_per_frame=%ld direct=%d",
                  rs_names[rs], ok, bulk ? "bulk" : "bulk-shifted", bw, tw, rp, K, KX, ntiles, nstrips, slot_stride, D.stage_bytes,
                  nstages, pl->smem, fused_block_threads(D), pl->ctas_per_sm, (int)g.crop, (int)g.resize, fetch, D.direct);
    pl->desc = buf;
    // bilinear goes through the integer kernel where its 4-column blocks and once-per-row staging pay: wide outputs
    // that are not strong downscales (upscales, mild downscales); TSVPP_MARCH_BILINEAR=1/0 forces / forbids it
    const int mb = env_int("TSVPP_MARCH_BILINEAR", -1);
    const bool march_bilinear = rs == RS_BILINEAR && (mb == 1 || (mb != 0 && dw >= 512 && dw % 4 == 0 && xr <= 2.0f && yr <= 2.0f));
    if (bulk && (rs == RS_BICUBIC || rs == RS_AREA || march_bilinear) && !ctx->no_march) {
        MarchInputs mi;
        mi.g = g; mi.rs = rs; mi.ok = ok; mi.need_uv = need_uv;
        mi.W = W; mi.H = H; mi.pitch_y = pitch_y; mi.pitch_uv = pitch_uv;
        mi.dw = dw; mi.dh = dh; mi.cw2 = cw2; mi.ch2 = ch2; mi.swap_rb = D.swap_rb;
        mi.xs = &xs; mi.ys = &ys; mi.tap_l = &tap_l; mi.tap_c = &tap_c; mi.lo_c = &lo_c;
        mi.lo_l = &lo_l; mi.hi_l = &hi_l; mi.hi_c = &hi_c;
        mi.cxw = &cxw; mi.cyw = &cyw;
        mi.K = K; mi.KX = KX; mi.PX = PX; mi.PY = PY; mi.patx = &patx; mi.paty = &paty;
        mi.smem_cap = smem_cap;
        MarchPlanHost mh;
        const int mrc = build_march(mi, ctx->sm_count, mh);
        if (mrc == TSVPP_OK) {
            pl->march = true;
            pl->mdev = mh.dev;
            pl->march_tables = mh.dev_tables;
            pl->march_smem = mh.smem;
            pl->march_ctas_per_sm = cap_ctas > 0 ? std::min(mh.ctas_per_sm, cap_ctas) : mh.ctas_per_sm;
            pl->fetch_bytes_per_frame = mh.fetch_bytes_per_frame;
            pl->vec_ok_geom = (dw % march_bw(mh.dev.mode) == 0) ? 1 : 0;
            pl->desc = mh.desc;
        } else {
            if (mh.dev_tables) cudaFree(mh.dev_tables);
            if (mrc > 0) return mrc;  // a CUDA error, not "not eligible"
        }
    }
    out = std::move(pl);
    return TSVPP_OK;
}

bool frames_bulk_ok(const TsvppFrame *f, int n) {
    for (int i = 0; i < n; ++i) {
        const int py = f[i].pitch_y ? f[i].pitch_y : f[i].width;
        const int puv = f[i].pitch_uv ? f[i].pitch_uv : f[i].width;
        if (((uintptr_t)f[i].y & 15) || ((uintptr_t)f[i].uv & 15) || (py & 15) || (puv & 15)) return false;
    }
    return true;
}

int get_plan(TsvppContext *ctx, const TsvppFrame &f0, const TsvppParams &p, bool bulk, std::shared_ptr<Plan> *out) {
    PlanKey key;
    std::memset(&key, 0, sizeof(key));
    key.w = f0.width;
    key.h = f0.height;
    key.pitch_y = f0.pitch_y ? f0.pitch_y : f0.width;
    key.pitch_uv = f0.pitch_uv ? f0.pitch_uv : f0.width;
    key.bulk = bulk ? 1 : 0;
    key.p = p;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        auto it = ctx->plans.find(key);
        if (it != ctx->plans.end()) {
            // move to the front of the LRU list unless it is there already (the steady state of a pipeline)
            if (!(ctx->lru.front() == key)) {
                ctx->lru.remove(key);
                ctx->lru.push_front(key);
            }
            *out = it->second;
            return TSVPP_OK;
        }
    }
    // Build OUTSIDE the context lock (table upload, for bicubic a small kernel): other consumers keep
    // converting meanwhile.  Two threads may build the same plan once; the first insert wins.
    std::unique_ptr<Plan> pl;
    int rc = build_plan(ctx, key.w, key.h, key.pitch_y, key.pitch_uv, p, bulk, pl);
    if (rc) return rc;
    std::shared_ptr<Plan> sp(pl.release());
    std::vector<std::shared_ptr<Plan>> evicted;  // destroyed after the lock is dropped
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        auto ins = ctx->plans.emplace(key, sp);
        if (ins.second) ctx->lru.push_front(key);
        *out = ins.first->second;
        while (ctx->plans.size() > ctx->max_plans && !ctx->lru.empty()) {
            const PlanKey victim = ctx->lru.back();
            ctx->lru.pop_back();
            auto v = ctx->plans.find(victim);
            if (v != ctx->plans.end()) {
                // ~Plan frees the device tables with cudaFree, which waits for every launch that may still
                // read them; holders of the shared_ptr (batches, calls in flight) delay it further
                evicted.push_back(std::move(v->second));
                ctx->plans.erase(v);
            }
        }
    }
    return TSVPP_OK;
}

int check_uniform(const TsvppFrame *f, int n) {
    for (int i = 1; i < n; ++i)
        if (f[i].width != f[0].width || f[i].height != f[0].height || f[i].pitch_y != f[0].pitch_y ||
            f[i].pitch_uv != f[0].pitch_uv)
            return TSVPP_UNSUPPORTED;
    return TSVPP_OK;
}

int vec_ok_for(const Plan *pl, void *const *out, int n) {
    if (!pl->vec_ok_geom) return 0;
    for (int i = 0; i < n; ++i)
        if ((uintptr_t)out[i] & 15) return 0;
    return 1;
}

// device-resident frames of a no-resize plan skip the staging ring (vpp_direct_kernel)
int launch_plan(const Plan *pl, const FrameDev *frames_dev, int n, int vec_ok, int grid, cudaStream_t stream,
                bool allow_direct) {
    if (pl->march) return launch_march(pl->mdev, pl->ok, frames_dev, n, vec_ok, grid, pl->march_smem, stream);
    if (pl->dev.direct && allow_direct) return launch_direct(pl->ok, pl->dev, frames_dev, n, vec_ok, stream);
    return launch_fused(pl->rs, pl->ok, pl->bulk, pl->dev, frames_dev, n, vec_ok, grid, pl->smem, stream);
}

int grid_for(const TsvppContext *ctx, const Plan *pl, int n) {
    const int ntiles = pl->march ? pl->mdev.ntiles : pl->dev.ntiles;
    const long items = (long)n * (pl->march ? pl->mdev.nstrips : pl->dev.nstrips) * ntiles;
    long cap = (long)ctx->sm_count * (pl->march ? pl->march_ctas_per_sm : pl->ctas_per_sm);
    // a CTA's items advance by gridDim.x and the tile is the fastest-varying item coordinate: with
    // the grid a multiple of ntiles every CTA keeps ONE column tile for its whole life, so its
    // column tap tables are loaded once instead of once per item
    if (cap > ntiles) cap -= cap % ntiles;
    return (int)std::max(1L, std::min(items, cap));
}

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

const char *tsvpp_version(void) { return "tsvpp-b200 0.2 (sm_100a)"; }

const char *tsvpp_last_error(void) { return g_last_error; }

int tsvpp_create(int device, TsvppContext **out) {
    if (!out) return TSVPP_ERROR;
    return abi_guard([&]() -> int {
        int ndev = 0;
        TSVPP_CUDA(cudaGetDeviceCount(&ndev));
        if (device < 0 || device >= ndev) return (int)cudaErrorInvalidDevice;
        DeviceGuard g(device);  // the caller's current device is restored on return
        if (g.err != cudaSuccess) return note_error((int)g.err, "cudaSetDevice", __LINE__);
        std::unique_ptr<TsvppContext> c(new TsvppContext());
        c->device = device;
        int v = 0;
        TSVPP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
        c->sm_count = v;
        TSVPP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        c->smem_optin = (size_t)v;
        c->max_plans = (size_t)std::max(1, env_int("TSVPP_MAX_PLANS", 64));
        c->no_march = env_int("TSVPP_NO_MARCH", 0) != 0;
        *out = c.release();
        return TSVPP_OK;
    }, __LINE__);
}

void tsvpp_destroy(TsvppContext *ctx) {
    if (!ctx) return;
    if (ctx->co_thread.joinable()) {
        ctx->co_stop.store(true, std::memory_order_release);
        {
            std::lock_guard<std::mutex> lk(ctx->co_cv_mu);
            ctx->co_cv.notify_all();
        }
        ctx->co_thread.join();
    }
    {
        DeviceGuard g(ctx->device);
        cudaDeviceSynchronize();
        for (int i = 0; i < TsvppContext::kRing; ++i) {
            if (ctx->ring_host[i]) cudaFreeHost(ctx->ring_host[i]);
            if (ctx->ring_dev[i]) cudaFree(ctx->ring_dev[i]);
            if (ctx->ring_ev[i]) cudaEventDestroy(ctx->ring_ev[i]);
        }
        for (int i = 0; i < TsvppContext::kCoEvents; ++i)
            if (ctx->co_ev[i]) cudaEventDestroy(ctx->co_ev[i]);
        if (ctx->co_stream) cudaStreamDestroy(ctx->co_stream);
        if (ctx->co_done_host) cudaFreeHost((void *)ctx->co_done_host);
        for (auto &sl : ctx->co_slot) {
            if (sl.host) cudaFreeHost(sl.host);
            if (sl.dev) cudaFree(sl.dev);
        }
        for (int i = 0; i < 2; ++i) {
            if (ctx->stage_in[i]) cudaFree(ctx->stage_in[i]);
            if (ctx->stage_out[i]) cudaFree(ctx->stage_out[i]);
            if (ctx->hstream[i]) cudaStreamDestroy(ctx->hstream[i]);
            if (ctx->hev[i]) cudaEventDestroy(ctx->hev[i]);
        }
        ctx->plans.clear();
        ctx->lru.clear();
    }
    delete ctx;
}

int tsvpp_output_info(int src_width, int src_height, const TsvppParams *p, int *out_w, int *out_h, size_t *elems,
                      int *elem_size) {
    if (!p) return TSVPP_ERROR;
    if (int rc = validate_geometry(src_width, src_height, *p)) return note_error(rc, "invalid frame size / parameters", __LINE__);
    const Geometry g = resolve_geometry(src_width, src_height, *p);
    if (out_w) *out_w = g.dw;
    if (out_h) *out_h = g.dh;
    if (elems) *elems = elems_of(p->fourcc, g.dw, g.dh);
    if (elem_size) *elem_size = (p->normalization || p->fourcc == TSVPP_HSV) ? 4 : 1;
    return TSVPP_OK;
}

// Descriptor upload through a small ring of pinned buffers: stream-ordered, no host sync unless every ring
// slot is still in flight.  A slot belongs to exactly one call from claim to release: it is claimed under the
// lock (busy), its previous user's event is awaited OUTSIDE the lock, and it is released (with its new event
// recorded) by release_slot() -- so a later call can neither overwrite the pinned descriptors while the H2D
// copy is in flight nor wait on an event that was never recorded.
static int claim_slot(TsvppContext *ctx, int n, int *slot_out) {
    int slot = -1;
    {
        std::unique_lock<std::mutex> lk(ctx->mu);
        for (;;) {
            for (int k = 0; k < TsvppContext::kRing; ++k) {
                const int c = (ctx->ring_next + k) % TsvppContext::kRing;
                if (!ctx->ring_busy[c]) {
                    slot = c;
                    break;
                }
            }
            if (slot >= 0) break;
            ctx->ring_cv.wait(lk);  // kRing calls are between claim and release: wait for one of them
        }
        ctx->ring_busy[slot] = true;
        ctx->ring_next = (slot + 1) % TsvppContext::kRing;
    }
    auto fail = [&](int rc) {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->ring_busy[slot] = false;
        ctx->ring_cv.notify_one();
        return rc;
    };
    if (!ctx->ring_ev[slot]) {
        cudaError_t e = cudaEventCreateWithFlags(&ctx->ring_ev[slot], cudaEventDisableTiming);
        if (e != cudaSuccess) return fail(note_error((int)e, "cudaEventCreateWithFlags", __LINE__));
    } else if (ctx->ring_recorded[slot]) {
        cudaError_t e = cudaEventSynchronize(ctx->ring_ev[slot]);  // previous user's launch has consumed the descriptors
        if (e != cudaSuccess) return fail(note_error((int)e, "cudaEventSynchronize", __LINE__));
    }
    if (ctx->ring_cap[slot] < n) {
        if (ctx->ring_host[slot]) cudaFreeHost(ctx->ring_host[slot]);
        if (ctx->ring_dev[slot]) cudaFree(ctx->ring_dev[slot]);
        ctx->ring_host[slot] = nullptr;
        ctx->ring_dev[slot] = nullptr;
        ctx->ring_cap[slot] = 0;
        const int cap = std::max(n, 64);
        cudaError_t e = cudaMallocHost(&ctx->ring_host[slot], sizeof(FrameDev) * cap);
        if (e == cudaSuccess) e = cudaMalloc(&ctx->ring_dev[slot], sizeof(FrameDev) * cap);
        if (e != cudaSuccess) return fail(note_error((int)e, "descriptor ring allocation", __LINE__));
        ctx->ring_cap[slot] = cap;
    }
    *slot_out = slot;
    return TSVPP_OK;
}

// records the slot's event on `stream` (after the launch that reads the descriptors) and frees the slot
static void release_slot(TsvppContext *ctx, int slot, cudaStream_t stream) {
    const bool rec = cudaEventRecord(ctx->ring_ev[slot], stream) == cudaSuccess;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->ring_recorded[slot] = rec;
    ctx->ring_busy[slot] = false;
    ctx->ring_cv.notify_one();
}

static int upload_descriptors(TsvppContext *ctx, const FrameDev *src, int n, cudaStream_t stream, int *slot_out) {
    int slot = -1;
    int rc = claim_slot(ctx, n, &slot);
    if (rc) return rc;
    std::memcpy(ctx->ring_host[slot], src, sizeof(FrameDev) * n);
    cudaError_t e = cudaMemcpyAsync(ctx->ring_dev[slot], ctx->ring_host[slot], sizeof(FrameDev) * n,
                                    cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) {
        release_slot(ctx, slot, stream);
        return note_error((int)e, "cudaMemcpyAsync(descriptors)", __LINE__);
    }
    *slot_out = slot;
    return TSVPP_OK;
}

// UYVY / YUV444: (fused crop+resize -> tight NV12 u8, only when needed) + one packing kernel
// (reference src/ColorConversion.cu:336-352)
static int convert_batch_impl(TsvppContext *ctx, const TsvppFrame *frames, int n, const TsvppParams *p,
                              void *const *out, cudaStream_t stream, bool frames_on_host);

static int convert_packed_yuv(TsvppContext *ctx, const TsvppFrame *frames, int n, const TsvppParams *p,
                              void *const *out, cudaStream_t stream, bool frames_on_host) {
    const int W = frames[0].width, H = frames[0].height;
    if (int vrc = validate_geometry(W, H, *p)) return note_error(vrc, "invalid frame size / parameters", __LINE__);
    const Geometry g = resolve_geometry(W, H, *p);
    const int is_float = p->normalization ? 1 : 0;
    std::vector<FrameDev> h(n);
    unsigned char *tmp = nullptr;
    int w = g.dw, hgt = g.dh, pitch;
    long y_bytes, uv_bytes;
    if (g.crop || g.resize) {
        TsvppParams p1 = *p;
        p1.fourcc = TSVPP_NV12;
        p1.normalization = 0;
        const size_t ybytes = (size_t)w * hgt;
        const size_t fbytes = (ybytes + (size_t)w * (hgt / 2) + 255) / 256 * 256 + 256;
        TSVPP_CUDA(cudaMallocAsync(&tmp, fbytes * n, stream));
        // chroma rows that do not exist for odd sizes must read as 0, like the oracle's
        TSVPP_CUDA(cudaMemsetAsync(tmp, 0, fbytes * n, stream));
        std::vector<void *> o1(n);
        for (int i = 0; i < n; ++i) o1[i] = tmp + fbytes * i;
        int rc = convert_batch_impl(ctx, frames, n, &p1, o1.data(), stream, frames_on_host);
        if (rc) {
            cudaFreeAsync(tmp, stream);
            return rc;
        }
        for (int i = 0; i < n; ++i) {
            h[i].y = tmp + fbytes * i;
            h[i].uv = tmp + fbytes * i + ybytes;
            h[i].out = out[i];
        }
        pitch = w;
        y_bytes = (long)ybytes;
        uv_bytes = (long)(fbytes - ybytes);
    } else {
        // straight from the input frame; BOTH planes indexed with linesize[0] (ColorConversion.cu:301)
        pitch = frames[0].pitch_y ? frames[0].pitch_y : W;
        const int puv = frames[0].pitch_uv ? frames[0].pitch_uv : W;
        y_bytes = (long)pitch * H;
        uv_bytes = (long)puv * (H / 2);
        for (int i = 0; i < n; ++i) {
            h[i].y = frames[i].y;
            h[i].uv = frames[i].uv;
            h[i].out = out[i];
        }
    }
    int slot = 0;
    int rc = upload_descriptors(ctx, h.data(), n, stream, &slot);
    if (rc == TSVPP_OK) {
        (void)cudaGetLastError();
        // the vectorised kernels need regular geometry and alignment; everything else takes the generic ones
        bool fast = !env_int("TSVPP_YUV_GENERIC", 0) && w % 8 == 0 && w >= 16 && hgt % 2 == 0 && hgt >= 2 && pitch % 8 == 0 &&
                    y_bytes >= (long)pitch * (hgt - 1) + w && uv_bytes >= (long)pitch * (hgt / 2 - 1) + w &&
                    (size_t)(4 + 2) * (size_t)w <= 48 * 1024;
        for (int i = 0; i < n && fast; ++i)
            if (((uintptr_t)h[i].y & 7) || ((uintptr_t)h[i].uv & 7) || ((uintptr_t)h[i].out & 15)) fast = false;
        rc = launch_packed_yuv(p->fourcc == TSVPP_YUV444, is_float, ctx->ring_dev[slot], n, w, hgt, pitch, y_bytes,
                               uv_bytes, is_float, fast ? 1 : 0, stream);
        release_slot(ctx, slot, stream);
        ctx->launches.fetch_add(1);
        if (rc) note_error(rc, "launch_packed_yuv", __LINE__);
    }
    if (tmp) cudaFreeAsync(tmp, stream);
    return rc;
}

// AREA downscales whose tap window exceeds the fused kernel's shared-memory stages (ratios beyond ~50x: 1080p ->
// 8x8): the reference's own two stages -- resize into a tight NV12 intermediate (area_wide_kernel), then the
// colour stage on that intermediate through the ordinary plans.
static bool area_wide_applies(const TsvppFrame &f0, const TsvppParams &p) {
    if (validate_geometry(f0.width, f0.height, p)) return false;
    const Geometry g = resolve_geometry(f0.width, f0.height, p);
    if (!g.resize || p.resize_type != TSVPP_AREA) return false;
    const float xr = (float)g.sw / g.dw, yr = (float)g.sh / g.dh;
    return xr > 1 && yr > 1 && g.dh >= 2;  // the colour stage needs a chroma row under its source (validate_geometry)
}

static int convert_area_wide(TsvppContext *ctx, const TsvppFrame *frames, int n, const TsvppParams *p,
                             void *const *out, cudaStream_t stream, bool frames_on_host) {
    const int W = frames[0].width, H = frames[0].height;
    const Geometry g = resolve_geometry(W, H, *p);
    AreaWideParams A{};
    A.dw = g.dw; A.dh = g.dh; A.sw = g.sw; A.sh = g.sh;
    A.crop = g.crop ? 1 : 0; A.x0 = g.x0; A.y0 = g.y0;
    A.pitch_y = frames[0].pitch_y ? frames[0].pitch_y : W;    // Resize.cu:420-421
    A.pitch_uv = frames[0].pitch_uv ? frames[0].pitch_uv : W;
    A.plane_y = (long)A.pitch_y * H;
    A.plane_uv = (long)A.pitch_uv * (H / 2);
    A.xr = (float)g.sw / g.dw;
    A.yr = (float)g.sh / g.dh;
    A.KX = (int)std::ceil(A.xr);
    A.K = (int)std::ceil(A.yr);
    std::vector<float> patx, paty;
    if (!area_pattern(A.xr, g.dw, patx, A.PX) || !area_pattern(A.yr, g.dh, paty, A.PY)) return TSVPP_UNSUPPORTED;
    const size_t nx = patx.size(), ny = paty.size();
    const size_t ybytes = (size_t)g.dw * g.dh;
    const size_t fbytes = (ybytes + ybytes / 2 + 255) / 256 * 256 + 256;
    const size_t tab_off = fbytes * (size_t)n;
    unsigned char *tmp = nullptr;
    TSVPP_CUDA(cudaMallocAsync(&tmp, tab_off + (nx + ny) * sizeof(float), stream));
    auto bail = [&](int rc) {
        cudaFreeAsync(tmp, stream);
        return rc;
    };
    // odd sizes: chroma bytes the resize stage never writes (last column, the row under the last luma row)
    // must read as 0, like the oracle's intermediate
    if ((g.dw | g.dh) & 1) {
        cudaError_t me = cudaMemsetAsync(tmp, 0, tab_off, stream);
        if (me != cudaSuccess) return bail(note_error((int)me, "cudaMemsetAsync(area intermediate)", __LINE__));
    }
    float *tab = reinterpret_cast<float *>(tmp + tab_off);
    // pageable sources: the runtime stages them before returning, the vectors may go out of scope
    cudaError_t e = cudaMemcpyAsync(tab, patx.data(), nx * sizeof(float), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(tab + nx, paty.data(), ny * sizeof(float), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return bail(note_error((int)e, "cudaMemcpyAsync(area patterns)", __LINE__));
    A.patx = tab;
    A.paty = tab + nx;
    std::vector<FrameDev> h(n);
    std::vector<TsvppFrame> f2(n);
    for (int i = 0; i < n; ++i) {
        h[i].y = frames[i].y;
        h[i].uv = frames[i].uv;
        h[i].out = tmp + fbytes * i;
        f2[i].y = tmp + fbytes * i;
        f2[i].uv = tmp + fbytes * i + ybytes;
        f2[i].pitch_y = f2[i].pitch_uv = g.dw;
        f2[i].width = g.dw;
        f2[i].height = g.dh;
    }
    int slot = 0;
    int rc = upload_descriptors(ctx, h.data(), n, stream, &slot);
    if (rc) return bail(rc);
    (void)cudaGetLastError();
    rc = launch_area_wide(A, ctx->ring_dev[slot], n, stream);
    release_slot(ctx, slot, stream);
    ctx->launches.fetch_add(1);
    if (rc) return bail(note_error(rc, "launch_area_wide", __LINE__));
    TsvppParams p2 = *p;  // colour stage on the intermediate: no crop, no resize
    p2.crop_x0 = p2.crop_y0 = p2.crop_x1 = p2.crop_y1 = 0;
    p2.dst_width = p2.dst_height = 0;
    p2.resize_type = TSVPP_NEAREST;
    (void)frames_on_host;  // the intermediate lives in device memory
    rc = convert_batch_impl(ctx, f2.data(), n, &p2, out, stream, false);
    return bail(rc);
}

// frames_on_host: the planes are device-mapped pinned host memory (zero-copy host path): keep the
// row-segment bulk copies of the fused kernel instead of the direct kernel's 32-bit loads
static int convert_batch_impl(TsvppContext *ctx, const TsvppFrame *frames, int n, const TsvppParams *p,
                              void *const *out, cudaStream_t stream, bool frames_on_host) {
    if (!ctx || !frames || !p || !out || n <= 0) return TSVPP_ERROR;
    int rc = check_uniform(frames, n);
    if (rc) return rc;
    if (p->fourcc == TSVPP_UYVY || p->fourcc == TSVPP_YUV444)
        return convert_packed_yuv(ctx, frames, n, p, out, stream, frames_on_host);
    const bool bulk = frames_bulk_ok(frames, n) && !env_int("TSVPP_FORCE_LDG", 0);
    std::shared_ptr<Plan> pl;
    rc = get_plan(ctx, frames[0], *p, bulk, &pl);
    if (rc == TSVPP_UNSUPPORTED && area_wide_applies(frames[0], *p))
        return convert_area_wide(ctx, frames, n, p, out, stream, frames_on_host);
    if (rc) return rc;
    std::vector<FrameDev> h(n);
    for (int i = 0; i < n; ++i) {
        h[i].y = frames[i].y;
        h[i].uv = frames[i].uv;
        h[i].out = out[i];
    }
    int slot = 0;
    rc = upload_descriptors(ctx, h.data(), n, stream, &slot);
    if (rc) return rc;
    (void)cudaGetLastError();  // drop stale non-sticky errors of unrelated earlier calls
    rc = launch_plan(pl.get(), ctx->ring_dev[slot], n, vec_ok_for(pl.get(), out, n), grid_for(ctx, pl.get(), n), stream,
                     !frames_on_host);
    release_slot(ctx, slot, stream);
    ctx->launches.fetch_add(1);
    if (rc) return note_error(rc, "launch_fused", __LINE__);
    return rc;
}

int tsvpp_convert_batch(TsvppContext *ctx, const TsvppFrame *frames, int n, const TsvppParams *p, void *const *out,
                        void *stream) {
    if (!ctx) return TSVPP_ERROR;
    return abi_guard([&]() -> int {
        TSVPP_ON_DEVICE(ctx);
        return convert_batch_impl(ctx, frames, n, p, out, (cudaStream_t)stream, false);
    }, __LINE__);
}


// ---------------------------------------------------------------------------------------------------------
// Coalescing of per-frame calls.  The reference's pipeline calls VideoProcessor::Convert once per
// Decoder::GetFrame from up to maxConsumers threads (src/Wrappers/WrapperPython.cpp:265-350, src/Decoder.cpp:97-131);
// a launch per frame tops out far below what one launch over a batch delivers.  Here every caller posts its
// frame to a queue and whichever caller holds the launch lock converts everything pending with ONE fused launch
// per (geometry, parameters) group on the context's coalescing stream ("flat combining": no timer, no extra
// thread -- while one launch is being issued the next batch accumulates).
// ---------------------------------------------------------------------------------------------------------
static void co_thread_main(TsvppContext *ctx);

static int co_init(TsvppContext *ctx) {
    std::call_once(ctx->co_once, [ctx] {
        // a blocking stream, like the reference's per-consumer streams (src/VideoProcessor.cpp:86): implicitly ordered
        // against the legacy default stream
        cudaError_t e = cudaStreamCreate(&ctx->co_stream);
        for (int i = 0; i < TsvppContext::kCoEvents && e == cudaSuccess; ++i)
            e = cudaEventCreateWithFlags(&ctx->co_ev[i], cudaEventDisableTiming);
        if (e == cudaSuccess && !env_int("TSVPP_CO_EVENTS", 0)) {
            void *h = nullptr, *d = nullptr;
            if (cudaHostAlloc(&h, 64, cudaHostAllocMapped) == cudaSuccess && cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) {
                *(volatile unsigned *)h = 0u;
                void *fn = nullptr;
                cudaDriverEntryPointQueryResult qres = cudaDriverEntryPointSymbolNotFound;
                if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
                    qres != cudaDriverEntryPointSuccess)
                    fn = nullptr;
                ctx->co_write32 = reinterpret_cast<CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int)>(fn);
                // one trial write decides between tickets and events
                if (ctx->co_write32 && ctx->co_write32((CUstream)ctx->co_stream, (CUdeviceptr)d, 0u, 0) == CUDA_SUCCESS &&
                    cudaStreamSynchronize(ctx->co_stream) == cudaSuccess) {
                    ctx->co_done_host = (volatile unsigned *)h;
                    ctx->co_done_dev = (CUdeviceptr)d;
                    ctx->co_memops = true;
                } else {
                    (void)cudaGetLastError();
                    cudaFreeHost(h);
                }
            } else {
                (void)cudaGetLastError();
                if (h) cudaFreeHost(h);
            }
        }
        ctx->co_init_rc = (int)e;
        if (e == cudaSuccess && env_int("TSVPP_CO_THREAD", 1)) {
            ctx->co_use_thread = true;
            ctx->co_thread = std::thread(co_thread_main, ctx);
        }
    });
    return ctx->co_init_rc;
}

static void co_drain(TsvppContext *ctx);
static inline bool co_reached(const TsvppContext *ctx, unsigned seq);

static void co_thread_main(TsvppContext *ctx) {
    cudaSetDevice(ctx->device);
    const unsigned depth = (unsigned)std::max(1, env_int("TSVPP_CO_DEPTH", 2));
    // Blocking callers come back in cohorts: the frames of one launch complete together and are resubmitted within
    // a microsecond or two of each other.  Draining the moment the first of them arrives splits the callers into
    // `depth + 1` small cohorts for good (8 consumers: batches of 2.6).  So when fewer frames are pending than recent
    // launches carried, the launcher lingers a few microseconds for the rest of the cohort.
    const long linger_ns = 1000L * std::max(0, env_int("TSVPP_CO_LINGER_US", 4));
    unsigned idle = 0;
    int expected = 1;
    while (!ctx->co_stop.load(std::memory_order_acquire)) {
        if (ctx->co_pending.load(std::memory_order_acquire) > 0) {
            // at most `depth` launches in flight: while the stream is that far behind, let the queue grow
            if (ctx->co_memops && !co_reached(ctx, ctx->co_seq - (depth - 1))) continue;
            if (linger_ns > 0 && ctx->co_pending.load(std::memory_order_acquire) < expected) {
                const auto t0 = std::chrono::steady_clock::now();
                while (ctx->co_pending.load(std::memory_order_acquire) < expected &&
                       (std::chrono::steady_clock::now() - t0).count() < linger_ns) {
                }
            }
            const int before = ctx->co_pending.load(std::memory_order_acquire);
            {
                std::lock_guard<std::mutex> lk(ctx->co_launch);
                co_drain(ctx);
            }
            expected = std::max(before, expected * 3 / 4);
            if (expected < 1) expected = 1;
            idle = 0;
        } else if (++idle > 20000) {  // ~100 us without work: sleep until the next frame is posted
            std::unique_lock<std::mutex> lk(ctx->co_cv_mu);
            ctx->co_sleeping.store(1, std::memory_order_release);
            ctx->co_cv.wait_for(lk, std::chrono::milliseconds(50), [ctx] {
                return ctx->co_pending.load(std::memory_order_acquire) > 0 || ctx->co_stop.load(std::memory_order_acquire);
            });
            ctx->co_sleeping.store(0, std::memory_order_release);
            idle = 0;
            expected = 1;
        }
    }
}

static inline bool co_reached(const TsvppContext *ctx, unsigned seq) {
    return (int)(*ctx->co_done_host - seq) >= 0;  // wrap-safe
}
static inline void co_spin_until(const TsvppContext *ctx, unsigned seq) {
    for (unsigned spins = 0; !co_reached(ctx, seq); ++spins)
        if (spins > 2000) std::this_thread::yield();
}

static bool co_same_group(const TsvppContext::CoReq *a, const TsvppContext::CoReq *b) {
    return a->f.width == b->f.width && a->f.height == b->f.height && a->f.pitch_y == b->f.pitch_y &&
           a->f.pitch_uv == b->f.pitch_uv && std::memcmp(&a->p, &b->p, sizeof(TsvppParams)) == 0;
}

// one group through the ticket path: descriptor slot of the coalescer's own ring, one launch, one stream write
static int co_launch_group(TsvppContext *ctx, const TsvppFrame *frames, void *const *outs, int n, const TsvppParams *p,
                           unsigned *seq_out) {
    int rc = check_uniform(frames, n);
    if (rc) return rc;
    const bool bulk = frames_bulk_ok(frames, n) && !env_int("TSVPP_FORCE_LDG", 0);
    std::shared_ptr<Plan> pl;
    rc = get_plan(ctx, frames[0], *p, bulk, &pl);
    if (rc) return rc;
    TsvppContext::CoSlot &sl = ctx->co_slot[ctx->co_slotnext++ % TsvppContext::kCoSlots];
    if (sl.used) co_spin_until(ctx, sl.seq);  // the launch that read this slot has completed
    if (sl.cap < n) {
        if (sl.host) cudaFreeHost(sl.host);
        if (sl.dev) cudaFree(sl.dev);
        sl.host = sl.dev = sl.host_dev = nullptr;
        sl.cap = 0;
        const int cap = std::max(64, n);
        TSVPP_CUDA(cudaHostAlloc(&sl.host, sizeof(FrameDev) * cap, cudaHostAllocMapped));
        TSVPP_CUDA(cudaHostGetDevicePointer(&sl.host_dev, sl.host, 0));
        TSVPP_CUDA(cudaMalloc(&sl.dev, sizeof(FrameDev) * cap));
        sl.cap = cap;
    }
    for (int i = 0; i < n; ++i) {
        sl.host[i].y = frames[i].y;
        sl.host[i].uv = frames[i].uv;
        sl.host[i].out = outs[i];
    }
    // Small batches: the kernel reads the few descriptors straight from the mapped pinned copy (one PCIe read per
    // CTA, ~1.5 us) -- a stream-ordered H2D copy in front of every launch costs more than that on the GPU side and is
    // one more driver call.  Large batches amortise the copy.
    const FrameDev *descs = sl.host_dev;
    if (n > env_int("TSVPP_CO_ZEROCOPY_MAX", 0)) {
        TSVPP_CUDA(cudaMemcpyAsync(sl.dev, sl.host, sizeof(FrameDev) * n, cudaMemcpyHostToDevice, ctx->co_stream));
        descs = sl.dev;
    }
    (void)cudaGetLastError();
    rc = launch_plan(pl.get(), descs, n, vec_ok_for(pl.get(), outs, n), grid_for(ctx, pl.get(), n), ctx->co_stream, true);
    ctx->launches.fetch_add(1);
    const unsigned seq = ++ctx->co_seq;
    ctx->co_seq_pub.store(seq, std::memory_order_relaxed);
    // the ticket is written even after a failed launch: waiters and the slot must not hang on it
    if (ctx->co_write32((CUstream)ctx->co_stream, ctx->co_done_dev, seq, 0) != CUDA_SUCCESS) {
        cudaStreamSynchronize(ctx->co_stream);
        *ctx->co_done_host = seq;
    }
    sl.seq = seq;
    sl.used = true;
    *seq_out = seq;
    if (rc) return note_error(rc, "launch (coalesced)", __LINE__);
    return TSVPP_OK;
}

// called with ctx->co_launch held
static void co_drain(TsvppContext *ctx) {
    std::vector<TsvppContext::CoReq *> batch;
    std::vector<TsvppContext::CoPosted> posted;
    {
        std::lock_guard<std::mutex> lk(ctx->co_qmu);
        batch.swap(ctx->co_q);
        posted.swap(ctx->co_qp);
    }
    if (batch.empty() && posted.empty()) return;
    // posted frames join the same groups through stack requests of the draining thread
    std::vector<TsvppContext::CoReq> carrier(posted.size());
    for (size_t i = 0; i < posted.size(); ++i) {
        carrier[i].f = posted[i].f;
        carrier[i].p = posted[i].p;
        carrier[i].out = posted[i].out;
        batch.push_back(&carrier[i]);
    }
    const size_t n_waited = batch.size() - posted.size();
    const auto t_begin = std::chrono::steady_clock::now();
    std::vector<char> taken(batch.size(), 0);
    std::vector<TsvppFrame> frames;
    std::vector<void *> outs;
    std::vector<size_t> members;
    for (size_t i = 0; i < batch.size(); ++i) {
        if (taken[i]) continue;
        frames.clear(); outs.clear(); members.clear();
        for (size_t k = i; k < batch.size(); ++k)
            if (!taken[k] && co_same_group(batch[i], batch[k])) {
                taken[k] = 1;
                frames.push_back(batch[k]->f);
                outs.push_back(batch[k]->out);
                members.push_back(k);
            }
        int rc;
        unsigned seq = 0;
        cudaEvent_t ev = nullptr;
        const bool packed = batch[i]->p.fourcc == TSVPP_UYVY || batch[i]->p.fourcc == TSVPP_YUV444;
        try {
            rc = TSVPP_UNSUPPORTED;
            if (ctx->co_memops && !packed)
                rc = co_launch_group(ctx, frames.data(), outs.data(), (int)frames.size(), &batch[i]->p, &seq);
            if (rc == TSVPP_UNSUPPORTED && (packed || !ctx->co_memops || area_wide_applies(frames[0], batch[i]->p))) {
                // multi-stage conversions (packed YUV, wide AREA windows) and the event fallback
                rc = convert_batch_impl(ctx, frames.data(), (int)frames.size(), &batch[i]->p, outs.data(), ctx->co_stream, false);
                ev = ctx->co_ev[ctx->co_evnext++ % TsvppContext::kCoEvents];
                if (rc == TSVPP_OK && cudaEventRecord(ev, ctx->co_stream) != cudaSuccess) rc = (int)cudaGetLastError();
            }
        } catch (...) {
            rc = note_error(TSVPP_ERROR, "C++ exception in a coalesced launch", __LINE__);
        }
        ctx->co_frames.fetch_add(frames.size());
        ctx->co_launches.fetch_add(1);
        for (size_t k : members) {
            batch[k]->rc = rc;
            batch[k]->seq = seq;
            batch[k]->ev = ev;
            batch[k]->done.store(1, std::memory_order_release);  // the request lives on its caller's stack: last touch
        }
    }
    for (size_t i = n_waited; i < batch.size(); ++i)
        if (batch[i]->rc) {
            int expected = 0;
            ctx->co_posted_rc.compare_exchange_strong(expected, batch[i]->rc);
        }
    ctx->co_drain_ns.fetch_add((unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
                                   std::chrono::steady_clock::now() - t_begin).count());
    ctx->co_posted_issued.fetch_add(posted.size(), std::memory_order_release);  // after their tickets were issued
    ctx->co_pending.fetch_sub((int)batch.size(), std::memory_order_acq_rel);
}

void *tsvpp_coalesce_stream(TsvppContext *ctx) {
    if (!ctx) return nullptr;
    DeviceGuard g(ctx->device);
    if (g.err != cudaSuccess || co_init(ctx) != 0) return nullptr;
    return (void *)ctx->co_stream;
}

int tsvpp_convert_coalesced(TsvppContext *ctx, const TsvppFrame *frame, const TsvppParams *p, void *out, int wait) {
    if (!ctx || !frame || !p || !out) return TSVPP_ERROR;
    return abi_guard([&]() -> int {
        TSVPP_ON_DEVICE(ctx);
        if (int rc = co_init(ctx)) return note_error(rc, "coalescing stream", __LINE__);
        if (wait < 0 && ctx->co_use_thread) {
            // posted: the launcher thread picks the frame up; nothing to wait for here
            if (int vrc = validate_geometry(frame->width, frame->height, *p)) return vrc;
            ctx->co_pending.fetch_add(1, std::memory_order_acq_rel);
            {
                std::lock_guard<std::mutex> lk(ctx->co_qmu);
                ctx->co_qp.push_back(TsvppContext::CoPosted{*frame, *p, out});
                ctx->co_posted_total.fetch_add(1, std::memory_order_release);
            }
            if (ctx->co_sleeping.load(std::memory_order_acquire)) {
                std::lock_guard<std::mutex> lk(ctx->co_cv_mu);
                ctx->co_cv.notify_one();
            }
            return TSVPP_OK;
        }
        if (wait < 0) wait = 0;
        TsvppContext::CoReq r;
        r.f = *frame;
        r.p = *p;
        r.out = out;
        ctx->co_pending.fetch_add(1, std::memory_order_acq_rel);
        {
            std::lock_guard<std::mutex> lk(ctx->co_qmu);
            ctx->co_q.push_back(&r);
        }
        if (ctx->co_use_thread) {
            if (ctx->co_sleeping.load(std::memory_order_acquire)) {
                std::lock_guard<std::mutex> lk(ctx->co_cv_mu);
                ctx->co_cv.notify_one();
            }
            for (unsigned spins = 0; !r.done.load(std::memory_order_acquire); ++spins)
                if (spins > 4000) std::this_thread::yield();
            if (r.rc) return r.rc;
            if (wait) {
                if (r.ev) TSVPP_CUDA(cudaEventSynchronize(r.ev));
                else co_spin_until(ctx, r.seq);
            }
            return TSVPP_OK;
        }
        // Whoever gets the launch lock converts everything pending.  At most `depth` launches are kept in flight: a
        // launch issued while the stream is still `depth` launches behind would only queue up, so the would-be
        // leader lets the queue grow instead -- the batch size adapts to the GPU's per-launch latency (without
        // this every frame gets its own launch as soon as the issue path is shorter than that latency).
        const unsigned depth = (unsigned)std::max(1, env_int("TSVPP_CO_DEPTH", 2));
        for (unsigned spins = 0; !r.done.load(std::memory_order_acquire); ++spins) {
            const bool room = !ctx->co_memops || co_reached(ctx, ctx->co_seq_pub.load(std::memory_order_relaxed) - (depth - 1));
            if (room && ctx->co_launch.try_lock()) {
                co_drain(ctx);
                ctx->co_launch.unlock();
            } else if (spins > 2000) {
                std::this_thread::yield();
            }
        }
        if (r.rc) return r.rc;
        if (wait) {
            if (r.ev) TSVPP_CUDA(cudaEventSynchronize(r.ev));
            else co_spin_until(ctx, r.seq);  // poll the ticket word: no driver call
        }
        return TSVPP_OK;
    }, __LINE__);
}

int tsvpp_coalesce_wait(TsvppContext *ctx) {
    if (!ctx) return TSVPP_ERROR;
    TSVPP_ON_DEVICE(ctx);
    if (!ctx->co_stream) return TSVPP_OK;
    // frames posted before this call: wait until the launcher has issued them (other threads may keep posting),
    // then for the newest ticket issued by then
    const unsigned long long target = ctx->co_posted_total.load(std::memory_order_acquire);
    for (unsigned spins = 0; ctx->co_posted_issued.load(std::memory_order_acquire) < target; ++spins)
        if (spins > 2000) std::this_thread::yield();
    if (ctx->co_memops) {
        co_spin_until(ctx, ctx->co_seq_pub.load(std::memory_order_acquire));
    } else {
        TSVPP_CUDA(cudaStreamSynchronize(ctx->co_stream));
    }
    return ctx->co_posted_rc.exchange(0);
}

int tsvpp_coalesce_stats(const TsvppContext *ctx, unsigned long long *frames, unsigned long long *launches) {
    if (!ctx) return TSVPP_ERROR;
    if (frames) *frames = ctx->co_frames.load();
    if (launches) *launches = ctx->co_launches.load();
    return TSVPP_OK;
}

/* nanoseconds spent issuing coalesced launches so far (plan lookup + descriptor upload + launch + event) */
unsigned long long tsvpp_coalesce_issue_ns(const TsvppContext *ctx) { return ctx ? ctx->co_drain_ns.load() : 0ull; }

int tsvpp_batch_create(TsvppContext *ctx, const TsvppFrame *frames, int n, const TsvppParams *p, void *const *out,
                       TsvppBatch **batch) {
    if (!ctx || !frames || !p || !out || !batch || n <= 0) return TSVPP_ERROR;
    return abi_guard([&]() -> int {
        TSVPP_ON_DEVICE(ctx);
        int rc = check_uniform(frames, n);
        if (rc) return rc;
        if (p->fourcc == TSVPP_UYVY || p->fourcc == TSVPP_YUV444) {
            if (int vrc = validate_geometry(frames[0].width, frames[0].height, *p)) return vrc;
            std::unique_ptr<TsvppBatch> b(new TsvppBatch());
            b->ctx = ctx;
            b->packed_yuv = true;
            b->params = *p;
            b->n = n;
            b->frames.assign(frames, frames + n);
            b->outs.assign(out, out + n);
            *batch = b.release();
            return TSVPP_OK;
        }
        const bool bulk = frames_bulk_ok(frames, n) && !env_int("TSVPP_FORCE_LDG", 0);
        std::shared_ptr<Plan> pl;
        rc = get_plan(ctx, frames[0], *p, bulk, &pl);
        if (rc == TSVPP_UNSUPPORTED && area_wide_applies(frames[0], *p)) {  // multi-stage, like packed YUV
            std::unique_ptr<TsvppBatch> b(new TsvppBatch());
            b->ctx = ctx;
            b->packed_yuv = true;
            b->params = *p;
            b->n = n;
            b->frames.assign(frames, frames + n);
            b->outs.assign(out, out + n);
            *batch = b.release();
            return TSVPP_OK;
        }
        if (rc) return rc;
        std::unique_ptr<TsvppBatch> b(new TsvppBatch());
        b->ctx = ctx;
        b->plan = pl;
        b->n = n;
        b->vec_ok = vec_ok_for(pl.get(), out, n);
        b->grid = grid_for(ctx, pl.get(), n);
        std::vector<FrameDev> h(n);
        for (int i = 0; i < n; ++i) {
            h[i].y = frames[i].y;
            h[i].uv = frames[i].uv;
            h[i].out = out[i];
        }
        TSVPP_CUDA(cudaMalloc(&b->frames_dev, sizeof(FrameDev) * n));
        cudaError_t e = cudaMemcpy(b->frames_dev, h.data(), sizeof(FrameDev) * n, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(b->frames_dev);
            return note_error((int)e, "cudaMemcpy(batch descriptors)", __LINE__);
        }
        *batch = b.release();
        return TSVPP_OK;
    }, __LINE__);
}

int tsvpp_batch_run(TsvppBatch *b, void *stream) {
    if (!b) return TSVPP_ERROR;
    if (b->packed_yuv)
        return tsvpp_convert_batch(b->ctx, b->frames.data(), b->n, &b->params, b->outs.data(), stream);
    TSVPP_ON_DEVICE(b->ctx);
    const Plan *pl = b->plan.get();
    (void)cudaGetLastError();
    int rc = launch_plan(pl, b->frames_dev, b->n, b->vec_ok, b->grid, (cudaStream_t)stream, true);
    b->ctx->launches.fetch_add(1);
    if (rc) return note_error(rc, "launch_fused", __LINE__);
    return rc;
}

void tsvpp_batch_destroy(TsvppBatch *b) {
    if (!b) return;
    {
        DeviceGuard g(b->ctx->device);
        if (b->frames_dev) cudaFree(b->frames_dev);
        b->plan.reset();
    }
    delete b;
}

static int convert_batch_host_impl(TsvppContext *ctx, const TsvppFrame *frames, int n, const TsvppParams *p,
                                   void *const *out) {
    int rc = check_uniform(frames, n);
    if (rc) return rc;
    const int W = frames[0].width, H = frames[0].height;
    const int py = frames[0].pitch_y ? frames[0].pitch_y : W, puv = frames[0].pitch_uv ? frames[0].pitch_uv : W;
    size_t elems = 0;
    int es = 1, ow = 0, oh = 0;
    rc = tsvpp_output_info(W, H, p, &ow, &oh, &elems, &es);
    if (rc) return rc;
    const size_t in_y = (size_t)py * H, in_uv = (size_t)puv * (H / 2);
    const size_t in_y_pad = (in_y + 255) / 256 * 256, in_uv_pad = (in_uv + 255) / 256 * 256;
    const size_t in_frame = in_y_pad + in_uv_pad;
    const size_t out_frame = (elems * es + 255) / 256 * 256;
    for (int s = 0; s < 2; ++s) {
        if (!ctx->hstream[s]) TSVPP_CUDA(cudaStreamCreateWithFlags(&ctx->hstream[s], cudaStreamNonBlocking));
        if (!ctx->hev[s]) TSVPP_CUDA(cudaEventCreateWithFlags(&ctx->hev[s], cudaEventDisableTiming));
    }
    // Zero-copy input: when every input plane lives in pinned (page-locked, device-mapped) host
    // memory the fused kernel reads it in place over PCIe -- its bulk copies fetch only the source
    // rows the taps touch (1.47 of 3.11 MB per 1080p frame for the 256x256 bilinear case), which a
    // whole-frame cudaMemcpy cannot do.  Pageable inputs go through staged H2D copies.
    bool zero_copy = env_int("TSVPP_HOST_ZEROCOPY", 1) != 0;
    std::vector<const uint8_t *> zy, zuv;
    if (zero_copy) {
        zy.resize(n);
        zuv.resize(n);
        for (int i = 0; i < n && zero_copy; ++i) {
            cudaPointerAttributes a;
            if (cudaPointerGetAttributes(&a, frames[i].y) != cudaSuccess || a.type != cudaMemoryTypeHost ||
                !a.devicePointer) {
                zero_copy = false;
                break;
            }
            zy[i] = (const uint8_t *)a.devicePointer;
            if (cudaPointerGetAttributes(&a, frames[i].uv) != cudaSuccess || a.type != cudaMemoryTypeHost ||
                !a.devicePointer) {
                zero_copy = false;
                break;
            }
            zuv[i] = (const uint8_t *)a.devicePointer;
        }
        (void)cudaGetLastError();
    }
    // Zero-copy output: pinned host output buffers are written by the kernel's own (coalesced,
    // streaming) stores, full-duplex with the input reads -- no staging buffer, no D2H copy.
    bool zc_out = env_int("TSVPP_HOST_ZEROCOPY_OUT", 1) != 0;
    std::vector<void *> zout;
    if (zc_out) {
        zout.resize(n);
        for (int i = 0; i < n; ++i) {
            cudaPointerAttributes a;
            if (cudaPointerGetAttributes(&a, out[i]) != cudaSuccess || a.type != cudaMemoryTypeHost || !a.devicePointer) {
                zc_out = false;
                break;
            }
            zout[i] = a.devicePointer;
        }
        (void)cudaGetLastError();
    }
    ctx->last_host_zero_copy = (zero_copy ? 1 : 0) | (zc_out ? 2 : 0);
    if (zero_copy && zc_out) {
        // Nothing to stage on either side.  The batch is split over the two internal streams so that the
        // (latency-bound) PCIe reads of one half overlap the posted writes of the other.
        std::vector<TsvppFrame> f(n);
        for (int i = 0; i < n; ++i) {
            f[i] = frames[i];
            f[i].y = zy[i];
            f[i].uv = zuv[i];
        }
        const int nsplit = (n >= 8 && env_int("TSVPP_HOST_SPLIT", 1)) ? 2 : 1;
        const int half = (n + nsplit - 1) / nsplit;
        for (int k = 0; k < nsplit; ++k) {
            const int b0 = k * half, m = std::min(half, n - b0);
            if (m <= 0) break;
            rc = convert_batch_impl(ctx, f.data() + b0, m, p, zout.data() + b0, ctx->hstream[k], true);
            if (rc) break;
        }
        cudaError_t e0 = cudaStreamSynchronize(ctx->hstream[0]);
        cudaError_t e1 = cudaStreamSynchronize(ctx->hstream[1]);
        if (rc) return rc;
        if (e0 != cudaSuccess) return note_error((int)e0, "cudaStreamSynchronize", __LINE__);
        if (e1 != cudaSuccess) return note_error((int)e1, "cudaStreamSynchronize", __LINE__);
        return TSVPP_OK;
    }
    // chunks of frames ping-pong over two streams so that H2D, kernel and D2H overlap
    const int chunk = std::max(1, std::min(n, env_int("TSVPP_HOST_CHUNK", 16)));
    for (int s = 0; s < 2; ++s) {
        if (!zero_copy && ctx->stage_in_cap[s] < in_frame * chunk) {
            if (ctx->stage_in[s]) cudaFree(ctx->stage_in[s]);
            ctx->stage_in[s] = nullptr;
            ctx->stage_in_cap[s] = 0;
            TSVPP_CUDA(cudaMalloc(&ctx->stage_in[s], in_frame * chunk));
            ctx->stage_in_cap[s] = in_frame * chunk;
        }
        if (!zc_out && ctx->stage_out_cap[s] < out_frame * chunk) {
            if (ctx->stage_out[s]) cudaFree(ctx->stage_out[s]);
            ctx->stage_out[s] = nullptr;
            ctx->stage_out_cap[s] = 0;
            TSVPP_CUDA(cudaMalloc(&ctx->stage_out[s], out_frame * chunk));
            ctx->stage_out_cap[s] = out_frame * chunk;
        }
    }
    std::vector<TsvppFrame> dframes(chunk);
    std::vector<void *> douts(chunk);
    int which = 0;
    for (int base = 0; base < n; base += chunk, which ^= 1) {
        const int m = std::min(chunk, n - base);
        cudaStream_t st = ctx->hstream[which];
        for (int k = 0; k < m; ++k) {
            dframes[k] = frames[base + k];
            if (zero_copy) {
                dframes[k].y = zy[base + k];
                dframes[k].uv = zuv[base + k];
            } else {
                unsigned char *dy = ctx->stage_in[which] + in_frame * k;
                unsigned char *duv = dy + in_y_pad;
                TSVPP_CUDA(cudaMemcpyAsync(dy, frames[base + k].y, in_y, cudaMemcpyHostToDevice, st));
                TSVPP_CUDA(cudaMemcpyAsync(duv, frames[base + k].uv, in_uv, cudaMemcpyHostToDevice, st));
                dframes[k].y = dy;
                dframes[k].uv = duv;
            }
            douts[k] = zc_out ? zout[base + k] : (void *)(ctx->stage_out[which] + out_frame * k);
        }
        rc = convert_batch_impl(ctx, dframes.data(), m, p, douts.data(), st, zero_copy);
        if (rc) break;
        if (!zc_out)
            for (int k = 0; k < m; ++k)
                TSVPP_CUDA(cudaMemcpyAsync(out[base + k], douts[k], elems * es, cudaMemcpyDeviceToHost, st));
        // the staging buffers of this stream are reused two chunks later: the stream order
        // already serialises that, nothing to wait for on the host here
    }
    cudaError_t e0 = cudaStreamSynchronize(ctx->hstream[0]);
    cudaError_t e1 = cudaStreamSynchronize(ctx->hstream[1]);
    if (rc) return rc;
    if (e0 != cudaSuccess) return note_error((int)e0, "cudaStreamSynchronize", __LINE__);
    if (e1 != cudaSuccess) return note_error((int)e1, "cudaStreamSynchronize", __LINE__);
    return TSVPP_OK;
}

int tsvpp_convert_batch_host(TsvppContext *ctx, const TsvppFrame *frames, int n, const TsvppParams *p,
                             void *const *out) {
    if (!ctx || !frames || !p || !out || n <= 0) return TSVPP_ERROR;
    return abi_guard([&]() -> int {
        TSVPP_ON_DEVICE(ctx);
        // one host-buffer call at a time per context: it owns the two internal streams and the staging buffers
        std::lock_guard<std::mutex> lk(ctx->host_mu);
        return convert_batch_host_impl(ctx, frames, n, p, out);
    }, __LINE__);
}

int tsvpp_debug_pow23(const double *w, int n, double *w2, double *w3) {
    if (!w || !w2 || !w3 || n <= 0) return TSVPP_ERROR;
    double *d = nullptr;
    TSVPP_CUDA(cudaMalloc(&d, sizeof(double) * 3 * (size_t)n));
    cudaError_t e = cudaMemcpy(d, w, sizeof(double) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        launch_pow23(d, d + n, d + 2 * (size_t)n, n, 0);
        e = cudaMemcpy(w2, d + n, sizeof(double) * n, cudaMemcpyDeviceToHost);
    }
    if (e == cudaSuccess) e = cudaMemcpy(w3, d + 2 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return note_error((int)e, "tsvpp_debug_pow23", __LINE__);
    return TSVPP_OK;
}

int tsvpp_last_host_path(const TsvppContext *ctx) { return ctx ? ctx->last_host_zero_copy : 0; }

unsigned long long tsvpp_launch_count(const TsvppContext *ctx) { return ctx ? ctx->launches.load() : 0ull; }

int tsvpp_plan_cache_size(TsvppContext *ctx) {
    if (!ctx) return 0;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return (int)ctx->plans.size();
}

int tsvpp_describe_plan(TsvppContext *ctx, const TsvppFrame *f0, const TsvppParams *p, char *buf, size_t buflen) {
    if (!ctx || !f0 || !p || !buf || !buflen) return TSVPP_ERROR;
    return abi_guard([&]() -> int {
        TSVPP_ON_DEVICE(ctx);
        std::shared_ptr<Plan> pl;
        const bool bulk = frames_bulk_ok(f0, 1) && !env_int("TSVPP_FORCE_LDG", 0);
        int rc = get_plan(ctx, *f0, *p, bulk, &pl);
        if (rc == TSVPP_UNSUPPORTED && area_wide_applies(*f0, *p)) {
            std::snprintf(buf, buflen, "kernel=area-wide resample=area (window too large for the fused stages: resize into a "
                                       "tight NV12 intermediate, then the colour stage)");
            return TSVPP_OK;
        }
        if (rc) return rc;
        std::snprintf(buf, buflen, "%s", pl->desc.c_str());
        return TSVPP_OK;
    }, __LINE__);
}

}  // extern "C"
