// VideoProcessor.cpp -- the reference's VideoProcessor API (include/VideoProcessor.h) on top of
// the fused B200 path.  Behavioural contract kept from reference src/VideoProcessor.cpp:
//   * one CUDA stream per consumer name, bound on first use (:98-104); no free slot -> VREADER_ERROR
//   * crop / resize / colour dispatch rules (:106-143) -- implemented once, in tsvpp_capi.cpp
//   * options.resize is filled in with the source size when it was 0 (:132-135)
//   * output->opaque = fresh device buffer owned by the caller; output->width/height = final size
//   * optional raw dump "Processed_<consumer>.yuv", append mode (:152-163)
//   * av_frame_unref(input) (:164)
// What is gone: the 3-7 cudaMalloc and 2-4 synchronising cudaFree per frame, the NV12
// intermediates in HBM and the per-frame AREA pattern upload.
#include "VideoProcessor.h"

#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <map>
#include <mutex>
#include <thread>

#include "tsvpp.h"

// NVTX ranges with the reference's names and colour (include/Common.h:72-105 PUSH_RANGE; pushed at
// src/VideoProcessor.cpp:51,80,95) when the Logger asks for them.
namespace {
class NVTXTracer {
public:
    void trace(const char *name, uint32_t argb) {
        nvtxEventAttributes_t a = {};
        a.version = NVTX_VERSION;
        a.size = NVTX_EVENT_ATTRIB_STRUCT_SIZE;
        a.colorType = NVTX_COLOR_ARGB;
        a.color = argb;
        a.messageType = NVTX_MESSAGE_TYPE_ASCII;
        a.message.ascii = name;
        nvtxRangePushEx(&a);
        pushed = true;
    }
    ~NVTXTracer() {
        if (pushed) nvtxRangePop();
    }

private:
    bool pushed = false;
};
constexpr uint32_t kNvtxYellow = 0xffffff00u;  // NVTXColors::YELLOW
}  // namespace
#define PUSH_RANGE(name, colorID)             \
    NVTXTracer tracer;                        \
    if (logger && logger->enableNVTX) {       \
        tracer.trace(name, colorID);          \
    }

// ---------------------------------------------------------------------------------------------
// AVFrame layout agreement between this library and its callers (see include/VideoProcessor.h)
// ---------------------------------------------------------------------------------------------
namespace {
std::atomic<int> g_avframe_mismatch{0};
}
extern "C" const char *tsvpp_vp_avframe_flavour(void) { return TSVPP_AVFRAME_FLAVOUR; }
extern "C" int tsvpp_vp_register_avframe_layout(size_t off_linesize, size_t off_width, size_t off_height, size_t off_opaque,
                                                size_t size) {
    const bool same = off_linesize == offsetof(AVFrame, linesize) && off_width == offsetof(AVFrame, width) &&
                      off_height == offsetof(AVFrame, height) && off_opaque == offsetof(AVFrame, opaque) &&
                      size == sizeof(AVFrame);
    if (!same) {
        g_avframe_mismatch.store(1);
        std::cout << "libtsvpp: AVFrame layout mismatch: the library was built against the '" << TSVPP_AVFRAME_FLAVOUR
                  << "' AVFrame (opaque at " << offsetof(AVFrame, opaque) << ", size " << sizeof(AVFrame)
                  << "), the caller has opaque at " << off_opaque << ", size " << size
                  << "; rebuild libtsvpp.so with the caller's headers\n" << std::flush;
    }
    return same ? 0 : -1;
}

float channelsByFourCC(FourCC fourCC) {
    switch (fourCC) {
        case Y800: return 1.f;
        case UYVY: return 2.f;
        case NV12: return 1.5f;
        default: return 3.f;
    }
}

float channelsByFourCC(std::string fourCC) {
    if (fourCC == "Y800") return 1.f;
    if (fourCC == "UYVY") return 2.f;
    if (fourCC == "NV12") return 1.5f;
    return 3.f;
}

namespace {
TsvppParams flatten(const FrameParameters &o) {
    TsvppParams p;
    p.crop_x0 = std::get<0>(o.crop.leftTopCorner);
    p.crop_y0 = std::get<1>(o.crop.leftTopCorner);
    p.crop_x1 = std::get<0>(o.crop.rightBottomCorner);
    p.crop_y1 = std::get<1>(o.crop.rightBottomCorner);
    p.dst_width = (int)o.resize.width;
    p.dst_height = (int)o.resize.height;
    p.resize_type = (int)o.resize.type;
    p.fourcc = (int)o.color.dstFourCC;
    p.planes = (int)o.color.planesPos;
    p.normalization = o.color.normalization ? 1 : 0;
    return p;
}

// size written by a dump: crop size, overridden by the resize size (reference :28-47)
void dump_size(const FrameParameters &o, int &w, int &h) {
    w = h = 0;
    const int cw = std::get<0>(o.crop.rightBottomCorner) - std::get<0>(o.crop.leftTopCorner);
    const int ch = std::get<1>(o.crop.rightBottomCorner) - std::get<1>(o.crop.leftTopCorner);
    if (cw > 0 && ch > 0) {
        w = cw;
        h = ch;
    }
    if (o.resize.width > 0 && o.resize.height > 0) {
        w = (int)o.resize.width;
        h = (int)o.resize.height;
    }
}
}  // namespace

// ---------------------------------------------------------------------------------------------
// Stage-level entry points (reference include/VideoProcessor.h:109-115)
// ---------------------------------------------------------------------------------------------
namespace {
// one library context per device for callers that have no VideoProcessor object; lives as long as the process
TsvppContext *stage_context() {
    static std::mutex mu;
    static std::map<int, TsvppContext *> by_device;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    auto it = by_device.find(dev);
    if (it != by_device.end()) return it->second;
    TsvppContext *c = nullptr;
    if (tsvpp_create(dev, &c) != TSVPP_OK) return nullptr;
    by_device[dev] = c;
    return c;
}

TsvppFrame frame_of(const AVFrame *a) {
    TsvppFrame f;
    f.y = a->data[0];
    f.uv = a->data[1];
    f.pitch_y = a->linesize[0];  // 0 => width, resolved by the library like the reference (src/Resize.cu:420-421)
    f.pitch_uv = a->linesize[1];
    f.width = a->width;
    f.height = a->height;
    return f;
}

// crop and/or resize to a tight NV12 image in two separately cudaFree-able buffers
int nv12_stage(const AVFrame *src, const TsvppParams &p, unsigned char **y_out, unsigned char **uv_out,
               cudaStream_t stream) {
    TsvppContext *ctx = stage_context();
    if (!ctx) return VREADER_ERROR;
    const TsvppFrame f = frame_of(src);
    int w = 0, h = 0, es = 1;
    size_t elems = 0;
    int rc = tsvpp_output_info(f.width, f.height, &p, &w, &h, &elems, &es);
    if (rc) return rc;
    const size_t y_bytes = (size_t)w * h, uv_bytes = (size_t)w * (h / 2);
    unsigned char *both = nullptr, *uv = nullptr;
    cudaError_t e = cudaMalloc(&both, elems * (size_t)es);  // Y plane, then the UV plane the fused NV12 output appends
    if (e != cudaSuccess) return (int)e;
    e = cudaMalloc(&uv, uv_bytes ? uv_bytes : 1);
    if (e != cudaSuccess) {
        cudaFree(both);
        return (int)e;
    }
    void *outs[1] = {both};
    rc = tsvpp_convert_batch(ctx, &f, 1, &p, outs, stream);
    if (rc == TSVPP_OK && uv_bytes)
        rc = (int)cudaMemcpyAsync(uv, both + y_bytes, uv_bytes, cudaMemcpyDeviceToDevice, stream);
    if (rc) {
        cudaFree(both);
        cudaFree(uv);
        return rc;
    }
    *y_out = both;  // the caller's cudaFree(data[0]) releases the whole allocation
    *uv_out = uv;
    return VREADER_OK;
}
}  // namespace

int cropHost(AVFrame *src, AVFrame *dst, CropOptions crop, int /*maxThreadsPerBlock*/, cudaStream_t *stream) {
    if (!src || !dst || !stream) return VREADER_ERROR;
    const int x0 = std::get<0>(crop.leftTopCorner), y0 = std::get<1>(crop.leftTopCorner);
    const int x1 = std::get<0>(crop.rightBottomCorner), y1 = std::get<1>(crop.rightBottomCorner);
    const int cw = x1 - x0, ch = y1 - y0;
    if (cw <= 0 || ch <= 0 || x0 < 0 || y0 < 0 || x1 > src->width || y1 > src->height) return VREADER_ERROR;
    // the library applies a crop under Convert's rule (0 < cw < W && 0 < ch < H, src/VideoProcessor.cpp:109);
    // a box that spans a full dimension is only the identity here
    const bool active = cw < src->width && ch < src->height;
    if (!active && (x0 != 0 || y0 != 0 || cw != src->width || ch != src->height)) return VREADER_UNSUPPORTED;
    TsvppParams p;
    p.crop_x0 = active ? x0 : 0;
    p.crop_y0 = active ? y0 : 0;
    p.crop_x1 = active ? x1 : 0;
    p.crop_y1 = active ? y1 : 0;
    p.dst_width = p.dst_height = 0;
    p.resize_type = (int)NEAREST;
    p.fourcc = (int)NV12;
    p.planes = (int)MERGED;
    p.normalization = 0;
    unsigned char *y = nullptr, *uv = nullptr;
    const int rc = nv12_stage(src, p, &y, &uv, *stream);
    if (rc) return rc;
    dst->data[0] = y;
    dst->data[1] = uv;
    return VREADER_OK;
}

int resizeKernel(AVFrame *src, AVFrame *dst, bool crop, ResizeOptions resize, int /*maxThreadsPerBlock*/,
                 cudaStream_t *stream) {
    if (!src || !dst || !stream || resize.width == 0 || resize.height == 0) return VREADER_ERROR;
    TsvppParams p;
    p.crop_x0 = p.crop_y0 = p.crop_x1 = p.crop_y1 = 0;
    p.dst_width = (int)resize.width;
    p.dst_height = (int)resize.height;
    p.resize_type = (int)resize.type;
    p.fourcc = (int)NV12;
    p.planes = (int)MERGED;
    p.normalization = 0;
    unsigned char *y = nullptr, *uv = nullptr;
    const int rc = nv12_stage(src, p, &y, &uv, *stream);
    if (rc) return rc;
    if (crop) {  // the crop intermediates (dst may be src): cudaFree waits for the work queued above
        cudaFree(dst->data[0]);
        cudaFree(dst->data[1]);
    }
    dst->data[0] = y;
    dst->data[1] = uv;
    return VREADER_OK;
}

template <class T>
int colorConversionKernel(AVFrame *src, AVFrame *dst, ColorOptions color, int /*maxThreadsPerBlock*/,
                          cudaStream_t *stream) {
    if (!src || !dst || !stream) return VREADER_ERROR;
    TsvppContext *ctx = stage_context();
    if (!ctx) return VREADER_ERROR;
    TsvppParams p;
    p.crop_x0 = p.crop_y0 = p.crop_x1 = p.crop_y1 = 0;
    p.dst_width = p.dst_height = 0;
    p.resize_type = (int)NEAREST;
    p.fourcc = (int)color.dstFourCC;
    p.planes = (int)color.planesPos;
    p.normalization = color.normalization ? 1 : 0;
    const TsvppFrame f = frame_of(src);
    int w = 0, h = 0, es = 1;
    size_t elems = 0;
    int rc = tsvpp_output_info(f.width, f.height, &p, &w, &h, &elems, &es);
    if (rc) return rc;  // unknown FourCC: cudaErrorMissingConfiguration (src/ColorConversion.cu:376)
    if ((size_t)es != sizeof(T)) return VREADER_UNSUPPORTED;
    void *buf = nullptr;
    cudaError_t e = cudaMalloc(&buf, elems * sizeof(T));
    if (e != cudaSuccess) return (int)e;
    void *outs[1] = {buf};
    rc = tsvpp_convert_batch(ctx, &f, 1, &p, outs, *stream);
    if (rc) {
        cudaFree(buf);
        return rc;
    }
    dst->opaque = buf;
    return VREADER_OK;
}
template int colorConversionKernel<unsigned char>(AVFrame *, AVFrame *, ColorOptions, int, cudaStream_t *);
template int colorConversionKernel<float>(AVFrame *, AVFrame *, ColorOptions, int, cudaStream_t *);

int VideoProcessor::Init(std::shared_ptr<Logger> lg, uint8_t maxConsumers, bool _enableDumps) {
    enableDumps = _enableDumps;
    logger = lg;
    PUSH_RANGE("VideoProcessor::Init", kNvtxYellow);
    if (g_avframe_mismatch.load()) return VREADER_ERROR;  // see tsvpp_vp_register_avframe_layout
    // the reference always queried device 0 (src/VideoProcessor.cpp:83); use the current one
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return (int)e;
    int rc = tsvpp_create(device, &ctx);
    if (rc) return rc;
    // Results come from the device's default memory pool (cudaMallocAsync) and callers release them
    // with cudaFree, which synchronises; keep the freed blocks in the pool instead of handing them
    // back to the OS at every synchronisation, or each Convert would pay a fresh OS allocation.
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    for (int i = 0; i < maxConsumers; ++i) {
        cudaStream_t s;
        // blocking streams like the reference (:86): implicitly ordered against the legacy default stream
        e = cudaStreamCreate(&s);
        if (e != cudaSuccess) return (int)e;
        streamArr.push_back(std::make_pair(std::string("empty"), s));
    }
    {
        const char *e = std::getenv("TSVPP_COALESCE");
        coalesce = !(e && *e == '0');
        if (coalesce) {
            coStream = (cudaStream_t)tsvpp_coalesce_stream(ctx);
            if (!coStream) return VREADER_ERROR;
        }
    }
    isClosed = false;
    return VREADER_OK;
}

int VideoProcessor::Convert(AVFrame *input, AVFrame *output, FrameParameters &options, std::string consumerName) {
    PUSH_RANGE("VideoProcessor::Convert", kNvtxYellow);
    return convertImpl(input, output, nullptr, options, consumerName);
}

int VideoProcessor::ConvertInto(AVFrame *input, AVFrame *output, void *dst, FrameParameters &options, std::string consumerName) {
    PUSH_RANGE("VideoProcessor::ConvertInto", kNvtxYellow);
    if (!dst) return VREADER_ERROR;
    return convertImpl(input, output, dst, options, consumerName);
}

int VideoProcessor::WaitCoalesced() { return (ctx && coalesce) ? tsvpp_coalesce_wait(ctx) : VREADER_OK; }

int VideoProcessor::convertImpl(AVFrame *input, AVFrame *output, void *callerDst, FrameParameters &options,
                                const std::string &consumerName) {
    if (isClosed || !ctx) return VREADER_ERROR;
    cudaStream_t stream;
    {
        std::unique_lock<std::mutex> locker(streamSync);
        stream = findFree<cudaStream_t>(consumerName, streamArr);
        if (stream == nullptr) return VREADER_ERROR;
    }
    TsvppFrame f;
    f.y = input->data[0];
    f.uv = input->data[1];
    f.pitch_y = input->linesize[0];
    f.pitch_uv = input->linesize[1];
    f.width = input->width;
    f.height = input->height;
    const TsvppParams p = flatten(options);
    int ow = 0, oh = 0, es = 1;
    size_t elems = 0;
    int rc = tsvpp_output_info(f.width, f.height, &p, &ow, &oh, &elems, &es);
    if (rc) return rc;
    // The reference's Convert returns with the work complete in every resize/crop path (its
    // cudaFree calls synchronise the device).  Keep that guarantee unless the caller opts out.
    // TSVPP_ASYNC: 0 (default) complete on return; 1 return once the launch is issued (stream-ordered); 2 posted --
    // return at once, results are valid after WaitCoalesced() (inputs must stay alive until then)
    static const int async_level = [] {
        const char *s = std::getenv("TSVPP_ASYNC");
        return (s && *s >= '0' && *s <= '2') ? *s - '0' : 0;
    }();
    const bool async_mode = async_level != 0;
    void *dst = callerDst;
    cudaError_t e;
    if (coalesce) {
        // the result buffer is allocated on the coalescing stream by the caller's own thread (stream-ordered, no
        // lock of ours), then the frame joins whatever launch is being assembled
        if (!callerDst) {
            e = cudaMallocAsync(&dst, elems * (size_t)es, coStream);
            if (e != cudaSuccess) return (int)e;
        }
        rc = tsvpp_convert_coalesced(ctx, &f, &p, dst, async_level == 0 ? 1 : async_level == 1 ? 0 : -1);
        if (rc) {
            if (!callerDst) cudaFreeAsync(dst, coStream);
            return rc;
        }
    } else {
        if (!callerDst) {
            e = cudaMallocAsync(&dst, elems * (size_t)es, stream);
            if (e != cudaSuccess) return (int)e;
        }
        void *outs[1] = {dst};
        rc = tsvpp_convert_batch(ctx, &f, 1, &p, outs, stream);
        if (rc) {
            if (!callerDst) cudaFreeAsync(dst, stream);
            return rc;
        }
        if (!async_mode) {
            e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess) return (int)e;
        }
    }
    output->opaque = dst;
    output->width = ow;
    output->height = oh;
    output->data[0] = output->data[1] = nullptr;  // no dangling intermediates
    if (options.resize.width == 0 || options.resize.height == 0) {  // reference :132-135
        options.resize.width = (unsigned)ow;
        options.resize.height = (unsigned)oh;
    }
    if (enableDumps) {
        const std::string fileName = std::string("Processed_") + consumerName + std::string(".yuv");
        std::shared_ptr<FILE> dumpFile(fopen(fileName.c_str(), "ab"), std::fclose);
        std::unique_lock<std::mutex> locker(dumpSync);
        if (es == 4)
            DumpFrame(static_cast<float *>(output->opaque), options, dumpFile);
        else
            DumpFrame(static_cast<unsigned char *>(output->opaque), options, dumpFile);
    }
    av_frame_unref(input);
    return VREADER_OK;
}

template <class T>
int VideoProcessor::DumpFrame(T *output, FrameParameters options, std::shared_ptr<FILE> dumpFile) {
    PUSH_RANGE("VideoProcessor::DumpFrame", kNvtxYellow);
    int w, h;
    dump_size(options, w, h);
    const size_t n = (size_t)(channelsByFourCC(options.color.dstFourCC) * w * h);
    std::unique_ptr<T[]> host(new T[n ? n : 1]);
    cudaError_t e = cudaMemcpy(host.get(), output, n * sizeof(T), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return (int)e;
    if (!dumpFile) return VREADER_ERROR;
    fwrite(host.get(), n, sizeof(T), dumpFile.get());
    fflush(dumpFile.get());
    return VREADER_OK;
}
template int VideoProcessor::DumpFrame(float *, FrameParameters, std::shared_ptr<FILE>);
template int VideoProcessor::DumpFrame(uint8_t *, FrameParameters, std::shared_ptr<FILE>);

void VideoProcessor::Close() {
    if (isClosed) return;
    if (ctx && std::getenv("TSVPP_COALESCE_STATS")) {
        unsigned long long frames = 0, launches = 0;
        tsvpp_coalesce_stats(ctx, &frames, &launches);
        std::cout << "coalesce_stats frames " << frames << " launches " << launches << " mean_batch "
                  << (launches ? (double)frames / (double)launches : 0.0) << " issue_us_per_launch "
                  << (launches ? (double)tsvpp_coalesce_issue_ns(ctx) / 1e3 / (double)launches : 0.0) << "\n" << std::flush;
    }
    for (auto &item : streamArr) cudaStreamDestroy(item.second);
    streamArr.clear();
    if (ctx) tsvpp_destroy(ctx);
    ctx = nullptr;
    isClosed = true;
}

VideoProcessor::~VideoProcessor() { Close(); }
