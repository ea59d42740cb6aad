// VideoProcessor.h -- drop-in replacement of the reference's include/VideoProcessor.h for the
// VPP hot path.  Same enum values, same option structs, same VideoProcessor member signatures
// (reference include/VideoProcessor.h:20-149), so TensorStream::getFrame
// (src/Wrappers/WrapperPython.cpp:312, src/Wrappers/WrapperC.cpp:291) and tests written against
// the reference (tests/src/VPPTests.cpp) compile unchanged -- but Convert() runs ONE fused
// sm_100a kernel through the C ABI in tsvpp.h instead of crop + resize + colour kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdio>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "Common.h"
#include "tsvpp_avframe.h"

struct TsvppContext;

enum FourCC { Y800 = 0, RGB24, BGR24, NV12, UYVY, YUV444, HSV };
enum Planes { PLANAR = 0, MERGED };
enum ResizeType { NEAREST = 0, BILINEAR, BICUBIC, AREA };

struct ColorOptions {
    ColorOptions(FourCC fourCC = FourCC::RGB24)
        : normalization(fourCC == FourCC::HSV), planesPos(Planes::MERGED), dstFourCC(fourCC) {}
    bool normalization;
    Planes planesPos;
    FourCC dstFourCC;
};

struct ResizeOptions {
    ResizeOptions(int w = 0, int h = 0) : width((unsigned)w), height((unsigned)h), type(ResizeType::NEAREST) {}
    unsigned int width;  // 0 => keep source size
    unsigned int height;
    ResizeType type;
};

struct CropOptions {
    CropOptions(std::tuple<int, int> lt = {0, 0}, std::tuple<int, int> rb = {0, 0})
        : leftTopCorner(lt), rightBottomCorner(rb) {}
    std::tuple<int, int> leftTopCorner;  // zero-area box => no crop
    std::tuple<int, int> rightBottomCorner;
};

struct FrameParameters {
    FrameParameters(ResizeOptions r = ResizeOptions(), ColorOptions c = ColorOptions(), CropOptions k = CropOptions())
        : resize(r), color(c), crop(k) {}
    ResizeOptions resize;
    ColorOptions color;
    CropOptions crop;
};

float channelsByFourCC(FourCC fourCC);
float channelsByFourCC(std::string fourCC);

// The reference's stage-level entry points (reference include/VideoProcessor.h:109-115; bodies in
// src/ColorConversion.cu:280-382, src/Resize.cu:408-473, src/Crop.cu:23-48).  Convert() does NOT go through
// them here -- it runs crop + resize + colour as one kernel -- but code written against the individual
// stages keeps working: each function is the fused path restricted to that stage, with the reference's
// buffer contract (fresh cudaMalloc'd, tight, cudaFree-able buffers in dst->data[0], dst->data[1] or
// dst->opaque; resizeKernel frees the crop intermediates when `crop` is set; dst->width/height are left to the
// caller, as in the reference).  A cropHost -> resizeKernel -> colorConversionKernel chain yields the bytes
// Convert() yields.  maxThreadsPerBlock is accepted and ignored.  colorConversionKernel<float> needs
// color.normalization (or HSV) and <unsigned char> needs it off -- the pairing Convert() uses
// (src/VideoProcessor.cpp:139-142); the other two pairings return VREADER_UNSUPPORTED.
template <class T>
int colorConversionKernel(AVFrame *src, AVFrame *dst, ColorOptions color, int maxThreadsPerBlock, cudaStream_t *stream);
int resizeKernel(AVFrame *src, AVFrame *dst, bool crop, ResizeOptions resize, int maxThreadsPerBlock, cudaStream_t *stream);
int cropHost(AVFrame *src, AVFrame *dst, CropOptions crop, int maxThreadsPerBlock, cudaStream_t *stream);

// Which AVFrame does the code that INCLUDES this header see?  libtsvpp.so is compiled against either FFmpeg's
// <libavutil/frame.h> or, where no FFmpeg headers exist, the minimal struct of tsvpp_avframe.h; the two lay
// `width`, `height` and `opaque` out differently.  Every translation unit that includes this header reports its own
// layout when the program starts; after a mismatch VideoProcessor::Init refuses to run (VREADER_ERROR) instead of
// writing output->opaque at the wrong offset.  Rebuild the library with the same headers as the caller
// (python -m tensor_stream_b200.build, TSVPP_FFMPEG_INCLUDE=<dir>), see INTEGRATION.md.
extern "C" int tsvpp_vp_register_avframe_layout(size_t off_linesize, size_t off_width, size_t off_height, size_t off_opaque,
                                                size_t size);
extern "C" const char *tsvpp_vp_avframe_flavour(void);  // "ffmpeg" or "shim": what the library was built against
namespace tsvpp_detail {
struct AvFrameLayoutCheck {
    AvFrameLayoutCheck() {
        tsvpp_vp_register_avframe_layout(offsetof(AVFrame, linesize), offsetof(AVFrame, width), offsetof(AVFrame, height),
                                         offsetof(AVFrame, opaque), sizeof(AVFrame));
    }
};
static AvFrameLayoutCheck avframe_layout_check_instance;
}  // namespace tsvpp_detail

class VideoProcessor {
public:
    int Init(std::shared_ptr<Logger> logger, uint8_t maxConsumers = 5, bool _enableDumps = false);
    // input: NV12 in device memory (data[0]=Y, data[1]=UV, linesize[0..1], width, height).
    // output->opaque receives a fresh device buffer (stream-ordered allocation, cudaFree-able
    // by the caller exactly like the reference's), output->width/height the final size.
    int Convert(AVFrame *input, AVFrame *output, FrameParameters &options, std::string consumerName);
    // --- extensions (not in the reference) for pipelines beyond ~1e5 frames/s ---------------------------------
    // Convert() hands out a fresh allocation per frame which the caller releases with cudaFree(): two driver calls
    // per frame that serialise across threads.  ConvertInto() writes into a buffer the CALLER owns (a tensor from
    // torch's caching allocator in the ATen hand-off, a recycled pool in tests/cpp/ring_main.cpp): `dst` must hold
    // channelsByFourCC * w * h elements and be usable on CoalescingStream().  Same dispatch, same bits, same
    // coalescing; no CUDA call on the caller's thread.  Returns like Convert(); output->opaque = dst.
    int ConvertInto(AVFrame *input, AVFrame *output, void *dst, FrameParameters &options, std::string consumerName);
    // the stream coalesced conversions run on (blocking stream), and a wait for everything issued on it so far
    cudaStream_t CoalescingStream() const { return coStream; }
    int WaitCoalesced();
    template <class T>
    int DumpFrame(T *output, FrameParameters options, std::shared_ptr<FILE> dumpFile);
    void Close();
    ~VideoProcessor();

private:
    bool enableDumps = false;
    int device = 0;
    TsvppContext *ctx = nullptr;
    std::vector<std::pair<std::string, cudaStream_t>> streamArr;  // one stream per consumer
    std::mutex streamSync;
    std::mutex dumpSync;
    bool isClosed = true;
    std::shared_ptr<Logger> logger;
    // Frames of consumers that call Convert at the same time are converted by one fused launch
    // (tsvpp_convert_coalesced); TSVPP_COALESCE=0 at Init keeps one launch per call on the consumer's own stream.
    bool coalesce = true;
    cudaStream_t coStream = nullptr;
    int convertImpl(AVFrame *input, AVFrame *output, void *callerDst, FrameParameters &options, const std::string &consumerName);
};
