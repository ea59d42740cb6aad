// stages_main.cpp -- the reference's stage-level entry points (include/VideoProcessor.h: cropHost,
// resizeKernel, colorConversionKernel<T>) chained the way the reference's own Convert chains them
// (reference src/VideoProcessor.cpp:106-150), compared byte for byte with the fused Convert() of the
// drop-in.  Compiled against include/VideoProcessor.h and linked with libtsvpp.so by
// tests/test_zz_stage_functions.py.
//
//   stages_main <nv12.raw> <w> <h> <pitch> <fourcc> <planes> <norm> <dstw> <dsth> <resize> <x0> <y0> <x1> <y1>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "VideoProcessor.h"

template <class T>
static int chain(AVFrame *in, AVFrame *out, FrameParameters &o, cudaStream_t *stream, bool *staged) {
    const int cw = std::get<0>(o.crop.rightBottomCorner) - std::get<0>(o.crop.leftTopCorner);
    const int ch = std::get<1>(o.crop.rightBottomCorner) - std::get<1>(o.crop.leftTopCorner);
    bool crop = false, resize = false;
    int sts = 0;
    if (cw > 0 && ch > 0 && cw < in->width && ch < in->height) {
        crop = true;
        if ((sts = cropHost(in, out, o.crop, 1024, stream))) return sts;
        out->width = cw;
        out->height = ch;
    }
    if (o.resize.width && o.resize.height) {
        const AVFrame *cur = crop ? out : in;
        if ((int)o.resize.width != cur->width || (int)o.resize.height != cur->height) resize = true;
        if (resize) {
            if ((sts = resizeKernel(crop ? out : in, out, crop, o.resize, 1024, stream))) return sts;
            out->width = (int)o.resize.width;
            out->height = (int)o.resize.height;
        }
    }
    if (out->width == 0 || out->height == 0) {
        out->width = in->width;
        out->height = in->height;
    }
    sts = colorConversionKernel<T>(resize || crop ? out : in, out, o.color, 1024, stream);
    *staged = resize || crop;
    return sts;
}

int main(int argc, char **argv) {
    if (argc < 15) {
        std::fprintf(stderr, "usage: see source\n");
        return 2;
    }
    const char *path = argv[1];
    const int w = std::atoi(argv[2]), h = std::atoi(argv[3]), pitch = std::atoi(argv[4]);
    ColorOptions color((FourCC)std::atoi(argv[5]));
    color.planesPos = (Planes)std::atoi(argv[6]);
    color.normalization = std::atoi(argv[7]) != 0 || color.dstFourCC == HSV;
    ResizeOptions resize(std::atoi(argv[8]), std::atoi(argv[9]));
    resize.type = (ResizeType)std::atoi(argv[10]);
    CropOptions crop({std::atoi(argv[11]), std::atoi(argv[12])}, {std::atoi(argv[13]), std::atoi(argv[14])});

    std::vector<uint8_t> host((size_t)pitch * h * 3 / 2);
    FILE *f = std::fopen(path, "rb");
    if (!f || std::fread(host.data(), 1, host.size(), f) != host.size()) {
        std::fprintf(stderr, "cannot read %s\n", path);
        return 2;
    }
    std::fclose(f);
    uint8_t *dY = nullptr, *dUV = nullptr;
    cudaMalloc(&dY, (size_t)pitch * h);
    cudaMalloc(&dUV, (size_t)pitch * h / 2);
    cudaMemcpy(dY, host.data(), (size_t)pitch * h, cudaMemcpyHostToDevice);
    cudaMemcpy(dUV, host.data() + (size_t)pitch * h, (size_t)pitch * h / 2, cudaMemcpyHostToDevice);

    auto make_input = [&](AVFrame &a) {
        std::memset(&a, 0, sizeof(a));
        a.data[0] = dY;
        a.data[1] = dUV;
        a.linesize[0] = a.linesize[1] = pitch;
        a.width = w;
        a.height = h;
    };

    // fused path
    VideoProcessor vpp;
    if (vpp.Init(std::make_shared<Logger>(), 2, false)) return 3;
    AVFrame in1, out1;
    make_input(in1);
    std::memset(&out1, 0, sizeof(out1));
    FrameParameters p1 = {resize, color, crop};
    int sts = vpp.Convert(&in1, &out1, p1, "stages");
    std::printf("convert %d\n", sts);
    if (sts) return 4;

    // staged path
    cudaStream_t stream;
    cudaStreamCreate(&stream);
    AVFrame in2, out2;
    make_input(in2);
    std::memset(&out2, 0, sizeof(out2));
    FrameParameters p2 = {resize, color, crop};
    bool staged = false;
    sts = color.normalization ? chain<float>(&in2, &out2, p2, &stream, &staged)
                              : chain<unsigned char>(&in2, &out2, p2, &stream, &staged);
    std::printf("chain %d\n", sts);
    if (sts) return 5;
    cudaStreamSynchronize(stream);
    std::printf("size %dx%d vs %dx%d\n", out1.width, out1.height, out2.width, out2.height);

    const size_t n = (size_t)(channelsByFourCC(color.dstFourCC) * out1.width * out1.height) * (color.normalization ? 4 : 1);
    std::vector<uint8_t> a(n), b(n);
    cudaMemcpy(a.data(), out1.opaque, n, cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), out2.opaque, n, cudaMemcpyDeviceToHost);
    size_t diff = 0;
    for (size_t i = 0; i < n; ++i) diff += a[i] != b[i];
    std::printf("bytes %zu\ndiff %zu\n", n, diff);

    int fr = 0;
    if (staged) {  // what the reference's Convert does with the NV12 intermediates (src/VideoProcessor.cpp:144-150)
        fr |= (int)cudaFree(out2.data[0]);
        fr |= (int)cudaFree(out2.data[1]);
    }
    fr |= (int)cudaFree(out2.opaque);
    fr |= (int)cudaFree(out1.opaque);
    std::printf("free %d\n", fr);

    // pairings Convert never uses
    AVFrame in3, out3;
    make_input(in3);
    std::memset(&out3, 0, sizeof(out3));
    out3.width = w;
    out3.height = h;
    ColorOptions odd(RGB24);
    odd.normalization = false;
    std::printf("float_without_normalization %d\n", colorConversionKernel<float>(&in3, &out3, odd, 1024, &stream));
    ColorOptions unknown((FourCC)42);
    std::printf("unknown_fourcc %d\n", colorConversionKernel<unsigned char>(&in3, &out3, unknown, 1024, &stream));
    cudaStreamDestroy(stream);
    vpp.Close();
    cudaFree(dY);
    cudaFree(dUV);
    return 0;
}
