// dropin_main.cpp -- exercises the C++ drop-in exactly the way the reference's own VPP tests do
// (tests/src/VPPTests.cpp:101-132 fourCCTest, :566-590 getFrame): build an AVFrame by hand around
// NV12 in device memory, call VideoProcessor::Convert, copy the result back, print its av_crc
// (AV_CRC_32_IEEE, init -1) and exercise DumpFrame.  Compiled against include/VideoProcessor.h and
// linked with libtsvpp.so by tests/test_cpp_dropin.py.
//
//   dropin_main <nv12.raw> <w> <h> <pitch> <fourcc> <planes> <norm> <dstw> <dsth> <resize> <x0> <y0> <x1> <y1>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "VideoProcessor.h"

static uint32_t av_crc32_ieee(uint32_t crc, const uint8_t *buf, size_t n) {
    static uint32_t tab[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i << 24;
            for (int k = 0; k < 8; k++) c = (c << 1) ^ ((c & 0x80000000u) ? 0x04C11DB7u : 0);
            tab[i] = ((c >> 24) & 0xFF) | ((c >> 8) & 0xFF00) | ((c << 8) & 0xFF0000) | (c << 24);
        }
        init = true;
    }
    while (n--) crc = tab[((uint8_t)crc) ^ *buf++] ^ (crc >> 8);
    return crc;
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
    color.normalization = std::atoi(argv[7]) != 0;
    ResizeOptions resize(std::atoi(argv[8]), std::atoi(argv[9]));
    resize.type = (ResizeType)std::atoi(argv[10]);
    CropOptions crop({std::atoi(argv[11]), std::atoi(argv[12])}, {std::atoi(argv[13]), std::atoi(argv[14])});
    FrameParameters frameArgs = {resize, color, crop};

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

    VideoProcessor VPP;
    if (VPP.Init(std::make_shared<Logger>()) != 0) return 3;
    AVFrame in = {}, out = {};
    in.data[0] = dY;
    in.data[1] = dUV;
    in.linesize[0] = in.linesize[1] = pitch;
    in.width = w;
    in.height = h;
    const int sts = VPP.Convert(&in, &out, frameArgs, "visualize");
    if (sts != VREADER_OK) {
        std::printf("status %d\n", sts);
        return 4;
    }
    const float channels = channelsByFourCC(color.dstFourCC);
    const size_t elems = (size_t)(channels * out.width * out.height);
    const size_t es = (color.normalization || color.dstFourCC == HSV) ? 4 : 1;
    std::vector<uint8_t> res(elems * es);
    if (cudaMemcpy(res.data(), out.opaque, res.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return 5;
    std::printf("size %d %d\n", out.width, out.height);
    std::printf("crc %u\n", av_crc32_ieee(0xFFFFFFFFu, res.data(), res.size()));
    std::printf("resize_backfilled %u %u\n", frameArgs.resize.width, frameArgs.resize.height);
    // DumpFrame round trip (VPPTests.cpp:114-125)
    {
        std::shared_ptr<FILE> wf(std::fopen("DumpFrame.yuv", "wb"), std::fclose);
        const int d = es == 4 ? VPP.DumpFrame(static_cast<float *>(out.opaque), frameArgs, wf)
                              : VPP.DumpFrame(static_cast<uint8_t *>(out.opaque), frameArgs, wf);
        if (d != VREADER_OK) return 6;
    }
    {
        std::vector<uint8_t> back(res.size());
        std::shared_ptr<FILE> rf(std::fopen("DumpFrame.yuv", "rb"), std::fclose);
        const size_t got = std::fread(back.data(), 1, back.size(), rf.get());
        std::printf("dump_crc %u\n", av_crc32_ieee(0xFFFFFFFFu, back.data(), got));
    }
    std::remove("DumpFrame.yuv");
    // the result is a real CUDA allocation the caller frees (c_examples/src/Sample.cpp:27,36)
    std::printf("free %d\n", (int)cudaFree(out.opaque));
    // sixth consumer name on a 5-stream processor: VREADER_ERROR (VideoProcessor.cpp:101-103)
    int last = 0;
    for (int c = 0; c < 6; ++c) {
        AVFrame o2 = {};
        FrameParameters fa = {resize, color, crop};
        last = VPP.Convert(&in, &o2, fa, std::string("consumer") + std::to_string(c));
        if (last == VREADER_OK) cudaFree(o2.opaque);
    }
    std::printf("sixth_consumer %d\n", last);
    // optional: per-frame Convert rate through the drop-in API (argv[15] = iterations)
    if (argc > 15) {
        const int iters = std::atoi(argv[15]);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int k = 0; k < iters; ++k) {
            AVFrame o2 = {};
            FrameParameters fa = {resize, color, crop};
            if (VPP.Convert(&in, &o2, fa, "visualize") != VREADER_OK) return 7;
            cudaFree(o2.opaque);
        }
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        std::printf("convert_per_frame_fps %.0f\n", iters / (ms * 1e-3));
    }
    VPP.Close();
    cudaFree(dY);
    cudaFree(dUV);
    return 0;
}
