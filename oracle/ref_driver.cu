/* ref_driver.cu -- OUR extern "C" driver around the reference's unmodified
 * VideoProcessor::Convert (reference src/VideoProcessor.cpp:94-166), so that tests and
 * bench.py --impl reference can call the reference VPP through ctypes.  Test / baseline
 * infrastructure; linked only into oracle/_ref/libtsref_vpp.so. */
#include <cuda_runtime.h>

#include <memory>
#include <string>
#include <tuple>

#include "VideoProcessor.h"

struct TsrefCtx {
    VideoProcessor vpp;
    std::shared_ptr<Logger> logger;
};

extern "C" {

void *tsref_create(int maxConsumers) {
    auto *c = new TsrefCtx();
    c->logger = std::make_shared<Logger>();
    if (c->vpp.Init(c->logger, (uint8_t)maxConsumers, false) != 0) {
        delete c;
        return nullptr;
    }
    return c;
}

void tsref_destroy(void *h) {
    auto *c = (TsrefCtx *)h;
    if (!c) return;
    c->vpp.Close();
    delete c;
}

/* One frame through the reference's public entry point.  *out_dev receives the device buffer
 * the reference cudaMalloc'ed (caller releases it with tsref_free, as a C++ user of the
 * reference would with cudaFree: c_examples/src/Sample.cpp:27,36). */
int tsref_convert(void *h, const uint8_t *dY, const uint8_t *dUV, int pitchY, int pitchUV, int width,
                  int height, int crop_x0, int crop_y0, int crop_x1, int crop_y1, int dst_w, int dst_h,
                  int resize_type, int fourcc, int planes, int normalization, const char *consumer,
                  void **out_dev, int *out_w, int *out_h) {
    auto *c = (TsrefCtx *)h;
    AVFrame in = {}, out = {};
    in.data[0] = const_cast<uint8_t *>(dY);
    in.data[1] = const_cast<uint8_t *>(dUV);
    in.linesize[0] = pitchY;
    in.linesize[1] = pitchUV;
    in.width = width;
    in.height = height;
    ColorOptions color((FourCC)fourcc);
    color.planesPos = (Planes)planes;
    color.normalization = normalization != 0;
    ResizeOptions resize(dst_w, dst_h);
    resize.type = (ResizeType)resize_type;
    CropOptions crop(std::make_tuple(crop_x0, crop_y0), std::make_tuple(crop_x1, crop_y1));
    FrameParameters fp(resize, color, crop);
    int sts = c->vpp.Convert(&in, &out, fp, std::string(consumer ? consumer : "bench"));
    *out_dev = out.opaque;
    *out_w = out.width;
    *out_h = out.height;
    return sts;
}

int tsref_free(void *p) { return (int)cudaFree(p); }
int tsref_memcpy_d2h(void *host, const void *dev, size_t n) { return (int)cudaMemcpy(host, dev, n, cudaMemcpyDeviceToHost); }
int tsref_sync(void) { return (int)cudaDeviceSynchronize(); }

}  // extern "C"
