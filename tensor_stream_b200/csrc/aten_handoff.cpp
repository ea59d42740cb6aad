// aten_handoff.cpp -- the ATen hand-off of the VPP path as a PyTorch C++ extension.
//
// Replaces reference src/Wrappers/WrapperPython.cpp:317-350 (torch::from_blob over a cudaMalloc'ed
// buffer + the use_count()-polling garbage collector at :173-184): the fused kernel writes
// straight into tensors that come from torch's caching allocator, on torch's current stream.
// This is the only piece of the path that needs libtorch; the arithmetic lives behind the C ABI
// (include/tsvpp.h, libtsvpp.so).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "tsvpp.h"

namespace {

std::mutex g_mu;
std::map<int, TsvppContext *> g_ctx;  // one context per device, created on first use

TsvppContext *context_for(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_ctx.find(device);
    if (it != g_ctx.end()) return it->second;
    TsvppContext *c = nullptr;
    const int rc = tsvpp_create(device, &c);
    if (rc != 0) throw std::runtime_error(std::to_string(rc));  // like CHECK_STATUS_THROW (Common.h:116-123)
    g_ctx[device] = c;
    return c;
}

// tensor shape by FourCC -- reference src/Wrappers/WrapperPython.cpp:318-341
std::vector<int64_t> shape_by_fourcc(const TsvppParams &p, int w, int h) {
    switch (p.fourcc) {
        case TSVPP_RGB24:
        case TSVPP_BGR24:
            if (p.planes == TSVPP_MERGED) return {h, w, 3};
            return {3, h, w};
        case TSVPP_YUV444:
        case TSVPP_HSV: return {h, w, 3};
        case TSVPP_UYVY: return {1, (int64_t)(h * 2), w};
        case TSVPP_NV12: return {1, (int64_t)(h * 1.5), w};
        default: return {1, h, w};  // Y800
    }
}

}  // namespace

// frames: list of (Y, UV) CUDA uint8 tensors (2-D, unit inner stride; the row stride is the pitch).
// Returns one tensor per frame, allocated by torch, filled by ONE fused launch on the current stream.
std::vector<at::Tensor> convert_batch(const std::vector<at::Tensor> &ys, const std::vector<at::Tensor> &uvs,
                                      int64_t width, int64_t height, std::vector<int64_t> crop, int64_t dst_w,
                                      int64_t dst_h, int64_t resize_type, int64_t fourcc, int64_t planes,
                                      bool normalization) {
    TORCH_CHECK(!ys.empty() && ys.size() == uvs.size(), "need as many Y planes as UV planes");
    TORCH_CHECK(crop.size() == 4, "crop = (x0, y0, x1, y1)");
    const auto dev = ys[0].device();
    TORCH_CHECK(dev.is_cuda(), "tensor_stream_b200 has no CPU path: frames must be CUDA tensors");
    c10::cuda::CUDAGuard guard(dev);
    const int n = (int)ys.size();
    std::vector<TsvppFrame> f(n);
    for (int i = 0; i < n; ++i) {
        const at::Tensor &y = ys[i], &uv = uvs[i];
        TORCH_CHECK(y.device() == dev && uv.device() == dev && y.scalar_type() == at::kByte && uv.scalar_type() == at::kByte &&
                        y.dim() == 2 && uv.dim() == 2 && y.stride(1) == 1 && uv.stride(1) == 1,
                    "frames must be 2-D uint8 CUDA tensors with unit inner stride");
        f[i].y = y.data_ptr<uint8_t>();
        f[i].uv = uv.data_ptr<uint8_t>();
        f[i].pitch_y = (int)y.stride(0);
        f[i].pitch_uv = (int)uv.stride(0);
        f[i].width = (int)(width > 0 ? width : y.size(1));
        f[i].height = (int)(height > 0 ? height : y.size(0));
    }
    TsvppParams p;
    p.crop_x0 = (int)crop[0]; p.crop_y0 = (int)crop[1]; p.crop_x1 = (int)crop[2]; p.crop_y1 = (int)crop[3];
    p.dst_width = (int)dst_w; p.dst_height = (int)dst_h;
    p.resize_type = (int)resize_type; p.fourcc = (int)fourcc; p.planes = (int)planes;
    p.normalization = normalization ? 1 : 0;
    int w = 0, h = 0, es = 1;
    size_t elems = 0;
    int rc = tsvpp_output_info(f[0].width, f[0].height, &p, &w, &h, &elems, &es);
    if (rc != 0) throw std::runtime_error(std::to_string(rc));
    const auto opts = at::TensorOptions(es == 4 ? at::kFloat : at::kByte).device(dev);
    std::vector<at::Tensor> outs(n);
    std::vector<void *> optr(n);
    const auto shape = shape_by_fourcc(p, w, h);
    for (int i = 0; i < n; ++i) {
        outs[i] = at::empty(shape, opts);
        optr[i] = outs[i].data_ptr();
    }
    rc = tsvpp_convert_batch(context_for(dev.index()), f.data(), n, &p, optr.data(),
                             at::cuda::getCurrentCUDAStream(dev.index()).stream());
    if (rc != 0) throw std::runtime_error(std::to_string(rc) + std::string(" ") + tsvpp_last_error());
    return outs;
}

// One frame per call, from any number of threads -- TensorStream::getFrame as the reference's Python wrapper runs it
// (src/Wrappers/WrapperPython.cpp:265-350), minus its cudaMalloc per frame and the `tensors` list that polls
// use_count() to cudaFree results (:173-184): the tensor comes from torch's caching allocator and lives exactly as
// long as Python references it.  Calls in flight at the same time share ONE fused launch (tsvpp_convert_coalesced).
// The block is allocated in the pool of the coalescing stream, so the allocator never hands it to another frame
// before the launch that wrote it has run; that stream is a blocking stream, so consumers on torch's default stream
// are ordered after it without further calls (consumers on other streams: Tensor.record_stream).
at::Tensor convert_frame(const at::Tensor &y, const at::Tensor &uv, int64_t width, int64_t height, std::vector<int64_t> crop,
                         int64_t dst_w, int64_t dst_h, int64_t resize_type, int64_t fourcc, int64_t planes, bool normalization,
                         bool wait) {
    TORCH_CHECK(crop.size() == 4, "crop = (x0, y0, x1, y1)");
    const auto dev = y.device();
    TORCH_CHECK(dev.is_cuda() && uv.device() == dev, "tensor_stream_b200 has no CPU path: frames must be CUDA tensors");
    TORCH_CHECK(y.scalar_type() == at::kByte && uv.scalar_type() == at::kByte && y.dim() == 2 && uv.dim() == 2 &&
                    y.stride(1) == 1 && uv.stride(1) == 1,
                "frames must be 2-D uint8 CUDA tensors with unit inner stride");
    c10::cuda::CUDAGuard guard(dev);
    TsvppFrame f;
    f.y = y.data_ptr<uint8_t>();
    f.uv = uv.data_ptr<uint8_t>();
    f.pitch_y = (int)y.stride(0);
    f.pitch_uv = (int)uv.stride(0);
    f.width = (int)(width > 0 ? width : y.size(1));
    f.height = (int)(height > 0 ? height : y.size(0));
    TsvppParams p;
    p.crop_x0 = (int)crop[0]; p.crop_y0 = (int)crop[1]; p.crop_x1 = (int)crop[2]; p.crop_y1 = (int)crop[3];
    p.dst_width = (int)dst_w; p.dst_height = (int)dst_h;
    p.resize_type = (int)resize_type; p.fourcc = (int)fourcc; p.planes = (int)planes;
    p.normalization = normalization ? 1 : 0;
    int w = 0, h = 0, es = 1;
    size_t elems = 0;
    int rc = tsvpp_output_info(f.width, f.height, &p, &w, &h, &elems, &es);
    if (rc != 0) throw std::runtime_error(std::to_string(rc));
    TsvppContext *ctx = context_for(dev.index());
    cudaStream_t cs = (cudaStream_t)tsvpp_coalesce_stream(ctx);
    if (!cs) throw std::runtime_error(std::string("no coalescing stream: ") + tsvpp_last_error());
    at::Tensor out;
    {
        c10::cuda::CUDAStreamGuard sg(c10::cuda::getStreamFromExternal(cs, dev.index()));
        out = at::empty(shape_by_fourcc(p, w, h), at::TensorOptions(es == 4 ? at::kFloat : at::kByte).device(dev));
    }
    rc = tsvpp_convert_coalesced(ctx, &f, &p, out.data_ptr(), wait ? 1 : 0);
    if (rc != 0) throw std::runtime_error(std::to_string(rc) + std::string(" ") + tsvpp_last_error());
    return out;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "ATen hand-off of the B200 VPP path (libtsvpp.so behind it)";
    m.def("convert_batch", &convert_batch, "fused VPP over a batch of NV12 frames -> list of torch tensors",
          py::arg("ys"), py::arg("uvs"), py::arg("width") = 0, py::arg("height") = 0,
          py::arg("crop") = std::vector<int64_t>{0, 0, 0, 0}, py::arg("dst_w") = 0, py::arg("dst_h") = 0,
          py::arg("resize_type") = 0, py::arg("fourcc") = 1, py::arg("planes") = 1, py::arg("normalization") = false,
          py::call_guard<py::gil_scoped_release>());
    m.def("convert_frame", &convert_frame, "one NV12 frame -> torch tensor; concurrent calls share a fused launch",
          py::arg("y"), py::arg("uv"), py::arg("width") = 0, py::arg("height") = 0,
          py::arg("crop") = std::vector<int64_t>{0, 0, 0, 0}, py::arg("dst_w") = 0, py::arg("dst_h") = 0,
          py::arg("resize_type") = 0, py::arg("fourcc") = 1, py::arg("planes") = 1, py::arg("normalization") = false,
          py::arg("wait") = true, py::call_guard<py::gil_scoped_release>());
    m.def("version", []() { return std::string(tsvpp_version()); });
}
