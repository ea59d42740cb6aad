"""Python face of the native VPP library (ctypes over include/tsvpp.h).

Mirrors the reference's Python-visible types for this path -- enums FourCC / Planes /
ResizeType (reference tensor_stream/tensor_stream.py:47-84) and FrameParameters
(:100-141) -- and exposes VideoProcessor.convert() on torch CUDA tensors.  PyTorch is used for
device memory and streams only; the arithmetic is in csrc/vpp_kernels.cu.  There is no CPU
fallback: a missing library or a non-CUDA tensor is an error.
"""
import ctypes as C
import os
from enum import Enum

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libtsvpp.so")
_lib = None


class FourCC(Enum):  # reference tensor_stream/tensor_stream.py:47-62
    Y800 = 0
    RGB24 = 1
    BGR24 = 2
    NV12 = 3
    UYVY = 4
    YUV444 = 5
    HSV = 6


class ResizeType(Enum):  # :66-75
    NEAREST = 0
    BILINEAR = 1
    BICUBIC = 2
    AREA = 3


class Planes(Enum):  # :79-83
    PLANAR = 0
    MERGED = 1


class TsvppFrame(C.Structure):
    _fields_ = [("y", C.c_void_p), ("uv", C.c_void_p), ("pitch_y", C.c_int), ("pitch_uv", C.c_int),
                ("width", C.c_int), ("height", C.c_int)]


class TsvppParams(C.Structure):
    _fields_ = [("crop_x0", C.c_int), ("crop_y0", C.c_int), ("crop_x1", C.c_int), ("crop_y1", C.c_int),
                ("dst_width", C.c_int), ("dst_height", C.c_int), ("resize_type", C.c_int),
                ("fourcc", C.c_int), ("planes", C.c_int), ("normalization", C.c_int)]


class VppError(RuntimeError):
    """Raised with the integer status as message, like the reference's
    CHECK_STATUS_THROW -> std::runtime_error(std::to_string(status)) (include/Common.h:116-123)."""


def load_library():
    """dlopen tensor_stream_b200/libtsvpp.so.  Fails loudly when it has not been built
    (python -m tensor_stream_b200.build); nothing else can run the hot path."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("TSVPP_LIB", _LIBPATH)  # developer override (A/B builds)
    if not os.path.exists(path):
        raise ImportError("native library %s is missing: run `python -m tensor_stream_b200.build` "
                          "(there is no CPU fallback)" % path)
    L = C.CDLL(path)
    vp = C.c_void_p
    L.tsvpp_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.tsvpp_create.restype = C.c_int
    L.tsvpp_destroy.argtypes = [vp]
    L.tsvpp_destroy.restype = None
    L.tsvpp_output_info.argtypes = [C.c_int, C.c_int, C.POINTER(TsvppParams), C.POINTER(C.c_int),
                                    C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
    L.tsvpp_output_info.restype = C.c_int
    L.tsvpp_convert_batch.argtypes = [vp, C.POINTER(TsvppFrame), C.c_int, C.POINTER(TsvppParams),
                                      C.POINTER(vp), vp]
    L.tsvpp_convert_batch.restype = C.c_int
    L.tsvpp_batch_create.argtypes = [vp, C.POINTER(TsvppFrame), C.c_int, C.POINTER(TsvppParams),
                                     C.POINTER(vp), C.POINTER(vp)]
    L.tsvpp_batch_create.restype = C.c_int
    L.tsvpp_batch_run.argtypes = [vp, vp]
    L.tsvpp_batch_run.restype = C.c_int
    L.tsvpp_batch_destroy.argtypes = [vp]
    L.tsvpp_batch_destroy.restype = None
    L.tsvpp_convert_batch_host.argtypes = [vp, C.POINTER(TsvppFrame), C.c_int, C.POINTER(TsvppParams),
                                           C.POINTER(vp)]
    L.tsvpp_convert_batch_host.restype = C.c_int
    L.tsvpp_convert_coalesced.argtypes = [vp, C.POINTER(TsvppFrame), C.POINTER(TsvppParams), vp, C.c_int]
    L.tsvpp_convert_coalesced.restype = C.c_int
    L.tsvpp_coalesce_stream.argtypes = [vp]
    L.tsvpp_coalesce_stream.restype = vp
    L.tsvpp_coalesce_stats.argtypes = [vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    L.tsvpp_coalesce_stats.restype = C.c_int
    L.tsvpp_plan_cache_size.argtypes = [vp]
    L.tsvpp_plan_cache_size.restype = C.c_int
    L.tsvpp_last_host_path.argtypes = [vp]
    L.tsvpp_last_host_path.restype = C.c_int
    L.tsvpp_launch_count.argtypes = [vp]
    L.tsvpp_launch_count.restype = C.c_ulonglong
    L.tsvpp_describe_plan.argtypes = [vp, C.POINTER(TsvppFrame), C.POINTER(TsvppParams), C.c_char_p, C.c_size_t]
    L.tsvpp_describe_plan.restype = C.c_int
    L.tsvpp_debug_pow23.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.tsvpp_debug_pow23.restype = C.c_int
    L.tsvpp_last_error.argtypes = []
    L.tsvpp_last_error.restype = C.c_char_p
    L.tsvpp_version.argtypes = []
    L.tsvpp_version.restype = C.c_char_p
    _lib = L
    return L


def _enum_val(v):
    return int(v.value) if isinstance(v, Enum) else int(v)


class FrameParameters:
    """Same constructor as the reference's Python FrameParameters
    (tensor_stream/tensor_stream.py:100-141); `.c` is the flattened TsvppParams."""

    def __init__(self, width=0, height=0, crop_coords=(0, 0, 0, 0), resize_type=ResizeType.NEAREST,
                 pixel_format=FourCC.RGB24, planes_pos=Planes.MERGED, normalization=None):
        fourcc = _enum_val(pixel_format)
        if normalization is None:
            # ColorOptions ctor: normalization=false unless HSV (reference include/VideoProcessor.h:40-47)
            normalization = fourcc == FourCC.HSV.value
        self.width, self.height = int(width), int(height)
        self.crop_coords = tuple(int(v) for v in crop_coords)
        self.resize_type = ResizeType(_enum_val(resize_type))
        self.pixel_format = FourCC(fourcc)
        self.planes_pos = Planes(_enum_val(planes_pos))
        self.normalization = bool(normalization)

    @property
    def c(self):
        cc = self.crop_coords
        return TsvppParams(cc[0], cc[1], cc[2], cc[3], self.width, self.height, self.resize_type.value,
                           self.pixel_format.value, self.planes_pos.value, 1 if self.normalization else 0)


def _check(rc):
    if rc != 0:
        e = VppError(str(rc))  # message is the bare status, like the reference's runtime_error
        try:
            e.detail = load_library().tsvpp_last_error().decode()
        except Exception:
            e.detail = ""
        if e.detail:
            e.add_note(e.detail)
        raise e


def device_pow23(w):
    """pow(w,2), pow(w,3) evaluated on the current CUDA device (test hook)."""
    import numpy as np
    L = load_library()
    w = np.ascontiguousarray(w, dtype=np.float64)
    w2, w3 = np.empty_like(w), np.empty_like(w)
    _check(L.tsvpp_debug_pow23(w.ctypes.data, int(w.size), w2.ctypes.data, w3.ctypes.data))
    return w2, w3


def output_info(src_w, src_h, params):
    """(w, h, elems, elem_size) after the dispatch rules of reference VideoProcessor.cpp:106-135."""
    L = load_library()
    p = params.c if isinstance(params, FrameParameters) else params
    w, h, n, es = C.c_int(), C.c_int(), C.c_size_t(), C.c_int()
    _check(L.tsvpp_output_info(src_w, src_h, C.byref(p), C.byref(w), C.byref(h), C.byref(n), C.byref(es)))
    return w.value, h.value, n.value, es.value


def output_shape(src_w, src_h, params):
    """Tensor shape by FourCC, as the reference's from_blob calls (src/Wrappers/WrapperPython.cpp:318-341)."""
    w, h, n, es = output_info(src_w, src_h, params)
    f = params.pixel_format
    if f in (FourCC.RGB24, FourCC.BGR24):
        shape = (h, w, 3) if params.planes_pos == Planes.MERGED else (3, h, w)
    elif f in (FourCC.YUV444, FourCC.HSV):
        shape = (h, w, 3)
    else:  # UYVY / NV12 / Y800
        ch = {FourCC.UYVY: 2.0, FourCC.NV12: 1.5, FourCC.Y800: 1.0}[f]
        shape = (1, int(h * ch), w)
    return shape, es


class VideoProcessor:
    """Python mirror of the reference's VideoProcessor (include/VideoProcessor.h:118-149) for the
    hot path: Init -> __init__, Convert -> convert / convert_batch, Close -> close."""

    def __init__(self, device=0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("tensor_stream_b200 needs a CUDA device (no CPU fallback)")
        self._torch = torch
        self._L = load_library()
        self.device = int(device)
        h = C.c_void_p()
        _check(self._L.tsvpp_create(self.device, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.tsvpp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -------------------------------------------------------------------------
    def _frame_struct(self, y, uv, width=None, height=None):
        torch = self._torch
        assert y.is_cuda and uv.is_cuda and y.dtype == torch.uint8 and uv.dtype == torch.uint8
        assert y.dim() == 2 and uv.dim() == 2 and y.stride(1) == 1 and uv.stride(1) == 1
        h = int(height if height is not None else y.shape[0])
        w = int(width if width is not None else y.shape[1])
        return TsvppFrame(y.data_ptr(), uv.data_ptr(), int(y.stride(0)), int(uv.stride(0)), w, h)

    def describe_plan(self, y, uv, params, width=None, height=None):
        f = self._frame_struct(y, uv, width, height)
        buf = C.create_string_buffer(1024)
        p = params.c
        _check(self._L.tsvpp_describe_plan(self._h, C.byref(f), C.byref(p), buf, 1024))
        return buf.value.decode()

    def last_host_path(self):
        """How the last convert_batch_host call moved data: inputs read in place ("zero-copy") or
        staged with H2D copies; "+out" when the outputs were written straight to pinned host memory."""
        m = self._L.tsvpp_last_host_path(self._h)
        return ("zero-copy" if m & 1 else "staged-h2d") + ("+out" if m & 2 else "")

    def launch_count(self):
        return int(self._L.tsvpp_launch_count(self._h))

    def plan_cache_size(self):
        return int(self._L.tsvpp_plan_cache_size(self._h))

    def coalesce_stats(self):
        """(frames, launches) of convert_coalesced so far."""
        f, l = C.c_ulonglong(), C.c_ulonglong()
        _check(self._L.tsvpp_coalesce_stats(self._h, C.byref(f), C.byref(l)))
        return int(f.value), int(l.value)

    def convert_coalesced(self, y, uv, params, out, width=None, height=None, wait=True):
        """One frame per call, callable from many threads at once (the reference's getFrame pattern): calls
        pending at the same time share ONE fused launch on the context's coalescing stream.  `out` is a
        preallocated CUDA tensor; with wait=False the call returns once the covering launch is enqueued."""
        f = self._frame_struct(y, uv, width, height)
        p = params.c
        _check(self._L.tsvpp_convert_coalesced(self._h, C.byref(f), C.byref(p), C.c_void_p(out.data_ptr()), 1 if wait else 0))
        return out

    # -- the hot path ---------------------------------------------------------------------
    def convert_batch(self, frames, params, out=None, widths=None, stream=None):
        """frames: list of (Y, UV) CUDA uint8 tensors (Y: [H, pitch], UV: [H/2, pitch]; the
        logical width may be smaller than the pitch -> pass `widths`=(w, h)).  Returns a list of
        output tensors (allocated from torch's caching allocator unless `out` is given).
        One fused kernel launch for the whole batch, on torch's current stream."""
        torch = self._torch
        n = len(frames)
        w, h = widths if widths is not None else (None, None)
        fs = (TsvppFrame * n)()
        for i, (y, uv) in enumerate(frames):
            fs[i] = self._frame_struct(y, uv, w, h)
        shape, es = output_shape(fs[0].width, fs[0].height, params)
        dtype = torch.float32 if es == 4 else torch.uint8
        if out is None:
            dev = frames[0][0].device
            out = [torch.empty(shape, dtype=dtype, device=dev) for _ in range(n)]
        outs = (C.c_void_p * n)(*[o.data_ptr() for o in out])
        p = params.c
        s = stream if stream is not None else torch.cuda.current_stream(frames[0][0].device)
        _check(self._L.tsvpp_convert_batch(self._h, fs, n, C.byref(p), outs, C.c_void_p(s.cuda_stream)))
        return out

    def convert(self, y, uv, params, width=None, height=None, out=None):
        """One frame (the reference's VideoProcessor::Convert, src/VideoProcessor.cpp:94-166)."""
        o = self.convert_batch([(y, uv)], params, None if out is None else [out],
                               None if width is None else (width, height))
        return o[0]

    def make_batch(self, frames, params, out, widths=None):
        """Pre-uploaded batch (descriptor ring of a streaming pipeline); returns a callable
        run(stream=None)."""
        torch = self._torch
        n = len(frames)
        w, h = widths if widths is not None else (None, None)
        fs = (TsvppFrame * n)()
        for i, (y, uv) in enumerate(frames):
            fs[i] = self._frame_struct(y, uv, w, h)
        outs = (C.c_void_p * n)(*[o.data_ptr() for o in out])
        p = params.c
        b = C.c_void_p()
        _check(self._L.tsvpp_batch_create(self._h, fs, n, C.byref(p), outs, C.byref(b)))
        L = self._L
        keep = (frames, out)

        class _Batch:
            def __init__(self):
                self._b = b
                self._keep = keep

            def run(self, stream=None):
                s = stream if stream is not None else torch.cuda.current_stream()
                _check(L.tsvpp_batch_run(self._b, C.c_void_p(s.cuda_stream)))

            def destroy(self):
                if self._b:
                    L.tsvpp_batch_destroy(self._b)
                    self._b = None

        return _Batch()

    def convert_batch_host(self, frames_np, params, width, height, out_np=None):
        """End-to-end path with HOST buffers (numpy arrays or pinned torch CPU tensors exposing
        ctypes pointers): H2D + fused kernel + D2H inside the call."""
        import numpy as np
        n = len(frames_np)
        fs = (TsvppFrame * n)()
        for i, (y, uv) in enumerate(frames_np):
            fs[i] = TsvppFrame(_host_ptr(y), _host_ptr(uv), _host_pitch(y), _host_pitch(uv), width, height)
        shape, es = output_shape(width, height, params)
        if out_np is None:
            out_np = [np.empty(shape, dtype=np.float32 if es == 4 else np.uint8) for _ in range(n)]
        outs = (C.c_void_p * n)(*[_host_ptr(o) for o in out_np])
        p = params.c
        _check(self._L.tsvpp_convert_batch_host(self._h, fs, n, C.byref(p), outs))
        return out_np


def convert_batch_aten(frames, params, widths=None):
    """Same result as VideoProcessor.convert_batch, through the PyTorch C++ extension
    (csrc/aten_handoff.cpp): outputs are at::empty tensors from torch's caching allocator, the
    launch goes to torch's current stream, no Python-side pointer juggling."""
    from . import aten_build
    mod = aten_build.load()
    w, h = widths if widths is not None else (0, 0)
    return mod.convert_batch([f[0] for f in frames], [f[1] for f in frames], w, h, list(params.crop_coords),
                             params.width, params.height, params.resize_type.value, params.pixel_format.value,
                             params.planes_pos.value, params.normalization)


def convert_frame_aten(y, uv, params, width=0, height=0, wait=True):
    """One frame per call through the PyTorch C++ extension, callable from many threads (the GIL is released):
    the result is an at::empty tensor owned by Python (no use_count() garbage collector, no cudaFree), calls in
    flight together share one fused launch.  This is the wrapper-shaped replacement of the reference's
    TensorStream.read() hand-off (src/Wrappers/WrapperPython.cpp:265-350)."""
    from . import aten_build
    mod = aten_build.load()
    return mod.convert_frame(y, uv, int(width or 0), int(height or 0), list(params.crop_coords), params.width, params.height,
                             params.resize_type.value, params.pixel_format.value, params.planes_pos.value,
                             params.normalization, bool(wait))


def _host_ptr(a):
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data


def _host_pitch(a):
    if hasattr(a, "stride") and callable(a.stride):
        return int(a.stride(0))
    return int(a.strides[0])
