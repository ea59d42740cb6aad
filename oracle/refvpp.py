"""ctypes front-end of oracle/_ref/libtsref_vpp.so: the reference's OWN, unmodified VPP sources
(src/Crop.cu, Resize.cu, ColorConversion.cu, VideoProcessor.cpp, Common.cpp) compiled in place
for sm_100a and driven through VideoProcessor::Convert by oracle/ref_driver.cu.

TEST / BASELINE INFRASTRUCTURE ONLY (the GPU oracle of record and the "reference CUDA VPP on one
B200" baseline of bench.py --impl reference).  Needs a GPU; the .so is prebuilt in the build
container (oracle/Makefile) and travels with the gpurun snapshot.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libtsref_vpp.so")
_lib = None


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise ImportError("%s missing: run `make -C oracle ref` where /root/reference is mounted" % _SO)
        L = C.CDLL(_SO)
        L.tsref_create.argtypes = [C.c_int]
        L.tsref_create.restype = C.c_void_p
        L.tsref_destroy.argtypes = [C.c_void_p]
        L.tsref_convert.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 14 + \
                                   [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.tsref_convert.restype = C.c_int
        L.tsref_free.argtypes = [C.c_void_p]
        L.tsref_free.restype = C.c_int
        L.tsref_sync.restype = C.c_int
        L.tsref_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.tsref_memcpy_d2h.restype = C.c_int
        _lib = L
    return _lib


class RefVPP:
    """One reference VideoProcessor instance (Init with maxConsumers streams)."""

    def __init__(self, max_consumers=5):
        self.L = lib()
        self.h = self.L.tsref_create(max_consumers)
        if not self.h:
            raise RuntimeError("reference VideoProcessor::Init failed")

    def close(self):
        if self.h:
            self.L.tsref_destroy(self.h)
            self.h = None

    def convert_raw(self, y_ptr, uv_ptr, pitch_y, pitch_uv, w, h, crop=(0, 0, 0, 0), dst=(0, 0), resize_type=0,
                    fourcc=1, planes=1, normalization=False, consumer=b"bench"):
        """Calls the reference Convert; returns (device pointer, w, h).  Caller frees with free()."""
        out = C.c_void_p()
        ow, oh = C.c_int(), C.c_int()
        rc = self.L.tsref_convert(self.h, y_ptr, uv_ptr, pitch_y, pitch_uv, w, h, crop[0], crop[1], crop[2],
                                  crop[3], dst[0], dst[1], int(resize_type), int(fourcc), int(planes),
                                  1 if normalization else 0, consumer, C.byref(out), C.byref(ow), C.byref(oh))
        if rc != 0:
            raise RuntimeError("reference Convert returned %d" % rc)
        return out.value, ow.value, oh.value

    def free(self, ptr):
        self.L.tsref_free(C.c_void_p(ptr))

    def sync(self):
        self.L.tsref_sync()

    def convert(self, y, uv, width=None, height=None, **kw):
        """torch CUDA uint8 tensors in -> flat numpy array out (copied out of the reference's
        buffer, which is then cudaFree'd like a C++ user of the reference would)."""
        import numpy as np
        import torch
        from . import oracle as O
        h = int(height if height is not None else y.shape[0])
        w = int(width if width is not None else y.shape[1])
        fourcc = kw.get("fourcc", 1)
        norm = bool(kw.get("normalization", False))
        torch.cuda.synchronize()
        ptr, ow, oh = self.convert_raw(y.data_ptr(), uv.data_ptr(), int(y.stride(0)), int(uv.stride(0)), w, h, **kw)
        self.sync()
        n = int(O.channels(fourcc) * ow * oh)
        is_float = norm or fourcc == 6
        out = np.empty(n, dtype=np.float32 if is_float else np.uint8)
        rc = self.L.tsref_memcpy_d2h(out.ctypes.data, ptr, out.nbytes)
        assert rc == 0, rc
        self.free(ptr)
        return out, (ow, oh)
