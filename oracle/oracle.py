"""ctypes front-end of the CPU oracle (oracle/vpp_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of vpp_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg; never by tensor_stream_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

Y800, RGB24, BGR24, NV12, UYVY, YUV444, HSV = range(7)
PLANAR, MERGED = 0, 1
NEAREST, BILINEAR, BICUBIC, AREA = range(4)


class Params(C.Structure):
    _fields_ = [("crop_x0", C.c_int), ("crop_y0", C.c_int), ("crop_x1", C.c_int), ("crop_y1", C.c_int),
                ("dst_width", C.c_int), ("dst_height", C.c_int), ("resize_type", C.c_int),
                ("fourcc", C.c_int), ("planes", C.c_int), ("normalization", C.c_int)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "vpp_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.vpo_convert.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(Params), C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.vpo_convert.restype = C.c_int
        L.vpo_channels.argtypes = [C.c_int]
        L.vpo_channels.restype = C.c_float
        L.vpo_yuv_to_rgb.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.vpo_rgb_to_hsv.argtypes = [C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float)]
        L.vpo_bicubic_weights.argtypes = [C.c_double, C.POINTER(C.c_double)]
        L.vpo_set_bicubic_fma.argtypes = [C.c_int]
        L.vpo_area_pattern.argtypes = [C.c_float, C.c_void_p, C.c_int]
        L.vpo_area_pattern.restype = C.c_int
        L.vpo_oob_reads.restype = C.c_long
        L.vpo_pow_table_set.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.vpo_pow_table_set.restype = C.c_int
        L.vpo_pow_record.argtypes = [C.c_int]
        L.vpo_pow_recorded.argtypes = [C.c_void_p, C.c_int]
        L.vpo_pow_recorded.restype = C.c_int
        L.vpo_av_crc32.argtypes = [C.c_uint32, C.c_char_p, C.c_size_t]
        L.vpo_av_crc32.restype = C.c_uint32
        _LIB = L
    return _LIB


def channels(fourcc):
    return float(lib().vpo_channels(int(fourcc)))


def out_geometry(width, height, crop=(0, 0, 0, 0), dst=(0, 0)):
    """Final (w, h) after the dispatch rules of reference src/VideoProcessor.cpp:106-135."""
    w, h = width, height
    cw, ch = crop[2] - crop[0], crop[3] - crop[1]
    if cw > 0 and ch > 0 and cw < width and ch < height:
        w, h = cw, ch
    if dst[0] and dst[1]:
        w, h = dst
    return w, h


def convert(y, uv, width=None, height=None, crop=(0, 0, 0, 0), dst=(0, 0), resize_type=NEAREST,
            fourcc=RGB24, planes=MERGED, normalization=False, pitch_y=None, pitch_uv=None,
            return_oob=False):
    """Run the whole chain on one frame.  `y` is (H, pitchY) uint8, `uv` is (H/2, pitchUV) uint8
    (C-contiguous).  Returns a flat numpy array (float32 if normalization or HSV else uint8) and
    the output (w, h)."""
    L = lib()
    y = np.ascontiguousarray(y, dtype=np.uint8)
    uv = np.ascontiguousarray(uv, dtype=np.uint8)
    if height is None:
        height = y.shape[0]
    if pitch_y is None:
        pitch_y = y.shape[1] if y.ndim == 2 else width
    if pitch_uv is None:
        pitch_uv = uv.shape[1] if uv.ndim == 2 else width
    if width is None:
        width = pitch_y
    p = Params(crop[0], crop[1], crop[2], crop[3], dst[0], dst[1], int(resize_type), int(fourcc),
               int(planes), 1 if normalization else 0)
    ow, oh = out_geometry(width, height, crop, dst)
    n = int(channels(fourcc) * ow * oh)
    is_float = bool(normalization) or fourcc == HSV
    out = np.zeros(n, dtype=np.float32 if is_float else np.uint8)
    cw, chh = C.c_int(0), C.c_int(0)
    L.vpo_reset_oob()
    rc = L.vpo_convert(y.ctypes.data, uv.ctypes.data, pitch_y, pitch_uv, width, height, C.byref(p),
                       out.ctypes.data, C.byref(cw), C.byref(chh))
    if rc != 0:
        raise RuntimeError("vpo_convert failed: %d" % rc)
    assert (cw.value, chh.value) == (ow, oh)
    if return_oob:
        return out, (ow, oh), int(L.vpo_oob_reads())
    return out, (ow, oh)


def av_crc32(buf, init=0xFFFFFFFF):
    b = bytes(buf)
    return int(lib().vpo_av_crc32(init, b, len(b)))


def yuv_to_rgb(Y, U, V):
    a = (C.c_int * 3)()
    lib().vpo_yuv_to_rgb(Y, U, V, a)
    return tuple(a)


def rgb_to_hsv(r, g, b):
    a = (C.c_float * 3)()
    lib().vpo_rgb_to_hsv(r, g, b, a)
    return tuple(a)


def bicubic_weights(w):
    a = (C.c_double * 4)()
    lib().vpo_bicubic_weights(w, a)
    return tuple(a)


def area_pattern(scale, max_rows=8192):
    import math
    stride = int(math.ceil(np.float32(scale)))
    buf = np.zeros((max_rows, stride), dtype=np.float32)
    n = lib().vpo_area_pattern(C.c_float(scale), buf.ctypes.data, max_rows)
    if n <= 0:
        raise RuntimeError("pattern generation failed: %d" % n)
    return buf[:n].copy()


def pow_record(on):
    lib().vpo_pow_record(1 if on else 0)


def pow_recorded():
    """Distinct-ish list of the w arguments bicubic asked pow() for since pow_record(True)."""
    buf = np.zeros(16384, dtype=np.float64)
    n = lib().vpo_pow_recorded(buf.ctypes.data, buf.size)
    return np.unique(buf[:min(n, buf.size)])


def pow_table_set(w, w2, w3):
    w = np.ascontiguousarray(w, dtype=np.float64)
    w2 = np.ascontiguousarray(w2, dtype=np.float64)
    w3 = np.ascontiguousarray(w3, dtype=np.float64)
    rc = lib().vpo_pow_table_set(w.ctypes.data, w2.ctypes.data, w3.ctypes.data, int(w.size))
    if rc:
        raise RuntimeError("pow table overflow")


def pow_table_clear():
    lib().vpo_pow_table_clear()
