"""Software H.264 decode to NV12 through the FFmpeg 8 shared objects bundled with
opencv-python-headless (no FFmpeg headers / binaries exist in this image).

This is *plumbing around* the VPP hot path, not part of it: the reference obtains NV12
frames from NVDEC (reference src/Decoder.cpp:97-184); H.264 decode is bit-exact, so a
software decode yields the very same NV12 bytes (CRC 2957341121 of
frame 1 of bbb_1080x608_420_10.h264 == reference tests/src/VPPTests.cpp:211-216).

Used by: tests/golden/make_golden.py (fixture generation), the CPU-side tests that pin the
oracle against the reference's golden CRCs, and LocalFileSource in stream.py.
"""
import ctypes as C
import glob
import os

import numpy as np

_LIBS = None

# x86-64 struct offsets for lavf 62 / lavc 62 / lavu 60
_OFF_FMT_STREAMS = 48
_OFF_STREAM_CODECPAR = 16
_OFF_PKT_STREAM_INDEX = 36
_OFF_FRAME_DATA = 0
_OFF_FRAME_LINESIZE = 64
_OFF_FRAME_WIDTH = 104
_OFF_FRAME_HEIGHT = 108
_OFF_FRAME_FORMAT = 116
_AVERROR_EAGAIN = -11
_AVERROR_EOF = -541478725


def _libdir():
    import cv2  # noqa: F401  (makes the bundled libs resolvable)
    site = os.path.dirname(os.path.dirname(cv2.__file__))
    return os.path.join(site, "opencv_python_headless.libs")


def load_ffmpeg():
    """dlopen libavutil/libswresample/libavcodec/libavformat/libswscale (RTLD_GLOBAL, in
    dependency order) and declare the handful of prototypes we call."""
    global _LIBS
    if _LIBS is not None:
        return _LIBS
    d = _libdir()

    def one(pat):
        hits = sorted(glob.glob(os.path.join(d, pat)))
        if not hits:
            raise OSError("bundled FFmpeg library %s not found in %s" % (pat, d))
        return C.CDLL(hits[0], mode=C.RTLD_GLOBAL)

    avutil = one("libavutil-*.so*")
    one("libswresample-*.so*")
    avcodec = one("libavcodec-*.so*")
    avformat = one("libavformat-*.so*")
    swscale = one("libswscale-*.so*")

    vp = C.c_void_p
    avformat.avformat_open_input.argtypes = [C.POINTER(vp), C.c_char_p, vp, vp]
    avformat.avformat_open_input.restype = C.c_int
    avformat.avformat_find_stream_info.argtypes = [vp, vp]
    avformat.avformat_find_stream_info.restype = C.c_int
    avformat.av_find_best_stream.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp), C.c_int]
    avformat.av_find_best_stream.restype = C.c_int
    avformat.av_read_frame.argtypes = [vp, vp]
    avformat.av_read_frame.restype = C.c_int
    avformat.avformat_close_input.argtypes = [C.POINTER(vp)]
    avformat.avformat_close_input.restype = None
    avcodec.avcodec_alloc_context3.argtypes = [vp]
    avcodec.avcodec_alloc_context3.restype = vp
    avcodec.avcodec_parameters_to_context.argtypes = [vp, vp]
    avcodec.avcodec_parameters_to_context.restype = C.c_int
    avcodec.avcodec_open2.argtypes = [vp, vp, vp]
    avcodec.avcodec_open2.restype = C.c_int
    avcodec.avcodec_send_packet.argtypes = [vp, vp]
    avcodec.avcodec_send_packet.restype = C.c_int
    avcodec.avcodec_receive_frame.argtypes = [vp, vp]
    avcodec.avcodec_receive_frame.restype = C.c_int
    avcodec.avcodec_free_context.argtypes = [C.POINTER(vp)]
    avcodec.avcodec_free_context.restype = None
    avcodec.av_packet_alloc.argtypes = []
    avcodec.av_packet_alloc.restype = vp
    avcodec.av_packet_unref.argtypes = [vp]
    avcodec.av_packet_unref.restype = None
    avcodec.av_packet_free.argtypes = [C.POINTER(vp)]
    avcodec.av_packet_free.restype = None
    avutil.av_frame_alloc.argtypes = []
    avutil.av_frame_alloc.restype = vp
    avutil.av_frame_unref.argtypes = [vp]
    avutil.av_frame_unref.restype = None
    avutil.av_frame_free.argtypes = [C.POINTER(vp)]
    avutil.av_frame_free.restype = None
    avutil.av_crc_get_table.argtypes = [C.c_int]
    avutil.av_crc_get_table.restype = vp
    avutil.av_crc.argtypes = [vp, C.c_uint32, C.c_char_p, C.c_size_t]
    avutil.av_crc.restype = C.c_uint32
    swscale.sws_getContext.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, vp, vp, vp]
    swscale.sws_getContext.restype = vp
    swscale.sws_scale.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int), C.c_int, C.c_int,
                                  C.POINTER(vp), C.POINTER(C.c_int)]
    swscale.sws_scale.restype = C.c_int
    swscale.sws_freeContext.argtypes = [vp]
    swscale.sws_freeContext.restype = None
    _LIBS = dict(avutil=avutil, avcodec=avcodec, avformat=avformat, swscale=swscale)
    return _LIBS


def _rd_i32(addr):
    return C.c_int.from_address(addr).value


def _rd_ptr(addr):
    return C.c_void_p.from_address(addr).value


def _frame_to_nv12(fr):
    """AVFrame (yuv420p, format 0) -> (Y[h,w] u8, UV[h/2,w] u8 interleaved)."""
    w = _rd_i32(fr + _OFF_FRAME_WIDTH)
    h = _rd_i32(fr + _OFF_FRAME_HEIGHT)
    fmt = _rd_i32(fr + _OFF_FRAME_FORMAT)
    if fmt not in (0, 12):  # yuv420p / yuvj420p
        raise RuntimeError("unexpected pix_fmt %d (want yuv420p)" % fmt)
    planes = []
    for p, (pw, ph) in enumerate(((w, h), (w // 2, h // 2), (w // 2, h // 2))):
        ptr = _rd_ptr(fr + _OFF_FRAME_DATA + 8 * p)
        ls = _rd_i32(fr + _OFF_FRAME_LINESIZE + 4 * p)
        buf = (C.c_uint8 * (ls * ph)).from_address(ptr)
        a = np.frombuffer(buf, dtype=np.uint8).reshape(ph, ls)[:, :pw].copy()
        planes.append(a)
    y, u, v = planes
    uv = np.empty((h // 2, w), dtype=np.uint8)
    uv[:, 0::2] = u
    uv[:, 1::2] = v
    return y, uv


def decode_nv12(path, max_frames=1):
    """Decode the first `max_frames` frames of a video file; returns a list of (Y, UV)."""
    L = load_ffmpeg()
    avf, avc, avu = L["avformat"], L["avcodec"], L["avutil"]
    fmt = C.c_void_p(None)
    if avf.avformat_open_input(C.byref(fmt), path.encode(), None, None) < 0:
        raise OSError("cannot open %s" % path)
    out = []
    ctx = C.c_void_p(None)
    pkt = C.c_void_p(avc.av_packet_alloc())
    fr = C.c_void_p(avu.av_frame_alloc())
    try:
        if avf.avformat_find_stream_info(fmt, None) < 0:
            raise OSError("no stream info in %s" % path)
        codec = C.c_void_p(None)
        sidx = avf.av_find_best_stream(fmt, 0, -1, -1, C.byref(codec), 0)
        if sidx < 0:
            raise OSError("no video stream in %s" % path)
        streams = _rd_ptr(fmt.value + _OFF_FMT_STREAMS)
        st = _rd_ptr(streams + 8 * sidx)
        par = _rd_ptr(st + _OFF_STREAM_CODECPAR)
        ctx = C.c_void_p(avc.avcodec_alloc_context3(codec))
        avc.avcodec_parameters_to_context(ctx, par)
        if avc.avcodec_open2(ctx, codec, None) < 0:
            raise OSError("cannot open decoder")

        def drain():
            while len(out) < max_frames:
                r = avc.avcodec_receive_frame(ctx, fr)
                if r in (_AVERROR_EAGAIN, _AVERROR_EOF) or r < 0:
                    return
                out.append(_frame_to_nv12(fr.value))
                avu.av_frame_unref(fr)

        while len(out) < max_frames:
            r = avf.av_read_frame(fmt, pkt)
            if r < 0:
                avc.avcodec_send_packet(ctx, None)  # flush
                drain()
                break
            if _rd_i32(pkt.value + _OFF_PKT_STREAM_INDEX) == sidx:
                avc.avcodec_send_packet(ctx, pkt)
                drain()
            avc.av_packet_unref(pkt)
    finally:
        avu.av_frame_free(C.byref(fr))
        avc.av_packet_free(C.byref(pkt))
        if ctx:
            avc.avcodec_free_context(C.byref(ctx))
        avf.avformat_close_input(C.byref(fmt))
    return out


def av_crc32(buf):
    """FFmpeg av_crc(AV_CRC_32_IEEE, init=0xFFFFFFFF) - what the reference's tests use
    (tests/src/VPPTests.cpp:113). MSB-first poly 0x04C11DB7, no final xor; not zlib.crc32."""
    L = load_ffmpeg()
    b = bytes(buf)
    tab = L["avutil"].av_crc_get_table(3)
    return L["avutil"].av_crc(tab, 0xFFFFFFFF, b, len(b))


def sws_nv12_to_rgb24(y, uv, dst_w=None, dst_h=None, bgr=False, flags=2):
    """FFmpeg swscale NV12 -> RGB24/BGR24 (uint8, merged), the CPU baseline BASELINE.md names
    (R-CPU).  flags: SWS_BILINEAR=2 (| SWS_FULL_CHR_H_INT=0x2000 | SWS_ACCURATE_RND=0x40000)."""
    L = load_ffmpeg()
    sws = L["swscale"]
    h, w = y.shape
    dst_w, dst_h = dst_w or w, dst_h or h
    ctx = sws.sws_getContext(w, h, 23, dst_w, dst_h, 3 if bgr else 2, flags, None, None, None)
    if not ctx:
        raise RuntimeError("sws_getContext failed")
    y = np.ascontiguousarray(y)
    uv = np.ascontiguousarray(uv)
    dst = np.empty((dst_h, dst_w, 3), dtype=np.uint8)
    src = (C.c_void_p * 4)(y.ctypes.data, uv.ctypes.data, None, None)
    sst = (C.c_int * 4)(y.strides[0], uv.strides[0], 0, 0)
    dp = (C.c_void_p * 4)(dst.ctypes.data, None, None, None)
    dst_st = (C.c_int * 4)(dst_w * 3, 0, 0, 0)
    sws.sws_scale(ctx, src, sst, 0, h, dp, dst_st)
    sws.sws_freeContext(ctx)
    return dst
