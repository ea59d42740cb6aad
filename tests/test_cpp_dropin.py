"""The C++ drop-in (include/VideoProcessor.h + libtsvpp.so) used the way the reference's own
gtest VPP suite uses VideoProcessor (tests/src/VPPTests.cpp): compile a program against the header
(CPU check: the reference-style code compiles and links), run it on the GPU against golden CRCs."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from tests import golden_cases as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "dropin_main")


def _build():
    from tensor_stream_b200 import build as B
    lib = B.build()
    src = os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp")
    if os.path.exists(BIN) and os.path.getmtime(BIN) > max(os.path.getmtime(src), os.path.getmtime(lib)):
        return BIN
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not available")
    libdir = os.path.dirname(lib)
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", BIN,
                           "-L", libdir, "-l:libtsvpp.so", "-Xlinker", "-rpath=" + libdir, "-Xlinker", "--allow-shlib-undefined"])
    return BIN


def test_reference_style_cpp_compiles_and_links():
    assert os.path.exists(_build())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rgb_merged", "bgr_planar", "rgb_bilinear_480x360", "rgb_bicubic_540x304",
                                  "rgb_area_480x360", "crop_down_right", "crop_area_down_center", "yuv444", "uyvy_720x480"])
def test_cpp_dropin_golden_crc(golden_frame, tmp_path, name):
    kw, crcs = next((c[1], c[2]) for c in G.CRC_CASES if c[0] == name)
    y, uv = golden_frame
    pitch = 1088
    raw = np.zeros((608 * 3 // 2, pitch), dtype=np.uint8)
    raw[:608, :1080] = y
    raw[608:, :1080] = uv
    p = tmp_path / "frame.nv12"
    raw.tofile(str(p))
    dst = kw.get("dst", (0, 0))
    crop = kw.get("crop", (0, 0, 0, 0))
    out = subprocess.check_output([_build(), str(p), "1080", "608", str(pitch), str(kw.get("fourcc", G.RGB24)),
                                   str(kw.get("planes", G.MERGED)), "0", str(dst[0]), str(dst[1]),
                                   str(kw.get("resize_type", G.NEAREST))] + [str(v) for v in crop],
                                  cwd=str(tmp_path), text=True)
    d = dict(ln.split(" ", 1) for ln in out.strip().splitlines())
    assert int(d["crc"]) in crcs and int(d["dump_crc"]) in crcs, out
    assert d["free"] == "0"                      # cudaFree accepts the returned buffer
    assert d["sixth_consumer"].strip() == "-3"   # VREADER_ERROR: no free stream slot
    if dst == (0, 0):
        assert d["resize_backfilled"] == d["size"]
