"""The reference's stage-level entry points kept by the drop-in header (cropHost, resizeKernel,
colorConversionKernel<T>; reference include/VideoProcessor.h:109-115): a reference-style chain of the three
must give the bytes the fused Convert() gives, with the reference's buffer contract (separately
cudaFree-able Y / UV intermediates, result in dst->opaque)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests import golden_cases as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "stages_main")


def _build():
    from tensor_stream_b200 import build as B
    lib = B.build()
    src = os.path.join(ROOT, "tests", "cpp", "stages_main.cpp")
    if os.path.exists(BIN) and os.path.getmtime(BIN) > max(os.path.getmtime(src), os.path.getmtime(lib)):
        return BIN
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not available")
    libdir = os.path.dirname(lib)
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", BIN,
                           "-L", libdir, "-l:libtsvpp.so", "-Xlinker", "-rpath=" + libdir, "-Xlinker", "--allow-shlib-undefined"])
    return BIN


def test_stage_functions_compile_and_link():
    """CPU check: code written against the reference's stage functions compiles against the drop-in
    header and links with libtsvpp.so (all four symbols exist with the reference's signatures)."""
    assert os.path.exists(_build())


CASES = [
    # name, fourcc, planes, norm, dst, resize type, crop
    ("crop_bilinear_bgr_planar_f32", G.BGR24, G.PLANAR, 1, (320, 240), G.BILINEAR, (64, 32, 704, 512)),
    ("resize_bicubic_rgb_merged_u8", G.RGB24, G.MERGED, 0, (540, 304), G.BICUBIC, (0, 0, 0, 0)),
    ("crop_only_y800", G.Y800, G.MERGED, 0, (0, 0), G.NEAREST, (100, 60, 740, 420)),
    ("crop_area_nv12_f32", G.NV12, G.MERGED, 1, (250, 150), G.AREA, (40, 20, 1040, 580)),
    ("colour_only_hsv", G.HSV, G.MERGED, 1, (0, 0), G.NEAREST, (0, 0, 0, 0)),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,fourcc,planes,norm,dst,rtype,crop", CASES, ids=[c[0] for c in CASES])
def test_stage_chain_equals_fused_convert(golden_frame, tmp_path, name, fourcc, planes, norm, dst, rtype, crop):
    y, uv = golden_frame
    pitch = 1088
    raw = np.zeros((608 * 3 // 2, pitch), dtype=np.uint8)
    raw[:608, :1080] = y
    raw[608:, :1080] = uv
    p = tmp_path / "frame.nv12"
    raw.tofile(str(p))
    out = subprocess.check_output([_build(), str(p), "1080", "608", str(pitch), str(fourcc), str(planes), str(norm),
                                   str(dst[0]), str(dst[1]), str(rtype)] + [str(v) for v in crop],
                                  cwd=str(tmp_path), text=True)
    d = dict(ln.split(" ", 1) for ln in out.strip().splitlines())
    assert d["convert"] == "0" and d["chain"] == "0", out
    a, b = d["size"].split(" vs ")
    assert a == b, out
    assert int(d["bytes"]) > 0 and d["diff"] == "0", out
    assert d["free"] == "0", out                          # every buffer the stages handed out is cudaFree-able
    assert d["float_without_normalization"] == "-2", out  # VREADER_UNSUPPORTED
    assert d["unknown_fourcc"] == "52", out               # cudaErrorMissingConfiguration (src/ColorConversion.cu:376)
