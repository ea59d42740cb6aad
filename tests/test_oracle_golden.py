"""Pins the CPU oracle (oracle/vpp_oracle.c) against the reference's own golden vectors
(SURVEY.md 4.1): 39 CRC cases on the committed fixture frame, the 7 float32 golden dumps, and -
when the reference's streams are available (/root/reference, build container only) - the
multi-frame CRCs."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import golden_cases as G

HERE = os.path.dirname(os.path.abspath(__file__))
RES = "/root/reference/tests/resources"


def test_crc_matches_ffmpeg_definition():
    # CRC-32/MPEG-2("123456789") = 0x0376E6E7; FFmpeg keeps the non-reflected CRC byte-swapped
    # on little-endian hosts (libavutil crc.c), hence 0xE7E67603
    assert O.av_crc32(b"123456789") == 0xE7E67603
    assert O.av_crc32(b"") == 0xFFFFFFFF


def test_fixture_frame_is_the_reference_frame(golden_frame):
    y, uv = golden_frame
    assert y.shape == (608, 1080) and uv.shape == (304, 1080)
    assert O.av_crc32(y.tobytes()) == 3265466497  # DecoderTests.cpp:63-65
    assert O.av_crc32(uv.tobytes()) == 2183362287


@pytest.mark.parametrize("name,kw,crcs,cite", G.CRC_CASES, ids=[c[0] for c in G.CRC_CASES])
def test_golden_crc(golden_frame, name, kw, crcs, cite):
    y, uv = golden_frame
    out, _ = O.convert(y, uv, **kw)
    assert O.av_crc32(out.tobytes()) in crcs, "oracle diverges from reference golden (%s)" % cite


@pytest.mark.parametrize("fname,kw", G.FLOAT_CASES, ids=[c[0] for c in G.FLOAT_CASES])
def test_golden_float_dump(golden_frame, fname, kw):
    y, uv = golden_frame
    out, _ = O.convert(y, uv, **kw)
    meta = json.load(open(os.path.join(HERE, "golden", "float_goldens.json")))[fname]
    b = out.astype(np.float32).tobytes()
    assert len(b) == meta["bytes"]
    assert hashlib.sha256(b).hexdigest() == meta["sha256"]
    ref = os.path.join(RES, "test_references", fname)
    if os.path.exists(ref):  # the file itself, when the reference tree is mounted
        assert b == open(ref, "rb").read()


@pytest.mark.skipif(not os.path.isdir(RES), reason="reference streams not mounted")
@pytest.mark.parametrize("stream,nframes,kw,crc,cite", G.MULTIFRAME_CASES,
                         ids=["%s_%dx%d" % (c[0][:8], c[2]["dst"][0], c[2]["dst"][1]) for c in G.MULTIFRAME_CASES])
def test_golden_multiframe(stream, nframes, kw, crc, cite):
    from tensor_stream_b200.swdecode import decode_nv12
    frames = decode_nv12(os.path.join(RES, stream), nframes)
    assert len(frames) == nframes
    blob = b"".join(O.convert(y, uv, **kw)[0].tobytes() for (y, uv) in frames)
    assert O.av_crc32(blob) == crc, cite


def test_oracle_crc_equals_libavutil():
    sw = pytest.importorskip("tensor_stream_b200.swdecode")
    try:
        sw.load_ffmpeg()
    except OSError:
        pytest.skip("bundled FFmpeg not present")
    rng = np.random.default_rng(7)
    for n in (0, 1, 3, 255, 4096, 100001):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert O.av_crc32(b) == sw.av_crc32(b)
