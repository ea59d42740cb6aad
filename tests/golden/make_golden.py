"""Generates the committed fixtures under tests/golden/ from the reference's own test
resources.  Run in the build container (needs /root/reference); the GPU box only reads the
outputs.

  bbb_1080x608_f1_nv12.npz   frame #1 of tests/resources/bbb_1080x608_420_10.h264 decoded to NV12
                             (software decode; CRC == the reference's NVDEC golden 2957341121)
  float_goldens.json         sha256 + size of the reference's seven float32 golden dumps
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tensor_stream_b200.swdecode import av_crc32, decode_nv12  # noqa: E402

RES = "/root/reference/tests/resources"


def main():
    y, uv = decode_nv12(os.path.join(RES, "bbb_1080x608_420_10.h264"), 1)[0]
    assert av_crc32(y.tobytes() + uv.tobytes()) == 2957341121  # VPPTests.cpp:211-216
    assert av_crc32(y.tobytes()) == 3265466497 and av_crc32(uv.tobytes()) == 2183362287  # DecoderTests.cpp:63-65
    np.savez_compressed(os.path.join(HERE, "bbb_1080x608_f1_nv12.npz"), y=y, uv=uv)
    d = {}
    refdir = os.path.join(RES, "test_references")
    for f in sorted(os.listdir(refdir)):
        b = open(os.path.join(refdir, f), "rb").read()
        d[f] = {"sha256": hashlib.sha256(b).hexdigest(), "bytes": len(b)}
    json.dump(d, open(os.path.join(HERE, "float_goldens.json"), "w"), indent=1, sort_keys=True)
    print("wrote fixtures:", os.listdir(HERE))


if __name__ == "__main__":
    main()
