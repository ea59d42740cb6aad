"""BASELINE.json configs[0]: tests/resources/bunny.mp4, 10 frames, FFmpeg software decode + swscale
NV12->RGB24 on CPU against the oracle -- golden-harness / plumbing validation, no GPU
(SURVEY.md 8d "Config 1": |delta| <= 2 LSB with default flags, <= 1 with accurate rounding)."""
import os

import numpy as np
import pytest

BUNNY = "/root/reference/tests/resources/bunny.mp4"


@pytest.mark.skipif(not os.path.exists(BUNNY), reason="reference resources not mounted")
def test_bunny_10_frames_oracle_vs_swscale():
    from oracle import oracle as O
    from tensor_stream_b200 import swdecode as S
    frames = S.decode_nv12(BUNNY, 10)
    assert len(frames) == 10 and frames[0][0].shape == (720, 1280)
    worst_default, worst_accurate = 0, 0
    for y, uv in frames:
        ours = O.convert(y, uv, fourcc=O.RGB24)[0].reshape(720, 1280, 3).astype(np.int16)
        d = S.sws_nv12_to_rgb24(y, uv).astype(np.int16)
        a = S.sws_nv12_to_rgb24(y, uv, flags=2 | 0x2000 | 0x40000).astype(np.int16)
        worst_default = max(worst_default, int(np.abs(ours - d).max()))
        worst_accurate = max(worst_accurate, int(np.abs(ours - a).max()))
    assert worst_default <= 2 and worst_accurate <= 2, (worst_default, worst_accurate)
