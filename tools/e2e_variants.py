"""Developer tool: end-to-end (host buffers) throughput under different host-path knobs."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensor_stream_b200 import vpp  # noqa: E402

W, H, NE = 1920, 1080, 64
v = vpp.VideoProcessor(0)
fp = vpp.FrameParameters(256, 256, resize_type=vpp.ResizeType.BILINEAR, pixel_format=vpp.FourCC.BGR24,
                         planes_pos=vpp.Planes.PLANAR, normalization=True)
hy = [torch.randint(0, 256, (H, W), dtype=torch.uint8).pin_memory() for _ in range(NE)]
huv = [torch.randint(0, 256, (H // 2, W), dtype=torch.uint8).pin_memory() for _ in range(NE)]
hout = [torch.empty((3, 256, 256), dtype=torch.float32).pin_memory() for _ in range(NE)]
frames = list(zip(hy, huv))
QUICK = os.environ.get("E2E_QUICK") == "1"  # zero-copy in/out and staged, default chunk only
for zc in ("1", "0"):
    for zo in (("1",) if QUICK else ("1", "0")):
        for chunk in ((16,) if QUICK else (8, 16, 32, 64)):
            os.environ.update(TSVPP_HOST_ZEROCOPY=zc, TSVPP_HOST_ZEROCOPY_OUT=zo, TSVPP_HOST_CHUNK=str(chunk))
            for _ in range(2):
                v.convert_batch_host(frames, fp, W, H, hout)
            t0 = time.perf_counter()
            for _ in range(8):
                v.convert_batch_host(frames, fp, W, H, hout)
            dt = time.perf_counter() - t0
            print("zero_copy_in=%s zero_copy_out=%s chunk=%2d -> %.0f fps" % (zc, zo, chunk, NE * 8 / dt), flush=True)
