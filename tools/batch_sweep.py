"""Batch-size sweep for the C2 workload: frames per fused launch = 1, 8, 64, 256, 768.
One 1080p->256x256 frame is ~0.4 us of device work, below a kernel launch, so the per-frame API is
launch-bound and the batched entry is what reaches the roofline."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensor_stream_b200 import vpp  # noqa: E402

W, H = 1920, 1080
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(1)
N = 768
ys = [torch.randint(0, 256, (H, W), dtype=torch.uint8, device=dev, generator=g) for _ in range(N)]
uvs = [torch.randint(0, 256, (H // 2, W), dtype=torch.uint8, device=dev, generator=g) for _ in range(N)]
v = vpp.VideoProcessor(0)
fp = vpp.FrameParameters(256, 256, resize_type=vpp.ResizeType.BILINEAR, pixel_format=vpp.FourCC.BGR24,
                         planes_pos=vpp.Planes.PLANAR, normalization=True)
outs = [torch.empty((3, 256, 256), dtype=torch.float32, device=dev) for _ in range(N)]
print("| frames per launch | launches | ms per launch | frames/s | B_full GB/s |")
print("|---|---|---|---|---|")
for bs in (1, 8, 64, 256, 768):
    batches = [v.make_batch(list(zip(ys[i:i + bs], uvs[i:i + bs])), fp, outs[i:i + bs]) for i in range(0, N, bs)]
    for b in batches[:8]:
        b.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3 if bs < 64 else 10
    e0.record()
    for _ in range(reps):
        for b in batches:
            b.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * len(batches))
    fps = bs / (ms * 1e-3)
    print("| %d | %d | %.4f | %.0f | %.0f |" % (bs, len(batches), ms, fps, fps * 3896832 / 1e9), flush=True)
    for b in batches:
        b.destroy()
# the per-frame public call (allocation + descriptor upload + launch, no sync), as VideoProcessor.convert
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(2000):
    v.convert(ys[i % N], uvs[i % N], fp)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("\nper-frame VideoProcessor.convert() through Python/ctypes (torch.empty + upload + launch): %.0f frames/s" % (2000 / dt))
