"""BASELINE.json configs[2] (C3) and configs[4] (C5: resize sweep nearest/bilinear/bicubic/area x output
128^2 .. 2160p): frames/sec and B_full GB/s of the fused path next to the reference's own CUDA VPP
(oracle/_ref, per-frame VideoProcessor::Convert).  Developer / profiles tool, not the contract bench.
Writes a markdown table to stdout."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensor_stream_b200 import vpp  # noqa: E402

PEAK = 6572.2
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def pool(W, H, n, pitch=None):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(99)
    pitch = pitch or W
    ys = [torch.randint(0, 256, (H, pitch), dtype=torch.uint8, device=dev, generator=g) for _ in range(n)]
    uvs = [torch.randint(0, 256, (H // 2, pitch), dtype=torch.uint8, device=dev, generator=g) for _ in range(n)]
    return ys, uvs


def time_ours(v, ys, uvs, fp, W, H, iters):
    n = len(ys)
    shape, es = vpp.output_shape(W, H, fp)
    outs = [torch.empty(shape, dtype=torch.float32 if es == 4 else torch.uint8, device="cuda:0") for _ in range(n)]
    b = v.make_batch(list(zip(ys, uvs)), fp, outs, widths=(W, H))
    for _ in range(3):
        b.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        b.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    b.destroy()
    return n / (ms * 1e-3), int(np.prod(shape)) * es


def time_ref(ref, ys, uvs, W, H, kw, max_s=2.0):
    n = min(len(ys), 48)
    for i in range(4):
        p, _, _ = ref.convert_raw(ys[i].data_ptr(), uvs[i].data_ptr(), int(ys[i].stride(0)), int(uvs[i].stride(0)), W, H, **kw)
        ref.free(p)
    ref.sync()
    t0 = time.perf_counter()
    done = 0
    while time.perf_counter() - t0 < max_s:
        for i in range(n):
            p, _, _ = ref.convert_raw(ys[i].data_ptr(), uvs[i].data_ptr(), int(ys[i].stride(0)), int(uvs[i].stride(0)), W, H, **kw)
            ref.free(p)
        done += n
    ref.sync()
    return done / (time.perf_counter() - t0)


def main():
    from oracle import refvpp as R
    v = vpp.VideoProcessor(0)
    ref = R.RefVPP(5) if R.available() else None
    print("| config | ours fps | ours B_full GB/s | frac of %.0f GB/s | plan | reference CUDA VPP fps | speed-up |" % PEAK)
    print("|---|---|---|---|---|---|---|")
    W, H = 1920, 1080
    ys, uvs = pool(W, H, 160)
    rows = []
    for rt in vpp.ResizeType:
        for (dw, dh) in ((128, 128), (256, 256), (512, 512), (1024, 1024), (1920, 1080), (3840, 2160)):
            for norm, fc, pl in ((False, vpp.FourCC.RGB24, vpp.Planes.MERGED), (True, vpp.FourCC.BGR24, vpp.Planes.PLANAR)):
                if (dw, dh) == (1920, 1080) and rt != vpp.ResizeType.NEAREST:
                    continue  # same size: no resize kernel runs whatever the type
                name = "1080p->%dx%d %s %s" % (dw, dh, rt.name.lower(), "bgr planar f32" if norm else "rgb merged u8")
                if os.environ.get("CS_ONLY") and os.environ["CS_ONLY"] not in name:
                    continue
                fp = vpp.FrameParameters(dw, dh, resize_type=rt, pixel_format=fc, planes_pos=pl, normalization=norm)
                nfr = 160 if dw * dh <= 1024 * 1024 else (48 if dw * dh <= 1920 * 1080 else 16)
                fps, outb = time_ours(v, ys[:nfr], uvs[:nfr], fp, W, H, 10)
                bfull = W * H * 1.5 + outb
                plan = v.describe_plan(ys[0], uvs[0], fp)
                keep = " ".join(x for x in plan.split() if x.split("=")[0] in ("tw", "rp", "K", "nstages", "ctas_per_sm"))
                kw = dict(dst=(dw, dh), resize_type=rt.value, fourcc=fc.value, planes=pl.value, normalization=norm)
                rfps = time_ref(ref, ys, uvs, W, H, kw, 1.0) if (ref and not os.environ.get("CS_NOREF")) else float("nan")
                print("| %s | %.0f | %.0f | %.2f | %s | %.0f | %.0fx |" % (name, fps, fps * bfull / 1e9, fps * bfull / 1e9 / PEAK, keep, rfps, fps / rfps), flush=True)
    # C3: 3840x2160 NV12, crop (0,0,1920,1080) -> RGB24 merged 1280x720 bicubic uint8
    W, H = 3840, 2160
    del ys, uvs
    torch.cuda.empty_cache()
    ys, uvs = pool(W, H, 40)
    fp = vpp.FrameParameters(1280, 720, crop_coords=(0, 0, 1920, 1080), resize_type=vpp.ResizeType.BICUBIC)
    fps, outb = time_ours(v, ys, uvs, fp, W, H, 10)
    bfull = 1920 * 1080 * 1.5 + outb
    plan = v.describe_plan(ys[0], uvs[0], fp)
    keep = " ".join(x for x in plan.split() if x.split("=")[0] in ("tw", "rp", "K", "nstages", "ctas_per_sm"))
    kw = dict(crop=(0, 0, 1920, 1080), dst=(1280, 720), resize_type=2, fourcc=1, planes=1, normalization=False)
    rfps = time_ref(ref, ys, uvs, W, H, kw, 2.0) if (ref and not os.environ.get("CS_NOREF")) else float("nan")
    print("| C3 4K crop 1080p->1280x720 bicubic rgb merged u8 | %.0f | %.0f | %.2f | %s | %.0f | %.0fx |" % (fps, fps * bfull / 1e9, fps * bfull / 1e9 / PEAK, keep, rfps, fps / rfps))


if __name__ == "__main__":
    main()
