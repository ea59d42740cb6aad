"""Developer timing loop (not the contract bench): device-resident pool, CUDA events."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensor_stream_b200 import vpp  # noqa: E402


def main():
    n = int(os.environ.get("QB_FRAMES", "160"))
    iters = int(os.environ.get("QB_ITERS", "20"))
    W, H = 1920, 1080
    cfg = os.environ.get("QB_CFG", "c2")
    if cfg == "c3":
        W, H = 3840, 2160
        n = min(n, 40)
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    ys = [torch.randint(0, 256, (H, W), dtype=torch.uint8, device=dev, generator=g) for _ in range(n)]
    uvs = [torch.randint(0, 256, (H // 2, W), dtype=torch.uint8, device=dev, generator=g) for _ in range(n)]
    v = vpp.VideoProcessor(0)
    cfgs = [("C2 bilinear 256 bgr planar f32", vpp.FrameParameters(256, 256, resize_type=vpp.ResizeType.BILINEAR,
             pixel_format=vpp.FourCC.BGR24, planes_pos=vpp.Planes.PLANAR, normalization=True))]
    if cfg == "c3":
        cfgs = [("C3 4K crop 1080p -> 1280x720 bicubic rgb merged u8", vpp.FrameParameters(
            1280, 720, crop_coords=(0, 0, 1920, 1080), resize_type=vpp.ResizeType.BICUBIC))]
    elif cfg == "noresize":
        cfgs = [("1080p NV12 -> RGB24 merged u8, native size (the reference's default read())", vpp.FrameParameters())]
    elif cfg == "bil4k":
        cfgs = [("1080p -> 3840x2160 bilinear bgr planar f32", vpp.FrameParameters(3840, 2160, resize_type=vpp.ResizeType.BILINEAR,
                 pixel_format=vpp.FourCC.BGR24, planes_pos=vpp.Planes.PLANAR, normalization=True))]
    elif cfg == "area":
        cfgs = [("area 256 bgr planar f32", vpp.FrameParameters(256, 256, resize_type=vpp.ResizeType.AREA,
                 pixel_format=vpp.FourCC.BGR24, planes_pos=vpp.Planes.PLANAR, normalization=True))]
    elif cfg.startswith("x:"):
        # x:<resize>:<dw>x<dh>:<fourcc>:<planes>:<f32|u8>, e.g. x:AREA:128x128:BGR24:PLANAR:f32
        _, rt, dst, fcc, pl, ty = cfg.split(":")
        dw_, dh_ = (int(v) for v in dst.split("x"))
        cfgs = [(cfg, vpp.FrameParameters(dw_, dh_, resize_type=getattr(vpp.ResizeType, rt), pixel_format=getattr(vpp.FourCC, fcc),
                                          planes_pos=getattr(vpp.Planes, pl), normalization=(ty == "f32")))]
    if os.environ.get("QB_ALL"):
        for rt in vpp.ResizeType:
            cfgs.append(("%s 256 rgb merged u8" % rt.name, vpp.FrameParameters(256, 256, resize_type=rt)))
    for name, fp in cfgs:
        shape, es = vpp.output_shape(W, H, fp)
        outs = [torch.empty(shape, dtype=torch.float32 if es == 4 else torch.uint8, device=dev) for _ in range(n)]
        try:
            print(name, "|", v.describe_plan(ys[0], uvs[0], fp))
        except vpp.VppError:
            print(name, "| (packed YUV: no fused plan)")
        b = v.make_batch(list(zip(ys, uvs)), fp, outs)
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
        fps = n / (ms * 1e-3)
        sw, sh = (1920, 1080) if cfg == "c3" else (W, H)
        bfull = sw * sh * 1.5 + int(np.prod(shape)) * es
        print("  %.3f ms / %d frames -> %.0f fps, B_full %.1f GB/s" % (ms, n, fps, fps * bfull / 1e9))
        b.destroy()


if __name__ == "__main__":
    main()
