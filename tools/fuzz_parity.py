#!/usr/bin/env python
"""Randomised parity soak: CUDA path (through the C-ABI) vs the CPU oracle on random geometries, pitches,
crops, resize types, FourCCs, layouts and normalisation.  Test infrastructure (it imports oracle/ as the
checker, like tests/); run it on the GPU box:

    python tools/fuzz_parity.py --seconds 300 --seed 1 > gpurun_out/fuzz.log

Every case is reproducible from its printed line (seed + index).  Exit code 1 if any case differs.
"""
import argparse
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import golden_cases as G  # noqa: E402
from tests import test_parity_gpu as T  # noqa: E402


def even(x):
    return max(2, int(x) & ~1)


def pick_edge_case(rng):
    """Tiny sources and outputs, extreme ratios, one-block crops: the clamps and the partial tiles."""
    w, h = even(rng.choice([2, 4, 6, 8, 16, 18, 30, 34, 62, 66, 130, 258]) + rng.choice([0, 0, 2])), \
        even(rng.choice([2, 4, 6, 8, 16, 18, 30, 34, 62, 66, 130]) + rng.choice([0, 0, 2]))
    if rng.random() < 0.2:
        w, h = rng.choice([(1920, 1080), (1280, 720), (2048, 64), (64, 1200)])
    pitch = rng.choice([w, w, (w + 15) // 16 * 16, w + rng.randint(1, 40)])
    kw = {}
    cw, ch = w, h
    if rng.random() < 0.4 and w >= 6 and h >= 6:
        cw, ch = even(rng.randint(2, w - 2)), even(rng.randint(2, h - 2))
        x0, y0 = even(rng.randint(0, w - cw)), even(rng.randint(0, h - ch))
        if x0 + cw > w:
            x0 = 0
        if y0 + ch > h:
            y0 = 0
        if rng.random() < 0.25:
            x0 += 1 if x0 + cw + 1 <= w else 0
            y0 += 1 if y0 + ch + 1 <= h and rng.random() < 0.5 else 0
        kw["crop"] = (x0, y0, x0 + cw, y0 + ch)
    if rng.random() < 0.85:
        dw = rng.choice([2, 4, 6, 8, 10, 16, 30, 32, 34, 64, 66, 100, 128, 256, 258, 510, 1024])
        dh = rng.choice([2, 4, 6, 8, 10, 16, 30, 32, 34, 64, 66, 100, 128, 256, 258, 510])
        if rng.random() < 0.15:
            dw, dh = dw + 1, dh + rng.choice([0, 1])
        kw["dst"] = (dw, dh)
        kw["resize_type"] = rng.choice([G.NEAREST, G.BILINEAR, G.BICUBIC, G.AREA])
    kw["fourcc"] = rng.choice([G.RGB24, G.BGR24, G.Y800, G.HSV, G.NV12, G.UYVY, G.YUV444])
    kw["planes"] = rng.choice([G.PLANAR, G.MERGED])
    kw["normalization"] = rng.choice([False, True])
    if kw["fourcc"] == G.HSV:
        kw["normalization"] = True
    dw, dh = kw.get("dst", (cw, ch))
    if (dw % 2 or dh % 2) and kw["fourcc"] in (G.NV12, G.UYVY, G.YUV444):
        kw["fourcc"] = G.BGR24
    return w, h, pitch, kw


def pick_wide_area_case(rng):
    """AREA downscales of 30x and more: the narrow-tile plans and the two-stage wide-window path."""
    w, h = rng.choice([(1920, 1080), (1280, 720), (3840, 2160), (2048, 64), (640, 1200), (1000, 600)])
    pitch = rng.choice([w, (w + 255) // 256 * 256, w + rng.randint(1, 40)])
    kw = {}
    if rng.random() < 0.3:
        cw, ch = even(rng.randint(w // 2, w - 2)), even(rng.randint(h // 2, h - 2))
        x0, y0 = even(rng.randint(0, w - cw)), even(rng.randint(0, h - ch))
        kw["crop"] = (x0, y0, x0 + cw, y0 + ch)
    kw["dst"] = (rng.randint(1, 40), rng.randint(2, 24))  # one-row outputs of such ratios are TSVPP_UNSUPPORTED
    kw["resize_type"] = G.AREA
    kw["fourcc"] = rng.choice([G.RGB24, G.BGR24, G.Y800, G.HSV, G.NV12, G.UYVY, G.YUV444])
    kw["planes"] = rng.choice([G.PLANAR, G.MERGED])
    kw["normalization"] = rng.choice([False, True])
    if kw["fourcc"] == G.HSV:
        kw["normalization"] = True
    dw, dh = kw["dst"]
    if (dw % 2 or dh % 2) and kw["fourcc"] in (G.NV12, G.UYVY, G.YUV444):
        kw["fourcc"] = G.BGR24
    return w, h, pitch, kw


def pick_case(rng):
    r = rng.random()
    if r < 0.25:
        w, h = rng.choice([(1920, 1080), (1280, 720), (1080, 608), (640, 360), (960, 540)])
    elif r < 0.30:
        w, h = 3840, 2160
    else:
        w, h = even(rng.randint(32, 2048)), even(rng.randint(32, 1200))
    pr = rng.random()
    if pr < 0.35:
        pitch = w
    elif pr < 0.6:
        a = rng.choice([16, 64, 256, 512])
        pitch = (w + a - 1) // a * a
    else:
        pitch = w + rng.randint(1, 200)
    kw = {}
    cw, ch = w, h
    if rng.random() < 0.4 and w >= 64 and h >= 64:
        cw, ch = even(rng.randint(16, w - 2)), even(rng.randint(16, h - 2))
        x0, y0 = even(rng.randint(0, w - cw)) if w - cw >= 2 else 0, even(rng.randint(0, h - ch)) if h - ch >= 2 else 0
        if x0 + cw > w:
            x0 = 0
        if y0 + ch > h:
            y0 = 0
        if rng.random() < 0.25:  # odd origins are defined behaviour in the reference (Crop.cu:9-16), even sizes only
            x0 += 1 if x0 + cw + 1 <= w else 0
            y0 += 1 if y0 + ch + 1 <= h and rng.random() < 0.5 else 0
        kw["crop"] = (x0, y0, x0 + cw, y0 + ch)
    dr = rng.random()
    if dr < 0.15:
        pass  # no resize
    else:
        if dr < 0.45:  # dyadic-friendly targets (integer march kernels)
            fx, fy = rng.choice([0.25, 0.5, 2.0, 2 / 3, 1.5, 0.125, 4 / 3, 0.75]), rng.choice([0.25, 0.5, 2.0, 2 / 3, 1.5, 0.125, 0.75])
            dw, dh = even(cw * fx), even(ch * fy)
        elif dr < 0.6:
            dw, dh = rng.choice([(256, 256), (128, 128), (224, 224), (299, 299), (512, 512), (720, 480), (1280, 720), (64, 64)])
        elif dr < 0.9:
            dw, dh = even(rng.randint(8, 1400)), even(rng.randint(8, 900))
        else:
            dw, dh = rng.randint(8, 700), rng.randint(8, 500)  # odd sizes included
        while dw * dh > 2_500_000:
            dw, dh = even(dw // 2), even(dh // 2)
        kw["dst"] = (dw, dh)
        kw["resize_type"] = rng.choice([G.NEAREST, G.BILINEAR, G.BICUBIC, G.AREA])
    kw["fourcc"] = rng.choice([G.RGB24, G.BGR24, G.RGB24, G.BGR24, G.Y800, G.HSV, G.NV12, G.UYVY, G.YUV444])
    kw["planes"] = rng.choice([G.PLANAR, G.MERGED])
    kw["normalization"] = rng.choice([False, True])
    if kw["fourcc"] == G.HSV:
        kw["normalization"] = True
    dw, dh = kw.get("dst", (cw, ch))
    if (dw % 2 or dh % 2) and kw["fourcc"] in (G.NV12, G.UYVY, G.YUV444):
        # odd sizes with the 4:2:0 / 4:2:2 outputs are races / out-of-bounds stores in the reference
        kw["fourcc"] = G.RGB24
    return w, h, pitch, kw


def run_api(v, api, y, uv, kw, pitch):
    """The same conversion through one of the library's entry points; returns the flat result(s)."""
    import torch
    from tensor_stream_b200 import vpp
    h, w = y.shape
    fp = T._params(kw)
    if api == "device":
        return [T._run_cuda(v, y, uv, kw, pitch)]
    if api in ("host", "pinned"):
        py = np.zeros((h, pitch), dtype=np.uint8)
        puv = np.zeros((h // 2, pitch), dtype=np.uint8)
        py[:, :w] = y
        puv[:, :w] = uv
        if api == "pinned":
            ty, tuv = torch.from_numpy(py).pin_memory(), torch.from_numpy(puv).pin_memory()
            outs = v.convert_batch_host([(ty, tuv), (ty, tuv)], fp, w, h)
        else:
            outs = v.convert_batch_host([(py, puv), (py, puv)], fp, w, h)
        return [np.asarray(o).reshape(-1) for o in outs]
    ty, tuv, _ = T._to_dev(y, uv, pitch)
    shape, es = vpp.output_shape(w, h, fp)
    dt = torch.float32 if es == 4 else torch.uint8
    if api == "batch":
        outs = [torch.zeros(shape, dtype=dt, device="cuda") for _ in range(3)]
        b = v.make_batch([(ty, tuv)] * 3, fp, outs, widths=(w, h))
        b.run()
        torch.cuda.synchronize()
        b.destroy()
        return [o.cpu().numpy().reshape(-1) for o in outs]
    if api == "coalesced":
        out = torch.zeros(shape, dtype=dt, device="cuda")
        v.convert_coalesced(ty, tuv, fp, out, width=w, height=h)
        torch.cuda.synchronize()
        return [out.cpu().numpy().reshape(-1)]
    raise ValueError(api)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-cases", type=int, default=100000)
    ap.add_argument("--profile", default="broad", choices=["broad", "edge", "wide-area"])
    ap.add_argument("--api", default="device", choices=["device", "host", "pinned", "batch", "coalesced", "all"],
                    help="entry point: tsvpp_convert_batch (default), the host-buffer entry with pageable or pinned buffers, "
                         "a prebuilt batch, the per-frame coalescer, or a random one per case")
    args = ap.parse_args()
    import torch
    from tensor_stream_b200 import vpp
    assert torch.cuda.is_available(), "needs a GPU"
    v = vpp.VideoProcessor(0)
    rng = random.Random(args.seed)
    t0 = time.time()
    n = ok = bad = rejected = both_rejected = 0
    kernels = {}
    for idx in range(args.max_cases):
        if time.time() - t0 > args.seconds:
            break
        w, h, pitch, kw = {"edge": pick_edge_case, "wide-area": pick_wide_area_case, "broad": pick_case}[args.profile](rng)
        nrng = np.random.default_rng(args.seed * 1000003 + idx)
        if rng.random() < 0.3:
            y, uv = T._synthetic(w, h, args.seed * 7919 + idx, video_like=True)
        else:
            y = nrng.integers(0, 256, (h, w), dtype=np.uint8)
            uv = nrng.integers(0, 256, (h // 2, w), dtype=np.uint8)
        n += 1
        tag = "case %d seed %d: %dx%d pitch %d %s" % (idx, args.seed, w, h, pitch, kw)
        try:
            want = T._run_oracle(y, uv, kw)
            oracle_err = None
        except Exception as e:  # noqa: BLE001
            want, oracle_err = None, e
        api = args.api if args.api != "all" else rng.choice(["device", "host", "pinned", "batch", "coalesced"])
        tag += " api=" + api
        try:
            gots = run_api(v, api, y, uv, kw, pitch)
            got, cuda_err = gots[0], None
        except vpp.VppError as e:
            gots, got, cuda_err = [], None, e
        if cuda_err is not None or oracle_err is not None:
            if cuda_err is not None and oracle_err is not None:
                both_rejected += 1
            else:
                rejected += 1
                print("REJECTED by %s only: %s  (%r)" % ("cuda" if cuda_err else "oracle", tag, cuda_err or oracle_err), flush=True)
            continue
        try:
            ty, tuv, _ = T._to_dev(y, uv, pitch)
            plan = v.describe_plan(ty, tuv, T._params(kw), width=w, height=h)
            k = " ".join(p for p in plan.split() if p.startswith(("kernel=", "resample=", "bulk=")))
        except Exception:  # noqa: BLE001
            k = "no fused plan"
        kernels[k] = kernels.get(k, 0) + 1
        b = T._mask_undefined(want, kw, w, h)
        good = True
        for got in gots:
            a = T._mask_undefined(got, kw, w, h)
            if not T._same(a, b):
                good = False
                print("MISMATCH %s [%s]: %s" % (tag, k, T._diff_report(np.asarray(a), np.asarray(b))), flush=True)
                break
        ok += good
        bad += not good
    print("cases %d ok %d mismatches %d rejected-by-one-side %d rejected-by-both %d in %.0f s" %
          (n, ok, bad, rejected, both_rejected, time.time() - t0))
    for k in sorted(kernels, key=kernels.get, reverse=True):
        print("  %5d  %s" % (kernels[k], k))
    v.close()
    sys.exit(1 if bad or rejected else 0)


if __name__ == "__main__":
    main()
