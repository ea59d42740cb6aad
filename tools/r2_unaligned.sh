mkdir -p gpurun_out/r2k; rm -f gpurun_out/r2k/*
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2k/pytest.log 2>&1; echo "rc=$?"
grep -E 'passed|failed|FAILED|Error' gpurun_out/r2k/pytest.log | tail -20
python - <<'PY' > gpurun_out/r2k/unaligned.txt 2>&1
import torch, time, os, sys
sys.path.insert(0, os.getcwd())
from tensor_stream_b200 import vpp
v = vpp.VideoProcessor(0)
dev = torch.device("cuda:0")
def run(w, h, pitch, off, fp, n=160, iters=20):
    g = torch.Generator(device=dev); g.manual_seed(1)
    bufs = [torch.randint(0, 256, (h * 3 // 2 * pitch + 64,), dtype=torch.uint8, device=dev, generator=g) for _ in range(n)]
    ys = [b[off:off + h * pitch].view(h, pitch) for b in bufs]
    uvs = [b[off + h * pitch: off + h * pitch + (h // 2) * pitch].view(h // 2, pitch) for b in bufs]
    shape, es = vpp.output_shape(w, h, fp)
    outs = [torch.empty(shape, dtype=torch.float32 if es == 4 else torch.uint8, device=dev) for _ in range(n)]
    plan = v.describe_plan(ys[0], uvs[0], fp, width=w, height=h)
    b = v.make_batch(list(zip(ys, uvs)), fp, outs, widths=(w, h))
    for _ in range(3): b.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): b.run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("%dx%d pitch %d offset %d: %.0f fps | %s" % (w, h, pitch, off, n / (ms * 1e-3), plan.split(" tw=")[0]))
    b.destroy()
fp = vpp.FrameParameters(256, 256, resize_type=vpp.ResizeType.BILINEAR, pixel_format=vpp.FourCC.BGR24, planes_pos=vpp.Planes.PLANAR, normalization=True)
run(1080, 608, 1088, 0, fp)      # aligned pitch (bulk)
run(1080, 608, 1080, 0, fp)      # tight clip: pitch = 8 mod 16
run(1080, 608, 1088, 4, fp)      # aligned pitch, base offset 4
run(1918, 1080, 1918, 0, fp)     # pitch = 14 mod 16
fp2 = vpp.FrameParameters(0, 0, pixel_format=vpp.FourCC.RGB24)
run(1080, 608, 1080, 0, fp2)     # no resize from a tight clip
PY
cat gpurun_out/r2k/unaligned.txt
