"""Decoder -> VPP hand-over: the reference calls VideoProcessor::Convert once per
Decoder::GetFrame from up to maxConsumers threads (src/Wrappers/WrapperPython.cpp:265-350, src/Decoder.cpp:97-131).
Here those per-frame calls are coalesced into batched launches behind the unchanged API; the ring keeps the
reference's GetFrame index semantics.  tests/cpp/ring_main.cpp is the reference-shaped C++ consumer."""
import os
import shutil
import subprocess
import threading
import zlib

import numpy as np
import pytest

from tests import golden_cases as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "ring_main")


def _build():
    from tensor_stream_b200 import build as B
    lib = B.build()
    src = os.path.join(ROOT, "tests", "cpp", "ring_main.cpp")
    hdrs = [os.path.join(ROOT, "include", h) for h in ("FrameRing.h", "VideoProcessor.h", "tsvpp.h", "tsvpp_avframe.h")]
    if os.path.exists(BIN) and os.path.getmtime(BIN) > max([os.path.getmtime(src), os.path.getmtime(lib)] + [os.path.getmtime(h) for h in hdrs]):
        return BIN
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not available")
    libdir = os.path.dirname(lib)
    subprocess.check_call(["nvcc", "-std=c++17", "-O2", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "include"), src,
                           "-o", BIN, "-L", libdir, "-l:libtsvpp.so", "-Xlinker", "-rpath=" + libdir, "-Xlinker", "--allow-shlib-undefined"])
    return BIN


def test_ring_consumer_compiles_and_links():
    assert os.path.exists(_build())


@pytest.mark.gpu
def test_getframe_index_semantics():
    """index in [-bufferDeep, 0], > 0 clamped, no wrap-around (VREADER_REPEAT), blocking until a new frame,
    'Decoding finished' on end of stream -- reference src/Decoder.cpp:97-131."""
    out = subprocess.check_output([_build(), "semantics"], text=True)
    assert "semantics ok" in out, out


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{}, {"TSVPP_ASYNC": "1"}, {"TSVPP_COALESCE": "0"}], ids=["coalesced_sync", "coalesced_async", "per_call"])
def test_ring_consumers_bit_exact(env):
    """8 consumer threads, unchanged getFrame sequence, every result compared with a one-frame-per-launch
    conversion of the same surface."""
    e = dict(os.environ, TSVPP_COALESCE_STATS="1", **env)
    r = subprocess.run([_build(), "run", "8", "96", "1"], text=True, capture_output=True, env=e)
    assert r.returncode == 0 and " bad 0 " in r.stdout, r.stdout + r.stderr
    if env.get("TSVPP_COALESCE") != "0":
        stats = [ln for ln in r.stdout.splitlines() if ln.startswith("coalesce_stats")][0].split()
        assert int(stats[2]) == 8 * 96 and int(stats[4]) <= 8 * 96


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"RING_OUT": "pool"}, {"RING_OUT": "pool", "TSVPP_ASYNC": "1"},
                                 {"RING_OUT": "pool", "TSVPP_ASYNC": "1", "TSVPP_CO_EVENTS": "1"},
                                 {"RING_OUT": "pool", "TSVPP_ASYNC": "2"}, {"RING_OUT": "pool", "TSVPP_CO_THREAD": "0"}],
                         ids=["into_sync", "into_async", "into_async_event_fallback", "into_posted", "into_sync_flat_combining"])
def test_ring_consumers_convert_into_recycled_buffers(env):
    """VideoProcessor::ConvertInto (caller-owned result buffers, no CUDA call per frame) through the same ring:
    bit-exact, with completion tickets and with the event fallback."""
    r = subprocess.run([_build(), "run", "8", "128", "1"], text=True, capture_output=True, env=dict(os.environ, **env))
    assert r.returncode == 0 and " bad 0 " in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_ring_older_frames_of_the_buffer():
    r = subprocess.run([_build(), "run", "4", "64", "1", "0", "-3"], text=True, capture_output=True)
    assert r.returncode == 0 and " bad 0 " in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_coalesced_c_abi_from_python_threads():
    """tsvpp_convert_coalesced through ctypes from 8 threads (the GIL is released inside the call): mixed
    parameter sets are grouped per (geometry, parameters), every result equals the oracle."""
    torch = pytest.importorskip("torch")
    from oracle import oracle as O
    from tensor_stream_b200 import vpp
    from tests.test_parity_gpu import _params, _run_oracle, _synthetic
    v = vpp.VideoProcessor(0)
    kws = [dict(dst=(256, 256), resize_type=G.BILINEAR, fourcc=G.BGR24, planes=G.PLANAR, normalization=True),
           dict(dst=(320, 180), resize_type=G.AREA, fourcc=G.RGB24)]
    nthreads, per = 8, 12
    frames = [_synthetic(1280, 720, seed=zlib.crc32(b"co%d" % i)) for i in range(6)]
    dev = [(torch.from_numpy(y).cuda(), torch.from_numpy(uv).cuda()) for y, uv in frames]
    want = {(fi, ki): _run_oracle(frames[fi][0], frames[fi][1], kws[ki]) for fi in range(len(frames)) for ki in range(2)}
    errors = []

    def work(t):
        try:
            for n in range(per):
                fi, ki = (t + n) % len(frames), (t + n) % 2
                fp = _params(kws[ki])
                shape, es = vpp.output_shape(1280, 720, fp)
                out = torch.empty(shape, dtype=torch.float32 if es == 4 else torch.uint8, device="cuda")
                torch.cuda.synchronize()  # `out` comes from torch's allocator on another stream: make it usable
                v.convert_coalesced(dev[fi][0], dev[fi][1], fp, out, wait=(n % 2 == 0))
                torch.cuda.synchronize()
                got = out.cpu().numpy().reshape(-1)
                if got.tobytes() != want[(fi, ki)].tobytes():
                    errors.append((t, n))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    ths = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    frames_done, launches = v.coalesce_stats()
    v.close()
    assert not errors, errors[:5]
    assert frames_done == nthreads * per and launches <= frames_done
