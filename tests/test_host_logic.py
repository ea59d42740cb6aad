"""CPU-side tests: the C-ABI library loads and exports every symbol include/tsvpp.h declares,
the host dispatch rules equal the oracle's, helper arithmetic used by the kernels is exact, and
the multi-rank layout (world_size 2, gloo) behaves.  No GPU compute here."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from tensor_stream_b200 import build as B
    from tensor_stream_b200 import vpp
    B.build()
    return vpp.load_library()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "tsvpp.h")).read()
    names = set(re.findall(r"\b(tsvpp_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 11
    for n in sorted(names):
        assert hasattr(lib, n), "libtsvpp.so does not export %s" % n
    assert b"sm_100a" in lib.tsvpp_version()


def test_enum_numbering_is_the_references():
    from tensor_stream_b200 import vpp
    # reference include/VideoProcessor.h:20-28,32-35,59-64
    assert [f.value for f in vpp.FourCC] == [0, 1, 2, 3, 4, 5, 6]
    assert [f.name for f in vpp.FourCC] == ["Y800", "RGB24", "BGR24", "NV12", "UYVY", "YUV444", "HSV"]
    assert vpp.Planes.PLANAR.value == 0 and vpp.Planes.MERGED.value == 1
    assert [r.name for r in vpp.ResizeType] == ["NEAREST", "BILINEAR", "BICUBIC", "AREA"]


def test_frame_parameters_defaults_match_reference():
    from tensor_stream_b200 import vpp
    fp = vpp.FrameParameters()
    # ColorOptions(): RGB24, MERGED, normalization=false; ResizeOptions(): NEAREST (include/VideoProcessor.h:38-78)
    assert (fp.pixel_format, fp.planes_pos, fp.normalization, fp.resize_type) == \
        (vpp.FourCC.RGB24, vpp.Planes.MERGED, False, vpp.ResizeType.NEAREST)
    assert vpp.FrameParameters(pixel_format=vpp.FourCC.HSV).normalization is True  # HSV ctor rule (:45-46)


@pytest.mark.parametrize("crop,dst,exp", [
    ((0, 0, 0, 0), (0, 0), (1080, 608)),
    ((0, 0, 320, 240), (0, 0), (320, 240)),
    ((0, 0, 1080, 608), (0, 0), (1080, 608)),      # crop == frame: ignored (strictly smaller rule)
    ((0, 0, 1080, 300), (0, 0), (1080, 608)),      # one side not smaller: ignored
    ((0, 0, 320, 240), (720, 480), (720, 480)),
    ((0, 0, 0, 0), (540, 304), (540, 304)),
    ((10, 10, 5, 5), (0, 0), (1080, 608)),          # negative area: ignored
])
def test_output_info_follows_reference_dispatch(lib, crop, dst, exp):
    from oracle import oracle as O
    from tensor_stream_b200 import vpp
    for fourcc, ch in ((vpp.FourCC.RGB24, 3), (vpp.FourCC.Y800, 1), (vpp.FourCC.NV12, 1.5), (vpp.FourCC.UYVY, 2)):
        fp = vpp.FrameParameters(dst[0], dst[1], crop_coords=crop, pixel_format=fourcc)
        w, h, n, es = vpp.output_info(1080, 608, fp)
        assert (w, h) == exp == O.out_geometry(1080, 608, crop, dst)
        assert n == int(ch * w * h) and es == 1
    fp = vpp.FrameParameters(dst[0], dst[1], crop_coords=crop, pixel_format=vpp.FourCC.HSV)
    assert vpp.output_info(1080, 608, fp)[3] == 4


def test_output_shapes_match_reference_from_blob():
    from tensor_stream_b200 import vpp
    # src/Wrappers/WrapperPython.cpp:318-341
    mk = lambda **k: vpp.FrameParameters(256, 128, **k)
    assert vpp.output_shape(1920, 1080, mk())[0] == (128, 256, 3)
    assert vpp.output_shape(1920, 1080, mk(planes_pos=vpp.Planes.PLANAR))[0] == (3, 128, 256)
    assert vpp.output_shape(1920, 1080, mk(pixel_format=vpp.FourCC.Y800))[0] == (1, 128, 256)
    assert vpp.output_shape(1920, 1080, mk(pixel_format=vpp.FourCC.NV12))[0] == (1, 192, 256)
    assert vpp.output_shape(1920, 1080, mk(pixel_format=vpp.FourCC.UYVY))[0] == (1, 256, 256)
    assert vpp.output_shape(1920, 1080, mk(pixel_format=vpp.FourCC.HSV)) == ((128, 256, 3), 4)


def test_div255_reciprocal_trick_is_exact():
    """csrc/vpp_kernels.cu div255(): q=x*r; e=fma(-q,255,x); q+=e*r must equal IEEE x/255.f for
    all 256 inputs (the reference divides, ColorConversion.cu:50)."""
    src = r'''
#include <math.h>
#include <stdio.h>
int main(){ const float r = 0x1.010102p-8f; int bad = 0;
 for (int c = 0; c < 256; c++) { float x = (float)c, q = x * r, e = fmaf(-q, 255.0f, x), v = fmaf(e, r, q);
   if (v != x / 255.0f) bad++; }
 printf("%d\n", bad); return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-o", os.path.join(d, "t"),
                               os.path.join(d, "t.c"), "-lm"])
        assert subprocess.check_output([os.path.join(d, "t")]).strip() == b"0"


def test_area_pattern_examples():
    """generateResizePattern (reference src/Resize.cu:359-386) on the examples of SURVEY.md A.5."""
    from oracle import oracle as O
    np.testing.assert_array_equal(O.area_pattern(1.5), np.array([[1, .5], [.5, 1]], dtype=np.float32))
    assert O.area_pattern(7.5).shape == (2, 8)
    assert O.area_pattern(1080 / 256).shape == (32, 5)
    assert O.area_pattern(2.0).shape == (1, 2)
    p = O.area_pattern(np.float32(1920) / np.float32(299))
    assert p.shape == (299, 7)  # period = dst/gcd(src,dst); rows may drop a surplus tail weight


def test_oracle_edge_cases_do_not_read_out_of_bounds():
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    y = rng.integers(0, 256, (64, 96), dtype=np.uint8)
    uv = rng.integers(0, 256, (32, 96), dtype=np.uint8)
    for rt in (O.NEAREST, O.BILINEAR, O.BICUBIC, O.AREA):
        for dst in ((48, 32), (192, 128), (96, 64), (32, 100)):
            out, (w, h), oob = O.convert(y, uv, dst=dst, resize_type=rt, return_oob=True)
            assert (w, h) == dst and out.size == 3 * w * h and oob == 0, (rt, dst, oob)


def test_oracle_cases_found_by_the_parity_fuzzer():
    """tools/fuzz_parity.py, round 2: (1) NV12 output with an odd width -- the reference's last column stores its V past
    the row / past the allocation; the oracle used to corrupt the heap; (2) an AREA scale whose f32 pattern only closes
    after thousands of rows (2034/508) used to hit the oracle's 8192-row cap."""
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    y = rng.integers(0, 256, (974, 1198), dtype=np.uint8)
    uv = rng.integers(0, 256, (487, 1198), dtype=np.uint8)
    for norm in (False, True):
        out, (w, h) = O.convert(y, uv, dst=(645, 268), resize_type=O.NEAREST, fourcc=O.NV12, normalization=norm)[:2]
        assert (w, h) == (645, 268) and out.size == 645 * 268 * 3 // 2
    y = rng.integers(0, 256, (260, 2034), dtype=np.uint8)
    uv = rng.integers(0, 256, (130, 2034), dtype=np.uint8)
    out, (w, h) = O.convert(y, uv, dst=(508, 64), resize_type=O.AREA, fourcc=O.NV12)[:2]
    assert (w, h) == (508, 64)
    # a pattern as long as the output is all the kernel can index: the long scale agrees with the full pattern on
    # the rows both have
    full = O.area_pattern(np.float32(2034) / np.float32(508), max_rows=60000)
    assert full.shape[0] > 508


def test_stream_sharding_is_a_partition():
    from tensor_stream_b200 import sharding as S
    for world in (1, 2, 4, 8):
        owned = [S.streams_of_rank(19, world, r) for r in range(world)]
        assert sorted(sum(owned, [])) == list(range(19))
        assert all(S.rank_of_stream(s, world) == r for r, ss in enumerate(owned) for s in ss)
    assert S.aggregate_fps(192, 20, 10.0, 8) == 8 * 192 * 20 / 0.010


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tensor_stream_b200 import sharding as S
    mine = S.streams_of_rank(8, world, rank)
    ms = [10.0 + 5.0 * rank, 100.0 - rank]
    q.put((rank, mine, S.max_over_ranks(ms)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_layout_over_gloo():
    """world_size 2 on CPU: each rank owns a disjoint half of the streams and the timing
    reduction returns the max over ranks on every rank (what bench.py --gpus N relies on)."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5, 7]
    assert res[0][2] == res[1][2] == [15.0, 100.0]


def test_no_cpu_fallback_without_gpu():
    """The product path must fail loudly, not fall back, when there is no CUDA device."""
    import torch
    from tensor_stream_b200 import vpp
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        vpp.VideoProcessor(0)
    h = C.c_void_p()
    assert vpp.load_library().tsvpp_create(0, C.byref(h)) != 0  # a cudaError_t, no context


def test_product_code_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under tensor_stream_b200/ or include/ may refer to it."""
    bad = []
    for base in ("tensor_stream_b200", "include"):
        for root, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in root.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cpp", ".h")):
                    txt = open(os.path.join(root, f), errors="ignore").read()
                    if re.search(r"^\s*(from|import)\s+oracle\b|#include\s+\"[^\"]*oracle|liboracle|vpp_oracle\.h", txt, re.M):
                        bad.append(os.path.join(root, f))
    assert not bad, bad


def test_c_abi_header_is_plain_c(lib, tmp_path):
    """include/tsvpp.h must be consumable from C (cgo / JNI / FFI generators): C99, no C++ types."""
    src = tmp_path / "use.c"
    src.write_text('#include "tsvpp.h"\n'
                   "int probe(void) { TsvppParams p = {0,0,0,0, 256,256, TSVPP_BILINEAR, TSVPP_BGR24, TSVPP_PLANAR, 1};\n"
                   "  int w, h, es; size_t n; return tsvpp_output_info(1920, 1080, &p, &w, &h, &n, &es) + (int)n + w + h + es; }\n")
    so = tmp_path / "libuse.so"
    libdir = os.path.join(ROOT, "tensor_stream_b200")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-shared", "-fPIC", "-I",
                           os.path.join(ROOT, "include"), str(src), "-o", str(so), "-L", libdir, "-l:libtsvpp.so",
                           "-Wl,-rpath," + libdir])
    probe = C.CDLL(str(so)).probe
    probe.restype = C.c_int
    assert probe() == 0 + 3 * 256 * 256 + 256 + 256 + 4


def test_numa_binding_uses_the_gpu_node_cpus(tmp_path, monkeypatch):
    """tensor_stream_b200/numa.py: pinned buffers for the zero-copy host path are allocated with the
    thread bound to the GPU's NUMA node; the affinity is restored afterwards and nothing happens when
    the topology is unknown."""
    import os
    from tensor_stream_b200 import numa
    assert numa._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    mine = sorted(os.sched_getaffinity(0))
    pick = mine[: max(1, len(mine) // 2)]
    sysfs = tmp_path
    (sysfs / "bus/pci/devices/0000:d1:00.0").mkdir(parents=True)
    (sysfs / "bus/pci/devices/0000:d1:00.0/numa_node").write_text("1\n")
    (sysfs / "devices/system/node/node1").mkdir(parents=True)
    (sysfs / "devices/system/node/node1/cpulist").write_text(",".join(str(c) for c in pick) + "\n")
    monkeypatch.setattr(numa, "gpu_pci_bus_id", lambda index: "0000:d1:00.0")
    assert numa.gpu_numa_node(0, str(sysfs)) == 1
    with numa.bound_to_gpu(0, str(sysfs)) as node:
        assert node == 1
        assert os.sched_getaffinity(0) == set(pick)
    assert sorted(os.sched_getaffinity(0)) == mine
    # unknown topology (numa_node = -1, as in VMs) -> no-op
    (sysfs / "bus/pci/devices/0000:d1:00.0/numa_node").write_text("-1\n")
    with numa.bound_to_gpu(0, str(sysfs)) as node:
        assert node is None
        assert sorted(os.sched_getaffinity(0)) == mine


def test_numa_node_listing_and_binding(tmp_path):
    import os
    from tensor_stream_b200 import numa
    mine = sorted(os.sched_getaffinity(0))
    sysfs = tmp_path
    for node, cpus in ((0, mine[:1]), (1, mine[-1:]), (2, [])):
        (sysfs / ("devices/system/node/node%d" % node)).mkdir(parents=True)
        (sysfs / ("devices/system/node/node%d/cpulist" % node)).write_text(",".join(str(c) for c in cpus) + "\n")
    (sysfs / "devices/system/node/possible").write_text("0-2\n")  # non-node entries are ignored
    assert numa.all_nodes(str(sysfs)) == [0, 1]                     # node2 has no CPUs
    with numa.bound_to_node(1, str(sysfs)) as node:
        assert node == 1 and os.sched_getaffinity(0) == {mine[-1]}
    with numa.bound_to_node(2, str(sysfs)) as node:                 # nothing to bind to
        assert node is None and sorted(os.sched_getaffinity(0)) == mine
    with numa.bound_to_node(None, str(sysfs)) as node:
        assert node is None
    assert sorted(os.sched_getaffinity(0)) == mine


def test_block_index_division_by_multiply_shift():
    """The direct kernel decodes a block index n into (row pair, column group) with
    __umulhi(n, M) >> (L - 1), M = floor(2^(31+L) / d) + 1, L = ceil(log2 d)  (tsvpp_capi.cpp builds M and
    re-checks it per plan).  Exact for every n < 2^31: e = M*d - 2^(31+L) is in (0, d], so the quotient
    error n*e / (d * 2^(31+L)) stays below 1/d."""
    import random
    rng = random.Random(7)
    ds = list(range(2, 70)) + [120, 255, 256, 257, 480, 960, 1920, 2048, 4095, 4096, 16384, 65535, 65536, 1 << 20]
    for d in ds:
        L = (d - 1).bit_length()
        M = ((1 << (31 + L)) // d) + 1
        assert M < (1 << 32)
        ns = [0, 1, d - 1, d, d + 1, (1 << 31) - 1, ((1 << 31) - 1) // d * d, ((1 << 31) - 1) // d * d - 1]
        ns += [rng.randrange(1 << 31) for _ in range(200)] + [k * d - 1 for k in range(1, 200)] + [k * d for k in range(1, 200)]
        for n in ns:
            if 0 <= n < (1 << 31):
                assert ((n * M) >> 32) >> (L - 1) == n // d, (d, n)


def test_output_element_counts_are_exact_for_large_odd_sizes(lib):
    """channels * w * h is computed in integers: a float product (the reference's (size_t)(channels*w*h)) is
    exact only below 2^24 elements and would round 3*5001*4001 to a neighbouring multiple of 4."""
    from tensor_stream_b200 import vpp
    for fourcc, want in ((vpp.FourCC.RGB24, 3 * 5001 * 4001), (vpp.FourCC.UYVY, 2 * 5001 * 4001),
                         (vpp.FourCC.NV12, 5001 * 4001 + (5001 * 4001) // 2), (vpp.FourCC.Y800, 5001 * 4001)):
        w, h, n, es = vpp.output_info(5001, 4001, vpp.FrameParameters(pixel_format=fourcc))
        assert (w, h, n, es) == (5001, 4001, want, 1)
