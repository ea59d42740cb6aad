"""In-tree build of the native library (nvcc, sm_100a only).

  python -m tensor_stream_b200.build        # builds tensor_stream_b200/libtsvpp.so

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc
cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtsvpp.so")

SOURCES = ["vpp_kernels.cu", "vpp_march_kernels.cu", "vpp_yuv_kernels.cu", "tsvpp_capi.cpp", "VideoProcessor.cpp"]
HEADERS = ["vpp_plan.h", "vpp_launch.h", "vpp_device.cuh", "vpp_march.h", "march_plan.inl", "../../include/tsvpp.h", "../../include/VideoProcessor.h",
           "../../include/tsvpp_avframe.h"]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall",
    "-I", os.path.join(HERE, "..", "include"),
]


# The C++ drop-in (VideoProcessor.cpp) must see the AVFrame its callers see: point TSVPP_FFMPEG_INCLUDE at the FFmpeg
# headers the wrapper is built with (include/tsvpp_avframe.h; without them a stand-in struct is compiled in and
# VideoProcessor::Init refuses callers with another layout).
if os.environ.get("TSVPP_FFMPEG_INCLUDE"):
    NVCC_FLAGS += ["-I", os.environ["TSVPP_FFMPEG_INCLUDE"], "-DTSVPP_HAVE_FFMPEG"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defs=(), out=None):
    """defs / out: developer A/B builds (extra -D flags into another .so, see tools/ab_build.sh)."""
    global LIB
    if out:
        LIB, force = out, True
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    if not force and not _stale(LIB, deps):
        return LIB
    objs, procs = [], []
    for s in srcs:  # the translation units compile in parallel
        o = os.path.join(CSRC, os.path.basename(s) + ".o")
        if out:
            o = os.path.join("/tmp", os.path.basename(out) + "." + os.path.basename(s) + ".o")
        cmd = ["nvcc"] + NVCC_FLAGS + list(defs) + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        if s.endswith(".cpp"):
            cmd.insert(1, "-x")
            cmd.insert(2, "cu")
        procs.append((cmd, subprocess.Popen(cmd)))
        objs.append(o)
    for cmd, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs)
    for o in objs:
        os.remove(o)
    return LIB


if __name__ == "__main__":
    _defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    _out = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defs=_defs, out=_out[0] if _out else None))
