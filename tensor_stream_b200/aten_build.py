"""In-tree build of the ATen hand-off extension (tensor_stream_b200/_tsvpp_torch.so).

  python -m tensor_stream_b200.aten_build

Links against tensor_stream_b200/libtsvpp.so (rpath $ORIGIN), so build that first.  No CUDA code in
this extension: it only needs the CUDA runtime headers for c10/cuda.
"""
import glob
import importlib.util
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
NAME = "_tsvpp_torch"


def so_path():
    hits = glob.glob(os.path.join(HERE, NAME + "*.so"))
    return hits[0] if hits else None


def build(force=False):
    from tensor_stream_b200 import build as B
    B.build()
    src = os.path.join(HERE, "csrc", "aten_handoff.cpp")
    cur = so_path()
    if cur and not force and os.path.getmtime(cur) > max(os.path.getmtime(src), os.path.getmtime(B.LIB)):
        return cur
    from torch.utils import cpp_extension as CE
    bdir = os.path.join(HERE, "build", NAME)
    os.makedirs(bdir, exist_ok=True)
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    built = os.path.join(bdir, NAME + ".so")
    try:
        # rpath: $ORIGIN for the in-tree copy (next to libtsvpp.so), the absolute directory for the
        # copy in build/ that torch imports right after linking
        CE.load(name=NAME, sources=[src], build_directory=bdir, verbose=False, is_python_module=False,
                extra_include_paths=[os.path.join(HERE, "..", "include"), cuda_inc],
                extra_cflags=["-O2", "-std=c++17"],
                extra_ldflags=["-L" + HERE, "-l:libtsvpp.so", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + HERE,
                               "-L" + os.path.join(os.path.dirname(CE.__file__), "..", "lib"), "-lc10_cuda"])
    except OSError:
        if not os.path.exists(built):
            raise
    dst = os.path.join(HERE, NAME + ".so")
    shutil.copy2(built, dst)
    return dst


def load():
    """Import the extension from its in-tree .so (fails loudly when it was not built)."""
    p = so_path()
    if not p:
        raise ImportError("%s.so missing: run `python -m tensor_stream_b200.aten_build`" % NAME)
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
