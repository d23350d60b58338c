"""Builds csrc/*.cu into lib/libmaxk_b200.so (sm_100a only, in-tree).

The shared library is plain CUDA runtime + the C ABI of include/maxk_b200.h; it has no
torch or Python dependency, so the same file serves ctypes (this package), cgo, JNI ...
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libmaxk_b200.so")
SOURCES = ["topk.cu", "spgemm_fwd.cu", "sspmm_bwd.cu", "meta.cu", "plan.cu", "wide.cu"]
HEADERS = [os.path.join(CSRC, "maxk_common.cuh"), os.path.join(CSRC, "slots.cuh"), os.path.join(CSRC, "recip31_table.inc"),
           os.path.join(os.path.dirname(HERE), "include", "maxk_b200.h")]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-shared", "--use_fast_math=false",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile if sources are newer than the library. Returns the library path."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    cmd = [_nvcc()] + flags + ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports a CC wrapper that nvcc must not pick up
    env.pop("CXX", None)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libmaxk_b200.so (exit %d)" % res.returncode)
    with open(os.path.join(LIB_DIR, "ptxas_info.txt"), "w") as f:
        f.write(res.stdout)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
