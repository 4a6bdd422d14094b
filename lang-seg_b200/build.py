"""Build liblseg_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The library is a single translation unit (csrc/lseg_b200.cu) linked against the static CUDA runtime
only; the driver API entry point it needs (cuTensorMapEncodeTiled) is resolved at run time, so the
.so loads on machines without libcuda (the C-ABI export test relies on that).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblseg_b200.so")
SOURCES = ["lseg_b200.cu"]
HEADERS = ["common.cuh", "gemm_tc.cuh", "mhsa.cuh", "mhsa2.cuh", "mhsa3.cuh", "text_attn.cuh", "elementwise.cuh", "engine.cuh"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(HERE, "..", "include", "lseg_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [
        nvcc_path(),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-O3", "-std=c++17", "-lineinfo",
        "-shared", "-Xcompiler", "-fPIC",
        "-cudart", "static",
        "-o", LIB,
    ]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building liblseg_b200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
