"""Builds the two in-tree shared libraries (no JIT cache, so they travel to the GPU box):

  token_hawk_b200/lib/libthk_sm100a.so   hand-written sm_100a kernels + C ABI (include/thk_cabi.h)
  token_hawk_b200/lib/libth_b200.so      host C++: th:: op surface, LLaMA graph, ggjt loader, capi_*

nvcc cross-compiles for sm_100a without a GPU.  Usage: python -m token_hawk_b200.build [--force]
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
# A/B variants: THK_VARIANT=<name> builds into lib_<name>/ with the extra nvcc flags in THK_DEFINES ("-DX -DY=2");
# load one with THK_LIBDIR=lib_<name> (token_hawk_b200/__init__.py)
_VAR = os.environ.get("THK_VARIANT", "")
LIB = os.path.join(PKG, "lib" + ("_" + _VAR if _VAR else ""))
INC = os.path.join(ROOT, "include")

CU_SOURCES = ["context.cu", "ops.cu", "decoder.cu", "gemm_tc.cu"]
HOST_SOURCES = ["host/th.cpp", "host/th_llama.cpp", "host/th_llama_loader.cpp", "host/capi.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I" + INC, "-I" + CSRC] + os.environ.get("THK_DEFINES", "").split()
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
CXX_FLAGS = ["-std=c++17", "-O2", "-fPIC", "-Wall", "-Wextra", "-I" + INC]


def _nvcc():
    for c in ("/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _headers():
    hs = [os.path.join(INC, "thk_cabi.h"), os.path.join(CSRC, "common.cuh")]
    hs += [os.path.join(INC, "th", f) for f in os.listdir(os.path.join(INC, "th"))]
    return hs


def build(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    objdir = os.path.join(LIB, "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    cu = [s for s in CU_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []
    for s in cu:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace("/", "_") + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + _headers()):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    kern = os.path.join(LIB, "libthk_sm100a.so")
    if force or _stale(kern, objs):
        subprocess.check_call([nvcc, "-shared", "-o", kern] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    hobjs = []
    for s in HOST_SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace("/", "_") + ".o")
        hobjs.append(obj)
        if force or _stale(obj, [src] + _headers()):
            subprocess.check_call([CXX] + CXX_FLAGS + ["-c", src, "-o", obj])
    host = os.path.join(LIB, "libth_b200.so")
    if force or _stale(host, hobjs + [kern]):
        subprocess.check_call([CXX, "-shared", "-o", host] + hobjs +
                              ["-L" + LIB, "-lthk_sm100a", "-Wl,-rpath,$ORIGIN"])
    return kern, host


if __name__ == "__main__":
    k, h = build(force="--force" in sys.argv, verbose=True)
    print("built", k, "and", h)
