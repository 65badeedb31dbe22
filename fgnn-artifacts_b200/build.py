"""In-tree build of the native code (no GPU needed: nvcc cross-compiles sm_100a).

  python fgnn-artifacts_b200/build.py [--force]

Outputs (git-ignored, shipped to the GPU box by gpurun):
  fgnn-artifacts_b200/lib/libfgnn_kernels.so    kernels + kernel C-ABI (include/fgnn_kernels.h)
  fgnn-artifacts_b200/samgraph/torch/c_lib.so   host runtime + samgraph_* C-ABI + Python module
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "lib")
INCLUDE = os.path.join(ROOT, "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
                     "-ccbin", HOST_CXX, "-I", INCLUDE, "-I", os.path.join(CSRC, "kernels"),
                     "-I", os.path.join(CSRC, "runtime")]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _headers():
    hs = []
    for base in (INCLUDE, CSRC):
        for dp, _, fs in os.walk(base):
            hs += [os.path.join(dp, f) for f in fs if f.endswith((".h", ".cuh", ".hpp"))]
    return hs


def _compile(src, force, extra=()):
    rel = os.path.relpath(src, CSRC).replace(os.sep, "_")
    obj = os.path.join(OBJ, rel + ".o")
    if force or _newer(obj, [src] + _headers()):
        cmd = [NVCC] + NVCC_FLAGS + list(extra) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj


def _sources(sub, exts):
    d = os.path.join(CSRC, sub)
    if not os.path.isdir(d):
        return []
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIB, exist_ok=True)
    kernel_srcs = _sources("kernels", (".cu",))
    runtime_srcs = _sources("runtime", (".cc", ".cu"))
    py_inc = sysconfig.get_paths()["include"]
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        kobjs = list(ex.map(lambda s: _compile(s, force), kernel_srcs))
        robjs = list(ex.map(lambda s: _compile(s, force, ("-I", py_inc, "-x", "cu")), runtime_srcs))

    outs = []
    klib = os.path.join(LIB, "libfgnn_kernels.so")
    if force or _newer(klib, kobjs):
        subprocess.check_call([NVCC] + ARCH + ["-shared", "-cudart", "static", "-ccbin", HOST_CXX,
                                               "-o", klib] + kobjs)
    outs.append(klib)

    if robjs:
        ext = os.path.join(HERE, "samgraph", "torch", "c_lib.so")
        if force or _newer(ext, kobjs + robjs):
            subprocess.check_call([NVCC] + ARCH + ["-shared", "-cudart", "static", "-ccbin", HOST_CXX,
                                                   "-Xlinker", "--version-script=" + os.path.join(CSRC, "c_lib_exports.map"),
                                                   "-o", ext] + kobjs + robjs + ["-lpthread", "-lrt"])
        outs.append(ext)
    if verbose:
        for o in outs:
            print("built", os.path.relpath(o, ROOT))
    return outs


if __name__ == "__main__":
    build(force="--force" in sys.argv)
