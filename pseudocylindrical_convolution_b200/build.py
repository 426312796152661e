"""Builds libpcx.so (hand-written CUDA for sm_100a + the host range coder) in-tree with nvcc.

The shared library is the product's only native artefact; it is git-ignored but travels to the GPU box
with the repository snapshot.  `python -m pseudocylindrical_convolution_b200.build` rebuilds it.
"""
import concurrent.futures
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(PKG, "libpcx.so")

SOURCES = ["pcx_geometry.cu", "pcx_tile.cu", "pcx_tile_nhwc.cu", "pcx_quant.cu", "pcx_dense.cu", "pcx_conv_tc.cu", "pcx_nhwc_ops.cu", "pcx_ctx.cu", "pcx_flow.cu", "pcx_metrics.cu",
           "pcx_coder.cpp"]
HEADERS = [os.path.join(CSRC, "pcx_common.cuh"), os.path.join(CSRC, "pcx_ctx_step.cuh"), os.path.join(CSRC, "pcx_flow.h"),
           os.path.join(PKG, "..", "include", "pcx.h")]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall", "--expt-relaxed-constexpr"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src):
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    path = os.path.join(CSRC, src)
    if _stale(obj, [path] + HEADERS):
        cmd = [NVCC] + NVCC_FLAGS + (["-x", "cu"] if src.endswith(".cpp") else []) + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj


def build_library(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(_compile, SOURCES))
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print("[pcx] built", LIB)
    return LIB


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose=True)
