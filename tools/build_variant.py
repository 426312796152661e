#!/usr/bin/env python
"""Build a variant of libpcx.so with extra -D flags on pcx_flow.cu (A/B experiments on the decoder kernel).

usage: tools/build_variant.py NAME -DFOO=1 [-DBAR=2 ...]   ->  pseudocylindrical_convolution_b200/_variants/libpcx_NAME.so
       run with PCX_LIB=<that path>
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pseudocylindrical_convolution_b200 import build as B  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    B.build_library()
    out_dir = os.path.join(ROOT, "pseudocylindrical_convolution_b200", "_variants")
    os.makedirs(out_dir, exist_ok=True)
    obj = os.path.join(out_dir, "pcx_flow_%s.o" % name)
    subprocess.run([B.NVCC] + B.NVCC_FLAGS + flags + ["-c", os.path.join(B.CSRC, "pcx_flow.cu"), "-o", obj], check=True)
    objs = [os.path.join(B.OBJ, os.path.splitext(s)[0] + ".o") for s in B.SOURCES if s != "pcx_flow.cu"] + [obj]
    lib = os.path.join(out_dir, "libpcx_%s.so" % name)
    subprocess.run([B.NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"], check=True)
    os.remove(obj)
    print(lib)


if __name__ == "__main__":
    main()
