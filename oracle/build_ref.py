"""Build the UNMODIFIED reference extensions from /root/reference into oracle/_ref/.

TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is imported by the product package.

Outputs (git-ignored, but they travel to the GPU box with the gpurun snapshot):
  oracle/_ref/PCONV_ref.so  - the reference CUDA extension (extension/*.cu, main.cpp), compiled
                              for sm_100 exactly as SURVEY.md fact 9 describes: same sources,
                              nvcc defaults (-fmad=true), only -std=c++17 and the gencode differ
                              from the reference's setup.py:10-17.
  oracle/_ref/coder_ref.so  - the reference host arithmetic coder (coder/*.cpp).
  oracle/_ref/refpy/        - the reference's Python layer (PCONV_operator/, model_zoo_v2, pseudo_codec) as marshalled
                              code objects (.pcb), compiled from where it lies: tests/test_gpu_reference_python.py runs the reference's
                              own PseudoEncoder / PseudoDecoder on the GPU box against (a) the unmodified PCONV_ref / coder_ref
                              extensions - the real reference end to end - and (b) this repository's `PCONV` / `coder` mirrors
                              (INTEGRATION.md route A).

The sources are compiled where they lie; no reference source is copied into this repository.
The modules are renamed through -DTORCH_EXTENSION_NAME so they can be imported next to the
product's own `PCONV` / `coder` mirrors.
"""
import os
import sys

REF = os.environ.get("PCX_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

EXT_SOURCES = [
    "main.cpp", "math_cuda.cu", "projects_cuda.cu", "dtow_cuda.cu", "context_reshape_cuda.cu",
    "entropy_gmm_cuda.cu", "mask_constrain_cuda.cu", "sphere_slice_cuda.cu", "sphere_uslice_cuda.cu",
    "entropy_gmm_table_cuda.cu", "entropy_context_cuda.cu", "entropy_ctx_pad_run2_cuda.cu",
    "d_extract_cuda_v2.cu", "d_input_cuda_v2.cu", "entropy_conv_cuda_v2.cu", "pseudo_context_cuda.cu",
    "pseudo_pad.cu", "pseudo_fill_cuda.cu", "pseudo_entropy_context_cuda.cu", "pseudo_entropy_pad_cuda.cu",
    "pseudo_quant_cuda.cu", "pseudo_dquant_cuda.cu", "string2class.cc", "entropy_add_cuda.cu",
]
CODER_SOURCES = ["python.cpp", "ArithmeticCoder.cpp", "BitIoStream.cpp"]


PY_MODULES = ["model_zoo_v2.py", "pseudo_codec.py"]


def build_refpy():
    """compile the reference's Python layer into oracle/_ref/refpy as marshalled code objects (`.pcb`; no source is copied, and
    unlike `.pyc` the files are not filtered out of the snapshot that travels to the GPU box).  tests/ref_runner.py imports them
    through a small meta-path finder."""
    import marshal
    dst = os.path.join(OUT, "refpy")
    os.makedirs(os.path.join(dst, "PCONV_operator"), exist_ok=True)
    pairs = [(os.path.join(REF, m), os.path.join(dst, m[:-3] + ".pcb")) for m in PY_MODULES]
    opdir = os.path.join(REF, "PCONV_operator")
    pairs += [(os.path.join(opdir, f), os.path.join(dst, "PCONV_operator", f[:-3] + ".pcb")) for f in sorted(os.listdir(opdir)) if f.endswith(".py")]
    for src, out in pairs:
        with open(src, "r", encoding="utf-8", errors="replace") as f:
            code = compile(f.read(), os.path.relpath(src, REF), "exec", dont_inherit=True)
        with open(out, "wb") as f:
            f.write(marshal.dumps(code))
    with open(os.path.join(dst, "PYTHON_VERSION"), "w") as f:
        f.write("%d.%d" % sys.version_info[:2])
    print(f"[build_ref] {dst} ({len(pairs)} modules)")


def build(which=("coder", "pconv", "py"), verbose=False):
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} not present - keeping whatever is prebuilt in {OUT}")
        return False
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    # this image exports CXX=/opt/gcc/bin/g++, a wrapper that links libstdc++ STATICALLY; a private libstdc++ inside
    # a Python extension crashes in std::stringstream (uninitialised locale) - use the system compiler instead
    for var, exe in (("CXX", "/usr/bin/g++"), ("CC", "/usr/bin/gcc")):
        if os.path.exists(exe):
            os.environ[var] = exe
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT, exist_ok=True)
    if "py" in which:
        build_refpy()
    if "coder" in which:
        bd = os.path.join(OUT, "build_coder")
        os.makedirs(bd, exist_ok=True)
        ce.load(name="coder_ref", sources=[os.path.join(REF, "coder", s) for s in CODER_SOURCES],
                extra_cflags=["-O2", "-std=c++17"], build_directory=bd, is_python_module=False,
                verbose=verbose)
        _publish(bd, "coder_ref")
    if "pconv" in which:
        bd = os.path.join(OUT, "build_pconv")
        os.makedirs(bd, exist_ok=True)
        ce.load(name="PCONV_ref", sources=[os.path.join(REF, "extension", s) for s in EXT_SOURCES],
                extra_include_paths=[os.path.join(REF, "extension")],
                extra_cflags=["-std=c++17", "-DOK"],
                extra_cuda_cflags=["-std=c++17", "-D__CUDA_NO_HALF_OPERATORS__",
                                   "-gencode", "arch=compute_100,code=sm_100"],
                build_directory=bd, is_python_module=False, with_cuda=True, verbose=verbose)
        _publish(bd, "PCONV_ref")
    return True


def _publish(build_dir, name):
    import shutil
    src = os.path.join(build_dir, name + ".so")
    dst = os.path.join(OUT, name + ".so")
    shutil.copyfile(src, dst)
    print(f"[build_ref] {dst}")


if __name__ == "__main__":
    which = tuple(sys.argv[1:]) or ("coder", "pconv", "py")
    build(which, verbose=True)
