"""The drop-in boundary: include/pcx.h <-> libpcx.so <-> the ctypes table in _lib.py.  CPU only: the library must
load without a GPU, export every symbol the header declares (and nothing else), and fail loudly - not fall
back - when a compute entry point is called without a device."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "pcx.h")


def _declared():
    """name -> number of parameters, parsed from the header's prototypes."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"typedef struct pcx_conv_desc \{.*?\} pcx_conv_desc;", " ", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(pcx_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


@pytest.fixture(scope="module")
def lib():
    from pseudocylindrical_convolution_b200 import _lib, build
    build.build_library()
    return _lib


def test_header_parses_to_a_non_trivial_abi():
    d = _declared()
    assert len(d) >= 40
    for must in ("pcx_slice_fwd", "pcx_pad_fwd", "pcx_conv2d_fwd", "pcx_ctx_conv_step", "pcx_gmm_table", "pcx_coder_encodes"):
        assert must in d


def test_every_declared_symbol_is_bound_and_exported(lib):
    declared = _declared()
    table = lib.PROTOTYPES
    assert set(declared) == set(table), (sorted(set(declared) - set(table)), sorted(set(table) - set(declared)))
    handle = lib.load()
    for name, nargs in declared.items():
        assert hasattr(handle, name), name
        assert len(table[name][1]) == nargs, "%s: header has %d parameters, ctypes table %d" % (name, nargs, len(table[name][1]))


def test_library_exports_only_the_abi(lib):
    """-fvisibility=hidden + the header's visibility push: no C++ internals leak into the dynamic symbol table."""
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    names = {ln.split()[-1] for ln in out.splitlines() if ln.strip() and ln.split()[-2] in ("T", "t")}
    names = {n for n in names if not n.startswith("_") or n.startswith("_Z")}
    extra = {n for n in names if not n.startswith("pcx_")}
    assert not extra, sorted(extra)[:10]
    assert set(_declared()) <= names


def test_no_torch_types_in_the_boundary():
    src = open(HEADER).read()
    for banned in ("at::", "torch::", "Tensor", "#include <torch", "c10::"):
        assert banned not in src


def test_conv_desc_layout_matches_header(lib):
    fields = [f for f, _ in lib.ConvDesc._fields_]
    src = open(HEADER).read()
    body = re.search(r"typedef struct pcx_conv_desc \{(.*?)\} pcx_conv_desc;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", " ", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        assert decl.startswith("int "), decl
        names += [re.sub(r"\[.*\]", "", t).strip() for t in decl[4:].split(",")]
    assert names == fields
    assert C.sizeof(lib.ConvDesc) == 4 * (len(fields) - 1) + 4 * lib.PCX_MAX_PART


def test_host_only_entry_points_work_without_a_gpu(lib):
    h = lib.load()
    assert h.pcx_abi_version() >= 1
    w = (C.c_float * 16)(*[15, 31, 54, 63, 63, 64, 64, 64, 64, 64, 64, 63, 63, 54, 31, 15])
    out = (C.c_int * 16)()
    assert h.pcx_band_widths(w, 16, 512, 1024, out) == 0
    assert list(out) == [240, 496, 864, 1008, 1008, 1024, 1024, 1024, 1024, 1024, 1024, 1008, 1008, 864, 496, 240]


def test_compute_fails_loudly_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lib.PcxError):
        lib.call("pcx_device_check", 0, None, None)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import or dlopen it."""
    pkg = os.path.join(ROOT, "pseudocylindrical_convolution_b200")
    bad = []
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "libpcx_oracle" in txt or "oracle/_ref" in txt:
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_engine_options_round_trip_without_a_device(lib):
    """pcx_wave_set_option is plain host state: it returns the previous value and rejects unknown names loudly."""
    handle = lib.load()
    for name, default in ((b"slabs", 0), (b"tsplit", 0), (b"smem", 1)):
        prev = handle.pcx_wave_set_option(name, 5)
        assert prev == default, (name, prev)
        assert handle.pcx_wave_set_option(name, prev) == 5
    assert handle.pcx_wave_set_option(b"no_such_option", 1) < 0
    assert b"no_such_option" in handle.pcx_last_error()
