#!/usr/bin/env python
"""Runs the REFERENCE's own Python layer (pseudo_codec.PseudoEncoder / PseudoDecoder, code objects staged by oracle/build_ref.py
in oracle/_ref/refpy - no reference source lives in this repository) in a fresh interpreter, over one of two native back ends:

    --backend ref     the unmodified reference extensions compiled for sm_100 (oracle/_ref/PCONV_ref.so, coder_ref.so):
                      the real reference, end to end
    --backend mirror  this repository's `PCONV` / `coder` mirrors (INTEGRATION.md route A: the drop-in)

Test infrastructure (tests/test_gpu_reference_python.py drives it).  Inputs: checkpoint files in the reference's layout, an
image (.npy, (1,3,512,1024) float32 in [0,1]).  Outputs into --out: enc.bin (the reference encoder's bitstream), sym.npy (the
symbol tensor it coded), and for every --decode NAME=PATH: NAME_rec.npy (reconstruction) and NAME_sym.npy (decoded symbols)."""
import argparse
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_so(name):
    path = os.path.join(ROOT, "oracle", "_ref", name + ".so")
    import torch  # noqa: F401
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _BlobFinder:
    """meta-path finder for the marshalled code objects of oracle/_ref/refpy (`name.pcb`, packages as `name/__init__.pcb`)"""

    def __init__(self, base):
        self.base = base

    def _path(self, fullname):
        rel = fullname.replace(".", os.sep)
        pkg = os.path.join(self.base, rel, "__init__.pcb")
        if os.path.exists(pkg):
            return pkg, True
        mod = os.path.join(self.base, rel + ".pcb")
        return (mod, False) if os.path.exists(mod) else (None, False)

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery
        p, is_pkg = self._path(fullname)
        if p is None:
            return None
        spec = importlib.machinery.ModuleSpec(fullname, self, origin=p, is_package=is_pkg)
        if is_pkg:
            spec.submodule_search_locations = [os.path.dirname(p)]
        return spec

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        import marshal
        with open(module.__spec__.origin, "rb") as f:
            code = marshal.loads(f.read())
        module.__file__ = module.__spec__.origin
        exec(code, module.__dict__)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", choices=["ref", "mirror"], required=True)
    ap.add_argument("--models", required=True, help="directory with 4_56_encoder.pt / _decoder.pt / _ent.pt")
    ap.add_argument("--image", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--decode", action="append", default=[], help="NAME=bitstream path to decode with the reference decoder")
    args = ap.parse_args()
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    if args.backend == "ref":
        sys.modules["PCONV"] = _load_so("PCONV_ref")
        sys.modules["coder"] = _load_so("coder_ref")
    else:
        from pseudocylindrical_convolution_b200 import PCONV as mirror, coder as mycoder
        sys.modules["PCONV"] = mirror
        sys.modules["coder"] = mycoder
    # drift the reference needs absorbed: numpy.lib.function_base is gone in NumPy 2 (the names are unused)
    fb = types.ModuleType("numpy.lib.function_base")
    fb.average, fb.interp = np.average, np.interp
    sys.modules["numpy.lib.function_base"] = fb
    sys.meta_path.insert(0, _BlobFinder(os.path.join(ROOT, "oracle", "_ref", "refpy")))
    import pseudo_codec as ref_pc                      # the reference's module (marshalled code object, no source here)
    assert ref_pc.__file__.endswith(".pcb"), ref_pc.__file__
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    torch.cuda.set_device(0)
    cuda = "cuda:0"
    os.makedirs(args.out, exist_ok=True)
    p = lambda n: os.path.join(args.models, "4_56_%s.pt" % n)
    x = torch.from_numpy(np.load(args.image)).to(cuda)
    enc = ref_pc.PseudoEncoder(56, 0).to(cuda)
    ref_pc.load_models(enc, p("encoder"), p("ent"), cuda)
    with torch.no_grad():
        t = enc.slice(x)
        code = enc.encoder(t)
        _, code_i = enc.quant(code)
        hcode = enc.dtw(enc.ext(code_i))
        np.save(os.path.join(args.out, "sym.npy"), hcode.cpu().numpy())
        np.save(os.path.join(args.out, "latent.npy"), code.cpu().numpy())
    enc(x, os.path.join(args.out, "enc.bin"))
    del enc
    dec = ref_pc.PseudoDecoder(56, 0).to(cuda)
    ref_pc.load_models(dec, p("decoder"), p("ent"), cuda)
    for item in args.decode:
        name, path = item.split("=", 1)
        rec = dec(path)
        np.save(os.path.join(args.out, name + "_rec.npy"), rec.cpu().numpy())
        dec.ent.start(path)
        sym = dec.ent(4, 128)
        np.save(os.path.join(args.out, name + "_sym.npy"), sym.cpu().numpy())
    print("ref_runner ok:", args.backend)


if __name__ == "__main__":
    main()
