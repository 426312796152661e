#!/usr/bin/env python
"""Runs the REFERENCE's own Python layer (pseudo_codec.PseudoEncoder / PseudoDecoder, byte code staged by oracle/build_ref.py
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
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "refpy"))
    import pseudo_codec as ref_pc                      # the reference's module (sourceless byte code)
    assert ref_pc.__file__.endswith(".pyc"), ref_pc.__file__
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
