import sys, faulthandler, importlib.util
faulthandler.enable()
import torch
spec = importlib.util.spec_from_file_location("PCONV_ref", "oracle/_ref/PCONV_ref.so")
R = importlib.util.module_from_spec(spec); spec.loader.exec_module(R)
WEIGHT = [float(v) for v in [15, 31, 54, 63, 63, 64, 64, 64, 64, 64, 64, 63, 63, 54, 31, 15]]
torch.zeros(1, device='cuda:0')
for name, args in [("DtowOp", (2, True, 0, False)), ("SphereSliceOp", (16, 0, 0, WEIGHT, 0, False)),
                   ("EntropyGmmTableOp", (8, 3.5, 3, 65536.0, 1e-6, 0, False)),
                   ("EntropyContextOp", (16, 18, WEIGHT, 0, False)),
                   ("PseudoEntropyContextOp", (16, 20, 1, WEIGHT, 0, False)),
                   ("PseudoContextOp", (16, 20, WEIGHT, 0, False))]:
    print("constructing", name, flush=True)
    o = getattr(R, name)(*args)
    print("  ok", flush=True)
    if hasattr(o, "addr"):
        print("  addr", o.addr(), flush=True)
