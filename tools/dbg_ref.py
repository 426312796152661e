import sys, faulthandler, importlib.util
faulthandler.enable()
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, numpy as np
spec = importlib.util.spec_from_file_location("PCONV_ref", "oracle/_ref/PCONV_ref.so")
R = importlib.util.module_from_spec(spec); spec.loader.exec_module(R)
W64 = [15, 31, 54, 63, 63, 64, 64, 64, 64, 64, 64, 63, 63, 54, 31, 15]
WEIGHT = [float(v) for v in W64]
dev = torch.device('cuda:0')
x = torch.randn(16, 3, 4, 128, device=dev)
pr = R.PseudoContextOp(16, 20, WEIGHT, 0, False)
print("ctx addr", pr.addr(), flush=True)
h = pr.produce_fill_param(4, 128)
print("fill param", h, flush=True)
op = R.PseudoFillOp(0, 16, 0, 0, pr.addr(), 0, 0, False)
print("fill op built", flush=True)
y = op.forward(x.clone())
print("fill ok", y[0].shape, flush=True)
pad = R.PseudoPadOp(1, 16, pr.addr(), 0, False)
print("pad built", flush=True)
z = pad.forward(y[0])
torch.cuda.synchronize()
print("pad ok", z[0].shape, flush=True)
