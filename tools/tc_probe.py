"""Runs one tensor-core convolution against the direct kernel (debug helper)."""
import ctypes as C, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from pseudocylindrical_convolution_b200._lib import call, ConvDesc
Ci, Co, k, s, h, W = [int(v) for v in (sys.argv[1:7] if len(sys.argv) > 6 else (96, 96, 3, 1, 8, 128))]
rng = np.random.default_rng(0)
halo = 2 if k == 3 else 0
Hi, Wi = h + halo, W + halo
pitch = (Wi + 3) // 4 * 4
ho, wo = (Hi - k) // s + 1, (Wi - k) // s + 1
opitch = (wo + 3) // 4 * 4
x = np.zeros((16, Ci, Hi, pitch), np.float32); x[..., :Wi] = rng.standard_normal((16, Ci, Hi, Wi))
w = (rng.standard_normal((Co, Ci, k, k)) / np.sqrt(Ci * k * k)).astype(np.float32)
dev = torch.device('cuda:0')
dx, dw = torch.from_numpy(x).to(dev), torch.from_numpy(w).to(dev)
outs = []
for impl in (1, 0):
    d = ConvDesc()
    d.N, d.npart, d.Ci, d.Hi, d.in_pitch = 1, 16, Ci, Hi, pitch
    d.Co, d.Ho, d.Wo, d.out_rows, d.out_pitch, d.out_y0, d.out_x0 = Co, ho, wo, ho, opitch, 0, 0
    d.k, d.stride, d.act, d.impl = k, s, 0, impl
    d.aux_rows, d.aux_pitch, d.aux_y0, d.aux_x0 = ho, opitch, 0, 0
    for g in range(16): d.wl_out[g] = wo
    y = torch.zeros((16, Co, ho, opitch), device=dev)
    call("pcx_conv2d_fwd", C.byref(d), C.c_void_p(dx.data_ptr()), C.c_void_p(dw.data_ptr()), None, None, None, None, C.c_void_p(y.data_ptr()), None)
    torch.cuda.synchronize()
    outs.append(y.cpu().numpy())
    print("impl", impl, "done", flush=True)
err = np.abs(outs[0] - outs[1])
print("max err", err.max(), "rms", np.sqrt((err**2).mean()), "ref rms", np.sqrt((outs[0]**2).mean()))
