#!/usr/bin/env python
"""Achieved HBM bandwidth of the memory-bound operator kernels behind the PCONV classes (NCHW, reference layout), each at a
config-scale shape, against the measured copy bandwidth of MEASURED_PEAKS.json.  Algorithmic bytes per SURVEY.md 8(d).
One JSON line per operator; `--only NAME` runs a single operator (for ncu).

    python tools/hbm_ops_bench.py [--only pad1] [--reps 20]
"""
import argparse, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pseudocylindrical_convolution_b200 import PCONV

W64 = [15, 31, 54, 63, 63, 64, 64, 64, 64, 64, 64, 63, 63, 54, 31, 15]
WEIGHT = [float(v) for v in W64]


def widths(W):
    return [int(float(np.float32(np.float32(w) / np.float32(64) * np.float32(W))) + 0.5) for w in W64]


def timed(fn, reps, flush):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.add_(1.0)                        # 512 MB write: evicts the 126 MB L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--reps", type=int, default=15)
    a = ap.parse_args()
    dev = torch.device("cuda:0"); torch.cuda.set_device(0)
    peak = 6536.7
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    flush = torch.zeros(128 * 1024 * 1024, dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev); g.manual_seed(0)
    ctx = PCONV.PseudoContextOp(16, 20, WEIGHT, 0, False)
    ectx = PCONV.PseudoEntropyContextOp(16, 20, 1, WEIGHT, 0, False)
    cases = []

    # tile tensors at the first scale of a 2048x4096 image: (16, 192, 64, 2048)
    NB, C, h, W = 16, 192, 64, 2048
    wl = widths(W)
    valid = sum(h * w for w in wl)
    tiles = torch.rand((NB, C, h, W), generator=g, device=dev)
    fill0 = PCONV.PseudoFillOp(0, 16, 0, 0, ctx.addr(), 0, 0, False)
    fill0.forward(tiles)
    for p in (1, 2):
        op = PCONV.PseudoPadOp(p, 16, ctx.addr(), 0, False)
        # algorithmic: read the valid interior + write the whole padded tile (zero fill included), per (band, channel)
        by = 4.0 * C * (valid + 16 * (h + 2 * p) * (W + 2 * p))
        cases.append(("pad%d" % p, "PseudoPadV2 pad=%d (16,192,64,2048)" % p, lambda op=op: op.forward(tiles), by))
    by = 4.0 * C * (16 * h * W - valid)
    cases.append(("fill", "PseudoFillV2 in place (16,192,64,2048): stores to the invalid columns only", lambda: fill0.forward(tiles), by))
    ep = PCONV.PseudoEntropyPadOp(2, 16, ectx.addr(), 0, False)
    cases.append(("entropy_pad2", "PseudoEntropyPad pad=2 (16,192,64,2048)", lambda: ep.forward(tiles),
                  4.0 * C * (valid + 16 * (h + 4) * (W + 4))))
    # slice / uslice of a 192-channel 1024x2048 ERP tensor
    erp = torch.rand((1, 192, 1024, 2048), generator=g, device=dev)
    sl = PCONV.SphereSliceOp(16, 0, 0, WEIGHT, 0, False)
    us = PCONV.SphereUsliceOp(16, 0, 0, WEIGHT, 0, False)
    t2 = sl.forward(erp)[0]
    cases.append(("slice", "SphereSlice (1,192,1024,2048) -> (16,192,64,2048)", lambda: sl.forward(erp), 8.0 * erp.numel()))
    cases.append(("uslice", "SphereUslice (16,192,64,2048) -> (1,192,1024,2048)", lambda: us.forward(t2), 8.0 * erp.numel()))
    # quant / dquant / dtow at code resolution of a batch of 8 x 4096x8192 images: (128, 192, 16, 512)
    code = torch.rand((128, 192, 16, 512), generator=g, device=dev)
    qw = torch.zeros((192, 8), device=dev); qw[:, 0] = 0.05; qw[:, 1:] = float(np.log(0.12))
    cnt = torch.zeros((192, 8), device=dev)
    q = PCONV.PseudoQuantOp(192, 8, 16, 0.0, 1, 2, 0.0, ctx.addr(), 0, False)
    dq = PCONV.PseudoDQuantOp(16, 192, 8, ctx.addr(), 0, False)
    symt = q.forward(code, qw, cnt)[1]
    cases.append(("quant", "PseudoQUANTV2 (128,192,16,512): value + symbol out", lambda: q.forward(code, qw, cnt), 12.0 * code.numel()))
    cases.append(("dquant", "PseudoDQUANT (128,192,16,512)", lambda: dq.forward(symt, qw), 8.0 * code.numel()))
    d2w = PCONV.DtowOp(2, True, 0, False)
    cases.append(("dtow", "Dtow d2w (128,192,16,512) -> (128,48,32,1024)", lambda: d2w.forward(code), 8.0 * code.numel()))
    # GMM CDF tables for 8 M symbols (one launch): 9 fp32 in + 9 fp32 out per symbol
    n = 8 * 1024 * 1024
    data = torch.randn((3, 3, n // 1024, 1024), generator=g, device=dev)
    gm = PCONV.EntropyGmmTableOp(8, 3.5, 3, 65536.0, 1e-6, 0, False)
    tn = torch.tensor([n], dtype=torch.int32)
    cases.append(("gmm_table", "EntropyBatchGmmTable, 8 Mi symbols", lambda: gm.forward_batch(data, tn), 72.0 * n))

    for name, desc, fn, by in cases:
        if a.only and a.only != name:
            continue
        ms = timed(fn, a.reps, flush)
        gbs = by / ms / 1e6
        print(json.dumps({"op": name, "what": desc, "ms": round(ms, 4), "algorithmic_bytes": by, "GB/s": round(gbs, 1),
                          "peak_GB/s": peak, "frac": round(gbs / peak, 3)}), flush=True)


if __name__ == "__main__":
    main()
