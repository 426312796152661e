#!/usr/bin/env python
"""Debug aid: decode one stream with engine 1 (step kernel) and engine 2 (dataflow kernel), dump every CDF row (PCX_WAVE_DUMP) and
report the first row where the two differ."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from conftest import smooth_images
    from pseudocylindrical_convolution_b200 import _lib, pseudo_codec as pc
    from pseudocylindrical_convolution_b200.random_init import synthesize_checkpoints
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    d = "/tmp/pcx_flow_debug"
    p_enc, p_dec, p_ent = synthesize_checkpoints(d, "4_56", 56, 0, seed=0)
    enc = pc.PseudoEncoder(56, 0).to(dev)
    dec = pc.PseudoDecoder(56, 0).to(dev)
    pc.load_models(enc, p_enc, p_ent, "cuda:0")
    pc.load_models(dec, p_dec, p_ent, "cuda:0")
    lib = _lib.load()
    H, W = 512, 1024
    xs = torch.from_numpy(smooth_images(1, 3, H, W, seed=5)).to(dev)
    names = [os.path.join(d, "b0.bin")]
    sym = enc.symbols(xs)
    enc.ent.encode_batch(sym.clone(), names)
    out = {}
    for eng in (1, 2):
        lib.pcx_wave_set_fused(eng)
        path = os.path.join(ROOT, "gpurun_out", "dump_e%d.txt" % eng)
        os.environ["PCX_WAVE_DUMP"] = path
        try:
            dec.ent.decode_batch(H // 128, W // 8, names)
            print("engine", eng, "ok")
        except Exception as e:
            print("engine", eng, "failed:", str(e)[:300])
        out[eng] = open(path).read().splitlines()
    a, b = out[1], out[2]
    print("rows", len(a), len(b))
    order, start = dec.ent.ctx2.op[0].order(H // 128, W // 8)
    order = order.cpu().numpy()
    Wc = W // 8
    bad = 0
    for i, (x, y) in enumerate(zip(a, b)):
        if x != y:
            st, k = int(x.split()[1]), int(x.split()[2])
            p0 = max(0, st - 13)
            hw = int(order[int(start[p0]) + k])
            print("first difference at line", i, "cell row %d col %d plane %d tc %d" % (hw // Wc, hw % Wc, hw // Wc + hw % Wc, st - hw // Wc - hw % Wc),
                  "\n  e1:", x, "\n  e2:", y)
            bad += 1
            if bad >= 12:
                break
    if not bad:
        print("common prefix identical")


if __name__ == "__main__":
    main()
