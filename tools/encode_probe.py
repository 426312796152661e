#!/usr/bin/env python
"""Where the entropy-encode stage of a batch spends its time: Python set-up, the native call, closing the bitstreams."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from conftest import smooth_images
    from pseudocylindrical_convolution_b200 import _lib, coder, pseudo_codec as pc
    from pseudocylindrical_convolution_b200.pseudo_codec import restart_entropy_network
    from pseudocylindrical_convolution_b200.random_init import synthesize_checkpoints
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    d = "/tmp/pcx_encode_probe"
    p_enc, p_dec, p_ent = synthesize_checkpoints(d, "4_56", 56, 0, seed=0)
    enc = pc.PseudoEncoder(56, 0).to(dev)
    pc.load_models(enc, p_enc, p_ent, "cuda:0")
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    H, W = 512, 1024
    xs = torch.from_numpy(smooth_images(N, 3, H, W, seed=5)).to(dev)
    names = [os.path.join(d, "b%d.bin" % i) for i in range(N)]
    sym = enc.symbols(xs)
    ent = enc.ent
    for it in range(5):
        torch.cuda.synchronize()
        t = [time.perf_counter()]
        with torch.no_grad():
            ent.apply(restart_entropy_network)
            data = ent.fill(sym.clone())
            ent.ctx2.setup_context(data.shape[3])
            torch.cuda.synchronize(); t.append(time.perf_counter())
            coders = [coder.coder(n) for n in names]
            for c in coders:
                c.start_encoder()
            t.append(time.perf_counter())
            ent.engine().encode(data, coders)
            t.append(time.perf_counter())
            torch.cuda.synchronize(); t.append(time.perf_counter())
            for c in coders:
                c.end_encoder()
            t.append(time.perf_counter())
        ms = [(b - a) * 1e3 for a, b in zip(t, t[1:])]
        print("pass %d: setup %.2f | coders %.2f | native call %.2f | sync %.2f | close %.2f | total %.2f ms" % (it, *ms, (t[-1] - t[0]) * 1e3), flush=True)


if __name__ == "__main__":
    main()
