#!/usr/bin/env python
"""One encode + decode round trip at a large ERP size (default 4096 x 8192): symbols must come back bit for bit."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from conftest import smooth_images
    from pseudocylindrical_convolution_b200 import pseudo_codec as pc
    from pseudocylindrical_convolution_b200.random_init import synthesize_checkpoints
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    d = "/tmp/pcx_big"
    p_enc, p_dec, p_ent = synthesize_checkpoints(d, "4_56", 56, 0, seed=0)
    enc = pc.PseudoEncoder(56, 0).to(dev)
    dec = pc.PseudoDecoder(56, 0).to(dev)
    pc.load_models(enc, p_enc, p_ent, "cuda:0")
    pc.load_models(dec, p_dec, p_ent, "cuda:0")
    x = torch.from_numpy(smooth_images(1, 3, H, W, seed=3)).to(dev)
    name = os.path.join(d, "big.bin")
    sym = enc.symbols(x)
    t0 = time.perf_counter()
    enc.ent.encode_batch(sym.clone(), [name])
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    got = dec.ent.decode_batch(H // 128, W // 8, [name])
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    ok = bool(torch.equal(got, enc.ent.fill(sym.clone())))
    rec = dec.reconstruct(got)
    torch.cuda.synchronize()
    print("%dx%d: entropy encode %.1f ms, entropy decode %.1f ms, %d bytes, symbols equal: %s, reconstruction finite: %s, peak memory %.1f GB"
          % (H, W, (t1 - t0) * 1e3, (t2 - t1) * 1e3, os.path.getsize(name), ok, bool(torch.isfinite(rec).all()),
             torch.cuda.max_memory_allocated() / 2 ** 30))


if __name__ == "__main__":
    main()
