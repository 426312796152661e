#!/usr/bin/env python
"""Debug aid: repeat an entropy decode with the dataflow engine and print the outcome (error text carries image / step / row)."""
import os
import sys
import time

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
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    H, W = 512, 1024
    xs = torch.from_numpy(smooth_images(N, 3, H, W, seed=5)).to(dev)
    names = [os.path.join(d, "b%d.bin" % i) for i in range(N)]
    sym = enc.symbols(xs)
    enc.ent.encode_batch(sym.clone(), names)
    want = enc.ent.fill(sym.clone())
    lib.pcx_wave_set_fused(2)
    for it in range(reps):
        t0 = time.perf_counter()
        try:
            got = dec.ent.decode_batch(H // 128, W // 8, names)
            torch.cuda.synchronize()
            print("try %d: %.2f ms, symbols equal: %s" % (it, (time.perf_counter() - t0) * 1e3, bool(torch.equal(got, want))), flush=True)
            if not torch.equal(got, want):
                bad = (got != want).nonzero()
                print("  mismatches:", bad.shape[0], "of", got.numel())
                for r in bad[:12].tolist():
                    n, g, y, x = r
                    band = n % 16
                    plane = band * (H // 128) + y + x
                    print("   plane-tensor idx", r, "band", band, "step", plane + g, "got", float(got[n, g, y, x]), "want", float(want[n, g, y, x]))
                steps = sorted(set(int((r[0] % 16) * (H // 128) + r[2] + r[3] + r[1]) for r in bad.tolist()))
                print("   steps with mismatches:", steps[:40], "...", len(steps))
        except Exception as e:
            print("try %d failed: %s" % (it, str(e)[-220:]), flush=True)


if __name__ == "__main__":
    main()
