#!/usr/bin/env python
"""Entropy decode (wavefront + host range decoder, no transforms) per engine: 2 = persistent dataflow kernel (pcx_flow.cu),
1 = one cooperative launch per step.  One JSON line per (engine, batch, size); symbols are checked against the encoder's.
PCX_WAVE_TRACE=<file> additionally dumps the per-step trace of the last decode."""
import json
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
    d = "/tmp/pcx_decode_bench"
    p_enc, p_dec, p_ent = synthesize_checkpoints(d, "4_56", 56, 0, seed=0)
    enc = pc.PseudoEncoder(56, 0).to(dev)
    dec = pc.PseudoDecoder(56, 0).to(dev)
    pc.load_models(enc, p_enc, p_ent, "cuda:0")
    pc.load_models(dec, p_dec, p_ent, "cuda:0")
    lib = _lib.load()
    cases = [(1, 512, 1024), (8, 512, 1024), (16, 512, 1024), (1, 1024, 2048), (1, 2048, 4096), (2, 2048, 4096)]
    engines = [int(a) for a in sys.argv[1:] if a.isdigit()] or [2, 1]
    for N, H, W in cases:
        xs = torch.from_numpy(smooth_images(N, 3, H, W, seed=5)).to(dev)
        names = [os.path.join(d, "b%d.bin" % i) for i in range(N)]
        sym = enc.symbols(xs)
        enc.ent.encode_batch(sym.clone(), names)
        want = enc.ent.fill(sym.clone())
        nsym = int(sum(os.path.getsize(n) for n in names))
        for eng in engines:
            lib.pcx_wave_set_fused(eng)
            ts = []
            ok = True
            for it in range(4):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                got = dec.ent.decode_batch(H // 128, W // 8, names)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
                ok = ok and bool(torch.equal(got, want))
            ts = sorted(ts[1:])
            steps = H // 8 + W // 8 + 14 - 2
            print(json.dumps({"stage": "entropy decode", "engine": eng, "batch": N, "H": H, "W": W, "p50_ms": ts[1] * 1e3,
                              "us_per_step": ts[1] * 1e6 / steps, "MP/s": N * H * W / 1e6 / ts[1], "symbols_ok": ok,
                              "stream_bytes": nsym}), flush=True)
        lib.pcx_wave_set_fused(2)
        del xs, sym, want
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
