#!/usr/bin/env python
"""Timing of the codec stages on one GPU (not the headline bench): analysis / synthesis transforms on the tensor-core path at
several ERP sizes, and full encode / decode (wavefront + host coder) at 512x1024.  Prints one JSON line per measurement."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    from conftest import smooth_images
    from pseudocylindrical_convolution_b200 import _lib, pseudo_codec as pc
    from pseudocylindrical_convolution_b200.random_init import synthesize_checkpoints
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    d = "/tmp/pcx_codec_bench"
    p_enc, p_dec, p_ent = synthesize_checkpoints(d, "4_56", 56, 0, seed=0)
    enc = pc.PseudoEncoder(56, 0).to(dev)
    dec = pc.PseudoDecoder(56, 0).to(dev)
    pc.load_models(enc, p_enc, p_ent, "cuda:0")
    pc.load_models(dec, p_dec, p_ent, "cuda:0")
    lib = _lib.load()
    sizes = [(512, 1024, 1), (1024, 2048, 1), (2048, 4096, 1), (2048, 4096, 4)]
    if "--big" in sys.argv:
        sizes.append((4096, 8192, 1))
    for H, W, N in sizes:
        x = torch.from_numpy(smooth_images(1, 3, H, W, seed=1)).to(dev).repeat(N, 1, 1, 1).contiguous()
        n0 = lib.pcx_launch_count()
        ms = timed(lambda: enc.latent(x), 3, 5)        # call 2 of a new problem captures the CUDA graph: keep it out of the timing
        launches = lib.pcx_launch_count() - n0            # launches issued by the host: 2 eager passes, then graph replays
        mp = N * H * W / 1e6
        # dense-grid FLOP counts of SURVEY.md 8d scaled to valid cells
        print(json.dumps({"stage": "analysis transform (TC)", "H": H, "W": W, "batch": N, "ms": ms, "MP/s": mp / ms * 1e3,
                          "TFLOP/s": 2 * 420552 * 0.8164 * mp * 1e6 / ms / 1e9, "launches": launches}))
        lat = enc.latent(x)
        sym = enc.dtw(enc.ext(enc.quant(lat)[1]))
        ms = timed(lambda: dec.reconstruct(sym), 3, 5)
        print(json.dumps({"stage": "synthesis transform (TC)", "H": H, "W": W, "batch": N, "ms": ms, "MP/s": mp / ms * 1e3,
                          "TFLOP/s": 2 * 517752 * 0.8164 * mp * 1e6 / ms / 1e9}))
        del lat, sym, x
        torch.cuda.empty_cache()
    if "--transforms-only" in sys.argv:
        return
    for H, W in ((512, 1024), (2048, 4096)):
        x = torch.from_numpy(smooth_images(1, 3, H, W, seed=1)).to(dev)
        path = os.path.join(d, "img.bin")
        steps = H // 8 + W // 8 + 14 - 2
        for name, fn in (("full encode %dx%d (transform + %d-step wavefront + host coder)" % (H, W, steps), lambda: enc(x, path)),
                         ("full decode %dx%d" % (H, W), lambda: dec(path, H, W))):
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            ts.sort()
            print(json.dumps({"stage": name, "p50_ms": ts[1] * 1e3, "MP/s": H * W / 1e6 / ts[1], "bytes": os.path.getsize(path)}), flush=True)
        del x
    H, W = 512, 1024

    for N, H, W in ((4, 512, 1024), (8, 512, 1024), (16, 512, 1024), (2, 2048, 4096)):
        xs = torch.from_numpy(smooth_images(N, 3, H, W, seed=5)).to(dev)
        names = [os.path.join(d, "b%d.bin" % i) for i in range(N)]
        for name, fn in (("batched encode %d x %dx%d" % (N, H, W), lambda: enc.encode_batch(xs, names)),
                         ("batched decode %d x %dx%d" % (N, H, W), lambda: dec.decode_batch(names, H, W))):
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            ts.sort()
            print(json.dumps({"stage": name, "p50_ms": ts[1] * 1e3, "MP/s": N * H * W / 1e6 / ts[1]}), flush=True)
        del xs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
