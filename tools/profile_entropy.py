"""Entropy stage alone (context model + GMM tables + host coder) on precomputed symbols: wall-clock per call, for
`enc` / `dec` / `both`, at a given code size.  Wrap in `ncu --metrics gpu__time_duration.sum` for the launch list.

    python tools/profile_entropy.py H W NIMG [enc|dec|both] [reps]
"""
import json, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import smooth_images
from pseudocylindrical_convolution_b200 import _lib, pseudo_codec as pc
from pseudocylindrical_convolution_b200.random_init import synthesize_checkpoints
H, W, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
what = sys.argv[4] if len(sys.argv) > 4 else "both"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
d = "/tmp/pcx_prof"
p_enc, p_dec, p_ent = synthesize_checkpoints(d, "4_56", 56, 0, seed=0)
enc = pc.PseudoEncoder(56, 0).to(dev); dec = pc.PseudoDecoder(56, 0).to(dev)
pc.load_models(enc, p_enc, p_ent, "cuda:0"); pc.load_models(dec, p_dec, p_ent, "cuda:0")
x = torch.from_numpy(smooth_images(N, 3, H, W, seed=1)).to(dev)
sym = enc.symbols(x)
names = [os.path.join(d, "e%d.bin" % i) for i in range(N)]
lib = _lib.load()


def run(tag, fn):
    ts = []
    n0 = 0
    for i in range(reps + 1):
        torch.cuda.synchronize()
        n0 = lib.pcx_launch_count()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts = sorted(ts[1:])
    p50 = ts[len(ts) // 2]
    print(json.dumps({"stage": tag, "H": H, "W": W, "nimg": N, "p50_ms": p50 * 1e3, "MP/s": N * H * W / 1e6 / p50,
                      "launches": int(lib.pcx_launch_count() - n0), "symbols": int(sym.numel() * 0.8164)}), flush=True)


if what in ("enc", "both"):
    run("entropy encode", lambda: enc.ent.encode_batch(sym, names))
else:
    enc.ent.encode_batch(sym, names)
if what in ("dec", "both"):
    got = None
    def f():
        global got
        got = dec.ent.decode_batch(H // 128, W // 8, names)
    run("entropy decode", f)
    assert torch.equal(got, sym)
print("bytes", [os.path.getsize(n) for n in names][:4])
