"""Runs the analysis + synthesis transforms once (after a warm-up) at a given ERP size - meant to be wrapped in
ncu --metrics gpu__time_duration.sum to get the per-kernel launch list of the tensor-core path."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import smooth_images
from pseudocylindrical_convolution_b200 import pseudo_codec as pc
from pseudocylindrical_convolution_b200.random_init import synthesize_checkpoints
H, W = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
p_enc, p_dec, p_ent = synthesize_checkpoints("/tmp/pcx_prof", "4_56", 56, 0, seed=0)
enc = pc.PseudoEncoder(56, 0).to(dev); dec = pc.PseudoDecoder(56, 0).to(dev)
pc.load_models(enc, p_enc, p_ent, "cuda:0"); pc.load_models(dec, p_dec, p_ent, "cuda:0")
x = torch.from_numpy(smooth_images(1, 3, H, W, seed=1)).to(dev)
for it in range(2):
    torch.cuda.synchronize()
    if it == 1: torch.cuda.nvtx.range_push("timed")
    lat = enc.latent(x)
    sym = enc.dtw(enc.ext(enc.quant(lat)[1]))
    rec = dec.reconstruct(sym)
    torch.cuda.synchronize()
print("done", float(rec.mean()))
