"""Analysis + synthesis of a batch of 512x1024 images, eager (no CUDA graph) - to be wrapped in ncu for one kernel of the batch-8 shapes."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import smooth_images
from pseudocylindrical_convolution_b200 import config, pseudo_codec as pc
from pseudocylindrical_convolution_b200.random_init import synthesize_checkpoints
config.CUDA_GRAPHS = False
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
p_enc, p_dec, p_ent = synthesize_checkpoints("/tmp/pcx_prof8", "4_56", 56, 0, seed=0)
enc = pc.PseudoEncoder(56, 0).to(dev); dec = pc.PseudoDecoder(56, 0).to(dev)
pc.load_models(enc, p_enc, p_ent, "cuda:0"); pc.load_models(dec, p_dec, p_ent, "cuda:0")
x = torch.from_numpy(smooth_images(N, 3, 512, 1024, seed=1)).to(dev)
for it in range(2):
    sym = enc.symbols(x)
    rec = dec.reconstruct(sym)
    torch.cuda.synchronize()
print("done", float(rec.mean()))
