import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from conftest import smooth_images
from pseudocylindrical_convolution_b200 import pseudo_codec as pc
from pseudocylindrical_convolution_b200.random_init import synthesize_checkpoints
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
p_enc, p_dec, p_ent = synthesize_checkpoints("/tmp/pcx_enc8", "4_56", 56, 0, seed=0)
enc = pc.PseudoEncoder(56, 0).to(dev)
pc.load_models(enc, p_enc, p_ent, "cuda:0")
x = torch.from_numpy(smooth_images(8, 3, 512, 1024, seed=1)).to(dev)
sym = enc.symbols(x)
names = ["/tmp/pcx_enc8/b%d.bin" % i for i in range(8)]
for it in range(2):
    enc.ent.encode_batch(sym.clone(), names)
    torch.cuda.synchronize()
print("done")
