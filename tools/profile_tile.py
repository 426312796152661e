"""Runs the configs[1] tile pipeline (slice+pad -> conv_pair_kernel<192> -> uslice) on ONE 192 x 1024 x 2048 image a few times -
meant to be wrapped in `ncu --set full -k regex:conv_pair_kernel` (DRAM traffic / tensor-pipe activity of the bench's dominant
tensor kernel; profiles/traffic.json is regenerated from that capture)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pseudocylindrical_convolution_b200.tile_pipeline import TilePipeline
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
torch.manual_seed(0)
pipe = TilePipeline(192, 192, npart=16, opt=True, act=True, device=0)
x = torch.rand((1, 192, 1024, 2048), device=dev)
out = torch.empty_like(x)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    pipe(x, out)
torch.cuda.synchronize()
print("done", float(out[0, 0, 0, :4].sum()))
