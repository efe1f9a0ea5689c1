"""Profiling target: the dense evaluator (box kernel) on a 64-wide slab of a 256^3 lattice, one 128^3 scene."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200

torch.manual_seed(0)
svr_b200.configure(net_res=128)
net = svr_b200.IFNet().cuda().eval()
x = (torch.rand(1, 1, 128, 128, 128) < 0.05).float().cuda()
with torch.no_grad():
    for _ in range(2):
        g = net.evaluate_grid(x, (256, 256, 256), scenes=[0], x_range=(0, 64))
torch.cuda.synchronize()
print("ok", float(g.sum()))
