"""Profiling target: the fused query forward at BASELINE configs[1] shape (4 scenes x 50k points)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200
from svr_b200 import ops

torch.manual_seed(0)
svr_b200.configure(net_res=128)
net = svr_b200.IFNet().cuda().eval()
B, N, D = 4, 50000, 128
x = (torch.rand(B, 1, D, D, D) < 0.05).float().cuda()
pts = (torch.rand(B, N, 3) - 0.5).cuda()
mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
with torch.no_grad():
    vols = net.ifnet_feature_extractor.encode(x)
if mode == "fwd":
    with torch.no_grad():
        for _ in range(4):
            out = net.query(x, vols, pts)
else:
    for _ in range(3):
        out = net.query(x, [v.clone().requires_grad_(True) for v in vols], pts)
        out.sum().backward()
torch.cuda.synchronize()
print("ok", float(out.sum()))
