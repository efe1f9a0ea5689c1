"""Small explicit-points run of the box kernel (svr_debug_fq_interp(2)) against the gather kernel."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200
from svr_b200 import _abi

torch.manual_seed(0)
svr_b200.configure(net_res=128)
net = svr_b200.IFNet().cuda().eval()
lib = _abi.load()
B, N, D = 1, 4096, 64
x = (torch.rand(B, 1, D, D, D) < 0.05).float().cuda()
pts = ((torch.rand(B, N, 3) - 0.5) * 1.03).cuda()
outs = []
for mode in (0, 4, 2):
    lib.svr_debug_fq_interp(mode)
    with torch.no_grad():
        vols = net.ifnet_feature_extractor.encode(x)
        outs.append(net.query(x, vols, pts).clone())
    torch.cuda.synchronize()
    print("mode", mode, "ok", flush=True)
print("max|d|/max|ref|", float((outs[0] - outs[1]).abs().max() / outs[0].abs().max()))
lib.svr_debug_fq_interp(1)
