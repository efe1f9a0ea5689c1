"""Times svr_gather_bwd's kernels inside one training backward at the bench shape."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200
from svr_b200 import _abi

torch.manual_seed(0)
svr_b200.configure(net_res=128)
net = svr_b200.IFNet().cuda().train()
B, N, D = 4, 50000, 128
x = (torch.rand(B, 1, D, D, D) < 0.05).float().cuda()
pts = (torch.rand(B, N, 3) - 0.5).cuda()
for i in range(4):
    if i == 3:
        _abi.PROFILE.reset(with_events=True)
    net.zero_grad(set_to_none=True)
    net(x, pts).sum().backward()
torch.cuda.synchronize()
for k, v in _abi.PROFILE.events.items():
    if "gather_bwd" in k or "query" in k:
        print(k, sum(a.elapsed_time(b) for a, b in v))
