"""Per-kernel times of one config-2 training backward (CUDA events around the library calls)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200
from svr_b200 import _abi

svr_b200.configure(net_res=128)
net = svr_b200.IFNet().cuda().train()
g = torch.Generator().manual_seed(0)
x = (torch.rand((4, 1, 128, 128, 128), generator=g) < 0.05).float().cuda()
pts = (torch.rand((4, 50000, 3), generator=g) - 0.5).cuda()
occ = (torch.rand((4, 50000), generator=g) < 0.5).float().cuda()
lossf = torch.nn.BCEWithLogitsLoss()
for it in range(6):
    if it == 3:
        torch.cuda.synchronize()
        _abi.load()
        _abi.PROFILE.reset(with_events=True)
    net.zero_grad(set_to_none=True)
    lossf(net(x, pts), occ).backward()
torch.cuda.synchronize()
for name, (calls, ms) in sorted(_abi.PROFILE.kernel_ms().items(), key=lambda kv: -kv[1][1]):
    print(f"{name:40s} calls {calls:3d}  {ms / 3:8.4f} ms/step")
