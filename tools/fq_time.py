"""Times the fused query forward (inference mode and training mode with saves) at the bench shape."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200

torch.manual_seed(0)
svr_b200.configure(net_res=128)
net = svr_b200.IFNet().cuda().eval()
B, N, D = 4, 50000, 128
x = (torch.rand(B, 1, D, D, D) < 0.05).float().cuda()
pts = (torch.rand(B, N, 3) - 0.5).cuda()
with torch.no_grad():
    vols = net.ifnet_feature_extractor.encode(x)
    for _ in range(3):
        out = net.query(x, vols, pts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = net.query(x, vols, pts)
    e1.record()
    torch.cuda.synchronize()
print("query (sort + pack + fused fwd) ms", e0.elapsed_time(e1) / 10)
from svr_b200 import _abi
_abi.PROFILE.reset(with_events=True)
with torch.no_grad():
    for _ in range(5):
        out = net.query(x, vols, pts)
torch.cuda.synchronize()
for k, v in _abi.PROFILE.events.items():
    print(k, sum(a.elapsed_time(b) for a, b in v) / 5)
