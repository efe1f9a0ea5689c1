"""Phase timeline of the fused query forward kernel (block 0): SM-clock records per role (gather warp 0, MMA thread,
epilogue warp 0).  Prints per-tile summaries: chunk period of the gather / MMA, epilogue durations, idle gaps."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200
from svr_b200 import _abi

torch.manual_seed(0)
svr_b200.configure(net_res=128)
net = svr_b200.IFNet().cuda().eval()
B, N, D = 4, 50000, 128
x = (torch.rand(B, 1, D, D, D) < 0.05).float().cuda()
pts = (torch.rand(B, N, 3) - 0.5).cuda()
with torch.no_grad():
    vols = net.ifnet_feature_extractor.encode(x)
    for _ in range(2):
        net.query(x, vols, pts)
    buf = torch.zeros((4, 1024, 2), dtype=torch.int64, device="cuda")
    _abi.load().svr_debug_fq_trace(buf.data_ptr())
    net.query(x, vols, pts)
    torch.cuda.synchronize()
    _abi.load().svr_debug_fq_trace(None)
b = buf.cpu()
t0 = int(b[b[:, :, 1] > 0][:, 1].min())
recs = {}
for role, name in enumerate(("gather", "mma", "epilogue")):
    recs[name] = [(int(t), int(c) - t0) for t, c in b[role] if c > 0]
    print(name, len(recs[name]))
    print("  " + " ".join(f"{t}:{c}" for t, c in recs[name][:260]))
