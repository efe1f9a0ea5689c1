"""Phase timeline of the fused kernel in dense-evaluation (lattice) mode, block 0."""
import sys, re
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200
from svr_b200 import _abi

torch.manual_seed(0)
svr_b200.configure(net_res=128)
net = svr_b200.IFNet().cuda().eval()
x = (torch.rand(1, 1, 128, 128, 128) < 0.05).float().cuda()
with torch.no_grad():
    net.evaluate_grid(x, (256, 256, 256), scenes=[0], x_range=(0, 64))
    buf = torch.zeros((4, 1024, 2), dtype=torch.int64, device="cuda")
    _abi.load().svr_debug_fq_trace(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    net.evaluate_grid(x, (256, 256, 256), scenes=[0], x_range=(0, 64))
    e1.record()
    torch.cuda.synchronize()
    _abi.load().svr_debug_fq_trace(None)
print("quarter scene ms", e0.elapsed_time(e1))
b = buf.cpu()
t0 = int(b[b[:, :, 1] > 0][:, 1].min())
for role, name in enumerate(("producer warp 0", "mma", "epilogue warp 0", "prep warp")):
    recs = [(int(t), int(c) - t0) for t, c in b[role] if c > 0]
    print(name, len(recs))
    print("  " + " ".join(f"{t}:{c}" for t, c in recs[:150]))
