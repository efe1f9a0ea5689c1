"""Attributes aten::copy_/fill_/add kernels of one training step to python call sites (torch.profiler with stacks)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200
from torch.profiler import profile, ProfilerActivity

torch.manual_seed(0)
torch.backends.cudnn.benchmark = True
svr_b200.configure(net_res=128)
net = svr_b200.IFNet().cuda().train()
opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
B, N, D = 4, 50000, 128
x = (torch.rand(B, 1, D, D, D) < 0.05).float().cuda()
pts = (torch.rand(B, N, 3) - 0.5).cuda()
occ = (torch.rand(B, N) < 0.5).float().cuda()

def step():
    opt.zero_grad(set_to_none=True)
    logits = net(x, pts)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, occ, reduction="none").sum(-1).mean()
    loss.backward()
    opt.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
rows = []
for e in prof.events():
    if e.name in ("aten::copy_", "aten::fill_", "aten::add", "aten::add_", "aten::zero_", "aten::contiguous", "aten::sum", "aten::clone", "aten::to", "aten::_to_copy") and e.device_time_total > 15:
        chain, q = [], e.cpu_parent
        while q is not None and len(chain) < 5:
            chain.append(q.name)
            q = q.cpu_parent
        rows.append((e.device_time_total, e.name, str(e.input_shapes)[:70], " <- ".join(chain)))
for r in sorted(rows, reverse=True)[:40]:
    print(f"{r[0]:8.1f} us  {r[1]:16s} {r[2]:90s} {r[3]}")
