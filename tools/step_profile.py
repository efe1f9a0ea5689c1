"""torch.profiler view of one training step (which torch-side ops surround the hot path)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
import svr_b200
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda:0")
svr_b200.configure(net_res=128, channels_last=True)
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = svr_b200.IFNet().to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
x, pts, occ = bench.synthetic_inputs(4, 100, dev)
pts, occ = pts.to(dev), occ.to(dev)


def step():
    opt.zero_grad(set_to_none=True)
    logits = net(x, pts)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, occ, reduction="none").sum(-1).mean()
    loss.backward()
    opt.step()


for _ in range(4):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))


evs = sorted([e for e in prof.events() if e.device_type.name == "CPU"], key=lambda e: e.time_range.start)
for i, ev in enumerate(evs):
    if ev.name == "aten::to" and ev.input_shapes and ev.input_shapes[0] == [4, 16, 128, 128, 128]:
        for e in evs[max(0, i - 25):i + 6]:
            print("   ", e.time_range.start, e.name, e.input_shapes[:2] if e.input_shapes else "")
        break
