"""Host-side cost of one training step (config 2 shape): cProfile of 20 enqueue-only steps."""
import cProfile
import pstats
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200

torch.manual_seed(0)
svr_b200.configure(net_res=128, precision=16)
torch.backends.cudnn.benchmark = True
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


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e3 * (t1 - t0) / 20:.2f} ms/step, total {1e3 * (t2 - t0) / 20:.2f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
st.sort_stats("tottime").print_stats(30)
