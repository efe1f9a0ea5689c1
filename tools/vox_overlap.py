"""Timeline of svr_voxelize_fwd (64 maps, 256^3): does the grid's zero fill (helper stream) overlap the bucketing kernels?"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200
from torch.profiler import profile, ProfilerActivity

S, B = (int(sys.argv[1]) if len(sys.argv) > 1 else 256), 64
scale = 1 if S == 128 else 0.5
proj = svr_b200.project((S, S, S), [3, 3, 3], torch.tensor([1.5, 1.5, 1.5])).cuda()
if len(sys.argv) > 2 and sys.argv[2] == "smooth":
    # a smooth surface per map (a tilted, gently waving sheet): neighbouring pixels land in neighbouring cells, no cell is
    # isolated -- the opposite extreme of the iid-random depths of BASELINE's synthetic configuration
    v, u = torch.meshgrid(torch.linspace(0, 1, 256), torch.linspace(0, 1, 256), indexing="ij")
    ph = torch.arange(B).view(B, 1, 1) * 0.37
    depth = (2.5 + 0.8 * u + 0.5 * torch.sin(6.0 * v + ph) + 0.3 * torch.cos(9.0 * u - ph)).cuda()
else:
    depth = (torch.rand((B, 256, 256)) * 5.0 + 0.5).cuda()
with torch.no_grad():
    pts = proj.depthmap_to_normed_points(depth, scale)
    for _ in range(3):
        proj.pc_voxels(pts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        proj.pc_voxels(pts)
    e1.record()
    torch.cuda.synchronize()
    print("pc_voxels ms/call", e0.elapsed_time(e1) / 20)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            proj.pc_voxels(pts)
        torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_time > 0], key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev:
    print(f"{(e.time_range.start - t0):10.1f} us  +{e.device_time:8.1f} us  {e.name[:70]}")
