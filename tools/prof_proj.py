"""Profiling target: project.forward (unproject + voxelise + blur) for 64 depth maps; argv[1] = 128 | 256."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, svr_b200
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
depth = (torch.rand((64, 256, 256), generator=g) * 5.0 + 0.5).to(dev)
dims, scale = ((256,) * 3, 0.5) if len(sys.argv) > 1 and sys.argv[1] == "256" else ((128,) * 3, 1)
proj = svr_b200.project(dims, [3, 3, 3], torch.tensor([1.5, 1.5, 1.5])).to(dev)
with torch.no_grad():
    for _ in range(3):
        out = proj(proj.depthmap_to_normed_points(depth, scale))
torch.cuda.synchronize()
print("ok", float(out.sum()))
