"""svr_blur_fwd (3x3x3) on 64 grids of S^3: ms per call and GB/s of compulsory traffic; argv[1] = S.  (The z-split sweep
recorded in csrc/projection.cu / DESIGN.md 4.2 was run with a temporary override of `zsplit` in svr_blur_fwd.)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200

S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
proj = svr_b200.project((S, S, S), [3, 3, 3], torch.tensor([1.5, 1.5, 1.5])).cuda()
vox = torch.rand((64, S, S, S), device="cuda")
with torch.no_grad():
    k = proj.smoothing_kernel()
    for _ in range(3):
        proj.voxels_smooth(vox, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        proj.voxels_smooth(vox, k)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"blur {S}^3 x 64: {ms:.4f} ms/call, {2 * vox.numel() * 4 / ms / 1e6:.0f} GB/s")
