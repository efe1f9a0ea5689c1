import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.nn.functional as F
import svr_b200
from svr_b200 import ops
from oracle import ref_torch as R
torch.manual_seed(0)
chans, delta, ac = (1, 16, 32, 64, 128, 128), R.DISPLACEMENT_128, False
B, N = 1, 64
dims = [(32, 24, 16), (32, 24, 16), (16, 12, 8), (8, 6, 4), (4, 3, 2), (2, 1, 1)]
x = torch.rand(B, 1, *dims[0]).cuda()
vols = [torch.randn(B, c, *d).cuda().bfloat16().float() for c, d in zip(chans[1:], dims[1:])]
pts = ((torch.rand(B, N, 3) - 0.5) * 0.9).cuda()
pyr = ops.PyramidSpec(chans, dims, ac, delta)
idx = ops.feature_index_map(pyr, "cuda")
for j in range(7):
    for need_p in (False, True):
        xr = x.clone().requires_grad_(True)
        grid = R.stencil_grid(pts, delta)
        ref = F.grid_sample(xr, grid, align_corners=ac)       # (B,1,1,7,N)
        gk = torch.zeros_like(ref)
        gk[:, :, :, j] = 1.0
        ref.backward(gk)
        x2 = x.clone().requires_grad_(True)
        p2 = pts.clone().requires_grad_(need_p)
        feat = ops.gather(pyr, p2, x2, vols)
        gfeat = torch.zeros(B * N, pyr.kp, device="cuda")
        gfeat[:, j] = 1.0
        feat.backward(gfeat.bfloat16())
        print(j, need_p, "sum mine %.4f ref %.4f  maxdiff %.4f nnz mine %d ref %d" % (float(x2.grad.sum()), float(xr.grad.sum()), float((x2.grad - xr.grad).abs().max()), int((x2.grad != 0).sum()), int((xr.grad != 0).sum())))
