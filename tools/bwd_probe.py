"""Component-level probe of the decoder/gather backward against torch ops on the device."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import ctypes as C

import torch
import torch.nn.functional as F

import svr_b200
from svr_b200 import _abi, ops
from oracle import ref_torch as R

lib = _abi.load()
st = torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return "inf %.5f l2 %.5f" % (float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm()))


# ---- head bwd
M, Hd = 600, 256
dl = torch.randn(M).cuda()
h2 = torch.relu(torch.randn(M, Hd)).cuda().bfloat16()
wout = torch.randn(Hd).cuda()
dz2 = torch.empty(M, Hd, device="cuda", dtype=torch.bfloat16)
gw = torch.zeros(Hd, device="cuda")
gb = torch.zeros(1, device="cuda")
_abi.check(lib.svr_decoder_head_bwd(dl.data_ptr(), None, h2.data_ptr(), wout.data_ptr(), M, Hd, dz2.data_ptr(), gw.data_ptr(), gb.data_ptr(), st))
torch.cuda.synchronize()
print("head dz2 ", rel(dz2, (dl[:, None] * wout[None]) * (h2.float() > 0)))
print("head gw  ", rel(gw, (dl[:, None] * h2.float()).sum(0)))
print("head gb  ", rel(gb, dl.sum()[None]))
# ---- colsum
a = torch.randn(M, 512).cuda().bfloat16()
out = torch.empty(512, device="cuda")
_abi.check(lib.svr_colsum_bf16(a.data_ptr(), M, 512, 512, out.data_ptr(), 0, st))
print("colsum   ", rel(out, a.float().sum(0)))
# ---- gather fwd / bwd vs grid_sample autograd (fp32, on bf16-rounded volumes)
for net_res, chans, delta, ac in ((128, (1, 16, 32, 64, 128, 128), R.DISPLACEMENT_128, False), (32, (1, 64, 128, 128), R.DISPLACEMENT_32, True)):
    B, N = 2, 300
    dims = [(32, 24, 16), (32, 24, 16), (16, 12, 8), (8, 6, 4), (4, 3, 2), (2, 1, 1)][:len(chans)]
    x = torch.rand(B, 1, *dims[0]).cuda()
    vols = [torch.randn(B, c, *d).cuda().bfloat16().float() for c, d in zip(chans[1:], dims[1:])]
    pts = ((torch.rand(B, N, 3) - 0.5) * 1.1).cuda()
    pyr = ops.PyramidSpec(chans, dims, ac, delta)
    xr = x.clone().requires_grad_(True)
    vr = [v.clone().requires_grad_(True) for v in vols]
    pr = pts.clone().requires_grad_(True)
    grid = R.stencil_grid(pr, delta)
    ref = torch.cat([F.grid_sample(v, grid, align_corners=ac) for v in [xr] + vr], dim=1)   # (B,C,1,7,N)
    refk = ref[:, :, 0].permute(0, 3, 1, 2).reshape(B * N, -1)                                # (BN, C*7) k=c*7+d
    x2 = x.clone().requires_grad_(True)
    v2 = [v.clone().requires_grad_(True) for v in vols]
    p2 = pts.clone().requires_grad_(True)
    feat = ops.gather(pyr, p2, x2, v2)
    idx = ops.feature_index_map(pyr, "cuda")
    mine = feat.float().index_select(1, idx)
    print(net_res, "gather fwd", rel(mine, refk))
    gk = torch.randn_like(refk)
    refk.backward(gk)
    gfeat = torch.zeros(B * N, pyr.kp, device="cuda")
    gfeat[:, idx] = gk
    feat.backward(gfeat.bfloat16())
    print(net_res, "dx   ", rel(x2.grad, xr.grad))
    print(net_res, "dpts ", rel(p2.grad, pr.grad))
    for i, (a_, b_) in enumerate(zip(v2, vr)):
        print(net_res, f"dvol{i + 1}", rel(a_.grad, b_.grad))
