"""Fused forward kernel vs the unfused (already parity-checked) pipeline, plus timings."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200
from svr_b200 import ops
from oracle import ref_torch as R

torch.manual_seed(0)
svr_b200.configure(net_res=128)
net = svr_b200.IFNet().cuda().eval()
net.load_state_dict(R.synthetic_state_dict(3, 128), strict=False)
for (B, N, D) in [(1, 100, 16), (2, 300, 32), (1, 1000, 48), (4, 50000, 128)]:
    x = (torch.rand(B, 1, D, D, D) < 0.1).float().cuda()
    pts = (torch.rand(B, N, 3) - 0.5).cuda() * 1.05
    with torch.no_grad():
        vols = net.ifnet_feature_extractor.encode(x)
        ops.USE_FUSED = False
        a = net.query(x, vols, pts)
        ops.USE_FUSED = True
        b = net.query(x, vols, pts)
        torch.cuda.synchronize()
    print(B, N, D, "fused vs unfused max|d|", float((a - b).abs().max()), "max|ref|", float(a.abs().max()), flush=True)
    # training-mode saves
    pr = pts.clone().requires_grad_(True)
    ops.USE_FUSED = True
    l1 = net.query(x, [v.clone().requires_grad_(True) for v in vols], pr)
    l1.sum().backward()
    g1 = pr.grad.clone()
    ops.USE_FUSED = False
    pr2 = pts.clone().requires_grad_(True)
    l2 = net.query(x, [v.clone().requires_grad_(True) for v in vols], pr2)
    l2.sum().backward()
    print("   train-mode logits d", float((l1 - l2).abs().max()), "dpts rel", float((g1 - pr2.grad).norm() / pr2.grad.norm()), flush=True)
# timing at config-2 shape
for fused, sortmin in ((False, 1 << 30), (True, 1 << 30), (True, 2048)):
    ops.USE_FUSED = fused
    ops.SORT_MIN_POINTS = sortmin
    with torch.no_grad():
        for _ in range(3):
            net.query(x, vols, pts)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            net.query(x, vols, pts)
        e1.record()
        torch.cuda.synchronize()
    print("fused" if fused else "unfused", "sorted" if sortmin < 1 << 30 else "unsorted", "fwd ms (incl. volume + weight packing)", e0.elapsed_time(e1) / 10, flush=True)

