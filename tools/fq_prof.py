"""Kernel-level durations (CUPTI through torch.profiler) of one fused query forward at the bench shape."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200
from svr_b200 import _abi
from torch.profiler import profile, ProfilerActivity

torch.manual_seed(0)
svr_b200.configure(net_res=128)
net = svr_b200.IFNet().cuda().eval()
B, N, D = 4, 50000, 128
x = (torch.rand(B, 1, D, D, D) < 0.05).float().cuda()
pts = ((torch.rand(B, N, 3) - 0.5) * 1.03).cuda()
lib = _abi.load()
for interp in (0, 1):
    lib.svr_debug_fq_interp(interp)
    with torch.no_grad():
        vols = net.ifnet_feature_extractor.encode(x)
        for _ in range(3):
            net.query(x, vols, pts)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                net.query(x, vols, pts)
            torch.cuda.synchronize()
    print(f"interp={interp}")
    for e in prof.key_averages():
        if e.device_time_total > 0:
            print(f"  {e.key[:90]:90s} n={e.count:3d} avg {e.device_time_total / e.count:9.1f} us")
