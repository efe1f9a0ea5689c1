"""Per-block duration of the fused query forward (first to last trace record) for a few blocks: load balance check."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200
from svr_b200 import _abi

torch.manual_seed(0)
svr_b200.configure(net_res=128)
net = svr_b200.IFNet().cuda().eval()
B, N, D = 4, 50000, 128
x = (torch.rand(B, 1, D, D, D) < 0.05).float().cuda()
pts = ((torch.rand(B, N, 3) - 0.5) * 1.03).cuda()
lib = _abi.load()
with torch.no_grad():
    vols = net.ifnet_feature_extractor.encode(x)
    for _ in range(2):
        net.query(x, vols, pts)
    for blk in (0, 1, 37, 74, 100, 146, 147):
        buf = torch.zeros((4, 1024, 2), dtype=torch.int64, device="cuda")
        lib.svr_debug_fq_trace_block(blk)
        lib.svr_debug_fq_trace(buf.data_ptr())
        net.query(x, vols, pts)
        torch.cuda.synchronize()
        lib.svr_debug_fq_trace(None)
        b = buf.cpu()
        ok = b[:, :, 1] > 0
        t0, t1 = int(b[:, :, 1][ok].min()), int(b[:, :, 1][ok].max())
        ep = [(int(t), int(c) - t0) for t, c in b[2] if c > 0]
        ends = [c for t, c in ep if t == 612]
        nr = [sum(1 for t, c in ep if 900 <= t < 1000)]
        print(f"block {blk}: span {t1 - t0} clocks, tiles {len(ends)}, rounds {nr}, tile ends {ends}")
