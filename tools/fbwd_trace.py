"""Phase timeline of the fused decoder backward kernel (block 0): SM-clock records per role."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200
from svr_b200 import ops, _abi

M, KP = 200000, 2624
g = torch.Generator().manual_seed(0)
dz2 = torch.randn((M, 256), generator=g).cuda().bfloat16()
h1 = torch.relu(torch.randn((M, 256), generator=g)).cuda().bfloat16()
h0 = torch.relu(torch.randn((M, 256), generator=g)).cuda().bfloat16()
ws = [(torch.randn(s, generator=g) * 0.06).cuda().bfloat16() for s in ((256, 256), (256, 256), (KP, 256))]
imgs = [ops.swizzled_image(w) for w in ws]
for _ in range(3):
    ops.decoder_bwd_fused(dz2, h1, h0, *imgs, KP)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.decoder_bwd_fused(dz2, h1, h0, *imgs, KP)
e1.record()
torch.cuda.synchronize()
print("ms per launch", e0.elapsed_time(e1) / 5)
buf = torch.zeros((3, 1024, 2), dtype=torch.int64, device="cuda")
_abi.load().svr_debug_fb_trace(buf.data_ptr())
ops.decoder_bwd_fused(dz2, h1, h0, *imgs, KP)
torch.cuda.synchronize()
_abi.load().svr_debug_fb_trace(None)
b = buf.cpu()
t0 = int(b[b[:, :, 1] > 0][:, 1].min())
for role, name in enumerate(("worker", "mma", "loader")):
    rec = [(int(t), int(c) - t0) for t, c in b[role] if c > 0]
    print(name, len(rec))
    print("  " + " ".join(f"{t}:{c}" for t, c in rec[:110]))
