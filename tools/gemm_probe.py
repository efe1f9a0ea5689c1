"""First-contact probe for the tcgen05 descriptors: runs the NT and TN GEMMs on small problems and
prints the error against torch, so that a wrong LBO/SBO/K-step hypothesis is visible at once."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import svr_b200

ops = svr_b200.ops
torch.manual_seed(0)
for (M, N, K) in [(128, 256, 64), (128, 256, 128), (256, 512, 256)]:
    A = torch.randn(M, K).cuda().bfloat16()
    B = torch.randn(N, K).cuda().bfloat16()
    out = torch.zeros(M, N, device="cuda")
    ops._gemm_nt(A, B, None, M, N, K, ops.ST_F32, c_f32=out, ldc=N)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    print("NT", M, N, K, "rel err", float((out - ref).abs().max() / ref.abs().max()), flush=True)
for (P, M, N) in [(64, 128, 256), (128, 128, 256), (256, 256, 512)]:
    A = torch.randn(P, M).cuda().bfloat16()
    B = torch.randn(P, N).cuda().bfloat16()
    out = torch.zeros(M, N, device="cuda")
    ops._gemm_tn(A, B, M, N, P, out)
    torch.cuda.synchronize()
    ref = A.float().t() @ B.float()
    print("TN", P, M, N, "rel err", float((out - ref).abs().max() / ref.abs().max()), flush=True)
