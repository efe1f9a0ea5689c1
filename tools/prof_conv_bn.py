"""Profiling target: fused conv_in + ReLU + BN stage at BASELINE configs[1] shape (4 x 128^3), fwd + bwd."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200
from svr_b200 import ops

torch.manual_seed(0)
B, D = 4, 128
x = (torch.rand(B, 1, D, D, D) < 0.05).float().cuda()
conv = torch.nn.Conv3d(1, 16, 3, padding=1).cuda()
bn = torch.nn.BatchNorm3d(16).cuda()
cot = torch.randn((B, D, D, D, 16), device="cuda").permute(0, 4, 1, 2, 3)
for _ in range(3):
    y = ops.conv1_relu_bn_channels_last(x, conv, bn)
    y.backward(cot)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ev[0].record()
y = ops.conv1_relu_bn_channels_last(x, conv, bn)
ev[1].record()
y.backward(cot)
ev[2].record()
torch.cuda.synchronize()
print("fwd ms", ev[0].elapsed_time(ev[1]), "bwd ms", ev[1].elapsed_time(ev[2]))
