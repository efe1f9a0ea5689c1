"""Repeats the fused conv_in+ReLU+BN stage check of tests/test_gpu_ifnet.py over many seeds and prints the worst
relative gradient deviation per tensor, and whether a second run on the same inputs is bit-identical."""
import copy
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch

import svr_b200
from svr_b200 import ops

torch.backends.cudnn.allow_tf32 = False


def rel(a, b):
    return float((a.detach() - b.detach()).abs().max() / b.detach().abs().max().clamp_min(1e-12))


def one(seed, shape):
    B, D, H, W = shape
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    x = (torch.rand((B, 1, D, H, W), generator=g) < 0.3).float().cuda()
    conv = torch.nn.Conv3d(1, 16, 3, padding=1).cuda()
    bn = torch.nn.BatchNorm3d(16).cuda()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(16, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(16, generator=g) * 0.1)
    conv_r, bn_r = copy.deepcopy(conv), copy.deepcopy(bn)
    conv_2, bn_2 = copy.deepcopy(conv), copy.deepcopy(bn)
    cot = torch.randn((B, 16, D, H, W), generator=g).cuda()
    cot_p = torch.randn((B, 16, D // 2, H // 2, W // 2), generator=g).cuda()
    y, yp = ops.conv1_relu_bn_channels_last(x, conv, bn, with_pool=True)
    ((y * cot).sum() + (yp * cot_p).sum()).backward()
    y2, yp2 = ops.conv1_relu_bn_channels_last(x, conv_2, bn_2, with_pool=True)
    ((y2 * cot).sum() + (yp2 * cot_p).sum()).backward()
    yr = bn_r(torch.relu(conv_r(x)))
    ypr = torch.nn.functional.max_pool3d(yr, 2)
    ((yr * cot).sum() + (ypr * cot_p).sum()).backward()
    pre = conv_r(x).detach()
    near = int((pre.abs() < 1e-6).sum())
    tie = int((torch.nn.functional.max_pool3d(yr, 2, return_indices=False).detach().repeat_interleave(2, 2).repeat_interleave(2, 3)
               .repeat_interleave(2, 4)[:, :, :D // 2 * 2, :H // 2 * 2, :W // 2 * 2] == yr.detach()[:, :, :D // 2 * 2, :H // 2 * 2, :W // 2 * 2]).sum()
              - ypr.numel())
    r = [rel(conv.weight.grad, conv_r.weight.grad), rel(conv.bias.grad, conv_r.bias.grad), rel(bn.weight.grad, bn_r.weight.grad),
         rel(bn.bias.grad, bn_r.bias.grad)]
    same = all(torch.equal(a.grad, b.grad) for a, b in ((conv.weight, conv_2.weight), (conv.bias, conv_2.bias), (bn.weight, bn_2.weight)))
    return r, rel(y, yr), same, near, tie


worst = 0.0
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    for shape in ((2, 24, 20, 40), (1, 16, 16, 32), (3, 8, 9, 70)):
        r, ry, same, near, tie = one(seed, shape)
        flag = " <<<" if max(r) > 2e-4 else ""
        worst = max(worst, max(r))
        if flag or not same or seed < 2:
            print(seed, shape, " ".join(f"{v:.2e}" for v in r), f"y {ry:.1e} rerun_identical={same} near_zero_pre={near} pool_ties={tie}{flag}")
print("worst", worst)
