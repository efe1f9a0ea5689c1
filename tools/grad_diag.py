"""Diagnostic: gradient errors (rel-inf and rel-L2) of the device path against the fp32 golden
vectors, and against a bf16-emulating torch restatement of the same pipeline (isolates kernel bugs
from bf16 rounding / ReLU-flip noise)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import svr_b200
from oracle import ref_torch as R


def rel(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm())


for net_res in (128, 32):
    g = np.load(f"tests/golden/ifnet{net_res}.npz")
    sd = R.synthetic_state_dict(100 + net_res, net_res)
    svr_b200.configure(net_res=net_res)
    net = svr_b200.IFNet().cuda()
    net.load_state_dict(sd, strict=False)
    for mode in ("eval", "train"):
        net.load_state_dict(sd, strict=False)
        net.train(mode == "train")
        net.zero_grad()
        x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
        pts = torch.from_numpy(g["pts"]).cuda().requires_grad_(True)
        occ = torch.from_numpy(g["occ"]).cuda()
        logits = net(x, pts)
        logits.backward(torch.from_numpy(g["cot"]).cuda())
        first = "conv_in" if net_res == 128 else "conv_1"
        checks = {"dx": x.grad, "dpts": pts.grad, f"{first}_w": getattr(net.ifnet_feature_extractor, first).weight.grad}
        for nm in ("fc_out", "fc_2", "fc_1", "fc_0"):
            checks[f"{nm}_w"] = getattr(net, nm).weight.grad[:8]
            checks[f"{nm}_b"] = getattr(net, nm).bias.grad
        print(f"== net {net_res} {mode}: logits", rel(logits.detach(), g[f"{mode}_logits"]))
        for k, v in checks.items():
            print(f"   vjp {k:12s} inf %.4f  l2 %.4f" % rel(v, g[f"{mode}_vjp_{k}"]))
        # same forward in torch with fp32 everywhere on the SAME device volumes: isolates the encoder
        with torch.no_grad():
            vols = net.ifnet_feature_extractor.encode(x)
        sdc = {k: v.cuda() for k, v in sd.items()}
        xr = x.detach().clone().requires_grad_(True)
        pr = pts.detach().clone().requires_grad_(True)
        vr = [v.detach().clone().requires_grad_(True) for v in vols]
        ws = {k: net.state_dict()[k].detach().clone().requires_grad_(True) for k in sd if k.startswith("fc_")}
        lr = R.query_from_volumes(ws, [xr] + vr, pr, net_res)
        lr.backward(torch.from_numpy(g["cot"]).cuda())
        # device path on the same volumes
        net.zero_grad()
        x2 = x.detach().clone().requires_grad_(True)
        p2 = pts.detach().clone().requires_grad_(True)
        v2 = [v.detach().clone().requires_grad_(True) for v in vols]
        l2 = net.query(x2, v2, p2)
        l2.backward(torch.from_numpy(g["cot"]).cuda())
        print("   -- hot path only (same volumes), device vs torch fp32 autograd")
        print("   logits           inf %.4f  l2 %.4f" % rel(l2.detach(), lr.detach()))
        for name in ("fc_out.weight", "fc_out.bias", "fc_2.weight", "fc_2.bias", "fc_1.weight", "fc_1.bias", "fc_0.weight", "fc_0.bias"):
            mod, attr = name.split(".")
            print(f"   {name:16s} inf %.4f  l2 %.4f" % rel(getattr(getattr(net, mod), attr).grad, ws[name].grad))
        print("   dx               inf %.4f  l2 %.4f" % rel(x2.grad, xr.grad))
        print("   dpts             inf %.4f  l2 %.4f" % rel(p2.grad, pr.grad))
        for i, (a, b) in enumerate(zip(v2, vr)):
            print(f"   dvol{i + 1}            inf %.4f  l2 %.4f" % rel(a.grad, b.grad))
