"""Stage-by-stage check of the fp32 tier (ops._Query32) against torch fp64 on the GPU: locates which stage loses accuracy."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.nn.functional as F
import svr_b200
from svr_b200 import ops, _abi
from oracle import ref_torch as R
import ctypes as C

torch.backends.cuda.matmul.allow_tf32 = False
svr_b200.configure(net_res=128, precision=32)
sd = R.synthetic_state_dict(41, 128)
net = svr_b200.IFNet().cuda().train()
net.load_state_dict(sd, strict=False)
g = torch.Generator().manual_seed(17)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = ((torch.rand((2, 1, S, S, S), generator=g) < 0.05).float() * torch.rand((2, 1, S, S, S), generator=g)).cuda()
pts = ((torch.rand((2, 4096, 3), generator=g) - 0.5) * 1.02).cuda()
cot = torch.randn((2, 4096), generator=g).cuda()
with torch.no_grad():
    vols = net.encode(x)
pyr = net.ifnet_feature_extractor.pyramid(x, vols)
M, KP = 8192, pyr.kp
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

# reference pipeline in fp64 on the GPU
xr = x.double().requires_grad_(True)
vr = [v.double().contiguous().requires_grad_(True) for v in vols]
pr = pts.double().requires_grad_(True)
grid = R.stencil_grid(pr, R.DISPLACEMENT_128)
feat = torch.cat([F.grid_sample(v, grid, align_corners=False) for v in [xr] + vr], 1)
B, Cc, _, S7, N = feat.shape
f = feat[:, :, 0].permute(0, 3, 1, 2).reshape(B * N, Cc * S7)
W = {k: sd[k].cuda().double().reshape(sd[k].shape[0], -1).requires_grad_(True) for k in ("fc_0.weight", "fc_1.weight", "fc_2.weight", "fc_out.weight")}
bb = {k: sd[k].cuda().double() for k in ("fc_0.bias", "fc_1.bias", "fc_2.bias", "fc_out.bias")}
z0 = f @ W["fc_0.weight"].t() + bb["fc_0.bias"]; h0 = torch.relu(z0)
z1 = h0 @ W["fc_1.weight"].t() + bb["fc_1.bias"]; h1 = torch.relu(z1)
z2 = h1 @ W["fc_2.weight"].t() + bb["fc_2.bias"]; h2 = torch.relu(z2)
logits = (h2 @ W["fc_out.weight"].t()).squeeze(1) + bb["fc_out.bias"]
for t in (f, z0, z1, z2):
    t.retain_grad()
logits.backward(cot.reshape(-1).double())

# device pipeline pieces
idx = ops.feature_index_map(pyr, x.device)
vf = [ops._ndhwc_f32(v) for v in vols]
featd = torch.empty((M, KP), device=x.device, dtype=torch.float32)
vt = _abi.ptr_table([None] + [v.data_ptr() for v in vf])
_abi.check(_abi.load().svr_gather_fwd_f32(pts.data_ptr(), 2, 4096, x.data_ptr(), vt, C.byref(pyr.c), featd.data_ptr(), ops._stream()), "g")
print("features      ", rel(featd.index_select(1, idx), f))
xx = x.clone().requires_grad_(True); pp = pts.clone().requires_grad_(True)
vv = [v.clone().requires_grad_(True) for v in vols]
out = net.query(xx, vv, pp)
print("logits        ", rel(out.reshape(-1), logits))
out.backward(cot)
print("dx            ", rel(xx.grad, xr.grad))
print("dpts          ", rel(pp.grad, pr.grad))
for i in range(5):
    print(f"dvol{i+1}         ", rel(vv[i].grad, vr[i].grad))
for nm in ("fc_0", "fc_1", "fc_2", "fc_out"):
    print(nm, "w", rel(getattr(net, nm).weight.grad.reshape(W[nm + ".weight"].shape), W[nm + ".weight"].grad))
# dfeat directly: split product of the true dz0 with W0
dz0 = z0.grad.float().contiguous()
w0p = torch.empty((256, KP), device=x.device, dtype=torch.float32)
w0f = sd["fc_0.weight"].cuda().reshape(256, -1).contiguous()
_abi.check(_abi.load().svr_pack_w0_f32(w0f.data_ptr(), 256, C.byref(pyr.c), w0p.data_ptr(), ops._stream()), "p")
print("w0p permutation", rel(w0p.index_select(1, idx), w0f))
dfeat = torch.empty((M, KP), device=x.device, dtype=torch.float32)
ops._mm_nt3(ops._split(dz0), ops._split(w0p.t().contiguous()), None, M, KP, 256, dfeat)
print("dfeat (split product of the exact dz0)", rel(dfeat.index_select(1, idx), f.grad))
print("dfeat level-0 columns", rel(dfeat[:, :7], f.grad[:, :7]))
# scatter of the exact dfeat
dfe = torch.zeros((M, KP), device=x.device, dtype=torch.float32)
dfe.index_copy_(1, idx, f.grad.float())
gx = torch.zeros_like(x)
gb = [torch.zeros_like(v) for v in vf]
gt = _abi.ptr_table([None] + [t.data_ptr() for t in gb])
_abi.check(_abi.load().svr_gather_bwd_f32(pts.data_ptr(), 2, 4096, x.data_ptr(), vt, C.byref(pyr.c), dfe.data_ptr(), gx.data_ptr(), gt, None, ops._stream()), "s")
print("scatter(exact dfeat): dx", rel(gx, xr.grad), " dvol1", rel(gb[0].permute(0, 4, 1, 2, 3), vr[0].grad))
