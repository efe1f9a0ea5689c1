"""A/B of the fused query forward at the bench shape: generic gather vs the halo'd wide path (ops.USE_HALO).
Checks that both give bit-identical logits / saved features, then times inference and training-mode launches."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200
from svr_b200 import _abi, ops

torch.manual_seed(0)
svr_b200.configure(net_res=128)
net = svr_b200.IFNet().cuda().eval()
B, N, D = 4, 50000, 128
x = (torch.rand(B, 1, D, D, D) < 0.05).float().cuda()
pts = ((torch.rand(B, N, 3) - 0.5) * 1.03).cuda()
res = {}
for halo in (False, True):
    ops.USE_HALO = halo
    with torch.no_grad():
        vols = net.ifnet_feature_extractor.encode(x)
        out = net.query(x, vols, pts)
    pp = pts.clone().requires_grad_(True)
    vv = [v.clone().requires_grad_(True) for v in vols]
    out_t = net.query(x, vv, pp)           # training-mode launch (features / activations saved)
    feat = out_t.grad_fn.saved_tensors[2].clone()
    res[halo] = (out.clone(), out_t.detach().clone(), feat)
    _abi.PROFILE.reset(with_events=True)
    with torch.no_grad():
        for _ in range(5):
            net.query(x, vols, pts)
    for _ in range(5):
        net.query(x, vv, pp)
    torch.cuda.synchronize()
    ev = _abi.PROFILE.events["svr_query_fwd_fused"]
    print(f"halo={halo}: fused fwd inference {sum(a.elapsed_time(b) for a, b in ev[:5]) / 5:.4f} ms, training {sum(a.elapsed_time(b) for a, b in ev[5:]) / 5:.4f} ms",
          {k: round(sum(a.elapsed_time(b) for a, b in v) / 10, 4) for k, v in _abi.PROFILE.events.items() if k != "svr_query_fwd_fused"}, flush=True)
    _abi.PROFILE.reset()
print("logits bit-identical:", torch.equal(res[False][0], res[True][0]), torch.equal(res[False][1], res[True][1]),
      " saved features bit-identical:", torch.equal(res[False][2], res[True][2]),
      " max |dlogit|:", float((res[False][0] - res[True][0]).abs().max()))
# small non-cubic scene (runtime-stride wide path) and out-of-range points
x2 = (torch.rand(2, 1, 48, 40, 56) < 0.1).float().cuda()
p2 = ((torch.rand(2, 5000, 3) - 0.5) * 1.6).cuda()
outs = []
for halo in (False, True):
    ops.USE_HALO = halo
    with torch.no_grad():
        outs.append(net.query(x2, net.ifnet_feature_extractor.encode(x2), p2))
print("non-cubic bit-identical:", torch.equal(outs[0], outs[1]), float((outs[0] - outs[1]).abs().max()))
