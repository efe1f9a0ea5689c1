"""A/B of the fused query forward at the bench shape: coarse levels gathered on the CUDA cores (svr_debug_fq_interp(0))
vs interpolated on the tensor cores from voxel boxes in shared memory (default).  Logits / saved features compared,
inference and training-mode launches timed, then one dense-evaluation slab."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import svr_b200
from svr_b200 import _abi, ops

torch.manual_seed(0)
svr_b200.configure(net_res=128)
net = svr_b200.IFNet().cuda().eval()
lib = _abi.load()
B, N, D = 4, 50000, 128
x = (torch.rand(B, 1, D, D, D) < 0.05).float().cuda()
pts = ((torch.rand(B, N, 3) - 0.5) * 1.03).cuda()
res = {}
for interp in (0, 2):
    lib.svr_debug_fq_interp(interp)
    with torch.no_grad():
        vols = net.ifnet_feature_extractor.encode(x)
        out = net.query(x, vols, pts)
    torch.cuda.synchronize()
    print(f"interp={interp}: inference ok", flush=True)
    pp = pts.clone().requires_grad_(True)
    vv = [v.clone().requires_grad_(True) for v in vols]
    out_t = net.query(x, vv, pp)           # training-mode launch (features / activations saved)
    feat = out_t.grad_fn.saved_tensors[2].clone()
    hs = [out_t.grad_fn.saved_tensors[i].clone() for i in (3, 4, 5)]
    torch.cuda.synchronize()
    res[interp] = (out.clone(), out_t.detach().clone(), feat, hs)
    _abi.PROFILE.reset(with_events=True)
    with torch.no_grad():
        for _ in range(5):
            net.query(x, vols, pts)
    for _ in range(5):
        net.query(x, vv, pp)
    torch.cuda.synchronize()
    ev = _abi.PROFILE.events["svr_query_fwd_fused"]
    print(f"interp={interp}: fused fwd inference {sum(a.elapsed_time(b) for a, b in ev[:5]) / 5:.4f} ms, training {sum(a.elapsed_time(b) for a, b in ev[5:]) / 5:.4f} ms", flush=True)
    _abi.PROFILE.reset()
a, b = res[0], res[2]
sc = float(a[0].abs().max())
print("logits  max|d|/max|ref| inference:", float((a[0] - b[0]).abs().max()) / sc, " training:", float((a[1] - b[1]).abs().max()) / sc)
print("inference vs training launch (interp):", float((b[0] - b[1]).abs().max()))
fa, fb = a[2].float(), b[2].float()
print("saved features rel L2:", float((fa - fb).norm() / fa.norm()), " max|d|:", float((fa - fb).abs().max()), " max|ref|:", float(fa.abs().max()))
kp = fa.shape[1]
for c0 in range(0, kp, 64 * 7):
    sl = slice(c0, min(c0 + 64 * 7, kp))
    print(f"  cols {c0:5d}..: rel L2 {float((fa[:, sl] - fb[:, sl]).norm() / fa[:, sl].norm().clamp_min(1e-20)):.3e}")
for i in range(3):
    print(f"saved h{i} rel L2:", float((a[3][i].float() - b[3][i].float()).norm() / a[3][i].float().norm()))
# dense evaluation, one scene, 256^3 lattice: gather kernel (0), box kernel with TMA-staged boxes (1), with cp.async-staged boxes (3)
for interp in (0, 1, 3):
    lib.svr_debug_fq_interp(interp)
    with torch.no_grad():
        g = net.evaluate_grid(x[:1], (256, 256, 256))
    torch.cuda.synchronize()
    _abi.PROFILE.reset(with_events=True)
    with torch.no_grad():
        g = net.evaluate_grid(x[:1], (256, 256, 256))
    torch.cuda.synchronize()
    ev = _abi.PROFILE.events["svr_dense_eval"]
    res[("d", interp)] = g.clone()
    print(f"interp={interp}: dense 256^3 scene {sum(a.elapsed_time(b) for a, b in ev):.2f} ms", flush=True)
    _abi.PROFILE.reset()
print("dense max|d| box(TMA) vs gather:", float((res[("d", 0)] - res[("d", 1)]).abs().max()),
      " box(TMA) vs box(cp.async):", float((res[("d", 1)] - res[("d", 3)]).abs().max()))
lib.svr_debug_fq_interp(1)
