"""GPU parity, IF-Net query path (through the C ABI) against the oracle and the golden vectors of
the unmodified reference.  Tolerances (BASELINE.json north_star): logits 1e-2 relative for the bf16
tensor-core path, measured as max|d| / max|ref| (element-wise relative error is ill-defined near
zero logits); gradients get the same per-tensor bound."""
import numpy as np
import copy
import pytest
import torch
import torch.nn.functional as F

from oracle import c_oracle as CO
from oracle import ref_torch as R

pytestmark = pytest.mark.gpu
TOL_BF16 = 1e-2
# bf16 tier, gradients against the fp32 reference (relative L2 per tensor): north_star gives no number; the bound is
# set by ReLU decisions on bf16-rounded pre-activations (see tests/test_gpu_parity_r2.py); the fp32 tier meets 1e-3
BF16_GRAD_TOL = 2.5e-1


def _record(name, value):
    import json
    from pathlib import Path
    d = Path(__file__).resolve().parent.parent / "gpurun_out"
    if d.is_dir():
        with open(d / "parity_r2.jsonl", "a") as f:
            f.write(json.dumps({"name": name, "value": float(value)}) + "\n")


def _rel(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _net(net_res, sd):
    import svr_b200
    svr_b200.configure(net_res=net_res)
    net = svr_b200.IFNet().cuda()
    net.load_state_dict(sd, strict=False)
    return net


def _case(golden, net_res):
    g = golden[f"ifnet{net_res}"]
    sd = R.synthetic_state_dict(100 + net_res, net_res)
    return g, sd


@pytest.mark.parametrize("net_res", [128, 32])
def test_feature_gather_vs_golden_and_oracle(golden, net_res):
    g, sd = _case(golden, net_res)
    net = _net(net_res, sd).eval()
    x, pts = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["pts"]).cuda()
    with torch.no_grad():
        feat = net.ifnet_feature_extractor(x, pts)                    # (B,C,1,7,N) like the reference
    assert feat.shape[2:4] == (1, 7) and feat.shape[-1] == pts.shape[1]
    head = feat[:, :, 0, :, :8].cpu().numpy()
    ref = g["eval_feat_head"]
    # bf16 volumes + bf16 features: 2^-8 relative to the largest magnitude
    assert np.abs(head - ref).max() / np.abs(ref).max() < 8e-3
    # level 0 is sampled in fp32 from the fp32 grid and only rounded at the end
    assert np.abs(head[:, 0] - ref[:, 0]).max() < 4e-3


def _rel_l2(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def _cos(a, b):
    a, b = torch.as_tensor(a).float().cpu().reshape(-1), torch.as_tensor(b).float().cpu().reshape(-1)
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))


def _bf(t):
    return t.bfloat16().float()


def _bf16_pipeline_reference(sd, x, vols, pts, cot, net_res):
    """torch restatement of the device pipeline WITH its rounding points (volumes, features, weights,
    hidden activations and dz in bf16; fp32 accumulation): F.grid_sample + fp32 matmuls on the GPU.
    Returns logits and every gradient the hot path produces."""
    delta = R.DISPLACEMENT_128 if net_res == 128 else R.DISPLACEMENT_32
    ac = net_res != 128
    xr = x.detach().clone().requires_grad_(True)
    pr = pts.detach().clone().requires_grad_(True)
    vr = [_bf(v.detach()).requires_grad_(True) for v in vols]
    grid = R.stencil_grid(pr, delta)
    feat = torch.cat([F.grid_sample(v, grid, align_corners=ac) for v in [xr] + vr], dim=1)       # (B,C,1,7,N)
    B, C, _, S, N = feat.shape
    f32 = feat[:, :, 0].permute(0, 3, 1, 2).reshape(B * N, C * S)                                    # k = c*7+d
    Fb = _bf(f32.detach())
    W = {k: _bf(sd[k].cuda().reshape(sd[k].shape[0], -1)) for k in ("fc_0.weight", "fc_1.weight", "fc_2.weight")}
    b = {k: sd[k].cuda() for k in ("fc_0.bias", "fc_1.bias", "fc_2.bias", "fc_out.bias")}
    wo = sd["fc_out.weight"].cuda().reshape(-1)
    h0 = torch.relu(Fb @ W["fc_0.weight"].t() + b["fc_0.bias"]); h0b = _bf(h0)
    h1 = torch.relu(h0b @ W["fc_1.weight"].t() + b["fc_1.bias"]); h1b = _bf(h1)
    h2 = torch.relu(h1b @ W["fc_2.weight"].t() + b["fc_2.bias"]); h2b = _bf(h2)
    logits = (h2 @ wo + b["fc_out.bias"]).view(B, N)
    dl = cot.reshape(-1)
    g = {"fc_out.weight": (dl[:, None] * h2b).sum(0), "fc_out.bias": dl.sum()[None]}
    dz2 = _bf(dl[:, None] * wo[None] * (h2b > 0))
    g["fc_2.weight"], g["fc_2.bias"] = dz2.t() @ h1b, dz2.sum(0)
    dz1 = _bf((dz2 @ W["fc_2.weight"]) * (h1b > 0))
    g["fc_1.weight"], g["fc_1.bias"] = dz1.t() @ h0b, dz1.sum(0)
    dz0 = _bf((dz1 @ W["fc_1.weight"]) * (h0b > 0))
    g["fc_0.weight"], g["fc_0.bias"] = dz0.t() @ Fb, dz0.sum(0)
    dF = _bf(dz0 @ W["fc_0.weight"])
    f32.backward(dF)
    g["dx"], g["dpts"] = xr.grad, pr.grad
    for i, v in enumerate(vr):
        g[f"dvol{i + 1}"] = v.grad
    return logits.detach(), g


@pytest.mark.parametrize("net_res", [128, 32])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_logits_and_gradients_vs_golden(golden, net_res, mode):
    """Forward: logits within 1e-2 (max|d|/max|ref|) of the UNMODIFIED reference (measured 1e-3 .. 5e-3).

    Backward, two checks per tensor (north_star states no number for gradients):
      (1) against a torch restatement of the same pipeline with the same bf16 rounding points:
          <= 3e-2 relative L2 (measured 1e-3 .. 1.4e-2; the residue is ReLU flips caused by 1-bf16-ulp
          differences between F.grid_sample and the gather) -- the parity bar for the backward kernels;
      (2) against the unmodified fp32 reference (golden VJPs, full tensors): relative L2 <= BF16_GRAD_TOL and cosine
          similarity >= 0.98.  ReLU decisions taken on bf16-rounded pre-activations flip for ~0.3 % of the units,
          which alone is several per cent relative L2 (the reference's own AMP path has the same property); the
          fp32 tier (test_gpu_parity_r2.py) meets 1e-3 on every tensor."""
    g, sd = _case(golden, net_res)
    net = _net(net_res, sd)
    net.train(mode == "train")
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    pts = torch.from_numpy(g["pts"]).cuda().requires_grad_(True)
    occ = torch.from_numpy(g["occ"]).cuda()
    cot = torch.from_numpy(g["cot"]).cuda()
    logits = net(x, pts)
    assert logits.shape == occ.shape and logits.dtype == torch.float32
    assert _rel(logits.detach(), g[f"{mode}_logits"]) < TOL_BF16
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, occ, reduction="none").sum(-1).mean()
    assert abs(float(loss.detach()) - float(g[f"{mode}_loss"])) / float(g[f"{mode}_loss"]) < TOL_BF16
    logits.backward(cot)
    first = "conv_in" if net_res == 128 else "conv_1"
    checks = {"dx": x.grad, "dpts": pts.grad, f"{first}_w": getattr(net.ifnet_feature_extractor, first).weight.grad}
    for nm in ("fc_out", "fc_2", "fc_1", "fc_0"):
        checks[f"{nm}_w"] = getattr(net, nm).weight.grad[:8]
        checks[f"{nm}_b"] = getattr(net, nm).bias.grad
    for name, got in checks.items():
        c = _cos(got, g[f"{mode}_vjp_{name}"])
        assert c > 0.98, (name, c)
        e = _rel_l2(got, g[f"{mode}_vjp_{name}"])
        _record(f"bf16/{net_res}/{mode}/{name}", e)
        assert e < BF16_GRAD_TOL, (name, e)
    for nm in ("fc_out", "fc_2", "fc_1"):              # full weight-gradient tensors (fc_0 in full: fp32-tier test)
        e = _rel_l2(getattr(net, nm).weight.grad, g[f"{mode}_vjpfull_{nm}_w"])
        _record(f"bf16/{net_res}/{mode}/{nm}_w_full", e)
        assert e < BF16_GRAD_TOL, (nm, e)
    # (1) hot path on the same device volumes vs the bf16-pipeline restatement; in train mode the volumes come from a
    # copy of the module so that the running statistics of `net` stay untouched
    with torch.no_grad():
        vols = copy.deepcopy(net).ifnet_feature_extractor.encode(x.detach())
    net.zero_grad()
    x2 = x.detach().clone().requires_grad_(True)
    p2 = pts.detach().clone().requires_grad_(True)
    v2 = [_bf(v).requires_grad_(True) for v in vols]
    l2 = net.query(x2, v2, p2)
    l2.backward(cot)
    ref_logits, rg = _bf16_pipeline_reference(sd, x, vols, pts, cot, net_res)
    assert _rel(l2.detach(), ref_logits) < 4e-3      # train-mode BatchNorm volumes: measured 2.1e-3
    for nm in ("fc_out", "fc_2", "fc_1", "fc_0"):
        mod = getattr(net, nm)
        assert _rel_l2(mod.weight.grad.reshape(mod.weight.shape[0], -1), rg[f"{nm}.weight"].reshape(mod.weight.shape[0], -1)) < 3e-2, nm
        assert _rel_l2(mod.bias.grad, rg[f"{nm}.bias"]) < 3e-2, nm
    assert _rel_l2(x2.grad, rg["dx"]) < 3e-2
    assert _rel_l2(p2.grad, rg["dpts"]) < 3e-2
    for i, v in enumerate(v2):
        assert _rel_l2(v.grad, rg[f"dvol{i + 1}"]) < 3e-2, i


def test_out_of_range_points_and_zero_padding():
    sd = R.synthetic_state_dict(5, 128)
    net = _net(128, sd).eval()
    g = torch.Generator().manual_seed(1)
    x = torch.rand((1, 1, 16, 16, 16), generator=g)
    pts = torch.cat([(torch.rand((1, 64, 3), generator=g) - 0.5) * 3.0, torch.tensor([[[0.5, 0.5, 0.5], [-0.5, -0.5, -0.5], [10., 0., 0.]]])], 1)
    with torch.no_grad():
        got = net(x.cuda(), pts.cuda()).cpu()
        ref = R.ifnet_forward({k: v.clone() for k, v in sd.items()}, x, pts, 128, training=False)
    assert _rel(got, ref) < TOL_BF16


def test_c_oracle_cross_check_on_device_volumes():
    """Sampler + decoder of the plain-C oracle on the SAME (device-computed) encoder volumes: isolates
    the hot path from cuDNN-vs-oneDNN differences in the encoder."""
    sd = R.synthetic_state_dict(9, 128)
    net = _net(128, sd).eval()
    g = torch.Generator().manual_seed(2)
    x = (torch.rand((2, 1, 32, 16, 24), generator=g) < 0.2).float()
    pts = (torch.rand((2, 150, 3), generator=g) - 0.5) * 1.05
    with torch.no_grad():
        xc = x.cuda()
        vols = net.ifnet_feature_extractor.encode(xc)
        got = net.query(xc, vols, pts.cuda()).cpu().numpy()
    sdn = {k: v.numpy() for k, v in sd.items()}
    for b in range(2):
        feat = CO.sample_features([x[b].numpy()] + [v[b].cpu().numpy() for v in vols], pts[b].numpy(), R.DISPLACEMENT_128, False)
        ref = CO.decoder(feat, sdn)
        assert np.abs(got[b] - ref).max() / np.abs(ref).max() < TOL_BF16


def test_dense_grid_evaluation_vs_golden(golden):
    import svr_b200
    for net_res in (128, 32):
        g, sd = _case(golden, net_res)
        net = _net(net_res, sd).eval()
        svr_b200.configure(num_points=64, batch_size=2)
        val = svr_b200.evaluate_network_on_grid(net, torch.from_numpy(g["x"][:1]).cuda(), g["grid_res"], 2)
        assert val.shape == g["grid_val"].shape
        assert np.abs(val - g["grid_val"]).max() < TOL_BF16
        assert torch.equal(svr_b200.make_3d_grid((-0.5,) * 3, (0.5,) * 3, g["grid_res"], 2), torch.from_numpy(g["grid_pts"]))
    svr_b200.configure(net_res=128, num_points=2048, batch_size=16)


def test_training_config_shape_properties():
    """BASELINE config 2 shape (B=4 x 50k points, 128^3): runs, finite, deterministic forward,
    linear in dlogits (backward), and matches the oracle on a 2k-point subsample."""
    sd = R.synthetic_state_dict(11, 128)
    net = _net(128, sd).eval()
    g = torch.Generator().manual_seed(4)
    x = (torch.rand((4, 1, 128, 128, 128), generator=g) < 0.05).float().cuda()
    pts = (torch.rand((4, 50000, 3), generator=g) - 0.5).cuda()
    with torch.no_grad():
        vols = net.ifnet_feature_extractor.encode(x)
        a = net.query(x, vols, pts)
        b = net.query(x, vols, pts)
    assert torch.isfinite(a).all() and torch.equal(a, b)
    sub = pts[:1, :2000].contiguous()
    with torch.no_grad():
        got = net.query(x[:1], [v[:1] for v in vols], sub).cpu()
    assert torch.equal(got, a[:1, :2000].cpu())
    sdn = {k: v.numpy() for k, v in sd.items()}
    feat = CO.sample_features([x[0].cpu().numpy()] + [v[0].cpu().numpy() for v in vols], sub[0].cpu().numpy(), R.DISPLACEMENT_128, False)
    ref = CO.decoder(feat, sdn)
    assert np.abs(got[0].numpy() - ref).max() / np.abs(ref).max() < TOL_BF16


def test_sorted_aggregated_backward_matches_direct():
    """The spatially sorted processing order + tensor-core scatter (coarse levels) must give the same
    logits (bit-identical: rows are independent) and the same gradients as the direct path: weights and
    point/grid gradients up to fp32 summation order (1e-4), coarse-level volume gradients to 5e-3
    relative L2 (the tensor-core scatter carries the trilinear weights in bf16)."""
    import svr_b200
    ops = svr_b200.ops
    sd = R.synthetic_state_dict(21, 128)
    net = _net(128, sd).eval()
    g = torch.Generator().manual_seed(8)
    x = (torch.rand((2, 1, 64, 48, 32), generator=g) < 0.1).float().cuda()
    pts = ((torch.rand((2, 6000, 3), generator=g) - 0.5) * 1.04).cuda()
    cot = torch.randn((2, 6000), generator=g).cuda()
    with torch.no_grad():
        vols = net.ifnet_feature_extractor.encode(x)
    res = {}
    old = ops.SORT_MIN_POINTS
    try:
        for tag, thr in (("direct", 1 << 30), ("sorted", 1)):
            ops.SORT_MIN_POINTS = thr
            net.zero_grad()
            xx = x.clone().requires_grad_(True)
            pp = pts.clone().requires_grad_(True)
            vv = [v.clone().requires_grad_(True) for v in vols]
            out = net.query(xx, vv, pp)
            out.backward(cot)
            res[tag] = (out.detach(), xx.grad, pp.grad, [v.grad for v in vv], net.fc_0.weight.grad.clone(), net.fc_2.bias.grad.clone())
    finally:
        ops.SORT_MIN_POINTS = old
    a, b = res["direct"], res["sorted"]
    assert torch.equal(a[0], b[0])
    assert _rel_l2(b[1], a[1]) < 1e-4 and _rel_l2(b[2], a[2]) < 1e-4
    for va, vb in zip(a[3], b[3]):
        assert _rel_l2(vb, va) < 5e-3
    assert _rel_l2(b[4], a[4]) < 1e-4 and _rel_l2(b[5], a[5]) < 1e-4


@pytest.mark.parametrize("kernel", ["gather", "box", "box-cpasync"])
def test_fused_dense_evaluator_matches_chunked_query(kernel):
    """svr_dense_eval (lattice generated in-kernel, brick order) against the chunked point path on the
    same make_3d_grid points, incl. non-multiple-of-brick sizes, several scenes and an x-slab.  The gather kernel
    (svr_debug_fq_interp(0)) runs the arithmetic of the point path: 1e-5.  The box kernel (default) interpolates the
    coarse levels on the tensor cores with bf16 trilinear weights: 2e-4 on the occupancy probabilities here (the 1e-2
    bound on the logits against the reference is tests/test_gpu_parity_r2.py::test_dense_eval_256cube_slab_vs_c_oracle);
    its voxel boxes are staged by TMA tensor copies ("box") or by cp.async ("box-cpasync", svr_debug_fq_interp(3))."""
    import svr_b200
    from svr_b200 import _abi
    sd = R.synthetic_state_dict(31, 128)
    net = _net(128, sd).eval()
    g = torch.Generator().manual_seed(12)
    x = (torch.rand((2, 1, 32, 24, 16), generator=g) < 0.15).float().cuda()
    _abi.load().svr_debug_fq_interp({"gather": 0, "box": 1, "box-cpasync": 3}[kernel])
    try:
        for lattice in ((19, 10, 7), (32, 24, 16)):
            dense = net.evaluate_grid(x, lattice)
            assert dense.shape == (2, *lattice)
            pts = svr_b200.make_3d_grid((-0.5,) * 3, (0.5,) * 3, lattice, 1).cuda()
            with torch.no_grad():
                vols = net.ifnet_feature_extractor.encode(x)
                ref = torch.sigmoid(net.query(x, vols, pts[None].expand(2, -1, -1).contiguous())).view(2, *lattice)
            assert float((dense - ref).abs().max()) < (1e-5 if kernel == "gather" else 2e-4)
            slab = net.evaluate_grid(x, lattice, scenes=[1], x_range=(8, 16))
            assert torch.equal(slab[0, 8:16], dense[1, 8:16]) and float(slab[0, :8].abs().max()) == 0.0
    finally:
        _abi.load().svr_debug_fq_interp(1)


@pytest.mark.parametrize("dims,lattice", [((70, 52, 56), (140, 104, 112)), ((48, 80, 64), (48, 80, 64)), ((64, 64, 64), (193, 67, 131))])
def test_box_kernel_dense_evaluator_odd_shapes(dims, lattice):
    """Dense evaluator, box kernel against gather kernel on non-cubic / non-power-of-two scenes (the reference's production
    grids are (139,104,112) and (70,52,56), trainer_scene_net.py:30-31) and lattices that are not multiples of the brick:
    box extents, tensor-map line lengths and the per-tile schedule all vary from tile to tile."""
    from svr_b200 import _abi
    sd = R.synthetic_state_dict(35, 128)
    net = _net(128, sd).eval()
    g = torch.Generator().manual_seed(14)
    x = (torch.rand((1, 1, *dims), generator=g) < 0.1).float().cuda()
    out = {}
    try:
        for mode in (0, 1, 3):
            _abi.load().svr_debug_fq_interp(mode)
            out[mode] = net.evaluate_grid(x, lattice)
            torch.cuda.synchronize()
    finally:
        _abi.load().svr_debug_fq_interp(1)
    assert float((out[1] - out[0]).abs().max()) < 2e-4
    assert float((out[3] - out[0]).abs().max()) < 2e-4


def test_box_kernel_on_sorted_points_matches_gather_kernel():
    """The box kernel on explicit (sorted) query points -- svr_debug_fq_interp(2): row tiles cut at sort-cell group
    boundaries, 16^3 / 8^3 levels interpolated on the tensor cores -- against the gather kernel on a 128^3 scene pair
    with out-of-range points and ragged tiles: logits to 3e-3 of the range, saved features to 5e-3 relative L2 (bf16
    trilinear weights), and a backward pass through the features it saved."""
    import svr_b200
    from svr_b200 import _abi
    sd = R.synthetic_state_dict(33, 128)
    net = _net(128, sd).eval()
    g = torch.Generator().manual_seed(13)
    x = (torch.rand((2, 1, 128, 128, 128), generator=g) < 0.05).float().cuda()
    pts = ((torch.rand((2, 6000, 3), generator=g) - 0.5) * 1.06).cuda()
    pts[1, :1500] = pts[1, 0]                     # 1500 rows in one sort cell: several tiles with the same voxel box
    res = {}
    try:
        for mode in (0, 2):
            _abi.load().svr_debug_fq_interp(mode)
            with torch.no_grad():
                vols = net.ifnet_feature_extractor.encode(x)
                out_i = net.query(x, vols, pts)
            vv = [v.clone().requires_grad_(True) for v in vols]
            out = net.query(x, vv, pts)
            feat = out.grad_fn.saved_tensors[2].float()
            out.backward(torch.ones_like(out))
            res[mode] = (out_i, out.detach(), feat, [v.grad for v in vv])
    finally:
        _abi.load().svr_debug_fq_interp(1)
    a, b = res[0], res[2]
    scale = float(a[0].abs().max())
    assert torch.equal(b[0], b[1])                # inference and training launches agree
    assert float((a[0] - b[0]).abs().max()) / scale < 3e-3
    assert _rel_l2(b[2], a[2]) < 5e-3
    for ga, gb in zip(a[3], b[3]):
        assert _rel_l2(gb, ga) < 2e-2


def test_channels_last_maxpool_matches_torch():
    import svr_b200
    g = torch.Generator().manual_seed(3)
    for shape in ((2, 16, 8, 6, 10), (1, 32, 9, 7, 5), (2, 128, 2, 2, 2)):
        x = torch.randn(shape, generator=g).cuda()
        x[0, :, :2, :2, :2] = 0.25                      # ties: the first maximum must win, like torch
        xa = x.clone().contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
        xb = x.clone().requires_grad_(True)
        ya = svr_b200.ops.maxpool2_channels_last(xa)
        yb = torch.nn.functional.max_pool3d(xb, 2)
        assert torch.equal(ya, yb) and ya.is_contiguous(memory_format=torch.channels_last_3d)
        cot = torch.randn(yb.shape, generator=g).cuda()
        ya.backward(cot)
        yb.backward(cot)
        assert torch.equal(xa.grad, xb.grad)


@pytest.mark.parametrize("co", [16, 32])
def test_first_conv_relu_matches_torch(co):
    """csrc/conv_in.cu against F.conv3d + relu in fp32 (TF32 disabled for the comparison)."""
    import svr_b200
    g = torch.Generator().manual_seed(co)
    x = torch.rand((2, 1, 9, 12, 10), generator=g).cuda()
    w = (torch.randn((co, 1, 3, 3, 3), generator=g) * 0.3).cuda()
    b = (torch.randn((co,), generator=g) * 0.1).cuda()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        xa, wa, ba = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        xb, wb, bb = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        ya = svr_b200.ops.conv1_relu_channels_last(xa, wa, ba)
        yb = torch.relu(torch.nn.functional.conv3d(xb, wb, bb, padding=1))
        assert ya.is_contiguous(memory_format=torch.channels_last_3d)
        assert float((ya - yb).abs().max()) < 1e-5
        cot = torch.randn(yb.shape, generator=g).cuda()
        ya.backward(cot)
        yb.backward(cot)
        for p, q in ((xa, xb), (wa, wb), (ba, bb)):
            assert float((p.grad - q.grad).abs().max() / q.grad.abs().max()) < 1e-4
    finally:
        torch.backends.cudnn.allow_tf32 = prev


@pytest.mark.parametrize("shape,zero_gamma", [((2, 24, 20, 40), False), ((1, 16, 16, 32), False), ((3, 8, 9, 70), False),
                                              ((2, 12, 10, 36), True)])
def test_first_stage_fused_conv_relu_bn(shape, zero_gamma):
    """conv_in + ReLU + BatchNorm3d fused stage (pre-BN activation recomputed, never stored) against the two torch
    modules in fp32 (TF32 off): output, running statistics, and the four parameter gradients.  Tolerances: 1e-4 of
    the largest entry (summation order of the batch statistics / weight gradient only)."""
    import svr_b200
    from svr_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    try:
        B, D, H, W = shape
        g = torch.Generator().manual_seed(sum(shape))
        x = (torch.rand((B, 1, D, H, W), generator=g) < 0.3).float().cuda()
        torch.manual_seed(sum(shape))        # the modules' default initialisation draws from the global generator
        conv = torch.nn.Conv3d(1, 16, 3, padding=1).cuda()
        bn = torch.nn.BatchNorm3d(16).cuda()
        with torch.no_grad():
            # keep every pre-activation away from the ReLU kink: on an occupancy grid a pre-activation is a sum of a subset
            # of the 27 taps, and a draw that lands within 1e-6 of zero flips the ReLU mask between fp32 and the fp64 truth
            # (seen once: fused stage AND cuDNN both 1.05e-3 off the fp64 gradient, min |pre-activation| 9.6e-7)
            for _ in range(16):
                pre = torch.nn.functional.conv3d(x.double(), conv.weight.double(), conv.bias.double(), padding=1)
                if float(pre.abs().min()) > 1e-4:
                    break
                conv.bias.add_(3.1e-4)
            bn.weight.copy_(torch.rand(16, generator=g) + 0.5)
            bn.bias.copy_(torch.randn(16, generator=g) * 0.1)
            if zero_gamma:       # a BN scale of exactly 0 cannot be inverted: the backward falls back to recomputing relu(conv(x))
                bn.weight[3] = 0.0
        conv_r, bn_r = copy.deepcopy(conv), copy.deepcopy(bn)
        cot = torch.randn((B, 16, D, H, W), generator=g).cuda()
        cot_p = torch.randn((B, 16, D // 2, H // 2, W // 2), generator=g).cuda()
        # fused stage incl. the following MaxPool3d(2): y feeds the query path, pooled the next convolution
        y, yp = ops.conv1_relu_bn_channels_last(x, conv, bn, with_pool=True)
        packed = ops.pack_volume(y)
        ((y * cot).sum() + (yp * cot_p).sum()).backward()
        yr = bn_r(torch.relu(conv_r(x)))
        ypr = torch.nn.functional.max_pool3d(yr, 2)
        ((yr * cot).sum() + (ypr * cot_p).sum()).backward()
        assert torch.equal(yp, torch.nn.functional.max_pool3d(y, 2))
        assert torch.equal(packed, y.permute(0, 2, 3, 4, 1).contiguous().bfloat16())     # bf16 copy written by the stage

        def rel(a, b):
            return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))

        assert rel(y, yr) < 1e-4
        assert rel(bn.running_mean, bn_r.running_mean) < 1e-5 and rel(bn.running_var, bn_r.running_var) < 1e-5
        assert int(bn.num_batches_tracked) == int(bn_r.num_batches_tracked) == 1
        # the gradients are judged against an fp64 evaluation of the same two modules (ground truth for both fp32
        # implementations); the message also carries cuDNN's own fp32 deviation and how close any pre-activation is to
        # the ReLU kink, so that a failure says which side moved
        c64, b64 = copy.deepcopy(conv_r).double(), copy.deepcopy(bn_r).double()
        b64.reset_running_stats()
        for p in list(c64.parameters()) + list(b64.parameters()):
            p.grad = None
        pre64 = c64(x.double())
        y64 = b64(torch.relu(pre64))
        ((y64 * cot.double()).sum() + (torch.nn.functional.max_pool3d(y64, 2) * cot_p.double()).sum()).backward()
        names = ("conv.weight", "conv.bias", "bn.weight", "bn.bias")
        mine = (conv.weight.grad, conv.bias.grad, bn.weight.grad, bn.bias.grad)
        cudnn = (conv_r.weight.grad, conv_r.bias.grad, bn_r.weight.grad, bn_r.bias.grad)
        truth = (c64.weight.grad, c64.bias.grad, b64.weight.grad, b64.bias.grad)
        for n, a, b, t in zip(names, mine, cudnn, truth):
            ra, rb = rel(a.double(), t), rel(b.double(), t)
            assert ra < 2e-4, (f"{n}: fused stage vs fp64 {ra:.3e}, torch fp32 vs fp64 {rb:.3e}, "
                               f"min |pre-activation| {float(pre64.detach().abs().min()):.3e}, cudnn.allow_tf32={torch.backends.cudnn.allow_tf32}")
        # eval mode: running statistics, forward only
        bn.eval(), bn_r.eval()
        with torch.no_grad():
            assert rel(ops.conv1_relu_bn_channels_last(x, conv, bn), bn_r(torch.relu(conv_r(x)))) < 1e-5
    finally:
        torch.backends.cudnn.allow_tf32 = True


def test_conv3d_bf16_backward_matches_fp32():
    """Encoder convolution with bf16-operand backward kernels: forward identical to nn.Conv3d, gradients within
    1e-2 relative L2 of the fp32 (TF32 off) gradients (bf16 rounding of activation and incoming gradient)."""
    from svr_b200 import ops
    g = torch.Generator().manual_seed(3)
    conv = torch.nn.Conv3d(16, 32, 3, padding=1).cuda().to(memory_format=torch.channels_last_3d)
    ref = copy.deepcopy(conv)
    x = torch.randn((2, 16, 24, 24, 24), generator=g).cuda().contiguous(memory_format=torch.channels_last_3d)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    cot = torch.randn((2, 32, 24, 24, 24), generator=g).cuda()
    y = ops.conv3d_bf16_backward(xa, conv)
    torch.backends.cudnn.allow_tf32 = False
    try:
        yr = ref(xb)
        (yr * cot).sum().backward()
    finally:
        torch.backends.cudnn.allow_tf32 = True
    (y * cot).sum().backward()
    rel = lambda a, b: float((a - b).norm() / b.norm())
    assert rel(y, yr) < 2e-3                      # TF32 forward
    assert rel(xa.grad, xb.grad) < 1e-2
    assert rel(conv.weight.grad, ref.weight.grad) < 1e-2
    assert rel(conv.bias.grad, ref.bias.grad) < 1e-5


@pytest.mark.parametrize("bf16_backward", [False, True])
def test_conv_bias_relu_fused(bf16_backward):
    """relu(conv(x)) with the fused bias/ReLU forward pass and mask / bias-gradient / cast backward pass against the
    torch modules run with the same (TF32) cuDNN forward, so that both sides take the same ReLU decisions.
    fp32-operand backward: 2e-3 relative L2; bf16-operand backward: 2e-2 (bf16 rounding of gradient, activation
    and weight; measured 1.3e-2)."""
    from svr_b200 import ops
    g = torch.Generator().manual_seed(5)
    conv = torch.nn.Conv3d(32, 64, 3, padding=1).cuda().to(memory_format=torch.channels_last_3d)
    ref = copy.deepcopy(conv)
    x = torch.randn((2, 32, 12, 10, 14), generator=g).cuda().contiguous(memory_format=torch.channels_last_3d)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    cot = torch.randn((2, 64, 12, 10, 14), generator=g).cuda()
    y = ops.conv3d_bias_relu(xa, conv, bf16_backward)
    # reference through the SAME bias-free cuDNN call (a fused-bias convolution may pick another TF32 algorithm)
    yr = torch.relu(torch.nn.functional.conv3d(xb, ref.weight, None, padding=1) + ref.bias.view(1, -1, 1, 1, 1))
    (yr * cot).sum().backward()
    (y * cot).sum().backward()
    rel = lambda a, b: float((a - b).norm() / b.norm())
    tol = 2e-2 if bf16_backward else 2e-3
    assert rel(y, yr) < 1e-5
    assert float(((y > 0) != (yr > 0)).float().mean()) < 1e-4
    assert rel(xa.grad, xb.grad) < tol, rel(xa.grad, xb.grad)
    assert rel(conv.weight.grad, ref.weight.grad) < tol, rel(conv.weight.grad, ref.weight.grad)
    assert rel(conv.bias.grad, ref.bias.grad) < 2e-3


@pytest.mark.parametrize("n_points", [0, 1, 127, 129])
def test_query_ragged_point_counts(golden, n_points):
    """Empty and ragged query sets (not a multiple of the 128-row tile): shapes, finiteness, and agreement with the same
    points evaluated inside a larger batch (rows are independent)."""
    g, sd = _case(golden, 128)
    net = _net(128, sd).eval()
    x = torch.from_numpy(g["x"]).cuda()
    pts = torch.from_numpy(g["pts"]).cuda()
    with torch.no_grad():
        full = net(x, pts)
        part = net(x, pts[:, :n_points].contiguous())
    assert part.shape == (x.shape[0], n_points)
    if n_points:
        assert torch.isfinite(part).all()
        assert float((part - full[:, :n_points]).abs().max()) <= 1e-6 * max(1.0, float(full.abs().max()))
