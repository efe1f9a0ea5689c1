"""TEST INFRASTRUCTURE -- golden-vector generator.  Runs ONLY in the build container.

Imports the UNMODIFIED reference from /root/reference (read-only) on CPU and writes small
input/output fixtures to tests/golden/.  The reference cannot travel to the GPU box, the
fixtures can.  Recipe: preset ``sys.argv`` (model/ifnet.py:8 parses it at import),
stub the three missing visualisation imports, ``chdir`` to the reference root
(model/projection.py:211 reads a CWD-relative intrinsic file).

    python oracle/make_golden.py            # regenerates everything (spawns itself for --net_res 32)
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import types
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("SVR_REFERENCE", "/root/reference"))
OUT = REPO / "tests" / "golden"
AMBIGUITY_TAU = 1e-4      # see gen_ifnet: rows with a pre-activation this close to zero (relative to the layer's rms) are excluded


def _import_reference(net_res: int):
    import torch  # noqa: F401
    for m in ("marching_cubes", "trimesh", "pyexr"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.argv = ["oracle", "--net_res", str(net_res), "--num_points", "64", "--batch_size", "2"]
    os.chdir(REF)
    sys.path.insert(0, str(REF))
    import model.ifnet as ref_ifnet
    import model.projection as ref_proj
    return ref_ifnet, ref_proj


def _sparse(grid: np.ndarray):
    flat = grid.reshape(-1)
    idx = np.flatnonzero(flat).astype(np.int64)
    return idx, flat[idx]


def gen_projection(ref_proj):
    import torch
    sys.path.insert(0, str(REPO))
    from oracle import ref_torch as R

    torch.set_num_threads(1)  # serial scatter == deterministic mode
    out = {}
    # ---- (P1) depth -> grid-space -> normalised -> raw voxel grid, production dims, 2 scale factors
    for tag, dims, scale, hw, depth_kind in (
        ("p1", (139, 104, 112), 1, (48, 64), "uniform"),
        ("p2", (70, 52, 56), 2, (40, 56), "uniform"),
        ("p3", (139, 104, 112), 1, (48, 64), "nearwall"),
    ):
        g = torch.Generator().manual_seed({"p1": 11, "p2": 12, "p3": 13}[tag])
        if depth_kind == "uniform":
            depth = torch.rand((2,) + hw, generator=g) * 5.0 + 0.5
        else:  # many pixels per voxel: heavy collisions, saturation and long per-voxel sums
            depth = 0.55 + 0.02 * torch.rand((2,) + hw, generator=g)
        mod = ref_proj.project(torch.tensor(dims), [3, 3, 3], torch.tensor([1.5, 1.5, 1.5]))
        pc_grid = mod.depthmap_to_gridspace(depth, scale)
        pc_norm = mod.norm_grid_space(pc_grid.clone())
        raw = mod.pc_voxels(pc_norm)
        idx, val = _sparse(raw.numpy())
        out[f"{tag}_dims"] = np.array(dims, dtype=np.int64)
        out[f"{tag}_scale"] = np.array(scale)
        out[f"{tag}_depth"] = depth.numpy()
        out[f"{tag}_pc_grid"] = pc_grid.numpy()
        out[f"{tag}_pc_norm"] = pc_norm.numpy()
        out[f"{tag}_raw_idx"] = idx
        out[f"{tag}_raw_val_bits"] = val.view(np.uint32)
        # restatement must be bit-identical here, too
        mine_grid = R.depthmap_to_gridspace(depth, R.intrinsic_matrix(), scale)
        assert torch.equal(mine_grid, pc_grid), tag
        mine_norm = R.norm_grid_space(mine_grid.clone(), torch.tensor(dims))
        assert torch.equal(mine_norm, pc_norm), tag
        assert torch.equal(R.pc_voxels(mine_norm, torch.tensor(dims)), raw), tag
        print(tag, "nnz", idx.size, "saturated", int((val == 1).sum()))

    # ---- (P4) small grid, direct normalised points: raw + blur + gradients (sigma, points)
    for tag, dims, ks, sg in (("b1", (24, 20, 28), [3, 3, 3], [1.5, 1.5, 1.5]),
                              ("b2", (18, 22, 17), [5, 3, 7], [1.5, 0.8, 2.0])):
        g = torch.Generator().manual_seed({"b1": 21, "b2": 22}[tag])
        pts = (torch.rand((2, 1500, 3), generator=g) - 0.5) * 1.04   # a few outside the valid box
        pts.requires_grad_(True)
        mod = ref_proj.project(torch.tensor(dims), ks, torch.tensor(sg))
        raw = mod.pc_voxels(pts)
        occ = mod(pts)
        wgt = torch.rand(occ.shape, generator=g)
        (occ * wgt).sum().backward()
        out[f"{tag}_dims"] = np.array(dims, dtype=np.int64)
        out[f"{tag}_ks"] = np.array(ks)
        out[f"{tag}_sigma"] = np.array(sg, dtype=np.float32)
        out[f"{tag}_pts"] = pts.detach().numpy()
        out[f"{tag}_raw"] = raw.detach().numpy()
        out[f"{tag}_occ"] = occ.detach().numpy()
        out[f"{tag}_wgt"] = wgt.numpy()
        out[f"{tag}_dsigma"] = mod.sigma.grad.numpy()
        out[f"{tag}_dpts"] = pts.grad.numpy()
        k1, k2, k3 = mod.smoothing_kernel()
        out[f"{tag}_taps_w"] = k1.detach().numpy().reshape(-1)
        out[f"{tag}_taps_h"] = k2.detach().numpy().reshape(-1)
        out[f"{tag}_taps_d"] = k3.detach().numpy().reshape(-1)
        # restatement check
        p2 = pts.detach().clone().requires_grad_(True)
        s2 = torch.tensor(sg, requires_grad=True)
        occ2 = R.project_forward(p2, torch.tensor(dims), s2, ks)
        assert torch.equal(occ2, occ), tag
        (occ2 * wgt).sum().backward()
        assert torch.allclose(s2.grad, mod.sigma.grad, rtol=1e-6, atol=1e-7)
        assert torch.allclose(p2.grad, pts.grad, rtol=1e-6, atol=1e-7)
        print(tag, "occ sum", float(occ.sum()))
    np.savez_compressed(OUT / "projection.npz", **out)

    # ---- (K) the reference's own data fixtures: known-answer tests
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    import cv2
    from data_processing.distance_to_depth import FromDistanceToDepth  # pyexr is stubbed; class is pure torch
    dist = cv2.imread(str(REF / "data/raw/overfit/00000/distance.exr"), cv2.IMREAD_UNCHANGED)
    dist = torch.from_numpy(np.ascontiguousarray(dist[:, :, 0] if dist.ndim == 3 else dist))
    depth = FromDistanceToDepth(ref_proj.project.get_intrinsic()[0][0])(dist)       # (240,320)
    mod = ref_proj.project(torch.tensor((139, 104, 112)), [3, 3, 3], torch.tensor([1.5, 1.5, 1.5]))
    pc_grid = mod.depthmap_to_gridspace(depth[None], 1)                                # (1,76800,3)
    hard = np.zeros((139, 104, 112))
    r = np.round(pc_grid[0].numpy()).astype(np.int32)
    hard[r[:, 0], r[:, 1], r[:, 2]] = 1
    fix_hard = np.load(REF / "data/processed/overfit/00000/depth_grid.npz")["grid"]
    print("depth_grid.npz mismatches:", int((hard != fix_hard).sum()), "ones:", int(fix_hard.sum()))
    soft = mod.pc_voxels(mod.norm_grid_space(pc_grid.clone())).numpy()[0]
    fix_soft = np.load(REF / "data/processed/overfit/00000/diffable_depth_grid.npz")["grid"]
    print("diffable_depth_grid.npz max|d|:", float(np.abs(soft - fix_soft).max()),
          "support equal:", bool(((soft != 0) == (fix_soft != 0)).all()))
    hi, _ = _sparse(fix_hard)
    si, sv = _sparse(fix_soft)
    oi, ov = _sparse(soft)
    np.savez_compressed(OUT / "known_answer.npz",
                        depth=depth.numpy().astype(np.float32),
                        hard_idx=hi, soft_idx=si, soft_val=sv,          # the reference's fixtures, sparse
                        ref_soft_idx=oi, ref_soft_val_bits=ov.view(np.uint32))  # reference run here, bit pattern


def gen_ifnet(ref_ifnet, net_res: int):
    import torch
    sys.path.insert(0, str(REPO))
    from oracle import ref_torch as R

    torch.set_num_threads(8)
    sd = R.synthetic_state_dict(100 + net_res, net_res)
    net = ref_ifnet.IFNet()
    missing = net.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=False)
    assert not missing.missing_keys, missing
    g = torch.Generator().manual_seed(7 + net_res)
    dims = (32, 24, 16) if net_res == 128 else (20, 12, 16)
    B, N = 2, 300
    x = (torch.rand((B, 1) + dims, generator=g) < 0.15).float() * torch.rand((B, 1) + dims, generator=g)
    pts = (torch.rand((B, N, 3), generator=g) - 0.5) * 1.1
    occ = (torch.rand((B, N), generator=g) < 0.5).float()
    cot = torch.randn((B, N), generator=g)          # fixed upstream gradient for the VJP checks
    out = {"dims": np.array(dims), "x": x.numpy(), "pts": pts.numpy(), "occ": occ.numpy(), "cot": cot.numpy()}
    for mode in ("train", "eval"):
        net.train(mode == "train")
        for k, v in sd.items():   # restore BN running stats mutated by the train pass
            dict(net.state_dict())[k].copy_(v)
        xx = x.clone().requires_grad_(True)
        pp = pts.clone().requires_grad_(True)
        net.zero_grad()
        feat = net.ifnet_feature_extractor(xx, pp)
        logits = net(xx, pp)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, occ, reduction="none").sum(-1).mean()
        loss.backward()
        out[f"{mode}_logits"] = logits.detach().numpy()
        out[f"{mode}_feat_head"] = feat.detach().numpy()[:, :, 0, :, :8]     # (B,C,7,8)
        out[f"{mode}_loss"] = np.array(float(loss))
        out[f"{mode}_dx"] = xx.grad.numpy()
        out[f"{mode}_dpts"] = pp.grad.numpy()
        out[f"{mode}_d_fc_out_w"] = net.fc_out.weight.grad.numpy().copy()
        out[f"{mode}_d_fc_out_b"] = net.fc_out.bias.grad.numpy().copy()
        out[f"{mode}_d_fc_2_w_head"] = net.fc_2.weight.grad.numpy()[:8].copy()
        out[f"{mode}_d_fc_1_b"] = net.fc_1.bias.grad.numpy().copy()
        out[f"{mode}_d_fc_0_w_head"] = net.fc_0.weight.grad.numpy()[:4].copy()
        out[f"{mode}_d_fc_0_b"] = net.fc_0.bias.grad.numpy().copy()
        first = "conv_in" if net_res == 128 else "conv_1"
        out[f"{mode}_d_{first}_w"] = getattr(net.ifnet_feature_extractor, first).weight.grad.numpy().copy()
        # vector-Jacobian products for a FIXED cotangent (isolates the backward kernels from the
        # loss non-linearity, which amplifies bf16 forward error)
        for k, v in sd.items():
            dict(net.state_dict())[k].copy_(v)
        xx = x.clone().requires_grad_(True)
        pp = pts.clone().requires_grad_(True)
        net.zero_grad()
        net(xx, pp).backward(cot)
        out[f"{mode}_vjp_dx"] = xx.grad.numpy()
        out[f"{mode}_vjp_dpts"] = pp.grad.numpy()
        for nm in ("fc_out", "fc_2", "fc_1", "fc_0"):
            out[f"{mode}_vjp_{nm}_w"] = getattr(net, nm).weight.grad.numpy()[:8].copy()
            out[f"{mode}_vjp_{nm}_b"] = getattr(net, nm).bias.grad.numpy().copy()
            if nm != "fc_0":      # full tensors of the small layers for the raw cotangent, too
                out[f"{mode}_vjpfull_{nm}_w"] = getattr(net, nm).weight.grad.numpy().copy()
        out[f"{mode}_vjp_{first}_w"] = getattr(net.ifnet_feature_extractor, first).weight.grad.numpy().copy()
        # ---- "safe" cotangent: zero on the rows whose ReLU decisions are ambiguous.
        # d relu/dz jumps at 0: a row with a pre-activation within rounding noise of zero has an ill-conditioned
        # gradient (two fp32 implementations with different summation orders -- oneDNN here, cuDNN / any GPU kernel
        # there -- disagree on the unit's mask, and ONE flipped unit among ~1e5 moves a gradient tensor by ~3e-3 in
        # relative L2).  The pre-activations come from the reference's own layers (ifnet.py:43-58 re-traced); a row is
        # ambiguous when any |z| < AMBIGUITY_TAU * rms(z of its layer).  Gradients are linear in the cotangent and rows
        # are independent, so zeroing the cotangent on those rows removes them from both sides of the comparison.
        for k, v in sd.items():
            dict(net.state_dict())[k].copy_(v)
        with torch.no_grad():
            f5 = net.ifnet_feature_extractor(x, pts)
            shp = f5.shape
            h = torch.reshape(f5, (shp[0], shp[1] * shp[3], shp[4]))
            amb = torch.zeros((B, N), dtype=torch.bool)
            for fc in (net.fc_0, net.fc_1, net.fc_2):
                z = fc(h)
                amb |= (z.abs() < AMBIGUITY_TAU * z.pow(2).mean().sqrt()).any(1)
                h = torch.relu(z)
        cot_safe = cot * (~amb).float()
        out[f"{mode}_cot_safe"] = cot_safe.numpy()
        print(net_res, mode, "ambiguous rows:", int(amb.sum()), "of", amb.numel())
        for k, v in sd.items():
            dict(net.state_dict())[k].copy_(v)
        xx = x.clone().requires_grad_(True)
        pp = pts.clone().requires_grad_(True)
        net.zero_grad()
        net(xx, pp).backward(cot_safe)
        out[f"{mode}_safe_dx"] = xx.grad.numpy()
        out[f"{mode}_safe_dpts"] = pp.grad.numpy()
        for pn, pv in net.named_parameters():       # decoder: every tensor in full; encoder: first conv, biases, BatchNorm
            if pn.startswith("fc_") or pn.endswith(".bias") or "_bn." in pn or pn.endswith(f"{first}.weight"):
                out[f"{mode}_safe_{pn}"] = pv.grad.numpy().copy()
        # the restatement must agree
        sd2 = {k: v.clone() for k, v in sd.items()}
        mine = R.ifnet_forward(sd2, x, pts, net_res, training=(mode == "train"))
        err = float((mine - logits.detach()).abs().max())
        print(net_res, mode, "restatement max|dlogit|", err, "loss", float(loss))
        assert err <= 1e-5
    # dense-grid evaluation (make_3d_grid ordering + chunking), eval mode
    net.eval()
    for k, v in sd.items():
        dict(net.state_dict())[k].copy_(v)
    res = np.array((6, 5, 4), dtype=np.int32)   # the trainers pass a numpy array (trainer_ifnet.py:53); a tuple breaks ifnet.py:208
    grid_pts = ref_ifnet.make_3d_grid((-0.5,) * 3, (0.5,) * 3, res, 2)
    val = ref_ifnet.evaluate_network_on_grid(net, x[:1], res, 2)
    out["grid_res"] = np.array(res)
    out["grid_pts"] = grid_pts.numpy()
    out["grid_val"] = val
    assert torch.equal(R.make_3d_grid((-0.5,) * 3, (0.5,) * 3, res, 2), grid_pts)
    mine = R.evaluate_on_grid({k: v.clone() for k, v in sd.items()}, x[:1], res, 2, chunk=64 * 2, net_res=net_res)
    print("grid eval max|d|", float(np.abs(mine - val).max()))
    np.savez_compressed(OUT / f"ifnet{net_res}.npz", **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None, choices=[None, "proj", "ifnet128", "ifnet32"])
    a = ap.parse_args()
    OUT.mkdir(parents=True, exist_ok=True)
    if a.only is None:
        for part in ("proj", "ifnet128", "ifnet32"):
            subprocess.check_call([sys.executable, __file__, "--only", part])
        return
    net_res = 32 if a.only == "ifnet32" else 128
    ref_ifnet, ref_proj = _import_reference(net_res)
    if a.only == "proj":
        gen_projection(ref_proj)
    else:
        gen_ifnet(ref_ifnet, net_res)


if __name__ == "__main__":
    main()
