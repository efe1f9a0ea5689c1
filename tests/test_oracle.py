"""CPU suite: pins oracle/ (torch restatement + plain-C restatement) against the golden vectors
written by the unmodified reference (oracle/make_golden.py) and against the reference's own two
data fixtures."""
import numpy as np
import torch

from oracle import c_oracle as CO
from oracle import ref_torch as R


def _affine(scale_factor):
    _, c2f = R.frustum_transform(R.intrinsic_matrix(), scale_factor)
    return [float(c2f[k, k]) for k in range(3)], [float(c2f[k, 3]) for k in range(3)]


def _dense(idx, bits, shape):
    g = np.zeros(int(np.prod(shape)), np.float32)
    g[idx] = bits.view(np.float32)
    return g.reshape(shape)


def _check_canonical_vs_reference(canon, ref):
    """The canonical order (sequential 8-fold self-sum) may differ from the
    CPU reference only inside the vectorised-sum remainder (last numel % 64 elements on AVX-512,
    see oracle/svr_oracle.c) and there by at most 2 ulp."""
    a, b = canon.reshape(-1).view(np.int32).astype(np.int64), ref.reshape(-1).view(np.int32).astype(np.int64)
    bad = np.flatnonzero(a != b)
    if bad.size:
        assert bad.min() >= a.size - a.size % 64
        assert np.abs(a[bad] - b[bad]).max() <= 2


def test_frustum_constants():
    # diag 20, offsets (69.0655, 51.7450, -8), dims (139,104,112) at scale 1
    dims, c2f = R.frustum_transform(R.intrinsic_matrix(), 1)
    assert [int(d) for d in dims] == [139, 104, 112]
    assert float(c2f[0, 0]) == 20.0
    np.testing.assert_allclose([float(c2f[k, 3]) for k in range(3)], [69.06552124, 51.74501419, -8.0], rtol=1e-7)
    dims2, c2f2 = R.frustum_transform(R.intrinsic_matrix(), 2)
    assert [int(d) for d in dims2] == [70, 52, 56] and float(c2f2[1, 1]) == 10.0


def test_unproject_and_voxelize_bit_exact(golden):
    g = golden["projection"]
    for tag in ("p1", "p2", "p3"):
        dims = g[f"{tag}_dims"]
        scale = int(g[f"{tag}_scale"])
        depth = g[f"{tag}_depth"]
        a, t = _affine(scale)
        f32 = np.float32
        # C restatement
        pts_grid = CO.unproject(depth, f32(R.FOCAL), f32(R.CX), f32(R.CY), a, t, dims, norm=False)
        assert np.array_equal(pts_grid.view(np.uint32), g[f"{tag}_pc_grid"].view(np.uint32)), tag
        pts = CO.unproject(depth, f32(R.FOCAL), f32(R.CX), f32(R.CY), a, t, dims, norm=True)
        assert np.array_equal(pts.view(np.uint32), g[f"{tag}_pc_norm"].view(np.uint32)), tag
        want = _dense(g[f"{tag}_raw_idx"], g[f"{tag}_raw_val_bits"], (2, *dims))
        got = CO.pc_voxels(pts, dims, tail_start="avx512")
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), tag
        _check_canonical_vs_reference(CO.pc_voxels(pts, dims), want)
        # torch restatement
        tp = R.norm_grid_space(R.depthmap_to_gridspace(torch.from_numpy(depth), R.intrinsic_matrix(), scale),
                               torch.from_numpy(dims))
        assert np.array_equal(tp.numpy().view(np.uint32), pts.view(np.uint32)), tag
        tg = R.pc_voxels(tp, torch.from_numpy(dims)).numpy()
        assert np.array_equal(tg.view(np.uint32), want.view(np.uint32)), tag


def test_known_answer_fixtures(golden):
    """The reference's own depth_grid.npz (exact) and diffable_depth_grid.npz (support exact,
    values to 1.2e-4: it was produced by a different accumulation order) -- SURVEY.md section 4."""
    k = golden["known_answer"]
    dims = np.array([139, 104, 112], np.int64)
    a, t = _affine(1)
    depth = k["depth"][None]
    f32 = np.float32
    pg = CO.unproject(depth, f32(R.FOCAL), f32(R.CX), f32(R.CY), a, t, dims, norm=False)[0]
    r = np.round(pg).astype(np.int64)
    hard = np.zeros(tuple(dims), np.float64)
    hard[r[:, 0], r[:, 1], r[:, 2]] = 1
    assert np.array_equal(np.flatnonzero(hard.reshape(-1)), k["hard_idx"])
    pn = CO.unproject(depth, f32(R.FOCAL), f32(R.CX), f32(R.CY), a, t, dims, norm=True)
    soft = CO.pc_voxels(pn, dims)[0].reshape(-1)
    assert np.array_equal(np.flatnonzero(soft), k["soft_idx"])
    assert np.abs(soft[k["soft_idx"]] - k["soft_val"]).max() < 2e-4
    # and bit-exact against the reference run in the build container
    assert np.array_equal(np.flatnonzero(soft), k["ref_soft_idx"])
    assert np.array_equal(soft[k["ref_soft_idx"]].view(np.uint32), k["ref_soft_val_bits"])


def test_blur_and_project_forward(golden):
    g = golden["projection"]
    for tag in ("b1", "b2"):
        dims, ks, sg = g[f"{tag}_dims"], [int(v) for v in g[f"{tag}_ks"]], g[f"{tag}_sigma"]
        raw = CO.pc_voxels(g[f"{tag}_pts"], dims, tail_start="avx512")
        assert np.array_equal(raw.view(np.uint32), g[f"{tag}_raw"].view(np.uint32))
        _check_canonical_vs_reference(CO.pc_voxels(g[f"{tag}_pts"], dims), g[f"{tag}_raw"])
        taps = [CO.gauss_taps(float(sg[i]), ks[i]) for i in range(3)]
        np.testing.assert_allclose(taps[0], g[f"{tag}_taps_w"], rtol=2e-7)
        np.testing.assert_allclose(taps[1], g[f"{tag}_taps_h"], rtol=2e-7)
        np.testing.assert_allclose(taps[2], g[f"{tag}_taps_d"], rtol=2e-7)
        occ = CO.blur(raw, *taps)
        np.testing.assert_allclose(occ, g[f"{tag}_occ"][:, 0], atol=1e-6, rtol=0)
        # torch restatement incl. gradients
        p = torch.from_numpy(g[f"{tag}_pts"]).requires_grad_(True)
        s = torch.from_numpy(sg.copy()).requires_grad_(True)
        o = R.project_forward(p, torch.from_numpy(dims), s, ks)
        np.testing.assert_allclose(o.detach().numpy(), g[f"{tag}_occ"], atol=1e-6, rtol=0)
        (o * torch.from_numpy(g[f"{tag}_wgt"])).sum().backward()
        np.testing.assert_allclose(s.grad.numpy(), g[f"{tag}_dsigma"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(p.grad.numpy(), g[f"{tag}_dpts"], rtol=1e-4, atol=1e-4)


def _ifnet_case(golden, net_res):
    g = golden[f"ifnet{net_res}"]
    sd = R.synthetic_state_dict(100 + net_res, net_res)
    return g, sd, torch.from_numpy(g["x"]), torch.from_numpy(g["pts"]), torch.from_numpy(g["occ"])


def test_ifnet_torch_restatement(golden):
    for net_res in (128, 32):
        g, sd, x, pts, occ = _ifnet_case(golden, net_res)
        for mode in ("train", "eval"):
            sdm = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k)
                   for k, v in sd.items()}
            xx = x.clone().requires_grad_(True)
            pp = pts.clone().requires_grad_(True)
            logits, feat, _ = R.ifnet_forward(sdm, xx, pp, net_res, training=(mode == "train"), return_all=True)
            np.testing.assert_allclose(logits.detach().numpy(), g[f"{mode}_logits"], rtol=1e-4, atol=1e-4)
            np.testing.assert_allclose(feat.detach().numpy()[:, :, 0, :, :8], g[f"{mode}_feat_head"],
                                       rtol=1e-4, atol=1e-5)
            loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, occ, reduction="none").sum(-1).mean()
            loss.backward()
            np.testing.assert_allclose(float(loss), float(g[f"{mode}_loss"]), rtol=1e-5)
            np.testing.assert_allclose(xx.grad.numpy(), g[f"{mode}_dx"], rtol=1e-3, atol=1e-4)
            np.testing.assert_allclose(pp.grad.numpy(), g[f"{mode}_dpts"], rtol=1e-3, atol=1e-3)
            np.testing.assert_allclose(sdm["fc_out.weight"].grad.numpy(), g[f"{mode}_d_fc_out_w"], rtol=1e-3, atol=1e-4)
            np.testing.assert_allclose(sdm["fc_0.weight"].grad.numpy()[:4], g[f"{mode}_d_fc_0_w_head"],
                                       rtol=1e-3, atol=1e-4)
            np.testing.assert_allclose(sdm["fc_1.bias"].grad.numpy(), g[f"{mode}_d_fc_1_b"], rtol=1e-3, atol=1e-4)


def test_ifnet_c_restatement(golden):
    """Plain-C sampler + decoder on the torch encoder's volumes vs the reference's logits/features."""
    for net_res, delta, ac in ((128, R.DISPLACEMENT_128, False), (32, R.DISPLACEMENT_32, True)):
        g, sd, x, pts, _ = _ifnet_case(golden, net_res)
        with torch.no_grad():
            vols = R.encoder_volumes({k: v.clone() for k, v in sd.items()}, x, net_res, training=False)
        sdn = {k: v.numpy() for k, v in sd.items()}
        for b in range(x.shape[0]):
            feat = CO.sample_features([v[b].numpy() for v in vols], pts[b].numpy(), delta, ac)
            c_tot = feat.shape[1] // 7
            head = feat[:8].reshape(8, c_tot, 7).transpose(1, 2, 0)          # (C,7,8)
            np.testing.assert_allclose(head, g["eval_feat_head"][b], rtol=1e-5, atol=1e-6)
            logits = CO.decoder(feat, sdn)
            np.testing.assert_allclose(logits, g["eval_logits"][b], rtol=1e-4, atol=1e-4)


def test_make_3d_grid_and_dense_eval(golden):
    for net_res in (128, 32):
        g, sd, x, _, _ = _ifnet_case(golden, net_res)
        res = g["grid_res"]
        s = [int(2 * r) for r in res]
        assert np.array_equal(CO.make_3d_grid(-0.5, 0.5, *s).view(np.uint32), g["grid_pts"].view(np.uint32))
        assert torch.equal(R.make_3d_grid((-0.5,) * 3, (0.5,) * 3, res, 2), torch.from_numpy(g["grid_pts"]))
        val = R.evaluate_on_grid({k: v.clone() for k, v in sd.items()}, x[:1], res, 2, chunk=128, net_res=net_res)
        np.testing.assert_allclose(val, g["grid_val"], rtol=1e-5, atol=1e-6)
