"""GPU parity, IF-Net query path (through the C ABI) against the oracle and the golden vectors of
the unmodified reference.  Tolerances (BASELINE.json north_star): logits 1e-2 relative for the bf16
tensor-core path, measured as max|d| / max|ref| (element-wise relative error is ill-defined near
zero logits); gradients get the same per-tensor bound."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as CO
from oracle import ref_torch as R

pytestmark = pytest.mark.gpu
TOL_BF16 = 1e-2


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


@pytest.mark.parametrize("net_res", [128, 32])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_logits_and_gradients_vs_golden(golden, net_res, mode):
    g, sd = _case(golden, net_res)
    net = _net(net_res, sd)
    net.train(mode == "train")
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    pts = torch.from_numpy(g["pts"]).cuda().requires_grad_(True)
    occ = torch.from_numpy(g["occ"]).cuda()
    logits = net(x, pts)
    assert logits.shape == occ.shape and logits.dtype == torch.float32
    assert _rel(logits.detach(), g[f"{mode}_logits"]) < TOL_BF16
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, occ, reduction="none").sum(-1).mean()
    assert abs(float(loss) - float(g[f"{mode}_loss"])) / float(g[f"{mode}_loss"]) < TOL_BF16
    loss.backward()
    first = "conv_in" if net_res == 128 else "conv_1"
    checks = {
        "d_fc_out_w": net.fc_out.weight.grad, "d_fc_out_b": net.fc_out.bias.grad, "d_fc_2_w_head": net.fc_2.weight.grad[:8],
        "d_fc_1_b": net.fc_1.bias.grad, "d_fc_0_w_head": net.fc_0.weight.grad[:4], "d_fc_0_b": net.fc_0.bias.grad,
        f"d_{first}_w": getattr(net.ifnet_feature_extractor, first).weight.grad, "dx": x.grad, "dpts": pts.grad,
    }
    for name, got in checks.items():
        r = _rel(got, g[f"{mode}_{name}"])
        assert r < 3 * TOL_BF16, (name, r)


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
