"""GPU parity, projection path (through the C ABI): unprojection and voxel grids bit-exact against
the oracle and the golden vectors written by the unmodified reference; blur within 1e-6 abs
(the reference itself is not self-consistent below that); gradients 1e-4."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as CO
from oracle import ref_torch as R

pytestmark = pytest.mark.gpu


def _proj(dims, ks=(3, 3, 3), sigma=(1.5, 1.5, 1.5)):
    import svr_b200
    return svr_b200.project(tuple(int(d) for d in dims), list(ks), torch.tensor(sigma)).cuda()


def _dense(idx, bits, shape):
    g = np.zeros(int(np.prod(shape)), np.float32)
    g[idx] = bits.view(np.float32)
    return g.reshape(shape)


def _bits(t):
    return t.detach().cpu().contiguous().numpy().view(np.uint32)


def test_golden_unproject_voxelize_bit_exact(golden):
    g = golden["projection"]
    for tag in ("p1", "p2", "p3"):
        dims, scale = g[f"{tag}_dims"], int(g[f"{tag}_scale"])
        proj = _proj(dims)
        depth = torch.from_numpy(g[f"{tag}_depth"]).cuda()
        pc = proj.depthmap_to_gridspace(depth, scale)
        assert np.array_equal(_bits(pc), g[f"{tag}_pc_grid"].view(np.uint32)), tag
        pn = proj.norm_grid_space(pc.clone())
        assert np.array_equal(_bits(pn), g[f"{tag}_pc_norm"].view(np.uint32)), tag
        assert np.array_equal(_bits(proj.depthmap_to_normed_points(depth, scale)), g[f"{tag}_pc_norm"].view(np.uint32))
        want = _dense(g[f"{tag}_raw_idx"], g[f"{tag}_raw_val_bits"], (2, *dims))
        proj.cpu_sum_tail = "avx512"          # the golden run: single-threaded AVX-512 CPU reference
        assert np.array_equal(_bits(proj.pc_voxels(pn)), want.view(np.uint32)), tag
        proj.cpu_sum_tail = None              # canonical order == plain-C oracle
        canon = CO.pc_voxels(g[f"{tag}_pc_norm"], dims)
        assert np.array_equal(_bits(proj.pc_voxels(pn)), canon.view(np.uint32)), tag


def test_known_answer_fixture(golden):
    k = golden["known_answer"]
    proj = _proj((139, 104, 112))
    depth = torch.from_numpy(k["depth"])[None].cuda()
    pg = proj.depthmap_to_gridspace(depth, 1)
    r = np.round(pg[0].cpu().numpy()).astype(np.int64)
    hard = np.zeros((139, 104, 112))
    hard[r[:, 0], r[:, 1], r[:, 2]] = 1
    assert np.array_equal(np.flatnonzero(hard.reshape(-1)), k["hard_idx"])          # depth_grid.npz, exact
    soft = proj.pc_voxels(proj.norm_grid_space(pg.clone()))[0].cpu().numpy().reshape(-1)
    assert np.array_equal(np.flatnonzero(soft), k["soft_idx"])                       # diffable_depth_grid.npz support
    assert np.abs(soft[k["soft_idx"]] - k["soft_val"]).max() < 2e-4
    assert np.array_equal(np.flatnonzero(soft), k["ref_soft_idx"])
    proj.cpu_sum_tail = "avx512"
    soft2 = proj.pc_voxels(proj.norm_grid_space(pg.clone()))[0].cpu().numpy().reshape(-1)
    assert np.array_equal(soft2[k["ref_soft_idx"]].view(np.uint32), k["ref_soft_val_bits"])


@pytest.mark.parametrize("case", ["uniform", "room", "nearwall", "edge", "empty"])
@pytest.mark.parametrize("dims", [(128, 128, 128), (70, 52, 56)])
def test_voxelize_vs_oracle_seeded(case, dims):
    g = torch.Generator().manual_seed(hash((case, dims)) % 1000)
    B, H, W = 3, 64, 96
    if case == "uniform":
        depth = torch.rand((B, H, W), generator=g) * 5.0 + 0.5
    elif case == "room":
        u = torch.arange(W)[None, None, :] / W
        v = torch.arange(H)[None, :, None] / H
        depth = (3.0 + 1.5 * torch.sin(2 * np.pi * u) * torch.cos(2 * np.pi * v)).expand(B, H, W).contiguous()
    elif case == "nearwall":
        depth = 0.5 + 0.01 * torch.rand((B, H, W), generator=g)
    elif case == "edge":   # many points outside the valid box / exactly on planes
        depth = torch.rand((B, H, W), generator=g) * 9.0
        depth[0, :4] = 0.0
    else:
        depth = torch.full((B, H, W), 50.0)     # nothing lands inside the grid
    proj = _proj(dims)
    _, c2f = R.frustum_transform(R.intrinsic_matrix(), 1)
    a = [float(c2f[k, k]) for k in range(3)]
    t = [float(c2f[k, 3]) for k in range(3)]
    want_pts = CO.unproject(depth.numpy(), np.float32(R.FOCAL), np.float32(R.CX), np.float32(R.CY), a, t, dims, norm=True)
    pts = proj.depthmap_to_normed_points(depth.cuda(), 1)
    assert np.array_equal(_bits(pts), want_pts.view(np.uint32))
    want = CO.pc_voxels(want_pts, dims)
    got = proj.pc_voxels(pts)
    assert np.array_equal(_bits(got), want.view(np.uint32))
    # deterministic run to run
    assert torch.equal(got, proj.pc_voxels(pts))
    if case == "empty":
        assert float(got.sum()) == 0.0


@pytest.mark.parametrize("hw", [(37, 50), (21, 33), (16, 68), (1, 4)])
def test_unproject_widths_bit_exact(hw):
    """Depth maps whose width is / is not a multiple of 4 (the 4-pixels-per-thread kernel and the scalar one), grid space
    and normalised: bit-exact against the plain-C oracle."""
    H, W = hw
    dims = (70, 52, 56)
    g = torch.Generator().manual_seed(H * 100 + W)
    depth = torch.rand((3, H, W), generator=g) * 6.0 + 0.3
    proj = _proj(dims)
    _, c2f = R.frustum_transform(R.intrinsic_matrix(), 1)
    a = [float(c2f[k, k]) for k in range(3)]
    t = [float(c2f[k, 3]) for k in range(3)]
    f, cx, cy = np.float32(R.FOCAL), np.float32(R.CX), np.float32(R.CY)
    want_grid = CO.unproject(depth.numpy(), f, cx, cy, a, t, dims, norm=False)
    want_norm = CO.unproject(depth.numpy(), f, cx, cy, a, t, dims, norm=True)
    assert np.array_equal(_bits(proj.depthmap_to_gridspace(depth.cuda(), 1)), want_grid.view(np.uint32))
    assert np.array_equal(_bits(proj.depthmap_to_normed_points(depth.cuda(), 1)), want_norm.view(np.uint32))


def test_voxelize_direct_points_and_ragged():
    dims = (24, 20, 28)
    proj = _proj(dims)
    for n in (0, 1, 7, 1500):
        g = torch.Generator().manual_seed(n)
        pts = (torch.rand((2, n, 3), generator=g) - 0.5) * 1.04
        want = CO.pc_voxels(pts.numpy(), dims) if n else np.zeros((2, *dims), np.float32)
        got = proj.pc_voxels(pts.cuda())
        assert np.array_equal(_bits(got), want.view(np.uint32)), n
    # all points in one cell: long serial sums, order matters
    g = torch.Generator().manual_seed(5)
    pts = torch.rand((1, 4000, 3), generator=g) * 0.01 + 0.1
    want = CO.pc_voxels(pts.numpy(), dims)
    assert np.array_equal(_bits(proj.pc_voxels(pts.cuda())), want.view(np.uint32))
    tiny = torch.rand((1, 4000, 3), generator=g) * 1e-4 + torch.tensor([0.1, 0.1, 0.1 + 0.99 / 27])
    want = CO.pc_voxels(tiny.numpy(), dims)
    assert np.array_equal(_bits(proj.pc_voxels(tiny.cuda())), want.view(np.uint32))


def test_blur_forward_backward_golden(golden):
    g = golden["projection"]
    for tag in ("b1", "b2"):
        dims, ks, sg = g[f"{tag}_dims"], [int(v) for v in g[f"{tag}_ks"]], g[f"{tag}_sigma"]
        proj = _proj(dims, ks, tuple(float(s) for s in sg))
        proj.cpu_sum_tail = "avx512"
        pts = torch.from_numpy(g[f"{tag}_pts"]).cuda().requires_grad_(True)
        raw = proj.pc_voxels(pts)
        assert np.array_equal(_bits(raw), g[f"{tag}_raw"].view(np.uint32))
        occ = proj(pts)
        assert occ.shape == (2, 1, *dims)
        np.testing.assert_allclose(occ.detach().cpu().numpy(), g[f"{tag}_occ"], atol=1e-6, rtol=0)
        (occ * torch.from_numpy(g[f"{tag}_wgt"]).cuda()).sum().backward()
        np.testing.assert_allclose(proj.sigma.grad.cpu().numpy(), g[f"{tag}_dsigma"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(pts.grad.cpu().numpy(), g[f"{tag}_dpts"], rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("shape", [(2, 5, 7, 128), (1, 3, 9, 260), (3, 1, 4, 4), (1, 17, 6, 384), (2, 40, 10, 112),
                                   (1, 64, 8, 128), (1, 70, 3, 516), (2, 6, 5, 30)])
def test_blur_333_shapes_against_oracle(shape):
    """The 3x3x3 blur on grids that exercise both kernels: W % 4 == 0 runs the register-only warp kernel (single and
    several 128-wide x tiles with edge halos, partial tiles, H not a multiple of the 4 rows a warp owns, z cut into
    segments when there are few columns, D = 1), W = 30 the block kernel.  Tolerance 1e-6 abs (test module header)."""
    B, D, H, W = shape
    g = torch.Generator().manual_seed(sum(shape))
    vox = (torch.rand((B, D, H, W), generator=g) * 1.2).clamp(0, 1)     # includes saturated voxels
    proj = _proj((D, H, W), (3, 3, 3), (1.5, 0.9, 2.1))
    got = proj.voxels_smooth(vox.cuda(), proj.smoothing_kernel())
    want = R.voxels_smooth(vox, R.smoothing_kernels(torch.tensor([1.5, 0.9, 2.1]), [3, 3, 3]))
    assert got.shape == want.shape
    np.testing.assert_allclose(got.detach().cpu().numpy(), want.numpy(), atol=1e-6, rtol=0)


def test_depth_gradient_matches_oracle_autograd():
    dims = (70, 52, 56)
    proj = _proj(dims)
    g = torch.Generator().manual_seed(3)
    depth = (torch.rand((2, 40, 56), generator=g) * 4.0 + 0.6)
    wgt = torch.rand((2, 1, *dims), generator=g)
    d1 = depth.clone().cuda().requires_grad_(True)
    pts = proj.norm_grid_space(proj.depthmap_to_gridspace(d1, 2))
    (proj(pts) * wgt.cuda()).sum().backward()
    d2 = depth.clone().requires_grad_(True)
    p2 = R.norm_grid_space(R.depthmap_to_gridspace(d2, R.intrinsic_matrix(), 2), torch.tensor(dims))
    (R.project_forward(p2, torch.tensor(dims), torch.tensor([1.5, 1.5, 1.5]), [3, 3, 3]) * wgt).sum().backward()
    np.testing.assert_allclose(d1.grad.cpu().numpy(), d2.grad.numpy(), rtol=1e-3, atol=1e-2)


def test_full_size_properties():
    """BASELINE config 3 shape (one slice): 256x256 depth maps into 128^3 and 256^3 grids --
    size-independent properties: determinism, range, support bound, batch independence."""
    for dims, scale in (((128, 128, 128), 1), ((256, 256, 256), 0.5)):
        proj = _proj(dims)
        g = torch.Generator().manual_seed(0)
        depth = (torch.rand((4, 256, 256), generator=g) * 5.0 + 0.5).cuda()
        pts = proj.depthmap_to_normed_points(depth, scale)
        a = proj.pc_voxels(pts)
        assert torch.equal(a, proj.pc_voxels(pts))
        assert float(a.min()) >= 0.0 and float(a.max()) <= 1.0
        assert int((a != 0).sum()) <= 8 * 4 * 256 * 256
        b = proj.pc_voxels(pts[1:2])
        assert torch.equal(a[1:2], b)
        # oracle on one map (seconds on CPU)
        want = CO.pc_voxels(pts[:1].cpu().numpy(), dims)
        assert np.array_equal(_bits(a[:1]), want.view(np.uint32))
