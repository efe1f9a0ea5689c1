"""svr_b200.mesh: the on-device mesher behind implicit_to_mesh (reference: util/visualize.py:23-25, third-party marching
cubes on the CPU).  The case table is generated, so it is checked from first principles: every one of the 256 cube
configurations, a sphere and random fields must give a CLOSED, consistently ORIENTED 2-manifold (every directed edge
exactly once, its reverse exactly once) that encloses the inside (value < level) with outward normals, with the vertices
on the grid edges at the interpolated level."""
import numpy as np
import pytest
import torch

import svr_b200
from svr_b200 import mesh as M


def _closed_oriented(v, f):
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    key = e[:, 0].astype(np.int64) * len(v) + e[:, 1]
    rkey = e[:, 1].astype(np.int64) * len(v) + e[:, 0]
    _, cnt = np.unique(key, return_counts=True)
    return cnt.max() == 1 and np.array_equal(np.sort(key), np.sort(rkey))


def _volume(v, f):
    p0, p1, p2 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    return float(np.einsum("ij,ij->i", p0, np.cross(p1, p2)).sum() / 6.0)


def test_all_256_cube_configurations_are_closed_and_outward():
    assert M._TRI_TABLE.shape == (256, 5, 3) and int(M._TRI_COUNT[0]) == 0 and int(M._TRI_COUNT[255]) == 0
    for case in range(1, 256):
        c = torch.full((4, 4, 4), 1.0)
        for k in range(8):
            if (case >> k) & 1:
                c[1 + (k & 1), 1 + ((k >> 1) & 1), 1 + ((k >> 2) & 1)] = -1.0
        v, f = M.marching_cubes(c, 0.0)
        v, f = v.numpy(), f.numpy()
        assert len(f) > 0 and _closed_oriented(v, f), case
        assert _volume(v, f) > 0, case


def test_sphere_volume_and_vertices_on_the_level_set():
    n = 40
    ax = torch.arange(n, dtype=torch.float32)
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    ctr = torch.tensor([19.3, 20.1, 18.7])
    r = torch.sqrt((X - ctr[0]) ** 2 + (Y - ctr[1]) ** 2 + (Z - ctr[2]) ** 2)
    v, f = M.marching_cubes(r, 12.0)
    v, f = v.numpy(), f.numpy()
    assert _closed_oriented(v, f)
    assert _volume(v, f) == pytest.approx(4 / 3 * np.pi * 12 ** 3, rel=1e-2)
    assert np.abs(np.linalg.norm(v - ctr.numpy(), axis=1) - 12.0).max() < 0.05      # linear interpolation of a distance field
    frac = v - np.floor(v)
    assert ((frac > 0).sum(1) <= 1).all()                                            # every vertex lies on a grid edge


def test_random_fields_are_watertight():
    torch.manual_seed(0)
    g = torch.randn(1, 1, 24, 24, 24)
    g = torch.nn.functional.avg_pool3d(torch.nn.functional.pad(g, (2,) * 6), 5, 1)[0, 0]
    g = torch.nn.functional.pad(g, (1,) * 6, value=10.0)
    for lev in (-0.1, 0.0, 0.13):
        v, f = M.marching_cubes(g, lev)
        assert _closed_oriented(v.numpy(), f.numpy())
        assert _volume(v.numpy(), f.numpy()) > 0
    v, f = M.marching_cubes(torch.full((5, 6, 7), 1.0), 0.5)
    assert v.shape == (0, 3) and f.shape == (0, 3)


def test_export_obj_format(tmp_path):
    g = torch.full((3, 3, 3), 1.0)
    g[1, 1, 1] = 0.0
    v, f = svr_b200.marching_cubes(g, 0.5)
    p = tmp_path / "m.obj"
    svr_b200.export_obj(v, f, p)
    lines = p.read_text().splitlines()
    assert sum(l.startswith("v ") for l in lines) == 6 and sum(l.startswith("f ") for l in lines) == 8     # an octahedron
    idx = [int(t) for l in lines if l.startswith("f ") for t in l.split()[1:]]
    assert min(idx) == 1 and max(idx) == 6


@pytest.mark.gpu
def test_device_mesher_matches_host_and_implicit_to_mesh_writes_obj(tmp_path):
    from oracle import ref_torch as R
    torch.manual_seed(1)
    g = torch.nn.functional.avg_pool3d(torch.randn(1, 1, 40, 36, 44), 3, 1)[0, 0]
    vc, fc = M.marching_cubes(g, 0.05)
    vg, fg = M.marching_cubes(g.cuda(), 0.05)
    assert vg.is_cuda and torch.equal(fg.cpu(), fc) and torch.allclose(vg.cpu(), vc, atol=1e-6)
    svr_b200.configure(net_res=128, precision=16)
    net = svr_b200.IFNet().cuda()
    net.load_state_dict(R.synthetic_state_dict(51, 128), strict=False)
    net.eval()
    x = (torch.rand(1, 1, 32, 32, 32) < 0.2).float().cuda()
    occ = net.evaluate_grid(x, (64, 64, 64), scenes=[0])[0]
    thr = float(1.0 - occ.median())                                  # a level the occupancy field actually crosses
    out = tmp_path / "scene.obj"
    svr_b200.implicit_to_mesh(net, x, (32, 32, 32), thr, str(out), 2)
    lines = out.read_text().splitlines()
    nv, nf = sum(l.startswith("v ") for l in lines), sum(l.startswith("f ") for l in lines)
    v, f = M.marching_cubes(1.0 - occ, thr)
    assert nv == v.shape[0] > 0 and nf == f.shape[0] > 0
