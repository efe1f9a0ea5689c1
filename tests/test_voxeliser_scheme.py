"""CPU: the voxeliser's OWNERSHIP SCHEME (csrc/projection.cu, vox_accumulate_kernel) restated in numpy and checked bit for bit
against the plain-C oracle of the reference's eight `index_put_(accumulate=True)` passes (projection.py:39-80).

The CUDA kernel never adds floating-point numbers atomically.  Points are bucketed by floor cell and kept in point order; the
voxel cell+sh receives pass ps from the cell cell+sh-ps, it is owned by the occupied source cell with the smallest pass index,
and the owner alone sums all contributions pass by pass, point by point -- the reference's serial order.  This test runs that
scheme on the host (float32 adds, one voxel at a time), so the algorithm is pinned independently of the GPU."""
import numpy as np
import pytest

from oracle import c_oracle as CO

f32 = np.float32


def _scheme(pts, dims, eps=1e-6):
    S = np.array(dims)
    sm1 = (S - 1).astype(f32)
    lo, hi = f32(-0.5 + eps), f32(0.5 - eps)
    B = pts.shape[0]
    out = np.zeros((B, *dims), f32)
    for b in range(B):
        p = pts[b].astype(f32)
        ok = np.all((p < hi) & (p > lo), axis=1)
        g = (p + f32(0.5)) * sm1                                   # fp32: add, then multiply (projection.py:49-51)
        f0 = np.floor(g)
        r = (g - f0).astype(f32)
        m = (f32(1.0) - r).astype(f32)
        cells = {}
        for n in np.flatnonzero(ok):                               # point order inside a cell = point index order
            cells.setdefault(tuple(int(v) for v in f0[n]), []).append(n)
        for cell, members in cells.items():
            for sh in range(8):
                s = ((sh >> 2) & 1, (sh >> 1) & 1, sh & 1)
                # owner test: no occupied source cell with a smaller pass index
                if any(tuple(c + a - ((ps >> k) & 1) for c, a, k in zip(cell, s, (2, 1, 0))) in cells for ps in range(sh)):
                    continue
                acc = f32(0.0)
                for ps in range(sh, 8):                            # ascending pass order
                    src = tuple(c + a - ((ps >> k) & 1) for c, a, k in zip(cell, s, (2, 1, 0)))
                    for n in cells.get(src, ()):
                        wx = r[n, 0] if ps & 4 else m[n, 0]
                        wy = r[n, 1] if ps & 2 else m[n, 1]
                        wz = r[n, 2] if ps & 1 else m[n, 2]
                        acc = f32(acc + f32(f32(wx * wy) * wz))
                s8 = acc
                for _ in range(7):                                 # torch.stack([grid] * 8).sum(0): row after row
                    s8 = f32(s8 + acc)
                v = tuple(c + a for c, a in zip(cell, s))
                out[(b, *v)] = min(max(s8, f32(0.0)), f32(1.0))
    return out


@pytest.mark.parametrize("seed,n,dims,spread", [(0, 400, (9, 7, 8), 1.04), (1, 1500, (6, 5, 7), 1.0), (2, 300, (12, 12, 12), 0.2)])
def test_ownership_scheme_equals_serial_passes(seed, n, dims, spread):
    rng = np.random.default_rng(seed)
    pts = ((rng.random((2, n, 3)) - 0.5) * spread).astype(f32)      # spread > 1: some points outside the valid box
    want = CO.pc_voxels(pts, dims)
    got = _scheme(pts, dims)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
