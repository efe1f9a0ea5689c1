"""On-device iso-surface extraction for ``implicit_to_mesh`` (reference: model/ifnet.py:232-234 ->
util/visualize.py:23-25 ``marching_cubes(1 - occupancy, level)`` + ``export_obj`` on the CPU, after a 67 MB
device-to-host copy of the 256^3 grid).

Marching cubes on the device the grid already lives on (torch tensor ops: this is the boundary of the hot path, not
one of its kernels), so that only the mesh crosses PCIe.  The 256-case table is GENERATED at import (no table of the
reference's third-party ``marching_cubes`` package is available offline, and none is copied): on every cube face the
crossed edges are joined pairwise -- on an ambiguous face (two diagonal corners inside) so that each inside corner is
cut off on its own, a rule that depends on the face's corners only, hence the same in both cells sharing the face: the
surface is watertight --, the joined edges form closed loops around the cube, and each loop is fan-triangulated with
its normal pointing from the inside (value < level) to the outside.  Vertices sit on the grid edges at the linear
interpolation of the level, in index coordinates (first array axis = x), one per crossed edge (shared by the
triangles around it), like the reference's mesher."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

# cube corner k = (k & 1, (k >> 1) & 1, (k >> 2) & 1) along (x, y, z); edges as (corner a, corner b) with a < b
_CORNERS = np.array([[(k >> 0) & 1, (k >> 1) & 1, (k >> 2) & 1] for k in range(8)], dtype=np.int64)
_EDGES = [(a, b) for a in range(8) for b in range(a + 1, 8) if bin(a ^ b).count("1") == 1]        # 12 edges
_EDGE_ID = {e: i for i, e in enumerate(_EDGES)}
_EDGE_AXIS = np.array([int(np.log2(a ^ b)) for a, b in _EDGES], dtype=np.int64)
_EDGE_BASE = np.array([_CORNERS[a] for a, _ in _EDGES], dtype=np.int64)                             # lower corner of each edge
# faces: 4 corners in cyclic order
_FACES = []
for axis in range(3):
    for side in range(2):
        u, v = [a for a in range(3) if a != axis]
        cyc = []
        for du, dv in ((0, 0), (1, 0), (1, 1), (0, 1)):
            c = [0, 0, 0]
            c[axis], c[u], c[v] = side, du, dv
            cyc.append(c[0] | c[1] << 1 | c[2] << 2)
        _FACES.append(cyc)


def _edge(a: int, b: int) -> int:
    return _EDGE_ID[(min(a, b), max(a, b))]


def _share_face(e0: int, e1: int) -> bool:
    c = set(_EDGES[e0]) | set(_EDGES[e1])
    return any(c <= set(f) for f in _FACES)


def _triangulate(loop):
    """Triangles (same orientation as the loop) of the polygon through the loop's edge vertices such that no DIAGONAL joins
    two vertices lying on a common cube face: such a chord would lie in that (ambiguous) face, where the neighbouring cell
    may put the same chord -- an edge with four triangles.  Polygons have at most 7 vertices: all triangulations are
    enumerated, the first admissible one is taken (a fan if there is none)."""
    n = len(loop)

    def rec(idx):                      # idx: indices (into loop) of a sub-polygon, in order
        if len(idx) == 3:
            yield [tuple(idx)]
            return
        a, b = idx[0], idx[-1]         # the side (b, a) belongs to the sub-polygon; pick its apex
        for k in range(1, len(idx) - 1):
            m = idx[k]
            ok = True
            for (u, v) in ((a, m), (m, b)):
                adjacent = (abs(u - v) == 1) or ({u, v} == {0, n - 1})
                if not adjacent and _share_face(loop[u], loop[v]):
                    ok = False
            if not ok:
                continue
            lefts = list(rec(idx[:k + 1])) if k >= 2 else [[]]
            rights = list(rec(idx[k:])) if len(idx) - k >= 3 else [[]]
            for lt in lefts:
                for rt in rights:
                    yield lt + [(a, m, b)] + rt

    for tri in rec(list(range(n))):
        return [(loop[i], loop[j], loop[k]) for i, j, k in tri]
    return [(loop[0], loop[i], loop[i + 1]) for i in range(1, n - 1)]


def _build_table():
    tris = [[] for _ in range(256)]
    face_normal = []
    for axis in range(3):
        for side in range(2):
            nrm = np.zeros(3)
            nrm[axis] = 1.0 if side else -1.0
            face_normal.append(nrm)
    for case in range(256):
        inside = [(case >> k) & 1 for k in range(8)]
        link = {}                                   # crossed edge -> [(neighbour on the surface polygon, face they share)]

        def join(e0, e1, face):
            link.setdefault(e0, []).append((e1, face))
            link.setdefault(e1, []).append((e0, face))

        for fi, cyc in enumerate(_FACES):
            st = [inside[c] for c in cyc]
            crossed = [i for i in range(4) if st[i] != st[(i + 1) % 4]]          # face edge i joins cyc[i], cyc[i+1]
            fe = [_edge(cyc[i], cyc[(i + 1) % 4]) for i in range(4)]
            if len(crossed) == 2:
                join(fe[crossed[0]], fe[crossed[1]], fi)
            elif len(crossed) == 4:
                # ambiguous face: cut off each INSIDE corner (edges i-1 and i meet at corner cyc[i])
                for i in range(4):
                    if st[i]:
                        join(fe[(i - 1) % 4], fe[i], fi)
        seen = set()
        for e0 in sorted(link):
            if e0 in seen:
                continue
            loop, faces, prev_face, cur = [e0], [], None, e0
            seen.add(e0)
            while True:
                # leave `cur` through the face it was not entered by (an edge lies on exactly two faces)
                (n, f) = [nf for nf in link[cur] if nf[1] != prev_face][0]
                faces.append(f)
                if n == e0:
                    break
                loop.append(n)
                seen.add(n)
                prev_face, cur = f, n
            # Orientation, decided on the face shared by the first two edges of the loop: walking A -> B along the polygon's
            # boundary with the polygon on the left (seen from its normal side), the polygon lies towards the cube's interior
            # (-N_face), so the normal is (B - A) x (-N_face); it must point from the inside end point of edge A to its
            # outside end point.
            mid = [(_CORNERS[_EDGES[e][0]] + _CORNERS[_EDGES[e][1]]) / 2.0 for e in loop]
            a, b = _EDGES[loop[0]]
            out_dir = (_CORNERS[b] - _CORNERS[a]) * (1.0 if inside[a] else -1.0)
            nrm = np.cross(mid[1] - mid[0], -face_normal[faces[0]])
            if np.dot(nrm, out_dir) < 0:
                loop = loop[::-1]
            tris[case].extend(_triangulate(loop))
    nmax = max(len(t) for t in tris)
    tab = np.zeros((256, nmax, 3), dtype=np.int64)
    cnt = np.zeros((256,), dtype=np.int64)
    for c, t in enumerate(tris):
        cnt[c] = len(t)
        for i, tri in enumerate(t):
            tab[c, i] = tri
    return tab, cnt


_TRI_TABLE, _TRI_COUNT = _build_table()


def marching_cubes(grid: torch.Tensor, level: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """Iso-surface ``grid == level`` of a (X, Y, Z) tensor on ITS device.  Returns vertices (V, 3) float32 in index
    coordinates and triangles (T, 3) int64 into them; value < level is the inside, normals point outwards."""
    if grid.dim() != 3:
        raise ValueError(f"marching_cubes expects a 3-D grid, got {tuple(grid.shape)}")
    g = grid.detach().to(torch.float32)
    dev = g.device
    X, Y, Z = g.shape
    if min(X, Y, Z) < 2:
        return torch.zeros((0, 3), device=dev), torch.zeros((0, 3), dtype=torch.int64, device=dev)
    inside = g < level
    # ---- vertices: one per crossed grid edge, ids by a running count over (axis, flat index)
    vid, verts, n0 = [], [], 0
    for axis in range(3):
        sl0 = [slice(None)] * 3
        sl1 = [slice(None)] * 3
        sl0[axis], sl1[axis] = slice(0, -1), slice(1, None)
        v0, v1 = g[tuple(sl0)], g[tuple(sl1)]
        cross = inside[tuple(sl0)] != inside[tuple(sl1)]
        ids = torch.cumsum(cross.reshape(-1).to(torch.int64), 0) - 1 + n0
        ids = torch.where(cross.reshape(-1), ids, torch.full_like(ids, -1)).reshape(cross.shape)
        vid.append(ids)
        idx = cross.nonzero()                                  # (n, 3) lower end point of the edge
        a, b = v0[cross], v1[cross]
        t = ((level - a) / (b - a)).clamp(0.0, 1.0)
        p = idx.to(torch.float32)
        p[:, axis] += t
        verts.append(p)
        n0 += int(idx.shape[0])
    vertices = torch.cat(verts, 0)
    # ---- cells
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], device=dev, dtype=torch.int64)
    case = torch.zeros((X - 1, Y - 1, Z - 1), dtype=torch.int64, device=dev)
    for k in range(8):
        dx, dy, dz = (int(c) for c in _CORNERS[k])
        case += inside[dx:X - 1 + dx, dy:Y - 1 + dy, dz:Z - 1 + dz].to(torch.int64) * w[k]
    cnt_t = torch.as_tensor(_TRI_COUNT, device=dev)
    tab_t = torch.as_tensor(_TRI_TABLE, device=dev)
    active = (cnt_t[case] > 0).nonzero()                       # (n, 3) cells the surface passes through
    if active.shape[0] == 0:
        return vertices, torch.zeros((0, 3), dtype=torch.int64, device=dev)
    ccase = case[active[:, 0], active[:, 1], active[:, 2]]
    ntri = cnt_t[ccase]
    nmax = tab_t.shape[1]
    keep = torch.arange(nmax, device=dev)[None, :] < ntri[:, None]                  # (n, nmax)
    local = tab_t[ccase]                                                            # (n, nmax, 3) local edge ids
    cell = active[:, None, None, :].expand(-1, nmax, 3, -1)                         # (n, nmax, 3, 3)
    base = torch.as_tensor(_EDGE_BASE, device=dev)[local] + cell                    # lower end point of each edge
    axis = torch.as_tensor(_EDGE_AXIS, device=dev)[local]
    tri = torch.full(local.shape, -1, dtype=torch.int64, device=dev)
    for a in range(3):
        m = (axis == a) & keep[:, :, None]
        b = base[m]
        tri[m] = vid[a][b[:, 0], b[:, 1], b[:, 2]]
    triangles = tri[keep]
    return vertices, triangles


def export_obj(vertices: torch.Tensor, triangles: torch.Tensor, path) -> None:
    """Wavefront OBJ in the reference mesher's format: ``v x y z`` lines, then 1-based ``f a b c`` lines."""
    v = vertices.detach().cpu().numpy()
    f = triangles.detach().cpu().numpy() + 1
    with open(path, "w") as fh:
        if len(v):
            fh.write("\n".join("v %f %f %f" % (p[0], p[1], p[2]) for p in v))
            fh.write("\n")
        if len(f):
            fh.write("\n".join("f %d %d %d" % (t[0], t[1], t[2]) for t in f))
            fh.write("\n")
