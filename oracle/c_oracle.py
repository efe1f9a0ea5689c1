"""TEST INFRASTRUCTURE -- ctypes binding of oracle/svr_oracle.c (numpy in, numpy out).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build() -> Path:
    so = _HERE / "libsvr_oracle.so"
    src = _HERE / "svr_oracle.c"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-s", "-B", "libsvr_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
    return _LIB


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


def unproject(depth, f, cx, cy, scale3, offset3, dims3, norm=True):
    depth = _f(depth)
    B, H, W = depth.shape
    out = np.empty((B, H * W, 3), np.float32)
    s, o = _f(scale3), _f(offset3)
    d = np.ascontiguousarray(dims3, dtype=np.int64)
    lib().orc_unproject(_p(depth), B, H, W, C.c_float(f), C.c_float(cx), C.c_float(cy), _p(s), _p(o),
                        _p(d, C.c_int64), int(norm), _p(out))
    return out


def pc_voxels(pts, dims3, eps=1e-6, want_pre=False, tail_start=None):
    """tail_start=None: canonical sequential 8-fold sum everywhere; "avx512": emulate the
    single-threaded AVX-512 CPU reference (see svr_oracle.c); or an explicit flat index."""
    pts = _f(pts)
    B, N, _ = pts.shape
    d = np.ascontiguousarray(dims3, dtype=np.int64)
    grid = np.empty((B, int(d[0]), int(d[1]), int(d[2])), np.float32)
    pre = np.empty_like(grid) if want_pre else None
    numel = grid.size
    if tail_start is None:
        tail_start = numel
    elif tail_start == "avx512":
        tail_start = numel - numel % 64
    lib().orc_pc_voxels(_p(pts), B, N, _p(d, C.c_int64), C.c_float(eps), C.c_int64(tail_start), _p(grid),
                        _p(pre) if want_pre else None)
    return (grid, pre) if want_pre else grid


def gauss_taps(sigma, ksize):
    n = (ksize // 2 + 1) - (-ksize // 2 + 1)
    t = np.empty(n, np.float32)
    lib().orc_gauss_taps(C.c_float(sigma), int(ksize), _p(t))
    return t


def blur(grid, taps_w, taps_h, taps_d):
    grid = _f(grid)
    B, D, H, W = grid.shape
    tw, th, td = _f(taps_w), _f(taps_h), _f(taps_d)
    out = np.empty_like(grid)
    lib().orc_blur(_p(grid), B, D, H, W, _p(tw), len(tw), _p(th), len(th), _p(td), len(td), _p(out))
    return out


def sample_features(vols, pts, delta, align_corners):
    """vols: list of (C,D,H,W) fp32 arrays of ONE scene; pts (n,3) -> feat (n, 7*sum C), k=c*7+d."""
    vols = [_f(v) for v in vols]
    pts = _f(pts)
    n = pts.shape[0]
    chans = np.array([v.shape[0] for v in vols], np.int32)
    dims = np.array([s for v in vols for s in v.shape[1:]], np.int32)
    feat = np.empty((n, int(chans.sum()) * 7), np.float32)
    arr = (C.POINTER(C.c_float) * len(vols))(*[_p(v) for v in vols])
    lib().orc_sample_features(arr, _p(chans, C.c_int), _p(dims, C.c_int), len(vols), _p(pts), n,
                              C.c_float(delta), int(align_corners), _p(feat))
    return feat


def decoder(feat, sd):
    """feat (n,K0); sd: dict with fc_0..fc_out weight/bias numpy arrays (reference shapes)."""
    feat = _f(feat)
    n, k0 = feat.shape
    w = {k: _f(np.asarray(v).reshape(np.asarray(v).shape[0], -1)) if k.endswith("weight") else _f(v)
         for k, v in sd.items() if k.startswith("fc_")}
    h0, h1, h2 = w["fc_0.weight"].shape[0], w["fc_1.weight"].shape[0], w["fc_2.weight"].shape[0]
    out = np.empty(n, np.float32)
    lib().orc_decoder(_p(feat), n, k0, h0, h1, h2, _p(w["fc_0.weight"]), _p(w["fc_0.bias"]),
                      _p(w["fc_1.weight"]), _p(w["fc_1.bias"]), _p(w["fc_2.weight"]), _p(w["fc_2.bias"]),
                      _p(w["fc_out.weight"]), _p(w["fc_out.bias"]), _p(out))
    return out


def make_3d_grid(lo, hi, sx, sy, sz):
    out = np.empty((sx * sy * sz, 3), np.float32)
    lib().orc_make_3d_grid(C.c_float(lo), C.c_float(hi), sx, sy, sz, _p(out))
    return out
