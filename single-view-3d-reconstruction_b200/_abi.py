"""ctypes binding of the C-ABI library (include/svr_b200.h).

This is the stub a maintainer of the reference would add: plain pointers and sizes in, int status
out.  There is deliberately NO fallback: if ``libsvr_b200.so`` is missing, cannot be loaded, or a
call returns non-zero, a RuntimeError is raised."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("SVR_B200_LIB", _HERE / "libsvr_b200.so"))

MAX_LEVELS = 6
MAX_TAPS = 15

c_f32p = C.POINTER(C.c_float)
c_i64p = C.POINTER(C.c_int64)
vp = C.c_void_p


class Pyramid(C.Structure):
    _fields_ = [("n_levels", C.c_int), ("channels", C.c_int * MAX_LEVELS), ("dims", (C.c_int * 3) * MAX_LEVELS),
                ("align_corners", C.c_int), ("displacement", C.c_float)]


class DecoderWeights(C.Structure):
    _fields_ = [("w0p", vp), ("w1", vp), ("w2", vp), ("b0", vp), ("b1", vp), ("b2", vp), ("wout", vp),
                ("bout", vp), ("h0", C.c_int), ("h1", C.c_int), ("h2", C.c_int)]


_SIGS = {
    "svr_abi_version": (C.c_int, []),
    "svr_last_error": (C.c_char_p, []),
    "svr_device_info": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "svr_unproject_fwd": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, c_f32p, c_f32p, c_i64p,
                                    C.c_int, vp, vp]),
    "svr_unproject_bwd": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, c_f32p, c_i64p,
                                    C.c_int, vp, vp]),
    "svr_norm_grid_space": (C.c_int, [vp, C.c_int64, c_i64p, vp]),
    "svr_voxelize_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, c_i64p]),
    "svr_voxelize_fwd": (C.c_int, [vp, C.c_int, C.c_int, c_i64p, C.c_double, C.c_int64, vp, vp, vp, C.c_size_t, vp]),
    "svr_voxelize_bwd": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, c_i64p, C.c_double, vp, vp]),
    "svr_blur_fwd": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int,
                               vp, vp, vp, vp]),
    "svr_blur_bwd": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, vp,
                               C.c_int, vp, vp, vp, vp]),
    "svr_feature_kp": (C.c_int, [C.POINTER(Pyramid)]),
    "svr_pack_volume": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64,
                                  C.c_int64, C.c_int64, vp, vp]),
    "svr_pack_volume_halo": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64,
                                       C.c_int64, C.c_int64, vp, vp]),
    "svr_unpack_volume_grad": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64,
                                         C.c_int64, C.c_int64, vp, C.c_int, vp]),
    "svr_pack_w0": (C.c_int, [vp, C.c_int, C.POINTER(Pyramid), vp, vp, vp]),
    "svr_unpack_w0_grad": (C.c_int, [vp, C.c_int, C.POINTER(Pyramid), vp, vp]),
    "svr_pack_matrix": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp]),
    "svr_gather_fwd": (C.c_int, [vp, C.c_int, C.c_int, vp, C.POINTER(vp), C.POINTER(Pyramid), vp, vp]),
    "svr_gather_bwd": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp, C.POINTER(vp), C.POINTER(Pyramid), vp, vp, C.POINTER(vp), vp, vp]),
    "svr_conv1_relu_fwd": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "svr_conv1_relu_bwd_workspace_bytes": (C.c_size_t, [C.c_int]),
    "svr_conv1_relu_bwd": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_size_t, vp]),
    "svr_maxpool2_cl_fwd": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    "svr_maxpool2_cl_bwd": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "svr_sort_points_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "svr_sort_cells_per_scene": (C.c_int, []),
    "svr_sort_points": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, C.c_size_t, vp]),
    "svr_gemm_nt": (C.c_int, [vp, C.c_int64, vp, C.c_int64, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int64, vp,
                              vp, vp, vp, vp]),
    "svr_gemm_tn_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "svr_gemm_tn": (C.c_int, [vp, C.c_int64, vp, C.c_int64, C.c_int, C.c_int, C.c_int, vp, C.c_int64, C.c_int, vp,
                              C.c_size_t, vp]),
    "svr_decoder_head_bwd": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp]),
    "svr_colsum_bf16": (C.c_int, [vp, C.c_int, C.c_int, C.c_int64, vp, C.c_int, vp]),
    "svr_pack_decoder_image": (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
    "svr_query_fwd_fused": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(Pyramid),
                                      C.POINTER(DecoderWeights), vp, vp, vp, C.c_int, vp]),
    "svr_conv1_bn_workspace_bytes": (C.c_size_t, []),
    "svr_conv1_relu_bn_stats": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, vp, vp, vp, vp, vp,
                                          C.c_size_t, vp]),
    "svr_conv1_relu_bn_apply": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]),
    "svr_conv1_relu_bn_bwd": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp,
                                        C.c_size_t, vp]),
    "svr_bias_relu_cl": (C.c_int, [vp, vp, C.c_int64, C.c_int, vp]),
    "svr_widen_bf16": (C.c_int, [vp, C.c_int64, vp, vp]),
    "svr_relu_bwd_cl_workspace_bytes": (C.c_size_t, [C.c_int]),
    "svr_relu_bwd_cl": (C.c_int, [vp, vp, C.c_int64, C.c_int, vp, vp, vp, vp, C.c_size_t, vp]),
    "svr_decoder_bwd_fused": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int64, C.c_int, vp, vp, vp, vp]),
    "svr_debug_fb_trace": (C.c_int, [vp]),
    "svr_debug_fq_trace": (C.c_int, [vp]),
    "svr_debug_fq_interp": (C.c_int, [C.c_int]),
    "svr_debug_fq_trace_block": (C.c_int, [C.c_int]),
    "svr_dense_eval": (C.c_int, [C.c_int, C.c_int, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(Pyramid), C.POINTER(DecoderWeights),
                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    # fp32-accurate tier (csrc/precise.cu)
    "svr_split_bf16": (C.c_int, [vp, C.c_int64, vp, vp, vp]),
    "svr_pack_w0_f32": (C.c_int, [vp, C.c_int, C.POINTER(Pyramid), vp, vp]),
    "svr_gather_fwd_f32": (C.c_int, [vp, C.c_int, C.c_int, vp, C.POINTER(vp), C.POINTER(Pyramid), vp, vp]),
    "svr_gather_bwd_f32": (C.c_int, [vp, C.c_int, C.c_int, vp, C.POINTER(vp), C.POINTER(Pyramid), vp, vp, C.POINTER(vp), vp, vp]),
    "svr_decoder_head_bwd_f32": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp]),
    "svr_colsum_f32": (C.c_int, [vp, C.c_int, C.c_int, C.c_int64, vp, vp]),
}

EXPORTS = tuple(_SIGS)
_lib = None

# kernels launched per entry point (memsets are not kernels of ours and are not counted)
KERNELS_PER_CALL = {
    "svr_unproject_fwd": 1, "svr_unproject_bwd": 1, "svr_norm_grid_space": 1, "svr_voxelize_fwd": 9, "svr_voxelize_bwd": 1,
    "svr_blur_fwd": 3, "svr_blur_bwd": 12, "svr_pack_volume": 1, "svr_pack_volume_halo": 1, "svr_unpack_volume_grad": 1, "svr_pack_w0": 1,
    "svr_unpack_w0_grad": 1, "svr_pack_matrix": 1, "svr_gather_fwd": 1, "svr_gather_bwd": 1, "svr_gather_bwd[tensor-core]": 2,
    "svr_gemm_nt": 1, "svr_gemm_tn": 2,
    "svr_decoder_head_bwd": 2, "svr_colsum_bf16": 2, "svr_query_fwd_fused": 1, "svr_dense_eval": 1, "svr_decoder_bwd_fused": 1, "svr_pack_decoder_image": 1, "svr_sort_points": 4, "svr_bias_relu_cl": 1, "svr_widen_bf16": 1, "svr_relu_bwd_cl": 2, "svr_conv1_relu_fwd": 1, "svr_conv1_relu_bwd": 2, "svr_conv1_relu_bn_stats": 2, "svr_conv1_relu_bn_apply": 1, "svr_conv1_relu_bn_bwd": 4, "svr_maxpool2_cl_fwd": 1, "svr_maxpool2_cl_bwd": 1,
    "svr_split_bf16": 1, "svr_pack_w0_f32": 1, "svr_gather_fwd_f32": 1, "svr_gather_bwd_f32": 1, "svr_decoder_head_bwd_f32": 3, "svr_colsum_f32": 2,
}


class Profile:
    """Optional instrumentation used by bench.py: counts kernel launches per entry point and, when
    ``events`` is enabled, brackets every call with CUDA events on the current torch stream."""

    def __init__(self):
        self.launches = {}
        self.calls = {}
        self.events = None          # None = off; dict name -> [(start, end)]
        self.label = None           # one-shot name override for the next call (entry points used for several kernels)

    def reset(self, with_events=False):
        self.launches, self.calls = {}, {}
        self.events = {} if with_events else None

    def total_launches(self):
        return sum(self.launches.values())

    def kernel_ms(self):
        """name -> (calls, total ms); needs a device synchronize before."""
        out = {}
        for name, evs in (self.events or {}).items():
            out[name] = (len(evs), sum(a.elapsed_time(b) for a, b in evs))
        return out


PROFILE = Profile()


class _Lib:
    """Thin proxy over the CDLL: same call syntax, plus the Profile bookkeeping."""

    def __init__(self, cdll):
        self._cdll = cdll
        for name in _SIGS:
            setattr(self, name, self._wrap(name, getattr(cdll, name)))

    @staticmethod
    def _wrap(name, fn):
        n_k = KERNELS_PER_CALL.get(name, 0)
        if n_k == 0:
            return fn

        def call(*a):
            prof = PROFILE
            key, prof.label = (prof.label or name), None
            prof.launches[key] = prof.launches.get(key, 0) + KERNELS_PER_CALL.get(key, n_k)
            prof.calls[key] = prof.calls.get(key, 0) + 1
            if prof.events is None:
                return fn(*a)
            import torch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*a)
            e1.record()
            prof.events.setdefault(key, []).append((e0, e1))
            return rc
        return call


def load():
    """Load the library (once) and attach the signatures.  Raises if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"svr_b200: CUDA library {LIB_PATH} not found.  Build it with `python __graft_entry__.py` "
            f"(or `python single-view-3d-reconstruction_b200/build.py`).  There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)     # AttributeError here == the library does not match the header
        fn.restype = res
        fn.argtypes = args
    if lib.svr_abi_version() != 1:
        raise RuntimeError(f"svr_b200: ABI version mismatch ({lib.svr_abi_version()} != 1)")
    _lib = _Lib(lib)
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().svr_last_error().decode(errors="replace")
        raise RuntimeError(f"svr_b200 {what} failed (status {rc}): {msg}")


def f32x3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def i64x3(v):
    return (C.c_int64 * 3)(*[int(x) for x in v])


def f32arr(v):
    return (C.c_float * len(v))(*[float(x) for x in v])


def ptr_table(ptrs):
    return (vp * MAX_LEVELS)(*([int(p) if p else None for p in ptrs] + [None] * (MAX_LEVELS - len(ptrs))))


def make_pyramid(channels, dims, align_corners: bool, displacement: float) -> Pyramid:
    p = Pyramid()
    p.n_levels = len(channels)
    for l, (c, d) in enumerate(zip(channels, dims)):
        p.channels[l] = int(c)
        for a in range(3):
            p.dims[l][a] = int(d[a])
    p.align_corners = int(bool(align_corners))
    p.displacement = float(displacement)
    return p
