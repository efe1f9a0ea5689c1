"""B200-native drop-in for the reference's ``model/projection.py``.

Same class name, constructor, methods, attributes and state_dict (only ``sigma``) as the
reference; the device work is done by the kernels in csrc/projection.cu through the C ABI.
Citations (file:line) refer to the reference root."""
from __future__ import annotations

import re
from pathlib import Path

import torch
import torch.nn as nn

from .. import ops

# data/intrinsics.txt:1-4 -- used when the CWD-relative file of projection.py:211 is absent
_DEFAULT_INTRINSIC = (277.1281435, 159.5, 119.5)
_FRUSTUM_IMAGE = (320, 240)          # projection.py:156 (hard-coded in the reference)
_DEPTH_RANGE = (0.4, 6.0)            # projection.py:156
_VOXEL = 0.05                        # projection.py:157


class _NormGridSpace(torch.autograd.Function):
    """projection.py:124-132, in place (the reference mutates and returns its argument)."""

    @staticmethod
    def forward(ctx, pc, dims):
        ops.norm_grid_space_(pc, dims)
        ctx.mark_dirty(pc)
        ctx.dims = dims
        return pc

    @staticmethod
    def backward(ctx, g):
        size = torch.tensor([float(d) for d in ctx.dims], device=g.device, dtype=g.dtype)
        return g / size, None


class project(nn.Module):
    """Projection from depth map to point cloud & differentiable voxelisation of the point cloud
    (projection.py:21-37)."""

    def __init__(self, dims, kernel_size, sigma):
        super().__init__()
        self.kernel_size = kernel_size
        self.sigma = torch.nn.Parameter(torch.as_tensor(sigma, dtype=torch.float32).clone())
        self.sigma.requires_grad = True
        self._dims = tuple(int(d) for d in (dims.tolist() if torch.is_tensor(dims) else dims))
        self.register_buffer("vox_size", torch.tensor(self._dims, dtype=torch.int64), persistent=False)
        self.register_buffer("intrinsic", self.get_intrinsic(), persistent=False)
        self._affine_cache = {}
        # None: canonical order (sequential 8-fold self-sum everywhere).
        # "avx512": reproduce the CPU reference's vectorised-sum remainder ("sum(0) remainder").
        self.cpu_sum_tail = None

    # ------------------------------------------------------------------ forward (projection.py:34-37)
    def forward(self, point_cloud):
        return self.voxel_occ_from_pc(point_cloud)

    def voxel_occ_from_pc(self, point_cloud):
        raw = self.pc_voxels(point_cloud)
        smooth = self.voxels_smooth(raw, kernels=self.smoothing_kernel())
        return smooth.unsqueeze(1)

    # ------------------------------------------------------------------ projection.py:39-80
    def pc_voxels(self, points, eps=1e-6):
        numel = points.shape[0] * self._dims[0] * self._dims[1] * self._dims[2]
        tail = -1
        if self.cpu_sum_tail == "avx512":
            tail = numel - numel % 64
        elif isinstance(self.cpu_sum_tail, int):
            tail = self.cpu_sum_tail
        return ops.voxelize(points, self._dims, eps, tail)

    # ------------------------------------------------------------------ projection.py:82-100
    def smoothing_kernel(self):
        dev = self.sigma.device
        ks = self.kernel_size
        out = []
        shapes = ((1, 1, 1, 1, -1), (1, 1, 1, -1, 1), (1, 1, -1, 1, 1))
        if ks[0] == ks[1] == ks[2]:
            # equal tap counts (the reference's default 3/3/3): the three Gaussians as ONE (3, k) evaluation of the same
            # elementwise expressions -- 5 launches instead of ~25 (arange, pow, neg, mul, div, exp, sum, div per axis), which
            # were 0.08 ms of a 0.97 ms project.forward at 64 maps (ncu launch list profiles/r2c_bench_launches_c3.txt).
            # -t^2 is a cached constant; autograd reaches sigma as before.
            cache = self.__dict__.setdefault("_neg_t2", {})
            key = (int(ks[0]), dev)
            if key not in cache:
                t = torch.arange(-ks[0] // 2 + 1., ks[0] // 2 + 1., device=dev)
                cache[key] = (-t ** 2).unsqueeze(0)
            g = torch.exp(cache[key] / (2. * self.sigma ** 2).unsqueeze(1))
            w = g / g.sum(1, keepdim=True)
            return [w[a].view(*shapes[a]) for a in range(3)]
        for a in range(3):
            t = torch.arange(-ks[a] // 2 + 1., ks[a] // 2 + 1., device=dev)
            g = torch.exp(-t ** 2 / (2. * self.sigma[a] ** 2))
            out.append((g / g.sum()).view(*shapes[a]))
        return out

    # ------------------------------------------------------------------ projection.py:102-117
    def voxels_smooth(self, voxels, kernels):
        assert isinstance(kernels, list)
        return ops.blur(voxels, kernels[0], kernels[1], kernels[2])

    # ------------------------------------------------------------------ projection.py:124-148
    def norm_grid_space(self, pc):
        if pc.is_cuda and pc.dtype == torch.float32 and pc.is_contiguous():
            return _NormGridSpace.apply(pc, self._dims)
        raise RuntimeError("svr_b200: norm_grid_space needs a contiguous fp32 CUDA tensor (no CPU path)")

    def un_norm_grid_space(self, point_cloud):
        # off the hot path (visualisation only in the reference); same in-place semantics
        for k in range(3):
            point_cloud[:, :, k] = point_cloud[:, :, k] * self.vox_size[k]
        for k in range(3):
            point_cloud[:, :, k] = point_cloud[:, :, k] + (self.vox_size[k] / 2)
        return point_cloud

    # ------------------------------------------------------------------ projection.py:150-163
    def _affine(self, scale_factor):
        """camera2frustum of projection.py:155-157 -- constant per (intrinsic, scale); the reference
        recomputes inverse + mm every forward, here it is computed once on the host with the same
        torch CPU ops (bit-identical constants) and cached."""
        key = float(scale_factor)
        if key not in self._affine_cache:
            K = self.intrinsic.detach().cpu()
            frustum = self.generate_frustum(list(_FRUSTUM_IMAGE), torch.inverse(K), *_DEPTH_RANGE)
            _, c2f = self.generate_frustum_volume(frustum, _VOXEL * scale_factor)
            self._affine_cache[key] = ([float(c2f[k, k]) for k in range(3)], [float(c2f[k, 3]) for k in range(3)])
        return self._affine_cache[key]

    def _intr(self):
        if not hasattr(self, "_fcxcy"):
            Kc = self.intrinsic.detach().cpu()
            self._fcxcy = (float(Kc[0][0]), float(Kc[0][2]), float(Kc[1][2]))
        return self._fcxcy

    def depthmap_to_gridspace(self, depthmap, scale_factor=1):
        f, cx, cy = self._intr()
        scale, offset = self._affine(scale_factor)
        bs = depthmap.shape[0]
        d = depthmap.reshape(bs, depthmap.shape[-2], depthmap.shape[-1])
        return ops.unproject(d, f, cx, cy, scale, offset, self._dims, normalise=False)

    def depthmap_to_normed_points(self, depthmap, scale_factor=1):
        """depthmap_to_gridspace + norm_grid_space in one kernel (same bits as the two calls)."""
        f, cx, cy = self._intr()
        scale, offset = self._affine(scale_factor)
        bs = depthmap.shape[0]
        d = depthmap.reshape(bs, depthmap.shape[-2], depthmap.shape[-1])
        return ops.unproject(d, f, cx, cy, scale, offset, self._dims, normalise=True)

    # ------------------------------------------------------------------ static helpers
    @staticmethod
    def generate_frustum(image_size, intrinsic_inv, depth_min, depth_max):
        """projection.py:166-179."""
        w, h = image_size[0], image_size[1]
        pts = [[px * d, py * d, d, 1.0] for d in (depth_min, depth_max) for (px, py) in ((0, 0), (0, h), (w, h), (w, 0))]
        corners = torch.tensor(pts, device=intrinsic_inv.device).transpose(1, 0)
        return torch.mm(intrinsic_inv, corners).transpose(1, 0)[:, :3]

    @staticmethod
    def generate_frustum_volume(frustum, voxelsize):
        """projection.py:182-198."""
        lo = [torch.min(frustum[:, a]) / voxelsize for a in range(3)]
        hi = [torch.max(frustum[:, a]) / voxelsize for a in range(3)]
        dims = tuple(torch.ceil(hi[a] - lo[a]) for a in range(3))
        c2f = torch.tensor([[1.0 / voxelsize, 0, 0, -lo[0]], [0, 1.0 / voxelsize, 0, -lo[1]],
                            [0, 0, 1.0 / voxelsize, -lo[2]], [0, 0, 0, 1.0]], device=frustum.device)
        return dims, c2f

    @staticmethod
    def depth_to_camera(depth_map, f, cx, cy):
        """projection.py:201-206 -- flattened camera-space X, Y, Z."""
        d = depth_map.reshape(-1, depth_map.shape[-2], depth_map.shape[-1])
        pts = ops.unproject(d, float(f), float(cx), float(cy), (1.0, 1.0, 1.0), (0.0, 0.0, 0.0), (2, 2, 2), normalise=False)
        flat = pts.reshape(-1, 3)
        return flat[:, 0], flat[:, 1], flat[:, 2]

    @staticmethod
    def get_intrinsic(intrinsic_path=None):
        """projection.py:209-218: focal length, cx, cy from the first two rows of the text matrix."""
        if intrinsic_path is None:
            intrinsic_path = Path("data") / "raw" / "overfit" / "00000" / "intrinsic.txt"
        f, cx, cy = _DEFAULT_INTRINSIC
        p = Path(intrinsic_path)
        if p.exists():
            rows = p.read_text().splitlines()[:2]
            num = r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?"
            r0 = [float(v) for v in re.findall(num, rows[0])]
            r1 = [float(v) for v in re.findall(num, rows[1])]
            f, cx, cy = r0[0], r0[2], r1[2]
        return torch.tensor([[f, 0, cx, 0], [0, f, cy, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float32)
