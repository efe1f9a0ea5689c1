"""Host side above the C ABI: torch.autograd.Functions that hand raw device pointers of torch
tensors to libsvr_b200.so on the current CUDA stream.  PyTorch is plumbing here (device memory,
streams, autograd graph); every kernel on the path is ours.  No CPU fallback: non-CUDA tensors are
an error."""
from __future__ import annotations

import ctypes as C
import functools
from typing import List, Optional, Sequence

import torch

from . import _abi

_BF16 = torch.bfloat16


def _lib():
    return _abi.load()


try:      # raw handle of the current stream of the current device: ~0.3 us instead of ~16 us through torch.cuda.current_stream()
    _raw_stream, _cur_device = torch._C._cuda_getCurrentRawStream, torch._C._cuda_getDevice
except Exception:                                            # pragma: no cover (older / newer torch without the private hooks)
    _raw_stream = _cur_device = None


def _stream() -> int:
    if _raw_stream is not None:
        return _raw_stream(_cur_device())
    return torch.cuda.current_stream().cuda_stream


def _cuda_device_of(args):
    """The one CUDA device all tensor arguments live on (None when there is no CUDA tensor: the callee then raises
    its own "must be a CUDA tensor" error).  Tensors on different GPUs are an error, as they are for torch ops."""
    dev = None
    for a in args:
        if isinstance(a, (list, tuple)):
            d = _cuda_device_of(a)
        elif torch.is_tensor(a) and a.is_cuda:
            d = a.device
        else:
            continue
        if d is None:
            continue
        if dev is not None and d != dev:
            raise RuntimeError(f"svr_b200: tensor arguments live on different devices ({dev} and {d})")
        dev = d
    return dev


def _entry(fn):
    """Every call into the C ABI runs (a) with the tensors' device current, so that the stream handed to the library,
    its per-device state (SM count, function attributes, scratch pool) and the allocations all belong to the device
    that owns the data -- a module on cuda:1 works without torch.cuda.set_device, like torch's own ops; and (b) with
    autocast disabled: the kernels take raw fp32 / bf16 pointers, so a tensor silently produced in half precision by
    an autocast region (trainer_scene_net.py:230 passes precision=16 = native AMP) must never reach them."""
    @functools.wraps(fn)
    def wrapper(*args, **kw):
        dev = _cuda_device_of(args)
        if dev is None:
            return fn(*args, **kw)
        # fast path (a training step makes ~60 of these calls: the two context managers were 1 ms of host time per step)
        same_dev = _cur_device is not None and dev.index == _cur_device()
        if same_dev and not torch.is_autocast_enabled("cuda"):
            return fn(*args, **kw)
        with torch.cuda.device(dev), torch.autocast("cuda", enabled=False):
            return fn(*args, **kw)
    return wrapper


def _dev_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"svr_b200: `{name}` must be a CUDA tensor (got {t.device}); there is no CPU path")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# =================================================================================================
# projection
# =================================================================================================
class _Unproject(torch.autograd.Function):
    """depth (B,H,W) -> points (B,H*W,3); projection.py:150-163,201-206 (+ :124-132 when normalise)."""

    @staticmethod
    @_entry
    def forward(ctx, depth, f, cx, cy, scale, offset, dims, normalise):
        depth = _dev_f32(depth, "depthmap")
        B, H, W = depth.shape
        pts = torch.empty((B, H * W, 3), device=depth.device, dtype=torch.float32)
        _abi.check(_lib().svr_unproject_fwd(depth.data_ptr(), B, H, W, f, cx, cy, _abi.f32x3(scale), _abi.f32x3(offset),
                                            _abi.i64x3(dims), int(normalise), pts.data_ptr(), _stream()), "unproject_fwd")
        ctx.meta = (B, H, W, f, cx, cy, tuple(scale), tuple(dims), int(normalise))
        return pts

    @staticmethod
    @_entry
    def backward(ctx, gpts):
        B, H, W, f, cx, cy, scale, dims, normalise = ctx.meta
        gpts = _dev_f32(gpts, "grad")
        gd = torch.empty((B, H, W), device=gpts.device, dtype=torch.float32)
        _abi.check(_lib().svr_unproject_bwd(gpts.data_ptr(), B, H, W, f, cx, cy, _abi.f32x3(scale), _abi.i64x3(dims),
                                            normalise, gd.data_ptr(), _stream()), "unproject_bwd")
        return gd, None, None, None, None, None, None, None


def unproject(depth, f, cx, cy, scale, offset, dims, normalise=False):
    return _Unproject.apply(depth, float(f), float(cx), float(cy), scale, offset, dims, normalise)


@_entry
def norm_grid_space_(pc: torch.Tensor, dims) -> torch.Tensor:
    """In-place projection.py:124-132 on a contiguous CUDA tensor (B,P,3)."""
    if not (pc.is_cuda and pc.dtype == torch.float32 and pc.is_contiguous()):
        raise RuntimeError("svr_b200: norm_grid_space needs a contiguous fp32 CUDA tensor")
    _abi.check(_lib().svr_norm_grid_space(pc.data_ptr(), pc.numel() // 3, _abi.i64x3(dims), _stream()), "norm_grid_space")
    return pc


class _Voxelize(torch.autograd.Function):
    """project.pc_voxels (projection.py:39-80), bit-exact."""

    @staticmethod
    @_entry
    def forward(ctx, points, dims, eps, tail_start):
        pts = _dev_f32(points, "points")
        B, N, _ = pts.shape
        d = _abi.i64x3(dims)
        grid = torch.empty((B, int(dims[0]), int(dims[1]), int(dims[2])), device=pts.device, dtype=torch.float32)
        need_grad = ctx.needs_input_grad[0]
        sat = torch.empty(((grid.numel() + 31) // 32,), device=pts.device, dtype=torch.int32) if need_grad else None
        ws_bytes = _lib().svr_voxelize_workspace_bytes(B, N, d)
        ws = torch.empty((ws_bytes,), device=pts.device, dtype=torch.uint8)
        _abi.check(_lib().svr_voxelize_fwd(pts.data_ptr(), B, N, d, float(eps), int(tail_start), grid.data_ptr(), _ptr(sat),
                                           ws.data_ptr(), ws_bytes, _stream()), "voxelize_fwd")
        ctx.save_for_backward(pts, sat)
        ctx.meta = (B, N, tuple(int(x) for x in dims), float(eps))
        return grid

    @staticmethod
    @_entry
    def backward(ctx, ggrid):
        pts, sat = ctx.saved_tensors
        B, N, dims, eps = ctx.meta
        ggrid = _dev_f32(ggrid, "grad")
        gp = torch.empty_like(pts)
        _abi.check(_lib().svr_voxelize_bwd(pts.data_ptr(), ggrid.data_ptr(), _ptr(sat), B, N, _abi.i64x3(dims), eps,
                                           gp.data_ptr(), _stream()), "voxelize_bwd")
        return gp, None, None, None


def voxelize(points, dims, eps=1e-6, tail_start=-1):
    return _Voxelize.apply(points, dims, eps, tail_start)


class _Blur(torch.autograd.Function):
    """project.voxels_smooth (projection.py:102-117): taps_w acts on the last axis, taps_d on the first."""

    @staticmethod
    @_entry
    def forward(ctx, grid, taps_w, taps_h, taps_d):
        g = _dev_f32(grid, "voxels")
        tw, th, td = (_dev_f32(t.reshape(-1), "taps") for t in (taps_w, taps_h, taps_d))
        B, D, H, W = g.shape
        out = torch.empty_like(g)
        fused = max(tw.numel(), th.numel(), td.numel()) <= 7      # single-pass kernel: no scratch grids needed
        tmp = out if fused else torch.empty((2,) + tuple(g.shape), device=g.device, dtype=torch.float32)
        t0, t1 = (out, out) if fused else (tmp[0], tmp[1])
        _abi.check(_lib().svr_blur_fwd(g.data_ptr(), B, D, H, W, tw.data_ptr(), tw.numel(), th.data_ptr(), th.numel(),
                                       td.data_ptr(), td.numel(), out.data_ptr(), t0.data_ptr(), t1.data_ptr(),
                                       _stream()), "blur_fwd")
        ctx.save_for_backward(g, tw, th, td)
        ctx.shapes = (taps_w.shape, taps_h.shape, taps_d.shape)
        return out

    @staticmethod
    @_entry
    def backward(ctx, gout):
        g, tw, th, td = ctx.saved_tensors
        gout = _dev_f32(gout, "grad")
        B, D, H, W = g.shape
        gin = torch.empty_like(g)
        gt = torch.empty((tw.numel() + th.numel() + td.numel(),), device=g.device, dtype=torch.float32)
        tmp = torch.empty((4,) + tuple(g.shape), device=g.device, dtype=torch.float32)
        _abi.check(_lib().svr_blur_bwd(g.data_ptr(), gout.data_ptr(), B, D, H, W, tw.data_ptr(), tw.numel(), th.data_ptr(),
                                       th.numel(), td.data_ptr(), td.numel(), gin.data_ptr(), gt.data_ptr(), tmp.data_ptr(),
                                       _stream()), "blur_bwd")
        kw, kh = tw.numel(), th.numel()
        sw, sh, sd = ctx.shapes
        return gin, gt[:kw].reshape(sw), gt[kw:kw + kh].reshape(sh), gt[kw + kh:].reshape(sd)


def blur(grid, taps_w, taps_h, taps_d):
    return _Blur.apply(grid, taps_w, taps_h, taps_d)


# =================================================================================================
# encoder helper: channels-last 2x2x2 max pooling
# =================================================================================================
class _MaxPool2CL(torch.autograd.Function):
    """nn.MaxPool3d(2) on a channels_last_3d activation without leaving NDHWC."""

    @staticmethod
    @_entry
    def forward(ctx, x):
        B, Cc, D, H, W = x.shape
        xp = _dev_f32(x.permute(0, 2, 3, 4, 1), "x")         # NDHWC view of the channels-last tensor (fp32, dense)
        out = torch.empty((B, D // 2, H // 2, W // 2, Cc), device=x.device, dtype=torch.float32)
        idx = torch.empty((out.numel() // 4,), device=x.device, dtype=torch.int32)
        _abi.check(_lib().svr_maxpool2_cl_fwd(xp.data_ptr(), B, D, H, W, Cc, out.data_ptr(), idx.data_ptr(), _stream()), "maxpool_fwd")
        ctx.save_for_backward(idx)
        ctx.shape = (B, Cc, D, H, W)
        return out.permute(0, 4, 1, 2, 3)                  # logical NCDHW, physically channels-last

    @staticmethod
    @_entry
    def backward(ctx, gout):
        (idx,) = ctx.saved_tensors
        B, Cc, D, H, W = ctx.shape
        g = _dev_f32(gout.permute(0, 2, 3, 4, 1), "grad")
        gin = torch.empty((B, D, H, W, Cc), device=gout.device, dtype=torch.float32)
        _abi.check(_lib().svr_maxpool2_cl_bwd(g.data_ptr(), idx.data_ptr(), B, D, H, W, Cc, gin.data_ptr(), _stream()), "maxpool_bwd")
        return gin.permute(0, 4, 1, 2, 3)


class _Conv1ReLU(torch.autograd.Function):
    """relu(Conv3d(1 -> Co, 3, padding=1)(x)) with a channels-last output (first encoder layer)."""

    @staticmethod
    @_entry
    def forward(ctx, x, weight, bias):
        x0 = _dev_f32(x, "x")
        B, _, D, H, W = x0.shape
        Co = weight.shape[0]
        w = _dev_f32(weight.detach().reshape(Co, 27), "weight")
        b = _dev_f32(bias.detach(), "bias") if bias is not None else None
        y = torch.empty((B, D, H, W, Co), device=x0.device, dtype=torch.float32)
        _abi.check(_lib().svr_conv1_relu_fwd(x0.data_ptr(), w.data_ptr(), _ptr(b), B, D, H, W, Co, y.data_ptr(), _stream()), "conv1_relu_fwd")
        ctx.save_for_backward(x0, y, w)
        ctx.has_bias = bias is not None
        ctx.wshape = weight.shape
        return y.permute(0, 4, 1, 2, 3)

    @staticmethod
    @_entry
    def backward(ctx, gy):
        x0, y, w = ctx.saved_tensors
        B, _, D, H, W = x0.shape
        Co = w.shape[0]
        g = _dev_f32(gy.permute(0, 2, 3, 4, 1), "grad")
        gw = torch.empty((Co, 27), device=x0.device, dtype=torch.float32)
        gb = torch.empty((Co,), device=x0.device, dtype=torch.float32)
        gx = torch.empty_like(x0) if ctx.needs_input_grad[0] else None
        nbytes = _lib().svr_conv1_relu_bwd_workspace_bytes(Co)
        ws = torch.empty((nbytes,), device=x0.device, dtype=torch.uint8)
        _abi.check(_lib().svr_conv1_relu_bwd(x0.data_ptr(), y.data_ptr(), g.data_ptr(), w.data_ptr(), B, D, H, W, Co, gw.data_ptr(),
                                             gb.data_ptr(), _ptr(gx), ws.data_ptr(), nbytes, _stream()), "conv1_relu_bwd")
        return gx, gw.view(ctx.wshape), (gb if ctx.has_bias else None)


def conv1_relu_channels_last(x, weight, bias):
    return _Conv1ReLU.apply(x, weight, bias)


class _Conv1ReLUBN(torch.autograd.Function):
    """BatchNorm3d(relu(Conv3d(1 -> 16, 3, padding=1)(x))) without ever storing the pre-BN activation
    (csrc/conv_in_bn.cu): training mode = batch statistics (running statistics updated in place), eval mode =
    running statistics (forward only).  With ``with_pool`` the following nn.MaxPool3d(2) is part of the stage:
    outputs (y, pooled, y_bf16); y then has a single consumer (the query path), so autograd never sums the pooling
    branch into its gradient -- the backward kernels route the pooled gradient through the winner codes on the fly.
    y_bf16 is the gather's bf16 NDHWC copy of y (non-differentiable, see ``pack_volume``)."""

    @staticmethod
    @_entry
    def forward(ctx, x, weight, bias, gamma, beta, running_mean, running_var, training, update_running, momentum, eps, with_pool,
                keep_for_backward):
        x0 = _dev_f32(x, "x")
        B, _, D, H, W = x0.shape
        Co = weight.shape[0]
        dev = x0.device
        w = _dev_f32(weight.detach().reshape(Co, 27), "weight")
        b = _dev_f32(bias.detach(), "bias") if bias is not None else None
        ga = _dev_f32(gamma.detach(), "bn.weight") if gamma is not None else None
        be = _dev_f32(beta.detach(), "bn.bias") if beta is not None else None
        if training:
            mean = torch.empty((Co,), device=dev, dtype=torch.float32)
            invstd = torch.empty((Co,), device=dev, dtype=torch.float32)
            nbytes = _lib().svr_conv1_bn_workspace_bytes()
            ws = torch.empty((nbytes,), device=dev, dtype=torch.uint8)
            _abi.check(_lib().svr_conv1_relu_bn_stats(x0.data_ptr(), w.data_ptr(), _ptr(b), B, D, H, W, Co, float(eps), float(momentum),
                                                      _ptr(running_mean if update_running else None),
                                                      _ptr(running_var if update_running else None), mean.data_ptr(), invstd.data_ptr(),
                                                      ws.data_ptr(), nbytes, _stream()), "conv1_relu_bn_stats")
        else:
            mean = running_mean.detach().float().contiguous()
            invstd = torch.rsqrt(running_var.detach().float() + eps).contiguous()
        y = torch.empty((B, D, H, W, Co), device=dev, dtype=torch.float32)
        y_bf16 = torch.empty((B, D, H, W, Co), device=dev, dtype=_BF16)
        keep = bool(training) and bool(keep_for_backward)        # backward reads y and the ReLU bits instead of recomputing them
        mask = torch.empty((B, D, H, W), device=dev, dtype=torch.int16) if keep else None
        _abi.check(_lib().svr_conv1_relu_bn_apply(x0.data_ptr(), w.data_ptr(), _ptr(b), mean.data_ptr(), invstd.data_ptr(), _ptr(ga), _ptr(be),
                                                  B, D, H, W, Co, y.data_ptr(), y_bf16.data_ptr(), _ptr(mask), _stream()), "conv1_relu_bn_apply")
        pooled = idx = None
        if with_pool:
            pooled = torch.empty((B, D // 2, H // 2, W // 2, Co), device=dev, dtype=torch.float32)
            idx = torch.empty((pooled.numel() // 4,), device=dev, dtype=torch.int32)
            _abi.check(_lib().svr_maxpool2_cl_fwd(y.data_ptr(), B, D, H, W, Co, pooled.data_ptr(), idx.data_ptr(), _stream()), "maxpool_fwd")
        ctx.save_for_backward(x0, w, b, ga, be, mean, invstd, idx, y if keep else None, mask)
        ctx.training = bool(training)
        ctx.wshape = weight.shape
        ctx.has = (bias is not None, gamma is not None, beta is not None)
        ctx.mark_non_differentiable(y_bf16)
        ctx.set_materialize_grads(False)     # no zero-filled stand-ins for the gradients of unused / non-differentiable outputs
        yv = y.permute(0, 4, 1, 2, 3)
        if with_pool:
            return yv, pooled.permute(0, 4, 1, 2, 3), y_bf16
        return yv, y_bf16

    @staticmethod
    @_entry
    def backward(ctx, gy, *rest):
        if not ctx.training:
            raise RuntimeError("svr_b200: the fused conv_in+ReLU+BN stage has no eval-mode backward; use the unfused modules")
        if ctx.needs_input_grad[0]:
            raise RuntimeError("svr_b200: the fused conv_in+ReLU+BN stage does not produce an input gradient")
        x0, w, b, ga, be, mean, invstd, idx, y, mask = ctx.saved_tensors
        gpool = rest[0] if idx is not None else None
        B, _, D, H, W = x0.shape
        Co = w.shape[0]
        dev = x0.device

        def ndhwc(t):
            if t is None:
                return None
            t = t.permute(0, 2, 3, 4, 1)
            return _dev_f32(t, "grad")       # contiguous NDHWC (a no-op for channels-last gradients)

        g, gp = ndhwc(gy), ndhwc(gpool)
        gw = torch.empty((Co, 27), device=dev, dtype=torch.float32)
        gb, gga, gbe = (torch.empty((Co,), device=dev, dtype=torch.float32) for _ in range(3))
        nbytes = _lib().svr_conv1_bn_workspace_bytes()
        ws = torch.empty((nbytes,), device=dev, dtype=torch.uint8)
        _abi.check(_lib().svr_conv1_relu_bn_bwd(x0.data_ptr(), w.data_ptr(), _ptr(b), mean.data_ptr(), invstd.data_ptr(), _ptr(ga), _ptr(be),
                                                _ptr(y), _ptr(mask), _ptr(g), _ptr(gp), _ptr(idx) if gp is not None else None, B, D, H, W, Co, gw.data_ptr(), gb.data_ptr(),
                                                gga.data_ptr(), gbe.data_ptr(), ws.data_ptr(), nbytes, _stream()), "conv1_relu_bn_bwd")
        has_b, has_g, has_be = ctx.has
        return (None, gw.view(ctx.wshape), gb if has_b else None, gga if has_g else None, gbe if has_be else None, None, None, None, None, None,
                None, None, None)


# bf16 NDHWC copy of a volume produced by the stage that wrote the fp32 one (weak reference to the fp32 tensor OBJECT ->
# packed copy): pack_volume() returns it instead of re-reading the fp32 volume.
_PREPACKED = {"src": None, "packed": None}


def conv1_relu_bn_channels_last(x, conv, bn, with_pool=False):
    """bn(relu(conv(x))) for the (Conv3d(1,16,3,padding=1), BatchNorm3d(16)) pair; mirrors nn.BatchNorm3d's side
    effects (running statistics and num_batches_tracked in training mode).  Returns y, or (y, maxpool2(y))."""
    import weakref
    track = bn.track_running_stats and bn.running_mean is not None
    update = bn.training and track
    if update and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    use_batch = bn.training or not track
    rm, rv = (bn.running_mean, bn.running_var) if track else (None, None)
    out = _Conv1ReLUBN.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, rm, rv, use_batch, update,
                             bn.momentum if bn.momentum is not None else 0.0, bn.eps, bool(with_pool), torch.is_grad_enabled())
    y, packed = out[0], out[-1]
    _PREPACKED["src"], _PREPACKED["packed"] = weakref.ref(y), packed
    return (y, out[1]) if with_pool else y


class _ConvBf16Backward(torch.autograd.Function):
    """Conv3d whose forward is cuDNN's fp32/TF32 kernel and whose backward-data / backward-weight run on bf16
    copies of the saved activation and of the incoming gradient (fp32 accumulation).  cuDNN's TF32 wgrad for the
    encoder's 64^3 layers is 0.6-0.9 ms per layer at batch 4; the bf16 kernels take 0.2-0.3 ms (tools/conv_probe.py),
    at a relative gradient error of 3e-3 (bf16 rounding of the operands) instead of 3e-4 (TF32)."""

    @staticmethod
    @_entry
    def forward(ctx, x, weight, bias, stride, padding, dilation, groups):
        y = torch.nn.functional.conv3d(x, weight, bias, stride, padding, dilation, groups)
        cl = torch.channels_last_3d
        ctx.save_for_backward(_to_bf16_channels_last(x.detach()), weight)
        ctx.conf = (stride, padding, dilation, groups, bias is not None)
        return y

    @staticmethod
    @_entry
    def backward(ctx, gy):
        x_bf, weight = ctx.saved_tensors
        stride, padding, dilation, groups, has_bias = ctx.conf
        cl = torch.channels_last_3d
        g_bf = gy.to(torch.bfloat16).contiguous(memory_format=cl)
        w_bf = weight.detach().to(torch.bfloat16).contiguous(memory_format=cl)
        gx, gw, _ = torch.ops.aten.convolution_backward(g_bf, x_bf, w_bf, None, list(stride), list(padding), list(dilation), False, [0, 0, 0],
                                                        groups, [ctx.needs_input_grad[0], ctx.needs_input_grad[1], False])
        gb = gy.sum((0, 2, 3, 4)) if has_bias and ctx.needs_input_grad[2] else None
        return (_widen_bf16(gx) if gx is not None else None, gw.to(weight.dtype) if gw is not None else None, gb, None, None, None, None)


class _ConvBiasReLU(torch.autograd.Function):
    """relu(Conv3d(x)) for channels-last fp32 activations: cuDNN convolution without bias, then ONE in-place
    bias+ReLU pass (csrc/encoder_glue.cu); the backward masks the gradient, sums the bias gradient and emits the
    bf16 (large layers) or fp32 operand of the convolution backward kernels in ONE pass."""

    @staticmethod
    @_entry
    def forward(ctx, x, weight, bias, stride, padding, dilation, groups, bf16_backward):
        cl = torch.channels_last_3d
        if x.dtype != torch.float32 or weight.dtype != torch.float32:      # half-precision activations (AMP callers)
            x, weight = x.float(), weight.float()
        y = torch.nn.functional.conv3d(x, weight, None, stride, padding, dilation, groups)
        if y.dtype != torch.float32:
            raise RuntimeError(f"svr_b200: conv3d returned {y.dtype}; the bias/ReLU kernel needs fp32")
        if not y.is_contiguous(memory_format=cl):
            y = y.contiguous(memory_format=cl)
        Co = y.shape[1]
        rows = y.numel() // Co
        _abi.check(_lib().svr_bias_relu_cl(y.data_ptr(), _ptr(bias.detach() if bias is not None else None), rows, Co, _stream()), "bias_relu_cl")
        xs = x.detach()
        ctx.save_for_backward(_to_bf16_channels_last(xs) if bf16_backward else xs, weight, y)
        ctx.conf = (stride, padding, dilation, groups, bias is not None, bool(bf16_backward))
        return y

    @staticmethod
    @_entry
    def backward(ctx, gy):
        xs, weight, y = ctx.saved_tensors
        stride, padding, dilation, groups, has_bias, bf16_backward = ctx.conf
        cl = torch.channels_last_3d
        if gy.dtype != torch.float32:
            gy = gy.float()
        if not gy.is_contiguous(memory_format=cl):   # the kernel reads (rows, C): NDHWC storage
            gy = gy.contiguous(memory_format=cl)
        Co = y.shape[1]
        rows = y.numel() // Co
        dev = y.device
        g = torch.empty_like(y, dtype=torch.bfloat16 if bf16_backward else torch.float32, memory_format=cl)
        gb = torch.empty((Co,), device=dev, dtype=torch.float32) if has_bias else None
        nbytes = _lib().svr_relu_bwd_cl_workspace_bytes(Co)
        ws = torch.empty((nbytes,), device=dev, dtype=torch.uint8)
        _abi.check(_lib().svr_relu_bwd_cl(gy.data_ptr(), y.data_ptr(), rows, Co, None if bf16_backward else g.data_ptr(),
                                          g.data_ptr() if bf16_backward else None, _ptr(gb), ws.data_ptr(), nbytes, _stream()), "relu_bwd_cl")
        w = weight.detach()
        if bf16_backward:
            w = w.to(torch.bfloat16).contiguous(memory_format=cl)
        gx, gw, _ = torch.ops.aten.convolution_backward(g, xs, w, None, list(stride), list(padding), list(dilation), False, [0, 0, 0], groups,
                                                        [ctx.needs_input_grad[0], ctx.needs_input_grad[1], False])
        if gx is not None and gx.dtype != torch.float32:
            gx = _widen_bf16(gx)
        if gw is not None and gw.dtype != weight.dtype:
            gw = gw.to(weight.dtype)
        return gx, gw, gb, None, None, None, None, None


def conv3d_bias_relu(x, conv, bf16_backward):
    return _ConvBiasReLU.apply(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups, bf16_backward)


@_entry
def _to_bf16_channels_last(x):
    """bf16 copy of a channels-last fp32 activation through the vectorised convert kernel (torch's .to() falls back to
    a strided element-wise copy on permuted views: 2 TB/s)."""
    if x.dtype == torch.float32 and x.is_cuda and x.dim() == 5 and x.is_contiguous(memory_format=torch.channels_last_3d):
        return pack_volume(x).permute(0, 4, 1, 2, 3)
    return x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)


@_entry
def _widen_bf16(t):
    """fp32 copy of a dense bf16 tensor with the same strides (vectorised; torch's .float() on a channels-last view is a
    strided element-wise copy)."""
    if (t.dtype == _BF16 and t.is_cuda and t.numel() % 8 == 0 and t.data_ptr() % 16 == 0
            and (t.is_contiguous() or (t.dim() == 5 and t.is_contiguous(memory_format=torch.channels_last_3d)))):
        out = torch.empty_like(t, dtype=torch.float32)          # preserve_format: same dense strides
        if out.stride() == t.stride():
            _abi.check(_lib().svr_widen_bf16(t.data_ptr(), t.numel(), out.data_ptr(), _stream()), "widen_bf16")
            return out
    return t.float()


def conv3d_bf16_backward(x, conv):
    return _ConvBf16Backward.apply(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)


def maxpool2_channels_last(x: torch.Tensor) -> torch.Tensor:
    return _MaxPool2CL.apply(x)


# =================================================================================================
# IF-Net sampling + decoder
# =================================================================================================
class PyramidSpec:
    """Static description of the sampled volumes of one forward call."""

    def __init__(self, channels: Sequence[int], dims: Sequence[Sequence[int]], align_corners: bool, displacement: float):
        self.channels = tuple(int(c) for c in channels)
        self.dims = tuple(tuple(int(v) for v in d) for d in dims)
        self.align_corners = bool(align_corners)
        self.displacement = float(displacement)
        self.c = _abi.make_pyramid(self.channels, self.dims, align_corners, displacement)
        self.kp = _lib().svr_feature_kp(C.byref(self.c))
        if self.kp <= 0:
            raise RuntimeError("svr_b200: invalid pyramid: " + _lib().svr_last_error().decode())
        self.k = 7 * sum(self.channels)


@_entry
def pack_volume(v: torch.Tensor) -> torch.Tensor:
    """fp32 (B,C,D,H,W), any strides -> bf16 NDHWC contiguous (B,D,H,W,C)."""
    if not v.is_cuda:
        raise RuntimeError("svr_b200: feature volumes must be CUDA tensors; there is no CPU path")
    src = _PREPACKED["src"]
    if src is not None and src() is v:            # the producing stage already wrote the bf16 copy
        return _PREPACKED["packed"]
    if v.dtype != torch.float32:
        v = v.float()
    B, Cc, D, H, W = v.shape
    out = torch.empty((B, D, H, W, Cc), device=v.device, dtype=_BF16)
    s = v.stride()
    _abi.check(_lib().svr_pack_volume(v.data_ptr(), B, Cc, D, H, W, s[0], s[1], s[2], s[3], s[4], out.data_ptr(), _stream()),
               "pack_volume")
    return out


USE_HALO = True      # halo'd copies of the wide levels for the fused kernel's bounds-check-free path


@_entry
def pack_volume_halo(v: torch.Tensor) -> torch.Tensor:
    """fp32 (B,C,D,H,W), any strides -> bf16 (B, D+2, H+2, W+2, C) with a one-voxel zero halo."""
    if not v.is_cuda:
        raise RuntimeError("svr_b200: feature volumes must be CUDA tensors; there is no CPU path")
    if v.dtype != torch.float32:
        v = v.float()
    B, Cc, D, H, W = v.shape
    out = torch.empty((B, D + 2, H + 2, W + 2, Cc), device=v.device, dtype=_BF16)
    s = v.stride()
    _abi.check(_lib().svr_pack_volume_halo(v.data_ptr(), B, Cc, D, H, W, s[0], s[1], s[2], s[3], s[4], out.data_ptr(), _stream()),
               "pack_volume_halo")
    return out


def halo_volumes(vols):
    """Halo'd copies of the trailing run of levels with C % 64 == 0 (None for the others)."""
    out = [None] * len(vols)
    if USE_HALO:
        for i in range(len(vols) - 1, -1, -1):
            if vols[i].shape[1] % 64 != 0:
                break
            out[i] = pack_volume_halo(vols[i])
    return out


class PackedDecoder:
    """bf16 copies of the decoder weights in kernel layout.  Re-packed on EVERY call (three tiny
    kernels, ~10 us) unless ``frozen`` is set: fused optimisers update parameters without bumping
    the tensors' version counters, so no cheap staleness test is reliable.  ``frozen`` is used by
    the chunked dense evaluation, where the weights cannot change between chunks."""

    def __init__(self):
        self.key = None
        self.t = {}
        self.frozen = False

    def get(self, pyr: PyramidSpec, w0, w1, w2, backward=True):
        """``backward=False`` (inference): the swizzled images of the TRANSPOSED weights, which only the fused decoder
        backward streams, are not built (three launches of the nine)."""
        key = (pyr.channels, pyr.align_corners, w0.data_ptr(), w1.data_ptr(), w2.data_ptr(), bool(backward))
        if key != self.key or not self.frozen:
            dev = w0.device
            h0, h1, h2 = w0.shape[0], w1.shape[0], w2.shape[0]
            w0f, w1f, w2f = (_dev_f32(w.detach().reshape(w.shape[0], -1), "weight") for w in (w0, w1, w2))
            if w0f.shape[1] != pyr.k:
                raise RuntimeError(f"svr_b200: fc_0 expects {w0f.shape[1]} features, pyramid provides {pyr.k}")
            t = {"w0p": torch.empty((h0, pyr.kp), device=dev, dtype=_BF16), "w0pT": torch.empty((pyr.kp, h0), device=dev, dtype=_BF16),
                 "w1": torch.empty((h1, h0), device=dev, dtype=_BF16), "w1T": torch.empty((h0, h1), device=dev, dtype=_BF16),
                 "w2": torch.empty((h2, h1), device=dev, dtype=_BF16), "w2T": torch.empty((h1, h2), device=dev, dtype=_BF16)}
            st = _stream()
            _abi.check(_lib().svr_pack_w0(w0f.data_ptr(), h0, C.byref(pyr.c), t["w0p"].data_ptr(), t["w0pT"].data_ptr(), st), "pack_w0")
            _abi.check(_lib().svr_pack_matrix(w1f.data_ptr(), h1, h0, t["w1"].data_ptr(), t["w1T"].data_ptr(), st), "pack_matrix")
            _abi.check(_lib().svr_pack_matrix(w2f.data_ptr(), h2, h1, t["w2"].data_ptr(), t["w2T"].data_ptr(), st), "pack_matrix")
            if h0 == h1 == h2 == 256:    # pre-swizzled UMMA chunk images for the fused forward kernel
                for name in ("w0p", "w1", "w2") + (("w0pT", "w1T", "w2T") if backward else ()):
                    t[name + "_img"] = swizzled_image(t[name])
            self.t, self.key = t, key
        return self.t


@_entry
def swizzled_image(w):
    """(R, K) bf16 row-major matrix -> the K-chunked 128-byte-swizzled operand image the fused kernels stream."""
    img = torch.empty((w.numel() * 2,), device=w.device, dtype=torch.uint8)
    _abi.check(_lib().svr_pack_decoder_image(w.data_ptr(), w.shape[0], w.shape[1], img.data_ptr(), _stream()), "pack_decoder_image")
    return img


def decoder_bwd_fused(dz2, h1, h0, w2T_img, w1T_img, w0pT_img, kp):
    """dz1 = (dz2 W2)*[h1>0], dz0 = (dz1 W1)*[h0>0], dfeat = dz0 W0' (bf16) in one persistent kernel (hidden size 256)."""
    M = dz2.shape[0]
    dz1, dz0 = torch.empty_like(dz2), torch.empty_like(dz2)
    dfeat = torch.empty((M, kp), device=dz2.device, dtype=_BF16)
    _abi.check(_lib().svr_decoder_bwd_fused(dz2.data_ptr(), h1.data_ptr(), h0.data_ptr(), w2T_img.data_ptr(), w1T_img.data_ptr(),
                                            w0pT_img.data_ptr(), M, kp, dz1.data_ptr(), dz0.data_ptr(), dfeat.data_ptr(), _stream()),
               "decoder_bwd_fused")
    return dz1, dz0, dfeat


def _gemm_nt(A, B, bias, M, N, K, flags, c_bf16=None, c_f32=None, ldc=0, mask=None, dot_w=None, dot_b=None, out_dot=None):
    _abi.check(_lib().svr_gemm_nt(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), _ptr(bias), M, N, K, flags, _ptr(c_bf16),
                                  _ptr(c_f32), ldc, _ptr(mask), _ptr(dot_w), _ptr(dot_b), _ptr(out_dot), _stream()), "gemm_nt")


def _gemm_tn(A, B, M, N, P, out, accumulate=False):
    nbytes = _lib().svr_gemm_tn_workspace_bytes(M, N, P)
    ws = torch.empty((nbytes,), device=A.device, dtype=torch.uint8)
    _abi.check(_lib().svr_gemm_tn(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), M, N, P, out.data_ptr(), out.stride(0),
                                  int(accumulate), ws.data_ptr(), nbytes, _stream()), "gemm_tn")


RELU, ST_BF16, ST_F32, MASK, DOT = 1, 2, 4, 8, 16
OVERLAP_WGRAD = True   # decoder weight-gradient GEMMs on a side stream, concurrent with the scatter kernels
OVERLAP_PREP = True    # point sort + weight packing on the side stream, concurrent with the encoder
_SIDE_STREAMS = {}


def _side_stream(dev) -> "torch.cuda.Stream":
    key = torch.device(dev).index
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(dev)
    return _SIDE_STREAMS[key]
USE_FUSED = True     # fused gather+decoder forward kernel when the decoder is 256/256/256
USE_FUSED_BWD = True  # fused decoder backward-data chain (dz1, dz0, dfeat) when the decoder is 256/256/256
SORT_MIN_POINTS = 2048   # spatially sort the query points of a scene when it has at least this many


@_entry
def sort_points(pts: torch.Tensor) -> torch.Tensor:
    """(B,N,3) -> int32 (B*N,) processing order (scene-major, Morton order of 16^3 cells inside a scene)."""
    B, N, _ = pts.shape
    # one buffer: the order, then the first sorted row of every (scene, cell) + the total (read by the scatter, which
    # cuts its row tiles at cell-group boundaries; see sort_cells_ptr)
    buf = torch.empty((B * N + B * _lib().svr_sort_cells_per_scene() + 1,), device=pts.device, dtype=torch.int32)
    perm = buf[:B * N]
    nbytes = _lib().svr_sort_points_workspace_bytes(B, N)
    ws = torch.empty((nbytes,), device=pts.device, dtype=torch.uint8)
    _abi.check(_lib().svr_sort_points(pts.data_ptr(), B, N, perm.data_ptr(), perm.data_ptr() + 4 * B * N, ws.data_ptr(), nbytes, _stream()),
               "sort_points")
    return perm


def sort_cells_ptr(perm, B: int):
    """Address of the cell table that ``sort_points`` stores behind the order (None for an order from elsewhere)."""
    if perm is None:
        return None
    M, cells = perm.numel(), B * _lib().svr_sort_cells_per_scene() + 1
    st = perm.untyped_storage()
    if perm.storage_offset() != 0 or st.nbytes() < 4 * (M + cells):
        return None
    return perm.data_ptr() + 4 * M


def fused_forward(pyr, W, pts, x0, packed, b0f, b1f, b2f, wof, bof, save: bool, sigmoid: bool = False, perm=None, halo=None):
    """One launch: gather -> tcgen05 decoder.  Returns (logits (M,), h (3,M,256) or None, feat (M,KP) or None)."""
    B, N, _ = pts.shape
    M = B * N
    dev = pts.device
    dw = _abi.DecoderWeights()
    dw.w0p, dw.w1, dw.w2 = W["w0p_img"].data_ptr(), W["w1_img"].data_ptr(), W["w2_img"].data_ptr()
    dw.b0, dw.b1, dw.b2 = b0f.data_ptr(), b1f.data_ptr(), b2f.data_ptr()
    dw.wout, dw.bout = wof.data_ptr(), bof.data_ptr()
    dw.h0 = dw.h1 = dw.h2 = 256
    logits = torch.empty((M,), device=dev, dtype=torch.float32)
    h = torch.empty((3, M, 256), device=dev, dtype=_BF16) if save else None
    feat = torch.empty((M, pyr.kp), device=dev, dtype=_BF16) if save else None
    tbl = _abi.ptr_table([None] + [v.data_ptr() for v in packed])
    htbl = _abi.ptr_table([None] + [_ptr(v) for v in (halo or [])])
    _abi.check(_lib().svr_query_fwd_fused(pts.data_ptr(), _ptr(perm), sort_cells_ptr(perm, B), B, N, x0.data_ptr(), tbl, htbl, C.byref(pyr.c), C.byref(dw),
                                          logits.data_ptr(), _ptr(h), _ptr(feat), int(sigmoid), _stream()), "query_fwd_fused")
    return logits, h, feat


def gather_features(points, x0, packed_vols: List[torch.Tensor], pyr: PyramidSpec) -> torch.Tensor:
    """(B,N,3), (B,1,D,H,W), packed bf16 volumes -> (B*N, KP) bf16 features in kernel order."""
    B, N, _ = points.shape
    feat = torch.empty((B * N, pyr.kp), device=points.device, dtype=_BF16)
    tbl = _abi.ptr_table([None] + [v.data_ptr() for v in packed_vols])
    _abi.check(_lib().svr_gather_fwd(points.data_ptr(), B, N, x0.data_ptr(), tbl, C.byref(pyr.c), feat.data_ptr(), _stream()),
               "gather_fwd")
    return feat


def _decoder_struct(W, b0f, b1f, b2f, wof, bof):
    dw = _abi.DecoderWeights()
    dw.w0p, dw.w1, dw.w2 = W["w0p_img"].data_ptr(), W["w1_img"].data_ptr(), W["w2_img"].data_ptr()
    dw.b0, dw.b1, dw.b2 = b0f.data_ptr(), b1f.data_ptr(), b2f.data_ptr()
    dw.wout, dw.bout = wof.data_ptr(), bof.data_ptr()
    dw.h0 = dw.h1 = dw.h2 = 256
    return dw


@_entry
def dense_eval(pyr, cache, x, vols, w0, b0, w1, b1, w2, b2, wo, bo, lattice, scenes=None, x_range=None):
    """sigmoid(decoder(sample(x, make_3d_grid lattice))) for whole scenes in ONE launch per scene
    (evaluate_network_on_grid, ifnet.py:215-229): the lattice points are generated inside the
    kernel in brick order, nothing but the (sx,sy,sz) result is written.  Returns (len(scenes), sx, sy, sz).
    ``x_range=(begin, end)`` restricts the work to a slab of the first lattice axis (point-block
    sharding across GPUs); the rest of the output is left at zero."""
    x0 = _dev_f32(x, "x")
    sx, sy, sz = (int(v) for v in lattice)
    scenes = list(range(x0.shape[0])) if scenes is None else list(scenes)
    packed = [pack_volume(v) for v in vols]
    W = cache.get(pyr, w0, w1, w2, backward=False)
    if "w0p_img" not in W:
        raise RuntimeError("svr_b200: the fused dense evaluator needs a 256/256/256 decoder")
    b0f, b1f, b2f, bof = (_dev_f32(b.detach(), "bias") for b in (b0, b1, b2, bo))
    wof = _dev_f32(wo.detach().reshape(-1), "fc_out.weight")
    dw = _decoder_struct(W, b0f, b1f, b2f, wof, bof)
    tbl = _abi.ptr_table([None] + [v.data_ptr() for v in packed])
    halo = halo_volumes(vols)
    htbl = _abi.ptr_table([None] + [_ptr(v) for v in halo])
    xb, xe = (0, sx) if x_range is None else x_range
    out = torch.zeros((len(scenes), sx, sy, sz), device=x0.device, dtype=torch.float32)
    for i, sc in enumerate(scenes):
        _abi.check(_lib().svr_dense_eval(int(sc), x0.shape[0], x0.data_ptr(), tbl, htbl, C.byref(pyr.c), C.byref(dw), sx, sy, sz, int(xb), int(xe),
                                         out[i].data_ptr(), _stream()), "dense_eval")
    return out


class QueryPrefetch:
    """Work of the query path that does not depend on the encoder -- the spatial sort of the points and the bf16 / swizzled
    copies of the decoder weights -- issued on a side stream so that it runs NEXT TO the encoder instead of after it."""

    def __init__(self, pyr, cache: "PackedDecoder", points, w0, w1, w2):
        pts = _dev_f32(points.detach(), "points")
        dev = pts.device
        self.pyr, self.pts_key = pyr, (pts.data_ptr(), tuple(pts.shape))
        self._pts = pts                        # alive until the consumer has waited on `event`
        main, side = torch.cuda.current_stream(dev), _side_stream(dev)
        grad_enabled = torch.is_grad_enabled()
        ev = main.record_event()               # the points / weights are ready in main-stream order
        with torch.cuda.device(dev), torch.cuda.stream(side), torch.autocast("cuda", enabled=False):
            side.wait_event(ev)
            self.perm = sort_points(pts) if pts.shape[1] >= SORT_MIN_POINTS else None
            self.W = cache.get(pyr, w0, w1, w2, backward=grad_enabled)
            self.event = side.record_event()
        for t in ([self.perm] if self.perm is not None else []) + list(self.W.values()):
            t.record_stream(main)

    def matches(self, pyr, pts) -> bool:
        return self.pyr is pyr and self.pts_key == (pts.data_ptr(), tuple(pts.shape))


class _Query(torch.autograd.Function):
    """IFNet.forward given the encoder's volumes (ifnet.py:38-61,156-197): stencil gather + decoder.

    inputs: points (B,N,3), x (B,1,D,H,W), vols[1..L-1] fp32 (B,C,D,H,W), then the 8 decoder
    parameters.  Returns logits (B,N) fp32."""

    @staticmethod
    @_entry
    def forward(ctx, pyr: PyramidSpec, cache: PackedDecoder, grad_mode, points, x, w0, b0, w1, b1, w2, b2, wo, bo, *vols):
        grad_mode, prefetched = grad_mode if isinstance(grad_mode, tuple) else (grad_mode, None)
        pts = _dev_f32(points, "points")
        x0 = _dev_f32(x, "x")
        B, N, _ = pts.shape
        M = B * N
        dev = pts.device
        packed = [pack_volume(v) for v in vols]
        if prefetched is not None and not prefetched.matches(pyr, pts):
            prefetched = None
        if prefetched is not None:
            torch.cuda.current_stream(dev).wait_event(prefetched.event)
            W = prefetched.W
        else:
            W = cache.get(pyr, w0, w1, w2, backward=bool(grad_mode))
        h0n, h1n, h2n = w0.shape[0], w1.shape[0], w2.shape[0]
        b0f, b1f, b2f, bof = (_dev_f32(b.detach(), "bias") for b in (b0, b1, b2, bo))
        wof = _dev_f32(wo.detach().reshape(-1), "fc_out.weight")
        # needs_input_grad is set for parameters even under torch.no_grad(); nothing is saved (and the kernels skip the
        # 1.3 GB of feature / activation stores per 200k points) unless a graph is being recorded
        needs_bwd = bool(grad_mode) and any(ctx.needs_input_grad)
        if needs_bwd and "w0p_img" in W and "w0pT_img" not in W:      # prefetched without gradients enabled
            W = cache.get(pyr, w0, w1, w2, backward=True)
        perm = None
        if USE_FUSED and "w0p_img" in W:
            if prefetched is not None:
                perm = prefetched.perm
            elif N >= SORT_MIN_POINTS:
                perm = sort_points(pts)
            logits, hs, feat = fused_forward(pyr, W, pts, x0, packed, b0f, b1f, b2f, wof, bof, save=needs_bwd, perm=perm,
                                             halo=halo_volumes(vols))
            if not needs_bwd:
                return logits.view(B, N)
            h0, h1, h2 = hs[0], hs[1], hs[2]
        else:
            feat = gather_features(pts, x0, packed, pyr)
            h0 = torch.empty((M, h0n), device=dev, dtype=_BF16)
            h1 = torch.empty((M, h1n), device=dev, dtype=_BF16)
            h2 = torch.empty((M, h2n), device=dev, dtype=_BF16)
            logits = torch.empty((M,), device=dev, dtype=torch.float32)
            _gemm_nt(feat, W["w0p"], b0f, M, h0n, pyr.kp, RELU | ST_BF16, c_bf16=h0, ldc=h0n)
            _gemm_nt(h0, W["w1"], b1f, M, h1n, h0n, RELU | ST_BF16, c_bf16=h1, ldc=h1n)
            _gemm_nt(h1, W["w2"], b2f, M, h2n, h1n, RELU | ST_BF16 | DOT, c_bf16=h2, ldc=h2n, dot_w=wof, dot_b=bof, out_dot=logits)
        ctx.pyr, ctx.cache_t = pyr, W
        ctx.vol_meta = [(v.shape, v.stride(), ctx.needs_input_grad[13 + i]) for i, v in enumerate(vols)]
        ctx.x_needs, ctx.p_needs = ctx.needs_input_grad[4], ctx.needs_input_grad[3]
        ctx.shapes = (B, N, w0.shape, w1.shape, w2.shape, wo.shape)
        ctx.has_perm = perm is not None
        ctx.save_for_backward(pts, x0, feat, h0, h1, h2, wof, perm if perm is not None else pts.new_empty(0), *packed)
        return logits.view(B, N)

    @staticmethod
    @_entry
    def backward(ctx, glogits):
        pts, x0, feat, h0, h1, h2, wof, perm, *packed = ctx.saved_tensors
        perm = perm if ctx.has_perm else None
        pyr, W = ctx.pyr, ctx.cache_t
        B, N, s0, s1, s2, so = ctx.shapes
        M = B * N
        dev = pts.device
        h0n, h1n, h2n = s0[0], s1[0], s2[0]
        dl = _dev_f32(glogits, "grad").reshape(-1)
        st = _stream()
        # fc_out + relu(fc_2) backward
        dz2 = torch.empty((M, h2n), device=dev, dtype=_BF16)
        gwo = torch.zeros((h2n,), device=dev, dtype=torch.float32)
        gbo = torch.zeros((1,), device=dev, dtype=torch.float32)
        _abi.check(_lib().svr_decoder_head_bwd(dl.data_ptr(), _ptr(perm), h2.data_ptr(), wof.data_ptr(), M, h2n, dz2.data_ptr(),
                                               gwo.data_ptr(), gbo.data_ptr(), st), "decoder_head_bwd")

        def colsum(a, n):
            out = torch.empty((n,), device=dev, dtype=torch.float32)
            _abi.check(_lib().svr_colsum_bf16(a.data_ptr(), M, n, a.stride(0), out.data_ptr(), 0, _stream()), "colsum")
            return out

        def weight_grads(which):
            """Weight / bias gradients of fc_2 (which = 2) or of fc_1 and fc_0 (which = 1) on the CURRENT stream."""
            if which == 2:
                g = torch.empty((h2n, h1n), device=dev, dtype=torch.float32)
                _gemm_tn(dz2, h1, h2n, h1n, M, g)
                return g, colsum(dz2, h2n)
            g1 = torch.empty((h1n, h0n), device=dev, dtype=torch.float32)
            _gemm_tn(dz1, h0, h1n, h0n, M, g1)
            b1_ = colsum(dz1, h1n)
            g0p = torch.empty((h0n, pyr.kp), device=dev, dtype=torch.float32)
            _gemm_tn(dz0, feat, h0n, pyr.kp, M, g0p)
            g0 = torch.empty((h0n, pyr.k), device=dev, dtype=torch.float32)
            _abi.check(_lib().svr_unpack_w0_grad(g0p.data_ptr(), h0n, C.byref(pyr.c), g0.data_ptr(), _stream()), "unpack_w0_grad")
            return g1, b1_, g0, colsum(dz0, h0n)

        need_dfeat = any(m[2] for m in ctx.vol_meta) or ctx.x_needs or ctx.p_needs
        dfeat = None
        fused_bwd = USE_FUSED_BWD and "w0pT_img" in W
        # The weight-gradient GEMMs (tensor / HBM bound, small grids) are independent of the scatter (atomics / latency
        # bound): with OVERLAP_WGRAD they run on a side stream next to it instead of in front of it.
        overlap = OVERLAP_WGRAD and fused_bwd and need_dfeat
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev) if overlap else None
        if overlap:
            ev_head = main.record_event()
        else:
            gw2, gb2 = weight_grads(2)
        if fused_bwd:   # dz1, dz0 and dfeat in one persistent kernel
            dz1, dz0, dfeat = decoder_bwd_fused(dz2, h1, h0, W["w2T_img"], W["w1T_img"], W["w0pT_img"], pyr.kp)
        else:
            dz1 = torch.empty((M, h1n), device=dev, dtype=_BF16)
            dz0 = torch.empty((M, h0n), device=dev, dtype=_BF16)
            _gemm_nt(dz2, W["w2T"], None, M, h1n, h2n, ST_BF16 | MASK, c_bf16=dz1, ldc=h1n, mask=h1)
            _gemm_nt(dz1, W["w1T"], None, M, h0n, h1n, ST_BF16 | MASK, c_bf16=dz0, ldc=h0n, mask=h0)
        if overlap:
            ev_chain = main.record_event()
            with torch.cuda.stream(side):
                side.wait_event(ev_head)
                gw2, gb2 = weight_grads(2)
                side.wait_event(ev_chain)
                gw1, gb1, gw0, gb0 = weight_grads(1)
                ev_side = side.record_event()
            for t in (gw2, gb2, gw1, gb1, gw0, gb0):      # allocated on the side stream, consumed on the main stream
                t.record_stream(main)
        else:
            gw1, gb1, gw0, gb0 = weight_grads(1)
        # d features, then scatter-add into the volumes
        any_vol = any(m[2] for m in ctx.vol_meta)
        gvols_out: List[Optional[torch.Tensor]] = [None] * len(packed)
        gx = gp = None
        if any_vol or ctx.x_needs or ctx.p_needs:
            if dfeat is None:
                dfeat = torch.empty((M, pyr.kp), device=dev, dtype=_BF16)
                _gemm_nt(dz0, W["w0pT"], None, M, pyr.kp, h0n, ST_BF16, c_bf16=dfeat, ldc=pyr.kp)
            gbufs = []
            for i, (shape, _, needs) in enumerate(ctx.vol_meta):
                Bv, Cv, Dv, Hv, Wv = shape
                gbufs.append(torch.zeros((Bv, Dv, Hv, Wv, Cv), device=dev, dtype=torch.float32) if needs else None)
            if ctx.x_needs:
                gx = torch.zeros_like(x0)
            if ctx.p_needs:
                gp = torch.zeros_like(pts)
            vt = _abi.ptr_table([None] + [v.data_ptr() for v in packed])
            # With spatially sorted rows the library scatters the coarse levels (<= 40^3 voxels) on the tensor cores and
            # the rest with direct vector reductions: two kernels.  They are issued as two calls (coarse-level table /
            # everything else) so that each kernel is bracketed on its own by the profiler; the work is identical.
            coarse = [perm is not None and g is not None and g.shape[1] * g.shape[2] * g.shape[3] <= 40 * 40 * 40 for g in gbufs]
            fine_t = _abi.ptr_table([None] + [None if c else _ptr(g) for g, c in zip(gbufs, coarse)])
            if gx is not None or gp is not None or any(g is not None and not c for g, c in zip(gbufs, coarse)):
                _abi.PROFILE.label = "svr_gather_bwd[direct]"
                _abi.check(_lib().svr_gather_bwd(pts.data_ptr(), _ptr(perm), None, B, N, x0.data_ptr(), vt, C.byref(pyr.c), dfeat.data_ptr(),
                                                 _ptr(gx), fine_t, _ptr(gp), st), "gather_bwd")
            if any(coarse):
                coarse_t = _abi.ptr_table([None] + [_ptr(g) if c else None for g, c in zip(gbufs, coarse)])
                _abi.PROFILE.label = "svr_gather_bwd[tensor-core]"
                _abi.check(_lib().svr_gather_bwd(pts.data_ptr(), _ptr(perm), sort_cells_ptr(perm, B), B, N, x0.data_ptr(), vt, C.byref(pyr.c),
                                                 dfeat.data_ptr(), None, coarse_t, None, st), "gather_bwd")
            gvols_out = [g.permute(0, 4, 1, 2, 3) if g is not None else None for g in gbufs]
        if overlap:
            main.wait_event(ev_side)
        return (None, None, None, gp, gx, gw0.view(s0), gb0, gw1.view(s1), gb1, gw2.view(s2), gb2, gwo.view(so), gbo, *gvols_out)


# =================================================================================================
# fp32-accurate tier (csrc/precise.cu): fp32 tensors in HBM, hi/lo bf16 split products on the tensor cores
# =================================================================================================
ACCUM = 32


def _split(x: torch.Tensor):
    """fp32 (contiguous) -> (hi, lo) bf16 with x ~= hi + lo to 2^-17 relative."""
    x = x.contiguous()
    if x.numel() % 8:
        raise RuntimeError("svr_b200: split needs a multiple of 8 elements")
    hi, lo = torch.empty_like(x, dtype=_BF16), torch.empty_like(x, dtype=_BF16)
    _abi.check(_lib().svr_split_bf16(x.data_ptr(), x.numel(), hi.data_ptr(), lo.data_ptr(), _stream()), "split_bf16")
    return hi, lo


def _mm_nt3(A, Bm, bias, M, N, K, out, flags=0, **kw):
    """out (M,N) fp32 = epi(A . B^T + bias) with A = (hi, lo), B = (hi, lo): three tensor-core passes, fp32 running sum."""
    _gemm_nt(A[0], Bm[0], None, M, N, K, ST_F32, c_f32=out, ldc=N)
    _gemm_nt(A[1], Bm[0], None, M, N, K, ST_F32 | ACCUM, c_f32=out, ldc=N)
    _gemm_nt(A[0], Bm[1], bias, M, N, K, ST_F32 | ACCUM | flags, c_f32=out, ldc=N, **kw)


def _mm_tn3(A, Bm, M, N, P, out):
    """out (M,N) fp32 = A^T . B with A (P,M) = (hi, lo), B (P,N) = (hi, lo) (contraction over the rows)."""
    _gemm_tn(A[0], Bm[0], M, N, P, out, accumulate=False)
    _gemm_tn(A[1], Bm[0], M, N, P, out, accumulate=True)
    _gemm_tn(A[0], Bm[1], M, N, P, out, accumulate=True)


def _colsum_f32(a: torch.Tensor) -> torch.Tensor:
    out = torch.empty((a.shape[1],), device=a.device, dtype=torch.float32)
    _abi.check(_lib().svr_colsum_f32(a.data_ptr(), a.shape[0], a.shape[1], a.stride(0), out.data_ptr(), _stream()), "colsum_f32")
    return out


def _ndhwc_f32(v: torch.Tensor) -> torch.Tensor:
    """(B,C,D,H,W) -> contiguous fp32 (B,D,H,W,C); a view when the encoder ran channels-last."""
    if not v.is_cuda:
        raise RuntimeError("svr_b200: feature volumes must be CUDA tensors; there is no CPU path")
    return _dev_f32(v.permute(0, 2, 3, 4, 1), "volume")


class _Query32(torch.autograd.Function):
    """IFNet.forward given the encoder's volumes, fp32-accurate tier (ifnet.py:38-61,156-197 in fp32)."""

    @staticmethod
    @_entry
    def forward(ctx, pyr: PyramidSpec, points, x, w0, b0, w1, b1, w2, b2, wo, bo, *vols):
        pts = _dev_f32(points, "points")
        x0 = _dev_f32(x, "x")
        B, N, _ = pts.shape
        M, dev, KP = B * N, pts.device, pyr.kp
        vf = [_ndhwc_f32(v.detach()) for v in vols]
        h0n, h1n, h2n = w0.shape[0], w1.shape[0], w2.shape[0]
        w0f = _dev_f32(w0.detach().reshape(h0n, -1), "weight")
        if w0f.shape[1] != pyr.k:
            raise RuntimeError(f"svr_b200: fc_0 expects {w0f.shape[1]} features, pyramid provides {pyr.k}")
        w0p = torch.empty((h0n, KP), device=dev, dtype=torch.float32)
        _abi.check(_lib().svr_pack_w0_f32(w0f.data_ptr(), h0n, C.byref(pyr.c), w0p.data_ptr(), _stream()), "pack_w0_f32")
        w1f = _dev_f32(w1.detach().reshape(h1n, -1), "weight")
        w2f = _dev_f32(w2.detach().reshape(h2n, -1), "weight")
        b0f, b1f, b2f, bof = (_dev_f32(b.detach(), "bias") for b in (b0, b1, b2, bo))
        wof = _dev_f32(wo.detach().reshape(-1), "fc_out.weight")
        feat = torch.empty((M, KP), device=dev, dtype=torch.float32)
        vt = _abi.ptr_table([None] + [v.data_ptr() for v in vf])
        _abi.check(_lib().svr_gather_fwd_f32(pts.data_ptr(), B, N, x0.data_ptr(), vt, C.byref(pyr.c), feat.data_ptr(), _stream()), "gather_fwd_f32")
        F2 = _split(feat)
        del feat
        h0 = torch.empty((M, h0n), device=dev, dtype=torch.float32)
        _mm_nt3(F2, _split(w0p), b0f, M, h0n, KP, h0, RELU)
        H0 = _split(h0)
        h1 = torch.empty((M, h1n), device=dev, dtype=torch.float32)
        _mm_nt3(H0, _split(w1f), b1f, M, h1n, h0n, h1, RELU)
        H1 = _split(h1)
        h2 = torch.empty((M, h2n), device=dev, dtype=torch.float32)
        logits = torch.empty((M,), device=dev, dtype=torch.float32)
        _mm_nt3(H1, _split(w2f), b2f, M, h2n, h1n, h2, RELU | DOT, dot_w=wof, dot_b=bof, out_dot=logits)
        if any(ctx.needs_input_grad):
            ctx.pyr = pyr
            ctx.vol_meta = [(v.shape, ctx.needs_input_grad[11 + i]) for i, v in enumerate(vols)]
            ctx.x_needs, ctx.p_needs = ctx.needs_input_grad[2], ctx.needs_input_grad[1]
            ctx.shapes = (B, N, w0.shape, w1.shape, w2.shape, wo.shape)
            ctx.save_for_backward(pts, x0, F2[0], F2[1], H0[0], H0[1], H1[0], H1[1], h2, w0p, w1f, w2f, wof, *vf)
        return logits.view(B, N)

    @staticmethod
    @_entry
    def backward(ctx, glogits):
        pts, x0, fhi, flo, h0hi, h0lo, h1hi, h1lo, h2, w0p, w1f, w2f, wof, *vf = ctx.saved_tensors
        pyr = ctx.pyr
        B, N, s0, s1, s2, so = ctx.shapes
        M, dev, KP = B * N, pts.device, pyr.kp
        h0n, h1n, h2n = s0[0], s1[0], s2[0]
        st = _stream()
        dl = _dev_f32(glogits, "grad").reshape(-1)
        dz2 = torch.empty((M, h2n), device=dev, dtype=torch.float32)
        gwo = torch.empty((h2n,), device=dev, dtype=torch.float32)
        gbo = torch.empty((1,), device=dev, dtype=torch.float32)
        _abi.check(_lib().svr_decoder_head_bwd_f32(dl.data_ptr(), h2.data_ptr(), wof.data_ptr(), M, h2n, dz2.data_ptr(), gwo.data_ptr(),
                                                   gbo.data_ptr(), st), "decoder_head_bwd_f32")
        del h2
        DZ2 = _split(dz2)
        gw2 = torch.empty((h2n, h1n), device=dev, dtype=torch.float32)
        _mm_tn3(DZ2, (h1hi, h1lo), h2n, h1n, M, gw2)
        gb2 = _colsum_f32(dz2)
        dz1 = torch.empty((M, h1n), device=dev, dtype=torch.float32)
        _mm_nt3(DZ2, _split(w2f.t().contiguous()), None, M, h1n, h2n, dz1, MASK, mask=h1hi)     # (dz2 . W2) * [h1 > 0]
        del dz2, DZ2
        DZ1 = _split(dz1)
        gw1 = torch.empty((h1n, h0n), device=dev, dtype=torch.float32)
        _mm_tn3(DZ1, (h0hi, h0lo), h1n, h0n, M, gw1)
        gb1 = _colsum_f32(dz1)
        dz0 = torch.empty((M, h0n), device=dev, dtype=torch.float32)
        _mm_nt3(DZ1, _split(w1f.t().contiguous()), None, M, h0n, h1n, dz0, MASK, mask=h0hi)
        del dz1, DZ1
        DZ0 = _split(dz0)
        gw0p = torch.empty((h0n, KP), device=dev, dtype=torch.float32)
        _mm_tn3(DZ0, (fhi, flo), h0n, KP, M, gw0p)
        gw0 = torch.empty((h0n, pyr.k), device=dev, dtype=torch.float32)
        _abi.check(_lib().svr_unpack_w0_grad(gw0p.data_ptr(), h0n, C.byref(pyr.c), gw0.data_ptr(), st), "unpack_w0_grad")
        gb0 = _colsum_f32(dz0)
        gvols_out: List[Optional[torch.Tensor]] = [None] * len(vf)
        gx = gp = None
        if any(m[1] for m in ctx.vol_meta) or ctx.x_needs or ctx.p_needs:
            dfeat = torch.empty((M, KP), device=dev, dtype=torch.float32)
            _mm_nt3(DZ0, _split(w0p.t().contiguous()), None, M, KP, h0n, dfeat)
            gbufs = [torch.zeros((s[0], s[2], s[3], s[4], s[1]), device=dev, dtype=torch.float32) if needs else None
                     for (s, needs) in ctx.vol_meta]
            gx = torch.zeros_like(x0) if ctx.x_needs else None
            gp = torch.zeros_like(pts) if ctx.p_needs else None
            vt = _abi.ptr_table([None] + [v.data_ptr() for v in vf])
            gt = _abi.ptr_table([None] + [_ptr(g) for g in gbufs])
            _abi.check(_lib().svr_gather_bwd_f32(pts.data_ptr(), B, N, x0.data_ptr(), vt, C.byref(pyr.c), dfeat.data_ptr(), _ptr(gx), gt,
                                                 _ptr(gp), st), "gather_bwd_f32")
            gvols_out = [g.permute(0, 4, 1, 2, 3) if g is not None else None for g in gbufs]
        return (None, gp, gx, gw0.view(s0), gb0, gw1.view(s1), gb1, gw2.view(s2), gb2, gwo.view(so), gbo, *gvols_out)


def query(pyr, cache, points, x, w0, b0, w1, b1, w2, b2, wo, bo, vols, precision=16, prefetched=None):
    """``precision`` 16: bf16 operands on the tensor cores (logits within 1e-2 of the fp32 reference); 32: the fp32-accurate
    tier (1e-3)."""
    if points.shape[0] * points.shape[1] == 0:          # empty query set: nothing to launch (ifnet.py:38-61 returns (B, 0))
        if not points.is_cuda:
            raise RuntimeError("svr_b200: `points` must be a CUDA tensor; there is no CPU path")
        return points.new_zeros((points.shape[0], points.shape[1]), dtype=torch.float32)
    if int(precision) == 32:
        return _Query32.apply(pyr, points, x, w0, b0, w1, b1, w2, b2, wo, bo, *vols)
    return _Query.apply(pyr, cache, (torch.is_grad_enabled(), prefetched), points, x, w0, b0, w1, b1, w2, b2, wo, bo, *vols)


class _Gather(torch.autograd.Function):
    """IFNetFeatureExtractor*.forward's sampling part as a standalone differentiable op returning
    the kernel-order bf16 feature rows (B*N, KP)."""

    @staticmethod
    @_entry
    def forward(ctx, pyr: PyramidSpec, points, x, *vols):
        pts = _dev_f32(points, "points")
        x0 = _dev_f32(x, "x")
        packed = [pack_volume(v) for v in vols]
        feat = gather_features(pts, x0, packed, pyr)
        ctx.pyr = pyr
        ctx.vol_meta = [(v.shape, ctx.needs_input_grad[3 + i]) for i, v in enumerate(vols)]
        ctx.x_needs, ctx.p_needs = ctx.needs_input_grad[2], ctx.needs_input_grad[1]
        ctx.save_for_backward(pts, x0, *packed)
        return feat

    @staticmethod
    @_entry
    def backward(ctx, gfeat):
        pts, x0, *packed = ctx.saved_tensors
        pyr = ctx.pyr
        B, N, _ = pts.shape
        dev = pts.device
        dfeat = gfeat.to(_BF16).contiguous()
        gbufs = [torch.zeros((s[0], s[2], s[3], s[4], s[1]), device=dev, dtype=torch.float32) if needs else None
                 for (s, needs) in ctx.vol_meta]
        gx = torch.zeros_like(x0) if ctx.x_needs else None
        gp = torch.zeros_like(pts) if ctx.p_needs else None
        vt = _abi.ptr_table([None] + [v.data_ptr() for v in packed])
        gt = _abi.ptr_table([None] + [_ptr(g) for g in gbufs])
        _abi.check(_lib().svr_gather_bwd(pts.data_ptr(), None, None, B, N, x0.data_ptr(), vt, C.byref(pyr.c), dfeat.data_ptr(), _ptr(gx), gt,
                                         _ptr(gp), _stream()), "gather_bwd")
        return (None, gp, gx, *[g.permute(0, 4, 1, 2, 3) if g is not None else None for g in gbufs])


def gather(pyr, points, x, vols):
    if points.shape[0] * points.shape[1] == 0:
        if not points.is_cuda:
            raise RuntimeError("svr_b200: `points` must be a CUDA tensor; there is no CPU path")
        return torch.zeros((0, pyr.kp), device=points.device, dtype=_BF16)
    return _Gather.apply(pyr, points, x, *vols)


def feature_index_map(pyr: PyramidSpec, device) -> torch.Tensor:
    """index tensor `idx` (K,) with reference feature k = c*7+d  ==  kernel-order column idx[k]."""
    ks = []
    ubase = 1
    # level 0: channel 0, stencil d -> column d
    cols = {}
    for d in range(7):
        cols[(0, d)] = d
    coff = 1
    for l in range(1, len(pyr.channels)):
        upd = pyr.channels[l] // 8
        if pyr.channels[l] % 64 == 0:
            ubase = (ubase + 7) // 8 * 8          # same chunk alignment as csrc/sampling.cuh make_pyr
        for d in range(7):
            for cc in range(pyr.channels[l]):
                cols[(coff + cc, d)] = (ubase + d * upd + cc // 8) * 8 + cc % 8
        ubase += 7 * upd
        coff += pyr.channels[l]
    for c in range(coff):
        for d in range(7):
            ks.append(cols[(c, d)])
    return torch.tensor(ks, device=device, dtype=torch.long)
