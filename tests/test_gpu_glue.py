"""GPU: the element-wise encoder glue entry points called directly through the C ABI (they are otherwise only reached
through the conv+ReLU autograd function): exact against torch on the same fp32 data."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from svr_b200 import _abi
    return _abi, _abi.load()


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("rows,C", [(1000, 32), (4097, 64), (77, 128), (3, 4), (0, 32)])
def test_bias_relu_cl(rows, C):
    _abi, lib = _lib()
    g = torch.Generator().manual_seed(rows + C)
    y = torch.randn((rows, C), generator=g).cuda()
    b = torch.randn((C,), generator=g).cuda()
    ref = torch.relu(y + b)
    _abi.check(lib.svr_bias_relu_cl(y.data_ptr(), b.data_ptr(), rows, C, _stream()), "bias_relu_cl")
    assert torch.equal(y, ref)
    y2 = torch.randn((rows, C), generator=g).cuda()
    ref2 = torch.relu(y2)
    _abi.check(lib.svr_bias_relu_cl(y2.data_ptr(), None, rows, C, _stream()), "bias_relu_cl")     # no bias
    assert torch.equal(y2, ref2)


def test_bias_relu_cl_rejects_bad_channel_count():
    _abi, lib = _lib()
    y = torch.zeros((8, 24), device="cuda")          # 256 % (24 / 4) != 0
    assert lib.svr_bias_relu_cl(y.data_ptr(), None, 8, 24, _stream()) < 0
    assert b"channel count" in lib.svr_last_error()


@pytest.mark.parametrize("rows,C", [(1000, 32), (5000, 64), (129, 128), (0, 32)])
@pytest.mark.parametrize("bf16_out", [False, True])
def test_relu_bwd_cl(rows, C, bf16_out):
    _abi, lib = _lib()
    g = torch.Generator().manual_seed(rows * 3 + C)
    gy = torch.randn((rows, C), generator=g).cuda()
    y = torch.relu(torch.randn((rows, C), generator=g)).cuda()
    ref = gy * (y > 0)
    out = torch.empty((rows, C), device="cuda", dtype=torch.bfloat16 if bf16_out else torch.float32)
    gb = torch.full((C,), float("nan"), device="cuda")
    nbytes = lib.svr_relu_bwd_cl_workspace_bytes(C)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device="cuda")
    _abi.check(lib.svr_relu_bwd_cl(gy.data_ptr(), y.data_ptr(), rows, C, None if bf16_out else out.data_ptr(),
                                   out.data_ptr() if bf16_out else None, gb.data_ptr(), ws.data_ptr(), nbytes, _stream()), "relu_bwd_cl")
    assert torch.equal(out, ref.to(out.dtype))
    refb = ref.double().sum(0)
    assert float((gb.double() - refb).abs().max()) <= 1e-4 * max(1.0, float(refb.abs().max()))


@pytest.mark.parametrize("n", [8, 4096, 1000000, 0])
def test_widen_bf16(n):
    _abi, lib = _lib()
    src = torch.randn((n,), generator=torch.Generator().manual_seed(n)).cuda().bfloat16()
    dst = torch.empty((n,), device="cuda", dtype=torch.float32)
    _abi.check(lib.svr_widen_bf16(src.data_ptr(), n, dst.data_ptr(), _stream()), "widen_bf16")
    assert torch.equal(dst, src.float())
    assert lib.svr_widen_bf16(src.data_ptr(), 12, dst.data_ptr(), _stream()) < 0     # not a multiple of 8


def test_host_prefetcher_roundtrip():
    import svr_b200
    pf = svr_b200.HostPrefetcher("cuda:0")
    a = torch.arange(1 << 20, dtype=torch.float32).pin_memory()
    b = torch.ones((17, 3)).pin_memory()
    h = pf.issue((a, b))
    da, db = pf.wait(h)
    torch.cuda.synchronize()
    assert torch.equal(da.cpu(), a) and torch.equal(db.cpu(), b)
    with pytest.raises(RuntimeError):
        pf.issue((torch.zeros(4),))          # not pinned
