"""GPU: the tcgen05 GEMMs (Conv1d k=1 layers and their backward) against torch fp32 matmul on the
same bf16-rounded operands.  Tolerance: fp32 accumulation of bf16 products, 2e-3 relative to the
largest output (accumulation order only)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    import svr_b200
    return svr_b200.ops


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 256, 256), (1000, 2624, 256), (517, 256, 2624), (77, 512, 2304)])
def test_gemm_nt(M, N, K):
    ops = _ops()
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn((M, K), generator=g).cuda().bfloat16()
    B = (torch.randn((N, K), generator=g) * 0.1).cuda().bfloat16()
    bias = torch.randn((N,), generator=g).cuda()
    out = torch.empty((M, N), device="cuda", dtype=torch.float32)
    ops._gemm_nt(A, B, bias, M, N, K, ops.ST_F32, c_f32=out, ldc=N)
    ref = A.float() @ B.float().t() + bias
    err = (out - ref).abs().max() / ref.abs().max()
    assert float(err) < 2e-3, float(err)
    # relu + bf16 store + mask + row-dot
    outb = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    ops._gemm_nt(A, B, bias, M, N, K, ops.RELU | ops.ST_BF16, c_bf16=outb, ldc=N)
    refr = torch.relu(ref)
    assert float((outb.float() - refr).abs().max() / refr.abs().max()) < 1e-2
    mask = (torch.randn((M, N), generator=g) > 0).cuda().bfloat16()
    outm = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    ops._gemm_nt(A, B, None, M, N, K, ops.ST_BF16 | ops.MASK, c_bf16=outm, ldc=N, mask=mask)
    refm = (A.float() @ B.float().t()) * mask.float()
    assert float((outm.float() - refm).abs().max() / refm.abs().max()) < 1e-2
    if N <= 256:
        w = torch.randn((N,), generator=g).cuda()
        b = torch.randn((1,), generator=g).cuda()
        dot = torch.empty((M,), device="cuda")
        ops._gemm_nt(A, B, bias, M, N, K, ops.RELU | ops.DOT, dot_w=w, dot_b=b, out_dot=dot)
        refd = refr @ w + b
        assert float((dot - refd).abs().max() / refd.abs().max()) < 2e-3


@pytest.mark.parametrize("P,M,N", [(64, 128, 256), (1000, 256, 256), (4097, 256, 2624), (333, 512, 2304)])
def test_gemm_tn(P, M, N):
    ops = _ops()
    g = torch.Generator().manual_seed(P + M + N)
    A = torch.randn((P, M), generator=g).cuda().bfloat16()
    B = torch.randn((P, N), generator=g).cuda().bfloat16()
    out = torch.empty((M, N), device="cuda", dtype=torch.float32)
    ops._gemm_tn(A, B, M, N, P, out)
    ref = A.float().t() @ B.float()
    err = (out - ref).abs().max() / ref.abs().max()
    assert float(err) < 2e-3, float(err)
    ops._gemm_tn(A, B, M, N, P, out, accumulate=True)
    assert float((out - 2 * ref).abs().max() / ref.abs().max()) < 4e-3


@pytest.mark.parametrize("M,KP", [(128, 64), (1000, 2624), (5000, 2304), (77, 128), (20000, 2624)])
def test_decoder_bwd_fused(M, KP):
    """Fused dz2 -> dz1 -> dz0 -> dfeat chain (TMA tiles, ragged last tile) against torch fp32 matmuls with the
    same bf16 rounding points; inputs to each stage are taken from the kernel's own previous stage so that the
    comparison is per GEMM (accumulation order only)."""
    ops = _ops()
    g = torch.Generator().manual_seed(M + KP)
    dz2 = torch.randn((M, 256), generator=g).cuda().bfloat16()
    h1 = torch.relu(torch.randn((M, 256), generator=g)).cuda().bfloat16()
    h0 = torch.relu(torch.randn((M, 256), generator=g)).cuda().bfloat16()
    w2 = (torch.randn((256, 256), generator=g) * 0.06).cuda().bfloat16()     # (out, in)
    w1 = (torch.randn((256, 256), generator=g) * 0.06).cuda().bfloat16()
    w0p = (torch.randn((256, KP), generator=g) * 0.06).cuda().bfloat16()
    imgs = [ops.swizzled_image(w.t().contiguous()) for w in (w2, w1, w0p)]
    dz1, dz0, dfeat = ops.decoder_bwd_fused(dz2, h1, h0, *imgs, KP)
    torch.cuda.synchronize()
    r1 = (dz2.float() @ w2.float()) * (h1 > 0)
    r0 = (dz1.float() @ w1.float()) * (h0 > 0)
    rf = dz0.float() @ w0p.float()
    for got, ref in ((dz1, r1), (dz0, r0), (dfeat, rf)):
        assert float((got.float() - ref).abs().max() / ref.abs().max()) < 1e-2
    # masked entries are exactly zero
    assert float(dz1.float()[h1 == 0].abs().max()) == 0.0 and float(dz0.float()[h0 == 0].abs().max()) == 0.0
