"""GPU suite, 2 ranks over NCCL (skipped on a single-GPU box): the data-parallel step's averaged gradients equal the
mean of the per-shard single-process gradients (BatchNorm uses per-shard batch statistics: DDP-without-SyncBN
semantics), bucket by bucket, with the hooks-launched early buckets in flight during backward;
and the dense evaluation sharded by first-axis slab over two ranks reassembles the single-GPU grid bit-for-bit."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs():
    g = torch.Generator().manual_seed(5)
    x = (torch.rand((4, 1, 32, 32, 32), generator=g) < 0.1).float()
    pts = torch.rand((4, 3000, 3), generator=g) - 0.5
    occ = (torch.rand((4, 3000), generator=g) < 0.5).float()
    return x, pts, occ


def _step(net, x, pts, occ):
    net.zero_grad(set_to_none=True)
    logits = net(x, pts)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, occ, reduction="none").sum(-1).mean()
    loss.backward()
    return loss


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(REPO))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import svr_b200
    from oracle import ref_torch as R
    from svr_b200 import dist as svr_dist
    svr_b200.configure(net_res=128, precision=16)
    net = svr_b200.IFNet().to(dev).train()
    net.load_state_dict(R.synthetic_state_dict(61, 128), strict=False)
    red = svr_dist.GradReducer(net, buckets=svr_dist.FINE_BUCKETS)      # four buckets: every hook-launched path is exercised
    assert len(red.buckets) == 4
    x, pts, occ = _inputs()
    b, e = svr_dist.shard_range(4, rank, world)
    for _ in range(2):       # two steps: bucket state must reset, p.grad views are replaced by fresh gradients
        _step(net, x[b:e].to(dev), pts[b:e].to(dev), occ[b:e].to(dev))
        red.allreduce()
    torch.cuda.synchronize()
    torch.save({n: p.grad.detach().cpu().clone() for n, p in net.named_parameters()}, Path(out_dir) / f"g{rank}.pt")
    # the reduced gradients (views of the bucket buffers) must be usable by a fused optimiser: same strides as the
    # (channels-last) parameters
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
    assert all(p.grad.stride() == p.stride() for p in net.parameters())
    opt.step()
    torch.cuda.synchronize()
    # dense evaluation sharded by first-axis slab (a fresh module: the training steps above moved the BatchNorm statistics)
    net = svr_b200.IFNet().to(dev).eval()
    net.load_state_dict(R.synthetic_state_dict(61, 128), strict=False)
    lattice = (32, 24, 16)
    xb, xe = svr_dist.shard_range(lattice[0] // 8, rank, world)
    slab = net.evaluate_grid(x[:1].to(dev), lattice, scenes=[0], x_range=(xb * 8, xe * 8))
    torch.save(slab.cpu(), Path(out_dir) / f"d{rank}.pt")
    dist.destroy_process_group()


def test_dp_gradients_and_sharded_dense_eval_two_ranks(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    import svr_b200
    from oracle import ref_torch as R
    from svr_b200 import dist as svr_dist
    svr_b200.configure(net_res=128, precision=16)
    x, pts, occ = _inputs()
    sd = R.synthetic_state_dict(61, 128)
    want = None
    for rank in range(world):
        net = svr_b200.IFNet().cuda().train()
        net.load_state_dict(sd, strict=False)
        b, e = svr_dist.shard_range(4, rank, world)
        _step(net, x[b:e].cuda(), pts[b:e].cuda(), occ[b:e].cuda())
        g = {n: p.grad.detach().double().cpu() for n, p in net.named_parameters()}
        want = g if want is None else {n: want[n] + g[n] for n in g}
    want = {n: v / world for n, v in want.items()}
    g0, g1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    for n, w in want.items():
        assert torch.equal(g0[n], g1[n]), n                       # every rank holds the same reduced gradient
        err = float((g0[n].double() - w).norm() / w.norm().clamp_min(1e-30))
        assert err < 2e-3, (n, err)       # fp32 atomics order in the volume-gradient scatter + bf16 dz rounding of re-run kernels
    # dense evaluation: the two slabs reassemble the single-GPU result
    net = svr_b200.IFNet().cuda().eval()
    net.load_state_dict(sd, strict=False)
    full = net.evaluate_grid(x[:1].cuda(), (32, 24, 16), scenes=[0])[0].cpu()
    d0, d1 = torch.load(tmp_path / "d0.pt")[0], torch.load(tmp_path / "d1.pt")[0]
    assert torch.equal(d0 + d1, full)
