"""CPU suite, world_size 2 over gloo: the N>1 host logic (scene sharding, bucketed gradient
all-reduce with the early decoder bucket) gives the single-process large-batch gradient."""
import os
import socket
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.enc = torch.nn.Linear(6, 8)
        self.fc_0 = torch.nn.Linear(8, 8)
        self.fc_out = torch.nn.Linear(8, 1)

    def forward(self, x):
        return self.fc_out(torch.relu(self.fc_0(torch.relu(self.enc(x))))).squeeze(-1)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(REPO))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from svr_b200 import dist as svr_dist
    torch.manual_seed(0)
    net = _Toy()
    red = svr_dist.GradReducer(net)
    g = torch.Generator().manual_seed(1)
    batch = {"x": torch.randn((6, 5, 6), generator=g), "y": torch.randn((6, 5), generator=g), "meta": "keep"}
    mine = svr_dist.shard_scenes(batch, rank, world)
    assert mine["meta"] == "keep" and mine["x"].shape[0] == 3
    for _ in range(2):   # two steps: the hook/bucket state must reset
        net.zero_grad()
        loss = ((net(mine["x"]) - mine["y"]) ** 2).sum(-1).mean()
        loss.backward()
        red.allreduce()
    torch.save({n: p.grad.clone() for n, p in net.named_parameters()}, Path(out_dir) / f"g{rank}.pt")
    # gradient accumulation: two micro-batches (backward twice) before ONE allreduce -- the bucket launched from the
    # hooks of the first backward is stale and must be reduced again
    for keep_views in (False, True):
        net.zero_grad(set_to_none=not keep_views)     # keep_views: p.grad stays a view of the bucket buffer
        for half in (slice(0, 2), slice(2, 3)):
            loss = ((net(mine["x"][half]) - mine["y"][half]) ** 2).sum(-1).sum() / 3.0
            loss.backward()
        red.allreduce()
        torch.save({n: p.grad.clone() for n, p in net.named_parameters()}, Path(out_dir) / f"acc{int(keep_views)}_{rank}.pt")
    # unequal shards: rank 0 takes 4 scenes, rank 1 takes 2; weights n_local / n_global
    n_loc = 4 if rank == 0 else 2
    sl = slice(0, 4) if rank == 0 else slice(4, 6)
    red_w = svr_dist.GradReducer(net, weight=n_loc / 6.0)
    for h in red._hooks:
        h.remove()
    net.zero_grad()
    ((net(batch["x"][sl]) - batch["y"][sl]) ** 2).sum(-1).mean().backward()
    red_w.allreduce()
    torch.save({n: p.grad.clone() for n, p in net.named_parameters()}, Path(out_dir) / f"w{rank}.pt")
    dist.destroy_process_group()


def test_grad_reducer_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g0, g1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    torch.manual_seed(0)
    net = _Toy()
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn((6, 5, 6), generator=g), torch.randn((6, 5), generator=g)
    ((net(x) - y) ** 2).sum(-1).mean().backward()
    for n, p in net.named_parameters():
        assert torch.allclose(g0[n], g1[n])
        assert torch.allclose(g0[n], p.grad, rtol=1e-5, atol=1e-6), n
        for tag in ("acc0", "acc1", "w"):
            a0, a1 = torch.load(tmp_path / f"{tag}_{0}.pt" if tag != "w" else tmp_path / "w0.pt"), torch.load(
                tmp_path / f"{tag}_{1}.pt" if tag != "w" else tmp_path / "w1.pt")
            assert torch.allclose(a0[n], a1[n]), (tag, n)
            assert torch.allclose(a0[n], p.grad, rtol=1e-5, atol=1e-6), (tag, n)


def test_shard_range_covers_everything():
    sys.path.insert(0, str(REPO))
    from svr_b200 import dist as svr_dist
    for n in (0, 1, 7, 8, 256):
        for w in (1, 2, 3, 8):
            blocks = [svr_dist.shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in blocks) - min(e - b for b, e in blocks) <= 1
