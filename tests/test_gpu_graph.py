"""svr_b200.GraphedStep: a whole training step (forward, BCE, backward, fused Adam) recorded as one CUDA graph must follow
the eager trajectory.  The step's kernels are deterministic (stable sort, fixed-order reductions, no floating-point atomics
on the weights' path except the volume scatter), so the losses agree to round-off."""
import copy

import pytest
import torch

from oracle import ref_torch as R

pytestmark = pytest.mark.gpu


def test_graphed_training_step_follows_eager_trajectory():
    import svr_b200
    svr_b200.configure(net_res=128, precision=16)
    sd = R.synthetic_state_dict(41, 128)
    g = torch.Generator().manual_seed(5)
    x = (torch.rand((2, 1, 64, 64, 64), generator=g) < 0.08).float().cuda()
    pts = (torch.rand((2, 4096, 3), generator=g) - 0.5).cuda()
    occ = (torch.rand((2, 4096), generator=g) < 0.5).float().cuda()
    losses = {}
    for mode in ("eager", "graph"):
        net = svr_b200.IFNet().cuda()
        net.load_state_dict(sd, strict=False)
        net.train()
        opt = torch.optim.Adam(net.parameters(), lr=2e-5, fused=True, capturable=True)

        def step(xd, pd, od):
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.binary_cross_entropy_with_logits(net(xd, pd), od, reduction="none").sum(-1).mean()
            loss.backward()
            opt.step()
            return loss

        run = step if mode == "eager" else svr_b200.GraphedStep(step, (x, pts, occ), warmup=0)
        out = []
        for _ in range(4):
            out.append(float(run(x, pts, occ)))
        losses[mode] = out
    # the graph is captured on its first call (which executes nothing), so its n-th replay is the eager n-th step
    assert losses["graph"][0] == pytest.approx(losses["eager"][0], rel=1e-5)
    # later steps: Adam's first updates are sign-like, so the round-off of the volume scatter's atomics (the one
    # order-dependent reduction of the step) is amplified from step to step -- in eager mode as well
    for a, b in zip(losses["eager"], losses["graph"]):
        assert b == pytest.approx(a, rel=5e-3)
    assert losses["graph"][-1] != losses["graph"][0]          # the replays do update the weights
