import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, svr_b200
from svr_b200 import _abi
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
depth = (torch.rand((64, 256, 256), generator=g) * 5.0 + 0.5).to(dev)
for dims, scale in (((128,) * 3, 1), ((256,) * 3, 0.5)):
    proj = svr_b200.project(dims, [3, 3, 3], torch.tensor([1.5, 1.5, 1.5])).to(dev)
    with torch.no_grad():
        for _ in range(3):
            out = proj(proj.depthmap_to_normed_points(depth, scale))
        torch.cuda.synchronize()
        _abi.PROFILE.reset(with_events=True)
        for _ in range(5):
            out = proj(proj.depthmap_to_normed_points(depth, scale))
        torch.cuda.synchronize()
        print(dims[0], {k: round(t / 5, 3) for k, (c, t) in _abi.PROFILE.kernel_ms().items()})
        _abi.PROFILE.reset()
    del proj, out
    torch.cuda.empty_cache()
