"""Secondary BASELINE configs measured on one B200 (numbers quoted in DESIGN.md):
  config 3: depth -> voxel projection sweep, batch 64 maps of 256x256 into 128^3 and 256^3 grids
  config 5: dense occupancy query for mesh extraction, 256^3 lattice (16.7 M points) per scene
CUDA-event timing, 3 warm-ups, inputs resident in HBM; prints one JSON line per measurement."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import svr_b200

PEAK = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text()) if (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
dev = torch.device("cuda:0")


def timed(fn, iters=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# ---------------- config 3
B = 64
g = torch.Generator().manual_seed(0)
depth = (torch.rand((B, 256, 256), generator=g) * 5.0 + 0.5).to(dev)
for dims, scale in (((128, 128, 128), 1), ((256, 256, 256), 0.5)):
    proj = svr_b200.project(dims, [3, 3, 3], torch.tensor([1.5, 1.5, 1.5])).to(dev)
    with torch.no_grad():
        pts = proj.depthmap_to_normed_points(depth, scale)
        ms_vox = timed(lambda: proj.pc_voxels(pts))
        ms_full = timed(lambda: proj(proj.depthmap_to_normed_points(depth, scale)))
    grid_bytes = B * dims[0] * dims[1] * dims[2] * 4
    print(json.dumps({"config": 3, "grid": dims[0], "maps": B, "pc_voxels_ms": ms_vox, "pc_voxels_maps_per_s": B / ms_vox * 1e3,
                      "pc_voxels_hbm_frac": (grid_bytes + B * 65536 * 12) / (ms_vox * 1e-3) / 1e9 / PEAK["hbm_gbs"],
                      "project_forward_ms": ms_full, "project_forward_maps_per_s": B / ms_full * 1e3,
                      "project_forward_hbm_frac": (grid_bytes + B * 65536 * 4) / (ms_full * 1e-3) / 1e9 / PEAK["hbm_gbs"]}), flush=True)
    del proj, pts
    torch.cuda.empty_cache()

# ---------------- config 5
svr_b200.configure(net_res=128)
torch.manual_seed(0)
net = svr_b200.IFNet().to(dev).eval()
x = (torch.rand((2, 1, 128, 128, 128), generator=g) < 0.05).float().to(dev)
lattice = (256, 256, 256)
with torch.no_grad():
    ms = timed(lambda: net.evaluate_grid(x, lattice, scenes=[0]), iters=3, warm=2)
    vols = net.ifnet_feature_extractor.encode(x[:1])
    ms_enc = timed(lambda: net.ifnet_feature_extractor.encode(x[:1]), iters=3, warm=1)
npts = 256 ** 3
flops = 2 * npts * (2624 * 256 + 2 * 256 * 256 + 256)
print(json.dumps({"config": 5, "lattice": 256, "points": npts, "ms_per_scene": ms, "points_per_s": npts / ms * 1e3, "encoder_ms": ms_enc,
                  "tensor_frac": flops / ((ms - ms_enc) * 1e-3) / 1e12 / PEAK["bf16_tflops_sustained"]}), flush=True)
