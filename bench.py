"""Benchmark of the IF-Net implicit query hot path (BASELINE.json metric: IF-Net query-points/sec,
fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE.json configs[1] -- IF-Net training step fwd+bwd, batch 4
scenes x 50 000 query points, 128^3 occupancy grid per GPU (weak scaling: N GPUs = 4N scenes, which at
N=8 is configs[3]'s 32 scenes).  A step = IFNet.forward (torch/cuDNN encoder + our gather/decoder
kernels) + BCE loss + backward + DP gradient all-reduce (N>1) + Adam step.

  value : whole-job query-points/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through the public module API with HOST (pinned) inputs; the H2D copies of the
          voxel grids / points / labels and the D2H read of the loss are inside the timed region
  roofline : dominant kernel of the step, timed live with CUDA events on the launching stream
  cpu_baseline : the oracle port (oracle/ref_torch.py, torch CPU ops == the reference's arithmetic)
          on the host cores, bounded sample, rank 0 / N=1 only
  --impl reference : the reference's CPU implementation of the same step (the oracle port: the
          reference is Python and does not exist on the GPU box), all host threads, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SCENES_PER_GPU = 4
POINTS = 50_000
GRID = (128, 128, 128)
DEPTH_HW = (256, 256)
METRIC = "ifnet_query_points_per_sec_fwd_bwd"
UNIT = "points/s"


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []          # (arrival time, csv line)
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_begin(self):
        """Start of the timed region.  The sampler process is launched BEFORE the warm-up so that neither the fork
        nor nvidia-smi's NVML start-up falls inside the timed steps; only samples between the marks are reported."""
        self.t0 = time.time()

    def stop(self):
        self.t1 = time.time()
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        inside = [ln for t, ln in self.lines if self.t0 is not None and self.t0 <= t <= self.t1 + 0.1]
        if not inside:            # region shorter than the 100 ms period: take the samples closest to it
            inside = [ln for _, ln in self.lines[-2:]]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def synthetic_inputs(n_scenes: int, seed: int, device):
    """BASELINE/SURVEY 8d synthetic inputs: uniform depth maps U(0.5,5.5) -> our projection -> 128^3
    occupancy grids; query points U(-0.5,0.5)^3; Bernoulli(0.5) labels."""
    import torch
    import svr_b200
    g = torch.Generator().manual_seed(seed)
    depth = torch.rand((n_scenes,) + DEPTH_HW, generator=g) * 5.0 + 0.5
    pts = torch.rand((n_scenes, POINTS, 3), generator=g) - 0.5
    occ = (torch.rand((n_scenes, POINTS), generator=g) < 0.5).float()
    proj = svr_b200.project(GRID, [3, 3, 3], torch.tensor([1.5, 1.5, 1.5])).to(device)
    with torch.no_grad():
        x = proj(proj.depthmap_to_normed_points(depth.to(device), 1))
    return x.contiguous(), pts, occ


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# ------------------------------------------------------------------------------------------------
def cpu_step_factory(n_points: int, seed: int = 0):
    import torch
    from oracle import ref_torch as R
    g = torch.Generator().manual_seed(seed)
    sd = R.synthetic_state_dict(0, 128)
    params = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    x = (torch.rand((1, 1) + GRID, generator=g) < 0.03).float()
    pts = torch.rand((1, n_points, 3), generator=g) - 0.5
    occ = (torch.rand((1, n_points), generator=g) < 0.5).float()

    def step():
        for p in params.values():
            p.grad = None
        logits = R.ifnet_forward(params, x, pts, 128, training=True)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, occ, reduction="none").sum(-1).mean()
        loss.backward()
        return float(loss)
    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    n_pts = 25_000
    step = cpu_step_factory(n_pts)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = n_pts * args.steps / dt
    sample = f"1 scene x {n_pts} points per step (128^3 grid, encoder + sampling + decoder fwd+bwd), torch CPU ops"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "IF-Net training step fwd+bwd, 128^3 grid, 50k query points/scene (BASELINE configs[1])",
                       "scenes_per_gpu": SCENES_PER_GPU, "points_per_scene": POINTS, "grid": list(GRID)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count()},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import svr_b200
    from svr_b200 import _abi
    from svr_b200 import dist as svr_dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _abi.check(_abi.load().svr_device_info(None, None, None, None), "device_info")

    svr_b200.configure(net_res=128, channels_last=os.environ.get("SVR_CHANNELS_LAST", "1") == "1")
    torch.backends.cudnn.benchmark = os.environ.get("SVR_CUDNN_BENCHMARK", "1") == "1"   # reference: trainer_ifnet.py:64
    torch.manual_seed(0)
    net = svr_b200.IFNet().to(dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
    reducer = svr_dist.GradReducer(net) if world > 1 else None
    x, pts_h, occ_h = synthetic_inputs(SCENES_PER_GPU, 100 + rank, dev)
    pts, occ = pts_h.to(dev), occ_h.to(dev)
    # host copies for the e2e leg
    x_pin, pts_pin, occ_pin = x.cpu().pin_memory(), pts_h.pin_memory(), occ_h.pin_memory()
    n_pts_step = SCENES_PER_GPU * POINTS

    def step(xd, pd, od):
        opt.zero_grad(set_to_none=True)
        logits = net(xd, pd)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, od, reduction="none").sum(-1).mean()
        loss.backward()
        if reducer is not None:
            reducer.allreduce()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # L2 hygiene: the step streams > 1 GB of volumes/features per iteration (inputs larger than the 126 MB L2)
    for _ in range(max(args.warmup, 3)):
        step(x, pts, occ)
    # ---------------- timed: device-resident inputs
    _abi.PROFILE.reset(with_events=False)
    barrier()
    clocks.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.profile_mode:
        torch.cuda.profiler.start()      # ncu --profile-from-start off: only the timed steps are captured
    e0.record()
    for _ in range(args.steps):
        step(x, pts, occ)
    e1.record()
    barrier()
    if args.profile_mode:
        torch.cuda.profiler.stop()
    clk = clocks.stop() if rank == 0 else None
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = _abi.PROFILE.total_launches()
    value = world * n_pts_step * args.steps / (ms * 1e-3)
    if args.profile_mode:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": ms / args.steps, "gpu_launches": launches,
                              "profile_mode": True}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- timed: end to end from pinned host memory through the public API
    # Every step copies ITS inputs from pinned host memory (svr_b200.HostPrefetcher: the copy of step i+1 is issued on a
    # side stream while step i computes -- what a pinned DataLoader gives a trainer) and reads its loss back to the host
    # (4 bytes D2H into pinned memory).  The read is asynchronous with a lag of two steps, like a trainer that logs the
    # loss without stalling the launch queue: the host consumes the loss of step i-2 while step i is being enqueued, and
    # the last two are consumed before the timed region ends.
    pf = svr_b200.HostPrefetcher(dev)
    host_batch = (x_pin, pts_pin, occ_pin)
    loss_pin = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(n):
        losses = []
        nxt = pf.issue(host_batch)
        for i in range(n):
            xd, pd, od = pf.wait(nxt)
            if i + 1 < n:
                nxt = pf.issue(host_batch)
            if i >= 2:                                  # result of step i-2 (its slot is reused now)
                loss_ev[i & 1].synchronize()
                losses.append(float(loss_pin[i & 1]))
            loss = step(xd, pd, od)
            loss_pin[i & 1].copy_(loss.detach(), non_blocking=True)
            loss_ev[i & 1].record()
        for i in range(max(n - 2, 0), n):
            loss_ev[i & 1].synchronize()
            losses.append(float(loss_pin[i & 1]))
        assert len(losses) == n and all(l == l for l in losses)
        return losses

    e2e_loop(3)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * n_pts_step * args.steps / (ms_e2e * 1e-3)
    h2d = x_pin.numel() * 4 + pts_pin.numel() * 4 + occ_pin.numel() * 4

    # ---------------- per-kernel timing (CUDA events on the launching stream, separate pass)
    _abi.PROFILE.reset(with_events=True)
    barrier()
    prof_steps = min(args.steps, 5)
    e0.record()
    for _ in range(prof_steps):
        step(x, pts, occ)
    e1.record()
    barrier()
    step_ms_prof = e0.elapsed_time(e1) / prof_steps
    kms = {k: (c / prof_steps, t / prof_steps) for k, (c, t) in _abi.PROFILE.kernel_ms().items()}
    _abi.PROFILE.reset(with_events=False)

    if rank == 0:
        peaks = _peaks()
        roof = roofline(kms, peaks, net)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": "IF-Net training step fwd+bwd, batch 4 scenes x 50k query points, 128^3 grid per GPU "
                                       "(BASELINE configs[1]; N GPUs = 4N scenes, N=8 is configs[3])",
                           "scenes_per_gpu": SCENES_PER_GPU, "points_per_scene": POINTS, "grid": list(GRID),
                           "step": "encoder(torch/cuDNN)+gather+decoder fwd, BCE, bwd, allreduce(N>1), Adam",
                           "l2": "inputs larger than L2 (volumes + features > 1 GB per step)", "parallelism": f"dp{world}"},
                "clocks": clk, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / args.steps,
                        "how": "public API (svr_b200.IFNet + HostPrefetcher): every step's inputs are copied from pinned host memory "
                               "inside the timed region (the copy of step i+1 overlaps step i) and every step's loss is read back "
                               "(async D2H, consumed with a lag of two steps, all consumed before the region ends)"},
                "roofline": roof,
                "kernels_ms_per_step": {k: {"calls": c, "ms": round(t, 4)} for k, (c, t) in sorted(kms.items(), key=lambda kv: -kv[1][1])},
                "hot_path_ms_per_step": round(sum(t for _, t in kms.values()), 4), "step_ms_profiled": round(step_ms_prof, 4)}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _ncu_traffic(profile_name: str, kernel_substr: str):
    """dram read + write bytes per launch of a kernel from a committed `ncu --set full` summary (profiles/)."""
    f = ROOT / "profiles" / profile_name
    if not f.exists():
        return None
    unit = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
    cur, rd, wr = False, None, None
    for line in f.read_text().splitlines():
        if line.startswith("== kernel"):
            cur = kernel_substr in line
        elif cur and line.startswith("dram__bytes_read.sum "):
            v = line.split()
            rd = float(v[-2]) * unit.get(v[-1], 1.0)
        elif cur and line.startswith("dram__bytes_write.sum "):
            v = line.split()
            wr = float(v[-2]) * unit.get(v[-1], 1.0)
    return None if rd is None or wr is None else rd + wr


def roofline(kms, peaks, net):
    """Roofline of the dominant hot-path kernel (largest share of the step among the IF-Net query kernels).
    Algorithmic bytes / flops per launch are the figures of DESIGN.md section 4."""
    hot = {k: v for k, v in kms.items() if k in ("svr_query_fwd_fused", "svr_gather_fwd", "svr_gather_bwd", "svr_gemm_nt", "svr_gemm_tn")}
    if not hot:
        return None
    name, (calls, ms) = max(hot.items(), key=lambda kv: kv[1][1])
    M = SCENES_PER_GPU * POINTS
    kp = 2624
    vol_elems = sum(c * (GRID[0] >> s) ** 3 for c, s in ((16, 0), (32, 1), (64, 2), (128, 3), (128, 4)))
    x_bytes = SCENES_PER_GPU * GRID[0] ** 3 * 4
    vols_bf16 = SCENES_PER_GPU * vol_elems * 2
    fwd_flops = 2 * M * (kp * 256 + 2 * 256 * 256 + 256)
    out = {"kernel": name, "ms_per_step": ms, "launches_per_step": calls, "traffic": None, "peak_source": peaks["source"]}
    if name == "svr_query_fwd_fused":
        # one launch: gather + fc_0..fc_out.  Tensor-core bound by the decoder FLOPs;
        # the achieved HBM rate on its compulsory bytes is reported next to it.
        nbytes = vols_bf16 + x_bytes + M * 16 + M * kp * 2 + 3 * M * 256 * 2     # + saved features / activations (training)
        ach = fwd_flops / (ms * 1e-3) / 1e12
        out.update({"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach / peaks["bf16_tflops_sustained"], "algorithmic_flops": fwd_flops, "algorithmic_bytes": nbytes,
                    "hbm_achieved_gbs": nbytes / (ms * 1e-3) / 1e9, "hbm_frac": nbytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                    "peak_source": peaks["source"] + " (sustained bf16)",
                    "traffic": _ncu_traffic("r1_query_path_v4.txt", "fused_query_kernel"),
                    "traffic_note": "dram read+write per launch, ncu --set full of the same training launch "
                                    "(profiles/r1_query_path_v4.txt)"})
    elif name in ("svr_gather_fwd", "svr_gather_bwd"):
        # gather: volumes + grid in, feature rows out; scatter: d-feature rows in, fp32 gradient volumes written once
        nbytes = (vols_bf16 + x_bytes + M * 12 + M * kp * 2) if name == "svr_gather_fwd" else (M * kp * 2 + 2 * vols_bf16 + M * 12)
        ach = nbytes / (ms * 1e-3) / 1e9
        out.update({"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                    "algorithmic_bytes": nbytes})
        if name == "svr_gather_bwd":       # one call = tensor-core scatter (coarse levels) + direct scatter (fine levels)
            parts = [_ncu_traffic("r1_query_path_v4.txt", k) for k in ("scatter_tc_kernel", "gather_bwd_kernel")]
            if all(v is not None for v in parts):
                out["traffic"] = sum(parts)
                out["traffic_note"] = ("dram read+write of scatter_tc_kernel + gather_bwd_kernel, ncu --set full of the same training step "
                                       "(profiles/r1_query_path_v4.txt); above the algorithmic bytes because the fp32 gradient volumes are "
                                       "read-modify-written by the L2 atomic units")
    else:
        flops = fwd_flops   # backward-data (dz1, dz0, dfeat) resp. weight-gradient (dW2, dW1, dW0) GEMMs: one forward's worth each
        ach = flops / (ms * 1e-3) / 1e12
        out.update({"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach / peaks["bf16_tflops_sustained"], "algorithmic_flops": flops,
                    "peak_source": peaks["source"] + " (sustained bf16)"})
    return out


def cpu_baseline():
    import torch
    step = cpu_step_factory(POINTS)
    step()                      # warm-up (allocators, oneDNN primitives)
    t0 = time.perf_counter()
    step()
    dt = time.perf_counter() - t0
    return {"value": POINTS / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "host_cpus": os.cpu_count(),
            "sample": f"1 scene x {POINTS} points, 1 timed step after 1 warm-up (of the 4-scene workload; scenes are independent), "
                      f"oracle/ref_torch.py on torch CPU ops", "seconds": dt}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--profile-mode", dest="profile_mode", action="store_true",
                    help="warm-up + timed device-resident steps only (the command profiled under ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
