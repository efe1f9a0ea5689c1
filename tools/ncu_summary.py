"""Summarise ncu outputs into small text files that can be committed under profiles/.

  python tools/ncu_summary.py rep  gpurun_out/x.ncu-rep  profiles/x.txt      # --set full capture(s)
  python tools/ncu_summary.py list gpurun_out/launches.csv profiles/y.txt    # launch list (shares)
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none summary of {path}\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"\n== kernel {d.get('Kernel Name')}  grid {d.get('Grid Size')}  block {d.get('Block Size')}\n")
            for k in KEYS:
                if k in d:
                    f.write(f"{k:90s} {d[k]:>18s} {units[hdr.index(k)]}\n")
            rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
            f.write(f"{'traffic = dram read + write (units as above)':90s} {rd} + {wr}\n")


def launch_list(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v_us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        name = r["Kernel Name"]
        tot[name][0] += 1
        tot[name][1] += v_us
    total = sum(t for _, t in tot.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list ({path}): per-kernel totals, cold-cache serialised times -- compare SHARES\n")
        f.write(f"# total {total:.1f} us over {sum(c for c, _ in tot.values())} launches\n")
        for name, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"{100 * t / total:6.2f} %  {t:12.1f} us  {c:5d} x  {name[:110]}\n")


if __name__ == "__main__":
    {"rep": rep, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
