"""Summarise ncu artefacts into profiles/ (text):  python tools/ncu_summary.py rep <file.ncu-rep> | launches <file.csv> [steps]"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]


def rep(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print(f"## {d.get('Kernel Name', ('', '?'))[1]}")
        for k in KEYS:
            if k in d:
                print(f"{k:90s} {d[k][1]:>18s} {d[k][0]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if hi:
        h = rows[hi[0]]
        si, ni = h.index("Source"), h.index("# Samples")
        data = [(int(r[ni]), r[si].strip()) for r in rows[hi[0] + 1:] if len(r) > ni and r[ni].isdigit()]
        tot = sum(n for n, _ in data) or 1
        print("\n## hottest SASS instructions by warp-stall samples")
        for n, s in sorted(data, reverse=True)[:18]:
            print(f"{100 * n / tot:6.2f}%  {s[:110]}")


def launches(path, steps=1):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        m = re.search(r"(\w+_kernel(<[^>]*>)?)", r[ki])
        name = m.group(1) if m else r[ki][:60]
        v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ui], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# per-step kernel time from ncu gpu__time_duration (cold-cache, serialised; {steps} step(s) captured)")
    print(f"{'ms/step':>10s} {'share':>7s} {'launches/step':>14s} {'avg us':>10s}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / steps:10.3f} {100 * v[1] / tot:6.1f}% {v[0] / steps:14.1f} {v[1] / v[0] * 1000:10.1f}  {k}")
    print(f"{tot / steps:10.3f}  total per step")


if __name__ == "__main__":
    if sys.argv[1] == "rep":
        rep(sys.argv[2])
    else:
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 1)
