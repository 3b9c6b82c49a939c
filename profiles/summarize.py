#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (tracked).
usage: summarize.py <tag> [launches.csv] [prof_a.ncu-rep | prof_a_raw.csv ...]"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def launches(path, out):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        name = r[ki].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write("%-36s launches %4d  total %9.3f ms  share %5.1f%%\n" % (k, v[0], v[1], 100 * v[1] / tot))


def report(path, out):
    """a .ncu-rep, or the `ncu -i rep --page raw --csv` export of one (what travels back from the GPU box)"""
    if path.endswith(".csv"):
        txt = "".join(l for l in open(path) if not l.startswith("=="))
    else:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units = rows[0], rows[1]
    for v in rows[2:]:
        out.write("\n# %s : kernel %s\n" % (path.split("/")[-1], v[h.index("Kernel Name")]))
        for w in WANT:
            if w in h:
                i = h.index(w)
                out.write("%-90s %14s %s\n" % (w, v[i], units[i]))


if __name__ == "__main__":
    tag = sys.argv[1]
    with open("profiles/%s.txt" % tag, "w") as out:
        for a in sys.argv[2:]:
            (launches if a.endswith(".csv") and not a.endswith("_raw.csv") else report)(a, out)
    print(open("profiles/%s.txt" % tag).read())
