#!/usr/bin/env python
"""Per-CUDA-line hot spots of one kernel in an ncu report (needs -lineinfo and --import-source on).
usage: srcview.py report.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
data, fname, h = [], "", None
for r in csv.reader(txt.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        h = r
        ci = {n: h.index(n) for n in ("# Samples", "Instructions Executed", "Thread Instructions Executed")}
    elif h and r[0].isdigit() and len(r) > ci["Thread Instructions Executed"]:
        try:
            data.append((int(r[ci["Instructions Executed"]] or 0), int(r[ci["# Samples"]] or 0), int(r[ci["Thread Instructions Executed"]] or 0), "%s:%s" % (fname, r[0]), r[1].strip()))
        except ValueError:
            pass
ti = sum(d[0] for d in data) or 1
ts = sum(d[1] for d in data) or 1
print("kernel %s: %d warp instructions, %d samples" % (kern, ti, ts))
for d in sorted(data, key=lambda d: -d[1])[:top]:
    print("%5.1f%% inst %5.1f%% smpl thr/inst %4.1f  %-20s %s" % (100 * d[0] / ti, 100 * d[1] / ts, d[2] / max(d[0], 1), d[3], d[4][:100]))
