#!/usr/bin/env python
"""dram bytes per launch of every stage kernel from an `ncu --set full` report -> profiles/traffic.json
(read by bench.py for roofline.traffic).  usage: mk_traffic.py report.ncu-rep [workload tag] [stage launches covered by the capture]"""
import collections, csv, json, os, subprocess, sys

STAGE = [("sketch_kernel", "sketch"), ("seed_kernel", "seed"), ("anchor_filter_kernel", "expand"), ("expand_kernel", "expand"),
         ("radix_sort_kernel", "sort"), ("sort_kernel", "sort"), ("chain_dp_kernel", "chain_dp"), ("backtrack_kernel", "backtrack"),
         ("rechain_kernel", "rechain"), ("regs_kernel", "regs"), ("ext_fill_kernel", "extend"), ("ext_dp_kernel", "extend"), ("ext_prep_kernel", "extend"), ("ext_stitch_kernel", "extend"), ("ext_job_scan_kernel", "extend")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep = sys.argv[1]
    tag = sys.argv[2] if len(sys.argv) > 2 else ""
    if rep.endswith(".csv"):   # the raw-page export of a report
        txt = "".join(l for l in open(rep) if not l.startswith("=="))
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units = rows[0], rows[1]
    kn, rd, wr, du = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
    per = collections.defaultdict(lambda: [0, 0.0, 0.0])
    missing = collections.defaultdict(list)
    n_chunks = int(sys.argv[3]) if len(sys.argv) > 3 else 1   # stage launches (chunks) the capture covers
    seen_sketch = 0
    for r in rows[2:]:
        if "sketch_kernel" in r[kn]:      # a chunk starts with its sketch: stop after the chunks asked for
            seen_sketch += 1
            if seen_sketch > n_chunks:
                break
        st = next((s for k, s in STAGE if k in r[kn]), None)
        if st is None:
            continue
        b = float(r[rd].replace(",", "")) * UNIT.get(units[rd], 1.0) + float(r[wr].replace(",", "")) * UNIT.get(units[wr], 1.0)
        per[st][0] += 1
        if b != b:   # ncu returned no dram counters for this launch (seen on ext_fill_kernel<6>): say so instead of summing a NaN
            missing[st].append(r[kn].split("(")[0])
            b = 0.0
        per[st][1] += b
        per[st][2] += float(r[du].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[du], 1.0)
    out = {st: {"dram_bytes_per_launch": v[1] / n_chunks, "kernels_captured": v[0], "stage_launches_captured": n_chunks, "ms_under_ncu_per_launch": v[2] / n_chunks,
                "report": os.path.basename(rep), "workload": tag, "kernels_without_dram_counters": missing.get(st, [])} for st, v in per.items()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
