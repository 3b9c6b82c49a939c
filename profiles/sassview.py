#!/usr/bin/env python
"""Hot spots of one kernel from the SASS source page of an ncu report, exported with
`ncu -i report.ncu-rep --page source --csv --print-source sass -k regex:NAME > NAME_sass.csv` (the export travels back from the
GPU box, the report does not).  Prints, per kernel in the file: the split of warp samples / executed instructions between the
main loop (the most executed backward branch) and the code before / after it, and the most sampled instructions with their
top stall reasons.   usage: sassview.py NAME_sass.csv [top_n]"""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kernels.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
done = set()
for k in kernels:
    h, data = k["rows"][0], k["rows"][1:]
    ia, isrc, ismp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall = {n: h.index(n) for n in h if n.startswith("stall_") and "(Not" not in n}
    recs, seen = [], set()
    for r in data:
        if len(r) <= iex or not r[ia].startswith("0x"):
            continue
        a = int(r[ia], 16)
        if a in seen:
            break
        seen.add(a)
        recs.append((a, r[isrc].strip(), int(r[ismp] or 0), int(r[iex] or 0), {n: int(r[i] or 0) for n, i in stall.items()}))
    if not recs:
        continue
    base = recs[0][0]
    ts, te = sum(r[2] for r in recs) or 1, sum(r[3] for r in recs) or 1
    name = re.sub(r"\((?!int\))[^()]*\)$", "", k["name"])          # drop the parameter list, keep template arguments
    if (name, ts, te) in done:                                      # the export repeats a kernel once per launch it matched
        continue
    done.add((name, ts, te))
    print("kernel %s\n  %d SASS instructions, %d warp samples, %d warp instructions executed" % (name, len(recs), ts, te))
    loop = None
    for a, src, smp, ex, _ in recs:
        m = re.search(r"BRA\s+0x([0-9a-f]+)", src)
        if m and int(m.group(1), 16) < a and a - int(m.group(1), 16) > 0x400 and (loop is None or ex > loop[2]):
            loop = (int(m.group(1), 16), a, ex)
    if loop:
        lo, hi, iters = loop
        print("  main loop 0x%x..0x%x, %d iterations" % (lo - base, hi - base, iters))
        for name, f in (("before the loop", lambda a: a < lo), ("loop", lambda a: lo <= a <= hi), ("after the loop", lambda a: a > hi)):
            s_ = sum(r[2] for r in recs if f(r[0]))
            e_ = sum(r[3] for r in recs if f(r[0]))
            print("    %-16s %5.1f %% of samples  %5.1f %% of instructions  %7.1f instructions per iteration" % (name, 100.0 * s_ / ts, 100.0 * e_ / te, e_ / max(iters, 1)))
    print("  most sampled instructions:")
    for a, src, smp, ex, st in sorted(recs, key=lambda r: -r[2])[:top_n]:
        why = ", ".join("%s %.1f" % (n[6:], 100.0 * v / ts) for n, v in collections.Counter(st).most_common(2) if v)
        print("    0x%05x %5.1f %%  executed %12d  %-58s %s" % (a - base, 100.0 * smp / ts, ex, src[:58], why))
    print()
