#!/usr/bin/env python
"""Multi-device aligner vs one device on a multi-chunk batch (fault isolation / timing).
usage: group_check.py [n_reads] [cigar 0|1] [mode: all|single|half|group]"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "oracle", "mappy-rs_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import torch
import data_gen, parity
from mappy_rs import _mmg
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
cigar = int(sys.argv[2]) if len(sys.argv) > 2 else 1
mode = sys.argv[3] if len(sys.argv) > 3 else "all"
lib = _mmg.Lib()
ref, coff, names = data_gen.config1_reference()
io, mo = _mmg.IdxOpt(), _mmg.MapOpt()
lib.check(lib.L.mmg_set_opt(None, ctypes.byref(io), ctypes.byref(mo)))
mo.flag = 4 if cigar else 0
idx = _mmg.Index.build(lib, io, names, [ref.tobytes()], device=0)
lib.check(lib.L.mmg_mapopt_update(ctypes.byref(mo), idx.h))
buf, offs, _ = data_gen.config1_reads(ref, coff, n_reads)
hb = torch.empty(len(buf), dtype=torch.uint8, pin_memory=True); hb.numpy()[:] = buf
nd = torch.cuda.device_count()
a = None
if mode in ("all", "single", "half"):
    one = _mmg.DeviceAligner(lib, idx, mo, device=0)
    m = n_reads // 2 if mode == "half" else n_reads
    for rep in range(3):
        t0 = time.time(); a = one.map_batch(hb.numpy()[:int(offs[m])], offs[:m + 1]); t1 = time.time() - t0
        print("one device, %d reads, rep %d: %d hits in %.2f s" % (m, rep, len(a.hits), t1), flush=True)
    one.close()
if mode in ("keep", "small_first"):   # variations of "all": the one-device aligner stays open / maps only a small batch first
    one = _mmg.DeviceAligner(lib, idx, mo, device=0)
    m = 2000 if mode == "small_first" else n_reads
    a0 = one.map_batch(hb.numpy()[:int(offs[m])], offs[:m + 1])
    print("one device: %d hits" % len(a0.hits), flush=True)
    if mode == "small_first":
        one.close()
    grp = _mmg.DeviceAligner(lib, idx, mo, devices=list(range(nd)))
    for rep in range(2):
        b = grp.map_batch(hb.numpy(), offs)
        print("%d devices, rep %d: %d hits" % (nd, rep, len(b.hits)), flush=True)
    grp.close()
    if mode == "keep":
        one.close()
if mode == "twice":   # a second one-device aligner after the first was closed
    for k in range(2):
        one = _mmg.DeviceAligner(lib, idx, mo, device=0)
        for rep in range(2):
            t0 = time.time(); a = one.map_batch(hb.numpy(), offs); t1 = time.time() - t0
            print("aligner %d, rep %d: %d hits in %.2f s" % (k, rep, len(a.hits), t1), flush=True)
        one.close()
if mode in ("all", "group"):
    grp = _mmg.DeviceAligner(lib, idx, mo, devices=list(range(nd)))
    for rep in range(3):
        t0 = time.time(); b = grp.map_batch(hb.numpy(), offs); t2 = time.time() - t0
        print("%d devices, rep %d: %d hits in %.2f s; equal: %s" % (nd, rep, len(b.hits), t2, a is None or parity.compare_hits(b, a) == []), flush=True)
    grp.close()
idx.close()
