import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "mappy-rs_b200", "oracle"): sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, data_gen, parity, mm2oracle as mo
from mappy_rs import _mmg
lib = _mmg.Lib()
mode = sys.argv[1] if len(sys.argv) > 1 else "map"
if mode == "map":      # many random hits: filter, radix sort, tie replay, bulk DP, rechain
    ref, coff, names, seqs = parity.random_reference(101, [20_000_000])
    io, mopt = _mmg.IdxOpt(), _mmg.MapOpt()
    lib.check(lib.L.mmg_set_opt(None, ctypes.byref(io), ctypes.byref(mopt)))
    io.k, io.w = 13, 5
    mopt.flag = 0
    idx = _mmg.Index.build(lib, io, names, seqs)
    lib.check(lib.L.mmg_mapopt_update(ctypes.byref(mopt), idx.h))
    al = _mmg.DeviceAligner(lib, idx, mopt)
    buf, offs, _ = data_gen.make_reads(102, ref, coff, 600, 1000, 9000)
    b2, o2 = data_gen.make_sv_reads(51, ref, coff, 100)
    buf = np.concatenate([buf, b2]); offs = np.concatenate([offs, o2[1:] + offs[-1]])
    r = al.map_batch(buf, offs)
    print("map ok", len(r.hits), r.stats)
else:                  # CIGAR mode on plain + SV reads
    ref, coff, names, seqs = parity.random_reference(41, [300000, 150000], n_repeats=60, rep_min=300, rep_max=4000, rep_div=0.03)
    c = parity.Case(lib, names, seqs, cigar=True)
    buf, offs = data_gen.make_sv_reads(51, ref, coff, 150)
    r = c.aligner.map_batch(buf, offs)
    print("cigar ok", len(r.hits), len(r.cigar))
