import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "mappy-rs_b200", "oracle"): sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, data_gen
from mappy_rs import _mmg
lib = _mmg.Lib()
ref, coff, names = data_gen.make_reference(3, data_gen.config2_contig_lens())
io, mo = _mmg.IdxOpt(), _mmg.MapOpt()
lib.check(lib.L.mmg_set_opt(None, ctypes.byref(io), ctypes.byref(mo))); mo.flag = 0
idx = _mmg.Index.build(lib, io, names, [ref[int(coff[i]):int(coff[i + 1])].tobytes() for i in range(len(names))])
lib.check(lib.L.mmg_mapopt_update(ctypes.byref(mo), idx.h))
al = _mmg.DeviceAligner(lib, idx, mo)
n = 60000
buf, offs, truth = data_gen.make_reads(4, ref, coff, n, 1000, 10000, p_sub=0.03, p_ins=0.02, p_del=0.03)
runs = [al.map_batch(buf, offs) for _ in range(3)]
a = runs[0]
for k, b in enumerate(runs[1:]):
    print("run", k + 1, "hit_off equal", np.array_equal(a.hit_off, b.hit_off), "stats", {s: (a.stats[s], b.stats[s]) for s in a.stats if a.stats[s] != b.stats[s]})
    for f in _mmg.HIT_DTYPE.names:
        d = np.nonzero(a.hits[f] != b.hits[f])[0]
        if len(d):
            reads = np.searchsorted(a.hit_off, d[:5], side="right") - 1
            print("  field", f, "differs in", len(d), "hits; first", d[:5].tolist(), "reads", reads.tolist(), a.hits[f][d[:5]].tolist(), b.hits[f][d[:5]].tolist(), "read lens", [int(offs[r+1]-offs[r]) for r in reads])
    raw_a = a.hits.view(np.uint8).reshape(len(a.hits), -1); raw_b = b.hits.view(np.uint8).reshape(len(b.hits), -1)
    cols = np.nonzero((raw_a != raw_b).any(axis=0))[0]
    print("  differing byte columns", cols.tolist())
