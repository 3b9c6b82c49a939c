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
n = 4000
buf, offs, truth = data_gen.make_reads(4, ref, coff, n, 1000, 10000, p_sub=0.03, p_ins=0.02, p_del=0.03)
for flt in (1, 0):
    al.set("anchor_filter", flt)
    r = al.map_batch(buf, offs, keep_handle=True)
    cap = int(r.stats["n_anchor"]) + 16
    _, _, off1 = al.debug_dump(r.handle, 1, cap, n)   # anchors that reached the sort
    _, _, off2 = al.debug_dump(r.handle, 2, cap, n)   # chained anchors
    al.free(r.handle)
    k1, k2 = np.diff(off1.astype(np.int64)), np.diff(off2.astype(np.int64))
    ln = np.diff(offs.astype(np.int64))
    print("filter", flt, "sorted/read", k1.mean(), "chained/read", k2.mean(), "dropped", r.stats["n_dropped"] / n)
    if flt:
        extra = k1 - k2
        for lo, hi in ((1000, 3000), (3000, 6000), (6000, 8000), (8000, 10001)):
            m = (ln >= lo) & (ln < hi)
            print("  len", lo, hi, "reads", m.sum(), "sorted", k1[m].mean(), "chained", k2[m].mean(), "extra", extra[m].mean(), "extra median", np.median(extra[m]), "max", extra[m].max())
        big = np.argsort(-extra)[:8]
        print("  largest extras", [(int(ln[i]), int(k1[i]), int(k2[i])) for i in big])
        print("  reads with extra > 1000:", int((extra > 1000).sum()), "their share of all extra", extra[extra > 1000].sum() / extra.sum())
al.set("anchor_filter", 1)
r = al.map_batch(buf, offs, keep_handle=True)
cap = int(r.stats["n_anchor"]) + 16
mx, my, moff = al.debug_dump(r.handle, 0, int(offs[-1]) + 16, n)
_, _, off1 = al.debug_dump(r.handle, 1, cap, n)
_, _, off2 = al.debug_dump(r.handle, 2, cap, n)
al.free(r.handle)
k1, k2 = np.diff(off1.astype(np.int64)), np.diff(off2.astype(np.int64))
unf = np.nonzero(k1 - k2 > 1000)[0]
print("unfiltered-looking reads", len(unf))
for i in unf[:10]:
    h = mx[int(moff[i]):int(moff[i + 1])] >> np.uint64(8)
    u, cnt = np.unique(h, return_counts=True)
    print("  read", i, "len", int(offs[i+1]-offs[i]), "n_mz", len(h), "distinct", len(u), "max mult", cnt.max(), "sorted", int(k1[i]))
nm = np.diff(moff.astype(np.int64))
dup = np.array([len(np.unique(mx[int(moff[i]):int(moff[i+1])] >> np.uint64(8))) < nm[i] for i in range(n)])
print("reads with a repeated minimizer hash:", dup.sum(), "of", n, "; among unfiltered:", dup[unf].sum())
