"""Differential harness: device library (real GPU or emulated test build) vs the oracle.

Both sides get the same index contents, the same options and the same reads;
hits are compared field by field (bit-exact), and on request the intermediate
stages (minimizers, sorted anchors, chained anchors, chains) are compared too so
that a mismatch is attributed to a stage.
"""
import ctypes

import numpy as np

import data_gen
import mm2oracle as mo
from mappy_rs import _mmg

MAPQ_FIELDS = _mmg.HIT_DTYPE.names


class Case:
    """One reference + option set, open on both sides."""

    def __init__(self, lib, names, seqs, preset=None, cigar=False, overrides=None, fasta=None, mmi=None):
        self.lib = lib
        overrides = dict(overrides or {})
        self.io = _mmg.IdxOpt()
        self.mopt = _mmg.MapOpt()
        lib.check(lib.L.mmg_set_opt(None, ctypes.byref(self.io), ctypes.byref(self.mopt)))
        if preset:
            lib.check(lib.L.mmg_set_opt(preset.encode(), ctypes.byref(self.io), ctypes.byref(self.mopt)))
        if mmi is not None:
            self.oracle = mo.Oracle(mmi, preset=preset)
            self.index = _mmg.Index.open(lib, mmi, self.io)
        elif fasta is not None:
            self.oracle = mo.Oracle(fasta, preset=preset)
            self.index = _mmg.Index.open(lib, fasta, self.io)
        else:
            self.oracle = mo.Oracle(names=names, seqs=seqs, preset=preset)
            self.index = _mmg.Index.build(lib, self.io, names, seqs)
        flag = 4 if cigar else 0
        self.mopt.flag = flag
        self.oracle.set_opt("flag", flag)
        for k, v in overrides.items():
            setattr(self.mopt, k, v)
            self.oracle.set_opt(k, v)
        lib.check(lib.L.mmg_mapopt_update(ctypes.byref(self.mopt), self.index.h))
        assert self.mopt.mid_occ == self.oracle.get_opt("mid_occ")
        self.aligner = _mmg.DeviceAligner(lib, self.index, self.mopt)

    def close(self):
        self.aligner.close()
        self.index.close()
        self.oracle.close()


def compare_hits(dev, ora, limit=5):
    """Returns a list of human-readable differences (empty = bit-exact)."""
    diffs = []
    if not np.array_equal(dev.hit_off, ora.hit_off):
        bad = np.nonzero(np.diff(dev.hit_off.astype(np.int64)) != np.diff(ora.hit_off.astype(np.int64)))[0]
        diffs.append("hit counts differ for %d reads, first: %s" % (len(bad), bad[:limit].tolist()))
        return diffs
    for f in MAPQ_FIELDS:
        if f == "cigar_off":
            continue
        a, b = dev.hits[f], ora.hits[f]
        if f == "div":
            a, b = a.view(np.uint32), b.view(np.uint32)
        if not np.array_equal(a, b):
            d = np.nonzero(a != b)[0]
            reads = np.searchsorted(ora.hit_off, d[:limit], side="right") - 1
            diffs.append("field %s differs in %d hits; first hits %s (reads %s): dev %s oracle %s" %
                         (f, len(d), d[:limit].tolist(), reads.tolist(), dev.hits[f][d[:limit]].tolist(), ora.hits[f][d[:limit]].tolist()))
    # CIGARs
    for i in range(len(ora.hits)):
        if ora.hits["n_cigar"][i] and not np.array_equal(dev.hit_cigar(dev.hits[i]), ora.hit_cigar(ora.hits[i])):
            diffs.append("cigar differs at hit %d" % i)
            if len(diffs) > limit:
                break
    return diffs


def compare_stats(dev, ora, keys=("n_bases", "n_mz", "n_seed", "n_hit", "n_anchor", "n_iter", "n_kept", "n_regs")):
    return ["stat %s: dev %d oracle %d" % (k, dev.stats[k], ora.stats[k]) for k in keys if dev.stats[k] != ora.stats[k]]


def compare_stages(case, buf, offs, max_reads=200):
    """Stage-by-stage comparison against oracle traces (single-chunk batches only)."""
    n = len(offs) - 1
    res = case.aligner.map_batch(buf, offs, keep_handle=True)
    diffs = []
    try:
        cap = int(offs[-1]) + 16
        mx, my, moff = case.aligner.debug_dump(res.handle, 0, cap, n)
        acap = int(res.stats["n_anchor"]) + 16
        sx, sy, soff = case.aligner.debug_dump(res.handle, 1, acap, n)
        cx, cy, coff = case.aligner.debug_dump(res.handle, 2, acap, n)
        ux, _, uoff = case.aligner.debug_dump(res.handle, 3, acap, n)
        for i in range(min(n, max_reads)):
            s = buf[int(offs[i]):int(offs[i + 1])].tobytes()
            tr = case.oracle.trace(s)
            sl = slice(int(moff[i]), int(moff[i + 1]))
            if not (np.array_equal(mx[sl], tr["mv"]["x"]) and np.array_equal(my[sl], tr["mv"]["y"])):
                diffs.append("read %d: minimizers differ (%d vs %d)" % (i, sl.stop - sl.start, len(tr["mv"])))
                continue
            # sorted anchors are overwritten by the chained anchors on the device; compare the final chains
            sl = slice(int(coff[i]), int(coff[i + 1]))
            if not (np.array_equal(cx[sl], tr["a"]["x"]) and np.array_equal(cy[sl], tr["a"]["y"])):
                diffs.append("read %d: chained anchors differ (%d vs %d, rechained=%d)" % (i, sl.stop - sl.start, len(tr["a"]), tr["rechained"]))
                continue
            sl = slice(int(uoff[i]), int(uoff[i + 1]))
            if not np.array_equal(ux[sl], tr["u"]):
                diffs.append("read %d: chains u[] differ" % i)
    finally:
        case.aligner.free(res.handle)
    return res, diffs


def random_reference(seed, contig_lens, n_repeats=0, **kw):
    ref, coff, names = data_gen.make_reference(seed, contig_lens, n_repeats=n_repeats, **kw)
    seqs = [ref[int(coff[i]):int(coff[i + 1])].tobytes() for i in range(len(names))]
    return ref, coff, names, seqs


def logf_mismatches(lib, n):
    """dev_logf (regs.cu) vs the host libm logf over the integers and ratios mm_set_mapq can produce."""
    libm = ctypes.CDLL("libm.so.6")
    libm.logf.restype = ctypes.c_float
    libm.logf.argtypes = [ctypes.c_float]
    rs = np.random.RandomState(1)
    x = np.concatenate([np.arange(1, n // 2 + 1, dtype=np.float32),
                        (rs.randint(1, 100000, n // 2) / rs.randint(1, 3, n // 2)).astype(np.float32)])
    out = np.zeros_like(x)
    lib.L.mmg_debug_logf.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
    lib.check(lib.L.mmg_debug_logf(x.ctypes.data, out.ctypes.data, len(x)))
    want = np.array([libm.logf(float(v)) for v in x], dtype=np.float32)
    return int(np.count_nonzero(out.view(np.uint32) != want.view(np.uint32)))


def compare_tags(case, buf, offs, dev, which, limit=5):
    """mmg_gen_tags (cs short form / MD, host marshalling of the product) against the oracle's mm_gen_cs / mm_gen_MD
    (oracle/mm2o_align.cpp) for every hit of the batch.  which: 0 = cs (no_iden = 1, what crate minimap2 passes), 1 = MD."""
    ours = _mmg.gen_tags(case.lib, case.index, buf, offs, dev, which)
    ora = case.oracle.map_batch(buf, offs, 8, cs=2 if which else 1)
    diffs = []
    if len(ours) != len(ora.hits):
        return ["tag count %d vs %d hits" % (len(ours), len(ora.hits))]
    for i in range(len(ours)):
        want = ora.cs[int(ora.cs_off[i]):int(ora.cs_off[i + 1])] if ora.hits["n_cigar"][i] else None
        if ours[i] != want:
            diffs.append("hit %d: %s differs: %r vs oracle %r" % (i, "MD" if which else "cs", (ours[i] or b"")[:60], (want or b"")[:60]))
            if len(diffs) >= limit:
                break
    return diffs
