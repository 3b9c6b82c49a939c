"""ctypes binding of the ORACLE (oracle/libmm2oracle.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never by the product.
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmm2oracle.so")

HIT_DTYPE = np.dtype([
    ("rid", "<i4"), ("rs", "<i4"), ("re", "<i4"), ("qs", "<i4"), ("qe", "<i4"),
    ("mlen", "<i4"), ("blen", "<i4"),
    ("score", "<i4"), ("score0", "<i4"), ("cnt", "<i4"), ("subsc", "<i4"), ("n_sub", "<i4"),
    ("parent", "<i4"), ("id", "<i4"),
    ("dp_score", "<i4"), ("dp_max", "<i4"), ("dp_max2", "<i4"),
    ("nm", "<i4"), ("n_ambi", "<i4"),
    ("hash", "<u4"), ("div", "<f4"),
    ("rev", "u1"), ("mapq", "u1"), ("is_primary", "u1"), ("flags", "u1"),
    ("n_cigar", "<u4"), ("cigar_off", "<u8"),
], align=True)

STAT_NAMES = ["n_bases", "n_mz", "n_seed", "n_hit", "n_anchor", "n_iter", "n_kept", "n_cell", "n_regs", "n_rechain"]


def build(force=False):
    if force or not os.path.exists(_SO):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-j8"])
    else:
        subprocess.check_call(["make", "-s", "-q", "-C", _HERE]) if False else subprocess.call(["make", "-s", "-C", _HERE, "-j8"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = ctypes.CDLL(_SO)
    vp, cp, i32, u32, u64, i64 = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int64
    L.mm2o_open.restype = vp; L.mm2o_open.argtypes = [cp, cp, i32]
    L.mm2o_build.restype = vp; L.mm2o_build.argtypes = [cp, i32, vp, vp, vp]
    L.mm2o_build_kw.restype = vp; L.mm2o_build_kw.argtypes = [cp, i32, i32, i32, vp, vp, vp]
    L.mm2o_close.argtypes = [vp]
    L.mm2o_dump_index.argtypes = [vp, cp]
    L.mm2o_set_opt_int.argtypes = [vp, cp, i64]
    L.mm2o_get_opt_int.restype = i64; L.mm2o_get_opt_int.argtypes = [vp, cp]
    L.mm2o_seq_name.restype = cp; L.mm2o_seq_name.argtypes = [vp, i32]
    L.mm2o_seq_len.restype = u32; L.mm2o_seq_len.argtypes = [vp, i32]
    L.mm2o_name2id.argtypes = [vp, cp]
    L.mm2o_getseq.argtypes = [vp, u32, u32, u32, vp]
    L.mm2o_index_entries.restype = u64; L.mm2o_index_entries.argtypes = [vp, vp, vp, u64]
    L.mm2o_sketch.restype = u64; L.mm2o_sketch.argtypes = [cp, i32, i32, i32, u32, i32, vp, vp, u64]
    L.mm2o_map_batch.restype = vp; L.mm2o_map_batch.argtypes = [vp, vp, vp, u32, i32, i32]
    for nm in ("mm2o_result_n_hits", "mm2o_result_n_cigar"):
        getattr(L, nm).restype = u64; getattr(L, nm).argtypes = [vp]
    for nm in ("mm2o_result_hit_off", "mm2o_result_hits", "mm2o_result_cigar", "mm2o_result_cs_off", "mm2o_result_cs"):
        getattr(L, nm).restype = vp; getattr(L, nm).argtypes = [vp]
    L.mm2o_result_stats.argtypes = [vp, vp]
    L.mm2o_result_free.argtypes = [vp]
    L.mm2o_set_ksw_scalar.argtypes = [i32]
    L.mm2o_trace.restype = vp; L.mm2o_trace.argtypes = [vp, cp, i32]
    L.mm2o_trace_n.restype = u64; L.mm2o_trace_n.argtypes = [vp, i32]
    L.mm2o_trace_ptr.restype = vp; L.mm2o_trace_ptr.argtypes = [vp, i32]
    L.mm2o_trace_free.argtypes = [vp]
    assert L.mm2o_sizeof_hit() == HIT_DTYPE.itemsize, (L.mm2o_sizeof_hit(), HIT_DTYPE.itemsize)
    _lib = L
    return L


def _np_from(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    nbytes = int(n) * np.dtype(dtype).itemsize
    buf = (ctypes.c_char * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=int(n)).copy()


def pack_reads(seqs):
    """list[str|bytes] -> (uint8 buffer, uint64 offsets[n+1])"""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    offs = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offs[1:] = np.cumsum([len(b) for b in bs])
    buf = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, dtype=np.uint8)
    return buf, offs


class BatchResult:
    def __init__(self, hit_off, hits, cigar, stats, cs_off=None, cs=None):
        self.hit_off, self.hits, self.cigar, self.stats = hit_off, hits, cigar, stats
        self.cs_off, self.cs = cs_off, cs

    def read_hits(self, i):
        return self.hits[int(self.hit_off[i]):int(self.hit_off[i + 1])]

    def hit_cigar(self, h):
        return self.cigar[int(h["cigar_off"]):int(h["cigar_off"]) + int(h["n_cigar"])]


class Oracle:
    """CPU restatement of minimap2 v2.26 behind mappy-rs' option plumbing
    (/root/reference/src/lib.rs:311-436)."""

    def __init__(self, fn_idx_in=None, preset=None, names=None, seqs=None, k=0, w=0, **overrides):
        L = lib()
        p = preset.encode() if preset else None
        if fn_idx_in is not None:
            with open(fn_idx_in, "rb") as fh:
                is_fasta = fh.read(4) != b"MMI\x02"
            self.h = L.mm2o_open(str(fn_idx_in).encode(), p, int(is_fasta))
        else:
            n = len(names)
            self._keep = [s if isinstance(s, bytes) else (s.tobytes() if hasattr(s, "tobytes") else s.encode()) for s in seqs]
            nm = (ctypes.c_char_p * n)(*[x.encode() for x in names])
            sq = (ctypes.c_char_p * n)(*self._keep)
            ln = np.array([len(s) for s in self._keep], dtype=np.uint32)
            self.h = L.mm2o_build_kw(p, int(k), int(w), n, nm, sq, ln.ctypes.data) if (k or w) else L.mm2o_build(p, n, nm, sq, ln.ctypes.data)
        if not self.h:
            raise RuntimeError("oracle: failed to open/build index")
        for k, v in overrides.items():
            self.set_opt(k, v)

    def close(self):
        if getattr(self, "h", None):
            lib().mm2o_close(self.h)
            self.h = None

    __del__ = close

    def set_opt(self, name, v):
        if lib().mm2o_set_opt_int(self.h, name.encode(), int(v)) != 0:
            raise KeyError(name)

    def get_opt(self, name):
        return int(lib().mm2o_get_opt_int(self.h, name.encode()))

    @property
    def k(self): return self.get_opt("k")
    @property
    def w(self): return self.get_opt("w")
    @property
    def n_seq(self): return self.get_opt("n_seq")
    @property
    def seq_names(self): return [lib().mm2o_seq_name(self.h, i).decode() for i in range(self.n_seq)]
    @property
    def seq_lens(self): return [int(lib().mm2o_seq_len(self.h, i)) for i in range(self.n_seq)]

    def dump_index(self, path):
        return lib().mm2o_dump_index(self.h, str(path).encode())

    def seq(self, name, start=0, end=0x7fffffff):
        rid = lib().mm2o_name2id(self.h, name.encode())
        if rid < 0:
            return None
        ln = self.seq_lens[rid]
        if start >= ln or start >= end:
            return None
        if end < 0 or end > ln:
            end = ln
        buf = np.zeros(end - start, dtype=np.uint8)
        lib().mm2o_getseq(self.h, rid, start, end, buf.ctypes.data)
        return bytes(b"ACGTN"[c] for c in buf).decode()

    def index_entries(self):
        n = lib().mm2o_index_entries(self.h, None, None, 0)
        mz = np.zeros(n, dtype=np.uint64); y = np.zeros(n, dtype=np.uint64)
        lib().mm2o_index_entries(self.h, mz.ctypes.data, y.ctypes.data, n)
        return mz, y

    def map_batch(self, buf, offs, n_threads=1, cs=False):
        L = lib()
        buf = np.ascontiguousarray(buf, dtype=np.uint8); offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = len(offs) - 1
        r = L.mm2o_map_batch(self.h, buf.ctypes.data, offs.ctypes.data, n, n_threads, int(cs))
        try:
            nh = L.mm2o_result_n_hits(r); nc = L.mm2o_result_n_cigar(r)
            hit_off = _np_from(L.mm2o_result_hit_off(r), n + 1, np.uint64)
            hits = _np_from(L.mm2o_result_hits(r), nh, HIT_DTYPE)
            cigar = _np_from(L.mm2o_result_cigar(r), nc, np.uint32)
            st = np.zeros(len(STAT_NAMES), dtype=np.uint64)
            L.mm2o_result_stats(r, st.ctypes.data)
            cs_off = cs_s = None
            if cs:
                cs_off = _np_from(L.mm2o_result_cs_off(r), nh + 1, np.uint64)
                cs_s = ctypes.string_at(L.mm2o_result_cs(r), int(cs_off[-1])) if nh else b""
            return BatchResult(hit_off, hits, cigar, dict(zip(STAT_NAMES, (int(x) for x in st))), cs_off, cs_s)
        finally:
            L.mm2o_result_free(r)

    def map(self, seq, cs=False):
        buf, offs = pack_reads([seq])
        return self.map_batch(buf, offs, 1, cs)

    def trace(self, seq):
        L = lib()
        s = seq.encode() if isinstance(seq, str) else bytes(seq)
        h = L.mm2o_trace(self.h, s, len(s))
        try:
            out = {}
            rec = np.dtype([("x", "<u8"), ("y", "<u8")])
            for nm, w in (("mv", 0), ("a_sorted", 1), ("a_dp", 2), ("a", 3)):
                out[nm] = _np_from(L.mm2o_trace_ptr(h, w), L.mm2o_trace_n(h, w), rec)
            for nm, w in (("u_dp", 4), ("u", 5)):
                out[nm] = _np_from(L.mm2o_trace_ptr(h, w), L.mm2o_trace_n(h, w), np.uint64)
            for nm, w in (("regs_gen", 6), ("regs_chain", 7), ("regs_final", 8)):
                out[nm] = _np_from(L.mm2o_trace_ptr(h, w), L.mm2o_trace_n(h, w), HIT_DTYPE)
            out["rechained"] = int(L.mm2o_trace_n(h, 9))
            out["rep_len"] = int(np.int32(np.uint32(L.mm2o_trace_n(h, 10) & 0xffffffff)))
            return out
        finally:
            L.mm2o_trace_free(h)


def sketch(seq, w, k, rid=0, is_hpc=0):
    s = seq.encode() if isinstance(seq, str) else bytes(seq)
    cap = len(s) + 16
    x = np.zeros(cap, dtype=np.uint64); y = np.zeros(cap, dtype=np.uint64)
    n = lib().mm2o_sketch(s, len(s), w, k, rid, is_hpc, x.ctypes.data, y.ctypes.data, cap)
    return x[:n].copy(), y[:n].copy()
