"""ctypes binding of the C ABI in include/mmg.h (libmmg.so).

`Lib()` loads the product library built by nvcc for sm_100a.  There is no CPU
fallback: if the library is missing, or no CUDA device is present, errors are
raised.  (The CPU test-suite passes an explicit path to its SIMT-emulated test
build of the same sources; the product never does.)
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(os.path.dirname(_HERE), "libmmg.so")

c_int, c_i64, c_u64, c_u32, c_vp, c_cp = ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_char_p


class IdxOpt(ctypes.Structure):
    _fields_ = [("k", ctypes.c_short), ("w", ctypes.c_short), ("flag", ctypes.c_short), ("bucket_bits", ctypes.c_short),
                ("mini_batch_size", c_i64), ("batch_size", c_u64)]


class MapOpt(ctypes.Structure):
    _fields_ = [("flag", c_i64), ("seed", c_int), ("sdust_thres", c_int), ("max_qlen", c_int),
                ("bw", c_int), ("bw_long", c_int), ("max_gap", c_int), ("max_gap_ref", c_int), ("max_frag_len", c_int),
                ("max_chain_skip", c_int), ("max_chain_iter", c_int), ("min_cnt", c_int), ("min_chain_score", c_int),
                ("chain_gap_scale", ctypes.c_float), ("chain_skip_scale", ctypes.c_float),
                ("rmq_size_cap", c_int), ("rmq_inner_dist", c_int), ("rmq_rescue_size", c_int), ("rmq_rescue_ratio", ctypes.c_float),
                ("mask_level", ctypes.c_float), ("mask_len", c_int), ("pri_ratio", ctypes.c_float), ("best_n", c_int),
                ("alt_drop", ctypes.c_float),
                ("a", c_int), ("b", c_int), ("q", c_int), ("e", c_int), ("q2", c_int), ("e2", c_int),
                ("transition", c_int), ("sc_ambi", c_int), ("noncan", c_int), ("junc_bonus", c_int),
                ("zdrop", c_int), ("zdrop_inv", c_int), ("end_bonus", c_int), ("min_dp_max", c_int), ("min_ksw_len", c_int),
                ("anchor_ext_len", c_int), ("anchor_ext_shift", c_int), ("max_clip_ratio", ctypes.c_float),
                ("rank_min_len", c_int), ("rank_frac", ctypes.c_float), ("pe_ori", c_int), ("pe_bonus", c_int),
                ("mid_occ_frac", ctypes.c_float), ("q_occ_frac", ctypes.c_float),
                ("min_mid_occ", ctypes.c_int32), ("max_mid_occ", ctypes.c_int32), ("mid_occ", ctypes.c_int32),
                ("max_occ", ctypes.c_int32), ("max_max_occ", ctypes.c_int32), ("occ_dist", ctypes.c_int32),
                ("mini_batch_size", c_i64), ("max_sw_mat", c_i64), ("cap_kalloc", c_i64), ("split_prefix", c_cp)]


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

STAT_NAMES = ["n_bases", "n_mz", "n_seed", "n_hit", "n_anchor", "n_iter", "n_kept", "n_cell", "n_regs", "n_rechain", "n_dropped", "n_cell_fill"]
N_STAGES = 12

EXPORTS = ["mmg_host_alloc", "mmg_host_free", "mmg_set_opt", "mmg_mapopt_update", "mmg_index_open", "mmg_index_build", "mmg_index_build_on", "mmg_debug_int32_peak", "mmg_index_dump", "mmg_index_destroy",
           "mmg_index_info", "mmg_index_seq_name", "mmg_index_seq_len", "mmg_index_name2id", "mmg_index_getseq",
           "mmg_index_entries", "mmg_aligner_create", "mmg_aligner_create_multi", "mmg_aligner_destroy", "mmg_aligner_set", "mmg_map_batch",
           "mmg_batch_upload", "mmg_batch_run", "mmg_batch_fetch", "mmg_batch_n_reads", "mmg_batch_n_hits",
           "mmg_batch_hit_off", "mmg_batch_hits", "mmg_batch_n_cigar", "mmg_batch_cigar", "mmg_gen_cs",
           "mmg_gen_md", "mmg_gen_tags", "mmg_debug_logf", "mmg_batch_destroy", "mmg_batch_stats", "mmg_stage_times", "mmg_stage_name",
           "mmg_last_run_ms", "mmg_debug_dump", "mmg_last_error", "mmg_version", "mmg_sizeof_hit",
           "mmg_submit", "mmg_flush", "mmg_next", "mmg_result_release"]


class Result(ctypes.Structure):
    _fields_ = [("read_id", c_u64), ("n_hits", c_u32), ("hits", c_vp), ("cigar", c_vp), ("owner", c_vp)]


class MmgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libmmg error %d: %s" % (code, msg))
        self.code = code


class Lib:
    def __init__(self, path=None):
        path = path or DEFAULT_LIB
        if not os.path.exists(path):
            raise ImportError("libmmg.so not found at %s: build it with `make -C mappy-rs_b200` (nvcc, sm_100a). "
                              "There is no CPU fallback." % path)
        L = self.L = ctypes.CDLL(path)
        self.path = path
        P = ctypes.POINTER
        L.mmg_set_opt.argtypes = [c_cp, P(IdxOpt), P(MapOpt)]
        L.mmg_mapopt_update.argtypes = [P(MapOpt), c_vp]
        L.mmg_index_open.argtypes = [c_cp, P(IdxOpt), c_int, P(c_vp)]
        L.mmg_index_build.argtypes = [P(IdxOpt), c_int, c_vp, c_vp, c_vp, c_int, P(c_vp)]
        L.mmg_index_build_on.argtypes = [P(IdxOpt), c_int, c_vp, c_vp, c_vp, c_int, c_int, P(c_vp)]
        L.mmg_debug_int32_peak.argtypes = [c_int, P(ctypes.c_double)]
        L.mmg_host_alloc.argtypes = [ctypes.c_size_t, P(c_vp)]
        L.mmg_host_free.argtypes = [c_vp]
        L.mmg_index_dump.argtypes = [c_vp, c_cp]
        L.mmg_index_destroy.argtypes = [c_vp]
        L.mmg_index_info.argtypes = [c_vp, c_vp]
        L.mmg_index_seq_name.restype = c_cp; L.mmg_index_seq_name.argtypes = [c_vp, c_u32]
        L.mmg_index_seq_len.restype = c_u32; L.mmg_index_seq_len.argtypes = [c_vp, c_u32]
        L.mmg_index_name2id.argtypes = [c_vp, c_cp]
        L.mmg_index_getseq.argtypes = [c_vp, c_u32, c_u32, c_u32, c_vp]
        L.mmg_index_entries.restype = c_u64; L.mmg_index_entries.argtypes = [c_vp, c_vp, c_vp, c_u64]
        L.mmg_aligner_create.argtypes = [c_vp, P(MapOpt), c_int, P(c_vp)]
        L.mmg_aligner_create_multi.argtypes = [c_vp, P(MapOpt), c_vp, c_int, P(c_vp)]
        L.mmg_aligner_destroy.argtypes = [c_vp]
        L.mmg_aligner_set.argtypes = [c_vp, c_cp, c_i64]
        L.mmg_map_batch.argtypes = [c_vp, c_vp, c_vp, c_u32, P(c_vp)]
        L.mmg_batch_upload.argtypes = [c_vp, c_vp, c_vp, c_u32, P(c_vp)]
        L.mmg_batch_run.argtypes = [c_vp, c_vp]
        L.mmg_batch_fetch.argtypes = [c_vp, c_vp]
        L.mmg_batch_n_reads.restype = c_u32; L.mmg_batch_n_reads.argtypes = [c_vp]
        for nm in ("mmg_batch_n_hits", "mmg_batch_n_cigar"):
            getattr(L, nm).restype = c_u64; getattr(L, nm).argtypes = [c_vp]
        for nm in ("mmg_batch_hit_off", "mmg_batch_hits", "mmg_batch_cigar"):
            getattr(L, nm).restype = c_vp; getattr(L, nm).argtypes = [c_vp]
        L.mmg_gen_cs.argtypes = [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, ctypes.c_size_t]
        L.mmg_gen_md.argtypes = [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, ctypes.c_size_t]
        L.mmg_gen_tags.restype = c_i64
        L.mmg_gen_tags.argtypes = [c_vp, c_vp, c_vp, c_u32, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_u64, c_vp]
        L.mmg_batch_destroy.argtypes = [c_vp]
        L.mmg_batch_stats.argtypes = [c_vp, c_vp]
        L.mmg_stage_times.argtypes = [c_vp, c_vp, c_vp]
        L.mmg_stage_name.restype = c_cp; L.mmg_stage_name.argtypes = [c_int]
        L.mmg_last_run_ms.restype = ctypes.c_double; L.mmg_last_run_ms.argtypes = [c_vp]
        L.mmg_debug_dump.restype = c_i64; L.mmg_debug_dump.argtypes = [c_vp, c_vp, c_int, c_vp, c_vp, c_u64, c_vp]
        L.mmg_submit.argtypes = [c_vp, c_vp, c_vp, c_u32, c_u64]
        L.mmg_flush.argtypes = [c_vp]
        L.mmg_next.argtypes = [c_vp, P(Result), c_int]
        L.mmg_result_release.argtypes = [c_vp, P(Result)]
        L.mmg_result_release.restype = None
        L.mmg_last_error.restype = c_cp
        L.mmg_version.restype = c_cp
        assert L.mmg_sizeof_hit() == HIT_DTYPE.itemsize, (L.mmg_sizeof_hit(), HIT_DTYPE.itemsize)

    def check(self, rc):
        if rc < 0:
            raise MmgError(rc, self.L.mmg_last_error().decode(errors="replace"))
        return rc

    def version(self):
        return self.L.mmg_version().decode()


def np_from(ptr, n, dtype, copy=True):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    nbytes = int(n) * np.dtype(dtype).itemsize
    buf = (ctypes.c_char * nbytes).from_address(ptr)
    a = np.frombuffer(buf, dtype=dtype, count=int(n))
    return a.copy() if copy else a


class PinnedBuffer:
    """Grow-only page-locked byte buffer (mmg_host_alloc) that the batch host assembles its reads in."""

    def __init__(self, lib):
        self.lib, self.ptr, self.cap, self.arr = lib, None, 0, None

    def view(self, n):
        if n > self.cap:
            self.close()
            cap = max(1 << 20, int(n * 1.5))
            p = c_vp()
            self.lib.check(self.lib.L.mmg_host_alloc(cap, ctypes.byref(p)))
            self.ptr, self.cap = p, cap
            self.arr = np.frombuffer((ctypes.c_char * cap).from_address(p.value), dtype=np.uint8)
        return self.arr[:max(n, 1)]

    def close(self):
        if self.ptr is not None:
            self.arr = None
            self.lib.L.mmg_host_free(self.ptr)
            self.ptr, self.cap = None, 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Index:
    def __init__(self, lib, handle):
        self.lib, self.h = lib, handle
        info = np.zeros(5, dtype=np.int32)
        lib.L.mmg_index_info(handle, info.ctypes.data)
        self.k, self.w, self.b, self.flag, self.n_seq = (int(x) for x in info)

    @classmethod
    def open(cls, lib, path, io, n_threads=3):
        h = c_vp()
        lib.check(lib.L.mmg_index_open(str(path).encode(), ctypes.byref(io), n_threads, ctypes.byref(h)))
        return cls(lib, h)

    @classmethod
    def build(cls, lib, io, names, seqs, n_threads=8, device=0):
        n = len(names)
        keep = [s if isinstance(s, bytes) else (s.tobytes() if hasattr(s, "tobytes") else s.encode()) for s in seqs]
        nm = (ctypes.c_char_p * n)(*[x.encode() for x in names])
        sq = (ctypes.c_char_p * n)(*keep)
        ln = np.array([len(s) for s in keep], dtype=np.uint32)
        h = c_vp()
        lib.check(lib.L.mmg_index_build_on(ctypes.byref(io), n, nm, sq, ln.ctypes.data, n_threads, device, ctypes.byref(h)))
        return cls(lib, h)

    def close(self):
        if self.h:
            self.lib.L.mmg_index_destroy(self.h)
            self.h = None

    def seq_name(self, i): return self.lib.L.mmg_index_seq_name(self.h, i).decode()
    def seq_len(self, i): return int(self.lib.L.mmg_index_seq_len(self.h, i))
    def name2id(self, name): return self.lib.L.mmg_index_name2id(self.h, name.encode())

    def getseq(self, rid, st, en):
        buf = np.zeros(max(en - st, 0), dtype=np.uint8)
        n = self.lib.L.mmg_index_getseq(self.h, rid, st, en, buf.ctypes.data)
        return None if n < 0 else buf[:n]

    def entries(self):
        n = self.lib.L.mmg_index_entries(self.h, None, None, 0)
        mz = np.zeros(n, dtype=np.uint64); y = np.zeros(n, dtype=np.uint64)
        self.lib.L.mmg_index_entries(self.h, mz.ctypes.data, y.ctypes.data, n)
        return mz, y

    def dump(self, path):
        self.lib.check(self.lib.L.mmg_index_dump(self.h, str(path).encode()))


class Batch:
    """Results of one mmg_map_batch call as numpy arrays: copies by default; with own=True the arrays are
    views of the library's result memory and this object keeps the batch handle alive (freed on close/GC)."""

    def __init__(self, lib, h, n_reads, own=False):
        L = lib.L
        self._lib, self._owned = lib, h if own else None
        nh = L.mmg_batch_n_hits(h); nc = L.mmg_batch_n_cigar(h)
        self.hit_off = np_from(L.mmg_batch_hit_off(h), n_reads + 1, np.uint64, not own)
        self.hits = np_from(L.mmg_batch_hits(h), nh, HIT_DTYPE, not own)
        self.cigar = np_from(L.mmg_batch_cigar(h), nc, np.uint32, not own)
        st = np.zeros(len(STAT_NAMES), dtype=np.uint64)
        L.mmg_batch_stats(h, st.ctypes.data)
        self.stats = dict(zip(STAT_NAMES, (int(x) for x in st)))

    def close(self):
        if self._owned is not None:
            self.hit_off = self.hits = self.cigar = None
            self._lib.L.mmg_batch_destroy(self._owned)
            self._owned = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def read_hits(self, i):
        return self.hits[int(self.hit_off[i]):int(self.hit_off[i + 1])]

    def hit_cigar(self, h):
        return self.cigar[int(h["cigar_off"]):int(h["cigar_off"]) + int(h["n_cigar"])]


class DeviceAligner:
    """One aligner on one GPU (`device`), or - with `devices=[...]` - one aligner that replicates the index on several
    GPUs of the box, shards every batch by bases over them and gathers the results in read order."""

    def __init__(self, lib, index, mo, device=0, devices=None):
        self.lib, self.index = lib, index
        h = c_vp()
        if devices is not None and len(devices) > 1:
            dv = np.ascontiguousarray(devices, dtype=np.int32)
            lib.check(lib.L.mmg_aligner_create_multi(index.h, ctypes.byref(mo), dv.ctypes.data, len(dv), ctypes.byref(h)))
        else:
            if devices is not None and len(devices) == 1:
                device = int(devices[0])
            lib.check(lib.L.mmg_aligner_create(index.h, ctypes.byref(mo), device, ctypes.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.L.mmg_aligner_destroy(self.h)
            self.h = None

    def set(self, key, v):
        self.lib.check(self.lib.L.mmg_aligner_set(self.h, key.encode(), int(v)))

    def map_batch(self, buf, offs, keep_handle=False, zero_copy=False):
        buf = np.ascontiguousarray(buf, dtype=np.uint8); offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = len(offs) - 1
        b = c_vp()
        self.lib.check(self.lib.L.mmg_map_batch(self.h, buf.ctypes.data, offs.ctypes.data, n, ctypes.byref(b)))
        if zero_copy:
            return Batch(self.lib, b, n, own=True)
        res = Batch(self.lib, b, n)
        if keep_handle:
            res.handle = b
        else:
            self.lib.L.mmg_batch_destroy(b)
        return res

    def upload(self, buf, offs):
        b = c_vp()
        self.lib.check(self.lib.L.mmg_batch_upload(self.h, buf.ctypes.data, offs.ctypes.data, len(offs) - 1, ctypes.byref(b)))
        return b

    def run(self, b): self.lib.check(self.lib.L.mmg_batch_run(self.h, b))
    def fetch(self, b): self.lib.check(self.lib.L.mmg_batch_fetch(self.h, b))
    def free(self, b): self.lib.L.mmg_batch_destroy(b)

    def last_run_ms(self): return float(self.lib.L.mmg_last_run_ms(self.h))

    def int32_peak(self, device=0):
        """Measured INT32 issue peak (Gop/s) of the device: the denominator of the integer rooflines."""
        v = ctypes.c_double(0)
        self.lib.check(self.lib.L.mmg_debug_int32_peak(device, ctypes.byref(v)))
        return float(v.value)

    def stage_times(self):
        ms = np.zeros(N_STAGES, dtype=np.float64); ln = np.zeros(N_STAGES, dtype=np.uint64)
        self.lib.L.mmg_stage_times(self.h, ms.ctypes.data, ln.ctypes.data)
        names = [self.lib.L.mmg_stage_name(i).decode() for i in range(N_STAGES)]
        return {n: (float(m), int(l)) for n, m, l in zip(names, ms, ln)}

    def debug_dump(self, b, which, cap, n_reads):
        x = np.zeros(cap, dtype=np.uint64); y = np.zeros(cap, dtype=np.uint64); off = np.zeros(n_reads + 1, dtype=np.uint64)
        n = self.lib.L.mmg_debug_dump(self.h, b, which, x.ctypes.data, y.ctypes.data, cap, off.ctypes.data)
        self.lib.check(int(n))
        return x[:n], y[:n], off


def gen_tags(lib, index, buf, offs, res, which, n_threads=0):
    """cs (which=0) or MD (which=1) strings of every hit of a Batch: list[bytes|None]."""
    n_reads, n_hits = len(offs) - 1, len(res.hits)
    if n_hits == 0:
        return []
    if n_threads <= 0:
        n_threads = min(32, os.cpu_count() or 4)
    buf = np.ascontiguousarray(buf, dtype=np.uint8); offs = np.ascontiguousarray(offs, dtype=np.uint64)
    hit_off = np.ascontiguousarray(res.hit_off, dtype=np.uint64)
    hits = np.ascontiguousarray(res.hits); cig = np.ascontiguousarray(res.cigar, dtype=np.uint32)
    so = np.zeros(n_hits + 1, dtype=np.uint64)
    args = [index.h, buf.ctypes.data, offs.ctypes.data, n_reads, hit_off.ctypes.data, hits.ctypes.data, cig.ctypes.data if len(cig) else None, which, n_threads]
    # one call in the common case: a tag is rarely longer than its alignment block; a second call with the exact size otherwise
    cap = int(hits["blen"].astype(np.int64).sum()) + 64 * n_hits + 1024
    out = np.empty(cap, dtype=np.uint8)
    tot = lib.check(int(lib.L.mmg_gen_tags(*args, out.ctypes.data, cap, so.ctypes.data)))
    if tot > cap:
        out = np.empty(tot, dtype=np.uint8)
        lib.check(int(lib.L.mmg_gen_tags(*args, out.ctypes.data, tot, so.ctypes.data)))
    raw = out[:tot].tobytes()
    has = ((res.hits["flags"] & 32) != 0).tolist()
    sl = so.tolist()
    return [raw[sl[i]:sl[i + 1]] if has[i] else None for i in range(n_hits)]
