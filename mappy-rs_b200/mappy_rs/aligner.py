"""Host-side mirror of the reference's Python module `mappy_rs` (src/lib.rs), above the C ABI.

The reference is a Rust/PyO3 extension; no Rust toolchain exists in this image,
so the operator interface is mirrored here with the same names, argument
meaning, defaults and error behaviour, and every data-path call goes through
libmmg.so (include/mmg.h).  Citations are to /root/reference/src/lib.rs.

  Aligner(...)               lib.rs:311-436   constructor / option plumbing
  Aligner.map                lib.rs:472-514   one read, blocking
  Aligner.enable_threading   lib.rs:541-636   (worker pool -> device batch pipeline)
  Aligner.map_batch          lib.rs:640-648, 771-906  streaming iterator of (mappings, dict)
  Aligner.seq / seq_names / k / w / n_seq / __bool__   lib.rs:439-470, 651-670
  Mapping                    lib.rs:106-285
There is no CPU mapping path: without libmmg.so or a CUDA device construction fails.
"""
import collections
import ctypes
import threading
import time

import numpy as np

from . import _mmg

_utf8_and_size = ctypes.pythonapi.PyUnicode_AsUTF8AndSize     # a str's UTF-8 bytes in place (ASCII: the str's own buffer)
_utf8_and_size.restype = ctypes.c_void_p
_utf8_and_size.argtypes = [ctypes.py_object, ctypes.POINTER(ctypes.c_ssize_t)]
_STR_ONLY = frozenset((str,))
_WORK_QUEUE_CAP = 50000   # lib.rs:429-430
_DRAIN_BASES = 32 << 20   # the worker starts a device batch as soon as this many bases are queued ...
_IDLE_S = 0.02            # ... or the producer has pushed nothing for this long (a slow generator still streams)
_CIGAR_OPS = "MIDNSHP=X"


class Mapping:
    """Result of an alignment (lib.rs:106-154).  `cigar` is a list of (length, op) tuples."""
    __slots__ = ("query_start", "query_end", "_strand", "target_name", "target_len", "target_start", "target_end",
                 "match_len", "block_len", "mapq", "is_primary", "_cigar", "NM", "_MD", "_cs")
    _FIELDS = ("query_start", "query_end", "_strand", "target_name", "target_len", "target_start", "target_end",
               "match_len", "block_len", "mapq", "is_primary", "cigar", "NM", "MD", "cs")

    def __init__(self, query_start, query_end, strand, target_name, target_len, target_start, target_end,
                 match_len, block_len, mapq, is_primary, cigar, NM, MD, cs):
        self.query_start, self.query_end, self._strand = query_start, query_end, strand
        self.target_name, self.target_len, self.target_start, self.target_end = target_name, target_len, target_start, target_end
        self.match_len, self.block_len, self.mapq, self.is_primary = match_len, block_len, mapq, is_primary
        # cigar may arrive as a packed uint32 array (len << 4 | op), MD / cs as bytes: they become the list of (len, op)
        # tuples / str of the reference on first access (building ~10^3 tuples per long read is most of the host time)
        self._cigar, self.NM, self._MD, self._cs = cigar, NM, MD, cs

    @property
    def cigar(self):
        c = self._cigar
        if not isinstance(c, list):
            c = self._cigar = list(zip((c >> 4).tolist(), (c & 0xf).tolist()))
        return c

    @property
    def MD(self):
        if isinstance(self._MD, bytes):
            self._MD = self._MD.decode()
        return self._MD

    @property
    def cs(self):
        if isinstance(self._cs, bytes):
            self._cs = self._cs.decode()
        return self._cs

    # mappy-style aliases (lib.rs:196-284)
    ctg = property(lambda s: s.target_name)
    ctg_len = property(lambda s: s.target_len)
    r_st = property(lambda s: s.target_start)
    r_en = property(lambda s: s.target_end)
    q_st = property(lambda s: s.query_start)
    q_en = property(lambda s: s.query_end)
    strand = property(lambda s: s._strand)          # +1 / -1 (lib.rs:231-237)
    blen = property(lambda s: s.block_len)
    mlen = property(lambda s: s.match_len)

    @property
    def cigar_str(self):
        out = []
        for n, op in self.cigar:
            if not 0 <= op <= 8:
                raise ValueError("Invalid CIGAR code `{op}`")   # lib.rs:269
            out.append("%d%s" % (n, _CIGAR_OPS[op]))
        return "".join(out)

    def __str__(self):   # lib.rs:159-179: PAF without query name / length
        return "\t".join(str(x) for x in (self.query_start, self.query_end, "+" if self._strand > 0 else "-", self.target_name,
                                            self.target_len, self.target_start, self.target_end, self.match_len, self.block_len,
                                            self.mapq, "tp:A:P" if self.is_primary else "tp:A:S", "cg:Z:" + self.cigar_str))

    def __repr__(self):  # lib.rs:186-188 ({self:#?})
        f = [("query_start", self.query_start), ("query_end", self.query_end), ("strand", "Forward" if self._strand > 0 else "Reverse"),
             ("target_name", '"%s"' % self.target_name), ("target_len", self.target_len), ("target_start", self.target_start),
             ("target_end", self.target_end), ("match_len", self.match_len), ("block_len", self.block_len), ("mapq", self.mapq),
             ("is_primary", "true" if self.is_primary else "false"), ("cigar", self.cigar), ("NM", self.NM),
             ("MD", "None" if self.MD is None else 'Some("%s")' % self.MD), ("cs", "None" if self.cs is None else 'Some("%s")' % self.cs)]
        return "Mapping {\n" + "".join("    %s: %s,\n" % kv for kv in f) + "}"

    def __eq__(self, o):
        return isinstance(o, Mapping) and all(getattr(self, k) == getattr(o, k) for k in self._FIELDS)


def _mappings_of_batch(res, names, lens, cs_list, md_list, n_reads):
    """All hits of a Batch -> per-read lists of Mapping (field mapping of crate minimap2 Aligner::map, lib.rs:493-509).
    The record arrays are converted column by column (one `tolist()` per field) instead of element by element: at
    device speed the per-hit Python work is the bottleneck of this host layer."""
    hits, cig = res.hits, res.cigar
    n = len(hits)
    col = {f: hits[f].tolist() for f in ("qs", "qe", "rev", "rid", "rs", "re", "mlen", "blen", "mapq", "is_primary", "nm", "cigar_off", "n_cigar")}
    out_hits = []
    for i in range(n):
        c0, rid = col["cigar_off"][i], col["rid"][i]
        c1 = c0 + col["n_cigar"][i]
        md = md_list[i] if md_list is not None else None
        cs = cs_list[i] if cs_list is not None else None
        out_hits.append(Mapping(col["qs"][i], col["qe"][i], -1 if col["rev"][i] else 1, names[rid], lens[rid], col["rs"][i], col["re"][i],
                                col["mlen"][i], col["blen"][i], col["mapq"][i], bool(col["is_primary"][i]),
                                cig[c0:c1], col["nm"][i], md, cs))
    ho = res.hit_off.tolist()
    return [out_hits[ho[i]:ho[i + 1]] for i in range(n_reads)]


class BatchMappings:
    """What `Aligner.map_arrays` returns: the hits of one device batch, kept as the library returned them - `hits`
    (structured array of mmg_hit_t), `cigar` (packed uint32, len << 4 | op), `hit_off` (hits of read i are
    hit_off[i] .. hit_off[i + 1]) are views of the page-locked result block, nothing is copied or converted.  Indexing
    or iterating materialises `Mapping` objects for the reads asked for only."""

    def __init__(self, res, names, lens, cs_list, md_list, n_reads):
        self._res, self._names, self._lens, self._cs, self._md, self._n = res, names, lens, cs_list, md_list, n_reads
        self.hits, self.cigar, self.hit_off = res.hits, res.cigar, res.hit_off

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        out = []
        for k in range(int(self.hit_off[i]), int(self.hit_off[i + 1])):
            h = self.hits[k]
            c0, rid = int(h["cigar_off"]), int(h["rid"])
            out.append(Mapping(int(h["qs"]), int(h["qe"]), -1 if h["rev"] else 1, self._names[rid], self._lens[rid], int(h["rs"]), int(h["re"]),
                               int(h["mlen"]), int(h["blen"]), int(h["mapq"]), bool(h["is_primary"]),
                               self.cigar[c0:c0 + int(h["n_cigar"])].copy(),   # a Mapping may outlive the result block
                               int(h["nm"]), self._md[k] if self._md is not None else None, self._cs[k] if self._cs is not None else None))
        return out

    def __iter__(self):
        return (self[i] for i in range(self._n))

    def cs(self, k):
        """cs string of hit k (bytes as generated; None if not requested)"""
        return None if self._cs is None else self._cs[k]

    def close(self):
        """returns the result block to the aligner's pool (also done when the object is collected)"""
        self.hits = self.cigar = self.hit_off = None
        if self._res is not None:
            self._res.close()
            self._res = None


class AlignmentBatchResultIter:
    """Iterator returned by map_batch (lib.rs:923-992): yields (list[Mapping], dict) in completion order."""

    def __init__(self):
        self._q = collections.deque()
        self._cv = threading.Condition()
        self._finished = False
        self._error = None

    def _put(self, items):
        with self._cv:
            self._q.extend(items)
            self._cv.notify_all()

    def _finish(self, error=None):
        with self._cv:
            self._finished, self._error = True, error
            self._cv.notify_all()

    def __iter__(self):
        return self

    def __next__(self):
        try:
            return self._q.popleft()       # deque.popleft is atomic: no lock while results are waiting
        except IndexError:
            pass
        with self._cv:
            while not self._q and not self._finished:
                self._cv.wait(0.05)
            if self._q:
                return self._q.popleft()
            if self._error is not None:
                err, self._error = self._error, None
                raise RuntimeError("device mapping failed: %s" % err)
            raise StopIteration("Finished")   # lib.rs:976


class Aligner:
    """Aligner mimicking mappy / mappy-rs (lib.rs:288-671) on the B200 mapping library."""

    def __init__(self, fn_idx_in=None, preset=None, k=None, w=None, min_cnt=None, min_chain_score=None, min_dp_score=None,
                 bw=None, best_n=None, n_threads=3, fn_idx_out=None, max_frag_len=None, extra_flags=None, seq=None, scoring=None,
                 device=0, devices=None, _lib=None, _tune=None):
        lib = self._lib = _lib or _mmg.Lib()
        io, mo = _mmg.IdxOpt(), _mmg.MapOpt()
        lib.check(lib.L.mmg_set_opt(None, ctypes.byref(io), ctypes.byref(mo)))                     # lib.rs:333
        if preset is not None:
            lib.check(lib.L.mmg_set_opt(str(preset).encode(), ctypes.byref(io), ctypes.byref(mo)))  # lib.rs:336
        mo.flag |= 4                                                                                # lib.rs:339
        io.batch_size |= 0x7fffffffffffffff                                                         # lib.rs:340
        if k is not None: io.k = k
        if w is not None: io.w = w
        if min_cnt is not None: mo.min_cnt = min_cnt
        if min_chain_score is not None: mo.min_chain_score = min_chain_score
        if min_dp_score is not None: mo.min_dp_max = min_dp_score
        if bw is not None: mo.bw = bw
        if best_n is not None: mo.best_n = best_n
        if max_frag_len is not None: mo.max_frag_len = max_frag_len
        if extra_flags is not None: mo.flag |= extra_flags
        if scoring is not None and len(scoring) >= 4:                                               # lib.rs:369-385
            mo.a, mo.b, mo.q, mo.e = (int(x) for x in scoring[:4])
            mo.q2, mo.e2 = mo.q, mo.e
            if len(scoring) >= 6:
                mo.q2, mo.e2 = int(scoring[4]), int(scoring[5])
                if len(scoring) >= 7:
                    mo.sc_ambi = int(scoring[6])
        if seq is not None:
            raise NotImplementedError("Not Implemented")                                            # lib.rs:388-390
        if fn_idx_out is not None:
            raise NotImplementedError("Not Implemented")                                            # lib.rs:391-394
        self._index = self._aligner = None
        self.n_threads = 0                                                                          # lib.rs:426
        if fn_idx_in is None:
            raise RuntimeError("Did not create or open an index")                                   # lib.rs:435
        self._index = _mmg.Index.open(lib, str(fn_idx_in), io, n_threads)                           # lib.rs:398-412
        lib.check(lib.L.mmg_mapopt_update(ctypes.byref(mo), self._index.h))                         # lib.rs:414
        self._io, self._mo = io, mo
        self._names = [self._index.seq_name(i) for i in range(self._index.n_seq)]
        self._lens = [self._index.seq_len(i) for i in range(self._index.n_seq)]
        # uploads the index once (devices=[...]: replicated over NVLink, batches sharded by bases); raises without a GPU
        self._aligner = _mmg.DeviceAligner(lib, self._index, mo, device=device, devices=devices)
        for key, val in (_tune or {}).items():   # device arena sizes (not mapping semantics)
            self._aligner.set(key, val)
        self._lock = threading.Lock()
        self._pinned = _mmg.PinnedBuffer(lib)

    # ---- index accessors ------------------------------------------------------------
    def __bool__(self):
        return self._index is not None                                                              # lib.rs:651-653

    @property
    def seq_names(self):
        if self._index is None:
            raise RuntimeError("Index hasn't loaded")                                               # lib.rs:441-443
        return list(self._names)

    @property
    def k(self): return self._index.k
    @property
    def w(self): return self._index.w
    @property
    def n_seq(self): return self._index.n_seq

    def seq(self, name, start=0, end=2147483647):
        """(Sub)sequence of a contig, or None (lib.rs:464-470, 706-766)."""
        if self._index is None or ((self._mo.flag & 4) and (self._index.flag & 2)):
            return None
        rid = self._index.name2id(name)
        if rid < 0 or rid >= self._index.n_seq:
            return None
        ln = self._lens[rid]
        if start >= ln or start >= end:
            return None
        if end < 0 or end > ln:
            end = ln
        codes = self._index.getseq(rid, start, end)
        if codes is None or (codes > 4).any():
            return None
        return np.frombuffer(b"ACGTN", dtype=np.uint8)[codes].tobytes().decode()

    # ---- mapping ----------------------------------------------------------------------
    def _map_reads(self, seqs, cs, md):
        """list[str] -> per-read lists of Mapping, through mmg_map_batch (host buffers in, results out)."""
        n = len(seqs)
        offs = np.zeros(n + 1, dtype=np.uint64)
        # ASCII reads (every real one): one join, then one copy out of the str's own buffer straight into the pinned block
        # (ctypes releases the GIL for it); anything else goes through encode() and is sized by its UTF-8 bytes
        joined = "".join(seqs)
        size = ctypes.c_ssize_t(0)
        src = _utf8_and_size(joined, ctypes.byref(size))
        if size.value == len(joined):
            np.cumsum(np.fromiter(map(len, seqs), dtype=np.uint64, count=n), out=offs[1:])
            bs = None
        else:
            bs = [s.encode() for s in seqs]
            np.cumsum(np.fromiter(map(len, bs), dtype=np.uint64, count=n), out=offs[1:])
        with self._lock:
            # the reads are assembled in page-locked memory: the library's chunked host->device copies then overlap its kernels
            buf = self._pinned.view(int(offs[-1]))
            if offs[-1]:
                if bs is None:
                    ctypes.memmove(buf.ctypes.data, src, size.value)
                else:
                    buf[:] = np.frombuffer(b"".join(bs), dtype=np.uint8)
            res = self._aligner.map_batch(buf, offs)
            cs_l = _mmg.gen_tags(self._lib, self._index, buf, offs, res, 0) if cs else None   # reads the shared pinned buffer
            md_l = _mmg.gen_tags(self._lib, self._index, buf, offs, res, 1) if md else None
        return _mappings_of_batch(res, self._names, self._lens, cs_l, md_l, n)

    def pinned_buffer(self, n_bytes):
        """A page-locked numpy uint8 array of n_bytes the caller can assemble a batch in (a FASTQ reader writing its
        records back to back): `map_arrays` on it copies nothing on the host.  Valid until the next call."""
        return self._pinned.view(int(n_bytes))[:int(n_bytes)]

    def map_arrays(self, bases, offsets, cs=True, MD=False):
        """Bulk entry next to the reference's API: `bases` holds the reads back to back
        (bytes, bytearray, memoryview or numpy uint8; ASCII), `offsets` has n + 1 entries.  Returns a `BatchMappings`:
        raw result arrays plus `Mapping` objects on demand.  No per-read dict, str or Mapping is created here."""
        if self._index is None:
            raise RuntimeError("No index")
        offs = np.ascontiguousarray(offsets, dtype=np.uint64)
        if offs.ndim != 1 or len(offs) < 1 or (len(offs) > 1 and (np.diff(offs.astype(np.int64)) <= 0).any()):
            raise ValueError("`offsets` must be increasing, one more than the number of reads (an empty read: 'Sequence is empty')")
        src = bases if isinstance(bases, np.ndarray) else np.frombuffer(bases, dtype=np.uint8)
        if src.dtype != np.uint8 or src.ndim != 1 or int(offs[-1]) > len(src):
            raise ValueError("`bases` must be a flat uint8 / bytes buffer that covers offsets[-1]")
        n = len(offs) - 1
        with self._lock:
            mine = self._pinned.arr is not None and src.ctypes.data == self._pinned.arr.ctypes.data
            buf = src if mine else self._pinned.view(int(offs[-1]))
            if not mine and offs[-1]:
                buf[:int(offs[-1])] = src[:int(offs[-1])]
            res = self._aligner.map_batch(buf, offs, zero_copy=True)
            cs_l = _mmg.gen_tags(self._lib, self._index, buf, offs, res, 0) if cs else None
            md_l = _mmg.gen_tags(self._lib, self._index, buf, offs, res, 1) if MD else None
        return BatchMappings(res, self._names, self._lens, cs_l, md_l, n)

    def map(self, seq, seq2=None, cs=False, MD=False):
        """Map a single read, blocking (lib.rs:472-514)."""
        if seq2 is not None:
            raise NotImplementedError("Using `seq2` is not implemented")
        if self._index is None:
            raise RuntimeError("No index")
        if len(seq) == 0:
            raise RuntimeError("Sequence is empty")   # crate minimap2 Aligner::map
        return self._map_reads([seq], cs, MD)[0]

    def map_no_op(self, _seq, seq2=None, _cs=False, _MD=False):
        """Canned mapping used to isolate wrapper overhead (lib.rs:517-533, 675-693)."""
        if seq2 is not None:
            raise NotImplementedError("Using `seq2` is not implemented")
        return [Mapping(0, 1000, 1, "Hello", 101010, 10, 1010, 1000, 1000, 60, True, [], 0, None, "Cigar string")]

    def enable_threading(self, n_threads):
        """lib.rs:541-636 spawns CPU workers; here it arms the device batch pipeline (the GPU replaces the pool)."""
        self.n_threads = int(n_threads)

    def map_batch(self, seqs, back_off=True):
        """Align a batch of dicts with a `seq` key; returns an iterator of (list[Mapping], dict) (lib.rs:640-648, 771-906)."""
        if self.n_threads == 0:
            raise RuntimeError("Multi threading not enabled on this instance. Please call `.enable_threading()`")
        if isinstance(seqs, (dict, set, frozenset)) or not (isinstance(seqs, (list, tuple)) or hasattr(seqs, "__next__") or
                                                            (hasattr(seqs, "__getitem__") and hasattr(seqs, "__len__"))):
            raise TypeError("Unsupported batch type, pass a list, iter, generator or tuple")
        res_iter = AlignmentBatchResultIter()
        self._streamed_probe = lambda: len(res_iter._q)   # results already waiting for the consumer (tests)
        work = collections.deque()          # the bounded work queue of lib.rs:301,429
        cv = threading.Condition()
        state = {"done": False, "abort": False, "bases": 0, "last_push": time.monotonic()}

        def worker():
            """The reference's N workers pop reads while the producer is still pushing (lib.rs:559-633).  Here one
            worker hands device-sized batches to the GPU as soon as enough bases are queued, the queue is full, the
            producer pauses, or the input ends - results reach the consumer while the input is still being produced."""
            try:
                while True:
                    with cv:
                        while not state["abort"]:
                            n = len(work)
                            if state["done"] or n >= _WORK_QUEUE_CAP or state["bases"] >= _DRAIN_BASES or \
                               (n and time.monotonic() - state["last_push"] >= _IDLE_S):
                                break
                            cv.wait(_IDLE_S / 2 if n else 0.05)
                        if state["abort"]:   # the producer raised: nothing is returned to the caller (lib.rs:847-866 return early)
                            break
                        items = []
                        try:                       # item by item: the producer appends without the lock
                            while True:
                                items.append(work.popleft())
                        except IndexError:
                            pass
                        state["bases"] = 0
                        done = state["done"]
                        cv.notify_all()
                    if items:
                        per_read = self._map_reads([s for _, s in items], True, False)   # cs=true, md=false: lib.rs:587-593
                        res_iter._put([(m, d) for m, (d, _) in zip(per_read, items)])
                    if done and not items:
                        break
                res_iter._finish()
            except Exception as e:   # surfaced to the consumer instead of silently dropping reads (lib.rs:621-623 logs and drops)
                res_iter._finish(error=str(e))

        th = threading.Thread(target=worker, daemon=True)
        th.start()
        self._workers = [t for t in getattr(self, "_workers", []) if t.is_alive()] + [th]
        ok = False
        try:
            for id_num, item in enumerate(iter(seqs)):
                if not isinstance(item, dict) or not _STR_ONLY.issuperset(map(type, item)) and not all(isinstance(k_, str) for k_ in item):
                    raise TypeError("Element in iterable is not a dictionary")
                data = dict(item)                      # lib.rs:847-855: a new dict with the same keys / values
                if "seq" not in item:
                    raise KeyError("AHHH Key 🗝️  not found in iterated dictionary")
                seq = item["seq"]
                if not isinstance(seq, str):
                    raise ValueError("`seq` must be a string")
                if len(work) < _WORK_QUEUE_CAP - 1 and state["bases"] + len(seq) < _DRAIN_BASES:
                    # the common case takes no lock: deque.append is atomic, `bases` has one writer while the queue fills,
                    # and the worker wakes on its own timer if a notify is missed
                    work.append((data, seq))
                    state["bases"] += len(seq)
                    state["last_push"] = time.monotonic()
                    continue
                with cv:
                    if len(work) >= _WORK_QUEUE_CAP:
                        if not back_off:
                            raise RuntimeError(
                                "Internal error adding data to work queue, without backoff. Work(({id_num}, ..)) {id_num}. "
                                "Is your fastq batch larger than 50000? Perhaps try `map_batch` with back_off=True?".format(id_num=id_num))
                        # lib.rs:871-885 sleeps 50 ms doubling, six times, and then DROPS the read with a message on
                        # stderr; here the producer waits until the worker has taken the queue: the 50000-entry limit
                        # is honoured and no read is lost
                        cv.notify_all()
                        while len(work) >= _WORK_QUEUE_CAP and not res_iter._finished:
                            cv.wait(0.05)
                    work.append((data, seq))
                    state["bases"] += len(seq)
                    state["last_push"] = time.monotonic()
                    if state["bases"] >= _DRAIN_BASES or len(work) >= _WORK_QUEUE_CAP:
                        cv.notify_all()
            ok = True
        finally:
            with cv:
                state["done"] = True
                state["abort"] = not ok
                cv.notify_all()
        return res_iter

    def close(self):
        for t in getattr(self, "_workers", []):   # device buffers must outlive in-flight batches
            t.join()
        self._workers = []
        if getattr(self, "_pinned", None) is not None:
            self._pinned.close()
            self._pinned = None
        if self._aligner is not None:
            self._aligner.close()
            self._aligner = None
        if self._index is not None:
            self._index.close()
            self._index = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
