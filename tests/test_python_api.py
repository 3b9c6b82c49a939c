"""Behavioural contract of the Python surface (`mappy_rs.Aligner`, `Mapping`, the `map_batch` iterator).

What a user of the reference observes is listed in SURVEY.md section 8(b) with the reference lines that pin it
(/root/reference/src/lib.rs and the behaviours its tests/python_test.py exercises); this file checks each item of that
list against the mirror in mappy-rs_b200/mappy_rs.  Two back ends: `emu` (CPU: the kernel sources under the SIMT
emulator, with the work-queue sizes scaled down) and `gpu` (the product library on a B200, real sizes).
"""
import itertools
import os

import pytest

from conftest import EMU_LIB, GOLDEN
from test_oracle_fixtures import BACILLUS, ENTEROCOCCUS, read_fasta

MMI = os.path.join(GOLDEN, "test.mmi")
FASTA = os.path.join(GOLDEN, "test.fa")
SMALL_ARENAS = {"tb_cap": 1 << 26, "cigar_cap": 1 << 22, "jobs_cap": 1 << 14, "chunk_bases": 1 << 20, "anchor_cap": 1 << 18,
                "chunk_reads": 1 << 16, "regs_cap": 1 << 17, "big_per_warp": 1 << 18}

# messages a caller may match on (src/lib.rs:777-792, 847-885)
MSG_NO_THREADS = "Multi threading not enabled on this instance. Please call `.enable_threading()`"
MSG_BAD_BATCH = "Unsupported batch type, pass a list, iter, generator or tuple"
MSG_NOT_DICT = "Element in iterable is not a dictionary"
MSG_NO_KEY = "AHHH Key 🗝️  not found in iterated dictionary"
MSG_NOT_STR = "`seq` must be a string"
MSG_QUEUE_FULL = ("Internal error adding data to work queue, without backoff",
                  "Is your fastq batch larger than 50000? Perhaps try `map_batch` with back_off=True?")


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def aligner(request):
    """(Aligner on test.mmi, number of reads that overflows the work queue of lib.rs:429)"""
    import mappy_rs
    from mappy_rs import _mmg, aligner as impl
    if request.param == "emu":   # ~10^2 reads/s: a 500-entry queue and 1000 reads stand in for 50 000 and 100 000
        request.getfixturevalue("emu_lib")
        request.getfixturevalue("monkeypatch").setattr(impl, "_WORK_QUEUE_CAP", 500)
        a, flood = mappy_rs.Aligner(MMI, _lib=_mmg.Lib(EMU_LIB), _tune=SMALL_ARENAS), 1000
    else:
        a, flood = mappy_rs.Aligner(MMI), 100000
    yield a, flood
    a.close()


def contig_records(copies=10):
    """the four contigs of test.fa, `copies` times, as the dicts map_batch takes"""
    seqs = [s for _, s in read_fasta(FASTA)]
    return [{"id": i, "seq": s} for i, s in enumerate(seqs * copies)]


# ---- index accessors (src/lib.rs:439-470, 651-670) ---------------------------------------------------------

def test_index_accessors(aligner):
    a, _ = aligner
    assert bool(a)
    assert (a.k, a.w, a.n_seq) == (15, 10, 4)
    assert sorted(a.seq_names) == ["Bacillus_subtilis", "Enterococcus_faecalis", "Escherichia_coli_1", "Escherichia_coli_2"]


def test_seq_slices(aligner):
    a, _ = aligner
    assert a.seq("Bacillus_subtilis") == BACILLUS
    assert a.seq("Bacillus_subtilis", 10, 20) == BACILLUS[10:20]
    assert a.seq("no such contig") is None
    assert a.seq("Bacillus_subtilis", 500, 600) is None          # start past the end


# ---- Aligner.map (src/lib.rs:472-514) ------------------------------------------------------------------------

def test_single_read_maps_end_to_end(aligner):
    a, _ = aligner
    hits = a.map(ENTEROCOCCUS, cs=True)
    assert len(hits) == 1
    m = hits[0]
    assert (m.target_start, m.target_end) == (0, 400)            # what the reference's own test pins
    # the remaining fields of this exact 400-mer
    assert (m.ctg, m.ctg_len, m.q_st, m.q_en, m.strand, m.mapq, m.is_primary) == ("Enterococcus_faecalis", 400, 0, 400, 1, 60, True)
    assert (m.cigar, m.cigar_str, m.NM, m.cs) == ([(400, 0)], "400M", 0, ":400")
    assert str(m) == "\t".join(["0", "400", "+", "Enterococcus_faecalis", "400", "0", "400", "400", "400", "60", "tp:A:P", "cg:Z:400M"])


# ---- map_batch: accepted containers, results, ordering of the checks (src/lib.rs:771-906) ------------------------

CONTAINERS = {
    "list": list,
    "tuple": tuple,
    "iterator": iter,
    "generator": lambda recs: (r for r in recs),
}


@pytest.mark.parametrize("kind", sorted(CONTAINERS))
def test_batch_containers(aligner, kind):
    a, _ = aligner
    a.enable_threading(2)
    n = 0
    for hits, meta in a.map_batch(CONTAINERS[kind](contig_records())):
        n += 1
        assert len(hits) == 1 and (hits[0].r_st, hits[0].r_en, hits[0].cs) == (0, 400, ":400")
        assert "id" in meta and "seq" in meta                     # the caller's dict comes back with the hits
    assert n == 40


def test_batch_before_enable_threading(aligner):
    a, _ = aligner
    with pytest.raises(RuntimeError) as err:
        a.map_batch(contig_records())
    assert MSG_NO_THREADS in str(err.value)


def _one_dict(recs): return recs[0]                                # a dict is not a batch
def _dict_of_dicts(recs): return dict(enumerate(recs))
def _bare_strings(recs): return [r["seq"] for r in recs]
def _wrong_key(recs): return [{"SEQ": r["seq"]} for r in recs]
def _bytes_seq(recs): return [{"seq": r["seq"].encode()} for r in recs]


@pytest.mark.parametrize("make,exc,msg", [
    (_one_dict, TypeError, MSG_BAD_BATCH),
    (_dict_of_dicts, TypeError, MSG_BAD_BATCH),
    (_bare_strings, TypeError, MSG_NOT_DICT),
    (_wrong_key, KeyError, MSG_NO_KEY),
    (_bytes_seq, ValueError, MSG_NOT_STR),
], ids=["dict", "dict-of-dicts", "strings", "no-seq-key", "bytes-seq"])
def test_batch_input_validation(aligner, make, exc, msg):
    a, _ = aligner
    a.enable_threading(2)
    with pytest.raises(exc) as err:
        a.map_batch(make(contig_records()))
    assert msg in str(err.value) or msg in str(err)


def test_batch_from_exhausted_iterator_is_empty(aligner):
    a, _ = aligner
    a.enable_threading(2)
    it = iter(contig_records())
    for _ in it:
        pass
    assert list(a.map_batch(it)) == []


def test_batch_larger_than_the_work_queue(aligner):
    """with back-off every read comes back; without it the producer fails with the documented message"""
    a, flood = aligner
    a.enable_threading(4)
    rec = contig_records(1)[0]
    assert sum(1 for _ in a.map_batch(itertools.repeat(rec, flood), back_off=True)) == flood
    with pytest.raises(RuntimeError) as err:
        for _ in a.map_batch(itertools.repeat(rec, flood), back_off=False):
            pass
    assert all(part in str(err.value) for part in MSG_QUEUE_FULL)


# ---- constructor and the not-implemented corners (src/lib.rs:388-394, 435, 476-480, 517-533) ------------------------

def test_constructor_and_unimplemented_corners(emu_lib):
    import mappy_rs
    with pytest.raises(RuntimeError, match="Did not create or open an index"):
        mappy_rs.Aligner(_lib=emu_lib)
    for kw in ({"seq": "ACGT"}, {"fn_idx_out": "x.mmi"}):
        with pytest.raises(NotImplementedError):
            mappy_rs.Aligner(MMI, _lib=emu_lib, **kw)
    a = mappy_rs.Aligner(MMI, _lib=emu_lib, _tune=SMALL_ARENAS)
    try:
        with pytest.raises(NotImplementedError, match="Using `seq2` is not implemented"):
            a.map("ACGT", seq2="ACGT")
        assert a.map_no_op("ACGT")[0].ctg == "Hello"
    finally:
        a.close()


def test_results_stream_while_the_producer_is_still_feeding(aligner):
    """/root/reference/src/lib.rs:559-633, 972-991: workers map while the producer is still pushing, so a consumer can
    read results before a slow generator is exhausted.  The device worker starts a batch when the producer pauses."""
    import threading
    import time
    a, _ = aligner
    a.enable_threading(2)
    recs = contig_records(2)
    gate = threading.Event()
    produced = []

    def slow():
        for i, r in enumerate(recs):
            if i == 4:
                gate.wait(60)          # the producer stalls until the consumer has seen the first results
            produced.append(i)
            yield r

    out = []
    box = {}

    def feed():
        box["it"] = a.map_batch(slow())

    th = threading.Thread(target=feed)
    th.start()
    deadline = time.time() + 60
    while "it" not in box and len(produced) < 4 and time.time() < deadline:
        time.sleep(0.01)
    # map_batch returns only when the producer is done (lib.rs:771-906 runs it on the caller's thread); the iterator it
    # will return is fed by the worker meanwhile: wait until the first four reads have been mapped, then open the gate
    while time.time() < deadline:
        w = [t for t in getattr(a, "_workers", []) if t.is_alive()]
        if len(produced) >= 4 and getattr(a, "_streamed_probe", lambda: 0)() >= 4:
            break
        time.sleep(0.02)
    early = getattr(a, "_streamed_probe", lambda: 0)()
    gate.set()
    th.join(120)
    out = list(box["it"])
    assert len(out) == len(recs) and early >= 4, (len(out), early)
    assert sorted(d["id"] for _, d in out) == list(range(len(recs)))


def test_map_arrays_is_the_same_answer_without_per_read_objects(aligner):
    """a bulk entry next to the reference's API.  Reads back to back in one buffer (here
    assembled in the aligner's page-locked buffer: no host copy), results as raw arrays with `Mapping`s on demand; every
    field, CIGAR and cs must equal what `map()` returns read by read."""
    import numpy as np
    aligner, _ = aligner
    reads = [r["seq"] for r in contig_records(2)] + [BACILLUS[:300], ENTEROCOCCUS[50:350]]
    lens = [len(s) for s in reads]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    buf = aligner.pinned_buffer(int(offs[-1]))
    buf[:] = np.frombuffer("".join(reads).encode(), dtype=np.uint8)
    res = aligner.map_arrays(buf, offs, cs=True, MD=True)
    assert len(res) == len(reads) and len(res.hits) == int(res.hit_off[-1])
    for i, s in enumerate(reads):
        assert res[i] == aligner.map(s, cs=True, MD=True), i
    assert [len(m) for m in res] == [len(res[i]) for i in range(len(res))]
    res.close()
    # from plain bytes (copied once into the page-locked buffer), without tags
    res2 = aligner.map_arrays("".join(reads).encode(), offs, cs=False)
    assert [[(m.target_name, m.target_start, m.target_end, m.NM) for m in ms] for ms in res2] == \
           [[(m.target_name, m.target_start, m.target_end, m.NM) for m in aligner.map(s)] for s in reads]
    assert all(m.cs is None for ms in res2 for m in ms)
    with pytest.raises(ValueError):
        aligner.map_arrays(buf, np.array([0, 10, 10], dtype=np.uint64))     # an empty read
    with pytest.raises(ValueError):
        aligner.map_arrays(b"ACGT", np.array([0, 400], dtype=np.uint64))    # offsets beyond the buffer
