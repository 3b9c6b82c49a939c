"""Parity tests proper: the product library (nvcc, sm_100a) on a real B200 through
the C ABI, against the oracle on the same seeded inputs.  Bit-exact on every hit
field; stage dumps attribute a mismatch to a kernel."""
import os

import numpy as np
import pytest

import data_gen
import parity
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
MMI = os.path.join(GOLDEN, "test.mmi")


@pytest.fixture(scope="module")
def case5mb(gpu_lib, oracle_mod):
    """BASELINE.json configs[1]: 5 Mb random reference, map-ont, mapping-only."""
    ref, coff, names = data_gen.config1_reference()
    seqs = [ref.tobytes()]
    c = parity.Case(gpu_lib, names, seqs)
    c.ref, c.coff = ref, coff
    yield c
    c.close()


def test_native_library_is_loaded(gpu_lib):
    assert "sm_100a" in gpu_lib.version()
    assert gpu_lib.path.endswith("mappy-rs_b200/libmmg.so")


def test_config1_sample_bit_exact(case5mb):
    buf, offs, _ = data_gen.config1_reads(case5mb.ref, case5mb.coff, 20000)
    dev = case5mb.aligner.map_batch(buf, offs)
    ora = case5mb.oracle.map_batch(buf, offs, os.cpu_count() or 8)
    assert parity.compare_stats(dev, ora) == []
    assert parity.compare_hits(dev, ora) == []
    assert len(dev.hits) >= 19900


def test_config1_stages(case5mb):
    buf, offs, _ = data_gen.make_reads(77, case5mb.ref, case5mb.coff, 2000, 1000, 10000)
    dev, diffs = parity.compare_stages(case5mb, buf, offs, max_reads=400)
    assert diffs == []


def test_edge_case_reads(case5mb, oracle_mod):
    ref = case5mb.ref
    rs = np.random.RandomState(5)
    base = ref[1000:4000].tobytes().decode()
    reads = [
        "", "A", "ACGT", base[:14], base[:15], base[:24], base[:25], base[:40],
        "N" * 50, base[:300] + "N" * 7 + base[300:900], "".join(c if i % 10 else "N" for i, c in enumerate(base[:1200])),
        "A" * 600, "AT" * 300, "ACGT" * 200, base[:500] + "A" * 300 + base[500:1000],
        base[:700].lower(), base[100:900][::-1], "ACGTTGCA" * 60 + base[:400] + "TGCAACGT" * 40,
        base[:800] + ref[50000:50800].tobytes().decode(), "".join(rs.choice(list("ACGT"), 2000)),
        base, base[:1000] + "N" * 32 + base[1000:], ref[200000:260000].tobytes().decode(),
    ]
    buf, offs = oracle_mod.pack_reads(reads)
    dev, diffs = parity.compare_stages(case5mb, buf, offs)
    ora = case5mb.oracle.map_batch(buf, offs, 1)
    assert diffs == []
    assert parity.compare_hits(dev, ora) == []


def test_reference_fixture(gpu_lib, oracle_mod):
    c = parity.Case(gpu_lib, None, None, mmi=MMI)
    try:
        seqs = [c.oracle.seq(n) for n in c.oracle.seq_names] * 10
        buf, offs = oracle_mod.pack_reads(seqs)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, 1)
        assert parity.compare_hits(dev, ora) == [] and len(dev.hits) == 40
    finally:
        c.close()


def test_hifi_preset(gpu_lib, oracle_mod):
    ref, coff, names, seqs = parity.random_reference(21, [2000000])
    c = parity.Case(gpu_lib, names, seqs, preset="map-hifi")
    try:
        buf, offs, _ = data_gen.make_reads(22, ref, coff, 2000, 10000, 25000, len_mean=15000, len_sd=2000, p_sub=0.002, p_ins=0.0015, p_del=0.0015)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, os.cpu_count() or 8)
        assert parity.compare_hits(dev, ora) == [] and parity.compare_stats(dev, ora) == []
    finally:
        c.close()


def test_multi_chunk_equals_oracle(gpu_lib, oracle_mod):
    ref, coff, names, seqs = parity.random_reference(31, [1000000])
    c = parity.Case(gpu_lib, names, seqs)
    try:
        c.aligner.set("chunk_bases", 2_000_000)
        c.aligner.set("chunk_reads", 512)
        c.aligner.set("anchor_cap", 150_000)
        buf, offs, _ = data_gen.make_reads(32, ref, coff, 5000, 300, 8000)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, os.cpu_count() or 8)
        assert parity.compare_hits(dev, ora) == [] and parity.compare_stats(dev, ora) == []
    finally:
        c.close()


def test_device_logf_matches_host_libm(gpu_lib):
    """mapq uses logf(); the device restatement of glibc's algorithm is compared with the host libm."""
    assert parity.logf_mismatches(gpu_lib, 300000) == 0


def test_repeats_chimeras_and_rechain(gpu_lib, oracle_mod):
    ref, coff, names, seqs = parity.random_reference(41, [3000000, 1500000], n_repeats=600, rep_min=300, rep_max=6000, rep_div=0.03)
    c = parity.Case(gpu_lib, names, seqs)
    try:
        buf, offs = data_gen.make_sv_reads(51, ref, coff, 6000, 400, 8000)
        dev, stage_diffs = parity.compare_stages(c, buf, offs, max_reads=300)
        ora = c.oracle.map_batch(buf, offs, os.cpu_count() or 8)
        assert ora.stats["n_rechain"] > 1000 and dev.stats["n_rechain"] == ora.stats["n_rechain"]
        assert stage_diffs == []
        assert parity.compare_stats(dev, ora) == []
        assert parity.compare_hits(dev, ora) == []
    finally:
        c.close()


def test_cigar_mode_config1_sample(case5mb, gpu_lib, oracle_mod):
    """CIGAR on (what mappy-rs always runs): 3000 config-1 reads, every field and every CIGAR op."""
    c = parity.Case(gpu_lib, ["chr1"], [case5mb.ref.tobytes()], cigar=True)
    try:
        buf, offs, _ = data_gen.config1_reads(case5mb.ref, case5mb.coff, 3000)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, os.cpu_count() or 8)
        assert parity.compare_hits(dev, ora) == []
        assert len(dev.cigar) == len(ora.cigar)
    finally:
        c.close()


def test_cigar_mode_splits_and_inversions(gpu_lib, oracle_mod):
    ref, coff, names, seqs = parity.random_reference(41, [3000000, 1500000], n_repeats=600, rep_min=300, rep_max=6000, rep_div=0.03)
    c = parity.Case(gpu_lib, names, seqs, cigar=True)
    try:
        buf, offs = data_gen.make_sv_reads(51, ref, coff, 1500, 400, 6000)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, os.cpu_count() or 8)
        assert parity.compare_hits(dev, ora) == []
        f = ora.hits["flags"]
        assert ((f & 8) > 0).sum() > 50 and ((f & 2) > 0).sum() > 20
    finally:
        c.close()


def test_cigar_mode_hifi(gpu_lib, oracle_mod):
    ref, coff, names, seqs = parity.random_reference(21, [2000000])
    c = parity.Case(gpu_lib, names, seqs, preset="map-hifi", cigar=True)
    try:
        buf, offs, _ = data_gen.make_reads(22, ref, coff, 300, 10000, 25000, len_mean=15000, len_sd=2000, p_sub=0.002, p_ins=0.0015, p_del=0.0015)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, os.cpu_count() or 8)
        assert parity.compare_hits(dev, ora) == []
    finally:
        c.close()


def test_cigar_mode_fixture_map_one(gpu_lib, oracle_mod):
    c = parity.Case(gpu_lib, None, None, mmi=MMI, cigar=True)
    try:
        seqs = [c.oracle.seq(n) for n in c.oracle.seq_names] * 10
        buf, offs = oracle_mod.pack_reads(seqs)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, 1)
        assert parity.compare_hits(dev, ora) == []
        assert [(int(h["rs"]), int(h["re"])) for h in dev.hits] == [(0, 400)] * 40
    finally:
        c.close()


def _random_hit_case(lib, oracle_mod, ref_len, seed, k, w, max_gap):
    """Short k-mers / a large reference: most anchors of a read are isolated random hits."""
    import ctypes
    from mappy_rs import _mmg
    ref, coff, names, seqs = parity.random_reference(seed, [ref_len])
    c = parity.Case.__new__(parity.Case)
    c.lib = lib
    c.io, c.mopt = _mmg.IdxOpt(), _mmg.MapOpt()
    lib.check(lib.L.mmg_set_opt(None, ctypes.byref(c.io), ctypes.byref(c.mopt)))
    c.io.k, c.io.w = k, w
    c.oracle = oracle_mod.Oracle(names=names, seqs=seqs, k=k, w=w)
    c.index = _mmg.Index.build(lib, c.io, names, seqs)
    c.mopt.flag = 0
    c.oracle.set_opt("flag", 0)
    if max_gap:
        for k_, v in (("max_gap", max_gap), ("bw", 400), ("bw_long", 400)):
            setattr(c.mopt, k_, v)
            c.oracle.set_opt(k_, v)
    lib.check(lib.L.mmg_mapopt_update(ctypes.byref(c.mopt), c.index.h))
    assert c.mopt.mid_occ == c.oracle.get_opt("mid_occ")
    c.aligner = _mmg.DeviceAligner(lib, c.index, c.mopt)
    return c, ref, coff


def test_isolated_anchor_filter_and_radix_sort_at_scale(gpu_lib, oracle_mod):
    """Thousands of anchors per read, mostly isolated random hits (13-mers on 60 Mb): the filter drops them, the
    reads that keep everything (repeated minimizer hashes) go through the CTA radix sort, and every hit still
    equals the oracle's, which sorts and chains all anchors."""
    c, ref, coff = _random_hit_case(gpu_lib, oracle_mod, 60_000_000, 101, 13, 5, 0)
    try:
        buf, offs, _ = data_gen.make_reads(102, ref, coff, 3000, 1000, 9000)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, os.cpu_count() or 8)
        assert ora.stats["n_anchor"] > 1500 * 3000 and dev.stats["n_dropped"] > 0.3 * ora.stats["n_anchor"]
        assert parity.compare_stats(dev, ora) == [] and parity.compare_hits(dev, ora) == []
        c.aligner.set("anchor_filter", 0)
        dev2 = c.aligner.map_batch(buf, offs)
        assert dev2.stats["n_dropped"] == 0 and parity.compare_hits(dev2, ora) == []
    finally:
        c.close()


def _primary(dev, n):
    """index of the first (best) primary hit of every read, -1 if unmapped"""
    first = dev.hit_off[:-1].astype(np.int64)
    has = np.diff(dev.hit_off.astype(np.int64)) > 0
    return np.where(has, first, -1)


def _same_hits(x, y):
    """field by field (numpy does not copy the padding bytes of an aligned record)"""
    return x.shape == y.shape and all(np.array_equal(x[f], y[f]) for f in x.dtype.names)


def test_reverse_complement_symmetry(case5mb):
    """A read and its reverse complement map to the same target interval on opposite strands with the same
    chain score (minimizers are canonical); 20 000 reads of configs[1]."""
    buf, offs, _ = data_gen.config1_reads(case5mb.ref, case5mb.coff, 20000)
    comp = np.zeros(256, dtype=np.uint8)
    comp[:] = ord("N")
    for x, y in zip(b"ACGTacgt", b"TGCAtgca"):
        comp[x] = y
    n = len(offs) - 1
    rbuf = np.empty_like(buf)
    for i in range(n):
        s, e = int(offs[i]), int(offs[i + 1])
        rbuf[s:e] = comp[buf[s:e]][::-1]
    a = case5mb.aligner.map_batch(buf, offs)
    b = case5mb.aligner.map_batch(rbuf, offs)
    pa, pb = _primary(a, n), _primary(b, n)
    both = (pa >= 0) & (pb >= 0)
    assert both.mean() > 0.995
    ha, hb = a.hits[pa[both]], b.hits[pb[both]]
    same = (ha["rid"] == hb["rid"]) & (ha["rs"] == hb["rs"]) & (ha["re"] == hb["re"]) & (ha["rev"] != hb["rev"]) & (ha["score"] == hb["score"])
    assert same.mean() > 0.99, same.mean()


def test_no_dependence_on_what_cudamalloc_returns(gpu_lib, oracle_mod):
    """MMG_POISON fills every device arena with 0xCD right after allocation: mapping-only (filter, radix sort, re-chain)
    and CIGAR mode (splits, inversions) must still equal the oracle, i.e. nothing reads memory it has not written."""
    os.environ["MMG_POISON"] = "1"
    try:
        ref, coff, names, seqs = parity.random_reference(41, [300000, 150000], n_repeats=60, rep_min=300, rep_max=4000, rep_div=0.03)
        for cigar in (False, True):
            c = parity.Case(gpu_lib, names, seqs, cigar=cigar)
            try:
                buf, offs = data_gen.make_sv_reads(57, ref, coff, 400)
                dev = c.aligner.map_batch(buf, offs)
                ora = c.oracle.map_batch(buf, offs, os.cpu_count() or 8)
                assert parity.compare_hits(dev, ora) == []
            finally:
                c.close()
        c, ref, coff = _random_hit_case(gpu_lib, oracle_mod, 60_000_000, 103, 13, 5, 0)
        try:
            buf, offs, _ = data_gen.make_reads(104, ref, coff, 1500, 1000, 8000)
            dev = c.aligner.map_batch(buf, offs)
            ora = c.oracle.map_batch(buf, offs, os.cpu_count() or 8)
            assert dev.stats["n_dropped"] > 0 and parity.compare_stats(dev, ora) == [] and parity.compare_hits(dev, ora) == []
        finally:
            c.close()
    finally:
        del os.environ["MMG_POISON"]
