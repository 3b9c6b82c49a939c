"""Kernel-source parity on CPU: the product's .cu files compiled against the
test-only SIMT emulator (tests/emu) and compared with the oracle, hit for hit
and stage for stage.  Small inputs (the emulator runs ~10^2-10^3 reads/s); the
same cases at full size are the `gpu` tests."""
import os

import numpy as np
import pytest

import data_gen
import parity
from conftest import GOLDEN

MMI = os.path.join(GOLDEN, "test.mmi")


@pytest.fixture(scope="module")
def plain_case(emu_lib, oracle_mod):
    ref, coff, names, seqs = parity.random_reference(11, [150000, 60000])
    c = parity.Case(emu_lib, names, seqs)
    c.ref, c.coff = ref, coff
    yield c
    c.close()


def test_mapping_only_reads(plain_case):
    buf, offs, _ = data_gen.make_reads(12, plain_case.ref, plain_case.coff, 300, 300, 5000)
    dev, stage_diffs = parity.compare_stages(plain_case, buf, offs, max_reads=60)
    ora = plain_case.oracle.map_batch(buf, offs, 4)
    assert stage_diffs == []
    assert parity.compare_stats(dev, ora) == []
    assert parity.compare_hits(dev, ora) == []
    assert len(dev.hits) >= 290


def test_edge_case_reads(plain_case):
    """SURVEY.md appendix D edge corpus: short reads, N runs, homopolymers, palindromes, empty read."""
    ref = plain_case.ref
    rs = np.random.RandomState(5)
    base = ref[1000:3000].tobytes().decode()
    reads = [
        "", "A", "ACGT", base[:14], base[:15], base[:24], base[:25], base[:40],
        "N" * 50, base[:300] + "N" * 7 + base[300:900], "".join(c if i % 10 else "N" for i, c in enumerate(base[:1200])),
        "A" * 600, "AT" * 300, "ACGT" * 200, base[:500] + "A" * 300 + base[500:1000],
        base[:700].lower(), base[100:900][::-1], "ACGTTGCA" * 60 + base[:400] + "TGCAACGT" * 40,
        base[:800] + ref[50000:50800].tobytes().decode(), "".join(rs.choice(list("ACGT"), 2000)),
        base, base[:1000] + "NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN" + base[1000:],
    ]
    buf, offs = plain_case.oracle and __import__("mm2oracle").pack_reads(reads)
    dev, stage_diffs = parity.compare_stages(plain_case, buf, offs)
    ora = plain_case.oracle.map_batch(buf, offs, 1)
    assert stage_diffs == []
    assert parity.compare_stats(dev, ora) == []
    assert parity.compare_hits(dev, ora) == []


def test_reference_fixture_reads(emu_lib, oracle_mod):
    """The reference's own workload: the 4 contigs of test.fa mapped to test.mmi (tests/python_test.py:167-178)."""
    c = parity.Case(emu_lib, None, None, mmi=MMI)
    try:
        seqs = [c.oracle.seq(n) for n in c.oracle.seq_names] * 3
        buf, offs = oracle_mod.pack_reads(seqs)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, 1)
        assert parity.compare_hits(dev, ora) == []
        assert len(dev.hits) == 12 and set(dev.hits["mapq"].tolist()) == {60}
    finally:
        c.close()


def test_hifi_preset(emu_lib, oracle_mod):
    ref, coff, names, seqs = parity.random_reference(21, [120000])
    c = parity.Case(emu_lib, names, seqs, preset="map-hifi")
    try:
        buf, offs, _ = data_gen.make_reads(22, ref, coff, 60, 2000, 9000, p_sub=0.002, p_ins=0.0015, p_del=0.0015)
        dev, stage_diffs = parity.compare_stages(c, buf, offs, max_reads=30)
        ora = c.oracle.map_batch(buf, offs, 4)
        assert stage_diffs == [] and parity.compare_hits(dev, ora) == []
    finally:
        c.close()


def test_multi_chunk_equals_single_chunk(emu_lib, oracle_mod):
    """Small arenas force several chunks and anchor sub-ranges; results must not change."""
    ref, coff, names, seqs = parity.random_reference(31, [80000])
    c = parity.Case(emu_lib, names, seqs)
    try:
        c.aligner.set("chunk_bases", 20000)
        c.aligner.set("chunk_reads", 16)
        c.aligner.set("anchor_cap", 3000)
        buf, offs, _ = data_gen.make_reads(32, ref, coff, 120, 300, 4000)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, 4)
        assert parity.compare_hits(dev, ora) == [] and parity.compare_stats(dev, ora) == []
    finally:
        c.close()


def test_device_logf_source_matches_host_libm(emu_lib):
    """The glibc-logf restatement in regs.cu (tables typed into the product source) vs the host libm."""
    assert parity.logf_mismatches(emu_lib, 100000) == 0


def test_repeats_chimeras_and_rechain(emu_lib, oracle_mod):
    """Planted repeats + chimeric / SV reads: secondaries, several chains per read and the RMQ
    long-join re-chain (krmq AVL replay) must all agree with the oracle."""
    ref, coff, names, seqs = parity.random_reference(41, [300000, 150000], n_repeats=60, rep_min=300, rep_max=4000, rep_div=0.03)
    c = parity.Case(emu_lib, names, seqs)
    try:
        buf, offs = data_gen.make_sv_reads(51, ref, coff, 250)
        dev, stage_diffs = parity.compare_stages(c, buf, offs, max_reads=250)
        ora = c.oracle.map_batch(buf, offs, 4)
        assert ora.stats["n_rechain"] > 100 and dev.stats["n_rechain"] == ora.stats["n_rechain"]
        assert stage_diffs == []
        assert parity.compare_stats(dev, ora) == []
        assert parity.compare_hits(dev, ora) == []
        assert (ora.hits["is_primary"] == 0).sum() > 0
    finally:
        c.close()


@pytest.mark.parametrize("mode", ["1", "14", "200"])
def test_rechain_tree_replay_paths(emu_lib, oracle_mod, mode, monkeypatch):
    """mm_lchain_rmq on the device answers the outer range-minimum with a warp scan and falls back to the replay of
    krmq's AVL tree when the minimum is tied.  MMG_RMQ_SERIAL=1 sends every read through the tree; 2k makes the warp
    form give up at its k-th anchor (early / deep inside a read), so the abandon-and-replay path is walked on purpose.
    All three must give what the default gives: the oracle's chains."""
    monkeypatch.setenv("MMG_RMQ_SERIAL", mode)
    ref, coff, names, seqs = parity.random_reference(41, [300000, 150000], n_repeats=60, rep_min=300, rep_max=4000, rep_div=0.03)
    c = parity.Case(emu_lib, names, seqs)
    try:
        buf, offs = data_gen.make_sv_reads(51, ref, coff, 120)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, 4)
        assert ora.stats["n_rechain"] > 40
        assert parity.compare_stats(dev, ora) == []
        assert parity.compare_hits(dev, ora) == []
    finally:
        c.close()


def test_arenas_step_down_when_the_device_is_short_of_memory(emu_lib, oracle_mod, monkeypatch):
    """The default arenas assume the aligner owns the GPU.  With less memory left (another aligner on the device; here
    MMG_ALLOC_LIMIT plays that part) the allocation is rolled back and repeated one size down until it fits - smaller
    chunks, then a smaller traceback arena - and the results are the same; sizes the caller set are not second-guessed."""
    ref, coff, names, seqs = parity.random_reference(17, [200000])
    buf, offs, _ = data_gen.make_reads(18, ref, coff, 40, 500, 5000)
    c = parity.Case(emu_lib, names, seqs, cigar=True)
    try:
        ora = c.oracle.map_batch(buf, offs, 4)
        monkeypatch.setenv("MMG_ALLOC_LIMIT", str(14 << 30))      # the stock sizes of the test device need ~60 GB
        dev = c.aligner.map_batch(buf, offs)
        assert parity.compare_hits(dev, ora) == []
    finally:
        c.close()
    c = parity.Case(emu_lib, names, seqs, cigar=True)
    try:
        c.aligner.set("tb_cap", 1 << 34)
        monkeypatch.setenv("MMG_ALLOC_LIMIT", str(1 << 30))
        with pytest.raises(RuntimeError, match="MMG_ALLOC_LIMIT"):
            c.aligner.map_batch(buf, offs)
    finally:
        c.close()


def _small_arenas(c):
    for k, v in (("tb_cap", 1 << 28), ("cigar_cap", 1 << 24), ("jobs_cap", 1 << 16), ("chunk_bases", 1 << 22),
                 ("anchor_cap", 1 << 20), ("chunk_reads", 4096), ("regs_cap", 1 << 16), ("big_per_warp", 1 << 18)):
        c.aligner.set(k, v)


def test_cigar_mode_plain_reads(emu_lib, oracle_mod):
    """mappy-rs' real mode (MM_F_CIGAR forced, src/lib.rs:339): extension + gap filling + CIGAR, bit-exact."""
    ref, coff, names, seqs = parity.random_reference(61, [200000])
    c = parity.Case(emu_lib, names, seqs, cigar=True)
    try:
        _small_arenas(c)
        buf, offs, _ = data_gen.make_reads(71, ref, coff, 80, 300, 3000)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, 4)
        assert parity.compare_hits(dev, ora) == []
        assert len(dev.cigar) == len(ora.cigar) > 1000
    finally:
        c.close()


def test_cigar_mode_splits_and_inversions(emu_lib, oracle_mod):
    """z-drop splits (mm_split_reg), inversion alignments (mm_align1_inv + ksw_ll_i16), inv mapq, secondaries."""
    ref, coff, names, seqs = parity.random_reference(41, [300000, 150000], n_repeats=60, rep_min=300, rep_max=4000, rep_div=0.03)
    c = parity.Case(emu_lib, names, seqs, cigar=True)
    try:
        _small_arenas(c)
        buf, offs = data_gen.make_sv_reads(51, ref, coff, 90)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, 4)
        assert parity.compare_hits(dev, ora) == []
        f = ora.hits["flags"]
        assert ((f & 8) > 0).sum() > 5 and ((f & 2) > 0).sum() > 2   # splits and inversions really occur
    finally:
        c.close()


def test_cigar_mode_hifi_scoring(emu_lib, oracle_mod):
    """map-hifi (k=19, w=19; a=1 b=4 q=6 e=2 q2=26 e2=1): the packed two-cells-per-register DP under the second
    scoring set, left- and right-aligned extensions and gap fills, bit-exact CIGARs."""
    ref, coff, names, seqs = parity.random_reference(23, [120000])
    c = parity.Case(emu_lib, names, seqs, preset="map-hifi", cigar=True)
    try:
        _small_arenas(c)
        buf, offs, _ = data_gen.make_reads(24, ref, coff, 24, 1500, 4000, p_sub=0.004, p_ins=0.003, p_del=0.003)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, 4)
        assert parity.compare_hits(dev, ora) == []
        assert len(dev.cigar) == len(ora.cigar) > 100
    finally:
        c.close()


def test_cigar_mode_ambiguous_bases_and_long_gaps(emu_lib, oracle_mod):
    """N bases in reads and reference (the sc_ambi score and the n_ambi counts of mm_update_extra) and reads with
    30-120 base insertions / deletions (the long-gap x2/y2 states of ksw_extd2): CIGAR, NM, dp_max, mapq bit-exact."""
    import numpy as np
    rng = np.random.default_rng(77)
    ref, coff, names, seqs = parity.random_reference(63, [150000])
    refb = bytearray(seqs[0].encode() if isinstance(seqs[0], str) else seqs[0])
    for p0 in rng.integers(1000, len(refb) - 1000, 40):
        refb[p0:p0 + int(rng.integers(1, 6))] = b"N" * 5
    seqs = [bytes(refb[:len(seqs[0])]).decode()]
    c = parity.Case(emu_lib, names, seqs, cigar=True)
    try:
        _small_arenas(c)
        buf, offs, _ = data_gen.make_reads(78, ref, coff, 60, 800, 3000)
        buf = np.array(np.frombuffer(bytes(buf), dtype=np.uint8))
        offs = np.asarray(offs)
        buf[rng.integers(0, len(buf), len(buf) // 150)] = ord("N")
        pieces = []
        for i in range(len(offs) - 1):                     # a long deletion or insertion in every other read
            r = buf[offs[i]:offs[i + 1]]
            if i % 2 == 0 and len(r) > 600:
                cut = int(rng.integers(200, len(r) - 300)); g = int(rng.integers(30, 120))
                r = np.concatenate([r[:cut], r[cut + g:]]) if i % 4 == 0 else np.concatenate([r[:cut], rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), g), r[cut:]])
            pieces.append(r)
        offs2 = np.zeros(len(pieces) + 1, dtype=np.uint64)
        offs2[1:] = np.cumsum([len(x) for x in pieces])
        buf2 = np.concatenate(pieces)
        dev = c.aligner.map_batch(buf2, offs2)
        ora = c.oracle.map_batch(buf2, offs2, 4)
        assert parity.compare_hits(dev, ora) == []
        assert len(dev.cigar) == len(ora.cigar) > 500
        assert any(((int(x) & 15) in (1, 2)) and (int(x) >> 4) >= 30 for x in ora.cigar)   # long gaps really occur
    finally:
        c.close()


@pytest.mark.parametrize("ov", [dict(sc_ambi=0), dict(sc_ambi=3, a=3, b=5), dict(q=0, e=3, q2=10, e2=1), dict(a=1, b=9, q=16, e=2, q2=41, e2=1)],
                         ids=["sc_ambi0", "sc_ambi3", "free_gap_open", "asm_like"])
def test_fill_score_from_cigar_under_odd_scorings(emu_lib, oracle_mod, ov):
    """The gap-fill kernel re-derives H(tlen-1, qlen-1) from the walked CIGAR: matches, mismatches and N pairs as the DP
    scored them (an N pair counts -e2 inside ksw_extd2 when sc_ambi is 0, not 0), each gap at the cheaper of the two
    affine costs.  That has to hold for every scoring a caller can pass (src/lib.rs:360-385): zero gap-open (adjacent
    gaps of the two kinds tie), sc_ambi 0 and large, a long_thres far from map-ont's.  dp_score, dp_max and mapq of every
    hit depend on it."""
    import numpy as np
    rng = np.random.default_rng(5)
    ref, coff, names, seqs = parity.random_reference(64, [120000])
    refb = bytearray(seqs[0].encode() if isinstance(seqs[0], str) else seqs[0])
    for p0 in rng.integers(1000, len(refb) - 1000, 60):
        refb[p0:p0 + 4] = b"NNNN"
    seqs = [bytes(refb).decode()]
    c = parity.Case(emu_lib, names, seqs, cigar=True, overrides=ov)
    try:
        _small_arenas(c)
        buf, offs, _ = data_gen.make_reads(79, ref, coff, 40, 800, 3000)
        buf = data_gen.sprinkle_n(buf, 9, 0.002)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, 4)
        assert parity.compare_hits(dev, ora) == []
        assert dev.stats["n_cell_fill"] > 0 and len(ora.hits) >= 30
    finally:
        c.close()


def test_cigar_mode_fixture_map_one(emu_lib, oracle_mod):
    """`map_one` through the product's kernel source: 1 hit, 0..400, 400M (src/lib.rs:1094-1106)."""
    c = parity.Case(emu_lib, None, None, mmi=MMI, cigar=True)
    try:
        _small_arenas(c)
        seqs = [c.oracle.seq(n) for n in c.oracle.seq_names]
        buf, offs = oracle_mod.pack_reads(seqs)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, 1)
        assert parity.compare_hits(dev, ora) == []
        assert [(int(h["rs"]), int(h["re"]), int(h["mapq"])) for h in dev.hits] == [(0, 400, 60)] * 4
        assert all([(int(x) >> 4, int(x) & 15) for x in dev.hit_cigar(h)] == [(400, 0)] for h in dev.hits)
    finally:
        c.close()


def test_two_stream_halves_equal_oracle(emu_lib, oracle_mod):
    """dual_min = 2 sends even a small chunk through the two-stream split of expand..re-chain (interleaved halves,
    separate work counters and deferred-read lists); repeats, chimeras and equal-key reads included."""
    ref, coff, names, seqs = parity.random_reference(41, [300000, 150000], n_repeats=60, rep_min=300, rep_max=4000, rep_div=0.03)
    c = parity.Case(emu_lib, names, seqs)
    try:
        c.aligner.set("dual_stream", 1)
        c.aligner.set("dual_min", 2)
        c.aligner.set("sort_small_max", 300)
        b1, o1 = data_gen.make_sv_reads(52, ref, coff, 120)
        b2, o2 = oracle_mod.pack_reads(_dup_reads(ref[:300000], 9, 53))
        buf = np.concatenate([b1, b2]); offs = np.concatenate([o1, o2[1:] + o1[-1]])
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, 4)
        assert parity.compare_stats(dev, ora) == [] and parity.compare_hits(dev, ora) == []
    finally:
        c.aligner.set("sort_small_max", 512)
        c.close()


def test_long_reads_use_large_tiles(emu_lib, oracle_mod):
    """Reads whose anchors exceed the small shared-memory tiles (sort: 2048 records) take the
    deferred large-tile passes; results must not change."""
    ref, coff, names, seqs = parity.random_reference(81, [200000])
    c = parity.Case(emu_lib, names, seqs)
    try:
        buf, offs, _ = data_gen.make_reads(82, ref, coff, 6, 20000, 60000, p_sub=0.01, p_ins=0.005, p_del=0.005)
        buf2, offs2, _ = data_gen.make_reads(83, ref, coff, 20, 300, 3000)
        buf = np.concatenate([buf2, buf]); offs = np.concatenate([offs2, offs[1:] + offs2[-1]])
        dev, stage_diffs = parity.compare_stages(c, buf, offs)
        ora = c.oracle.map_batch(buf, offs, 4)
        assert ora.stats["n_anchor"] > 6 * 2048
        assert stage_diffs == [] and parity.compare_stats(dev, ora) == [] and parity.compare_hits(dev, ora) == []
    finally:
        c.close()


def _dup_reads(ref, n, seed):
    """Reads with a tandem duplication: two query minimizers hit the SAME target position, i.e. equal anchor keys."""
    rs = np.random.RandomState(seed)
    comp = {65: 84, 67: 71, 71: 67, 84: 65}
    out = []
    for _ in range(n):
        a = int(rs.randint(0, len(ref) - 9000))
        l1, d = int(rs.randint(1500, 6000)), int(rs.randint(200, 1200))
        s = np.concatenate([ref[a:a + l1], ref[a + d:a + l1]])
        if rs.rand() < 0.5:
            s = np.array([comp[int(c)] for c in s[::-1]], dtype=np.uint8)
        out.append(s.tobytes().decode())
    return out


def test_equal_anchor_keys_replay_upstream_order(emu_lib, oracle_mod):
    """Equal anchor keys above 64 anchors: the order left by upstream's unstable radix sort reaches the
    chaining DP, so the sort stage must reproduce it (sort.cu tie path)."""
    ref, coff, names, seqs = parity.random_reference(91, [120000])
    c = parity.Case(emu_lib, names, seqs)
    try:
        buf, offs = oracle_mod.pack_reads(_dup_reads(ref, 24, 92))
        n_tie = 0
        for i in range(len(offs) - 1):
            x = c.oracle.trace(buf[int(offs[i]):int(offs[i + 1])].tobytes())["a_sorted"]["x"]
            n_tie += int(len(x) > 64 and (np.diff(x) == 0).any())
        assert n_tie >= 20
        dev, stage_diffs = parity.compare_stages(c, buf, offs)
        ora = c.oracle.map_batch(buf, offs, 4)
        assert stage_diffs == [] and parity.compare_stats(dev, ora) == [] and parity.compare_hits(dev, ora) == []
    finally:
        c.close()


def test_radix_sort_pass_for_every_read(emu_lib, oracle_mod):
    """sort_small_max = 0 sends every read through the CTA radix sort (the path of reads with > 512 anchors),
    including its equal-key replay; two contigs so that the linear target coordinate spans contigs."""
    ref, coff, names, seqs = parity.random_reference(93, [90000, 50000])
    c = parity.Case(emu_lib, names, seqs)
    try:
        c.aligner.set("sort_small_max", 0)
        b1, o1 = oracle_mod.pack_reads(_dup_reads(ref[:90000], 10, 94))
        b2, o2, _ = data_gen.make_reads(95, ref, coff, 60, 200, 6000)
        buf = np.concatenate([b1, b2]); offs = np.concatenate([o1, o2[1:] + o1[-1]])
        dev, stage_diffs = parity.compare_stages(c, buf, offs)
        ora = c.oracle.map_batch(buf, offs, 4)
        assert stage_diffs == [] and parity.compare_stats(dev, ora) == [] and parity.compare_hits(dev, ora) == []
    finally:
        c.aligner.set("sort_small_max", 512)
        c.close()


def _random_hit_case(lib, oracle_mod, ref_len, seed):
    """Short k-mers on a large reference: most anchors of a read are random, isolated hits (what a human-scale
    index does to 15-mers).  A 60-copy tandem array gives some seeds more than 32 hits (the filter's per-warp path)."""
    ref, coff, names, _ = parity.random_reference(seed, [ref_len])
    ref = ref.copy()
    ref[1000000:1000000 + 60 * 180] = np.tile(ref[2000000:2000180], 60)
    seqs = [ref.tobytes()]
    c = parity.Case.__new__(parity.Case)
    import ctypes
    from mappy_rs import _mmg
    c.lib = lib
    c.io, c.mopt = _mmg.IdxOpt(), _mmg.MapOpt()
    lib.check(lib.L.mmg_set_opt(None, ctypes.byref(c.io), ctypes.byref(c.mopt)))
    c.io.k, c.io.w = 11, 5
    c.oracle = oracle_mod.Oracle(names=names, seqs=seqs, k=11, w=5)
    c.index = _mmg.Index.build(lib, c.io, names, seqs)
    c.mopt.flag = 0
    c.oracle.set_opt("flag", 0)
    for k_, v in (("max_gap", 1000), ("bw", 400), ("bw_long", 400)):
        setattr(c.mopt, k_, v)
        c.oracle.set_opt(k_, v)
    lib.check(lib.L.mmg_mapopt_update(ctypes.byref(c.mopt), c.index.h))
    assert c.mopt.mid_occ == c.oracle.get_opt("mid_occ")
    c.aligner = _mmg.DeviceAligner(lib, c.index, c.mopt)
    return c, ref, coff


def test_isolated_anchor_filter_is_exact(emu_lib, oracle_mod):
    """Reads whose anchors are mostly isolated random hits: they are dropped before the sort (n_dropped > 0) and
    every chain, region and mapq still equals the oracle's, which sorts and chains all of them."""
    c, ref, coff = _random_hit_case(emu_lib, oracle_mod, 6000000, 97)
    try:
        buf, offs, _ = data_gen.make_reads(98, ref, coff, 70, 900, 1900)   # short: few repeated 11-mer hashes per read
        b2, o2 = oracle_mod.pack_reads(_dup_reads(ref, 4, 99))             # repeated minimizers: these keep every anchor
        b3, o3 = oracle_mod.pack_reads([ref[a:a + 1500].tobytes().decode() for a in (998600, 999300, 1010200, 1010900, 2000000 - 700)])
        buf = np.concatenate([buf, b2, b3]); offs = np.concatenate([offs, o2[1:] + offs[-1], o3[1:] + offs[-1] + o2[-1]])
        dev, stage_diffs = parity.compare_stages(c, buf, offs)
        ora = c.oracle.map_batch(buf, offs, 4)
        assert dev.stats["n_dropped"] > 0.2 * ora.stats["n_anchor"]
        assert stage_diffs == [] and parity.compare_stats(dev, ora) == [] and parity.compare_hits(dev, ora) == []
        c.aligner.set("anchor_filter", 0)
        dev2 = c.aligner.map_batch(buf, offs)
        assert dev2.stats["n_dropped"] == 0 and parity.compare_hits(dev2, ora) == []
    finally:
        c.close()


@pytest.mark.parametrize("k,w", [(15, 1), (15, 2), (15, 3), (15, 7), (13, 16), (15, 31), (15, 32), (17, 19), (21, 11), (15, 40)])
def test_sketch_window_sizes(emu_lib, oracle_mod, k, w):
    """The doubling window minimum of the sketch kernel (w <= 32) decomposes w by its bits; w > 32 takes the scan.
    Minimizers (order included) must equal mm_sketch for every shape of w, with N runs, homopolymers and short reads."""
    import ctypes
    from mappy_rs import _mmg
    ref, coff, names, seqs = parity.random_reference(200 + w, [40000])
    io, mo = _mmg.IdxOpt(), _mmg.MapOpt()
    emu_lib.check(emu_lib.L.mmg_set_opt(None, ctypes.byref(io), ctypes.byref(mo)))
    io.k, io.w = k, w
    mo.flag = 0
    idx = _mmg.Index.build(emu_lib, io, names, seqs)
    emu_lib.check(emu_lib.L.mmg_mapopt_update(ctypes.byref(mo), idx.h))
    al = _mmg.DeviceAligner(emu_lib, idx, mo)
    ora = oracle_mod.Oracle(names=names, seqs=seqs, k=k, w=w)
    ora.set_opt("flag", 0)
    try:
        base = ref[3000:9000].tobytes().decode()
        reads = [base, base[:k - 1], base[:k], base[:k + w - 1], base[:k + w], base[:k + w + 1], base[:700] + "N" * 3 + base[700:1500] + "N" + base[1500:1600],
                 "A" * 300 + base[:400] + "ACAC" * 50, base[:2000].lower(), "".join(c if i % 37 else "N" for i, c in enumerate(base[:3000])), base[::-1][:1234]]
        buf, offs = oracle_mod.pack_reads(reads)
        res = al.map_batch(buf, offs, keep_handle=True)
        mx, my, moff = al.debug_dump(res.handle, 0, int(offs[-1]) + 16, len(reads))
        al.free(res.handle)
        for i, s in enumerate(reads):
            mv = ora.trace(s.encode())["mv"]      # mm_sketch followed by mm_seed_mz_flt, as the device dump
            sl = slice(int(moff[i]), int(moff[i + 1]))
            assert np.array_equal(mx[sl], mv["x"]) and np.array_equal(my[sl], mv["y"]), (k, w, i)
    finally:
        al.close()
        idx.close()
        ora.close()


def _fixture_contigs(oracle):
    return [oracle.seq(n) for n in oracle.seq_names]


def test_config0_flanked_noisy_reads_with_cigar(emu_lib, oracle_mod):
    """BASELINE.json configs[0] in mappy-rs' real mode (CIGAR on): random flank + 8 %-error substring of a contig of
    the reference's test.mmi + random flank; every field and CIGAR operation against the oracle."""
    c = parity.Case(emu_lib, None, None, mmi=MMI, cigar=True)
    try:
        buf, offs = data_gen.config0_reads(_fixture_contigs(c.oracle), 80, len_max=3000)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, 4)
        assert parity.compare_hits(dev, ora) == []
        assert len(ora.hits) >= 50 and (ora.hits["rev"] != 0).sum() > 10 and (ora.hits["nm"] > 0).sum() > 30
    finally:
        c.close()


@pytest.mark.parametrize("scoring", [(2, 4, 4, 2), (1, 4, 6, 2)])
def test_four_tuple_scoring_equals_ksw_extz2(emu_lib, oracle_mod, scoring):
    """A 4-tuple `scoring` (/root/reference/src/lib.rs:369-376) makes q2 = q, e2 = e, and upstream then runs
    ksw_extz2_sse.  The oracle restates that kernel separately (oracle/mm2o_ksw2.cpp ksw_extz2); the device runs its
    dual-gap kernel with equal gap pairs and must produce the same CIGARs and scores."""
    ref, coff, names, seqs = parity.random_reference(41, [300000, 150000], n_repeats=60, rep_min=300, rep_max=4000, rep_div=0.03)
    a, b, q, e = scoring
    c = parity.Case(emu_lib, names, seqs, cigar=True, overrides=dict(a=a, b=b, q=q, e=e, q2=q, e2=e))
    try:
        buf, offs = data_gen.make_sv_reads(57, ref, coff, 40)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, 4)
        assert parity.compare_hits(dev, ora) == []
        assert (ora.hits["n_cigar"] > 3).sum() > 15
    finally:
        c.close()


def test_cs_and_md_tags_match_oracle(emu_lib, oracle_mod):
    """cs (short form, what map_batch always returns: /root/reference/src/lib.rs:589) and MD of every hit against the
    oracle's mm_gen_cs / mm_gen_MD on noisy reads of both strands that contain N bases."""
    ref, coff, names, seqs = parity.random_reference(61, [200000])
    c = parity.Case(emu_lib, names, seqs, cigar=True)
    try:
        buf, offs, truth = data_gen.make_reads(62, ref, coff, 60, 300, 3000)
        buf = data_gen.sprinkle_n(buf, 63, 0.004)
        dev = c.aligner.map_batch(buf, offs)
        assert (dev.hits["rev"] != 0).sum() > 10 and (dev.hits["rev"] == 0).sum() > 10 and (dev.hits["n_ambi"] > 0).sum() > 10
        assert parity.compare_tags(c, buf, offs, dev, 0) == []
        assert parity.compare_tags(c, buf, offs, dev, 1) == []
    finally:
        c.close()


def test_multi_device_aligner_equals_single_device(emu_lib, oracle_mod):
    """mmg_aligner_create_multi: the index replicated on three (emulated) devices, a batch sharded by bases over them on
    three host threads and gathered in read order must give the bytes of the one-device result (and of the oracle),
    CIGAR offsets rebased, in both modes; empty and tiny batches included."""
    import subprocess
    import sys
    code = r'''
import os, sys, ctypes
import numpy as np
sys.path[:0] = [os.path.join(%(root)r, "oracle"), os.path.join(%(root)r, "tests"), os.path.join(%(root)r, "mappy-rs_b200")]
import data_gen, parity
from mappy_rs import _mmg
lib = _mmg.Lib(%(emu)r)
ref, coff, names, seqs = parity.random_reference(11, [150000, 60000])
buf, offs, _ = data_gen.make_reads(12, ref, coff, 61, 300, 5000)
for cigar in (False, True):
    c = parity.Case(lib, names, seqs, cigar=cigar)
    one = c.aligner.map_batch(buf, offs)
    multi = _mmg.DeviceAligner(lib, c.index, c.mopt, devices=[0, 1, 2])
    for n in (61, 2, 1, 0):
        sub_b, sub_o = buf[:int(offs[n])], offs[:n + 1]
        want = c.aligner.map_batch(sub_b, sub_o) if n != 61 else one
        got = multi.map_batch(sub_b, sub_o)
        assert parity.compare_hits(got, want) == [], (cigar, n)
        assert np.array_equal(got.hit_off, want.hit_off) and got.stats["n_bases"] == want.stats["n_bases"]
    ora = c.oracle.map_batch(buf, offs, 4)
    assert parity.compare_hits(multi.map_batch(buf, offs), ora) == []
    multi.close()
    c.close()
print("multi-device ok")
'''
    from conftest import ROOT, EMU_LIB
    env = dict(os.environ, MMG_EMU_DEVICES="3")
    r = subprocess.run([sys.executable, "-c", code % {"root": ROOT, "emu": EMU_LIB}], env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and "multi-device ok" in r.stdout, (r.stdout + r.stderr)[-2000:]


@pytest.mark.parametrize("preset,cigar", [("ava-ont", False), ("ava-ont", True), ("asm20", True), ("asm5", True)])
def test_other_presets_reachable_by_string(emu_lib, oracle_mod, preset, cigar):
    """Presets a caller can name through `Aligner(preset=...)` (/root/reference/src/lib.rs:334-337): ava-ont (all
    chains, no long join, w = 5; its NO_DIAG / NO_DUAL flags need a query name and are inert here) and asm5/10/20
    (MM_F_RMQ: chaining by range-minimum query instead of mm_lchain_dp, best_n = 50, wide bands)."""
    ref, coff, names, seqs = parity.random_reference(41, [300000, 150000], n_repeats=60, rep_min=300, rep_max=4000, rep_div=0.03)
    c = parity.Case(emu_lib, names, seqs, cigar=cigar, preset=preset)
    try:
        if preset.startswith("asm"):
            buf, offs, _ = data_gen.make_reads(58, ref, coff, 16, 3000, 20000, p_sub=0.01, p_ins=0.003, p_del=0.003)
        else:
            buf, offs = data_gen.make_sv_reads(57, ref, coff, 30)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, 4)
        assert parity.compare_stats(dev, ora) == []
        assert parity.compare_hits(dev, ora) == []
        assert len(ora.hits) >= len(offs) - 1
    finally:
        c.close()


def test_unsupported_presets_are_rejected_not_approximated(emu_lib):
    import ctypes
    from mappy_rs import _mmg
    io, mo = _mmg.IdxOpt(), _mmg.MapOpt()
    emu_lib.check(emu_lib.L.mmg_set_opt(None, ctypes.byref(io), ctypes.byref(mo)))
    for preset in (b"sr", b"splice", b"map-pb", b"nonsense"):
        assert emu_lib.L.mmg_set_opt(preset, ctypes.byref(io), ctypes.byref(mo)) < 0


def test_scorings_outside_minimap2s_8bit_range_are_refused(emu_lib):
    """ksw_extd2_sse wraps beyond (O1+E1)+(O2+E2) <= 127 and returns nothing when b > 2*(O1+E1): such a scoring=
    (src/lib.rs:360-381) is refused at aligner creation, not answered with something else"""
    import ctypes
    from mappy_rs import _mmg
    idx = None
    for a, b, q, e, q2, e2 in ((2, 4, 60, 10, 60, 9), (2, 30, 4, 2, 24, 1), (120, 4, 4, 2, 24, 1), (0, 4, 4, 2, 24, 1)):
        io, mo = _mmg.IdxOpt(), _mmg.MapOpt()
        emu_lib.check(emu_lib.L.mmg_set_opt(None, ctypes.byref(io), ctypes.byref(mo)))
        idx = idx or _mmg.Index.open(emu_lib, MMI, io)
        emu_lib.check(emu_lib.L.mmg_mapopt_update(ctypes.byref(mo), idx.h))
        mo.flag |= 4
        mo.a, mo.b, mo.q, mo.e, mo.q2, mo.e2 = a, b, q, e, q2, e2
        with pytest.raises(RuntimeError, match="scoring"):
            _mmg.DeviceAligner(emu_lib, idx, mo)
        mo.flag &= ~4       # without CIGAR no DP kernel runs: chaining does not look at the scoring
        _mmg.DeviceAligner(emu_lib, idx, mo).close()


def _stream_all(lib, aligner, buf, offs, pieces, hit_dtype):
    """feeds the reads through mmg_submit in `pieces` calls, then collects every result with mmg_next"""
    import ctypes
    from mappy_rs import _mmg
    n = len(offs) - 1
    cuts = [n * k // pieces for k in range(pieces + 1)]
    for a, b in zip(cuts[:-1], cuts[1:]):
        o = np.ascontiguousarray(offs[a:b + 1])
        lib.check(lib.L.mmg_submit(aligner.h, buf.ctypes.data, o.ctypes.data, b - a, a))
    lib.check(lib.L.mmg_flush(aligner.h))
    got = {}
    res = _mmg.Result()
    while True:
        rc = lib.check(lib.L.mmg_next(aligner.h, ctypes.byref(res), 30000))
        if rc == 0:
            break
        hits = _mmg.np_from(res.hits, res.n_hits, hit_dtype)
        cig = [_mmg.np_from(res.cigar + 4 * int(h["cigar_off"]), int(h["n_cigar"]), np.uint32) for h in hits] if res.cigar else []
        got[int(res.read_id)] = (hits, cig)
        lib.L.mmg_result_release(aligner.h, ctypes.byref(res))
    return got


@pytest.mark.parametrize("cigar", [False, True])
def test_submit_next_streaming_equals_map_batch(emu_lib, oracle_mod, cigar):
    """mmg_submit / mmg_flush / mmg_next / mmg_result_release (the work queue + workers + result queue of
    /root/reference/src/lib.rs:541-636, 972-991 behind the C ABI): every read comes back exactly once, with the hits
    mmg_map_batch gives, whatever the submission granularity."""
    from mappy_rs import _mmg
    ref, coff, names, seqs = parity.random_reference(11, [150000, 60000])
    c = parity.Case(emu_lib, names, seqs, cigar=cigar)
    try:
        buf, offs, _ = data_gen.make_reads(12, ref, coff, 40, 300, 3000)
        want = c.aligner.map_batch(buf, offs)
        for pieces in (1, 7):
            got = _stream_all(emu_lib, c.aligner, buf, offs, pieces, _mmg.HIT_DTYPE)
            assert sorted(got) == list(range(40))
            for i in range(40):
                w = want.read_hits(i)
                h, cg = got[i]
                assert len(h) == len(w)
                for f in _mmg.HIT_DTYPE.names:
                    if f != "cigar_off":
                        assert np.array_equal(h[f], w[f]), (i, f)
                for k in range(len(w)):
                    if w["n_cigar"][k]:
                        assert np.array_equal(cg[k], want.hit_cigar(w[k]))
        import ctypes
        res = _mmg.Result()
        assert emu_lib.L.mmg_next(c.aligner.h, ctypes.byref(res), 10) == 0     # nothing outstanding
    finally:
        c.close()
