"""The oracle against every fixture the reference's own tests hold for this path
(SURVEY.md section 8c): resources/test/test.{fa,mmi} copied to tests/golden/ and
the assertions of /root/reference/src/lib.rs:1040-1106 and tests/python_test.py."""
import os

import numpy as np

from conftest import GOLDEN

MMI = os.path.join(GOLDEN, "test.mmi")
FA = os.path.join(GOLDEN, "test.fa")

ENTEROCOCCUS = (
    "AGAGCAGGTAGGATCGTTGAAAAAAGAGTACTCAGGATTCCATTCAACTTTTACTGATTTGAAGCGTACTGTTTATGGCC"
    "AAGAATATTTACGTCTTTACAACCAATACGCAAAAAAAGGTTCATTGAGTTTGGTTGTGATTTGATGAAAATTACTGAGA"
    "ATAACAGGATTATTAAGCTGATTGATGAACTAAATCAGCTTAATAAATATTCTTTGCAGATAGGAATATTTGGGGAAAAT"
    "GATTCTTTTATGGCGATGTTGGCCCAAGTTCATGAATTTGGGGTGACTATTCGTCCCAAAGGTCGTTTTCTTGTTATACC"
    "ACTTATGAAAAAGTATAGAGGTAAAAGTCCACGTCAATTTGATTTGTTTTTTATGCAAACTAAAGAAAATCACAAGTTTT")
BACILLUS = (
    "AGAGTGAAGCCAATATTCCGATAACGATTGCTTTCATGATATCCCTCATTCTGGCATTATTTTTTTATACTATACTATTC"
    "GATATCGCACAGATCAATGGAGTCGTGAGAAAATAAACATGTTTTGCGAACCGCTATGTGTGGAAGACAAAAAATGGAGG"
    "TGAAATTGATGGAAGCAAAGACACAGGCGTACTTTTTTCAGGATGATGGCAGGATTCCGAATCACCCTGATTTTCCGCTC"
    "GTTGTGTATCAAAACGCACTCAAGGACACCGGTCAGGCAGAGCGGATCGTCAACCGGCATGGCTGGTCAAACAGCTGGTC"
    "GGGGAGTGTTTTTCCATACCATCATTATCACAGCAATACGCATGAAGTCCTGATTGCAGTTCGGGGAGAGGCTGTGATTC")


def read_fasta(path):
    out, name = [], None
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            name = line[1:].split()[0]
            out.append([name, ""])
        elif line:
            out[-1][1] += line
    return out


def test_index_properties(oracle_mod):
    o = oracle_mod.Oracle(MMI)
    assert (o.k, o.w, o.n_seq) == (15, 10, 4)            # src/lib.rs:1046-1061
    assert sorted(o.seq_names) == ["Bacillus_subtilis", "Enterococcus_faecalis", "Escherichia_coli_1", "Escherichia_coli_2"]
    assert o.seq("Bacillus_subtilis") == BACILLUS         # src/lib.rs:1078-1091
    assert o.get_opt("mid_occ") == 10


def test_mmi_equals_index_built_from_fasta(oracle_mod):
    """hash64 + mm_sketch + bucket/key split: all 280 entries of test.mmi are rebuilt from test.fa."""
    a = oracle_mod.Oracle(MMI).index_entries()
    b = oracle_mod.Oracle(FA).index_entries()
    sa = sorted(zip(a[0].tolist(), a[1].tolist()))
    sb = sorted(zip(b[0].tolist(), b[1].tolist()))
    assert len(sa) == 280 and sa == sb


def test_mmi_roundtrip_bytes(oracle_mod, tmp_path):
    """mm_idx_dump of the loaded index reproduces the file except for khash slot order."""
    o = oracle_mod.Oracle(MMI)
    out = tmp_path / "o.mmi"
    o.dump_index(out)
    assert os.path.getsize(out) == os.path.getsize(MMI) == 136470
    o2 = oracle_mod.Oracle(str(out))
    a, b = o.index_entries(), o2.index_entries()
    assert np.array_equal(np.sort(a[0]), np.sort(b[0])) and o2.seq("Bacillus_subtilis") == BACILLUS


def test_sketch_of_contigs_matches_mmi(oracle_mod):
    o = oracle_mod.Oracle(MMI)
    mz, y = o.index_entries()
    want = set(zip(mz.tolist(), y.tolist()))
    got = set()
    for rid, (name, seq) in enumerate(read_fasta(FA)):
        x, yy = oracle_mod.sketch(seq, 10, 15, rid=rid)
        got |= set(zip((x >> np.uint64(8)).tolist(), yy.tolist()))
    assert got == want


def test_map_one_mapping_only(oracle_mod):
    """Chain-level coordinates predicted in SURVEY.md appendix E (no CIGAR): 1..394, 75 anchors, score 393."""
    o = oracle_mod.Oracle(MMI)
    o.set_opt("flag", 0)
    r = o.map(ENTEROCOCCUS)
    assert len(r.hits) == 1
    h = r.hits[0]
    assert (h["rid"], h["rs"], h["re"], h["qs"], h["qe"], h["cnt"], h["score"], h["mapq"], h["rev"]) == (1, 1, 394, 1, 394, 75, 393, 60, 0)


def test_map_one_reference_assertions(oracle_mod):
    """`map_one` (src/lib.rs:1094-1106, tests/python_test.py:124-137): exactly one mapping, 0..400.
    mappy-rs always aligns (flag |= 4, src/lib.rs:339), so this pins that extension reaches both ends."""
    o = oracle_mod.Oracle(MMI)
    assert o.get_opt("flag") & 4
    r = o.map(ENTEROCOCCUS, cs=True)
    assert len(r.hits) == 1
    h = r.hits[0]
    assert (h["rs"], h["re"]) == (0, 400)
    # SURVEY.md appendix E prediction for the remaining fields
    assert (h["qs"], h["qe"], h["rev"], h["mapq"], h["mlen"], h["blen"], h["nm"], h["dp_max"], h["is_primary"]) == (0, 400, 0, 60, 400, 400, 0, 800, 1)
    assert [(int(c) >> 4, int(c) & 15) for c in r.hit_cigar(h)] == [(400, 0)]
    assert r.cs == b":400"


def test_all_fixture_contigs_and_reverse_complements(oracle_mod):
    """tests/python_test.py:167-178 maps the four contigs 10x each and expects one result per read."""
    o = oracle_mod.Oracle(MMI)
    comp = str.maketrans("ACGT", "TGCA")
    for name, seq in read_fasta(FA):
        for s, rev in ((seq, 0), (seq.translate(comp)[::-1], 1)):
            r = o.map(s)
            assert len(r.hits) == 1
            h = r.hits[0]
            assert (o.seq_names[h["rid"]], h["rs"], h["re"], h["rev"], h["mapq"]) == (name, 0, 400, rev, 60)


def test_ksw_sse_blocks_equal_scalar_restatement(oracle_mod):
    """The oracle evaluates ksw_extd2's 16-lane blocks with the SSE2/SSE4.1 intrinsics upstream uses (what SIMDe maps
    to on x86-64).  The byte-by-byte restatement of the same blocks is kept as a self-check: both must give the same
    hits and CIGARs (left- and right-aligned gaps, exact and approximate maximum, z-drop splits, inversions)."""
    import numpy as np
    import data_gen
    import parity
    ref, coff, names, seqs = parity.random_reference(41, [300000, 150000], n_repeats=60, rep_min=300, rep_max=4000, rep_div=0.03)
    buf, offs = data_gen.make_sv_reads(57, ref, coff, 120)
    L = oracle_mod.lib()
    for preset in (None, "map-hifi"):
        o = oracle_mod.Oracle(names=names, seqs=seqs, preset=preset)
        o.set_opt("flag", 4)
        try:
            L.mm2o_set_ksw_scalar(1)
            a = o.map_batch(buf, offs, 8)
            L.mm2o_set_ksw_scalar(0)
            b = o.map_batch(buf, offs, 8)
        finally:
            L.mm2o_set_ksw_scalar(0)
            o.close()
        assert len(a.hits) > 100 and ((a.hits["flags"] & 8) > 0).sum() > 10
        assert all(np.array_equal(a.hits[f], b.hits[f]) for f in a.hits.dtype.names) and np.array_equal(a.cigar, b.cigar)
