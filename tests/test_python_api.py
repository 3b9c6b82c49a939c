"""The reference's own Python tests (/root/reference/tests/python_test.py) against the `mappy_rs` mirror.

Same assertions, same messages.  On CPU they run over the SIMT-emulated test build of the kernels
(`emu` parametrisation); on the B200 box the `gpu` parametrisation runs them over the product library."""
import copy
import os
from itertools import repeat

import pytest

from conftest import EMU_LIB, GOLDEN, PRODUCT_LIB

MMI_FILE = os.path.join(GOLDEN, "test.mmi")
FA_FILE = os.path.join(GOLDEN, "test.fa")
TUNE = {"tb_cap": 1 << 26, "cigar_cap": 1 << 22, "jobs_cap": 1 << 14, "chunk_bases": 1 << 20, "anchor_cap": 1 << 18,
        "chunk_reads": 1 << 16, "regs_cap": 1 << 17, "big_per_warp": 1 << 18}


def read_fasta(fh):
    for line in fh:
        if line.startswith(">"):
            name = line[1:].strip()
            break
    fa_lines = []
    for line in fh:
        if line.startswith(">"):
            yield name, "".join(fa_lines)
            fa_lines = []
            name = line[1:].strip()
            continue
        fa_lines.append(line.strip())
    yield name, "".join(fa_lines)


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def al(request):
    import mappy_rs
    from mappy_rs import _mmg
    from mappy_rs import aligner as aligner_mod
    if request.param == "emu":
        # the emulator maps ~10^2 reads/s: the 50000-entry work queue of lib.rs:429 is scaled to 500 entries and
        # the 100000-read batches of the reference tests to 1000 reads; the `gpu` run uses the real sizes
        request.getfixturevalue("emu_lib")
        monkeypatch = request.getfixturevalue("monkeypatch")
        monkeypatch.setattr(aligner_mod, "_WORK_QUEUE_CAP", 500)
        a = mappy_rs.Aligner(MMI_FILE, _lib=_mmg.Lib(EMU_LIB), _tune=TUNE)
        a._big_n = 1000
    else:
        a = mappy_rs.Aligner(MMI_FILE)
        a._big_n = 100000
    yield a
    a.close()


@pytest.fixture
def fasta_list():
    with open(FA_FILE, "rt") as fh:
        seqs = [s for _, s in read_fasta(fh)]
    return [{"id": i, "seq": seq} for i, seq in enumerate(copy.copy(s) for _ in range(10) for s in seqs)]


@pytest.fixture
def fasta_iter(fasta_list):
    return iter(fasta_list)


@pytest.fixture
def fasta_tuple(fasta_list):
    return tuple(fasta_list)


@pytest.fixture
def fasta_generator(fasta_list):
    return (item for item in fasta_list)


@pytest.fixture
def fasta(request):
    return request.getfixturevalue(request.param)


def test_test(al):
    assert al


def test_property_k(al):
    assert al.k == 15


def test_property_n_seq(al):
    assert al.n_seq == 4


def test_property_w(al):
    assert al.w == 10


def test_property_seq_names(al):
    expected = ["Bacillus_subtilis", "Enterococcus_faecalis", "Escherichia_coli_1", "Escherichia_coli_2"]
    seq_names = al.seq_names
    seq_names.sort()
    assert seq_names == expected


def test_get_seq(al):
    expected = (
        "AGAGTGAAGCCAATATTCCGATAACGATTGCTTTCATGATATCCCTCATTCTGGCATTATTTTTTTATA"
        "CTATACTATTCGATATCGCACAGATCAATGGAGTCGTGAGAAAATAAACATGTTTTGCGAACCGCTATG"
        "TGTGGAAGACAAAAAATGGAGGTGAAATTGATGGAAGCAAAGACACAGGCGTACTTTTTTCAGGATGAT"
        "GGCAGGATTCCGAATCACCCTGATTTTCCGCTCGTTGTGTATCAAAACGCACTCAAGGACACCGGTCAG"
        "GCAGAGCGGATCGTCAACCGGCATGGCTGGTCAAACAGCTGGTCGGGGAGTGTTTTTCCATACCATCAT"
        "TATCACAGCAATACGCATGAAGTCCTGATTGCAGTTCGGGGAGAGGCTGTGATTC")
    assert al.seq("Bacillus_subtilis") == expected
    assert al.seq("Bacillus_subtilis", 10, 20) == expected[10:20]
    assert al.seq("nope") is None and al.seq("Bacillus_subtilis", 500, 600) is None


def test_map_one(al):
    mappings = al.map(
        "AGAGCAGGTAGGATCGTTGAAAAAAGAGTACTCAGGATTCCATTCAACTTTTACTGATTTGAAGCGTAC"
        "TGTTTATGGCCAAGAATATTTACGTCTTTACAACCAATACGCAAAAAAAGGTTCATTGAGTTTGGTTGT"
        "GATTTGATGAAAATTACTGAGAATAACAGGATTATTAAGCTGATTGATGAACTAAATCAGCTTAATAAA"
        "TATTCTTTGCAGATAGGAATATTTGGGGAAAATGATTCTTTTATGGCGATGTTGGCCCAAGTTCATGAA"
        "TTTGGGGTGACTATTCGTCCCAAAGGTCGTTTTCTTGTTATACCACTTATGAAAAAGTATAGAGGTAAA"
        "AGTCCACGTCAATTTGATTTGTTTTTTATGCAAACTAAAGAAAATCACAAGTTTT",
        cs=True,
    )
    assert len(mappings) == 1
    mapping = mappings[0]
    assert mapping.target_start == 0
    assert mapping.target_end == 400
    # fields beyond the reference's assertions
    assert (mapping.ctg, mapping.ctg_len, mapping.q_st, mapping.q_en, mapping.strand, mapping.mapq) == ("Enterococcus_faecalis", 400, 0, 400, 1, 60)
    assert mapping.cigar == [(400, 0)] and mapping.cigar_str == "400M" and mapping.NM == 0 and mapping.cs == ":400" and mapping.is_primary
    assert str(mapping) == "0\t400\t+\tEnterococcus_faecalis\t400\t0\t400\t400\t400\t60\ttp:A:P\tcg:Z:400M"


def test_map_batch_100000(al, fasta_iter):
    al.enable_threading(4)
    iter_ = repeat(next(fasta_iter), al._big_n)
    mappings = al.map_batch(iter_, back_off=True)
    n = 0
    for res in mappings:
        n += 1
    assert n == al._big_n


def test_map_batch_100000_no_backoff(al, fasta_iter):
    al.enable_threading(4)
    iter_ = repeat(next(fasta_iter), al._big_n)
    with pytest.raises(RuntimeError) as excinfo:
        mappings = al.map_batch(iter_, back_off=False)
        n = 0
        for res in mappings:
            n += 1
    assert "Internal error adding data to work queue, without backoff" in str(excinfo)
    assert "Is your fastq batch larger than 50000? Perhaps try `map_batch` with back_off=True?" in str(excinfo)


@pytest.mark.parametrize("fasta", ["fasta_iter", "fasta_list", "fasta_tuple", "fasta_generator"], indirect=True)
def test_map_batch(al, fasta):
    al.enable_threading(2)
    mappings = al.map_batch(fasta)
    n = 0
    for res in mappings:
        n += 1
        hits, data = res
        assert len(hits) == 1 and hits[0].r_st == 0 and hits[0].r_en == 400 and hits[0].cs == ":400" and "id" in data
    assert n == 40


def test_map_batch_needs_threading(al, fasta_list):
    with pytest.raises(RuntimeError) as excinfo:
        al.map_batch(fasta_list)
    assert "Multi threading not enabled on this instance. Please call `.enable_threading()`" in str(excinfo.value)


def test_map_batch_fail_dict_single(al, fasta_iter):
    fasta = next(fasta_iter)
    al.enable_threading(2)
    with pytest.raises(TypeError) as excinfo:
        _ = al.map_batch(fasta)
    assert "Unsupported batch type, pass a list, iter, generator or tuple" in str(excinfo)


def test_map_batch_fail_dict_many(al, fasta_iter):
    fasta = {i: dct for i, dct in enumerate(fasta_iter)}
    al.enable_threading(2)
    with pytest.raises(TypeError) as excinfo:
        _ = al.map_batch(fasta)
    assert "Unsupported batch type, pass a list, iter, generator or tuple" in str(excinfo)


def test_map_batch_fail_list_str(al, fasta_iter):
    fasta = [dct["seq"] for dct in fasta_iter]
    al.enable_threading(2)
    with pytest.raises(TypeError) as excinfo:
        _ = al.map_batch(fasta)
    assert "Element in iterable is not a dictionary" in str(excinfo.value)


def test_map_batch_fail_no_seq_key(al, fasta_iter):
    fasta = [{"SEQ": dct["seq"]} for dct in fasta_iter]
    al.enable_threading(2)
    with pytest.raises(KeyError) as excinfo:
        _ = al.map_batch(fasta)
    assert "AHHH Key 🗝️  not found in iterated dictionary" in str(excinfo)


def test_map_batch_fail_seq_not_str(al, fasta_iter):
    fasta = [{"seq": dct["seq"].encode()} for dct in fasta_iter]
    al.enable_threading(2)
    with pytest.raises(ValueError) as excinfo:
        _ = al.map_batch(fasta)
    assert "`seq` must be a string" in str(excinfo)


def test_map_batch_fail_exhausted_iter(al, fasta_iter):
    _ = list(fasta_iter)
    al.enable_threading(2)
    mappings = al.map_batch(fasta_iter)
    assert len(list(mappings)) == 0


def test_constructor_errors(emu_lib):
    import mappy_rs
    with pytest.raises(RuntimeError, match="Did not create or open an index"):
        mappy_rs.Aligner(_lib=emu_lib)
    with pytest.raises(NotImplementedError):
        mappy_rs.Aligner(MMI_FILE, seq="ACGT", _lib=emu_lib)
    with pytest.raises(NotImplementedError):
        mappy_rs.Aligner(MMI_FILE, fn_idx_out="x.mmi", _lib=emu_lib)
    a = mappy_rs.Aligner(MMI_FILE, _lib=emu_lib, _tune=TUNE)
    with pytest.raises(NotImplementedError, match="Using `seq2` is not implemented"):
        a.map("ACGT", seq2="ACGT")
    assert a.map_no_op("ACGT")[0].ctg == "Hello"
    a.close()
