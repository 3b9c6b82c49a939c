"""Device index construction (index_dev.cu) against the golden .mmi of the
reference (resources/test/test.mmi <- test.fa), the oracle's builder and the host builder: same entries, same
mid_occ, same .mmi bytes.  CPU: the kernel source under the SIMT emulator (its library sort is replaced by
std::sort there); GPU: the product library."""
import ctypes
import os

import numpy as np
import pytest

import data_gen
import parity
from conftest import GOLDEN
from mappy_rs import _mmg

MMI = os.path.join(GOLDEN, "test.mmi")
FA = os.path.join(GOLDEN, "test.fa")


def _entries(ix):
    mz, y = ix.entries()
    o = np.lexsort((y, mz))
    return mz[o], y[o]


def _opts(lib, preset=None):
    io, mo = _mmg.IdxOpt(), _mmg.MapOpt()
    lib.check(lib.L.mmg_set_opt(None, ctypes.byref(io), ctypes.byref(mo)))
    if preset:
        lib.check(lib.L.mmg_set_opt(preset.encode(), ctypes.byref(io), ctypes.byref(mo)))
    return io, mo


def _host_build(lib, io, names, seqs):
    os.environ["MMG_HOST_INDEX_BUILD"] = "1"
    try:
        return _mmg.Index.build(lib, io, names, seqs)
    finally:
        del os.environ["MMG_HOST_INDEX_BUILD"]


def _check_golden(lib, tmp_path):
    io, mo = _opts(lib)
    dev = _mmg.Index.open(lib, FA, io)          # FASTA -> device build
    gold = _mmg.Index.open(lib, MMI, io)        # the reference's own index file
    try:
        a, b = _entries(dev), _entries(gold)
        assert len(a[0]) == 280 and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert (dev.k, dev.w, dev.n_seq) == (15, 10, 4)
        assert np.array_equal(dev.getseq(1, 0, 400), gold.getseq(1, 0, 400))
        lib.check(lib.L.mmg_mapopt_update(ctypes.byref(mo), dev.h))
        assert mo.mid_occ == 10
        p1, p2 = tmp_path / "dev.mmi", tmp_path / "gold.mmi"
        dev.dump(p1), gold.dump(p2)
        assert open(p1, "rb").read() == open(p2, "rb").read() and os.path.getsize(p1) == os.path.getsize(MMI)
    finally:
        dev.close(), gold.close()


def _check_vs_host(lib, oracle_mod, preset, lens, seed, n_repeats):
    """Contigs longer than a segment, N runs, repeats (multi-occurrence keys), a contig shorter than k."""
    ref, coff, names = data_gen.make_reference(seed, lens, n_repeats=n_repeats, rep_min=200, rep_max=3000, rep_div=0.01)
    ref = ref.copy()
    rs = np.random.RandomState(seed)
    for _ in range(12):                                   # ambiguous bases, some across segment borders
        p = int(rs.randint(0, len(ref) - 50))
        ref[p:p + int(rs.randint(1, 40))] = ord("N")
    for b in (32768, 65536):
        if b + 4 < len(ref):
            ref[b - 3:b + 2] = ord("N")
    seqs = [ref[int(coff[i]):int(coff[i + 1])].tobytes() for i in range(len(names))]
    io, mo = _opts(lib, preset)
    dev = _mmg.Index.build(lib, io, names, seqs)
    host = _host_build(lib, io, names, seqs)
    try:
        a, b = _entries(dev), _entries(host)
        assert len(a[0]) == len(b[0]) > 1000
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        o = oracle_mod.Oracle(names=names, seqs=seqs, preset=preset)
        mz, y = o.index_entries()
        oo = np.lexsort((y, mz))
        assert np.array_equal(a[0], mz[oo]) and np.array_equal(a[1], y[oo])
        mo2 = _mmg.MapOpt.from_buffer_copy(mo)
        lib.check(lib.L.mmg_mapopt_update(ctypes.byref(mo), dev.h))
        lib.check(lib.L.mmg_mapopt_update(ctypes.byref(mo2), host.h))
        assert mo.mid_occ == mo2.mid_occ == o.get_opt("mid_occ")
        for rid in range(len(names)):
            n = dev.seq_len(rid)
            assert np.array_equal(dev.getseq(rid, 0, n), host.getseq(rid, 0, n))
        o.close()
    finally:
        dev.close(), host.close()


def test_emu_device_build_matches_reference_mmi(emu_lib, tmp_path):
    _check_golden(emu_lib, tmp_path)


def test_emu_device_build_matches_host_and_oracle(emu_lib, oracle_mod):
    _check_vs_host(emu_lib, oracle_mod, None, [70000, 33000, 9, 40000], 7, 25)
    _check_vs_host(emu_lib, oracle_mod, "map-hifi", [70000, 12000], 8, 10)


@pytest.mark.gpu
def test_gpu_device_build_matches_reference_mmi(gpu_lib, tmp_path):
    _check_golden(gpu_lib, tmp_path)


@pytest.mark.gpu
def test_gpu_device_build_matches_host_and_oracle(gpu_lib, oracle_mod):
    _check_vs_host(gpu_lib, oracle_mod, None, [3000000, 700000, 9, 1200000], 7, 200)
    _check_vs_host(gpu_lib, oracle_mod, "map-hifi", [2000000, 300000], 8, 50)


def test_gzip_fasta_fastq_and_multipart_mmi(emu_lib, tmp_path):
    """mm_idx_reader_open reads FASTA / FASTQ, plain or gzip-compressed, and `Aligner` keeps only the FIRST part of a
    multi-part .mmi (/root/reference/src/lib.rs:398-412)."""
    import gzip
    import shutil
    io, _ = _opts(emu_lib)
    gz = str(tmp_path / "test.fa.gz")
    with open(FA, "rb") as f, gzip.open(gz, "wb") as g:
        shutil.copyfileobj(f, g)
    recs = []
    for line in open(FA):
        if line.startswith(">"):
            recs.append([line[1:].strip(), ""])
        else:
            recs[-1][1] += line.strip()
    fq = str(tmp_path / "test.fq.gz")
    with gzip.open(fq, "wt") as g:
        for n, s in recs:
            g.write("@%s some description\n%s\n+\n%s\n" % (n, s, ">@+" * (len(s) // 3) + "I" * (len(s) % 3)))   # quality lines may start with > @ +
    mmi2 = str(tmp_path / "two_parts.mmi")
    with open(mmi2, "wb") as out:
        blob = open(MMI, "rb").read()
        out.write(blob + blob)
    want = None
    for path in (FA, gz, fq, MMI, mmi2):
        ix = _mmg.Index.open(emu_lib, path, io)
        try:
            got = (_entries(ix), ix.n_seq, [ix.seq_name(i) for i in range(ix.n_seq)], [ix.seq_len(i) for i in range(ix.n_seq)])
            if want is None:
                want = got
            assert got[1:] == want[1:] and np.array_equal(got[0][0], want[0][0]) and np.array_equal(got[0][1], want[0][1]), path
        finally:
            ix.close()
