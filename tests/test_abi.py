"""The C-ABI library exports every symbol include/mmg.h declares, and the host-side
entry points (options, index) behave like the reference interfaces they replace.
No device work here: aligner creation is exercised only by the `gpu` tests."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, PRODUCT_LIB, ROOT
from mappy_rs import _mmg

MMI = os.path.join(GOLDEN, "test.mmi")
FA = os.path.join(GOLDEN, "test.fa")


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "mmg.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mmg_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    import shutil
    import subprocess
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "mappy-rs_b200"), "-j8"], stderr=subprocess.DEVNULL)
    return _mmg.Lib(PRODUCT_LIB)


def test_exports_every_declared_symbol(lib):
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib.L, s), "libmmg.so does not export %s" % s
    assert set(_mmg.EXPORTS) <= set(syms)


def test_options_match_preset_tables(lib):
    io, mo_ = _mmg.IdxOpt(), _mmg.MapOpt()
    lib.check(lib.L.mmg_set_opt(None, ctypes.byref(io), ctypes.byref(mo_)))
    assert (io.k, io.w, io.bucket_bits) == (15, 10, 14)
    assert (mo_.bw, mo_.bw_long, mo_.max_gap, mo_.min_cnt, mo_.min_chain_score, mo_.best_n) == (500, 20000, 5000, 3, 40, 5)
    assert (mo_.a, mo_.b, mo_.q, mo_.e, mo_.q2, mo_.e2, mo_.min_dp_max) == (2, 4, 4, 2, 24, 1, 80)
    lib.check(lib.L.mmg_set_opt(b"map-hifi", ctypes.byref(io), ctypes.byref(mo_)))
    assert (io.k, io.w, mo_.max_gap, mo_.min_mid_occ, mo_.max_mid_occ, mo_.min_dp_max) == (19, 19, 10000, 50, 500, 200)
    assert lib.L.mmg_set_opt(b"no-such-preset", ctypes.byref(io), ctypes.byref(mo_)) < 0


def test_index_load_and_accessors(lib, oracle_mod):
    io, mo_ = _mmg.IdxOpt(), _mmg.MapOpt()
    lib.check(lib.L.mmg_set_opt(None, ctypes.byref(io), ctypes.byref(mo_)))
    idx = _mmg.Index.open(lib, MMI, io)
    assert (idx.k, idx.w, idx.n_seq) == (15, 10, 4)
    names = [idx.seq_name(i) for i in range(4)]
    assert sorted(names) == ["Bacillus_subtilis", "Enterococcus_faecalis", "Escherichia_coli_1", "Escherichia_coli_2"]
    rid = idx.name2id("Bacillus_subtilis")
    codes = idx.getseq(rid, 0, 400)
    o = oracle_mod.Oracle(MMI)
    assert bytes(b"ACGTN"[c] for c in codes).decode() == o.seq("Bacillus_subtilis")
    assert idx.name2id("nope") < 0 and idx.getseq(rid, 400, 500) is None
    lib.check(lib.L.mmg_mapopt_update(ctypes.byref(mo_), idx.h))
    assert mo_.mid_occ == 10
    # FASTA input builds the same table as the .mmi (mm_idx_reader_open sniffs the magic, src/lib.rs:398)
    idx2 = _mmg.Index.open(lib, FA, io)
    a, b = idx.entries(), idx2.entries()
    assert sorted(zip(a[0].tolist(), a[1].tolist())) == sorted(zip(b[0].tolist(), b[1].tolist()))
    om, oy = o.index_entries()
    assert sorted(zip(a[0].tolist(), a[1].tolist())) == sorted(zip(om.tolist(), oy.tolist()))


def test_index_dump_roundtrip(lib, tmp_path):
    io, mo_ = _mmg.IdxOpt(), _mmg.MapOpt()
    lib.check(lib.L.mmg_set_opt(None, ctypes.byref(io), ctypes.byref(mo_)))
    idx = _mmg.Index.open(lib, FA, io)
    out = tmp_path / "x.mmi"
    idx.dump(out)
    assert os.path.getsize(out) == 136470
    idx2 = _mmg.Index.open(lib, out, io)
    a, b = idx.entries(), idx2.entries()
    assert sorted(zip(a[0].tolist(), a[1].tolist())) == sorted(zip(b[0].tolist(), b[1].tolist()))


def test_no_device_is_a_loud_error(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    io, mo_ = _mmg.IdxOpt(), _mmg.MapOpt()
    lib.check(lib.L.mmg_set_opt(None, ctypes.byref(io), ctypes.byref(mo_)))
    idx = _mmg.Index.open(lib, MMI, io)
    with pytest.raises(_mmg.MmgError) as e:
        _mmg.DeviceAligner(lib, idx, mo_)
    assert e.value.code == -3 and "no CPU" in str(e.value)


def test_missing_library_is_a_loud_error(tmp_path):
    with pytest.raises(ImportError):
        _mmg.Lib(str(tmp_path / "libmmg.so"))
