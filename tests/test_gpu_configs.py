"""BASELINE.json configurations at their real sizes on the B200, against the oracle on the same seeded inputs:
configs[0] (flanked noisy reads on the reference's test.mmi, CIGAR on), configs[2] (3.1 Gb reference; mapping-only
and CIGAR on), configs[3] (400-base prefixes), configs[4] (map-hifi on the 3.1 Gb reference, CIGAR on), cs / MD tags,
and 4-tuple scoring (upstream's ksw_extz2).  Bit-exact on every hit field and CIGAR operation."""
import ctypes
import os

import numpy as np
import pytest

import data_gen
import parity
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
MMI = os.path.join(GOLDEN, "test.mmi")
NT = os.cpu_count() or 8


class _Human:
    """3.1 Gb reference (24 contigs, seed 3) open on both sides; aligners are created per option set on the shared
    device index (built once on the device, used in place)."""

    def __init__(self, lib, mo_mod, preset=None):
        from mappy_rs import _mmg
        self.lib, self.preset = lib, preset
        self.ref, self.coff, self.names = data_gen.make_reference(3, data_gen.config2_contig_lens())
        seqs = [self.ref[int(self.coff[i]):int(self.coff[i + 1])].tobytes() for i in range(len(self.names))]
        self.io, mo = _mmg.IdxOpt(), _mmg.MapOpt()
        lib.check(lib.L.mmg_set_opt(None, ctypes.byref(self.io), ctypes.byref(mo)))
        if preset:
            lib.check(lib.L.mmg_set_opt(preset.encode(), ctypes.byref(self.io), ctypes.byref(mo)))
        self.index = _mmg.Index.build(lib, self.io, self.names, seqs)
        self.oracle = mo_mod.Oracle(names=self.names, seqs=seqs, preset=preset)
        del seqs

    def case(self, cigar):
        """a parity.Case-like object (aligner + oracle with the same options)"""
        from mappy_rs import _mmg
        c = parity.Case.__new__(parity.Case)
        c.lib, c.index, c.oracle, c.io = self.lib, self.index, self.oracle, self.io
        c.mopt = _mmg.MapOpt()
        io = _mmg.IdxOpt()
        self.lib.check(self.lib.L.mmg_set_opt(None, ctypes.byref(io), ctypes.byref(c.mopt)))
        if self.preset:
            self.lib.check(self.lib.L.mmg_set_opt(self.preset.encode(), ctypes.byref(io), ctypes.byref(c.mopt)))
        c.mopt.flag = 4 if cigar else 0
        self.oracle.set_opt("flag", 4 if cigar else 0)
        self.lib.check(self.lib.L.mmg_mapopt_update(ctypes.byref(c.mopt), self.index.h))
        assert c.mopt.mid_occ == self.oracle.get_opt("mid_occ")
        c.aligner = _mmg.DeviceAligner(self.lib, self.index, c.mopt)
        return c

    def close(self):
        self.index.close()
        self.oracle.close()


@pytest.fixture(scope="module")
def human(gpu_lib, oracle_mod):
    h = _Human(gpu_lib, oracle_mod)
    yield h
    h.close()


def _primary(dev, n):
    first = dev.hit_off[:-1].astype(np.int64)
    return np.where(np.diff(dev.hit_off.astype(np.int64)) > 0, first, -1)


def _same_hits(x, y):
    return x.shape == y.shape and all(np.array_equal(x[f], y[f]) for f in x.dtype.names)


def test_config2_mapping_only_20k_reads_bit_exact(human):
    """configs[2], mapping-only: 20 000 reads against the oracle (all fields, stats), then the oracle-free properties at
    60 000 reads: reads return to their origin with mapq 60, idempotence, independence of how the batch is cut."""
    c = human.case(cigar=False)
    try:
        n = 60000
        buf, offs, truth = data_gen.make_reads(4, human.ref, human.coff, n, 1000, 10000, p_sub=0.03, p_ins=0.02, p_del=0.03)
        m = 20000
        sb, so = buf[:int(offs[m])], offs[:m + 1]
        dev = c.aligner.map_batch(sb, so)
        ora = c.oracle.map_batch(sb, so, NT)
        assert dev.stats["n_dropped"] > 0.5 * dev.stats["n_anchor"]          # the isolated-anchor filter is at work
        assert parity.compare_stats(dev, ora) == []
        assert parity.compare_hits(dev, ora) == []
        a = c.aligner.map_batch(buf, offs)
        p = _primary(a, n)
        h = a.hits[np.maximum(p, 0)]
        ok = (p >= 0) & (h["rid"] == truth[:, 0]) & (h["rev"] == (truth[:, 3] != 0)) & \
             (np.minimum(h["re"], truth[:, 2]) - np.maximum(h["rs"], truth[:, 1]) > 0.8 * (truth[:, 2] - truth[:, 1]))
        assert ok.mean() > 0.99, ok.mean()
        assert (h["mapq"][ok] == 60).mean() > 0.98
        b = c.aligner.map_batch(buf, offs)
        assert np.array_equal(a.hit_off, b.hit_off) and _same_hits(a.hits, b.hits)
        half = n // 2
        c1 = c.aligner.map_batch(buf[:int(offs[half])], offs[:half + 1])
        c2 = c.aligner.map_batch(buf[int(offs[half]):], offs[half:] - offs[half])
        assert _same_hits(a.hits, np.concatenate([c1.hits, c2.hits]))
    finally:
        c.aligner.close()


def test_config2_cigar_3k_reads_bit_exact(human):
    """configs[2] in the only mode mappy-rs runs (MM_F_CIGAR, /root/reference/src/lib.rs:339): 3 000 reads, every hit
    field, every CIGAR operation, cs and MD tags."""
    c = human.case(cigar=True)
    try:
        buf, offs, _ = data_gen.make_reads(4, human.ref, human.coff, 3000, 1000, 10000, p_sub=0.03, p_ins=0.02, p_del=0.03)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, NT)
        assert parity.compare_hits(dev, ora) == []
        assert len(dev.cigar) == len(ora.cigar) and len(dev.hits) >= 2990
        assert parity.compare_tags(c, buf, offs, dev, 0) == []
        assert parity.compare_tags(c, buf, offs, dev, 1) == []
    finally:
        c.aligner.close()


@pytest.mark.parametrize("cigar", [False, True])
def test_config3_prefix_batches_bit_exact(human, cigar):
    """configs[3]: the first 400 bases of configs[2] reads in a batch of 20 000 (readfish-style), on the 3.1 Gb reference."""
    c = human.case(cigar=cigar)
    try:
        buf, offs, _ = data_gen.make_reads(4, human.ref, human.coff, 20000, 1000, 10000, p_sub=0.03, p_ins=0.02, p_del=0.03)
        pb, po = data_gen.prefixes(buf, offs, 400)
        assert int(po[-1]) == 400 * 20000
        dev = c.aligner.map_batch(pb, po)
        ora = c.oracle.map_batch(pb, po, NT)
        assert parity.compare_stats(dev, ora) == []
        assert parity.compare_hits(dev, ora) == []
        assert len(dev.hits) > 15000
    finally:
        c.aligner.close()


def test_config0_flanked_noisy_reads_on_reference_fixture(gpu_lib, oracle_mod):
    """configs[0]: 20 000 reads of 1-10 kb = random flank + 8 %-error substring of a contig of the reference's
    test.mmi + random flank, CIGAR on (mappy-rs' mode), against the oracle; plus cs / MD."""
    c = parity.Case(gpu_lib, None, None, mmi=MMI, cigar=True)
    try:
        contigs = [c.oracle.seq(n) for n in c.oracle.seq_names]
        buf, offs = data_gen.config0_reads(contigs, 20000)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, NT)
        assert parity.compare_hits(dev, ora) == []
        assert len(ora.hits) > 15000 and (ora.hits["rev"] != 0).sum() > 5000
        sb, so = buf[:int(offs[3000])], offs[:3001]
        d2 = c.aligner.map_batch(sb, so)
        assert parity.compare_tags(c, sb, so, d2, 0) == [] and parity.compare_tags(c, sb, so, d2, 1) == []
    finally:
        c.close()


def test_cs_md_tags_noisy_reads_with_n_bases(gpu_lib, oracle_mod):
    ref, coff, names, seqs = parity.random_reference(61, [2000000])
    c = parity.Case(gpu_lib, names, seqs, cigar=True)
    try:
        buf, offs, _ = data_gen.make_reads(62, ref, coff, 3000, 300, 8000)
        buf = data_gen.sprinkle_n(buf, 63, 0.004)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, NT)
        assert parity.compare_hits(dev, ora) == []
        assert (dev.hits["rev"] != 0).sum() > 1000 and (dev.hits["n_ambi"] > 0).sum() > 1000
        assert parity.compare_tags(c, buf, offs, dev, 0) == []
        assert parity.compare_tags(c, buf, offs, dev, 1) == []
    finally:
        c.close()


@pytest.mark.parametrize("scoring", [(2, 4, 4, 2), (1, 4, 6, 2), (4, 8, 8, 4)])
def test_four_tuple_scoring_equals_ksw_extz2(gpu_lib, oracle_mod, scoring):
    """4-tuple `scoring` (/root/reference/src/lib.rs:369-376): upstream dispatches to ksw_extz2_sse, restated
    separately in the oracle; the device's dual-gap kernel with equal gap pairs must agree on every CIGAR."""
    ref, coff, names, seqs = parity.random_reference(41, [3000000, 1500000], n_repeats=600, rep_min=300, rep_max=6000, rep_div=0.03)
    a, b, q, e = scoring
    c = parity.Case(gpu_lib, names, seqs, cigar=True, overrides=dict(a=a, b=b, q=q, e=e, q2=q, e2=e))
    try:
        buf, offs = data_gen.make_sv_reads(51, ref, coff, 1500, 400, 6000)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, NT)
        assert parity.compare_hits(dev, ora) == []
        f = ora.hits["flags"]
        assert ((f & 8) > 0).sum() > 50 and ((f & 2) > 0).sum() > 10
        buf2, offs2, _ = data_gen.make_reads(52, ref, coff, 1500, 1000, 8000)
        dev2 = c.aligner.map_batch(buf2, offs2)
        ora2 = c.oracle.map_batch(buf2, offs2, NT)
        assert parity.compare_hits(dev2, ora2) == []
    finally:
        c.close()


@pytest.mark.parametrize("preset", ["ava-ont", "asm20", "asm10", "asm5"])
def test_other_presets_reachable_by_string(gpu_lib, oracle_mod, preset):
    """ava-ont and the asm presets (MM_F_RMQ chaining) with CIGAR, against the oracle."""
    ref, coff, names, seqs = parity.random_reference(41, [3000000, 1500000], n_repeats=600, rep_min=300, rep_max=6000, rep_div=0.03)
    c = parity.Case(gpu_lib, names, seqs, cigar=True, preset=preset)
    try:
        if preset.startswith("asm"):
            buf, offs, _ = data_gen.make_reads(58, ref, coff, 300, 5000, 60000, p_sub=0.01, p_ins=0.003, p_del=0.003)
        else:
            buf, offs = data_gen.make_sv_reads(51, ref, coff, 1500, 400, 6000)
        dev = c.aligner.map_batch(buf, offs)
        ora = c.oracle.map_batch(buf, offs, NT)
        assert parity.compare_stats(dev, ora) == []
        assert parity.compare_hits(dev, ora) == []
    finally:
        c.close()
