"""The multi-device aligner (mmg_aligner_create_multi) on real GPUs: index built on device 0 and replicated to the
others over NVLink with peer copies, batches sharded by bases, results gathered in read order.  Needs >= 2 GPUs
(`gpurun --gpus 2 ...`); skipped on a one-GPU box."""
import os

import numpy as np
import pytest

import data_gen
import parity

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least two GPUs")
@pytest.mark.parametrize("cigar", [False, True])
def test_n_device_aligner_equals_one_device(gpu_lib, oracle_mod, cigar):
    from mappy_rs import _mmg
    ref, coff, names = data_gen.config1_reference()
    c = parity.Case(gpu_lib, names, [ref.tobytes()], cigar=cigar)
    try:
        buf, offs, _ = data_gen.config1_reads(ref, coff, 6000 if cigar else 30000)
        one = c.aligner.map_batch(buf, offs)
        devs = list(range(_n_gpus()))
        multi = _mmg.DeviceAligner(gpu_lib, c.index, c.mopt, devices=devs)
        try:
            got = multi.map_batch(buf, offs)
            assert parity.compare_hits(got, one) == []
            assert np.array_equal(got.hit_off, one.hit_off) and got.stats["n_bases"] == one.stats["n_bases"]
            ora = c.oracle.map_batch(buf[:int(offs[2000])], offs[:2001], os.cpu_count() or 8)
            sub = multi.map_batch(buf[:int(offs[2000])], offs[:2001])
            assert parity.compare_hits(sub, ora) == []
            empty = multi.map_batch(buf[:0], offs[:1])
            assert len(empty.hits) == 0
        finally:
            multi.close()
    finally:
        c.close()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least two GPUs")
def test_python_aligner_over_all_devices(gpu_lib):
    """mappy_rs.Aligner(devices=[...]) behind the reference's API: same mappings as the one-device aligner."""
    import mappy_rs
    from conftest import GOLDEN
    mmi = os.path.join(GOLDEN, "test.mmi")
    a1 = mappy_rs.Aligner(mmi)
    an = mappy_rs.Aligner(mmi, devices=list(range(_n_gpus())))
    try:
        seqs = [a1.seq(n) for n in a1.seq_names] * 50
        a1.enable_threading(2), an.enable_threading(2)
        r1 = sorted((d["i"], [str(m) for m in ms]) for ms, d in a1.map_batch([{"seq": s, "i": i} for i, s in enumerate(seqs)]))
        rn = sorted((d["i"], [str(m) for m in ms]) for ms, d in an.map_batch([{"seq": s, "i": i} for i, s in enumerate(seqs)]))
        assert r1 == rn and len(r1) == 200
    finally:
        a1.close(), an.close()
