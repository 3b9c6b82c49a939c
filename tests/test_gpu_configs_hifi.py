"""BASELINE.json configs[4] at its real reference size on the B200: map-hifi (k = 19, w = 19) with CIGAR on the 3.1 Gb
reference, against the oracle.  (Its own module so that the map-ont index of test_gpu_configs.py is released first.)"""
import os

import pytest

import data_gen
import parity
from test_gpu_configs import _Human

pytestmark = pytest.mark.gpu
NT = os.cpu_count() or 8


def test_config4_hifi_cigar_human_reference(gpu_lib, oracle_mod):
    h = _Human(gpu_lib, oracle_mod, preset="map-hifi")
    try:
        c = h.case(cigar=True)
        try:
            buf, offs, _ = data_gen.make_reads(5, h.ref, h.coff, 600, 10000, 25000, len_mean=15000.0, len_sd=2000.0, p_sub=0.002, p_ins=0.0015, p_del=0.0015)
            dev = c.aligner.map_batch(buf, offs)
            ora = c.oracle.map_batch(buf, offs, NT)
            assert parity.compare_hits(dev, ora) == []
            assert len(dev.hits) >= 598 and parity.compare_tags(c, buf, offs, dev, 0) == []
        finally:
            c.aligner.close()
    finally:
        h.close()
