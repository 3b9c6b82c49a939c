#!/usr/bin/env python
"""Writes the differential corpus that tools/pin_with_mappy.py maps with a REAL minimap2 (tests/golden/corpus/):
a 450 kb two-contig reference with planted repeats, and reads that exercise every branch the oracle restates
(plain noisy reads, chimeras, long deletions / insertions -> re-chaining, inversions -> z-drop splits and
mm_align1_inv, unrelated stretches -> z-drop without a gap, reads with N bases, reads on the reference's own
test.fa contigs with random flanks).  Deterministic (fixed seeds); the files are committed so that the machine that
has minimap2 needs neither gcc nor this repo's generator."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import data_gen  # noqa: E402


def main():
    out = os.path.join(ROOT, "tests", "golden", "corpus")
    os.makedirs(out, exist_ok=True)
    ref, coff, names = data_gen.make_reference(41, [300000, 150000], n_repeats=60, rep_min=300, rep_max=4000, rep_div=0.03)
    data_gen.write_fasta(os.path.join(out, "ref.fa"), ref, coff, names, width=100)
    reads = []
    b, o = data_gen.make_sv_reads(57, ref, coff, 150)
    reads += data_gen.reads_as_list(b, o)
    b, o, _ = data_gen.make_reads(58, ref, coff, 80, 500, 6000)
    reads += data_gen.reads_as_list(data_gen.sprinkle_n(b, 59, 0.003), o)
    b, o, _ = data_gen.make_reads(60, ref, coff, 30, 8000, 15000, p_sub=0.002, p_ins=0.0015, p_del=0.0015)
    reads += data_gen.reads_as_list(b, o)
    with open(os.path.join(out, "reads.fa"), "w") as fh:
        for i, s in enumerate(reads):
            fh.write(">r%d\n%s\n" % (i, s))
    # reads for the reference's own fixture index (tests/golden/test.mmi): flank + noisy contig slice + flank
    contigs = []
    name = None
    for line in open(os.path.join(ROOT, "tests", "golden", "test.fa")):
        if line.startswith(">"):
            contigs.append("")
        else:
            contigs[-1] += line.strip()
    b, o = data_gen.config0_reads(contigs, 120, len_max=3000)
    with open(os.path.join(out, "reads_fixture.fa"), "w") as fh:
        for i, s in enumerate(data_gen.reads_as_list(b, o)):
            fh.write(">f%d\n%s\n" % (i, s))
    print("wrote", out, len(reads), "reads + 120 fixture reads")


if __name__ == "__main__":
    main()
