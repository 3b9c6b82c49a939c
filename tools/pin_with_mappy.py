#!/usr/bin/env python
"""Pins the oracle against a REAL minimap2.

Neither minimap2 nor mappy / mappy_rs exists in the build container or on the GPU pool, so the
oracle (oracle/*.cpp) is a restatement that only the reference's few fixtures pin.  Run this script on ANY machine
where `import mappy` (minimap2's own Python binding, the module mappy-rs mimics) or `import mappy_rs` (the reference)
works - no GPU, no gcc needed:

    python tools/pin_with_mappy.py            # writes tests/golden/minimap2/*.json
    python -m pytest tests/test_pinned_by_minimap2.py -q

It maps the committed differential corpus (tests/golden/corpus/, written by tools/make_corpus.py) with the options
mappy-rs uses (MM_F_CIGAR on, cs requested: /root/reference/src/lib.rs:339, 589) and stores, per read, every field of
every hit.  tests/test_pinned_by_minimap2.py then requires the oracle to reproduce those files bit for bit; until
they exist that test is skipped and parity stays "unpinned".  Version 2.26 is what the reference pins
(Cargo.toml:24,29); other versions are recorded but flagged.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CORPUS = os.path.join(ROOT, "tests", "golden", "corpus")
OUT = os.path.join(ROOT, "tests", "golden", "minimap2")
CASES = [  # (name, index source, preset, reads)
    ("map-ont", "ref.fa", None, "reads.fa"),
    ("map-hifi", "ref.fa", "map-hifi", "reads.fa"),
    ("fixture-mmi", os.path.join("..", "test.mmi"), None, "reads_fixture.fa"),
]


def read_fasta(path):
    name, out = None, []
    for line in open(path):
        if line.startswith(">"):
            name = line[1:].strip()
        else:
            out.append((name, line.strip()))
    return out


def load_backend():
    try:
        import mappy
        return "mappy", getattr(mappy, "__version__", "?"), lambda idx, preset: mappy.Aligner(idx, preset=preset) if preset else mappy.Aligner(idx)
    except ImportError:
        pass
    import mappy_rs  # the reference itself
    if "mappy-rs_b200" in (getattr(mappy_rs, "__file__", "") or ""):
        raise ImportError("the importable mappy_rs is this repository's drop-in, not the reference")
    return "mappy_rs", getattr(mappy_rs, "__version__", "?"), lambda idx, preset: mappy_rs.Aligner(idx, preset=preset) if preset else mappy_rs.Aligner(idx)


def main():
    try:
        backend, version, make = load_backend()
    except ImportError as e:
        print("no real minimap2 binding here (%s): nothing written; the oracle stays pinned by the reference's fixtures only" % e)
        return 1
    os.makedirs(OUT, exist_ok=True)
    for name, idx, preset, reads in CASES:
        al = make(os.path.join(CORPUS, idx), preset)
        if not al:
            raise SystemExit("cannot open index " + idx)
        rec = []
        for rname, seq in read_fasta(os.path.join(CORPUS, reads)):
            hits = []
            for h in al.map(seq, cs=True):
                g = lambda *names: next(getattr(h, n) for n in names if hasattr(h, n))
                strand = g("strand")
                hits.append({"ctg": g("ctg", "target_name"), "ctg_len": g("ctg_len", "target_len"), "r_st": g("r_st", "target_start"), "r_en": g("r_en", "target_end"),
                             "q_st": g("q_st", "query_start"), "q_en": g("q_en", "query_end"), "strand": int(strand) if not isinstance(strand, str) else (1 if strand == "+" else -1),
                             "mapq": g("mapq"), "mlen": g("mlen", "match_len"), "blen": g("blen", "block_len"), "NM": g("NM"),
                             "is_primary": bool(g("is_primary")), "cigar": g("cigar_str"), "cs": g("cs")})
            rec.append({"read": rname, "hits": hits})
        with open(os.path.join(OUT, name + ".json"), "w") as fh:
            json.dump({"backend": backend, "version": version, "pinned_version": "2.26", "preset": preset, "index": idx, "reads": reads, "records": rec}, fh)
        print("%s: %d reads, %d hits (%s %s)" % (name, len(rec), sum(len(r["hits"]) for r in rec), backend, version))
    return 0


if __name__ == "__main__":
    sys.exit(main())
