"""Throughput of the drop-in Python API (mappy_rs.Aligner.map_batch: CIGAR + cs per hit, one dict per read) on the GPU box."""
import sys, os, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "mappy-rs_b200", "oracle"): sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, data_gen
import mappy_rs
ref, coff, names = data_gen.config1_reference()
fa = os.path.join(tempfile.mkdtemp(), "ref.fa")
data_gen.write_fasta(fa, ref, coff, names)
t0 = time.perf_counter()
al = mappy_rs.Aligner(fa)
print("Aligner(fasta) s", round(time.perf_counter() - t0, 2))
buf, offs, _ = data_gen.config1_reads(ref, coff, 20000)
reads = data_gen.reads_as_list(buf, offs)
al.enable_threading(4)
for rep in range(2):
    t0 = time.perf_counter()
    n = nh = 0
    for maps, meta in al.map_batch({"seq": s, "i": i} for i, s in enumerate(reads)):
        n += 1; nh += len(maps)
    dt = time.perf_counter() - t0
    print("map_batch: %d reads, %d hits in %.2f s = %.0f reads/s (%.1f Mbases/s)" % (n, nh, dt, n / dt, int(offs[-1]) / dt / 1e6))
m = al.map(reads[0])
print(m[0])
