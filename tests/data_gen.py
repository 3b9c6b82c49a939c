"""Synthetic reference / read generator.

Test and benchmark infrastructure: a small C generator (tests/csrc/simgen.c)
driven through ctypes.  Not part of the product path.
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "csrc", "libsimgen.so")


def build(force=False):
    src = os.path.join(_HERE, "csrc", "simgen.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _SO, src, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.sim_reference.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]
        _lib.sim_plant_families.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double]
        _lib.sim_plant_repeats.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_double]
        _lib.sim_reads.restype = ctypes.c_uint64
        _lib.sim_reads.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                   ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def make_reference(seed, contig_lens, n_repeats=0, rep_min=300, rep_max=6000, rep_div=0.05, n_families=0, fam_copies=0, fam_div=0.1):
    """Returns (ref uint8 array of ASCII bases, contig offsets uint64[n+1], names).  n_repeats: two-copy duplications;
    n_families x fam_copies: high-copy repeat families (consensus of rep_min..rep_max bases, fam_div divergence per copy)."""
    contig_lens = [int(x) for x in contig_lens]
    coff = np.zeros(len(contig_lens) + 1, dtype=np.uint64)
    coff[1:] = np.cumsum(contig_lens)
    total = int(coff[-1])
    ref = np.empty(total, dtype=np.uint8)
    lib().sim_reference(seed, total, ref.ctypes.data)
    if n_repeats:
        lib().sim_plant_repeats(seed + 1000, ref.ctypes.data, total, n_repeats, rep_min, rep_max, rep_div)
    if n_families and fam_copies:
        lib().sim_plant_families(seed + 2000, ref.ctypes.data, total, n_families, fam_copies, rep_min, rep_max, fam_div)
    names = ["chr%d" % (i + 1) for i in range(len(contig_lens))]
    return ref, coff, names


def make_reads(seed, ref, coff, n_reads, len_min=1000, len_max=10000, len_mean=0.0, len_sd=0.0,
               p_sub=0.03, p_ins=0.02, p_del=0.03):
    """Returns (buf uint8, offsets uint64[n+1], truth int64[n,4] = ctg,start,end,strand)."""
    cap = int(n_reads * (len_max * 1.25 + 64))
    buf = np.empty(cap, dtype=np.uint8)
    offs = np.empty(n_reads + 1, dtype=np.uint64)
    truth = np.empty((n_reads, 4), dtype=np.int64)
    n = lib().sim_reads(seed, ref.ctypes.data, len(coff) - 1, coff.ctypes.data, n_reads, len_min, len_max,
                        len_mean, len_sd, p_sub, p_ins, p_del, buf.ctypes.data, offs.ctypes.data, truth.ctypes.data)
    return buf[:n].copy(), offs, truth


def write_fasta(path, ref, coff, names, width=80):
    with open(path, "wb") as fh:
        for i, nm in enumerate(names):
            fh.write(b">" + nm.encode() + b"\n")
            s = ref[int(coff[i]):int(coff[i + 1])]
            for j in range(0, len(s), width):
                fh.write(s[j:j + width].tobytes() + b"\n")


def reads_as_list(buf, offs):
    b = buf.tobytes()
    return [b[int(offs[i]):int(offs[i + 1])].decode() for i in range(len(offs) - 1)]


# BASELINE.md section 3 configurations ------------------------------------------------
GRCH38_LENS = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
               138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
               83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415]


def config1_reference():
    return make_reference(1, [5_000_000])


def config1_reads(ref, coff, n_reads=200_000):
    return make_reads(2, ref, coff, n_reads, 1000, 10000, p_sub=0.03, p_ins=0.02, p_del=0.03)


def config2_contig_lens(total=3_100_000_000):
    s = float(sum(GRCH38_LENS))
    return [int(round(x / s * total)) for x in GRCH38_LENS]


def make_sv_reads(seed, ref, coff, n_reads, len_min=400, len_max=4000):
    """Reads that exercise multi-chain logic: plain reads, chimeras (two loci), reads with a 0.6-4 kb
    deletion or insertion relative to the reference (broken chains -> RMQ long-join re-chaining)."""
    rs = np.random.RandomState(seed)
    base_buf, base_offs, _ = make_reads(seed + 7, ref, coff, n_reads * 2, len_min, len_max)
    reads = reads_as_list(base_buf, base_offs)
    comp = str.maketrans("ACGT", "TGCA")
    out = []
    n_ctg = len(coff) - 1
    for i in range(n_reads):
        kind = rs.randint(6)
        c = rs.randint(n_ctg)
        c0, L = int(coff[c]), int(coff[c + 1] - coff[c])
        if kind == 0 or L < 14000:
            out.append(reads[2 * i])
        elif kind == 1:
            out.append(reads[2 * i] + reads[2 * i + 1])
        elif kind == 2:
            st, gap = rs.randint(0, L - 12000), rs.randint(600, 4000)
            s = ref[c0 + st:c0 + st + 2500].tobytes().decode() + ref[c0 + st + 2500 + gap:c0 + st + 5000 + gap].tobytes().decode()
            out.append(s.translate(comp)[::-1] if rs.randint(2) else s)
        elif kind == 3:
            st = rs.randint(0, L - 12000)
            ins = "".join(rs.choice(list("ACGT"), rs.randint(600, 3000)))
            out.append(ref[c0 + st:c0 + st + 2000].tobytes().decode() + ins + ref[c0 + st + 2000:c0 + st + 4500].tobytes().decode())
        elif kind == 4:  # inversion: the middle stretch is reverse-complemented (z-drop split + inversion alignment)
            st, il = rs.randint(0, L - 12000), rs.randint(200, 1500)
            mid = ref[c0 + st + 2000:c0 + st + 2000 + il].tobytes().decode().translate(comp)[::-1]
            s = ref[c0 + st:c0 + st + 2000].tobytes().decode() + mid + ref[c0 + st + 2000 + il:c0 + st + 4500 + il].tobytes().decode()
            out.append(s.translate(comp)[::-1] if rs.randint(2) else s)
        else:  # a stretch replaced by unrelated sequence of the same length (z-drop without a gap)
            st, il = rs.randint(0, L - 12000), rs.randint(300, 1500)
            mid = "".join(rs.choice(list("ACGT"), il))
            s = ref[c0 + st:c0 + st + 2000].tobytes().decode() + mid + ref[c0 + st + 2000 + il:c0 + st + 4500 + il].tobytes().decode()
            out.append(s.translate(comp)[::-1] if rs.randint(2) else s)
    bs = [s.encode() for s in out]
    offs = np.zeros(len(bs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(b) for b in bs])
    return np.frombuffer(b"".join(bs), dtype=np.uint8).copy(), offs


def config0_reads(contigs, n_reads=20000, seed=20261018, len_min=1000, len_max=10000, err=0.08, n_frac=0.0):
    """BASELINE.json configs[0]: the reference's 4 x 400 bp fixture cannot host 1-10 kb reads,
    so every read is random flank + a substring of a contig (200-400 bp, random contig and strand, `err` errors
    split 3:2:3 into substitutions / insertions / deletions) + random flank.  `contigs`: list of str."""
    rs = np.random.RandomState(seed % (2 ** 32))
    comp = str.maketrans("ACGT", "TGCA")
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    p_sub, p_ins, p_del = err * 3 / 8, err * 2 / 8, err * 3 / 8
    out = []
    for _ in range(n_reads):
        total = int(rs.randint(len_min, len_max + 1))
        ctg = contigs[rs.randint(len(contigs))]
        sl = int(rs.randint(200, min(400, len(ctg)) + 1))
        st = int(rs.randint(0, len(ctg) - sl + 1))
        core = ctg[st:st + sl]
        if rs.randint(2):
            core = core.translate(comp)[::-1]
        c = np.frombuffer(core.encode(), dtype=np.uint8)
        u = rs.random_sample(len(c))
        pieces = []
        for b, x in zip(c.tolist(), u.tolist()):
            if x < p_del:
                continue
            if x < p_del + p_sub:
                b = int(acgt[(b"ACGT".find(bytes([b])) + 1 + rs.randint(3)) % 4]) if bytes([b]) in b"ACGT" else b
            pieces.append(b)
            if x > 1.0 - p_ins:
                pieces.append(int(acgt[rs.randint(4)]))
        core_b = bytes(pieces)
        left = int(rs.randint(0, max(1, total - len(core_b))))
        right = max(0, total - len(core_b) - left)
        s = acgt[rs.randint(0, 4, left)].tobytes() + core_b + acgt[rs.randint(0, 4, right)].tobytes()
        if n_frac > 0:
            a = np.frombuffer(s, dtype=np.uint8).copy()
            a[rs.random_sample(len(a)) < n_frac] = ord("N")
            s = a.tobytes()
        out.append(s)
    offs = np.zeros(len(out) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(b) for b in out])
    return np.frombuffer(b"".join(out), dtype=np.uint8).copy(), offs


def prefixes(buf, offs, n=400):
    """BASELINE.json configs[3]: the first `n` bases of every read (readfish-style adaptive sampling)."""
    ln = np.minimum(np.diff(offs.astype(np.int64)), n)
    noffs = np.zeros(len(offs), dtype=np.uint64)
    noffs[1:] = np.cumsum(ln)
    idx = np.repeat(offs[:-1].astype(np.int64) - noffs[:-1].astype(np.int64), ln) + np.arange(int(noffs[-1]))
    return buf[idx].copy(), noffs


def sprinkle_n(buf, seed, frac=0.002):
    """A copy of the reads with a fraction of the bases replaced by N (ambiguity codes in cs / MD / sc_ambi paths)."""
    rs = np.random.RandomState(seed)
    a = buf.copy()
    a[rs.random_sample(len(a)) < frac] = ord("N")
    return a
