"""Read sharding for multi-GPU runs: the index is replicated on every GPU, a batch is cut into
contiguous shards of (nearly) equal BASES (not read counts), every rank maps its shard, results return to the host in
read order.  There is no collective on the data path."""
import numpy as np


def split_by_bases(offs, n_parts):
    """offs: uint64[n+1] read offsets. Returns n_parts+1 read indices: shard p = reads [b[p], b[p+1])."""
    offs = np.asarray(offs, dtype=np.uint64)
    n = len(offs) - 1
    total = int(offs[-1] - offs[0])
    bounds = [0]
    for p in range(1, n_parts):
        target = int(offs[0]) + total * p // n_parts
        i = int(np.searchsorted(offs, target, side="left"))
        bounds.append(min(max(i, bounds[-1]), n))
    bounds.append(n)
    return bounds


def shard(buf, offs, rank, world):
    """(buffer view, rebased offsets, first read index) of this rank's shard."""
    b = split_by_bases(offs, world)
    lo, hi = b[rank], b[rank + 1]
    o = np.asarray(offs, dtype=np.uint64)
    return buf[int(o[lo]):int(o[hi])], (o[lo:hi + 1] - o[lo]).astype(np.uint64), lo
