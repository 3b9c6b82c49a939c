"""Test helper: how a launcher with one process per GPU cuts a batch: contiguous shards of (nearly)
equal BASES, index replicated, results back in read order, no collective on the data path.  The product does this
split itself inside `mmg_map_batch` of a multi-device aligner (csrc/pipeline.cu map_batch_group); this copy lets the
world_size-2 gloo test drive two single-device aligners the way torchrun ranks would."""
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
