"""N > 1 host logic on CPU: two `gloo` ranks each map their shard of a batch (index replicated, reads split by bases,
no data-path collective); the gathered result must equal the unsharded oracle.  Uses the SIMT-emulated test build."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from conftest import EMU_LIB, ROOT


def test_split_by_bases_balances_and_covers():
    from shard_util import split_by_bases
    rs = np.random.RandomState(3)
    lens = rs.randint(100, 10000, 1000)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    for n in (1, 2, 3, 8):
        b = split_by_bases(offs, n)
        assert b[0] == 0 and b[-1] == 1000 and all(x <= y for x, y in zip(b, b[1:]))
        per = [int(offs[b[i + 1]] - offs[b[i]]) for i in range(n)]
        assert max(per) - min(per) <= 2 * lens.max()
    assert split_by_bases(np.array([0], dtype=np.uint64), 4) == [0, 0, 0, 0, 0]


WORKER = textwrap.dedent('''
    import os, sys, json
    sys.path[:0] = [r"{root}/tests", r"{root}/mappy-rs_b200", r"{root}/oracle"]
    import numpy as np, torch.distributed as dist
    import data_gen, parity
    from mappy_rs import _mmg
    from shard_util import shard
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = _mmg.Lib(r"{emu}")
    ref, coff, names, seqs = parity.random_reference(5, [120000])
    buf, offs, _ = data_gen.make_reads(6, ref, coff, 90, 300, 3000)
    c = parity.Case(lib, names, seqs)                      # index replicated on every rank
    sbuf, soffs, first = shard(buf, offs, rank, world)
    dev = c.aligner.map_batch(sbuf, soffs)
    mine = [(first + i, [tuple(int(h[f]) for f in ("rid", "rs", "re", "qs", "qe", "rev", "mapq")) for h in dev.read_hits(i)])
            for i in range(len(soffs) - 1)]
    out = [None] * world
    dist.all_gather_object(out, mine)                      # results to the host side only; no data-path collective
    if rank == 0:
        got = dict(x for part in out for x in part)
        ora = c.oracle.map_batch(buf, offs, 2)
        want = {{i: [tuple(int(h[f]) for f in ("rid", "rs", "re", "qs", "qe", "rev", "mapq")) for h in ora.read_hits(i)] for i in range(90)}}
        assert got == want, "sharded result differs from the oracle"
        print("SHARD_OK", len(got))
    dist.barrier()
    dist.destroy_process_group()
''')


def test_two_rank_gloo_sharded_mapping(emu_lib, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, emu=EMU_LIB))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, MMG_EMU_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARD_OK 90" in r.stdout
