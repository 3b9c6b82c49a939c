#!/usr/bin/env python
"""bench.py -- reads/s of the minimap2 mapping path behind mappy-rs `map_batch` on B200.

Default workload = the configuration BASELINE.json's metric is quoted on (configs[2]):
3.1 Gb synthetic reference (24 contigs, seed 3; the index is built on the device),
250 000 simulated ONT reads of 1-10 kb with 8 % errors per GPU (2 M / 8; seed 4), preset
map-ont, **MM_F_CIGAR on** - the only mode mappy-rs can run (/root/reference/src/lib.rs:339
ORs the flag in unconditionally).  One "step" = one pass of the whole hot path (sketch ->
seed -> anchor filter/expand -> sort -> chain -> select -> ksw_extd2 extension / CIGAR ->
mapq) over the batch.  The same line carries a `mapping_only` object: the same reads through
the chain-level path without base alignment (BASELINE.json configs[1]'s mode), measured the
same way.  --workload config1 / prefix / hifi select configs[1] (5 Mb reference, 200 000
reads), configs[3] (400-base prefixes streamed in 20k batches, with batch latency) and
configs[4] (map-hifi); --mapping-only makes the chain-level mode the primary one.

  value     : reads/s with the reads already resident in HBM, timed by CUDA
              events on the library's stream around all kernels of a step.
  e2e       : reads/s through the C-ABI call a host makes (mmg_map_batch) with
              HOST (pinned) buffers: H2D + kernels + D2H inside the timed region.
  roofline  : dominant stage, algorithmic bytes (anchors counted AFTER the isolated-anchor
              filter, i.e. what the kernel processes) / measured event time vs the measured HBM copy
              peak (MEASURED_PEAKS.json).  The dominant kernels (ext_dp, chain_dp) are integer-issue
              bound: `int32_roofline` (24 int-ops per DP cell / predecessor over the INT32 issue
              peak this library measures on the device, mmg_debug_int32_peak) is the bound that
              explains them.
  cpu_baseline / --impl reference : the oracle (CPU restatement of minimap2
              2.26; the reference itself cannot be built here)
              on all host cores over a bounded sample of the same reads.
N > 1 (torchrun): index replicated per GPU (each rank builds it on its device), every rank maps its own reads
(weak scaling), no collective on the data path; timing = max over ranks.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "mappy-rs_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

def _kept(s):
    """anchors that reach the sort and the chaining DP: the isolated-anchor filter (seed.cu) removes n_dropped of the
    n_anchor anchors upstream would have sorted, before they are ever written"""
    return s["n_anchor"] - s.get("n_dropped", 0)


ALGO_BYTES = {  # BASELINE.md "Roofline accounting": algorithmic bytes per step from the batch counters
    "sketch": lambda s: s["n_bases"] + 16 * s["n_mz"],
    "seed": lambda s: 16 * s["n_mz"] + 16 * s["n_mz"],
    "expand": lambda s: 8 * s["n_hit"] + 16 * _kept(s),
    "sort": lambda s: 32 * _kept(s),
    "chain_dp": lambda s: 32 * _kept(s),
    "backtrack": lambda s: 32 * s["n_kept"],
    "rechain": lambda s: 32 * s["n_kept"],
    "regs": lambda s: 32 * s["n_kept"],
    "extend": lambda s: 2 * s["n_cell"],
}
INT_OPS = {  # integer operations of the two issue-bound stages
    "chain_dp": lambda s: 24 * s["n_iter"],
    "extend": lambda s: 24 * s["n_cell"],
}


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None, "reasons": reasons}


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


WORKLOADS = {
    # name: (description, reference builder, read simulator kwargs, preset, reads per GPU default)
    "config1": "BASELINE.json configs[1]: 5 Mb synthetic reference (seed 1), simulated 1-10 kb ONT reads (3% sub, 2% ins, 3% del; seed 2), map-ont",
    "human": "BASELINE.json configs[2]: 3.1 Gb synthetic reference (24 contigs, GRCh38-proportional, seed 3), simulated 1-10 kb ONT reads (3% sub, 2% ins, 3% del; seed 4), map-ont, index replicated per GPU, reads sharded",
    "human-repeats": "BASELINE.json configs[2] stress variant (SURVEY.md section 8d): the 3.1 Gb reference with 5 % planted repeat families (50 families x 1000 copies of 0.3-6 kb, 10 % divergence per copy), same reads, map-ont",
    "prefix": "BASELINE.json configs[3]: 400-base read prefixes streamed in batches of 20 000 (readfish-style), map-ont",
    "hifi": "BASELINE.json configs[4]: simulated HiFi reads N(15 kb, 2 kb) clipped to [10 kb, 25 kb], 0.5% errors (seed 5), map-hifi",
}


def workload(args, rank):
    """Returns (ref, coff, names, buf, offs, preset) for this rank's shard (weak scaling: args.reads per GPU)."""
    import data_gen
    big = args.workload in ("human", "human-repeats") or args.ref == "human"
    if big:
        fam = dict(n_families=50, fam_copies=int(1000 * args.ref_bases / 3.1e9) or 1, rep_min=300, rep_max=6000, fam_div=0.1) if args.workload == "human-repeats" else {}
        ref, coff, names = data_gen.make_reference(3, data_gen.config2_contig_lens(args.ref_bases), **fam)
    else:
        ref, coff, names = data_gen.config1_reference()
    seed = (4 if big else 2) + 1000 * rank
    preset = None
    if args.workload == "hifi":
        buf, offs, _ = data_gen.make_reads(5 + 1000 * rank, ref, coff, args.reads, 10000, 25000, len_mean=15000.0, len_sd=2000.0, p_sub=0.002, p_ins=0.0015, p_del=0.0015)
        preset = "map-hifi"
    else:
        buf, offs, _ = data_gen.make_reads(seed, ref, coff, args.reads, 1000, 10000, p_sub=0.03, p_ins=0.02, p_del=0.03)
        if args.workload == "prefix":   # the first 400 bases of every read
            ln = np.minimum(np.diff(offs.astype(np.int64)), 400)
            noffs = np.zeros(len(offs), dtype=np.uint64)
            noffs[1:] = np.cumsum(ln)
            idx = np.repeat(offs[:-1].astype(np.int64) - noffs[:-1].astype(np.int64), ln) + np.arange(int(noffs[-1]))
            buf, offs = buf[idx].copy(), noffs
    return ref, coff, names, buf, offs, preset


def contig_seqs(ref, coff, names):
    return [ref[int(coff[i]):int(coff[i + 1])].tobytes() for i in range(len(names))]


def make_oracle(args, ref, coff, names, preset, cigar):
    import mm2oracle as mo
    o = mo.Oracle(names=names, seqs=contig_seqs(ref, coff, names), preset=preset)
    o.set_opt("flag", 4 if cigar else 0)
    return o


def config_of(args):
    """Static description of the workload: identical in the GPU arm and in `--impl reference`."""
    return {"workload": WORKLOADS[args.workload], "mode": "mapping-only" if args.mapping_only else "CIGAR on (MM_F_CIGAR, mappy-rs' only mode)",
            "reads_per_gpu": args.reads, "ref": "3.1 Gb" if (args.workload in ("human", "human-repeats") or args.ref == "human") else "5 Mb",
            "parallelism": "reads sharded, index replicated", "l2": "inputs and per-chunk arenas are far larger than L2 (126 MB); no flush needed"}


def run_reference(args, rank, world):
    """CPU arm: the oracle (kind 'port': the reference is Rust over un-vendored minimap2 C and cannot be built here)
    with every host thread, its ksw_extd2 on the SSE4.1 intrinsics upstream uses, on a bounded sample per step."""
    if rank != 0:
        return
    cfg = config_of(args)
    n_full = args.reads
    args.reads = min(args.reads, args.cpu_sample)
    ref, coff, names, buf, offs, preset = workload(args, 0)
    n_sample = len(offs) - 1
    o = make_oracle(args, ref, coff, names, preset, not args.mapping_only)
    cores = os.cpu_count() or 1
    nw = min(2000, n_sample)
    for _ in range(args.warmup):
        o.map_batch(buf[:int(offs[nw])], offs[:nw + 1], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.map_batch(buf, offs, cores, cs=not args.mapping_only)   # map_batch workers always build cs (src/lib.rs:589)
    dt = (time.perf_counter() - t0) / args.steps
    v = n_sample / dt
    sample = "first %d of the %d reads per step (%.1f Mbases), all host threads" % (n_sample, n_full, int(offs[-1]) / 1e6)
    print(json.dumps({
        "impl": "reference", "metric": "reads_per_s", "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic", "mbases_per_s": int(offs[-1]) / dt / 1e6,
        "config": cfg,
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def load_traffic():
    """dram bytes per launch of each stage kernel from the committed `ncu --set full` capture (profiles/)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


def measure_mode(args, lib, idx, mopt_base, cigar, views, n_reads, n_bases, local_rank, world, barrier, maxrank, allranks):
    """Device-resident and end-to-end timing of one mode (CIGAR on / mapping-only) on this rank's reads."""
    from mappy_rs import _mmg
    mopt = _mmg.MapOpt.from_buffer_copy(mopt_base)
    mopt.flag = 4 if cigar else 0
    al = _mmg.DeviceAligner(lib, idx, mopt, device=local_rank)
    for key in ("dual_stream", "ramp_shift", "sort_small_max", "chunk_bases", "chunk_reads", "anchor_cap", "regs_cap", "cigar_cap", "jobs_cap", "keep_words"):
        if os.environ.get("MMG_" + key.upper()):
            al.set(key, int(os.environ["MMG_" + key.upper()]))
    # ---- device-resident timing -------------------------------------------------
    al.set("profile", 0)
    handles = [al.upload(v, o) for v, o in views]
    for _ in range(args.warmup):
        for b in handles:
            al.run(b)
    barrier()
    dev_ms, stage_ms, stage_ln, launches = 0.0, {}, {}, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for b in handles:
            al.run(b)
            dev_ms += al.last_run_ms()
            for k, (ms, ln) in al.stage_times().items():
                stage_ln[k] = stage_ln.get(k, 0) + ln
                launches += ln
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    # per-stage kernel times: one more pass with per-stage CUDA events; the timed steps run without them
    al.set("profile", 1)
    for b in handles:
        al.run(b)
        for k, (ms, ln) in al.stage_times().items():
            stage_ms[k] = stage_ms.get(k, 0.0) + ms
    prof_run_ms = al.last_run_ms() if len(handles) == 1 else None
    al.set("profile", 0)
    stats, res0 = {}, None
    for b, (v, o) in zip(handles, views):
        al.fetch(b)
        r = _mmg.Batch(lib, b, len(o) - 1)
        res0 = res0 or r
        for k, x in r.stats.items():
            stats[k] = stats.get(k, 0) + x
        al.free(b)
    per_rank_ms = allranks(dev_ms / args.steps)
    dev_ms = maxrank(dev_ms)
    # ---- end to end through the C ABI with host buffers ---------------------------
    for _ in range(max(min(args.warmup, 2), 1)):
        for v, o in views:
            al.map_batch(v, o)
    barrier()
    lat, d2h = [], 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        d2h = 0
        for v, o in views:
            t1 = time.perf_counter()
            r = al.map_batch(v, o, zero_copy=True)   # results are read in place (views of the library's pinned host memory)
            lat.append(time.perf_counter() - t1)
            d2h += r.hits.nbytes + r.cigar.nbytes + (len(o) - 1) * 4 + 80
            r.close()
    barrier()
    e2e_s = maxrank((time.perf_counter() - t0) / args.steps)
    int_peak = al.int32_peak(local_rank)
    al.close()
    ms_per_step = dev_ms / args.steps
    out = {"value": world * n_reads / (ms_per_step / 1e3), "ms_per_step": ms_per_step, "mbases_per_s": world * n_bases / (ms_per_step / 1e3) / 1e6,
           "e2e": {"value": world * n_reads / e2e_s, "unit": "reads/s", "h2d_bytes_per_step": n_bases + (n_reads + len(views)) * 8, "d2h_bytes_per_step": int(d2h),
                   "ms_per_step": e2e_s * 1e3, "mbases_per_s": world * n_bases / e2e_s / 1e6},
           "gpu_launches": int(launches), "stage_ms_per_step": dict(stage_ms), "counters": stats, "wall_ms_per_step": wall_ms / args.steps,
           "per_rank_ms_per_step": per_rank_ms,
           "host_gap_ms_per_step": (prof_run_ms - sum(stage_ms.values())) if prof_run_ms is not None else None,
           "host_gap_note": "profiled pass: CUDA-event time of the whole step minus the sum of its per-stage event times = time the stream idles between kernels (host round trips)"}
    stage_only = {k: v for k, v in stage_ms.items() if k in ALGO_BYTES and v > 0}
    top = max(stage_only, key=stage_only.get) if stage_only else "chain_dp"
    top_ms = stage_only.get(top, 0.0)
    peak, how = measured_peak()
    algo = ALGO_BYTES[top](stats)
    achieved = algo / (top_ms / 1e3) / 1e9 if top_ms > 0 else 0.0
    # one "launch" of a stage = its kernels over one chunk (what profiles/traffic.json sums, too); every chunk runs the sketch stage once
    n_launch = max(1, stage_ln.get("sketch", stage_ln.get(top, 1)) // args.steps)
    tr = load_traffic().get(("cigar:" if cigar else "") + top) or load_traffic().get(top)
    out["roofline"] = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                       "traffic": tr["dram_bytes_per_launch"] if tr else None, "peak_source": "of " + how, "ms_per_step": top_ms, "launches_per_step": n_launch,
                       "algorithmic_bytes_per_launch": algo / n_launch, "ms_per_launch": top_ms / n_launch, "traffic_source": (tr or {}).get("report"), "traffic_excludes": (tr or {}).get("kernels_without_dram_counters") or None,
                       "note": "achieved = algorithmic bytes of the stage per step (BASELINE.md; anchors after the isolated-anchor filter) / its summed event time. "
                               "This stage is integer-issue bound: see int32_roofline"}
    out["int32_roofline"] = {}
    for st, fn in INT_OPS.items():
        ms = stage_ms.get(st, 0.0)
        if ms > 0 and int_peak > 0:
            a = fn(stats) / (ms / 1e3) / 1e9
            out["int32_roofline"][st] = {"achieved": a, "peak": int_peak, "unit": "Gop/s", "frac": a / int_peak, "ms_per_step": ms,
                                         "gcups" if st == "extend" else "giter_per_s": (stats["n_cell"] if st == "extend" else stats["n_iter"]) / (ms / 1e3) / 1e9}
    out["int32_roofline"]["note"] = "24 int-ops per DP cell (extend) / per predecessor evaluation (chain_dp), BASELINE.md; peak = INT32 IADD3/LOP3 lane-op issue rate measured on this device by mmg_debug_int32_peak (self-measured: not in MEASURED_PEAKS.json)"
    return out, res0, lat


def run_group(args, lib, io, mopt, idx, ref, coff, names, preset, hbuf, offs, rank, local_rank, world, barrier, host_pg):
    """N > 1 only, after the per-rank (weak-scaling) legs: ONE process - rank 0 - drives all N GPUs through the
    product's multi-device aligner (mmg_aligner_create_multi): the index is built once and replicated over NVLink with
    peer copies, one common batch of N x reads_per_gpu reads (the shards the ranks mapped, concatenated) is sharded by
    bases inside mmg_map_batch and the results are gathered on the host in read order.  The other ranks release their
    GPUs and wait.  This is the reference's own shape (one aligner, N workers on a shared index, lib.rs:545-553)."""
    import torch
    import torch.distributed as dist
    from mappy_rs import _mmg
    import data_gen
    idx.close()
    torch.cuda.empty_cache()
    barrier()
    out = None
    if rank == 0:
        t0 = time.perf_counter()
        gidx = _mmg.Index.build(lib, io, names, contig_seqs(ref, coff, names), device=0)
        m2 = _mmg.MapOpt.from_buffer_copy(mopt)
        m2.flag = 0 if args.mapping_only else 4
        lib.check(lib.L.mmg_mapopt_update(ctypes.byref(m2), gidx.h))
        al = _mmg.DeviceAligner(lib, gidx, m2, devices=list(range(world)))
        setup_s = time.perf_counter() - t0
        bufs, lens = [hbuf.numpy()], [np.diff(offs.astype(np.int64))]
        big = args.workload in ("human", "human-repeats") or args.ref == "human"
        def reads_of(r):   # the reads rank r mapped (same seeds); the simulator is C code that releases the GIL
            if args.workload == "hifi":
                b, o, _ = data_gen.make_reads(5 + 1000 * r, ref, coff, args.reads, 10000, 25000, len_mean=15000.0, len_sd=2000.0, p_sub=0.002, p_ins=0.0015, p_del=0.0015)
            else:
                b, o, _ = data_gen.make_reads((4 if big else 2) + 1000 * r, ref, coff, args.reads, 1000, 10000, p_sub=0.03, p_ins=0.02, p_del=0.03)
                if args.workload == "prefix":
                    b, o = data_gen.prefixes(b, o, 400)
            return b, np.diff(o.astype(np.int64))
        import concurrent.futures
        with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(world - 1, 8))) as ex:
            for b, l in ex.map(reads_of, range(1, world)):
                bufs.append(b), lens.append(l)
        n_all = int(sum(len(x) for x in lens))
        goffs = np.zeros(n_all + 1, dtype=np.uint64)
        goffs[1:] = np.cumsum(np.concatenate(lens))
        gbuf = torch.empty(int(goffs[-1]), dtype=torch.uint8, pin_memory=True)
        pos = 0
        for b in bufs:
            gbuf.numpy()[pos:pos + len(b)] = b
            pos += len(b)
        gh = gbuf.numpy()
        steps = max(2, min(args.steps, 3))
        al.map_batch(gh, goffs, zero_copy=True).close()
        t0 = time.perf_counter()
        nh = 0
        for _ in range(steps):
            r = al.map_batch(gh, goffs, zero_copy=True)
            nh = len(r.hits)
            r.close()
        dt = (time.perf_counter() - t0) / steps
        out = {"value": n_all / dt, "unit": "reads/s", "ms_per_step": dt * 1e3, "n_gpus": world, "reads_per_step": n_all, "bases_per_step": int(goffs[-1]), "hits": nh, "steps": steps,
               "mbases_per_s": int(goffs[-1]) / dt / 1e6, "setup_s": setup_s,
               "what": "one process, one multi-device aligner (mmg_aligner_create_multi): index built once on GPU 0 and peer-copied to the others, ONE batch of all ranks' reads sharded by bases, results gathered on the host in read order; host buffers in, host results out (end to end)"}
        al.close()
        gidx.close()
    dist.barrier(group=host_pg)   # the waiting ranks block on the host (gloo): no NCCL kernel spins on the GPUs rank 0 is timing
    return out


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from mappy_rs import _mmg
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the mapping path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    host_pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_pg = dist.new_group(backend="gloo")   # host-side barrier for the one-process multi-device leg
    lib = _mmg.Lib()
    cfg = config_of(args)
    ref, coff, names, buf, offs, preset = workload(args, rank)
    n_reads, n_bases = len(offs) - 1, int(offs[-1])
    io, mopt = _mmg.IdxOpt(), _mmg.MapOpt()
    lib.check(lib.L.mmg_set_opt(None, ctypes.byref(io), ctypes.byref(mopt)))
    if preset:
        lib.check(lib.L.mmg_set_opt(preset.encode(), ctypes.byref(io), ctypes.byref(mopt)))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx = _mmg.Index.build(lib, io, names, contig_seqs(ref, coff, names), device=local_rank)
    index_build_s = time.perf_counter() - t0
    lib.check(lib.L.mmg_mapopt_update(ctypes.byref(mopt), idx.h))
    if os.environ.get("MMG_BENCH_PROFILER_RANGE"):
        # `ncu --profile-from-start off`: profile the mapping kernels only, not the one-off index build before them
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    # pinned host staging (torch is plumbing here: pinned memory + process group)
    hbuf = torch.empty(n_bases, dtype=torch.uint8, pin_memory=True)
    hbuf.numpy()[:] = buf
    hptr = hbuf.numpy()
    offs = np.ascontiguousarray(offs)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxrank(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allranks(x):
        if world == 1:
            return [x]
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    # batches of one step: the whole shard, or 20k-read batches in the streaming workload
    cuts = (list(range(0, n_reads, args.batch)) + [n_reads]) if args.workload == "prefix" else [0, n_reads]
    views = [(hptr[int(offs[a]):int(offs[b])], offs[a:b + 1] - offs[a]) for a, b in zip(cuts[:-1], cuts[1:])]
    sampler = ClockSampler(local_rank) if rank == 0 else None   # one NVML poller per job, not one per rank
    if sampler:
        sampler.start()
    primary_cigar = not args.mapping_only
    prim, res0, lat = measure_mode(args, lib, idx, mopt, primary_cigar, views, n_reads, n_bases, local_rank, world, barrier, maxrank, allranks)
    second = None
    if primary_cigar and not args.no_secondary:
        second, _, _ = measure_mode(args, lib, idx, mopt, False, views, n_reads, n_bases, local_rank, world, barrier, maxrank, allranks)
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    group = None
    if world > 1 and not args.no_group:
        try:
            group = run_group(args, lib, io, mopt, idx, ref, coff, names, preset, hbuf, offs, rank, local_rank, world, barrier, host_pg)
        except Exception as e:   # the per-rank lines above stay valid; the failure is reported, not hidden
            group = {"error": str(e)[:300]}
            try:
                dist.barrier(group=host_pg)
            except Exception:
                pass
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    out = {
        "metric": "reads_per_s", "value": prim["value"], "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": prim["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "mbases_per_s": prim["mbases_per_s"], "config": cfg,
        "setup": {"index_build_s": index_build_s, "index": "built on the device (index_dev.cu)", "reads_this_rank": n_reads, "bases_this_rank": n_bases, "batches_per_step": len(views)},
        "e2e": prim["e2e"], "gpu_launches": prim["gpu_launches"], "roofline": prim["roofline"], "int32_roofline": prim["int32_roofline"],
        "stage_ms_per_step": prim["stage_ms_per_step"],
        "stage_ms_note": "from one extra pass with per-stage CUDA events; the timed steps run without them",
        "counters": prim["counters"], "wall_ms_per_step": prim["wall_ms_per_step"], "per_rank_ms_per_step": prim["per_rank_ms_per_step"],
        "host_gap_ms_per_step": prim["host_gap_ms_per_step"], "host_gap_note": prim["host_gap_note"],
        "clocks": sampler.summary(),
    }
    if group is not None:
        out["group"] = group
    if second is not None:
        out["mapping_only"] = {k: second[k] for k in ("value", "ms_per_step", "mbases_per_s", "e2e", "gpu_launches", "roofline", "int32_roofline", "stage_ms_per_step",
                                                      "counters", "per_rank_ms_per_step", "host_gap_ms_per_step")}
        out["mapping_only"]["note"] = "same reads, same index, MM_F_CIGAR off (chain-level coordinates; BASELINE.json configs[1]'s mode, not reachable through mappy-rs' API)"
    if args.workload == "prefix":
        ls = sorted(lat)
        out["latency_ms"] = {"batch_reads": args.batch, "p50": 1e3 * ls[len(ls) // 2], "p99": 1e3 * ls[min(len(ls) - 1, int(len(ls) * 0.99))], "max": 1e3 * ls[-1], "n": len(ls)}
    if world == 1 and not args.no_cpu_baseline:
        import parity
        ns = min(n_reads, args.cpu_sample)
        t0 = time.perf_counter()
        o = make_oracle(args, ref, coff, names, preset, primary_cigar)
        t_idx = time.perf_counter() - t0
        cores = os.cpu_count() or 1
        t0 = time.perf_counter()
        ores = o.map_batch(buf[:int(offs[ns])], offs[:ns + 1], cores, cs=primary_cigar)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": ns / dt, "unit": "reads/s", "cores": cores, "kind": "port", "index_build_s": t_idx,
                               "sample": "first %d reads (%.1f Mbases) of the same workload, oracle (SSE4.1 ksw_extd2, cs strings built) on all host threads" % (ns, int(offs[ns]) / 1e6)}
        # the sample is also a parity check of the timed GPU results: every hit field and every CIGAR operation
        nh = int(ores.hit_off[-1])
        class _V:  # the GPU results restricted to the sampled reads (first batch)
            pass
        dv = _V()
        m = min(ns, len(res0.hit_off) - 1)
        dv.hit_off, dv.hits, dv.cigar, dv.hit_cigar = res0.hit_off[:m + 1], res0.hits[:int(res0.hit_off[m])], res0.cigar, res0.hit_cigar
        if m < ns:
            ov = _V()
            ov.hit_off, ov.hits, ov.cigar, ov.hit_cigar = ores.hit_off[:m + 1], ores.hits[:int(ores.hit_off[m])], ores.cigar, ores.hit_cigar
        else:
            ov = ores
        diffs = parity.compare_hits(dv, ov)
        out["cpu_baseline"]["sample_matches_gpu"] = not diffs
        out["cpu_baseline"]["sample_check"] = "all %d hit fields + every CIGAR op of %d hits (tests/parity.compare_hits)" % (len(parity.MAPQ_FIELDS) - 1, nh)
        if diffs:
            out["cpu_baseline"]["sample_diffs"] = diffs[:3]
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (default 250000; prefix workload 400000 = 20 batches of 20000)")
    ap.add_argument("--workload", default="human", choices=sorted(WORKLOADS),
                    help="default: the configuration BASELINE.json's metric is quoted on (configs[2], 3.1 Gb reference; it fits one GPU)")
    ap.add_argument("--ref", default="", choices=["", "human"], help="prefix/hifi workloads: use the 3.1 Gb reference instead of the 5 Mb one")
    ap.add_argument("--ref-bases", type=int, default=3_100_000_000)
    ap.add_argument("--batch", type=int, default=20000, help="prefix workload: reads per streamed batch")
    ap.add_argument("--cpu-sample", type=int, default=20000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cigar", action="store_true", help="(default) MM_F_CIGAR on: what mappy-rs itself always runs")
    ap.add_argument("--mapping-only", action="store_true", help="primary mode = chain-level mapping without base alignment (configs[1]'s mode)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the mapping_only leg of the default line")
    ap.add_argument("--no-group", action="store_true", help="N > 1: skip the one-process multi-device leg (one common batch sharded over all GPUs, gathered on rank 0)")
    args = ap.parse_args()
    if args.reads <= 0:
        args.reads = {"config1": 200000, "human": 250000, "human-repeats": 250000, "prefix": 400000, "hifi": 40000}[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
