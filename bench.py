#!/usr/bin/env python
"""bench.py -- reads/s of the minimap2 mapping path behind mappy-rs `map_batch` on B200.

Workload (BASELINE.json configs[1]): 5 Mb synthetic reference (seed 1), 200 000
simulated ONT reads of 1-10 kb with 8 % errors (seed 2), preset map-ont,
mapping-only, one GPU.  One "step" = one pass of the whole hot path (sketch ->
seed -> sort -> chain -> select/mapq) over the batch.

  value     : reads/s with the reads already resident in HBM, timed by CUDA
              events on the library's stream around all kernels of a step.
  e2e       : reads/s through the C-ABI call a host makes (mmg_map_batch) with
              HOST (pinned) buffers: H2D + kernels + D2H inside the timed region.
  roofline  : dominant kernel, algorithmic bytes / measured launch time vs the
              measured HBM copy peak (MEASURED_PEAKS.json).
  cpu_baseline / --impl reference : the oracle (CPU restatement of minimap2
              2.26; the reference itself cannot be built here)
              on all host cores over a bounded sample of the same reads.
N > 1 (torchrun): index replicated per GPU, every rank maps its own 200k reads
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

ALGO_BYTES = {  # BASELINE.md "Roofline accounting": algorithmic bytes per launch from the batch counters
    "sketch": lambda s: s["n_bases"] + 16 * s["n_mz"],
    "seed": lambda s: 16 * s["n_mz"] + 16 * s["n_mz"],
    "expand": lambda s: 8 * s["n_hit"] + 16 * s["n_anchor"],
    "sort": lambda s: 32 * s["n_anchor"],
    "chain_dp": lambda s: 32 * s["n_anchor"],
    "backtrack": lambda s: 32 * s["n_kept"],
    "rechain": lambda s: 32 * s["n_kept"],
    "regs": lambda s: 32 * s["n_kept"],
    "extend": lambda s: 2 * s["n_cell"],
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


def workload(n_reads, rank):
    import data_gen
    ref, coff, names = data_gen.config1_reference()
    buf, offs, _ = data_gen.make_reads(2 + 1000 * rank, ref, coff, n_reads, 1000, 10000, p_sub=0.03, p_ins=0.02, p_del=0.03)
    return ref, coff, names, buf, offs


def run_reference(args, rank, world):
    """CPU arm: the oracle (kind 'port') with every host thread, on a bounded sample per step."""
    if rank != 0:
        return
    import mm2oracle as mo
    n_sample = min(args.reads, args.cpu_sample)
    ref, coff, names, buf, offs = workload(n_sample, 0)
    o = mo.Oracle(names=names, seqs=[ref.tobytes()])
    o.set_opt("flag", 4 if args.cigar else 0)
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        o.map_batch(buf[:int(offs[2000])], offs[:2001], cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.map_batch(buf, offs, cores)
    dt = (time.perf_counter() - t0) / args.steps
    v = n_sample / dt
    sample = "%d reads (%.1f Mbases) of the same workload per step" % (n_sample, int(offs[-1]) / 1e6)
    print(json.dumps({
        "impl": "reference", "metric": "reads_per_s", "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic", "mbases_per_s": int(offs[-1]) / dt / 1e6,
        "config": {"workload": "BASELINE.json configs[1]: 5 Mb synthetic reference, simulated 1-10 kb ONT reads (8% error), map-ont, mapping-only", "reads_per_step": n_sample},
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from mappy_rs import _mmg
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the mapping path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _mmg.Lib()
    ref, coff, names, buf, offs = workload(args.reads, rank)
    n_reads, n_bases = len(offs) - 1, int(offs[-1])
    io, mopt = _mmg.IdxOpt(), _mmg.MapOpt()
    lib.check(lib.L.mmg_set_opt(None, ctypes.byref(io), ctypes.byref(mopt)))
    mopt.flag = 4 if args.cigar else 0  # configs[1] is mapping-only
    idx = _mmg.Index.build(lib, io, names, [ref.tobytes()])
    lib.check(lib.L.mmg_mapopt_update(ctypes.byref(mopt), idx.h))
    al = _mmg.DeviceAligner(lib, idx, mopt, device=local_rank)
    al.set("profile", 1)
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

    # ---- device-resident timing -------------------------------------------------
    b = al.upload(hptr, offs)
    for _ in range(args.warmup):
        al.run(b)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    dev_ms, stage_ms, launches = 0.0, {}, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        al.run(b)
        dev_ms += al.last_run_ms()
        for k, (ms, ln) in al.stage_times().items():
            stage_ms[k] = stage_ms.get(k, 0.0) + ms
            launches += ln
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    al.fetch(b)
    res = _mmg.Batch(lib, b, n_reads)
    al.free(b)
    dev_ms = maxrank(dev_ms)
    # ---- end to end through the C ABI with host buffers ---------------------------
    al.set("profile", 0)
    for _ in range(max(min(args.warmup, 2), 1)):
        al.map_batch(hptr, offs)
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        r = al.map_batch(hptr, offs)
        d2h = r.hits.nbytes + r.cigar.nbytes + n_reads * 4 + 80
    barrier()
    e2e_s = maxrank((time.perf_counter() - t0) / args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if rank != 0:
        return
    stats = res.stats
    ms_per_step = dev_ms / args.steps
    value = world * n_reads / (ms_per_step / 1e3)
    stage_only = {k: v for k, v in stage_ms.items() if k in ALGO_BYTES and v > 0}
    top = max(stage_only, key=stage_only.get) if stage_only else "chain_dp"
    top_ms = stage_only.get(top, 0.0) / args.steps
    peak, how = measured_peak()
    achieved = ALGO_BYTES[top](stats) / (top_ms / 1e3) / 1e9 if top_ms > 0 else 0.0
    out = {
        "metric": "reads_per_s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "mbases_per_s": world * n_bases / (ms_per_step / 1e3) / 1e6,
        "config": {"workload": "BASELINE.json configs[1]: 5 Mb synthetic reference (seed 1), %d simulated 1-10 kb ONT reads (3%% sub, 2%% ins, 3%% del; seed 2), map-ont, %s" % (n_reads, "CIGAR on" if args.cigar else "mapping-only"),
                   "reads_per_gpu": n_reads, "bases_per_gpu": n_bases, "l2": "inputs (%.0f MB) larger than L2" % (n_bases / 1e6), "parallelism": "reads sharded, index replicated"},
        "e2e": {"value": world * n_reads / e2e_s, "unit": "reads/s", "h2d_bytes_per_step": n_bases + (n_reads + 1) * 8, "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": "of " + how, "ms_per_launch": top_ms, "note": "integer/latency-bound stage; algorithmic bytes per BASELINE.md"},
        "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
        "counters": stats, "wall_ms_per_step": wall_ms / args.steps,
        "clocks": sampler.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        import mm2oracle as mo
        ns = min(n_reads, args.cpu_sample)
        o = mo.Oracle(names=names, seqs=[ref.tobytes()])
        o.set_opt("flag", 4 if args.cigar else 0)
        cores = os.cpu_count() or 1
        t0 = time.perf_counter()
        ores = o.map_batch(buf[:int(offs[ns])], offs[:ns + 1], cores)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": ns / dt, "unit": "reads/s", "cores": cores, "kind": "port",
                               "sample": "first %d reads (%.1f Mbases) of the same workload, oracle on all host threads" % (ns, int(offs[ns]) / 1e6)}
        same = bool(np.array_equal(ores.hits["rs"], res.hits["rs"][:len(ores.hits)]) and np.array_equal(ores.hits["mapq"], res.hits["mapq"][:len(ores.hits)]))
        out["cpu_baseline"]["sample_matches_gpu"] = same
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=200000)
    ap.add_argument("--cpu-sample", type=int, default=20000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cigar", action="store_true", help="MM_F_CIGAR on (what mappy-rs itself always runs); default is configs[1] mapping-only")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
