#!/usr/bin/env python3
"""
bench.py -- long-read Gbp/s sketched+mapped on B200 (BASELINE.json metric), one JSON line on stdout.

Workloads (config.workload; BASELINE.json `configs`):
  --gpus 1/2/4  configs[2]: synthetic 100 Mbp assembly (1-200 kbp contigs), 40x ONT-like reads (4 Gbp) PER GPU,
                k=24 w=250 --sensitive (weak scaling: every rank maps its own 40x read set against the same target)
  --gpus 8      configs[3]: synthetic 3.1 Gbp human-size assembly, 30x ONT-like reads = 93 Gbp sharded over the 8 GPUs
                (11.6 Gbp each), k=32 w=250; the target is sketched in contig shards, the minimizer triples are
                all-gathered over NCCL and every GPU builds the full (DRAM-resident) replicated index
  --config c1   configs[1] (5 Mbp / 30x / k32 w100), the round-1 workload; --scale shrinks any of them for smoke runs
Inputs are generated on the device by the library's counter-based simulator (csrc/synth_logic.cuh); the same bytes can
be generated on the host (tests/test_gpu_synth.py), which is how the CPU arm and the parity check get their inputs.

One "step" = the whole job: target sketch + index build + read sketch + lookup + chaining + pair events (+ NCCL event
gather at N > 1) + pair tally.
  value     inputs resident in HBM; device time (CUDA events on the library's stream) or host wall time of the
            bracketed region, whichever is larger; max over ranks
  e2e       the same job through the C ABI with pinned HOST buffers: H2D of target + reads and D2H of all mapping
            results inside the timed region
  roofline  the dominant kernel: algorithmic bytes of the sketch (1.0 B/base + 13 B/minimizer) / its mean
            launch time, CUDA events around every launch inside the timed region, vs the measured HBM peak
  parity    after the timed region, on rank 0: the first reads of the run are mapped again (a) through the timed entry
            points on a second context and (b) through ntl_map_reads, and verbose_mapping / PAF / pairs.tsv /
            scaffold.dot bytes are compared with the CPU pipeline (the unmodified reference mapper when oracle/_ref is
            staged); the pair table of the resident arm must also equal the one of the e2e arm
  cpu_baseline  (N = 1) the CPU pipeline on a scale model of the workload: the contigs and reads of a genome window
            (see cpu_window_run), `indexlr -t <all cores> | ntlink_pair.py` exactly as the make recipe

`--impl reference` times that CPU pipeline instead (rank 0 only).
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

SEED = 20240502
Z = 1000
CONFIGS = {
    "c1": dict(label="configs[1]", genome=5_000_000, coverage=30, reads_per_gpu=150_000_000, k=32, w=100, sensitive=False,
               what="synthetic 5 Mbp genome, 30x ONT-like reads per GPU"),
    "c2": dict(label="configs[2]", genome=100_000_000, coverage=40, reads_per_gpu=4_000_000_000, k=24, w=250, sensitive=True,
               what="synthetic 100 Mbp assembly, 40x ONT-like reads (4 Gbp) per GPU"),
    "c3": dict(label="configs[3]", genome=3_100_000_000, coverage=30, reads_per_gpu=93_000_000_000 // 8, k=32, w=250, sensitive=False,
               what="synthetic 3.1 Gbp human-size assembly, 30x ONT-like reads (93 Gbp) in shards of 11.6 Gbp per GPU"),
}


def pick_config(args, world):
    name = args.config or ("c3" if world >= 8 else "c2")
    cfg = dict(CONFIGS[name], name=name)
    if args.scale != 1.0:
        cfg["genome"] = max(200_000, int(cfg["genome"] * args.scale))
        cfg["reads_per_gpu"] = max(2_000_000, int(cfg["reads_per_gpu"] * args.scale))
    if getattr(args, "reads_per_gpu", None):
        cfg["reads_per_gpu"] = int(args.reads_per_gpu)
        cfg["what"] += f" [reads per GPU overridden: {int(args.reads_per_gpu)} bp]"
    cfg["workload"] = (f"{cfg['label']}: {cfg['what']}, 1-200 kbp contigs, reads ~12 kbp lognormal with 4% sub / 3% del / 3% ins, "
                       f"k={cfg['k']} w={cfg['w']} z={Z}{' --sensitive' if cfg['sensitive'] else ''}"
                       + (f" [scaled x{args.scale}]" if args.scale != 1.0 else ""))
    return cfg


def plans(cfg, rank):
    from ntlink_b200 import synth
    cplan, names = synth.plan_assembly(cfg["genome"], SEED + 7)
    rplan = synth.plan_reads(cfg["genome"], cfg["reads_per_gpu"], SEED + 1 + 1000 * rank, first_id=rank * 100_000_000)
    return cplan, names, rplan


class ClockSampler(threading.Thread):
    "samples nvidia-smi clocks / throttle reasons while the timed region runs"

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.stop_flag, self.samples = gpu, threading.Event(), []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.check_output(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}",
                                               "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fin:
            return float(json.load(fin)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def scratch_dir():
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    return tempfile.mkdtemp(prefix="ntl_bench_", dir=base)


# ------------------------------------------------------------------------------------------- CPU pipeline
def cpu_pipeline():
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import cpu_pipeline as cp
    return cp


def cpu_window_inputs(cfg, window_bp, tmp):
    """Scale model of the workload for the CPU arm: the contigs of the assembly that lie inside the first `window_bp` bases
    of the genome and reads drawn from that window with the workload's own coverage and distributions. Returns
    (target.fa, reads.fa, read bases, number of reads, number of contigs)."""
    from ntlink_b200 import synth
    cp = cpu_pipeline()
    cplan, names = synth.plan_assembly(cfg["genome"], SEED + 7)
    window_bp = min(window_bp, cfg["genome"])
    inside = (cplan["start"] + cplan["len"]) <= window_bp
    cplan, names = cplan[inside], [n for n, keep in zip(names, inside) if keep]
    rplan = synth.plan_reads(window_bp, int(cfg["coverage"] * window_bp), SEED + 99)
    threads = os.cpu_count() or 1
    contigs = synth.host_contigs(SEED, cplan, names, threads)
    reads = synth.host_reads(SEED, rplan, threads=threads)
    tf, rf = os.path.join(tmp, "target.fa"), os.path.join(tmp, "reads.fa")
    cp.write_fasta(tf, contigs)
    cp.write_fasta(rf, reads)
    return tf, rf, int(reads.offsets[-1]), len(reads), len(contigs)


def cpu_window_run(cfg, tf, rf, tmp, threads):
    "one step of the CPU pipeline on the window: target sketch + (read sketch | mapper) as the make recipe; seconds"
    cp = cpu_pipeline()
    tsv, t_target = cp.sketch_target(tf, cfg["k"], cfg["w"], threads)
    t_map = cp.map_reads(tf, tsv, rf, os.path.join(tmp, "cpu"), cfg["k"], cfg["w"], Z, threads, sensitive=cfg["sensitive"], verbose=True)
    return t_target + t_map


def cpu_sample_text(cfg, window_bp, nreads, nbases, ncontig, kind):
    return (f"scale model of the workload: the {ncontig} contigs inside the first {window_bp / 1e6:.0f} Mbp of the genome + {nreads} reads "
            f"({nbases} bp, the workload's coverage and error model) drawn from that window; per step: indexlr (C restatement of btllib's, all "
            f"host threads) on the target, then indexlr --len on the reads piped into "
            + ("the UNMODIFIED bin/ntlink_pair.py of the reference (oracle/_ref, igraph stand-in)" if kind == "reference"
               else "oracle/pair_oracle.py (port of bin/ntlink_pair.py)")
            + " --verbose, single-threaded like the reference (no -t); FASTA files on tmpfs")


def run_reference(args, rank, world):
    if rank != 0:
        return
    cfg = pick_config(args, world)
    cp = cpu_pipeline()
    threads = os.cpu_count() or 1
    tmp = scratch_dir()
    try:
        window = int(args.cpu_window)
        tf, rf, nbases, nreads, ncontig = cpu_window_inputs(cfg, window, tmp)
        for _ in range(min(args.warmup, 1)):
            cpu_window_run(cfg, tf, rf, tmp, threads)
        t = 0.0
        for _ in range(args.steps):
            t += cpu_window_run(cfg, tf, rf, tmp, threads)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    val = nbases * args.steps / t / 1e9
    kind = cp.mapper_kind()
    line = {"impl": "reference", "metric": "long_read_gbp_per_s_sketched_mapped", "value": val, "unit": "Gbp/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": bench_config(cfg, world),
            "cpu_baseline": {"value": val, "unit": "Gbp/s", "cores": threads, "kind": kind,
                             "sample": cpu_sample_text(cfg, window, nreads, nbases, ncontig, kind)},
            "e2e": {"value": val, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def bench_config(cfg, world):
    "identical in both arms"
    return {"workload": cfg["workload"], "genome_bp": cfg["genome"], "read_bases_per_gpu_nominal": cfg["reads_per_gpu"],
            "k": cfg["k"], "w": cfg["w"], "z": Z, "sensitive": bool(cfg["sensitive"]),
            "step": "target sketch + index build + read sketch + lookup + chain + pair events + tally",
            "multi_gpu": "reads sharded over the ranks, target index replicated" if world > 1 else "n/a"}


# ------------------------------------------------------------------------------------------- parity
def pairs_digest(raw, gaps):
    "digest of the pair table in first-seen order with every pair's gap list in read order (the content of pairs.tsv)"
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(raw[:, :5]).tobytes())      # src, tgt, flags, n, anchor
    if len(raw):
        off = raw[:, 6].astype(np.int64) | (raw[:, 7].astype(np.int64) << 32)
        n = raw[:, 3].astype(np.int64)
        # the gap lists sit in hash-table slot order inside `gaps`: gather them pair by pair
        idx = np.repeat(off - np.concatenate(([0], np.cumsum(n)[:-1])), n) + np.arange(int(n.sum()))
        h.update(np.ascontiguousarray(gaps[idx]).tobytes())
    return h.hexdigest()[:16]


def parity_check(cfg, ctx, contigs_host, n_sub, local_rank):
    """rank 0, outside the timed regions: the first n_sub resident reads through (a) the timed entry points (resident target +
    resident reads on a second context) and (b) ntl_map_reads, against the CPU pipeline on the same FASTA files."""
    from ntlink_b200 import Context, pair, synth
    cp = cpu_pipeline()
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import util
    k, w = cfg["k"], cfg["w"]
    nreads, _ = ctx.resident_info(1)
    n_sub = min(n_sub, nreads)
    sub = ctx.resident_download(1, 0, n_sub, [f"read{i:09d}" for i in range(n_sub)])
    lengths = {n: int(l) for n, l in zip(contigs_host.names, contigs_host.lengths)}
    out = {"reads": n_sub, "read_bases": int(sub.offsets[-1]), "checker": cp.mapper_kind() + " mapper + C indexlr restatement"}
    tmp = scratch_dir()
    try:
        tf, rf = os.path.join(tmp, "target.fa"), os.path.join(tmp, "reads.fa")
        cp.write_fasta(tf, contigs_host)
        cp.write_fasta(rf, sub)
        threads = os.cpu_count() or 1
        tsv, _ = cp.sketch_target(tf, k, w, threads)
        cp.map_reads(tf, tsv, rf, os.path.join(tmp, "cpu"), k, w, Z, threads, sensitive=cfg["sensitive"], verbose=True, pairs=True, paf=True)
        want = cp.outputs(os.path.join(tmp, "cpu"))
        want_tsv = open(tsv, "rb").read()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    prm = ctx.params(k, w, Z, 10, 0.0, cfg["sensitive"], False)

    def files(c):
        prs = pair.filter_weak_anchor_pairs(pair.filter_pairs_distances(pair.pairs_dict(c.pairs(), contigs_host.names), lengths), 1)
        return pair.pairs_tsv(prs).encode(), pair.scaffold_dot(prs, lengths, 1).encode()

    # (b) the C-ABI call a user makes, host arrays in, all files out
    ctx.events_reset()
    tsk = ctx.build_index_from_sequences(contigs_host, k, w, want_sketch=True)
    res = ctx.map_reads(sub, prm, 0)
    ptsv, dot = files(ctx)
    out["target_tsv"] = tsk.to_tsv(contigs_host) == want_tsv
    out["verbose_mapping"] = res.verbose_bytes(sub, contigs_host) == want["verbose"]
    out["paf"] = res.paf_bytes(sub, sub.lengths.astype(np.uint32), contigs_host, k) == want["paf"]
    out["pairs_tsv"] = ptsv == want["pairs"]
    out["scaffold_dot"] = util.dot_parts(dot) == util.dot_parts(want["dot"])
    # (a) the timed entry points on the same reads: resident target + resident reads, second context
    c2 = Context(local_rank)
    try:
        c2.set_option("resident_chunk_bases", max(65536, int(sub.offsets[-1]) // 3))      # several chunks
        c2.target_upload(contigs_host)
        c2.reads_upload(sub)
        c2.events_reset()
        c2.index_build_resident(k, w)
        st = c2.map_resident(prm, 0)
        ptsv2, dot2 = files(c2)
        out["resident_pairs_tsv"] = ptsv2 == want["pairs"]
        out["resident_scaffold_dot"] = util.dot_parts(dot2) == util.dot_parts(want["dot"])
        out["resident_counts"] = (st["mx"], st["hits"], st["runs"], st["events"]) == (res.n_mx, res.n_hits, res.n_runs, res.n_events)
    finally:
        c2.close()
    out["ok"] = all(v for kk, v in out.items() if isinstance(v, bool))
    return out


# ------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args, rank, world, local_rank):
    import torch
    from ntlink_b200 import Context
    from ntlink_b200 import dist as nd
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist = nd.init_nccl(local_rank)                              # NCCL's own banner must not share stdout with the JSON line

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = pick_config(args, world)
    K, W = cfg["k"], cfg["w"]
    cplan, cnames, rplan = plans(cfg, rank)
    target_bases = int(cplan["len"].sum())
    shard_target = world > 1 and (args.shard_target == "always" or (args.shard_target == "auto" and target_bases >= (256 << 20)))
    ctx = Context(local_rank)
    for opt, env in (("cand_c", "NTL_CAND_C"), ("strip_len", "NTL_STRIP_LEN"), ("pipeline_min_bases", "NTL_PIPE_MIN"), ("async", "NTL_ASYNC"),
                     ("graph", "NTL_GRAPH"), ("resident_chunk_bases", "NTL_RESIDENT_CHUNK"), ("tile", "NTL_TILE")):   # sweeps / profiling only
        if os.environ.get(env):
            ctx.set_option(opt, float(os.environ[env]))
    prm = ctx.params(K, W, Z, 10, 0.0, cfg["sensitive"], False)
    t_setup = time.perf_counter()
    ctx.synth_target_resident(SEED, cplan, cnames)
    read_bases = ctx.synth_reads_resident(SEED, rplan)
    n_reads = len(rplan)
    t_setup = time.perf_counter() - t_setup
    # global read ordinals: rank blocks in rank order
    first_ordinal, total_bases, total_reads = 0, read_bases, n_reads
    if dist is not None:
        mine = torch.tensor([n_reads, read_bases], device="cuda", dtype=torch.int64)
        allc = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)
        first_ordinal = int(sum(int(c[0]) for c in allc[:rank]))
        total_reads = int(sum(int(c[0]) for c in allc))
        total_bases = int(sum(int(c[1]) for c in allc))
    xch = nd.GpuExchange(ctx, dist, rank, world) if world > 1 else None
    force_sync = bool(os.environ.get("NTL_GATHER_SYNC"))

    # ---------------- resident arm (value)
    stats = {}

    def step_resident(want_digest=False):
        t_a = time.perf_counter()
        ctx.events_reset()
        if shard_target:
            stats["index_mx"] = xch.build_index_sharded_resident(K, W)
        else:
            ctx.index_build_resident(K, W)
        st = ctx.map_resident(prm, first_ordinal=first_ordinal)
        t_b = time.perf_counter()
        if world > 1:
            xch.gather_events(force_sync)
        t_c = time.perf_counter()
        if rank == 0:
            raw, gaps = ctx.pairs_raw()
            stats["pairs"] = len(raw)
            if want_digest:
                stats["digest"] = pairs_digest(raw, gaps)
        t_d = time.perf_counter()
        for key, dt in (("t_map", t_b - t_a), ("t_gather", t_c - t_b), ("t_pairs", t_d - t_c)):
            stats[key] = stats.get(key, 0.0) + dt
        stats.update(st)

    for _ in range(args.warmup):
        step_resident()
    if world > 1 and not force_sync:
        xch.agree_capacity()
    step_resident(want_digest=True)                    # one more untimed step on the final exchange path (and the pair-table digest)
    for key in ("t_map", "t_gather", "t_pairs"):
        stats[key] = 0.0
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.timing_reset()
    t0 = time.perf_counter()
    ctx.mark(0)
    for _ in range(args.steps):
        step_resident()
    ctx.mark(1)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ctx.mark_elapsed_ms()
    tm = ctx.timing()
    # the timed region is bracketed by syncs; use the larger of (device events, host wall) so that host gaps count
    t_res = max(dev_ms / 1e3, wall)
    resident_digest = stats.get("digest")

    # ---------------- end-to-end arm (e2e): pinned host buffers in, host results out
    import psutil
    avail = psutil.virtual_memory().available
    e2e_reads = n_reads
    budget = int(0.35 * avail / max(1, world)) - target_bases
    if read_bases > budget:                             # host memory bound: a prefix of the rank's reads
        e2e_reads = max(1, int(n_reads * max(budget, 1 << 28) / read_bases))
    pc = ctx.resident_download(0, 0, len(cplan), cnames, pinned=True)
    pr = ctx.resident_download(1, 0, e2e_reads, [None] * e2e_reads, pinned=True)
    e2e_bases = int(pr.offsets[-1])
    d2h = {}

    def step_e2e(want_digest=False):
        import ctypes as C
        from ntlink_b200 import _lib
        ctx.events_reset()
        if shard_target:
            xch.build_index_sharded(pc, K, W)
        else:
            ctx.build_index_from_sequences(pc, K, W, want_sketch=False)
        mo = _lib.MapOut()
        ctx._check(ctx.lib.ntl_map_reads(ctx.h, pr.seq.ctypes.data, pr.offsets.ctypes.data, len(pr), first_ordinal, C.byref(prm), C.byref(mo)),
                   "ntl_map_reads")
        d2h["bytes"] = int(mo.n_hits) * 24 + int(mo.n_events) * 24 + (len(pr) + 1) * 16
        if world > 1:
            xch.gather_events(True)                     # buffers sized for the resident arm's counts; keep the checked path here
        if rank == 0:
            raw, gaps = ctx.pairs_raw()
            d2h["pairs"] = len(raw)
            if want_digest:
                d2h["digest"] = pairs_digest(raw, gaps)

    for _ in range(max(1, min(2, args.warmup // 2))):
        step_e2e(want_digest=True)
    barrier()
    graphs0, fallbacks0 = ctx.stat("graph_launches"), ctx.stat("async_fallbacks")
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    t_e2e = time.perf_counter() - t0
    clocks = sampler.summary()

    # ---------------- max over ranks
    e2e_total_bases = e2e_bases
    if dist is not None:
        t = torch.tensor([t_res, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_res, t_e2e = float(t[0]), float(t[1])
        b = torch.tensor([e2e_bases], device="cuda", dtype=torch.int64)
        dist.all_reduce(b)
        e2e_total_bases = int(b[0])
    if rank == 0:
        value = total_bases * args.steps / t_res / 1e9
        e2e = e2e_total_bases * args.steps / t_e2e / 1e9
        peak, peak_kind = measured_peak()
        n_mx = stats["mx"]
        # roofline of the dominant kernel over the read batches (target launches excluded). Algorithmic bytes per launch
        # (SURVEY.md 8d): 1.0 B/base of ASCII + 13 B per minimizer.
        dense_ms = tm["big_dense_ms"] / max(1, tm["big_dense_launches"])
        bases_per_launch = tm["big_dense_bases"] / max(1, tm["big_dense_launches"])
        alg_bytes = bases_per_launch * (1.0 + 13.0 * n_mx / read_bases)
        achieved = alg_bytes / (dense_ms * 1e-3) / 1e9 if dense_ms > 0 else 0.0
        traffic, ncu_pipes = None, None
        tp = os.path.join(REPO, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as fin:
                prof = json.load(fin)
            if prof.get("bases_per_launch"):
                traffic = prof.get("dram_bytes_per_launch", 0) / prof["bases_per_launch"] * bases_per_launch
            ncu_pipes = prof.get("pipes")
        config = bench_config(cfg, world)
        config.update({"contigs": len(cplan), "target_bases": target_bases, "read_bases_rank0": read_bases, "reads_rank0": n_reads,
                       "read_bases_all_ranks": total_bases, "reads_all_ranks": total_reads,
                       "l2": "inputs larger than L2 (%.1f GB of ASCII reads per GPU and step)" % (read_bases / 1e9),
                       "index": ("target sketched in contig shards + NCCL all-gather of the minimizer triples, replicated index built on every GPU"
                                 if shard_target else "target sketched and indexed on every GPU") if world > 1 else "built on the GPU every step",
                       "index_bytes": int(ctx.index_stats()["slots"]) * 16,
                       "setup_s": round(t_setup, 2)})
        line = {"metric": "long_read_gbp_per_s_sketched_mapped", "value": value, "unit": "Gbp/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": config,
                "e2e": {"value": e2e, "unit": "Gbp/s",
                        "h2d_bytes_per_step": int(len(pc.seq) + len(pr.seq) + 8 * (len(pc) + len(pr) + 2)),
                        "d2h_bytes_per_step": int(d2h.get("bytes", 0)), "ms_per_step": 1e3 * t_e2e / args.steps,
                        "read_bases_per_step_all_ranks": e2e_total_bases,
                        "reads": "all reads of the rank" if e2e_reads == n_reads else f"the first {e2e_reads} of {n_reads} reads of every rank (pinned host memory bound)",
                        "path": "sync-free call: %d chunk graphs launched, %d calls fell back to the synchronous path"
                                % (int(ctx.stat("graph_launches") - graphs0), int(ctx.stat("async_fallbacks") - fallbacks0))},
                "gpu_launches": int(tm["launches"]),
                "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": args.dominant_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "peak_kind": peak_kind,
                             "ms_per_launch": dense_ms, "launches": int(tm["big_dense_launches"]), "ncu": ncu_pipes,
                             "algorithmic_bytes_per_launch": alg_bytes, "bases_per_launch": bases_per_launch,
                             "whole_step": {"algorithmic_bytes": read_bases * (1.0 + 13.0 * n_mx / read_bases) + 44.0 * n_mx + 24.0 * stats["hits"],
                                            "frac": (read_bases * 1.0 + 57.0 * n_mx + 24.0 * stats["hits"]) / (t_res / args.steps) / 1e9 / peak},
                             "note": "integer-ALU bound kernel (rolling ntHash); HBM fraction reported as the metric asks; traffic = dram "
                                     "read+write bytes per launch from the ncu --set full capture summarised in profiles/traffic.json, scaled to this launch size"},
                "stage_ms_per_step": {k: tm[k] / args.steps for k in ("pack", "dense", "select", "gap", "emit", "lookup",
                                                                      "chain", "tally", "index")},
                "host_ms_per_step_rank0": {"index+map_resident": round(1e3 * stats["t_map"] / args.steps, 3),
                                           "event_exchange": round(1e3 * stats["t_gather"] / args.steps, 3),
                                           "tally+pairs": round(1e3 * stats["t_pairs"] / args.steps, 3)},
                "counts": {"read_minimizers": int(n_mx), "hits": int(stats["hits"]), "runs": int(stats["runs"]),
                           "events": int(stats["events"]), "pairs": int(stats.get("pairs", 0))}}
        parity = {"resident_vs_e2e_pair_table": (resident_digest == d2h.get("digest")) if (e2e_reads == n_reads and world == 1) else None}
        if not args.no_parity:
            try:
                if target_bases <= 400_000_000:
                    parity.update(parity_check(cfg, ctx, pc, args.parity_reads, local_rank))
                else:
                    parity["skipped"] = ("CPU checker on a 3.1 Gbp target takes tens of minutes; this configuration is checked by "
                                         "tests/test_gpu_synth.py (same k/w, N = 1 vs 2 GPUs vs CPU pipeline) and by the pair-table digests")
            except Exception as exc:       # a failed check must be visible in the line, not kill the measurement
                parity.update({"ok": False, "error": repr(exc)[:300]})
        line["parity"] = parity.get("ok", parity.get("resident_vs_e2e_pair_table")) and parity.get("resident_vs_e2e_pair_table") is not False
        line["parity_detail"] = parity
        if world == 1 and not args.no_cpu:
            tmp = scratch_dir()
            try:
                threads = os.cpu_count() or 1
                tf, rf, nb, nr, nc = cpu_window_inputs(cfg, int(args.cpu_window), tmp)
                dt = cpu_window_run(cfg, tf, rf, tmp, threads)
                kind = cpu_pipeline().mapper_kind()
                line["cpu_baseline"] = {"value": nb / dt / 1e9, "unit": "Gbp/s", "cores": threads, "kind": kind, "seconds": round(dt, 2),
                                        "sample": cpu_sample_text(cfg, int(args.cpu_window), nr, nb, nc, kind)}
            finally:
                shutil.rmtree(tmp, ignore_errors=True)
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 70 (configs[2]: 70 x 15 ms = a timed region of ~1 s for the resident arm); 10 for --impl reference (a CPU step is ~2 s)")
    ap.add_argument("--warmup", type=int, default=None, help="default: 5 (3 for --impl reference)")
    ap.add_argument("--impl", default="ntlink_b200", choices=["ntlink_b200", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS), help="default: c2 (configs[2]) up to 4 GPUs, c3 (configs[3]) at 8")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink genome and reads (smoke runs)")
    ap.add_argument("--reads-per-gpu", type=float, default=None, help="override the read bases per GPU (profiling runs)")
    ap.add_argument("--cpu-window", type=float, default=10e6, help="genome window (bp) of the CPU arm's scale model")
    ap.add_argument("--parity-reads", type=int, default=3000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--dominant-kernel", default="k_dense")
    ap.add_argument("--shard-target", default="auto", choices=["auto", "always", "never"],
                    help="N>1: sketch the target in contig shards + NCCL all-gather (auto: targets >= 256 Mbp)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 10 if args.impl == "reference" else 70
    if args.warmup is None:
        args.warmup = 3 if args.impl == "reference" else 5
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # started without a launcher: one process per GPU through torchrun, as the harness does
        import socket
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
