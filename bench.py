#!/usr/bin/env python3
"""
bench.py -- long-read Gbp/s sketched+mapped on B200 (BASELINE.json metric), one JSON line on stdout.

Workload (config.workload): BASELINE.json configs[1] -- synthetic 5 Mbp genome cut into 1-200 kbp contigs,
30x simulated ONT reads (~10 % error), k=32 w=100, 1 GPU. One "step" = the whole job on that input:
target sketch + index build + read sketch + lookup + chaining + pair events + pair tally.

  value   inputs resident in HBM, device-side CUDA-event time on the library's stream, max over ranks
  e2e     the same job through the public API with pinned HOST buffers: H2D of target+reads and D2H of all
          mapping results inside the timed region
  roofline  the dominant kernel (k_dense): algorithmic bytes of the sketch (1.0 B/base + 13 B/minimizer,
          SURVEY.md 8d) / its mean launch time (CUDA events inside the library, same run) vs the measured HBM peak
  cpu_baseline  the CPU oracle (C sketcher on all host threads + the Python mapper, i.e. the shape of the
          reference pipeline `indexlr -t N | ntlink_pair.py`) on a bounded sample of the same reads

N > 1 (torchrun): reads are sharded (weak scaling: every rank maps its own 30x read set against the same target);
the target sketch is split over the ranks by contig and all-gathered with NCCL so every GPU builds the full
replicated index; pair events are gathered to rank 0 with NCCL and tallied there.

`--impl reference` times the CPU oracle port instead (the reference itself is Python + btllib and cannot run on
the GPU box: /root/reference is absent there and btllib is not installed).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

GENOME_BP = 5_000_000
COVERAGE = 30
K, W, Z = 32, 100, 1000
SEED = 20240502
WORKLOAD = ("configs[1]: synthetic 5 Mbp genome, 1-200 kbp contigs, 30x ONT-like reads (4% sub, 3% ins, 3% del), "
            "k=32 w=100 z=1000")


def make_inputs(rank, world):
    from ntlink_b200 import synth
    gen = synth.genome(GENOME_BP, SEED)
    contigs = synth.assembly(gen, SEED + 7)
    reads = synth.reads(gen, COVERAGE, SEED + 1 + 1000 * rank, first_id=rank * 10_000_000)
    return contigs, reads


def pinned_copy(batch):
    "same SeqBatch with its arrays in pinned host memory (torch is used for buffer management only)"
    import torch
    from ntlink_b200 import SeqBatch
    seq = torch.empty(len(batch.seq) + 64, dtype=torch.uint8, pin_memory=True)
    off = torch.empty(len(batch.offsets), dtype=torch.int64, pin_memory=True)
    s = seq.numpy()
    s[:len(batch.seq)] = batch.seq
    s[len(batch.seq):] = ord("N")
    o = off.numpy().view(np.uint64)
    o[:] = batch.offsets
    out = SeqBatch.__new__(SeqBatch)
    out.seq, out.offsets, out.names, out._name_blob = s[:len(batch.seq)], o, batch.names, None
    out._keep = (seq, off)
    return out


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


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_pipeline(contigs, reads, n_reads, threads):
    """oracle port of the reference pipeline on the first n_reads reads: C sketcher (all threads) for target and
    reads, TSV text in between, Python mapper (single thread, like bin/ntlink_pair.py). Returns (seconds, bases)."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import util
    import pair_oracle as po
    n_reads = min(n_reads, len(reads))
    sub_off = reads.offsets[:n_reads + 1]
    sub_seq = reads.seq[:int(sub_off[-1])]
    t0 = time.perf_counter()
    th, tp, ts, to = util.oracle_sketch_batch(contigs.seq, contigs.offsets, K, W, threads=threads)
    rh, rp, rs, ro = util.oracle_sketch_batch(sub_seq, sub_off, K, W, threads=threads)

    def tsv(names, h, p, s, off, lens=None):
        out = []
        for i, n in enumerate(names):
            a, b = int(off[i]), int(off[i + 1])
            toks = " ".join(f"{x}:{y}:{'+' if z else '-'}" for x, y, z in zip(h[a:b].tolist(), p[a:b].tolist(), s[a:b].tolist()))
            out.append(n + (f"\t{lens[i]}" if lens is not None else "") + "\t" + toks + "\n")
        return out

    t_lines = tsv(contigs.names, th, tp, ts, to)
    r_lines = tsv(reads.names[:n_reads], rh, rp, rs, ro, np.diff(sub_off))
    index = po.read_target_index(t_lines)
    lengths = {n: int(l) for n, l in zip(contigs.names, contigs.lengths)}
    prm = po.default_params(K, z=Z)
    pairs = po.filter_pairs(po.map_reads(r_lines, index, lengths, prm), lengths, 1)
    dt = time.perf_counter() - t0
    return dt, int(sub_off[-1]), len(pairs)


def run_reference(args, rank, world):
    if rank != 0:
        return
    contigs, reads = make_inputs(0, 1)
    threads = os.cpu_count() or 1
    n_reads = args.cpu_reads
    for _ in range(args.warmup):
        cpu_pipeline(contigs, reads, max(50, n_reads // 10), threads)
    t, bases = 0.0, 0
    for _ in range(args.steps):
        dt, nb, _ = cpu_pipeline(contigs, reads, n_reads, threads)
        t += dt
        bases += nb
    val = bases / t / 1e9
    line = {"impl": "reference", "metric": "long_read_gbp_per_s_sketched_mapped", "value": val, "unit": "Gbp/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "read_bases_per_gpu": int(reads.offsets[-1]), "reads_per_gpu": len(reads),
                       "contigs": len(contigs), "k": K, "w": W, "z": Z,
                       "step": "target sketch + index build + read sketch + lookup + chain + events + tally (CPU port of the reference path)"},
            "cpu_baseline": {"value": val, "unit": "Gbp/s", "cores": threads, "kind": "port",
                             "sample": f"{min(n_reads, len(reads))} of {len(reads)} reads ({bases // args.steps} bp) per step against the full 5 Mbp target; "
                                       "C oracle sketcher on all threads + single-threaded Python mapper (the reference's "
                                       "ntlink_pair.py has no -t)"},
            "e2e": {"value": val, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
def build_index_distributed(ctx, contigs, rank, world, dist, torch):
    """target sketch split over the ranks by contig, all-gathered over NCCL, index built on every GPU"""
    import ctypes as C
    from ntlink_b200 import SeqBatch, name_ranks
    from ntlink_b200 import dist as nd
    n = len(contigs)
    cum = contigs.offsets.astype(np.int64)
    a, b = nd.contig_shard(contigs.offsets, rank, world)
    part = SeqBatch(contigs.seq[int(cum[a]):int(cum[b])], contigs.offsets[a:b + 1] - contigs.offsets[a], contigs.names[a:b])
    sk = ctx.sketch(part, K, W)            # the triples also stay on the device; the host copy gives the contig ids
    m = len(sk.hash)
    dev = torch.device("cuda", torch.cuda.current_device())
    h = torch.empty(m, device=dev, dtype=torch.int64)
    nmx, dh, dp, do = C.c_uint64(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    ctx._check(ctx.lib.ntl_device_sketch_arrays(ctx.h, C.byref(nmx), C.byref(dh), C.byref(dp), C.byref(do)), "arrays")
    if m:
        ctx._check(ctx.lib.ntl_copy_device(ctx.h, h.data_ptr(), dh, m * 8), "copy")
    ctg = torch.from_numpy(np.repeat(np.arange(a, b, dtype=np.int32), np.diff(sk.seq_off).astype(np.int64))).to(dev)
    posf = torch.from_numpy(sk.pos_strand.view(np.int32)).to(dev)
    hashes, ctgs, posfs = nd.gather_triples(h, ctg, posf, dist)
    torch.cuda.synchronize()
    cl = contigs.lengths.astype(np.uint32)
    rk = name_ranks(contigs.names)
    ctx._check(ctx.lib.ntl_index_build_device(ctx.h, hashes.data_ptr(), ctgs.data_ptr(), posfs.data_ptr(), int(hashes.numel()),
                                              cl.ctypes.data, rk.ctypes.data, n), "ntl_index_build_device")


_gather_buf = {}


def gather_events_sync(ctx, rank, world, dist, torch):
    "the same exchange with the synchronising entry points (ntl_events_export / ntl_events_import_gathered); NTL_GATHER_SYNC=1"
    import ctypes as C
    dev = torch.device("cuda", torch.cuda.current_device())
    while True:
        cap = _gather_buf.get("cap", 8192)
        if _gather_buf.get("send") is None or _gather_buf["send"].shape[0] != (cap + 1) * 6:
            _gather_buf["send"] = torch.zeros((cap + 1) * 6, device=dev, dtype=torch.int32)
            _gather_buf["recv"] = torch.zeros(world * (cap + 1) * 6, device=dev, dtype=torch.int32)
        n = C.c_uint64()
        ctx._check(ctx.lib.ntl_events_export(ctx.h, _gather_buf["send"].data_ptr(), cap, C.byref(n)), "ntl_events_export")
        dist.all_gather_into_tensor(_gather_buf["recv"], _gather_buf["send"])
        counts = _gather_buf["recv"].view(world, -1)[:, 0].tolist()
        if max(counts) <= cap:
            if rank == 0:
                ovf = C.c_int(0)
                ctx._check(ctx.lib.ntl_events_import_gathered(ctx.h, _gather_buf["recv"].data_ptr(), world, cap, C.byref(ovf)),
                           "ntl_events_import_gathered")
            return
        _gather_buf["cap"] = int(max(counts)) * 2
        _gather_buf["send"] = None


def gather_events(ctx, rank, world, dist, torch):
    """pair events of every rank -> rank 0's device event log, in rank order = global read order: the library writes
    {count, events} into a fixed-capacity device buffer, ONE NCCL all_gather moves them, rank 0 imports the result.
    No host synchronisation at all: the library's stream and the collective's stream are ordered with events, the
    per-rank counts are interpreted on the device (ntl_events_import_device) and rank 0 learns the exact total together
    with the pair table. The capacity was agreed on by every rank during the warm-up steps (gather_events_sync reads
    the counts and grows the buffers); an overflow in a later step makes ntl_pairs_finish fail loudly."""
    import ctypes as C
    if os.environ.get("NTL_GATHER_SYNC") or not _gather_buf.get("agreed"):
        gather_events_sync(ctx, rank, world, dist, torch)
        return
    dev = torch.device("cuda", torch.cuda.current_device())
    if "stream" not in _gather_buf:
        sp = C.c_void_p()
        ctx._check(ctx.lib.ntl_stream(ctx.h, C.byref(sp)), "ntl_stream")
        _gather_buf["stream"] = torch.cuda.ExternalStream(sp.value, device=dev)
    lib_stream = _gather_buf["stream"]
    cap = _gather_buf["cap"]
    n = C.c_uint64()
    ctx._check(ctx.lib.ntl_events_export_async(ctx.h, _gather_buf["send"].data_ptr(), cap, C.byref(n)), "ntl_events_export_async")
    if n.value > cap:
        raise RuntimeError("event exchange buffer too small for this step; run more warm-up steps")
    torch.cuda.current_stream().wait_stream(lib_stream)
    dist.all_gather_into_tensor(_gather_buf["recv"], _gather_buf["send"])
    lib_stream.wait_stream(torch.cuda.current_stream())
    if rank == 0:
        ctx._check(ctx.lib.ntl_events_import_device(ctx.h, _gather_buf["recv"].data_ptr(), world, cap), "ntl_events_import_device")


def agree_on_gather_capacity(world, dist, torch):
    "after the warm-up: every rank takes the same capacity (4x the largest count seen, at least 8192 events)"
    if world == 1 or os.environ.get("NTL_GATHER_SYNC"):
        return
    dev = torch.device("cuda", torch.cuda.current_device())
    seen = int(_gather_buf["recv"].view(world, -1)[:, 0].max().item()) if _gather_buf.get("recv") is not None else 0
    t = torch.tensor([seen], device=dev, dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cap = max(8192, 4 * int(t.item()))
    _gather_buf["cap"] = cap
    _gather_buf["send"] = torch.zeros((cap + 1) * 6, device=dev, dtype=torch.int32)
    _gather_buf["recv"] = torch.zeros(world * (cap + 1) * 6, device=dev, dtype=torch.int32)
    torch.cuda.synchronize()
    _gather_buf["agreed"] = True


def run_gpu(args, rank, world, local_rank):
    import torch
    from ntlink_b200 import Context
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    contigs, reads = make_inputs(rank, world)
    # replicated target index: small targets are sketched redundantly on every GPU (cheaper than a collective);
    # large ones are sketched in contig shards and the minimizer triples all-gathered over NCCL
    shard_target = args.shard_target == "always" or (args.shard_target == "auto" and int(contigs.offsets[-1]) >= (256 << 20))
    ctx = Context(local_rank)
    for opt, env in (("cand_c", "NTL_CAND_C"), ("strip_len", "NTL_STRIP_LEN"),
                     ("pipeline_min_bases", "NTL_PIPE_MIN"), ("async", "NTL_ASYNC"), ("graph", "NTL_GRAPH")):   # sweeps / profiling only
        if os.environ.get(env):
            ctx.set_option(opt, float(os.environ[env]))
    prm = ctx.params(K, W, Z)
    read_bases = int(reads.offsets[-1])

    # ---------------- resident arm (value)
    ctx.target_upload(contigs)
    ctx.reads_upload(reads)
    stats = {}

    def step_resident():
        t_a = time.perf_counter()
        ctx.events_reset()
        if world == 1 or not shard_target:
            ctx.index_build_resident(K, W)
        else:
            build_index_distributed(ctx, contigs, rank, world, dist, torch)
        st = ctx.map_resident(prm, first_ordinal=rank * len(reads))
        t_b = time.perf_counter()
        if world > 1:
            gather_events(ctx, rank, world, dist, torch)
        t_c = time.perf_counter()
        if rank == 0:
            stats["pairs"] = len(ctx.pairs_raw()[0])
        t_d = time.perf_counter()
        for key, dt in (("t_map", t_b - t_a), ("t_gather", t_c - t_b), ("t_pairs", t_d - t_c)):
            stats[key] = stats.get(key, 0.0) + dt
        stats.update(st)

    for _ in range(args.warmup):
        step_resident()
    agree_on_gather_capacity(world, dist, torch)
    step_resident()                                    # one more untimed step on the final exchange path
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

    # ---------------- end-to-end arm (e2e): pinned host buffers in, host results out
    pc, pr = pinned_copy(contigs), pinned_copy(reads)
    d2h = {}

    def step_e2e():
        ctx.events_reset()
        if world == 1 or not shard_target:
            ctx.build_index_from_sequences(pc, K, W, want_sketch=False)
        else:
            build_index_distributed(ctx, pc, rank, world, dist, torch)
        import ctypes as C
        from ntlink_b200 import _lib
        mo = _lib.MapOut()
        ctx._check(ctx.lib.ntl_map_reads(ctx.h, pr.seq.ctypes.data, pr.offsets.ctypes.data, len(pr), rank * len(pr), C.byref(prm), C.byref(mo)),
                   "ntl_map_reads")
        d2h["bytes"] = int(mo.n_hits) * 24 + int(mo.n_events) * 24 + (len(pr) + 1) * 16
        if world > 1:
            gather_events(ctx, rank, world, dist, torch)
        if rank == 0:
            d2h["pairs"] = len(ctx.pairs_raw()[0])

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    graphs0, fallbacks0 = ctx.stat("graph_launches"), ctx.stat("async_fallbacks")
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    t_e2e = time.perf_counter() - t0
    clocks = sampler.summary()

    # ---------------- max over ranks
    if dist is not None:
        t = torch.tensor([t_res, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_res, t_e2e = float(t[0]), float(t[1])
    if rank == 0:
        total_bases = read_bases * world          # weak scaling: same-size shard per rank (rank 0's size x N)
        value = total_bases * args.steps / t_res / 1e9
        e2e = total_bases * args.steps / t_e2e / 1e9
        peak, peak_kind = measured_peak()
        n_mx = stats["mx"]
        # roofline of the dominant kernel: k_dense over the read batch (one launch per step; the 5 Mbp target launch is
        # excluded). Algorithmic bytes per launch: 1.0 B/base of ASCII + 13 B per minimizer.
        dense_ms = tm["big_dense_ms"] / max(1, tm["big_dense_launches"])
        bases_per_launch = tm["big_dense_bases"] / max(1, tm["big_dense_launches"])
        mx_per_base = n_mx / read_bases
        alg_bytes = bases_per_launch * (1.0 + 13.0 * mx_per_base)
        achieved = alg_bytes / (dense_ms * 1e-3) / 1e9 if dense_ms > 0 else 0.0
        traffic, ncu_pipes = None, None
        tp = os.path.join(REPO, "profiles", "r1_traffic.json")
        if os.path.exists(tp):
            with open(tp) as fin:
                prof = json.load(fin)
            traffic = prof.get("k_dense_reads_launch_dram_bytes")
            ncu_pipes = {"alu_pipe_pct": prof.get("k_dense_alu_pipe_pct"), "issue_slots_pct": prof.get("k_dense_issue_slots_pct"),
                         "dram_throughput_pct": prof.get("k_dense_dram_throughput_pct"), "source": "profiles/r1_k_dense_ncu_full.txt"}
        line = {"metric": "long_read_gbp_per_s_sketched_mapped", "value": value, "unit": "Gbp/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": WORKLOAD,
                           "read_bases_per_gpu": read_bases, "reads_per_gpu": len(reads), "contigs": len(contigs),
                           "k": K, "w": W, "z": Z, "l2": "inputs larger than L2 (150 MB ASCII reads per step)",
                           "step": "target sketch + index build + read sketch + lookup + chain + events + tally",
                           "multi_gpu": ("reads sharded, index replicated (%s), events gathered with NCCL to rank 0"
                                         % ("target sketched in contig shards + NCCL all-gather" if shard_target else
                                            "small target sketched on every GPU")) if world > 1 else "n/a"},
                "e2e": {"value": e2e, "unit": "Gbp/s",
                        "h2d_bytes_per_step": int(len(pc.seq) + len(pr.seq) + 8 * (len(pc) + len(pr) + 2)),
                        "d2h_bytes_per_step": int(d2h.get("bytes", 0)), "ms_per_step": 1e3 * t_e2e / args.steps,
                        "path": "sync-free call: %d chunk graphs launched, %d calls fell back to the synchronous path"
                                % (int(ctx.stat("graph_launches") - graphs0), int(ctx.stat("async_fallbacks") - fallbacks0))},
                "gpu_launches": int(tm["launches"]),
                "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "k_dense", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "peak_kind": peak_kind,
                             "ms_per_launch": dense_ms, "launches": int(tm["big_dense_launches"]), "ncu": ncu_pipes,
                             "algorithmic_bytes_per_launch": alg_bytes,
                             "note": "integer-ALU bound kernel (rolling ntHash: ncu alu pipe 80 %, issue slots 73 %, DRAM 11 %, "
                                     "profiles/r1_k_dense_ncu_full.txt); HBM fraction reported as the metric asks; traffic = "
                                     "dram read+write bytes of the same launch from ncu --set full (profiles/r1_traffic.json)"},
                "stage_ms_per_step": {k: tm[k] / args.steps for k in ("pack", "dense", "select", "gap", "emit", "lookup",
                                                                      "chain", "tally", "index")},
                "host_ms_per_step_rank0": {"index+map_resident": round(1e3 * stats["t_map"] / args.steps, 3),
                                           "event_exchange": round(1e3 * stats["t_gather"] / args.steps, 3),
                                           "tally+pairs": round(1e3 * stats["t_pairs"] / args.steps, 3)},
                "counts": {"read_minimizers": int(n_mx), "hits": int(stats["hits"]), "runs": int(stats["runs"]),
                           "events": int(stats["events"]), "pairs": int(stats.get("pairs", 0))}}
        if world == 1 and not args.no_cpu:
            dt, nb, _ = cpu_pipeline(contigs, reads, args.cpu_reads, os.cpu_count() or 1)
            line["cpu_baseline"] = {"value": nb / dt / 1e9, "unit": "Gbp/s", "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": f"{min(args.cpu_reads, len(reads))} of {len(reads)} reads ({nb} bp) against the full target, "
                                              f"{dt:.1f} s; C oracle sketcher on all threads + single-threaded Python mapper"}
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ntlink_b200", choices=["ntlink_b200", "reference"])
    ap.add_argument("--cpu-reads", type=int, default=1000000, help="reads in the bounded CPU sample (default: the whole 150 Mbp workload, ~2 s of CPU)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--shard-target", default="auto", choices=["auto", "always", "never"],
                    help="N>1: sketch the target in contig shards + NCCL all-gather (auto: targets >= 256 Mbp)")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # started without a launcher: one process per GPU through torchrun, as the harness does
        import socket
        import subprocess
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
