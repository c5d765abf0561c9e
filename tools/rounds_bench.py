#!/usr/bin/env python3
"""BASELINE.json configs[4]: three rounds of ntLink (ntLink_rounds) with mapping liftover and paf=True on a synthetic
500 Mbp assembly, the hot-path share of every round timed from HOST buffers through the public API:

  round 1   (ntLink:198-225)          target sketch + index, reads sketched and mapped, verbose_mapping.tsv and .paf text
                                      produced, pair table
  round 2,3 (ntLink_rounds:122-145)   sketch + index of the round's scaffolds, liftover of the previous round's mappings
                                      through the round's AGP (ntl_liftover_mappings: arrays in, lifted arrays + the lifted
                                      verbose_mapping.tsv text out), checkpoint tally of the lifted mappings on the device
                                      (ntl_tally_mappings), pair table

What joins the contigs between rounds (abyss-scaffold + ntlink_stitch_paths.py, SURVEY.md 8: out of scope) is replaced
by an untimed stand-in: greedy paths over the round's strongest pairs written as an AGP in the format of
<prefix>.trimmed_scafs.agp, and the scaffold sequences built from it on the host.

Several GPUs (torchrun): reads are sharded, every rank lifts and tallies its own reads' mappings, pair events are
gathered on rank 0 (ntlink_b200.dist.GpuExchange), rank 0 makes the AGP and broadcasts it.

Parity (inside the run, untimed): the first --parity-reads reads go through the same three rounds on the GPU on their
own and through the CPU (unmodified ntlink_pair.py / ntlink_liftover_mappings.py staged in oracle/_ref when present,
else the port); verbose_mapping.tsv and pairs.tsv of every round must be byte-identical. The CPU times of that sample
are the `cpu_baseline` of every round.

Prints one JSON line (rank 0)."""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402  (plans / seeds / scratch dir shared with bench.py)

COMP = np.zeros(256, np.uint8)
COMP[:] = ord("N")
for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
    COMP[a] = b


def greedy_agp(pairs, names, lengths, seed, min_n=2):
    """Stand-in for abyss-scaffold + stitch: contigs joined into paths along the strongest pairs (every contig at most two
    neighbours, no cycles), written like <prefix>.trimmed_scafs.agp. Contigs outside every path get an identity entry
    (path id == contig id), a few get no entry at all (both cases of liftover:63-66 / :84-85).
    Returns (agp lines, {new name: length})."""
    import random
    rng = random.Random(seed)
    idx = {n: i for i, n in enumerate(names)}
    edges = []
    for (src, _, tgt, _), (gaps, _) in pairs.items():
        if src != tgt and len(gaps) >= min_n:
            edges.append((len(gaps), idx[src], idx[tgt], int(np.median(gaps))))
    edges.sort(key=lambda e: (-e[0], e[1], e[2]))
    parent = list(range(len(names)))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    nbr = [[] for _ in names]
    for _, s, t, gap in edges:
        if len(nbr[s]) < 2 and len(nbr[t]) < 2 and find(s) != find(t):
            parent[find(s)] = find(t)
            nbr[s].append((t, gap))
            nbr[t].append((s, gap))
    from ntlink_b200.pair import largest_ntlink_scaffold_id
    largest = largest_ntlink_scaffold_id(names)                     # new path ids continue after the existing ones (pair:118-131)
    lines, new_len, seen, path_no = [], {}, [False] * len(names), 0 if largest is None else largest + 1
    for start in range(len(names)):
        if seen[start] or len(nbr[start]) != 1:
            continue
        members, prev, cur, gap_in = [], -1, start, 0
        while True:
            seen[cur] = True
            members.append((cur, gap_in))
            nxt = [(t, g) for t, g in nbr[cur] if t != prev]
            if not nxt:
                break
            prev, (cur, gap_in) = cur, nxt[0]
        path, pos, comp = f"ntLink_{path_no}", 1, 1
        path_no += 1
        for j, (c, gap) in enumerate(members):
            L = int(lengths[names[c]])
            if j:
                gap = min(max(gap, 20), 5000)
                lines.append(f"{path}\t{pos}\t{pos + gap - 1}\t{comp}\tN\t{gap}\tscaffold\tyes\tpaired-ends\n")
                pos += gap
                comp += 1
            a = 1 + (rng.randint(0, min(300, L // 4)) if rng.random() < 0.3 else 0)
            b = L - (rng.randint(0, min(300, L // 4)) if rng.random() < 0.3 else 0)
            lines.append(f"{path}\t{pos}\t{pos + b - a}\t{comp}\tW\t{names[c]}\t{a}\t{b}\t{rng.choice('+-')}\n")
            pos += b - a + 1
            comp += 1
        new_len[path] = pos - 1
    for c, name in enumerate(names):
        if seen[c]:
            continue
        L = int(lengths[name])
        new_len[name] = L
        if rng.random() >= 0.05:
            lines.append(f"{name}\t1\t{L}\t1\tW\t{name}\t1\t{L}\t+\n")
    return lines, new_len


def scaffold_sequences(old, agp_lines, new_names):
    "the round's scaffolds as a SeqBatch (order = new_names): pieces cut, reverse-complemented and N-joined as the AGP says"
    from ntlink_b200 import SeqBatch
    at = {n: i for i, n in enumerate(old.names)}
    off = old.offsets.astype(np.int64)
    pieces = {}
    for line in agp_lines:
        f = line.rstrip("\n").split("\t")
        if f[4] == "N":
            pieces.setdefault(f[0], []).append(np.full(int(f[5]), ord("N"), np.uint8))
            continue
        i = at[f[5]]
        s = old.seq[off[i] + int(f[6]) - 1: off[i] + int(f[7])]
        pieces.setdefault(f[0], []).append(COMP[s[::-1]] if f[8] == "-" else s)
    parts = []
    for name in new_names:
        if name in pieces:
            parts += [np.concatenate(pieces[name])] if len(pieces[name]) > 1 else pieces[name]
        else:
            i = at[name]
            parts.append(old.seq[off[i]:off[i + 1]])
    lens = np.array([len(p) for p in parts], np.uint64)
    offs = np.zeros(len(parts) + 1, np.uint64)
    offs[1:] = np.cumsum(lens)
    seq = np.empty(int(offs[-1]) + 64, np.uint8)
    seq[int(offs[-1]):] = ord("N")
    np.concatenate(parts, out=seq[:int(offs[-1])])
    return SeqBatch(seq[:int(offs[-1])], offs, list(new_names))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=float, default=500e6)
    ap.add_argument("--coverage", type=float, default=30.0, help="read bases = coverage x genome, split over the GPUs")
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("-k", type=int, default=32)
    ap.add_argument("-w", type=int, default=100)
    ap.add_argument("--parity-reads", type=int, default=3000)
    ap.add_argument("--threads", type=int, default=16, help="host threads of the text emitters")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()

    import torch
    from ntlink_b200 import Context, SeqBatch, liftover, pair
    from ntlink_b200 import dist as nd
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        dist = nd.init_nccl(local)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    K, W, Z = args.k, args.w, bench.Z
    cfg = dict(genome=int(args.genome), reads_per_gpu=int(args.genome * args.coverage / world))
    cplan, cnames, rplan = bench.plans(cfg, rank)
    ctx = Context(local)
    prm = ctx.params(K, W, Z)
    xch = nd.GpuExchange(ctx, dist, rank, world) if world > 1 else None

    # ---- inputs: generated on the device, then held in pinned host memory (what a reader would hand over)
    t0 = time.perf_counter()
    ctx.synth_target_resident(bench.SEED, cplan, cnames)
    read_bases = ctx.synth_reads_resident(bench.SEED, rplan)
    from ntlink_b200 import synth
    target = ctx.resident_download(0, 0, len(cplan), cnames, pinned=True)
    reads = ctx.resident_download(1, 0, len(rplan), synth.read_names(rplan), pinned=True)
    read_len = reads.lengths.astype(np.uint32)
    first_ordinal, total_bases, total_reads = 0, read_bases, len(rplan)
    if dist is not None:
        mine = torch.tensor([len(rplan), read_bases], device="cuda", dtype=torch.int64)
        allc = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)
        first_ordinal = int(sum(int(c[0]) for c in allc[:rank]))
        total_reads, total_bases = int(sum(int(c[0]) for c in allc)), int(sum(int(c[1]) for c in allc))
    t_setup = time.perf_counter() - t0

    def finish_pairs(names, lengths):
        "events of all ranks -> rank 0's filtered pairs dict (pair:241-255)"
        if xch is not None:
            xch.gather_events(True)
        if rank != 0:
            return None
        return pair.filter_weak_anchor_pairs(pair.filter_pairs_distances(pair.pairs_dict(ctx.pairs(), names), lengths), 1)

    def share(obj):
        if dist is None:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def timed(fn):
        barrier()
        t = time.perf_counter()
        out = fn()
        barrier()
        dt = time.perf_counter() - t
        if dist is not None:
            tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt[0])
        return out, dt

    def round_one(tgt, rds, rlen, ordinal, want_pairs=True):
        ctx.events_reset()
        ctx.build_index_from_sequences(tgt, K, W, want_sketch=False)
        res = ctx.map_reads(rds, prm, ordinal)
        verbose = res.verbose_bytes(rds, tgt, threads=args.threads)
        paf = res.paf_bytes(rds, rlen, tgt, K, threads=args.threads)
        lengths = dict(zip(tgt.names, (int(x) for x in tgt.lengths)))
        return res, verbose, paf, (finish_pairs(tgt.names, lengths) if want_pairs else None), lengths

    def later_round(prev, ids, scaffolds, rows, ordinal, want_pairs=True):
        ctx.events_reset()
        ctx.build_index_from_sequences(scaffolds, K, W, want_sketch=False)
        res = ctx.liftover_mappings(prev.hit_off, prev.nruns, prev.runs, prev.hits, rows, K)
        verbose = res.verbose_bytes(ids, scaffolds, threads=args.threads)
        ctx.tally_lifted(res.n_reads, prm, ordinal)
        lengths = dict(zip(scaffolds.names, (int(x) for x in scaffolds.lengths)))
        return res, verbose, (finish_pairs(scaffolds.names, lengths) if want_pairs else None), lengths

    rounds, agps, namespaces = [], [], [list(cnames)]
    # ---- round 1 (one warm-up pass: first-touch allocations of the workspaces, as in bench.py)
    round_one(target, reads, read_len, first_ordinal)
    (res, verbose, paf, pairs, lengths), dt = timed(lambda: round_one(target, reads, read_len, first_ordinal))
    rounds.append({"round": 1, "seconds": dt, "read_gbp_per_s": total_bases / dt / 1e9, "target_sequences": len(target),
                   "verbose_mapping_bytes_rank0": len(verbose), "paf_bytes_rank0": len(paf), "hits_rank0": int(res.n_hits),
                   "pairs": len(pairs) if pairs is not None else None})
    cur_target, cur_res = target, res
    ids = SeqBatch(np.empty(0, np.uint8), np.zeros(len(reads) + 1, np.uint64), reads.names)
    ids._name_blob = reads.name_blob()
    for rnd in range(2, args.rounds + 1):
        made = None
        if rank == 0:
            made = greedy_agp(pairs, cur_target.names, lengths, bench.SEED + rnd)
        agp_lines, new_len = share(made)
        agps.append(agp_lines)
        rows, new_names = liftover.agp_table(cur_target.names, liftover.read_agp(agp_lines))
        scaffolds = scaffold_sequences(cur_target, agp_lines, new_names)
        assert all(int(l) == new_len[n] for n, l in zip(scaffolds.names, scaffolds.lengths))
        namespaces.append(list(new_names))
        (res, verbose, pairs, lengths), dt = timed(lambda: later_round(cur_res, ids, scaffolds, rows, first_ordinal))
        rounds.append({"round": rnd, "seconds": dt, "read_gbp_per_s": total_bases / dt / 1e9, "target_sequences": len(scaffolds),
                       "agp_lines": len(agp_lines), "verbose_mapping_bytes_rank0": len(verbose), "hits_rank0": int(res.n_hits),
                       "pairs": len(pairs) if pairs is not None else None})
        cur_target, cur_res = scaffolds, res
    launches = ctx.timing()["launches"]

    # ---- parity + CPU times on the first reads of rank 0 (same AGPs, same scaffolds)
    parity, cpu = None, None
    if rank == 0 and not args.no_parity:
        sys.path.insert(0, os.path.join(REPO, "oracle"))
        import cpu_pipeline as cp
        n_sub = min(args.parity_reads, len(reads))
        end = int(reads.offsets[n_sub])
        sub = SeqBatch(reads.seq[:end], reads.offsets[:n_sub + 1].copy(), reads.names[:n_sub])
        sub_ids = SeqBatch(np.empty(0, np.uint8), np.zeros(n_sub + 1, np.uint64), sub.names)
        tmp = tempfile.mkdtemp(prefix="ntl_rounds_", dir=bench.scratch_dir())
        try:
            tf, rf = os.path.join(tmp, "t.fa"), os.path.join(tmp, "r.fa")
            cp.write_fasta(tf, target)
            cp.write_fasta(rf, sub)
            threads = min(32, os.cpu_count() or 1)
            tsv, t_sk = cp.sketch_target(tf, K, W, threads)
            p1 = os.path.join(tmp, "round1")
            t_map = cp.map_reads(tf, tsv, rf, p1, K, W, Z, threads, verbose=True, pairs=True, paf=True)
            out = cp.outputs(p1)
            xch_was = xch
            xch = None                                              # the sample runs on this GPU alone
            g_res, g_verbose, g_paf, g_pairs, g_len = round_one(target, sub, read_len[:n_sub], 0)
            parity = {"reads": n_sub, "read_bases": end, "mapper": cp.mapper_kind(),
                      "round1_verbose": out["verbose"] == g_verbose, "round1_paf": out["paf"] == g_paf,
                      "round1_pairs": out["pairs"] == pair.pairs_tsv(g_pairs).encode()}
            cpu = [{"round": 1, "seconds": t_map, "target_sketch_seconds": t_sk, "read_gbp_per_s": end / t_map / 1e9, "cores": threads}]
            prev_verbose, prev_target = p1 + ".verbose_mapping.tsv", target
            for rnd in range(2, args.rounds + 1):
                agp_lines, names = agps[rnd - 2], namespaces[rnd - 1]
                agp_path = os.path.join(tmp, f"round{rnd - 1}.agp")
                with open(agp_path, "w") as fout:
                    fout.writelines(agp_lines)
                scaffolds = scaffold_sequences(prev_target, agp_lines, names)
                sf = os.path.join(tmp, f"round{rnd}.fa")
                with open(sf, "w") as fout:                         # the checkpoint path only needs names and lengths
                    for name, L in zip(scaffolds.names, scaffolds.lengths):
                        fout.write(f">{name}\n{'N' * int(L)}\n")
                pfx = os.path.join(tmp, f"round{rnd}")
                t_lift, t_pair = cp.next_round(prev_verbose, agp_path, sf, pfx, K, Z)
                out = cp.outputs(pfx)
                rows, new_names = liftover.agp_table(prev_target.names, liftover.read_agp(agp_lines))
                assert new_names == names
                g_res, g_verbose, g_pairs, g_len = later_round(g_res, sub_ids, scaffolds, rows, 0)
                parity[f"round{rnd}_verbose"] = out["verbose"] == g_verbose
                parity[f"round{rnd}_pairs"] = out["pairs"] == pair.pairs_tsv(g_pairs).encode()
                cpu.append({"round": rnd, "seconds": t_lift + t_pair, "liftover_seconds": t_lift, "pair_seconds": t_pair,
                            "read_gbp_per_s": end / (t_lift + t_pair) / 1e9, "cores": 1})
                prev_verbose, prev_target = pfx + ".verbose_mapping.tsv", scaffolds
            parity["all"] = all(v for k, v in parity.items() if k.startswith("round"))
            xch = xch_was
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    if rank == 0:
        total = sum(r["seconds"] for r in rounds)
        print(json.dumps({"metric": "ntlink_rounds_hot_path_seconds", "value": total, "unit": "s", "higher_is_better": False,
                          "n_gpus": world, "rounds": rounds, "parity": parity, "cpu_baseline": cpu,
                          "gpu_launches": int(launches), "data": "synthetic",
                          "config": {"workload": f"configs[4]: ntLink_rounds x{args.rounds} with mapping liftover + paf on a synthetic "
                                                 f"{args.genome / 1e6:.0f} Mbp assembly, {args.coverage:g}x ONT-like reads, k={K} w={W} z={Z}",
                                     "contigs": len(cplan), "read_bases_all_ranks": total_bases, "reads_all_ranks": total_reads,
                                     "timed": "host buffers in, host results + text out, per round; joining stage between rounds untimed",
                                     "setup_s": round(t_setup, 2)}}))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
