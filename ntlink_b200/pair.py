#!/usr/bin/env python3
"""
GPU-backed drop-in for the pairing stage of ntLink (bin/ntlink_pair.py of bcgsc/ntLink v1.3.11).

Same command line, same output files (<p>.n<n>.scaffold.dot, <p>.verbose_mapping.tsv, <p>.paf, <p>.pairs.tsv),
but the hot loops -- NtLink.read_minimizers (pair:189-211), find_scaffold_pairs (pair:336-414),
ntlink_utils.get_accepted_anchor_contigs (utils:200-294), tally_pairs_from_mappings / add_pair (pair:315-334,416-435)
-- run as CUDA kernels behind libntlink_b200.so. What stays in Python is what the reference does once per run
on O(#pairs) data: the median gap estimate and the two pair filters (pair:70-74,241-255), pairs.tsv (pair:490-496)
and the scaffold graph in dot format (pair:118-155,263-305,498-506).

Two ways to feed reads:
  * FILES = indexlr TSV (`indexlr --long --pos --strand --len`), `-` for stdin: exactly the reference interface;
  * --reads-fasta FASTA/FASTQ[.gz] ...: the fused path -- reads are sketched on the GPU too and the TSV text
    never exists (replaces `gzip -cd reads | indexlr ... | ntlink_pair.py ... -`, ntLink:221-225).
The target is given by -m (its indexlr TSV) as in the reference, or sketched on the GPU from -s when
--sketch-target is set (then <target>.k<k>.w<w>.tsv can be written with --write-target-tsv PATH).
"""
import argparse
import datetime
import os
import re
import sys

import numpy as np

from . import api


class NtlinkPairError(Exception):
    "ntLink pair exception (same name as the reference's, bin/ntlink_pair.py:28-30)"


def reverse_orientation(ori):
    return "-" if ori == "+" else "+"


def gap_estimate(gaps):
    """PairInfo.get_gap_estimate (pair:70-74): int() of numpy's median -> truncation toward zero. Computed without numpy:
    the lists are a few integers long and np.median costs ~20 us a call, which at human scale (2 x 10^5 pairs, three
    calls each) was most of the pairing stage's host time. For an even count numpy averages the two middle values in
    float64, exactly what (a + b) / 2 does for integers of this size."""
    n = len(gaps)
    if n == 1:
        return int(gaps[0])
    s = sorted(gaps)
    m = n >> 1
    if n & 1:
        return int(s[m])
    return int((s[m - 1] + s[m]) / 2)


def pairs_dict(gpu_pairs, names):
    """Device pair table -> the reference's insertion-ordered `pairs` dict:
    (source, source_ori, target, target_ori) -> (gap list, anchor)"""
    out = {}
    for src, tgt, flags, _, anchor, gaps in gpu_pairs:
        key = (names[src], "+" if flags & 1 else "-", names[tgt], "+" if flags & 2 else "-")
        out[key] = (gaps.tolist() if hasattr(gaps, "tolist") else [int(g) for g in gaps], int(anchor))
    return out


def filter_pairs_distances(pairs, lengths):
    "pair:246-255"
    return {k: v for k, v in pairs.items()
            if not (gap_estimate(v[0]) <= -lengths[k[0]] or gap_estimate(v[0]) <= -lengths[k[2]])}


def filter_weak_anchor_pairs(pairs, min_anchor):
    "pair:241-244"
    return {k: v for k, v in pairs.items() if v[1] >= min_anchor}


def pairs_tsv(pairs):
    "pair:490-496 with PairInfo.__str__ (pair:80-83)"
    return "".join(f"{k[0]}{k[1]}\t{k[2]}{k[3]}\tn={len(v[0])}, gap_estimates={v[0]}, anchor={v[1]}\n"
                   for k, v in pairs.items())


def largest_ntlink_scaffold_id(names):
    "pair:118-131"
    largest = None
    rx = re.compile(r"^ntLink_(\d+)$")
    for name in names:
        m = rx.search(name)
        if m and (largest is None or int(m.group(1)) > largest):
            largest = int(m.group(1))
    return largest


def scaffold_dot(pairs, lengths, min_weight):
    """build_scaffold_graph + filter_graph_global + print_directed_graph (pair:263-305,498-506,133-155) without
    igraph: every pair contributes the edge source->target and the reverse-complement edge; edges are listed
    grouped by source in first-seen order; edges with n < min_weight are dropped; node lines are listed in
    first-seen order (the reference lists them in Python-set order, which is arbitrary)."""
    edges, nodes = {}, {}
    for key, info in pairs.items():
        s, t = key[0] + key[1], key[2] + key[3]
        rs, rt = key[2] + reverse_orientation(key[3]), key[0] + reverse_orientation(key[1])
        for v in (s, t, rs, rt):
            nodes.setdefault(v, None)
        if (s in edges and t in edges[s]) or (rs in edges and rt in edges[rs]):
            raise NtlinkPairError(f"duplicate edge {s} -> {t}")
        edges.setdefault(s, {})[t] = info
        edges.setdefault(rs, {})[rt] = info
    out = ["digraph G {\n", f"graph [scaf_num={largest_ntlink_scaffold_id(lengths)}]\n"]
    out.extend(f"\"{v}\" [l={lengths[v[:-1]]}]\n" for v in nodes)
    for s, targets in edges.items():
        for t, info in targets.items():
            if len(info[0]) >= min_weight:
                out.append(f"\"{s}\" -> \"{t}\" [d={gap_estimate(info[0])} e=100 n={len(info[0])}]\n")
    out.append("}\n")
    return "".join(out)


def parse_sketch_tsv(lines, with_len):
    """indexlr TSV -> (names, lengths or None, hash u64, pos_strand u32, seq_off u64). Records without a
    minimizer column are kept as empty sketches (the reference skips them, pair:200,357)."""
    names, lens, hashes, posf, off = [], [], [], [], [0]
    for line in lines:
        fields = line.rstrip("\n").split("\t")
        if not fields or fields == [""]:
            continue
        names.append(fields[0])
        col = 1
        if with_len:
            lens.append(int(fields[1]) if len(fields) > 1 and fields[1] else 0)
            col = 2
        if len(fields) > col and fields[col].strip():
            for tok in fields[col].strip().split(" "):
                h, p, s = tok.split(":")
                hashes.append(int(h))
                posf.append(int(p) | (0x80000000 if s == "+" else 0))
        off.append(len(hashes))
    return (names, np.array(lens, np.uint32) if with_len else None, np.array(hashes, np.uint64),
            np.array(posf, np.uint32), np.array(off, np.uint64))


def parse_verbose_mappings(lines, contig_index, share_repeated=True, with_ids=False):
    """verbose_mapping.tsv -> the arrays of Context.tally_mappings, following parse_verbose_entries (pair:466-488):
    one run per line, hit_count = number of listed hits (the num_hits column is not used), 'read length' = the largest
    first/last read position over the read's runs, reads = maximal blocks of consecutive lines with the same read id.
    share_repeated: a contig listed twice within a read refers both times to its LAST listing (the reference keeps one
    dict entry, pair:470-472); the liftover reads the same file line by line and passes False.
    with_ids: also return the read id of every block."""
    hit_off, nruns, read_len, runs, hits, ids = [0], [], [], [], [], []
    cur, group = None, []

    def flush():
        if not group:
            return
        base = hit_off[-1]
        last = {}
        local = []
        for n_line, (ctg, toks) in enumerate(group):
            start = len(hits) - base
            for tok in toks:
                c, r = tok.split("_")
                cp, cs = c.split(":")
                rp, rs = r.split(":")
                hits.append((ctg, int(cp) | (0x80000000 if cs == "+" else 0), int(rp) | (0x80000000 if rs == "+" else 0)))
            key = ctg if share_repeated else n_line
            last[key] = (ctg, start, len(toks))
            local.append(key)
        positions = []
        for key in local:
            run = last[key]
            runs.append(run)
            positions += [hits[base + run[1]][2] & 0x7FFFFFFF, hits[base + run[1] + run[2] - 1][2] & 0x7FFFFFFF]
        # holey layout: the read's region holds max(#hits, #runs) slots in both arrays
        width = max(len(hits) - base, len(local))
        hits.extend([(0, 0, 0)] * (base + width - len(hits)))
        runs.extend([(0, 0, 0)] * (base + width - len(runs)))
        nruns.append(len(local))
        read_len.append(max(positions))
        hit_off.append(base + width)
        ids.append(cur)

    for line in lines:
        read_id, contig_id, _, mx_hits = line.strip().split("\t")
        if read_id != cur:
            flush()
            cur, group = read_id, []
        group.append((contig_index[contig_id], mx_hits.split(" ")))
    flush()
    out = (np.array(hit_off, np.uint32), np.array(nruns, np.uint32), np.array(runs, np.uint32).reshape(-1, 3),
           np.array(hits, np.uint32).reshape(-1, 3), np.array(read_len, np.uint32))
    return out + (ids,) if with_ids else out


def parse_arguments(argv=None):
    "same options as bin/ntlink_pair.py:508-536, plus the fused-path options"
    p = argparse.ArgumentParser(description="ntLink pairing stage on a B200 (drop-in for ntlink_pair.py)")
    p.add_argument("FILES", nargs="*", help="Long read minimizer TSV files ('-' = stdin)")
    p.add_argument("-s", help="Target scaffolds fasta file", required=True)
    p.add_argument("-m", help="Target scaffolds minimizer TSV file", required=False)
    p.add_argument("-p", help="Output prefix [out]", default="out", type=str)
    p.add_argument("-n", help="Minimum edge weight [1]", default=1, type=int)
    p.add_argument("-k", help="Kmer size used for minimizer step", required=True, type=int)
    p.add_argument("-z", help="Minimum size of contig to scaffold", default=500, type=int)
    p.add_argument("-a", help="Minimum number of anchoring long reads for an edge", type=int, default=1)
    p.add_argument("-f", help="Maximum number of contigs in a run for full transitive edge addition", default=10, type=int)
    p.add_argument("-x", help="Fudge factor allowed between mapping block lengths on read and assembly", type=float, default=0)
    p.add_argument("-c", "--checkpoint", help="Mappings checkpoint file", required=False)
    p.add_argument("--pairs", help="Output pairs TSV file", action="store_true")
    p.add_argument("--paf", help="Output mappings in PAF-like format", action="store_true")
    p.add_argument("--sensitive", help="Run more sensitive read mapping", action="store_true")
    p.add_argument("--repeat-filter", help="Remove repetitive minimizers within a long read's sketch", action="store_true")
    p.add_argument("-v", "--version", action="version", version="ntLink v1.3.11 (ntlink_b200 GPU path)")
    p.add_argument("--verbose", help="Verbose output logging", action="store_true")
    # fused path
    p.add_argument("-w", help="Window size (needed when sketching on the GPU)", type=int, default=None)
    p.add_argument("--reads-fasta", nargs="+", help="Long reads FASTA/FASTQ[.gz]: sketch them on the GPU", default=None)
    p.add_argument("--sketch-target", action="store_true", help="Sketch the target (-s) on the GPU instead of reading -m")
    p.add_argument("--write-target-tsv", help="Also write the target sketch TSV here", default=None)
    p.add_argument("--device", type=int, default=0)
    p.add_argument("--gpus", type=int, default=1, help="GPUs of this node: read batches are dealt to the ranks, the target index is "
                                                       "replicated, pair events are gathered over NCCL (one process per GPU, torchrun)")
    p.add_argument("--batch-bases", type=float, default=1e9, help="bases per streamed read batch with --reads-fasta [1e9]")
    p.add_argument("--batch-minimizers", type=float, default=5e7, help="minimizers per streamed batch of a read TSV [5e7]")
    p.add_argument("-t", type=int, default=4, help="host threads for text output")
    a = p.parse_args(argv)
    # the reference declares FILES with nargs='+' (pair:510); here reads may also come from --reads-fasta or a checkpoint
    if not a.FILES and not a.reads_fasta and not a.checkpoint and not os.path.isfile(a.p + ".verbose_mapping.tsv"):
        p.error("the following arguments are required: FILES (or --reads-fasta)")
    return a


class NtLink:
    "GPU-backed counterpart of the reference's NtLink class (bin/ntlink_pair.py:115-617)"

    def __init__(self, args, ctx=None, rank=0, world=1, dist=None):
        self.args = args
        self.rank, self.world, self.dist = rank, world, dist
        self.ctx = ctx or api.Context(args.device)
        self.contigs = None          # SeqBatch of the target (names/lengths)
        self.lengths = None          # name -> length
        self.xch = None
        self.created = []            # output files this run created (removed again if the run fails)
        if world > 1:
            from . import dist as nd
            self.xch = nd.GpuExchange(self.ctx, dist, rank, world)

    # -- M1
    def read_minimizers(self):
        "builds the device index (replaces pair:189-211); returns the number of unique target minimizers"
        a = self.args
        print(datetime.datetime.today(), ": Reading minimizers", a.s, file=sys.stdout)
        self.contigs = api.read_sequences(a.s)
        self.lengths = {n: int(l) for n, l in zip(self.contigs.names, self.contigs.lengths)}
        if a.sketch_target or not a.m:
            if self.world > 1 and not a.write_target_tsv:
                # every rank sketches its contig shard, the triples are all-gathered, every GPU builds the same index
                self.xch.build_index_sharded(self.contigs, a.k, a.w)
            else:
                sk = self.ctx.build_index_from_sequences(self.contigs, a.k, a.w, want_sketch=bool(a.write_target_tsv))
                if a.write_target_tsv and self.rank == 0:
                    with open(a.write_target_tsv, "wb") as fout:
                        fout.write(sk.to_tsv(self.contigs, with_len=False, threads=a.t, copy=False))
        else:
            idx = {n: i for i, n in enumerate(self.contigs.names)}
            hs, ps, cs = [], [], []
            try:
                for names, _, sk in api.read_sketch_tsv(a.m, with_len=False, max_mx=int(a.batch_minimizers)):
                    hs.append(sk.hash)
                    ps.append(sk.pos_strand)
                    cs.append(np.repeat(np.array([idx[n] for n in names], np.uint32), np.diff(sk.seq_off).astype(np.int64)))
            except (ValueError, KeyError) as exc:
                raise NtlinkPairError(f"target minimizer file {a.m}: {exc}") from exc
            cat = lambda parts, dt: np.concatenate(parts) if parts else np.empty(0, dt)   # noqa: E731
            self.ctx.build_index(cat(hs, np.uint64), cat(cs, np.uint32), cat(ps, np.uint32), self.contigs.lengths.astype(np.uint32),
                                 self.contigs.names)
        return self.ctx.index_stats()["unique"]

    def _read_batches(self):
        "(reads SeqBatch, read lengths, Sketch or None) of every batch of the run, in input order -- the same on every rank"
        a = self.args
        if a.reads_fasta:
            # streamed: batches of ~batch_bases, the next one is decoded by a background thread while this one is on the
            # GPU (a 60x human read set does not fit in host memory, and gzip decoding is the slowest stage)
            for reads in api.prefetch_batches(a.reads_fasta, int(a.batch_bases)):
                yield reads, reads.lengths.astype(np.uint32), None
        else:
            # the reference's text interface (indexlr --len TSV, '-' = stdin), parsed natively in bounded batches
            for path in a.FILES:
                try:
                    for names, lens, sk in api.read_sketch_tsv(path, with_len=True, max_mx=int(a.batch_minimizers)):
                        yield api.SeqBatch(np.empty(0, np.uint8), np.zeros(len(names) + 1, np.uint64), names), lens, sk
                except ValueError as exc:
                    raise NtlinkPairError(f"{path}: {exc}") from exc

    # -- M2..M6, M8, M9
    def find_scaffold_pairs(self):
        "replaces pair:336-414; returns the ordered pairs dict (rank 0; None on the other ranks)"
        a = self.args
        print(datetime.datetime.today(), ": Finding pairs", file=sys.stdout)
        prm = self.ctx.params(a.k, a.w or 1, a.z, a.f, a.x, a.sensitive, a.repeat_filter)
        self.ctx.events_reset()
        single = self.world == 1
        outs = {}
        for key, suffix, on in (("v", ".verbose_mapping.tsv", a.verbose), ("p", ".paf", a.paf)):
            if on and (single or self.rank == 0):
                outs[key] = open(a.p + suffix, "wb")
                self.created.append(a.p + suffix)
        ordinal, parts = 0, []
        try:
            for i, (reads, lens, sk) in enumerate(self._read_batches()):
                if i % self.world == self.rank:          # batches are dealt round robin; global read ordinals keep the order
                    res = self.ctx.map_reads(reads, prm, ordinal) if sk is None else self.ctx.map_sketch(sk, lens, prm, ordinal)
                    if single:
                        self._emit(res, reads, lens, outs.get("v"), outs.get("p"))
                    else:
                        files = {}
                        try:
                            for key, on in (("v", a.verbose), ("p", a.paf)):
                                if on:
                                    path = f"{a.p}.part{i:06d}.{key}"
                                    self.created.append(path)
                                    files[key] = open(path, "wb")
                            self._emit(res, reads, lens, files.get("v"), files.get("p"))
                        finally:
                            for f in files.values():
                                f.close()
                parts.append(i)
                ordinal += len(reads)
            if not single:
                self.xch.gather_events(force_sync=True)      # every rank's pair events -> rank 0's event log
                self.dist.barrier()                          # all part files are complete
                if self.rank == 0:
                    for key, suffix in (("v", ".verbose_mapping.tsv"), ("p", ".paf")):
                        if key in outs:
                            for i in parts:
                                with open(f"{a.p}.part{i:06d}.{key}", "rb") as fin:
                                    while True:
                                        blk = fin.read(1 << 24)
                                        if not blk:
                                            break
                                        outs[key].write(blk)
                self.dist.barrier()
                for i in parts:
                    if i % self.world == self.rank:
                        for key in ("v", "p"):
                            path = f"{a.p}.part{i:06d}.{key}"
                            if os.path.isfile(path):
                                os.remove(path)
        finally:
            for f in outs.values():
                f.close()
        if self.rank != 0:
            return None
        return pairs_dict(self.ctx.pairs(), self.contigs.names)

    def find_scaffold_pairs_checkpoints(self, chunk_hits=50_000_000):
        """replaces pair:437-464: re-tally the pairs from the checkpoint verbose_mapping.tsv; the accepted runs go to
        the GPU in chunks cut at read boundaries and only the pair events are computed there"""
        a = self.args
        print(datetime.datetime.today(), ": Finding pairs", file=sys.stdout)
        self.contigs = api.read_sequences(a.s)
        self.lengths = {n: int(l) for n, l in zip(self.contigs.names, self.contigs.lengths)}
        # lengths and name ranks only: the checkpoint path never touches the target minimizers (pair:571-575)
        none = np.empty(0, np.uint32)
        self.ctx.build_index(np.empty(0, np.uint64), none, none, self.contigs.lengths.astype(np.uint32), self.contigs.names)
        prm = self.ctx.params(a.k, a.w or 1, a.z, a.f, a.x, a.sensitive, a.repeat_filter)
        self.ctx.events_reset()
        ordinal = 0
        # the file is parsed natively (ntl_verbose_*), in batches cut at read boundaries
        try:
            for hit_off, nruns, runs, hits, read_len, _, _ in api.read_verbose_mappings(a.checkpoint, self.contigs.names, share_repeated=True,
                                                                                        max_hits=chunk_hits):
                self.ctx.tally_mappings(hit_off, nruns, runs, hits, read_len, prm, ordinal)
                ordinal += len(nruns)
        except ValueError as exc:
            raise NtlinkPairError(str(exc)) from exc
        return pairs_dict(self.ctx.pairs(), self.contigs.names)

    def _emit(self, res, reads, read_len, vf, pf):
        if vf is not None:
            vf.write(res.verbose_bytes(reads, self.contigs, threads=self.args.t, copy=False))
        if pf is not None:
            pf.write(res.paf_bytes(reads, read_len, self.contigs, self.args.k, threads=self.args.t, copy=False))

    def main(self):
        a = self.args
        print("Running pairing stage of ntLink ...\n")
        ok = False
        try:
            if os.path.isfile(a.p + ".verbose_mapping.tsv"):       # pair:565-566
                a.checkpoint = a.p + ".verbose_mapping.tsv"
            if self.world > 1:
                # every rank must have looked before rank 0 creates this run's verbose_mapping.tsv, or a late rank would
                # take the new file for a checkpoint and never join the collectives
                self.dist.barrier()
            if not a.checkpoint and a.w is None and (a.reads_fasta or a.sketch_target or not a.m):
                raise NtlinkPairError("-w is required to sketch on the GPU (--reads-fasta / --sketch-target / no -m)")
            if a.checkpoint:
                print("Found checkpoint file, bypassing read mapping...\n")
                if a.paf:
                    print("Warning: --paf specified, but not compatible with checkpoint")
                pairs = self.find_scaffold_pairs_checkpoints() if self.rank == 0 else None
            else:
                self.read_minimizers()
                pairs = self.find_scaffold_pairs()
            if self.rank == 0:
                pairs = filter_pairs_distances(pairs, self.lengths)
                pairs = filter_weak_anchor_pairs(pairs, a.a)
                if a.pairs:
                    with open(a.p + ".pairs.tsv", "w") as fout:
                        fout.write(pairs_tsv(pairs))
                print(datetime.datetime.today(), ": Building scaffold graph", file=sys.stdout)
                out_graph = f"{a.p}.n{a.n}.scaffold.dot"
                print(datetime.datetime.today(), ": Printing graph", out_graph, sep=" ", file=sys.stdout)
                with open(out_graph, "w") as fout:
                    fout.write(scaffold_dot(pairs, self.lengths, int(a.n)))
                print(datetime.datetime.today(), ": DONE!", file=sys.stdout)
            ok = True
        except BaseException as exc:
            # The reference's bare `except:` (pair:608-613): whatever ends the run -- an error, a failed assertion of the
            # PAF writer, Ctrl-C, a kill -- must not leave a partial verbose_mapping.tsv behind, because the next run would
            # take it for a checkpoint (pair:565-566) and silently write a truncated graph.
            if isinstance(exc, (KeyboardInterrupt, SystemExit)):
                raise
            raise NtlinkPairError("ntLink pairing stage encountered an error..") from exc
        finally:
            if not ok:
                for path in self.created:
                    if os.path.isfile(path):
                        os.remove(path)


def main(argv=None):
    argv = sys.argv[1:] if argv is None else list(argv)
    args = parse_arguments(argv)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # one process per GPU through torchrun, same argv
        import socket
        import subprocess
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), "-m", "ntlink_b200.pair"] + argv
        rc = subprocess.call(cmd)
        if rc != 0:
            raise NtlinkPairError(f"multi-GPU run failed (torchrun exit code {rc})")
        return
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        from . import dist as nd
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        args.device = local
        dist = nd.init_nccl(local)
        try:
            NtLink(args, rank=rank, world=world, dist=dist).main()
        finally:
            dist.destroy_process_group()
        return
    NtLink(args).main()


if __name__ == "__main__":
    main()
