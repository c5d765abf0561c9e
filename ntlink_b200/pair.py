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
    "PairInfo.get_gap_estimate (pair:70-74): int() of numpy's median -> truncation toward zero"
    return int(np.median(gaps))


def pairs_dict(gpu_pairs, names):
    """Device pair table -> the reference's insertion-ordered `pairs` dict:
    (source, source_ori, target, target_ori) -> (gap list, anchor)"""
    out = {}
    for src, tgt, flags, _, anchor, gaps in gpu_pairs:
        key = (names[src], "+" if flags & 1 else "-", names[tgt], "+" if flags & 2 else "-")
        out[key] = ([int(g) for g in gaps], int(anchor))
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
    p.add_argument("--batch-bases", type=float, default=1e9, help="bases per streamed read batch with --reads-fasta [1e9]")
    p.add_argument("-t", type=int, default=4, help="host threads for text output")
    return p.parse_args(argv)


class NtLink:
    "GPU-backed counterpart of the reference's NtLink class (bin/ntlink_pair.py:115-617)"

    def __init__(self, args, ctx=None):
        self.args = args
        self.ctx = ctx or api.Context(args.device)
        self.contigs = None          # SeqBatch of the target (names/lengths)
        self.lengths = None          # name -> length

    # -- M1
    def read_minimizers(self):
        "builds the device index (replaces pair:189-211); returns the number of unique target minimizers"
        a = self.args
        print(datetime.datetime.today(), ": Reading minimizers", a.s, file=sys.stdout)
        self.contigs = api.read_sequences(a.s)
        self.lengths = {n: int(l) for n, l in zip(self.contigs.names, self.contigs.lengths)}
        if a.sketch_target or not a.m:
            if a.w is None:
                raise NtlinkPairError("-w is required to sketch the target on the GPU")
            sk = self.ctx.build_index_from_sequences(self.contigs, a.k, a.w, want_sketch=bool(a.write_target_tsv))
            if a.write_target_tsv:
                with open(a.write_target_tsv, "wb") as fout:
                    fout.write(sk.to_tsv(self.contigs, with_len=False, threads=a.t))
        else:
            with (sys.stdin if a.m == "-" else open(a.m)) as fin:
                names, _, hashes, posf, off = parse_sketch_tsv(fin, with_len=False)
            idx = {n: i for i, n in enumerate(self.contigs.names)}
            ctg = np.repeat(np.array([idx[n] for n in names], np.uint32), np.diff(off).astype(np.int64))
            self.ctx.build_index(hashes, ctg, posf, self.contigs.lengths.astype(np.uint32), self.contigs.names)
        return self.ctx.index_stats()["unique"]

    # -- M2..M6, M8, M9
    def find_scaffold_pairs(self):
        "replaces pair:336-414; returns the ordered pairs dict"
        a = self.args
        print(datetime.datetime.today(), ": Finding pairs", file=sys.stdout)
        prm = self.ctx.params(a.k, a.w or 1, a.z, a.f, a.x, a.sensitive, a.repeat_filter)
        self.ctx.events_reset()
        vf = open(a.p + ".verbose_mapping.tsv", "wb") if a.verbose else None
        pf = open(a.p + ".paf", "wb") if a.paf else None
        ordinal = 0
        try:
            if a.reads_fasta:
                if a.w is None:
                    raise NtlinkPairError("-w is required with --reads-fasta")
                # streamed: batches of ~batch_bases, the next one is decoded by a background thread while this one is
                # on the GPU (a 60x human read set does not fit in host memory, and gzip decoding is the slowest stage)
                for reads in api.prefetch_batches(a.reads_fasta, int(a.batch_bases)):
                    res = self.ctx.map_reads(reads, prm, ordinal)
                    self._emit(res, reads, reads.lengths.astype(np.uint32), vf, pf)
                    ordinal += len(reads)
            else:
                for path in a.FILES:
                    with (sys.stdin if path == "-" else open(path)) as fin:
                        names, lens, hashes, posf, off = parse_sketch_tsv(fin, with_len=True)
                    reads = api.SeqBatch(np.empty(0, np.uint8), np.zeros(len(names) + 1, np.uint64), names)
                    res = self.ctx.map_sketch(api.Sketch(hashes, posf, off), lens, prm, ordinal)
                    self._emit(res, reads, lens, vf, pf)
                    ordinal += len(names)
        finally:
            if vf:
                vf.close()
            if pf:
                pf.close()
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
            vf.write(res.verbose_bytes(reads, self.contigs, threads=self.args.t))
        if pf is not None:
            pf.write(res.paf_bytes(reads, read_len, self.contigs, self.args.k, threads=self.args.t))

    def main(self):
        a = self.args
        print("Running pairing stage of ntLink ...\n")
        try:
            if os.path.isfile(a.p + ".verbose_mapping.tsv"):       # pair:565-566
                a.checkpoint = a.p + ".verbose_mapping.tsv"
            if a.checkpoint:
                print("Found checkpoint file, bypassing read mapping...\n")
                if a.paf:
                    print("Warning: --paf specified, but not compatible with checkpoint")
                pairs = self.find_scaffold_pairs_checkpoints()
            else:
                self.read_minimizers()
                pairs = self.find_scaffold_pairs()
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
        except Exception as exc:
            # same clean-up as the reference (pair:608-613): never leave a partial checkpoint behind
            for suffix, on in ((".verbose_mapping.tsv", a.verbose), (".paf", a.paf)):
                if on and not a.checkpoint and os.path.isfile(a.p + suffix) and not isinstance(exc, NtlinkPairError):
                    os.remove(a.p + suffix)
            raise NtlinkPairError("ntLink pairing stage encountered an error..") from exc


def main(argv=None):
    NtLink(parse_arguments(argv)).main()


if __name__ == "__main__":
    main()
