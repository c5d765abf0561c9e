#!/usr/bin/env python3
"""
GPU-backed drop-in for bin/ntlink_liftover_mappings.py of bcgsc/ntLink v1.3.11 (called between rounds,
ntLink_rounds:122-125): lifts <round N>.verbose_mapping.tsv over to the coordinates of the round-N scaffolds through
the round's AGP, so that round N+1 starts from the checkpoint path instead of mapping again.

Same command line (-m mappings, -a agp, -o output, -k k) and byte-identical output. The per-hit coordinate transform
(liftover:61-88) and the per-read regrouping (liftover:90-124) run in the CUDA kernel k_liftover behind
ntl_liftover_mappings; the host parses the two text files and prints the result with the library's verbose emitter.

`liftover_and_tally` is the fused form SURVEY.md 8(f) N1 asks for: lifted mappings stay on the device and are tallied
there (ntl_tally_mappings with resident inputs), i.e. rounds >= 2 of ntLink_rounds without the text round trip.
"""
import argparse
import os
import sys

import numpy as np

from . import api
from .pair import parse_verbose_mappings

AGP_IN, AGP_MINUS, AGP_KEEP = 1, 2, 4


def read_agp(lines):
    "liftover:40-50: contig id -> (path_id, scaf_start, ctg_start, ctg_end, orientation); gap lines skipped, last entry wins"
    agp = {}
    for line in lines:
        path_id, scaf_start, _, _, ctype, ctg_id, ctg_start, ctg_end, orientation = line.strip().split("\t")
        if ctype in ("N", "P"):
            continue
        agp[ctg_id] = (path_id, int(scaf_start), int(ctg_start), int(ctg_end), orientation)
    return agp


def agp_table(old_names, agp):
    """One ntl_agp_row per old contig + the names of the new namespace. A contig without an entry keeps its own name
    (liftover:65-66); names are matched as strings, exactly like the reference's groupby/dict keys."""
    new_index, new_names = {}, []

    def new_id(name):
        if name not in new_index:
            new_index[name] = len(new_names)
            new_names.append(name)
        return new_index[name]

    rows = np.zeros((len(old_names), 5), np.uint32)
    for i, name in enumerate(old_names):
        entry = agp.get(name)
        if entry is None:
            rows[i, 0] = new_id(name)
            continue
        path_id, scaf_start, ctg_start, ctg_end, orientation = entry
        flags = AGP_IN
        if orientation == "-":
            flags |= AGP_MINUS
        if path_id == name or orientation not in "+-" or not orientation:
            flags |= AGP_KEEP
        rows[i] = (new_id(path_id), flags, scaf_start, ctg_start, ctg_end)
    return rows, new_names


def load_mappings(mappings):
    """verbose_mapping.tsv -> (old contig names, arrays, read ids); contig ids are assigned in order of appearance.
    `mappings` is a path (parsed natively by the library) or an iterable of lines (parsed in Python)."""
    if isinstance(mappings, (str, bytes, os.PathLike)):
        batches = list(api.read_verbose_mappings(os.fspath(mappings), None, share_repeated=False))
        if not batches:
            empty = np.zeros((0, 3), np.uint32)
            return [], (np.zeros(1, np.uint32), np.zeros(0, np.uint32), empty, empty), []
        hit_off, nruns, runs, hits, _, ids, names = batches[0]
        return names, (hit_off, nruns, runs, hits), ids
    mapping_lines = list(mappings)
    old_index = {}
    for line in mapping_lines:
        old_index.setdefault(line.split("\t", 2)[1], len(old_index))
    hit_off, nruns, runs, hits, _, ids = parse_verbose_mappings(mapping_lines, old_index, share_repeated=False, with_ids=True)
    return list(old_index), (hit_off, nruns, runs, hits), ids


def liftover(ctx, mapping_lines, agp_lines, k, threads=4):
    "-> the lifted verbose_mapping.tsv as bytes"
    old_names, arrays, ids = load_mappings(mapping_lines)
    rows, new_names = agp_table(old_names, read_agp(agp_lines))
    res = ctx.liftover_mappings(*arrays, rows, k)
    reads = api.SeqBatch(np.empty(0, np.uint8), np.zeros(len(ids) + 1, np.uint64), ids)
    contigs = api.SeqBatch(np.empty(0, np.uint8), np.zeros(len(new_names) + 1, np.uint64), new_names)
    return res.verbose_bytes(reads, contigs, threads=threads)


def liftover_file(ctx, mappings_path, agp_lines, k, out_path, threads=4, batch_hits=50_000_000):
    """Streaming form for genome-scale files: the mappings are parsed natively in batches of ~batch_hits hits (cut at read
    boundaries), every batch is lifted on the GPU and its text appended to out_path. Contig ids are assigned in order of
    appearance and stay stable across batches; the AGP table is extended as new contigs show up."""
    agp = read_agp(agp_lines)
    with open(out_path, "wb") as fout:
        for hit_off, nruns, runs, hits, _, ids, old_names in api.read_verbose_mappings(os.fspath(mappings_path), None, share_repeated=False,
                                                                                     max_hits=batch_hits):
            rows, new_names = agp_table(old_names, agp)
            res = ctx.liftover_mappings(hit_off, nruns, runs, hits, rows, k)
            reads = api.SeqBatch(np.empty(0, np.uint8), np.zeros(len(ids) + 1, np.uint64), ids)
            contigs = api.SeqBatch(np.empty(0, np.uint8), np.zeros(len(new_names) + 1, np.uint64), new_names)
            fout.write(res.verbose_bytes(reads, contigs, threads=threads, copy=False))


def liftover_and_tally(ctx, mapping_lines, agp_lines, k, scaffold_lengths, prm):
    """Round N+1 without text in between: liftover, then the checkpoint tally (pair:437-488) of the lifted mappings, both
    on the device. scaffold_lengths: name -> length of the round-N scaffolds. Returns pair.pairs_dict-style raw pairs
    (before the two filters) keyed by scaffold names."""
    from .pair import pairs_dict
    old_names, arrays, ids = load_mappings(mapping_lines)
    rows, new_names = agp_table(old_names, read_agp(agp_lines))
    ctx.liftover_mappings(*arrays, rows, k, want_result=False)
    lengths = np.array([scaffold_lengths.get(n, 0) for n in new_names], np.uint32)
    none = np.empty(0, np.uint32)
    ctx.build_index(np.empty(0, np.uint64), none, none, lengths, new_names)
    ctx.events_reset()
    ctx.tally_lifted(len(ids), prm, 0)
    return pairs_dict(ctx.pairs(), new_names)


def main(argv=None):
    "same options as bin/ntlink_liftover_mappings.py:150-158"
    p = argparse.ArgumentParser(description="Liftover of ntLink mappings on a B200 (drop-in for ntlink_liftover_mappings.py)")
    p.add_argument("-m", "--mappings", help="Path to the verbose mappings file", required=True)
    p.add_argument("-a", "--agp", help="Path to the AGP file", required=True)
    p.add_argument("-o", "--output", help="Output file name", required=True)
    p.add_argument("-k", "--kmer", help="Kmer size", required=True, type=int)
    p.add_argument("-v", "--version", action="version", version="ntLink v1.3.11 (ntlink_b200 GPU path)")
    p.add_argument("--device", type=int, default=0)
    p.add_argument("--batch-hits", type=float, default=5e7, help="hits per streamed batch [5e7]")
    p.add_argument("-t", type=int, default=4, help="host threads for text output")
    args = p.parse_args(argv)
    ctx = api.Context(args.device)
    try:
        with open(args.agp, encoding="utf-8") as fin:
            agp_lines = fin.readlines()
        liftover_file(ctx, args.mappings, agp_lines, args.kmer, args.output, threads=args.t, batch_hits=int(args.batch_hits))
    finally:
        ctx.close()


if __name__ == "__main__":
    main(sys.argv[1:])
