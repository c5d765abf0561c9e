"""The round-1 probe workload (20 Mbp genome, 30x ONT-like reads, k32 w100) for the small probes in this directory
(copy_interference, e2e_parts, pageable_probe, tally_probe). bench.py itself now runs BASELINE.json's configurations."""
import numpy as np

SEED, GENOME_BP, COVERAGE, K, W, Z = 20240502, 20_000_000, 30, 32, 100, 1000


def make_inputs(rank, world):
    from ntlink_b200 import synth
    gen = synth.genome(GENOME_BP, SEED)
    contigs = synth.assembly(gen, SEED + 7)
    reads = synth.reads(gen, COVERAGE, SEED + 1 + 1000 * rank, first_id=rank * 10_000_000)
    return contigs, reads


def pinned_copy(batch):
    "same SeqBatch with its arrays in pinned host memory"
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
