"""Host-side text emitters (csrc/emit.cpp: ntl_format_verbose / _paf / _sketch_tsv) without a GPU: the same bytes for every
thread count, for the copying and the zero-copy form, across calls of changing size (the emitters recycle their part and
output buffers), and equal to a plain Python rendering of the arrays (bin/ntlink_pair.py:382-388, indexlr's TSV)."""
import gc

import numpy as np
import pytest

from ntlink_b200.api import MapResult, SeqBatch, Sketch

STRAND = np.uint32(1 << 31)


def synthetic_result(rng, n_reads, ncontig, hits_per_run=6):
    runs_per = rng.integers(0, 4, n_reads).astype(np.uint32)
    nh = (runs_per * hits_per_run).astype(np.uint32)
    hit_off = np.zeros(n_reads + 1, np.uint32)
    hit_off[1:] = np.cumsum(nh)
    tot = int(hit_off[-1])
    runs = np.zeros((tot, 3), np.uint32)
    hits = np.zeros((tot, 3), np.uint32)
    j = np.arange(tot, dtype=np.uint32) - np.repeat(hit_off[:-1], nh)
    first = j < np.repeat(runs_per, nh)
    runs[first, 0] = rng.integers(0, ncontig, int(first.sum()))
    runs[first, 1] = j[first] * hits_per_run
    runs[first, 2] = hits_per_run
    cpos = np.sort(rng.integers(0, 30000, (max(1, tot // hits_per_run), hits_per_run)), axis=1).reshape(-1)[:tot]
    flip = (rng.integers(0, 2, tot).astype(np.uint32) << np.uint32(31))
    hits[:, 0] = rng.integers(0, ncontig, tot)
    hits[:, 1] = cpos.astype(np.uint32) | flip
    hits[:, 2] = (cpos + 100).astype(np.uint32) | STRAND
    m = MapResult.__new__(MapResult)
    m.n_reads, m.n_mx, m.n_hits, m.n_runs, m.n_events = n_reads, 0, tot, int(runs_per.sum()), 0
    m.hit_off, m.nruns, m.runs, m.hits = hit_off, runs_per, runs, hits
    m.ev_off, m.ev_cnt, m.events = np.zeros(n_reads + 1, np.uint32), np.zeros(n_reads, np.uint32), np.zeros((0, 6), np.uint32)
    reads = SeqBatch(np.empty(0, np.uint8), np.zeros(n_reads + 1, np.uint64), [f"read{i:07d}" for i in range(n_reads)])
    contigs = SeqBatch(np.empty(0, np.uint8), np.arange(ncontig + 1, dtype=np.uint64) * 40000, [f"ctg{i:05d}" for i in range(ncontig)])
    return m, reads, contigs


def python_verbose(m, reads, contigs):
    out = []
    for r in range(m.n_reads):
        base = int(m.hit_off[r])
        for i in range(int(m.nruns[r])):
            ctg, start, count = (int(x) for x in m.runs[base + i])
            toks = []
            for h in m.hits[base + start: base + start + count]:
                cp, rp = int(h[1]), int(h[2])
                toks.append(f"{cp & 0x7FFFFFFF}:{'+' if cp >> 31 else '-'}_{rp & 0x7FFFFFFF}:{'+' if rp >> 31 else '-'}")
            out.append(f"{reads.names[r]}\t{contigs.names[ctg]}\t{count}\t{' '.join(toks)}\n")
    return "".join(out).encode()


def test_verbose_and_paf_bytes_do_not_depend_on_threads_copies_or_buffer_reuse():
    rng = np.random.default_rng(12)
    last_view = None
    for n_reads in (3000, 60000, 500, 60000, 0, 20000):          # sizes go up and down: recycled buffers of every fit
        m, reads, contigs = synthetic_result(rng, n_reads, 300)
        want = python_verbose(m, reads, contigs) if n_reads <= 3000 else None
        ref = m.verbose_bytes(reads, contigs, threads=1)
        if want is not None:
            assert ref == want
        rl = np.full(n_reads, 200000, np.uint32)
        paf_ref = m.paf_bytes(reads, rl, contigs, 32, threads=1)
        for threads in (2, 8):
            assert m.verbose_bytes(reads, contigs, threads=threads) == ref
            view = m.verbose_bytes(reads, contigs, threads=threads, copy=False)
            assert bytes(view) == ref
            last_view = view                                          # a live view keeps its buffer out of the pool
            assert m.paf_bytes(reads, rl, contigs, 32, threads=threads) == paf_ref
            assert bytes(m.paf_bytes(reads, rl, contigs, 32, threads=threads, copy=False)) == paf_ref
        assert bytes(last_view) == ref                               # still intact after the other calls
    del last_view
    gc.collect()


def test_sketch_tsv_threads_and_views():
    rng = np.random.default_rng(5)
    for nseq in (1, 700, 40, 5000):
        counts = rng.integers(0, 40, nseq)
        off = np.zeros(nseq + 1, np.uint64)
        off[1:] = np.cumsum(counts)
        n = int(off[-1])
        sk = Sketch(rng.integers(0, 2**63, n, dtype=np.uint64), rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32), off)
        lens = rng.integers(1, 10**6, nseq)
        batch = SeqBatch(np.empty(0, np.uint8), np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64), [f"s{i} x"[: 2 + i % 5] for i in range(nseq)])
        want = []
        for i in range(nseq):
            toks = " ".join(f"{int(sk.hash[j])}:{int(sk.pos_strand[j]) & 0x7FFFFFFF}:{'+' if int(sk.pos_strand[j]) >> 31 else '-'}"
                            for j in range(int(off[i]), int(off[i + 1])))
            want.append(f"{batch.names[i]}\t{int(lens[i])}\t{toks}\n")
        want = "".join(want).encode()
        for threads in (1, 3, 8):
            assert sk.to_tsv(batch, with_len=True, threads=threads) == want
            assert bytes(sk.to_tsv(batch, with_len=True, threads=threads, copy=False)) == want
