"""Kernel LOGIC of the sketch pipeline (ntlink_b200/csrc/sketch_logic.cuh compiled for the host by tests/emu)
against the C oracle. CPU only -- the real CUDA path is checked by tests/test_gpu_*.py."""
import numpy as np
import pytest

import util


def check(seq, offs, k, w, **kw):
    eh, ep, eo, st = util.emu_sketch(seq, offs, k, w, **kw)
    oh, op, os_, oo = util.oracle_sketch_batch(seq, offs, k, w)
    assert np.array_equal(eo, oo), "per-sequence minimizer counts differ"
    assert np.array_equal(eh, oh), "hashes differ"
    assert np.array_equal(ep & 0x7FFFFFFF, op), "positions differ"
    assert np.array_equal(ep >> 31, os_), "strands differ"
    return st


@pytest.mark.parametrize("name,k,w", [("scaffolds_1.fa", 32, 250), ("scaffolds_2.fa", 32, 100),
                                      ("scaffolds_3.fa", 24, 250), ("scaffolds_4.fa", 40, 100),
                                      ("scaffolds_1.fa", 32, 100), ("scaffolds_3.fa", 20, 10), ("scaffolds_2.fa", 15, 5)])
def test_fixture_targets(tmp_path, name, k, w):
    _, seq, offs = util.load_fasta_batch(util.fixture_file(tmp_path, name))
    check(seq, offs, k, w)
    check(seq, offs, k, w, S=64)
    check(seq, offs, k, w, S=8, c=3.0)


def test_fixture_reads(tmp_path):
    _, seq, offs = util.load_fasta_batch(util.fixture_file(tmp_path, "long_reads_4.fa"))
    check(seq, offs, 40, 100)
    check(seq, offs, 32, 250, S=128)


def random_batch(rng, nseq, lo, hi, p_n=0.0, n_run=0, lower=False):
    seqs = []
    for _ in range(nseq):
        L = int(rng.integers(lo, hi))
        s = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=L)
        if p_n > 0:
            mask = rng.random(L) < p_n
            s[mask] = ord("N")
        for _ in range(n_run):
            if L > 10:
                a = int(rng.integers(0, L - 1))
                b = min(L, a + int(rng.integers(1, 400)))
                s[a:b] = ord("N")
        if lower:
            m = rng.random(L) < 0.3
            s[m] |= 0x20
        seqs.append(s)
    offs = np.zeros(nseq + 1, np.uint64)
    offs[1:] = np.cumsum([len(s) for s in seqs])
    return (np.concatenate(seqs) if seqs else np.empty(0, np.uint8)), offs


@pytest.mark.parametrize("k,w,S,c", [(32, 100, 256, 10.0), (32, 250, 64, 10.0), (24, 250, 256, 4.0), (40, 100, 16, 10.0),
                                     (15, 5, 256, 10.0), (20, 10, 32, 10.0), (7, 3, 8, 2.0), (33, 50, 24, 6.0),
                                     (32, 100, 256, 1.0), (100, 40, 64, 10.0), (5, 1, 8, 10.0)])
def test_random_sequences(k, w, S, c):
    rng = np.random.default_rng(k * 1000 + w)
    seq, offs = random_batch(rng, 40, 1, 6000)
    check(seq, offs, k, w, S=S, c=c)


@pytest.mark.parametrize("k,w,S", [(32, 100, 256), (24, 250, 64), (20, 10, 16), (9, 30, 8)])
def test_invalid_bases_span_windows(k, w, S):
    rng = np.random.default_rng(7 + k)
    seq, offs = random_batch(rng, 30, 50, 8000, p_n=0.002, n_run=2, lower=True)
    check(seq, offs, k, w, S=S)
    seq, offs = random_batch(rng, 10, 50, 3000, p_n=0.05)      # N every ~20 bases: few valid k-mers
    check(seq, offs, 9, 4, S=S)


def test_gaps_and_overflow_paths_are_exercised():
    rng = np.random.default_rng(99)
    seq, offs = random_batch(rng, 20, 3000, 9000)
    st = check(seq, offs, 32, 100, S=64, c=0.5)               # tiny threshold: most windows have no candidate
    assert st["gaps"] > 50
    st = check(seq, offs, 32, 100, S=64, c=10.0, cap_override=2)   # tiny slots: strips overflow into the pool
    assert st["ovf"] > 50
    st = check(seq, offs, 32, 100, S=64, c=0.001)             # whole sequences without a candidate
    assert st["gaps"] >= 20


def test_low_complexity_and_ties():
    parts = [b"A" * 3000, b"ACACACACAC" * 300, b"ACGT" * 700, (b"AAAAAAAAAACCCCCCCCCCGGGGGGGGGGTTTTTTTTTT" * 80),
             b"ACGTTGCA" * 400 + b"N" * 50 + b"TTTTTTTTTTTTTTTT" * 100]
    seq = np.frombuffer(b"".join(parts), np.uint8)
    offs = np.zeros(len(parts) + 1, np.uint64)
    offs[1:] = np.cumsum([len(p) for p in parts])
    for k, w, S in [(32, 100, 256), (16, 20, 32), (8, 5, 8), (32, 250, 64)]:
        check(seq, offs, k, w, S=S)
        check(seq, offs, k, w, S=S, c=100.0)                  # threshold saturated: every k-mer is a candidate


def test_short_and_empty_sequences():
    parts = [b"", b"ACGT", b"ACGTACGTAC", b"N" * 100, b"ACGTTGCATGCATGCAAGCTTGCA", b"ACGTTGCATGCATGCAAGCTTGCAT"]
    seq = np.frombuffer(b"".join(parts), np.uint8)
    offs = np.zeros(len(parts) + 1, np.uint64)
    offs[1:] = np.cumsum([len(p) for p in parts])
    for k, w in [(8, 4), (4, 2), (24, 1), (24, 2), (10, 15)]:
        check(seq, offs, k, w, S=8)
        check(seq, offs, k, w, S=256)
