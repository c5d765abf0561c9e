"""The single-pass sketch kernel (csrc/tile_kernel.cuh) against the C oracle, with the counters that prove it ran: every
case checks that the batch took the tile path (stat tile_batches) and did not fall back to the multi-pass pipeline."""
import numpy as np
import pytest

import util
from test_emu_sketch import random_batch

pytestmark = pytest.mark.gpu
ACGT = np.frombuffer(b"ACGT", np.uint8)


@pytest.fixture(scope="module")
def ctx():
    from ntlink_b200 import Context
    c = Context(0)
    yield c
    c.close()


def check(ctx, seq, offs, k, w, c=7.0, allow_fallback=False):
    from ntlink_b200 import SeqBatch
    ctx.set_option("cand_c", c)
    ctx.set_option("tile", 1)
    ctx.set_option("small", 0)                 # w <= 16 would otherwise go to the dense-mode kernels (tests/test_gpu_small.py)
    t0, f0 = ctx.stat("tile_batches"), ctx.stat("tile_fallbacks")
    batch = SeqBatch(seq, offs, [f"s{i}" for i in range(len(offs) - 1)])
    sk = ctx.sketch(batch, k, w)
    assert ctx.stat("tile_batches") > t0, "the tile path did not run"
    if not allow_fallback:
        assert ctx.stat("tile_fallbacks") == f0, "the tile path fell back to the multi-pass pipeline"
    oh, op, os_, oo = util.oracle_sketch_batch(seq, offs, k, w)
    assert np.array_equal(sk.seq_off, oo), ("offsets", len(sk.hash), len(oh))
    assert np.array_equal(sk.hash, oh) and np.array_equal(sk.pos, op) and np.array_equal(sk.strand, os_)
    ctx.set_option("cand_c", 7.0)
    return len(oh)


def long_batch(rng, lens, p_n=0.0, runs=(), lower=False):
    seqs = []
    for L in lens:
        s = ACGT[rng.integers(0, 4, L)]
        if p_n:
            s[rng.random(L) < p_n] = ord("N")
        for run in runs:
            for _ in range(max(1, L // 40000)):
                a = int(rng.integers(0, max(1, L - run)))
                s[a:a + run] = ord("N")
        if lower:
            s[rng.random(L) < 0.3] |= 0x20
        seqs.append(s)
    offs = np.zeros(len(seqs) + 1, np.uint64)
    offs[1:] = np.cumsum([len(s) for s in seqs])
    return np.concatenate(seqs), offs


@pytest.mark.parametrize("k,w,c", [(32, 100, 7.0), (24, 250, 7.0), (32, 250, 7.0), (40, 100, 4.0), (15, 13, 3.0), (20, 26, 7.0), (24, 52, 7.0),
                                   (32, 256, 7.0), (32, 257, 7.0), (32, 300, 7.0), (28, 600, 7.0), (100, 64, 7.0), (32, 100, 1.0),
                                   (24, 250, 0.3), (32, 100, 0.01), (31, 33, 7.0)])
def test_random_multi_tile(ctx, k, w, c):
    rng = np.random.default_rng(k * 977 + w)
    lens = [int(x) for x in rng.integers(1, 5000, 30)] + [250000, 70000, 33000, 32279, 32280, 1, 0, 131072, 12000, 800]
    seq, offs = long_batch(rng, lens)
    n = check(ctx, seq, offs, k, w, c, allow_fallback=c < 4.0)     # few candidates = many exact-scan records: a tile may run out of them
    assert n > 0


@pytest.mark.parametrize("k,w,p_n,runs", [(32, 100, 0.001, (50, 300)), (24, 250, 0.0005, (1000, 10000)), (32, 250, 0.0, (200, 257, 5000)),
                                          (20, 26, 0.01, (10,)), (32, 100, 0.05, ()), (24, 250, 0.0, (40000,))])
def test_invalid_bases_and_scaffold_gaps(ctx, k, w, p_n, runs):
    rng = np.random.default_rng(k + w + len(runs))
    lens = [300000, 90000, 40000, 65536, 5000, 700, 200000]
    seq, offs = long_batch(rng, lens, p_n=p_n, runs=runs, lower=True)
    check(ctx, seq, offs, k, w)
    check(ctx, seq, offs, k, w, c=1.5, allow_fallback=True)


def test_low_complexity_and_tiny(ctx):
    rng = np.random.default_rng(3)
    parts = [b"A" * 70000, b"", b"ACACACACAC" * 9000, b"ACGT", b"ACGT" * 20000, b"N" * 1000,
             (b"AAAAAAAAAACCCCCCCCCCGGGGGGGGGGTTTTTTTTTT" * 2500), b"ACGTTGCA" * 400 + b"N" * 50 + b"T" * 40000,
             ACGT[rng.integers(0, 4, 50000)].tobytes() + b"G" * 3000 + ACGT[rng.integers(0, 4, 50000)].tobytes(),
             (ACGT[rng.integers(0, 4, 171)].tobytes() * 600)]
    seq = np.frombuffer(b"".join(parts), np.uint8)
    offs = np.zeros(len(parts) + 1, np.uint64)
    offs[1:] = np.cumsum([len(p) for p in parts])
    for k, w in [(32, 100), (24, 250), (16, 20)]:
        check(ctx, seq, offs, k, w, allow_fallback=True)


def test_many_short_sequences(ctx):
    rng = np.random.default_rng(9)
    seq, offs = random_batch(rng, 4000, 1, 900, p_n=0.001)
    check(ctx, seq, offs, 20, 26)
    check(ctx, seq, offs, 32, 100)
    seq, offs = random_batch(rng, 300, 200, 3000)
    check(ctx, seq, offs, 24, 250)


def test_deferred_path_and_mapping_use_the_tile_kernel(ctx, tmp_path):
    "ntl_map_reads (sync-free, chunk graphs) on a fixture: golden bytes, tile path, no fallback"
    from ntlink_b200 import read_sequences
    case = util.manifest()["f3_default"]
    contigs = read_sequences(util.fixture_file(tmp_path, case["target"]))
    reads = read_sequences(util.fixture_file(tmp_path, case["reads"]))
    ctx.set_option("pipeline_min_bases", 4e6)
    try:
        t0, f0, a0 = ctx.stat("tile_batches"), ctx.stat("tile_fallbacks"), ctx.stat("async_fallbacks")
        ctx.build_index_from_sequences(contigs, case["k"], case["w"], want_sketch=False)
        prm = ctx.params(case["k"], case["w"], case["z"], case["f"], case["x"])
        ctx.events_reset()
        res = ctx.map_reads(reads, prm, 0)
        assert res.verbose_bytes(reads, contigs) == util.golden_case("f3_default", "verbose_mapping.tsv")
        assert ctx.stat("tile_batches") >= t0 + 3 and ctx.stat("tile_fallbacks") == f0 and ctx.stat("async_fallbacks") == a0
    finally:
        ctx.set_option("pipeline_min_bases", 80 << 20)
