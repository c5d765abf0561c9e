"""The dense-mode sketch kernels for small windows (csrc/small_kernel.cuh: k_small, option small=2, and k_stream, small=3; small=1 picks by w; w <= 16: the overlap stage k15/w5 and gap
filling k20/w10) against the C oracle: tile boundaries (tiles of 4096 k-mer positions), runs of N in and next to the halo,
low-complexity sequence that overflows the staging segments, sequences shorter than a window, and the generic sparse
path on the same input."""
import numpy as np
import pytest

import util
from test_emu_sketch import random_batch
from test_gpu_sketch import compare

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from ntlink_b200 import Context
    c = Context(0)
    yield c
    c.set_option("small", 1)
    c.close()


def small_batches(ctx):
    v = ctx.stat("small_batches")
    return v


@pytest.mark.parametrize("k,w", [(15, 5), (20, 10), (32, 16), (11, 2), (64, 7), (5, 3)])
def test_small_windows_random_and_tile_edges(ctx, k, w):
    rng = np.random.default_rng(k * 100 + w)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    # lengths around multiples of the tile: np = L - k + 1 in {4096 - 1, 4096, 4096 + 1, 8192, 8192 + w - 1, ...}
    lens = [4096 + k - 2, 4096 + k - 1, 4096 + k, 8192 + k - 1, 8192 + k - 1 + w - 1, 3 * 4096 + k + w, k + w - 2, k + w - 1, k + w, k - 1, 0, 1]
    parts = [acgt[rng.integers(0, 4, n)] for n in lens]
    seq, offs = random_batch(rng, 40, 1, 30000)
    offs = np.concatenate([offs, offs[-1] + np.cumsum([len(p) for p in parts], dtype=np.uint64)]).astype(np.uint64)
    seq = np.concatenate([seq] + parts)
    before = small_batches(ctx)
    compare(ctx, seq, offs, k, w, small=1)
    assert small_batches(ctx) > before, "the dense-mode kernel did not run"
    compare(ctx, seq, offs, k, w, small=2)                      # the tile form of the dense-mode kernel
    compare(ctx, seq, offs, k, w, small=3)                      # the streaming form
    compare(ctx, seq, offs, k, w, small=0)                      # the sparse path on the same input


@pytest.mark.parametrize("k,w", [(15, 5), (20, 10), (9, 16)])
def test_small_windows_with_runs_of_n(ctx, k, w):
    rng = np.random.default_rng(k + w)
    seq, offs = random_batch(rng, 30, 50, 20000, p_n=0.001, n_run=3, lower=True)
    compare(ctx, seq, offs, k, w, small=1)
    # N runs placed right at tile boundaries, inside the halo, and long enough to hide whole tiles
    acgt = np.frombuffer(b"ACGT", np.uint8)
    s = acgt[rng.integers(0, 4, 40000)].copy()
    for at, n in [(4096 - 3, 2), (4096 + k, 1), (8192 - w, w), (12288 - 1, 1), (12288 + k + w, 5), (16384, 9000), (30000, k)]:
        s[at:at + n] = ord("N")
    compare(ctx, s, np.array([0, len(s)], np.uint64), k, w, small=1)
    dense_n = acgt[rng.integers(0, 4, 20000)].copy()
    dense_n[rng.random(20000) < 0.03] = ord("N")
    compare(ctx, dense_n, np.array([0, 5000, 5000, 20000], np.uint64), k, w, small=1)
    compare(ctx, dense_n, np.array([0, 5000, 5000, 20000], np.uint64), k, w, small=2)
    compare(ctx, dense_n, np.array([0, 5000, 5000, 20000], np.uint64), k, w, small=3)
    compare(ctx, s, np.array([0, len(s)], np.uint64), k, w, small=3)
    compare(ctx, s, np.array([0, len(s)], np.uint64), k, w, small=2)


def test_small_windows_low_complexity_overflows_staging(ctx):
    "poly-A and short tandem repeats: every position is a minimizer candidate, the staging bound is exceeded and grown"
    parts = [b"A" * 20000, b"AC" * 9000, b"ACGTTGCA" * 2000 + b"N" * 50 + b"T" * 9000, b"ACG" * 5000]
    seq = np.frombuffer(b"".join(parts), np.uint8)
    offs = np.zeros(len(parts) + 1, np.uint64)
    offs[1:] = np.cumsum([len(p) for p in parts])
    for k, w in [(15, 5), (8, 3), (20, 10)]:
        compare(ctx, seq, offs, k, w, small=1)
        compare(ctx, seq, offs, k, w, small=2)
        compare(ctx, seq, offs, k, w, small=3)


def test_small_windows_larger_batch_equals_sparse_path(ctx):
    from ntlink_b200 import synth
    gen = synth.genome(3_000_000, 77)
    reads = synth.reads(gen, 4, 78)
    for k, w in [(15, 5), (20, 10)]:
        ctx.set_option("small", 0)
        b = ctx.sketch(reads, k, w)
        for mode in (2, 3, 1):
            ctx.set_option("small", mode)
            a = ctx.sketch(reads, k, w)
            assert np.array_equal(a.seq_off, b.seq_off) and np.array_equal(a.hash, b.hash) and np.array_equal(a.pos_strand, b.pos_strand)
        assert len(a.hash) > 0.25 * len(reads.seq) * 2 / (w + 1)
