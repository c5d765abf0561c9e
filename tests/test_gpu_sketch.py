"""The CUDA sketch path (through the C ABI) against the C oracle and the reference's golden target sketches."""
import numpy as np
import pytest

import util
from test_emu_sketch import random_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from ntlink_b200 import Context
    c = Context(0)
    yield c
    c.close()


def compare(ctx, seq, offs, k, w, names=None, **opts):
    from ntlink_b200 import SeqBatch
    defaults = {"strip_len": 256, "cand_c": 10.0, "batch_bases": 1 << 30}
    defaults.update(opts)
    for name, v in defaults.items():
        ctx.set_option(name, v)
    batch = SeqBatch(seq, offs, names or [f"s{i}" for i in range(len(offs) - 1)])
    sk = ctx.sketch(batch, k, w)
    oh, op, os_, oo = util.oracle_sketch_batch(seq, offs, k, w)
    assert np.array_equal(sk.seq_off, oo), _first_diff(sk.seq_off, oo, "per-sequence offsets")
    assert np.array_equal(sk.hash, oh), _first_diff(sk.hash, oh, "hash")
    assert np.array_equal(sk.pos, op), _first_diff(sk.pos, op, "pos")
    assert np.array_equal(sk.strand, os_), _first_diff(sk.strand, os_, "strand")
    return sk, batch


def _first_diff(a, b, what):
    n = min(len(a), len(b))
    d = np.nonzero(a[:n] != b[:n])[0]
    i = int(d[0]) if len(d) else n
    return f"{what}: len {len(a)} vs {len(b)}, first diff at {i}: {a[i:i+4]} vs {b[i:i+4]}"


@pytest.mark.parametrize("name,k,w", [("scaffolds_1.fa", 32, 250), ("scaffolds_2.fa", 32, 100),
                                      ("scaffolds_3.fa", 24, 250), ("scaffolds_4.fa", 40, 100)])
def test_golden_target_tsv_byte_exact(ctx, tmp_path, name, k, w):
    from ntlink_b200 import read_sequences
    batch = read_sequences(util.fixture_file(tmp_path, name))
    for opt in ({"strip_len": 256}, {"strip_len": 64, "cand_c": 4.0}):
        for o, v in opt.items():
            ctx.set_option(o, v)
        sk = ctx.sketch(batch, k, w)
        assert sk.to_tsv(batch) == util.expected_output(f"{name}.k{k}.w{w}.tsv")
    ctx.set_option("strip_len", 256)
    ctx.set_option("cand_c", 10.0)


@pytest.mark.parametrize("reads,k,w", [("long_reads_1.fa", 32, 250), ("long_reads_2.fq", 32, 100),
                                       ("long_reads_4.fa", 40, 100), ("long_reads_3.fa", 20, 10)])
def test_fixture_reads_vs_oracle(ctx, tmp_path, reads, k, w):
    from ntlink_b200 import read_sequences
    path = util.fixture_file(tmp_path, reads)
    batch = read_sequences(path)
    names, seq, offs = util.load_fasta_batch(path)
    assert batch.names == names and np.array_equal(batch.offsets, offs) and np.array_equal(batch.seq, seq)
    _, b = compare(ctx, seq, offs, k, w, names=names)
    # the --len TSV of the reads equals the oracle executable's bytes
    sk = ctx.sketch(b, k, w)
    assert sk.to_tsv(b, with_len=True) == util.oracle_indexlr(path, k, w, length=True)


@pytest.mark.parametrize("k,w,S,c", [(32, 100, 256, 10.0), (32, 250, 64, 10.0), (24, 250, 256, 4.0), (40, 100, 16, 10.0),
                                     (15, 5, 256, 10.0), (20, 10, 32, 10.0), (7, 3, 8, 2.0), (33, 50, 24, 6.0),
                                     (32, 100, 256, 1.0), (100, 40, 64, 10.0), (5, 1, 8, 10.0), (32, 100, 64, 0.01)])
def test_random_sequences(ctx, k, w, S, c):
    rng = np.random.default_rng(k * 1000 + w)
    seq, offs = random_batch(rng, 60, 1, 9000)
    compare(ctx, seq, offs, k, w, strip_len=S, cand_c=c)


@pytest.mark.parametrize("k,w,S", [(32, 100, 256), (24, 250, 64), (20, 10, 16), (9, 30, 8)])
def test_invalid_bases(ctx, k, w, S):
    rng = np.random.default_rng(7 + k)
    seq, offs = random_batch(rng, 40, 50, 12000, p_n=0.002, n_run=2, lower=True)
    compare(ctx, seq, offs, k, w, strip_len=S)
    seq, offs = random_batch(rng, 10, 50, 3000, p_n=0.05)
    compare(ctx, seq, offs, 9, 4, strip_len=S)


def test_low_complexity_empty_and_short(ctx):
    parts = [b"A" * 3000, b"", b"ACACACACAC" * 300, b"ACGT", b"ACGT" * 700, b"N" * 100,
             (b"AAAAAAAAAACCCCCCCCCCGGGGGGGGGGTTTTTTTTTT" * 80), b"ACGTTGCA" * 400 + b"N" * 50 + b"T" * 1600]
    seq = np.frombuffer(b"".join(parts), np.uint8)
    offs = np.zeros(len(parts) + 1, np.uint64)
    offs[1:] = np.cumsum([len(p) for p in parts])
    for k, w, S in [(32, 100, 256), (16, 20, 32), (8, 5, 8), (32, 250, 64)]:
        compare(ctx, seq, offs, k, w, strip_len=S)
        compare(ctx, seq, offs, k, w, strip_len=S, cand_c=100.0)


def test_multi_batch_equals_single_batch(ctx):
    rng = np.random.default_rng(5)
    seq, offs = random_batch(rng, 200, 500, 20000)
    compare(ctx, seq, offs, 32, 100, batch_bases=200000)        # ~10 device batches
    compare(ctx, seq, offs, 32, 100, batch_bases=1 << 30)


def test_larger_than_l2_properties(ctx):
    "300 Mbp of reads (bigger than the 126 MB L2): checksums equal the multi-threaded oracle's"
    rng = np.random.default_rng(11)
    nseq = 20000
    lens = np.clip(rng.lognormal(np.log(12000), 0.7, nseq), 1000, 200000).astype(np.int64)
    total = int(lens.sum())
    seq = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=total)
    offs = np.zeros(nseq + 1, np.uint64)
    offs[1:] = np.cumsum(lens)
    from ntlink_b200 import SeqBatch
    ctx.set_option("strip_len", 256); ctx.set_option("cand_c", 10.0); ctx.set_option("batch_bases", 1 << 30)
    sk = ctx.sketch(SeqBatch(seq, offs, [str(i) for i in range(nseq)]), 32, 250)
    oh, op, os_, oo = util.oracle_sketch_batch(seq, offs, 32, 250, threads=32)
    assert len(sk.hash) == len(oh)
    assert np.array_equal(sk.seq_off, oo)
    assert int(np.bitwise_xor.reduce(sk.hash)) == int(np.bitwise_xor.reduce(oh))
    assert int(sk.pos.astype(np.uint64).sum()) == int(op.astype(np.uint64).sum())
    assert int(sk.strand.sum()) == int(os_.sum())
    # density ~ 2/(w+1) and strictly increasing positions inside every sequence
    assert abs(len(oh) / total - 2 / 251) < 2e-4
    d = np.diff(sk.pos.astype(np.int64))
    starts = sk.seq_off[1:-1].astype(np.int64)
    interior = np.ones(len(d), bool)
    interior[starts[(starts > 0) & (starts < len(sk.pos))] - 1] = False
    assert (d[interior] > 0).all()


def test_indexlr_cli_drop_in(tmp_path):
    "the indexlr-compatible executable: same argv as ntLink:199/223, TSV on stdout, file or stdin (gzip too)"
    import gzip
    import subprocess
    import sys
    env = dict(__import__("os").environ, PYTHONPATH=util.REPO)
    tgt = util.fixture_file(tmp_path, "scaffolds_1.fa")
    out = subprocess.check_output([sys.executable, util.REPO + "/bin/indexlr", "--long", "--pos", "--strand", "-k", "32",
                                   "-w", "250", "-t", "4", tgt], env=env)
    assert out == util.expected_output("scaffolds_1.fa.k32.w250.tsv")
    reads = util.fixture_file(tmp_path, "long_reads_4_top5.fa")
    gz = gzip.compress(open(reads, "rb").read())
    out = subprocess.run([sys.executable, "-m", "ntlink_b200.indexlr", "--long", "--pos", "--strand", "--len", "-k", "40",
                          "-w", "100", "-t", "2", "--chunk-bases", "30000", "-"], input=gz, env=env, cwd=util.REPO,
                         stdout=subprocess.PIPE, check=True).stdout
    assert out == util.oracle_indexlr(reads, 40, 100, length=True)
    out = subprocess.check_output([sys.executable, "-m", "ntlink_b200.indexlr", "--long", "--pos", "-k", "15", "-w", "5",
                                   tgt], env=env, cwd=util.REPO)
    assert out == util.oracle_indexlr(tgt, 15, 5, strand=False)


def test_btllib_shaped_iterator(ctx, tmp_path):
    "btllib.Indexlr / SeqReader shapes used by ntlink_patch_gaps.py (k20 w10) on top of the GPU sketcher"
    from ntlink_b200 import btllib_shim as btllib
    path = util.fixture_file(tmp_path, "long_reads_4_top5.fa")
    names, seq, offs = util.load_fasta_batch(path)
    oh, op, os_, oo = util.oracle_sketch_batch(seq, offs, 20, 10)
    with btllib.Indexlr(path, 20, 10, btllib.IndexlrFlag.LONG_MODE, 4, ctx=ctx) as idx:
        recs = list(idx)
    assert [r.id for r in recs] == names and [r.readlen for r in recs] == np.diff(offs).astype(int).tolist()
    for i, r in enumerate(recs):
        a, b = int(oo[i]), int(oo[i + 1])
        assert [m.out_hash for m in r.minimizers] == oh[a:b].tolist()
        assert [m.pos for m in r.minimizers] == op[a:b].tolist()
        assert [m.forward for m in r.minimizers] == [bool(x) for x in os_[a:b]]
    with btllib.SeqReader(path, btllib.SeqReaderFlag.LONG_MODE) as rd:
        got = [(r.id, r.seq) for r in rd]
    raw = seq.tobytes().decode()
    assert got == [(n, raw[int(offs[i]):int(offs[i + 1])]) for i, n in enumerate(names)]
