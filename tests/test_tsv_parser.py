"""Native indexlr-TSV parser (ntl_tsv_*, the reference's text interface ntLink:221-225 / bin/ntlink_pair.py:196-203,355-366)
against the Python restatement, and the host side of the synthetic-input generator. CPU only."""
import gzip
import os

import numpy as np
import pytest

import util


def test_native_tsv_parser_equals_python_restatement(tmp_path):
    from ntlink_b200 import api, pair
    reads = util.fixture_file(tmp_path, "long_reads_3.fa")
    tgt = util.fixture_file(tmp_path, "scaffolds_3.fa")
    tsv = util.oracle_indexlr(reads, 24, 250, length=True)
    p = str(tmp_path / "r.tsv")
    open(p, "wb").write(tsv)
    pn, pl, ph, pp, po = pair.parse_sketch_tsv(tsv.decode().splitlines(True), with_len=True)
    (n, l, sk), = list(api.read_sketch_tsv(p, True, 0))
    assert n == pn and np.array_equal(l, pl) and np.array_equal(sk.hash, ph) and np.array_equal(sk.pos_strand, pp) and np.array_equal(sk.seq_off, po)
    names, hs, offs = [], [], []
    for n, l, sk in api.read_sketch_tsv(p, True, 3000):          # bounded batches, whole records
        assert len(sk.hash) >= 1 and len(n) == len(l) == len(sk.seq_off) - 1
        names += n
        hs.append(sk.hash)
        offs.append(len(sk.hash))
    assert names == pn and np.array_equal(np.concatenate(hs), ph) and len(offs) > 5
    # gzip, no length column, no strand column
    t2 = util.oracle_indexlr(tgt, 15, 5, strand=False)
    pg = str(tmp_path / "t.tsv.gz")
    open(pg, "wb").write(gzip.compress(t2))
    (n, l, sk), = list(api.read_sketch_tsv(pg, False, 0))
    want = [ln.split("\t") for ln in t2.decode().splitlines()]
    assert n == [w[0] for w in want] and l is None
    toks = [t for w in want if len(w) > 1 for t in w[1].split(" ") if t]
    assert sk.hash.tolist() == [int(t.split(":")[0]) for t in toks] and sk.pos_strand.tolist() == [int(t.split(":")[1]) for t in toks]


@pytest.mark.parametrize("text", ["a\t12\t5:x:+\n", "a\t1z\t5:1:+\n", "a\t12\t5:1:?\n", "a\t12\t:1:+\n", "a\t12\t5:99999999999:+\n"])
def test_native_tsv_parser_rejects_malformed_input(tmp_path, text):
    from ntlink_b200 import api
    p = str(tmp_path / "bad.tsv")
    open(p, "w").write("ok\t100\t7:3:+ 9:10:-\n" + text)
    with pytest.raises(ValueError):
        list(api.read_sketch_tsv(p, True, 0))


def test_records_without_minimizers_and_blank_lines(tmp_path):
    from ntlink_b200 import api
    p = str(tmp_path / "e.tsv")
    open(p, "w").write("r1\t500\t\n\nr2\t10\nr3\t900\t11:0:+ 12:5:-   \nr4\n")
    (n, l, sk), = list(api.read_sketch_tsv(p, True, 0))
    assert n == ["r1", "r2", "r3", "r4"] and l.tolist() == [500, 10, 900, 0]
    assert sk.seq_off.tolist() == [0, 0, 0, 2, 2] and sk.hash.tolist() == [11, 12] and sk.pos_strand.tolist() == [0x80000000, 5]


def test_host_generator_is_deterministic_and_shaped_like_the_plan():
    from ntlink_b200 import synth
    cplan, names = synth.plan_assembly(2_000_000, 5, n_frac=0.5)
    rplan = synth.plan_reads(2_000_000, 6_000_000, 6, first_id=100)
    a, b = synth.host_contigs(9, cplan, names, threads=1), synth.host_contigs(9, cplan, names, threads=4)
    assert np.array_equal(a.seq, b.seq) and np.array_equal(a.offsets, b.offsets) and a.lengths.tolist() == cplan["len"].tolist()
    assert sorted(names) == [f"ctg{i:07d}" for i in range(len(names))] and names != sorted(names)
    r1, r2 = synth.host_reads(9, rplan, threads=1), synth.host_reads(9, rplan, threads=3)
    assert np.array_equal(r1.seq, r2.seq) and r1.names[0] == "read000000100"
    assert set(np.unique(r1.seq).tolist()) == {65, 67, 71, 84} and ord("N") in a.seq
    # a read is its genome slice with ~10 % edits: most 32-mers of an error-free read are found in the contig it came from
    clean = synth.host_reads(9, rplan[:20], err=(0, 0, 0))
    assert clean.lengths.tolist() == rplan["len"][:20].tolist()
    other = synth.host_reads(10, rplan[:20], err=(0, 0, 0))
    assert not np.array_equal(clean.seq, other.seq)
