"""Mapping liftover on the GPU (ntl_liftover_mappings) and round 2 on top of it, against goldens made by the unmodified
bin/ntlink_liftover_mappings.py and bin/ntlink_pair.py (checkpoint path)."""
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
MAN = util.lift_manifest()


@pytest.fixture(scope="module")
def ctx():
    from ntlink_b200 import Context
    c = Context(0)
    yield c
    c.close()


def write_inputs(tmp_path, case):
    m = os.path.join(str(tmp_path), "in.verbose_mapping.tsv")
    a = os.path.join(str(tmp_path), "in.agp")
    with open(m, "wb") as fout:
        fout.write(util.lift_golden(case, "in.verbose_mapping.tsv"))
    with open(a, "wb") as fout:
        fout.write(util.lift_golden(case, "in.agp"))
    return m, a


@pytest.mark.parametrize("case", sorted(MAN))
def test_liftover_cli_against_reference(tmp_path, case):
    from ntlink_b200 import liftover
    m, a = write_inputs(tmp_path, case)
    out = os.path.join(str(tmp_path), "lifted.tsv")
    liftover.main(["-m", m, "-a", a, "-o", out, "-k", str(MAN[case]["k"])])
    assert open(out, "rb").read() == util.lift_golden(case, "lifted.verbose_mapping.tsv")
    liftover.main(["-m", m, "-a", a, "-o", out, "-k", str(MAN[case]["k"]), "--batch-hits", "300"])      # streamed in many batches
    assert open(out, "rb").read() == util.lift_golden(case, "lifted.verbose_mapping.tsv")


def round2_lengths(case):
    return {n: int(l) for n, l in (x.split("\t") for x in util.lift_golden(case, "round2.lengths.tsv").decode().splitlines())}


@pytest.mark.parametrize("case", sorted(c for c in MAN if MAN[c]["round2_ok"]))
def test_round2_fused_liftover_and_tally(ctx, case):
    "liftover + checkpoint tally without leaving the device == reference liftover file fed to the reference's checkpoint path"
    from ntlink_b200 import liftover, pair
    m = MAN[case]
    lengths = round2_lengths(case)
    prm = ctx.params(m["k"], 1, m["z"], m["f"], m["x"])
    pairs = liftover.liftover_and_tally(ctx, util.lift_golden(case, "in.verbose_mapping.tsv").decode().splitlines(True),
                                        util.lift_golden(case, "in.agp").decode().splitlines(True), m["k"], lengths, prm)
    pairs = pair.filter_weak_anchor_pairs(pair.filter_pairs_distances(pairs, lengths), m["a"])
    assert pair.pairs_tsv(pairs).encode() == util.lift_golden(case, "round2.pairs.tsv")
    assert util.dot_parts(pair.scaffold_dot(pairs, lengths, m["n"]).encode()) == \
        util.dot_parts(util.lift_golden(case, "round2.scaffold.dot"))


@pytest.mark.parametrize("case", ["lift_syn_f3", "lift_syn_f2"])
def test_round2_through_the_files(tmp_path, case):
    "ntLink_rounds:122-145 as files: liftover CLI writes <prefix>.verbose_mapping.tsv, pair CLI takes its checkpoint path"
    from ntlink_b200 import liftover, pair
    m, a = write_inputs(tmp_path, case)
    prefix = os.path.join(str(tmp_path), "round2")
    liftover.main(["-m", m, "-a", a, "-o", prefix + ".verbose_mapping.tsv", "-k", str(MAN[case]["k"])])
    fa = os.path.join(str(tmp_path), "round2.fa")
    with open(fa, "w") as fout:
        for name, length in round2_lengths(case).items():
            fout.write(f">{name}\n{'A' * length}\n")
    c = MAN[case]
    pair.main(["-p", prefix, "-n", str(c["n"]), "-s", fa, "-k", str(c["k"]), "-a", str(c["a"]), "-z", str(c["z"]), "-f", str(c["f"]),
               "-x", str(c["x"]), "--pairs", "-m", "unused.tsv", "unused_reads.tsv"])
    assert open(prefix + ".pairs.tsv", "rb").read() == util.lift_golden(case, "round2.pairs.tsv")


def test_liftover_rejects_malformed_arrays(ctx):
    from ntlink_b200.api import NtlError
    hit_off = np.array([0, 2], np.uint32)
    runs = np.array([[0, 1, 1], [0, 0, 1]], np.uint32)
    rows = np.array([[0, 1, 1, 1, 100]], np.uint32)
    with pytest.raises(NtlError):
        ctx.liftover_mappings(hit_off, np.array([2], np.uint32), runs, np.zeros((2, 3), np.uint32), rows, 10)
    with pytest.raises(NtlError):     # contig id outside of the table
        ctx.liftover_mappings(hit_off, np.array([1], np.uint32), np.array([[5, 0, 1], [0, 0, 0]], np.uint32),
                              np.zeros((2, 3), np.uint32), rows, 10)
    # empty input is fine
    res = ctx.liftover_mappings(np.zeros(1, np.uint32), np.zeros(0, np.uint32), np.zeros((0, 3), np.uint32),
                                np.zeros((0, 3), np.uint32), rows, 10)
    assert res.n_reads == 0
