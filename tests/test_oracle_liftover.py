"""oracle/liftover_oracle.py against the goldens made by the unmodified reference (bin/ntlink_liftover_mappings.py)."""
import gzip
import json
import os
import sys

import pytest

import util

sys.path.insert(0, os.path.join(util.REPO, "oracle"))
import liftover_oracle  # noqa: E402

LIFT = os.path.join(util.GOLD, "liftover")
with open(os.path.join(LIFT, "manifest.json")) as _f:
    MAN = json.load(_f)


def lift_file(case, what):
    return gzip.open(os.path.join(LIFT, case, what + ".gz"), "rt").read()


@pytest.mark.parametrize("case", sorted(MAN))
def test_liftover_oracle_against_reference(case):
    out = liftover_oracle.liftover(lift_file(case, "in.verbose_mapping.tsv").splitlines(True),
                                   lift_file(case, "in.agp").splitlines(True), MAN[case]["k"])
    assert "".join(out) == lift_file(case, "lifted.verbose_mapping.tsv")


def test_goldens_exercise_every_branch():
    for key in liftover_oracle.COVERAGE:
        liftover_oracle.COVERAGE[key] = 0
    for case in MAN:
        liftover_oracle.liftover(lift_file(case, "in.verbose_mapping.tsv").splitlines(True),
                                 lift_file(case, "in.agp").splitlines(True), MAN[case]["k"])
    assert all(v > 0 for v in liftover_oracle.COVERAGE.values()), liftover_oracle.COVERAGE


@pytest.mark.parametrize("case", sorted(c for c in MAN if MAN[c]["round2_ok"]))
def test_round2_tally_of_lifted_mappings(case):
    "ntLink_rounds:137-145: the lifted file is the checkpoint of the next round's ntlink_pair.py"
    import pair_oracle as po
    m = MAN[case]
    prm = po.Params(m["k"], m["z"], m["a"], m["f"], m["x"], m["n"], False, False)
    lengths = {n: int(l) for n, l in (x.split("\t") for x in lift_file(case, "round2.lengths.tsv").splitlines())}
    lines = lift_file(case, "lifted.verbose_mapping.tsv").splitlines(True)
    pairs = po.filter_pairs(po.retally_from_verbose(lines, lengths, prm), lengths, prm.a)
    assert "".join(po.pairs_tsv_lines(pairs)) == lift_file(case, "round2.pairs.tsv")
    assert util.dot_parts("".join(po.dot_lines(pairs, lengths, prm.n)).encode()) == \
        util.dot_parts(lift_file(case, "round2.scaffold.dot").encode())
