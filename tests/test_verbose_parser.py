"""Native verbose_mapping.tsv parser (ntl_verbose_*) against the Python restatement in ntlink_b200/pair.py (which the
checkpoint goldens pin). No GPU."""
import gzip
import os

import numpy as np
import pytest

import util
from ntlink_b200 import api, pair


def same(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a[:5], b[:5])) and list(a[5]) == list(b[5])


@pytest.mark.parametrize("name,gz", [("f3_default", False), ("f2_f1_a3", True), ("f3k20w10_default", False)])
def test_native_parser_equals_python(tmp_path, name, gz):
    text = util.golden_case(name, "verbose_mapping.tsv")
    path = os.path.join(str(tmp_path), "v.tsv" + (".gz" if gz else ""))
    with (gzip.open(path, "wb") if gz else open(path, "wb")) as fout:
        fout.write(text)
    lines = text.decode().splitlines(True)
    # ids in order of appearance (liftover)
    idx = {}
    for line in lines:
        idx.setdefault(line.split("\t")[1], len(idx))
    for share in (True, False):
        want = pair.parse_verbose_mappings(lines, idx, share_repeated=share, with_ids=True)
        got = list(api.read_verbose_mappings(path, None, share_repeated=share))
        assert len(got) == 1 and same(got[0], want) and got[0][6] == list(idx)
    # a given contig table (checkpoint path): ids follow it
    table = sorted(idx)
    tidx = {n: i for i, n in enumerate(table)}
    want = pair.parse_verbose_mappings(lines, tidx, with_ids=True)
    got = list(api.read_verbose_mappings(path, table))
    assert same(got[0], want)
    # streamed in batches: concatenation of the batches == the whole file
    parts = list(api.read_verbose_mappings(path, table, max_hits=500))
    assert len(parts) > 3
    cat = lambda k: np.concatenate([p[k] for p in parts])
    assert np.array_equal(cat(1), want[1]) and np.array_equal(cat(4), want[4]) and sum((p[5] for p in parts), []) == want[5]
    off = 0
    for p in parts:
        a, b = off, off + int(p[0][-1])
        assert np.array_equal(p[2], want[2][a:b]) and np.array_equal(p[3], want[3][a:b])
        off = b
    assert off == int(want[0][-1])


def test_native_parser_quirks_and_errors(tmp_path):
    lines = ["r1\tA\t2\t5:+_10:+ 9:+_20:+\n", "r1\tB\t1\t7:-_40:+\n", "r2\tA\t1\t1:+_3:-\n", "r1\tB\t1\t2:+_8:+\n",
             "r3\tA\t1\t5:+_10:+\n", "r3\tB\t1\t7:-_40:+\n", "  r3\tA\t2\t6:+_50:+ 8:+_60:+  \r\n"]
    path = os.path.join(str(tmp_path), "q.tsv")
    with open(path, "w", newline="") as fout:
        fout.writelines(lines)
    for share in (True, False):
        want = pair.parse_verbose_mappings(lines, {"A": 0, "B": 1}, share_repeated=share, with_ids=True)
        got = list(api.read_verbose_mappings(path, ["A", "B"], share_repeated=share))[0]
        assert same(got, want)
    with pytest.raises(ValueError, match="not in the target"):
        list(api.read_verbose_mappings(path, ["A"]))
    for bad in ("r\tA\t1\n", "r\tA\t1\t5:+_x:+\n", "r\tA\t1\t5:+_1:+\textra\n", "r\tA\t1\t5:*_1:+\n"):
        with open(path, "w") as fout:
            fout.write(bad)
        with pytest.raises(ValueError):
            list(api.read_verbose_mappings(path))
    open(path, "w").close()
    assert list(api.read_verbose_mappings(path)) == []
