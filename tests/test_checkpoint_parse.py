"""Host logic of the checkpoint path (bin/ntlink_pair.py:437-488): verbose_mapping.tsv -> run/hit arrays."""
import numpy as np
import pytest

import util
from ntlink_b200 import pair

MAN = util.manifest()


@pytest.mark.parametrize("name", ["f3_default", "f2_f1_a3", "f3k20w10_default"])
def test_parse_verbose_mappings_roundtrip(name):
    lines = util.golden_case(name, "verbose_mapping.tsv").decode().splitlines(True)
    ctgs = sorted({l.split("\t")[1] for l in lines})
    idx = {n: i for i, n in enumerate(ctgs)}
    hit_off, nruns, runs, hits, read_len = pair.parse_verbose_mappings(lines, idx)
    assert len(hit_off) == len(nruns) + 1 == len(read_len) + 1
    assert nruns.sum() == len(lines)
    rebuilt, li = [], 0
    for r in range(len(nruns)):
        base = int(hit_off[r])
        first_last = []
        for i in range(int(nruns[r])):
            ctg, start, count = (int(v) for v in runs[base + i])
            toks = []
            for h in hits[base + start: base + start + count]:
                assert int(h[0]) == ctg
                toks.append(f"{int(h[1]) & 0x7FFFFFFF}:{'+' if h[1] >> 31 else '-'}_{int(h[2]) & 0x7FFFFFFF}:{'+' if h[2] >> 31 else '-'}")
            first_last += [int(hits[base + start][2]) & 0x7FFFFFFF, int(hits[base + start + count - 1][2]) & 0x7FFFFFFF]
            read_id, _, num, _ = lines[li].rstrip("\n").split("\t")
            rebuilt.append(f"{read_id}\t{ctgs[ctg]}\t{num}\t{' '.join(toks)}\n")
            li += 1
        assert int(read_len[r]) == max(first_last)       # pair:487
    assert rebuilt == lines


def test_parse_verbose_mappings_read_blocks_and_repeated_contig():
    lines = ["r1\tA\t2\t5:+_10:+ 9:+_20:+\n", "r1\tB\t1\t7:-_40:+\n", "r2\tA\t1\t1:+_3:-\n", "r1\tB\t1\t2:+_8:+\n",
             "r3\tA\t1\t5:+_10:+\n", "r3\tB\t1\t7:-_40:+\n", "r3\tA\t2\t6:+_50:+ 8:+_60:+\n"]
    hit_off, nruns, runs, hits, read_len = pair.parse_verbose_mappings(lines, {"A": 0, "B": 1})
    # a read is a block of consecutive lines with one id (pair:447-457): r1 comes back as a new read
    assert nruns.tolist() == [2, 1, 1, 3]
    assert read_len.tolist() == [40, 3, 8, 60]
    assert hit_off.tolist() == [0, 3, 4, 5, 9]
    # r3 lists contig A twice: both entries of contig_runs see the last listing (pair:470-472)
    base = int(hit_off[3])
    assert runs[base:base + 3].tolist() == [[0, 2, 2], [1, 1, 1], [0, 2, 2]]
    assert hits[base + 2].tolist() == [0, 6 | 0x80000000, 50 | 0x80000000]
