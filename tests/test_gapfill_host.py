"""Host logic of ntlink_b200.gapfill: pairing of the masked records (bin/ntlink_patch_gaps.py:414-434). No GPU."""
import pytest

from ntlink_b200 import gapfill


def test_pair_up_follows_the_reference_headers():
    scaf = ["ctg1+_source", "ctg2-_target", "a_b+_source", "x__y-_target"]
    reads = ["read_9__ctg1+__ctg2-", "m54__r/1__a_b+__x__y-"]
    # greedy groups like the reference's regex (patch:414): everything up to the LAST two "__" belongs to the read name
    assert gapfill.pair_up(scaf[:2], reads[:1]) == [("read_9", "ctg1+", "ctg2-")]
    m = gapfill.READ_HEADER_RE.search(reads[1]).groups()
    assert m == ("m54__r/1__a_b+", "x", "y-")
    for bad_scaf, bad_reads in ((["ctg1+_target", "ctg2-_source"], reads[:1]), (scaf[:1], reads[:1]), (scaf[:2], ["noseparators"]),
                                (["ctg9+_source", "ctg2-_target"], reads[:1])):
        with pytest.raises(ValueError):
            gapfill.pair_up(bad_scaf, bad_reads)
    assert gapfill.pair_up([], []) == []
