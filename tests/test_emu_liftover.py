"""The liftover kernel logic (ntlink_b200/csrc/lift_logic.cuh, compiled for the host by tests/emu) + the host side of
ntlink_b200/liftover.py against the goldens made by the unmodified bin/ntlink_liftover_mappings.py."""
import numpy as np
import pytest

import util
from ntlink_b200 import liftover

MAN = util.lift_manifest()


def emu_liftover(mapping_lines, agp_lines, k):
    old_names, (hit_off, nruns, runs, hits), ids = liftover.load_mappings(mapping_lines)
    rows, new_names = liftover.agp_table(old_names, liftover.read_agp(agp_lines))
    lib = util.emu_lib()
    nruns_out = np.zeros_like(nruns)
    runs_out = np.zeros_like(runs)
    hits_out = np.zeros_like(hits)
    err = lib.emu_liftover(hit_off.ctypes.data, nruns.ctypes.data, runs.ctypes.data, hits.ctypes.data, len(nruns),
                           rows.ctypes.data, len(rows), k, nruns_out.ctypes.data, runs_out.ctypes.data, hits_out.ctypes.data)
    return err, util.verbose_text(hit_off, nruns_out, runs_out, hits_out, ids, new_names)


@pytest.mark.parametrize("case", sorted(MAN))
def test_liftover_logic_against_reference(case):
    err, text = emu_liftover(util.lift_golden(case, "in.verbose_mapping.tsv").decode().splitlines(True),
                             util.lift_golden(case, "in.agp").decode().splitlines(True), MAN[case]["k"])
    assert err == 0
    assert text == util.lift_golden(case, "lifted.verbose_mapping.tsv").decode()


def test_subsumption_is_by_name_from_the_first_run():
    "A B A C A: the third A marks B, A and C (liftover:99-102 counts from the FIRST run of A), so nothing is left"
    agp = ["pA\t1\t1000\t1\tW\ta\t1\t1000\t+\n", "pB\t1\t1000\t1\tW\tb\t1\t1000\t+\n", "pC\t1\t1000\t1\tW\tc\t1\t1000\t+\n",
           "pD\t1\t1000\t1\tW\td\t1\t1000\t+\n"]
    rows = ["r\ta\t1\t10:+_1:+\n", "r\tb\t1\t20:+_2:+\n", "r\ta\t1\t30:+_3:+\n", "r\tc\t1\t40:+_4:+\n", "r\ta\t1\t50:+_5:+\n",
            "r\td\t1\t60:+_6:+\n"]
    err, text = emu_liftover(rows, agp, 10)
    assert err == 0 and text == "r\tpD\t1\t60:+_6:+\n"
    # A B A: B is dropped and the two A runs merge into one increasing run
    err, text = emu_liftover(rows[:3], agp, 10)
    assert err == 0 and text == "r\tpA\t2\t10:+_1:+ 30:+_3:+\n"
    import sys, os
    sys.path.insert(0, os.path.join(util.REPO, "oracle"))
    import liftover_oracle
    assert "".join(liftover_oracle.liftover(rows, agp, 10)) == "r\tpD\t1\t60:+_6:+\n"


def test_malformed_layout_is_reported():
    hit_off = np.array([0, 2], np.uint32)
    nruns = np.array([2], np.uint32)
    runs = np.array([[0, 1, 1], [0, 0, 1]], np.uint32)           # second run starts before the first ends
    hits = np.zeros((2, 3), np.uint32)
    rows = np.array([[0, 1, 1, 1, 100]], np.uint32)
    out = [np.zeros_like(nruns), np.zeros_like(runs), np.zeros_like(hits)]
    err = util.emu_lib().emu_liftover(hit_off.ctypes.data, nruns.ctypes.data, runs.ctypes.data, hits.ctypes.data, 1,
                                      rows.ctypes.data, 1, 10, *(o.ctypes.data for o in out))
    assert err == 2 and out[0][0] == 0
