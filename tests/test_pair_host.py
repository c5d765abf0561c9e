"""Host-side helpers of ntlink_b200/pair.py that do not need a GPU."""
import numpy as np

from ntlink_b200 import pair


def test_gap_estimate_is_int_of_numpy_median():
    "PairInfo.get_gap_estimate (bin/ntlink_pair.py:70-74) is int(np.median(gaps)); pair.gap_estimate computes it without numpy"
    rng = np.random.default_rng(3)
    for _ in range(5000):
        n = int(rng.integers(1, 14))
        gaps = rng.integers(-4000, 8000, n).tolist()
        if rng.random() < 0.2:
            gaps = [gaps[0]] * n
        assert pair.gap_estimate(gaps) == int(np.median(gaps)), gaps
    for gaps in ([-1, 0], [-3, -2], [-1, 2], [5], [2**31 - 1, 2**31 - 1], [-2**31, -2**31 + 1], [0, -1, -1, 0], [7, 7, 8, 8]):
        assert pair.gap_estimate(gaps) == int(np.median(gaps)), gaps
    assert pair.gap_estimate(np.array([4, -9, 3], np.int32)) == 3


def test_pairs_dict_takes_numpy_gap_arrays_and_lists():
    names = ["a", "b", "c"]
    raw = [(0, 1, 3, 2, 1, np.array([5, -2], np.int32)), (2, 0, 0, 1, 0, [7])]
    d = pair.pairs_dict(raw, names)
    assert d == {("a", "+", "b", "+"): ([5, -2], 1), ("c", "-", "a", "-"): ([7], 0)}
    assert all(isinstance(g, int) for v in d.values() for g in v[0])
