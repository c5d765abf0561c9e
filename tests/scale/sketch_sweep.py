#!/usr/bin/env python3
"""Sketch-only kernel time over (k, w) settings the reference uses (ntLink defaults, the overlap stage's k15/w5 and the
gap-fill stage's k20/w10), with a checksum comparison against the C oracle on a sample of the sequences.

    python tests/scale/sketch_sweep.py [--bases 100e6]"""
import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bases", type=float, default=100e6)
    a = ap.parse_args()
    from ntlink_b200 import Context, SeqBatch
    import util
    rng = np.random.default_rng(5)
    n = int(a.bases)
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
    lens = []
    left = n
    while left > 0:
        L = min(left, int(rng.integers(2000, 60000)))
        lens.append(L)
        left -= L
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    batch = SeqBatch(seq, offs, [f"s{i}" for i in range(len(lens))])
    ns = min(len(lens), 200)
    ctx = Context(0)
    if os.environ.get("NTL_STRIP_LEN"):
        ctx.set_option("strip_len", float(os.environ["NTL_STRIP_LEN"]))
    configs = [(32, 100), (24, 250), (40, 100), (20, 10), (15, 5)]
    if os.environ.get("SWEEP_KW"):
        configs = [tuple(int(v) for v in kw.split(",")) for kw in os.environ["SWEEP_KW"].split()]
    for k, w in configs:
        for _ in range(2):
            ctx.sketch(batch, k, w)
        ctx.timing_reset()
        reps = 3
        for _ in range(reps):
            sk = ctx.sketch(batch, k, w)
        tm = ctx.timing()
        stages = {s: round(tm[s] / reps, 3) for s in ("pack", "dense", "select", "gap", "emit")}
        total = sum(stages.values())
        oh, op, os_, omo = util.oracle_sketch_batch(seq[:int(offs[ns])], offs[:ns + 1], k, w)
        m = int(sk.seq_off[ns])
        ok = (int(omo[ns]) == m and np.array_equal(sk.hash[:m], oh) and
              np.array_equal(sk.pos_strand[:m], (op | (os_.astype(np.uint32) << 31)).astype(np.uint32)))
        print(json.dumps({"k": k, "w": w, "minimizers": int(len(sk.hash)), "kernel_ms": round(total, 3),
                          "gbp_per_s": round(n / total / 1e6, 1), "stages_ms": stages, "oracle_match_first_%d" % ns: bool(ok)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
