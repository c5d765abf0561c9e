#!/usr/bin/env python3
"""ntl_map_reads from PAGEABLE host memory (what the file reader hands over): driver staging vs the library's pinned
bounce buffers filled by N host threads.   python tools/pageable_probe.py"""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import r1_inputs as bench
    from ntlink_b200 import Context, SeqBatch
    contigs, reads = bench.make_inputs(0, 1)
    reps = 4
    seq = np.concatenate([reads.seq] * reps)
    offs = np.concatenate([[0]] + [reads.offsets[1:] + i * reads.offsets[-1] for i in range(reps)]).astype(np.uint64)
    big = SeqBatch(seq, offs, [f"r{i}" for i in range(len(offs) - 1)])
    ctx = Context(0)
    ctx.build_index_from_sequences(contigs, bench.K, bench.W, want_sketch=False)
    prm = ctx.params(bench.K, bench.W, bench.Z)
    for threads in (0, 1, 2, 4, 8, -1):
        ctx.set_option("copy_threads", threads)
        ts = []
        for _ in range(4):
            ctx.events_reset()
            t0 = time.perf_counter()
            ctx.map_reads(big, prm, 0)
            ts.append(time.perf_counter() - t0)
        best = min(ts[1:])
        print(json.dumps({"copy_threads": threads, "bases": int(offs[-1]), "best_s": round(best, 4), "gbp_per_s": round(int(offs[-1]) / best / 1e9, 2),
                          "all": [round(t, 3) for t in ts]}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
