#!/usr/bin/env python3
"""Time of ntl_pairs_finish (tally + pair table download) as a function of the number of events in the log -- what rank 0
pays at N GPUs.   python tools/tally_probe.py"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import r1_inputs as bench
    from ntlink_b200 import Context
    contigs, reads = bench.make_inputs(0, 1)
    ctx = Context(0)
    ctx.target_upload(contigs)
    ctx.reads_upload(reads)
    prm = ctx.params(bench.K, bench.W, bench.Z)
    ctx.index_build_resident(bench.K, bench.W)
    for copies in (1, 2, 4, 8, 16):
        ctx.events_reset()
        for i in range(copies):
            ctx.map_resident(prm, first_ordinal=i * len(reads))
        n = ctx.events_count()
        ts = []
        for _ in range(6):
            t0 = time.perf_counter()
            raw, gaps = ctx.pairs_raw()
            ts.append(time.perf_counter() - t0)
        ctx.timing_reset()
        ctx.pairs_raw()
        print(json.dumps({"events": int(n), "pairs": len(raw), "pairs_raw_ms_best": round(1e3 * min(ts[1:]), 3),
                          "tally_kernels_ms": round(ctx.timing()["tally"], 3)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
