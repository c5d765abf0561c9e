#!/usr/bin/env python3
"""Wall time of the parts of bench.py's end-to-end step (pinned host inputs): python tools/e2e_parts.py"""
import ctypes as C
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import r1_inputs as bench
    from ntlink_b200 import Context, _lib
    contigs, reads = bench.make_inputs(0, 1)
    pc, pr = bench.pinned_copy(contigs), bench.pinned_copy(reads)
    ctx = Context(0)
    if os.environ.get("NTL_PIPE_MIN"):
        ctx.set_option("pipeline_min_bases", float(os.environ["NTL_PIPE_MIN"]))
    prm = ctx.params(bench.K, bench.W, bench.Z)
    acc = {"reset": 0.0, "index": 0.0, "map_reads": 0.0, "pairs": 0.0}
    n = 10
    for it in range(n + 3):
        t = [time.perf_counter()]
        ctx.events_reset(); t.append(time.perf_counter())
        ctx.build_index_from_sequences(pc, bench.K, bench.W, want_sketch=False); t.append(time.perf_counter())
        mo = _lib.MapOut()
        ctx._check(ctx.lib.ntl_map_reads(ctx.h, pr.seq.ctypes.data, pr.offsets.ctypes.data, len(pr), 0, C.byref(prm), C.byref(mo)), "map")
        t.append(time.perf_counter())
        ctx.pairs_raw(); t.append(time.perf_counter())
        if it >= 3:
            for key, a, b in zip(acc, t, t[1:]):
                acc[key] += (b - a) * 1e3 / n
    print(json.dumps({k: round(v, 3) for k, v in acc.items()}), "total", round(sum(acc.values()), 3))
    ctx.close()


if __name__ == "__main__":
    main()
