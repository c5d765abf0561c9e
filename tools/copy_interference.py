#!/usr/bin/env python3
"""Does a host->device copy in flight slow the sketch/mapping kernels down? Runs the resident arm of bench.py's workload
with and without a saturating pinned H2D copy on another stream and prints the per-stage kernel times.

    python tools/copy_interference.py"""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    from ntlink_b200 import Context
    import r1_inputs as bench
    contigs, reads = bench.make_inputs(0, 1)
    ctx = Context(0)
    ctx.target_upload(contigs)
    ctx.reads_upload(reads)
    prm = ctx.params(32, 100, 1000)
    h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    side = torch.cuda.Stream()

    def run(label, with_copy):
        for _ in range(3):
            ctx.events_reset(); ctx.index_build_resident(32, 100); ctx.map_resident(prm)
        torch.cuda.synchronize()
        if with_copy:
            with torch.cuda.stream(side):
                for _ in range(40):                 # ~0.2 s of copy queued ahead
                    d.copy_(h, non_blocking=True)
        ctx.timing_reset()
        ctx.mark(0)
        n = 10
        for _ in range(n):
            ctx.events_reset(); ctx.index_build_resident(32, 100); ctx.map_resident(prm); ctx.pairs_raw()
        ctx.mark(1)
        ctx.sync()
        ms = ctx.mark_elapsed_ms() / n
        tm = ctx.timing()
        torch.cuda.synchronize()
        print(json.dumps({"case": label, "ms_per_step": round(ms, 3),
                          "stages": {k: round(tm[k] / n, 3) for k in ("pack", "dense", "select", "gap", "emit", "lookup", "chain", "tally", "index")}}))

    for mode in (1, 0):
        ctx.set_option("async", mode)
        run(f"async={mode} no copy", False)
        run(f"async={mode} H2D in flight", True)
    # host-side cost of enqueueing: time the enqueue of one step without waiting for it
    import time
    ctx.set_option("async", 1)
    for with_copy in (False, True):
        torch.cuda.synchronize()
        if with_copy:
            with torch.cuda.stream(side):
                for _ in range(40):
                    d.copy_(h, non_blocking=True)
        t0 = time.perf_counter()
        for _ in range(10):
            ctx.events_reset(); ctx.index_build_resident(32, 100); ctx.map_resident(prm)
        dt = (time.perf_counter() - t0) / 10
        torch.cuda.synchronize()
        print(json.dumps({"case": "wall per step, copy=%s" % with_copy, "ms": round(dt * 1e3, 3)}))
    ctx.close()


if __name__ == "__main__":
    main()
