#!/usr/bin/env python3
"""Large-scale parity + throughput check (not part of the default test-suite: takes minutes).

    python tests/scale/scale_check.py --genome 100e6 --coverage 10 -k 24 -w 250 --sensitive

Generates a synthetic assembly + ONT-like reads (ntlink_b200/synth.py), runs the whole GPU path through the C ABI,
and compares against the CPU oracle: sketch checksums of ALL reads (multi-threaded C oracle) and byte-equality of
verbose_mapping / PAF / pairs.tsv / scaffold.dot on the first --oracle-reads reads mapped by the Python oracle."""
import argparse
import io
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=float, default=100e6)
    ap.add_argument("--coverage", type=float, default=10)
    ap.add_argument("-k", type=int, default=24)
    ap.add_argument("-w", type=int, default=250)
    ap.add_argument("-z", type=int, default=1000)
    ap.add_argument("--sensitive", action="store_true")
    ap.add_argument("--oracle-reads", type=int, default=20000)
    ap.add_argument("--n-frac", type=float, default=0.02, help="fraction of contigs with an internal N run")
    a = ap.parse_args()
    from ntlink_b200 import Context, SeqBatch, pair, synth
    import util
    import pair_oracle as po

    t0 = time.time()
    gen = synth.genome(int(a.genome), 777)
    contigs = synth.assembly(gen, 778, n_frac=a.n_frac)
    reads = synth.reads(gen, a.coverage, 779)
    print(f"[gen] {len(contigs)} contigs {int(contigs.offsets[-1])} bp, {len(reads)} reads {int(reads.offsets[-1])} bp, {time.time()-t0:.1f}s", flush=True)

    ctx = Context(0)
    t0 = time.time()
    tsk = ctx.build_index_from_sequences(contigs, a.k, a.w, want_sketch=True)
    prm = ctx.params(a.k, a.w, a.z, 10, 0.0, a.sensitive, False)
    ctx.events_reset()
    ctx.timing_reset()
    t1 = time.time()
    res = ctx.map_reads(reads, prm, 0)
    t2 = time.time()
    gp = ctx.pairs()
    t3 = time.time()
    tm = ctx.timing()
    rb = int(reads.offsets[-1])
    print(f"[gpu] index {t1-t0:.2f}s  map_reads {t2-t1:.2f}s ({rb/(t2-t1)/1e9:.1f} Gbp/s incl. H2D/D2H from pageable numpy)  tally {t3-t2:.3f}s", flush=True)
    print(f"[gpu] device total {tm['total']:.1f} ms -> {rb/tm['total']/1e6:.1f} Gbp/s on-device; dense {tm['big_dense_ms']:.1f} ms over {tm['big_dense_launches']} launches", flush=True)
    print(f"[gpu] minimizers target {len(tsk)} reads {res.n_mx} hits {res.n_hits} runs {res.n_runs} events {res.n_events} pairs {len(gp)} index {ctx.index_stats()}", flush=True)

    # ---- sketch parity on everything (C oracle, all threads)
    t0 = time.time()
    oh, op, os_, oo = util.oracle_sketch_batch(contigs.seq, contigs.offsets, a.k, a.w, threads=os.cpu_count())
    assert np.array_equal(tsk.hash, oh) and np.array_equal(tsk.pos, op) and np.array_equal(tsk.strand, os_) and np.array_equal(tsk.seq_off, oo)
    rsk = ctx.sketch(reads, a.k, a.w)
    rh, rp, rs, ro = util.oracle_sketch_batch(reads.seq, reads.offsets, a.k, a.w, threads=os.cpu_count())
    assert np.array_equal(rsk.seq_off, ro), "read sketch offsets differ"
    assert np.array_equal(rsk.hash, rh) and np.array_equal(rsk.pos, rp) and np.array_equal(rsk.strand, rs), "read sketch differs"
    print(f"[parity] sketches of target ({len(oh)}) and ALL reads ({len(rh)} minimizers) identical to the oracle, {time.time()-t0:.1f}s", flush=True)

    # ---- mapping parity on the first N reads (Python oracle)
    n = min(a.oracle_reads, len(reads))
    sub = SeqBatch(reads.seq[:int(reads.offsets[n])], reads.offsets[:n + 1], reads.names[:n])
    ctx.events_reset()
    sres = ctx.map_reads(sub, prm, 0)
    lengths = {nm: int(l) for nm, l in zip(contigs.names, contigs.lengths)}
    gpairs = pair.filter_weak_anchor_pairs(pair.filter_pairs_distances(pair.pairs_dict(ctx.pairs(), contigs.names), lengths), 1)
    got = (sres.verbose_bytes(sub, contigs, threads=8), sres.paf_bytes(sub, sub.lengths.astype(np.uint32), contigs, a.k, threads=8),
           pair.pairs_tsv(gpairs).encode(), pair.scaffold_dot(gpairs, lengths, 1).encode())
    t0 = time.time()

    def tsv(names, h, p, s, off, lens=None):
        out = []
        for i, nm in enumerate(names):
            x, y = int(off[i]), int(off[i + 1])
            toks = " ".join(f"{a_}:{b_}:{'+' if c_ else '-'}" for a_, b_, c_ in zip(h[x:y].tolist(), p[x:y].tolist(), s[x:y].tolist()))
            out.append(nm + (f"\t{lens[i]}" if lens is not None else "") + "\t" + toks + "\n")
        return out
    index = po.read_target_index(tsv(contigs.names, oh, op, os_, oo))
    oprm = po.Params(a.k, a.z, 1, 10, 0.0, 1, a.sensitive, False)
    vb, pf = io.StringIO(), io.StringIO()
    opairs = po.filter_pairs(po.map_reads(tsv(sub.names, rh, rp, rs, ro[:n + 1], sub.lengths), index, lengths, oprm, vb, pf), lengths, 1)
    assert got[0] == vb.getvalue().encode(), "verbose_mapping differs"
    assert got[1] == pf.getvalue().encode(), "paf differs"
    assert got[2] == "".join(po.pairs_tsv_lines(opairs)).encode(), "pairs.tsv differs"
    assert util.dot_parts(got[3]) == util.dot_parts("".join(po.dot_lines(opairs, lengths, 1)).encode()), "dot differs"
    print(f"[parity] mapping of the first {n} reads: verbose ({len(got[0])} B), paf ({len(got[1])} B), pairs ({len(opairs)}), dot identical "
          f"to the oracle, oracle {time.time()-t0:.1f}s", flush=True)
    print(json.dumps({"scale_check": "ok", "genome_bp": int(a.genome), "read_bp": rb, "k": a.k, "w": a.w, "sensitive": a.sensitive,
                      "device_ms_total": tm["total"], "gbp_per_s_device": rb / tm["total"] / 1e6, "pairs": len(gp)}))


if __name__ == "__main__":
    main()
