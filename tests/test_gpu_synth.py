"""The synthetic-input generator (device == host bytes) and the TIMED entry points of bench.py (resident target, resident
reads in several chunks, sync-free multi-chunk ntl_map_reads) on simulated ONT reads against the CPU pipeline: the
reference's own make recipe with the C restatement of indexlr and the unmodified bin/ntlink_pair.py when oracle/_ref is
staged (the port oracle/pair_oracle.py otherwise)."""
import os
import sys

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
sys.path.insert(0, util.ORACLE_DIR)


@pytest.fixture(scope="module")
def ctx():
    from ntlink_b200 import Context
    c = Context(0)
    yield c
    c.close()


def test_device_generator_equals_host_generator(ctx):
    from ntlink_b200 import synth
    cplan, names = synth.plan_assembly(3_000_000, 11, n_frac=0.3)
    rplan = synth.plan_reads(3_000_000, 30_000_000, 12, first_id=777)
    hc, hr = synth.host_contigs(5, cplan, names), synth.host_reads(5, rplan)
    ctx.set_option("resident_chunk_bases", 7_000_000)            # reads end up in several device chunks
    try:
        ctx.synth_target_resident(5, cplan, names)
        nb = ctx.synth_reads_resident(5, rplan)
        assert nb == len(hr.seq) and ctx.resident_info(1) == (len(rplan), nb)
        dc = ctx.resident_download(0, 0, len(cplan), names)
        dr = ctx.resident_download(1, 0, len(rplan), hr.names)
        assert np.array_equal(dc.offsets, hc.offsets) and np.array_equal(dc.seq, hc.seq)
        assert np.array_equal(dr.offsets, hr.offsets) and np.array_equal(dr.seq, hr.seq)
        a, n = len(rplan) // 3, len(rplan) // 2                   # a range that crosses chunk boundaries
        part = ctx.resident_download(1, a, n, hr.names[a:a + n])
        assert np.array_equal(part.seq, hr.seq[int(hr.offsets[a]):int(hr.offsets[a + n])])
        assert (hc.seq == ord("N")).sum() > 0
        comp = np.bincount(hr.seq, minlength=256)[[65, 67, 71, 84]] / len(hr.seq)
        assert np.all(np.abs(comp - 0.25) < 0.01)
        ratio = len(hr.seq) / rplan["len"].sum()
        assert 0.99 < ratio < 1.01                                # 3 % deletions, 3 % insertions
    finally:
        ctx.set_option("resident_chunk_bases", 512 << 20)


@pytest.mark.parametrize("k,w,sens,genome,n_frac", [(24, 250, True, 6_000_000, 0.2), (32, 100, False, 3_000_000, 0.0), (32, 250, False, 5_000_000, 0.5)])
def test_timed_entry_points_against_cpu_pipeline(ctx, tmp_path, k, w, sens, genome, n_frac):
    import cpu_pipeline as cp
    from ntlink_b200 import Context, pair, synth
    cplan, names = synth.plan_assembly(genome, 21, n_frac=n_frac)
    rplan = synth.plan_reads(genome, 8 * genome, 22)
    contigs, reads = synth.host_contigs(9, cplan, names), synth.host_reads(9, rplan)
    tf, rf = str(tmp_path / "target.fa"), str(tmp_path / "reads.fa")
    cp.write_fasta(tf, contigs)
    cp.write_fasta(rf, reads)
    tsv, _ = cp.sketch_target(tf, k, w, 4)
    cp.map_reads(tf, tsv, rf, str(tmp_path / "cpu"), k, w, 1000, 4, sensitive=sens, verbose=True, pairs=True, paf=True)
    want = cp.outputs(str(tmp_path / "cpu"))
    assert len(want["pairs"]) > 100 and len(want["verbose"]) > 10000
    lengths = {n: int(l) for n, l in zip(contigs.names, contigs.lengths)}
    prm = ctx.params(k, w, 1000, 10, 0.0, sens, False)

    def files(c):
        prs = pair.filter_weak_anchor_pairs(pair.filter_pairs_distances(pair.pairs_dict(c.pairs(), contigs.names), lengths), 1)
        return pair.pairs_tsv(prs).encode(), util.dot_parts(pair.scaffold_dot(prs, lengths, 1).encode())

    # sync-free ntl_map_reads in several chunks: every file
    ctx.set_option("pipeline_min_bases", 16e6)
    try:
        ctx.events_reset()
        tsk = ctx.build_index_from_sequences(contigs, k, w, want_sketch=True)
        fb = ctx.stat("async_fallbacks")
        res = ctx.map_reads(reads, prm, 0)
        assert ctx.stat("async_fallbacks") == fb and ctx.stat("graph_launches") >= 2
    finally:
        ctx.set_option("pipeline_min_bases", 80 << 20)
    assert tsk.to_tsv(contigs) == open(tsv, "rb").read()
    assert res.verbose_bytes(reads, contigs) == want["verbose"]
    assert res.paf_bytes(reads, reads.lengths.astype(np.uint32), contigs, k) == want["paf"]
    ptsv, dot = files(ctx)
    assert ptsv == want["pairs"] and dot == util.dot_parts(want["dot"])
    # the resident entry points (what bench.py times), inputs generated on the device, several chunks
    c2 = Context(0)
    try:
        c2.set_option("resident_chunk_bases", int(reads.offsets[-1]) // 5)
        c2.synth_target_resident(9, cplan, names)
        c2.synth_reads_resident(9, rplan)
        c2.events_reset()
        c2.index_build_resident(k, w)
        st = c2.map_resident(prm, 0)
        assert (st["mx"], st["hits"], st["runs"], st["events"]) == (res.n_mx, res.n_hits, res.n_runs, res.n_events)
        assert c2.stat("async_fallbacks") == 0
        ptsv2, dot2 = files(c2)
        assert ptsv2 == want["pairs"] and dot2 == util.dot_parts(want["dot"])
    finally:
        c2.close()
