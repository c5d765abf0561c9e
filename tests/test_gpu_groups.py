"""Grouped mapping for gap filling (ntl_map_groups): many (a few target sequences + one read) problems in one call, against
the oracle's restatement of read_btllib_minimizers + get_accepted_anchor_contigs (bin/ntlink_patch_gaps.py:397-442,
bin/ntlink_utils.py:200-294)."""
import os
import sys

import numpy as np
import pytest

import util

sys.path.insert(0, os.path.join(util.REPO, "oracle"))
import pair_oracle as po  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from ntlink_b200 import Context
    c = Context(0)
    yield c
    c.close()


def oracle_group(th, tp, to, tnames, tlen, targets, rh, rp, r0, r1, read_len, prm):
    "index of the group's targets (hash seen twice -> dropped), membership filter, accepted anchor contigs"
    index, multi = {}, set()
    for t in targets:
        for i in range(int(to[t]), int(to[t + 1])):
            mx = int(th[i])
            if mx in index:
                multi.add(mx)
            else:
                index[mx] = (t, int(tp[i] & 0x7FFFFFFF), "+" if tp[i] >> 31 else "-")
    for mx in multi:
        del index[mx]
    hits = [(int(rh[i]), int(rp[i] & 0x7FFFFFFF), "+" if rp[i] >> 31 else "-") for i in range(r0, r1) if int(rh[i]) in index]
    if not hits:
        return []
    lengths = {t: int(tlen[t]) for t in targets}
    acc = po.accepted_anchor_contigs(hits, read_len, lengths, index, prm)
    return [(ctg, [(h.ctg_pos, h.ctg_strand, h.read_pos, h.read_strand) for h in hs]) for ctg, hs in acc]


@pytest.mark.parametrize("fixture,reads,k,w,sens,x,z", [("scaffolds_3.fa", "long_reads_3.fa", 20, 10, False, 0.0, 1000),
                                                        ("scaffolds_2.fa", "long_reads_2.fq", 20, 10, True, 0.0, 500),
                                                        ("scaffolds_3.fa", "long_reads_3.fa", 24, 50, False, 1.2, 1000)])
def test_grouped_mapping_against_oracle(ctx, tmp_path, fixture, reads, k, w, sens, x, z):
    from ntlink_b200 import Sketch
    tnames, tseq, toff = util.load_fasta_batch(util.fixture_file(tmp_path, fixture))
    rnames, rseq, roff = util.load_fasta_batch(util.fixture_file(tmp_path, reads))
    nreads = min(len(rnames), 160)
    roff = roff[:nreads + 1]
    rseq = rseq[:int(roff[-1])]
    th, tp, ts, tmo = util.oracle_sketch_batch(tseq, toff, k, w)
    rh, rp, rs, rmo = util.oracle_sketch_batch(rseq, roff, k, w)
    tpf = (tp | (ts.astype(np.uint32) << 31)).astype(np.uint32)
    rpf = (rp | (rs.astype(np.uint32) << 31)).astype(np.uint32)
    tlen = np.diff(toff).astype(np.uint32)
    rlen = np.diff(roff).astype(np.uint32)
    nt = len(tnames)
    # target sequences are laid out group by group: group g = two contigs (like the two ends of a gap); a few groups get
    # one, three or ALL contigs (the last kind does not fit the shared-memory table)
    rng = np.random.default_rng(k + w)
    order, g_off = [], [0]
    for g in range(nreads):
        kind = g % 17
        members = list(range(nt)) if kind == 5 else [int(v) for v in rng.choice(nt, size=1 if kind == 3 else 3 if kind == 7 else 2, replace=False)]
        order += members
        g_off.append(len(order))
    # materialise the permuted / repeated target sketch
    seg_h = [th[int(tmo[t]):int(tmo[t + 1])] for t in order]
    seg_p = [tpf[int(tmo[t]):int(tmo[t + 1])] for t in order]
    gh = np.concatenate(seg_h) if seg_h else np.empty(0, np.uint64)
    gp = np.concatenate(seg_p) if seg_p else np.empty(0, np.uint32)
    go = np.concatenate([[0], np.cumsum([len(s) for s in seg_h])]).astype(np.uint64)
    glen = tlen[order]
    prm_gpu = ctx.params(k, w, z, 10, x, sens, False)
    res = ctx.map_groups(Sketch(gh, gp, go), glen, np.array(g_off, np.uint32), Sketch(rh, rpf, rmo), rlen, prm_gpu)
    prm = po.Params(k, z, 1, 10, x, 1, sens, False)
    assert res.n_reads == nreads
    n_big = 0
    for g in range(nreads):
        targets = list(range(g_off[g], g_off[g + 1]))
        n_big += int(go[g_off[g + 1]] - go[g_off[g]]) > 4096
        want = oracle_group(gh, gp, go, None, glen, targets, rh, rpf, int(rmo[g]), int(rmo[g + 1]), int(rlen[g]), prm)
        base = int(res.hit_off[g])
        got = []
        for ctg, start, count in res.runs[base:base + int(res.nruns[g])]:
            hs = res.hits[base + int(start):base + int(start) + int(count)]
            assert all(int(h[0]) == int(ctg) for h in hs)
            got.append((int(ctg), [(int(h[1]) & 0x7FFFFFFF, "+" if int(h[1]) >> 31 else "-", int(h[2]) & 0x7FFFFFFF, "+" if int(h[2]) >> 31 else "-") for h in hs]))
        assert got == want, g
    assert n_big >= 3 and res.n_runs > 0


def test_grouped_mapping_edge_cases(ctx):
    from ntlink_b200 import Sketch
    prm = ctx.params(20, 10, 0, 10, 0.0)
    empty = Sketch(np.empty(0, np.uint64), np.empty(0, np.uint32), np.zeros(1, np.uint64))
    res = ctx.map_groups(empty, np.empty(0, np.uint32), np.zeros(1, np.uint32), empty, np.empty(0, np.uint32), prm)
    assert res.n_reads == 0
    # one group, one target with a duplicated hash, the all-ones hash (the table's EMPTY sentinel) and a normal one
    E = np.uint64(0xFFFFFFFFFFFFFFFF)
    t = Sketch(np.array([5, 7, 5, E, 9], np.uint64), np.array([10, 20, 30, 40, 50], np.uint32) | np.uint32(1 << 31), np.array([0, 5], np.uint64))
    r = Sketch(np.array([9, 5, E, 7, 11], np.uint64), np.array([1, 2, 3, 4, 5], np.uint32), np.array([0, 5], np.uint64))
    res = ctx.map_groups(t, np.array([1000], np.uint32), np.array([0, 1], np.uint32), r, np.array([100], np.uint32), prm)
    assert int(res.nruns[0]) == 1
    hits = res.hits[:int(res.runs[0][2])]
    assert [(int(h[1]) & 0x7FFFFFFF, int(h[2]) & 0x7FFFFFFF) for h in hits] == [(50, 1), (40, 3), (20, 4)]      # 5 is duplicated, 11 unknown
    with pytest.raises(ValueError):
        ctx.map_groups(t, np.array([1000], np.uint32), np.array([0, 2], np.uint32), r, np.array([100], np.uint32), prm)


def test_gapfill_front_end_on_masked_files(ctx, tmp_path):
    """ntlink_b200.gapfill.map_long_reads on files written the way print_masked_sequences does (patch:346-389): N-masked
    full-length scaffolds and reads, `<read>__<src>__<tgt>` / `<src>_source` / `<tgt>_target` headers; checked per gap
    against the oracle (index of the two ends with duplicates dropped + accepted_anchor_contigs)."""
    from ntlink_b200 import gapfill
    k, w = 20, 10
    tnames, tseq, toff = util.load_fasta_batch(util.fixture_file(tmp_path, "scaffolds_3.fa"))
    rnames, rseq, roff = util.load_fasta_batch(util.fixture_file(tmp_path, "long_reads_3.fa"))
    rng = np.random.default_rng(99)
    lengths = {n: int(toff[i + 1] - toff[i]) for i, n in enumerate(tnames)}
    big = [i for i, n in enumerate(tnames) if lengths[n] > 4000]
    # gaps where the read really joins the two contigs: taken from the reference-generated mapping of this fixture
    tidx, ridx, joined, seen = {n: i for i, n in enumerate(tnames)}, {n: i for i, n in enumerate(rnames)}, [], {}
    for line in util.golden_case("f3_default", "verbose_mapping.tsv").decode().splitlines():
        rd, ctg = line.split("\t")[:2]
        seen.setdefault(rd, []).append(ctg)
    for rd, ctgs in seen.items():
        if len(ctgs) >= 2 and all(lengths[c] > 4000 for c in ctgs[:2]):
            joined.append((ridx[rd], tidx[ctgs[0]], tidx[ctgs[1]]))
    assert len(joined) >= 5
    spath, rpath = os.path.join(str(tmp_path), "o.scaffolds.masked_temp.fa"), os.path.join(str(tmp_path), "o.reads.masked_temp.fa")
    gaps = []
    with open(spath, "w") as fs, open(rpath, "w") as fr:
        for g in range(40):
            if g < len(joined):
                r, a, b = joined[g]
            else:
                a, b = (int(v) for v in rng.choice(big, size=2, replace=False))
                r = int(rng.integers(0, len(rnames)))
            src, tgt = tnames[a] + "+-"[g % 2], tnames[b] + "-+"[g % 3 == 0]
            sa, sb = tseq[int(toff[a]):int(toff[a + 1])].tobytes().decode(), tseq[int(toff[b]):int(toff[b + 1])].tobytes().decode()
            cut_a, cut_b = int(rng.integers(500, len(sa) // 3)), int(rng.integers(500, len(sb) // 3))
            fs.write(f">{src}_source\n{'N' * cut_a}{sa[cut_a:]}\n" if src[-1] == "+" else f">{src}_source\n{sa[:cut_a]}{'N' * (len(sa) - cut_a)}\n")
            fs.write(f">{tgt}_target\n{sb[:cut_b]}{'N' * (len(sb) - cut_b)}\n" if tgt[-1] == "+" else f">{tgt}_target\n{'N' * cut_b}{sb[cut_b:]}\n")
            rs = rseq[int(roff[r]):int(roff[r + 1])].tobytes().decode()
            lo, hi = int(rng.integers(0, len(rs) // 8 + 1)), len(rs) - int(rng.integers(0, len(rs) // 8 + 1))
            fr.write(f">{rnames[r]}__{src}__{tgt}\n{'N' * lo}{rs[lo:hi]}{'N' * (len(rs) - hi)}\n")
            gaps.append((rnames[r], src, tgt))
    got = gapfill.map_long_reads(ctx, spath, rpath, lengths, k=k, w=w, z=1000)
    assert [(m.read, m.source, m.target) for m in got] == gaps
    # oracle on the same files
    _, sseq, soff = util.load_fasta_batch(spath)
    _, mseq, moff = util.load_fasta_batch(rpath)
    th, tp, ts, tmo = util.oracle_sketch_batch(sseq, soff, k, w)
    rh, rp, rs_, rmo = util.oracle_sketch_batch(mseq, moff, k, w)
    tpf = (tp | (ts.astype(np.uint32) << 31)).astype(np.uint32)
    rpf = (rp | (rs_.astype(np.uint32) << 31)).astype(np.uint32)
    prm = po.Params(k, 1000, 1, 10, 0.0, 1, False, False)
    n_with = 0
    for g, m in enumerate(got):
        ids = [m.source.strip("+-"), m.target.strip("+-")]
        tlen = [lengths[ids[0]], lengths[ids[1]]]
        want = oracle_group(th, tpf, tmo, None, {2 * g: tlen[0], 2 * g + 1: tlen[1]}, [2 * g, 2 * g + 1], rh, rpf, int(rmo[g]), int(rmo[g + 1]),
                            int(moff[g + 1] - moff[g]), prm)
        assert [(r.contig, [(h.ctg_pos, h.ctg_strand, h.read_pos, h.read_strand) for h in r.hits]) for r in m.accepted] == \
            [(ids[c - 2 * g], hs) for c, hs in want], g
        n_with += bool(m.accepted)
    assert n_with >= 3
    with pytest.raises(ValueError):
        gapfill.pair_up(["a+_source"], ["r__a+__b-"])
