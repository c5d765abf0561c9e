#!/usr/bin/env python3
"""
pair_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle / CPU baseline for the mapping stage).

A plain-Python restatement of the minimizer-mapping stage of bcgsc/ntLink v1.3.11, written from the
behaviour of the reference (not a copy of it); every function cites the reference lines it follows:

  M1  index build with multi-copy removal      bin/ntlink_pair.py:189-211
  M2  per-read lookup (+ --repeat-filter)      bin/ntlink_pair.py:352-376
  M3  anchor chaining                          bin/ntlink_utils.py:200-294
  M4  first / terminal minimizer               bin/ntlink_pair.py:395-406
  M5  pair tallying                            bin/ntlink_pair.py:416-435
  M6  pair info / gap estimate                 bin/ntlink_pair.py:157-187,213-239,315-334
  M7  median / filters                         bin/ntlink_pair.py:70-74,241-255
  M8  verbose_mapping.tsv lines                bin/ntlink_pair.py:307-313,382-388
  M9  PAF-like lines                           bin/ntlink_paf_output.py:9-135
  M10 pairs.tsv / scaffold.dot                 bin/ntlink_pair.py:118-155,263-305,490-506

Parity is PINNED: tests/golden/make_golden.py runs the unmodified reference (imported from
/root/reference/bin with a tiny igraph stand-in) on the reference fixtures for many option
combinations and commits its outputs under tests/golden/; tests/test_oracle_pair.py checks this
restatement against every one of them and against the reference's own tests/expected_outputs.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
module. The product path (ntlink_b200/) never does.
"""
import argparse
import datetime
import os
import re
import sys
from collections import namedtuple

Hit = namedtuple("Hit", ["mx", "ctg_pos", "ctg_strand", "read_pos", "read_strand"])
Params = namedtuple("Params", ["k", "z", "a", "f", "x", "n", "sensitive", "repeat_filter"])


def default_params(k, z=500, a=1, f=10, x=0.0, n=1, sensitive=False, repeat_filter=False):
    "argparse defaults of bin/ntlink_pair.py:508-536"
    return Params(k, z, a, f, x, n, sensitive, repeat_filter)


# ----------------------------------------------------------------------------------------- inputs
def read_fasta_lengths(path):
    "contig -> length, id = header up to the first whitespace (bin/ntlink_utils.py:65-73, bin/read_fasta.py:17-19)"
    import gzip
    lengths = {}
    name, total = None, 0
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as fin:
        in_qual, qual_left = False, 0
        for line in fin:
            line = line.rstrip("\n")
            if in_qual:
                qual_left -= len(line)
                if qual_left <= 0:
                    in_qual = False
                continue
            if line[:1] in (">", "@"):
                if name is not None:
                    lengths[name] = total
                name, total = line[1:].split(None, 1)[0], 0
            elif line[:1] == "+" and name is not None:
                in_qual, qual_left = True, total
                if total == 0:
                    in_qual = False
            elif name is not None:
                total += len(line)
    if name is not None:
        lengths[name] = total
    return lengths


def read_target_index(tsv_lines):
    """M1 (bin/ntlink_pair.py:189-211): mx -> (contig, pos, strand); a minimizer seen two or more times
    anywhere in the target is dropped entirely. Keys stay decimal strings as in the reference."""
    index, multi = {}, set()
    for line in tsv_lines:
        fields = line.strip().split("\t")
        if len(fields) < 2:
            continue
        ctg = fields[0]
        for tok in fields[1].split(" "):
            mx, pos, strand = tok.split(":")
            if mx in index:
                multi.add(mx)
            else:
                index[mx] = (ctg, int(pos), strand)
    for mx in multi:
        del index[mx]
    return index


def read_hits(line, index, repeat_filter):
    "M2 (bin/ntlink_pair.py:355-378): returns (read_name, read_len, [(mx,pos,strand) hits]) or None"
    fields = line.strip().split("\t")
    if len(fields) < 3:
        return None
    hits = []
    for tok in fields[2].split(" "):
        mx, pos, strand = tok.split(":")
        if mx in index:
            hits.append((mx, int(pos), strand))
    if repeat_filter:
        count = {}
        for mx, _, _ in hits:
            count[mx] = count.get(mx, 0) + 1
        hits = [h for h in hits if count[h[0]] == 1]
    if not hits:
        return None
    return fields[0], int(fields[1]), hits


# ----------------------------------------------------------------------------------------- M3
def accepted_anchor_contigs(read_hits_, read_len, lengths, index, prm):
    """M3 (bin/ntlink_utils.py:200-268). Returns the ordered list [(contig, [Hit...])] of accepted contig
    runs (contigs are unique in the result)."""
    kept = []                       # (contig, Hit) in read order
    per_ctg = {}
    for mx, rpos, rstrand in read_hits_:
        ctg, cpos, cstrand = index[mx]
        if lengths[ctg] >= prm.z:
            h = Hit(mx, cpos, cstrand, rpos, rstrand)
            kept.append((ctg, h))
            per_ctg.setdefault(ctg, []).append(h)

    # noisy-contig filter: span on the contig larger than the read allows (utils:217-234)
    noisy = set()
    for ctg, hs in per_ctg.items():
        if len(hs) < 2:
            continue
        cps = [h.ctg_pos for h in hs]
        lo = cps.index(min(cps))    # first occurrence, like numpy argmin / argmax
        hi = cps.index(max(cps))
        span = abs(hs[hi].ctg_pos - hs[lo].ctg_pos)
        if prm.x == 0:
            if span > read_len + prm.k:
                noisy.add(ctg)
        else:
            thr = min(read_len + prm.k, (prm.x * abs(hs[hi].read_pos - hs[lo].read_pos)) + prm.k)
            if span > thr:
                noisy.add(ctg)
    kept = [t for t in kept if t[0] not in noisy]

    # runs of consecutive hits on the same contig (utils:236-237)
    runs = []                       # [contig, [hits]]
    for ctg, h in kept:
        if runs and runs[-1][0] == ctg:
            runs[-1][1].append(h)
        else:
            runs.append([ctg, [h]])

    dropped = [False] * len(runs)
    if prm.sensitive:
        # utils:271-278 -- only the runs strictly between consecutive occurrences of a contig
        occ = {}
        for i, (ctg, _) in enumerate(runs):
            occ.setdefault(ctg, []).append(i)
        for idxs in occ.values():
            for i, j in zip(idxs, idxs[1:]):
                for t in range(i + 1, j):
                    dropped[t] = True
    else:
        # utils:280-294 -- every contig named between the FIRST occurrence and a later occurrence
        first, bad = {}, set()
        for i, (ctg, _) in enumerate(runs):
            if ctg in first:
                for t in range(first[ctg] + 1, i):
                    bad.add(runs[t][0])
            else:
                first[ctg] = i
        for i, (ctg, _) in enumerate(runs):
            if ctg in bad:
                dropped[i] = True
    runs = [r for r, d in zip(runs, dropped) if not d]

    merged = []                     # utils:255-258
    for ctg, hs in runs:
        if merged and merged[-1][0] == ctg:
            merged[-1][1].extend(hs)
        else:
            merged.append((ctg, list(hs)))
    assert len({c for c, _ in merged}) == len(merged)      # utils:262-266
    return merged


# ----------------------------------------------------------------------------------------- M9
def _consistent(srt, increasing, i1, i2, dups):
    "paf:9-16"
    if srt[i1].ctg_pos in dups or srt[i2].ctg_pos in dups:
        return True
    if increasing:
        return srt[i1].read_pos <= srt[i2].read_pos
    return srt[i1].read_pos >= srt[i2].read_pos


def _filter_and_break(trans, srt, dups, increasing):
    "paf:18-58"
    breaks, filters = set(), set()
    for i, ok in enumerate(trans):
        if ok:
            continue
        if srt[i].ctg_pos in dups or srt[i + 1].ctg_pos in dups:
            continue
        if i + 2 >= len(trans):
            breaks.add(i + 1)
        elif _consistent(srt, increasing, i, i + 2, dups):
            filters.add(i + 1)
        elif i > 0 and _consistent(srt, increasing, i - 1, i + 1, dups):
            filters.add(i)
        else:
            breaks.add(i + 1)
    if not breaks and not filters:
        return [srt]
    blocks, cur = [], []
    for i, h in enumerate(srt):
        if i in filters:
            continue
        if i in breaks:
            blocks.append(cur)
            cur = [h]
        else:
            cur.append(h)
    blocks.append(cur)
    return blocks


def _mapped_blocks(srt):
    "paf:60-93"
    seen, dups = set(), set()
    t_inc, t_dec = [], []
    for a, b in zip(srt, srt[1:]):
        t_inc.append(a.read_pos <= b.read_pos)
        t_dec.append(a.read_pos >= b.read_pos)
        if a.ctg_pos in seen:
            dups.add(a.ctg_pos)
        else:
            seen.add(a.ctg_pos)
    if srt[-1].ctg_pos in seen:
        dups.add(srt[-1].ctg_pos)
    if all(t_inc) or all(t_dec):
        return [srt]
    n_inc = t_inc.count(True)
    if n_inc / len(t_inc) >= 0.75:
        return _filter_and_break(t_inc, srt, dups, True)
    if (len(t_inc) - n_inc) / len(t_inc) >= 0.75:
        return _filter_and_break(t_dec, srt, dups, False)
    return []


def paf_lines(read_name, read_len, accepted, lengths, k):
    "M9 (paf:95-135): list of PAF-like lines (with trailing newline) for one read"
    out = []
    for ctg, hits in accepted:
        srt = sorted(hits, key=lambda h: (h.ctg_pos, h.read_pos))
        if hits == srt or hits == sorted(srt, key=lambda h: (h.ctg_pos, h.read_pos), reverse=True):
            blocks = [srt]
        else:
            blocks = _mapped_blocks(srt)
        for blk in blocks:
            same = sum(1 for h in blk if h.ctg_strand == h.read_strand)
            strand = "+" if same / len(blk) * 100 >= 50 else "-"
            first, last = blk[0], blk[-1]
            ts, te = min(first.ctg_pos, last.ctg_pos), max(first.ctg_pos, last.ctg_pos) + k
            qs, qe = min(first.read_pos, last.read_pos), max(first.read_pos, last.read_pos) + k
            assert qs < qe and qs >= 0 and qe <= read_len       # paf:127-129
            out.append(f"{read_name}\t{read_len}\t{qs}\t{qe}\t{strand}\t{ctg}\t{lengths[ctg]}\t"
                       f"{ts}\t{te}\t{len(blk)}\t{te - ts}\t255\n")
    return out


# ----------------------------------------------------------------------------------------- M5/M6
def _flip(o):
    return "-" if o == "+" else "+"


def pair_event(run_i, run_j, read_len, lengths, k):
    """M6 (bin/ntlink_pair.py:315-334 with 157-187, 213-239). run = (contig, hits); run_i precedes run_j in the
    read. Returns (pair_key, gap, anchored) or None when |gap| > read length."""
    ci, hi = run_i
    cj, hj = run_j
    ti, fj = hi[-1], hj[0]          # terminal minimizer of i, first minimizer of j (M4)
    assert ti.read_pos < fj.read_pos
    oi = "+" if ti.read_strand == ti.ctg_strand else "-"
    oj = "+" if fj.read_strand == fj.ctg_strand else "-"
    a = lengths[ci] - ti.ctg_pos - k if oi == "+" else ti.ctg_pos
    b = fj.ctg_pos if oj == "+" else lengths[cj] - fj.ctg_pos - k
    assert a >= 0 and b >= 0
    gap = int((fj.read_pos - ti.read_pos) - a - b)
    key = (ci, oi, cj, oj) if ci < cj else (cj, _flip(oj), ci, _flip(oi))
    if abs(gap) > read_len:
        return None
    return key, gap, (len(hi) > 1 and len(hj) > 1)


def tally(accepted, read_len, pairs, lengths, prm):
    "M5 (bin/ntlink_pair.py:416-435). pairs: insertion-ordered dict key -> [gap list, anchor count]"
    def add(ri, rj, already=None):
        ev = pair_event(ri, rj, read_len, lengths, prm.k)
        if ev is None:
            return None
        key, gap, anchored = ev
        if already is not None and key in already:
            return None
        info = pairs.setdefault(key, [[], 0])
        info[0].append(gap)
        if anchored:
            info[1] += 1
        return key

    n = len(accepted)
    if n <= prm.f:
        for i in range(n):
            for j in range(i + 1, n):
                add(accepted[i], accepted[j])
    else:
        added = set()
        for ri, rj in zip(accepted, accepted[1:]):
            added.add(add(ri, rj))
        strong = [r for r in accepted if len(r[1]) > 1]
        for ri, rj in zip(strong, strong[1:]):
            add(ri, rj, already=added)


# ----------------------------------------------------------------------------------------- M7/M10
def gap_estimate(gaps):
    "int(np.median(list)) -- mean of the two middles for even n, truncated toward zero (pair:70-74)"
    s = sorted(gaps)
    n = len(s)
    if n % 2:
        return int(s[n // 2])
    return int((s[n // 2 - 1] + s[n // 2]) / 2)


def filter_pairs(pairs, lengths, min_anchor):
    "pair:241-255"
    out = {}
    for key, info in pairs.items():
        d = gap_estimate(info[0])
        if d <= -lengths[key[0]] or d <= -lengths[key[2]]:
            continue
        if info[1] >= min_anchor:
            out[key] = info
    return out


def pairs_tsv_lines(pairs):
    "pair:490-496, 80-83"
    return [f"{k[0]}{k[1]}\t{k[2]}{k[3]}\tn={len(v[0])}, gap_estimates={v[0]}, anchor={v[1]}\n"
            for k, v in pairs.items()]


def dot_lines(pairs, lengths, min_n):
    """pair:263-305 + 498-506 + 133-155. Node lines are emitted in first-seen order (the reference's order is
    Python-set order, i.e. PYTHONHASHSEED dependent) -- compare node lines as a multiset."""
    edges = {}                      # src -> {tgt: info}, insertion ordered like the reference's defaultdict
    nodes = []
    for key, info in pairs.items():
        s, t = key[0] + key[1], key[2] + key[3]
        rs, rt = key[2] + _flip(key[3]), key[0] + _flip(key[1])
        for v in (s, t, rs, rt):
            if v not in nodes:
                nodes.append(v)
        assert not (s in edges and t in edges[s])
        assert not (rs in edges and rt in edges[rs])
        edges.setdefault(s, {})[t] = info
        edges.setdefault(rs, {})[rt] = info
    largest = None
    for name in lengths:
        m = re.search(r"^ntLink_(\d+)$", name)
        if m and (largest is None or int(m.group(1)) > largest):
            largest = int(m.group(1))
    out = ["digraph G {\n", f"graph [scaf_num={largest}]\n"]
    for v in nodes:
        out.append(f"\"{v}\" [l={lengths[v[:-1]]}]\n")
    for s in edges:
        for t, info in edges[s].items():
            if len(info[0]) < min_n:
                continue
            out.append(f"\"{s}\" -> \"{t}\" [d={gap_estimate(info[0])} e=100 n={len(info[0])}]\n")
    out.append("}\n")
    return out


def verbose_lines(read_name, accepted):
    "M8 (pair:307-313, 382-388)"
    return ["{}\t{}\t{}\t{}\n".format(
        read_name, ctg, len(hits),
        " ".join(f"{h.ctg_pos}:{h.ctg_strand}_{h.read_pos}:{h.read_strand}" for h in hits))
        for ctg, hits in accepted]


# ----------------------------------------------------------------------------------------- driver
def map_reads(read_tsv_lines, index, lengths, prm, verbose_out=None, paf_out=None):
    "find_scaffold_pairs (pair:336-414): returns the raw (unfiltered) ordered pairs dict"
    pairs = {}
    for line in read_tsv_lines:
        parsed = read_hits(line, index, prm.repeat_filter)
        if parsed is None:
            continue
        name, rlen, hits = parsed
        accepted = accepted_anchor_contigs(hits, rlen, lengths, index, prm)
        if accepted:
            if verbose_out is not None:
                verbose_out.writelines(verbose_lines(name, accepted))
            if paf_out is not None:
                paf_out.writelines(paf_lines(name, rlen, accepted, lengths, prm.k))
        tally(accepted, rlen, pairs, lengths, prm)
    return pairs


def retally_from_verbose(verbose_lines, lengths, prm):
    """Checkpoint path (bin/ntlink_pair.py:437-488): re-tally the pairs from a verbose_mapping.tsv. Quirks kept:
    hit_count is the number of parsed hits, the 'read length' is the largest first/last read position of the read's
    runs (:487), a contig listed twice keeps only its last run (:467)."""
    pairs = {}

    def flush(group):
        if not group:
            return
        order, by_ctg, positions = [], {}, []
        for ctg, hits_str in group:
            hits = []
            for tok in hits_str.split(" "):
                c, r = tok.split("_")
                cp, cs = c.split(":")
                rp, rs = r.split(":")
                hits.append(Hit(None, int(cp), cs, int(rp), rs))
            order.append(ctg)
            by_ctg[ctg] = hits
            positions += [hits[0].read_pos, hits[-1].read_pos]
        tally([(c, by_ctg[c]) for c in order], max(positions), pairs, lengths, prm)

    cur, group = None, []
    for line in verbose_lines:
        read_id, ctg, _, hits_str = line.strip().split("\t")
        if read_id != cur:
            flush(group)
            cur, group = read_id, []
        group.append((ctg, hits_str))
    flush(group)
    return pairs


def run(files, target_fasta, target_tsv, prefix, prm, verbose=False, pairs_out=False, paf=False):
    "NtLink.main without the checkpoint path (pair:560-607)"
    with (sys.stdin if target_tsv == "-" else open(target_tsv)) as fin:
        index = read_target_index(fin)
    lengths = read_fasta_lengths(target_fasta)
    vf = open(prefix + ".verbose_mapping.tsv", "w") if verbose else None
    pf = open(prefix + ".paf", "w") if paf else None
    try:
        pairs = {}
        for path in files:
            with (sys.stdin if path == "-" else open(path)) as fin:
                got = map_reads(fin, index, lengths, prm, vf, pf)
            for key, info in got.items():       # several files behave like one concatenated stream
                cur = pairs.setdefault(key, [[], 0])
                cur[0].extend(info[0])
                cur[1] += info[1]
    finally:
        if vf:
            vf.close()
        if pf:
            pf.close()
    pairs = filter_pairs(pairs, lengths, prm.a)
    if pairs_out:
        with open(prefix + ".pairs.tsv", "w") as fout:
            fout.writelines(pairs_tsv_lines(pairs))
    with open(f"{prefix}.n{prm.n}.scaffold.dot", "w") as fout:
        fout.writelines(dot_lines(pairs, lengths, prm.n))
    return pairs


def main():
    ap = argparse.ArgumentParser(description="CPU oracle of the ntLink pairing stage (test infrastructure)")
    ap.add_argument("FILES", nargs="+")
    ap.add_argument("-s", required=True)
    ap.add_argument("-m", required=True)
    ap.add_argument("-p", default="out")
    ap.add_argument("-n", default=1, type=int)
    ap.add_argument("-k", required=True, type=int)
    ap.add_argument("-z", default=500, type=int)
    ap.add_argument("-a", default=1, type=int)
    ap.add_argument("-f", default=10, type=int)
    ap.add_argument("-x", default=0, type=float)
    ap.add_argument("--pairs", action="store_true")
    ap.add_argument("--paf", action="store_true")
    ap.add_argument("--sensitive", action="store_true")
    ap.add_argument("--repeat-filter", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    prm = Params(a.k, a.z, a.a, a.f, a.x, a.n, a.sensitive, a.repeat_filter)
    print(datetime.datetime.today(), ": pair_oracle start", file=sys.stderr)
    run(a.FILES, a.s, a.m, a.p, prm, verbose=a.verbose, pairs_out=a.pairs, paf=a.paf)
    print(datetime.datetime.today(), ": pair_oracle done", file=sys.stderr)


if __name__ == "__main__":
    main()
