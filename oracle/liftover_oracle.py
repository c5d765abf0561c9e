"""
CPU restatement of ntLink's mapping liftover (bin/ntlink_liftover_mappings.py of bcgsc/ntLink v1.3.11).

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing else); the product path is the CUDA kernel k_liftover behind
ntl_liftover_mappings. Pinned on goldens produced by the unmodified reference script (tests/golden/liftover/, generated
by tests/golden/make_golden.py), see tests/test_oracle_liftover.py.
"""
from collections import namedtuple

AgpEntry = namedtuple("AgpEntry", "path_id scaf_start scaf_end ctg_start ctg_end orientation")
COVERAGE = {"lines": 0, "not_in_agp": 0, "hits_outside_region": 0, "identity": 0, "subsumed_lines": 0, "merged_groups": 0,
            "non_monotonic_groups": 0, "empty_groups": 0, "decreasing_groups": 0}


def read_agp(lines):
    "liftover:40-50: one entry per contig (the last one wins), gap lines (N, P) skipped"
    agp = {}
    for line in lines:
        f = line.strip().split("\t")
        path_id, scaf_start, scaf_end, _, ctype, ctg_id, ctg_start, ctg_end, orientation = f
        if ctype in ("N", "P"):
            continue
        agp[ctg_id] = AgpEntry(path_id, int(scaf_start), int(scaf_end), int(ctg_start), int(ctg_end), orientation)
    return agp


def lift_line(ctg, hits_str, agp, k):
    "liftover:61-88 -> (new contig id, [(ctg_pos, ctg_strand, read_pos, read_strand)])"
    COVERAGE["lines"] += 1
    if ctg not in agp:
        COVERAGE["not_in_agp"] += 1
        return ctg, []
    e = agp[ctg]
    out = []
    for tok in hits_str.split(" "):
        c, r = tok.split("_")
        cp, cs = c.split(":")
        rp, rs = r.split(":")
        cp, rp = int(cp), int(rp)
        if not e.ctg_start - 1 <= cp <= e.ctg_end - k:                      # :73
            COVERAGE["hits_outside_region"] += 1
            continue
        adjust = cp - (e.ctg_start - 1)
        offset = e.scaf_start - 1
        if e.orientation == "+" and e.path_id != ctg:
            out.append((offset + adjust, cs, rp, rs))
        elif e.orientation == "-" and e.path_id != ctg:
            out.append((offset + (e.ctg_end - e.ctg_start + 1 - adjust) - k, "-" if cs == "+" else "+", rp, rs))
        else:
            COVERAGE["identity"] += 1
            out.append((cp, cs, rp, rs))                                      # :85 untouched, not even re-based
    return e.path_id, out


def lift_read(read_id, lines, agp, k):
    "liftover:90-124 for the consecutive lines [(ctg, hits_str)] of one read -> output lines"
    lifted = [lift_line(ctg, hs, agp, k) for ctg, hs in lines]
    groups = []                                   # consecutive lines with the same new contig id
    for new_ctg, hits in lifted:
        if groups and groups[-1][0] == new_ctg:
            groups[-1][1].append(hits)
        else:
            groups.append((new_ctg, [hits]))
    first, subsumed = {}, set()
    for i, (ctg, _) in enumerate(groups):
        if ctg not in first:
            first[ctg] = i
        else:                                     # everything between the FIRST run of ctg and this one (:101-102)
            for j in range(first[ctg] + 1, i):
                subsumed.add(groups[j][0])
    COVERAGE["subsumed_lines"] += sum(1 for c, _ in lifted if c in subsumed)
    kept = [(c, h) for c, h in lifted if c not in subsumed]
    out, i = [], 0
    while i < len(kept):
        j = i
        hits = []
        while j < len(kept) and kept[j][0] == kept[i][0]:
            hits += kept[j][1]
            j += 1
        if j - i > 1:
            COVERAGE["merged_groups"] += 1
        ctg = kept[i][0]
        i = j
        if not hits:
            COVERAGE["empty_groups"] += 1
            continue
        inc = all(a[0] < b[0] for a, b in zip(hits, hits[1:]))
        if not inc:
            if not all(a[0] > b[0] for a, b in zip(hits, hits[1:])):
                COVERAGE["non_monotonic_groups"] += 1
                continue
            COVERAGE["decreasing_groups"] += 1
        mx = " ".join(f"{cp}:{cs}_{rp}:{rs}" for cp, cs, rp, rs in hits)
        out.append(f"{read_id}\t{ctg}\t{len(hits)}\t{mx}\n")
    return out


def liftover(verbose_lines, agp_lines, k):
    "liftover:128-147"
    agp = read_agp(agp_lines)
    out, cur, block = [], None, []
    for line in verbose_lines:
        read_id, ctg, _, hits_str = line.strip().split("\t")
        if read_id != cur:
            if cur is not None:
                out += lift_read(cur, block, agp, k)
            cur, block = read_id, []
        block.append((ctg, hits_str))
    if cur is not None:
        out += lift_read(cur, block, agp, k)
    return out
