"""
Batched GPU counterpart of the re-mapping step of ntLink's gap filling (bin/ntlink_patch_gaps.py:412-442 of bcgsc/ntLink
v1.3.11): for every gap, the two masked scaffold ends and the masked chosen read are sketched (k20/w10 by default) and the
read is chained against the two ends exactly like ntlink_utils.get_accepted_anchor_contigs does.

The reference walks the gaps one by one through two btllib.Indexlr iterators and a Python dict per gap; here both masked
FASTA files are sketched in one pass each and ALL gaps are mapped by one ntl_map_groups call (one CUDA block per gap).
What comes back has the shape the rest of map_long_reads reads (assess_accepted_anchor_contigs and the cut logic,
patch:443-489, stay as they are): per gap an ordered list of runs with .contig, .hits (MinimizerPositions) and .hit_count.
"""
import re
from collections import namedtuple

import numpy as np

from . import api

READ_HEADER_RE = re.compile(r"^(\S+)__(\S+)__(\S+)$")            # patch:414
SCAFFOLD_HEADER_RE = re.compile(r"^(\S+)_(source|target)$")      # patch:415

MinimizerPositions = namedtuple("MinimizerPositions", ["mx", "ctg_pos", "ctg_strand", "read_pos", "read_strand"])   # utils:19-21
AnchorRun = namedtuple("AnchorRun", ["contig", "hits", "hit_count"])
GapMapping = namedtuple("GapMapping", ["read", "source", "target", "accepted"])


def pair_up(scaffold_names, read_names):
    """The masked files are written gap by gap (patch:346-389): read `<read>__<source>__<target>`, scaffold records
    `<source>_source`, `<target>_target`. Returns [(read, source, target)] and checks the order like patch:424-434."""
    if len(scaffold_names) != 2 * len(read_names):
        raise ValueError("masked scaffold file must hold two records per masked read")
    gaps = []
    for g, rname in enumerate(read_names):
        m = READ_HEADER_RE.search(rname)
        if not m:
            raise ValueError(f"unexpected masked read header {rname!r}")
        read, source, target = m.groups()
        for name, want, label in ((scaffold_names[2 * g], source, "source"), (scaffold_names[2 * g + 1], target, "target")):
            ms = SCAFFOLD_HEADER_RE.search(name)
            if not ms or ms.group(1) != want or ms.group(2) != label:
                raise ValueError(f"masked scaffold record {name!r} does not match read {rname!r}")
        gaps.append((read, source, target))
    return gaps


def map_long_reads(ctx, scaffolds_masked_fa, reads_masked_fa, scaffold_lengths, k=20, w=10, z=1000, x=0.0, sensitive=False):
    """scaffold_lengths: scaffold id (no orientation) -> length of the full scaffold (what the z filter looks at, utils:206).
    Returns [GapMapping]; .accepted lists the accepted contigs in read order, contig = scaffold id without orientation."""
    ends = api.read_sequences(scaffolds_masked_fa)
    reads = api.read_sequences(reads_masked_fa)
    gaps = pair_up(ends.names, reads.names)
    if not gaps:
        return []
    ids = [SCAFFOLD_HEADER_RE.search(n).group(1).strip("+-") for n in ends.names]
    t_len = np.array([scaffold_lengths[i] for i in ids], np.uint32)
    t_sk = ctx.sketch(ends, k, w)
    r_sk = ctx.sketch(reads, k, w)
    prm = ctx.params(k, w, z, 10, x, sensitive, False)
    res = ctx.map_groups(t_sk, t_len, np.arange(0, len(ends) + 1, 2, dtype=np.uint32), r_sk, reads.lengths.astype(np.uint32), prm)
    out = []
    for g, (read, source, target) in enumerate(gaps):
        base = int(res.hit_off[g])
        accepted = []
        for ctg, start, count in res.runs[base:base + int(res.nruns[g])]:
            hs = res.hits[base + int(start):base + int(start) + int(count)]
            hits = [MinimizerPositions(None, int(h[1]) & 0x7FFFFFFF, "+" if int(h[1]) >> 31 else "-",
                                       int(h[2]) & 0x7FFFFFFF, "+" if int(h[2]) >> 31 else "-") for h in hs]
            accepted.append(AnchorRun(ids[int(ctg)], hits, len(hits)))
        if ids[2 * g] == ids[2 * g + 1] and len(accepted) > 1:
            # both ends of the gap belong to the same scaffold (A+ -> A-): the reference keys its minimizer table, the runs and
            # the accepted dict by scaffold id (patch:427-442), so it can never see two accepted contigs here and falls back;
            # the two runs are folded into one entry so that the caller's `len(accepted) != 2` test takes the same branch
            merged = [h for run in accepted for h in run.hits]
            accepted = [AnchorRun(ids[2 * g], merged, len(merged))]
        out.append(GapMapping(read, source, target, accepted))
    return out
