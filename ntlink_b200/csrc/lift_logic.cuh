// Mapping liftover, per-read logic (host+device): rewrites the accepted runs of one read from the coordinates of the
// round-N contigs to the coordinates of the round-N+1 scaffolds through the AGP of the round.
//
// Replaces bin/ntlink_liftover_mappings.py (liftover_ctg_mappings :61-88, print_adjusted_mappings :90-124). The
// per-line / per-hit transform is a table lookup + integer arithmetic; the per-read regrouping keeps the reference's
// exact rules (also its quirks: subsumption is by contig NAME, counted from the FIRST run of the repeated contig).
#pragma once
#include "map_logic.cuh"

namespace ntl {

// one row per round-N contig id
struct AgpRow {
    uint32_t new_id;      // id of the path (or of the contig's own name when it is not in the AGP) in the new namespace
    uint32_t flags;       // AGP_IN | AGP_MINUS | AGP_KEEP
    uint32_t scaf_start;  // 1-based start of the component on the path
    uint32_t ctg_start;   // 1-based first used base of the contig
    uint32_t ctg_end;     // 1-based last used base of the contig
};
enum : uint32_t {
    AGP_IN = 1,           // contig has an AGP entry (liftover:65-66: otherwise the line keeps its name and loses its hits)
    AGP_MINUS = 2,        // orientation '-'
    AGP_KEEP = 4          // path id == contig id, or an orientation other than +/-: hits are passed through (:84-85)
};
enum : uint32_t { LIFTERR_RANGE = 1, LIFTERR_LAYOUT = 2 };

// Lifts the `nr` runs of one read. Input: runs/hits of the read's region (Run.start relative to the region).
// `cap` = slots of the region, `ncontig` = rows of the AGP table. Output: runs_out/hits_out of the same region (same
// capacity, holes allowed), returns the number of output runs. tmp_id / tmp_kept: scratch of >= nr words each.
// hits_out may NOT alias hits. The runs must lie in the region in ascending, non-overlapping order (as every
// verbose_mapping.tsv parses to); anything else sets LIFTERR_LAYOUT and yields no runs.
NTL_HD uint32_t lift_read(const Hit* hits, const Run* runs, uint32_t nr, uint32_t cap, const AgpRow* agp, uint32_t ncontig,
                          int32_t k, Hit* hits_out, Run* runs_out, uint32_t* tmp_id, uint32_t* tmp_kept, uint32_t* err) {
    const uint32_t BETWEEN = 0x80000000u, SUBSUMED = 0x40000000u, CNT = 0x3FFFFFFFu;
    if (nr > cap) { *err |= LIFTERR_LAYOUT; return 0; }
    uint64_t end_prev = 0;
    for (uint32_t i = 0; i < nr; i++) {
        const Run ru = runs[i];
        if (ru.ctg >= ncontig || ru.start < end_prev || (uint64_t)ru.start + ru.count > cap) { *err |= LIFTERR_LAYOUT; return 0; }
        end_prev = (uint64_t)ru.start + ru.count;
    }
    // phase 1 (:61-88): per line, the new contig id and the surviving hits in the new coordinates (kept in place)
    for (uint32_t i = 0; i < nr; i++) {
        const Run ru = runs[i];
        const AgpRow row = agp[ru.ctg];
        tmp_id[i] = row.new_id;
        uint32_t kept = 0;
        if (row.flags & AGP_IN) {
            const int64_t lo = (int64_t)row.ctg_start - 1, hi = (int64_t)row.ctg_end - k;
            const int64_t offset = (int64_t)row.scaf_start - 1;
            const int64_t ctg_len = (int64_t)row.ctg_end - (int64_t)row.ctg_start + 1;
            for (uint32_t q = 0; q < ru.count; q++) {
                Hit h = hits[ru.start + q];
                const int64_t pos = pos_of(h.cposf);
                if (pos < lo || pos > hi) continue;                       // outside of the used contig region (:73)
                if (!(row.flags & AGP_KEEP)) {
                    const int64_t adjust = pos - lo;
                    int64_t np;
                    uint32_t fwd = fwd_of(h.cposf);
                    if (row.flags & AGP_MINUS) { np = offset + (ctg_len - adjust) - k; fwd ^= 1u; }
                    else np = offset + adjust;
                    if (np < 0 || np > 0x7FFFFFFF) { *err |= LIFTERR_RANGE; np = 0; }
                    h.cposf = (uint32_t)np | (fwd << 31);
                }
                h.ctg = row.new_id;
                hits_out[ru.start + kept++] = h;
            }
        }
        tmp_kept[i] = kept;
    }
    // phase 2 (:93-105): runs of consecutive lines with one id; a contig seen again marks every run after its FIRST run
    // and before this one, and a marked run subsumes its contig NAME (all lines carrying that id)
    for (uint32_t i = 1; i < nr; i++) {
        if (tmp_id[i] == tmp_id[i - 1]) continue;                         // not the start of a run
        uint32_t p = 0;
        while (tmp_id[p] != tmp_id[i]) p++;
        if (p == i) continue;                                              // first run of this contig
        uint32_t e = p;
        while (tmp_id[e + 1] == tmp_id[p]) e++;                            // e < i - 1 here, so e + 1 is in range
        for (uint32_t j = e + 1; j < i; j++) tmp_kept[j] |= BETWEEN;
    }
    for (uint32_t i = 0; i < nr; i++) {
        if (!(tmp_kept[i] & BETWEEN)) continue;
        for (uint32_t j = 0; j < nr; j++)
            if (tmp_id[j] == tmp_id[i]) tmp_kept[j] |= SUBSUMED;
    }
    // phase 3 (:107-124): regroup what is left, concatenate, keep strictly monotonic non-empty groups
    uint32_t nout = 0, cursor = 0;
    uint32_t i = 0;
    while (i < nr) {
        if (tmp_kept[i] & SUBSUMED) { i++; continue; }
        const uint32_t id = tmp_id[i];
        const uint32_t begin = cursor;
        bool inc = true, dec = true;
        uint32_t prev = 0;
        while (i < nr) {
            if (tmp_kept[i] & SUBSUMED) { i++; continue; }                 // removed lines do not split a group
            if (tmp_id[i] != id) break;
            const uint32_t src = runs[i].start, cnt = tmp_kept[i] & CNT;
            for (uint32_t q = 0; q < cnt; q++) {
                const Hit h = hits_out[src + q];                           // cursor <= src + q: ascending copy is safe
                const uint32_t pos = pos_of(h.cposf);
                if (cursor > begin) {
                    if (!(prev < pos)) inc = false;
                    if (!(prev > pos)) dec = false;
                }
                prev = pos;
                hits_out[cursor++] = h;
            }
            i++;
        }
        if (cursor == begin) continue;                                     // nothing survived (:113-114)
        if (!inc && !dec) { cursor = begin; continue; }                    // neither increasing nor decreasing (:115-119)
        Run o; o.ctg = id; o.start = begin; o.count = cursor - begin;
        runs_out[nout++] = o;
    }
    return nout;
}

}  // namespace ntl
