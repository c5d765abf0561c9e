// map_logic.cuh -- per-thread logic of the minimizer-mapping kernels (one thread = one read).
//
// Replaces, with identical results, the per-read part of the reference's pairing stage:
//   M2  lookup + --repeat-filter                  bin/ntlink_pair.py:352-376
//   M3  get_accepted_anchor_contigs               bin/ntlink_utils.py:200-294
//   M4  first / terminal minimizer                bin/ntlink_pair.py:395-406
//   M5  tally_pairs_from_mappings                 bin/ntlink_pair.py:416-435
//   M6  add_pair / calculate_pair_info / calculate_gap_size / normalize_pair
//                                                 bin/ntlink_pair.py:157-187,213-239,315-334
// Like sketch_logic.cuh it is sequential per-thread code that also compiles for the host so the CPU unit
// tests (tests/emu/) can check it against the oracle without a GPU. The product only runs it on the device.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <vector>
#include "nthash.cuh"

namespace ntl {

// ------------------------------------------------------------------------------------ target index (M1)
// Open-addressing table of the target minimizers that occur exactly once (bin/ntlink_pair.py:189-211).
struct IdxEntry {
    uint64_t key;      // printed hash h1; EMPTY_KEY = free slot
    uint32_t ctg;      // contig id (FASTA order); DUP_CTG = hash seen more than once -> treated as absent
    uint32_t posf;     // position | (forward-strand flag << 31)
};
#define NTL_EMPTY_KEY 0xFFFFFFFFFFFFFFFFULL
enum : uint32_t { DUP_CTG = 0xFFFFFFFFu, NONE32_M = 0xFFFFFFFFu };

NTL_HD uint64_t idx_slot(uint64_t key, uint64_t mask) {
    uint64_t x = key * 0x9E3779B97F4A7C15ULL;          // keys are already hashes; one multiply spreads low bits
    return (x >> 20) & mask;
}

// A hash equal to the EMPTY sentinel cannot live in the table; it is kept in a one-entry side slot.
struct IdxSpecial { uint32_t count; uint32_t ctg; uint32_t posf; uint32_t pad; };

struct IndexView {
    const IdxEntry* table;
    uint64_t mask;
    const IdxSpecial* special;
};

// returns true and fills (ctg, posf) when `key` is a unique target minimizer
NTL_HD bool index_lookup(const IndexView& ix, uint64_t key, uint32_t& ctg, uint32_t& posf) {
    if (key == NTL_EMPTY_KEY) {
        if (ix.special->count != 1) return false;
        ctg = ix.special->ctg; posf = ix.special->posf;
        return true;
    }
    uint64_t s = idx_slot(key, ix.mask);
    for (;;) {
        const IdxEntry e = ix.table[s];
        if (e.key == key) {
            if (e.ctg == DUP_CTG) return false;
            ctg = e.ctg; posf = e.posf;
            return true;
        }
        if (e.key == NTL_EMPTY_KEY) return false;
        s = (s + 1) & ix.mask;
    }
}

// ------------------------------------------------------------------------------------ chaining (M3)
struct Hit {
    uint32_t ctg;      // contig id
    uint32_t cposf;    // contig position | (contig strand forward << 31)
    uint32_t rposf;    // read position   | (read strand forward << 31)
};
struct Run {           // an accepted contig run of one read: hits [start, start+count) of the read's region
    uint32_t ctg, start, count;
};

struct MapParams {
    int32_t k, z, f;
    int32_t x_is_zero;
    double x;
    int32_t sensitive, repeat_filter;
};

NTL_HD uint32_t pos_of(uint32_t pf) { return pf & 0x7FFFFFFFu; }
NTL_HD uint32_t fwd_of(uint32_t pf) { return pf >> 31; }

#if defined(__CUDA_ARCH__)
#define NTL_DMUL(a, b) __dmul_rn((a), (b))
#define NTL_DADD(a, b) __dadd_rn((a), (b))
#else   /* host build (tests/emu only): compiled with -ffp-contract=off */
#define NTL_DMUL(a, b) ((a) * (b))
#define NTL_DADD(a, b) ((a) + (b))
#endif

// Chains the hits of one read. `hits[0..nh)` are the index hits in read order (input, clobbered: the
// accepted hits are compacted to the front in output order); `runs` and `mark` are scratch/output arrays
// with room for nh entries. Returns the number of accepted runs (contigs), unique per contig.
NTL_HD uint32_t chain_read(Hit* hits, uint32_t nh, Run* runs, uint8_t* mark, uint32_t read_len,
                           const uint32_t* ctg_len, const MapParams& P) {
    // ---- M2 repeat filter: drop every minimizer that hits more than once within this read
    //      (same index entry <=> same (contig, position))
    uint32_t m = nh;
    if (P.repeat_filter) {
        for (uint32_t i = 0; i < nh; i++) mark[i] = 0;
        for (uint32_t i = 0; i < nh; i++) {
            if (mark[i]) continue;
            for (uint32_t j = i + 1; j < nh; j++)
                if (hits[j].ctg == hits[i].ctg && hits[j].cposf == hits[i].cposf) { mark[i] = 1; mark[j] = 1; }
        }
        m = 0;
        for (uint32_t i = 0; i < nh; i++) if (!mark[i]) hits[m++] = hits[i];
    }
    // ---- contigs shorter than z are ignored (utils:206)
    {
        uint32_t o = 0;
        for (uint32_t i = 0; i < m; i++) if ((int64_t)ctg_len[hits[i].ctg] >= (int64_t)P.z) hits[o++] = hits[i];
        m = o;
    }
    if (m == 0) return 0;
    // ---- runs of consecutive hits on one contig
    uint32_t nr = 0;
    for (uint32_t i = 0; i < m; i++) {
        if (nr && runs[nr - 1].ctg == hits[i].ctg) runs[nr - 1].count++;
        else { runs[nr].ctg = hits[i].ctg; runs[nr].start = i; runs[nr].count = 1; nr++; }
    }
    // ---- noisy contigs (utils:217-234): span on the contig between the first-minimum and first-maximum
    //      contig position exceeds what the read can cover
    for (uint32_t r = 0; r < nr; r++) mark[r] = 0;
    for (uint32_t r = 0; r < nr; r++) {
        const uint32_t c = runs[r].ctg;
        bool seen_before = false;
        for (uint32_t q = 0; q < r; q++) if (runs[q].ctg == c) { seen_before = true; break; }
        if (seen_before) continue;
        uint32_t total = 0, mn = 0, mx = 0, mn_r = 0, mx_r = 0;
        bool any = false;
        for (uint32_t q = r; q < nr; q++) {
            if (runs[q].ctg != c) continue;
            for (uint32_t i = runs[q].start; i < runs[q].start + runs[q].count; i++) {
                const uint32_t cp = pos_of(hits[i].cposf), rp = pos_of(hits[i].rposf);
                if (!any) { mn = mx = cp; mn_r = mx_r = rp; any = true; }
                else {
                    if (cp < mn) { mn = cp; mn_r = rp; }       // strict: first occurrence wins (numpy argmin)
                    if (cp > mx) { mx = cp; mx_r = rp; }
                }
                total++;
            }
        }
        if (total < 2) continue;
        const int64_t span = (int64_t)mx - (int64_t)mn;
        bool noisy;
        if (P.x_is_zero) noisy = span > (int64_t)read_len + P.k;
        else {
            // threshold = min(read_len + k, x * |read_pos[hi] - read_pos[lo]| + k) in IEEE double, unfused
            const int64_t dr = mx_r > mn_r ? (int64_t)mx_r - mn_r : (int64_t)mn_r - mx_r;
            double thr = NTL_DADD(NTL_DMUL(P.x, (double)dr), (double)P.k);
            const double cap = (double)((int64_t)read_len + P.k);
            if (cap < thr) thr = cap;
            noisy = (double)span > thr;
        }
        if (noisy) for (uint32_t q = r; q < nr; q++) if (runs[q].ctg == c) mark[q] = 1;
    }
    // drop noisy runs, compact the hits and re-group (adjacent runs of one contig merge)
    {
        uint32_t o = 0, nr2 = 0;
        for (uint32_t r = 0; r < nr; r++) {
            if (mark[r]) continue;
            const Run in = runs[r];
            for (uint32_t i = 0; i < in.count; i++) hits[o + i] = hits[in.start + i];
            if (nr2 && runs[nr2 - 1].ctg == in.ctg) runs[nr2 - 1].count += in.count;
            else { runs[nr2].ctg = in.ctg; runs[nr2].start = o; runs[nr2].count = in.count; nr2++; }
            o += in.count;
        }
        nr = nr2; m = o;
    }
    if (nr == 0) return 0;
    // ---- subsumption
    for (uint32_t r = 0; r < nr; r++) mark[r] = 0;
    if (P.sensitive) {
        // utils:271-278: for consecutive occurrences (i, j) of a contig drop the runs strictly between
        for (uint32_t j = 1; j < nr; j++) {
            uint32_t prev = NONE32_M;
            for (uint32_t i = j; i-- > 0;) if (runs[i].ctg == runs[j].ctg) { prev = i; break; }
            if (prev == NONE32_M) continue;
            for (uint32_t t = prev + 1; t < j; t++) mark[t] = 1;
        }
    } else {
        // utils:280-294: every contig NAMED in a run strictly between the FIRST occurrence of a contig and any
        // later occurrence is dropped everywhere (bit 1 = named in between, bit 0 = dropped)
        for (uint32_t j = 1; j < nr; j++) {
            uint32_t first = NONE32_M;
            for (uint32_t i = 0; i < j; i++) if (runs[i].ctg == runs[j].ctg) { first = i; break; }
            if (first == NONE32_M) continue;
            for (uint32_t t = first + 1; t < j; t++) mark[t] |= 2;
        }
        for (uint32_t r = 0; r < nr; r++) {
            if (!(mark[r] & 2)) continue;
            for (uint32_t q = 0; q < nr; q++) if (runs[q].ctg == runs[r].ctg) mark[q] |= 1;
        }
        for (uint32_t r = 0; r < nr; r++) mark[r] &= 1;
    }
    // drop subsumed runs, compact, merge adjacent (utils:253-258)
    {
        uint32_t o = 0, nr2 = 0;
        for (uint32_t r = 0; r < nr; r++) {
            if (mark[r]) continue;
            const Run in = runs[r];
            for (uint32_t i = 0; i < in.count; i++) hits[o + i] = hits[in.start + i];
            if (nr2 && runs[nr2 - 1].ctg == in.ctg) runs[nr2 - 1].count += in.count;
            else { runs[nr2].ctg = in.ctg; runs[nr2].start = o; runs[nr2].count = in.count; nr2++; }
            o += in.count;
        }
        nr = nr2;
    }
    return nr;
}

// ------------------------------------------------------------------------------------ pair events (M5/M6)
struct Event {
    uint32_t read;      // global read ordinal
    uint32_t ord;       // ordinal of the event within its read (reference insertion order)
    uint32_t src, tgt;  // contig ids after normalisation (lexicographically smaller NAME first)
    int32_t gap;
    uint32_t flags;     // bit0 = source orientation '+', bit1 = target orientation '+', bit2 = anchored
};

// One candidate pair (run i precedes run j in the read). Returns false when |gap| > read length.
NTL_HD bool make_event(const Hit* hits, const Run& ri, const Run& rj, uint32_t read_len, const uint32_t* ctg_len,
                       const uint32_t* name_rank, int32_t k, Event& ev) {
    const Hit ti = hits[ri.start + ri.count - 1];   // terminal minimizer of i
    const Hit fj = hits[rj.start];                  // first minimizer of j
    const bool oi = fwd_of(ti.rposf) == fwd_of(ti.cposf);
    const bool oj = fwd_of(fj.rposf) == fwd_of(fj.cposf);
    const int64_t a = oi ? (int64_t)ctg_len[ri.ctg] - pos_of(ti.cposf) - k : (int64_t)pos_of(ti.cposf);
    const int64_t b = oj ? (int64_t)pos_of(fj.cposf) : (int64_t)ctg_len[rj.ctg] - pos_of(fj.cposf) - k;
    const int64_t gap = ((int64_t)pos_of(fj.rposf) - (int64_t)pos_of(ti.rposf)) - a - b;
    if (name_rank[ri.ctg] < name_rank[rj.ctg]) {
        ev.src = ri.ctg; ev.tgt = rj.ctg; ev.flags = (oi ? 1u : 0u) | (oj ? 2u : 0u);
    } else {                                        // swap and flip both orientations (pair:216-219)
        ev.src = rj.ctg; ev.tgt = ri.ctg; ev.flags = (oj ? 0u : 1u) | (oi ? 0u : 2u);
    }
    ev.gap = (int32_t)gap;
    if (ri.count > 1 && rj.count > 1) ev.flags |= 4u;
    // the reference asserts mx_i_pos < mx_j_pos (pair:225) and a, b >= 0 (pair:173-184); this cannot fail for
    // mappings produced by the chaining above, but a hand-made / lifted-over checkpoint file can violate it. Such an
    // event is flagged (bit 3) and always written so that the caller can abort like the reference does.
    if (!(pos_of(ti.rposf) < pos_of(fj.rposf)) || a < 0 || b < 0) { ev.flags |= 8u; return true; }
    const int64_t ag = gap < 0 ? -gap : gap;
    return ag <= (int64_t)read_len;
}

// upper bound of the number of events a read with nr accepted contigs can emit
// (64-bit arithmetic, saturated at MAX_EVENTS_PER_READ + 1: the order key of the pair tally holds the event's ordinal
// within its read in 24 bits; a read that could emit more is rejected with an explicit error instead of colliding keys)
enum : uint32_t { MAX_EVENTS_PER_READ = 0xFFFFFFu };
NTL_HD uint32_t max_events(uint32_t nr, int32_t f) {
    if (nr < 2) return 0;
    const uint64_t n = (int64_t)nr <= (int64_t)f ? (uint64_t)nr * (nr - 1) / 2 : 2ull * (nr - 1);
    return n > MAX_EVENTS_PER_READ ? MAX_EVENTS_PER_READ + 1u : (uint32_t)n;
}

// Writes the events of one read in reference order; returns how many were written.
NTL_HD uint32_t tally_read(const Hit* hits, const Run* runs, uint32_t nr, uint32_t read_len, uint32_t read_ord,
                           const uint32_t* ctg_len, const uint32_t* name_rank, const MapParams& P, Event* out) {
    uint32_t ne = 0;
    if ((int64_t)nr <= (int64_t)P.f) {
        for (uint32_t i = 0; i < nr; i++)
            for (uint32_t j = i + 1; j < nr; j++) {
                Event ev;
                if (make_event(hits, runs[i], runs[j], read_len, ctg_len, name_rank, P.k, ev)) {
                    ev.read = read_ord; ev.ord = ne; out[ne++] = ev;
                }
            }
        return ne;
    }
    for (uint32_t i = 0; i + 1 < nr; i++) {                        // adjacent pairs
        Event ev;
        if (make_event(hits, runs[i], runs[i + 1], read_len, ctg_len, name_rank, P.k, ev)) {
            ev.read = read_ord; ev.ord = ne; out[ne++] = ev;
        }
    }
    const uint32_t n_adj = ne;
    uint32_t prev = NONE32_M;                                      // transitive pairs over weak contigs
    for (uint32_t i = 0; i < nr; i++) {
        if (runs[i].count <= 1) continue;
        if (prev != NONE32_M) {
            Event ev;
            if (make_event(hits, runs[prev], runs[i], read_len, ctg_len, name_rank, P.k, ev)) {
                bool dup = false;                                  // check_added (pair:324-325)
                for (uint32_t q = 0; q < n_adj; q++)
                    if (out[q].src == ev.src && out[q].tgt == ev.tgt && ((out[q].flags ^ ev.flags) & 3u) == 0) { dup = true; break; }
                if (!dup || (ev.flags & 8u)) { ev.read = read_ord; ev.ord = ne; out[ne++] = ev; }
            }
        }
        prev = i;
    }
    return ne;
}

// Host side of the pair table: the permutation that puts the pairs in first-seen order (the reference's dict order),
// i.e. a stable ascending sort of their 64-bit first-seen keys. LSD radix sort on the bytes that actually differ: at
// human scale (2 x 10^5 pairs on rank 0) a comparison sort of the records was the largest part of the tally time.
inline void order_first_seen(const uint64_t* keys, uint32_t n, uint32_t* perm) {
    if (n < 2) { if (n) perm[0] = 0; return; }
    struct Rec { uint64_t k; uint64_t i; };
    std::vector<Rec> a(n), b(n);
    uint64_t differ = 0;
    for (uint32_t i = 0; i < n; i++) { a[i].k = keys[i]; a[i].i = i; differ |= keys[i] ^ keys[0]; }
    Rec* src = a.data();
    Rec* dst = b.data();
    int lo = 0, hi = 63;                                            // bits that differ between the keys
    while (lo < 64 && !((differ >> lo) & 1ull)) lo++;
    while (hi > lo && !((differ >> hi) & 1ull)) hi--;
    constexpr int DIGIT = 11;                                       // 2048 buckets: four passes over a 44-bit span
    std::vector<uint32_t> count((1u << DIGIT) + 1);
    for (int shift = lo; shift <= hi && differ; shift += DIGIT) {
        const uint64_t mask = (1ull << DIGIT) - 1;
        if (((differ >> shift) & mask) == 0) continue;
        std::fill(count.begin(), count.end(), 0u);
        for (uint32_t i = 0; i < n; i++) count[((src[i].k >> shift) & mask) + 1]++;
        for (uint32_t d = 0; d < (1u << DIGIT); d++) count[d + 1] += count[d];
        for (uint32_t i = 0; i < n; i++) dst[count[(src[i].k >> shift) & mask]++] = src[i];
        Rec* t = src; src = dst; dst = t;
    }
    for (uint32_t i = 0; i < n; i++) perm[i] = (uint32_t)src[i].i;
}

}  // namespace ntl
