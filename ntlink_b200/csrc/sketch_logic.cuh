// sketch_logic.cuh -- per-thread logic of the minimizer sketch (indexlr --long --pos --strand) kernels.
//
// Everything here is plain sequential per-thread code (no intra-block communication), written so that
// the same functions compile for the device (kernels_sketch.cu) and for the host (tests/emu/, a CPU
// emulation used ONLY by the CPU unit tests to debug kernel logic against the oracle; it is never
// loaded by the product library).
//
// Algorithm (replaces btllib Indexlr::minimize; reference call sites ntLink:198-199,
// 221-225). The reference slides one window sequentially over each sequence; that is the wrong shape for
// a GPU. The minimizer SET of a sequence is { rightmost-argmin of every window of w consecutive valid
// k-mers }, so we use the equivalent position-local characterisation:
//
//   position i is a minimizer  <=>  some window [j, j+w-1] (in valid-k-mer index space, 0 <= j <= n-w)
//                                   contains i, has only values >= h0[i] to the left of i and only
//                                   values > h0[i] to the right of i.
//
//   1. DENSE pass (process_strip): every thread rolls ntHash over one strip of S k-mer positions and keeps
//      only "candidates": k-mers whose canonical hash is below a threshold tau ~ c/w of the hash space
//      (a few % of all positions). A window that contains a candidate has its minimum among the candidates.
//   2. SPARSE pass (select_strip): per candidate, scan neighbouring candidates left (first smaller) and right
//      (first smaller-or-equal) and apply the characterisation above. Exact, including ties.
//   3. GAPS (gap_scan): a stretch of >= w consecutive valid k-mers without any candidate (probability
//      ~e^-c per window) is re-scanned exactly with the textbook sliding window; its minimizers are
//      attached behind the candidate that precedes the gap so output order stays position order.
//
// k-mers containing a non-ACGT base are skipped and windows run over consecutive VALID k-mers (they span
// N gaps), exactly like btllib: all distances above are measured in valid-k-mer index space
// (vbase[strip] + lord).
#pragma once
#include <stdint.h>
#include "nthash.cuh"

namespace ntl {

// A candidate k-mer. 16 bytes, written with one 128-bit store.
struct alignas(16) Cand {
    uint64_t h0;     // canonical hash (orders the window)
    uint32_t posf;   // k-mer start position | (forward-strand flag << 31)
    uint32_t lord;   // number of valid k-mers of the same strip before this one
};

enum : uint32_t { NONE32 = 0xFFFFFFFFu, POS_MASK = 0x7FFFFFFFu, FWD_BIT = 0x80000000u };

// 8 consecutive nibbles (bases g .. g+7) of the packed sequence as one 32-bit word
NTL_HD uint32_t fetch8(const uint32_t* __restrict__ packed, uint64_t g) {
    const uint64_t wi = g >> 3;
    const uint32_t sh = (uint32_t)(g & 7) * 4;
    const uint32_t lo = packed[wi], hi = packed[wi + 1];
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}
NTL_HD uint32_t fetch1(const uint32_t* __restrict__ packed, uint64_t g) {
    return (packed[g >> 3] >> ((uint32_t)(g & 7) * 4)) & 7u;
}

// ---------------------------------------------------------------------------------------------------
// DENSE pass over one strip: k-mer positions [p0, p0+n) of a sequence whose first base has global base
// index gseq. Calls emit(h0, pos, fwd, lord) for every valid k-mer with (h0 >> 32) < tau_hi, in position
// order, and returns the number of valid k-mers of the strip.
//   tbl      roll table; entry e lives at tbl[e * tstride] (per-lane copies in shared memory on the device)
// The hash state starts from zero and is rolled over k-1 lead-in bases with a virtual "zero" base leaving
// (code 4), which yields exactly the ntHash initial value (see nthash.cuh).
// ---------------------------------------------------------------------------------------------------
// With ALL = true every valid k-mer is emitted regardless of the threshold (used by the gap re-scan).
template <bool ALL = false, class Emit>
NTL_HD uint32_t process_strip(const uint32_t* __restrict__ packed, uint64_t gseq, uint32_t p0, uint32_t n,
                              uint32_t k, const RollEntry* tbl, uint32_t tstride, uint32_t tau_hi, Emit& emit) {
    const uint64_t g0 = gseq + p0;              // first base consumed
    const int32_t T = (int32_t)(k - 1 + n);     // roll steps; step s completes k-mer p0 + s - (k-1)
    const int32_t lead = (int32_t)k - 1;
    uint64_t fh = 0, rh = 0;
    int32_t last_bad = -(1 << 30);              // step index of the most recent invalid base
    uint32_t nv = 0;
    for (int32_t t = 0; t < T; t += 8) {
        const uint32_t wi = fetch8(packed, g0 + (uint32_t)t);
        uint32_t wo;
        const int32_t o = t - (int32_t)k;       // base index (relative to g0) leaving at step t
        if (o >= 0) wo = fetch8(packed, g0 + (uint32_t)o);
        else if (o <= -8) wo = 0x44444444u;
        else {                                  // -8 < o < 0: the first -o steps still push out virtual bases
            const uint32_t sh = 4u * (uint32_t)(-o);
            wo = (fetch8(packed, g0) << sh) | (0x44444444u & ((1u << sh) - 1u));
        }
        const bool fast = ((wi & 0x44444444u) == 0u) && (t - last_bad >= (int32_t)k);
        if (fast) {
            const int32_t first_out = t > lead ? t : lead;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t e = (((wi >> (4 * j)) & 7u) << 3) | ((wo >> (4 * j)) & 7u);
                const RollEntry re = tbl[e * tstride];
                fh = srol1(fh) ^ re.f;
                rh = sror1(rh ^ re.r);
                const uint64_t h0 = fh + rh;
                if (ALL || (uint32_t)(h0 >> 32) < tau_hi) {
                    const int32_t s = t + j;
                    if (s >= lead && s < T)
                        emit(h0, p0 + (uint32_t)(s - lead), fh <= rh, nv + (uint32_t)(s - first_out));
                }
            }
            const int32_t last = (t + 8 < T ? t + 8 : T);
            if (last > first_out) nv += (uint32_t)(last - first_out);
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int32_t s = t + j;
                const uint32_t cin = (wi >> (4 * j)) & 7u;
                const uint32_t e = (cin << 3) | ((wo >> (4 * j)) & 7u);
                const RollEntry re = tbl[e * tstride];
                fh = srol1(fh) ^ re.f;
                rh = sror1(rh ^ re.r);
                if (cin >= CODE_INVALID) last_bad = s;
                if (s >= lead && s < T && s - last_bad >= (int32_t)k) {
                    const uint64_t h0 = fh + rh;
                    if (ALL || (uint32_t)(h0 >> 32) < tau_hi) emit(h0, p0 + (uint32_t)(s - lead), fh <= rh, nv);
                    nv++;
                }
            }
        }
    }
    return nv;
}

// ---------------------------------------------------------------------------------------------------
// Walks the valid k-mers of one sequence base by base (slow path: gap re-scan only).
// ---------------------------------------------------------------------------------------------------
struct KmerWalker {
    const uint32_t* packed;
    const RollEntry* tbl;     // stride-1 table
    uint64_t gseq;
    uint32_t L, k;
    uint32_t pos;
    uint64_t fh, rh;
    // two cached 8-base windows of the packed sequence (entering / leaving base streams are sequential)
    uint32_t win, wout, in_base, out_base;

    NTL_HD void reset_cache() { in_base = NONE32; out_base = NONE32; win = 0; wout = 0; }
    NTL_HD uint32_t code(uint32_t i) const { return fetch1(packed, gseq + i); }
    NTL_HD uint32_t code_in(uint32_t i) {
        if (in_base == NONE32 || i - in_base >= 8u) { in_base = i; win = fetch8(packed, gseq + i); }
        return (win >> (4u * (i - in_base))) & 7u;
    }
    NTL_HD uint32_t code_out(uint32_t i) {
        if (out_base == NONE32 || i - out_base >= 8u) { out_base = i; wout = fetch8(packed, gseq + i); }
        return (wout >> (4u * (i - out_base))) & 7u;
    }

    // first valid k-mer starting at or after p; false if there is none
    NTL_HD bool seek(uint32_t p) {
        for (;;) {
            if ((uint64_t)p + k > L) return false;
            int64_t bad = -1;
            for (uint32_t j = k; j-- > 0;)
                if (code(p + j) >= CODE_INVALID) { bad = (int64_t)p + j; break; }
            if (bad >= 0) { p = (uint32_t)bad + 1; continue; }
            fh = 0; rh = 0;
            for (uint32_t j = 0; j < k; j++) {
                const RollEntry re = tbl[(code_in(p + j) << 3) | CODE_INVALID];
                fh = srol1(fh) ^ re.f;
                rh = sror1(rh ^ re.r);
            }
            pos = p;
            return true;
        }
    }
    // next valid k-mer after the current one
    NTL_HD bool next() {
        if ((uint64_t)pos + 1 + k > L) return false;
        const uint32_t cin = code_in(pos + k);
        if (cin >= CODE_INVALID) return seek(pos + k + 1);
        const RollEntry re = tbl[(cin << 3) | code_out(pos)];
        fh = srol1(fh) ^ re.f;
        rh = sror1(rh ^ re.r);
        pos++;
        return true;
    }
    NTL_HD uint64_t h0() const { return fh + rh; }
    NTL_HD bool fwd() const { return fh <= rh; }
};

// Exact sliding-window minimizers (rightmost argmin, btllib calc_minimizer semantics) over the valid k-mers
// with start position in [start_pos, end_pos) of one sequence; only windows lying entirely inside that range
// count. No buffer: when the current minimum leaves the window the window is re-hashed from its left end
// (tracked by a second walker). Calls out(h0, pos, fwd) in increasing position order; returns the count.
template <class Out>
NTL_HD uint32_t gap_scan(const uint32_t* packed, const RollEntry* tbl, uint64_t gseq, uint32_t L, uint32_t k,
                         uint32_t w, uint32_t start_pos, uint32_t end_pos, Out& out) {
    KmerWalker head;
    head.packed = packed; head.tbl = tbl; head.gseq = gseq; head.L = L; head.k = k; head.pos = 0; head.fh = 0; head.rh = 0;
    head.reset_cache();
    KmerWalker tail = head;
    if (!head.seek(start_pos) || head.pos >= end_pos) return 0;
    tail = head;                                  // tail = left end of the current window
    uint64_t t = 0;                               // index (within the range) of head's k-mer
    uint64_t cur_h = 0, cur_idx = 0; uint32_t cur_pos = 0; bool cur_fwd = false, have_cur = false;
    int64_t last_emitted = -1;
    uint32_t emitted = 0;
    for (;;) {
        if (t + 1 >= w) {
            const uint64_t left = t + 1 - w;      // tail sits on index `left`
            if (!have_cur || cur_idx < left) {
                KmerWalker sc = tail;             // rescan the whole window, <= keeps the rightmost minimum
                cur_h = sc.h0(); cur_idx = left; cur_pos = sc.pos; cur_fwd = sc.fwd(); have_cur = true;
                for (uint64_t q = left + 1; q <= t; q++) {
                    sc.next();
                    if (sc.h0() <= cur_h) { cur_h = sc.h0(); cur_idx = q; cur_pos = sc.pos; cur_fwd = sc.fwd(); }
                }
            } else if (head.h0() <= cur_h) {
                cur_h = head.h0(); cur_idx = t; cur_pos = head.pos; cur_fwd = head.fwd();
            }
            if ((int64_t)cur_pos > last_emitted && cur_h != 0xFFFFFFFFFFFFFFFFULL) {
                last_emitted = (int64_t)cur_pos;
                out(cur_h, cur_pos, cur_fwd);
                emitted++;
            }
            tail.next();                          // window of the next step starts one valid k-mer later
        }
        if (!head.next() || head.pos >= end_pos) break;
        t++;
    }
    return emitted;
}

// ---------------------------------------------------------------------------------------------------
// SPARSE pass helpers
// ---------------------------------------------------------------------------------------------------
struct CandView {
    const Cand* cands;        // [nstrips * cap] slot area followed by the overflow pool
    const uint32_t* cnt;      // [nstrips] number of candidates of a strip (may exceed cap -> lives in the pool)
    const uint32_t* ovf_off;  // [nstrips] pool offset of an overflowed strip
    const uint32_t* vbase;    // [nstrips + 1] exclusive prefix sum of valid k-mers per strip
    uint32_t cap;
    uint64_t pool_base;       // = nstrips * cap
};
NTL_HD uint64_t cand_gid(const CandView& v, uint32_t s, uint32_t j) {
    return v.cnt[s] <= v.cap ? (uint64_t)s * v.cap + j : v.pool_base + v.ovf_off[s] + j;
}

// A candidate-free stretch that needs an exact re-scan; chained per strip through `next`.
struct GapRec {
    uint32_t seq;
    uint32_t start_pos, end_pos;   // k-mer start positions [start_pos, end_pos)
    uint32_t strip;                // attach point: emitted right after candidate (strip, j); j = NONE32 -> before
    uint32_t j;                    //   all candidates of `strip`
    uint32_t next;                 // next GapRec of the same strip or NONE32
    uint32_t out_off, out_cnt;     // extras written by the gap kernel
    uint32_t max_out;              // reservation size = (#valid k-mers in the stretch) - w + 1
    uint32_t pad;
};

// Decide for candidate (s, j) whether it is a minimizer; also reports the candidate-free stretch that
// follows it (gap_len valid k-mers, ending at position gap_end) so the caller can queue a GapRec.
//   fs, es   first strip / one-past-last strip of the candidate's sequence
//   npos     number of k-mer positions of the sequence (end position when no candidate follows)
struct SelectResult { bool selected; uint32_t gap_len; uint32_t gap_end; };

NTL_HD SelectResult select_candidate(const CandView& v, uint32_t s, uint32_t j, uint32_t fs, uint32_t es,
                                     uint32_t w, uint32_t npos) {
    const Cand ci = v.cands[cand_gid(v, s, j)];
    const uint64_t val = ci.h0;
    const int64_t idx0 = v.vbase[fs];
    const int64_t n = (int64_t)v.vbase[es] - idx0;            // valid k-mers of the sequence
    const int64_t idx = (int64_t)v.vbase[s] + ci.lord;
    const int64_t rel = idx - idx0;
    const int64_t W1 = (int64_t)w - 1;

    // ---- left: A = number of consecutive valid k-mers before i with value >= val (capped at w-1)
    int64_t A;
    {
        uint32_t ss = s, jj = j;
        for (;;) {
            if (jj == 0) {
                bool found = false, capped = false;
                while (ss > fs) {
                    if (idx - (int64_t)v.vbase[ss] + 1 >= (int64_t)w) { capped = true; break; }
                    ss--;
                    if (v.cnt[ss]) { jj = v.cnt[ss]; found = true; break; }
                }
                if (capped) { A = W1; break; }
                if (!found) { A = rel; break; }                // reached the start of the sequence
            }
            jj--;
            const Cand c = v.cands[cand_gid(v, ss, jj)];
            const int64_t d = idx - ((int64_t)v.vbase[ss] + c.lord);
            if (d >= (int64_t)w) { A = W1; break; }
            if (c.h0 < val) { A = d - 1; break; }
        }
    }
    // ---- right: B = number of consecutive valid k-mers after i with value > val (capped at w-1);
    //      the first neighbour also bounds the candidate-free stretch after i
    int64_t B;
    SelectResult res; res.gap_len = 0; res.gap_end = npos;
    {
        uint32_t ss = s, jj = j;
        bool first = true;
        for (;;) {
            jj++;
            if (jj >= v.cnt[ss]) {
                bool found = false, capped = false;
                while (ss + 1 < es) {
                    if (!first && (int64_t)v.vbase[ss + 1] - idx >= (int64_t)w) { capped = true; break; }
                    ss++;
                    if (v.cnt[ss]) { jj = 0; found = true; break; }
                }
                if (capped) { B = W1; break; }
                if (!found) {                                  // reached the end of the sequence
                    B = idx0 + n - 1 - idx;
                    if (first) { res.gap_len = (uint32_t)(B < 0 ? 0 : B); res.gap_end = npos; }
                    break;
                }
            }
            const Cand c = v.cands[cand_gid(v, ss, jj)];
            const int64_t d = ((int64_t)v.vbase[ss] + c.lord) - idx;
            if (first) { res.gap_len = (uint32_t)(d - 1); res.gap_end = c.posf & POS_MASK; first = false; }
            if (d >= (int64_t)w) { B = W1; break; }
            if (c.h0 <= val) { B = d - 1; break; }
        }
    }
    // ---- is there a window start j with  max(0, rel-(w-1), rel-A) <= j <= min(rel, n-w, rel+B-(w-1)) ?
    int64_t lo = rel - W1; if (lo < 0) lo = 0; if (rel - A > lo) lo = rel - A;
    int64_t hi = rel; if (n - (int64_t)w < hi) hi = n - (int64_t)w; if (rel + B - W1 < hi) hi = rel + B - W1;
    res.selected = lo <= hi;
    return res;
}

// ---------------------------------------------------------------------------------------------------
// One pass over a whole strip: the same decisions as select_candidate for every candidate of strip s, but in
// O(#candidates) with a monotone stack (all-nearest-smaller-values): walk the candidates from (w-1) valid k-mers
// before the strip to (w-1) after it; when a candidate x arrives every stacked candidate with value >= x.h0 is
// popped -- x is its first smaller-or-equal neighbour on the right, and the entry below it on the stack is its
// first strictly smaller neighbour on the left. Candidates outside the strip only serve as context.
//   sel        flags indexed by candidate id (cand_gid)
//   gap(j, gap_len, gap_end)   called for every own candidate j followed by >= w candidate-free valid k-mers
// Returns the number of selected candidates, or NONE32 when the stack would overflow (caller falls back to
// select_candidate; only degenerate hash sequences nest deeper than SEL_STACK).
// ---------------------------------------------------------------------------------------------------
enum : uint32_t { SEL_STACK = 16 };

// stack storage: entry d of this thread lives at index d * stride (shared memory, one column per thread, on the
// device; a plain local array with stride 1 on the host)
struct SelStack { uint64_t* h; uint32_t* i; uint32_t* j; uint32_t stride; };

template <class GapFn>
NTL_HD uint32_t select_strip(const CandView& v, uint32_t s, uint32_t fs, uint32_t es, uint32_t w, uint32_t npos,
                             uint8_t* sel, GapFn& gap, const SelStack& S) {
    const uint32_t c = v.cnt[s];
    if (c == 0) return 0;
    const int64_t idx0 = v.vbase[fs];
    const int64_t n = (int64_t)v.vbase[es] - idx0;
    const int64_t W1 = (int64_t)w - 1;
    const uint64_t own_base = cand_gid(v, s, 0);
    const int64_t i_first = (int64_t)v.vbase[s] + v.cands[own_base].lord;
    const int64_t i_last = (int64_t)v.vbase[s] + v.cands[own_base + c - 1].lord;

    // ---- candidate-free stretches of >= w valid k-mers after an own candidate (independent of the stack walk, so
    //      a stack overflow below never loses or duplicates a gap)
    {
        int64_t prev = i_first;
        for (uint32_t j = 1; j < c; j++) {
            const Cand y = v.cands[own_base + j];
            const int64_t yi = (int64_t)v.vbase[s] + y.lord;
            if (yi - prev - 1 >= (int64_t)w) gap(j - 1, (uint32_t)(yi - prev - 1), y.posf & POS_MASK);
            prev = yi;
        }
        uint32_t ns = s;
        bool found = false;
        while (ns + 1 < es) { ns++; if (v.cnt[ns]) { found = true; break; } }
        if (found) {
            const Cand y = v.cands[cand_gid(v, ns, 0)];
            const int64_t g = ((int64_t)v.vbase[ns] + y.lord) - i_last - 1;
            if (g >= (int64_t)w) gap(c - 1, (uint32_t)g, y.posf & POS_MASK);
        } else {
            const int64_t g = idx0 + n - 1 - i_last;
            if (g >= (int64_t)w) gap(c - 1, (uint32_t)g, npos);
        }
    }

    // ---- where the left context starts: first candidate within w-1 valid k-mers of the first own candidate
    uint32_t cs = s, cj = 0;
    bool left_is_seq_start = false;
    for (;;) {
        if (cj == 0) {
            bool found = false, far = false;
            uint32_t ss = cs;
            while (ss > fs) {
                if (i_first - (int64_t)v.vbase[ss] + 1 >= (int64_t)w) { far = true; break; }
                ss--;
                if (v.cnt[ss]) { found = true; break; }
            }
            if (far) break;
            if (!found) { left_is_seq_start = true; break; }
            const Cand p = v.cands[cand_gid(v, ss, v.cnt[ss] - 1)];
            if (i_first - ((int64_t)v.vbase[ss] + p.lord) >= (int64_t)w) break;
            cs = ss; cj = v.cnt[ss] - 1;
        } else {
            const Cand p = v.cands[cand_gid(v, cs, cj - 1)];
            if (i_first - ((int64_t)v.vbase[cs] + p.lord) >= (int64_t)w) break;
            cj--;
        }
    }

    uint64_t* const st_h = S.h;        // value
    uint32_t* const st_i = S.i;        // valid-k-mer index (fits 32 bits: a device batch is < 2^32 bases)
    uint32_t* const st_j = S.j;        // own candidate number or NONE32 for context
    const uint32_t sd = S.stride;
    uint32_t depth = 0, nsel = 0;

    // decision for own candidate j once both neighbours are known (B_known: a right neighbour <= value was found)
    auto decide = [&](uint32_t j, int64_t idx, bool has_psv, int64_t psv_idx, bool has_nse, int64_t nse_idx,
                      bool right_is_seq_end) {
        const int64_t rel = idx - idx0;
        const int64_t A = has_psv ? idx - psv_idx - 1 : (left_is_seq_start ? rel : W1);
        const int64_t B = has_nse ? nse_idx - idx - 1 : (right_is_seq_end ? idx0 + n - 1 - idx : W1);
        int64_t lo = rel - W1; if (lo < 0) lo = 0; if (rel - A > lo) lo = rel - A;
        int64_t hi = rel; if (n - (int64_t)w < hi) hi = n - (int64_t)w; if (rel + B - W1 < hi) hi = rel + B - W1;
        const bool selected = lo <= hi;
        sel[own_base + j] = selected ? 1 : 0;
        nsel += selected ? 1u : 0u;
    };

    // ---- forward walk
    bool right_is_seq_end = false;
    for (;;) {
        const Cand x = v.cands[cand_gid(v, cs, cj)];
        const int64_t xi = (int64_t)v.vbase[cs] + x.lord;
        const bool own = (cs == s);
        if (!own && cs > s && xi - i_last > W1) break;          // beyond every own candidate's reach
        while (depth && st_h[(depth - 1) * sd] >= x.h0) {
            depth--;
            if (st_j[depth * sd] != NONE32)
                decide(st_j[depth * sd], (int64_t)st_i[depth * sd], depth > 0, depth > 0 ? (int64_t)st_i[(depth - 1) * sd] : 0,
                       true, xi, false);
        }
        if (depth == SEL_STACK) return NONE32;
        st_h[depth * sd] = x.h0; st_i[depth * sd] = (uint32_t)xi; st_j[depth * sd] = own ? cj : NONE32; depth++;
        // next candidate of the sequence
        uint32_t ns = cs, nj = cj + 1;
        bool found = nj < v.cnt[ns];
        if (!found) {
            for (;;) {
                if (ns + 1 >= es) { right_is_seq_end = true; break; }           // no candidate left in the sequence
                ns++;
                if (v.cnt[ns]) { nj = 0; found = true; break; }
                if (cs > s && (int64_t)v.vbase[ns + 1] - i_last > W1) break;    // the rest is out of reach
            }
        }
        if (!found) break;
        cs = ns; cj = nj;
    }
    // whatever is still stacked has no smaller-or-equal neighbour within reach on the right
    while (depth) {
        depth--;
        if (st_j[depth * sd] != NONE32)
            decide(st_j[depth * sd], (int64_t)st_i[depth * sd], depth > 0, depth > 0 ? (int64_t)st_i[(depth - 1) * sd] : 0,
                   false, 0, right_is_seq_end);
    }
    return nsel;
}

// candidate threshold on the high word of h0: about c/w of the hash space (everything when w <= c)
inline uint32_t candidate_threshold(uint32_t w, double c) {
    double f = c / (double)w;
    if (f >= 1.0) return 0xFFFFFFFFu;
    double t = f * 4294967296.0;
    return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}

}  // namespace ntl
