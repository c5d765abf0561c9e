// emu.cpp -- TEST INFRASTRUCTURE ONLY. Host emulation of the CUDA kernels' per-thread logic.
//
// The container that builds this repo has no GPU. To debug the kernel logic (ntlink_b200/csrc/sketch_logic.cuh,
// map_logic.cuh, lift_logic.cuh) against the oracle without one, this file compiles those SAME headers with g++ and drives them
// with plain loops that stand in for the grid ("for every strip", "for every read") and for the scans.
// It is built into tests/emu/_build/libntl_emu.so, loaded only by tests/test_emu_*.py, never by the product
// (ntlink_b200/ has no CPU path: libntlink_b200.so fails to initialise without a CUDA device).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <unordered_map>
#include <vector>

#include "../../ntlink_b200/csrc/lift_logic.cuh"
#include "../../ntlink_b200/csrc/map_logic.cuh"
#include "../../ntlink_b200/csrc/nthash.cuh"
#include "../../ntlink_b200/csrc/sketch_logic.cuh"

using namespace ntl;

namespace {

uint32_t base_code(unsigned char c) {
    switch (c & 0xDF) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 4; }
}
uint32_t seq_npos(uint64_t L, uint32_t k, uint32_t w) {
    if (L < k) return 0;
    uint64_t np = L - k + 1;
    return np < w ? 0u : (uint32_t)np;
}

struct SlotEmit {
    Cand* dst; uint32_t cap, count;
    void operator()(uint64_t h0, uint32_t pos, bool fwd, uint32_t lord) {
        if (count < cap) { Cand c; c.h0 = h0; c.posf = pos | (fwd ? FWD_BIT : 0u); c.lord = lord; dst[count] = c; }
        count++;
    }
};
struct VecEmit {
    std::vector<Cand>* v;
    void operator()(uint64_t h0, uint32_t pos, bool fwd) { Cand c; c.h0 = h0; c.posf = pos | (fwd ? FWD_BIT : 0u); c.lord = 0; v->push_back(c); }
};

struct Sketch { std::vector<uint64_t> hash; std::vector<uint32_t> posf; std::vector<uint64_t> off; uint64_t ncand = 0, ngaps = 0, novf = 0, nstackovf = 0; };

void emu_sketch_impl(const unsigned char* seq, const uint64_t* off, uint32_t nseq, uint32_t k, uint32_t w, uint32_t S,
                     double cc, uint32_t cap_override, Sketch& out) {
    const uint64_t total = off[nseq];
    std::vector<uint32_t> packed(total / 8 + 8, 0x44444444u);
    for (uint64_t i = 0; i < total; i++) {
        packed[i >> 3] &= ~(0xFu << ((i & 7) * 4));
        packed[i >> 3] |= base_code(seq[i]) << ((i & 7) * 4);
    }
    RollEntry tbl[ROLL_TABLE_ENTRIES];
    build_roll_table(k, tbl);
    const uint32_t tau = candidate_threshold(w, cc);
    const uint64_t mult = second_hash_multiplier(k);
    double mu = (double)S * cc / (double)w; if (mu > S) mu = S;
    uint32_t cap = (uint32_t)(mu + 6.0 * std::sqrt(mu) + 8.0); if (cap > S) cap = S; cap = (cap + 1) & ~1u;
    if (cap_override) cap = cap_override;
    std::vector<uint32_t> strip_off(nseq + 1, 0);
    for (uint32_t q = 0; q < nseq; q++) strip_off[q + 1] = strip_off[q] + (seq_npos(off[q + 1] - off[q], k, w) + S - 1) / S;
    const uint32_t nstrips = strip_off[nseq];
    std::vector<Cand> cands((size_t)nstrips * cap);
    std::vector<uint32_t> cnt(nstrips), nv(nstrips), ovf_off(nstrips, 0), vbase(nstrips + 1, 0);
    std::vector<uint8_t> has_cand(nseq, 0);
    const uint64_t pool_base = (uint64_t)nstrips * cap;
    // k_dense + k_overflow
    for (uint32_t q = 0; q < nseq; q++) {
        const uint32_t np = seq_npos(off[q + 1] - off[q], k, w);
        for (uint32_t s = strip_off[q]; s < strip_off[q + 1]; s++) {
            const uint32_t p0 = (s - strip_off[q]) * S, n = std::min(S, np - p0);
            SlotEmit em{cands.data() + (size_t)s * cap, cap, 0};
            nv[s] = process_strip(packed.data(), off[q], p0, n, k, tbl, 1, tau, em);
            cnt[s] = em.count;
            if (em.count) has_cand[q] = 1;
        }
    }
    for (uint32_t q = 0; q < nseq; q++) {
        const uint32_t np = seq_npos(off[q + 1] - off[q], k, w);
        for (uint32_t s = strip_off[q]; s < strip_off[q + 1]; s++) {
            if (cnt[s] <= cap) continue;
            out.novf++;
            ovf_off[s] = (uint32_t)(cands.size() - pool_base);
            const size_t old = cands.size();
            cands.resize(old + cnt[s]);
            const uint32_t p0 = (s - strip_off[q]) * S, n = std::min(S, np - p0);
            SlotEmit em{cands.data() + old, cnt[s], 0};
            process_strip(packed.data(), off[q], p0, n, k, tbl, 1, tau, em);
        }
    }
    for (uint32_t s = 0; s < nstrips; s++) vbase[s + 1] = vbase[s] + nv[s];
    CandView V{cands.data(), cnt.data(), ovf_off.data(), vbase.data(), cap, pool_base};
    // k_select + k_seq_gaps
    std::vector<uint8_t> sel(cands.size(), 0);
    std::vector<GapRec> gaps;
    std::vector<uint32_t> gap_head(nstrips + 1, NONE32);
    auto queue_gap = [&](uint32_t q, uint32_t sp, uint32_t ep, uint32_t strip, uint32_t j, uint32_t nvalid) {
        GapRec g; g.seq = q; g.start_pos = sp; g.end_pos = ep; g.strip = strip; g.j = j; g.out_off = 0; g.out_cnt = 0;
        g.max_out = nvalid - w + 1; g.pad = 0; g.next = gap_head[strip];
        gap_head[strip] = (uint32_t)gaps.size();
        gaps.push_back(g);
    };
    for (uint32_t q = 0; q < nseq; q++) {
        const uint32_t fs = strip_off[q], es = strip_off[q + 1];
        const uint32_t np = seq_npos(off[q + 1] - off[q], k, w);
        for (uint32_t s = fs; s < es; s++) {
            if (!cnt[s]) continue;
            // k_select: one monotone-stack pass per strip; cross-checked here against the per-candidate scan
            struct GapCollect {
                std::vector<uint32_t> j, len, end;
                void operator()(uint32_t jj, uint32_t l, uint32_t e) { j.push_back(jj); len.push_back(l); end.push_back(e); }
            } gc;
            uint64_t stk_h[SEL_STACK]; uint32_t stk_i[SEL_STACK], stk_j[SEL_STACK];
            const SelStack stk{stk_h, stk_i, stk_j, 1};
            const uint32_t nsel = select_strip(V, s, fs, es, w, np, sel.data(), gc, stk);
            uint32_t nsel_ref = 0, ngap_ref = 0;
            for (uint32_t j = 0; j < cnt[s]; j++) {
                SelectResult r = select_candidate(V, s, j, fs, es, w, np);
                const uint64_t gid = cand_gid(V, s, j);
                if (nsel != NONE32 && sel[gid] != (r.selected ? 1 : 0)) abort();
                sel[gid] = r.selected;
                nsel_ref += r.selected;
                out.ncand++;
                if (r.gap_len >= w) {
                    if (ngap_ref >= gc.j.size() || gc.j[ngap_ref] != j || gc.len[ngap_ref] != r.gap_len || gc.end[ngap_ref] != r.gap_end) abort();
                    ngap_ref++;
                    queue_gap(q, (cands[gid].posf & POS_MASK) + 1, r.gap_end, s, j, r.gap_len);
                }
            }
            if (ngap_ref != gc.j.size()) abort();
            if (nsel != NONE32 && nsel != nsel_ref) abort();
            if (nsel == NONE32) out.nstackovf++;
        }
        if (fs == es) continue;
        const uint32_t nvalid = vbase[es] - vbase[fs];
        if (nvalid < w) continue;
        if (!has_cand[q]) { queue_gap(q, 0, np, fs, NONE32, nvalid); continue; }
        uint32_t fc = fs;
        while (fc < es && cnt[fc] == 0) fc++;
        const Cand c = cands[cand_gid(V, fc, 0)];
        const uint32_t lead = vbase[fc] + c.lord - vbase[fs];
        if (lead >= w) queue_gap(q, 0, c.posf & POS_MASK, fs, NONE32, lead);
    }
    // k_gap
    std::vector<std::vector<Cand>> extras(gaps.size());
    for (size_t g = 0; g < gaps.size(); g++) {
        VecEmit em{&extras[g]};
        const GapRec& G = gaps[g];
        gap_scan(packed.data(), tbl, off[G.seq], (uint32_t)(off[G.seq + 1] - off[G.seq]), k, w, G.start_pos, G.end_pos, em);
        if (extras[g].size() > G.max_out) abort();   // reservation bound must hold
    }
    out.ngaps = gaps.size();
    // k_emit
    out.off.assign(nseq + 1, 0);
    for (uint32_t q = 0; q < nseq; q++) {
        for (uint32_t s = strip_off[q]; s < strip_off[q + 1]; s++) {
            auto put_gap = [&](uint32_t j) {
                for (uint32_t g = gap_head[s]; g != NONE32; g = gaps[g].next)
                    if (gaps[g].j == j)
                        for (const Cand& e : extras[g]) { out.hash.push_back(second_hash(e.h0, mult)); out.posf.push_back(e.posf); }
            };
            put_gap(NONE32);
            for (uint32_t j = 0; j < cnt[s]; j++) {
                const uint64_t gid = cand_gid(V, s, j);
                if (sel[gid]) { out.hash.push_back(second_hash(cands[gid].h0, mult)); out.posf.push_back(cands[gid].posf); }
                put_gap(j);
            }
        }
        out.off[q + 1] = out.hash.size();
    }
}

}  // namespace

extern "C" {

// returns the number of minimizers (or -1 if cap too small); stats[0..2] = candidates, gaps, overflowed strips
int64_t emu_sketch(const unsigned char* seq, const uint64_t* off, uint32_t nseq, uint32_t k, uint32_t w, uint32_t S,
                   double cc, uint32_t cap_override, uint64_t* hash, uint32_t* posf, uint64_t cap, uint64_t* mx_off,
                   uint64_t* stats) {
    Sketch sk;
    emu_sketch_impl(seq, off, nseq, k, w, S, cc, cap_override, sk);
    if (stats) { stats[0] = sk.ncand; stats[1] = sk.ngaps; stats[2] = sk.novf; stats[3] = sk.nstackovf; }
    for (uint32_t q = 0; q <= nseq; q++) mx_off[q] = sk.off[q];
    if (sk.hash.size() > cap) return -1;
    if (!sk.hash.empty()) {
        memcpy(hash, sk.hash.data(), sk.hash.size() * 8);
        memcpy(posf, sk.posf.data(), sk.posf.size() * 4);
    }
    return (int64_t)sk.hash.size();
}

// Mapping emulation. Target / read sketches come in as arrays; results mirror ntl_map_out (holey layout).
// runs/hits: capacity = number of read minimizers. events: capacity ev_cap. Returns the number of events or -1.
int64_t emu_map(const uint64_t* t_hash, const uint32_t* t_ctg, const uint32_t* t_posf, uint64_t t_n,
                const uint32_t* ctg_len, const uint32_t* name_rank, uint32_t ncontig,
                const uint64_t* r_hash, const uint32_t* r_posf, const uint64_t* r_off, const uint32_t* r_len, uint32_t nreads,
                int32_t k, int32_t z, int32_t f, double x, int32_t sensitive, int32_t repeat_filter,
                uint32_t* hit_off /* nreads+1 */, uint32_t* nruns, Run* runs, Hit* hits,
                uint32_t* ev_off /* nreads+1 */, Event* events, uint64_t ev_cap) {
    (void)ncontig;
    // index with the same open-addressing layout and duplicate rule as k_index_insert / k_index_finalize
    uint64_t slots = 1024;
    while (slots < 2 * t_n) slots <<= 1;
    std::vector<IdxEntry> table(slots);
    for (auto& e : table) { e.key = NTL_EMPTY_KEY; e.ctg = DUP_CTG; e.posf = 0xFFFFFFFFu; }
    std::vector<uint8_t> dup(slots, 0);
    IdxSpecial special{0, 0, 0, 0};
    for (uint64_t i = 0; i < t_n; i++) {
        const uint64_t key = t_hash[i];
        if (key == NTL_EMPTY_KEY) { if (special.count++ == 0) { special.ctg = t_ctg[i]; special.posf = t_posf[i]; } continue; }
        uint64_t s = idx_slot(key, slots - 1);
        for (;;) {
            if (table[s].key == NTL_EMPTY_KEY) { table[s].key = key; table[s].ctg = t_ctg[i]; table[s].posf = t_posf[i]; break; }
            if (table[s].key == key) { dup[s] = 1; break; }
            s = (s + 1) & (slots - 1);
        }
    }
    for (uint64_t s = 0; s < slots; s++) if (table[s].key != NTL_EMPTY_KEY && dup[s]) table[s].ctg = DUP_CTG;
    IndexView ix{table.data(), slots - 1, &special};
    MapParams P; P.k = k; P.z = z; P.f = f; P.x = x; P.x_is_zero = (x == 0.0); P.sensitive = sensitive; P.repeat_filter = repeat_filter;
    // lookup + ordered compaction
    uint32_t nh = 0;
    for (uint32_t r = 0; r < nreads; r++) {
        hit_off[r] = nh;
        for (uint64_t i = r_off[r]; i < r_off[r + 1]; i++) {
            uint32_t ctg, cposf;
            if (index_lookup(ix, r_hash[i], ctg, cposf)) { hits[nh].ctg = ctg; hits[nh].cposf = cposf; hits[nh].rposf = r_posf[i]; nh++; }
        }
    }
    hit_off[nreads] = nh;
    std::vector<uint8_t> mark(nh + 1);
    uint64_t ne = 0;
    for (uint32_t r = 0; r < nreads; r++) {
        const uint32_t o = hit_off[r], n = hit_off[r + 1] - o;
        uint32_t nr = 0;
        if (n) nr = chain_read(hits + o, n, runs + o, mark.data() + o, r_len[r], ctg_len, P);
        nruns[r] = nr;
        ev_off[r] = (uint32_t)ne;
        if (nr >= 2) {
            if (ne + max_events(nr, f) > ev_cap) return -1;
            ne += tally_read(hits + o, runs + o, nr, r_len[r], r, ctg_len, name_rank, P, events + ne);
        }
    }
    ev_off[nreads] = (uint32_t)ne;
    return (int64_t)ne;
}

// liftover (k_liftover): one loop iteration per read. Returns the error bits (0 = ok).
uint32_t emu_liftover(const uint32_t* hit_off, const uint32_t* nruns, const Run* runs, const Hit* hits, uint32_t nreads,
                      const AgpRow* agp, uint32_t ncontig, int32_t k, uint32_t* nruns_out, Run* runs_out, Hit* hits_out) {
    uint32_t err = 0;
    const uint32_t n = nreads ? hit_off[nreads] : 0;
    std::vector<uint32_t> tmp_id(n + 1), tmp_kept(n + 1);
    for (uint32_t r = 0; r < nreads; r++) {
        const uint32_t o = hit_off[r], cap = hit_off[r + 1] - o;
        nruns_out[r] = nruns[r] ? lift_read(hits + o, runs + o, nruns[r], cap, agp, ncontig, k, hits_out + o, runs_out + o,
                                            tmp_id.data() + o, tmp_kept.data() + o, &err)
                                : 0;
    }
    return err;
}

// the host-side ordering of the pair table (map_logic.cuh)
void emu_order_first_seen(const uint64_t* keys, uint32_t n, uint32_t* perm) { order_first_seen(keys, n, perm); }

}  // extern "C"
