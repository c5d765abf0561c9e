// map.cu -- target index, per-read lookup, anchor chaining, pair events and the pair tally on the device.
// Replaces the mapping stage of bin/ntlink_pair.py (M1-M7 of SURVEY.md 8a); per-thread logic in map_logic.cuh.
//
// Kernels:
//   k_index_insert / k_index_finalize   open-addressing table built with atomicCAS; a hash seen twice is flagged
//                                       and dropped entirely (bin/ntlink_pair.py:204-209)
//   k_expand_ctg                        contig id per target minimizer (from the per-contig offsets)
//   k_lookup                            one probe (one 32-byte sector) per read minimizer
//   k_compact_hits                      ordered compaction of the hits, per-read hit offsets
//   k_chain                             one thread per read: z filter, noisy filter, runs, subsumption, merge
//   k_events                            one thread per read: contig-pair observations in reference order
//   k_compact_events                    append to the device event log
//   k_tally_*                           edge table with atomics (n, anchor, first-seen), gap lists in read order
#include <algorithm>
#include <thread>

#include "common.cuh"
#include "lift_logic.cuh"

namespace ntl {

namespace {

inline uint32_t div_up(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------- index
__global__ void k_index_insert(const uint64_t* __restrict__ hash, const uint32_t* __restrict__ ctg,
                               const uint32_t* __restrict__ posf, uint64_t n, const uint32_t* __restrict__ n_src,
                               IdxEntry* __restrict__ table, uint64_t mask, uint8_t* __restrict__ dupflag,
                               IdxSpecial* __restrict__ special) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n_src && *n_src < n) n = *n_src;          // n is an upper bound after a deferred sketch; the exact count is on the device
    if (i >= n) return;
    const uint64_t key = hash[i];
    if (key == NTL_EMPTY_KEY) {
        if (atomicAdd(&special->count, 1u) == 0) { special->ctg = ctg[i]; special->posf = posf[i]; }
        return;
    }
    uint64_t s = idx_slot(key, mask);
    for (;;) {
        const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(&table[s].key),
                                                  (unsigned long long)NTL_EMPTY_KEY, (unsigned long long)key);
        if (prev == NTL_EMPTY_KEY) {
            *reinterpret_cast<uint2*>(&table[s].ctg) = make_uint2(ctg[i], posf[i]);
            return;
        }
        if (prev == key) { dupflag[s] = 1; return; }
        s = (s + 1) & mask;
    }
}

__global__ void k_index_finalize(IdxEntry* __restrict__ table, uint64_t slots, const uint8_t* __restrict__ dupflag,
                                 unsigned long long* __restrict__ n_unique) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= slots) return;
    if (table[s].key == NTL_EMPTY_KEY) return;
    if (dupflag[s]) table[s].ctg = DUP_CTG;
    else atomicAdd(n_unique, 1ull);
}

__global__ void k_expand_ctg(const uint32_t* __restrict__ mx_off, uint32_t nseq, uint32_t n_mx,
                             uint32_t* __restrict__ ctg) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_mx) return;
    uint32_t lo = 0, hi = nseq;            // mx_off[lo] <= i < mx_off[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (mx_off[mid] <= i) lo = mid; else hi = mid;
    }
    ctg[i] = lo;
}

// ------------------------------------------------------------------------------------------- lookup
// n_bound: host-side size (exact, or an upper bound after a deferred sketch); n_src: device word with the exact count
__global__ void __launch_bounds__(256) k_lookup(const uint64_t* __restrict__ hash, const uint32_t* __restrict__ posf,
                                                uint32_t n_bound, const uint32_t* __restrict__ n_src, IndexView ix,
                                                Hit* __restrict__ tmp, uint32_t* __restrict__ flag, uint32_t* __restrict__ n_dev) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = n_src ? min(*n_src, n_bound) : n_bound;
    if (i == 0) *n_dev = n;
    if (i >= n) return;
    uint32_t ctg, cposf;
    if (index_lookup(ix, hash[i], ctg, cposf)) {
        Hit h; h.ctg = ctg; h.cposf = cposf; h.rposf = posf[i];
        tmp[i] = h; flag[i] = 1;
    } else flag[i] = 0;
}

__global__ void __launch_bounds__(256) k_compact_hits(const Hit* __restrict__ tmp, const uint32_t* __restrict__ flag,
                                                      const uint32_t* __restrict__ pref, const uint32_t* __restrict__ n_dev,
                                                      Hit* __restrict__ hits) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_dev) return;
    if (flag[i]) hits[pref[i]] = tmp[i];
}

__global__ void k_hit_offsets(const uint32_t* __restrict__ mx_off, const uint32_t* __restrict__ pref, uint32_t nreads,
                              uint32_t* __restrict__ hit_off, MapStatus* __restrict__ st, uint32_t* __restrict__ nreads_dev) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) *nreads_dev = nreads;
    if (r > nreads) return;
    const uint32_t v = pref[mx_off[r]];
    hit_off[r] = v;
    if (r == nreads) st->n_hits = v;
}

__global__ void k_read_len(const uint64_t* __restrict__ seq_off, uint32_t nreads, uint32_t* __restrict__ read_len) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nreads) read_len[r] = (uint32_t)(seq_off[r + 1] - seq_off[r]);
}

// ------------------------------------------------------------------------------------------- chain
// One warp per read: the lanes stage the read's hits in shared memory (coalesced), lane 0 runs the sequential chaining
// rules there -- the same chain_read, on shared instead of global memory, where every step of its dependent chains
// costs ~30 cycles instead of an L2 round trip -- and the lanes write the accepted hits and runs back. Reads with more
// than CHAIN_CAP hits run chain_read on global memory.
constexpr int CHAIN_WARPS = 4, CHAIN_CAP = 192;
__global__ void __launch_bounds__(CHAIN_WARPS * 32) k_chain(Hit* __restrict__ hits, Run* __restrict__ runs, uint8_t* __restrict__ mark,
                                               const uint32_t* __restrict__ hit_off, const uint32_t* __restrict__ read_len,
                                               uint32_t nreads, const uint32_t* __restrict__ ctg_len, MapParams P,
                                               uint32_t* __restrict__ nruns, uint32_t* __restrict__ evmax,
                                               MapStatus* __restrict__ st) {
    __shared__ Hit s_hits[CHAIN_WARPS][CHAIN_CAP];
    __shared__ Run s_runs[CHAIN_WARPS][CHAIN_CAP];
    __shared__ uint8_t s_mark[CHAIN_WARPS][CHAIN_CAP];
    const uint32_t wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * CHAIN_WARPS + wid;
    if (r >= nreads) return;
    const uint32_t o = hit_off[r], nh = hit_off[r + 1] - o;
    uint32_t nr = 0;
    if (nh == 0) {
    } else if (nh <= CHAIN_CAP) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(hits + o);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&s_hits[wid][0]);
        for (uint32_t i = lane; i < nh * 3; i += 32) dst[i] = src[i];
        __syncwarp();
        uint32_t m = 0;
        if (lane == 0) {
            nr = chain_read(&s_hits[wid][0], nh, &s_runs[wid][0], &s_mark[wid][0], read_len[r], ctg_len, P);
            for (uint32_t i = 0; i < nr; i++) m += s_runs[wid][i].count;
        }
        nr = __shfl_sync(0xffffffffu, nr, 0);
        m = __shfl_sync(0xffffffffu, m, 0);
        __syncwarp();
        uint32_t* hout = reinterpret_cast<uint32_t*>(hits + o);
        for (uint32_t i = lane; i < m * 3; i += 32) hout[i] = dst[i];
        const uint32_t* rsrc = reinterpret_cast<const uint32_t*>(&s_runs[wid][0]);
        uint32_t* rout = reinterpret_cast<uint32_t*>(runs + o);
        for (uint32_t i = lane; i < nr * 3; i += 32) rout[i] = rsrc[i];
    } else {
        if (lane == 0) nr = chain_read(hits + o, nh, runs + o, mark + o, read_len[r], ctg_len, P);
        nr = __shfl_sync(0xffffffffu, nr, 0);
    }
    if (lane == 0) {
        nruns[r] = nr;
        evmax[r] = max_events(nr, P.f);
        if (nr) atomicAdd(&st->n_runs, nr);
    }
}

__global__ void __launch_bounds__(128) k_events(const Hit* __restrict__ hits, const Run* __restrict__ runs,
                                                const uint32_t* __restrict__ hit_off, const uint32_t* __restrict__ read_len,
                                                const uint32_t* __restrict__ nruns, const uint32_t* __restrict__ ev_off,
                                                uint32_t nreads, uint32_t first_ordinal, const uint32_t* __restrict__ ctg_len,
                                                const uint32_t* __restrict__ name_rank, MapParams P, uint32_t ev_cap,
                                                Event* __restrict__ events, uint32_t* __restrict__ ev_cnt,
                                                MapStatus* __restrict__ st) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreads) return;
    const uint32_t nr = nruns[r];
    uint32_t ne = 0;
    if (nr >= 2) {
        const uint32_t eo = ev_off[r];
        if (max_events(nr, P.f) > MAX_EVENTS_PER_READ) atomicOr(&st->err, MAPERR_READ_EVENTS);
        else if ((uint64_t)eo + max_events(nr, P.f) > ev_cap) atomicOr(&st->err, MAPERR_EVENTS);
        else {
            const uint32_t o = hit_off[r];
            ne = tally_read(hits + o, runs + o, nr, read_len[r], first_ordinal + r, ctg_len, name_rank, P, events + eo);
            for (uint32_t i = 0; i < ne; i++) if (events[eo + i].flags & 8u) atomicOr(&st->err, MAPERR_ASSERT);
        }
    }
    ev_cnt[r] = ne;
}

__global__ void __launch_bounds__(128) k_compact_events(const Event* __restrict__ events, const uint32_t* __restrict__ ev_off,
                                                        const uint32_t* __restrict__ ev_cnt, const uint32_t* __restrict__ ev_pref,
                                                        uint32_t nreads, Event* __restrict__ log, MapStatus* __restrict__ st,
                                                        const CallState* __restrict__ call) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) st->n_events = ev_pref[nreads];
    if (r >= nreads) return;
    if (call) log += call->log_n;                 // sync-free call: the log cursor lives on the device
    const uint32_t n = ev_cnt[r];
    for (uint32_t i = 0; i < n; i++) log[ev_pref[r] + i] = events[ev_off[r] + i];
}

// ---- sync-free call: results of a chunk -> pinned host arrays (device-addressable), then advance the call totals
__global__ void __launch_bounds__(128) k_results_host(const uint32_t* __restrict__ hit_off, const uint32_t* __restrict__ nruns,
                                                      const Run* __restrict__ runs, const Hit* __restrict__ hits,
                                                      const uint32_t* __restrict__ ev_cnt, const uint32_t* __restrict__ ev_pref,
                                                      const Event* __restrict__ log, uint32_t rb, uint32_t nreads,
                                                      HostResults H, CallState* __restrict__ call) {
    // one warp per read: the lanes move the read's runs, hits and events word by word (coalesced writes over PCIe)
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r > nreads) return;
    const uint32_t hits_total = call->hits_total, ev_total = call->ev_total;
    const uint32_t o = hit_off[r], ep = ev_pref[r];
    if (lane == 0) { H.hit_off[rb + r] = hits_total + o; H.ev_off[rb + r] = ev_total + ep; }
    if (r == nreads) return;
    const uint32_t nr = nruns[r], nh = hit_off[r + 1] - o, ne = ev_cnt[r];
    if (lane == 0) { H.nruns[rb + r] = nr; H.ev_cnt[rb + r] = ne; }
    if ((uint64_t)hits_total + o + nh > H.hits_cap || (uint64_t)ev_total + ep + ne > H.ev_cap) {
        if (lane == 0) atomicOr(&call->err, CALLERR_RESULTS);
        return;
    }
    const uint32_t* rs = reinterpret_cast<const uint32_t*>(runs + o);
    uint32_t* rd = reinterpret_cast<uint32_t*>(H.runs + hits_total + o);
    for (uint32_t i = lane; i < nr * 3; i += 32) rd[i] = rs[i];
    const uint32_t* hs = reinterpret_cast<const uint32_t*>(hits + o);
    uint32_t* hd = reinterpret_cast<uint32_t*>(H.hits + hits_total + o);
    for (uint32_t i = lane; i < nh * 3; i += 32) hd[i] = hs[i];
    const uint32_t* es = reinterpret_cast<const uint32_t*>(log + call->log_n + ep);
    uint32_t* ed = reinterpret_cast<uint32_t*>(H.events + ev_total + ep);
    for (uint32_t i = lane; i < ne * 6; i += 32) ed[i] = es[i];
}
__global__ void k_chunk_finish(const MapStatus* __restrict__ st, const uint32_t* __restrict__ n_mx, CallState* __restrict__ call) {
    if (st->err) atomicOr(&call->err, CALLERR_MAP | (st->err << 8));
    call->hits_total += st->n_hits; call->ev_total += st->n_events; call->log_n += st->n_events;
    call->runs_total += st->n_runs; call->mx_total += *n_mx;
}
__global__ void k_call_begin(CallState* __restrict__ call, uint32_t log_n) {
    CallState z = {};
    z.log_n = log_n;
    *call = z;
}
__global__ void k_call_publish(const CallState* __restrict__ call, uint32_t* __restrict__ host) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(call);
    if (threadIdx.x < sizeof(CallState) / 4) host[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}

// ------------------------------------------------------------------------------------------- tally
#define NTL_PAIR_EMPTY 0xFFFFFFFFFFFFFFFFULL
__device__ __forceinline__ uint64_t pair_key(const Event& e) {
    return ((uint64_t)e.src << 33) | ((uint64_t)e.tgt << 2) | (e.flags & 3u);
}
__device__ __forceinline__ uint64_t order_key(const Event& e) { return ((uint64_t)e.read << 24) | (e.ord & 0xFFFFFFu); }

// n is exact, or -- after an import whose counts stayed on the device -- an upper bound with the exact count in *n_src
__global__ void k_tally_insert(const Event* __restrict__ ev, uint64_t n, const uint32_t* __restrict__ n_src,
                               unsigned long long* __restrict__ keys,
                               uint64_t mask, uint32_t* __restrict__ pn, uint32_t* __restrict__ panchor,
                               unsigned long long* __restrict__ pfirst, uint32_t* __restrict__ ev_slot) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n_src && *n_src < n) n = *n_src;
    if (i >= n) return;
    const Event e = ev[i];
    const uint64_t key = pair_key(e);
    uint64_t s = idx_slot(key, mask);
    for (;;) {
        const unsigned long long prev = atomicCAS(&keys[s], (unsigned long long)NTL_PAIR_EMPTY, (unsigned long long)key);
        if (prev == NTL_PAIR_EMPTY || prev == key) break;
        s = (s + 1) & mask;
    }
    atomicAdd(&pn[s], 1u);
    if (e.flags & 4u) atomicAdd(&panchor[s], 1u);
    atomicMin(&pfirst[s], (unsigned long long)order_key(e));
    ev_slot[i] = (uint32_t)s;
}

__global__ void k_tally_scatter(const Event* __restrict__ ev, uint64_t n, const uint32_t* __restrict__ n_src,
                                const uint32_t* __restrict__ ev_slot,
                                const uint32_t* __restrict__ gap_off, uint32_t* __restrict__ cursor,
                                unsigned long long* __restrict__ gkey, int32_t* __restrict__ gval) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n_src && *n_src < n) n = *n_src;
    if (i >= n) return;
    const uint32_t s = ev_slot[i];
    const uint32_t p = gap_off[s] + atomicAdd(&cursor[s], 1u);
    gkey[p] = order_key(ev[i]);
    gval[p] = ev[i].gap;
}

// one thread per pair: insertion sort of its gap list by read order (lists are ~coverage long)
// One warp per pair: rank sort of its gap list by read order (the keys are unique) -- every lane counts, for its
// elements, how many keys are smaller, which is the element's final place. Lists of up to SORT_CAP gaps are staged in
// shared memory; longer ones are ranked on global memory into the event-slot scratch and copied back.
constexpr int SORT_WARPS = 4, SORT_CAP = 512;
__global__ void __launch_bounds__(SORT_WARPS * 32) k_tally_sort(const uint32_t* __restrict__ pn, const uint32_t* __restrict__ gap_off,
                                                               uint32_t slots, unsigned long long* __restrict__ gkey,
                                                               int32_t* __restrict__ gval, uint32_t* __restrict__ nonempty,
                                                               unsigned long long* __restrict__ tmp_key, int32_t* __restrict__ tmp_val) {
    __shared__ unsigned long long s_key[SORT_WARPS][SORT_CAP];
    __shared__ int32_t s_val[SORT_WARPS][SORT_CAP];
    const uint32_t wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t s = blockIdx.x * SORT_WARPS + wid;
    if (s >= slots) return;
    const uint32_t n = pn[s];
    if (lane == 0) nonempty[s] = n ? 1u : 0u;
    if (n < 2) return;
    unsigned long long* k = gkey + gap_off[s];
    int32_t* v = gval + gap_off[s];
    if (n <= SORT_CAP) {
        for (uint32_t i = lane; i < n; i += 32) { s_key[wid][i] = k[i]; s_val[wid][i] = v[i]; }
        __syncwarp();
        for (uint32_t i = lane; i < n; i += 32) {
            const unsigned long long kk = s_key[wid][i];
            uint32_t rank = 0;
            for (uint32_t j = 0; j < n; j++) rank += s_key[wid][j] < kk;
            k[rank] = kk; v[rank] = s_val[wid][i];
        }
    } else {
        unsigned long long* tk = tmp_key + gap_off[s];
        int32_t* tv = tmp_val + gap_off[s];
        for (uint32_t i = lane; i < n; i += 32) {
            const unsigned long long kk = k[i];
            uint32_t rank = 0;
            for (uint32_t j = 0; j < n; j++) rank += k[j] < kk;
            tk[rank] = kk; tv[rank] = v[i];
        }
        __syncwarp();
        for (uint32_t i = lane; i < n; i += 32) { k[i] = tk[i]; v[i] = tv[i]; }
    }
}

__global__ void k_pairs_compact(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ pn,
                                const uint32_t* __restrict__ panchor, const unsigned long long* __restrict__ pfirst,
                                const uint32_t* __restrict__ gap_off, const uint32_t* __restrict__ ppref, uint32_t slots,
                                ntl_pair* __restrict__ out) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= slots || pn[s] == 0) return;
    ntl_pair p;
    const uint64_t key = keys[s];
    p.src = (uint32_t)(key >> 33); p.tgt = (uint32_t)((key >> 2) & 0x7FFFFFFFu); p.flags = (uint32_t)(key & 3u);
    p.n = pn[s]; p.anchor = panchor[s]; p.reserved = 0; p.gap_off = gap_off[s]; p.first_key = pfirst[s];
    out[ppref[s]] = p;
}

__global__ void k_set_u32(uint32_t* p, uint32_t v) { *p = v; }

// liftover: one thread per read (bin/ntlink_liftover_mappings.py:61-124)
__global__ void __launch_bounds__(128) k_liftover(const Hit* __restrict__ hits, const Run* __restrict__ runs,
                                                  const uint32_t* __restrict__ hit_off, const uint32_t* __restrict__ nruns,
                                                  uint32_t nreads, const AgpRow* __restrict__ agp, uint32_t ncontig, int32_t k,
                                                  Hit* __restrict__ hits_out, Run* __restrict__ runs_out,
                                                  uint32_t* __restrict__ nruns_out, uint32_t* __restrict__ tmp_id,
                                                  uint32_t* __restrict__ tmp_kept, MapStatus* __restrict__ st) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreads) return;
    const uint32_t o = hit_off[r], cap = hit_off[r + 1] - o;
    uint32_t err = 0, nout = 0, nh = 0;
    const uint32_t nr = nruns[r];
    if (nr) nout = lift_read(hits + o, runs + o, nr, cap, agp, ncontig, k, hits_out + o, runs_out + o, tmp_id + o, tmp_kept + o, &err);
    nruns_out[r] = nout;
    for (uint32_t i = 0; i < nout; i++) nh += runs_out[o + i].count;
    if (nout) { atomicAdd(&st->n_runs, nout); atomicAdd(&st->n_hits, nh); }
    if (err) atomicOr(&st->err, err << 8);
}

// checkpoint path: event capacity per read from the uploaded run counts
// With read_len_out the reference's substitute read length is computed too: the largest first/last read position
// over the read's runs (bin/ntlink_pair.py:483-487).
__global__ void k_evmax(const uint32_t* __restrict__ nruns, uint32_t nreads, int32_t f, uint32_t* __restrict__ evmax,
                        MapStatus* __restrict__ st, uint32_t* __restrict__ nreads_dev, uint32_t n_hits,
                        const uint32_t* __restrict__ hit_off, const Run* __restrict__ runs, const Hit* __restrict__ hits,
                        uint32_t* __restrict__ read_len_out) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) { *nreads_dev = nreads; st->n_hits = n_hits; }
    if (r >= nreads) return;
    const uint32_t nr = nruns[r];
    evmax[r] = max_events(nr, f);
    if (nr) atomicAdd(&st->n_runs, nr);
    if (read_len_out) {
        const uint32_t o = hit_off[r];
        uint32_t m = 0;
        for (uint32_t i = 0; i < nr; i++) {
            const Run ru = runs[o + i];
            m = max(m, max(pos_of(hits[o + ru.start].rposf), pos_of(hits[o + ru.start + ru.count - 1].rposf)));
        }
        read_len_out[r] = m;
    }
}

// counters -> pinned (UVA-mapped) host memory without using a copy engine
__global__ void k_publish_map(const MapStatus* __restrict__ st, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                              uint32_t* __restrict__ host) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(st);
    if (threadIdx.x < sizeof(MapStatus) / 4) host[threadIdx.x] = src[threadIdx.x];
    if (threadIdx.x == 0) { host[16] = *a; host[17] = *b; }
    __threadfence_system();
}


// ------------------------------------------------------------------------------------------- grouped mapping (gap filling)
// Many small, independent mapping problems: group g = a few target sequences (the two scaffold ends around a gap) and
// ONE read (bin/ntlink_patch_gaps.py:412-442: read_btllib_minimizers + get_accepted_anchor_contigs per gap). One block
// per group: the group's unique-minimizer table is built in shared memory (global memory for the rare group whose
// targets have more minimizers than fit), the read's minimizers are looked up and compacted in read order, and thread 0
// chains the hits with the same chain_read as the main path.
constexpr int GM_THREADS = 128;
constexpr uint32_t GM_SLOTS = 8192;                       // shared-memory table: 64 KB keys + 32 KB values + 8 KB flags
struct GroupTables { unsigned long long* keys; uint32_t* vals; uint8_t* dup; };

__global__ void __launch_bounds__(GM_THREADS) k_group_map(const uint64_t* __restrict__ t_hash, const uint32_t* __restrict__ t_posf,
                                                          const uint32_t* __restrict__ t_mx_off, const uint32_t* __restrict__ t_len,
                                                          const uint32_t* __restrict__ g_t_off,
                                                          const uint64_t* __restrict__ r_hash, const uint32_t* __restrict__ r_posf,
                                                          const uint32_t* __restrict__ r_mx_off, const uint32_t* __restrict__ r_len,
                                                          uint32_t ngroups, uint32_t g_base, MapParams P, const uint32_t* __restrict__ g_slots,
                                                          const uint64_t* __restrict__ g_tab_off, GroupTables gtab,
                                                          Hit* __restrict__ hits, Run* __restrict__ runs, uint8_t* __restrict__ mark,
                                                          uint32_t* __restrict__ nruns, uint32_t* __restrict__ nhits,
                                                          MapStatus* __restrict__ st) {
    extern __shared__ unsigned long long gm_smem[];
    __shared__ uint32_t s_warp[GM_THREADS / 32];
    __shared__ uint32_t s_running, s_special_cnt, s_special_val;
    const uint32_t g = g_base + blockIdx.x, tid = threadIdx.x;
    if (g >= ngroups) return;
    const uint32_t ts0 = g_t_off[g], ts1 = g_t_off[g + 1];            // target sequences of the group
    const uint32_t t0 = t_mx_off[ts0], t1 = t_mx_off[ts1];            // their minimizers (contiguous)
    const uint32_t slots = g_slots[g];
    const uint64_t mask = slots - 1;
    unsigned long long* keys;
    uint32_t* vals;
    uint8_t* dup;
    if (g_tab_off[g] == ~0ull) {
        keys = gm_smem; vals = reinterpret_cast<uint32_t*>(gm_smem + GM_SLOTS); dup = reinterpret_cast<uint8_t*>(vals + GM_SLOTS);
    } else {
        keys = gtab.keys + g_tab_off[g]; vals = gtab.vals + g_tab_off[g]; dup = gtab.dup + g_tab_off[g];
    }
    for (uint32_t i = tid; i < slots; i += GM_THREADS) { keys[i] = NTL_EMPTY_KEY; dup[i] = 0; }
    if (tid == 0) { s_running = 0; s_special_cnt = 0; s_special_val = 0; }
    // the table may live in global memory: its plain stores must have reached L2, where the atomics below operate,
    // before any thread of the block starts inserting (a block-level barrier alone only orders what the block sees
    // through its L1)
    __threadfence();
    __syncthreads();
    // M1 within the group (patch:397-410): first occurrence keeps its place, any second occurrence removes the hash
    for (uint32_t i = t0 + tid; i < t1; i += GM_THREADS) {
        const unsigned long long key = t_hash[i];
        if (key == NTL_EMPTY_KEY) { if (atomicAdd(&s_special_cnt, 1u) == 0) s_special_val = i - t0; continue; }
        uint64_t s = idx_slot(key, mask);
        for (;;) {
            const unsigned long long prev = atomicCAS(&keys[s], (unsigned long long)NTL_EMPTY_KEY, key);
            if (prev == NTL_EMPTY_KEY) { vals[s] = i - t0; break; }
            if (prev == key) { dup[s] = 1; break; }
            s = (s + 1) & mask;
        }
    }
    __threadfence();
    __syncthreads();
    // Which occurrence "wins" a slot is a race, but only unique hashes are ever looked up successfully, so it does not matter.
    const uint32_t r0 = r_mx_off[g], r1 = r_mx_off[g + 1];
    for (uint32_t base = r0; base < r1; base += GM_THREADS) {
        const uint32_t i = base + tid;
        bool found = false;
        Hit h; h.ctg = 0; h.cposf = 0; h.rposf = 0;
        if (i < r1) {
            const unsigned long long key = r_hash[i];
            uint32_t v = 0xFFFFFFFFu;
            if (key == NTL_EMPTY_KEY) { if (s_special_cnt == 1) v = s_special_val; }
            else {
                // volatile: read what the atomics / stores of the insert phase left in memory, not an L1 copy
                const volatile unsigned long long* vkeys = keys;
                const volatile uint32_t* vvals = vals;
                const volatile uint8_t* vdup = dup;
                uint64_t s = idx_slot(key, mask);
                for (;;) {
                    const unsigned long long kk = vkeys[s];
                    if (kk == key) { if (!vdup[s]) v = vvals[s]; break; }
                    if (kk == NTL_EMPTY_KEY) break;
                    s = (s + 1) & mask;
                }
            }
            if (v != 0xFFFFFFFFu) {
                const uint32_t gi = t0 + v;
                uint32_t ctg = ts0;
                while (ctg + 1 < ts1 && t_mx_off[ctg + 1] <= gi) ctg++;           // a handful of targets per group
                h.ctg = ctg; h.cposf = t_posf[gi]; h.rposf = r_posf[i];
                found = true;
            }
        }
        const uint32_t b = __ballot_sync(0xffffffffu, found);
        if ((tid & 31) == 0) s_warp[tid >> 5] = __popc(b);
        __syncthreads();
        uint32_t before = s_running;
        for (uint32_t q = 0; q < (tid >> 5); q++) before += s_warp[q];
        if (found) hits[r0 + before + __popc(b & ((1u << (tid & 31)) - 1u))] = h;
        __syncthreads();
        if (tid == 0) { uint32_t tot = 0; for (int q = 0; q < GM_THREADS / 32; q++) tot += s_warp[q]; s_running += tot; }
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t nh = s_running;
        uint32_t nr = 0;
        if (nh) nr = chain_read(hits + r0, nh, runs + r0, mark + r0, r_len[g], t_len, P);
        nruns[g] = nr;
        nhits[g] = nh;
        if (nr) atomicAdd(&st->n_runs, nr);
        if (nh) atomicAdd(&st->n_hits, nh);
    }
}

}  // namespace

// grow a device buffer keeping its first `keep` bytes
static int grow_preserve(ntl_ctx* c, DevBuf& b, size_t keep, size_t need) {
    if (need <= b.cap) return NTL_OK;
    DevBuf nb;
    NTL_CUDA(c, nb.ensure(need + need / 2));
    if (keep && b.p) NTL_CUDA(c, cudaMemcpyAsync(nb.p, b.p, keep, cudaMemcpyDeviceToDevice, c->stream));
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    b.release();
    b = nb;
    return NTL_OK;
}

// Build the replicated target index from device-resident minimizer triples.
// n_dev != null: n is an upper bound, the exact count is read on the device. sync = false: nothing is waited for (the
// caller guarantees that h_ctg_len / h_name_rank stay valid until the stream has consumed them, e.g. pinned buffers it owns).
int index_build_device(ntl_ctx* c, const uint64_t* d_hash, const uint32_t* d_ctg, const uint32_t* d_posf, uint64_t n,
                       const uint32_t* h_ctg_len, const uint32_t* h_name_rank, uint32_t ncontig, const uint32_t* n_dev, bool sync) {
    TargetIndex& X = c->index;
    uint64_t slots = 1024;
    while (slots < 2 * n) slots <<= 1;
    X.slots = slots; X.ncontig = ncontig; X.n_inserted = n; X.built = false;
    NTL_CUDA(c, X.table.ensure(slots * sizeof(IdxEntry)));
    NTL_CUDA(c, X.dupflag.ensure(slots));
    NTL_CUDA(c, X.special.ensure(sizeof(IdxSpecial) + 16));
    NTL_CUDA(c, X.ctg_len.ensure(((size_t)ncontig + 1) * 4));
    NTL_CUDA(c, X.name_rank.ensure(((size_t)ncontig + 1) * 4));
    tick(c, T_INDEX);
    {
        FillSegs fs{};
        fs.p[0] = X.table.p; fs.n[0] = slots * sizeof(IdxEntry); fs.v[0] = 0xFF;
        fs.p[1] = X.dupflag.p; fs.n[1] = slots; fs.v[1] = 0;
        fs.p[2] = X.special.p; fs.n[2] = sizeof(IdxSpecial) + 16; fs.v[2] = 0;
        k_fill_segs<<<std::min<uint32_t>(div_up(slots * sizeof(IdxEntry), 256), 148 * 8), 256, 0, c->stream>>>(fs);
        c->launches++;
    }
    NTL_CUDA(c, cudaMemcpyAsync(X.ctg_len.p, h_ctg_len, (size_t)ncontig * 4, cudaMemcpyHostToDevice, c->stream));
    NTL_CUDA(c, cudaMemcpyAsync(X.name_rank.p, h_name_rank, (size_t)ncontig * 4, cudaMemcpyHostToDevice, c->stream));
    if (n) {
        k_index_insert<<<div_up(n, 256), 256, 0, c->stream>>>(d_hash, d_ctg, d_posf, n, n_dev, X.table.as<IdxEntry>(), slots - 1,
                                                             X.dupflag.as<uint8_t>(), X.special.as<IdxSpecial>());
        c->launches++;
    }
    k_index_finalize<<<div_up(slots, 256), 256, 0, c->stream>>>(X.table.as<IdxEntry>(), slots, X.dupflag.as<uint8_t>(),
                                                               (unsigned long long*)((char*)X.special.p + sizeof(IdxSpecial)));
    c->launches++;
    tock(c, T_INDEX);
    NTL_CUDA(c, cudaGetLastError());
    if (sync) NTL_CUDA(c, cudaStreamSynchronize(c->stream));   // h_ctg_len / h_name_rank are caller memory
    X.built = true;
    return NTL_OK;
}
// note the exact number of minimizers of a deferred index build in the call state
__global__ void k_call_note_mx(const uint32_t* __restrict__ n_mx, CallState* __restrict__ call) { call->mx_total = *n_mx; }
int call_note_mx(ntl_ctx* c, const uint32_t* n_dev, CallState* call) {
    k_call_note_mx<<<1, 1, 0, c->stream>>>(n_dev, call);
    c->launches++;
    NTL_CUDA(c, cudaGetLastError());
    return NTL_OK;
}

int expand_contig_ids(ntl_ctx* c, const DeviceSketch& sk, DevBuf& ctg_ids) {
    NTL_CUDA(c, ctg_ids.ensure((size_t)sk.n_mx * 4 + 4));
    if (sk.n_mx) {
        k_expand_ctg<<<div_up(sk.n_mx, 256), 256, 0, c->stream>>>(sk.mx_off.as<uint32_t>(), sk.nseq, sk.n_mx, ctg_ids.as<uint32_t>());
        c->launches++;
    }
    NTL_CUDA(c, cudaGetLastError());
    return NTL_OK;
}

int call_reserve_events(ntl_ctx* c, uint32_t nreads);
int events_resolve_count(ntl_ctx* c);

// Liftover of host mappings (ntl_map_out layout) through the AGP table. The lifted runs/hits stay in c->mw (hit_off,
// nruns, runs, hits) -- exactly where map_device(pre->resident) expects them -- and the counters are returned.
int liftover_device(ntl_ctx* c, const uint32_t* hit_off, const uint32_t* nruns, const Run* runs, const Hit* hits, uint32_t nreads,
                    const AgpRow* agp, uint32_t ncontig, int32_t k, MapStatus* counts_out) {
    MapWork& M = c->mw;
    const uint32_t n = nreads ? hit_off[nreads] : 0;
    NTL_CUDA(c, M.hits.ensure((size_t)n * sizeof(Hit) + 16));
    NTL_CUDA(c, M.runs.ensure((size_t)n * sizeof(Run) + 16));
    NTL_CUDA(c, M.hit_tmp.ensure((size_t)n * sizeof(Hit) + 16));          // input hits
    NTL_CUDA(c, M.lift_runs.ensure((size_t)n * sizeof(Run) + 16));        // input runs
    NTL_CUDA(c, M.hit_flag.ensure((size_t)n * 4 + 16));                    // scratch: new ids
    NTL_CUDA(c, M.hit_pref.ensure((size_t)n * 4 + 16));                    // scratch: kept counts + flags
    NTL_CUDA(c, M.hit_off.ensure(((size_t)nreads + 2) * 4));
    NTL_CUDA(c, M.nruns.ensure(((size_t)nreads + 2) * 4));
    NTL_CUDA(c, M.lift_nruns.ensure(((size_t)nreads + 2) * 4));
    NTL_CUDA(c, M.lift_agp.ensure(((size_t)ncontig + 1) * sizeof(AgpRow)));
    NTL_CUDA(c, M.status.ensure(sizeof(MapStatus) + 64));
    NTL_CUDA(c, c->h_status.ensure(256));
    MapStatus* st = M.status.as<MapStatus>();
    {
        FillSegs fs{};
        fs.p[0] = st; fs.n[0] = sizeof(MapStatus) + 64;
        k_fill_segs<<<1, 256, 0, c->stream>>>(fs);
        c->launches++;
    }
    NTL_CUDA(c, cudaMemcpyAsync(M.hit_off.p, hit_off, ((size_t)nreads + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    if (nreads) NTL_CUDA(c, cudaMemcpyAsync(M.lift_nruns.p, nruns, (size_t)nreads * 4, cudaMemcpyHostToDevice, c->stream));
    if (n) {
        NTL_CUDA(c, cudaMemcpyAsync(M.lift_runs.p, runs, (size_t)n * sizeof(Run), cudaMemcpyHostToDevice, c->stream));
        NTL_CUDA(c, cudaMemcpyAsync(M.hit_tmp.p, hits, (size_t)n * sizeof(Hit), cudaMemcpyHostToDevice, c->stream));
    }
    if (ncontig) NTL_CUDA(c, cudaMemcpyAsync(M.lift_agp.p, agp, (size_t)ncontig * sizeof(AgpRow), cudaMemcpyHostToDevice, c->stream));
    if (nreads) {
        k_liftover<<<div_up(nreads, 128), 128, 0, c->stream>>>(M.hit_tmp.as<Hit>(), M.lift_runs.as<Run>(), M.hit_off.as<uint32_t>(),
                                                             M.lift_nruns.as<uint32_t>(), nreads, M.lift_agp.as<AgpRow>(), ncontig, k,
                                                             M.hits.as<Hit>(), M.runs.as<Run>(), M.nruns.as<uint32_t>(),
                                                             M.hit_flag.as<uint32_t>(), M.hit_pref.as<uint32_t>(), st);
        c->launches++;
    }
    k_publish_map<<<1, 32, 0, c->stream>>>(st, &st->n_hits, &st->n_runs, c->h_status.as<uint32_t>());
    c->launches++;
    NTL_CUDA(c, cudaGetLastError());
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    const MapStatus hs = *c->h_status.as<MapStatus>();
    if (hs.err & (LIFTERR_LAYOUT << 8)) { c->err = "liftover: malformed mappings (runs out of order / out of their read's region, or unknown contig id)"; return NTL_ERR_ARG; }
    if (hs.err & (LIFTERR_RANGE << 8)) { c->err = "liftover: a lifted position does not fit in 31 bits"; return NTL_ERR_ARG; }
    *counts_out = hs;
    M.lifted_reads = nreads; M.lifted_hits = n; M.lifted_valid = true;
    return NTL_OK;
}


// Grouped mapping on host arrays (see k_group_map). Results: holey per-read regions at r_mx_off (a read cannot have more
// hits than minimizers), contig ids = target sequence indices.
int group_map_device(ntl_ctx* c, const uint64_t* t_hash, const uint32_t* t_posf, const uint64_t* t_mx_off, const uint32_t* t_len,
                     uint32_t ntargets, const uint32_t* g_t_off, const uint64_t* r_hash, const uint32_t* r_posf,
                     const uint64_t* r_mx_off, const uint32_t* r_len, uint32_t ngroups, const ntl_params* prm, MapStatus* counts_out) {
    MapWork& M = c->mw;
    const uint64_t nt = ntargets ? t_mx_off[ntargets] - t_mx_off[0] : 0, nr = ngroups ? r_mx_off[ngroups] - r_mx_off[0] : 0;
    if (nt >= (1ull << 31) || nr >= (1ull << 31)) { c->err = "ntl_map_groups: too many minimizers in one call"; return NTL_ERR_ARG; }
    MapParams P;
    P.k = prm->k; P.z = prm->z; P.f = prm->f; P.x = prm->x; P.x_is_zero = (prm->x == 0.0);
    P.sensitive = prm->sensitive; P.repeat_filter = prm->repeat_filter;
    // host-side planning: 32-bit offsets, table size per group, global table room for the groups too big for shared memory
    std::vector<uint32_t> t_off32((size_t)ntargets + 1), r_off32((size_t)ngroups + 1), slots(ngroups ? ngroups : 1);
    std::vector<uint64_t> tab_off(ngroups ? ngroups : 1);
    for (uint32_t i = 0; i <= ntargets; i++) t_off32[i] = (uint32_t)(t_mx_off[i] - t_mx_off[0]);
    for (uint32_t i = 0; i <= ngroups; i++) r_off32[i] = (uint32_t)(r_mx_off[i] - r_mx_off[0]);
    uint64_t gtab_total = 0;
    for (uint32_t g = 0; g < ngroups; g++) {
        if (g_t_off[g] > g_t_off[g + 1] || g_t_off[g + 1] > ntargets) { c->err = "ntl_map_groups: group_t_off must be non-decreasing and end at ntargets"; return NTL_ERR_ARG; }
        const uint64_t n = (uint64_t)t_off32[g_t_off[g + 1]] - t_off32[g_t_off[g]];
        uint64_t sl = 64;
        while (sl < 2 * n) sl <<= 1;
        slots[g] = (uint32_t)sl;
        if (sl <= GM_SLOTS) tab_off[g] = ~0ull;
        else { tab_off[g] = gtab_total; gtab_total += sl; }
    }
    DevBuf &d_th = M.gm[0], &d_tp = M.gm[1], &d_to = M.gm[2], &d_tl = M.gm[3], &d_go = M.gm[4], &d_rh = M.gm[5], &d_rp = M.gm[6], &d_ro = M.gm[7],
           &d_rl = M.gm[8], &d_sl = M.gm[9], &d_tb = M.gm[10], &d_gk = M.gm[11], &d_gv = M.gm[12], &d_gd = M.gm[13], &d_nh = M.gm[14];
    NTL_CUDA(c, d_th.ensure(nt * 8 + 8)); NTL_CUDA(c, d_tp.ensure(nt * 4 + 4)); NTL_CUDA(c, d_to.ensure(((size_t)ntargets + 1) * 4));
    NTL_CUDA(c, d_tl.ensure((size_t)ntargets * 4 + 4)); NTL_CUDA(c, d_go.ensure(((size_t)ngroups + 1) * 4));
    NTL_CUDA(c, d_rh.ensure(nr * 8 + 8)); NTL_CUDA(c, d_rp.ensure(nr * 4 + 4)); NTL_CUDA(c, d_ro.ensure(((size_t)ngroups + 1) * 4));
    NTL_CUDA(c, d_rl.ensure((size_t)ngroups * 4 + 4)); NTL_CUDA(c, d_sl.ensure((size_t)ngroups * 4 + 4)); NTL_CUDA(c, d_tb.ensure((size_t)ngroups * 8 + 8));
    NTL_CUDA(c, d_gk.ensure(gtab_total * 8 + 8)); NTL_CUDA(c, d_gv.ensure(gtab_total * 4 + 4)); NTL_CUDA(c, d_gd.ensure(gtab_total + 8));
    NTL_CUDA(c, d_nh.ensure((size_t)ngroups * 4 + 4));
    NTL_CUDA(c, M.hits.ensure((size_t)nr * sizeof(Hit) + 16)); NTL_CUDA(c, M.runs.ensure((size_t)nr * sizeof(Run) + 16));
    NTL_CUDA(c, M.mark.ensure((size_t)nr + 16)); NTL_CUDA(c, M.nruns.ensure(((size_t)ngroups + 2) * 4));
    NTL_CUDA(c, M.hit_off.ensure(((size_t)ngroups + 2) * 4));
    NTL_CUDA(c, M.status.ensure(sizeof(MapStatus) + 64));
    NTL_CUDA(c, c->h_status.ensure(256));
    MapStatus* st = M.status.as<MapStatus>();
    {
        FillSegs fs{};
        fs.p[0] = st; fs.n[0] = sizeof(MapStatus) + 64;
        k_fill_segs<<<1, 256, 0, c->stream>>>(fs);
        c->launches++;
    }
    auto up = [&](DevBuf& d, const void* src, size_t bytes) -> cudaError_t {
        return bytes ? cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, c->stream) : cudaSuccess;
    };
    NTL_CUDA(c, up(d_th, t_hash + t_mx_off[0], nt * 8)); NTL_CUDA(c, up(d_tp, t_posf + t_mx_off[0], nt * 4));
    NTL_CUDA(c, up(d_to, t_off32.data(), ((size_t)ntargets + 1) * 4)); NTL_CUDA(c, up(d_tl, t_len, (size_t)ntargets * 4));
    NTL_CUDA(c, up(d_go, g_t_off, ((size_t)ngroups + 1) * 4));
    NTL_CUDA(c, up(d_rh, r_hash + r_mx_off[0], nr * 8)); NTL_CUDA(c, up(d_rp, r_posf + r_mx_off[0], nr * 4));
    NTL_CUDA(c, up(d_ro, r_off32.data(), ((size_t)ngroups + 1) * 4)); NTL_CUDA(c, up(d_rl, r_len, (size_t)ngroups * 4));
    NTL_CUDA(c, up(d_sl, slots.data(), (size_t)ngroups * 4)); NTL_CUDA(c, up(d_tb, tab_off.data(), (size_t)ngroups * 8));
    NTL_CUDA(c, up(M.hit_off, r_off32.data(), ((size_t)ngroups + 1) * 4));
    if (ngroups) {
        static bool attr_set = false;
        const size_t smem = (size_t)GM_SLOTS * (8 + 4 + 1);
        if (!attr_set) { NTL_CUDA(c, cudaFuncSetAttribute(k_group_map, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_set = true; }
        tick(c, T_CHAIN);
        GroupTables gt{d_gk.as<unsigned long long>(), d_gv.as<uint32_t>(), d_gd.as<uint8_t>()};
        k_group_map<<<ngroups, GM_THREADS, smem, c->stream>>>(d_th.as<uint64_t>(), d_tp.as<uint32_t>(), d_to.as<uint32_t>(), d_tl.as<uint32_t>(),
                                                             d_go.as<uint32_t>(), d_rh.as<uint64_t>(), d_rp.as<uint32_t>(), d_ro.as<uint32_t>(),
                                                             d_rl.as<uint32_t>(), ngroups, 0u, P, d_sl.as<uint32_t>(), d_tb.as<uint64_t>(), gt,
                                                             M.hits.as<Hit>(), M.runs.as<Run>(), M.mark.as<uint8_t>(), M.nruns.as<uint32_t>(),
                                                             d_nh.as<uint32_t>(), st);
        c->launches++;
        tock(c, T_CHAIN);
    }
    k_publish_map<<<1, 32, 0, c->stream>>>(st, &st->n_hits, &st->n_runs, c->h_status.as<uint32_t>());
    c->launches++;
    NTL_CUDA(c, cudaGetLastError());
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));        // also: the host vectors above are done with
    *counts_out = *c->h_status.as<MapStatus>();
    c->mw.lifted_valid = false;
    return NTL_OK;
}

// Lookup + chain + events for the reads whose sketch is `sk` (device). d_read_len: device read lengths.
// Results are left in c->mw (hits, runs, nruns, hit_off, events log segment); the counters are returned.
// With `pre` set the lookup and chaining are skipped: the accepted runs/hits come from the host (checkpoint path,
// bin/ntlink_pair.py:437-488) and only the pair events are computed.
// With `call` set (sync-free call) nothing is waited for: the event buffer is sized from what was enough before, an
// overflow or a failed assertion sets CALLERR_MAP in call->err, the events go to the log at the device cursor call->log_n
// and the chunk's totals are added to *call by k_chunk_finish (call_chunk_finish); counts_out stays zero.
int map_device(ntl_ctx* c, const DeviceSketch& sk, const uint32_t* d_read_len, uint32_t nreads, uint64_t first_ordinal,
               const ntl_params* prm, MapStatus* counts_out, uint64_t* log_base_out, const PreMappings* pre, CallState* call) {
    if (!c->index.built) { c->err = "map: no target index (call ntl_index_build first)"; return NTL_ERR_STATE; }
    if (!call) NTL_TRY(events_resolve_count(c));
    MapWork& M = c->mw;
    const uint32_t n = pre ? pre->n_hits : sk.n_mx;
    MapParams P;
    P.k = prm->k; P.z = prm->z; P.f = prm->f; P.x = prm->x; P.x_is_zero = (prm->x == 0.0);
    P.sensitive = prm->sensitive; P.repeat_filter = prm->repeat_filter;
    NTL_CUDA(c, M.hit_tmp.ensure((size_t)n * sizeof(Hit) + 16));
    NTL_CUDA(c, M.hit_flag.ensure((size_t)n * 4 + 16));
    NTL_CUDA(c, M.hit_pref.ensure((size_t)n * 4 + 16));
    NTL_CUDA(c, M.hits.ensure((size_t)n * sizeof(Hit) + 16));
    NTL_CUDA(c, M.runs.ensure((size_t)n * sizeof(Run) + 16));
    NTL_CUDA(c, M.mark.ensure((size_t)n + 16));
    NTL_CUDA(c, M.hit_off.ensure(((size_t)nreads + 2) * 4));
    NTL_CUDA(c, M.nruns.ensure(((size_t)nreads + 2) * 4));
    NTL_CUDA(c, M.ev_cnt.ensure(((size_t)nreads + 2) * 4 * 4));   // evmax | ev_off | ev_cnt | ev_pref
    NTL_CUDA(c, M.status.ensure(sizeof(MapStatus) + 64));
    NTL_CUDA(c, c->h_status.ensure(256));
    MapStatus* st = M.status.as<MapStatus>();
    uint32_t* n_dev = (uint32_t*)((char*)M.status.p + sizeof(MapStatus));
    uint32_t* nreads_dev = n_dev + 1;
    uint32_t* evmax = M.ev_cnt.as<uint32_t>();
    uint32_t* ev_off = evmax + (nreads + 2);
    uint32_t* ev_cnt = ev_off + (nreads + 2);
    uint32_t* ev_pref = ev_cnt + (nreads + 2);
    uint32_t ev_cap = std::max<uint32_t>(std::max<uint32_t>(1 << 16, 4 * nreads), c->ev_cap_hint);
    int attempt = 0;

    {
        FillSegs fs{};
        fs.p[0] = st; fs.n[0] = sizeof(MapStatus) + 64;
        k_fill_segs<<<1, 256, 0, c->stream>>>(fs);
        c->launches++;
    }
    if (pre) {
        if (!pre->resident) {
            NTL_CUDA(c, cudaMemcpyAsync(M.hit_off.p, pre->hit_off, ((size_t)nreads + 1) * 4, cudaMemcpyHostToDevice, c->stream));
            NTL_CUDA(c, cudaMemcpyAsync(M.nruns.p, pre->nruns, (size_t)nreads * 4, cudaMemcpyHostToDevice, c->stream));
            if (n) {
                NTL_CUDA(c, cudaMemcpyAsync(M.runs.p, pre->runs, (size_t)n * sizeof(Run), cudaMemcpyHostToDevice, c->stream));
                NTL_CUDA(c, cudaMemcpyAsync(M.hits.p, pre->hits, (size_t)n * sizeof(Hit), cudaMemcpyHostToDevice, c->stream));
            }
        }
        tick(c, T_CHAIN);
        k_evmax<<<div_up(std::max<uint32_t>(nreads, 1), 256), 256, 0, c->stream>>>(
            M.nruns.as<uint32_t>(), nreads, P.f, evmax, st, nreads_dev, n, M.hit_off.as<uint32_t>(), M.runs.as<Run>(), M.hits.as<Hit>(),
            pre->compute_read_len ? const_cast<uint32_t*>(d_read_len) : nullptr);
        c->launches++;
        goto events_stage;
    }
    {
    IndexView ix{c->index.table.as<IdxEntry>(), c->index.slots - 1, c->index.special.as<IdxSpecial>()};
    tick(c, T_LOOKUP);
    k_lookup<<<div_up(std::max<uint32_t>(n, 1), 256), 256, 0, c->stream>>>(sk.hash.as<uint64_t>(), sk.posf.as<uint32_t>(), n, sk.n_dev, ix,
                                                                         M.hit_tmp.as<Hit>(), M.hit_flag.as<uint32_t>(), n_dev);
    c->launches++;
    NTL_TRY(exclusive_scan_u32(c, M.hit_flag.as<uint32_t>(), M.hit_pref.as<uint32_t>(), n_dev, n, M.blocksums));
    if (n) {
        k_compact_hits<<<div_up(n, 256), 256, 0, c->stream>>>(M.hit_tmp.as<Hit>(), M.hit_flag.as<uint32_t>(),
                                                            M.hit_pref.as<uint32_t>(), n_dev, M.hits.as<Hit>());
        c->launches++;
    }
    k_hit_offsets<<<div_up((uint64_t)nreads + 1, 256), 256, 0, c->stream>>>(sk.mx_off.as<uint32_t>(), M.hit_pref.as<uint32_t>(), nreads,
                                                                          M.hit_off.as<uint32_t>(), st, nreads_dev);
    c->launches++;
    tock(c, T_LOOKUP);

    tick(c, T_CHAIN);
    if (nreads) {
        k_chain<<<div_up(nreads, CHAIN_WARPS), CHAIN_WARPS * 32, 0, c->stream>>>(M.hits.as<Hit>(), M.runs.as<Run>(), M.mark.as<uint8_t>(),
                                                          M.hit_off.as<uint32_t>(), d_read_len, nreads,
                                                          c->index.ctg_len.as<uint32_t>(), P, M.nruns.as<uint32_t>(), evmax, st);
        c->launches++;
    }
    }
events_stage:
    NTL_TRY(exclusive_scan_u32(c, evmax, ev_off, nreads_dev, nreads, M.blocksums));
retry_events:
    NTL_CUDA(c, M.events.ensure((size_t)ev_cap * sizeof(Event)));
    if (nreads) {
        k_events<<<div_up(nreads, 128), 128, 0, c->stream>>>(M.hits.as<Hit>(), M.runs.as<Run>(), M.hit_off.as<uint32_t>(), d_read_len,
                                                           M.nruns.as<uint32_t>(), ev_off, nreads, (uint32_t)first_ordinal,
                                                           c->index.ctg_len.as<uint32_t>(), c->index.name_rank.as<uint32_t>(), P,
                                                           ev_cap, M.events.as<Event>(), ev_cnt, st);
        c->launches++;
    }
    NTL_TRY(exclusive_scan_u32(c, ev_cnt, ev_pref, nreads_dev, nreads, M.blocksums));
    if (call) {
        // room for everything the chunks in flight may append, then append at the device cursor
        NTL_TRY(call_reserve_events(c, nreads));
        c->tl_pending_bound += ev_cap;
        if (nreads) {
            k_compact_events<<<div_up(nreads, 128), 128, 0, c->stream>>>(M.events.as<Event>(), ev_off, ev_cnt, ev_pref, nreads,
                                                                       c->tl_events.as<Event>(), st, call);
            c->launches++;
        }
        tock(c, T_CHAIN);
        NTL_CUDA(c, cudaGetLastError());
        if (counts_out) *counts_out = MapStatus{};
        if (log_base_out) *log_base_out = 0;
        return NTL_OK;
    }
    // counters -> host (one synchronisation), then append the events to the device log
    k_publish_map<<<1, 32, 0, c->stream>>>(st, ev_pref + nreads, ev_off + nreads, c->h_status.as<uint32_t>());
    c->launches++;
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    MapStatus hs = *c->h_status.as<MapStatus>();
    const uint32_t n_events = nreads ? *(uint32_t*)(c->h_status.as<char>() + 64) : 0;
    const uint32_t ev_need = nreads ? *(uint32_t*)(c->h_status.as<char>() + 68) : 0;
    if (hs.err & MAPERR_ASSERT) {
        c->err = "an assertion of the reference would fail (bin/ntlink_pair.py:225 / :173-184): minimizer order or overhang";
        return NTL_ERR_ASSERT;
    }
    if (hs.err & MAPERR_READ_EVENTS) {
        c->err = "a read maps to so many contigs that it would emit more than 16,777,215 contig pairs (-f too large for this read)";
        return NTL_ERR_ARG;
    }
    if (hs.err & MAPERR_EVENTS) {
        if (++attempt > 2) { c->err = "map: event buffer exhausted"; return NTL_ERR_WORKSPACE; }
        ev_cap = ev_need + 1024;
        c->ev_cap_hint = std::max(c->ev_cap_hint, ev_cap + ev_cap / 4);
        k_set_u32<<<1, 1, 0, c->stream>>>(&st->err, 0u);
        c->launches++;
        goto retry_events;
    }
    NTL_TRY(grow_preserve(c, c->tl_events, c->tl_n_events * sizeof(Event), (c->tl_n_events + n_events + 1) * sizeof(Event)));
    if (nreads) {
        k_compact_events<<<div_up(nreads, 128), 128, 0, c->stream>>>(M.events.as<Event>(), ev_off, ev_cnt, ev_pref, nreads,
                                                                   c->tl_events.as<Event>() + c->tl_n_events, st, nullptr);
        c->launches++;
    }
    tock(c, T_CHAIN);
    NTL_CUDA(c, cudaGetLastError());
    hs.n_events = n_events;
    *counts_out = hs;
    *log_base_out = c->tl_n_events;
    c->tl_n_events += n_events;
    return NTL_OK;
}

// ---- sync-free call plumbing (used by capi.cu)
// room in the device event log for one more chunk of `nreads` reads (called by map_device; capi.cu calls it ahead of
// a stream capture, where growing the log -- a copy and a synchronisation -- is not possible)
int call_reserve_events(ntl_ctx* c, uint32_t nreads) {
    const uint32_t ev_cap = std::max<uint32_t>(std::max<uint32_t>(1 << 16, 4 * nreads), c->ev_cap_hint);
    const size_t need = (c->tl_n_events + c->tl_pending_bound + ev_cap + 1) * sizeof(Event);
    if (need <= c->tl_events.cap) return NTL_OK;
    if (c->capturing) { c->err = "internal: event log not reserved before capture"; return NTL_ERR_STATE; }
    return grow_preserve(c, c->tl_events, (c->tl_n_events + c->tl_pending_bound) * sizeof(Event), need);
}
int call_begin(ntl_ctx* c, CallState** call_out) {
    NTL_TRY(events_resolve_count(c));
    NTL_CUDA(c, c->call_state.ensure(sizeof(CallState)));
    NTL_CUDA(c, c->h_status.ensure(256));
    if (c->tl_n_events >= (1ull << 31)) { c->err = "event log too large"; return NTL_ERR_WORKSPACE; }
    c->tl_pending_bound = 0;
    k_call_begin<<<1, 1, 0, c->stream>>>(c->call_state.as<CallState>(), (uint32_t)c->tl_n_events);
    c->launches++;
    *call_out = c->call_state.as<CallState>();
    return NTL_OK;
}
// after map_device(call): chunk results -> host arrays (H may be null: counters only), totals advanced
int call_chunk_finish(ntl_ctx* c, CallState* call, uint32_t rb, uint32_t nreads, const HostResults* H) {
    MapWork& M = c->mw;
    MapStatus* st = M.status.as<MapStatus>();
    const uint32_t* n_dev = (const uint32_t*)((char*)M.status.p + sizeof(MapStatus));
    const uint32_t* evmax = M.ev_cnt.as<uint32_t>();
    const uint32_t* ev_cnt = evmax + 2 * ((size_t)nreads + 2);
    const uint32_t* ev_pref = evmax + 3 * ((size_t)nreads + 2);
    if (H) {
        k_results_host<<<div_up(((uint64_t)nreads + 1) * 32, 128), 128, 0, c->stream>>>(M.hit_off.as<uint32_t>(), M.nruns.as<uint32_t>(), M.runs.as<Run>(),
                                                                               M.hits.as<Hit>(), ev_cnt, ev_pref, c->tl_events.as<Event>(), rb,
                                                                               nreads, *H, call);
        c->launches++;
    }
    k_chunk_finish<<<1, 1, 0, c->stream>>>(st, n_dev, call);
    c->launches++;
    NTL_CUDA(c, cudaGetLastError());
    return NTL_OK;
}
// one synchronisation for the whole call; on success the host-side log size follows the device cursor
int call_end(ntl_ctx* c, CallState* call, CallState* host_out) {
    k_call_publish<<<1, 32, 0, c->stream>>>(call, c->h_status.as<uint32_t>());
    c->launches++;
    NTL_CUDA(c, cudaGetLastError());
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    *host_out = *c->h_status.as<CallState>();
    c->tl_pending_bound = 0;
    if (host_out->err == 0) {
        const uint32_t appended = host_out->log_n - (uint32_t)c->tl_n_events;
        c->ev_cap_hint = std::max(c->ev_cap_hint, std::min<uint32_t>(appended + appended / 2 + 1024, 1u << 28));
        c->tl_n_events = host_out->log_n;
    } else if (host_out->err & (MAPERR_EVENTS << 8)) {
        c->ev_cap_hint = std::max<uint32_t>(c->ev_cap_hint * 4, 1u << 18);
    }
    return NTL_OK;
}

// after ntl_events_import_device the host only knows a bound of the log size: fetch the exact count (one synchronisation)
int events_resolve_count(ntl_ctx* c) {
    if (!c->tl_count_on_device) return NTL_OK;
    uint32_t cw[2] = {0, 0};
    NTL_CUDA(c, cudaMemcpyAsync(cw, c->tl_count.p, 8, cudaMemcpyDeviceToHost, c->stream));
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    c->tl_count_on_device = false;
    if (cw[1]) { c->tl_n_events = 0; c->err = "a rank sent more events than the exchange buffer holds (ntl_events_import_device)"; return NTL_ERR_WORKSPACE; }
    c->tl_n_events = cw[0];
    return NTL_OK;
}

// Import of the gathered exchange buffers with the per-rank counts left on the device: rank r's events go behind those of
// the ranks before it; *count = {total, overflow flag}.
__global__ void k_import_gathered(const uint32_t* __restrict__ src, uint32_t world, uint32_t cap, Event* __restrict__ log,
                                  uint32_t* __restrict__ count) {
    const uint32_t r = blockIdx.y;
    const size_t stride = ((size_t)cap + 1) * 6;                    // words per rank: header row + cap events
    uint32_t before = 0, ovf = 0;
    for (uint32_t q = 0; q < world; q++) {
        const uint32_t cq = src[q * stride];
        if (cq > cap) ovf = 1;
        if (q < r) before += min(cq, cap);
    }
    const uint32_t mine = min(src[r * stride], cap);
    if (r == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        uint32_t total = 0;
        for (uint32_t q = 0; q < world; q++) total += min(src[q * stride], cap);
        count[0] = total; count[1] = ovf;
    }
    const uint32_t* ev = src + r * stride + 6;
    uint32_t* dst = reinterpret_cast<uint32_t*>(log + before);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < mine * 6; i += gridDim.x * blockDim.x) dst[i] = ev[i];
}
int events_import_device(ntl_ctx* c, const void* d_src, uint32_t world, uint64_t cap_events) {
    const uint64_t bound = (uint64_t)world * cap_events;
    if (bound >= (1ull << 31)) { c->err = "event exchange buffers too large"; return NTL_ERR_ARG; }
    NTL_CUDA(c, c->tl_events.ensure((bound + 1) * sizeof(Event)));
    NTL_CUDA(c, c->tl_count.ensure(16));
    k_import_gathered<<<dim3(std::max<uint32_t>(1, std::min<uint32_t>(div_up(cap_events * 6, 256), 64)), world), 256, 0, c->stream>>>(
        (const uint32_t*)d_src, world, (uint32_t)cap_events, c->tl_events.as<Event>(), c->tl_count.as<uint32_t>());
    c->launches++;
    NTL_CUDA(c, cudaGetLastError());
    c->tl_n_events = bound;                  // upper bound until the tally has read the exact count back
    c->tl_count_on_device = true;
    return NTL_OK;
}

// Pair table over the whole device event log.
int tally_device(ntl_ctx* c, std::vector<ntl_pair>& pairs, std::vector<int32_t>& gaps) {
    uint64_t n = c->tl_n_events;
    pairs.clear(); gaps.clear();
    if (n == 0) return NTL_OK;
    if (c->tl_count_on_device) {
        // After a sync-free event exchange the host only knows the capacity of the exchange buffers (ranks x 4 x the largest
        // count seen); the exact number of events is a word on the device. Fetch it first -- one short synchronisation, and
        // the tally ends with one anyway -- so that the pair table, its fills, scans and the per-slot kernels are sized by
        // the events and not by that bound (N = 8, configs[3]: 137 k events against a bound of ~1.1 M).
        NTL_CUDA(c, c->tw.h_stage.ensure(64));
        uint32_t* cw = c->tw.h_stage.as<uint32_t>();
        NTL_CUDA(c, cudaMemcpyAsync(cw, c->tl_count.p, 8, cudaMemcpyDeviceToHost, c->stream));
        NTL_CUDA(c, cudaStreamSynchronize(c->stream));
        c->tl_count_on_device = false;
        if (cw[1]) { c->tl_n_events = 0; c->err = "a rank sent more events than the exchange buffer holds (ntl_events_import_device)"; return NTL_ERR_WORKSPACE; }
        n = std::min<uint64_t>(n, cw[0]);
        c->tl_n_events = n;
        if (n == 0) return NTL_OK;
    }
    if (n >= (1ull << 31)) { c->err = "tally: too many events"; return NTL_ERR_WORKSPACE; }
    uint64_t slots = 1024;
    while (slots < 2 * n) slots <<= 1;
    // persistent work buffers (cudaMalloc/cudaFree per call would dominate a small tally)
    TallyWork& T = c->tw;
    DevBuf &keys = T.keys, &pn = T.pn, &panchor = T.panchor, &pfirst = T.pfirst, &ev_slot = T.ev_slot, &gap_off = T.gap_off,
           &cursor = T.cursor, &gkey = T.gkey, &gval = T.gval, &nonempty = T.nonempty, &ppref = T.ppref, &out = T.out,
           &ndev = T.ndev, &bs = T.bs;
    int rc = NTL_OK;
    auto cleanup = [&]() {};
#define TL_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { c->err = std::string("tally: ") + cudaGetErrorString(e__); cleanup(); return NTL_ERR_CUDA; } } while (0)
    TL_CUDA(keys.ensure(slots * 8)); TL_CUDA(pn.ensure(slots * 4)); TL_CUDA(panchor.ensure(slots * 4));
    TL_CUDA(pfirst.ensure(slots * 8)); TL_CUDA(ev_slot.ensure(n * 4)); TL_CUDA(gap_off.ensure((slots + 1) * 4));
    TL_CUDA(cursor.ensure(slots * 4)); TL_CUDA(gkey.ensure(n * 8)); TL_CUDA(gval.ensure(n * 4));
    TL_CUDA(nonempty.ensure(slots * 4)); TL_CUDA(ppref.ensure((slots + 1) * 4)); TL_CUDA(ndev.ensure(16));
    TL_CUDA(T.skey.ensure(n * 8)); TL_CUDA(T.sval.ensure(n * 4));
    tick(c, T_TALLY);
    {
        FillSegs fs{};
        fs.p[0] = keys.p; fs.n[0] = slots * 8; fs.v[0] = 0xFF;
        fs.p[1] = pfirst.p; fs.n[1] = slots * 8; fs.v[1] = 0xFF;
        fs.p[2] = pn.p; fs.n[2] = slots * 4; fs.v[2] = 0;
        fs.p[3] = panchor.p; fs.n[3] = slots * 4; fs.v[3] = 0;
        fs.p[4] = cursor.p; fs.n[4] = slots * 4; fs.v[4] = 0;
        k_fill_segs<<<std::min<uint32_t>(div_up(slots * 8, 256), 148 * 8), 256, 0, c->stream>>>(fs);
        c->launches++;
    }
    const Event* ev = c->tl_events.as<Event>();
    k_set_u32<<<1, 1, 0, c->stream>>>(ndev.as<uint32_t>(), (uint32_t)slots);
    const uint32_t* n_src = c->tl_count_on_device ? c->tl_count.as<uint32_t>() : nullptr;
    k_tally_insert<<<div_up(n, 256), 256, 0, c->stream>>>(ev, n, n_src, keys.as<unsigned long long>(), slots - 1, pn.as<uint32_t>(),
                                                         panchor.as<uint32_t>(), pfirst.as<unsigned long long>(), ev_slot.as<uint32_t>());
    c->launches += 2;
    rc = exclusive_scan_u32(c, pn.as<uint32_t>(), gap_off.as<uint32_t>(), ndev.as<uint32_t>(), (uint32_t)slots, bs);
    if (rc != NTL_OK) { cleanup(); return rc; }
    k_tally_scatter<<<div_up(n, 256), 256, 0, c->stream>>>(ev, n, n_src, ev_slot.as<uint32_t>(), gap_off.as<uint32_t>(), cursor.as<uint32_t>(),
                                                          gkey.as<unsigned long long>(), gval.as<int32_t>());
    k_tally_sort<<<div_up(slots, SORT_WARPS), SORT_WARPS * 32, 0, c->stream>>>(pn.as<uint32_t>(), gap_off.as<uint32_t>(), (uint32_t)slots,
                                                           gkey.as<unsigned long long>(), gval.as<int32_t>(), nonempty.as<uint32_t>(),
                                                           T.skey.as<unsigned long long>(), T.sval.as<int32_t>());
    c->launches += 2;
    rc = exclusive_scan_u32(c, nonempty.as<uint32_t>(), ppref.as<uint32_t>(), ndev.as<uint32_t>(), (uint32_t)slots, bs);
    if (rc != NTL_OK) { cleanup(); return rc; }
    uint32_t n_pairs = 0;
    // a pair needs at least one event: with few events the table is copied at that bound and one synchronisation is enough
    const bool one_sync = n * sizeof(ntl_pair) <= (8u << 20);
    if (!one_sync) {
        TL_CUDA(cudaMemcpyAsync(&n_pairs, ppref.as<uint32_t>() + slots, 4, cudaMemcpyDeviceToHost, c->stream));
        TL_CUDA(cudaStreamSynchronize(c->stream));
    }
    const size_t pair_cap = one_sync ? (size_t)n : (size_t)n_pairs;
    TL_CUDA(out.ensure(pair_cap * sizeof(ntl_pair) + 16));
    k_pairs_compact<<<div_up(slots, 128), 128, 0, c->stream>>>(keys.as<unsigned long long>(), pn.as<uint32_t>(), panchor.as<uint32_t>(),
                                                              pfirst.as<unsigned long long>(), gap_off.as<uint32_t>(), ppref.as<uint32_t>(),
                                                              (uint32_t)slots, (ntl_pair*)out.p);
    c->launches++;
    tock(c, T_TALLY);
    // through a pinned staging buffer: a device->host copy into pageable memory is several times slower, and the copy
    // is sized by the bound (pair_cap), not by the number of pairs
    const size_t pair_bytes = pair_cap * sizeof(ntl_pair), gap_bytes = (size_t)n * 4;
    size_t gap_bytes_copy = gap_bytes;
    const size_t cnt_at = (pair_bytes + gap_bytes + 15) & ~(size_t)15;
    TL_CUDA(T.h_stage.ensure(cnt_at + 64));
    char* hs = T.h_stage.as<char>();
    if (one_sync) TL_CUDA(cudaMemcpyAsync(hs + cnt_at, ppref.as<uint32_t>() + slots, 4, cudaMemcpyDeviceToHost, c->stream));
    if (n_src) TL_CUDA(cudaMemcpyAsync(hs + cnt_at + 8, n_src, 8, cudaMemcpyDeviceToHost, c->stream));   // {exact count, overflow flag}
    if (pair_cap) TL_CUDA(cudaMemcpyAsync(hs, out.p, pair_bytes, cudaMemcpyDeviceToHost, c->stream));
    TL_CUDA(cudaMemcpyAsync(hs + pair_bytes, gval.p, gap_bytes, cudaMemcpyDeviceToHost, c->stream));
    TL_CUDA(cudaStreamSynchronize(c->stream));
    if (one_sync) n_pairs = *reinterpret_cast<const uint32_t*>(hs + cnt_at);
    uint64_t n_exact = n;
    if (n_src) {
        const uint32_t* cw = reinterpret_cast<const uint32_t*>(hs + cnt_at + 8);
        c->tl_count_on_device = false;
        if (cw[1]) { c->tl_n_events = 0; c->err = "a rank sent more events than the exchange buffer holds (ntl_events_import_device)"; return NTL_ERR_WORKSPACE; }
        n_exact = cw[0];
        c->tl_n_events = n_exact;                                  // from here on the host knows the exact size again
    }
    pairs.resize(n_pairs); gaps.resize(n_exact);
    if (n_exact < n) gap_bytes_copy = (size_t)n_exact * 4;
    if (n_pairs) memcpy(pairs.data(), hs, (size_t)n_pairs * sizeof(ntl_pair));
    memcpy(gaps.data(), hs + pair_bytes, gap_bytes_copy);
#undef TL_CUDA
    cleanup();
    // first-seen order (the reference's dict order): radix-sorted permutation of the first-seen keys, then one pass that moves
    // the 40-byte rows (a comparison sort of the rows, later of (key, index) records on four threads, was most of rank 0's
    // tally time at human scale)
    {
        const size_t np_ = pairs.size();
        std::vector<uint64_t> key(np_);
        std::vector<uint32_t> perm(np_);
        for (size_t i = 0; i < np_; i++) key[i] = pairs[i].first_key;
        order_first_seen(key.data(), (uint32_t)np_, perm.data());
        std::vector<ntl_pair> sorted(np_);
        for (size_t i = 0; i < np_; i++) sorted[i] = pairs[perm[i]];
        pairs.swap(sorted);
    }
    return NTL_OK;
}

int read_len_device(ntl_ctx* c, const uint64_t* d_off, uint32_t nreads, DevBuf& out) {
    NTL_CUDA(c, out.ensure(((size_t)nreads + 1) * 4));
    if (nreads) { k_read_len<<<div_up(nreads, 256), 256, 0, c->stream>>>(d_off, nreads, out.as<uint32_t>()); c->launches++; }
    NTL_CUDA(c, cudaGetLastError());
    return NTL_OK;
}

}  // namespace ntl
