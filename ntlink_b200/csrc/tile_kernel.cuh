// tile_kernel.cuh -- k_tile: the minimizer sketch as ONE pass over the packed sequence (included by sketch.cu).
//
// Replaces k_dense + k_overflow + k_select + k_seq_gaps + k_gap + k_emit and the three scans between them for the
// window sizes ntLink runs with (w >= 13). A block owns a TILE of consecutive strips (a strip = S k-mer positions of one
// sequence, one thread each) plus H halo strips on either side, and everything stays in shared memory:
//
//   1 dense    every thread rolls ntHash over its strip (process_strip_dev, same hot loop as before) and keeps the
//              candidates -- k-mers with h0 < tau -- in its slot row in shared memory (12 bytes each, SoA);
//   2 select   every own candidate is decided with the position-local rule of sketch_logic.cuh, written with sentinels:
//              A = consecutive valid k-mers to the left with h0 >= mine, B = to the right with h0 > mine, both stopping
//              at the ends of the sequence; minimizer  <=>  min(A, w-1) + min(B, w-1) >= w-1. The neighbour scans walk
//              the slot rows of the view (shared memory, 32-bit arithmetic);
//   3 exact    whatever the candidates cannot answer becomes a RECORD handled by an exact windowed scan (one warp each):
//              candidate-free stretches of >= w valid k-mers that touch the own strips (their minimizers are
//              non-candidates), and candidates whose neighbourhood leaves the view (halo shortened by runs of N).
//              A tile answers for the k-mer POSITIONS of its own strips only, so tiles never depend on one another's
//              threshold: a tile whose slot rows overflow (low-complexity sequence) simply starts over with tau / 8;
//   4 emit     per-strip output counts -> block scan -> decoupled look-back over the tiles (one 64-bit word per tile)
//              -> second hash, pos|strand written in position order; per-sequence offsets on the way.
//
// DRAM traffic: the packed bases once (+ halo), the minimizers once. Nothing else leaves the SM.
#pragma once

namespace ntl {
namespace {

constexpr int TILE_THREADS = 128;
constexpr int TILE_MAXREC = 32;          // exact-scan records per tile (more -> the batch falls back to the multi-pass path)
constexpr int TILE_GT = 512;             // k-mer positions hashed per round of the exact scan (per warp)
constexpr int TILE_KPAD = 20;            // padding of the select key array (the unrolled neighbour scans look up to 16 entries away)
constexpr int TILE_SCAN_WARPS = 1;       // warps that run exact scans (their hash buffer aliases the select keys, dead by then)
constexpr uint32_t REC_LEADING = 0x80000000u;   // attach: before all candidates of view strip (attach & 0xFFFF)
constexpr uint32_t TILE_TBL_ENTRIES = 37;       // roll-table entries (in << 3 | out, codes 0..4) that exist
constexpr uint32_t TILE_TBL_BYTES = TILE_TBL_ENTRIES * TBL_STRIDE;

struct TileParams {
    uint32_t k, w, S, cap, tau_hi, nseq, H;      // H halo strips; own strips per tile = TILE_THREADS - 2H
    uint64_t mult;
    uint32_t out_cap, extras_cap, ntiles_max;
    uint32_t flat_cap, shared_at;                // candidates a view may hold; byte offset of TileShared in dynamic shared memory
    uint32_t tcap;                               // staging entries per tile
};

struct TileRec {                          // one exact scan
    uint32_t v;                           // view strip of the sequence (for gseq / L / np)
    uint32_t emit_lo, emit_hi;            // k-mer positions whose minimizers this record may emit
    uint32_t clip_lo, clip_hi;            // the scan never needs to look outside [clip_lo, clip_hi)
    uint32_t special;                     // a candidate position that may be emitted too (NONE32: non-candidates only)
    uint32_t attach;                      // outputs go right after flat candidate `attach`, or REC_LEADING | view strip
    uint32_t out_off, out_cnt;            // where the scan left them (extras)
    uint32_t out_at;                      // tile-local output rank of the first one
};

struct Scan4 { uint32_t a_ex, b_ex, a_tot, b_tot, mx, mn; };
struct Scan4Scratch { uint32_t a[4], b[4], mx[4], mn[4]; };
struct TileShared {
    uint32_t cnt[TILE_THREADS], nv[TILE_THREADS], vbase[TILE_THREADS + 1], foff[TILE_THREADS + 1];
    uint32_t seq[TILE_THREADS], p0[TILE_THREADS], npos_strip[TILE_THREADS], flags[TILE_THREADS];   // flags: 1 first, 2 last strip of its sequence
    uint32_t pre[TILE_THREADS], prex[TILE_THREADS];       // outputs of leading records of a strip; exclusive prefix over the strips
    Scan4Scratch scan4[2];
    TileRec rec[TILE_MAXREC];
    uint32_t nrec, overflow, tile, gbase, err, pool_n;
    uint8_t sfirst[TILE_THREADS], slast[TILE_THREADS];    // first / last view strip of the strip's sequence
};

// Candidates of the whole view go into ONE pool in shared memory (no per-strip slack): a slot is taken with a shared-memory
// atomic, every entry points at the previous candidate of the same strip; the position order is rebuilt afterwards
// (flat list). meta word at emission: offset in the strip (8 bits) | forward (bit 8) | view strip (7 bits, from bit 9) |
// ordinal among the strip's valid k-mers (from bit 16); the ordinal becomes tile-wide once the strips have been scanned.
struct SmemEmit {
    uint32_t *lo, *hi, *meta;
    uint16_t* prev;
    uint32_t* pool_n;
    uint32_t cap, count, p0, vbits, tail;
    // staging of the 8-step interior blocks: slot s of this thread lives at index s * TILE_THREADS (conflict-free columns)
    unsigned long long* m_h; uint32_t* m_m; uint32_t n;
    // slow path (blocks with bounds / validity checks): one candidate, straight into the pool
    __device__ __forceinline__ void operator()(uint64_t h0, uint32_t pos, bool fwd, uint32_t lord) {
        const uint32_t i = atomicAdd(pool_n, 1u);
        if (i < cap) {
            lo[i] = (uint32_t)h0; hi[i] = (uint32_t)(h0 >> 32); meta[i] = (pos - p0) | (fwd ? 0x100u : 0u) | vbits | (lord << 16);
            prev[i] = (uint16_t)tail; tail = i;
        }
        count++;
    }
    __device__ __forceinline__ void fast(uint64_t h0, uint32_t pos, bool fwd, uint32_t lord) { (*this)(h0, pos, fwd, lord); }
    __device__ __forceinline__ bool room_for_block() const { return true; }
    // fast path: the candidate test of every step costs a predicated store pair, nothing else; the pool slot (a
    // shared-memory atomic) is taken once per 8-step block for everything the block found
    __device__ __forceinline__ void push(bool c, uint64_t h0, uint32_t pos, bool fwd, uint32_t lord) {
        const uint32_t m = (pos - p0) | (fwd ? 0x100u : 0u) | vbits | (lord << 16);
        if (c) { m_h[n * TILE_THREADS] = h0; m_m[n * TILE_THREADS] = m; }
        n += c ? 1u : 0u;
    }
    __device__ __forceinline__ void flush_block() {
        if (n) {
            const uint32_t base = atomicAdd(pool_n, n);
            for (uint32_t s = 0; s < n; s++) {
                const uint32_t i = base + s;
                if (i < cap) {
                    const unsigned long long h0 = m_h[s * TILE_THREADS];
                    lo[i] = (uint32_t)h0; hi[i] = (uint32_t)(h0 >> 32); meta[i] = m_m[s * TILE_THREADS];
                    prev[i] = (uint16_t)tail; tail = i;
                }
            }
            count += n; n = 0;
        }
    }
};

struct GapTileEmit {
    unsigned long long* h; uint8_t* f; uint32_t base;
    __device__ __forceinline__ void operator()(uint64_t h0, uint32_t pos, bool fwd, uint32_t) { h[pos - base] = h0; f[pos - base] = fwd ? 1 : 0; }
};
struct RecSerialEmit {
    Cand* dst; uint32_t count, cap, lo, hi, special, tau_hi;
    __device__ __forceinline__ void operator()(uint64_t h0, uint32_t pos, bool fwd) {
        if (pos < lo || pos >= hi) return;
        if ((uint32_t)(h0 >> 32) < tau_hi && pos != special) return;
        if (count < cap) { Cand c; c.h0 = h0; c.posf = pos | (fwd ? FWD_BIT : 0u); c.lord = 0; dst[count] = c; }
        count++;
    }
};

__device__ __forceinline__ uint32_t tile_block_scan(uint32_t v, uint32_t* sm /* [34] */, uint32_t* total) {
    uint32_t incl = v;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += t; }
    if (lane == 31) sm[wid] = incl;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t s = 0; for (int q = 0; q < TILE_THREADS / 32; q++) { const uint32_t t = sm[q]; sm[q] = s; s += t; } sm[33] = s; }
    __syncthreads();
    const uint32_t r = sm[wid] + incl - v;
    *total = sm[33];
    __syncthreads();
    return r;
}

// number of valid k-mers among positions [a, b) of a CLEAN range = b - a; a range is clean when no base of it is invalid
__device__ __forceinline__ bool range_dirty_warp(const uint32_t* __restrict__ packed, uint64_t gseq, uint32_t base_lo, uint32_t nb, uint32_t lane) {
    bool dirty = false;
    for (uint32_t b = lane * 8; b < nb; b += 256) {
        uint32_t wd = fetch8(packed, gseq + base_lo + b);
        const uint32_t left = nb - b;
        if (left < 8) wd &= (1u << (4 * left)) - 1u;
        dirty |= (wd & 0x44444444u) != 0u;
    }
    return __any_sync(0xffffffffu, dirty);
}

// One record, one warp. Exact windowed minimizers (rightmost argmin of every window of w consecutive valid k-mers) over a
// range that covers w-1 valid k-mers on either side of [emit_lo, emit_hi) (or reaches the clip / the sequence ends);
// emitted: argmins inside the emit range that are non-candidates under tau (or the record's special position).
__device__ void tile_exact_scan(const uint32_t* __restrict__ packed, const RollEntry* __restrict__ tbl_g, const unsigned char* tbl_s,
                                uint64_t gseq, uint32_t L, uint32_t np, const TileParams& P, uint32_t tau_hi, TileRec& r,
                                Cand* __restrict__ extras, uint32_t* __restrict__ extras_used, uint32_t* __restrict__ err,
                                unsigned long long* H, uint8_t* F) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t w = P.w, k = P.k;
    uint32_t lo = r.emit_lo > w - 1 ? r.emit_lo - (w - 1) : 0;
    uint32_t hi = (uint64_t)r.emit_hi + (w - 1) < np ? r.emit_hi + (w - 1) : np;
    if (lo < r.clip_lo) lo = r.clip_lo;
    if (hi > r.clip_hi) hi = r.clip_hi;
    const bool dirty = hi <= lo || range_dirty_warp(packed, gseq, lo, hi - lo + k - 1, lane) || w > TILE_GT / 2;
    if (dirty) {
        // serial: widen the range until it holds w-1 valid k-mers on either side of the emit range (or hits a bound)
        uint32_t a = r.emit_lo, b = r.emit_hi;
        if (lane == 0) {
            {
                // nb = smallest invalid base index >= a that matters (below a + k), NONE32 if none
                uint32_t have = 0, nb = NONE32;
                for (uint32_t j = 0; j < k && a + j < L; j++) if (fetch1(packed, gseq + a + j) >= CODE_INVALID) { nb = a + j; break; }
                while (a > r.clip_lo && have < w - 1) {
                    a--;
                    if (fetch1(packed, gseq + a) >= CODE_INVALID) nb = a;
                    if (nb == NONE32 || nb >= a + k) have++;
                }
            }
            {
                // run = consecutive valid bases ending at base b + k - 1 (capped by the scan start): k-mer b valid <=> run >= k
                uint32_t have = 0, run = 0;
                const uint32_t bound = min(r.clip_hi, np);
                for (uint32_t j = 0; j < k && b + j < L; j++) run = fetch1(packed, gseq + b + j) >= CODE_INVALID ? 0u : run + 1;
                while (b < bound && have < w - 1) {
                    if (run >= k) have++;
                    b++;
                    run = (b + k - 1 < L && fetch1(packed, gseq + b + k - 1) < CODE_INVALID) ? run + 1 : 0u;
                }
            }
        }
        lo = __shfl_sync(0xffffffffu, a, 0); hi = __shfl_sync(0xffffffffu, b, 0);
    }
    // at most one output per window, and only inside the emit range
    const uint32_t max_out = min(r.emit_hi - r.emit_lo, hi - lo >= w ? hi - lo - w + 1 : 0u);
    if (max_out == 0) { if (lane == 0) { r.out_off = 0; r.out_cnt = 0; } return; }
    uint32_t off = 0;
    if (lane == 0) off = atomicAdd(extras_used, max_out);
    off = __shfl_sync(0xffffffffu, off, 0);
    if ((uint64_t)off + max_out > P.extras_cap) { if (lane == 0) { atomicOr(err, SKERR_EXTRAS); r.out_off = 0; r.out_cnt = 0; } return; }
    Cand* out = extras + off;
    uint32_t n_out = 0;
    if (dirty) {
        if (lane == 0) {
            RecSerialEmit em{out, 0, max_out, r.emit_lo, r.emit_hi, r.special, tau_hi};
            gap_scan(packed, tbl_g, gseq, L, k, w, lo, hi, em);
            n_out = min(em.count, max_out);
        }
        n_out = __shfl_sync(0xffffffffu, n_out, 0);
    } else {
        const uint32_t G = hi - lo;                                 // all valid
        const uint32_t nwin = G >= w ? G - w + 1 : 0;
        uint32_t prev_amin = NONE32;
        for (uint32_t wb = 0; wb < nwin; wb += TILE_GT - w + 1) {
            const uint32_t we = min(nwin, wb + (TILE_GT - w + 1));
            const uint32_t npt = (we - wb) + w - 1;
            const uint32_t base = lo + wb;
            const uint32_t chunk = (npt + 31) / 32;
            const uint32_t pa = min(npt, lane * chunk), pb = min(npt, pa + chunk);
            __syncwarp();
            if (pb > pa) {
                GapTileEmit te{H, F, base};
                process_strip<true>(packed, gseq, base + pa, pb - pa, k, reinterpret_cast<const RollEntry*>(tbl_s), TBL_STRIDE / 16, 0u, te);
            }
            __syncwarp();
            for (uint32_t j0 = wb; j0 < we; j0 += 32) {
                const uint32_t j = j0 + lane;
                uint32_t amin = NONE32;
                unsigned long long hmin = 0;
                if (j < we) {
                    const uint32_t o = j - wb;
                    hmin = H[o]; amin = o;
                    for (uint32_t q = 1; q < w; q++) {
                        const unsigned long long hv = H[o + q];
                        if (hv <= hmin) { hmin = hv; amin = o + q; }
                    }
                    amin += base;
                }
                uint32_t left = __shfl_up_sync(0xffffffffu, amin, 1);
                if (lane == 0) left = prev_amin;
                bool emit = (j < we) && (amin != left) && (hmin != 0xFFFFFFFFFFFFFFFFULL) && amin >= r.emit_lo && amin < r.emit_hi;
                if (emit && (uint32_t)(hmin >> 32) < tau_hi && amin != r.special) emit = false;
                const uint32_t ballot = __ballot_sync(0xffffffffu, emit);
                if (emit) {
                    const uint32_t at = n_out + __popc(ballot & ((1u << lane) - 1u));
                    if (at < max_out) { Cand c; c.h0 = hmin; c.posf = amin | (F[amin - base] ? FWD_BIT : 0u); c.lord = 0; out[at] = c; }
                }
                n_out += __popc(ballot);
                const uint32_t nvalid = min(32u, we - j0);
                prev_amin = __shfl_sync(0xffffffffu, amin, nvalid - 1);
            }
        }
        n_out = min(n_out, max_out);
    }
    if (lane == 0) { r.out_off = off; r.out_cnt = n_out; }
}

// tile state word: bits 63..62 = 0 empty, 1 aggregate, 2 inclusive prefix; low 62 bits = count
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) { return *reinterpret_cast<const volatile unsigned long long*>(p); }

// One barrier for four block-wide scans of per-thread values: exclusive sums of a and b, inclusive forward max of mx,
// inclusive reverse min of mn. Warp-level shuffles, the warp aggregates go through one of two scratch sets (alternating:
// a scratch set is only rewritten after the barrier of the call in between).
__device__ __forceinline__ Scan4 tile_scan4(uint32_t a, uint32_t b, uint32_t mx, uint32_t mn, Scan4Scratch* sc) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t ai = a, bi = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t ta = __shfl_up_sync(0xffffffffu, ai, d), tb = __shfl_up_sync(0xffffffffu, bi, d), tm = __shfl_up_sync(0xffffffffu, mx, d);
        const uint32_t tn = __shfl_down_sync(0xffffffffu, mn, d);
        if (lane >= (uint32_t)d) { ai += ta; bi += tb; mx = max(mx, tm); }
        if (lane + d < 32) mn = min(mn, tn);
    }
    if (lane == 31) { sc->a[wid] = ai; sc->b[wid] = bi; sc->mx[wid] = mx; }
    if (lane == 0) sc->mn[wid] = mn;
    __syncthreads();
    Scan4 r;
    uint32_t ca = 0, cb = 0, ta = 0, tb = 0, cm = 0, cn = 0xFFFFFFFFu;
#pragma unroll
    for (uint32_t q = 0; q < TILE_THREADS / 32; q++) {
        const uint32_t xa = sc->a[q], xb = sc->b[q];
        if (q < wid) { ca += xa; cb += xb; cm = max(cm, sc->mx[q]); }
        if (q > wid) cn = min(cn, sc->mn[q]);
        ta += xa; tb += xb;
    }
    r.a_ex = ca + ai - a; r.b_ex = cb + bi - b; r.a_tot = ta; r.b_tot = tb; r.mx = max(mx, cm); r.mn = min(mn, cn);
    return r;
}

// meta word of a candidate: offset in its strip (8 bits) | forward strand (bit 8) | view strip (bits 9..15) | valid-k-mer
// ordinal (bits 16..: within the strip at emission, tile-wide once the flat list is built)
#define TM_OFF(m) ((m) & 0xFFu)
#define TM_FWD(m) (((m) >> 8) & 1u)
#define TM_V(m) (((m) >> 9) & 0x7Fu)
#define TM_ORD(m) ((m) >> 16)

__global__ void __launch_bounds__(TILE_THREADS) k_tile(const uint32_t* __restrict__ packed, const uint64_t* __restrict__ seq_off,
                                                       const uint32_t* __restrict__ strip_off, const uint32_t* __restrict__ strip_seq,
                                                       TileParams P, const RollEntry* __restrict__ tbl_g,
                                                       uint32_t* __restrict__ tile_cnt, uint32_t* __restrict__ ticket,
                                                       Cand* __restrict__ extras, uint64_t* __restrict__ out_hash, uint32_t* __restrict__ out_posf,
                                                       uint32_t* __restrict__ mx_off, SketchStatus* __restrict__ st) {
    extern __shared__ __align__(256) unsigned char dyn[];
    unsigned char* tbl_s = dyn;                                                     // roll table (entry stride 256 B, 16 copies)
    uint32_t* p_lo = reinterpret_cast<uint32_t*>(dyn + TILE_TBL_BYTES);            // candidate pool (SoA)
    uint32_t* p_hi = p_lo + P.flat_cap;
    uint32_t* p_meta = p_hi + P.flat_cap;
    uint16_t* p_prev = reinterpret_cast<uint16_t*>(p_meta + P.flat_cap);            // previous candidate of the same strip
#define c_lo(i) p_lo[i]
#define c_hi(i) p_hi[i]
#define c_meta(i) p_meta[i]
#define c_prev(i) p_prev[i]
    uint16_t* f_idx = p_prev + P.flat_cap + 2;                                      // flat list: pool slot of the e-th candidate of the view
    uint16_t* f_rank = f_idx + P.flat_cap + 2;                                      // outputs before it (own candidates)
    uint8_t* f_sel = reinterpret_cast<uint8_t*>(f_rank + P.flat_cap + 2);           // 1: it is a minimizer
    // select keys in flat order (contiguous, so the neighbour scans are plain strided loads): high hash word, and the
    // ordinal in f_rank (the ranks are only computed after the select phase); TILE_KPAD entries of padding on both sides
    unsigned char* const key_region = dyn + ((TILE_TBL_BYTES + P.flat_cap * 12 + (P.flat_cap + 2) * 6 + (P.flat_cap + 2) + 7u) & ~7u);
    uint32_t* k_hi = reinterpret_cast<uint32_t*>(key_region) + TILE_KPAD;
    uint16_t* k_ord = f_rank;
    unsigned char* const mini_region = dyn + ((TILE_TBL_BYTES + P.flat_cap * 12 + P.flat_cap * 2 + 4 + 7u) & ~7u);     // from f_idx on
    unsigned long long* gap_h = reinterpret_cast<unsigned long long*>(key_region);  // exact-scan buffers: after the select phase
    uint8_t* gap_f = reinterpret_cast<uint8_t*>(gap_h + TILE_GT);
    TileShared& T = *reinterpret_cast<TileShared*>(dyn + P.shared_at);
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (uint32_t i = tid; i < TILE_TBL_ENTRIES * TBL_COPIES; i += TILE_THREADS) {
        const RollEntry e = tbl_g[i / TBL_COPIES];
        reinterpret_cast<uint4*>(tbl_s)[i] = make_uint4((uint32_t)e.f, (uint32_t)(e.f >> 32), (uint32_t)e.r, (uint32_t)(e.r >> 32));
    }
    const uint32_t nstrips = strip_off[P.nseq];
    const uint32_t H = P.H, NS = TILE_THREADS - 2 * H, w = P.w, W1 = P.w - 1;
    const uint32_t ntiles = (nstrips + NS - 1) / NS;
    if (blockIdx.x == 0 && tid == 0) st->nstrips = nstrips;
    if (tid == 0) T.err = 0;
    __syncthreads();

    for (;;) {
        // the previous tile is finished for thread 0 once it gets here; the others may still be writing its outputs, which
        // touches neither of the words reset below
        if (tid == 0) {
            if (T.err) { atomicOr(&st->err, T.err); T.err = 0; }
            T.tile = atomicAdd(ticket, 1u); T.nrec = 0; T.pool_n = 0;
        }
        __syncthreads();
        const uint32_t tile = T.tile;
        if (tile >= ntiles) break;
        const uint32_t b0 = tile * NS, b1 = min(b0 + NS, nstrips), nown = b1 - b0;
        // ---- view strip of this thread
        const int64_t sg = (int64_t)b0 - H + tid;
        const bool live = sg >= 0 && sg < (int64_t)nstrips;
        const bool own = live && tid >= H && (uint32_t)sg < b1;
        uint32_t q = NONE32, p0 = 0, n = 0, fl = 0, np = 0;
        uint64_t gseq = 0;
        if (live) {
            const uint32_t s = (uint32_t)sg;
            q = strip_seq[s];
            gseq = seq_off[q];
            np = seq_npos(seq_off[q + 1] - gseq, P.k, P.w);
            const uint32_t fs = strip_off[q], es = strip_off[q + 1];
            p0 = (s - fs) * P.S;
            n = min(P.S, np - p0);
            fl = (s == fs ? 1u : 0u) | (s + 1 == es ? 2u : 0u);
        }
        T.seq[tid] = q; T.p0[tid] = p0; T.flags[tid] = fl; T.npos_strip[tid] = np; T.pre[tid] = 0;
        uint32_t tau_hi = P.tau_hi;
        uint32_t vb = 0, fo = 0, tail = 0xFFFFu, sfirst = 0, slast = 0, count = 0;
        // ---- 1: dense (again with a smaller threshold while the candidate pool overflows)
        for (;;) {
            uint32_t nvalid = 0;
            count = 0;
            if (live) {
                // the staging columns use the flat-list / select-key region, which is dead during the dense phase
                SmemEmit em{p_lo, p_hi, p_meta, p_prev, &T.pool_n, P.flat_cap, 0u, p0, tid << 9, 0xFFFFu,
                            reinterpret_cast<unsigned long long*>(mini_region) + tid, reinterpret_cast<uint32_t*>(mini_region + 8 * 8 * TILE_THREADS) + tid, 0u};
                nvalid = process_strip_dev(packed, gseq, p0, n, P.k, tbl_s, (tid & 15u) << 4, tau_hi, em);
                count = em.count;
                tail = em.tail;
            }
            // valid k-mers and candidates before my strip; first / last view strip of my strip's sequence (the view may
            // cut the sequence on either side)
            const Scan4 sc = tile_scan4(nvalid, count, (fl & 1u) ? tid : 0u, (fl & 2u) ? tid : (live ? TILE_THREADS - 1 : tid), &T.scan4[0]);
            vb = sc.a_ex; fo = sc.b_ex; sfirst = sc.mx; slast = sc.mn;
            T.cnt[tid] = count; T.nv[tid] = nvalid; T.vbase[tid] = vb; T.foff[tid] = fo;
            T.sfirst[tid] = (uint8_t)sfirst; T.slast[tid] = (uint8_t)slast;
            if (tid == TILE_THREADS - 1) { T.vbase[TILE_THREADS] = sc.a_tot; T.foff[TILE_THREADS] = sc.b_tot; }
            if (sc.b_tot <= P.flat_cap) break;           // same value in every thread
            __syncthreads();
            if (tid == 0) T.pool_n = 0;
            __syncthreads();
            tau_hi = tau_hi >> 3;            // tau = 0: nothing is a candidate, the exact scan does everything
        }
        // flat list of the view's candidates in position order; ordinals into the meta words
        {
            uint32_t i = tail;
            for (uint32_t j = count; j-- > 0;) {                                      // my candidates, last one first
                const uint32_t m = c_meta(i) + (vb << 16);
                c_meta(i) = m;
                f_idx[fo + j] = (uint16_t)i; k_hi[fo + j] = c_hi(i); k_ord[fo + j] = (uint16_t)TM_ORD(m);
                i = c_prev(i);
            }
        }
        __syncthreads();

        // ---- 2: select, one candidate per thread + records
        auto add_rec = [&](uint32_t v, uint32_t elo, uint32_t ehi, uint32_t clo, uint32_t chi, uint32_t special, uint32_t attach) {
            if (ehi <= elo) return;
            const uint32_t id = atomicAdd(&T.nrec, 1u);
            if (id >= TILE_MAXREC) { T.err = SKERR_GAPS; return; }
            TileRec r; r.v = v; r.emit_lo = elo; r.emit_hi = ehi; r.clip_lo = clo; r.clip_hi = chi; r.special = special;
            r.attach = attach; r.out_off = 0; r.out_cnt = 0; r.out_at = 0;
            T.rec[id] = r;
        };
        // own k-mer positions of strip v's sequence segment end at seg_hi(v)
        auto seg_hi_of = [&](uint32_t v) { const uint32_t vz = min((uint32_t)T.slast[v], H + nown - 1); return (T.flags[vz] & 2u) ? T.npos_strip[v] : T.p0[vz] + P.S; };
        const uint32_t e_own0 = T.foff[H], e_own1 = T.foff[H + nown];
        // per-candidate neighbour scans (only where the stack walk below gives up): 0 no, 1 minimizer, 2 cannot tell
        auto decide_linear = [&](uint32_t e) -> uint32_t {
            const uint32_t idx0 = f_idx[e], v = TM_V(c_meta(idx0));
            const uint32_t me_hi = c_hi(idx0), me_lo = c_lo(idx0), ord = TM_ORD(c_meta(idx0));
            const uint32_t vf = T.sfirst[v], vl = T.slast[v], e_lo = T.foff[vf], e_hi = T.foff[vl + 1];
            uint32_t A = 0, B = 0; bool unknown = false;
            for (uint32_t p = e;;) {
                if (p == e_lo) { A = ord - T.vbase[vf]; unknown = !(T.flags[vf] & 1u) && A < W1; break; }
                p--;
                const uint32_t idx = f_idx[p], d = ord - TM_ORD(c_meta(idx)), oh = c_hi(idx);
                if (d >= w) { A = W1; break; }
                if (oh < me_hi || (oh == me_hi && c_lo(idx) < me_lo)) { A = d - 1; break; }
            }
            for (uint32_t p = e + 1;; p++) {
                if (p >= e_hi) { B = T.vbase[vl] + T.nv[vl] - 1 - ord; if (!(T.flags[vl] & 2u) && B < W1) unknown = true; break; }
                const uint32_t idx = f_idx[p], d = TM_ORD(c_meta(idx)) - ord, oh = c_hi(idx);
                if (d >= w) { B = W1; break; }
                if (oh < me_hi || (oh == me_hi && c_lo(idx) <= me_lo)) { B = d - 1; break; }
            }
            return unknown ? 2u : (min(A, W1) + min(B, W1) >= W1 ? 1u : 0u);
        };
        // one own candidate per thread: fixed-length predicated scans over the neighbouring keys (no data-dependent
        // control flow inside a group of four steps, all loads independent of one another); the few candidates with more
        // than 16 neighbours inside their window go through decide_linear
        for (uint32_t eb = e_own0; eb < e_own1; eb += TILE_THREADS) {
            const bool act = eb + tid < e_own1;
            const uint32_t e = act ? eb + tid : e_own0;                      // idle lanes re-read a valid entry
            uint32_t me_hi = 0, ord = 0, e_lo = 0, e_hi = 0, bndA = 0, bndB = 0;
            bool unkA = false, unkB = false;
            if (act) {
                const uint32_t v = TM_V(c_meta(f_idx[e]));
                const uint32_t vf = T.sfirst[v], vl = T.slast[v];
                me_hi = k_hi[e]; ord = k_ord[e];
                e_lo = T.foff[vf]; e_hi = T.foff[vl + 1];
                bndA = ord - T.vbase[vf]; unkA = !(T.flags[vf] & 1u) && bndA < W1;
                bndB = T.vbase[vl] + T.nv[vl] - 1 - ord; unkB = !(T.flags[vl] & 2u) && bndB < W1;
            }
            uint32_t A = 0, B = 0;
            bool fA = !act, fB = !act, hitA = false, hitB = false, tie = false;
            for (int d0 = 1; d0 <= 16; d0 += 4) {
#pragma unroll
                for (int dd = 0; dd < 4; dd++) {
                    const int d = d0 + dd;
                    {
                        const int p = (int)e - d;
                        const bool inb = p >= (int)e_lo;
                        const uint32_t oh = k_hi[p], dist = ord - k_ord[p];
                        tie |= !fA && inb && oh == me_hi && dist < w;
                        const bool stop = !inb || dist >= w || oh < me_hi;
                        if (!fA && stop) { A = inb ? (dist >= w ? W1 : dist - 1) : bndA; hitA = !inb; }
                        fA |= stop;
                    }
                    {
                        const int p = (int)e + d;
                        const bool inb = p < (int)e_hi;
                        const uint32_t oh = k_hi[p], dist = k_ord[p] - ord;
                        tie |= !fB && inb && oh == me_hi && dist < w;
                        const bool stop = !inb || dist >= w || oh < me_hi;
                        if (!fB && stop) { B = inb ? (dist >= w ? W1 : dist - 1) : bndB; hitB = !inb; }
                        fB |= stop;
                    }
                }
                if (__all_sync(0xffffffffu, fA && fB)) break;
            }
            if (act) {
                uint32_t r;
                if (tie || !fA || !fB) r = decide_linear(e);                 // equal high words or a crowded window: the exact walk
                else if ((hitA && unkA) || (hitB && unkB)) r = 2;
                else r = min(A, W1) + min(B, W1) >= W1 ? 1u : 0u;
                // records: a candidate the view cannot decide, the candidate-free stretch after this candidate
                const uint32_t m0 = c_meta(f_idx[e]), v = TM_V(m0), pos = T.p0[v] + TM_OFF(m0);
                const uint32_t vl = T.slast[v];
                if (r == 2) { r = 0; add_rec(v, pos, pos + 1, 0, T.npos_strip[v], pos, e); }
                f_sel[e] = (uint8_t)r;
                uint32_t gap_len, nxt_pos; bool nxt_known = true;
            if (e + 1 < e_hi) { const uint32_t m = c_meta(f_idx[e + 1]); gap_len = TM_ORD(m) - ord - 1; nxt_pos = T.p0[TM_V(m)] + TM_OFF(m); }
            else if (T.flags[vl] & 2u) { gap_len = T.vbase[vl] + T.nv[vl] - 1 - ord; nxt_pos = T.npos_strip[v]; }
            else { gap_len = 0xFFFFFFFFu; nxt_pos = 0; nxt_known = false; }
                if (gap_len >= w) {
                    const uint32_t seg_hi = seg_hi_of(v);
                    add_rec(v, pos + 1, min(nxt_known ? nxt_pos : seg_hi, seg_hi), pos + 1, nxt_known ? nxt_pos : T.npos_strip[v], NONE32, e);
                }
            }
        }
        // the stretch that reaches an own segment from the left: handled by the segment's first own strip
        if (own && (tid == H || (fl & 1u))) {
            const uint32_t vf = T.sfirst[tid], vl = T.slast[tid];
            const uint32_t e_lo = T.foff[vf], e_hi = T.foff[vl + 1];
            const uint32_t ef = T.foff[tid];                                         // first candidate at or after the segment start
            const uint32_t seg_hi = seg_hi_of(tid);
            uint32_t f_pos = 0, f_ord = 0; bool f_known = true;
            if (ef < e_hi) { const uint32_t m = c_meta(f_idx[ef]); f_pos = T.p0[TM_V(m)] + TM_OFF(m); f_ord = TM_ORD(m); }
            else if (T.flags[vl] & 2u) { f_pos = np; f_ord = T.vbase[vl] + T.nv[vl]; }
            else f_known = false;
            uint32_t pr_pos = 0, pr_ord = 0; bool pr_known = true, pr_is_start = false;
            if (ef > e_lo) { const uint32_t m = c_meta(f_idx[ef - 1]); pr_pos = T.p0[TM_V(m)] + TM_OFF(m); pr_ord = TM_ORD(m); }
            else if (T.flags[vf] & 1u) { pr_is_start = true; pr_ord = T.vbase[vf]; }
            else pr_known = false;
            // an own candidate's stretch is queued by the candidate itself (above); here only what starts left of the segment
            {
                bool stretch = true;                                                // a bound outside the view: let the exact scan decide
                if (pr_known && f_known) stretch = (f_ord - pr_ord - (pr_is_start ? 0u : 1u)) >= w;
                if (stretch) {
                    const uint32_t clo = pr_known ? (pr_is_start ? 0u : pr_pos + 1) : 0u;
                    const uint32_t chi = f_known ? f_pos : np;
                    add_rec(tid, max(p0, clo), min(seg_hi, chi), clo, chi, NONE32, REC_LEADING | tid);
                }
            }
        }
        __syncthreads();
        // ---- 3: exact scans, one warp per record
        const uint32_t nrec = min(T.nrec, (uint32_t)TILE_MAXREC);
        if (wid < TILE_SCAN_WARPS)
            for (uint32_t id = wid; id < nrec; id += TILE_SCAN_WARPS) {
                TileRec& r = T.rec[id];
                const uint32_t qq = T.seq[r.v];
                const uint64_t gs = seq_off[qq];
                tile_exact_scan(packed, tbl_g, tbl_s, gs, (uint32_t)(seq_off[qq + 1] - gs), T.npos_strip[r.v], P, tau_hi, r, extras, &st->extras_used,
                                &T.err, gap_h, gap_f);
            }
        if (nrec) __syncthreads();                                                 // same value in every thread
        // ---- 4: output ranks (own candidates in flat order, each followed by its records), tile offset, emit
        const uint32_t n_own = e_own1 - e_own0;
        const uint32_t chunk = (n_own + TILE_THREADS - 1) / TILE_THREADS;
        const uint32_t ca = min(n_own, tid * chunk), cb = min(n_own, ca + chunk);
        uint32_t mine = 0;
        for (uint32_t i = ca; i < cb; i++) mine += f_sel[e_own0 + i];
        if (nrec) {                                                               // rare: records attached to my candidates / my strip
            for (uint32_t id = 0; id < nrec; id++) {
                const TileRec& r = T.rec[id];
                if (r.attach & REC_LEADING) { if ((r.attach & 0xFFFFu) == tid) T.pre[tid] += r.out_cnt; }
                else if (r.attach - e_own0 >= ca && r.attach - e_own0 < cb) mine += r.out_cnt;
            }
        }
        // outputs of the chunks before mine; leading records of the strips before mine
        const Scan4 so = tile_scan4(mine, T.pre[tid], 0u, 0u, &T.scan4[1]);
        const uint32_t tot_c = so.a_tot;
        T.prex[tid] = so.b_ex;
        const uint32_t tile_total = so.a_tot + so.b_tot;
        {
            uint32_t r = so.a_ex;
            for (uint32_t i = ca; i < cb; i++) {
                f_rank[e_own0 + i] = (uint16_t)r;
                r += f_sel[e_own0 + i];
                if (nrec) for (uint32_t id = 0; id < nrec; id++) if (T.rec[id].attach == e_own0 + i) { T.rec[id].out_at = r; r += T.rec[id].out_cnt; }
            }
            if (tid == TILE_THREADS - 1) f_rank[e_own1] = (uint16_t)tot_c;
        }
        // the tile's minimizers go to its own staging segment (no waiting for other tiles); k_tile_gather packs the segments
        if (tid == 0) {
            tile_cnt[tile] = tile_total;
            if (tile_total > P.tcap) T.err |= SKERR_OUT;
        }
        __syncthreads();
        const uint32_t gb = tile * P.tcap;
        const bool writable = tile_total <= P.tcap;
        if (own && (fl & 1u)) {                                                      // a sequence starts here: its offset inside the
            const uint32_t o = f_rank[T.foff[tid]] + T.prex[tid];                    // tile, and that of the empty sequences before it
            mx_off[q] = o;
            for (uint32_t qe = q; qe-- > 0 && strip_off[qe] == strip_off[qe + 1];) mx_off[qe] = o;
        }
        if (writable) {
            for (uint32_t eb = e_own0; eb < e_own1; eb += TILE_THREADS) {
                const uint32_t e = eb + tid;
                if (e < e_own1 && f_sel[e]) {
                    const uint32_t idx = f_idx[e], m = c_meta(idx), v = TM_V(m);
                    const uint32_t o = gb + f_rank[e] + T.prex[v] + T.pre[v];
                    const uint64_t h0 = ((uint64_t)c_hi(idx) << 32) | c_lo(idx);
                    out_hash[o] = second_hash(h0, P.mult);
                    out_posf[o] = (T.p0[v] + TM_OFF(m)) | (TM_FWD(m) ? FWD_BIT : 0u);
                }
            }
            for (uint32_t id = wid; id < nrec; id += TILE_THREADS / 32) {            // record outputs, one warp each
                const TileRec& r = T.rec[id];
                uint32_t o;
                if (r.attach & REC_LEADING) { const uint32_t v = r.attach & 0xFFFFu; o = gb + f_rank[T.foff[v]] + T.prex[v]; }
                else { const uint32_t v = TM_V(c_meta(f_idx[r.attach])); o = gb + r.out_at + T.prex[v] + T.pre[v]; }
                for (uint32_t i = lane; i < r.out_cnt; i += 32) {
                    const Cand c = extras[r.out_off + i];
                    out_hash[o + i] = second_hash(c.h0, P.mult); out_posf[o + i] = c.posf;
                }
            }
        }
        __syncthreads();       // the pool and the flat list are free again (thread 0 may take the next ticket)
    }
}

// after k_tile: exclusive scan of the tiles' minimizer counts (one block; a batch has at most a few 10^5 tiles)
__global__ void __launch_bounds__(1024) k_tile_scan(const uint32_t* __restrict__ strip_off, TileParams P, const uint32_t* __restrict__ tile_cnt,
                                                    uint32_t* __restrict__ tile_base, SketchStatus* __restrict__ st) {
    __shared__ uint32_t sm[34];
    const uint32_t nstrips = strip_off[P.nseq];
    const uint32_t NS = TILE_THREADS - 2 * P.H;
    const uint32_t ntiles = (nstrips + NS - 1) / NS;
    const uint32_t per = (ntiles + 1023) / 1024;
    const uint32_t a = min(ntiles, threadIdx.x * per), b = min(ntiles, a + per);
    uint32_t s = 0;
    for (uint32_t i = a; i < b; i++) s += tile_cnt[i];
    // block scan over 1024 partial sums
    uint32_t incl = s;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += t; }
    if (lane == 31) sm[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t x = sm[lane], xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, xi, d); if (lane >= (uint32_t)d) xi += t; }
        sm[lane] = xi - x;
        if (lane == 31) sm[32] = xi;
    }
    __syncthreads();
    uint32_t r = sm[wid] + incl - s;
    for (uint32_t i = a; i < b; i++) { tile_base[i] = r; r += tile_cnt[i]; }
    if (threadIdx.x == 0) {
        const uint32_t total = sm[32];
        tile_base[ntiles] = total;
        if (total > P.out_cap) atomicOr(&st->err, SKERR_OUT);
    }
}

// staging segments -> the packed output; per-sequence offsets; totals; error gating. One block per tile (grid-stride).
__global__ void __launch_bounds__(256) k_tile_gather(const uint32_t* __restrict__ strip_off, TileParams P, const uint32_t* __restrict__ tile_cnt,
                                                     const uint32_t* __restrict__ tile_base, const uint64_t* __restrict__ st_hash,
                                                     const uint32_t* __restrict__ st_posf, uint64_t* __restrict__ out_hash,
                                                     uint32_t* __restrict__ out_posf, uint32_t* __restrict__ mx_off, SketchStatus* __restrict__ st,
                                                     CallState* __restrict__ call, uint32_t deferred) {
    const uint32_t nstrips = strip_off[P.nseq];
    const uint32_t NS = TILE_THREADS - 2 * P.H;
    const uint32_t ntiles = (nstrips + NS - 1) / NS;
    const uint32_t total = tile_base[ntiles];
    const bool bad = st->err != 0;
    if (!bad)
        for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const uint32_t n = tile_cnt[t], o = tile_base[t];
            const uint64_t from = (uint64_t)t * P.tcap;
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) { out_hash[o + i] = st_hash[from + i]; out_posf[o + i] = st_posf[from + i]; }
        }
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q <= P.nseq; q += gridDim.x * blockDim.x) {
        if (bad && deferred) mx_off[q] = 0;
        else if (q == P.nseq || strip_off[q] == nstrips) mx_off[q] = total;
        else mx_off[q] += tile_base[strip_off[q] / NS];                          // k_tile left the offset inside the tile
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->n_mx = (bad && deferred) ? 0u : total;
        if (bad && call) atomicOr(&call->err, CALLERR_SKETCH);
    }
}

}  // namespace
}  // namespace ntl
