// small_kernel.cuh -- k_small: the sketch for small windows (w <= SMALL_W_MAX: the overlap stage `indexlr -k15 -w5`,
// ntLink:243-251, and gap filling `-k20 -w10`, bin/ntlink_patch_gaps.py:417-420), included by sketch.cu.
//
// With windows this small a large share of all k-mers are minimizers (2 / (w + 1)) and every k-mer would be a
// "candidate" of the sparse path, whose 16-byte candidate round trip per position then dominates. This kernel keeps
// everything on chip instead. One block per TILE of SMALL_S consecutive k-mer positions of one sequence:
//   1. hash     the threads roll ntHash (process_strip_dev, the hot loop of k_dense) over sub-strips of the tile plus
//               w-1 positions of halo on either side and leave all canonical hashes in shared memory;
//   2. windows  one thread per window: rightmost argmin of its w hashes, marked in a shared-memory byte map (the marked
//               positions are exactly the positions btllib emits: the argmin sequence is monotone, so "emit when the
//               argmin changes" and "emit every position that is some window's argmin" are the same set);
//   3. output   marks of the tile's own positions are counted, scanned and written in position order to the tile's
//               staging segment; k_small_gather packs the segments after a scan over the tile counts.
// Windows are defined over VALID k-mers and therefore span runs of N; a tile whose hashed range contains an invalid
// base cannot use the position arithmetic above and is walked serially by one thread with gap_scan (exact for any
// content) over a range widened to w-1 valid k-mers on either side. Rare in reads; in scaffolds one tile per N run.
#pragma once

namespace ntl {
namespace {

constexpr int SMALL_THREADS = 128;                                   // threads of a GROUP: a group works on one tile at a time
constexpr int SMALL_GROUPS = 2;                                      // groups per block; they share the roll table only
constexpr uint32_t SMALL_S = 4096;                                   // k-mer positions per tile
constexpr uint32_t SMALL_W_MAX = 16;
constexpr uint32_t SMALL_R = SMALL_S + 2 * (SMALL_W_MAX - 1);        // hashed positions per tile at most
constexpr size_t SMALL_GROUP_SMEM = (size_t)SMALL_R * 8 + 2 * (((size_t)SMALL_R + 15) & ~(size_t)15) + 32;     // H, F, M of one tile
constexpr size_t SMALL_SMEM = (size_t)ROLL_TABLE_ENTRIES * TBL_STRIDE + SMALL_GROUPS * SMALL_GROUP_SMEM;

// barrier of one group (named barrier 1 + group): the groups of a block never wait for each other after the table is loaded
__device__ __forceinline__ void small_group_sync(uint32_t grp) { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(SMALL_THREADS) : "memory"); }

struct SmallParams {
    uint32_t k, w, nseq, tcap, out_cap;
    uint64_t mult;
};

struct SmallEmit {                     // process_strip_dev<ALL> emitter: hashes and strands of a sub-strip into shared memory
    unsigned long long* H; uint8_t* F; uint32_t base;
    __device__ __forceinline__ void operator()(uint64_t h0, uint32_t pos, bool fwd, uint32_t) { H[pos - base] = h0; F[pos - base] = fwd ? 1 : 0; }
    __device__ __forceinline__ bool room_for_block() const { return true; }
    __device__ __forceinline__ void push(bool, uint64_t h0, uint32_t pos, bool fwd, uint32_t) { H[pos - base] = h0; F[pos - base] = fwd ? 1 : 0; }
    __device__ __forceinline__ void flush_block() {}
};

struct SmallSerialEmit {               // gap_scan emitter of a dirty tile: own positions straight into the staging segment
    uint64_t* hash; uint32_t* posf; uint32_t count, cap, lo, hi; uint64_t mult; bool raw;      // raw: the gather applies the second hash
    __device__ __forceinline__ void operator()(uint64_t h0, uint32_t pos, bool fwd) {
        if (pos < lo || pos >= hi) return;
        if (count < cap) { hash[count] = raw ? h0 : second_hash(h0, mult); posf[count] = pos | (fwd ? FWD_BIT : 0u); }
        count++;
    }
};

__global__ void __launch_bounds__(SMALL_THREADS * SMALL_GROUPS) k_small(const uint32_t* __restrict__ packed, const uint64_t* __restrict__ seq_off,
                                                         const uint32_t* __restrict__ strip_off, const uint32_t* __restrict__ strip_seq,
                                                         SmallParams P, const RollEntry* __restrict__ tbl_g, uint32_t* __restrict__ tile_cnt,
                                                         uint32_t* __restrict__ ticket, uint64_t* __restrict__ st_hash,
                                                         uint32_t* __restrict__ st_posf, SketchStatus* __restrict__ st) {
    extern __shared__ __align__(256) unsigned char sm_raw[];
    unsigned char* tbl_s = sm_raw;                                                      // entry e, copy c at e*256 + c*16 (as in k_dense)
    const uint32_t grp = threadIdx.x / SMALL_THREADS;
    unsigned long long* H = reinterpret_cast<unsigned long long*>(sm_raw + (size_t)ROLL_TABLE_ENTRIES * TBL_STRIDE + grp * SMALL_GROUP_SMEM);
    uint8_t* F = reinterpret_cast<uint8_t*>(H + SMALL_R);
    uint8_t* M = F + ((SMALL_R + 15u) & ~15u);
    __shared__ uint32_t s_part_all[SMALL_GROUPS][SMALL_THREADS / 32 + 1];
    __shared__ uint32_t s_tile_all[SMALL_GROUPS];
    __shared__ int s_dirty_all[SMALL_GROUPS];
    uint32_t* const s_part = s_part_all[grp];
    uint32_t& s_tile = s_tile_all[grp];
    int& s_dirty = s_dirty_all[grp];
    const uint32_t tid = threadIdx.x % SMALL_THREADS, lane = tid & 31, wid = tid >> 5;
    const uint32_t nstrips = strip_off[P.nseq];
    const uint32_t k = P.k, w = P.w;
    if (blockIdx.x == 0 && threadIdx.x == 0) st->nstrips = nstrips;
    for (uint32_t i = threadIdx.x; i < ROLL_TABLE_ENTRIES * TBL_COPIES; i += SMALL_THREADS * SMALL_GROUPS) {
        const RollEntry e = tbl_g[i / TBL_COPIES];
        reinterpret_cast<uint4*>(tbl_s)[i] = make_uint4((uint32_t)e.f, (uint32_t)(e.f >> 32), (uint32_t)e.r, (uint32_t)(e.r >> 32));
    }
    // persistent blocks: tiles are handed out by a ticket counter (tiles at sequence ends are partial, tiles with invalid
    // bases are slow), the roll table is loaded once per block
    __syncthreads();
  for (;;) {
    small_group_sync(grp);                                           // the previous tile's buffers are free
    if (tid == 0) { s_tile = atomicAdd(ticket, 1u); s_dirty = 0; }
    for (uint32_t i = tid; i < ((SMALL_R + 15u) & ~15u) / 4; i += SMALL_THREADS) reinterpret_cast<uint32_t*>(M)[i] = 0u;
    small_group_sync(grp);
    const uint32_t t = s_tile;
    if (t >= nstrips) return;
    const uint32_t q = strip_seq[t];
    const uint64_t gseq = seq_off[q];
    const uint32_t L = (uint32_t)(seq_off[q + 1] - gseq);
    const uint32_t np = L - k + 1;                                   // a tile exists only when seq_npos > 0
    const uint32_t p0 = (t - strip_off[q]) * SMALL_S;
    const uint32_t n = min(SMALL_S, np - p0);
    const uint32_t r_lo = p0 >= w - 1 ? p0 - (w - 1) : 0;            // hashed positions [r_lo, r_hi)
    const uint32_t r_hi = min(np, p0 + n + (w - 1));
    const uint32_t R = r_hi - r_lo;
    uint64_t* const my_hash = st_hash + (uint64_t)t * P.tcap;
    uint32_t* const my_posf = st_posf + (uint64_t)t * P.tcap;
    {   // any invalid base among the bases of the hashed range?
        const uint32_t nb = R + k - 1;
        bool dirty = false;
        for (uint32_t b = tid * 8; b < nb; b += SMALL_THREADS * 8) {
            uint32_t wd = fetch8(packed, gseq + r_lo + b);
            const uint32_t left = nb - b;
            if (left < 8) wd &= (1u << (4 * left)) - 1u;
            dirty |= (wd & 0x44444444u) != 0u;
        }
        if (dirty) s_dirty = 1;
    }
    small_group_sync(grp);
    if (s_dirty) {
        if (tid == 0) {
            // widen until the range holds w-1 valid k-mers on either side of the own positions (or reaches the sequence ends)
            uint32_t a = p0, b = p0 + n;
            {
                uint32_t have = 0, nb = NONE32;                      // nb = smallest invalid base index in [a, a + k), if any
                for (uint32_t j = 0; j < k && a + j < L; j++) if (fetch1(packed, gseq + a + j) >= CODE_INVALID) { nb = a + j; break; }
                while (a > 0 && have < w - 1) {
                    a--;
                    if (fetch1(packed, gseq + a) >= CODE_INVALID) nb = a;
                    if (nb == NONE32 || nb >= a + k) have++;
                }
            }
            {
                uint32_t have = 0, run = 0;                          // run = valid bases ending at base b + k - 1: k-mer b valid <=> run >= k
                for (uint32_t j = 0; j < k && b + j < L; j++) run = fetch1(packed, gseq + b + j) >= CODE_INVALID ? 0u : run + 1;
                while (b < np && have < w - 1) {
                    if (run >= k) have++;
                    b++;
                    run = (b + k - 1 < L && fetch1(packed, gseq + b + k - 1) < CODE_INVALID) ? run + 1 : 0u;
                }
            }
            SmallSerialEmit em{my_hash, my_posf, 0u, P.tcap, p0, p0 + n, P.mult, false};
            gap_scan(packed, tbl_g, gseq, L, k, w, a, b, em);
            tile_cnt[t] = em.count;
            if (em.count > P.tcap) atomicOr(&st->err, SKERR_OUT);
        }
        continue;
    }
    // ---- 1: hashes of [r_lo, r_hi) (all valid), sub-strips of equal length
    {
        const uint32_t per = (R + SMALL_THREADS - 1) / SMALL_THREADS;
        const uint32_t pa = min(R, tid * per), pb = min(R, pa + per);
        if (pb > pa) {
            SmallEmit em{H, F, r_lo};
            process_strip_dev<true>(packed, gseq, r_lo + pa, pb - pa, k, tbl_s, (tid & 15u) << 4, 0xFFFFFFFFu, em);
        }
    }
    small_group_sync(grp);
    // ---- 2: one thread per window of w consecutive positions inside the range: mark its rightmost argmin. The scan
    //         compares the HIGH hash words only (one 32-bit load per probe, unrolled by four); a tie of two high words
    //         sends the window through the full 64-bit scan (identical k-mers: low-complexity sequence).
    if (R >= w) {
        const uint32_t nwin = R - w + 1;
        const uint32_t* const Hw = reinterpret_cast<const uint32_t*>(H);
        // two windows per thread and step: two independent compare chains keep the few resident warps issuing
        for (uint32_t j = tid; j < nwin; j += 2 * SMALL_THREADS) {
            const uint32_t j2 = j + SMALL_THREADS;
            const bool two = j2 < nwin;
            const uint32_t* hp = Hw + 2 * j + 1;                     // high word of H[j + x] at hp[2 * x]
            const uint32_t* hq = Hw + 2 * (two ? j2 : j) + 1;
            uint32_t m = hp[0], a = 0, m2 = hq[0], a2 = 0;
            bool tie = false, tie2 = false;
            for (uint32_t x = 1; x < w; x++) {
                const uint32_t hv = hp[2 * x], hv2 = hq[2 * x];
                tie |= hv == m; tie2 |= hv2 == m2;
                if (hv <= m) { m = hv; a = x; }
                if (hv2 <= m2) { m2 = hv2; a2 = x; }
            }
            unsigned long long mv = H[j + a], mv2 = H[(two ? j2 : j) + a2];
            if (tie | tie2) {
                mv = H[j]; a = 0; mv2 = H[two ? j2 : j]; a2 = 0;
                for (uint32_t y = 1; y < w; y++) {
                    const unsigned long long hv = H[j + y], hv2 = H[(two ? j2 : j) + y];
                    if (hv <= mv) { mv = hv; a = y; }
                    if (hv2 <= mv2) { mv2 = hv2; a2 = y; }
                }
            }
            if (mv != 0xFFFFFFFFFFFFFFFFull) M[j + a] = 1;           // btllib never reports the all-ones hash (its "no minimizer" value)
            if (two && mv2 != 0xFFFFFFFFFFFFFFFFull) M[j2 + a2] = 1;
        }
    }
    small_group_sync(grp);
    // ---- 3: own marks -> staging segment, in position order. A warp owns a contiguous share of the own positions:
    //         pass 1 turns the byte map into ballot words (32 positions each, one word per lane at the end), a warp
    //         scan and a four-entry block scan place every word, pass 2 lets every lane write the marks of its word
    const uint32_t own0 = p0 - r_lo;                                 // index of the first own position in H / F / M
    const uint32_t wchunk = ((n + SMALL_THREADS - 1) / SMALL_THREADS) * 32;        // <= 1024 positions = 32 words per warp
    const uint32_t wa = min(n, wid * wchunk), wb = min(n, wa + wchunk);
    uint32_t myword = 0;
    for (uint32_t base = wa, it = 0; base < wb; base += 32, it++) {
        const uint32_t i = base + lane;
        const uint32_t bal = __ballot_sync(0xffffffffu, i < wb && M[own0 + i]);
        if (lane == it) myword = bal;
    }
    const uint32_t mine = __popc(myword);
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += y; }
    if (lane == 31) s_part[wid] = incl;
    small_group_sync(grp);
    if (tid == 0) {
        uint32_t run = 0;
        for (uint32_t i = 0; i < SMALL_THREADS / 32; i++) { const uint32_t x = s_part[i]; s_part[i] = run; run += x; }
        s_part[SMALL_THREADS / 32] = run;
        tile_cnt[t] = run;
        if (run > P.tcap) atomicOr(&st->err, SKERR_OUT);
    }
    small_group_sync(grp);
    if (s_part[SMALL_THREADS / 32] <= P.tcap) {
        uint32_t at = s_part[wid] + incl - mine;
        const uint32_t first = wa + lane * 32;                       // own position of bit 0 of this lane's word
        for (uint32_t bits = myword; bits; bits &= bits - 1) {
            const uint32_t i = first + (uint32_t)__ffs(bits) - 1;
            my_hash[at] = second_hash(H[own0 + i], P.mult);
            my_posf[at] = (p0 + i) | (F[own0 + i] ? FWD_BIT : 0u);
            at++;
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_stream: the same result with no tile-wide hash array. One THREAD per strip of STREAM_S positions (the strip table of the
// sparse path with a small S): the thread rolls over its strip plus w-1 positions of halo on either side and decides the
// windows while it rolls, with the block decomposition of the sliding-window minimum: positions are grouped in blocks of
// w; a window is a suffix of the previous block plus a prefix of the current one, so its rightmost argmin is one comparison
// between the running prefix minimum (registers) and the suffix minimum of the previous block (its position comes from a
// table of 16 nibbles in a register pair, its value from the thread's ring of the last w hashes in shared memory, which
// the current block has not overwritten yet at that offset). The phase inside a block is the same for every lane of a
// warp, so the only divergent part is writing the marks. Marks of the strip's own positions go straight to the strip's
// staging segment in position order (the argmin sequence is monotone: a mark is new iff its position differs from the
// previous window's). Strips whose hashed range contains an invalid base are walked with gap_scan, as in k_small.
constexpr uint32_t STREAM_S = 256;
constexpr int STREAM_THREADS = 128;
constexpr uint32_t STREAM_RING = 32;                                 // hashes kept per thread: the previous block, the current one, 8 not yet decided
// a clean strip only meets base codes 0..3 and the virtual leaving base 4: roll-table entries (in << 3 | out) <= 28. Loading just
// those leaves room for five blocks per SM instead of four.
constexpr uint32_t STREAM_TBL_ENTRIES = 29;
constexpr size_t STREAM_SMEM = (size_t)STREAM_TBL_ENTRIES * TBL_STRIDE + (size_t)STREAM_RING * STREAM_THREADS * 8;

// The unrolled rolling loop only stores (push); the window logic runs once per 8-step block over the stored positions
// (flush_block) so that it exists once in the instruction stream -- inlined into every unrolled step the kernel outgrew the
// instruction cache (measured: 3.9 issue slots lost per instruction to instruction fetch).
struct StreamEmit {
    unsigned long long* ring;          // slot e of this thread at ring[e * STREAM_THREADS]
    uint32_t w, head, pend, fbits;     // next slot to store; stored but undecided positions; strand bit of every slot
    uint32_t c, blk_pos, blk_slot;     // phase inside the current block of w positions; sequence position / slot of its first position
    unsigned long long pm;             // prefix minimum of the current block ...
    uint32_t pm_c;                     // ... and its (rightmost) offset
    unsigned long long sfx;            // nibble r = offset of the rightmost minimum of the previous block's suffix [r, w)
    uint32_t prev_slot, have_prev, last;   // slot of the previous block's first position; argmin position of the previous window
    uint32_t own_lo, own_hi;
    uint64_t* out_hash; uint32_t* out_posf; uint32_t cnt, cap;     // out_* advance with every mark written

    __device__ __forceinline__ void store(uint64_t h0, uint32_t fwd) {
        ring[head * STREAM_THREADS] = h0;
        fbits = (fbits & ~(1u << head)) | (fwd << head);
        head = (head + 1u) & (STREAM_RING - 1u);
        pend++;
    }
    __device__ __forceinline__ void operator()(uint64_t h0, uint32_t, bool fwd, uint32_t) { store(h0, fwd ? 1u : 0u); }
    __device__ __forceinline__ bool room_for_block() const { return true; }
    __device__ __forceinline__ void push(bool, uint64_t h0, uint32_t, bool fwd, uint32_t) { store(h0, fwd ? 1u : 0u); }
    __device__ __forceinline__ void flush_block() {
#pragma unroll 1
        for (; pend; pend--) {
            const uint32_t slot = (head - pend) & (STREAM_RING - 1u);
            const unsigned long long h0 = ring[slot * STREAM_THREADS];
            if (c == 0 || h0 <= pm) { pm = h0; pm_c = c; }
            {
                // rightmost argmin of the window that ends here: prefix minimum of this block against the suffix minimum of the
                // previous one (a tie goes to the later position). No branch on the lane's own data: almost every step some
                // lane of the warp has a new mark, so the emission is two predicated stores of the RAW hash (k_stream_gather
                // applies the second hash) instead of a divergent path.
                const bool havewin = (c == w - 1) | (have_prev != 0);
                unsigned long long wv = pm;
                uint32_t wpos = blk_pos + pm_c, wslot = (blk_slot + pm_c) & (STREAM_RING - 1u);
                if (c != w - 1) {                                    // warp-uniform
                    const uint32_t q = (uint32_t)(sfx >> (4u * (c + 1u))) & 15u;
                    const uint32_t qs = (prev_slot + q) & (STREAM_RING - 1u);
                    const unsigned long long sv = ring[qs * STREAM_THREADS];
                    const bool prev_wins = sv < pm;
                    wv = prev_wins ? sv : wv; wpos = prev_wins ? blk_pos - w + q : wpos; wslot = prev_wins ? qs : wslot;
                }
                const bool fresh = havewin & (wpos != last);
                last = fresh ? wpos : last;
                const bool emit = fresh & (wpos - own_lo < own_hi - own_lo) & (wv != 0xFFFFFFFFFFFFFFFFull);
                const bool put = emit & (cnt < cap);
                if (put) { *out_hash = wv; *out_posf = wpos | ((fbits >> wslot) << 31); }
                out_hash += put ? 1 : 0; out_posf += put ? 1 : 0;
                cnt += emit ? 1u : 0u;
            }
            if (c == w - 1) {                                        // block complete: its suffix minima; it becomes the previous block
                unsigned long long m = h0, tab = (unsigned long long)(w - 1) << (4u * (w - 1u));
                uint32_t idx = w - 1;
                for (uint32_t r = w - 1; r-- > 0;) {
                    const unsigned long long v = ring[((blk_slot + r) & (STREAM_RING - 1u)) * STREAM_THREADS];
                    if (v < m) { m = v; idx = r; }
                    tab |= (unsigned long long)idx << (4u * r);
                }
                sfx = tab; have_prev = 1; prev_slot = blk_slot; blk_slot = (blk_slot + w) & (STREAM_RING - 1u); blk_pos += w; c = 0;
            } else {
                c++;
            }
        }
    }
};

__global__ void __launch_bounds__(STREAM_THREADS) k_stream(const uint32_t* __restrict__ packed, const uint64_t* __restrict__ seq_off,
                                                           const uint32_t* __restrict__ strip_off, const uint32_t* __restrict__ strip_seq,
                                                           SmallParams P, const RollEntry* __restrict__ tbl_g, uint32_t* __restrict__ strip_cnt,
                                                           uint64_t* __restrict__ st_hash, uint32_t* __restrict__ st_posf,
                                                           SketchStatus* __restrict__ st) {
    extern __shared__ __align__(256) unsigned char sm_raw[];
    unsigned char* tbl_s = sm_raw;
    unsigned long long* ring = reinterpret_cast<unsigned long long*>(sm_raw + (size_t)STREAM_TBL_ENTRIES * TBL_STRIDE) + threadIdx.x;
    for (uint32_t i = threadIdx.x; i < STREAM_TBL_ENTRIES * TBL_COPIES; i += STREAM_THREADS) {
        const RollEntry e = tbl_g[i / TBL_COPIES];
        reinterpret_cast<uint4*>(tbl_s)[i] = make_uint4((uint32_t)e.f, (uint32_t)(e.f >> 32), (uint32_t)e.r, (uint32_t)(e.r >> 32));
    }
    __syncthreads();
    const uint32_t nstrips = strip_off[P.nseq];
    if (blockIdx.x == 0 && threadIdx.x == 0) st->nstrips = nstrips;
    const uint32_t k = P.k, w = P.w;
    uint32_t overflow = 0;
    for (uint32_t s = blockIdx.x * STREAM_THREADS + threadIdx.x; s < nstrips; s += gridDim.x * STREAM_THREADS) {
        const uint32_t q = strip_seq[s];
        const uint64_t gseq = seq_off[q];
        const uint32_t L = (uint32_t)(seq_off[q + 1] - gseq);
        const uint32_t np = L - k + 1;
        const uint32_t p0 = (s - strip_off[q]) * STREAM_S;
        const uint32_t n = min(STREAM_S, np - p0);
        const uint32_t r_lo = p0 >= w - 1 ? p0 - (w - 1) : 0;
        const uint32_t r_hi = min(np, p0 + n + (w - 1));
        const uint32_t R = r_hi - r_lo;
        uint64_t* const my_hash = st_hash + (uint64_t)s * P.tcap;
        uint32_t* const my_posf = st_posf + (uint64_t)s * P.tcap;
        bool dirty = false;
        {
            const uint32_t nb = R + k - 1;
            for (uint32_t b = 0; b < nb; b += 8) {
                uint32_t wd = fetch8(packed, gseq + r_lo + b);
                const uint32_t left = nb - b;
                if (left < 8) wd &= (1u << (4 * left)) - 1u;
                dirty |= (wd & 0x44444444u) != 0u;
            }
        }
        uint32_t count;
        if (dirty) {
            uint32_t a = p0, b = p0 + n;
            {
                uint32_t have = 0, nb = NONE32;
                for (uint32_t j = 0; j < k && a + j < L; j++) if (fetch1(packed, gseq + a + j) >= CODE_INVALID) { nb = a + j; break; }
                while (a > 0 && have < w - 1) {
                    a--;
                    if (fetch1(packed, gseq + a) >= CODE_INVALID) nb = a;
                    if (nb == NONE32 || nb >= a + k) have++;
                }
            }
            {
                uint32_t have = 0, run = 0;
                for (uint32_t j = 0; j < k && b + j < L; j++) run = fetch1(packed, gseq + b + j) >= CODE_INVALID ? 0u : run + 1;
                while (b < np && have < w - 1) {
                    if (run >= k) have++;
                    b++;
                    run = (b + k - 1 < L && fetch1(packed, gseq + b + k - 1) < CODE_INVALID) ? run + 1 : 0u;
                }
            }
            SmallSerialEmit em{my_hash, my_posf, 0u, P.tcap, p0, p0 + n, P.mult, true};
            gap_scan(packed, tbl_g, gseq, L, k, w, a, b, em);
            count = em.count;
        } else {
            StreamEmit em;
            em.ring = ring; em.w = w; em.head = 0; em.pend = 0; em.fbits = 0; em.c = 0; em.blk_pos = r_lo; em.blk_slot = 0; em.pm = 0; em.pm_c = 0;
            em.sfx = 0; em.prev_slot = 0; em.have_prev = 0; em.last = NONE32; em.own_lo = p0; em.own_hi = p0 + n;
            em.out_hash = my_hash; em.out_posf = my_posf; em.cnt = 0; em.cap = P.tcap;
            process_strip_dev<true>(packed, gseq, r_lo, R, k, tbl_s, (threadIdx.x & 15u) << 4, 0xFFFFFFFFu, em);
            count = em.cnt;
        }
        strip_cnt[s] = count;
        overflow |= count > P.tcap;
    }
    if (overflow) atomicOr(&st->err, SKERR_OUT);
}

// staging segments of k_stream -> packed output: one warp per strip
__global__ void __launch_bounds__(256) k_stream_gather(const uint32_t* __restrict__ strip_off, SmallParams P, const uint32_t* __restrict__ strip_cnt,
                                                       const uint32_t* __restrict__ strip_base, const uint64_t* __restrict__ st_hash,
                                                       const uint32_t* __restrict__ st_posf, uint64_t* __restrict__ out_hash,
                                                       uint32_t* __restrict__ out_posf, uint32_t* __restrict__ mx_off, SketchStatus* __restrict__ st,
                                                       CallState* __restrict__ call, uint32_t deferred) {
    const uint32_t nstrips = strip_off[P.nseq];
    const uint32_t total = strip_base[nstrips];
    const bool bad = st->err != 0 || total > P.out_cap;
    const uint32_t lane = threadIdx.x & 31, nwarp = gridDim.x * (blockDim.x / 32);
    if (!bad)
        for (uint32_t s = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); s < nstrips; s += nwarp) {
            const uint32_t n = strip_cnt[s], o = strip_base[s];
            const uint64_t from = (uint64_t)s * P.tcap;
            for (uint32_t i = lane; i < n; i += 32) { out_hash[o + i] = second_hash(st_hash[from + i], P.mult); out_posf[o + i] = st_posf[from + i]; }
        }
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q <= P.nseq; q += gridDim.x * blockDim.x)
        mx_off[q] = (bad && deferred) ? 0u : strip_base[q == P.nseq ? nstrips : strip_off[q]];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->n_mx = (bad && deferred) ? 0u : total;
        if (total > P.out_cap) atomicOr(&st->err, SKERR_OUT);
        if (bad && call) atomicOr(&call->err, CALLERR_SKETCH);
    }
}

// staging segments -> packed output, per-sequence offsets, totals, error gating (tile_base = exclusive scan of tile_cnt)
__global__ void __launch_bounds__(256) k_small_gather(const uint32_t* __restrict__ strip_off, SmallParams P, const uint32_t* __restrict__ tile_cnt,
                                                      const uint32_t* __restrict__ tile_base, const uint64_t* __restrict__ st_hash,
                                                      const uint32_t* __restrict__ st_posf, uint64_t* __restrict__ out_hash,
                                                      uint32_t* __restrict__ out_posf, uint32_t* __restrict__ mx_off, SketchStatus* __restrict__ st,
                                                      CallState* __restrict__ call, uint32_t deferred) {
    const uint32_t nstrips = strip_off[P.nseq];
    const uint32_t total = tile_base[nstrips];
    const bool bad = st->err != 0 || total > P.out_cap;
    if (!bad)
        for (uint32_t t = blockIdx.x; t < nstrips; t += gridDim.x) {
            const uint32_t n = tile_cnt[t], o = tile_base[t];
            const uint64_t from = (uint64_t)t * P.tcap;
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) { out_hash[o + i] = st_hash[from + i]; out_posf[o + i] = st_posf[from + i]; }
        }
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q <= P.nseq; q += gridDim.x * blockDim.x)
        mx_off[q] = (bad && deferred) ? 0u : tile_base[q == P.nseq ? nstrips : strip_off[q]];      // an empty sequence starts where the next one does
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->n_mx = (bad && deferred) ? 0u : total;
        if (total > P.out_cap) atomicOr(&st->err, SKERR_OUT);
        if (bad && call) atomicOr(&call->err, CALLERR_SKETCH);
    }
}

}  // namespace
}  // namespace ntl
