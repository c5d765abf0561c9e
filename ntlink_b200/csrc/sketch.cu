// sketch.cu -- the minimizer sketch pipeline on the device (replaces btllib `indexlr --long --pos --strand`,
// invoked by the reference at ntLink:198-199 and ntLink:221-225). See sketch_logic.cuh for the algorithm.
//
// Kernels (all hand-written, sm_100a):
//   k_pack        ASCII -> 4-bit codes, 128-bit coalesced loads, 64-bit stores      (HBM streaming)
//   k_strip_count strips per sequence                                               (tiny)
//   k_dense       one thread per strip of S k-mer positions: rolling ntHash, candidate filter  (THE hot kernel,
//                 integer-ALU bound; roll table replicated 8x in shared memory for conflict-free LDS.128)
//   k_overflow    re-runs the rare strips whose candidates did not fit their slots
//   k_select      per candidate: exact minimizer decision from neighbouring candidates; queues gaps
//   k_seq_gaps    leading / whole-sequence candidate-free stretches
//   k_gap         exact sliding-window re-scan of the queued stretches
//   k_emit        ordered compaction: second hash, pos|strand, per-sequence offsets
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace ntl {

namespace {

// ------------------------------------------------------------------------------------------- pack
// four ASCII bases (one 32-bit word) -> four 4-bit codes in the low 16 bits: A/a=0 C/c=1 G/g=2 T/t=3, anything else 4.
// SIMD within the register: (x>>1 ^ x>>2) & 3 is the classic ACGT code; the byte is valid iff it equals the letter that
// code stands for (A 0x41, C 0x43, G 0x47, T 0x54 after case folding), checked for all four bytes at once.
__device__ __forceinline__ uint32_t pack4(uint32_t w) {
    const uint32_t x = w & 0xDFDFDFDFu;
    const uint32_t c2 = ((x >> 1) ^ (x >> 2)) & 0x03030303u;
    const uint32_t b0 = c2 & 0x01010101u, b1 = (c2 >> 1) & 0x01010101u;
    const uint32_t expect = 0x41414141u + 2u * c2 + 2u * b1 + 11u * (b0 & b1);
    const uint32_t d = x ^ expect;                                           // non-zero byte <=> invalid base
    const uint32_t nz = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
    const uint32_t m = nz >> 7;                                              // 0x01 per invalid byte
    const uint32_t code = (c2 & ~(m * 3u)) | (m << 2);
    const uint32_t t = code | (code >> 4);                                   // byte0 = c0|c1<<4, byte2 = c2|c3<<4
    return __byte_perm(t, 0u, 0x4420);
}

constexpr int PACK_CHUNKS = 4;   // 16-byte chunks per thread, loaded up front (memory-level parallelism)

__global__ void __launch_bounds__(256) k_pack(const uint8_t* __restrict__ seq, uint64_t nbases,
                                              uint32_t* __restrict__ packed /* 8 bases per word */) {
    const uint64_t chunk0 = (uint64_t)blockIdx.x * (blockDim.x * PACK_CHUNKS) + threadIdx.x;
    uint4 v[PACK_CHUNKS];
#pragma unroll
    for (int i = 0; i < PACK_CHUNKS; i++) {
        const uint64_t b0 = (chunk0 + (uint64_t)i * blockDim.x) * 16;
        if (b0 + 16 <= nbases) v[i] = __ldg(reinterpret_cast<const uint4*>(seq + b0));
        else {
            uint32_t w[4] = {0x4E4E4E4Eu, 0x4E4E4E4Eu, 0x4E4E4E4Eu, 0x4E4E4E4Eu};   // 'N'
            for (int j = 0; j < 16; j++) {
                const uint64_t b = b0 + j;
                if (b < nbases) w[j >> 2] = (w[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | ((uint32_t)seq[b] << (8 * (j & 3)));
            }
            v[i] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
#pragma unroll
    for (int i = 0; i < PACK_CHUNKS; i++) {
        const uint64_t chunk = chunk0 + (uint64_t)i * blockDim.x;
        if (chunk * 16 >= nbases) continue;
        const uint32_t o0 = pack4(v[i].x) | (pack4(v[i].y) << 16);
        const uint32_t o1 = pack4(v[i].z) | (pack4(v[i].w) << 16);
        *reinterpret_cast<uint2*>(packed + chunk * 2) = make_uint2(o0, o1);
    }
}

// ------------------------------------------------------------------------------------------- strips
struct SkParams {
    uint32_t k, w, S, cap, tau_hi, nseq;
    uint64_t mult;
    uint64_t pool_base;     // = nstrips_max * cap
    uint32_t pool_cap, gaps_cap, extras_cap, out_cap;
};

__device__ __forceinline__ uint32_t seq_npos(uint64_t L, uint32_t k, uint32_t w) {
    // btllib: k > L or w > L-k+1 -> nothing
    if (L < k) return 0;
    const uint64_t np = L - k + 1;
    return np < w ? 0u : (uint32_t)np;
}

__global__ void k_strip_count(const uint64_t* __restrict__ seq_off, uint32_t nseq, uint32_t k, uint32_t w, uint32_t S,
                              uint32_t* __restrict__ scnt, uint32_t* __restrict__ nseq_dev) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q == 0) *nseq_dev = nseq;
    if (q >= nseq) return;
    const uint32_t np = seq_npos(seq_off[q + 1] - seq_off[q], k, w);
    scnt[q] = (np + S - 1) / S;
}

// largest q with strip_off[q] <= s  (strip_off has nseq+1 entries, strip_off[nseq] = nstrips > s)
__device__ __forceinline__ uint32_t seq_of_strip(const uint32_t* __restrict__ strip_off, uint32_t nseq, uint32_t s) {
    uint32_t lo = 0, hi = nseq;           // invariant: strip_off[lo] <= s < strip_off[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (strip_off[mid] <= s) lo = mid; else hi = mid;
    }
    return lo;
}

// strip -> sequence table, computed once per batch (k_dense, k_overflow and k_select then need one load instead of a
// 14-step dependent binary search per thread)
__global__ void k_strip_seq(const uint32_t* __restrict__ strip_off, uint32_t nseq, uint32_t* __restrict__ strip_seq) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= strip_off[nseq]) return;
    strip_seq[s] = seq_of_strip(strip_off, nseq, s);
}

struct SlotEmit {
    uint4* base;       // slot array of the strip (a Cand is exactly one uint4: h0.lo, h0.hi, posf, lord)
    uint32_t cap;
    uint32_t count;    // keeps counting past cap so that the strip's total is known
    // checked form (generic blocks)
    __device__ __forceinline__ void operator()(uint64_t h0, uint32_t pos, bool fwd, uint32_t lord) {
        if (count < cap) base[count] = make_uint4((uint32_t)h0, (uint32_t)(h0 >> 32), pos | (fwd ? FWD_BIT : 0u), lord);
        count++;
    }
    // unchecked form: the caller guarantees room for the whole 8-step block
    __device__ __forceinline__ void fast(uint64_t h0, uint32_t pos, bool fwd, uint32_t lord) {
        base[count] = make_uint4((uint32_t)h0, (uint32_t)(h0 >> 32), pos | (fwd ? FWD_BIT : 0u), lord);
        count++;
    }
    __device__ __forceinline__ bool room_for_block() const { return count + 8 <= cap; }
    // interior blocks: the caller guaranteed room for 8, so the candidate test costs one predicated 128-bit store and a
    // predicated increment instead of a divergent branch (some lane of a warp finds a candidate at almost every step)
    __device__ __forceinline__ void push(bool c, uint64_t h0, uint32_t pos, bool fwd, uint32_t lord) {
        if (c) base[count] = make_uint4((uint32_t)h0, (uint32_t)(h0 >> 32), pos | (fwd ? FWD_BIT : 0u), lord);
        count += c ? 1u : 0u;
    }
    __device__ __forceinline__ void flush_block() {}
};

// ---- the hot loop, device-only formulation --------------------------------------------------------------------
// Same arithmetic as process_strip (sketch_logic.cuh) but shaped for the integer pipes of an SM:
//  * hash state in four 32-bit registers; the split 33/31-bit rotations are 3-input LOP3s and funnel shifts
//    (7 ALU ops per hash per base);
//  * roll table in shared memory with a 256-byte entry stride and 16 interleaved copies: the LDS.128 address of an
//    entry is produced by ONE byte-permute (PRMT) from a per-block word of precomputed (in<<3|out) bytes and the
//    lane's copy offset, and a quarter-warp always touches 8 different 16-byte bank groups (conflict-free);
//  * lead-in blocks never test for candidates, interior blocks carry no bounds or validity checks; everything
//    irregular (sequence tail, first output block, blocks containing N) goes through the generic block.
struct H32 { uint32_t flo, fhi, rlo, rhi; };

template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
}

// one base: fh' = srol(fh) ^ t.f ; rh' = sror(rh ^ t.r) on 32-bit halves, 7 integer ops per strand.
// LOP3 truth tables (A=0xF0, B=0xCC, C=0xAA): 0x6A = (A&B)^C, 0x28 = (A^B)&C, 0x96 = A^B^C, 0xD8 = (A&~C)|(B&C), 0xF8 = A|(B&C)
// (Tried: writing the shifts as mul/mulhi so that they issue on the idle FMA pipe -- IMAD.HI is slow enough that it
// was a net loss on B200; the funnel-shift form below is the fastest measured.)
__device__ __forceinline__ void roll32(H32& h, const uint4 t) {
    const uint32_t a = h.flo + h.flo;                              // bits 1..31 of the low word
    const uint32_t b = __funnelshift_l(h.flo, h.fhi, 1);           // high word shifted, bit 31 of lo carried in
    const uint32_t c = h.fhi >> 30;                                // bit63 lands on bit 1
    const uint32_t lo = a ^ lop3<0x6A>(h.fhi, 1u, t.x);            // bit32 -> bit0, xor table
    const uint32_t hi = lop3<0x96>(b, lop3<0x28>(b, c, 2u), t.y);  // bit63 -> bit33, xor table
    const uint32_t xlo = h.rlo ^ t.z, xhi = h.rhi ^ t.w;
    const uint32_t rl = __funnelshift_r(xlo, xhi, 1);
    const uint32_t u = lop3<0xD8>(xhi >> 1, xlo, 1u);              // bit0 -> bit32
    const uint32_t rh = lop3<0xF8>(u, xhi << 30, 0x80000000u);     // bit33 -> bit63
    h.flo = lo; h.fhi = hi; h.rlo = rl; h.rhi = rh;
}

enum : uint32_t { TBL_STRIDE = 256, TBL_COPIES = 16 };     // bytes per entry, copies per entry

__device__ __forceinline__ uint4 tbl_fetch(const unsigned char* tbl_s, uint32_t comb, uint32_t lanebase, uint32_t sel) {
    const uint32_t off = __byte_perm(comb, lanebase, sel); // (entry << 8) | (copy << 4)
    return *reinterpret_cast<const uint4*>(tbl_s + off);
}

// generic block (8 steps) with bounds and validity checks; identical to the slow branch of process_strip
template <bool ALL = false, class Emit>
__device__ __forceinline__ void generic_block(H32& h, uint32_t wi, uint32_t wo, int32_t t, int32_t lead, int32_t T, int32_t k,
                                              int32_t& last_bad, uint32_t& nv, uint32_t p0, uint32_t tau_hi,
                                              const unsigned char* tbl_s, uint32_t lanebase, Emit& emit) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int32_t s = t + j;
        const uint32_t cin = (wi >> (4 * j)) & 7u;
        const uint32_t e = (cin << 3) | ((wo >> (4 * j)) & 7u);
        roll32(h, *reinterpret_cast<const uint4*>(tbl_s + e * TBL_STRIDE + lanebase));
        if (cin >= CODE_INVALID) last_bad = s;
        if (s >= lead && s < T && s - last_bad >= k) {
            const uint64_t fh = ((uint64_t)h.fhi << 32) | h.flo, rh = ((uint64_t)h.rhi << 32) | h.rlo;
            const uint64_t h0 = fh + rh;
            if (ALL || (uint32_t)(h0 >> 32) < tau_hi) emit(h0, p0 + (uint32_t)(s - lead), fh <= rh, nv);
            nv++;
        }
    }
}

// five consecutive 32-bit words starting at word r (0..3) of the 8-word window {A, B}
__device__ __forceinline__ void pick5(const uint4 A, const uint4 B, uint32_t r, uint32_t (&x)[5]) {
    switch (r) {
        case 0: x[0] = A.x; x[1] = A.y; x[2] = A.z; x[3] = A.w; x[4] = B.x; break;
        case 1: x[0] = A.y; x[1] = A.z; x[2] = A.w; x[3] = B.x; x[4] = B.y; break;
        case 2: x[0] = A.z; x[1] = A.w; x[2] = B.x; x[3] = B.y; x[4] = B.z; break;
        default: x[0] = A.w; x[1] = B.x; x[2] = B.y; x[3] = B.z; x[4] = B.w; break;
    }
}

// The packed sequence is read with 128-bit loads, one per 32 bases and stream (entering / leaving bases): lanes of a
// warp work on strips that are 128 bytes apart, so every load instruction costs 32 L1 wavefronts whatever its
// width -- 16-byte loads cut the wavefront count (the limiter of this kernel before) by 4x compared to words.
// `packed` must be preceded by one readable 16-byte chunk (the leaving stream starts k bases before the strip).
// ALL = true: every valid k-mer is handed to the emitter (the dense-mode kernel for small windows), tau_hi is ignored.
template <bool ALL = false, class Emit>
__device__ __forceinline__ uint32_t process_strip_dev(const uint32_t* __restrict__ packed, uint64_t gseq, uint32_t p0, uint32_t n,
                                                      uint32_t k, const unsigned char* tbl_s, uint32_t lanebase,
                                                      uint32_t tau_hi, Emit& emit) {
    const uint64_t g0 = gseq + p0;
    const int32_t T = (int32_t)(k - 1 + n);
    const int32_t lead = (int32_t)k - 1;
    H32 h = {0, 0, 0, 0};
    int32_t last_bad = -(1 << 30);
    uint32_t nv = 0;
    const uint4* __restrict__ chunks = reinterpret_cast<const uint4*>(packed);
    // entering bases: base index g0 + t
    const int64_t ci_in = (int64_t)(g0 >> 5);
    const uint32_t r_in = (uint32_t)(g0 >> 3) & 3u, sh_in = (uint32_t)(g0 & 7) * 4;
    // leaving bases: base index g0 + t - k (virtual "zero" bases while t < k)
    const int64_t go = (int64_t)g0 - (int64_t)k;
    const int64_t ci_out = go >> 5;                                   // arithmetic shift: floor
    const uint32_t r_out = (uint32_t)(go >> 3) & 3u, sh_out = (uint32_t)(go & 7) * 4;
    uint4 Ain = chunks[ci_in];
    uint4 Aout = make_uint4(0, 0, 0, 0);
    bool out_live = false;
    for (int32_t t0 = 0; t0 < T; t0 += 32) {
        const int32_t sb = t0 >> 5;
        uint32_t xi[5], xo[5] = {0, 0, 0, 0, 0};
        const uint4 Bin = chunks[ci_in + sb + 1];
        pick5(Ain, Bin, r_in, xi);
        Ain = Bin;
        const bool need_out = t0 + 32 > (int32_t)k;                   // some real leaving base in this super-block
        if (need_out) {
            if (!out_live) { Aout = chunks[ci_out + sb]; out_live = true; }
            const uint4 Bout = chunks[ci_out + sb + 1];
            pick5(Aout, Bout, r_out, xo);
            Aout = Bout;
        }
        // the four 8-step blocks share ONE copy of the unrolled step code (keeps the kernel inside the instruction cache)
#pragma unroll 1
        for (int b = 0; b < 4; b++) {
            const int32_t t = t0 + 8 * b;
            if (t >= T) break;
            uint32_t xa, xb, ya, yb;
            switch (b) {
                case 0: xa = xi[0]; xb = xi[1]; ya = xo[0]; yb = xo[1]; break;
                case 1: xa = xi[1]; xb = xi[2]; ya = xo[1]; yb = xo[2]; break;
                case 2: xa = xi[2]; xb = xi[3]; ya = xo[2]; yb = xo[3]; break;
                default: xa = xi[3]; xb = xi[4]; ya = xo[3]; yb = xo[4]; break;
            }
            const uint32_t wi = __funnelshift_r(xa, xb, sh_in);
            uint32_t wo;
            const int32_t o = t - (int32_t)k;
            if (o <= -8) wo = 0x44444444u;
            else {
                wo = __funnelshift_r(ya, yb, sh_out);
                if (o < 0) {                                          // the first -o steps still push out virtual bases
                    const uint32_t m = (1u << (4u * (uint32_t)(-o))) - 1u;
                    wo = (wo & ~m) | (0x44444444u & m);
                }
            }
            const bool clean = ((wi & 0x44444444u) == 0u);
            if (clean && t + 8 <= lead) {
                // lead-in: no k-mer completes in this block -> roll only
                const uint32_t ce = (wi & 0x0F0F0F0Fu) * 8u + (wo & 0x0F0F0F0Fu);
                const uint32_t co = ((wi >> 4) & 0x0F0F0F0Fu) * 8u + ((wo >> 4) & 0x0F0F0F0Fu);
#pragma unroll
                for (int j = 0; j < 8; j++)
                    roll32(h, tbl_fetch(tbl_s, (j & 1) ? co : ce, lanebase, 0x7604u | ((uint32_t)(j >> 1) << 4)));
            } else if (clean && t >= lead && t + 8 <= T && t - last_bad >= (int32_t)k && emit.room_for_block()) {
                // interior: 8 valid, in-range k-mers, and room for 8 candidates
                const uint32_t ce = (wi & 0x0F0F0F0Fu) * 8u + (wo & 0x0F0F0F0Fu);
                const uint32_t co = ((wi >> 4) & 0x0F0F0F0Fu) * 8u + ((wo >> 4) & 0x0F0F0F0Fu);
                const uint32_t pos0 = p0 + (uint32_t)(t - lead);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    roll32(h, tbl_fetch(tbl_s, (j & 1) ? co : ce, lanebase, 0x7604u | ((uint32_t)(j >> 1) << 4)));
                    const uint64_t fh = ((uint64_t)h.fhi << 32) | h.flo, rh = ((uint64_t)h.rhi << 32) | h.rlo;
                    const uint64_t h0 = fh + rh;
                    emit.push(ALL || (uint32_t)(h0 >> 32) < tau_hi, h0, pos0 + j, fh <= rh, nv + j);
                }
                if (!ALL) emit.flush_block();
                nv += 8;
            } else {
                generic_block<ALL>(h, wi, wo, t, lead, T, (int32_t)k, last_bad, nv, p0, tau_hi, tbl_s, lanebase, emit);
            }
            if (ALL) emit.flush_block();                              // one copy of the emitter's per-block work for all three kinds of block
        }
    }
    return nv;
}

__global__ void __launch_bounds__(128) k_dense(const uint32_t* __restrict__ packed, const uint64_t* __restrict__ seq_off,
                                               const uint32_t* __restrict__ strip_off, const uint32_t* __restrict__ strip_seq, SkParams P,
                                               const RollEntry* __restrict__ tbl_g, Cand* __restrict__ slots,
                                               uint32_t* __restrict__ cnt, uint32_t* __restrict__ nv,
                                               uint8_t* __restrict__ has_cand, SketchStatus* __restrict__ st) {
    // entry e, copy c at byte offset e*256 + c*16
    __shared__ __align__(256) unsigned char tbl_s[ROLL_TABLE_ENTRIES * TBL_STRIDE];
    for (uint32_t i = threadIdx.x; i < ROLL_TABLE_ENTRIES * TBL_COPIES; i += blockDim.x) {
        const RollEntry e = tbl_g[i / TBL_COPIES];
        reinterpret_cast<uint4*>(tbl_s)[i] = make_uint4((uint32_t)e.f, (uint32_t)(e.f >> 32), (uint32_t)e.r, (uint32_t)(e.r >> 32));
    }
    __syncthreads();
    const uint32_t nstrips = strip_off[P.nseq];
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) st->nstrips = nstrips;
    if (s >= nstrips) return;
    const uint32_t q = strip_seq[s];
    const uint64_t gseq = seq_off[q];
    const uint32_t np = seq_npos(seq_off[q + 1] - gseq, P.k, P.w);
    const uint32_t p0 = (s - strip_off[q]) * P.S;
    const uint32_t n = min(P.S, np - p0);
    SlotEmit em{reinterpret_cast<uint4*>(slots + (uint64_t)s * P.cap), P.cap, 0u};
    const uint32_t nvalid = process_strip_dev(packed, gseq, p0, n, P.k, tbl_s, (threadIdx.x & 15u) << 4, P.tau_hi, em);
    const uint32_t count = em.count;
    cnt[s] = count;
    nv[s] = nvalid;
    if (count) has_cand[q] = 1;
    if (count > P.cap) atomicAdd(&st->n_ovf, 1u);
}

struct PoolEmit {
    Cand* dst;
    uint32_t count;
    __device__ __forceinline__ void operator()(uint64_t h0, uint32_t pos, bool fwd, uint32_t lord) {
        Cand c; c.h0 = h0; c.posf = pos | (fwd ? FWD_BIT : 0u); c.lord = lord;
        dst[count++] = c;
    }
};

// strips whose candidate count exceeded the slot capacity: reserve room in the pool and run them again
__global__ void __launch_bounds__(128) k_overflow(const uint32_t* __restrict__ packed, const uint64_t* __restrict__ seq_off,
                                                  const uint32_t* __restrict__ strip_off, const uint32_t* __restrict__ strip_seq, SkParams P,
                                                  const RollEntry* __restrict__ tbl_g, Cand* __restrict__ cands,
                                                  const uint32_t* __restrict__ cnt, uint32_t* __restrict__ ovf_off,
                                                  SketchStatus* __restrict__ st) {
    if (st->n_ovf == 0) return;
    const uint32_t nstrips = st->nstrips;
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nstrips || cnt[s] <= P.cap) return;
    const uint32_t off = atomicAdd(&st->pool_used, cnt[s]);
    if ((uint64_t)off + cnt[s] > P.pool_cap) { atomicOr(&st->err, SKERR_POOL); ovf_off[s] = 0; return; }
    ovf_off[s] = off;
    const uint32_t q = strip_seq[s];
    const uint64_t gseq = seq_off[q];
    const uint32_t np = seq_npos(seq_off[q + 1] - gseq, P.k, P.w);
    const uint32_t p0 = (s - strip_off[q]) * P.S;
    const uint32_t n = min(P.S, np - p0);
    PoolEmit em{cands + P.pool_base + off, 0};
    process_strip(packed, gseq, p0, n, P.k, tbl_g, 1, P.tau_hi, em);
}

// ------------------------------------------------------------------------------------------- select
__device__ __forceinline__ void queue_gap(GapRec* gaps, uint32_t* gap_head, SketchStatus* st, const SkParams& P,
                                          uint32_t seq, uint32_t start_pos, uint32_t end_pos, uint32_t strip,
                                          uint32_t j, uint32_t nvalid) {
    const uint32_t id = atomicAdd(&st->ngaps, 1u);
    if (id >= P.gaps_cap) { atomicOr(&st->err, SKERR_GAPS); return; }
    const uint32_t need = nvalid - P.w + 1;
    const uint32_t off = atomicAdd(&st->extras_used, need);
    GapRec g;
    g.seq = seq; g.start_pos = start_pos; g.end_pos = end_pos; g.strip = strip; g.j = j;
    g.out_off = off; g.out_cnt = 0; g.max_out = need; g.pad = 0;
    if ((uint64_t)off + need > P.extras_cap) { atomicOr(&st->err, SKERR_EXTRAS); g.max_out = 0; }
    g.next = atomicExch(&gap_head[strip], id);
    gaps[id] = g;
}

}  // namespace
}  // namespace ntl
#include "select_kernel.cuh"   // k_select (block-cooperative, shared-memory neighbour scans)
namespace ntl {
namespace {

// candidate-free stretch at the start of a sequence (or the whole sequence)
__global__ void k_seq_gaps(const uint64_t* __restrict__ seq_off, const uint32_t* __restrict__ strip_off, SkParams P,
                           CandView V, const uint8_t* __restrict__ has_cand, GapRec* __restrict__ gaps,
                           uint32_t* __restrict__ gap_head, SketchStatus* __restrict__ st) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.nseq) return;
    const uint32_t fs = strip_off[q], es = strip_off[q + 1];
    if (fs == es) return;
    const uint32_t np = seq_npos(seq_off[q + 1] - seq_off[q], P.k, P.w);
    const uint32_t idx0 = V.vbase[fs];
    const uint32_t nvalid = V.vbase[es] - idx0;
    if (nvalid < P.w) return;
    if (!has_cand[q]) { queue_gap(gaps, gap_head, st, P, q, 0, np, fs, NONE32, nvalid); return; }
    uint32_t fc = fs;                               // first candidate of the sequence
    while (fc < es && V.cnt[fc] == 0) fc++;
    if (fc >= es) return;   // cannot happen when has_cand is set
    const Cand c = V.cands[cand_gid(V, fc, 0)];
    const uint32_t lead = V.vbase[fc] + c.lord - idx0;
    if (lead >= P.w) queue_gap(gaps, gap_head, st, P, q, 0, c.posf & POS_MASK, fs, NONE32, lead);
}

}  // namespace
}  // namespace ntl
#include "gap_kernel.cuh"      // k_gap (one warp per candidate-free stretch)
#include "tile_kernel.cuh"
#include "small_kernel.cuh"     // k_tile (the single-pass sketch for w >= 13)
namespace ntl {
namespace {

// ------------------------------------------------------------------------------------------- emit
// G lanes per strip (G = 4, 8, 16 or 32, picked from the expected number of candidates per strip): the lanes of a group
// take the strip's candidates G at a time (coalesced 16-byte loads), the selected ones are ranked with a ballot and
// written next to each other. Strips with a re-scanned gap attached or with overflowed slots take the serial path on
// the group's first lane.
constexpr int EMIT_THREADS = 256;
template <int G>
__global__ void __launch_bounds__(EMIT_THREADS) k_emit(SkParams P, CandView V, const uint8_t* __restrict__ sel,
                                                       const unsigned long long* __restrict__ selmask,
                                                       const uint32_t* __restrict__ selbase, const GapRec* __restrict__ gaps,
                                                       const uint32_t* __restrict__ gap_head, const Cand* __restrict__ extras,
                                                       uint64_t* __restrict__ out_hash, uint32_t* __restrict__ out_posf,
                                                       SketchStatus* __restrict__ st) {
    const uint32_t nstrips = st->nstrips;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t s = gtid / G;
    const uint32_t lane = threadIdx.x & 31, gl = lane % G, gshift = lane - gl;
    const uint32_t gmask = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
    if (gtid == 0) st->n_mx = selbase[nstrips];
    const bool live = s < nstrips;
    // everything indexed by the strip is fetched up front: one memory latency instead of a chain of them
    uint32_t o = 0, end = 0, head = NONE32, c = 0;
    unsigned long long m = 0;
    if (live) { o = selbase[s]; end = selbase[s + 1]; head = gap_head[s]; c = V.cnt[s]; m = selmask[s]; }
    bool fast = live && o != end;
    if (fast && end > P.out_cap) { if (gl == 0) atomicOr(&st->err, SKERR_OUT); fast = false; c = 0; }
    const bool serial = fast && !(head == NONE32 && c <= V.cap);
    if (!fast || serial) c = 0;
    const uint64_t base = (uint64_t)s * V.cap;
    for (uint32_t j0 = 0; __any_sync(0xffffffffu, j0 < c); j0 += G) {
        const uint32_t j = j0 + gl;
        bool on = false;
        if (j < c) on = j < 64 ? ((m >> j) & 1ull) != 0 : sel[base + j] != 0;
        const uint32_t b = (__ballot_sync(0xffffffffu, on) >> gshift) & gmask;
        if (on) {
            const Cand e = V.cands[base + j];
            const uint32_t at = o + __popc(b & ((1u << gl) - 1u));
            out_hash[at] = second_hash(e.h0, P.mult); out_posf[at] = e.posf;
        }
        o += __popc(b);
    }
    if (!serial || gl != 0) return;
    c = V.cnt[s];
    if (head != NONE32) {
        for (uint32_t g = head; g != NONE32; g = gaps[g].next)
            if (gaps[g].j == NONE32)
                for (uint32_t i = 0; i < gaps[g].out_cnt; i++) {
                    const Cand e = extras[gaps[g].out_off + i];
                    out_hash[o] = second_hash(e.h0, P.mult); out_posf[o] = e.posf; o++;
                }
    }
    for (uint32_t j = 0; j < c; j++) {
        const uint64_t gid = cand_gid(V, s, j);
        if (sel[gid]) {
            const Cand e = V.cands[gid];
            out_hash[o] = second_hash(e.h0, P.mult); out_posf[o] = e.posf; o++;
        }
        if (head != NONE32) {
            for (uint32_t g = head; g != NONE32; g = gaps[g].next)
                if (gaps[g].j == j)
                    for (uint32_t i = 0; i < gaps[g].out_cnt; i++) {
                        const Cand e = extras[gaps[g].out_off + i];
                        out_hash[o] = second_hash(e.h0, P.mult); out_posf[o] = e.posf; o++;
                    }
        }
    }
}

// call != null (deferred pass): a sketch that hit a workspace limit is turned into an empty one and the call is flagged
__global__ void k_seq_offsets(const uint32_t* __restrict__ strip_off, const uint32_t* __restrict__ selbase,
                              uint32_t nseq, uint32_t* __restrict__ mx_off, SketchStatus* __restrict__ st,
                              CallState* __restrict__ call) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    const bool bad = call != nullptr && st->err != 0;
    if (q <= nseq) mx_off[q] = bad ? 0u : selbase[strip_off[q]];
    if (q == 0 && bad) { atomicOr(&call->err, CALLERR_SKETCH); st->n_mx = 0; }
}

// status block + number of strips + number of minimizers -> host-mapped memory
__global__ void k_publish_sketch(const SketchStatus* __restrict__ st, const uint32_t* __restrict__ nstrips_p,
                                 const uint32_t* __restrict__ selbase, uint32_t* __restrict__ host) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(st);
    if (threadIdx.x < sizeof(SketchStatus) / 4) host[threadIdx.x] = src[threadIdx.x];
    if (threadIdx.x == 0) { const uint32_t ns = *nstrips_p; host[16] = ns; host[17] = selbase ? selbase[ns] : st->n_mx; }
    __threadfence_system();
}

__global__ void k_fill_u32(uint32_t* p, uint32_t v, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}

inline uint32_t div_up(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

}  // namespace

int sketch_prepare(ntl_ctx* c, uint32_t k) {
    SketchWork& W = c->sw;
    if (W.tbl_k == k && W.tbl.p) return NTL_OK;
    if (c->capturing) { c->err = "internal: rolling table not prepared before capture"; return NTL_ERR_STATE; }
    NTL_CUDA(c, W.tbl.ensure(sizeof(RollEntry) * ROLL_TABLE_ENTRIES));
    RollEntry tbl[ROLL_TABLE_ENTRIES];
    build_roll_table(k, tbl);
    NTL_CUDA(c, cudaMemcpyAsync(W.tbl.p, tbl, sizeof tbl, cudaMemcpyHostToDevice, c->stream));
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));   // tbl is a stack array
    W.tbl_k = k;
    return NTL_OK;
}

// Shape of the single-pass path for a window size: strip length S (largest of 256 / 128 / 64 whose expected number of
// candidates per strip fits a slot row), slot row capacity (odd: conflict-free shared-memory rows), halo strips.
struct TileShape { bool ok; uint32_t S, cap, H, flat_cap, shared_at; size_t smem; };
static TileShape tile_shape(uint32_t k, uint32_t w, double cand_c) {
    TileShape t{false, 0, 0, 0, 0};
    if (w < 2 || k > 4096) return t;
    for (uint32_t S : {256u, 128u, 64u}) {
        const double mu = (double)S * cand_c / (double)w;
        if (mu > 36.0) continue;
        const uint32_t cap = 0;
        const uint32_t H = (w - 1 + S - 1) / S;
        if (H < 1 || 2 * H > TILE_THREADS / 2) continue;
        t.ok = true; t.S = S; t.cap = cap; t.H = H;
        const double view = TILE_THREADS * mu;
        t.flat_cap = ((uint32_t)(view + 5.0 * sqrt(view) + 48.0) + 3u) & ~3u;     // pool and flat list: mean + 5 sigma, no per-strip slack
        if (t.flat_cap > 16384 || S > 256) continue;
        size_t at = (size_t)TILE_TBL_BYTES + (size_t)t.flat_cap * 12 + ((size_t)t.flat_cap + 2) * 6 + ((size_t)t.flat_cap + 2);
        at = (at + 7) & ~(size_t)7;
        at += std::max<size_t>(((size_t)t.flat_cap + 2 * TILE_KPAD) * 4, (size_t)TILE_GT * 9);     // select keys, later the exact-scan buffers
        // the dense phase stages candidates (8 slots x 12 bytes per thread) in the region that starts at the flat list
        const size_t mini_from = ((size_t)TILE_TBL_BYTES + (size_t)t.flat_cap * 14 + 4 + 7) & ~(size_t)7;
        at = std::max(at, mini_from + (size_t)8 * 12 * TILE_THREADS);
        at = (at + 15) & ~(size_t)15;
        t.shared_at = (uint32_t)at;
        t.smem = at + sizeof(TileShared);
        return t;
    }
    return t;
}

// The single-pass sketch (tile_kernel.cuh). Returns NTL_OK with *done = false when the batch has to take the multi-pass
// path below (a tile ran out of exact-scan records).
static int sketch_device_tile(ntl_ctx* c, const TileShape& shape, const uint8_t* d_seq, const uint64_t* d_off, uint32_t nseq, uint64_t total_bases,
                              uint32_t k, uint32_t w, DeviceSketch& out, CallState* call_state, bool* done) {
    SketchWork& W = c->sw;
    *done = false;
    const uint32_t S = shape.S, NS = TILE_THREADS - 2 * shape.H;
    const uint32_t nstrips_max = (uint32_t)(total_bases / S + nseq + 1);
    const uint32_t ntiles_max = nstrips_max / NS + 1;
    uint32_t extras_cap = (uint32_t)std::min<uint64_t>(0xFFFF0000ull, total_bases / 64 + 65536);
    uint32_t out_cap = sketch_out_bound(total_bases, nseq, w, c->mx_density_factor);
    static int smem_set = 0;
    if (smem_set < (int)shape.smem) {
        NTL_CUDA(c, cudaFuncSetAttribute(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        smem_set = 227 * 1024;
    }
    int attempt = 0;
retry:
    TileParams P;
    P.k = k; P.w = w; P.S = S; P.cap = shape.cap; P.tau_hi = candidate_threshold(w, c->cand_c); P.nseq = nseq; P.H = shape.H;
    P.mult = second_hash_multiplier(k); P.out_cap = out_cap; P.extras_cap = extras_cap; P.ntiles_max = ntiles_max;
    P.flat_cap = shape.flat_cap; P.shared_at = shape.shared_at;
    NTL_CUDA(c, W.packed.ensure(total_bases / 2 + 512));
    NTL_CUDA(c, W.scnt.ensure(((size_t)nseq + 2) * 4));
    NTL_CUDA(c, W.strip_off.ensure(((size_t)nseq + 2) * 4));
    NTL_CUDA(c, W.strip_seq.ensure(((size_t)nstrips_max + 1) * 4));
    NTL_CUDA(c, W.extras.ensure((size_t)extras_cap * sizeof(Cand)));
    // staging: every tile writes its minimizers to its own segment of tcap entries (what random sequence needs x the density
    // head room of the output bound + 25 %); tile_state = counts [ntiles_max + 1] | bases [ntiles_max + 1]
    const uint32_t tcap = (uint32_t)((double)NS * S * c->mx_density_factor / ((double)w + 1.0) * 1.25) + 64;
    NTL_CUDA(c, W.tile_state.ensure(((size_t)ntiles_max + 2) * 8));
    NTL_CUDA(c, W.stage_hash.ensure((size_t)ntiles_max * tcap * 8 + 8));
    NTL_CUDA(c, W.stage_posf.ensure((size_t)ntiles_max * tcap * 4 + 4));
    P.tcap = tcap;
    uint32_t* tile_cnt = W.tile_state.as<uint32_t>();
    uint32_t* tile_base = tile_cnt + ntiles_max + 1;
    NTL_CUDA(c, W.status.ensure(sizeof(SketchStatus) + 64));
    NTL_CUDA(c, c->h_status.ensure(256));
    NTL_CUDA(c, out.hash.ensure((size_t)out_cap * 8 + 8));
    NTL_CUDA(c, out.posf.ensure((size_t)out_cap * 4 + 4));
    SketchStatus* st = W.status.as<SketchStatus>();
    uint32_t* nseq_dev = (uint32_t*)((char*)W.status.p + sizeof(SketchStatus));
    uint32_t* ticket = nseq_dev + 1;
    NTL_TRY(sketch_prepare(c, k));
    uint32_t* const d_packed = reinterpret_cast<uint32_t*>(W.packed.as<char>() + 64);
    {
        FillSegs fs{};
        fs.p[0] = st; fs.n[0] = sizeof(SketchStatus) + 64; fs.v[0] = 0;
        fs.p[1] = W.tile_state.p; fs.n[1] = ((size_t)ntiles_max + 2) * 8; fs.v[1] = 0;
        fs.p[2] = W.packed.p; fs.n[2] = 64; fs.v[2] = 0x44;
        fs.p[3] = W.packed.as<char>() + 64 + total_bases / 2; fs.n[3] = 192; fs.v[3] = 0x44;
        k_fill_segs<<<std::max<uint32_t>(1, std::min<uint32_t>(div_up((uint64_t)ntiles_max * 8, 256), 148)), 256, 0, c->stream>>>(fs);
        c->launches += 1;
    }
    tick(c, T_PACK);
    k_pack<<<div_up(div_up(total_bases, 16), 256 * PACK_CHUNKS), 256, 0, c->stream>>>(d_seq, total_bases, d_packed);
    k_strip_count<<<div_up(nseq, 256), 256, 0, c->stream>>>(d_off, nseq, k, w, S, W.scnt.as<uint32_t>(), nseq_dev);
    c->launches += 2;
    NTL_TRY(exclusive_scan_u32(c, W.scnt.as<uint32_t>(), W.strip_off.as<uint32_t>(), nseq_dev, nseq, W.blocksums));
    k_strip_seq<<<div_up(nstrips_max, 256), 256, 0, c->stream>>>(W.strip_off.as<uint32_t>(), nseq, W.strip_seq.as<uint32_t>());
    c->launches += 1;
    tock(c, T_PACK);

    tick(c, T_DENSE);
    {
        int per_sm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tile, TILE_THREADS, shape.smem);
        const uint32_t grid = std::max<uint32_t>(1, std::min<uint32_t>(ntiles_max, 148u * (uint32_t)std::max(per_sm, 1)));
        k_tile<<<grid, TILE_THREADS, shape.smem, c->stream>>>(d_packed, d_off, W.strip_off.as<uint32_t>(), W.strip_seq.as<uint32_t>(), P,
                                                              W.tbl.as<RollEntry>(), tile_cnt, ticket, W.extras.as<Cand>(),
                                                              W.stage_hash.as<uint64_t>(), W.stage_posf.as<uint32_t>(), out.mx_off.as<uint32_t>(), st);
    }
    tock(c, T_DENSE, total_bases);
    c->launches += 1; c->dense_launches += 1; c->dense_bases += total_bases;
    tick(c, T_EMIT);
    k_tile_scan<<<1, 1024, 0, c->stream>>>(W.strip_off.as<uint32_t>(), P, tile_cnt, tile_base, st);
    k_tile_gather<<<std::max<uint32_t>(1, std::min<uint32_t>(std::max<uint32_t>(ntiles_max, div_up((uint64_t)nseq + 1, 256)), 148 * 8)), 256, 0, c->stream>>>(
        W.strip_off.as<uint32_t>(), P, tile_cnt, tile_base, W.stage_hash.as<uint64_t>(), W.stage_posf.as<uint32_t>(), out.hash.as<uint64_t>(),
        out.posf.as<uint32_t>(), out.mx_off.as<uint32_t>(), st, call_state, call_state ? 1u : 0u);
    c->launches += 2;
    tock(c, T_EMIT);
    NTL_CUDA(c, cudaGetLastError());
    out.n_dev = &st->n_mx;
    if (call_state) { out.n_mx = out_cap; *done = true; return NTL_OK; }
    k_publish_sketch<<<1, 32, 0, c->stream>>>(st, W.strip_off.as<uint32_t>() + nseq, nullptr, c->h_status.as<uint32_t>());
    c->launches += 1;
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    {
        const SketchStatus hs = *c->h_status.as<SketchStatus>();
        if (hs.err) {
            if (hs.err & SKERR_GAPS) return NTL_OK;                        // *done stays false: multi-pass path
            if (++attempt > 6) { c->err = "sketch: device workspace exhausted"; return NTL_ERR_WORKSPACE; }
            if (hs.err & SKERR_EXTRAS) extras_cap = (uint32_t)std::min<uint64_t>(0xFFFF0000ull, std::max<uint64_t>((uint64_t)hs.extras_used + 1024, (uint64_t)extras_cap * 4));
            if (hs.err & SKERR_OUT) {                               // denser than the bound (low complexity): more room per tile and overall
                c->mx_density_factor *= 2.0;
                out_cap = (uint32_t)std::min<uint64_t>(total_bases, (uint64_t)out_cap * 2);
            }
            goto retry;
        }
        out.n_mx = hs.n_mx;
        note_mx_density(c, hs.n_mx, total_bases, w);
    }
    *done = true;
    return NTL_OK;
}

// The dense-mode sketch for small windows (small_kernel.cuh): pack, tile table, k_small / k_stream, scan of the counts, gather.
// Returns SMALL_SKIP when the batch should take the sparse path instead.
constexpr int SMALL_SKIP = 1;
constexpr double SMALL_STAGE_LIMIT = 48e9;
static int sketch_device_small(ntl_ctx* c, const uint8_t* d_seq, const uint64_t* d_off, uint32_t nseq, uint64_t total_bases,
                               uint32_t k, uint32_t w, DeviceSketch& out, CallState* call_state) {
    SketchWork& W = c->sw;
    // option small: 1 = automatic (measured, DESIGN.md 8: the tile kernel wins for w <= 6, where a third of all positions are
    // minimizers and the streaming kernel's per-position mark logic costs most), 2 = tile kernel, 3 = streaming kernel
    bool stream = c->small_mode == 3 || (c->small_mode == 1 && w > 6);
    // Staging is one segment per tile / strip, and every sequence adds a partial one: a batch of very many short sequences
    // (or a density head room that doubled after low-complexity input) would ask for more staging than the result is worth.
    // Beyond SMALL_STAGE_LIMIT the batch takes the other form if that fits (automatic mode), else the sparse path.
    auto stage_bytes = [&](bool st) {
        const double s_ = st ? (double)STREAM_S : (double)SMALL_S;
        const double n_ = (double)(total_bases / (uint64_t)s_ + nseq + 1);
        const double tc = std::min<double>(s_, s_ * c->mx_density_factor / ((double)w + 1.0) * 1.25 + (st ? 8.0 : 64.0));
        return n_ * tc * 12.0;
    };
    if (stage_bytes(stream) > SMALL_STAGE_LIMIT) {
        if (c->small_mode == 1 && stage_bytes(!stream) <= SMALL_STAGE_LIMIT) stream = !stream;
        else return SMALL_SKIP;
    }
    const uint32_t S = stream ? STREAM_S : SMALL_S;
    const uint32_t nstrips_max = (uint32_t)(total_bases / S + nseq + 1);
    static bool smem_set = false;
    if (!smem_set) {
        NTL_CUDA(c, cudaFuncSetAttribute(k_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM));
        NTL_CUDA(c, cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STREAM_SMEM));
        smem_set = true;
    }
    int attempt = 0;
retry:
    const uint32_t out_cap = sketch_out_bound(total_bases, nseq, w, c->mx_density_factor);
    SmallParams P;
    P.k = k; P.w = w; P.nseq = nseq; P.mult = second_hash_multiplier(k); P.out_cap = out_cap;
    // staging segment of a tile: what random sequence needs (2 / (w + 1) of the positions) x the density head room + 25 %
    P.tcap = (uint32_t)std::min<double>(S, (double)S * c->mx_density_factor / ((double)w + 1.0) * 1.25 + (stream ? 8.0 : 64.0));
    NTL_CUDA(c, W.packed.ensure(total_bases / 2 + 512));
    NTL_CUDA(c, W.scnt.ensure(((size_t)nseq + 2) * 4));
    NTL_CUDA(c, W.strip_off.ensure(((size_t)nseq + 2) * 4));
    NTL_CUDA(c, W.strip_seq.ensure(((size_t)nstrips_max + 1) * 4));
    NTL_CUDA(c, W.tile_state.ensure(((size_t)nstrips_max + 2) * 8));
    NTL_CUDA(c, W.stage_hash.ensure((size_t)nstrips_max * P.tcap * 8 + 8));
    NTL_CUDA(c, W.stage_posf.ensure((size_t)nstrips_max * P.tcap * 4 + 4));
    uint32_t* tile_cnt = W.tile_state.as<uint32_t>();
    uint32_t* tile_base = tile_cnt + nstrips_max + 1;
    NTL_CUDA(c, W.status.ensure(sizeof(SketchStatus) + 64));
    NTL_CUDA(c, c->h_status.ensure(256));
    NTL_CUDA(c, out.hash.ensure((size_t)out_cap * 8 + 8));
    NTL_CUDA(c, out.posf.ensure((size_t)out_cap * 4 + 4));
    SketchStatus* st = W.status.as<SketchStatus>();
    uint32_t* nseq_dev = (uint32_t*)((char*)W.status.p + sizeof(SketchStatus));
    NTL_TRY(sketch_prepare(c, k));
    uint32_t* const d_packed = reinterpret_cast<uint32_t*>(W.packed.as<char>() + 64);
    {
        FillSegs fs{};
        fs.p[0] = st; fs.n[0] = sizeof(SketchStatus) + 64; fs.v[0] = 0;
        fs.p[1] = W.packed.p; fs.n[1] = 64; fs.v[1] = 0x44;
        fs.p[2] = W.packed.as<char>() + 64 + total_bases / 2; fs.n[2] = 192; fs.v[2] = 0x44;
        k_fill_segs<<<1, 256, 0, c->stream>>>(fs);
        c->launches += 1;
    }
    tick(c, T_PACK);
    k_pack<<<div_up(div_up(total_bases, 16), 256 * PACK_CHUNKS), 256, 0, c->stream>>>(d_seq, total_bases, d_packed);
    k_strip_count<<<div_up(nseq, 256), 256, 0, c->stream>>>(d_off, nseq, k, w, S, W.scnt.as<uint32_t>(), nseq_dev);
    c->launches += 2;
    NTL_TRY(exclusive_scan_u32(c, W.scnt.as<uint32_t>(), W.strip_off.as<uint32_t>(), nseq_dev, nseq, W.blocksums));
    k_strip_seq<<<div_up(nstrips_max, 256), 256, 0, c->stream>>>(W.strip_off.as<uint32_t>(), nseq, W.strip_seq.as<uint32_t>());
    c->launches += 1;
    tock(c, T_PACK);
    tick(c, T_DENSE);
    if (stream) {
        static int per_sm = 0;
        if (!per_sm) {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_stream, STREAM_THREADS, STREAM_SMEM);
            per_sm = std::max(per_sm, 1);
        }
        const uint32_t grid = std::max<uint32_t>(1, std::min<uint32_t>(div_up(nstrips_max, STREAM_THREADS), 148u * (uint32_t)per_sm));
        k_stream<<<grid, STREAM_THREADS, STREAM_SMEM, c->stream>>>(d_packed, d_off, W.strip_off.as<uint32_t>(), W.strip_seq.as<uint32_t>(), P,
                                                                    W.tbl.as<RollEntry>(), tile_cnt, W.stage_hash.as<uint64_t>(),
                                                                    W.stage_posf.as<uint32_t>(), st);
    } else {
        static int per_sm = 0;
        if (!per_sm) {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_small, SMALL_THREADS * SMALL_GROUPS, SMALL_SMEM);
            per_sm = std::max(per_sm, 1);
        }
        const uint32_t grid = std::max<uint32_t>(1, std::min<uint32_t>(div_up(nstrips_max, SMALL_GROUPS), 148u * (uint32_t)per_sm));
        k_small<<<grid, SMALL_THREADS * SMALL_GROUPS, SMALL_SMEM, c->stream>>>(d_packed, d_off, W.strip_off.as<uint32_t>(), W.strip_seq.as<uint32_t>(), P,
                                                                 W.tbl.as<RollEntry>(), tile_cnt, nseq_dev + 1, W.stage_hash.as<uint64_t>(),
                                                                 W.stage_posf.as<uint32_t>(), st);
    }
    tock(c, T_DENSE, total_bases);
    c->launches += 1; c->dense_launches += 1; c->dense_bases += total_bases;
    tick(c, T_EMIT);
    NTL_TRY(exclusive_scan_u32(c, tile_cnt, tile_base, W.strip_off.as<uint32_t>() + nseq, nstrips_max, W.blocksums));
    if (stream)
        k_stream_gather<<<std::max<uint32_t>(1, std::min<uint32_t>(std::max<uint32_t>(div_up(nstrips_max, 8), div_up((uint64_t)nseq + 1, 256)), 148 * 8)), 256, 0, c->stream>>>(
            W.strip_off.as<uint32_t>(), P, tile_cnt, tile_base, W.stage_hash.as<uint64_t>(), W.stage_posf.as<uint32_t>(), out.hash.as<uint64_t>(),
            out.posf.as<uint32_t>(), out.mx_off.as<uint32_t>(), st, call_state, call_state ? 1u : 0u);
    else
        k_small_gather<<<std::max<uint32_t>(1, std::min<uint32_t>(std::max<uint32_t>(nstrips_max, div_up((uint64_t)nseq + 1, 256)), 148 * 8)), 256, 0, c->stream>>>(
            W.strip_off.as<uint32_t>(), P, tile_cnt, tile_base, W.stage_hash.as<uint64_t>(), W.stage_posf.as<uint32_t>(), out.hash.as<uint64_t>(),
            out.posf.as<uint32_t>(), out.mx_off.as<uint32_t>(), st, call_state, call_state ? 1u : 0u);
    c->launches += 1;
    tock(c, T_EMIT);
    NTL_CUDA(c, cudaGetLastError());
    out.n_dev = &st->n_mx;
    c->n_small_batches++;
    if (call_state) { out.n_mx = out_cap; return NTL_OK; }          // deferred: an overflow sets the call's error bit, the caller repeats synchronously
    k_publish_sketch<<<1, 32, 0, c->stream>>>(st, W.strip_off.as<uint32_t>() + nseq, nullptr, c->h_status.as<uint32_t>());
    c->launches += 1;
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    {
        const SketchStatus hs = *c->h_status.as<SketchStatus>();
        if (hs.err) {                                                // denser than the bound (low complexity): more room per tile and overall
            if (++attempt > 6) { c->err = "sketch: device workspace exhausted"; return NTL_ERR_WORKSPACE; }
            c->mx_density_factor *= 2.0;
            if (stage_bytes(stream) > SMALL_STAGE_LIMIT) return SMALL_SKIP;      // the sparse path sizes its buffers by what it finds
            goto retry;
        }
        out.n_mx = hs.n_mx;
        note_mx_density(c, hs.n_mx, total_bases, w);
    }
    return NTL_OK;
}

// Sketch `nseq` sequences that are already on the device (ASCII d_seq, offsets d_off). The result stays on the
// device in `out`. One host synchronisation (to learn the number of minimizers before the final compaction).
int sketch_device(ntl_ctx* c, const uint8_t* d_seq, const uint64_t* d_off, uint32_t nseq, uint64_t total_bases,
                  uint32_t k, uint32_t w, DeviceSketch& out, CallState* call_state) {
    if (k == 0 || w == 0 || k > 100000 || total_bases >= (1ull << 32)) { c->err = "sketch: bad k/w or batch too large"; return NTL_ERR_ARG; }
    SketchWork& W = c->sw;
    out.nseq = nseq; out.n_mx = 0; out.n_dev = nullptr;
    NTL_CUDA(c, out.mx_off.ensure(((size_t)nseq + 1) * 4));
    if (nseq == 0 || total_bases == 0) {
        NTL_CUDA(c, cudaMemsetAsync(out.mx_off.p, 0, ((size_t)nseq + 1) * 4, c->stream));
        return NTL_OK;
    }
    if (c->small_mode && w >= 2 && w <= SMALL_W_MAX && k <= 2048) {
        const int rc = sketch_device_small(c, d_seq, d_off, nseq, total_bases, k, w, out, call_state);
        if (rc != SMALL_SKIP) return rc;
    }
    if (c->tile_mode) {
        const TileShape shape = tile_shape(k, w, c->cand_c);
        if (shape.ok) {
            bool done = false;
            NTL_TRY(sketch_device_tile(c, shape, d_seq, d_off, nseq, total_bases, k, w, out, call_state, &done));
            c->n_tile_batches++;
            if (done) return NTL_OK;
            c->n_tile_fallbacks++;
            if (call_state) return NTL_OK;      // deferred: the error bit is set, the caller repeats the call synchronously
        }
    }
    // Strip length: 256 k-mer positions per thread keeps the k-1 lead-in at ~10 % and gives small batches enough threads;
    // with wide windows and batches of hundreds of Mbp 768 is ~9 % faster end to end (the select pass stages fewer context
    // strips per decided strip, the lead-in shrinks), measured on configs[2].
    const uint32_t S = c->strip_len ? c->strip_len : (w >= 200 && total_bases >= (256ull << 20) ? 768u : 256u);
    double mu = (double)S * c->cand_c / (double)w;
    if (mu > S) mu = S;
    uint32_t cap = (uint32_t)(mu + 6.0 * sqrt(mu) + 8.0);
    if (cap > S) cap = S;
    cap = (cap + 1) & ~1u;
    const uint32_t nstrips_max = (uint32_t)(total_bases / S + nseq + 1);

    int attempt = 0;
    uint32_t pool_cap = (uint32_t)std::max<uint64_t>(1 << 16, (uint64_t)nstrips_max * cap / 16);
    uint32_t gaps_cap = (uint32_t)std::max<uint64_t>(1 << 14, total_bases / (4ull * w) + nseq);
    uint32_t extras_cap = (uint32_t)std::max<uint64_t>(1 << 16, total_bases / (2ull * w) + nseq);
    uint32_t out_cap = 0;

retry:
    SkParams P;
    P.k = k; P.w = w; P.S = S; P.cap = cap; P.tau_hi = candidate_threshold(w, c->cand_c); P.nseq = nseq;
    P.mult = second_hash_multiplier(k);
    P.pool_base = (uint64_t)nstrips_max * cap;
    P.pool_cap = pool_cap; P.gaps_cap = gaps_cap; P.extras_cap = extras_cap; P.out_cap = 0;

    NTL_CUDA(c, W.packed.ensure(total_bases / 2 + 512));
    NTL_CUDA(c, W.scnt.ensure(((size_t)nseq + 2) * 4));
    NTL_CUDA(c, W.strip_off.ensure(((size_t)nseq + 2) * 4));
    NTL_CUDA(c, W.slots.ensure(((size_t)P.pool_base + pool_cap) * sizeof(Cand)));
    NTL_CUDA(c, W.sel.ensure((size_t)P.pool_base + pool_cap));
    NTL_CUDA(c, W.cnt.ensure(((size_t)nstrips_max + 1) * 4));
    NTL_CUDA(c, W.nv.ensure(((size_t)nstrips_max + 1) * 4));
    NTL_CUDA(c, W.vbase.ensure(((size_t)nstrips_max + 2) * 4));
    NTL_CUDA(c, W.ovf_off.ensure(((size_t)nstrips_max + 1) * 4));
    NTL_CUDA(c, W.selcnt.ensure(((size_t)nstrips_max + 1) * 4));
    NTL_CUDA(c, W.selmask.ensure(((size_t)nstrips_max + 1) * 8));
    NTL_CUDA(c, W.strip_seq.ensure(((size_t)nstrips_max + 1) * 4));
    NTL_CUDA(c, W.selbase.ensure(((size_t)nstrips_max + 2) * 4));
    NTL_CUDA(c, W.gap_head.ensure(((size_t)nstrips_max + 1) * 4 + 16));
    NTL_CUDA(c, W.gaps.ensure((size_t)gaps_cap * sizeof(GapRec)));
    NTL_CUDA(c, W.extras.ensure((size_t)extras_cap * sizeof(Cand)));
    NTL_CUDA(c, W.has_cand.ensure((size_t)nseq + 1));
    NTL_CUDA(c, W.status.ensure(sizeof(SketchStatus) + 64));
    NTL_CUDA(c, W.tbl.ensure(sizeof(RollEntry) * ROLL_TABLE_ENTRIES));
    NTL_CUDA(c, c->h_status.ensure(256));

    SketchStatus* st = W.status.as<SketchStatus>();
    uint32_t* nseq_dev = (uint32_t*)((char*)W.status.p + sizeof(SketchStatus));   // scan length for the strip table
    NTL_TRY(sketch_prepare(c, k));
    // layout: [64 B front pad | packed bases | >= 192 B tail pad]; d_packed points at the first real chunk
    uint32_t* const d_packed = reinterpret_cast<uint32_t*>(W.packed.as<char>() + 64);
    {
        FillSegs fs{};
        fs.p[0] = st; fs.n[0] = sizeof(SketchStatus) + 64; fs.v[0] = 0;
        fs.p[1] = W.has_cand.p; fs.n[1] = (size_t)nseq + 1; fs.v[1] = 0;
        fs.p[2] = W.packed.p; fs.n[2] = 64; fs.v[2] = 0x44;
        fs.p[3] = W.packed.as<char>() + 64 + total_bases / 2; fs.n[3] = 192; fs.v[3] = 0x44;
        fs.p[4] = W.gap_head.p; fs.n[4] = (((size_t)nstrips_max + 1) * 4 + 15) & ~(size_t)15; fs.v[4] = 0xFF;          // NONE32
        k_fill_segs<<<std::max<uint32_t>(1, std::min<uint32_t>(div_up((uint64_t)nstrips_max * 4, 4096), 296)), 256, 0, c->stream>>>(fs);
        c->launches += 1;
    }

    tick(c, T_PACK);
    k_pack<<<div_up(div_up(total_bases, 16), 256 * PACK_CHUNKS), 256, 0, c->stream>>>(d_seq, total_bases, d_packed);
    k_strip_count<<<div_up(nseq, 256), 256, 0, c->stream>>>(d_off, nseq, k, w, S, W.scnt.as<uint32_t>(), nseq_dev);
    c->launches += 2;
    NTL_TRY(exclusive_scan_u32(c, W.scnt.as<uint32_t>(), W.strip_off.as<uint32_t>(), nseq_dev, nseq, W.blocksums));
    k_strip_seq<<<div_up(nstrips_max, 256), 256, 0, c->stream>>>(W.strip_off.as<uint32_t>(), nseq, W.strip_seq.as<uint32_t>());
    c->launches += 1;
    tock(c, T_PACK);

    tick(c, T_DENSE);
    k_dense<<<div_up(nstrips_max, 128), 128, 0, c->stream>>>(d_packed, d_off, W.strip_off.as<uint32_t>(), W.strip_seq.as<uint32_t>(), P,
                                                            W.tbl.as<RollEntry>(), W.slots.as<Cand>(), W.cnt.as<uint32_t>(),
                                                            W.nv.as<uint32_t>(), W.has_cand.as<uint8_t>(), st);
    tock(c, T_DENSE, total_bases);
    c->launches += 1; c->dense_launches += 1; c->dense_bases += total_bases;

    tick(c, T_SELECT);
    k_overflow<<<div_up(nstrips_max, 128), 128, 0, c->stream>>>(d_packed, d_off, W.strip_off.as<uint32_t>(), W.strip_seq.as<uint32_t>(), P,
                                                               W.tbl.as<RollEntry>(), W.slots.as<Cand>(), W.cnt.as<uint32_t>(),
                                                               W.ovf_off.as<uint32_t>(), st);
    c->launches += 1;
    NTL_TRY(exclusive_scan_u32(c, W.nv.as<uint32_t>(), W.vbase.as<uint32_t>(), &st->nstrips, nstrips_max, W.blocksums));
    CandView V;
    V.cands = W.slots.as<Cand>(); V.cnt = W.cnt.as<uint32_t>(); V.ovf_off = W.ovf_off.as<uint32_t>();
    V.vbase = W.vbase.as<uint32_t>(); V.cap = cap; V.pool_base = P.pool_base;
    uint32_t nsb, nctx, scap;
    select_shape(mu, S, w, nsb, nctx, scap);
    k_select<<<div_up(nstrips_max, nsb), SEL_THREADS, (size_t)scap * sizeof(uint4), c->stream>>>(d_off, W.strip_off.as<uint32_t>(), W.strip_seq.as<uint32_t>(), P, V, W.sel.as<uint8_t>(),
                                                             W.selcnt.as<uint32_t>(), W.selmask.as<unsigned long long>(), W.gaps.as<GapRec>(),
                                                             W.gap_head.as<uint32_t>(), st, nsb, nctx, scap);
    k_seq_gaps<<<div_up(nseq, 128), 128, 0, c->stream>>>(d_off, W.strip_off.as<uint32_t>(), P, V, W.has_cand.as<uint8_t>(),
                                                        W.gaps.as<GapRec>(), W.gap_head.as<uint32_t>(), st);
    c->launches += 2;
    tock(c, T_SELECT);

    tick(c, T_GAP);
    k_gap<<<std::min<uint32_t>(div_up(gaps_cap, GAP_WARPS), 148 * 8), GAP_WARPS * 32, 0, c->stream>>>(
        d_packed, d_off, P, W.tbl.as<RollEntry>(), W.gaps.as<GapRec>(), W.extras.as<Cand>(), W.selcnt.as<uint32_t>(), st);
    c->launches += 1;
    tock(c, T_GAP);

    tick(c, T_EMIT);
    NTL_TRY(exclusive_scan_u32(c, W.selcnt.as<uint32_t>(), W.selbase.as<uint32_t>(), &st->nstrips, nstrips_max, W.blocksums));
    if (call_state) {
        // deferred pass: the output is sized by an upper bound (random sequence has 2/(w+1) minimizers per position;
        // 30 % head room, never more than one per position) and nobody waits for the counters
        out_cap = sketch_out_bound(total_bases, nseq, w, c->mx_density_factor);
        out.n_mx = out_cap;
        goto emit;
    }
    // total = selbase[nstrips]; fetch the counters to size the output
    // the counters reach the host through a tiny kernel that writes pinned (UVA-mapped) host memory: no copy engine
    // is involved, so this never queues behind a large host->device copy of the next batch
    k_publish_sketch<<<1, 32, 0, c->stream>>>(st, W.strip_off.as<uint32_t>() + nseq, W.selbase.as<uint32_t>(), c->h_status.as<uint32_t>());
    c->launches += 1;
    NTL_CUDA(c, cudaStreamSynchronize(c->stream));
    {
        SketchStatus hs = *c->h_status.as<SketchStatus>();
        const uint32_t nstrips = *(uint32_t*)(c->h_status.as<char>() + 64);
        if (hs.err) {
            if (++attempt > 6) { c->err = "sketch: device workspace exhausted"; return NTL_ERR_WORKSPACE; }
            if (hs.err & SKERR_POOL) pool_cap = (uint32_t)std::min<uint64_t>(0xFFFF0000ull, std::max<uint64_t>((uint64_t)hs.pool_used + 1024, (uint64_t)pool_cap * 4));
            if (hs.err & SKERR_GAPS) gaps_cap = (uint32_t)std::min<uint64_t>(0xFFFF0000ull, std::max<uint64_t>((uint64_t)hs.ngaps + 1024, (uint64_t)gaps_cap * 4));
            if (hs.err & SKERR_EXTRAS) extras_cap = (uint32_t)std::min<uint64_t>(0xFFFF0000ull, std::max<uint64_t>((uint64_t)hs.extras_used + 1024, (uint64_t)extras_cap * 4));
            goto retry;
        }
        (void)nstrips;
        const uint32_t total = *(uint32_t*)(c->h_status.as<char>() + 68);
        out_cap = total;
        out.n_mx = total;
        note_mx_density(c, total, total_bases, w);
    }
emit:
    NTL_CUDA(c, out.hash.ensure((size_t)out_cap * 8 + 8));
    NTL_CUDA(c, out.posf.ensure((size_t)out_cap * 4 + 4));
    P.out_cap = out_cap;
    out.n_dev = &st->n_mx;
    {
        int G = mu <= 24 ? 4 : mu <= 40 ? 8 : mu <= 96 ? 16 : 32;      // lanes per strip; configs[2] (mu = 21.5): G = 4 0.93 ms, 8 0.99, 16 1.22, 32 1.64
        if (const char* eg = getenv("NTL_EMIT_G")) G = atoi(eg);
#define NTL_EMIT(GG)                                                                                                              \
        k_emit<GG><<<div_up((uint64_t)nstrips_max * GG, EMIT_THREADS), EMIT_THREADS, 0, c->stream>>>(                               \
            P, V, W.sel.as<uint8_t>(), W.selmask.as<unsigned long long>(), W.selbase.as<uint32_t>(), W.gaps.as<GapRec>(),           \
            W.gap_head.as<uint32_t>(), W.extras.as<Cand>(), out.hash.as<uint64_t>(), out.posf.as<uint32_t>(), st)
        if (G == 4) NTL_EMIT(4); else if (G == 8) NTL_EMIT(8); else if (G == 16) NTL_EMIT(16); else NTL_EMIT(32);
#undef NTL_EMIT
    }
    k_seq_offsets<<<div_up((uint64_t)nseq + 1, 256), 256, 0, c->stream>>>(W.strip_off.as<uint32_t>(), W.selbase.as<uint32_t>(), nseq,
                                                                        out.mx_off.as<uint32_t>(), st, call_state);
    c->launches += 2;
    tock(c, T_EMIT);
    NTL_CUDA(c, cudaGetLastError());
    return NTL_OK;
}

}  // namespace ntl
