// select_kernel.cuh -- k_select: the exact minimizer decision for every candidate (included by sketch.cu).
//
// SIMT-friendly shape: a block stages the candidates of SEL_STRIPS consecutive strips plus SEL_CTX strips of
// context on each side in shared memory as two compact arrays (value, valid-k-mer index); then every thread takes
// one candidate and scans its neighbours left (first strictly smaller) and right (first smaller-or-equal) in shared
// memory with two short, simple loops. The rule applied is exactly select_candidate (sketch_logic.cuh); whatever
// the staged range cannot answer (context cut short by N runs, strips that overflowed their slots, more than
// SEL_CAP candidates) falls back to select_candidate on global memory, candidate by candidate.
#pragma once

namespace ntl {
namespace {

#ifndef SEL_MINB
#define SEL_MINB 6
#endif
constexpr int SEL_THREADS = 256;
constexpr int SEL_STRIPS = 32;     // strips decided per block (upper bound; the launch passes the actual number)
constexpr int SEL_CTX = 3;         // context strips staged on each side (upper bound, likewise)
constexpr int SEL_NL = SEL_STRIPS + 2 * SEL_CTX;
constexpr int SEL_CAP = 2560;      // most candidates a block can stage (40 KB); the launch sizes the buffer (scap)

__global__ void __launch_bounds__(SEL_THREADS, SEL_MINB) k_select(const uint64_t* __restrict__ seq_off,
                                                         const uint32_t* __restrict__ strip_off,
                                                         const uint32_t* __restrict__ strip_seq, SkParams P, CandView V,
                                                         uint8_t* __restrict__ sel, uint32_t* __restrict__ selcnt,
                                                         unsigned long long* __restrict__ selmask,
                                                         GapRec* __restrict__ gaps, uint32_t* __restrict__ gap_head,
                                                         SketchStatus* __restrict__ st, uint32_t nsb, uint32_t nctx, uint32_t scap) {
    // nsb strips are decided per block with nctx strips of context on each side: chosen by the host so that the
    // expected number of staged candidates fits SEL_CAP and the context covers w - 1 positions (select_shape)
    // staged candidates {h0.lo, h0.hi, valid-k-mer index, staged strip}: scap entries of dynamic shared memory, sized by
    // the host from the expected candidate density so that as many blocks as possible are resident (the kernel waits on
    // global loads while staging; more resident blocks hide that)
    extern __shared__ uint4 sh_c[];               // {h0.lo, h0.hi, valid-k-mer index, staged strip}
    __shared__ uint32_t sh_off[SEL_NL + 1];       // compact offset of every staged strip
    __shared__ uint32_t sh_cnt[SEL_NL];
    __shared__ uint32_t sh_q[SEL_NL], sh_fs[SEL_NL], sh_es[SEL_NL], sh_idx0[SEL_NL], sh_n[SEL_NL], sh_np[SEL_NL];
    __shared__ uint32_t sh_sel[SEL_STRIPS];
    __shared__ unsigned long long sh_mask[SEL_STRIPS];   // bit j = candidate j of the strip is a minimizer (j < 64)
    __shared__ uint32_t sh_bad;                   // the staged range cannot be used -> whole block falls back

    const uint32_t nstrips = st->nstrips;
    const uint32_t b0 = blockIdx.x * nsb;
    if (b0 >= nstrips) return;
    const uint32_t b1 = min(b0 + nsb, nstrips);
    const uint32_t l0 = b0 > nctx ? b0 - nctx : 0;
    const uint32_t l1 = min(b1 + nctx, nstrips);
    const uint32_t nl = l1 - l0;
    const uint32_t tid = threadIdx.x;

    if (tid == 0) sh_bad = 0;
    if (tid < SEL_STRIPS) { sh_sel[tid] = 0; sh_mask[tid] = 0ull; }
    __syncthreads();
    // per staged strip: candidate count and sequence bounds; offsets by a tiny serial scan (nl <= 38)
    if (tid < nl) {
        const uint32_t s = l0 + tid;
        uint32_t c = V.cnt[s];
        if (c > V.cap) { sh_bad = 1; c = 0; }
        sh_cnt[tid] = c;
        const uint32_t q = strip_seq[s];
        const uint32_t fs = strip_off[q], es = strip_off[q + 1];
        sh_q[tid] = q; sh_fs[tid] = fs; sh_es[tid] = es;
        sh_idx0[tid] = V.vbase[fs]; sh_n[tid] = V.vbase[es] - V.vbase[fs];
        sh_np[tid] = seq_npos(seq_off[q + 1] - seq_off[q], P.k, P.w);
    }
    __syncthreads();
    if (tid < 32) {                               // exclusive scan of <= 64 counts by one warp (two per lane)
        const uint32_t a = tid < nl ? sh_cnt[tid] : 0, b = tid + 32 < nl ? sh_cnt[tid + 32] : 0;
        uint32_t x = a, y = b;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t x2 = __shfl_up_sync(0xffffffffu, x, d), y2 = __shfl_up_sync(0xffffffffu, y, d);
            if (tid >= (uint32_t)d) { x += x2; y += y2; }
        }
        const uint32_t tot_a = __shfl_sync(0xffffffffu, x, 31);
        if (tid < nl) sh_off[tid] = x - a;
        if (tid + 32 < nl) sh_off[tid + 32] = tot_a + y - b;
        if (tid == 31) sh_off[nl] = tot_a + y;
    }
    __syncthreads();
    const uint32_t total = sh_off[nl];
    const bool staged = (sh_bad == 0) && (total <= scap);
    __syncthreads();

    if (staged) {
        // one warp per staged strip, lanes copy the strip's candidates (coalesced 16-byte loads, no searching). A warp
        // owns up to SEL_ROUNDS strips; the loads of all of them are issued before the first one is consumed, so the
        // block waits for ONE global-memory latency here, not for one per strip.
        constexpr int SEL_ROUNDS = (SEL_NL + SEL_THREADS / 32 - 1) / (SEL_THREADS / 32);
        const uint32_t lane = tid & 31, wrp = tid >> 5;
        Cand cd[SEL_ROUNDS];
        uint32_t vb[SEL_ROUNDS];
#pragma unroll
        for (int u = 0; u < SEL_ROUNDS; u++) {
            const uint32_t t = wrp + u * (SEL_THREADS / 32);
            cd[u].h0 = 0; cd[u].posf = 0; cd[u].lord = 0; vb[u] = 0;
            if (t < nl) {
                vb[u] = V.vbase[l0 + t];
                if (lane < sh_cnt[t]) cd[u] = V.cands[(uint64_t)(l0 + t) * V.cap + lane];
            }
        }
#pragma unroll
        for (int u = 0; u < SEL_ROUNDS; u++) {
            const uint32_t t = wrp + u * (SEL_THREADS / 32);
            if (t < nl && lane < sh_cnt[t])
                sh_c[sh_off[t] + lane] = make_uint4((uint32_t)cd[u].h0, (uint32_t)(cd[u].h0 >> 32), vb[u] + cd[u].lord, t);
        }
        // strips with more than 32 candidates (small w): the rest, 32 at a time
        for (uint32_t t = wrp; t < nl; t += SEL_THREADS / 32) {
            const uint32_t c = sh_cnt[t];
            if (c <= 32) continue;
            const uint32_t s = l0 + t, o = sh_off[t], vbs = V.vbase[s];
            for (uint32_t j = 32 + lane; j < c; j += 32) {
                const Cand x = V.cands[(uint64_t)s * V.cap + j];
                sh_c[o + j] = make_uint4((uint32_t)x.h0, (uint32_t)(x.h0 >> 32), vbs + x.lord, t);
            }
        }
    }
    __syncthreads();

    const uint32_t w = P.w;
    if (staged) {
        const uint32_t e_begin = sh_off[b0 - l0], e_end = sh_off[b1 - l0];
        const uint32_t first_idx = V.vbase[l0], end_idx = V.vbase[l1];   // valid k-mers covered by the staged strips
        const int32_t w32 = (int32_t)w, W1s = w32 - 1;
        // Uniform trip count for the whole block and explicit reconvergence (__syncwarp) after each scan: the scans
        // have data-dependent lengths, and without the barriers the lanes of a warp run the long tail below one
        // small group at a time. Positions inside a sequence are below 2^31 (POS_MASK), so the window arithmetic is 32-bit.
        for (uint32_t eb = e_begin; eb < e_end; eb += SEL_THREADS) {
            const uint32_t e = eb + tid;
            const bool act = e < e_end;
            uint32_t idxu = 0, t = 0, e_lo = 0, e_hi = 0;
            uint64_t val = 0;
            if (act) {
                const uint4 me = sh_c[e];
                val = ((uint64_t)me.y << 32) | me.x; idxu = me.z; t = me.w;
                const uint32_t fs0 = sh_fs[t], es0 = sh_es[t];
                e_lo = sh_off[(fs0 > l0 ? fs0 : l0) - l0];               // staged candidates of the same sequence
                e_hi = sh_off[(es0 < l1 ? es0 : l1) - l0];
            }
            // left: first strictly smaller value (single-exit loop)
            int32_t A32 = -1;
            {
                uint32_t p = e;
                bool go = act && p > e_lo;
                while (go) {
                    p--;
                    const uint4 cp = sh_c[p];
                    const int32_t d = (int32_t)(idxu - cp.z);
                    const bool smaller = (((uint64_t)cp.y << 32) | cp.x) < val;
                    const bool far = d >= w32;
                    if (far | smaller) { A32 = far ? W1s : d - 1; go = false; }
                    else go = p > e_lo;
                }
            }
            __syncwarp();
            // right: first smaller-or-equal value
            int32_t B32 = -1;
            {
                uint32_t p = e + 1;
                bool go = act && p < e_hi;
                while (go) {
                    const uint4 cp = sh_c[p];
                    const int32_t d = (int32_t)(cp.z - idxu);
                    const bool smaller = (((uint64_t)cp.y << 32) | cp.x) <= val;
                    const bool far = d >= w32;
                    if (far | smaller) { B32 = far ? W1s : d - 1; go = false; }
                    else { p++; go = p < e_hi; }
                }
            }
            __syncwarp();
            if (!act) continue;
            const uint32_t s = l0 + t, j = e - sh_off[t];
            const uint32_t fs = sh_fs[t], es = sh_es[t];
            const uint32_t idx0 = sh_idx0[t];
            const int32_t n = (int32_t)sh_n[t], rel = (int32_t)(idxu - idx0);
            bool need_fallback = false;
            int32_t A = A32, B = B32;
            if (A32 < 0) {
                if (fs >= l0) A = rel;                                   // the sequence starts inside the staged range
                else if (idxu - first_idx >= (uint32_t)W1s) A = W1s;     // everything earlier is out of reach
                else need_fallback = true;
            }
            // the immediate right neighbour bounds the candidate-free stretch
            uint32_t gap_len = 0, gap_end = 0;
            bool gap_known = true;
            if (e + 1 < e_hi) gap_len = sh_c[e + 1].z - idxu - 1;
            else if (es <= l1) { gap_len = (uint32_t)(n - 1 - rel); gap_end = sh_np[t]; }
            else gap_known = false;
            if (B32 < 0) {
                if (es <= l1) B = n - 1 - rel;                           // the sequence ends inside the staged range
                else if (end_idx - idxu > (uint32_t)W1s) B = W1s;
                else need_fallback = true;
            }
            const uint64_t gid = (uint64_t)s * V.cap + j;
            bool selected;
            if (need_fallback || !gap_known) {
                const SelectResult r = select_candidate(V, s, j, fs, es, w, sh_np[t]);
                selected = r.selected; gap_len = r.gap_len; gap_end = r.gap_end;
            } else {
                const int32_t lo_w = max(0, max(rel - W1s, rel - A));
                const int32_t hi_w = min(rel, min(n - w32, rel + B - W1s));
                selected = lo_w <= hi_w;
                if (gap_len >= w && e + 1 < e_hi) {                      // position of the neighbour that ends the stretch
                    const uint32_t t2 = sh_c[e + 1].w;
                    gap_end = V.cands[(uint64_t)(l0 + t2) * V.cap + (e + 1 - sh_off[t2])].posf & POS_MASK;
                }
            }
            sel[gid] = selected ? 1 : 0;
            if (selected) {
                atomicAdd(&sh_sel[t - (b0 - l0)], 1u);
                if (j < 64) atomicOr(&sh_mask[t - (b0 - l0)], 1ull << j);
            }
            if (gap_len >= w) queue_gap(gaps, gap_head, st, P, sh_q[t], (V.cands[gid].posf & POS_MASK) + 1, gap_end, s, j, gap_len);
        }
    } else {
        // fallback for the whole block: per-candidate neighbour scan on global memory
        for (uint32_t t = b0 - l0; t < b1 - l0; t++) {
            const uint32_t s = l0 + t;
            const uint32_t c = V.cnt[s];
            for (uint32_t j = tid; j < c; j += SEL_THREADS) {
                const SelectResult r = select_candidate(V, s, j, sh_fs[t], sh_es[t], w, sh_np[t]);
                const uint64_t gid = cand_gid(V, s, j);
                sel[gid] = r.selected ? 1 : 0;
                if (r.selected) {
                    atomicAdd(&sh_sel[t - (b0 - l0)], 1u);
                    if (j < 64) atomicOr(&sh_mask[t - (b0 - l0)], 1ull << j);
                }
                if (r.gap_len >= w)
                    queue_gap(gaps, gap_head, st, P, sh_q[t], (V.cands[gid].posf & POS_MASK) + 1, r.gap_end, s, j, r.gap_len);
            }
        }
    }
    __syncthreads();
    if (tid < b1 - b0) { selcnt[b0 + tid] = sh_sel[tid]; selmask[b0 + tid] = sh_mask[tid]; }
}

// strips per block / context strips for a given expected number of candidates per strip
inline void select_shape(double mu, uint32_t S, uint32_t w, uint32_t& nsb, uint32_t& nctx, uint32_t& scap) {
    nctx = std::min<uint32_t>(SEL_CTX, std::max<uint32_t>(1, (w - 1 + S - 1) / S));
    const double fit = 0.85 * SEL_CAP / std::max(1.0, mu) - 2.0 * nctx;
    nsb = (uint32_t)std::min<double>(SEL_STRIPS, std::max(1.0, fit));
    // room for the expected number of staged candidates + 6 sigma (Poisson) + slack; blocks that exceed it fall back
    const double staged = (nsb + 2.0 * nctx) * mu;
    scap = (uint32_t)std::min<double>(SEL_CAP, staged + 6.0 * std::sqrt(staged) + 64.0);
    scap = (scap + 63u) & ~63u;
}

}  // namespace
}  // namespace ntl
