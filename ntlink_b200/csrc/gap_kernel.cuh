// gap_kernel.cuh -- k_gap: exact re-scan of candidate-free stretches (included by sketch.cu).
//
// A stretch of >= w valid k-mers without a candidate (probability ~e^-c per window) needs the textbook windowed
// minimum. One WARP per stretch: the lanes hash disjoint chunks of the stretch in parallel (process_strip with
// ALL = true) into a shared-memory array of canonical hashes, then every lane takes windows j = lane, lane+32, ...
// and finds the rightmost argmin of its window; consecutive windows with the same argmin are dropped (same rule
// as btllib: emit when the minimizer position advances) and the survivors are written in window order with a
// ballot/popc compaction. Stretches that contain an invalid base (windows then span N gaps) take the serial
// walker gap_scan on lane 0 -- identical results, only slower.
#pragma once

namespace ntl {
namespace {

constexpr int GAP_WARPS = 4;
constexpr int GAP_TILE = 768;        // k-mer positions hashed per tile (6 KB of hashes per warp)

struct TileEmit {
    uint64_t* h; uint8_t* f; uint32_t base;
    __device__ __forceinline__ void operator()(uint64_t h0, uint32_t pos, bool fwd, uint32_t) {
        h[pos - base] = h0; f[pos - base] = fwd ? 1 : 0;
    }
};
struct ExtraEmit {
    Cand* dst;
    uint32_t count, cap;
    __device__ __forceinline__ void operator()(uint64_t h0, uint32_t pos, bool fwd) {
        if (count < cap) { Cand c; c.h0 = h0; c.posf = pos | (fwd ? FWD_BIT : 0u); c.lord = 0; dst[count] = c; }
        count++;
    }
};

__global__ void __launch_bounds__(GAP_WARPS * 32) k_gap(const uint32_t* __restrict__ packed, const uint64_t* __restrict__ seq_off,
                                                        SkParams P, const RollEntry* __restrict__ tbl_g, GapRec* __restrict__ gaps,
                                                        Cand* __restrict__ extras, uint32_t* __restrict__ selcnt,
                                                        SketchStatus* __restrict__ st) {
    __shared__ RollEntry tbl_s[ROLL_TABLE_ENTRIES];
    __shared__ uint64_t sm_h[GAP_WARPS][GAP_TILE];
    __shared__ uint8_t sm_f[GAP_WARPS][GAP_TILE];
    __shared__ uint64_t sm_gm[GAP_WARPS][GAP_TILE / 8];     // minimum of every group of 8 hashed positions ...
    __shared__ uint16_t sm_ga[GAP_WARPS][GAP_TILE / 8];     // ... and its rightmost position
    for (uint32_t i = threadIdx.x; i < ROLL_TABLE_ENTRIES; i += blockDim.x) tbl_s[i] = tbl_g[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
    const uint32_t ng = min(st->ngaps, P.gaps_cap);
    const uint32_t w = P.w, k = P.k;
    uint64_t* H = sm_h[wv];
    uint8_t* F = sm_f[wv];
    uint64_t* GM = sm_gm[wv];
    uint16_t* GA = sm_ga[wv];
    for (uint32_t id = blockIdx.x * GAP_WARPS + wv; id < ng; id += gridDim.x * GAP_WARPS) {
        const GapRec g = gaps[id];
        if (g.max_out == 0) continue;
        const uint64_t gseq = seq_off[g.seq];
        const uint32_t L = (uint32_t)(seq_off[g.seq + 1] - gseq);
        Cand* out = extras + g.out_off;
        uint32_t n_out = 0;
        // any invalid base among the bases of the stretch?
        const uint32_t nb = g.end_pos - g.start_pos + k - 1;
        bool dirty = false;
        for (uint32_t b = lane * 8; b < nb; b += 256) {
            uint32_t wd = fetch8(packed, gseq + g.start_pos + b);
            const uint32_t left = nb - b;
            if (left < 8) wd &= (1u << (4 * left)) - 1u;
            dirty |= (wd & 0x44444444u) != 0u;
        }
        dirty = __any_sync(0xffffffffu, dirty) || (w > GAP_TILE / 2);
        if (dirty) {
            if (lane == 0) {
                ExtraEmit em{out, 0, g.max_out};
                gap_scan(packed, tbl_s, gseq, L, k, w, g.start_pos, g.end_pos, em);
                n_out = min(em.count, g.max_out);
            }
            n_out = __shfl_sync(0xffffffffu, n_out, 0);
        } else {
            const uint32_t G = g.end_pos - g.start_pos;            // all valid
            const uint32_t nwin = G >= w ? G - w + 1 : 0;
            uint32_t prev_amin = NONE32;                           // argmin (sequence position) of the previous window
            for (uint32_t wb = 0; wb < nwin; wb += GAP_TILE - w + 1) {
                const uint32_t we = min(nwin, wb + (GAP_TILE - w + 1));   // windows [wb, we) of this tile
                const uint32_t npt = (we - wb) + w - 1;                 // positions hashed: start_pos + wb .. + npt
                const uint32_t base = g.start_pos + wb;
                const uint32_t chunk = (npt + 31) / 32;
                const uint32_t pa = min(npt, lane * chunk), pb = min(npt, pa + chunk);
                __syncwarp();
                if (pb > pa) {
                    TileEmit te{H, F, base};
                    process_strip<true>(packed, gseq, base + pa, pb - pa, k, tbl_s, 1, 0u, te);
                }
                __syncwarp();
                // minima of aligned groups of 8 positions: a window then costs ~w/8 + 14 probes instead of w
                for (uint32_t gi = lane; gi * 8 < npt; gi += 32) {
                    const uint32_t q0 = gi * 8, q1 = min(npt, q0 + 8);
                    uint64_t m = H[q0];
                    uint32_t a = q0;
                    for (uint32_t q = q0 + 1; q < q1; q++) {
                        const uint64_t hv = H[q];
                        if (hv <= m) { m = hv; a = q; }
                    }
                    GM[gi] = m; GA[gi] = (uint16_t)a;
                }
                __syncwarp();
                for (uint32_t j0 = wb; j0 < we; j0 += 32) {
                    const uint32_t j = j0 + lane;
                    uint32_t amin = NONE32;
                    uint64_t hmin = 0;
                    if (j < we) {
                        const uint32_t o = j - wb, end = o + w;             // window = hashed positions [o, end)
                        hmin = H[o]; amin = o;
                        uint32_t q = o + 1;
                        for (; q < end && (q & 7u); q++) {                  // up to the next group boundary
                            const uint64_t hv = H[q];
                            if (hv <= hmin) { hmin = hv; amin = q; }
                        }
                        for (; q + 8 <= end; q += 8) {                      // whole groups
                            const uint64_t hv = GM[q >> 3];
                            if (hv <= hmin) { hmin = hv; amin = GA[q >> 3]; }
                        }
                        for (; q < end; q++) {                              // the rest
                            const uint64_t hv = H[q];
                            if (hv <= hmin) { hmin = hv; amin = q; }
                        }
                        amin += base;                                   // sequence position
                    }
                    uint32_t left = __shfl_up_sync(0xffffffffu, amin, 1);
                    if (lane == 0) left = prev_amin;
                    const bool emit = (j < we) && (amin != left) && (hmin != 0xFFFFFFFFFFFFFFFFULL);
                    const uint32_t ballot = __ballot_sync(0xffffffffu, emit);
                    if (emit) {
                        const uint32_t r = n_out + __popc(ballot & ((1u << lane) - 1u));
                        if (r < g.max_out) {
                            Cand c; c.h0 = hmin; c.posf = amin | (F[amin - base] ? FWD_BIT : 0u); c.lord = 0;
                            out[r] = c;
                        }
                    }
                    n_out += __popc(ballot);
                    // argmin of the last window of this round (lanes beyond `we` carry NONE32: take the last valid lane)
                    const uint32_t nvalid = min(32u, we - j0);
                    prev_amin = __shfl_sync(0xffffffffu, amin, nvalid - 1);
                }
            }
            n_out = min(n_out, g.max_out);
        }
        if (lane == 0) {
            gaps[id].out_cnt = n_out;
            if (n_out) atomicAdd(&selcnt[g.strip], n_out);
        }
    }
}

}  // namespace
}  // namespace ntl
