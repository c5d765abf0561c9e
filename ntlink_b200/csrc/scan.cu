// scan.cu -- exclusive prefix sum of uint32 arrays whose length lives on the device (no host round trip): one kernel per
// scan (a one-block kernel for small arrays, a chained scan with decoupled look-back for large ones).
#include "common.cuh"

namespace ntl {

namespace {
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) >= d) v += t;
    }
    return v;
}
// exclusive scan of one value per thread across the block; returns the exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* smem /* [33] */, uint32_t* total) {
    const uint32_t incl = warp_incl_scan(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 31) smem[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < (int)(blockDim.x >> 5) ? smem[lane] : 0;
        uint32_t si = warp_incl_scan(s);
        smem[lane] = si - s;
        if (lane == 31) smem[32] = si;
    }
    __syncthreads();
    const uint32_t r = smem[wid] + incl - v;
    *total = smem[32];
    __syncthreads();
    return r;
}

// small arrays (the strip table, per-read counters, the pair table): one block does the whole scan in one launch
constexpr uint32_t SCAN_SMALL_MAX = 1u << 13;
__global__ void __launch_bounds__(1024) k_scan_small(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                     const uint32_t* __restrict__ n_dev) {
    __shared__ uint32_t sm[33];
    const uint32_t n = *n_dev;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n; base += 1024 * 4) {
        const uint32_t i0 = base + threadIdx.x * 4;
        uint32_t v[4], s = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) { v[i] = (i0 + i < n) ? in[i0 + i] : 0; s += v[i]; }
        uint32_t total;
        uint32_t ex = block_excl_scan(s, sm, &total) + carry;
#pragma unroll
        for (int i = 0; i < 4; i++) { if (i0 + i < n) out[i0 + i] = ex; ex += v[i]; }
        carry += total;
    }
    if (threadIdx.x == 0) out[n] = carry;
}
// Large arrays: ONE kernel (chained scan with decoupled look-back) instead of reduce / scan-of-sums / apply. A block takes
// its tile number from a ticket counter (blocks with a lower number are then guaranteed to be running or done), publishes
// its tile total, sums the totals / prefixes of the tiles before it, publishes its own prefix and writes its part of the
// output. The per-tile state words carry the epoch of the call, so nothing has to be cleared between calls:
//   bits 63..36 epoch | bits 35..34 flag (1 tile total, 2 inclusive prefix) | bits 33..0 value
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_chained(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                const uint32_t* __restrict__ n_dev, unsigned long long* __restrict__ state,
                                                                uint32_t* __restrict__ ticket, unsigned long long epoch) {
    __shared__ uint32_t sm[33];
    __shared__ uint32_t s_bid;
    __shared__ unsigned long long s_excl;
    if (threadIdx.x == 0) {
        const uint32_t t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) *ticket = 0;             // every block of this launch has its number: ready for the next launch
        s_bid = t;
    }
    __syncthreads();
    const uint32_t bid = s_bid;
    const uint32_t n = *n_dev;
    const uint32_t nblocks = n ? (uint32_t)(((uint64_t)n + SCAN_TILE - 1) / SCAN_TILE) : 1u;
    if (bid >= nblocks) return;                          // the grid is sized by the host-side bound
    const uint64_t base = (uint64_t)bid * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        const uint64_t idx = base + i;
        v[i] = idx < n ? in[idx] : 0;
        s += v[i];
    }
    uint32_t total;
    uint32_t ex = block_excl_scan(s, sm, &total);
    const unsigned long long VAL = (1ull << 34) - 1;
    if (threadIdx.x < 32) {
        const uint32_t lane = threadIdx.x;
        unsigned long long excl = 0;
        if (bid > 0) {
            if (lane == 0) { __threadfence(); atomicExch(&state[bid], (epoch << 36) | (1ull << 34) | total); }
            int64_t look = (int64_t)bid - 1;
            for (;;) {
                const int64_t idx = look - lane;
                unsigned long long w = (epoch << 36) | (2ull << 34);                    // before tile 0: prefix 0
                if (idx >= 0) {
                    const volatile unsigned long long* p = state + idx;
                    while ((((w = *p) >> 36) != epoch) || ((w >> 34) & 3ull) == 0) __nanosleep(32);
                }
                const uint32_t pref = __ballot_sync(0xffffffffu, ((w >> 34) & 3ull) == 2);
                const uint32_t upto = pref ? (uint32_t)__ffs(pref) - 1 : 31u;
                unsigned long long t = lane <= upto ? (w & VAL) : 0ull;
#pragma unroll
                for (int d = 16; d; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
                excl += t;
                if (pref) break;
                look -= 32;
            }
        }
        if (lane == 0) {
            __threadfence();
            atomicExch(&state[bid], (epoch << 36) | (2ull << 34) | ((excl + total) & VAL));
            s_excl = excl;
        }
    }
    __syncthreads();
    ex += (uint32_t)s_excl;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        const uint64_t idx = base + i;
        if (idx < n) out[idx] = ex;
        ex += v[i];
    }
    if (bid == nblocks - 1 && threadIdx.x == 0) out[n] = (uint32_t)s_excl + total;
}
}  // namespace

int exclusive_scan_u32(ntl_ctx* c, const uint32_t* in, uint32_t* out, const uint32_t* n_dev, uint32_t n_max,
                       DevBuf& blocksums) {
    if (n_max <= SCAN_SMALL_MAX) {
        k_scan_small<<<1, 1024, 0, c->stream>>>(in, out, n_dev);
        c->launches += 1;
        NTL_CUDA(c, cudaGetLastError());
        return NTL_OK;
    }
    const uint32_t nblocks = (n_max + SCAN_TILE - 1) / SCAN_TILE;
    // scratch: [ticket, padding to 16 bytes | one state word per tile]; zeroed once per allocation
    NTL_CUDA(c, blocksums.ensure(16 + ((size_t)nblocks + 1) * 8));
    if (blocksums.scan_ready_for != blocksums.p) {
        NTL_CUDA(c, cudaMemsetAsync(blocksums.p, 0, blocksums.cap, c->stream));
        blocksums.scan_ready_for = blocksums.p;
    }
    blocksums.scan_epoch = (blocksums.scan_epoch + 1) & ((1ull << 28) - 1);
    if (blocksums.scan_epoch == 0) blocksums.scan_epoch = 1;
    k_scan_chained<<<nblocks, SCAN_THREADS, 0, c->stream>>>(in, out, n_dev, reinterpret_cast<unsigned long long*>(blocksums.as<char>() + 16),
                                                            blocksums.as<uint32_t>(), blocksums.scan_epoch);
    c->launches += 1;
    NTL_CUDA(c, cudaGetLastError());
    return NTL_OK;
}

}  // namespace ntl
