// scan.cu -- exclusive prefix sum of uint32 arrays whose length lives on the device (no host round trip).
// Three small kernels: per-block reduce, one-block scan of the block sums, per-block scan + offset.
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

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const uint32_t* __restrict__ in,
                                                               const uint32_t* __restrict__ n_dev,
                                                               uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t sm[33];
    const uint32_t n = *n_dev;
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        const uint64_t idx = base + (uint64_t)i * SCAN_THREADS + threadIdx.x;
        if (idx < n) s += in[idx];
    }
    uint32_t total;
    block_excl_scan(s, sm, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(uint32_t* __restrict__ block_sums, uint32_t nblocks) {
    __shared__ uint32_t sm[33];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nblocks; base += SCAN_THREADS) {
        const uint32_t idx = base + threadIdx.x;
        const uint32_t v = idx < nblocks ? block_sums[idx] : 0;
        uint32_t total;
        const uint32_t ex = block_excl_scan(v, sm, &total);
        if (idx < nblocks) block_sums[idx] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) block_sums[nblocks] = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t* __restrict__ in,
                                                              uint32_t* __restrict__ out,
                                                              const uint32_t* __restrict__ n_dev,
                                                              const uint32_t* __restrict__ block_sums,
                                                              uint32_t nblocks) {
    __shared__ uint32_t sm[33];
    const uint32_t n = *n_dev;
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        const uint64_t idx = base + i;
        v[i] = idx < n ? in[idx] : 0;
        s += v[i];
    }
    uint32_t total;
    uint32_t ex = block_excl_scan(s, sm, &total) + block_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        const uint64_t idx = base + i;
        if (idx < n) out[idx] = ex;
        ex += v[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = block_sums[nblocks];
}
// small arrays (the strip table, per-read counters, the pair table): one block does the whole scan in one launch
constexpr uint32_t SCAN_SMALL_MAX = 1u << 16;
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
}  // namespace

int exclusive_scan_u32(ntl_ctx* c, const uint32_t* in, uint32_t* out, const uint32_t* n_dev, uint32_t n_max,
                       DevBuf& blocksums) {
    if (n_max <= SCAN_SMALL_MAX) {
        k_scan_small<<<1, 1024, 0, c->stream>>>(in, out, n_dev);
        c->launches += 1;
        NTL_CUDA(c, cudaGetLastError());
        return NTL_OK;
    }
    const uint32_t nblocks = n_max ? (n_max + SCAN_TILE - 1) / SCAN_TILE : 1;
    NTL_CUDA(c, blocksums.ensure(((size_t)nblocks + 1) * sizeof(uint32_t)));
    k_scan_reduce<<<nblocks, SCAN_THREADS, 0, c->stream>>>(in, n_dev, blocksums.as<uint32_t>());
    k_scan_sums<<<1, SCAN_THREADS, 0, c->stream>>>(blocksums.as<uint32_t>(), nblocks);
    k_scan_apply<<<nblocks, SCAN_THREADS, 0, c->stream>>>(in, out, n_dev, blocksums.as<uint32_t>(), nblocks);
    c->launches += 3;
    NTL_CUDA(c, cudaGetLastError());
    return NTL_OK;
}

}  // namespace ntl
