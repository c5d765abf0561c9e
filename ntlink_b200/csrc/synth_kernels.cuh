// synth_kernels.cuh -- synthetic benchmark inputs generated on the device (included by capi.cu; logic in synth_logic.cuh).
#pragma once
#include "synth_logic.cuh"

namespace ntl {
namespace {

// one block per contig slice: plain coalesced byte stores
__global__ void __launch_bounds__(256) k_synth_contigs(uint64_t seed, const SynthContig* __restrict__ ctg, const uint64_t* __restrict__ off,
                                                       uint32_t ncontig, uint8_t* __restrict__ seq) {
    for (uint32_t c = blockIdx.y; c < ncontig; c += gridDim.y) {
        const SynthContig sc = ctg[c];
        uint8_t* dst = seq + off[c];
        for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < sc.len; j += gridDim.x * blockDim.x) dst[j] = contig_base(seed, sc, j);
    }
}

// output length of every read: one warp per read
__global__ void __launch_bounds__(256) k_synth_read_len(uint64_t seed, const SynthRead* __restrict__ rd, uint32_t nreads, SynthErr e,
                                                        uint32_t* __restrict__ out_len) {
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= nreads) return;
    const SynthRead sr = rd[r];
    const uint64_t key = read_key(seed, sr.id);
    uint32_t n = 0;
    for (uint32_t j = lane; j < sr.len; j += 32) n += read_emit_count(key, j, e);
#pragma unroll
    for (int d = 16; d; d >>= 1) n += __shfl_xor_sync(0xffffffffu, n, d);
    if (lane == 0) out_len[r] = n;
}

// one warp per read; the source is walked in blocks of 32 lanes x SEG positions, every lane first counts what its
// segment emits, a warp scan places the segments, then the lane writes its bytes
constexpr uint32_t SYNTH_SEG = 32;
__global__ void __launch_bounds__(256) k_synth_reads(uint64_t seed, const SynthRead* __restrict__ rd, const uint64_t* __restrict__ off,
                                                     uint32_t nreads, SynthErr e, uint8_t* __restrict__ seq) {
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= nreads) return;
    const SynthRead sr = rd[r];
    const uint64_t key = read_key(seed, sr.id);
    uint8_t* dst = seq + off[r];
    uint64_t done = 0;
    for (uint32_t b0 = 0; b0 < sr.len; b0 += 32 * SYNTH_SEG) {
        const uint32_t ja = min(sr.len, b0 + lane * SYNTH_SEG), jb = min(sr.len, ja + SYNTH_SEG);
        uint32_t n = 0;
        for (uint32_t j = ja; j < jb; j++) n += read_emit_count(key, j, e);
        uint32_t x = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= (uint32_t)d) x += y; }
        uint8_t* p = dst + done + (x - n);
        for (uint32_t j = ja; j < jb; j++) {
            uint8_t o[2];
            const uint32_t m = read_emit(seed, key, sr, j, e, o);
            if (m) *p++ = o[0];
            if (m == 2) *p++ = o[1];
        }
        done += __shfl_sync(0xffffffffu, x, 31);
    }
}

}  // namespace
}  // namespace ntl
