// synth_logic.cuh -- deterministic synthetic genomes, assemblies and long reads, counter-based so that the
// same bytes come out of the CUDA kernels (synth.cu: inputs generated straight into HBM for the benchmark configurations,
// 4-93 Gbp of reads) and of the host functions (tests, the CPU reference arm). Nothing here is part of the mapping path.
//
//   genome    base i = 2 bits of mix64(seed, i / 32): iid uniform ACGT, never materialised
//   contig    genome[start, start+len), optionally reverse-complemented, optionally with one run of N inside
//   read      genome[start, start+len), optionally reverse-complemented, then per SOURCE position j one draw
//             u = mix64(read key, j):  substitute / delete / insert-after / copy  with 16-bit thresholds (ONT-like iid errors)
#pragma once
#include <stdint.h>
#include "nthash.cuh"

namespace ntl {

NTL_HD uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}
NTL_HD uint32_t genome_code(uint64_t seed, uint64_t i) {
    const uint64_t word = mix64(seed + 0x9e3779b97f4a7c15ULL * ((i >> 5) + 1));
    return (uint32_t)(word >> (2 * (i & 31))) & 3u;
}
NTL_HD uint8_t code_ascii(uint32_t c) { return (uint8_t)((0x54474341u >> (8 * c)) & 0xFFu); }   // "ACGT"

struct SynthContig { uint64_t start; uint32_t len, flip, n_start, n_len; uint32_t pad[2]; };      // 32 bytes
struct SynthRead { uint64_t start; uint32_t len, flip; uint64_t id; };                             // 24 bytes
struct SynthErr { uint32_t sub, del, ins; };     // thresholds out of 65536

NTL_HD uint32_t source_code(uint64_t seed, uint64_t start, uint32_t len, uint32_t flip, uint32_t j) {
    return flip ? 3u - genome_code(seed, start + (len - 1 - j)) : genome_code(seed, start + j);
}
NTL_HD uint8_t contig_base(uint64_t seed, const SynthContig& c, uint32_t j) {
    if (j - c.n_start < c.n_len) return (uint8_t)'N';
    return code_ascii(source_code(seed, c.start, c.len, c.flip, j));
}
NTL_HD uint64_t read_key(uint64_t seed, uint64_t id) { return mix64(seed ^ (0xd1342543de82ef95ULL * (id + 1))); }

// what source position j of a read contributes: 0, 1 or 2 output bases
NTL_HD uint32_t read_emit(uint64_t seed, uint64_t key, const SynthRead& r, uint32_t j, const SynthErr& e, uint8_t out[2]) {
    const uint64_t u = mix64(key + 0x9e3779b97f4a7c15ULL * ((uint64_t)j + 1));
    const uint32_t t = (uint32_t)u & 0xFFFFu;
    uint32_t c = source_code(seed, r.start, r.len, r.flip, j);
    if (t < e.sub) { out[0] = code_ascii((c + 1u + (uint32_t)((u >> 16) % 3u)) & 3u); return 1; }
    if (t < e.sub + e.del) return 0;
    out[0] = code_ascii(c);
    if (t < e.sub + e.del + e.ins) { out[1] = code_ascii((uint32_t)(u >> 20) & 3u); return 2; }
    return 1;
}
NTL_HD uint32_t read_emit_count(uint64_t key, uint32_t j, const SynthErr& e) {
    const uint32_t t = (uint32_t)mix64(key + 0x9e3779b97f4a7c15ULL * ((uint64_t)j + 1)) & 0xFFFFu;
    if (t < e.sub) return 1;
    if (t < e.sub + e.del) return 0;
    return t < e.sub + e.del + e.ins ? 2u : 1u;
}

}  // namespace ntl
