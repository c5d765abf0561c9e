// nthash.cuh -- ntHash arithmetic shared by the CUDA kernels and the host-side table builder.
//
// Restates btllib's ntHash (the reference's external sketcher `indexlr`, invoked at ntLink:198-199,221-225;
// SURVEY.md 8a S1-S2): 64-bit words whose low 33 bits and high 31 bits rotate separately, canonical
// hash = forward + reverse-complement, printed hash = ntHash "extra hash" number 1.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define NTL_HD __host__ __device__ __forceinline__
#else
#define NTL_HD inline
#endif

namespace ntl {

// base codes used everywhere on the device: A=0 C=1 G=2 T=3, 4 = invalid (N, IUPAC, anything else)
enum : uint32_t { CODE_INVALID = 4 };

NTL_HD uint64_t seed_of(uint32_t code) {
    switch (code) {
        case 0: return 0x3c8bfbb395c60474ULL;   // A
        case 1: return 0x3193c18562a02b4cULL;   // C
        case 2: return 0x20323ed082572324ULL;   // G
        case 3: return 0x295549f54be24456ULL;   // T
        default: return 0;                      // invalid bases contribute nothing
    }
}
NTL_HD uint64_t seed_rc_of(uint32_t code) { return code < 4 ? seed_of(3 - code) : 0; }

// split rotate left by one: bit32 -> bit0, bit63 -> bit33
NTL_HD uint64_t srol1(uint64_t x) {
    uint64_t m = ((x & 0x8000000000000000ULL) >> 30) | ((x & 0x100000000ULL) >> 32);
    return ((x << 1) & 0xFFFFFFFDFFFFFFFEULL) | m;
}
// split rotate right by one (inverse of srol1): bit0 -> bit32, bit33 -> bit63
NTL_HD uint64_t sror1(uint64_t x) {
    uint64_t m = ((x & 0x200000000ULL) << 30) | ((x & 1ULL) << 32);
    return ((x >> 1) & 0xFFFFFFFEFFFFFFFFULL) | m;
}
// split rotate left by d: the 33-bit part by d mod 33, the 31-bit part by d mod 31
NTL_HD uint64_t sroln(uint64_t x, uint32_t d) {
    const uint64_t M33 = (1ULL << 33) - 1, M31 = (1ULL << 31) - 1;
    uint64_t lo = x & M33, hi = x >> 33;
    uint32_t a = d % 33, b = d % 31;
    if (a) lo = ((lo << a) | (lo >> (33 - a))) & M33;
    if (b) hi = ((hi << b) | (hi >> (31 - b))) & M31;
    return lo | (hi << 33);
}

NTL_HD uint64_t second_hash_multiplier(uint32_t k) { return 1ULL ^ ((uint64_t)k * 0x90b45d39fb6da1faULL); }
NTL_HD uint64_t second_hash(uint64_t h0, uint64_t mult) {
    uint64_t t = h0 * mult;
    return t ^ (t >> 27);
}

// One entry of the roll table, indexed by (code_in << 3) | code_out, codes 0..4:
//   f = seed[in] ^ srol^k(seed[out])            forward:  fh' = srol(fh) ^ f
//   r = srol^k(seed_rc[in]) ^ seed_rc[out]      reverse:  rh' = sror(rh ^ r)
// Invalid bases (code 4) contribute zero, so a rolling hash that passed over an N is exact again as soon
// as the N has left the k-mer (no re-initialisation needed); k-mers containing an N are masked by position.
struct RollEntry { uint64_t f, r; };
enum : uint32_t { ROLL_TABLE_ENTRIES = 64 };

inline void build_roll_table(uint32_t k, RollEntry* tbl /* [64] */) {
    for (uint32_t i = 0; i < ROLL_TABLE_ENTRIES; i++) { tbl[i].f = 0; tbl[i].r = 0; }
    for (uint32_t in = 0; in <= 4; in++)
        for (uint32_t out = 0; out <= 4; out++) {
            RollEntry e;
            e.f = seed_of(in) ^ sroln(seed_of(out), k);
            e.r = sroln(seed_rc_of(in), k) ^ seed_rc_of(out);
            tbl[(in << 3) | out] = e;
        }
}

}  // namespace ntl
