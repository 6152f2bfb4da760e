// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into the product library.
//
// XXH3-64 (seeded) restated from the published xxHash specification (xxHash >= 0.8,
// "XXH3_64bits_withSeed").  The reference calls the third-party crate `xxh3 = "0.1.1"`
// (Cargo.toml:9) as `xxh3::hash64_with_seed(bytes, seed)` at simple_bloom.rs:22,30,
// batch_search_pe.rs:49,129, perfect_search.rs:29,86 and read_id_mt_pe.rs:79,120,144.
// The crate is NOT vendored under /root/reference and no reference test asserts a hash
// value, so hash parity with a real colorid binary is UNPINNED.
// What is pinned: this restatement == python `xxhash` 3.7.0 (libxxhash 0.8.2) for
// lengths 0..240 (tests/test_oracle_golden.py::test_xxh3_*).
//
// `variant` (0 = the stable algorithm above): the crate predates the XXH3 freeze, so the places where the published
// drafts of XXH3 are known to differ are switchable, one bit each, for inputs of 1..128 bytes (every k-mer / minimizer):
//   1  final avalanche multiplier PRIME64_3 (0.7.x) instead of PRIME_MX1     2  avalanche shift 29 (0.7.0) instead of 37
//   4  128-bit product folded by + (0.7.0) instead of ^                      8  seed enters once as (len + seed) * PRIME64_1
//   16 secret read as 32-bit words (kKey[] of 0.7.0)                            instead of secret +/- seed per lane (17+ bytes)
// 1 = the 0.7.1-0.7.3 drafts as recalled, 31 = the 0.7.0 draft as recalled.  None of them is pinned either: they exist so
// that a .bxi written by a real colorid binary can be matched (tools/pin_from_bxi.py).
#pragma once
#include <cstdint>
#include <cstring>
#include <cstddef>

namespace orc {

static const uint8_t kXxh3Secret[192] = {
    0xb8, 0xfe, 0x6c, 0x39, 0x23, 0xa4, 0x4b, 0xbe, 0x7c, 0x01, 0x81, 0x2c, 0xf7, 0x21, 0xad, 0x1c,
    0xde, 0xd4, 0x6d, 0xe9, 0x83, 0x90, 0x97, 0xdb, 0x72, 0x40, 0xa4, 0xa4, 0xb7, 0xb3, 0x67, 0x1f,
    0xcb, 0x79, 0xe6, 0x4e, 0xcc, 0xc0, 0xe5, 0x78, 0x82, 0x5a, 0xd0, 0x7d, 0xcc, 0xff, 0x72, 0x21,
    0xb8, 0x08, 0x46, 0x74, 0xf7, 0x43, 0x24, 0x8e, 0xe0, 0x35, 0x90, 0xe6, 0x81, 0x3a, 0x26, 0x4c,
    0x3c, 0x28, 0x52, 0xbb, 0x91, 0xc3, 0x00, 0xcb, 0x88, 0xd0, 0x65, 0x8b, 0x1b, 0x53, 0x2e, 0xa3,
    0x71, 0x64, 0x48, 0x97, 0xa2, 0x0d, 0xf9, 0x4e, 0x38, 0x19, 0xef, 0x46, 0xa9, 0xde, 0xac, 0xd8,
    0xa8, 0xfa, 0x76, 0x3f, 0xe3, 0x9c, 0x34, 0x3f, 0xf9, 0xdc, 0xbb, 0xc7, 0xc7, 0x0b, 0x4f, 0x1d,
    0x8a, 0x51, 0xe0, 0x4b, 0xcd, 0xb4, 0x59, 0x31, 0xc8, 0x9f, 0x7e, 0xc9, 0xd9, 0x78, 0x73, 0x64,
    0xea, 0xc5, 0xac, 0x83, 0x34, 0xd3, 0xeb, 0xc3, 0xc5, 0x81, 0xa0, 0xff, 0xfa, 0x13, 0x63, 0xeb,
    0x17, 0x0d, 0xdd, 0x51, 0xb7, 0xf0, 0xda, 0x49, 0xd3, 0x16, 0x55, 0x26, 0x29, 0xd4, 0x68, 0x9e,
    0x2b, 0x16, 0xbe, 0x58, 0x7d, 0x47, 0xa1, 0xfc, 0x8f, 0xf8, 0xb8, 0xd1, 0x7a, 0xd0, 0x31, 0xce,
    0x45, 0xcb, 0x3a, 0x8f, 0x95, 0x16, 0x04, 0x28, 0xaf, 0xd7, 0xfb, 0xca, 0xbb, 0x4b, 0x40, 0x7e,
};

static const uint64_t P64_1 = 0x9E3779B185EBCA87ULL;
static const uint64_t P64_2 = 0xC2B2AE3D27D4EB4FULL;
static const uint64_t P64_3 = 0x165667B19E3779F9ULL;
static const uint64_t P_MX1 = 0x165667919E3779F9ULL;  // XXH3 avalanche multiplier
static const uint64_t P_MX2 = 0x9FB21C651E98DF25ULL;  // rrmxmx multiplier

static inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t mul128_fold64(uint64_t a, uint64_t b) {
    unsigned __int128 p = (unsigned __int128)a * b;
    return (uint64_t)p ^ (uint64_t)(p >> 64);
}
static inline uint64_t xxh64_avalanche(uint64_t h) {
    h ^= h >> 33; h *= P64_2; h ^= h >> 29; h *= P64_3; h ^= h >> 32; return h;
}
static inline uint64_t xxh3_avalanche(uint64_t h) {
    h ^= h >> 37; h *= P_MX1; h ^= h >> 32; return h;
}
static inline uint64_t mix16(const uint8_t* in, const uint8_t* sec, uint64_t seed) {
    return mul128_fold64(rd64(in) ^ (rd64(sec) + seed), rd64(in + 8) ^ (rd64(sec + 8) - seed));
}

struct Xxh3Variant {
    uint32_t v;
    uint64_t sec64(const uint8_t* p) const {          // bit 4: the secret's bytes as an array of u32 constants
        if (!(v & 16u)) return rd64(p);
        return ((uint64_t)__builtin_bswap32(rd32(p + 4)) << 32) | __builtin_bswap32(rd32(p));
    }
    uint64_t fold(uint64_t a, uint64_t b) const {
        unsigned __int128 p = (unsigned __int128)a * b;
        return (v & 4u) ? (uint64_t)p + (uint64_t)(p >> 64) : (uint64_t)p ^ (uint64_t)(p >> 64);
    }
    uint64_t aval(uint64_t h) const { h ^= h >> ((v & 2u) ? 29 : 37); h *= (v & 1u) ? P64_3 : P_MX1; h ^= h >> 32; return h; }
    uint64_t mix(const uint8_t* in, const uint8_t* sec, uint64_t seed) const {
        const uint64_t sd = (v & 8u) ? 0 : seed;
        return fold(rd64(in) ^ (sec64(sec) + sd), rd64(in + 8) ^ (sec64(sec + 8) - sd));
    }
};
static inline uint64_t xxh3_64_variant(const uint8_t* in, size_t len, uint64_t seed, uint32_t variant) {
    const Xxh3Variant V{variant};
    const uint8_t* s = kXxh3Secret;
    if (len == 0 || len > 128) return 0;           // not produced by any k-mer path
    if (len <= 3) {
        uint8_t c1 = in[0], c2 = in[len >> 1], c3 = in[len - 1];
        uint32_t combined = ((uint32_t)c1 << 16) | ((uint32_t)c2 << 24) | (uint32_t)c3 | ((uint32_t)len << 8);
        const uint64_t s0 = V.sec64(s);
        uint64_t bitflip = (uint64_t)((uint32_t)s0 ^ (uint32_t)(s0 >> 32)) + seed;
        return xxh64_avalanche((uint64_t)combined ^ bitflip);
    }
    if (len <= 8) {
        seed ^= (uint64_t)__builtin_bswap32((uint32_t)seed) << 32;
        uint32_t in1 = rd32(in), in2 = rd32(in + len - 4);
        uint64_t bitflip = (V.sec64(s + 8) ^ V.sec64(s + 16)) - seed;
        uint64_t h = ((uint64_t)in2 + ((uint64_t)in1 << 32)) ^ bitflip;
        h ^= rotl64(h, 49) ^ rotl64(h, 24);
        h *= P_MX2;
        h ^= (h >> 35) + len;
        h *= P_MX2;
        return h ^ (h >> 28);
    }
    if (len <= 16) {
        uint64_t bf1 = (V.sec64(s + 24) ^ V.sec64(s + 32)) + seed;
        uint64_t bf2 = (V.sec64(s + 40) ^ V.sec64(s + 48)) - seed;
        uint64_t lo = rd64(in) ^ bf1, hi = rd64(in + len - 8) ^ bf2;
        return V.aval(len + __builtin_bswap64(lo) + hi + V.fold(lo, hi));
    }
    uint64_t acc = (variant & 8u) ? (len + seed) * P64_1 : len * P64_1;
    if (len > 32) {
        if (len > 64) {
            if (len > 96) {
                acc += V.mix(in + 48, s + 96, seed);
                acc += V.mix(in + len - 64, s + 112, seed);
            }
            acc += V.mix(in + 32, s + 64, seed);
            acc += V.mix(in + len - 48, s + 80, seed);
        }
        acc += V.mix(in + 16, s + 32, seed);
        acc += V.mix(in + len - 32, s + 48, seed);
    }
    acc += V.mix(in, s, seed);
    acc += V.mix(in + len - 16, s + 16, seed);
    return V.aval(acc);
}

// Lengths 0..240 (every k the k-mer paths can produce); longer inputs are out of scope.
static inline uint64_t xxh3_64_with_seed(const uint8_t* in, size_t len, uint64_t seed, uint32_t variant = 0) {
    if (variant) return xxh3_64_variant(in, len, seed, variant);
    const uint8_t* s = kXxh3Secret;
    if (len == 0) return xxh64_avalanche(seed ^ (rd64(s + 56) ^ rd64(s + 64)));
    if (len <= 3) {
        uint8_t c1 = in[0], c2 = in[len >> 1], c3 = in[len - 1];
        uint32_t combined = ((uint32_t)c1 << 16) | ((uint32_t)c2 << 24) | (uint32_t)c3 | ((uint32_t)len << 8);
        uint64_t bitflip = (uint64_t)(rd32(s) ^ rd32(s + 4)) + seed;
        return xxh64_avalanche((uint64_t)combined ^ bitflip);
    }
    if (len <= 8) {
        seed ^= (uint64_t)__builtin_bswap32((uint32_t)seed) << 32;
        uint32_t in1 = rd32(in), in2 = rd32(in + len - 4);
        uint64_t bitflip = (rd64(s + 8) ^ rd64(s + 16)) - seed;
        uint64_t h = ((uint64_t)in2 + ((uint64_t)in1 << 32)) ^ bitflip;
        h ^= rotl64(h, 49) ^ rotl64(h, 24);
        h *= P_MX2;
        h ^= (h >> 35) + len;
        h *= P_MX2;
        return h ^ (h >> 28);
    }
    if (len <= 16) {
        uint64_t bf1 = (rd64(s + 24) ^ rd64(s + 32)) + seed;
        uint64_t bf2 = (rd64(s + 40) ^ rd64(s + 48)) - seed;
        uint64_t lo = rd64(in) ^ bf1, hi = rd64(in + len - 8) ^ bf2;
        uint64_t acc = len + __builtin_bswap64(lo) + hi + mul128_fold64(lo, hi);
        return xxh3_avalanche(acc);
    }
    if (len <= 128) {
        uint64_t acc = len * P64_1;
        if (len > 32) {
            if (len > 64) {
                if (len > 96) {
                    acc += mix16(in + 48, s + 96, seed);
                    acc += mix16(in + len - 64, s + 112, seed);
                }
                acc += mix16(in + 32, s + 64, seed);
                acc += mix16(in + len - 48, s + 80, seed);
            }
            acc += mix16(in + 16, s + 32, seed);
            acc += mix16(in + len - 32, s + 48, seed);
        }
        acc += mix16(in, s, seed);
        acc += mix16(in + len - 16, s + 16, seed);
        return xxh3_avalanche(acc);
    }
    if (len <= 240) {
        uint64_t acc = len * P64_1;
        size_t rounds = len / 16;
        for (size_t i = 0; i < 8; i++) acc += mix16(in + 16 * i, s + 16 * i, seed);
        acc = xxh3_avalanche(acc);
        for (size_t i = 8; i < rounds; i++) acc += mix16(in + 16 * i, s + 16 * (i - 8) + 3, seed);
        acc += mix16(in + len - 16, s + 136 - 17, seed);
        return xxh3_avalanche(acc);
    }
    return 0;  // not reachable from the k-mer paths (k <= 240); callers reject longer k
}

}  // namespace orc
