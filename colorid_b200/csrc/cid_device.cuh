// Device-side building blocks shared by every kernel of the BIGSI hot path (sm_100a).
//
//  * XXH3-64 (seeded) of a canonical k-mer's ASCII bytes and `% bloom_size` — the reference's
//    hash/seed scheme, simple_bloom.rs:19-26: bit = xxh3::hash64_with_seed(kmer, i) % len.
//  * A shared-memory "tile" of bases (ASCII + 2-bit codes + validity/case/boundary bitmasks) from
//    which every k-mer window of kmer.rs is extracted: has_no_n (seq.rs:66-70), canonical choice
//    `fwd < revcomp ? fwd : revcomp` on raw bytes (kmer.rs:104), upper-casing (kmer.rs:106).
//  * FNV-1a low bits for the hashbrown iteration-order emulation used by read_id.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cid {

// ------------------------------------------------------------------ XXH3-64 with seed, len <= 32
// Secret words (little-endian u64 at byte offsets 0..64 of the default XXH3 secret).
#define CID_SEC0 0xbe4ba423396cfeb8ULL
#define CID_SEC8 0x1cad21f72c81017cULL
#define CID_SEC16 0xdb979083e96dd4deULL
#define CID_SEC24 0x1f67b3b7a4a44072ULL
#define CID_SEC32 0x78e5c0cc4ee679cbULL
#define CID_SEC40 0x2172ffcc7dd05a82ULL
#define CID_SEC48 0x8e2443f7744608b8ULL
#define CID_P64_1 0x9E3779B185EBCA87ULL
#define CID_P64_2 0xC2B2AE3D27D4EB4FULL
#define CID_P64_3 0x165667B19E3779F9ULL
#define CID_PMX1 0x165667919E3779F9ULL
#define CID_PMX2 0x9FB21C651E98DF25ULL

__device__ __forceinline__ uint64_t mul128_fold64(uint64_t a, uint64_t b) { return (a * b) ^ __umul64hi(a, b); }
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ uint64_t bswap64(uint64_t x) {
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
}

// The bytes of a k-mer that XXH3 actually reads, as little-endian words:
//   k >= 17: w0=in[0:8] w1=in[8:16] w2=in[k-16:k-8] w3=in[k-8:k]
//   9..16 : w0=in[0:8] w1=in[k-8:k]
//   4..8  : w0=in[0:4] w1=in[k-4:k]           (32-bit values)
//   1..3  : w0=in[0]   w1=in[k>>1] w2=in[k-1]
struct HashIn { uint64_t w0, w1, w2, w3; };

// Hash variants.  The reference pins the third-party crate `xxh3 = "0.1.1"` (Cargo.toml:9), which predates the XXH3 freeze and
// is not vendored, so which draft of XXH3 a real colorid binary computes is unpinned.  Variant 0 is
// stable XXH3 (xxHash >= 0.8); the other 31 are the combinations of the five places where the drafts are known to differ,
// one bit each, so that an index written by a real binary can be matched without touching a kernel (tools/pin_from_bxi.py):
//   bit 0  final avalanche multiplier PRIME64_3 instead of PRIME_MX1      bit 1  avalanche shift 29 instead of 37
//   bit 2  128-bit product folded by + instead of ^                       bit 3  seed enters as (len + seed) * PRIME64_1
//   bit 4  secret read as 32-bit words (kKey[] of the first draft)               instead of secret +/- seed per lane (len >= 17)
// The variant is resolved on the host into a HashCfg kept in device memory next to the index; only the out-of-line variant
// path reads it, variant 0 keeps its immediates.
#define CID_HASH_VARIANTS 32
struct HashCfg {
    uint64_t sec[7];        // secret words at byte offsets 0, 8, .., 48 (32-bit halves byte-swapped for bit 4)
    uint64_t av_mul;
    uint32_t av_shift, fold_add, seed_acc, var;
};
// h % S for run-time S: see mod_s below.  `cfg` points to the index's HashCfg in DEVICE memory (only the out-of-line
// variant path reads it), so a kernel that hashes never needs the address of one of its own parameters.
struct ModS { uint64_t S, M; uint32_t var; const HashCfg* cfg; };
__host__ __device__ __forceinline__ ModS make_mods(uint64_t S, uint32_t var = 0, const HashCfg* cfg = nullptr) {
    ModS m; m.S = S; m.var = var; m.cfg = cfg;
    m.M = S <= 1 ? 0 : (uint64_t)((((unsigned __int128)1) << 64) / S);
    return m;
}
inline HashCfg make_hashcfg(uint32_t var) {
    HashCfg h;
    const uint64_t sec[7] = {CID_SEC0, CID_SEC8, CID_SEC16, CID_SEC24, CID_SEC32, CID_SEC40, CID_SEC48};
    for (int i = 0; i < 7; i++) {
        uint64_t s = sec[i];
        if (var & 16u) {     // each 32-bit half byte-swapped: the secret's bytes read as an array of u32 constants
            auto sw = [](uint32_t x) { return (x >> 24) | ((x >> 8) & 0xFF00u) | ((x << 8) & 0xFF0000u) | (x << 24); };
            s = ((uint64_t)sw((uint32_t)(s >> 32)) << 32) | sw((uint32_t)s);
        }
        h.sec[i] = s;
    }
    h.av_mul = (var & 1u) ? CID_P64_3 : CID_PMX1;
    h.av_shift = (var & 2u) ? 29u : 37u;
    h.fold_add = (var & 4u) ? 1u : 0u;
    h.seed_acc = (var & 8u) ? 1u : 0u;
    h.var = var;
    return h;
}

// stable XXH3 (variant 0): every constant an immediate
__device__ __forceinline__ uint64_t xxh3_kmer_stable(const HashIn& in, uint32_t k, uint64_t seed) {
    if (k >= 17) {
        uint64_t acc = (uint64_t)k * CID_P64_1;
        acc += mul128_fold64(in.w0 ^ (CID_SEC0 + seed), in.w1 ^ (CID_SEC8 - seed));
        acc += mul128_fold64(in.w2 ^ (CID_SEC16 + seed), in.w3 ^ (CID_SEC24 - seed));
        acc ^= acc >> 37; acc *= CID_PMX1; acc ^= acc >> 32;
        return acc;
    } else if (k >= 9) {
        uint64_t lo = in.w0 ^ ((CID_SEC24 ^ CID_SEC32) + seed);
        uint64_t hi = in.w1 ^ ((CID_SEC40 ^ CID_SEC48) - seed);
        uint64_t acc = (uint64_t)k + bswap64(lo) + hi + mul128_fold64(lo, hi);
        acc ^= acc >> 37; acc *= CID_PMX1; acc ^= acc >> 32;
        return acc;
    } else if (k >= 4) {
        uint32_t s32 = (uint32_t)seed;
        seed ^= (uint64_t)__byte_perm(s32, 0, 0x0123) << 32;
        uint64_t h = (in.w1 + (in.w0 << 32)) ^ ((CID_SEC8 ^ CID_SEC16) - seed);
        h ^= rotl64(h, 49) ^ rotl64(h, 24);
        h *= CID_PMX2;
        h ^= (h >> 35) + k;
        h *= CID_PMX2;
        return h ^ (h >> 28);
    } else {
        uint32_t combined = ((uint32_t)in.w0 << 16) | ((uint32_t)in.w1 << 24) | (uint32_t)in.w2 | (k << 8);
        uint64_t h = (uint64_t)combined ^ ((uint64_t)(0x396cfeb8u ^ 0xbe4ba423u) + seed);
        h ^= h >> 33; h *= CID_P64_2; h ^= h >> 29; h *= CID_P64_3; h ^= h >> 32;
        return h;
    }
}
// any variant, from the resolved configuration (rarely run: kept out of line so that the kernels' register budgets are
// those of the stable path)
static __device__ __noinline__ uint64_t xxh3_kmer_cfg(uint64_t w0, uint64_t w1, uint64_t w2, uint64_t w3, uint32_t k, uint64_t seed, const HashCfg* cp) {
    const HashCfg& c = *cp;
    auto fold = [&](uint64_t a, uint64_t b) { const uint64_t lo = a * b, hi = __umul64hi(a, b); return c.fold_add ? lo + hi : lo ^ hi; };
    auto aval = [&](uint64_t h) { h ^= h >> c.av_shift; h *= c.av_mul; h ^= h >> 32; return h; };
    if (k >= 17) {
        const uint64_t sd = c.seed_acc ? 0ull : seed;
        uint64_t acc = ((uint64_t)k + (c.seed_acc ? seed : 0ull)) * CID_P64_1;
        acc += fold(w0 ^ (c.sec[0] + sd), w1 ^ (c.sec[1] - sd));
        acc += fold(w2 ^ (c.sec[2] + sd), w3 ^ (c.sec[3] - sd));
        return aval(acc);
    } else if (k >= 9) {
        const uint64_t lo = w0 ^ ((c.sec[3] ^ c.sec[4]) + seed);
        const uint64_t hi = w1 ^ ((c.sec[5] ^ c.sec[6]) - seed);
        return aval((uint64_t)k + bswap64(lo) + hi + fold(lo, hi));
    } else if (k >= 4) {
        const uint32_t s32 = (uint32_t)seed;
        seed ^= (uint64_t)__byte_perm(s32, 0, 0x0123) << 32;
        uint64_t h = (w1 + (w0 << 32)) ^ ((c.sec[1] ^ c.sec[2]) - seed);
        h ^= rotl64(h, 49) ^ rotl64(h, 24);
        h *= CID_PMX2;
        h ^= (h >> 35) + k;
        h *= CID_PMX2;
        return h ^ (h >> 28);
    } else {
        const uint32_t combined = ((uint32_t)w0 << 16) | ((uint32_t)w1 << 24) | (uint32_t)w2 | (k << 8);
        uint64_t h = (uint64_t)combined ^ ((uint64_t)((uint32_t)c.sec[0] ^ (uint32_t)(c.sec[0] >> 32)) + seed);
        h ^= h >> 33; h *= CID_P64_2; h ^= h >> 29; h *= CID_P64_3; h ^= h >> 32;
        return h;
    }
}
__device__ __forceinline__ uint64_t xxh3_kmer(const HashIn& in, uint32_t k, uint64_t seed, const ModS& m) {
    if (m.var == 0u) return xxh3_kmer_stable(in, k, seed);
    return xxh3_kmer_cfg(in.w0, in.w1, in.w2, in.w3, k, seed, m.cfg);
}

// h % S for run-time S (2 <= S < 2^63) with M = floor(2^64 / S): q = mulhi(h, M) is floor(h/S) or
// one less, so a single conditional subtract makes it exact.
__device__ __forceinline__ uint64_t mod_s(uint64_t h, const ModS& m) {
    if (m.S <= 1) return 0;
    uint64_t q = __umul64hi(h, m.M);
    uint64_t r = h - q * m.S;
    return r >= m.S ? r - m.S : r;
}

// Row index of hash function `seed` (simple_bloom.rs:21-24): xxh3(kmer, seed) % bloom_size, in the index's hash variant.
__device__ __forceinline__ uint64_t hash64(const HashIn& in, uint32_t k, uint64_t seed, const ModS& m) {
    return xxh3_kmer(in, k, seed, m);
}
__device__ __forceinline__ uint64_t hash_row(const HashIn& in, uint32_t k, uint64_t seed, const ModS& m) {
    return mod_s(hash64(in, k, seed, m), m);
}

// ------------------------------------------------------------------ 2-bit codes <-> ASCII
// code = A0 C1 G2 T3 (byte order == numeric order for upper case)
__device__ __forceinline__ uint32_t base_code(uint32_t c) { return ((c >> 1) ^ (c >> 2)) & 3u; }
__device__ __forceinline__ bool base_is_acgt(uint32_t c) {
    uint32_t u = c & 0xDFu;
    return u == 'A' || u == 'C' || u == 'G' || u == 'T';
}
__device__ __forceinline__ uint32_t code_ascii(uint32_t code) { return 'A' + ((0x13060200u >> (code * 8)) & 0xFFu); }

// 256-entry table: 4 bases (first base in bits 7:6) -> 4 upper-case ASCII bytes (first base in byte 0).
__device__ __forceinline__ void lut4_init(uint32_t* lut, int tid, int nthreads) {
    for (int i = tid; i < 256; i += nthreads)
        lut[i] = code_ascii((i >> 6) & 3) | (code_ascii((i >> 4) & 3) << 8) | (code_ascii((i >> 2) & 3) << 16) |
                 (code_ascii(i & 3) << 24);
}
__device__ __forceinline__ uint64_t ascii8(const uint32_t* lut, uint32_t c16) {
    return (uint64_t)lut[(c16 >> 8) & 0xFF] | ((uint64_t)lut[c16 & 0xFF] << 32);
}
// key: k bases, first base in the top of the low 2k bits.  Upper-case ASCII as XXH3 reads it.
__device__ __forceinline__ HashIn hashin_from_key(const uint32_t* lut, uint64_t key, uint32_t k) {
    HashIn in;
    uint64_t kk = key << (64 - 2 * k);   // left-aligned: first base in bits 63:62
    if (k >= 17) {
        in.w0 = ascii8(lut, (uint32_t)(kk >> 48));
        in.w1 = ascii8(lut, (uint32_t)(kk >> 32) & 0xFFFF);
        in.w2 = ascii8(lut, (uint32_t)(key >> 16) & 0xFFFF);
        in.w3 = ascii8(lut, (uint32_t)key & 0xFFFF);
    } else if (k >= 9) {
        in.w0 = ascii8(lut, (uint32_t)(kk >> 48));
        in.w1 = ascii8(lut, (uint32_t)key & 0xFFFF);
        in.w2 = in.w3 = 0;
    } else if (k >= 4) {
        in.w0 = lut[(uint32_t)(kk >> 56)];
        in.w1 = lut[(uint32_t)key & 0xFF];
        in.w2 = in.w3 = 0;
    } else {
        in.w0 = code_ascii((uint32_t)(kk >> 62));
        in.w1 = code_ascii((uint32_t)(key >> (2 * (k - 1 - (k >> 1)))) & 3);
        in.w2 = code_ascii((uint32_t)key & 3);
        in.w3 = 0;
    }
    return in;
}

// Raw-case inputs (kmers_from_fq_qual / kmers_fq_pe_qual, kmer.rs:461-510,581-655: no to_uppercase) keep lower-case bases in
// their k-mers: such a k-mer is the packed codes plus a case mask `cs` (bit j = byte j of the canonical string is lower case).
// Lower-case ASCII = upper-case | 0x20: spread 8 mask bits over the 8 bytes of a word.
__device__ __forceinline__ uint64_t case_bytes8(uint32_t bits8) {
    const uint64_t t = ((uint64_t)(bits8 & 0xFFu) * 0x0101010101010101ULL) & 0x8040201008040201ULL;    // byte i keeps bit i
    return (((t + 0x7F7F7F7F7F7F7F7FULL) >> 7) & 0x0101010101010101ULL) * 0x20ULL;
}
__device__ __forceinline__ void hashin_apply_case(HashIn& in, uint32_t cs, uint32_t k) {
    if (cs == 0u) return;
    if (k >= 17) {
        in.w0 |= case_bytes8(cs); in.w1 |= case_bytes8(cs >> 8);
        in.w2 |= case_bytes8(cs >> (k - 16)); in.w3 |= case_bytes8(cs >> (k - 8));
    } else if (k >= 9) {
        in.w0 |= case_bytes8(cs); in.w1 |= case_bytes8(cs >> (k - 8));
    } else if (k >= 4) {
        in.w0 |= case_bytes8(cs & 0xFu) & 0xFFFFFFFFULL; in.w1 |= case_bytes8((cs >> (k - 4)) & 0xFu) & 0xFFFFFFFFULL;
    } else {
        in.w0 |= (cs & 1u) << 5; in.w1 |= ((cs >> (k >> 1)) & 1u) << 5; in.w2 |= ((cs >> (k - 1)) & 1u) << 5;
    }
}

// Low 32 bits of FNV-1a-64 over the k upper-case ASCII bytes of `key` followed by 0xFF
// (`impl Hash for str`).  The low 32 bits of h*0x100000001b3 depend only on the low 32 bits of h.
__device__ __forceinline__ uint32_t fnv1a_low32_key(uint64_t key, uint32_t k) {
    uint32_t h = 0x84222325u;   // low 32 bits of 0xcbf29ce484222325
    uint64_t kk = key << (64 - 2 * k);
    for (uint32_t j = 0; j < k; j++) {
        h = (h ^ code_ascii((uint32_t)(kk >> 62))) * 0x1b3u;
        kk <<= 2;
    }
    return (h ^ 0xFFu) * 0x1b3u;
}

// Same, 4 bases per shared-memory LUT lookup (3 ops per byte instead of 7).
__device__ __forceinline__ uint32_t fnv1a_low32_key_lut(const uint32_t* lut, uint64_t key, uint32_t k) {
    uint32_t h = 0x84222325u;
    uint64_t kk = key << (64 - 2 * k);
    uint32_t j = 0;
    for (; j + 4 <= k; j += 4) {
        uint32_t w = lut[(uint32_t)(kk >> 56)];
        kk <<= 8;
        h = (h ^ (w & 0xFFu)) * 0x1b3u;
        h = (h ^ ((w >> 8) & 0xFFu)) * 0x1b3u;
        h = (h ^ ((w >> 16) & 0xFFu)) * 0x1b3u;
        h = (h ^ (w >> 24)) * 0x1b3u;
    }
    if (j < k) {
        uint32_t w = lut[(uint32_t)(kk >> 56)];
        for (; j < k; j++) { h = (h ^ (w & 0xFFu)) * 0x1b3u; w >>= 8; }
    }
    return (h ^ 0xFFu) * 0x1b3u;
}

// Reverse complement of a packed k-mer.
__device__ __forceinline__ uint64_t revcomp_key(uint64_t v, uint32_t k) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    uint64_t r = ((uint64_t)__brev(lo) << 32) | __brev(hi);          // bit-reverse 64
    r = ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);  // restore bit order inside each base
    r >>= (64 - 2 * k);
    return r ^ (k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1));
}

// ------------------------------------------------------------------ shared-memory tile of bases
// Layout (all arrays zero-padded so that window reads never leave the allocation):
//   ascii [cap]           raw bytes (after optional quality masking)
//   codes [cap/16 + 3]    16 bases per u32, first base in bits 31:30
//   bad   [cap/32 + 2]    bit j (LSB-first) = base j is not one of ACGTacgt (or beyond len)
//   lower [cap/32 + 2]    bit j = base j is a lower-case acgt
//   start [cap/32 + 2]    bit j = a sequence starts at base j (windows may not cross it)
struct Tile {
    uint8_t* ascii;
    uint32_t* codes;
    uint32_t* bad;
    uint32_t* lower;
    uint32_t* start;
    int len;
};
__host__ __device__ constexpr size_t tile_smem_bytes(int cap) {   // cap multiple of 32
    return (size_t)cap + 4 * (size_t)(cap / 16 + 3) + 3 * 4 * (size_t)(cap / 32 + 2);
}
__device__ __forceinline__ Tile tile_carve(uint8_t* smem, int cap) {
    Tile t;
    t.ascii = smem;
    t.codes = (uint32_t*)(smem + cap);
    t.bad = t.codes + (cap / 16 + 3);
    t.lower = t.bad + (cap / 32 + 2);
    t.start = t.lower + (cap / 32 + 2);
    t.len = 0;
    return t;
}
// 4 bases (one little-endian u32, first base in byte 0) -> 8 code bits (first base in bits 7:6),
// and 4-bit LSB-first masks of "not ACGTacgt" / "lower-case acgt".
__device__ __forceinline__ void pack4(uint32_t w, uint32_t& pk, uint32_t& bad4, uint32_t& low4) {
    const uint32_t u = w & 0xDFDFDFDFu;                                   // fold case
    const uint32_t t = ((u >> 1) ^ (u >> 2)) & 0x03030303u;               // A0 C1 G2 T3 per byte
    const uint32_t c0 = t & 0x01010101u, c1 = (t >> 1) & 0x01010101u, cc = c0 & c1;
    // the upper-case letter that code stands for: 'A' + {0,2,6,19}
    const uint32_t expect = 0x41414141u + (c0 << 1) + (c1 << 2) + (c1 << 1) + (cc << 3) + (cc << 1) + cc;
    const uint32_t x = u ^ expect;
    const uint32_t nz = (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;   // byte != 0 (exact per byte)
    const uint32_t badf = nz >> 7;                                                // 0/1 per byte
    const uint32_t lowf = (w >> 5) & 0x01010101u & ~badf;
    bad4 = (badf * 0x01020408u) >> 24;
    low4 = (lowf * 0x01020408u) >> 24;
    pk = ((t << 6) | (t >> 4) | (t >> 14) | (t >> 24)) & 0xFFu;
}
// Build codes/bad/lower from ascii[0..len) (ascii already in smem; caller syncs before and after).
// Every thread packs 16 bases per step with SIMD-in-register byte arithmetic; nthreads must be a
// multiple of 32 (mask words are assembled from lane pairs).
__device__ __forceinline__ void tile_pack(Tile& t, int cap, int tid, int nthreads) {
    const int ngroups = (cap / 16 + 3 + 1) & ~1;            // 16-base groups incl. padding, even
    const uint32_t* a4 = (const uint32_t*)t.ascii;
    for (int g0 = 0; g0 < ngroups; g0 += nthreads) {
        const int g = g0 + tid;
        uint32_t codes = 0, bad16 = 0xFFFFu, low16 = 0;
        const int nvalid = t.len - g * 16;                   // bases of this group inside the tile
        if (g < ngroups && nvalid > 0) {
            uint32_t pk, b4, l4;
            bad16 = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                pack4(a4[g * 4 + q], pk, b4, l4);
                codes |= pk << (24 - 8 * q);
                bad16 |= b4 << (4 * q);
                low16 |= l4 << (4 * q);
            }
            if (nvalid < 16) {
                const uint32_t keep = (1u << nvalid) - 1;
                bad16 |= ~keep & 0xFFFFu;
                low16 &= keep;
            }
        }
        const uint32_t pb = __shfl_down_sync(0xffffffffu, bad16, 1), pl = __shfl_down_sync(0xffffffffu, low16, 1);
        if (g < ngroups) {
            if (g < cap / 16 + 3) t.codes[g] = codes;
            if (!(g & 1) && (g >> 1) < cap / 32 + 2) {
                t.bad[g >> 1] = bad16 | (pb << 16);
                t.lower[g >> 1] = low16 | (pl << 16);
            }
        }
    }
}
__device__ __forceinline__ uint32_t mask_window(const uint32_t* m, int i, uint32_t k) {   // k <= 32
    uint32_t lo = m[i >> 5], hi = m[(i >> 5) + 1];
    uint32_t v = __funnelshift_r(lo, hi, i & 31);
    return k == 32 ? v : (v & ((1u << k) - 1));
}
__device__ __forceinline__ uint64_t codes_window(const uint32_t* c, int i, uint32_t k) {
    int w = i >> 4, s = 2 * (i & 15);
    uint32_t w0 = c[w], w1 = c[w + 1], w2 = c[w + 2];
    uint32_t hi = __funnelshift_l(w1, w0, s), lo = __funnelshift_l(w2, w1, s);
    return (((uint64_t)hi << 32) | lo) >> (64 - 2 * k);
}

// Case-preserving complement of a valid base (kmer.rs:847-863 switch_base on acgtACGT).
__device__ __forceinline__ uint32_t comp_base(uint32_t c) {
    uint32_t u = c & 0xDFu, cs = c & 0x20u;
    uint32_t r = (u == 'A') ? 'T' : (u == 'C') ? 'G' : (u == 'G') ? 'C' : 'A';
    return r | cs;
}

// One k-mer window of the tile.  Returns false when the window is not a k-mer of the reference
// (contains a non-ACGT byte: seq.rs has_no_n; or crosses a sequence boundary).
//   key      : packed codes of the chosen strand (after upper-casing)
//   took_fwd : kmer.rs:104 `l[i..i+k] < l_r[..]` on RAW bytes; ties (palindromes) take the rc branch
//   has_lower: the window contains lower-case bases (raw-case modes cannot represent the key)
__device__ __forceinline__ bool tile_kmer(const Tile& t, int i, uint32_t k, uint64_t& key, bool& took_fwd, bool& has_lower) {
    if (i + (int)k > t.len) return false;
    if (mask_window(t.bad, i, k) != 0u) return false;
    if ((mask_window(t.start, i, k) & ~1u) != 0u) return false;
    uint64_t f = codes_window(t.codes, i, k);
    uint64_t r = revcomp_key(f, k);
    uint32_t low = mask_window(t.lower, i, k);
    has_lower = low != 0u;
    if (!has_lower) {
        took_fwd = f < r;
    } else {
        // exact raw-byte comparison of fwd against its case-preserving reverse complement
        took_fwd = false;
        for (uint32_t j = 0; j < k; j++) {
            uint32_t a = t.ascii[i + j], b = comp_base(t.ascii[i + k - 1 - j]);
            if (a != b) { took_fwd = a < b; break; }
        }
    }
    key = took_fwd ? f : r;
    return true;
}

// ------------------------------------------------------------------ k-mers with bytes outside ACGTacgt (kmerize_string only)
// kmer.rs:271-299 kmerize_string has no has_no_n test: every window of a record is a k-mer, whatever its bytes.
// kmer.rs:847-863 switch_base on ANY byte: a<->t c<->g, u->a, n->n (case kept), everything else -> 'N'.
__device__ __forceinline__ uint32_t switch_base_any(uint32_t c) {
    switch (c) {
        case 'a': return 't'; case 'c': return 'g'; case 't': return 'a'; case 'g': return 'c'; case 'u': return 'a'; case 'n': return 'n';
        case 'A': return 'T'; case 'C': return 'G'; case 'T': return 'A'; case 'G': return 'C'; case 'U': return 'A'; case 'N': return 'N';
        default: return 'N';
    }
}
// The canonical upper-cased string of the window win[0..k): `fwd < revcomp ? fwd : revcomp` on raw bytes (kmer.rs:280),
// then to_uppercase (:283).  Returns true when the result is a plain ACGT string (then `key` holds its 2-bit codes).
__device__ __noinline__ static bool string_kmer(const uint8_t* win, uint32_t k, uint8_t* out, uint64_t& key) {
    bool took_fwd = false;
    for (uint32_t j = 0; j < k; j++) {
        const uint32_t a = win[j], b = switch_base_any(win[k - 1 - j]);
        if (a != b) { took_fwd = a < b; break; }
    }
    bool clean = true;
    key = 0;
    for (uint32_t j = 0; j < k; j++) {
        uint32_t c = took_fwd ? (uint32_t)win[j] : switch_base_any(win[k - 1 - j]);
        if (c >= 'a' && c <= 'z') c -= 32;
        out[j] = (uint8_t)c;
        clean = clean && (c == 'A' || c == 'C' || c == 'G' || c == 'T');
        key = (key << 2) | base_code(c);
    }
    return clean;
}
// The words XXH3 reads (see HashIn) straight from k bytes.
__device__ __forceinline__ uint64_t le_bytes(const uint8_t* p, uint32_t n) {
    uint64_t v = 0;
    for (uint32_t i = 0; i < n; i++) v |= (uint64_t)p[i] << (8 * i);
    return v;
}
__device__ __forceinline__ HashIn hashin_from_bytes(const uint8_t* s, uint32_t k) {
    HashIn in;
    in.w0 = in.w1 = in.w2 = in.w3 = 0;
    if (k >= 17) { in.w0 = le_bytes(s, 8); in.w1 = le_bytes(s + 8, 8); in.w2 = le_bytes(s + k - 16, 8); in.w3 = le_bytes(s + k - 8, 8); }
    else if (k >= 9) { in.w0 = le_bytes(s, 8); in.w1 = le_bytes(s + k - 8, 8); }
    else if (k >= 4) { in.w0 = le_bytes(s, 4); in.w1 = le_bytes(s + k - 4, 4); }
    else { in.w0 = s[0]; in.w1 = s[k >> 1]; in.w2 = s[k - 1]; }
    return in;
}

// ------------------------------------------------------------------ minimizers (.mxi indexes)
// kmer.rs:971-986 find_minimizer(seq, m) on the canonical k-mer `s` (packed, upper case) and its reverse
// complement `src`: the smallest of seq[0..m] and, for i in 1..=k-m, seq[i..i+m] and revcomp(seq[i..i+m])
// -- the reverse complement of the FIRST m-mer is never a candidate.  For upper-case (or all lower-case)
// bases byte order == numeric order of the 2-bit codes.  `which` = 2*i + (1 if the winner is a reverse
// complement).  Ties keep the earlier candidate like the reference's strict `<`; tied strings are equal anyway.
__device__ __forceinline__ uint64_t minimizer_packed(uint64_t s, uint64_t src, uint32_t k, uint32_t m, uint32_t& which) {
    const uint64_t mm = (1ULL << (2 * m)) - 1;          // m <= 31
    uint64_t best = (s >> (2 * (k - m))) & mm;
    uint32_t w = 0;
    for (uint32_t i = 1; i + m <= k; i++) {
        const uint64_t f = (s >> (2 * (k - m - i))) & mm;   // seq[i..i+m]
        const uint64_t r = (src >> (2 * i)) & mm;           // revcomp(seq)[k-(i+m)..k-i]
        if (f < best) { best = f; w = 2 * i; }
        if (r < best) { best = r; w = 2 * i + 1; }
    }
    which = w;
    return best;
}
// The same on the raw bytes of tile window i, for windows that mix upper and lower case (kmer.rs:328-394 choose
// the minimizer on the raw-case canonical k-mer and upper-case it afterwards).  Returns the upper-cased codes.
static __device__ __noinline__ uint64_t minimizer_exact(const Tile& t, int i, uint32_t k, uint32_t m, bool took_fwd, uint32_t& which) {
    // byte j of the canonical k-mer; byte x of candidate (a, rc): rc ? comp(sb(a+m-1-x)) : sb(a+x)
    auto sb = [&](uint32_t j) -> uint32_t { return took_fwd ? (uint32_t)t.ascii[i + j] : comp_base(t.ascii[i + k - 1 - j]); };
    auto cb = [&](uint32_t a, bool rc, uint32_t x) -> uint32_t { return rc ? comp_base(sb(a + m - 1 - x)) : sb(a + x); };
    uint32_t ba = 0; bool brc = false;
    for (uint32_t a = 1; a + m <= k; a++) {
        for (int rc = 0; rc < 2; rc++) {
            for (uint32_t x = 0; x < m; x++) {
                const uint32_t c = cb(a, rc != 0, x), b = cb(ba, brc, x);
                if (c != b) { if (c < b) { ba = a; brc = rc != 0; } break; }
            }
        }
    }
    uint64_t key = 0;
    for (uint32_t x = 0; x < m; x++) key = (key << 2) | base_code(cb(ba, brc, x));
    which = 2 * ba + (brc ? 1u : 0u);
    return key;
}
// Minimizer of the k-mer that tile_kmer() returned for window i (`key`, strand `took_fwd`), plus the forward-read
// window [mpos, mpos+m) that spells it (mfwd) or its reverse complement (!mfwd).
__device__ __forceinline__ uint64_t tile_minimizer(const Tile& t, int i, uint32_t k, uint32_t m, uint64_t key, bool took_fwd,
                                                   bool has_lower, uint32_t& mpos, bool& mfwd) {
    uint32_t which;
    uint64_t mk;
    const uint32_t low = has_lower ? mask_window(t.lower, i, k) : 0u;
    if (low != 0u && low != (k == 32 ? 0xFFFFFFFFu : ((1u << k) - 1))) mk = minimizer_exact(t, i, k, m, took_fwd, which);
    else mk = minimizer_packed(key, revcomp_key(key, k), k, m, which);
    const uint32_t j = which >> 1;
    mpos = (uint32_t)i + (took_fwd ? j : k - m - j);
    mfwd = ((which & 1u) == 0u) == took_fwd;
    return mk;
}

// ------------------------------------------------------------------ count table (open addressing)
// `pad` is the case mask of a raw-case k-mer (see hashin_apply_case) in tables filled by table_insert_cs, CID_CS_UNSET
// (what table_clear writes) everywhere else: slot_cs() is what the hashing kernels apply.
struct __align__(16) Slot { unsigned long long key; uint32_t count; uint32_t pad; };
#define CID_EMPTY_KEY 0xFFFFFFFFFFFFFFFFULL
#define CID_CS_UNSET 0xFFFFFFFFu
__device__ __forceinline__ uint32_t slot_cs(uint32_t pad) { return pad == CID_CS_UNSET ? 0u : pad; }

__device__ __forceinline__ uint64_t mix64(uint64_t x) {   // murmur3 finalizer: slot choice only
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
// Returns 1 if the key was new, 0 if it was already present, -1 if the region is (nearly) full: the probe
// sequence is cut after CID_MAX_PROBE slots so an optimistically sized table can never hang a kernel.
#define CID_MAX_PROBE 2048
__device__ __forceinline__ int table_insert(Slot* region, uint64_t mask, uint64_t key) {
    uint64_t h = mix64(key) & mask;
    for (int probes = 0; probes < CID_MAX_PROBE; probes++) {
        unsigned long long prev = atomicCAS(&region[h].key, CID_EMPTY_KEY, (unsigned long long)key);
        if (prev == CID_EMPTY_KEY || prev == key) { atomicAdd(&region[h].count, 1u); return prev == CID_EMPTY_KEY ? 1 : 0; }
        h = (h + 1) & mask;
    }
    return -1;
}

// Case-aware variant: the slot of (key, cs).  Two k-mers that differ only in case share their codes, so a slot whose codes
// match is ours only if its case mask does too; whoever reaches a freshly claimed slot first settles its mask (the claim of
// the codes and of the mask are two atomics, both idempotent for equal k-mers).  Returns 1 for the call that created the entry.
__device__ __forceinline__ int table_insert_cs(Slot* region, uint64_t mask, uint64_t key, uint32_t cs) {
    uint64_t h = mix64(key) & mask;
    for (int probes = 0; probes < CID_MAX_PROBE; probes++) {
        const unsigned long long prev = atomicCAS(&region[h].key, CID_EMPTY_KEY, (unsigned long long)key);
        if (prev == CID_EMPTY_KEY || prev == key) {
            const uint32_t c = atomicCAS(&region[h].pad, CID_CS_UNSET, cs);
            if (c == CID_CS_UNSET || c == cs) { atomicAdd(&region[h].count, 1u); return c == CID_CS_UNSET ? 1 : 0; }
        }
        h = (h + 1) & mask;
    }
    return -1;
}

// Packed variant for keys of <= 21 bases (42 bits): one 8-byte word holds key << 22 | count, so a read-set build's
// count table is half as large (more of it stays in L2) and an insert is one load + one atomic.  All ones = empty
// (never a canonical key: T..T's reverse complement A..A is smaller).  A count within 2^20 of the field's limit raises
// `overflow` instead of adding (more threads than that cannot be between their load and their add).
#define CID_PK_BITS 22
#define CID_PK_CMASK ((1ULL << CID_PK_BITS) - 1)
__device__ __forceinline__ int packed_insert(unsigned long long* tab, uint64_t mask, uint64_t key, bool& overflow) {
    uint64_t h = mix64(key) & mask;
    for (int probes = 0; probes < CID_MAX_PROBE; probes++) {
        unsigned long long cur = *(volatile unsigned long long*)&tab[h];
        if (cur == CID_EMPTY_KEY) {
            cur = atomicCAS(&tab[h], CID_EMPTY_KEY, (unsigned long long)((key << CID_PK_BITS) | 1ULL));
            if (cur == CID_EMPTY_KEY) return 1;
        }
        if ((cur >> CID_PK_BITS) == key) {
            if ((cur & CID_PK_CMASK) >= CID_PK_CMASK - (1ULL << 20)) overflow = true;
            else atomicAdd(&tab[h], 1ULL);
            return 0;
        }
        h = (h + 1) & mask;
    }
    return -1;
}
struct SlotView { uint64_t key; uint32_t count; uint32_t cs; bool used; };
template <bool PACKED> __device__ __forceinline__ SlotView slot_read(const void* table, uint64_t i) {
    SlotView v;
    if (PACKED) {
        const unsigned long long w = ((const unsigned long long*)table)[i];
        v.used = w != CID_EMPTY_KEY; v.key = w >> CID_PK_BITS; v.count = (uint32_t)(w & CID_PK_CMASK); v.cs = 0;
    } else {
        const Slot s = ((const Slot*)table)[i];
        v.used = s.key != CID_EMPTY_KEY; v.key = s.key; v.count = s.count; v.cs = slot_cs(s.pad);
    }
    return v;
}

// Key-only variant (8-byte slots) for builds that keep every k-mer (no count filter): the table of a 5-Mbp genome
// then fits the L2 cache.  Same return values.
__device__ __forceinline__ int set_insert(unsigned long long* keys, uint64_t mask, uint64_t key) {
    uint64_t h = mix64(key) & mask;
    for (int probes = 0; probes < CID_MAX_PROBE; probes++) {
        const unsigned long long prev = atomicCAS(&keys[h], CID_EMPTY_KEY, (unsigned long long)key);
        if (prev == CID_EMPTY_KEY) return 1;
        if (prev == key) return 0;
        h = (h + 1) & mask;
    }
    return -1;
}

}  // namespace cid
