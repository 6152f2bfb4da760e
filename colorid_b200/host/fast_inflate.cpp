// Streaming gzip decoder of the host layer (see fast_inflate.hpp).
#include "fast_inflate.hpp"

#include <immintrin.h>
#include <zlib.h>      // crc32(): the reference value of the tests and the fallback without PCLMULQDQ

#include <cstdlib>
#include <cstring>

#include "cid_host.hpp"

namespace cidh {

namespace {
// table entry (fast_inflate.hpp): value << 16 | flags << 12 | extra bits << 8 | code bits to consume
// F_LIT2: two literals in one first-level entry (value = first | second << 8, code bits = both codes): decoding literals is a
// chain of dependent table lookups (load latency per symbol), and FASTQ text is almost all literals with 2..6-bit codes
enum : uint32_t { F_LIT = GzInflater::F_LIT, F_BASE = GzInflater::F_BASE, F_EOB = GzInflater::F_EOB, F_SUB = GzInflater::F_SUB, F_LIT2 = GzInflater::F_LIT2 };
const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
inline uint32_t rev_bits(uint32_t code, unsigned len) {
    uint32_t r = 0;
    for (unsigned i = 0; i < len; i++) { r = (r << 1) | (code & 1u); code >>= 1; }
    return r;
}

// CRC-32 (gzip polynomial, reflected) of len bytes, len a multiple of 16 and >= 64, by carry-less multiplication:
// four 128-bit lanes folded by 512 bits per step (Gopal et al., "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ").
__attribute__((target("pclmul,sse4.1")))
static uint32_t crc32_clmul(const unsigned char* buf, size_t len, uint32_t crc) {
    alignas(16) static const uint64_t k1k2[2] = {0x0154442bd4ull, 0x01c6e41596ull};
    alignas(16) static const uint64_t k3k4[2] = {0x01751997d0ull, 0x00ccaa009eull};
    alignas(16) static const uint64_t k5k0[2] = {0x0163cd6124ull, 0x0000000000ull};
    alignas(16) static const uint64_t poly[2] = {0x01db710641ull, 0x01f7011641ull};
    __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
    x1 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
    x2 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
    x3 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
    x4 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
    x0 = _mm_load_si128((const __m128i*)k1k2);
    buf += 64; len -= 64;
    while (len >= 64) {
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
        x7 = _mm_clmulepi64_si128(x3, x0, 0x00); x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
        x3 = _mm_clmulepi64_si128(x3, x0, 0x11); x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
        y5 = _mm_loadu_si128((const __m128i*)(buf + 0x00)); y6 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
        y7 = _mm_loadu_si128((const __m128i*)(buf + 0x20)); y8 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5); x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
        x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7); x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
        buf += 64; len -= 64;
    }
    x0 = _mm_load_si128((const __m128i*)k3k4);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
    while (len >= 16) {
        x2 = _mm_loadu_si128((const __m128i*)buf);
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
        buf += 16; len -= 16;
    }
    x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
    x3 = _mm_setr_epi32(~0, 0, ~0, 0);
    x1 = _mm_srli_si128(x1, 8);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_loadl_epi64((const __m128i*)k5k0);
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, x3);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_load_si128((const __m128i*)poly);
    x2 = _mm_and_si128(x1, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
    x2 = _mm_and_si128(x2, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}
}  // namespace
// zlib-compatible: running crc in, running crc out
uint32_t GzInflater::crc32_fast(uint32_t crc, const unsigned char* p, size_t n) {
    if (n >= 64 && __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1")) {
        const size_t m = n & ~(size_t)15;
        crc = ~crc32_clmul(p, m, ~crc);
        p += m; n -= m;
    }
    return n ? (uint32_t)crc32(crc, p, (uInt)n) : crc;
}

GzInflater::GzInflater(const uint8_t* data, size_t n, const std::string& what)
    : base_(data), in_(data), end_(data + n), what_(what), obuf_((size_t)HIST + CHUNK + SLACK) {}

void GzInflater::fail(const char* why) const { throw Error("gzip stream of " + what_ + ": " + why); }

inline void GzInflater::refill() {
    if (in_ + 8 <= end_) {
        uint64_t w;
        memcpy(&w, in_, 8);
        bitbuf_ |= w << bitcnt_;
        in_ += (63 - bitcnt_) >> 3;
        bitcnt_ |= 56;
        return;
    }
    while (bitcnt_ <= 56) {
        uint64_t b = 0;
        if (in_ < end_) b = *in_++; else fed_zero_bytes_++;
        bitbuf_ |= b << bitcnt_;
        bitcnt_ += 8;
    }
}
inline uint32_t GzInflater::bits(unsigned n) {           // n <= 32, after a refill
    const uint32_t v = (uint32_t)(bitbuf_ & ((1ull << n) - 1));
    bitbuf_ >>= n;
    bitcnt_ -= n;
    return v;
}
// drop the bits up to the next byte boundary and hand the whole bytes still buffered back to the input
void GzInflater::byte_align() {
    const unsigned drop = bitcnt_ & 7;
    bitbuf_ >>= drop;
    bitcnt_ -= drop;
    const uint64_t nbytes = bitcnt_ >> 3;
    if (fed_zero_bytes_ > nbytes) fail("unexpected end of data");
    in_ -= (nbytes - fed_zero_bytes_);
    fed_zero_bytes_ = 0;
    bitbuf_ = 0;
    bitcnt_ = 0;
}

size_t GzInflater::header_size(const uint8_t* p, const uint8_t* end, const char** why) {
    // RFC 1952: ID1 ID2 CM FLG MTIME(4) XFL OS [FEXTRA] [FNAME] [FCOMMENT] [FHCRC]
    const uint8_t* in = p;
    auto bad = [&](const char* w) -> size_t { *why = w; return 0; };
    if (end - in < 10) return bad("truncated header");
    if (in[0] != 0x1f || in[1] != 0x8b) return bad("not a gzip member");
    if (in[2] != 8) return bad("unknown compression method");
    const unsigned flg = in[3];
    if (flg & 0xE0) return bad("reserved header flags set");
    in += 10;
    if (flg & 4) {
        if (end - in < 2) return bad("truncated header");
        const size_t xlen = in[0] | ((size_t)in[1] << 8);
        in += 2;
        if ((size_t)(end - in) < xlen) return bad("truncated header");
        in += xlen;
    }
    for (unsigned f = 8; f <= 16; f <<= 1)      // FNAME, FCOMMENT: zero-terminated
        if (flg & f) {
            const uint8_t* z = (const uint8_t*)memchr(in, 0, (size_t)(end - in));
            if (!z) return bad("truncated header");
            in = z + 1;
        }
    if (flg & 2) {
        if (end - in < 2) return bad("truncated header");
        in += 2;
    }
    return (size_t)(in - p);
}

void GzInflater::parse_header() {
    const char* why = nullptr;
    const size_t h = header_size(in_, end_, &why);
    if (!h) fail(why);
    in_ += h;
    crc_ = (uint32_t)crc32(0L, Z_NULL, 0);
    crc_from_ = out_;
    member_start_ = (int64_t)out_;
    last_block_ = false;
    st_ = ST_BLOCK;
}

void GzInflater::resume(uint64_t bitpos, bool at_header, const uint8_t* win, size_t nwin, uint32_t crc, uint64_t member_len) {
    in_ = base_ + (bitpos >> 3);
    bitbuf_ = 0; bitcnt_ = 0; fed_zero_bytes_ = 0;
    out_ = taken_ = crc_from_ = HIST;
    if (at_header) { st_ = ST_MEMBER; return; }
    refill();
    bits((unsigned)(bitpos & 7));
    if (nwin > (size_t)HIST) { win += nwin - HIST; nwin = HIST; }
    memcpy(obuf_.data() + HIST - nwin, win, nwin);
    crc_ = crc;
    member_start_ = (int64_t)HIST - (int64_t)member_len;
    last_block_ = false;
    st_ = ST_BLOCK;
}

void GzInflater::fold_crc() {
    if (out_ > crc_from_) crc_ = crc32_fast(crc_, obuf_.data() + crc_from_, out_ - crc_from_);     // (zlib's crc32 was 30 % of the decoder's time)
    crc_from_ = out_;
}

void GzInflater::parse_trailer() {
    byte_align();
    if (end_ - in_ < 8) fail("truncated trailer");
    fold_crc();
    uint32_t want_crc, want_len;
    memcpy(&want_crc, in_, 4);
    memcpy(&want_len, in_ + 4, 4);
    in_ += 8;
    if (want_crc != crc_) fail("CRC mismatch");
    if (want_len != (uint32_t)((int64_t)out_ - member_start_)) fail("length mismatch");
    // another member (cat a.gz b.gz, bgzip)?  anything else after the trailer is ignored, as gzread does
    st_ = (end_ - in_ >= 2 && in_[0] == 0x1f && in_[1] == 0x8b) ? ST_MEMBER : ST_DONE;
}

void GzInflater::build_table(const uint8_t* lens, unsigned n, unsigned tbits, bool dist, std::vector<uint32_t>& table, bool& ok) {
    unsigned count[16] = {0};
    for (unsigned i = 0; i < n; i++) count[lens[i]]++;
    count[0] = 0;
    int left = 1;
    for (unsigned l = 1; l <= 15; l++) {
        left <<= 1;
        left -= (int)count[l];
        if (left < 0) { ok = false; return; }          // over-subscribed
    }
    uint32_t next[16];
    uint32_t code = 0;
    for (unsigned l = 1; l <= 15; l++) { code = (code + count[l - 1]) << 1; next[l] = code; }
    table.assign((size_t)1 << tbits, 0u);
    auto entry = [&](unsigned sym, unsigned consume) -> uint32_t {
        if (!dist) {
            if (sym < 256) return ((uint32_t)sym << 16) | F_LIT | consume;
            if (sym == 256) return F_EOB | consume;
            if (sym > 285) return 0u;
            return ((uint32_t)kLenBase[sym - 257] << 16) | F_BASE | ((uint32_t)kLenExtra[sym - 257] << 8) | consume;
        }
        if (sym > 29) return 0u;
        return ((uint32_t)kDistBase[sym] << 16) | F_BASE | ((uint32_t)kDistExtra[sym] << 8) | consume;
    };
    // codes, bit-reversed (DEFLATE packs Huffman codes starting with their most significant bit)
    uint32_t rcode[320];
    static thread_local uint8_t sublen[1 << LBITS];
    static thread_local uint32_t suboff[1 << LBITS];
    memset(sublen, 0, (size_t)1 << tbits);
    for (unsigned s = 0; s < n; s++) {
        const unsigned l = lens[s];
        if (!l) continue;
        rcode[s] = rev_bits(next[l]++, l);
        if (l > tbits) {
            uint8_t& sl = sublen[rcode[s] & ((1u << tbits) - 1)];
            if (l - tbits > sl) sl = (uint8_t)(l - tbits);
        }
    }
    size_t off = (size_t)1 << tbits;
    for (size_t p = 0; p < ((size_t)1 << tbits); p++)
        if (sublen[p]) {
            suboff[p] = (uint32_t)off;
            table[p] = ((uint32_t)off << 16) | F_SUB | ((uint32_t)sublen[p] << 8) | tbits;
            off += (size_t)1 << sublen[p];
        }
    if (off > 65535) { ok = false; return; }
    table.resize(off, 0u);
    for (unsigned s = 0; s < n; s++) {
        const unsigned l = lens[s];
        if (!l) continue;
        if (l <= tbits) {
            const uint32_t e = entry(s, l);
            for (size_t i = rcode[s]; i < ((size_t)1 << tbits); i += (size_t)1 << l) table[i] = e;
        } else {
            const size_t p = rcode[s] & ((1u << tbits) - 1);
            const uint32_t e = entry(s, l - tbits);
            for (size_t i = rcode[s] >> tbits; i < ((size_t)1 << sublen[p]); i += (size_t)1 << (l - tbits)) table[suboff[p] + i] = e;
        }
    }
    if (!dist && tbits == LBITS) {
        // pair up literals: index i starts with a literal of l1 bits; if the bits after it (zero-extended) select a literal of at
        // most tbits - l1 bits, both are decided by the tbits in hand (descending i: entries i >> l1 are still single)
        for (size_t i = ((size_t)1 << tbits); i-- > 0;) {
            const uint32_t e1 = table[i];
            if (!(e1 & F_LIT)) continue;
            const unsigned l1 = e1 & 0x7F;
            if (l1 >= tbits) continue;
            const uint32_t e2 = table[i >> l1];
            if (!(e2 & F_LIT) || (e2 & F_LIT2)) continue;
            const unsigned l2 = e2 & 0x7F;
            if (l1 + l2 > tbits) continue;
            table[i] = ((e1 >> 16) << 16) | ((e2 >> 16) << 24) | F_LIT | F_LIT2 | (l1 + l2);
        }
    }
    ok = true;
}

void GzInflater::fixed_lengths(uint8_t lens[288], uint8_t dl[32]) {
    for (unsigned i = 0; i < 144; i++) lens[i] = 8;
    for (unsigned i = 144; i < 256; i++) lens[i] = 9;
    for (unsigned i = 256; i < 280; i++) lens[i] = 7;
    for (unsigned i = 280; i < 288; i++) lens[i] = 8;
    for (unsigned i = 0; i < 32; i++) dl[i] = 5;
}
void GzInflater::build_fixed() {
    uint8_t lens[288], dl[32];
    fixed_lengths(lens, dl);
    bool ok = false;
    build_table(lens, 288, LBITS, false, lt_, ok);
    build_table(dl, 32, DBITS, true, dt_, ok);
}

void GzInflater::build_dynamic() {
    refill();
    const unsigned hlit = bits(5) + 257, hdist = bits(5) + 1, hclen = bits(4) + 4;
    if (hlit > 286 || hdist > 30) fail("bad code counts");
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t cl[19] = {0};
    for (unsigned i = 0; i < hclen; i++) { refill(); cl[order[i]] = (uint8_t)bits(3); }
    std::vector<uint32_t> ct;
    bool ok = false;
    build_table(cl, 19, 7, false, ct, ok);       // (code-length alphabet: symbols 0..18 come back as "literals")
    if (!ok) fail("bad code-length code");
    uint8_t lens[320] = {0};
    unsigned i = 0;
    while (i < hlit + hdist) {
        refill();
        const uint32_t e = ct[bitbuf_ & 127];
        if (!(e & F_LIT)) fail("bad code-length symbol");
        bits(e & 0xFF);
        const unsigned sym = e >> 16;
        if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
        unsigned rep, val = 0;
        if (sym == 16) { if (!i) fail("repeat without a previous length"); val = lens[i - 1]; rep = 3 + bits(2); }
        else if (sym == 17) rep = 3 + bits(3);
        else rep = 11 + bits(7);
        if (i + rep > hlit + hdist) fail("code lengths overrun");
        while (rep--) lens[i++] = (uint8_t)val;
    }
    if (!lens[256]) fail("no end-of-block code");
    build_table(lens, hlit, LBITS, false, lt_, ok);
    if (!ok) fail("bad literal/length code");
    build_table(lens + hlit, hdist, DBITS, true, dt_, ok);
    if (!ok) fail("bad distance code");
}

void GzInflater::block_header() {
    refill();
    if (fed_zero_bytes_ * 8 > bitcnt_) fail("unexpected end of data");
    last_block_ = bits(1) != 0;
    const unsigned type = bits(2);
    if (type == 0) {
        byte_align();
        if (end_ - in_ < 4) fail("truncated stored block");
        const unsigned len = in_[0] | ((unsigned)in_[1] << 8), nlen = in_[2] | ((unsigned)in_[3] << 8);
        if ((len ^ 0xFFFFu) != nlen) fail("stored block length check failed");
        in_ += 4;
        stored_left_ = len;
        st_ = ST_STORED;
    } else if (type == 1) { build_fixed(); st_ = ST_HUFF; }
    else if (type == 2) { build_dynamic(); st_ = ST_HUFF; }
    else fail("reserved block type");
}

// Decodes symbols of the current Huffman block until the output reaches `limit` or the block ends.  The bit reader lives in
// locals here: the byte stores into the output may alias the object's members as far as the compiler knows, which would
// force a reload of the bit buffer after every literal.
void GzInflater::run_huff(size_t limit) {
    uint8_t* const ob = obuf_.data();
    const uint32_t* const lt = lt_.data();
    const uint32_t* const dt = dt_.data();
    size_t out = out_;
    uint64_t bb = bitbuf_;
    unsigned bc = bitcnt_;
    const uint8_t* in = in_;
    const uint8_t* const end = end_;
    uint64_t fedz = fed_zero_bytes_;
    const int64_t mstart = member_start_;
    const char* err = nullptr;
#define CID_REFILL()                                                                    \
    do {                                                                                \
        if (in + 8 <= end) {                                                            \
            uint64_t w_;                                                                \
            memcpy(&w_, in, 8);                                                         \
            bb |= w_ << bc;                                                             \
            in += (63 - bc) >> 3;                                                       \
            bc |= 56;                                                                   \
        } else {                                                                        \
            while (bc <= 56) {                                                          \
                uint64_t b_ = 0;                                                        \
                if (in < end) b_ = *in++; else fedz++;                                  \
                bb |= b_ << bc;                                                         \
                bc += 8;                                                                \
            }                                                                           \
        }                                                                               \
    } while (0)
    while (out < limit) {
        CID_REFILL();                                       // >= 56 bits
        uint32_t e = lt[bb & ((1u << LBITS) - 1)];
        if (e & F_SUB) {
            const unsigned sb = (e >> 8) & 15;
            e = lt[(e >> 16) + ((bb >> LBITS) & ((1u << sb) - 1))];
            bb >>= LBITS; bc -= LBITS;
        }
        if (e & F_LIT) {
            // literals, one or two per entry; more of them out of the bits in hand (first-level entries only: at most 11
            // bits each, at least 41 bits are left after the first)
            ob[out] = (uint8_t)(e >> 16); ob[out + 1] = (uint8_t)(e >> 24);
            out += 1 + ((e >> 7) & 1u);
            bb >>= (e & 0x7F); bc -= (e & 0x7F);
            e = lt[bb & ((1u << LBITS) - 1)];
            if (e & F_LIT) {
                ob[out] = (uint8_t)(e >> 16); ob[out + 1] = (uint8_t)(e >> 24);
                out += 1 + ((e >> 7) & 1u);
                bb >>= (e & 0x7F); bc -= (e & 0x7F);
                e = lt[bb & ((1u << LBITS) - 1)];
                if (e & F_LIT) {
                    ob[out] = (uint8_t)(e >> 16); ob[out + 1] = (uint8_t)(e >> 24);
                    out += 1 + ((e >> 7) & 1u);
                    bb >>= (e & 0x7F); bc -= (e & 0x7F);
                    e = lt[bb & ((1u << LBITS) - 1)];
                    if ((e & (F_LIT | F_SUB)) == F_LIT) {
                        ob[out] = (uint8_t)(e >> 16); ob[out + 1] = (uint8_t)(e >> 24);
                        out += 1 + ((e >> 7) & 1u);
                        bb >>= (e & 0x7F); bc -= (e & 0x7F);
                    }
                }
            }
            continue;
        }
        if (e & F_EOB) {
            bb >>= (e & 0xFF); bc -= (e & 0xFF);
            st_ = last_block_ ? ST_TRAILER : ST_BLOCK;
            break;
        }
        if (!(e & F_BASE)) { err = "invalid literal/length code"; break; }
        bb >>= (e & 0xFF); bc -= (e & 0xFF);
        const unsigned lx = (e >> 8) & 15;
        const unsigned len = (e >> 16) + (unsigned)(bb & ((1u << lx) - 1));
        bb >>= lx; bc -= lx;
        // (>= 36 bits are left here: 56 after the refill at the top, at most 15 + 5 taken by the length; a distance takes <= 15 + 13)
        uint32_t d = dt[bb & ((1u << DBITS) - 1)];
        if (d & F_SUB) {
            const unsigned sb = (d >> 8) & 15;
            d = dt[(d >> 16) + ((bb >> DBITS) & ((1u << sb) - 1))];
            bb >>= DBITS; bc -= DBITS;
        }
        if (!(d & F_BASE)) { err = "invalid distance code"; break; }
        bb >>= (d & 0xFF); bc -= (d & 0xFF);
        const unsigned dx = (d >> 8) & 15;
        const size_t dist = (d >> 16) + (size_t)(bb & ((1u << dx) - 1));
        bb >>= dx; bc -= dx;
        if ((int64_t)out - (int64_t)dist < mstart) { err = "distance reaches before the start of the data"; break; }
        const uint8_t* src = ob + out - dist;
        uint8_t* dst = ob + out;
        if (dist >= 8) {
            for (unsigned i = 0; i < len; i += 8) memcpy(dst + i, src + i, 8);      // (may write up to 7 bytes past len: SLACK)
        } else if (dist == 1) {
            memset(dst, src[0], len);                       // runs (a quality line of one character): not a byte-by-byte chain
        } else {
            // period 2..7: spell the first D = dist * ceil(8 / dist) >= 8 bytes one by one, then copy eight at a time from D back
            const unsigned D = dist * ((7 + (unsigned)dist) / (unsigned)dist);
            unsigned i = 0;
            for (; i < D && i < len; i++) dst[i] = src[i];
            for (; i < len; i += 8) memcpy(dst + i, dst + i - D, 8);
        }
        out += len;
    }
#undef CID_REFILL
    out_ = out; bitbuf_ = bb; bitcnt_ = bc; in_ = in; fed_zero_bytes_ = fedz;
    if (err) fail(err);
    if (fed_zero_bytes_ * 8 > bitcnt_) fail("unexpected end of data");
}

// Produces the next chunk of output: obuf_[HIST, out_).
void GzInflater::decode_chunk() {
    // slide: the last HIST bytes become the history of the new chunk
    if (out_ > (size_t)HIST) {
        fold_crc();
        const size_t shift = out_ - HIST;
        memmove(obuf_.data(), obuf_.data() + shift, HIST);
        member_start_ -= (int64_t)shift;
        out_ = taken_ = crc_from_ = HIST;
    }
    const size_t limit = (size_t)HIST + CHUNK - 272;
    while (out_ < limit && st_ != ST_DONE) {
        switch (st_) {
            case ST_MEMBER: parse_header(); break;
            case ST_BLOCK: block_header(); break;
            case ST_STORED: {
                const size_t n = std::min<size_t>(stored_left_, limit - out_);
                if ((size_t)(end_ - in_) < n) fail("truncated stored block");
                memcpy(obuf_.data() + out_, in_, n);
                in_ += n; out_ += n; stored_left_ -= (uint32_t)n;
                if (!stored_left_) st_ = last_block_ ? ST_TRAILER : ST_BLOCK;
                break;
            }
            case ST_HUFF: run_huff(limit); break;
            case ST_TRAILER: parse_trailer(); break;
            case ST_DONE: break;
        }
    }
}

size_t GzInflater::read(uint8_t* dst, size_t cap) {
    size_t got = 0;
    while (got < cap) {
        if (taken_ == out_) {
            if (st_ == ST_DONE) break;
            decode_chunk();
            if (taken_ == out_ && st_ == ST_DONE) break;
            continue;
        }
        const size_t n = std::min(cap - got, out_ - taken_);
        memcpy(dst + got, obuf_.data() + taken_, n);
        taken_ += n; got += n;
    }
    return got;
}

}  // namespace cidh
