// cid_pack_reads: host-side packing of reads for the read_id entry points that take packed input.
//
// The reference masks low-quality bases on the host (seq.rs:36-56 qual_mask, called from the FASTQ loops of
// read_id_mt_pe.rs:733-760) and hands ASCII strings to its classifier.  Shipping ASCII bases + qualities to a GPU costs
// 2 bytes per base over PCIe and makes every kernel mask and pack them again; this packer does the masking like the
// reference does -- on the host, while the reads are being parsed -- and emits, per read (all mates concatenated), exactly
// the planes the kernels' shared-memory tile holds:
//     codes[ceil(L/16)]  2 bits per base (A0 C1 G2 T3), 16 bases per word, first base in bits 31:30
//     bad  [ceil(L/32)]  bit j (LSB first) = base j is not one of ACGTacgt, or its quality is below the offset; bits >= L set
//     lower[ceil(L/32)]  bit j = base j is a lower-case acgt            (only when some read of the batch has one)
// 3 (or 4) bits per base instead of 16; lossless for read_id, which only asks of a non-base that it is one (kmer.rs:221-243,
// seq.rs:59-70).  AVX2 path: 32 bases per step; scalar table path elsewhere.  Host threads over read ranges.
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "cid_internal.h"

namespace {

struct Nib { uint8_t v[256]; Nib() { for (int c = 0; c < 256; c++) { const int u = c & 0xDF; const bool good = u == 'A' || u == 'C' || u == 'G' || u == 'T';
                                                                    v[c] = good ? (uint8_t)((((c >> 1) ^ (c >> 2)) & 3) | ((c & 0x20) ? 4 : 0)) : 8; } } };
const Nib kNib;

// bases [i0, L) of one read, scalar
inline void pack_tail(const uint8_t* c, const uint8_t* q, uint32_t maxq, uint32_t i0, uint32_t L, uint32_t* codes, uint32_t* bad, uint32_t* lower, bool& any_lower) {
    const uint32_t end = (L + 31) & ~31u;
    for (uint32_t i = i0; i < end; i++) {
        uint32_t n = 8;
        if (i < L) { n = kNib.v[c[i]]; if (q && q[i] < maxq) n = 8; }
        if ((i & 15) == 0 && (i >> 4) < ((L + 15) >> 4)) codes[i >> 4] = 0;
        if ((i & 31) == 0) { bad[i >> 5] = 0; if (lower) lower[i >> 5] = 0; }
        if (n & 8) bad[i >> 5] |= 1u << (i & 31);
        else {
            codes[i >> 4] |= (n & 3u) << (30 - 2 * (i & 15));
            if (n & 4) { any_lower = true; if (lower) lower[i >> 5] |= 1u << (i & 31); }
        }
    }
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) inline uint32_t squeeze8(uint64_t x) {      // 8 bytes of 2-bit codes -> 16 bits, first base on top
    uint64_t y = __builtin_bswap64(x);
    y = (y | (y >> 6)) & 0x000F000F000F000FULL;
    y = (y | (y >> 12)) & 0x000000FF000000FFULL;
    y = (y | (y >> 24)) & 0xFFFFULL;
    return (uint32_t)y;
}
__attribute__((target("avx2"))) uint32_t pack_blocks_avx2(const uint8_t* c, const uint8_t* q, uint32_t maxq, uint32_t L, uint32_t* codes, uint32_t* bad,
                                                          uint32_t* lower, bool& any_lower) {
    const __m256i vDF = _mm256_set1_epi8((char)0xDF), v20 = _mm256_set1_epi8(0x20), v03 = _mm256_set1_epi8(3), vN = _mm256_set1_epi8('N');
    const __m256i vA = _mm256_set1_epi8('A'), vC = _mm256_set1_epi8('C'), vG = _mm256_set1_epi8('G'), vT = _mm256_set1_epi8('T');
    const __m256i vq = _mm256_set1_epi8((char)(maxq > 255 ? 255 : maxq));
    uint32_t i = 0, lowacc = 0;
    for (; i + 32 <= L; i += 32) {
        __m256i b = _mm256_loadu_si256((const __m256i*)(c + i));
        if (q) {
            const __m256i qv = _mm256_loadu_si256((const __m256i*)(q + i));
            const __m256i ge = _mm256_cmpeq_epi8(_mm256_max_epu8(qv, vq), qv);      // quality >= offset
            b = _mm256_blendv_epi8(vN, b, ge);
        }
        const __m256i u = _mm256_and_si256(b, vDF);
        const __m256i good = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, vA), _mm256_cmpeq_epi8(u, vC)),
                                             _mm256_or_si256(_mm256_cmpeq_epi8(u, vG), _mm256_cmpeq_epi8(u, vT)));
        const uint32_t gm = (uint32_t)_mm256_movemask_epi8(good);
        const uint32_t lm = (uint32_t)_mm256_movemask_epi8(_mm256_and_si256(good, _mm256_cmpeq_epi8(_mm256_and_si256(b, v20), v20)));
        const __m256i t = _mm256_and_si256(_mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(b, 1), _mm256_srli_epi16(b, 2)), v03), good);
        alignas(32) uint64_t x[4];
        _mm256_store_si256((__m256i*)x, t);
        codes[(i >> 4)] = (squeeze8(x[0]) << 16) | squeeze8(x[1]);
        codes[(i >> 4) + 1] = (squeeze8(x[2]) << 16) | squeeze8(x[3]);
        bad[i >> 5] = ~gm;
        if (lower) lower[i >> 5] = lm;
        lowacc |= lm;
    }
    if (lowacc) any_lower = true;
    return i;
}
#endif

inline bool pack_read(const uint8_t* c, const uint8_t* q, uint32_t maxq, uint32_t L, uint32_t* codes, uint32_t* bad, uint32_t* lower, bool avx2) {
    bool any_lower = false;
    uint32_t i = 0;
#if defined(__x86_64__)
    if (avx2) i = pack_blocks_avx2(c, q, maxq, L, codes, bad, lower, any_lower);
#else
    (void)avx2;
#endif
    if (i < L) pack_tail(c, q, maxq, i, L, codes, bad, lower, any_lower);
    return any_lower;
}

inline uint64_t read_words(uint64_t L, bool with_lower) { return ((L + 15) >> 4) + ((L + 31) >> 5) * (with_lower ? 2 : 1); }

}  // namespace

using namespace cid;

extern "C" {

uint64_t cid_pack_words_bound(const uint64_t* seq_offs, const uint64_t* read_offs, uint64_t nreads, int with_lower) {
    if (!seq_offs || !read_offs) return 0;
    uint64_t total = 0;
    for (uint64_t r = 0; r < nreads; r++) total += read_words(seq_offs[read_offs[r + 1]] - seq_offs[read_offs[r]], with_lower != 0);
    return total;
}

int cid_pack_reads(const char* bases, const char* quals, const uint64_t* seq_offs, uint64_t nseq, const uint64_t* read_offs, uint64_t nreads,
                   uint32_t qual_offset, int threads, uint32_t* words, uint64_t words_cap, uint64_t* word_offs, uint32_t* pack_flags) {
    (void)nseq;
    if (!seq_offs || !read_offs || !word_offs || !pack_flags || (nreads && (!bases || !words))) { set_error("cid_pack_reads: null argument"); return CID_E_INVALID; }
    const uint32_t maxq = quals && qual_offset ? qual_offset + 33 : 0;      // seq.rs:45 `q < max_quality_offset + 33`
    const uint8_t* q = maxq ? (const uint8_t*)quals : nullptr;
#if defined(__x86_64__)
    const bool avx2 = __builtin_cpu_supports("avx2");
#else
    const bool avx2 = false;
#endif
    unsigned nt = threads > 0 ? (unsigned)threads : std::max(1u, std::thread::hardware_concurrency());
    nt = (unsigned)std::min<uint64_t>(nt, std::max<uint64_t>(1, nreads / 4096));
    bool with_lower = false;
    for (int pass = 0; pass < 2; pass++) {
        word_offs[0] = 0;
        for (uint64_t r = 0; r < nreads; r++) word_offs[r + 1] = word_offs[r] + read_words(seq_offs[read_offs[r + 1]] - seq_offs[read_offs[r]], with_lower);
        if (word_offs[nreads] > words_cap) { set_error("cid_pack_reads: %llu words needed, capacity %llu", (unsigned long long)word_offs[nreads], (unsigned long long)words_cap); return CID_E_CAPACITY; }
        std::atomic<int> saw_lower{0};
        auto work = [&](uint64_t r_lo, uint64_t r_hi) {
            bool any = false;
            for (uint64_t r = r_lo; r < r_hi; r++) {
                const uint64_t b0 = seq_offs[read_offs[r]], L = seq_offs[read_offs[r + 1]] - b0;
                uint32_t* codes = words + word_offs[r];
                uint32_t* bad = codes + ((L + 15) >> 4);
                uint32_t* lower = with_lower ? bad + ((L + 31) >> 5) : nullptr;
                any |= pack_read((const uint8_t*)bases + b0, q ? q + b0 : nullptr, maxq, (uint32_t)L, codes, bad, lower, avx2);
                if (any && !with_lower) break;           // this pass is void: the batch is packed again with the lower plane
            }
            if (any) saw_lower = 1;
        };
        if (nt <= 1) work(0, nreads);
        else {
            std::vector<std::thread> th;
            const uint64_t per = (nreads + nt - 1) / nt;
            for (unsigned t = 0; t < nt; t++) th.emplace_back(work, std::min(nreads, t * per), std::min(nreads, (t + 1) * per));
            for (auto& x : th) x.join();
        }
        if (!saw_lower.load() || with_lower) break;
        with_lower = true;                                // (rare) soft-masked reads: three planes
    }
    *pack_flags = with_lower ? CID_PACK_LOWER : 0u;
    return CID_OK;
}

}  // extern "C"
