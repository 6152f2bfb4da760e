// One gzip stream decoded by several threads (see GzParallel in fast_inflate.hpp).
//
// The reference inflates on the parsing thread (read_id_mt_pe.rs:727-761); with the GPU classifying a million read pairs in
// 20 ms, a single inflating thread per file is what `read_id` of .fastq.gz waits for.  DEFLATE has no index, so a thread
// that starts in the middle (a) has to find a block header by trying bit positions and (b) does not know the 32 KB of
// output before it.  (a) is checked, not trusted: a span counts only when the span before it arrives at exactly that bit.
// (b) is deferred: unknown bytes travel through the match copies as 16-bit markers and are filled in afterwards.
#include <sys/mman.h>
#include <zlib.h>      // crc32_combine

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>

#include "cid_host.hpp"
#include "fast_inflate.hpp"

namespace cidh {

namespace {

enum : size_t { WIN = 32768 };
enum : uint16_t { MARK = 0x8000 };       // symbol >= MARK: byte (symbol - MARK) of the window before the span
enum : uint32_t { F_LIT = GzInflater::F_LIT, F_BASE = GzInflater::F_BASE, F_EOB = GzInflater::F_EOB, F_SUB = GzInflater::F_SUB };
enum : unsigned { LBITS = GzInflater::LBITS, DBITS = GzInflater::DBITS };

struct BitIn {
    const uint8_t* base = nullptr; const uint8_t* in = nullptr; const uint8_t* end = nullptr;
    uint64_t bb = 0; unsigned bc = 0; uint64_t fedz = 0;
    inline void refill() {
        if (in + 8 <= end) {
            uint64_t w;
            memcpy(&w, in, 8);
            bb |= w << bc;
            in += (63 - bc) >> 3;
            bc |= 56;
            return;
        }
        while (bc <= 56) {
            uint64_t b = 0;
            if (in < end) b = *in++; else fedz++;
            bb |= b << bc;
            bc += 8;
        }
    }
    inline uint32_t bits(unsigned n) {
        const uint32_t v = (uint32_t)(bb & ((1ull << n) - 1));
        bb >>= n; bc -= n;
        return v;
    }
    void seek(const uint8_t* b, size_t n, uint64_t bitpos) {
        base = b; end = b + n; in = b + (bitpos >> 3); bb = 0; bc = 0; fedz = 0;
        refill();
        bits((unsigned)(bitpos & 7));
    }
    uint64_t bitpos() const { return ((uint64_t)(in - base) + fedz) * 8 - bc; }
    bool overrun() const { return fedz * 8 > bc; }          // consumed bits that were not in the input
};

// Kraft sum of a set of code lengths (<= 15): 0 complete, > 0 incomplete, < 0 over-subscribed; maxlen = longest code
int kraft_left(const uint8_t* lens, unsigned n, unsigned& maxlen) {
    unsigned count[16] = {0};
    for (unsigned i = 0; i < n; i++) count[lens[i]]++;
    int left = 1;
    maxlen = 0;
    for (unsigned l = 1; l <= 15; l++) {
        left <<= 1;
        left -= (int)count[l];
        if (left < 0) return -1;
        if (count[l]) maxlen = l;
    }
    return left;
}

// HLIT / HDIST / HCLEN and the code lengths of a dynamic block, the reader just past the 3 header bits.  `strict`: zlib's
// rules for a valid set (complete codes, or a single 1-bit code) -- what a block header found by search has to satisfy; else
// the rules of GzInflater::build_dynamic, so that both decoders take and refuse the same streams.
const char* read_dynamic(BitIn& br, uint8_t lens[320], unsigned& hlit, unsigned& hdist, bool strict) {
    br.refill();
    hlit = br.bits(5) + 257; hdist = br.bits(5) + 1;
    const unsigned hclen = br.bits(4) + 4;
    if (hlit > 286 || hdist > 30) return "bad code counts";
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t cl[19] = {0};
    for (unsigned i = 0; i < hclen; i++) { br.refill(); cl[order[i]] = (uint8_t)br.bits(3); }
    unsigned mx;
    if (strict && kraft_left(cl, 19, mx) != 0) return "bad code-length code";
    static thread_local std::vector<uint32_t> ct;
    bool ok = false;
    GzInflater::build_table(cl, 19, 7, false, ct, ok);
    if (!ok) return "bad code-length code";
    memset(lens, 0, 320);
    unsigned i = 0;
    while (i < hlit + hdist) {
        br.refill();
        const uint32_t e = ct[br.bb & 127];
        if (!(e & F_LIT)) return "bad code-length symbol";
        br.bits(e & 0xFF);
        const unsigned sym = e >> 16;
        if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
        unsigned rep, val = 0;
        if (sym == 16) { if (!i) return "repeat without a previous length"; val = lens[i - 1]; rep = 3 + br.bits(2); }
        else if (sym == 17) rep = 3 + br.bits(3);
        else rep = 11 + br.bits(7);
        if (i + rep > hlit + hdist) return "code lengths overrun";
        while (rep--) lens[i++] = (uint8_t)val;
    }
    if (!lens[256]) return "no end-of-block code";
    if (br.overrun()) return "unexpected end of data";
    if (strict) {
        int left = kraft_left(lens, hlit, mx);
        if (left < 0 || (left > 0 && mx != 1)) return "bad literal/length code";
        left = kraft_left(lens + hlit, hdist, mx);
        if (left < 0 || (left > 0 && mx > 1)) return "bad distance code";
    }
    return nullptr;
}

// First bit position in [from, to) that can be the header of a non-final dynamic-Huffman block going by its first 17 bits and
// its code-length code (which must be complete); UINT64_MAX if none.  Positions within 32 bytes of the end are not tried.
uint64_t find_candidate(const uint8_t* base, size_t n, uint64_t from, uint64_t to) {
    if (n < 32) return UINT64_MAX;
    const uint64_t last = (uint64_t)(n - 32) * 8;
    if (to > last) to = last;
    for (uint64_t p = from; p < to; p++) {
        const uint8_t* q = base + (p >> 3);
        const unsigned r = (unsigned)(p & 7);
        uint64_t w0;
        memcpy(&w0, q, 8);
        const uint64_t v = w0 >> r;                         // >= 57 bits
        if ((v & 7) != 4) continue;                         // BFINAL = 0, BTYPE = 2
        if (((v >> 3) & 31) > 29 || ((v >> 8) & 31) > 29) continue;
        const unsigned hclen = (unsigned)((v >> 13) & 15) + 4;
        uint64_t w1;
        memcpy(&w1, q + 2, 8);                              // bits 16 .. 79 of the byte-aligned view
        uint64_t c = w1 >> (r + 1);                         // bit p + 17 onwards, >= 56 bits
        unsigned sum = 0;
        unsigned i = 0;
        for (; i < hclen && i < 18; i++) { const unsigned l = (unsigned)(c & 7); c >>= 3; if (l) sum += 128u >> l; }
        if (i < hclen) {                                    // the 19th length: bits p + 17 + 54 ..
            uint64_t w2;
            memcpy(&w2, q + 8, 8);
            const unsigned sh = r + 17 + 54 - 64;
            const unsigned l = (unsigned)((w2 >> sh) & 7);
            if (l) sum += 128u >> l;
        }
        if (sum == 128) return p;
    }
    return UINT64_MAX;
}

struct Tables { std::vector<uint32_t> lt, dt; };

const Tables& fixed_tables() {
    static const Tables* t = [] {
        Tables* x = new Tables;
        uint8_t lens[288], dl[32];
        GzInflater::fixed_lengths(lens, dl);
        bool ok = false;
        GzInflater::build_table(lens, 288, LBITS, false, x->lt, ok);
        GzInflater::build_table(dl, 32, DBITS, true, x->dt, ok);
        return x;
    }();
    return *t;
}

// Symbols of one Huffman block into 16-bit cells.  0 = end of block, 1 = `limit` reached, 2 = error.
int run_huff16(BitIn& br, const uint32_t* lt, const uint32_t* dt, uint16_t* ob, size_t& out_io, size_t limit, size_t min_index, const char*& err) {
    size_t out = out_io;
    uint64_t bb = br.bb;
    unsigned bc = br.bc;
    const uint8_t* in = br.in;
    const uint8_t* const end = br.end;
    uint64_t fedz = br.fedz;
    int rc = 1;
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
#define CID_LIT()                                                                       \
    do {                                                                                \
        ob[out] = (uint16_t)((e >> 16) & 0xFF); ob[out + 1] = (uint16_t)(e >> 24);      \
        out += 1 + ((e >> 7) & 1u);                                                     \
        bb >>= (e & 0x7F); bc -= (e & 0x7F);                                            \
    } while (0)
    while (out < limit) {
        CID_REFILL();
        uint32_t e = lt[bb & ((1u << LBITS) - 1)];
        if (e & F_SUB) {
            const unsigned sb = (e >> 8) & 15;
            e = lt[(e >> 16) + ((bb >> LBITS) & ((1u << sb) - 1))];
            bb >>= LBITS; bc -= LBITS;
        }
        if (e & F_LIT) {
            CID_LIT();
            e = lt[bb & ((1u << LBITS) - 1)];
            if (e & F_LIT) {
                CID_LIT();
                e = lt[bb & ((1u << LBITS) - 1)];
                if (e & F_LIT) {
                    CID_LIT();
                    e = lt[bb & ((1u << LBITS) - 1)];
                    if ((e & (F_LIT | F_SUB)) == F_LIT) CID_LIT();
                }
            }
            continue;
        }
        if (e & F_EOB) {
            bb >>= (e & 0xFF); bc -= (e & 0xFF);
            rc = 0;
            break;
        }
        if (!(e & F_BASE)) { err = "invalid literal/length code"; rc = 2; break; }
        bb >>= (e & 0xFF); bc -= (e & 0xFF);
        const unsigned lx = (e >> 8) & 15;
        const unsigned len = (e >> 16) + (unsigned)(bb & ((1u << lx) - 1));
        bb >>= lx; bc -= lx;
        uint32_t d = dt[bb & ((1u << DBITS) - 1)];
        if (d & F_SUB) {
            const unsigned sb = (d >> 8) & 15;
            d = dt[(d >> 16) + ((bb >> DBITS) & ((1u << sb) - 1))];
            bb >>= DBITS; bc -= DBITS;
        }
        if (!(d & F_BASE)) { err = "invalid distance code"; rc = 2; break; }
        bb >>= (d & 0xFF); bc -= (d & 0xFF);
        const unsigned dx = (d >> 8) & 15;
        const size_t dist = (d >> 16) + (size_t)(bb & ((1u << dx) - 1));
        bb >>= dx; bc -= dx;
        if (out < dist + min_index) { err = "distance reaches before the start of the data"; rc = 2; break; }
        const uint16_t* src = ob + out - dist;
        uint16_t* dst = ob + out;
        if (dist >= 8) {
            for (unsigned i = 0; i < len; i += 8) memcpy(dst + i, src + i, 16);
        } else if (dist == 1) {
            const uint16_t v = src[0];
            for (unsigned i = 0; i < len; i++) dst[i] = v;
        } else {
            const unsigned D = (unsigned)dist * ((7 + (unsigned)dist) / (unsigned)dist);
            unsigned i = 0;
            for (; i < D && i < len; i++) dst[i] = src[i];
            for (; i < len; i += 8) memcpy(dst + i, dst + i - D, 16);
        }
        out += len;
    }
#undef CID_LIT
#undef CID_REFILL
    out_io = out; br.bb = bb; br.bc = bc; br.in = in; br.fedz = fedz;
    return rc;
}

// Buffers of tens of megabytes that are written once per round: 2 MB-aligned and advised as huge pages, because with 4 KB
// pages the first touch of every fresh buffer cost as much as resolving into it.
void* big_alloc(size_t bytes) {
    const size_t two_mb = (size_t)2 << 20;
    const size_t r = (bytes + two_mb - 1) & ~(two_mb - 1);
    void* p = nullptr;
    if (posix_memalign(&p, two_mb, r) != 0) throw std::bad_alloc();
    madvise(p, r, MADV_HUGEPAGE);
    return p;
}
struct BigFree { void operator()(void* p) const { free(p); } };

struct Cells {                               // 16-bit output of a span: [0, WIN) the window before it, then its symbols
    std::unique_ptr<uint16_t, BigFree> p;
    size_t cap = 0;
    void grow(size_t need, size_t keep) {
        if (need <= cap) return;
        size_t c = cap ? cap : ((size_t)1 << 20);
        while (c < need) c += c / 2;
        std::unique_ptr<uint16_t, BigFree> q((uint16_t*)big_alloc(c * sizeof(uint16_t)));
        if (keep) memcpy(q.get(), p.get(), keep * sizeof(uint16_t));
        p.swap(q);
        cap = c;
    }
};

struct Bytes {
    std::unique_ptr<uint8_t, BigFree> p;
    size_t cap = 0, n = 0;
    void need(size_t c) { if (c > cap) { p.reset((uint8_t*)big_alloc(c)); cap = c; } }
};

enum End { E_NONE, E_SYNC, E_ROUND_END, E_MEMBER_END, E_NOSYNC, E_GIVE_UP };

struct Span {
    // in
    uint64_t from_bit = 0, to_bit = 0;       // where the span may start (span 0: from_bit is the verified start)
    size_t min_index = 0;                    // span 0: WIN - valid history
    // out
    Cells cells;
    size_t out = WIN;                        // write position in cells
    uint64_t start_bit = 0, end_bit = 0;
    End end = E_NONE;
    int next = -1;                           // E_SYNC: the span whose start this one arrived at
    uint32_t crc = 0;                        // of the resolved bytes
    size_t at = 0;                           // offset of the resolved bytes in the round's buffer
    double t_search = 0, t_total = 0, t_cpu = 0;        // seconds (trace)
    unsigned n_cand = 0;
};
double cpu_s() { timespec ts; clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; }
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

struct GzParallel::Impl {
    const uint8_t* base; size_t n;
    std::string what;
    unsigned T; size_t span;
    size_t work_span;                        // span of the next round: `span`, less where the data inflates by more than 8x
    size_t soft_cap, hard_cap;               // output symbols of one span: stop at the next block / give up

    // coordinator state: the stream is verified up to here
    uint64_t pos_bit = 0;
    bool at_header = true;
    uint8_t win[WIN];
    size_t win_valid = 0;
    uint32_t crc = 0;
    uint64_t member_len = 0;
    uint64_t member_from_byte = 0;
    unsigned small_members = 0;
    bool trace = getenv("COLORID_B200_GZ_TRACE") != nullptr;

    std::vector<Span> sets[2];               // the spans of even and odd rounds (one set is decoded into while the other is resolved)
    Span* spans = nullptr;                   // the current round's
    std::vector<std::atomic<int64_t>> sync;  // -1 not known yet, -2 none, else the bit the span started at
    uint64_t round_end_bit = 0;
    int n_spans = 0;

    // output queue
    struct Item { Bytes data; bool eof = false; std::string error; };
    std::mutex mu;
    std::condition_variable cv_put, cv_get;
    std::deque<Item> queue;
    std::vector<Bytes> spare;
    std::atomic<bool> abort{false};
    std::thread coord;
    Item cur; size_t cur_pos = 0; bool done = false;

    Impl(const uint8_t* d, size_t len, const std::string& w, const Config& cfg)
        : base(d), n(len), what(w), T(cfg.threads < 1 ? 1 : cfg.threads), span(cfg.span < 4096 ? 4096 : cfg.span), sync(T) {
        sets[0] = std::vector<Span>(T);
        sets[1] = std::vector<Span>(T);
        work_span = span;
        soft_cap = span * 16;
        hard_cap = span * 48;
    }

    // ---------------------------------------------------------------- queue
    Bytes take_spare() {
        std::lock_guard<std::mutex> g(mu);
        if (spare.empty()) return Bytes();
        Bytes b = std::move(spare.back());
        spare.pop_back();
        return b;
    }
    bool put(Item&& it) {
        std::unique_lock<std::mutex> g(mu);
        cv_put.wait(g, [&] { return queue.size() < 2 || abort.load(); });
        if (abort.load()) return false;
        queue.push_back(std::move(it));
        cv_get.notify_one();
        return true;
    }

    // ---------------------------------------------------------------- one span
    // Decodes blocks from the reader's position (a block header) until a stop condition holds at a block boundary.
    // first_only: return after one block (true = it decoded cleanly).
    bool decode_blocks(BitIn& br, Span& s, int idx, bool first_strict, bool first_only, const char*& err) {
        Tables dyn;
        uint8_t lens[320];
        bool first = true;
        for (;;) {
            br.refill();
            if (br.overrun()) { err = "unexpected end of data"; return false; }
            const bool bfinal = br.bits(1) != 0;
            const unsigned type = br.bits(2);
            if (type == 0) {
                br.bits(br.bc & 7);
                br.refill();
                const unsigned len = br.bits(16), nlen = br.bits(16);
                if (br.overrun()) { err = "truncated stored block"; return false; }
                if ((len ^ 0xFFFFu) != nlen) { err = "stored block length check failed"; return false; }
                const uint64_t byte = br.bitpos() >> 3;
                if (byte + len > n) { err = "truncated stored block"; return false; }
                s.cells.grow(s.out + len + 512, s.out);
                uint16_t* ob = s.cells.p.get();
                for (unsigned i = 0; i < len; i++) ob[s.out + i] = base[byte + i];
                s.out += len;
                br.seek(base, n, (byte + len) * 8);
            } else if (type == 3) {
                err = "reserved block type";
                return false;
            } else {
                const Tables* t = &fixed_tables();
                if (type == 2) {
                    unsigned hlit, hdist;
                    if ((err = read_dynamic(br, lens, hlit, hdist, first && first_strict)) != nullptr) return false;
                    bool ok = false;
                    GzInflater::build_table(lens, hlit, LBITS, false, dyn.lt, ok);
                    if (!ok) { err = "bad literal/length code"; return false; }
                    GzInflater::build_table(lens + hlit, hdist, DBITS, true, dyn.dt, ok);
                    if (!ok) { err = "bad distance code"; return false; }
                    t = &dyn;
                }
                for (;;) {
                    s.cells.grow(s.out + ((size_t)1 << 16), s.out);
                    const int rc = run_huff16(br, t->lt.data(), t->dt.data(), s.cells.p.get(), s.out, s.cells.cap - 600, s.min_index, err);
                    if (rc == 2) return false;
                    if (br.overrun()) { err = "unexpected end of data"; return false; }
                    if (rc == 0) break;
                    if (s.out - WIN > hard_cap) { err = nullptr; s.end = E_GIVE_UP; return false; }
                    s.cells.grow(s.cells.cap + s.cells.cap / 2, s.out);
                }
            }
            if (first_only) return true;
            first = false;
            // ---- a block boundary
            const uint64_t b = br.bitpos();
            s.end_bit = b;
            if (bfinal) { s.end = E_MEMBER_END; return true; }
            if (abort.load(std::memory_order_relaxed)) { s.end = E_GIVE_UP; return false; }
            if (s.out - WIN > soft_cap) { s.end = E_ROUND_END; return true; }
            int& nxt = s.next;
            if (nxt < 0) nxt = idx + 1;
            for (;;) {
                if (nxt >= n_spans) {
                    if (b >= round_end_bit) { s.end = E_ROUND_END; nxt = -1; return true; }
                    break;
                }
                if (b < spans[nxt].from_bit) break;
                int64_t sy;
                while ((sy = sync[nxt].load(std::memory_order_acquire)) == -1) {
                    if (abort.load(std::memory_order_relaxed)) { s.end = E_GIVE_UP; return false; }
                    std::this_thread::sleep_for(std::chrono::microseconds(20));
                }
                if (sy == -2 || b > (uint64_t)sy) { nxt++; continue; }      // no start found there, or one this stream does not pass through
                if (b == (uint64_t)sy) { s.end = E_SYNC; return true; }
                break;
            }
        }
    }

    // a round's thread: its share of the round before, then its span.  Nothing may leave a thread as an exception (out of
    // memory while growing a buffer, say): the span then counts as given up and the stream continues from the one before it.
    std::atomic<bool> thread_failed{false};
    void round_thread(int idx) noexcept {
        try { job_work(); } catch (...) { thread_failed.store(true); }
        try { run_span(idx); } catch (...) {
            spans[idx].end = idx ? E_NOSYNC : E_GIVE_UP;
            if (idx) { int64_t e = -1; sync[idx].compare_exchange_strong(e, -2); }
        }
    }
    void run_span(int idx) {
        Span& s = spans[idx];
        s.end = E_NONE; s.next = -1; s.out = WIN; s.n_cand = 0;
        const double t0 = now_s();
        struct Stamp { Span& s; double t0, c0; ~Stamp() { s.t_total = now_s() - t0; s.t_cpu = cpu_s() - c0; } } stamp{s, t0, cpu_s()};
        BitIn br;
        const char* err = nullptr;
        if (idx == 0) {
            s.start_bit = s.from_bit;
            br.seek(base, n, s.start_bit);
            if (!decode_blocks(br, s, idx, false, false, err)) s.end = E_GIVE_UP;
            return;
        }
        struct Publish {                              // whatever happens, the spans before this one must not wait for ever
            std::atomic<int64_t>& a;
            ~Publish() { int64_t e = -1; a.compare_exchange_strong(e, -2); }
        } publish{sync[idx]};
        s.min_index = 0;
        s.cells.grow(WIN + span * 5, 0);
        {
            uint16_t* w = s.cells.p.get();
            for (size_t i = 0; i < WIN; i++) w[i] = (uint16_t)(MARK | i);
        }
        uint64_t p = s.from_bit;
        for (;;) {
            if (abort.load(std::memory_order_relaxed)) { s.end = E_NOSYNC; return; }
            p = find_candidate(base, n, p, s.to_bit);
            if (p == UINT64_MAX) { s.end = E_NOSYNC; return; }
            s.out = WIN;
            br.seek(base, n, p);
            s.n_cand++;
            if (decode_blocks(br, s, idx, true, true, err)) break;
            if (s.end == E_GIVE_UP) { s.end = E_NOSYNC; return; }
            p++;
        }
        s.start_bit = p;
        s.t_search = now_s() - t0;
        sync[idx].store((int64_t)p, std::memory_order_release);
        // the first block is in; its end is a block boundary like any other, but the stop conditions are only evaluated after a
        // block decoded by the loop below -- a span that found its start always continues by at least one more block, unless the
        // first one was final, which find_candidate excludes
        if (!decode_blocks(br, s, idx, false, false, err)) s.end = E_GIVE_UP;
    }

    // ---------------------------------------------------------------- resolve
    static void make_lut(const uint8_t* w, uint8_t* lut) {
        memset(lut, 0, 65536);
        for (unsigned i = 0; i < 256; i++) lut[i] = (uint8_t)i;
        memcpy(lut + MARK, w, WIN);
    }
    static void resolve(const uint16_t* __restrict src, size_t cnt, const uint8_t* __restrict lut, uint8_t* __restrict dst) {
        size_t i = 0;
        for (; i + 8 <= cnt; i += 8) {           // eight look-ups, one store
            const uint64_t r = (uint64_t)lut[src[i]] | ((uint64_t)lut[src[i + 1]] << 8) | ((uint64_t)lut[src[i + 2]] << 16) | ((uint64_t)lut[src[i + 3]] << 24) |
                               ((uint64_t)lut[src[i + 4]] << 32) | ((uint64_t)lut[src[i + 5]] << 40) | ((uint64_t)lut[src[i + 6]] << 48) | ((uint64_t)lut[src[i + 7]] << 56);
            memcpy(dst + i, &r, 8);
        }
        for (; i < cnt; i++) dst[i] = lut[src[i]];
    }

    [[noreturn]] void fail(const char* why) const { throw Error("gzip stream of " + what + ": " + why); }

    // ---------------------------------------------------------------- the rest of the stream on one thread
    void sequential() {
        GzInflater seq(base, n, what);
        seq.resume(pos_bit, at_header, win + (WIN - win_valid), win_valid, crc, member_len);
        for (;;) {
            Item it;
            it.data = take_spare();
            it.data.need((size_t)4 << 20);
            it.data.n = seq.read(it.data.p.get(), (size_t)4 << 20);
            if (!it.data.n) break;
            if (!put(std::move(it))) return;
        }
    }

    // ---------------------------------------------------------------- markers -> bytes of the round before
    // The spans of a round are resolved by the threads of the NEXT round before they start decoding (its cells are a second
    // set), so that no thread idles between rounds; a round that ends a member is resolved at once, for the trailer check.
    struct Job {
        std::vector<Span*> chain;
        std::vector<std::unique_ptr<uint8_t[]>> luts;
        Item item;
        std::atomic<size_t> nextk{0};
        bool active = false;
    } job;
    void job_work() {
        if (!job.active) return;
        uint8_t* dst = job.item.data.p.get();
        for (size_t k; (k = job.nextk.fetch_add(1)) < job.chain.size();) {
            Span& s = *job.chain[k];
            const size_t cnt = s.out - WIN;
            uint32_t c = (uint32_t)crc32(0L, Z_NULL, 0);
            for (size_t o = 0; o < cnt; o += (size_t)16 << 10) {             // (the CRC reads what was just written: 16 KB at a time, from L1)
                const size_t m = std::min<size_t>((size_t)16 << 10, cnt - o);
                resolve(s.cells.p.get() + WIN + o, m, job.luts[k].get(), dst + s.at + o);
                c = GzInflater::crc32_fast(c, dst + s.at + o, m);
            }
            s.crc = c;
        }
    }
    // all of the job's spans are resolved: fold the CRCs, hand the bytes on.  false = the reader is gone
    bool job_finish() {
        if (!job.active) return true;
        job.active = false;
        for (Span* s : job.chain) crc = (uint32_t)crc32_combine(crc, s->crc, (z_off_t)(s->out - WIN));
        if (!job.item.data.n) return true;
        return put(std::move(job.item));
    }
    bool job_flush() {                       // on the coordinator alone
        job_work();
        return job_finish();
    }

    // ---------------------------------------------------------------- coordinator
    void run() {
        try {
            for (unsigned round = 0;; round++) {
                if (abort.load()) return;
                if (at_header) {
                    const char* why = nullptr;
                    const uint64_t byte = pos_bit >> 3;
                    const size_t h = GzInflater::header_size(base + byte, base + n, &why);
                    if (!h) fail(why);
                    member_from_byte = byte;
                    pos_bit = (byte + h) * 8;
                    at_header = false;
                    crc = (uint32_t)crc32(0L, Z_NULL, 0);
                    member_len = 0;
                    win_valid = 0;
                }
                const uint64_t pos_byte = pos_bit >> 3;
                const size_t remaining = n - (size_t)pos_byte;
                if (T < 2 || remaining < work_span || small_members >= 4) {
                    if (!job_flush()) return;
                    sequential();
                    break;
                }
                // ---- lay out the round
                spans = sets[round & 1].data();
                size_t sp = work_span;
                n_spans = (int)T;
                if (remaining < (size_t)T * sp) {
                    n_spans = (int)std::max<size_t>(1, std::min<size_t>(T, remaining / (sp / 2)));
                    sp = (remaining + n_spans - 1) / n_spans;
                }
                for (int j = 0; j < n_spans; j++) {
                    spans[j].from_bit = j ? (pos_byte + (uint64_t)j * sp) * 8 : pos_bit;
                    spans[j].to_bit = (pos_byte + (uint64_t)(j + 1) * sp) * 8;
                    sync[j].store(-1);
                }
                round_end_bit = std::min<uint64_t>((uint64_t)n, pos_byte + (uint64_t)n_spans * sp) * 8;
                // span 0 knows its window
                Span& s0 = spans[0];
                s0.min_index = WIN - win_valid;
                s0.cells.grow(WIN + span * 5, 0);
                for (size_t i = WIN - win_valid; i < WIN; i++) s0.cells.p.get()[i] = win[i];
                // ---- phase A: resolve the round before, decode this one
                const double tA = now_s();
                {
                    std::vector<std::thread> th;
                    int started = 1;
                    try {
                        for (int j = 1; j < n_spans; j++) { th.emplace_back([this, j] { round_thread(j); }); started = j + 1; }
                    } catch (...) {
                        for (int j = started; j < n_spans; j++) { spans[j].end = E_NOSYNC; sync[j].store(-2); }
                    }
                    round_thread(0);
                    for (auto& t : th) t.join();
                }
                if (thread_failed.load()) fail("internal error while resolving a round");
                if (!job_finish()) return;
                if (abort.load()) return;
                // ---- the chain of spans that are the stream
                std::vector<Span*> chain;
                for (int c = 0;;) {
                    Span& s = spans[c];
                    if (s.end == E_GIVE_UP || s.end == E_NONE || s.end == E_NOSYNC) break;       // its output is not used
                    chain.push_back(&s);
                    if (s.end != E_SYNC) break;
                    c = s.next;
                }
                if (chain.empty()) { sequential(); break; }                                         // (also reports a damaged stream)
                const double tB = now_s();
                // ---- phase B: offsets, the window after every span, the tables that turn its markers into bytes
                size_t total = 0;
                for (Span* s : chain) { s->at = total; total += s->out - WIN; }
                job.item = Item();
                job.item.data = take_spare();
                job.item.data.need(total + total / 4 + 64);
                job.item.data.n = total;
                job.luts.resize(std::max(job.luts.size(), chain.size()));
                for (size_t k = 0; k < chain.size(); k++) {
                    Span& s = *chain[k];
                    const size_t cnt = s.out - WIN;
                    if (!job.luts[k]) job.luts[k].reset(new uint8_t[65536]);
                    make_lut(win, job.luts[k].get());
                    if (k > 0 && win_valid < WIN) {
                        // the member is younger than a window: a marker may point before its start
                        const uint16_t* src = s.cells.p.get() + WIN;
                        for (size_t i = 0; i < cnt; i++)
                            if (src[i] >= MARK && (size_t)(src[i] - MARK) < WIN - win_valid) fail("distance reaches before the start of the data");
                    }
                    if (cnt >= WIN) resolve(s.cells.p.get() + s.out - WIN, WIN, job.luts[k].get(), win);
                    else {
                        memmove(win, win + cnt, WIN - cnt);
                        resolve(s.cells.p.get() + WIN, cnt, job.luts[k].get(), win + (WIN - cnt));
                    }
                    member_len += cnt;
                    win_valid = (size_t)std::min<uint64_t>(WIN, member_len);
                }
                job.chain = chain;
                job.nextk.store(0);
                job.active = true;
                {
                    // data that inflates by more than 8x gets shorter spans, so that a span's output stays around 8 x `span` symbols
                    // and below the cap at which a span stops early (which would leave the spans after it unused)
                    const uint64_t used_bits = chain.back()->end_bit - pos_bit;
                    const double ratio = used_bits ? (double)total * 8.0 / (double)used_bits : 1.0;
                    const size_t lo = std::max<size_t>(4096, span / 32);
                    work_span = ratio > 8.0 ? std::max(lo, (size_t)((double)span * 8.0 / ratio)) : span;
                }
                const Span& last = *chain.back();
                const End how = last.end;
                const uint64_t end_bit = last.end_bit;
                if (trace) {
                    fprintf(stderr, "[gz] round at byte %llu: %d spans, chain %zu, %zu bytes out; resolve + decode %.1f ms, windows %.1f ms;", (unsigned long long)pos_byte, n_spans, chain.size(), total, (tB - tA) * 1e3, (now_s() - tB) * 1e3);
                    for (int j = 0; j < n_spans; j++) fprintf(stderr, " [%d: end %d cand %u +%llu search %.1f total %.1f cpu %.1f out %zu]", j, (int)spans[j].end, spans[j].n_cand, (unsigned long long)(spans[j].start_bit - spans[j].from_bit), spans[j].t_search * 1e3, spans[j].t_total * 1e3, spans[j].t_cpu * 1e3, spans[j].out - WIN);
                    fprintf(stderr, "\n");
                }
                if (how == E_MEMBER_END) {
                    if (!job_flush()) return;
                    const uint64_t byte = (end_bit + 7) >> 3;
                    if (byte + 8 > n) fail("truncated trailer");
                    uint32_t want_crc, want_len;
                    memcpy(&want_crc, base + byte, 4);
                    memcpy(&want_len, base + byte + 4, 4);
                    if (want_crc != crc) fail("CRC mismatch");
                    if (want_len != (uint32_t)member_len) fail("length mismatch");
                    const uint64_t nb = byte + 8;
                    if (!(n - nb >= 2 && base[nb] == 0x1f && base[nb + 1] == 0x8b)) break;          // the end (trailing bytes ignored, as gzread does)
                    if (nb - member_from_byte < span) small_members++;                             // bgzip-like files: not worth a round each
                    pos_bit = nb * 8;
                    at_header = true;
                } else {
                    pos_bit = end_bit;                  // E_ROUND_END / E_SYNC to a span that gave up: a block boundary of the stream
                }
            }
            if (!job_flush()) return;
            Item e;
            e.eof = true;
            put(std::move(e));
        } catch (const std::exception& ex) {
            // what was decoded before the damage is still delivered, as the sequential decoder would
            if (!thread_failed.load()) { try { job_flush(); } catch (...) {} }
            Item e;
            e.eof = true;
            e.error = ex.what();
            put(std::move(e));
        }
    }
};

GzParallel::GzParallel(const uint8_t* data, size_t n, const std::string& what, const Config& cfg) : impl_(new Impl(data, n, what, cfg)) {
    impl_->coord = std::thread([this] { impl_->run(); });
}

GzParallel::~GzParallel() {
    {
        std::lock_guard<std::mutex> g(impl_->mu);
        impl_->abort.store(true);
    }
    impl_->cv_put.notify_all();
    if (impl_->coord.joinable()) impl_->coord.join();
    delete impl_;
}

size_t GzParallel::read(uint8_t* dst, size_t cap) {
    Impl& m = *impl_;
    size_t got = 0;
    while (got < cap && !m.done) {
        if (m.cur_pos == m.cur.data.n) {
            std::unique_lock<std::mutex> g(m.mu);
            if (m.cur.data.cap) { m.spare.push_back(std::move(m.cur.data)); m.cur.data = Bytes(); }
            m.cur_pos = 0;
            m.cv_get.wait(g, [&] { return !m.queue.empty(); });
            m.cur = std::move(m.queue.front());
            m.queue.pop_front();
            m.cv_put.notify_one();
            if (m.cur.eof) {
                m.done = true;
                if (!m.cur.error.empty()) throw Error(m.cur.error);
                break;
            }
            continue;
        }
        const size_t k = std::min(cap - got, m.cur.data.n - m.cur_pos);
        memcpy(dst + got, m.cur.data.p.get() + m.cur_pos, k);
        m.cur_pos += k; got += k;
    }
    return got;
}

}  // namespace cidh
