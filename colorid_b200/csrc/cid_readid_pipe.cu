// Host-pointer read_id entry points: read_id_mt_pe.rs:282-363 parallel_vec (+ the streams that feed
// it, :701-951).  The reference classifies one batch of reads at a time with no overlap between
// parsing and classification (:762-790).  Here a batch is cut into chunks that flow through a
// small ring of slots, each with its own CUDA stream and device buffers:
//
//     H2D (bases, quals, offsets)  ->  readid_kmerize / readid_order / readid_vote  ->  D2H reports
//                                                                                   ->  host vote
//
// so the PCIe copies of chunk c+1, the kernels of chunk c and the host-side kmer_poll_plus of
// chunk c-1 run concurrently.  Kernels of neighbouring chunks also overlap on the GPU, which lets
// the DRAM-access-bound vote kernel hide under the issue-bound kmerize/order kernels.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "cid_internal.h"

struct cid_readid_pipe {
    enum { NS = 3 };
    struct Slot {
        cudaStream_t st = nullptr;
        cudaStream_t cst = nullptr;          // the chunk's input copies run on their own stream (see the issue loop)
        cudaEvent_t h2d_done = nullptr;
        cudaEvent_t done = nullptr;
        cid::DevBuf bases, quals, seq_offs, read_offs, entries, order, nocc, n_set, flags, rep_n, rep, ord_out, big, vp;
        cid::DevBuf kind, hits, n_top, top, list, cursor;      // fused vote: device classification + undecided list
        cid::PinBuf h_cursor, h_list, h_offs;
        cudaEvent_t offs_done = nullptr;     // the staged offsets have been consumed by their H2D copies
        cudaEvent_t kern_done = nullptr;     // this chunk's kernels have finished (the next chunk's kernels wait for it)
    } slot[NS];
    cid::DevBuf fp;            // false_prob per colour (f64), uploaded per call
    bool ready = false;
};

namespace cid {

void default_readid_params(cid_readid_params& p, const cid_readid_params* in, uint32_t N) {
    if (in) p = *in;
    else { p.downsample = 1; p.start_sample = 3; p.qual_offset = 0; p.group_width = 16; p.reserve_before_find = 1; p.rep_cap = 0; }
    if (p.downsample == 0) p.downsample = 1;
    if (p.group_width == 0) p.group_width = 16;
    if (p.rep_cap == 0) p.rep_cap = N + 1;
}

void readid_pipe_destroy(cid_ctx* ctx) {
    cid_readid_pipe* pp = ctx->pipe;
    if (!pp) return;
    for (auto& s : pp->slot) {
        if (s.cst) cudaStreamSynchronize(s.cst);
        if (s.st) cudaStreamSynchronize(s.st);
        for (DevBuf* b : {&s.bases, &s.quals, &s.seq_offs, &s.read_offs, &s.entries, &s.order, &s.nocc, &s.n_set, &s.flags,
                          &s.rep_n, &s.rep, &s.ord_out, &s.big, &s.vp, &s.kind, &s.hits, &s.n_top, &s.top, &s.list, &s.cursor})
            b->release();
        for (PinBuf* b : {&s.h_cursor, &s.h_list, &s.h_offs}) b->release();
        if (s.offs_done) cudaEventDestroy(s.offs_done);
        if (s.kern_done) cudaEventDestroy(s.kern_done);
        if (s.done) cudaEventDestroy(s.done);
        if (s.st) cudaStreamDestroy(s.st);
        if (s.cst) cudaStreamDestroy(s.cst);
        if (s.h2d_done) cudaEventDestroy(s.h2d_done);
    }
    pp->fp.release();
    delete pp;
    ctx->pipe = nullptr;
}

static int pipe_get(cid_ctx* ctx, cid_readid_pipe** out) {
    if (!ctx->pipe) ctx->pipe = new cid_readid_pipe();
    cid_readid_pipe* pp = ctx->pipe;
    if (!pp->ready) {
        for (auto& s : pp->slot) {
            CID_CUDA(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
            CID_CUDA(cudaStreamCreateWithFlags(&s.cst, cudaStreamNonBlocking));
            CID_CUDA(cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
            CID_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
            CID_CUDA(cudaEventCreateWithFlags(&s.offs_done, cudaEventDisableTiming));
            CID_CUDA(cudaEventCreateWithFlags(&s.kern_done, cudaEventDisableTiming));
        }
        pp->ready = true;
    }
    *out = pp;
    return CID_OK;
}

// host scan of the read geometry: longest read and most k-mer start positions, over all reads and over the reads the
// warp-per-read kernels hold (<= READID_FAST_BASES bases; the others take the general path)
struct Geometry { uint32_t max_bases, max_kmers, fast_bases, fast_kmers; };
static Geometry read_geometry(const uint64_t* seq_offs, const uint64_t* read_offs, uint64_t nreads, uint32_t k, uint32_t d) {
    uint64_t mb = 0, mk = 0, fb = 0, fk = 0;
    for (uint64_t r = 0; r < nreads; r++) {
        uint64_t b = seq_offs[read_offs[r + 1]] - seq_offs[read_offs[r]], kk = 0;
        for (uint64_t s = read_offs[r]; s < read_offs[r + 1]; s++) {
            uint64_t l = seq_offs[s + 1] - seq_offs[s];
            if (l >= k) kk += (l - k) / d + 1;
        }
        mb = std::max(mb, b); mk = std::max(mk, kk);
        if (b <= READID_FAST_BASES) { fb = std::max(fb, b); fk = std::max(fk, kk); }
    }
    Geometry g;
    g.max_bases = (uint32_t)std::min<uint64_t>(mb, 0xFFFFFFFFu);
    g.max_kmers = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(mk, 1), 0xFFFFFFFFu);
    g.fast_bases = (uint32_t)fb;
    g.fast_kmers = (uint32_t)std::max<uint64_t>(fk, 1);
    return g;
}
// scratch of the general path for one batch: up to READID_BIG_BUDGET bytes shared by as many CTAs as fit
static const size_t READID_BIG_BUDGET = 6ull << 30;
static int big_scratch(cid_index* ix, DevBuf& buf, uint32_t max_bases, uint32_t max_kmers, ReadIdScratch& scr) {
    size_t bytes; uint32_t ctas;
    readid_big_plan(ix, max_bases, max_kmers, READID_BIG_BUDGET, &bytes, &ctas);
    CID_TRY(buf.ensure(bytes));
    scr.big = buf.as<uint8_t>(); scr.big_ctas = ctas; scr.big_bases = max_bases; scr.big_kmers = max_kmers;
    return CID_OK;
}

struct HostOut {            // where a chunk's results go (user memory, indexed by absolute read)
    uint32_t *n_set = nullptr, *flags = nullptr, *rep_n = nullptr, *rep_colour = nullptr, *rep_count = nullptr;
    uint32_t order_cap = 0;
    uint32_t* order_n = nullptr; uint8_t* order_seq = nullptr; uint32_t* order_pos = nullptr;
    // fused host vote
    const VoteParams* vote = nullptr;
    int threads = 0;
    int32_t* kind = nullptr; uint32_t *hits = nullptr, *n_top = nullptr, *top = nullptr; uint32_t top_cap = 0;
};

struct HostPacked { const uint32_t* words; const uint64_t* word_offs; uint32_t lower; };    // cid_pack_reads output (host)
struct VoteJob { uint64_t r0, nr; int slot; };

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int read_id_pipeline(cid_index* ix, const char* bases, const char* quals, const uint64_t* seq_offs, uint64_t nseq,
                            const uint64_t* read_offs, uint64_t nreads, const cid_readid_params* p, const HostOut& out,
                            const HostPacked* hp = nullptr) {
    cid_ctx* ctx = ix->ctx;
    if (!seq_offs || !read_offs) { set_error("read_id: null argument"); return CID_E_INVALID; }
    CID_CUDA(cudaSetDevice(ctx->device));
    cid_readid_params pp;
    default_readid_params(pp, p, ix->N);
    if (nreads == 0) return CID_OK;
    (void)nseq;
    static const bool trace = getenv("CID_TRACE") != nullptr;     // stage timings on stderr (diagnostics only)
    const double t_start = now_ms();
    double t_vote = 0, t_wait = 0, t_evwait = 0;
    double t_geom = 0;
    const bool use_q = !hp && quals && pp.qual_offset;
    const bool want_rep = out.rep_n != nullptr || out.vote != nullptr;
    const bool fused = out.vote != nullptr;
    cid_readid_pipe* pipe;
    CID_TRY(pipe_get(ctx, &pipe));
    // work queued on the context's own stream (index build / upload) must be visible to the slot streams
    CID_CUDA(cudaStreamSynchronize(ctx->stream));

    // Chunk schedule: a geometric ramp.  Small first chunks start the kernels after a short H2D copy; large
    // later chunks keep the kernels efficient (the order kernel wants >= 10^5 reads per launch) while their
    // copies hide behind the kernels of the chunks before them.  The growth factor is 1.5 because the copy of
    // chunk c+1 must finish within the kernels of chunk c: PCIe delivers ~87 K read pairs/ms (600 B each at
    // 54 GB/s), the kernels consume ~56 K/ms, so a chunk may be at most 1.55x its predecessor.
    // Packed reads are 4x fewer bytes per read than ASCII bases + qualities: one rank alone gets 55 GB/s over PCIe, ~370 K read
    // pairs/ms, six times what the kernels consume, so the ramp can be steep (65,536 -> x4 -> up to 2^20 reads: three chunks
    // per million pairs instead of seven; the kernels are 8 % faster on a 672 K chunk than on seven chunks of 33 K .. 319 K:
    // 46.9 -> 49.4 M pairs/s).  Eight ranks of one node share the host's memory system and get ~16 GB/s each: the steep ramp
    // then starves the GPUs (328 M pairs/s on 8 GPUs against 345 M with x1.5), so it is only taken when the caller has not
    // said that more than two ranks share the host (option host_ranks; cid_mg sets it to its number of GPUs).
    const bool steep = hp != nullptr && ctx->opt_host_ranks <= 2;
    uint64_t chunk_max = ctx->opt_readid_chunk ? ctx->opt_readid_chunk : (steep ? 1048576 : 262144);
    {   // the per-chunk report (and, fused, the undecided-read list) is dense in the accession count: keep one slot's share of it
        // below ~1.5 GB so that a wide index shrinks the chunks instead of exhausting device memory (3 slots are in flight)
        const uint64_t per_read = (uint64_t)pp.rep_cap * 8 + (out.vote ? (2 + 2 * (uint64_t)pp.rep_cap) * 4 : 0);
        const uint64_t fit = std::max<uint64_t>(1024, (3ull << 29) / std::max<uint64_t>(per_read, 1));
        chunk_max = std::min(chunk_max, fit);
    }
    const uint64_t growth_pct = ctx->opt_readid_chunk_growth ? (uint64_t)ctx->opt_readid_chunk_growth : (steep ? 400 : 150);
    std::vector<uint64_t> cuts{0};
    for (uint64_t cur = std::min<uint64_t>(ctx->opt_readid_chunk0 ? ctx->opt_readid_chunk0 : (steep ? 65536 : 32768), chunk_max), at = 0; at < nreads;) {
        uint64_t size = std::min(cur, nreads - at);
        if (nreads - at - size < size / 4) size = nreads - at;      // absorb a short tail
        at += size;
        cuts.push_back(at);
        cur = std::min(cur * growth_pct / 100, chunk_max);
    }
    const uint64_t nchunks = cuts.size() - 1;
    const uint64_t chunk = chunk_max;
    const int NS = cid_readid_pipe::NS;
    const size_t rc = pp.rep_cap;

    // fused vote: a worker thread classifies chunk c while the GPU runs chunks c+1, c+2
    std::mutex mu;
    std::condition_variable cv;
    std::deque<VoteJob> jobs;
    bool slot_busy[cid_readid_pipe::NS] = {false, false, false};
    bool closing = false;
    int worker_rc = CID_OK;
    std::thread worker;
    std::vector<uint32_t> sub_idx, sub_n_set, sub_flags, sub_rep_n, sub_rc, sub_rv, sub_hits, sub_ntop, sub_top;
    std::vector<int32_t> sub_kind;
    uint64_t n_host_voted = 0;
    if (fused) {
        CID_TRY(pipe->fp.ensure((size_t)ix->N * 8));
        CID_CUDA(cudaMemcpy(pipe->fp.p, out.vote->fp.data(), (size_t)ix->N * 8, cudaMemcpyHostToDevice));
        worker = std::thread([&]() {
            cudaSetDevice(ctx->device);
            for (;;) {
                VoteJob j;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return !jobs.empty() || closing; });
                    if (jobs.empty()) return;
                    j = jobs.front();
                    jobs.pop_front();
                }
                auto& s = pipe->slot[j.slot];
                const double t0 = now_ms();
                const bool evok = cudaEventSynchronize(s.done) == cudaSuccess;
                const double t1 = now_ms();
                t_evwait += t1 - t0;
                if (!evok) worker_rc = CID_E_CUDA;
                else {
                    // the device decided every read it can decide bit-exactly; the rest come back as
                    // [read, n, (colour, count) x n] records and go through the host vote
                    const uint32_t nwords = *s.h_cursor.as<uint32_t>();
                    if (nwords) {
                        if (s.h_list.ensure((size_t)nwords * 4) != CID_OK ||
                            cudaMemcpyAsync(s.h_list.p, s.list.p, (size_t)nwords * 4, cudaMemcpyDeviceToHost, s.st) != cudaSuccess ||
                            cudaStreamSynchronize(s.st) != cudaSuccess)
                            worker_rc = CID_E_CUDA;
                        else {
                            const uint32_t* L = s.h_list.as<uint32_t>();
                            sub_idx.clear(); sub_n_set.clear(); sub_rep_n.clear(); sub_rc.clear(); sub_rv.clear();
                            for (uint32_t at = 0; at + 2 <= nwords;) {
                                const uint32_t i = L[at], n = L[at + 1];
                                if (n > rc || at + 2 + 2 * (size_t)n > nwords || i >= j.nr) { worker_rc = CID_E_CAPACITY; break; }
                                sub_idx.push_back(i);
                                sub_n_set.push_back(out.n_set[j.r0 + i]);
                                sub_rep_n.push_back(n);
                                const size_t base = sub_rc.size();
                                sub_rc.resize(base + rc, 0); sub_rv.resize(base + rc, 0);
                                for (uint32_t e = 0; e < n; e++) { sub_rc[base + e] = L[at + 2 + 2 * e]; sub_rv[base + e] = L[at + 3 + 2 * e]; }
                                at += 2 + 2 * n;
                            }
                            const size_t ns = sub_idx.size();
                            n_host_voted += ns;
                            sub_flags.assign(ns, 0);
                            sub_kind.resize(ns); sub_hits.resize(ns); sub_ntop.resize(ns);
                            sub_top.assign(ns * (size_t)std::max<uint32_t>(out.top_cap, 1), 0);
                            classify_chunk(*out.vote, ns, sub_n_set.data(), sub_flags.data(), sub_rep_n.data(), sub_rc.data(),
                                           sub_rv.data(), (uint32_t)rc, out.threads, sub_kind.data(), sub_hits.data(),
                                           sub_ntop.data(), out.top ? sub_top.data() : nullptr, out.top_cap);
                            for (size_t q = 0; q < ns; q++) {
                                const uint64_t r = j.r0 + sub_idx[q];
                                out.kind[r] = sub_kind[q]; out.hits[r] = sub_hits[q]; out.n_top[r] = sub_ntop[q];
                                if (out.top) memcpy(out.top + r * out.top_cap, sub_top.data() + q * out.top_cap, (size_t)out.top_cap * 4);
                            }
                        }
                    }
                    t_vote += now_ms() - t1;
                }
                {
                    std::lock_guard<std::mutex> lk(mu);
                    slot_busy[j.slot] = false;
                }
                cv.notify_all();
            }
        });
    }
    auto finish = [&](int rcode) {
        if (fused) {
            { std::lock_guard<std::mutex> lk(mu); closing = true; }
            cv.notify_all();
            worker.join();
        }
        for (auto& s : pipe->slot) { cudaStreamSynchronize(s.cst); cudaStreamSynchronize(s.st); }
        if (rcode == CID_OK && worker_rc != CID_OK) { set_error("read_id: host vote worker failed"); return worker_rc; }
        return rcode;
    };
#define PIPE_TRY(expr) do { int _rc = (expr); if (_rc != CID_OK) return finish(_rc); } while (0)
#define PIPE_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); return finish(CID_E_CUDA); } } while (0)

    // Host-side preparation of every chunk -- the scan of its read geometry and the copy of its offset arrays into page-locked
    // staging (pageable cudaMemcpyAsync would drain the stream) -- is ~10 ns per read on one thread: serial, it delays the copy
    // of a large chunk by milliseconds while the GPU idles.  A few helper threads take the chunks in order; the issue loop
    // below only waits for "chunk c is prepared".
    // (A chunk is cut into PREP_PARTS pieces, so that the first chunk -- nothing runs before it is prepared -- and the large
    // chunks of the steep ramp are scanned by all helpers at once: 0.45 -> 0.12 ms for 65,536 reads.)
    enum { PREP_PARTS = 4 };
    struct Prep { Geometry geo[PREP_PARTS]; size_t stage_off = 0; std::atomic<int> ready{0}; };
    std::vector<Prep> prep(nchunks);
    size_t stage_words = 0;
    for (uint64_t c = 0; c < nchunks; c++) {
        const uint64_t nso = read_offs[cuts[c + 1]] - read_offs[cuts[c]] + 1, nro = cuts[c + 1] - cuts[c] + 1;
        prep[c].stage_off = stage_words;
        stage_words += nso + nro + (hp ? nro : 0);
    }
    if (ctx->pinned[3].ensure(stage_words * 8 + 64) != CID_OK) return finish(CID_E_NOMEM);
    uint64_t* const stage = ctx->pinned[3].as<uint64_t>();
    std::atomic<uint64_t> prep_next{0};
    auto prep_work = [&]() {
        for (;;) {
            const uint64_t task = prep_next.fetch_add(1), c = task / PREP_PARTS, part = task % PREP_PARTS;
            if (c >= nchunks) return;
            const uint64_t r0 = cuts[c], r1 = cuts[c + 1], nr = r1 - r0, s0 = read_offs[r0], s1 = read_offs[r1];
            // reads [pa, pb) of the chunk: their geometry, and their share of the offset arrays (the +1 end entries go with the last part)
            const uint64_t pa = nr * part / PREP_PARTS, pb = nr * (part + 1) / PREP_PARTS;
            const bool last = part == PREP_PARTS - 1;
            prep[c].geo[part] = read_geometry(seq_offs, read_offs + r0 + pa, pb - pa, ix->k, pp.downsample);
            uint64_t* ho = stage + prep[c].stage_off;
            const size_t nso = s1 - s0 + 1, nro = nr + 1;
            const uint64_t sa = read_offs[r0 + pa], sb = read_offs[r0 + pb];
            memcpy(ho + (sa - s0), seq_offs + sa, (sb - sa + (last ? 1 : 0)) * 8);
            memcpy(ho + nso + pa, read_offs + r0 + pa, (pb - pa + (last ? 1 : 0)) * 8);
            if (hp) memcpy(ho + nso + nro + pa, hp->word_offs + r0 + pa, (pb - pa + (last ? 1 : 0)) * 8);
            prep[c].ready.fetch_add(1, std::memory_order_release);
        }
    };
    const unsigned n_prep = (unsigned)std::min<uint64_t>(nchunks * PREP_PARTS, std::min(4u, std::max(1u, std::thread::hardware_concurrency())));
    std::vector<std::thread> prep_threads;
    for (unsigned i = 0; i < n_prep; i++) prep_threads.emplace_back(prep_work);
    struct PrepJoin { std::vector<std::thread>& t; ~PrepJoin() { for (auto& x : t) if (x.joinable()) x.join(); } } prep_join{prep_threads};

    // CID_TRACE: per-chunk stage timeline from CUDA events on the slot streams (diagnostics only)
    std::vector<cudaEvent_t> tev;
    cudaEvent_t tev0 = nullptr;
    auto tmark = [&](cudaStream_t stx) {
        if (!trace) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, stx); tev.push_back(e);
    };
    if (trace) { cudaEventCreate(&tev0); cudaEventRecord(tev0, pipe->slot[0].st); }
    for (uint64_t c = 0; c < nchunks; c++) {
        const int si = (int)(c % NS);
        auto& s = pipe->slot[si];
        const uint64_t r0 = cuts[c], r1 = cuts[c + 1], nr = r1 - r0;
        const uint64_t s0 = read_offs[r0], s1 = read_offs[r1];
        const uint64_t b0 = seq_offs[s0], b1 = seq_offs[s1];
        if (fused) {          // the slot's pinned staging must have been consumed
            const double t0 = now_ms();
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return !slot_busy[si]; });
            t_wait += now_ms() - t0;
        }
        // longest read / most k-mer positions of THIS chunk size its kernels (scanned while the GPU runs earlier chunks)
        Geometry geo;
        {
            const double t0 = now_ms();
            while (prep[c].ready.load(std::memory_order_acquire) < PREP_PARTS) std::this_thread::yield();
            geo = prep[c].geo[0];
            for (int q = 1; q < PREP_PARTS; q++) {
                const Geometry& g2 = prep[c].geo[q];
                geo.max_bases = std::max(geo.max_bases, g2.max_bases); geo.max_kmers = std::max(geo.max_kmers, g2.max_kmers);
                geo.fast_bases = std::max(geo.fast_bases, g2.fast_bases); geo.fast_kmers = std::max(geo.fast_kmers, g2.fast_kmers);
            }
            t_geom += now_ms() - t0;
        }
        const uint32_t max_bases = geo.fast_bases, max_kmers = geo.fast_kmers;
        size_t eb, ob, nb;
        readid_scratch_bytes(ix, max_bases, max_kmers, nr, &eb, &ob, &nb);
        const uint64_t w0 = hp ? hp->word_offs[r0] : 0, w1 = hp ? hp->word_offs[r1] : 0;
        if (hp) { PIPE_TRY(s.bases.ensure((w1 - w0) * 4 + 64)); PIPE_TRY(s.quals.ensure((nr + 1) * 8 + 64)); }
        else PIPE_TRY(s.bases.ensure(b1 - b0 + 64));
        if (use_q) PIPE_TRY(s.quals.ensure(b1 - b0 + 64));
        PIPE_TRY(s.seq_offs.ensure((s1 - s0 + 1) * 8));
        PIPE_TRY(s.read_offs.ensure((nr + 1) * 8));
        PIPE_TRY(s.entries.ensure(eb));
        PIPE_TRY(s.order.ensure(ob));
        PIPE_TRY(s.nocc.ensure(nb));
        PIPE_TRY(s.n_set.ensure(nr * 4));
        PIPE_TRY(s.flags.ensure(nr * 4));
        if (want_rep) { PIPE_TRY(s.rep_n.ensure(nr * 4)); PIPE_TRY(s.rep.ensure(nr * rc * 8)); }
        if (out.order_n) PIPE_TRY(s.ord_out.ensure(nr * 4 + nr * (size_t)out.order_cap * 5 + 64));
        if (fused) {
            PIPE_TRY(s.kind.ensure(nr * 4)); PIPE_TRY(s.hits.ensure(nr * 4)); PIPE_TRY(s.n_top.ensure(nr * 4));
            PIPE_TRY(s.top.ensure(nr * (size_t)std::max<uint32_t>(out.top_cap, 1) * 4));
            PIPE_TRY(s.list.ensure(nr * (2 + 2 * rc) * 4)); PIPE_TRY(s.cursor.ensure(16)); PIPE_TRY(s.h_cursor.ensure(16));
        }
        // The caller's offset arrays are usually pageable (a pageable cudaMemcpyAsync first drains the
        // stream, i.e. waits for the big copies queued just before it): stage them in pinned memory.
        // The input copies go to a stream of their own and the kernels wait for them through an event.  Enqueued on the stream
        // that also carries the chunk's kernels, the copy of the reads ran 4-5x slower than alone (116 MB: 10.8 ms against
        // 2.1 ms, the GPU otherwise idle: CID_TRACE timelines of round 2) -- whatever the driver does with a copy that has
        // dependent launches queued behind it, a copy-only stream does not see it.
        if (c >= (uint64_t)NS) PIPE_CUDA(cudaStreamWaitEvent(s.cst, s.kern_done, 0));      // the slot's buffers are free again
        {
            const size_t nso = s1 - s0 + 1, nro = nr + 1;
            uint64_t* ho = stage + prep[c].stage_off;              // staged by the preparation threads
            PIPE_CUDA(cudaMemcpyAsync(s.seq_offs.p, ho, nso * 8, cudaMemcpyHostToDevice, s.cst));
            PIPE_CUDA(cudaMemcpyAsync(s.read_offs.p, ho + nso, nro * 8, cudaMemcpyHostToDevice, s.cst));
            // (packed reads: the word offset of every read, absolute -- the device word pointer is biased to match)
            if (hp) PIPE_CUDA(cudaMemcpyAsync(s.quals.p, ho + nso + nro, nro * 8, cudaMemcpyHostToDevice, s.cst));
        }
        tmark(s.cst);
        if (hp) { if (w1 > w0) PIPE_CUDA(cudaMemcpyAsync(s.bases.p, hp->words + w0, (w1 - w0) * 4, cudaMemcpyHostToDevice, s.cst)); }
        else if (b1 > b0) PIPE_CUDA(cudaMemcpyAsync(s.bases.p, bases + b0, b1 - b0, cudaMemcpyHostToDevice, s.cst));
        if (use_q && b1 > b0) PIPE_CUDA(cudaMemcpyAsync(s.quals.p, quals + b0, b1 - b0, cudaMemcpyHostToDevice, s.cst));
        tmark(s.cst);
        PIPE_CUDA(cudaEventRecord(s.h2d_done, s.cst));
        PIPE_CUDA(cudaStreamWaitEvent(s.st, s.h2d_done, 0));
        // device arrays are indexed by absolute base / sequence / read numbers: bias the chunk buffers
        PackedReads pk{hp ? s.bases.as<uint32_t>() - w0 : nullptr, hp ? s.quals.as<uint64_t>() - r0 : nullptr, hp ? hp->lower : 0u};
        const PackedReads* packed_ptr = hp ? &pk : nullptr;
        const uint8_t* d_bases = s.bases.as<uint8_t>() - b0;
        const uint8_t* d_quals = use_q ? s.quals.as<uint8_t>() - b0 : nullptr;
        const uint64_t* d_seq_offs = s.seq_offs.as<uint64_t>() - s0;
        const uint64_t* d_read_offs = s.read_offs.as<uint64_t>() - r0;
        uint32_t* d_n_set = s.n_set.as<uint32_t>() - r0;
        uint32_t* d_flags = s.flags.as<uint32_t>() - r0;
        uint32_t *d_rep_n = nullptr, *d_rc = nullptr, *d_rv = nullptr;
        if (want_rep) {
            d_rep_n = s.rep_n.as<uint32_t>() - r0;
            d_rc = s.rep.as<uint32_t>() - r0 * rc;
            d_rv = s.rep.as<uint32_t>() + nr * rc - r0 * rc;
        }
        uint32_t* d_on = nullptr; uint8_t* d_os = nullptr; uint32_t* d_op = nullptr;
        if (out.order_n) {
            d_on = s.ord_out.as<uint32_t>() - r0;
            d_op = s.ord_out.as<uint32_t>() + nr - r0 * (size_t)out.order_cap;
            d_os = (uint8_t*)(s.ord_out.as<uint32_t>() + nr + nr * (size_t)out.order_cap) - r0 * (size_t)out.order_cap;
        }
        ReadIdScratch scr{s.entries.as<uint32_t>(), s.order.as<uint16_t>(), s.nocc.as<uint32_t>(), nr, &s.vp};
        PIPE_TRY(big_scratch(ix, s.big, geo.max_bases, geo.max_kmers, scr));
        // Option readid_serialize: this chunk's kernels start only after the previous chunk's have finished (copies still
        // overlap).  Measured on B200: 44.8M pairs/s against 46.3M with free overlap across the slot streams, so it is off.
        if (c > 0 && ctx->opt_readid_serialize) PIPE_CUDA(cudaStreamWaitEvent(s.st, pipe->slot[(c - 1) % NS].kern_done, 0));
        PIPE_TRY(readid_run(ix, s.st, d_bases, d_quals, packed_ptr, d_seq_offs, d_read_offs, r0, nr, max_bases, max_kmers, pp, scr,
                            d_n_set, d_flags, d_rep_n, d_rc, d_rv, out.order_cap, d_on, d_os, d_op));
        tmark(s.st);
        if (fused) {
            PIPE_CUDA(cudaMemsetAsync(s.cursor.p, 0, 4, s.st));
            PIPE_TRY(launch_readid_classify(ctx, s.st, r0, nr, ix->N, (uint32_t)rc, d_n_set, d_flags, d_rep_n, d_rc, d_rv,
                                            pipe->fp.as<double>(), out.vote->fp_correct, out.vote->group_width,
                                            s.kind.as<int32_t>() - r0,
                                            s.hits.as<uint32_t>() - r0, s.n_top.as<uint32_t>() - r0,
                                            s.top.as<uint32_t>() - r0 * (size_t)out.top_cap, out.top_cap, s.list.as<uint32_t>(),
                                            s.cursor.as<uint32_t>()));
            PIPE_CUDA(cudaEventRecord(s.kern_done, s.st));
            PIPE_CUDA(cudaMemcpyAsync(out.kind + r0, s.kind.p, nr * 4, cudaMemcpyDeviceToHost, s.st));
            PIPE_CUDA(cudaMemcpyAsync(out.hits + r0, s.hits.p, nr * 4, cudaMemcpyDeviceToHost, s.st));
            PIPE_CUDA(cudaMemcpyAsync(out.n_top + r0, s.n_top.p, nr * 4, cudaMemcpyDeviceToHost, s.st));
            PIPE_CUDA(cudaMemcpyAsync(out.n_set + r0, s.n_set.p, nr * 4, cudaMemcpyDeviceToHost, s.st));
            if (out.flags) PIPE_CUDA(cudaMemcpyAsync(out.flags + r0, s.flags.p, nr * 4, cudaMemcpyDeviceToHost, s.st));
            if (out.top) PIPE_CUDA(cudaMemcpyAsync(out.top + r0 * (size_t)out.top_cap, s.top.p, nr * (size_t)out.top_cap * 4,
                                                   cudaMemcpyDeviceToHost, s.st));
            PIPE_CUDA(cudaMemcpyAsync(s.h_cursor.p, s.cursor.p, 4, cudaMemcpyDeviceToHost, s.st));
            PIPE_CUDA(cudaEventRecord(s.done, s.st));
            tmark(s.st);
            { std::lock_guard<std::mutex> lk(mu); slot_busy[si] = true; jobs.push_back(VoteJob{r0, nr, si}); }
            cv.notify_all();
        } else {
            PIPE_CUDA(cudaEventRecord(s.kern_done, s.st));
            if (out.n_set) PIPE_CUDA(cudaMemcpyAsync(out.n_set + r0, s.n_set.p, nr * 4, cudaMemcpyDeviceToHost, s.st));
            if (out.flags) PIPE_CUDA(cudaMemcpyAsync(out.flags + r0, s.flags.p, nr * 4, cudaMemcpyDeviceToHost, s.st));
            if (out.rep_n) {
                PIPE_CUDA(cudaMemcpyAsync(out.rep_n + r0, s.rep_n.p, nr * 4, cudaMemcpyDeviceToHost, s.st));
                PIPE_CUDA(cudaMemcpyAsync(out.rep_colour + r0 * rc, s.rep.p, nr * rc * 4, cudaMemcpyDeviceToHost, s.st));
                PIPE_CUDA(cudaMemcpyAsync(out.rep_count + r0 * rc, s.rep.as<uint32_t>() + nr * rc, nr * rc * 4,
                                          cudaMemcpyDeviceToHost, s.st));
            }
            if (out.order_n) {
                const size_t oc = out.order_cap;
                PIPE_CUDA(cudaMemcpyAsync(out.order_n + r0, s.ord_out.p, nr * 4, cudaMemcpyDeviceToHost, s.st));
                PIPE_CUDA(cudaMemcpyAsync(out.order_pos + r0 * oc, s.ord_out.as<uint32_t>() + nr, nr * oc * 4,
                                          cudaMemcpyDeviceToHost, s.st));
                PIPE_CUDA(cudaMemcpyAsync(out.order_seq + r0 * oc, s.ord_out.as<uint32_t>() + nr + nr * oc,
                                          nr * oc, cudaMemcpyDeviceToHost, s.st));
            }
        }
    }
#undef PIPE_TRY
#undef PIPE_CUDA
    const double t_issued = now_ms() - t_start;
    int rcode = finish(CID_OK);
    if (trace)
        fprintf(stderr, "[cid trace] read_id: %llu reads, %llu chunks (ramp to %llu): geometry %.2f ms, issue loop %.2f ms (slot wait %.2f), "
                        "host vote busy %.2f ms for %llu undecided reads, vote waiting on GPU %.2f ms, total %.2f ms\n",
                (unsigned long long)nreads, (unsigned long long)nchunks, (unsigned long long)chunk, t_geom, t_issued, t_wait,
                t_vote, (unsigned long long)n_host_voted, t_evwait, now_ms() - t_start);
    if (trace && fused && rcode == CID_OK) {
        // four marks per chunk: offsets copied | bases+quals copied | kernels done | classify + D2H done
        for (size_t c = 0; c + 3 < tev.size() && c / 4 < nchunks; c += 4) {
            float a = 0, b = 0, k2 = 0, d2 = 0;
            cudaEventElapsedTime(&a, tev0, tev[c]); cudaEventElapsedTime(&b, tev0, tev[c + 1]);
            cudaEventElapsedTime(&k2, tev0, tev[c + 2]); cudaEventElapsedTime(&d2, tev0, tev[c + 3]);
            fprintf(stderr, "[cid trace]   chunk %zu: offs@%.2f h2d@%.2f kernels@%.2f out@%.2f ms\n", c / 4, a, b, k2, d2);
        }
    }
    for (auto e : tev) cudaEventDestroy(e);
    if (tev0) cudaEventDestroy(tev0);
    if (rcode != CID_OK) return rcode;
    return check_err_flags(ctx, ctx->stream);
}

}  // namespace cid

using namespace cid;

extern "C" {

static int read_id_batch_dev_impl(cid_index* ix, const char* d_bases, const char* d_quals, const PackedReads* pkp_in, const uint64_t* d_seq_offs,
                                  const uint64_t* d_read_offs, uint64_t nreads, uint32_t h_max_read_bases, uint32_t h_max_kmers,
                                  const cid_readid_params* p, uint32_t* d_n_set, uint32_t* d_flags, uint32_t* d_rep_n,
                                  uint32_t* d_rep_colour, uint32_t* d_rep_count, void* stream) {
    cid_ctx* ctx = ix->ctx;
    CID_CUDA(cudaSetDevice(ctx->device));
    cid_readid_params pp;
    default_readid_params(pp, p, ix->N);
    cudaStream_t user = (cudaStream_t)stream;
    const int nst = (ctx->opt_readid_streams >= 2 && nreads >= 4 * 32768) ? 2 : 1;
    const uint32_t big_kmers = h_max_kmers ? h_max_kmers : std::max<uint32_t>(h_max_read_bases, 1);
    const PackedReads* pkp = pkp_in;
    if (nst == 1) {
        const uint64_t cap = std::min<uint64_t>(std::max<uint64_t>(nreads, 1), 1u << 20);
        size_t eb, ob, nb;
        readid_scratch_bytes(ix, h_max_read_bases, h_max_kmers, cap, &eb, &ob, &nb);
        CID_TRY(ctx->scratch[16].ensure(eb));
        CID_TRY(ctx->scratch[17].ensure(ob));
        CID_TRY(ctx->scratch[18].ensure(nb));
        ReadIdScratch scr{ctx->scratch[16].as<uint32_t>(), ctx->scratch[17].as<uint16_t>(), ctx->scratch[18].as<uint32_t>(), cap, &ctx->scratch[26]};
        CID_TRY(big_scratch(ix, ctx->scratch[24], h_max_read_bases, big_kmers, scr));
        return readid_run(ix, user, (const uint8_t*)d_bases, (const uint8_t*)d_quals, pkp, d_seq_offs, d_read_offs, 0,
                          nreads, h_max_read_bases, h_max_kmers, pp, scr, d_n_set, d_flags, d_rep_n, d_rep_colour, d_rep_count,
                          0, nullptr, nullptr, nullptr);
    }
    // fork: chunks alternate over two internal streams; join back into the caller's stream
    for (int i = 0; i < 2; i++) {
        if (!ctx->aux[i]) CID_CUDA(cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking));
        if (!ctx->aux_join[i]) CID_CUDA(cudaEventCreateWithFlags(&ctx->aux_join[i], cudaEventDisableTiming));
    }
    if (!ctx->aux_fork) CID_CUDA(cudaEventCreateWithFlags(&ctx->aux_fork, cudaEventDisableTiming));
    const uint64_t chunk = std::min<uint64_t>(131072, std::max<uint64_t>(32768, (nreads + 7) / 8));
    size_t eb, ob, nb;
    readid_scratch_bytes(ix, h_max_read_bases, h_max_kmers, chunk, &eb, &ob, &nb);
    for (int i = 0; i < 2; i++) {
        CID_TRY(ctx->scratch[16 + 3 * i].ensure(eb));
        CID_TRY(ctx->scratch[17 + 3 * i].ensure(ob));
        CID_TRY(ctx->scratch[18 + 3 * i].ensure(nb));
    }
    CID_CUDA(cudaEventRecord(ctx->aux_fork, user));
    for (int i = 0; i < 2; i++) CID_CUDA(cudaStreamWaitEvent(ctx->aux[i], ctx->aux_fork, 0));
    int rc = CID_OK;
    uint64_t c = 0;
    for (uint64_t r0 = 0; r0 < nreads && rc == CID_OK; r0 += chunk, c++) {
        const int i = (int)(c & 1);
        ReadIdScratch scr{ctx->scratch[16 + 3 * i].as<uint32_t>(), ctx->scratch[17 + 3 * i].as<uint16_t>(),
                          ctx->scratch[18 + 3 * i].as<uint32_t>(), chunk, &ctx->scratch[26 + i]};
        rc = big_scratch(ix, ctx->scratch[24 + i], h_max_read_bases, big_kmers, scr);
        if (rc == CID_OK) rc = readid_run(ix, ctx->aux[i], (const uint8_t*)d_bases, (const uint8_t*)d_quals, pkp, d_seq_offs, d_read_offs, r0,
                        std::min(chunk, nreads - r0), h_max_read_bases, h_max_kmers, pp, scr, d_n_set, d_flags, d_rep_n,
                        d_rep_colour, d_rep_count, 0, nullptr, nullptr, nullptr);
    }
    for (int i = 0; i < 2; i++) {       // always join, also after an error, so the caller's stream stays ordered
        cudaEventRecord(ctx->aux_join[i], ctx->aux[i]);
        cudaStreamWaitEvent(user, ctx->aux_join[i], 0);
    }
    return rc;
}

int cid_read_id_batch_dev(cid_index* ix, const char* d_bases, const char* d_quals, const uint64_t* d_seq_offs,
                          uint64_t nseq, uint64_t nbases, const uint64_t* d_read_offs, uint64_t nreads,
                          uint32_t h_max_read_bases, uint32_t h_max_kmers, const cid_readid_params* p,
                          uint32_t* d_n_set, uint32_t* d_flags, uint32_t* d_rep_n, uint32_t* d_rep_colour,
                          uint32_t* d_rep_count, void* stream) {
    (void)nseq; (void)nbases;
    return read_id_batch_dev_impl(ix, d_bases, d_quals, nullptr, d_seq_offs, d_read_offs, nreads, h_max_read_bases, h_max_kmers, p, d_n_set,
                                  d_flags, d_rep_n, d_rep_colour, d_rep_count, stream);
}
int cid_read_id_batch_packed_dev(cid_index* ix, const uint32_t* d_words, const uint64_t* d_word_offs, uint32_t pack_flags,
                                 const uint64_t* d_seq_offs, uint64_t nseq, const uint64_t* d_read_offs, uint64_t nreads,
                                 uint32_t h_max_read_bases, uint32_t h_max_kmers, const cid_readid_params* p, uint32_t* d_n_set,
                                 uint32_t* d_flags, uint32_t* d_rep_n, uint32_t* d_rep_colour, uint32_t* d_rep_count, void* stream) {
    (void)nseq;
    if (!d_words || !d_word_offs) { set_error("cid_read_id_batch_packed_dev: null argument"); return CID_E_INVALID; }
    const PackedReads pk{d_words, d_word_offs, pack_flags & CID_PACK_LOWER};
    return read_id_batch_dev_impl(ix, nullptr, nullptr, &pk, d_seq_offs, d_read_offs, nreads, h_max_read_bases, h_max_kmers, p, d_n_set,
                                  d_flags, d_rep_n, d_rep_colour, d_rep_count, stream);
}

int cid_read_id_batch(cid_index* ix, const char* bases, const char* quals, const uint64_t* seq_offs, uint64_t nseq,
                      const uint64_t* read_offs, uint64_t nreads, const cid_readid_params* p, uint32_t* n_set,
                      uint32_t* flags, uint32_t* rep_n, uint32_t* rep_colour, uint32_t* rep_count) {
    if (!rep_n || !rep_colour || !rep_count) { set_error("read_id: null report buffers"); return CID_E_INVALID; }
    HostOut o;
    o.n_set = n_set; o.flags = flags; o.rep_n = rep_n; o.rep_colour = rep_colour; o.rep_count = rep_count;
    return read_id_pipeline(ix, bases, quals, seq_offs, nseq, read_offs, nreads, p, o);
}

int cid_read_id_classify(cid_index* ix, const char* bases, const char* quals, const uint64_t* seq_offs, uint64_t nseq,
                         const uint64_t* read_offs, uint64_t nreads, const cid_readid_params* p,
                         const uint64_t* n_ref_by_colour, double fp_correct, int32_t* kind, uint32_t* hits,
                         uint32_t* n_set, uint32_t* n_top, uint32_t* top, uint32_t top_cap) {
    if (!ix || !n_ref_by_colour || !kind || !hits || !n_set || !n_top) { set_error("read_id_classify: null argument"); return CID_E_INVALID; }
    cid_readid_params pp;
    default_readid_params(pp, p, ix->N);
    pp.rep_cap = ix->N + 1;      // the vote needs the untruncated report
    VoteParams vp;
    vote_params_init(vp, ix->S, ix->H, ix->N, n_ref_by_colour, fp_correct, pp.group_width);
    HostOut o;
    o.n_set = n_set;
    o.vote = &vp; o.threads = ix->ctx->opt_host_threads;
    o.kind = kind; o.hits = hits; o.n_top = n_top; o.top = top; o.top_cap = top ? top_cap : 0;
    return read_id_pipeline(ix, bases, quals, seq_offs, nseq, read_offs, nreads, &pp, o);
}

int cid_read_id_classify_packed(cid_index* ix, const uint32_t* words, const uint64_t* word_offs, uint32_t pack_flags,
                                const uint64_t* seq_offs, uint64_t nseq, const uint64_t* read_offs, uint64_t nreads,
                                const cid_readid_params* p, const uint64_t* n_ref_by_colour, double fp_correct, int32_t* kind,
                                uint32_t* hits, uint32_t* n_set, uint32_t* n_top, uint32_t* top, uint32_t top_cap) {
    if (!ix || !words || !word_offs || !n_ref_by_colour || !kind || !hits || !n_set || !n_top) { set_error("read_id_classify_packed: null argument"); return CID_E_INVALID; }
    cid_readid_params pp;
    default_readid_params(pp, p, ix->N);
    pp.rep_cap = ix->N + 1;
    pp.qual_offset = 0;          // the packer applied seq.rs:36-56 qual_mask
    VoteParams vp;
    vote_params_init(vp, ix->S, ix->H, ix->N, n_ref_by_colour, fp_correct, pp.group_width);
    HostOut o;
    o.n_set = n_set;
    o.vote = &vp; o.threads = ix->ctx->opt_host_threads;
    o.kind = kind; o.hits = hits; o.n_top = n_top; o.top = top; o.top_cap = top ? top_cap : 0;
    const HostPacked hp{words, word_offs, pack_flags & CID_PACK_LOWER};
    return read_id_pipeline(ix, nullptr, nullptr, seq_offs, nseq, read_offs, nreads, &pp, o, &hp);
}

int cid_read_kmer_order32(cid_index* ix, const char* bases, const uint64_t* seq_offs, uint64_t nseq,
                          const uint64_t* read_offs, uint64_t nreads, const cid_readid_params* p, uint32_t order_cap,
                          uint32_t* order_n, uint8_t* order_seq, uint32_t* order_pos) {
    if (!order_n || !order_seq || !order_pos || order_cap == 0) { set_error("read_kmer_order: null argument"); return CID_E_INVALID; }
    HostOut o;
    o.order_cap = order_cap; o.order_n = order_n; o.order_seq = order_seq; o.order_pos = order_pos;
    return read_id_pipeline(ix, bases, nullptr, seq_offs, nseq, read_offs, nreads, p, o);
}

int cid_read_kmer_order(cid_index* ix, const char* bases, const uint64_t* seq_offs, uint64_t nseq,
                        const uint64_t* read_offs, uint64_t nreads, const cid_readid_params* p, uint32_t order_cap,
                        uint32_t* order_n, uint8_t* order_seq, uint16_t* order_pos) {
    if (!order_pos) { set_error("read_kmer_order: null argument"); return CID_E_INVALID; }
    std::vector<uint32_t> wide((size_t)nreads * order_cap);
    CID_TRY(cid_read_kmer_order32(ix, bases, seq_offs, nseq, read_offs, nreads, p, order_cap, order_n, order_seq, wide.data()));
    for (uint64_t r = 0; r < nreads; r++)
        for (uint32_t i = 0; i < std::min(order_n[r], order_cap); i++) {
            const uint32_t v = wide[r * order_cap + i];
            if (v > 0xFFFFu) { set_error("read_kmer_order: position %u does not fit 16 bits, use cid_read_kmer_order32", v); return CID_E_CAPACITY; }
            order_pos[r * order_cap + i] = (uint16_t)v;
        }
    return CID_OK;
}

}  // extern "C"
