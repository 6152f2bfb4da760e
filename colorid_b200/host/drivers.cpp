// Drivers of the host layer: the reference's build / search / read_id entry points re-expressed over
// the C ABI (include/colorid_b200.h).  File parsing, quality masking of the build/search inputs and
// report formatting stay on the host exactly as north_star prescribes; k-mer extraction, Bloom
// insert, transposition, row gathers and the read vote run on the GPU.  No CPU fallback.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <thread>
#include <zlib.h>      // crc32 of the `_host fastq_parse` digest

#include "../../include/colorid_b200.h"
#include "cid_host.hpp"

namespace cidh {

namespace {
void ck(int rc) { if (rc != CID_OK) throw Error(std::string("colorid_b200: ") + cid_last_error()); }
bool ends_with(const std::string& s, const char* suf) { size_t n = strlen(suf); return s.size() >= n && !s.compare(s.size() - n, n, suf); }
struct Timer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    unsigned long long secs() const { return (unsigned long long)std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - t0).count(); }
};
struct Trace {                    // COLORID_B200_TRACE=1: stage timings on stderr (diagnostics only)
    bool on = getenv("COLORID_B200_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void mark(const char* what) {
        auto t1 = std::chrono::steady_clock::now();
        if (on) fprintf(stderr, "[trace] %-28s %8.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    }
};
// The GPU side of a command: one device (cid_ctx + cid_index), or -- COLORID_B200_DEVICES=0,1,.. -- one index over several
// GPUs (cid_mg).  COLORID_B200_SHARD=columns|replicated picks the layout; by default the matrix is replicated when it fits
// one GPU comfortably and column-sharded otherwise (BASELINE north_star).  The same calls either way.
std::vector<int> env_devices() {
    std::vector<int> v;
    if (const char* ds = getenv("COLORID_B200_DEVICES"))
        for (const char* q = ds; *q;) { char* e; long x = strtol(q, &e, 10); if (e == q) break; v.push_back((int)x); q = *e ? e + 1 : e; }
    return v;
}
struct Gpu {
    cid_ctx* ctx = nullptr;
    cid_index* ix = nullptr;
    cid_mg* mg = nullptr;             // set instead of ctx/ix when several devices are in use
    std::vector<int> devs;
    // creating a CUDA context takes 1.5-2.5 s on a fresh process: do it on a helper thread while the caller
    // reads its input files, join on first use
    std::thread starter;
    int start_rc = CID_OK;
    std::string start_err;
    explicit Gpu(int device, bool multi = false) {
        if (multi) devs = env_devices();
        if (devs.size() > 1) return;                      // cid_mg_create once the index shape (hence the layout) is known
        if (devs.size() == 1) device = devs[0];
        starter = std::thread([this, device] {
            start_rc = cid_ctx_create(device, &ctx);
            if (start_rc != CID_OK) start_err = cid_last_error();       // cid_last_error is thread-local
        });
    }
    void ready() {
        if (starter.joinable()) starter.join();
        if (start_rc != CID_OK) { int rc = start_rc; start_rc = CID_OK; (void)rc; throw Error("colorid_b200: " + start_err); }
    }
    ~Gpu() {
        if (starter.joinable()) starter.join();
        if (mg) cid_mg_destroy(mg);
        if (ix) cid_index_destroy(ix);
        if (ctx) cid_ctx_destroy(ctx);
    }
    void create(uint64_t S, uint64_t H, uint64_t k, uint64_t N) {
        if (H > 0xFFFFFFFFull || k > 0xFFFFFFFFull || N > 0xFFFFFFFFull) throw Error("index parameters out of range");
        // The reference hashes with the crate `xxh3 = "0.1.1"` (Cargo.toml:9), a pre-freeze XXH3 that is not pinned here:
        // COLORID_B200_HASH_VARIANT selects the draft an index was / is to be hashed with (include/colorid_b200.h;
        // default 0 = stable XXH3; tools/pin_from_bxi.py finds the variant that reproduces a given .bxi).
        const char* hv = getenv("COLORID_B200_HASH_VARIANT");
        if (devs.size() > 1) {
            int mode = -1;
            if (const char* sm = getenv("COLORID_B200_SHARD")) mode = !strcmp(sm, "columns") ? CID_MG_COLUMNS : !strcmp(sm, "replicated") ? CID_MG_REPLICATED : -1;
            if (mode < 0) {
                const uint64_t W = (N + 31) / 32, Wp = W <= 2 ? W : (W + 3) / 4 * 4;
                mode = S * Wp * 4 > (48ull << 30) ? CID_MG_COLUMNS : CID_MG_REPLICATED;
            }
            ck(cid_mg_create(devs.data(), (int)devs.size(), mode, &mg));
            fprintf(stderr, "%zu GPUs, index %s\n", devs.size(), mode == CID_MG_COLUMNS ? "column-sharded" : "replicated");
            ck(cid_mg_index_create(mg, S, (uint32_t)H, (uint32_t)k, (uint32_t)N));
            if (hv) ck(cid_mg_index_set_hash_variant(mg, (uint32_t)strtoul(hv, nullptr, 10)));
            return;
        }
        ready();
        ck(cid_index_create(ctx, S, (uint32_t)H, (uint32_t)k, (uint32_t)N, &ix));
        if (hv) ck(cid_index_set_hash_variant(ix, (uint32_t)strtoul(hv, nullptr, 10)));
    }
    void set_minimizer(uint32_t m) { ck(mg ? cid_mg_index_set_minimizer(mg, m) : cid_index_set_minimizer(ix, m)); }
    void set_host_threads(uint64_t t) {
        if (mg) ck(cid_mg_set_option(mg, "host_threads", (int64_t)t));
        else { ready(); ck(cid_ctx_set_option(ctx, "host_threads", (int64_t)t)); }
    }
    void upload(const Bigsi& b) {   // main.rs:576 / :796 read_bigsi -> dense device matrix
        create(b.bloom_size, b.num_hash, b.k_size, b.n_colors());
        if (b.mini) {
            if (b.m_size == 0 || b.m_size > 0xFFFFFFFFull) throw Error("minimizer size out of range");
            set_minimizer((uint32_t)b.m_size);
        }
        uint64_t c = 0;
        for (auto& kv : b.colors) if (kv.first != c++) throw Error("index colours are not 0..N-1");
        ck(mg ? cid_mg_index_upload_rows(mg, b.row_ids.data(), b.words.data(), b.row_ids.size())
              : cid_index_upload_rows(ix, b.row_ids.data(), b.words.data(), b.row_ids.size()));
    }
    // accessions of different lanes may be built concurrently (one lane per GPU that owns columns)
    uint32_t lanes() const { return mg ? (uint32_t)cid_mg_n_shards(mg) : 1u; }
    uint32_t lane_of(uint32_t colour) const { return mg ? (uint32_t)cid_mg_shard_of_colour(mg, colour) : 0u; }
    void build_accession(uint32_t colour, const SeqBatch& sb, int mode, int64_t cutoff, int mini_variant, uint64_t* nref, int64_t* used) {
        if (mg) ck(cid_mg_build_accession(mg, colour, sb.bases.data(), sb.offs.data(), sb.n(), mode, cutoff, mini_variant, nref, used));
        else if (mini_variant >= 0) ck(cid_build_accession_mini(ix, colour, sb.bases.data(), sb.offs.data(), sb.n(), mode, cutoff, mini_variant, nref, used));
        else ck(cid_build_accession(ix, colour, sb.bases.data(), sb.offs.data(), sb.n(), mode, cutoff, nref, used));
    }
    void finalize() { ck(mg ? cid_mg_build_finalize(mg) : cid_build_finalize(ix)); }
    void nonzero_rows(Bigsi& out) {
        uint64_t nrows = 0, got = 0;
        ck(mg ? cid_mg_index_count_nonzero_rows(mg, &nrows) : cid_index_count_nonzero_rows(ix, &nrows));
        out.row_words = (out.n_colors() + 31) / 32;
        out.row_ids.resize(nrows);
        out.words.resize(nrows * out.row_words);
        ck(mg ? cid_mg_index_download_nonzero_rows(mg, out.row_ids.data(), out.words.data(), nrows, &got)
              : cid_index_download_nonzero_rows(ix, out.row_ids.data(), out.words.data(), nrows, &got));
    }
    void query_counts(const SeqBatch& sb, const std::vector<uint64_t>& qoffs, int mode, bool gene, int64_t filter, uint32_t* counts,
                      uint64_t* nk, uint64_t* un, uint64_t* us, uint64_t* um) {
        const uint64_t nq = qoffs.size() - 1;
        ck(mg ? cid_mg_query_counts(mg, sb.bases.data(), sb.offs.data(), sb.n(), qoffs.data(), nq, mode, gene ? 1 : 0, filter, counts, nk, un, us, um, nullptr)
              : cid_query_counts(ix, sb.bases.data(), sb.offs.data(), sb.n(), qoffs.data(), nq, mode, gene ? 1 : 0, filter, counts, nk, un, us, um, nullptr));
    }
    void query_perfect(const SeqBatch& sb, const std::vector<uint64_t>& qoffs, uint32_t* rows, uint8_t* status, uint64_t* nk) {
        const uint64_t nq = qoffs.size() - 1;
        ck(mg ? cid_mg_query_perfect(mg, sb.bases.data(), sb.offs.data(), sb.n(), qoffs.data(), nq, rows, status, nk)
              : cid_query_perfect(ix, sb.bases.data(), sb.offs.data(), sb.n(), qoffs.data(), nq, rows, status, nk));
    }
    void query_perfect_mf(const SeqBatch& sb, uint64_t nq, uint32_t* rows, uint8_t* status, uint64_t* nk) {
        ck(mg ? cid_mg_query_perfect_mf(mg, sb.bases.data(), sb.offs.data(), nq, rows, status, nk)
              : cid_query_perfect_mf(ix, sb.bases.data(), sb.offs.data(), nq, rows, status, nk));
    }
    void read_id_classify(const char* bases, const char* quals, const uint64_t* seq_offs, uint64_t nseq, const uint64_t* read_offs, uint64_t n,
                          const cid_readid_params* p, const uint64_t* n_ref, double fp, int32_t* kind, uint32_t* hits, uint32_t* n_set,
                          uint32_t* n_top, uint32_t* top, uint32_t top_cap) {
        ck(mg ? cid_mg_read_id_classify(mg, bases, quals, seq_offs, nseq, read_offs, n, p, n_ref, fp, kind, hits, n_set, n_top, top, top_cap)
              : cid_read_id_classify(ix, bases, quals, seq_offs, nseq, read_offs, n, p, n_ref, fp, kind, hits, n_set, n_top, top, top_cap));
    }
};
Bigsi read_index(const std::string& path) {     // main.rs:633-635,724-728,796-800: the suffix decides the struct
    // A .bxi does not say how its rows were hashed.  Indexes written by this program with the same variant are fine; one
    // written by a real colorid binary may need another variant -- say so once, unless the user has chosen one.
    static bool told = false;
    if (!told && !getenv("COLORID_B200_HASH_VARIANT")) {
        told = true;
        fprintf(stderr, "note: rows of %s are read as stable XXH3 (hash variant 0); an index written by the original colorid binary "
                        "(xxh3 crate 0.1.1) has not been verified against it -- see tools/pin_from_bxi.py / COLORID_B200_HASH_VARIANT\n", path.c_str());
    }
    return ends_with(path, ".mxi") ? read_bigsi_mini(path) : read_bigsi(path);
}
Bigsi load_index(const std::string& path) {
    Timer t;
    Bigsi b = read_index(path);
    fprintf(stderr, "Index loaded in %llu seconds\n", t.secs());
    return b;
}
}  // namespace

// ------------------------------------------------------------------ build
int build(const BuildOpts& o) {
    Trace tr;
    Gpu g(o.device, /*multi=*/true);
    const auto map = tab_to_map(o.ref_file);
    if (map.empty()) throw Error("reference file lists no accessions");
    g.create(o.bloom, o.hashes, o.k, map.size());
    if (o.minimizer) {            // main.rs:485-533
        printf("Build with minimizers, minimizer size: %llu\n", (unsigned long long)o.minimizer_value);
        if (o.minimizer_value == 0 || o.minimizer_value > 0xFFFFFFFFull) throw Error("minimizer size out of range");
        g.set_minimizer((uint32_t)o.minimizer_value);
    }
    // -t 1 -> build_single_mini (minimizers of the filtered k-mers), else build_multi_mini (counted minimizers)
    const int mini_variant = !o.minimizer ? -1 : (o.threads == 1 ? CID_MINI_OF_KMERS : CID_MINI_COUNTED);
    tr.mark("context + index create");
    Bigsi out;
    out.bloom_size = o.bloom; out.num_hash = o.hashes; out.k_size = o.k;
    out.mini = o.minimizer; out.m_size = o.minimizer ? o.minimizer_value : 0;
    // colours = rank of the accession in byte-wise sorted order (build.rs:102-113)
    struct Acc { std::string name; std::vector<std::string> files; uint32_t colour; };
    std::vector<Acc> accs;
    for (auto& kv : map) accs.push_back(Acc{kv.first, kv.second, (uint32_t)accs.size()});
    struct AccIn { SeqBatch sb; int mode = CID_SEQ_FASTA; };
    auto load = [&](const Acc& a) {                   // the reference's readers, on a helper thread
        AccIn in;
        if (a.files.size() == 2) {
            fastq_masked_pe(a.files[0], a.files[1], o.quality, in.sb);
            in.mode = CID_SEQ_FASTQ;
        } else if (ends_with(a.files[0], "gz")) {
            // build_multi hard-codes quality 15 for single-end reads (build.rs:187); build_single (:70) and both
            // minimizer builders (:322-325, :439) honour -Q
            fastq_masked_se(a.files[0], (o.threads == 1 || o.minimizer) ? o.quality : 15, in.sb);
            in.mode = CID_SEQ_FASTQ;
        } else {
            for (auto& s : read_fasta(a.files[0])) in.sb.add(s);
        }
        return in;
    };
    // One lane per GPU that owns accession columns (one lane on a single GPU).  In each lane the next accession's files are
    // read and parsed while the GPU builds the current one (the reference reads and builds serially, build.rs:47-99).
    const uint32_t lanes = g.lanes();
    std::vector<std::vector<size_t>> todo(lanes);
    for (size_t i = 0; i < accs.size(); i++) todo[std::min(g.lane_of(accs[i].colour), lanes - 1)].push_back(i);
    std::mutex out_mu;
    std::atomic<size_t> counter{1};
    std::atomic<long long> us_read{0}, us_gpu{0};
    std::vector<std::exception_ptr> errs(lanes);
    auto lane = [&](uint32_t l) {
        try {
            const auto& list = todo[l];
            if (list.empty()) return;
            AccIn next_in;
            std::exception_ptr next_err;
            std::thread reader([&] { try { next_in = load(accs[list[0]]); } catch (...) { next_err = std::current_exception(); } });
            for (size_t j = 0; j < list.size(); j++) {
                const auto ta = std::chrono::steady_clock::now();
                reader.join();
                if (next_err) std::rethrow_exception(next_err);
                AccIn cur = std::move(next_in);
                if (j + 1 < list.size())
                    reader = std::thread([&, j] { try { next_in = load(accs[list[j + 1]]); } catch (...) { next_err = std::current_exception(); } });
                const Acc& a = accs[list[j]];
                fprintf(stderr, "Adding %s to index (%zu/%zu)\n", a.name.c_str(), counter.fetch_add(1), accs.size());
                uint64_t nref = 0;
                int64_t used = 0;
                const auto tb = std::chrono::steady_clock::now();
                // -1: FASTQ -> auto_cutoff (build.rs:56-58), FASTA -> keep everything (:86-87)
                try { g.build_accession(a.colour, cur.sb, cur.mode, o.filter, mini_variant, &nref, &used); }
                catch (...) { if (reader.joinable()) reader.join(); throw; }
                us_read += std::chrono::duration_cast<std::chrono::microseconds>(tb - ta).count();
                us_gpu += std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - tb).count();
                std::lock_guard<std::mutex> lk(out_mu);
                out.colors[a.colour] = a.name;
                // build_single_mini records n_ref_kmers for FASTA accessions only (build.rs:450): FASTQ accessions have no entry
                if (!(o.minimizer && mini_variant == CID_MINI_OF_KMERS && cur.mode == CID_SEQ_FASTQ)) out.n_ref_kmers[a.name] = nref;
            }
        } catch (...) { errs[l] = std::current_exception(); }
    };
    if (lanes == 1) lane(0);
    else {
        std::vector<std::thread> th;
        for (uint32_t l = 0; l < lanes; l++) th.emplace_back(lane, l);
        for (auto& t : th) t.join();
    }
    for (auto& e : errs) if (e) std::rethrow_exception(e);
    if (tr.on) fprintf(stderr, "[trace] waiting for input %.3f s, cid_build_accession %.3f s (summed over %u lanes)\n", us_read / 1e6, us_gpu / 1e6, lanes);
    tr.mark("accessions");
    g.finalize();
    tr.mark("finalize (transpose)");
    printf("Saving BIGSI to file.\n");
    g.nonzero_rows(out);
    tr.mark("download non-zero rows");
    if (o.minimizer) save_bigsi_mini(o.prefix + ".mxi", out);
    else save_bigsi(o.prefix + ".bxi", out);
    tr.mark("save_bigsi");
    return 0;
}

// ------------------------------------------------------------------ search
namespace {
void perfect_lines(const Bigsi& b, const std::string& name, const uint32_t* and_row, uint64_t n_kmers) {
    uint64_t hits = 0;
    for (uint32_t c = 0; c < b.n_colors(); c++) hits += (and_row[c >> 5] >> (c & 31)) & 1u;
    fprintf(stderr, "%llu hits\n", (unsigned long long)hits);
    for (uint32_t c = 0; c < b.n_colors(); c++)
        if ((and_row[c >> 5] >> (c & 31)) & 1u)
            printf("%s\t%s\t%llu\t1.00\n", name.c_str(), b.colors.at(c).c_str(), (unsigned long long)n_kmers);
}

// perfect_search.rs:6-60: one query per FASTA file
void perfect_batch_search(Gpu& g, const Bigsi& b, const std::vector<std::string>& files) {
    SeqBatch sb;
    std::vector<uint64_t> qoffs{0};
    for (auto& f : files) {
        fprintf(stderr, "Counting k-mers, this may take a while!\n");
        for (auto& s : read_fasta(f)) sb.add(s);
        qoffs.push_back(sb.n());
    }
    const uint64_t nq = files.size();
    const uint32_t W = b.row_words;
    std::vector<uint32_t> rows(nq * W);
    std::vector<uint8_t> status(nq);
    std::vector<uint64_t> nk(nq);
    g.query_perfect(sb, qoffs, rows.data(), status.data(), nk.data());
    for (uint64_t q = 0; q < nq; q++) {
        fprintf(stderr, "%llu kmers in query\n", (unsigned long long)nk[q]);
        if (status[q] == 2) fprintf(stderr, "Warning! no kmers in query; maybe your kmer length is larger than your query length?\n");
        else if (status[q] == 1) fprintf(stderr, "No perfect hits!\n");
        else perfect_lines(b, files[q], rows.data() + q * W, nk[q]);
    }
}

// perfect_search.rs:62-120: one query per FASTA record, labelled by its header
void perfect_batch_search_mf(Gpu& g, const Bigsi& b, const std::vector<std::string>& files) {
    for (auto& f : files) {
        std::vector<std::string> labels, seqs;
        read_fasta_mf(f, labels, seqs);
        if (seqs.size() < labels.size()) throw Error("multi-fasta: a header without a sequence (the reference panics: index out of bounds)");
        SeqBatch sb;
        for (size_t i = 0; i < labels.size(); i++) sb.add(seqs[i]);
        const uint64_t nq = labels.size();
        const uint32_t W = b.row_words;
        std::vector<uint32_t> rows(nq * W);
        std::vector<uint8_t> status(nq);
        std::vector<uint64_t> nk(nq);
        if (nq) g.query_perfect_mf(sb, nq, rows.data(), status.data(), nk.data());
        for (uint64_t q = 0; q < nq; q++) {
            if (status[q] == 2) {
                printf("Warning! no kmers in query '%s'; maybe your kmer length is larger than your query length?\n", labels[q].c_str());
                continue;
            }
            fprintf(stderr, "%llu kmers in query\n", (unsigned long long)nk[q]);
            if (status[q] == 1) fprintf(stderr, "No perfect hits!\n");
            else perfect_lines(b, labels[q], rows.data() + q * W, nk[q]);
        }
    }
}

// batch_search_pe.rs:9-179: one query per file (pair)
void batch_search(Gpu& g, const Bigsi& b, const SearchOpts& o) {
    const uint32_t N = b.n_colors();
    size_t i = 0;
    while (i < o.files1.size()) {
        SeqBatch sb;
        std::vector<uint64_t> qoffs{0};
        std::vector<std::string> names;
        int mode;
        if (ends_with(o.files1[i], "gz")) {           // FASTQ query: one call per sample
            const std::string& f1 = o.files1[i];
            if (o.files2.empty()) {
                fprintf(stderr, "%s\nCounting k-mers, this may take a while!\n", f1.c_str());
                fastq_masked_se(f1, o.quality, sb);
            } else {
                if (i >= o.files2.size()) throw Error("fewer reverse files than forward files");
                fprintf(stderr, "Paired end: %s %s\nCounting k-mers, this may take a while!\n", f1.c_str(), o.files2[i].c_str());
                fastq_masked_pe(f1, o.files2[i], o.quality, sb);
            }
            qoffs.push_back(sb.n());
            names.push_back(f1);
            mode = CID_SEQ_FASTQ;
            i++;
        } else {                                       // a run of FASTA queries goes to the GPU as one batch
            while (i < o.files1.size() && !ends_with(o.files1[i], "gz")) {
                fprintf(stderr, "%s\nCounting k-mers, this may take a while!\n", o.files1[i].c_str());
                for (auto& s : read_fasta(o.files1[i])) sb.add(s);
                qoffs.push_back(sb.n());
                names.push_back(o.files1[i]);
                if (!o.gene_search && o.filter < 0) fprintf(stderr, "no gene search\n");
                i++;
            }
            mode = CID_SEQ_FASTA;
        }
        const uint64_t nq = names.size();
        std::vector<uint32_t> counts(nq * N);
        std::vector<uint64_t> nk(nq), un, us, um;
        if (!o.gene_search) { un.resize(nq * N); us.resize(nq * N); um.resize(nq * N); }
        g.query_counts(sb, qoffs, mode, o.gene_search, o.filter, counts.data(), nk.data(), o.gene_search ? nullptr : un.data(),
                       o.gene_search ? nullptr : us.data(), o.gene_search ? nullptr : um.data());
        for (uint64_t q = 0; q < nq; q++) {
            fprintf(stderr, "%llu k-mers in query\n", (unsigned long long)nk[q]);
            if (!o.gene_search)
                generate_report(stdout, names[q], b, counts.data() + q * N, un.data() + q * N, us.data() + q * N, um.data() + q * N,
                                nk[q], o.cov);
            else
                generate_report_gene(stdout, names[q], b, counts.data() + q * N, nk[q], o.cov);
        }
    }
}
}  // namespace

int search(const SearchOpts& o) {
    if (ends_with(o.bigsi, ".mxi")) {
        fprintf(stderr, "Error: An index with minimizers (.mxi) is used, but not available for this function\n");
        return 0;
    }
    fprintf(stderr, "Loading index\n");
    Trace tr;
    Gpu g(o.device, /*multi=*/true);
    const Bigsi b = load_index(o.bigsi);
    tr.mark("read_bigsi");
    g.upload(b);
    tr.mark("context + index upload");
    if (o.perfect_search) {
        if (o.multi_fasta) perfect_batch_search_mf(g, b, o.files1);
        else perfect_batch_search(g, b, o.files1);
    } else {
        batch_search(g, b, o);
    }
    tr.mark("queries");
    return 0;
}

// ------------------------------------------------------------------ read_id
namespace {
// The reference flushes every `-c` records (default 50,000; read_id_mt_pe.rs:762); batching never changes
// the output, so the GPU gets larger batches.
const uint64_t kGpuBatchReads = 1u << 17;
const uint64_t kGpuBatchBytes = 96ull << 20;

struct PinnedBytes {               // growable page-locked byte buffer (cid_host_alloc)
    char* p = nullptr;
    size_t size = 0, cap = 0;
    bool pageable = false;         // plain memory: `_host fastq_parse`, which runs without a GPU
    ~PinnedBytes() { release(p); }
    void release(char* q) { if (pageable) free(q); else cid_host_free(q); }
    void reserve(size_t want) {
        if (want <= cap) return;
        size_t ncap = std::max(want, cap + cap / 2);
        void* np = nullptr;
        if (pageable) { np = malloc(ncap); if (!np) throw Error("out of memory"); }
        else ck(cid_host_alloc(ncap, &np));
        if (size) memcpy(np, p, size);
        release(p);
        p = (char*)np; cap = ncap;
    }
    void append(const char* s, size_t n) { reserve(size + n); memcpy(p + size, s, n); size += n; }
    void fill(char c, size_t n) { reserve(size + n); memset(p + size, c, n); size += n; }
};
struct ReadBatch {
    std::string id_data;                       // read ids back to back (no string per read on the parsing thread)
    std::vector<uint32_t> id_off{0};
    PinnedBytes bases, quals;
    std::vector<uint64_t> seq_offs{0}, read_offs{0};
    bool any_qual = false;
    void add_mate(std::string_view seq, const std::string_view* qual, uint8_t off) {
        if (!bases.cap) { bases.reserve(kGpuBatchBytes + (8u << 20)); quals.reserve(kGpuBatchBytes + (8u << 20)); }   // page-locking is slow: once
        if (qual && off) {
            any_qual = true;
            if (qual->size() == seq.size()) { bases.append(seq.data(), seq.size()); quals.append(qual->data(), qual->size()); }   // masked on the device
            else { const std::string m = qual_mask(std::string(seq), std::string(*qual), off); bases.append(m.data(), m.size()); quals.fill('~', m.size()); }
        } else {
            bases.append(seq.data(), seq.size());
            quals.fill('~', seq.size());
        }
        seq_offs.push_back(bases.size);
    }
    void end_read(std::string_view id) { id_data.append(id.data(), id.size()); id_off.push_back((uint32_t)id_data.size()); read_offs.push_back(seq_offs.size() - 1); }
    uint64_t n() const { return id_off.size() - 1; }
    std::string id(uint64_t r) const { return id_data.substr(id_off[r], id_off[r + 1] - id_off[r]); }
    void clear() { id_data.clear(); id_off.assign(1, 0); bases.size = 0; quals.size = 0; seq_offs.assign(1, 0); read_offs.assign(1, 0); any_qual = false; }
};

// The record loop of per_read_stream_pe / per_read_stream_se (read_id_mt_pe.rs:701-832, :835-951) over one or two readers:
// line 1 of a record is the id (of file 1), line 2 the sequences, line 4 the qualities; paired input stops where file 2 does.
// Whole stretches of records are taken from the readers' line blocks at once: the lengths are in the line tables, so this
// thread lays out the offsets and the ids, and the bases and qualities -- 600 MB per million pairs, what bounded read_id of
// .fastq.gz once the inflating was parallel -- are copied by kCopyThreads threads.  A record whose quality line is not as long
// as its sequence (qual_mask on the host), the last lines of a file and COLORID_B200_PARSE_SLOW=1 go line by line.
const size_t kCopyThreads = 4;
// stretches / batches of fewer records than this stay on one thread; COLORID_B200_PAR_MIN overrides it (the tests set 1 so
// that small inputs take the threaded paths)
size_t par_min(size_t dflt) {
    static const long v = [] { const char* e = getenv("COLORID_B200_PAR_MIN"); return e && *e ? atol(e) : -1L; }();
    return v >= 0 ? (size_t)v : dflt;
}
size_t take_records(const AsyncLineReader::LineBlock* b, const size_t* at, int nfiles, size_t nrec, uint8_t quality, ReadBatch& rb) {
    if (rb.n() >= kGpuBatchReads) return 0;
    nrec = std::min<size_t>(nrec, kGpuBatchReads - rb.n());
    const bool use_q = quality != 0;
    const size_t seq0 = rb.seq_offs.size() - 1;           // index of the first new sequence
    size_t bytes = rb.bases.size, k = 0;
    for (; k < nrec; k++) {
        if (bytes >= kGpuBatchBytes) break;                // the batch is full (flushed by the caller)
        size_t ls[2];
        bool same = true;
        for (int f = 0; f < nfiles; f++) {
            const size_t i = at[f] + 4 * k;
            ls[f] = b[f].end[i + 1] - b[f].begin[i + 1];
            if (use_q && (size_t)(b[f].end[i + 3] - b[f].begin[i + 3]) != ls[f]) same = false;
        }
        if (!same) break;
        for (int f = 0; f < nfiles; f++) { bytes += ls[f]; rb.seq_offs.push_back(bytes); }
        const size_t i0 = at[0] + 4 * k;
        rb.id_data.append(b[0].data + b[0].begin[i0], b[0].end[i0] - b[0].begin[i0]);
        rb.id_off.push_back((uint32_t)rb.id_data.size());
        rb.read_offs.push_back(rb.seq_offs.size() - 1);
    }
    if (k == 0) return 0;
    if (!rb.bases.cap) { rb.bases.reserve(kGpuBatchBytes + (8u << 20)); rb.quals.reserve(kGpuBatchBytes + (8u << 20)); }
    rb.bases.reserve(bytes); rb.quals.reserve(bytes);
    if (use_q) rb.any_qual = true;
    char* const bases = rb.bases.p;
    char* const quals = rb.quals.p;
    const uint64_t* const so = rb.seq_offs.data();
    auto copy = [&](size_t r0, size_t r1) {
        for (size_t r = r0; r < r1; r++)
            for (int f = 0; f < nfiles; f++) {
                const size_t m = seq0 + r * (size_t)nfiles + (size_t)f, d = so[m], n = so[m + 1] - d, i = at[f] + 4 * r;
                memcpy(bases + d, b[f].data + b[f].begin[i + 1], n);
                if (use_q) memcpy(quals + d, b[f].data + b[f].begin[i + 3], n);
                else memset(quals + d, '~', n);
            }
    };
    const size_t nt = k >= par_min(4096) ? kCopyThreads : 1;
    std::vector<std::thread> th;
    size_t spawned = 0;
    try {
        for (size_t t = 1; t < nt; t++) { th.emplace_back(copy, k * t / nt, k * (t + 1) / nt); spawned = t; }
    } catch (...) {}                                          // no more threads to be had: their shares are copied here
    copy(0, k / nt);
    if (spawned + 1 < nt) copy(k * (spawned + 1) / nt, k);
    for (auto& t : th) t.join();
    rb.bases.size = rb.quals.size = bytes;
    return k;
}

template <class Cur, class Flush>
void parse_fastq_records(AsyncLineReader** rd, int nfiles, uint8_t quality, Cur&& cur_batch, Flush&& maybe_flush) {
    static const bool fast = [] { const char* e = getenv("COLORID_B200_PARSE_SLOW"); return !(e && *e && *e != '0'); }();
    std::string_view l[2], id, s[2];
    uint64_t line_count = 1;
    unsigned slow_lines = 0;                 // lines still to be taken one by one (a record the block path declined)
    for (;;) {
        if (fast && line_count % 4 == 1 && !slow_lines) {
            AsyncLineReader::LineBlock b[2];
            size_t at[2], nrec = SIZE_MAX;
            bool ok = true;
            for (int f = 0; f < nfiles && ok; f++) {
                ok = rd[f]->peek_block(b[f], at[f]);
                if (ok) nrec = std::min(nrec, (b[f].n - at[f]) / 4);
            }
            if (ok && nrec > 0) {
                const size_t done = take_records(b, at, nfiles, nrec, quality, cur_batch());
                if (done) {
                    for (int f = 0; f < nfiles; f++) rd[f]->skip_lines(4 * done);
                    line_count += 4 * done;
                    maybe_flush();
                    continue;
                }
            }
            slow_lines = 4;                  // the end of a file, fewer than four lines left in a block, or a record for qual_mask
        }
        if (!rd[0]->next_view(l[0])) break;
        const bool have2 = nfiles < 2 || rd[1]->next_view(l[1]);
        if (line_count % 4 == 1) id = l[0];
        else if (line_count % 4 == 2) { if (!have2) break; s[0] = l[0]; s[1] = l[1]; }
        else if (line_count % 4 == 0) {
            if (!have2) break;
            ReadBatch& rb = cur_batch();
            for (int f = 0; f < nfiles; f++) rb.add_mate(s[f], &l[f], quality);
            rb.end_read(id);
            maybe_flush();
        }
        line_count++;
        if (slow_lines) slow_lines--;
    }
}

struct ReadIdRun {
    Gpu& g; const Bigsi& b; const ReadIdOpts& o;
    FILE* out;
    std::vector<uint64_t> n_ref;
    std::vector<std::string> names;                      // accession by colour
    enum { kFormatThreads = 4 };
    std::vector<std::string> outbufs;
    double fp_correct;
    uint64_t threads = 0;                                // -t: host threads of the vote, applied before the first batch
    uint64_t read_count = 0;
    double t_gpu = 0, t_out = 0;
    std::vector<std::string> count_keys;                 // first-appearance order of the counts-file keys
    std::vector<uint64_t> slot_counts;                   // by position in count_keys
    int slot_too_short = -1, slot_no_hits = -1, slot_reject = -1;
    std::vector<int> slot_color;                         // by colour; -1 = not seen yet
    ReadIdRun(Gpu& g_, const Bigsi& b_, const ReadIdOpts& o_) : g(g_), b(b_), o(o_) {
        out = fopen((o.prefix + "_reads.txt").c_str(), "wb");
        if (!out) throw Error("could not create outfile!");
        for (auto& kv : b.colors) {
            auto it = b.n_ref_kmers.find(kv.second);
            if (it == b.n_ref_kmers.end()) throw Error("index has no k-mer count for accession " + kv.second);
            n_ref.push_back(it->second);
            names.push_back(kv.second);
        }
        slot_color.assign(names.size(), -1);
        fp_correct = std::pow(10.0, -o.correct);          // main.rs:711
        if (o.threads) threads = o.threads;
    }
    ~ReadIdRun() { if (out) fclose(out); }
    // one more read under `key` (an accession, too_short, no_hits or reject); `slot` caches where that key is counted
    void count_slot(int& slot, const std::string& key) {
        if (slot < 0) {
            const auto it = std::find(count_keys.begin(), count_keys.end(), key);      // (an accession may be called "reject")
            if (it == count_keys.end()) { count_keys.push_back(key); slot_counts.push_back(0); slot = (int)count_keys.size() - 1; }
            else slot = (int)(it - count_keys.begin());
        }
        slot_counts[(size_t)slot]++;
    }
    // read_id_mt_pe.rs:282-363 parallel_vec for one batch, then the :779-788 output lines in input order
    void flush(ReadBatch& rb) {
        const uint64_t n = rb.n();
        if (n == 0) return;
        const uint32_t N = b.n_colors();
        cid_readid_params p;
        p.downsample = (uint32_t)o.down_sample; p.start_sample = (uint32_t)o.bitvector_sample; p.qual_offset = o.quality;
        p.group_width = 16; p.reserve_before_find = 1; p.rep_cap = 0;
        uint32_t top_cap = N < 64 ? N : 64;
        std::vector<int32_t> kind(n);
        std::vector<uint32_t> hits(n), n_set(n), n_top(n), top((size_t)n * top_cap);
        const auto tg0 = std::chrono::steady_clock::now();
        if (threads) { g.set_host_threads(threads); threads = 0; }
        g.read_id_classify(rb.bases.p, rb.any_qual ? rb.quals.p : nullptr, rb.seq_offs.data(), rb.seq_offs.size() - 1,
                           rb.read_offs.data(), n, &p, n_ref.data(), fp_correct, kind.data(), hits.data(), n_set.data(), n_top.data(),
                           top.data(), top_cap);
        const auto tg1 = std::chrono::steady_clock::now();
        t_gpu += std::chrono::duration<double>(tg1 - tg0).count();
        // ---- tally (first appearance fixes the order of the counts file: one pass in read order, no strings) and the rare
        // read with more tied accessions than the batch call kept, which is classified again on its own
        std::map<uint64_t, std::string> wide_multi;
        for (uint64_t r = 0; r < n; r++) {
            switch (kind[r]) {
                case CID_CLS_TOO_SHORT: count_slot(slot_too_short, "too_short"); break;
                case CID_CLS_NO_HITS: count_slot(slot_no_hits, "no_hits"); break;
                case CID_CLS_NO_SIGNIFICANT: count_slot(slot_reject, "reject"); break;
                case CID_CLS_ACCEPT: { const uint32_t c = top[r * top_cap]; count_slot(slot_color.at(c), names[c]); break; }
                case CID_CLS_REJECT_MULTI:
                    count_slot(slot_reject, "reject");
                    if (n_top[r] > top_cap) {
                        std::vector<uint32_t> all(N);
                        int32_t k1; uint32_t h1, s1, t1;
                        const uint64_t s_lo = rb.read_offs[r], s_hi = rb.read_offs[r + 1];
                        std::vector<uint64_t> so, ro{0, s_hi - s_lo};
                        for (uint64_t q = s_lo; q <= s_hi; q++) so.push_back(rb.seq_offs[q] - rb.seq_offs[s_lo]);
                        g.read_id_classify(rb.bases.p + rb.seq_offs[s_lo], rb.any_qual ? rb.quals.p + rb.seq_offs[s_lo] : nullptr, so.data(),
                                           s_hi - s_lo, ro.data(), 1, &p, n_ref.data(), fp_correct, &k1, &h1, &s1, &t1, all.data(), N);
                        std::string& multi = wide_multi[r];
                        for (uint32_t j = 0; j < n_top[r]; j++) { if (j) multi += ','; multi += names.at(all[j]); }
                    }
                    break;
                default:
                    throw Error("read " + rb.id(r) + ": a later mate is shorter than k-1 (the reference panics in kmerize_vector_skip_n_set)");
            }
        }
        // ---- the read_id_mt_pe.rs:779-788 lines, formatted in kFormatThreads stretches of reads and written in input order
        auto format = [&](uint64_t r0, uint64_t r1, std::string& ob) {
            auto put_u = [&](uint32_t v) { char tmp[12]; int k = 0; do { tmp[k++] = (char)('0' + v % 10); v /= 10; } while (v); while (k) ob += tmp[--k]; };
            ob.clear();
            ob.reserve((size_t)(r1 - r0) * 48);
            std::string multi;
            for (uint64_t r = r0; r < r1; r++) {
                const std::string* cls = nullptr;
                static const std::string kTooShort = "too_short", kNoHits = "no_hits", kNoSig = "no_significant_hits";
                bool accept = true;
                uint32_t h = 0, ns = n_set[r], nt = 0;
                switch (kind[r]) {
                    case CID_CLS_TOO_SHORT: cls = &kTooShort; ns = 0; break;
                    case CID_CLS_NO_HITS: cls = &kNoHits; break;
                    case CID_CLS_NO_SIGNIFICANT: cls = &kNoSig; accept = false; break;
                    case CID_CLS_ACCEPT: cls = &names[top[r * top_cap]]; h = hits[r]; nt = 1; break;
                    default: {                               // CID_CLS_REJECT_MULTI (anything else was refused above)
                        const auto w = wide_multi.find(r);
                        if (w != wide_multi.end()) cls = &w->second;
                        else {
                            const uint32_t* t = top.data() + r * top_cap;
                            multi.clear();
                            for (uint32_t j = 0; j < n_top[r]; j++) { if (j) multi += ','; multi += names[t[j]]; }
                            cls = &multi;
                        }
                        h = hits[r]; nt = n_top[r]; accept = false;
                        break;
                    }
                }
                ob.append(rb.id_data, rb.id_off[r], rb.id_off[r + 1] - rb.id_off[r]); ob += '\t'; ob += *cls; ob += '\t'; put_u(h); ob += '\t'; put_u(ns);
                ob += accept ? "\taccept\t" : "\treject\t"; put_u(nt); ob += '\n';
            }
        };
        const size_t nt_fmt = n >= par_min(8192) ? (size_t)kFormatThreads : 1;
        outbufs.resize(kFormatThreads);
        {
            std::vector<std::thread> th;
            size_t spawned = 0;
            try {
                for (size_t t = 1; t < nt_fmt; t++) { th.emplace_back(format, n * t / nt_fmt, n * (t + 1) / nt_fmt, std::ref(outbufs[t])); spawned = t; }
            } catch (...) {}
            format(0, n / nt_fmt, outbufs[0]);
            for (auto& t : th) t.join();
            for (size_t t = spawned + 1; t < nt_fmt; t++) format(n * t / nt_fmt, n * (t + 1) / nt_fmt, outbufs[t]);      // (threads that could not be started)
        }
        for (size_t t = 0; t < nt_fmt; t++) fwrite(outbufs[t].data(), 1, outbufs[t].size(), out);
        read_count += n;
        rb.clear();
        t_out += std::chrono::duration<double>(std::chrono::steady_clock::now() - tg1).count();
    }
    void finish() {
        fclose(out); out = nullptr;
        std::map<std::string, uint64_t> counts;
        for (size_t i = 0; i < count_keys.size(); i++) counts[count_keys[i]] = slot_counts[i];
        write_counts_five_fields(o.prefix + "_counts.txt", count_keys, counts);   // main.rs:865
    }
};
}  // namespace

namespace {
// per_read_stream_pe / per_read_stream_se / stream_fasta (read_id_mt_pe.rs:440-951) for one sample, followed by
// reports::read_counts_five_fields (main.rs:865): o.query in, o.prefix_reads.txt + o.prefix_counts.txt out
void read_id_sample(Gpu& g, const Bigsi& b, const ReadIdOpts& o, Trace& tr, ReadBatch* batches) {
    if (o.query.empty()) throw Error("no query files");
    ReadIdRun run(g, b, o);
    // two batches (owned by the caller: their page-locked buffers take ~0.4 s to allocate and are reused across the
    // samples of batch_id): while the worker thread runs the GPU call and writes the lines of one, the parser fills the other
    batches[0].clear(); batches[1].clear();
    int cur = 0;
    std::thread worker;
    std::exception_ptr worker_err;
    auto join_worker = [&]() {
        if (worker.joinable()) worker.join();
        if (worker_err) { std::exception_ptr e = worker_err; worker_err = nullptr; std::rethrow_exception(e); }
    };
    auto submit = [&]() {
        join_worker();
        ReadBatch* full = &batches[cur];
        worker = std::thread([&run, &worker_err, full] {
            try { run.flush(*full); } catch (...) { worker_err = std::current_exception(); }
        });
        cur ^= 1;
    };
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{worker};
#define rb batches[cur]
    Timer t;
    auto maybe_flush = [&]() { if (rb.n() >= kGpuBatchReads || rb.bases.size >= kGpuBatchBytes) submit(); };
    if (ends_with(o.query[0], ".gz")) {
        auto cur_batch = [&]() -> ReadBatch& { return batches[cur]; };
        if (o.query.size() > 1) {                       // per_read_stream_pe, read_id_mt_pe.rs:701-832
            AsyncLineReader a(o.query[0]), c(o.query[1]);
            AsyncLineReader* rd[2] = {&a, &c};
            parse_fastq_records(rd, 2, o.quality, cur_batch, maybe_flush);
            submit(); join_worker();
            fprintf(stderr, "Classified %llu read pairs in %llu seconds\n", (unsigned long long)run.read_count, t.secs());
        } else {                                        // per_read_stream_se, :835-951
            AsyncLineReader a(o.query[0]);
            AsyncLineReader* rd[2] = {&a, nullptr};
            parse_fastq_records(rd, 1, o.quality, cur_batch, maybe_flush);
            submit(); join_worker();
            fprintf(stderr, "Classified %llu reads in %llu seconds\n", (unsigned long long)run.read_count, t.secs());
        }
    } else {                                            // stream_fasta, :440-570 (sequence keeps its line feeds)
        AsyncLineReader a(o.query[0], /*keep_eol=*/true);
        std::string l, id, sub;
        uint64_t count = 0;
        while (a.next(l)) {
            if (count == 0) id = l.substr(0, l.size() - 1);
            else if (l.find('>') != std::string::npos) {
                if (!sub.empty()) {
                    rb.add_mate(sub, nullptr, 0); rb.end_read(id); maybe_flush();
                    id = l.substr(0, l.size() - 1);
                    sub.clear();
                }
            } else sub += l;
            count++;
        }
        rb.add_mate(sub, nullptr, 0); rb.end_read(id);
        submit(); join_worker();
        fprintf(stderr, "Classified %llu reads in %llu seconds\n", (unsigned long long)run.read_count, t.secs());
    }
#undef rb
    tr.mark("parse + classify + write");
    if (tr.on) fprintf(stderr, "[trace]   of which GPU calls %.3f s, formatting+writing %.3f s\n", run.t_gpu, run.t_out);
    run.finish();
    tr.mark("counts file");
}
}  // namespace

// `_host fastq_parse`: the record loop of read_id over one or two FASTQ(.gz) files, every batch digested instead of classified
// (no GPU needed): what the tests compare between the block-wise and the line-by-line path.
int fastq_parse_digest(const std::vector<std::string>& files, uint8_t quality) {
    if (files.empty() || files.size() > 2) throw Error("one or two files");
    ReadBatch rb;
    rb.bases.pageable = rb.quals.pageable = true;
    uint64_t n_batches = 0, n_reads = 0;
    unsigned long crc = crc32(0L, Z_NULL, 0);
    const bool timing_only = getenv("COLORID_B200_PARSE_NODIGEST") != nullptr;          // (zlib's crc32 is slower than the loop it checks)
    auto fold = [&](const void* p, size_t n) { if (!timing_only) crc = crc32(crc, (const Bytef*)p, (uInt)n); };
    auto digest = [&]() {
        if (!rb.n()) return;
        n_batches++; n_reads += rb.n();
        const uint64_t n = rb.n();
        fold(&n, 8); fold(rb.bases.p, rb.bases.size);
        const unsigned char q = rb.any_qual ? 1 : 0;
        fold(&q, 1);
        if (rb.any_qual) fold(rb.quals.p, rb.quals.size);
        fold(rb.seq_offs.data(), rb.seq_offs.size() * 8); fold(rb.read_offs.data(), rb.read_offs.size() * 8);
        fold(rb.id_data.data(), rb.id_data.size()); fold(rb.id_off.data(), rb.id_off.size() * 4);
        rb.clear();
    };
    auto cur_batch = [&]() -> ReadBatch& { return rb; };
    auto maybe_flush = [&]() { if (rb.n() >= kGpuBatchReads || rb.bases.size >= kGpuBatchBytes) digest(); };
    std::unique_ptr<AsyncLineReader> a(new AsyncLineReader(files[0])), c(files.size() > 1 ? new AsyncLineReader(files[1]) : nullptr);
    AsyncLineReader* rd[2] = {a.get(), c.get()};
    parse_fastq_records(rd, (int)files.size(), quality, cur_batch, maybe_flush);
    digest();
    printf("reads\t%llu\nbatches\t%llu\ncrc32\t%08lx\n", (unsigned long long)n_reads, (unsigned long long)n_batches, crc);
    return 0;
}

int read_id(const ReadIdOpts& o) {
    if (o.query.empty()) throw Error("no query files");
    Timer tload;
    Trace tr;
    Gpu g(o.device, /*multi=*/true);
    const Bigsi b = read_index(o.bigsi);
    fprintf(stderr, "Index loaded in %llu seconds\n", tload.secs());
    tr.mark("read_bigsi");
    g.ready();
    tr.mark("context create (rest)");
    g.upload(b);
    tr.mark("index upload");
    ReadBatch batches[2];
    read_id_sample(g, b, o, tr, batches);
    return 0;
}

// read_id_batch.rs:7-181 read_id_batch: the index is loaded (here: uploaded to the GPU) once, then every sample of
// the tab-delimited list [sample_name reads1 reads2(optional)] is classified into SAMPLE_TAG_reads.txt / _counts.txt.
// Samples run in byte-wise sorted order of their names (the reference iterates an FnvHashMap; the files are the same).
int batch_id(const BatchIdOpts& o) {
    Timer tload;
    Trace tr;
    const std::vector<int> devices = o.devices.empty() ? std::vector<int>{o.device} : o.devices;
    const auto batch_map = tab_to_map(o.batch_samples);
    const Bigsi b = read_index(o.bigsi);
    fprintf(stderr, "Index loaded in %llu seconds\n", tload.secs());
    // samples are independent units (replicated index): one worker per GPU, each with its own replica of
    // the matrix, pulls the next sample; with one device this is the reference's sequential loop
    std::vector<std::pair<std::string, std::vector<std::string>>> samples(batch_map.begin(), batch_map.end());
    std::atomic<size_t> next{0};
    std::mutex log_mu;
    std::vector<std::exception_ptr> errs(devices.size());
    auto worker = [&](size_t w) {
        try {
            Gpu g(devices[w]);
            g.upload(b);
            Trace wtr;
            ReadBatch batches[2];
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= samples.size()) break;
                { std::lock_guard<std::mutex> lk(log_mu); fprintf(stderr, "Classifying %s\n", samples[i].first.c_str()); }
                ReadIdOpts s;
                s.bigsi = o.bigsi; s.prefix = samples[i].first + "_" + o.tag; s.query = samples[i].second;
                s.threads = o.threads; s.down_sample = o.down_sample; s.batch = o.batch; s.bitvector_sample = o.bitvector_sample;
                s.correct = o.correct; s.quality = o.quality; s.high_mem_load = o.high_mem_load; s.device = devices[w];
                read_id_sample(g, b, s, wtr, batches);
            }
        } catch (...) { errs[w] = std::current_exception(); }
    };
    if (devices.size() == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (size_t w = 0; w < devices.size(); w++) th.emplace_back(worker, w);
        for (auto& t : th) t.join();
    }
    for (auto& e : errs) if (e) std::rethrow_exception(e);
    tr.mark("batch_id");
    return 0;
}

// ------------------------------------------------------------------ info
int info(const std::string& path) {
    fprintf(stderr, "Loading index\n");
    const Bigsi b = load_index(path);
    if (b.mini)               // main.rs:645-648 (the blank line and the leading space are the reference's)
        printf("BIGSI parameters:\nBloomfilter-size: %llu\nNumber of hashes: %llu\nK-mer size: %llu\n minimizer size: %llu\n\n",
               (unsigned long long)b.bloom_size, (unsigned long long)b.num_hash, (unsigned long long)b.k_size, (unsigned long long)b.m_size);
    else
        printf("BIGSI parameters:\nBloomfilter-size: %llu\nNumber of hashes: %llu\nK-mer size: %llu\n", (unsigned long long)b.bloom_size,
               (unsigned long long)b.num_hash, (unsigned long long)b.k_size);
    printf("Number of accessions in index: %zu\n", b.colors.size());
    std::vector<std::string> acc;
    for (auto& kv : b.colors) acc.push_back(kv.second);
    std::sort(acc.begin(), acc.end());
    for (auto& a : acc) {
        auto it = b.n_ref_kmers.find(a);
        if (it == b.n_ref_kmers.end()) throw Error("index has no k-mer count for accession " + a);
        printf("%s %llu %.3f\n", a.c_str(), (unsigned long long)it->second,
               false_prob((double)b.bloom_size, (double)b.num_hash, (double)it->second));
    }
    return 0;
}

}  // namespace cidh
