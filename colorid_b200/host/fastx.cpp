// Sequence file readers of the host layer: FASTA (kmer.rs:10-84), FASTQ(.gz) record iteration
// (kmer.rs:461-475,581-612), seq.rs:36-56 qual_mask and build.rs:15-31 tab_to_map.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <fstream>
#include <memory>
#include <mutex>
#include <sstream>
#include <thread>

#include "cid_host.hpp"
#include "fast_inflate.hpp"

namespace cidh {

// ------------------------------------------------------------------ LineReader
bool MappedGz::eligible(const std::string& path) {
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    unsigned char magic[2] = {0, 0};
    const bool ok = fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size >= 18 && pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    close(fd);
    return ok;
}
std::atomic<int> MappedGz::live_{0};
MappedGz::MappedGz(const std::string& path) {
    struct Live { bool keep = false; Live() { live_++; } ~Live() { if (!keep) live_--; } } live;
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) throw Error("file not found: " + path);
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size == 0) { close(fd); throw Error("cannot map " + path); }
    void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) throw Error("cannot map " + path);
    madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
    map_ = m; len_ = (size_t)st.st_size;
    // Large files are decoded by several threads (GzParallel): COLORID_B200_GZ_THREADS (default: 3/4 of the cores shared by the open
    // files, at most 6 each; 1 = the sequential decoder), COLORID_B200_GZ_SPAN = compressed bytes per thread and round (default 2 MiB).
    GzParallel::Config cfg;
    const unsigned hw = std::thread::hardware_concurrency();
    // 3/4 of the cores are shared by the gzip files open at this moment (two per paired sample; `batch_id` on several GPUs has
    // several samples open), at most 6 threads each
    const unsigned open_now = (unsigned)std::max(1, live_.load());
    cfg.threads = std::min(6u, std::max(1u, hw * 3 / 4 / open_now));
    if (const char* e = getenv("COLORID_B200_GZ_THREADS")) { const long v = atol(e); if (v >= 1 && v <= 64) cfg.threads = (unsigned)v; }
    if (const char* e = getenv("COLORID_B200_GZ_SPAN")) { const long long v = atoll(e); if (v >= 4096) cfg.span = (size_t)v; }
    try {
        if (cfg.threads >= 2 && len_ >= 4 * cfg.span) par_ = new GzParallel((const uint8_t*)m, len_, path, cfg);
        else inf_ = new GzInflater((const uint8_t*)m, len_, path);
    } catch (...) {
        munmap(m, len_);
        throw;
    }
    live.keep = true;
}
MappedGz::~MappedGz() {
    delete par_;
    delete inf_;
    if (map_) munmap(map_, len_);
    live_--;
}
size_t MappedGz::read(char* dst, size_t cap) { return par_ ? par_->read((uint8_t*)dst, cap) : inf_->read((uint8_t*)dst, cap); }

LineReader::LineReader(const std::string& path) : buf_(1 << 20) {
    // A regular file that starts with the gzip magic is mapped and decoded by the host layer's own inflater (every member's
    // CRC checked); anything else -- plain text, pipes, or COLORID_B200_ZLIB=1 -- goes through zlib's gzread (transparent
    // for plain files, walks concatenated members).
    const char* force = getenv("COLORID_B200_ZLIB");
    if (!(force && *force && *force != '0') && MappedGz::eligible(path)) { fast_ = new MappedGz(path); return; }
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) throw Error("file not found: " + path);
    gzbuffer(f, 1 << 20);
    gz_ = f;
}
LineReader::~LineReader() {
    if (gz_) gzclose((gzFile)gz_);
    delete fast_;
}
bool LineReader::fill() {
    if (eof_) return false;
    size_t n;
    if (fast_) n = fast_->read(buf_.data(), buf_.size());
    else {
        const int r = gzread((gzFile)gz_, buf_.data(), (unsigned)buf_.size());
        if (r < 0) throw Error("gz read error");
        n = (size_t)r;
    }
    pos_ = 0; len_ = n;
    if (n == 0) eof_ = true;
    return n > 0;
}
bool LineReader::next(std::string& line, bool keep_eol) {
    line.clear();
    bool any = false;
    for (;;) {
        if (pos_ == len_ && !fill()) break;
        const char* b = buf_.data() + pos_;
        const char* nl = (const char*)memchr(b, '\n', len_ - pos_);
        if (nl) {
            line.append(b, nl - b + (keep_eol ? 1 : 0));
            pos_ += (size_t)(nl - b) + 1;
            if (!keep_eol && !line.empty() && line.back() == '\r') line.pop_back();   // BufRead::lines strips CRLF
            return true;
        }
        line.append(b, len_ - pos_);
        pos_ = len_;
        any = true;
    }
    return any || !line.empty();
}

size_t LineReader::next_lines(std::string& data, std::vector<uint32_t>& begin, std::vector<uint32_t>& end, size_t max_lines, bool keep_eol) {
    size_t added = 0;
    size_t line_start = data.size();            // where the line being assembled began (it may span refills of the buffer)
    while (added < max_lines) {
        if (pos_ == len_ && !fill()) {
            if (data.size() > line_start) { begin.push_back((uint32_t)line_start); end.push_back((uint32_t)data.size()); added++; }   // last line without EOL
            break;
        }
        const char* const b = buf_.data() + pos_;
        const char* const e = buf_.data() + len_;
        const size_t d0 = data.size();
        const char* s = b;
        while (added < max_lines) {
            const char* nl = (const char*)memchr(s, '\n', (size_t)(e - s));
            if (!nl) break;
            const size_t lend = d0 + (size_t)(nl - b);       // where this '\n' lands in `data`
            size_t le = keep_eol ? lend + 1 : lend;
            if (!keep_eol && lend > line_start) {
                const char prev = nl > b ? nl[-1] : data[d0 - 1];
                if (prev == '\r') le--;                      // BufRead::lines strips CRLF
            }
            begin.push_back((uint32_t)line_start);
            end.push_back((uint32_t)le);
            line_start = lend + 1;
            added++;
            s = nl + 1;
        }
        if (added == max_lines) { data.append(b, (size_t)(s - b)); pos_ += (size_t)(s - b); break; }
        data.append(b, (size_t)(e - b));                      // whole lines and the start of the next one
        pos_ = len_;
    }
    return added;
}

// ------------------------------------------------------------------ AsyncLineReader
struct AsyncLineReader::Impl {
    struct Block { std::string data; std::vector<uint32_t> begin, end; };       // line i = data[begin[i] .. end[i])
    enum { BLOCK_LINES = 1 << 16, DEPTH = 8 };
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::unique_ptr<Block>> q;
    std::vector<std::unique_ptr<Block>> spare;          // consumed blocks go back to the producer: a fresh 10 MB string per block is page faults
    bool done = false, stop = false;
    std::string error;
    std::thread th;
    std::unique_ptr<Block> cur;
    size_t at = 0;
    void produce(std::string path, bool keep_eol) {
        try {
            LineReader lr(path);
            for (;;) {
                std::unique_ptr<Block> b;
                {
                    std::lock_guard<std::mutex> lk(mu);
                    if (!spare.empty()) { b = std::move(spare.back()); spare.pop_back(); }
                }
                if (b) { b->data.clear(); b->begin.clear(); b->end.clear(); }
                else {
                    b.reset(new Block());
                    b->data.reserve(10 << 20);
                    b->begin.reserve(BLOCK_LINES);
                    b->end.reserve(BLOCK_LINES);
                }
                lr.next_lines(b->data, b->begin, b->end, BLOCK_LINES, keep_eol);     // whole lines, split with memchr: no per-line strings here
                const bool last = b->end.size() < BLOCK_LINES;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return q.size() < DEPTH || stop; });
                    if (stop) return;
                    if (!b->end.empty()) q.push_back(std::move(b));
                    if (last) done = true;
                }
                cv.notify_all();
                if (last) return;
            }
        } catch (const std::exception& e) {
            std::lock_guard<std::mutex> lk(mu);
            error = e.what();
            done = true;
            cv.notify_all();
        }
    }
};
AsyncLineReader::AsyncLineReader(const std::string& path, bool keep_eol) : p_(new Impl()) {
    {                                                // a missing file fails here, on the caller's thread
        const int fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) throw Error("file not found: " + path);
        close(fd);
    }
    p_->th = std::thread(&Impl::produce, p_, path, keep_eol);
}
AsyncLineReader::~AsyncLineReader() {
    { std::lock_guard<std::mutex> lk(p_->mu); p_->stop = true; }
    p_->cv.notify_all();
    if (p_->th.joinable()) p_->th.join();
    delete p_;
}
// makes the current block one with unread lines; false at the end of the file
bool AsyncLineReader::load() {
    Impl& s = *p_;
    if (!s.cur || s.at == s.cur->end.size()) {
        std::unique_lock<std::mutex> lk(s.mu);
        s.cv.wait(lk, [&] { return !s.q.empty() || s.done; });
        if (s.q.empty()) {
            if (!s.error.empty()) throw Error(s.error);
            return false;
        }
        if (s.cur) s.spare.push_back(std::move(s.cur));
        s.cur = std::move(s.q.front());
        s.q.pop_front();
        s.at = 0;
        lk.unlock();
        s.cv.notify_all();
    }
    return true;
}
bool AsyncLineReader::next_view(std::string_view& line) {
    if (!load()) return false;
    Impl& s = *p_;
    const uint32_t lo = s.cur->begin[s.at], hi = s.cur->end[s.at];
    line = std::string_view(s.cur->data.data() + lo, hi - lo);
    s.at++;
    return true;
}
bool AsyncLineReader::peek_block(LineBlock& b, size_t& at) {
    if (!load()) return false;
    Impl& s = *p_;
    b.data = s.cur->data.data(); b.begin = s.cur->begin.data(); b.end = s.cur->end.data(); b.n = s.cur->end.size();
    at = s.at;
    return true;
}
void AsyncLineReader::skip_lines(size_t k) { p_->at += k; }
bool AsyncLineReader::next(std::string& line) {
    std::string_view v;
    if (!next_view(v)) return false;
    line.assign(v.data(), v.size());
    return true;
}

// ------------------------------------------------------------------ FASTA
// str::lines(): split on '\n', a trailing '\r' is dropped, no empty line after a final newline.
static std::vector<std::string> file_lines(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error("file not found: " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string c = ss.str();
    std::vector<std::string> lines;
    size_t at = 0;
    while (at < c.size()) {
        size_t nl = c.find('\n', at);
        size_t end = nl == std::string::npos ? c.size() : nl;
        size_t e2 = end;
        if (nl != std::string::npos && e2 > at && c[e2 - 1] == '\r') e2--;
        lines.emplace_back(c, at, e2 - at);
        at = nl == std::string::npos ? c.size() : nl + 1;
    }
    return lines;
}

// A record ends at the next line holding a '>' (anywhere in the line) or at the last line of the
// file; empty records are dropped (kmer.rs:26-42), which is why labels and sequences of
// read_fasta_mf can fall out of step exactly as in the reference.
static void fasta_records(const std::string& path, std::vector<std::string>* labels, std::vector<std::string>& seqs) {
    const std::vector<std::string> lines = file_lines(path);
    std::string cur;
    for (size_t i = 0; i < lines.size(); i++) {
        const std::string& l = lines[i];
        if (l.find('>') != std::string::npos) {
            if (labels) labels->push_back(l.substr(1));
            if (!cur.empty()) seqs.push_back(cur);
            cur.clear();
        } else {
            cur += l;
            if (i + 1 == lines.size() && !cur.empty()) seqs.push_back(cur);
        }
    }
}
std::vector<std::string> read_fasta(const std::string& path) {
    std::vector<std::string> seqs;
    fasta_records(path, nullptr, seqs);
    return seqs;
}
void read_fasta_mf(const std::string& path, std::vector<std::string>& labels, std::vector<std::string>& seqs) {
    labels.clear(); seqs.clear();
    fasta_records(path, &labels, seqs);
}

// ------------------------------------------------------------------ qual_mask
// Output has one base per QUALITY character (a longer sequence is truncated; a shorter one makes
// the reference panic): 'N' where phred+33 < offset+33.  offset 0 returns the sequence untouched.
std::string qual_mask(const std::string& seq, const std::string& qual, uint8_t off) {
    if (off == 0) return seq;
    if (qual.size() > seq.size()) throw Error("ERROR: could not get the next nt in the sequence");
    const uint8_t maxq = (uint8_t)(off + 33);
    std::string out(qual.size(), 'N');
    for (size_t i = 0; i < qual.size(); i++)
        if ((uint8_t)qual[i] >= maxq) out[i] = seq[i];
    return out;
}

// ------------------------------------------------------------------ tab_to_map
std::map<std::string, std::vector<std::string>> tab_to_map(const std::string& path) {
    std::map<std::string, std::vector<std::string>> m;
    LineReader lr(path);
    std::string l;
    while (lr.next(l)) {
        std::vector<std::string> v;
        size_t at = 0;
        for (;;) {
            size_t t = l.find('\t', at);
            v.emplace_back(l, at, t == std::string::npos ? std::string::npos : t - at);
            if (t == std::string::npos) break;
            at = t + 1;
        }
        if (v.size() < 2) throw Error("reference file: line without a tab: '" + l + "'");   // v[1] panics in the reference
        if (v.size() == 2) m[v[0]] = {v[1]};
        else m[v[0]] = {v[1], v[2]};
    }
    return m;
}

// ------------------------------------------------------------------ FASTQ
// seq.rs:36-56 applied while appending to the batch (no per-record strings)
static void add_masked(SeqBatch& out, std::string_view seq, std::string_view qual, uint8_t off) {
    if (off == 0) { out.add(seq.data(), seq.size()); return; }
    if (qual.size() > seq.size()) throw Error("ERROR: could not get the next nt in the sequence");
    const uint8_t maxq = (uint8_t)(off + 33);
    const size_t at = out.bases.size(), n = qual.size();
    out.bases.append(seq.data(), n);
    char* p = &out.bases[at];
    const char* q = qual.data();
    for (size_t i = 0; i < n; i++) p[i] = (uint8_t)q[i] < maxq ? 'N' : p[i];       // (a select, not a branch: vectorised)
    out.offs.push_back(out.bases.size());
}
// (the lines are views into the readers' blocks: the four lines of a record share a block)
uint64_t fastq_masked_se(const std::string& path, uint8_t qual_offset, SeqBatch& out) {
    AsyncLineReader lr(path);
    std::string_view l, seq;
    uint64_t line_count = 1, n = 0;
    while (lr.next_view(l)) {
        if (line_count % 4 == 2) seq = l;
        else if (line_count % 4 == 0) { add_masked(out, seq, l, qual_offset); n++; }
        line_count++;
    }
    return n;
}
uint64_t fastq_masked_pe(const std::string& p1, const std::string& p2, uint8_t qual_offset, SeqBatch& out) {
    AsyncLineReader a(p1), b(p2);           // both files inflate on their own threads
    std::string_view l1, l2, s1, s2;
    uint64_t line_count = 1, n = 0;
    while (a.next_view(l1)) {
        const bool have2 = b.next_view(l2);
        if (line_count % 4 == 1) { if (!have2) break; }
        else if (line_count % 4 == 2) { if (!have2) break; s1 = l1; s2 = l2; }
        else if (line_count % 4 == 0) {
            if (!have2) break;
            add_masked(out, s1, l1, qual_offset);
            add_masked(out, s2, l2, qual_offset);
            n++;
        }
        line_count++;
    }
    return n;
}

}  // namespace cidh

// ------------------------------------------------------------------ read_filter (read_filter.rs)
namespace cidh {
namespace {
struct GzOut {
    gzFile f;
    explicit GzOut(const std::string& path, const char* what) : f(gzopen(path.c_str(), "wb6")) {   // Compression::default() = level 6
        if (!f) throw Error(what);
        gzbuffer(f, 1 << 20);
    }
    ~GzOut() { if (f) gzclose(f); }
    void record(const std::string& h, const std::string& s, const std::string& q) {
        std::string r = h + "\n" + s + "\n+\n" + q + "\n";
        if (gzwrite(f, r.data(), (unsigned)r.size()) != (int)r.size()) throw Error("could not write reads!");
    }
    void finish() { int rc = gzclose(f); f = nullptr; if (rc != Z_OK) throw Error("Could not close new read file"); }
};
std::string first_word(const std::string& s) { return s.substr(0, s.find(' ')); }   // header.split(' ')[0]
}  // namespace

int read_filter(const ReadFilterOpts& o) {
    // read_filter.rs:10-28 tab_to_map: read id (up to the first blank) of every line whose class CONTAINS the taxon
    std::map<std::string, std::string> cls;
    {
        LineReader lr(o.classification);
        std::string l;
        while (lr.next(l)) {
            const size_t t1 = l.find('\t');
            if (t1 == std::string::npos) throw Error("classification file: line without a class column (the reference panics)");
            const size_t t2 = l.find('\t', t1 + 1);
            const std::string c = l.substr(t1 + 1, t2 == std::string::npos ? std::string::npos : t2 - t1 - 1);
            if (c.find(o.taxon) != std::string::npos) cls[first_word(l.substr(0, t1))] = c;
        }
    }
    std::string cleaned = o.taxon;
    for (auto& ch : cleaned) if (ch == ' ') ch = '_';
    uint64_t kept = 0;
    if (o.files.size() == 1) {              // read_filter_se, read_filter.rs:139-191
        LineReader a(o.files[0]);
        GzOut out(o.prefix + "_" + cleaned + ".fq.gz", "could not create R1!");
        std::string l, h, s;
        uint64_t line_count = 1;
        while (a.next(l)) {
            if (line_count % 4 == 1) h = l;
            else if (line_count % 4 == 2) s = l;
            else if (line_count % 4 == 0 && (cls.count(first_word(h)) != 0) != o.exclude) { out.record(h, s, l); kept++; }
            line_count++;
        }
        out.finish();
    } else {                                // read_filter_pe, :31-136: stops at the shorter file; the R1 header decides
        if (o.files.size() < 2) throw Error("no read files");
        LineReader a(o.files[0]), b(o.files[1]);
        GzOut o1(o.prefix + "_" + cleaned + "_R1.fq.gz", "could not create R1!"), o2(o.prefix + "_" + cleaned + "_R2.fq.gz", "could not create R2!");
        std::string l1, l2, h1, h2, s1, s2;
        uint64_t line_count = 1;
        while (a.next(l1)) {
            if (!b.next(l2)) { if (line_count % 4 != 3) break; l2.clear(); }   // the '+' line of mate 2 is never looked at
            if (line_count % 4 == 1) { h1 = l1; h2 = l2; }
            else if (line_count % 4 == 2) { s1 = l1; s2 = l2; }
            else if (line_count % 4 == 0 && (cls.count(first_word(h1)) != 0) != o.exclude) { o1.record(h1, s1, l1); o2.record(h2, s2, l2); kept++; }
            line_count++;
        }
        o1.finish(); o2.finish();
    }
    if (o.exclude) fprintf(stderr, "Excluded %llu read pairs  with classification containing '%s' from output files\n", (unsigned long long)kept, o.taxon.c_str());
    else fprintf(stderr, "Wrote %llu read-pairs with classification containing '%s' to output files\n", (unsigned long long)kept, o.taxon.c_str());
    return 0;
}
}  // namespace cidh
