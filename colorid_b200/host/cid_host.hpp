// Host side above the C ABI (include/colorid_b200.h): the callers and data formats either side of the
// hot path, mirroring the reference's own `pub fn`s (same names, argument meaning, output bytes):
//
//   kmer.rs:10-84        read_fasta, read_fasta_mf            -> fastx.cpp
//   kmer.rs:461-475,581-612, read_id_mt_pe.rs:727-761,870-880  FASTQ(.gz) record iteration -> fastx.cpp
//   seq.rs:36-56         qual_mask                            -> fastx.cpp
//   build.rs:15-31       tab_to_map                           -> fastx.cpp
//   bigsi.rs:19-27,51-69 BigsyMapNew, save_bigsi, read_bigsi  -> bxi.cpp   (bincode 1.x layout)
//   bigsi.rs:40-49,71-89 BigsyMapMiniNew, save_bigsi_mini, read_bigsi_mini (.mxi) -> bxi.cpp
//   reports.rs:8-120     generate_report[_gene], mode, read_counts_five_fields -> reports.cpp
//   build.rs:33-492      build_single / build_multi [_mini]   -> drivers.cpp
//   batch_search_pe.rs:9-179, perfect_search.rs:6-120         -> drivers.cpp
//   read_id_mt_pe.rs:440-951 stream_fasta, per_read_stream_pe/se -> drivers.cpp
//   read_id_batch.rs:7-181 read_id_batch                      -> drivers.cpp (batch_id)
//   read_filter.rs:10-191 tab_to_map, read_filter_pe/se       -> fastx.cpp (read_filter)
//   main.rs              clap CLI (build / search / read_id / batch_id / read_filter / info) -> main.cpp
//
// The reference is Rust; this image has no Rust toolchain, so the host side is C++.
// Everything compute-heavy goes through the C ABI to the GPU; nothing here has a CPU fallback.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <map>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

namespace cidh {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

// ---- sequences -----------------------------------------------------------------------------------
// A list of sequences in the ABI's layout: one byte array + offsets[n+1].
struct SeqBatch {
    std::string bases;
    std::vector<uint64_t> offs{0};
    void add(const std::string& s) { bases += s; offs.push_back(bases.size()); }
    void add(const char* p, size_t n) { bases.append(p, n); offs.push_back(bases.size()); }
    uint64_t n() const { return offs.size() - 1; }
    void clear() { bases.clear(); offs.assign(1, 0); }
};

// Lines of a plain or gzip file (concatenated members like flate2::MultiGzDecoder); the line
// terminator ("\n" or "\r\n") is stripped like BufRead::lines() unless keep_eol.
// A gzip file mapped into memory and decoded by the host layer's own inflater (fast_inflate.hpp).
class MappedGz {
public:
    explicit MappedGz(const std::string& path);      // throws if the file cannot be mapped or is not gzip
    ~MappedGz();
    size_t read(char* dst, size_t cap);
    static bool eligible(const std::string& path);    // regular file that starts with the gzip magic
private:
    class GzInflater* inf_ = nullptr;    // one of the two
    class GzParallel* par_ = nullptr;
    static std::atomic<int> live_;       // MappedGz objects alive: they share the cores
    void* map_ = nullptr;
    size_t len_ = 0;
};

class LineReader {
public:
    explicit LineReader(const std::string& path);
    ~LineReader();
    bool next(std::string& line, bool keep_eol = false);
    // Up to max_lines lines at once: their raw bytes (EOLs included) are appended to `data`, line i of the call is
    // data[begin[i], end[i]) (EOL -- "\n" or "\r\n" -- excluded unless keep_eol).  Returns the number of lines added.
    size_t next_lines(std::string& data, std::vector<uint32_t>& begin, std::vector<uint32_t>& end, size_t max_lines, bool keep_eol = false);
private:
    void* gz_ = nullptr;                 // zlib gzFile: plain files, pipes, COLORID_B200_ZLIB=1
    MappedGz* fast_ = nullptr;           // regular gzip files: the host layer's own decoder over the mapped file
    std::vector<char> buf_;
    size_t pos_ = 0, len_ = 0;
    bool eof_ = false;
    bool fill();
};

// The same interface with decompression and line splitting on a producer thread (the
// reference gunzips and parses serially between batches, read_id_mt_pe.rs:727-761).  Blocks of lines travel
// through a bounded queue, so both mates' files inflate in parallel with each other, with batch assembly and
// with the GPU call of the previous batch.
class AsyncLineReader {
public:
    explicit AsyncLineReader(const std::string& path, bool keep_eol = false);
    ~AsyncLineReader();
    bool next(std::string& line);
    // The next line as a view into the reader's current block: valid until the call that crosses into the next block, i.e. --
    // blocks hold a multiple of four lines -- for the four lines of a FASTQ record at least.
    bool next_view(std::string_view& line);
    // The unread lines of the current block at once: line i is data[begin[i], end[i]) for at <= i < n.  false at the end of the
    // file.  skip_lines(k) marks k of them as read (k <= n - at).
    struct LineBlock { const char* data; const uint32_t* begin; const uint32_t* end; size_t n; };
    bool peek_block(LineBlock& b, size_t& at);
    void skip_lines(size_t k);
private:
    struct Impl;
    Impl* p_;
    bool load();
};

std::vector<std::string> read_fasta(const std::string& path);                                      // kmer.rs:10-45
void read_fasta_mf(const std::string& path, std::vector<std::string>& labels, std::vector<std::string>& seqs);  // kmer.rs:47-84
std::string qual_mask(const std::string& seq, const std::string& qual, uint8_t max_quality_offset);   // seq.rs:36-56
// accession -> 1 (FASTA or single-end FASTQ.gz) or 2 (paired FASTQ.gz) file names; a later line with
// the same accession replaces an earlier one; iteration is in byte-wise sorted order = colour order
std::map<std::string, std::vector<std::string>> tab_to_map(const std::string& path);                // build.rs:15-31
// All records of a single-end / paired FASTQ(.gz) as quality-masked sequences (mates are separate
// sequences); paired input stops at the shorter file (kmer.rs:604,612,649).  Returns records read.
int fastq_parse_digest(const std::vector<std::string>& files, uint8_t quality);                     // `_host fastq_parse` (tests, timing)
uint64_t fastq_masked_se(const std::string& path, uint8_t qual_offset, SeqBatch& out);              // kmer.rs:461-475
uint64_t fastq_masked_pe(const std::string& p1, const std::string& p2, uint8_t qual_offset, SeqBatch& out);   // kmer.rs:581-612

// ---- .bxi ------------------------------------------------------------------------------------------
struct Bigsi {                                  // bigsi.rs:19-27 BigsyMapNew / :40-49 BigsyMapMiniNew
    uint64_t bloom_size = 0, num_hash = 0, k_size = 0;
    uint64_t m_size = 0;                        // BigsyMapMiniNew only (.mxi)
    bool mini = false;
    std::map<uint64_t, std::string> colors;     // colour -> accession
    std::vector<uint64_t> row_ids;              // map keys (non-zero rows)
    std::vector<uint32_t> words;                // map values: row_words u32 per row (BitVec storage)
    uint32_t row_words = 0;
    std::map<std::string, uint64_t> n_ref_kmers;
    uint32_t n_colors() const { return (uint32_t)colors.size(); }
};
void save_bigsi(const std::string& path, const Bigsi& b);   // bigsi.rs:51-57
Bigsi read_bigsi(const std::string& path);                  // bigsi.rs:59-69 (both variants: same bytes)
void save_bigsi_mini(const std::string& path, const Bigsi& b);   // bigsi.rs:71-77 (.mxi: m_size after k_size)
Bigsi read_bigsi_mini(const std::string& path);             // bigsi.rs:79-89

// ---- reports ---------------------------------------------------------------------------------------
// reports.rs:8-48: one line per accession with hits/n_ref > cov (ascending colour; the reference's order
// is SipHash-random)
void generate_report(FILE* out, const std::string& query, const Bigsi& ix, const uint32_t* counts, const uint64_t* uniq_n,
                     const uint64_t* uniq_sum, const uint64_t* uniq_mode, uint64_t num_kmers, double cov);
// reports.rs:50-62
void generate_report_gene(FILE* out, const std::string& query, const Bigsi& ix, const uint32_t* counts, uint64_t num_kmers,
                          double cov);
// Iteration order of an FnvHashMap<String, _> filled with entry(key).or_insert() in the given order of
// first appearance (hashbrown, group width 16): reports.rs:98-120 writes PREFIX_counts.txt in it.
std::vector<size_t> fnv_string_map_order(const std::vector<std::string>& keys_in_insertion_order, uint32_t group_width = 16);
// reports.rs:98-120 from the (class, accept|reject) column pair of every output line
void write_counts_five_fields(const std::string& path, const std::vector<std::string>& insertion_keys,
                              const std::map<std::string, uint64_t>& counts);
double false_prob(double m, double k, double n);            // read_id_mt_pe.rs:695-698 (for `info`)

// ---- drivers (need a GPU: every one of them goes through the C ABI) -------------------------------
struct BuildOpts { std::string ref_file, prefix; uint64_t k = 31, bloom = 50000000, hashes = 4, threads = 1; uint8_t quality = 15; int64_t filter = -1; int device = 0;
                   bool minimizer = false; uint64_t minimizer_value = 15; };   // -m / -v (main.rs:480-482)
int build(const BuildOpts& o);                              // main.rs:467-553 + build.rs:33-256
struct SearchOpts { std::string bigsi; std::vector<std::string> files1, files2; int64_t filter = -1; double cov = 0.35;
                    bool gene_search = false, perfect_search = false, multi_fasta = false; uint8_t quality = 15; int device = 0; };
int search(const SearchOpts& o);                            // main.rs:555-628
struct ReadIdOpts { std::string bigsi, prefix; std::vector<std::string> query; uint64_t threads = 0, down_sample = 1, batch = 50000,
                    bitvector_sample = 3; double correct = 3.0; uint8_t quality = 15; bool high_mem_load = false; int device = 0; };
int read_id(const ReadIdOpts& o);                           // main.rs:704-866
struct BatchIdOpts { std::string bigsi, batch_samples, tag; uint64_t threads = 0, down_sample = 1, batch = 50000, bitvector_sample = 3;
                     double correct = 3.0; uint8_t quality = 15; bool high_mem_load = false; int device = 0;
                     std::vector<int> devices; };          // COLORID_B200_DEVICES=0,1,..: samples dealt to one worker per GPU
int batch_id(const BatchIdOpts& o);                         // main.rs:869-887 + read_id_batch.rs:7-181
int info(const std::string& bigsi);                         // main.rs:630-703
// read_filter.rs:10-191 + main.rs:888-900: keep (or with `exclude` drop) the reads whose PREFIX_reads.txt classification
// contains `taxon`; writes PREFIX_TAXON_R1.fq.gz / _R2.fq.gz (paired) or PREFIX_TAXON.fq.gz.  Pure file I/O, no GPU.
struct ReadFilterOpts { std::string classification, taxon, prefix; std::vector<std::string> files; bool exclude = false; };
int read_filter(const ReadFilterOpts& o);

}  // namespace cidh
