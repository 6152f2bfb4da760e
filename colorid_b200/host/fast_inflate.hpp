// A streaming gzip (RFC 1952 / DEFLATE RFC 1951) decoder for the read files of the host layer.
//
// The reference inflates its FASTQ.gz inputs with flate2 on the parsing thread (read_id_mt_pe.rs:727-761,
// kmer.rs:461-475); the host layer here did the same with zlib's gzread, and `read_id` of .fastq.gz pairs was bound by
// it: zlib decodes FASTQ text -- mostly literals, 2 bits of entropy per base -- at ~0.2-0.4 GB/s per stream.  This
// decoder keeps 64 bits of input in a register, decodes literal/length codes through an 11-bit first-level table (a
// short subtable for longer codes), takes up to three literals per refill and copies matches eight bytes at a time.
// Every member's CRC-32 and length are checked against its trailer (zlib's crc32()), so a decoding error cannot pass
// silently; concatenated members (bgzip, `cat a.gz b.gz`) are walked like gzread does.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace cidh {

class GzInflater {
public:
    // `data` must stay valid (the caller maps the file); throws cidh::Error on a corrupt stream
    GzInflater(const uint8_t* data, size_t n, const std::string& what);
    // up to `cap` decoded bytes into dst; 0 = end of the last member
    size_t read(uint8_t* dst, size_t cap);
    // Take over a stream that another decoder (GzParallel) has followed up to bit `bitpos` of the data: either a member header
    // (at_header, byte-aligned) or a block header inside a member, of which `win` holds the last nwin = min(member_len, 32768)
    // bytes of output, `crc` the CRC-32 and member_len the length so far.  Call before the first read().
    void resume(uint64_t bitpos, bool at_header, const uint8_t* win, size_t nwin, uint32_t crc, uint64_t member_len);

    // shared with GzParallel (par_inflate.cpp)
    enum { LBITS = 11, DBITS = 8 };
    // decode-table entry: value << 16 | flags << 12 | extra bits << 8 | code bits to consume (F_LIT2: bit 7, two literals in one entry)
    enum : uint32_t { F_LIT = 1u << 12, F_BASE = 2u << 12, F_EOB = 4u << 12, F_SUB = 8u << 12, F_LIT2 = 1u << 7 };
    static void build_table(const uint8_t* lens, unsigned n, unsigned tbits, bool dist, std::vector<uint32_t>& table, bool& ok);
    static void fixed_lengths(uint8_t litlen[288], uint8_t dist[32]);
    // size of the member header at p (RFC 1952), or 0 with *why set
    static size_t header_size(const uint8_t* p, const uint8_t* end, const char** why);
    static uint32_t crc32_fast(uint32_t crc, const unsigned char* p, size_t n);

private:
    enum { HIST = 32768, CHUNK = 1 << 22, SLACK = 320 };
    enum State { ST_MEMBER, ST_BLOCK, ST_STORED, ST_HUFF, ST_TRAILER, ST_DONE };
    const uint8_t* base_; const uint8_t* in_; const uint8_t* end_;
    std::string what_;
    uint64_t bitbuf_ = 0;
    unsigned bitcnt_ = 0;
    uint64_t fed_zero_bytes_ = 0;           // zero bytes supplied past the end of the input (consuming them is an error)
    State st_ = ST_MEMBER;
    bool last_block_ = false;
    uint32_t stored_left_ = 0;
    std::vector<uint8_t> obuf_;             // [0, HIST) history | [HIST, HIST + CHUNK) the chunk being produced
    size_t out_ = HIST;                     // write position
    size_t taken_ = HIST;                   // bytes of the current chunk already handed to the caller
    size_t crc_from_ = HIST;                // bytes of the current chunk already folded into the CRC
    int64_t member_start_ = HIST;           // where the current member's output began, in obuf_ coordinates (negative once slid out)
    uint32_t crc_ = 0;
    std::vector<uint32_t> lt_, dt_;         // decode tables (first level + subtables)

    void refill();
    uint32_t bits(unsigned n);
    void byte_align();
    void fail(const char* why) const;
    void parse_header();
    void parse_trailer();
    void block_header();
    void build_fixed();
    void build_dynamic();
    void run_huff(size_t limit);
    void decode_chunk();
    void fold_crc();
};

// The same stream decoded by several threads (par_inflate.cpp).  The compressed data is cut into spans; the first span of a
// round starts at a known block header with a known 32 KB window, every other one looks for the header of a dynamic-Huffman
// block near its cut and decodes from there into 16-bit symbols in which a back-reference into the window it does not know
// yet is kept as a marker (window index | 0x8000).  A span's output is used only if the span before it ended on exactly the
// bit its own decoding started at -- the chain of such spans is the true stream by construction -- and its markers are then
// replaced from the real window.  Whatever does not fit that scheme (stored-only stretches, runs of small members, a span
// that inflates beyond the buffers, any error) is handed to the sequential GzInflater from the last verified position, so
// results and error messages are those of the sequential decoder.  CRC-32 and length of every member are checked here too.
class GzParallel {
public:
    struct Config { unsigned threads = 4; size_t span = (size_t)2 << 20; };
    GzParallel(const uint8_t* data, size_t n, const std::string& what, const Config& cfg);
    ~GzParallel();
    size_t read(uint8_t* dst, size_t cap);
    GzParallel(const GzParallel&) = delete;
    GzParallel& operator=(const GzParallel&) = delete;
private:
    struct Impl;
    Impl* impl_;
};

}  // namespace cidh
