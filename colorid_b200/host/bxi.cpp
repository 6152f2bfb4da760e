// .bxi reader / writer: bigsi.rs:51-69 save_bigsi / read_bigsi[_highmem] of BigsyMapNew (bigsi.rs:19-27)
// in bincode 1.x default configuration: little-endian, fixed-width integers,
// usize as u64, sequences / maps / strings prefixed with a u64 length, struct fields in declaration order:
//
//   u64 bloom_size | u64 num_hash | u64 k_size | [.mxi only: u64 m_size]   (BigsyMapMiniNew, bigsi.rs:40-49,71-89)
//   u64 n | n x { u64 colour ; u64 len ; bytes }                               colors
//   u64 n | n x { u64 row ; u64 n_words ; n_words x u32 ; u64 nbits }          map   (BitVec {storage, nbits})
//   u64 n | n x { u64 len ; bytes ; u64 n_ref_kmers }                          n_ref_kmers
//
// Rows stream straight between the file and the flat (row_ids, words) arrays the C ABI takes
// (cid_index_upload_rows / cid_index_download_nonzero_rows): no per-row heap objects.
#include <algorithm>
#include <cstring>

#include "cid_host.hpp"

namespace cidh {

namespace {
struct Writer {
    FILE* f;
    std::vector<char> buf;
    explicit Writer(const std::string& path) : f(fopen(path.c_str(), "wb")) {
        if (!f) throw Error("could not create " + path);
        buf.reserve(1 << 22);
    }
    ~Writer() { if (f) fclose(f); }
    void flush() {
        if (!buf.empty() && fwrite(buf.data(), 1, buf.size(), f) != buf.size()) throw Error("problems preparing serialized data for writing");
        buf.clear();
    }
    void raw(const void* p, size_t n) {
        if (buf.size() + n > buf.capacity()) flush();
        if (n > buf.capacity()) { if (fwrite(p, 1, n, f) != n) throw Error("write failed"); return; }
        buf.insert(buf.end(), (const char*)p, (const char*)p + n);
    }
    void u64(uint64_t v) { raw(&v, 8); }          // host is little-endian (x86_64 / aarch64)
    void str(const std::string& s) { u64(s.size()); raw(s.data(), s.size()); }
};
struct Reader {
    FILE* f;
    explicit Reader(const std::string& path) : f(fopen(path.c_str(), "rb")) {
        if (!f) throw Error("Can't open index!");
        setvbuf(f, nullptr, _IOFBF, 1 << 22);
    }
    ~Reader() { if (f) fclose(f); }
    void raw(void* p, size_t n) { if (n && fread(p, 1, n, f) != n) throw Error("can't deserialize"); }
    uint64_t u64() { uint64_t v; raw(&v, 8); return v; }
    std::string str() {
        uint64_t n = u64();
        if (n > (1ull << 32)) throw Error("can't deserialize");
        std::string s(n, '\0');
        raw(&s[0], n);
        return s;
    }
};
}  // namespace

static void save_impl(const std::string& path, const Bigsi& b, bool mini) {
    Writer w(path);
    w.u64(b.bloom_size); w.u64(b.num_hash); w.u64(b.k_size);
    if (mini) w.u64(b.m_size);
    w.u64(b.colors.size());
    for (auto& kv : b.colors) { w.u64(kv.first); w.str(kv.second); }
    const uint64_t nbits = b.colors.size();
    const uint32_t W = b.row_words;
    if (W != (nbits + 31) / 32) throw Error("save_bigsi: row width does not match the number of colours");
    w.u64(b.row_ids.size());
    {   // rows in blocks: one record is {u64 row, u64 n_words, n_words x u32, u64 nbits}
        const size_t rec = 24 + (size_t)W * 4, block = 1 << 16;
        std::vector<char> out(rec * block);
        for (size_t i0 = 0; i0 < b.row_ids.size(); i0 += block) {
            const size_t n = std::min(block, b.row_ids.size() - i0);
            char* o = out.data();
            const uint64_t w64 = W;
            for (size_t i = i0; i < i0 + n; i++, o += rec) {
                memcpy(o, &b.row_ids[i], 8);
                memcpy(o + 8, &w64, 8);
                memcpy(o + 16, b.words.data() + i * W, (size_t)W * 4);
                memcpy(o + 16 + (size_t)W * 4, &nbits, 8);
            }
            w.flush();
            if (fwrite(out.data(), 1, rec * n, w.f) != rec * n) throw Error("problems preparing serialized data for writing");
        }
    }
    w.u64(b.n_ref_kmers.size());
    for (auto& kv : b.n_ref_kmers) { w.str(kv.first); w.u64(kv.second); }
    w.flush();
}

void save_bigsi(const std::string& path, const Bigsi& b) { save_impl(path, b, false); }
void save_bigsi_mini(const std::string& path, const Bigsi& b) { save_impl(path, b, true); }

static Bigsi read_impl(const std::string& path, bool mini) {
    Reader r(path);
    Bigsi b;
    b.bloom_size = r.u64(); b.num_hash = r.u64(); b.k_size = r.u64();
    if (mini) { b.m_size = r.u64(); b.mini = true; }
    const uint64_t nc = r.u64();
    if (nc > (1ull << 32)) throw Error("can't deserialize");
    for (uint64_t i = 0; i < nc; i++) { uint64_t c = r.u64(); b.colors[c] = r.str(); }
    b.row_words = (uint32_t)((nc + 31) / 32);
    const uint64_t nrows = r.u64();
    if (nrows > b.bloom_size) throw Error("can't deserialize");
    b.row_ids.resize(nrows);
    b.words.resize(nrows * b.row_words);
    {   // rows in blocks (a per-field fread costs ~10 ns x 4 x 50 M rows)
        const size_t rec = 24 + (size_t)b.row_words * 4, block = 1 << 16;
        std::vector<char> in(rec * block);
        for (uint64_t i0 = 0; i0 < nrows; i0 += block) {
            const size_t n = (size_t)std::min<uint64_t>(block, nrows - i0);
            r.raw(in.data(), rec * n);
            const char* p = in.data();
            for (uint64_t i = i0; i < i0 + n; i++, p += rec) {
                uint64_t nw, nbits;
                memcpy(&b.row_ids[i], p, 8);
                memcpy(&nw, p + 8, 8);
                if (nw != b.row_words) throw Error("can't deserialize: row width differs from the number of colours");
                memcpy(b.words.data() + i * b.row_words, p + 16, (size_t)nw * 4);
                memcpy(&nbits, p + 16 + (size_t)nw * 4, 8);
                if (nbits != nc) throw Error("can't deserialize: BitVec length differs from the number of colours");
            }
        }
    }
    const uint64_t nr = r.u64();
    for (uint64_t i = 0; i < nr; i++) { std::string a = r.str(); b.n_ref_kmers[a] = r.u64(); }
    return b;
}
Bigsi read_bigsi(const std::string& path) { return read_impl(path, false); }
Bigsi read_bigsi_mini(const std::string& path) { return read_impl(path, true); }

}  // namespace cidh
