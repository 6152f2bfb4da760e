// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of colorid's BIGSI hot path (build / search / read_id), written to follow the
// reference source line by line so the CUDA product can be checked for bit-exact parity.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product path (colorid_b200/) never links, imports or calls it.
//
// The Rust reference cannot be compiled in this image (no cargo/rustc, un-vendored crates), so
// there is no oracle/_ref binary.  Parity status per third-party dependency:
//   xxh3 0.1.1        -> restated from the xxHash spec, pinned against python-xxhash; vs. a real
//                        colorid binary: PARITY UNPINNED
//   fnv + hashbrown   -> restated (hashbrown_emul.hpp): PARITY UNPINNED
//   probability 0.15  -> Binomial::mass restated (Loader saddle point): PARITY UNPINNED
// Everything under src/ of the reference is followed literally; each function cites file:line
// relative to /root/reference.
//
// All strings are std::string k-mers exactly like the reference (no 2-bit packing here).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "hashbrown_emul.hpp"
#include "xxh3_ref.hpp"

namespace orc {

typedef std::unordered_map<std::string, uint64_t> KMap;

// ---------------------------------------------------------------- seq.rs / kmer.rs helpers

// seq.rs:59-64
static inline bool is_good_base(uint8_t c) {
    switch (c) {
        case 'a': case 'c': case 'g': case 't': case 'A': case 'C': case 'G': case 'T': return true;
        default: return false;
    }
}
// seq.rs:66-70
static inline bool has_no_n(const char* s, size_t n) {
    for (size_t i = 0; i < n; i++) if (!is_good_base((uint8_t)s[i])) return false;
    return true;
}
// kmer.rs:847-863
static inline char switch_base(char c) {
    switch (c) {
        case 'a': return 't'; case 'c': return 'g'; case 't': return 'a'; case 'g': return 'c';
        case 'u': return 'a'; case 'n': return 'n';
        case 'A': return 'T'; case 'C': return 'G'; case 'T': return 'A'; case 'G': return 'C';
        case 'U': return 'A'; case 'N': return 'N';
        default: return 'N';
    }
}
// kmer.rs:839-845
static std::string revcomp(const std::string& dna) {
    std::string rc;
    rc.reserve(dna.size());
    for (size_t i = dna.size(); i-- > 0;) rc.push_back(switch_base(dna[i]));
    return rc;
}
static inline std::string to_upper_ascii(std::string s) {
    for (auto& c : s) if (c >= 'a' && c <= 'z') c = (char)(c - 32);
    return s;
}
// byte-wise lexicographic `<` on equal-length slices (Rust str ordering)
static inline bool slice_lt(const char* a, const char* b, size_t n) { return memcmp(a, b, n) < 0; }

// seq.rs:36-56.  Returns false where the reference would panic (qual longer than seq).
static bool qual_mask(const std::string& seq, const std::string& qual, uint8_t off, std::string& out) {
    if (off == 0) { out = seq; return true; }
    uint8_t maxq = (uint8_t)(off + 33);
    out.clear();
    size_t si = 0;
    for (size_t i = 0; i < qual.size(); i++) {
        if (si >= seq.size()) return false;  // .expect("could not get the next nt")
        char nt = seq[si++];
        out.push_back(((uint8_t)qual[i] < maxq) ? 'N' : nt);
    }
    return true;
}

enum KmerMode {
    MODE_FASTA = 0,    // kmer.rs:87-125 kmerize_vector: len>=k, has_no_n, compare raw, then uppercase
    MODE_FASTQ = 1,    // kmer.rs:461-510 / 581-655: len>=k, has_no_n, compare raw, keep case
    MODE_STRING = 2,   // kmer.rs:271-299 kmerize_string: NO has_no_n, uppercase (one seq per call)
    MODE_READSET = 3,  // kmer.rs:221-243 kmerize_vector_skip_n_set: no length guard, keep case
};

// One sequence into a count map. `on_kmer(kmer_string, seq_pos, took_fwd)` is called per accepted
// position in order.  Returns false where the reference would panic (MODE_READSET with
// len < k-1: `l.len() - k + 1` wraps in release builds and the slice goes out of range).
template <class F>
static bool for_each_canonical(const std::string& l, size_t k, size_t d, int mode, F on_kmer) {
    size_t L = l.size();
    if (mode == MODE_READSET) {
        if (L + 1 == k) return true;   // wrapping (len-k)+1 == 0 in release: empty range
        if (L + 1 < k) return false;   // slice index out of range -> panic
    } else if (L < k) {
        return true;                   // kmer.rs:94 `continue` / kmer.rs:279 None handled by caller
    }
    std::string r = revcomp(l);
    for (size_t i = 0; i + k <= L; i += d) {
        const char* f = l.data() + i;
        const char* c = r.data() + (L - (i + k));
        if (mode != MODE_STRING && !has_no_n(f, k)) continue;
        bool fwd = slice_lt(f, c, k);           // `l[i..i+k] < l_r[..]`, else branch takes rc (ties too)
        std::string km(fwd ? f : c, k);
        if (mode == MODE_FASTA || mode == MODE_STRING) km = to_upper_ascii(km);
        on_kmer(km, i, fwd);
    }
    return true;
}

static bool kmap_add_seq(KMap& m, const std::string& l, size_t k, size_t d, int mode) {
    return for_each_canonical(l, k, d, mode, [&](const std::string& km, size_t, bool) { m[km] += 1; });
}

// kmer.rs:971-986 find_minimizer: the byte-wise smallest m-mer among seq[0..m] (forward only!), and for
// i in 1..=len-m both seq[i..i+m] and revcomp(seq)[len-(i+m)..len-i] (= revcomp of seq[i..i+m]).
// `which` (optional) reports the winner as 2*i + (1 if it came from the reverse complement).
static std::string find_minimizer(const std::string& seq, size_t m, size_t* which = nullptr) {
    std::string r = revcomp(seq);
    size_t L = seq.size();
    const char* best = seq.data();
    size_t w = 0;
    for (size_t i = 1; i + m <= L; i++) {
        const char* f = seq.data() + i;
        const char* c = r.data() + (L - (i + m));
        if (slice_lt(f, best, m)) { best = f; w = 2 * i; }
        if (slice_lt(c, best, m)) { best = c; w = 2 * i + 1; }
    }
    if (which) *which = w;
    return std::string(best, m);
}

// Minimizer count maps.  `upper`: kmer.rs:328-361 minimerize_vector_skip_n (FASTA, build_multi_mini) upper-cases the
// minimizer AFTER it was chosen on the raw-case canonical k-mer; kmer.rs:694-824 kmers_fq_pe_minimizer_qual /
// kmers_from_fq_minimizer_qual keep the case.  Both: len >= k guard, has_no_n on the forward window, canonical by raw bytes.
static void minimap_add_seq(KMap& mm, const std::string& l, size_t k, size_t m, size_t d, bool upper) {
    size_t L = l.size();
    if (L < k) return;
    std::string r = revcomp(l);
    for (size_t i = 0; i + k <= L; i += d) {
        const char* f = l.data() + i;
        const char* c = r.data() + (L - (i + k));
        if (!has_no_n(f, k)) continue;
        std::string mi = find_minimizer(std::string(slice_lt(f, c, k) ? f : c, k), m);
        mm[upper ? to_upper_ascii(mi) : mi] += 1;
    }
}

// kmer.rs:826-837
static void clean_map(KMap& m, uint64_t t) {
    for (auto it = m.begin(); it != m.end();) {
        if (it->second > t) ++it; else it = m.erase(it);
    }
}

// kmer.rs:866-942.  Returns -1 where the reference would panic (index/underflow on a degenerate
// histogram).  `histo` maps count -> number of distinct k-mers with that count.
static int64_t auto_cutoff_from_histo(const std::map<uint64_t, uint64_t>& histo, uint64_t n_distinct) {
    uint64_t max_cov = 0, sum = 0;
    for (auto& kv : histo) { if (kv.first > max_cov) max_cov = kv.first; sum += kv.first * kv.second; }
    double total_mean = (double)sum / (double)n_distinct;   // 0/0 = NaN -> `NaN < 1.5` false
    if (total_mean < 1.5) return 0;
    std::vector<uint64_t> coverages;
    for (uint64_t c = 1; c < max_cov; c++) {
        auto it = histo.find(c);
        coverages.push_back(it == histo.end() ? 0 : it->second);
    }
    if (coverages.size() < 1) return -1;         // `coverages.len() - 1` underflow
    std::vector<double> d1, d2;
    for (size_t i = 1; i + 1 < coverages.size(); i++)
        d1.push_back((double)coverages[i] / (double)coverages[i + 1]);
    if (d1.size() < 1) return -1;                // `d1.len() - 1` underflow
    for (size_t i = 0; i + 1 < d1.size(); i++) d2.push_back(d1[i] / d1[i + 1]);
    size_t first_pos_d1 = 0, first_pos_d2 = 0;
    for (size_t i = 0; i < d1.size(); i++) if (d1[i] < 1.0) { first_pos_d1 = i + 1; break; }
    for (size_t i = 0; i < d2.size(); i++) if (d2[i] < 1.0) { first_pos_d2 = i + 1; break; }
    uint64_t bigsum = 0, num = 0;
    for (size_t i = 0; i + 1 < coverages.size(); i++) { bigsum += i * coverages[1 + i]; num += coverages[1 + i]; }
    double mean = (double)bigsum / (double)num;
    if (first_pos_d1 > 0 && (double)first_pos_d1 < mean * 0.75) return (int64_t)first_pos_d1;
    if (first_pos_d2 > 0) return (int64_t)first_pos_d2;
    double h = std::ceil(mean / 2.0);
    // `(mean / 2.0).ceil() as usize`: NaN -> 0 (saturating cast), then max(1, _)
    uint64_t hv = (h != h) ? 0 : (uint64_t)h;
    return (int64_t)std::max<uint64_t>(1, hv);
}
static int64_t auto_cutoff(const KMap& m) {
    std::map<uint64_t, uint64_t> histo;
    for (auto& kv : m) histo[kv.second] += 1;
    return auto_cutoff_from_histo(histo, m.size());
}

// ---------------------------------------------------------------- BitVec<u32> (bit-vec_serde/src/lib.rs)

// lib.rs:218-224 storage Vec<u32> + nbits; get :465-474 / set :492-500 are LSB-first within u32
struct BitVec {
    std::vector<uint32_t> w;
    size_t nbits = 0;
    BitVec() {}
    BitVec(size_t n, bool v) : w((n + 31) / 32, v ? 0xFFFFFFFFu : 0u), nbits(n) {  // from_elem :295-303
        if (v && (n % 32)) w.back() &= (1u << (n % 32)) - 1;                          // fix_last_block
    }
    bool get(size_t i) const { return (w[i / 32] >> (i % 32)) & 1; }
    void set(size_t i, bool x) { if (x) w[i / 32] |= 1u << (i % 32); else w[i / 32] &= ~(1u << (i % 32)); }
    bool none() const { for (auto x : w) if (x) return false; return true; }            // :798-800
    void intersect(const BitVec& o) { for (size_t i = 0; i < w.size(); i++) w[i] &= o.w[i]; }  // :598-600
};

// simple_bloom.rs:19-26
static inline uint64_t bloom_bit(const std::string& kmer, uint64_t seed, uint64_t bloom_size, uint32_t variant) {
    return xxh3_64_with_seed((const uint8_t*)kmer.data(), kmer.size(), seed, variant) % bloom_size;
}

// ---------------------------------------------------------------- index model (bigsi.rs:19-27)

struct Index {
    uint64_t S; uint32_t H, k, N, W;
    uint32_t m = 0;                             // bigsi.rs:40-49 BigsyMapMiniNew.m_size (0 = k-mer index)
    uint32_t hv = 0;                            // hash variant (xxh3_ref.hpp; 0 = stable XXH3)
    std::vector<uint32_t> rows;                 // dense [S][W]; an all-zero row == absent row (build.rs:123-127)
    std::vector<std::vector<uint32_t>> bitsets; // phase-1 per-colour Bloom bitsets (build.rs:63-67)
    bool row_present(uint64_t r) const {
        const uint32_t* p = &rows[r * W];
        for (uint32_t i = 0; i < W; i++) if (p[i]) return true;
        return false;
    }
};

// ---------------------------------------------------------------- Binomial::mass (probability crate; App. E)

static double stirlerr(double n) {
    static const double S0 = 1.0 / 12.0, S1 = 1.0 / 360.0, S2 = 1.0 / 1260.0, S3 = 1.0 / 1680.0, S4 = 1.0 / 1188.0;
    static const double SFE[16] = {
        0.0, 0.081061466795327258219670264, 0.041340695955409294093822081, 0.0276779256849983391487892927,
        0.020790672103765093111522771, 0.0166446911898211921631948653, 0.013876128823070747998745727,
        0.0118967099458917700950557241, 0.010411265261972096497478567, 0.0092554621827127329177286366,
        0.008330563433362871256469318, 0.0075736754879518407949720242, 0.006942840107209529865664152,
        0.0064089941880042070684396310, 0.005951370112758847735624416, 0.0055547335519628013710386899};
    if (n < 16.0) return SFE[(size_t)n];
    double nn = n * n;
    if (n > 500.0) return (S0 - S1 / nn) / n;
    if (n > 80.0) return (S0 - (S1 - S2 / nn) / nn) / n;
    if (n > 35.0) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}
static double ln_d0(double x, double np) {
    if (std::fabs(x - np) < 0.1 * (x + np)) {
        double s = (x - np) * (x - np) / (x + np);
        double v = (x - np) / (x + np);
        double ej = 2.0 * x * v;
        for (int j = 1;; j++) {
            ej *= v * v;
            double s1 = s + ej / (double)(2 * j + 1);
            if (s1 == s) return s1;
            s = s1;
        }
    }
    return x * std::log(x / np) + np - x;
}
static double binomial_mass(uint64_t n_, double p, uint64_t x_) {
    double q = 1.0 - p;
    if (p == 0.0) return x_ == 0 ? 1.0 : 0.0;
    if (p == 1.0) return x_ == n_ ? 1.0 : 0.0;
    double n = (double)n_;
    if (x_ == 0) return std::exp(n * std::log(q));
    if (x_ == n_) return std::exp(n * std::log(p));
    double x = (double)x_, nmx = n - x;
    double ln_c = stirlerr(n) - stirlerr(x) - stirlerr(nmx) - ln_d0(x, n * p) - ln_d0(nmx, n * q);
    return std::exp(ln_c) * std::sqrt(n / (2.0 * M_PI * x * nmx));
}

// read_id_mt_pe.rs:695-698
static double false_prob(double m, double k, double n) {
    return std::pow(1.0 - std::pow(M_E, -((k * (n + 0.5)) / (m - 1.0))), k);
}
// read_id_mt_pe.rs:168-181
static bool not_fp_significant(uint64_t observations, double p_false, double fp_correct, uint64_t taxon_hits) {
    double critical = (double)observations * p_false;
    double mpf = binomial_mass(observations, p_false, taxon_hits);
    return ((double)taxon_hits < critical) || (((double)taxon_hits > critical) && (mpf >= fp_correct));
}

// ---------------------------------------------------------------- read_id (read_id_mt_pe.rs)

struct ReadKmer { std::string s; uint32_t seq, pos; uint8_t fwd; };

// kmer.rs:221-243 into an order-emulated FnvHashSet<String>.  Returns false on reference panic.
static bool read_kmer_set(const std::vector<std::string>& seqs, size_t k, size_t d, HbPolicy pol,
                          std::vector<ReadKmer>& keys, std::vector<int>& order) {
    HbTable tab(pol);
    keys.clear();
    bool ok = true;
    for (size_t si = 0; si < seqs.size() && ok; si++) {
        ok = for_each_canonical(seqs[si], k, d, MODE_READSET, [&](const std::string& km, size_t pos, bool fwd) {
            uint64_t h = fnv_hash_str(km);
            int idx = (int)keys.size();
            bool fresh = tab.insert(h, idx, [&](int slot_key) { return keys[slot_key].s == km; });
            if (fresh) keys.push_back(ReadKmer{km, (uint32_t)si, (uint32_t)pos, (uint8_t)fwd});
        });
    }
    order = tab.iter_order();
    return ok;
}

// kmer.rs:363-394 minimerize_vector_skip_n_set (read_id with an .mxi index): len >= k guard per mate, has_no_n,
// canonical by raw bytes, find_minimizer on the raw-case canonical k-mer, then to_uppercase; set of minimizers.
// pos/fwd of a key = forward-read window [pos, pos+m) that spells the minimizer (fwd) or its reverse complement.
static bool read_minimizer_set(const std::vector<std::string>& seqs, size_t k, size_t m, size_t d, HbPolicy pol,
                               std::vector<ReadKmer>& keys, std::vector<int>& order) {
    HbTable tab(pol);
    keys.clear();
    for (size_t si = 0; si < seqs.size(); si++) {
        const std::string& l = seqs[si];
        size_t L = l.size();
        if (L < k) continue;
        std::string r = revcomp(l);
        for (size_t i = 0; i + k <= L; i += d) {
            const char* f = l.data() + i;
            const char* c = r.data() + (L - (i + k));
            if (!has_no_n(f, k)) continue;
            bool fwd = slice_lt(f, c, k);
            size_t which = 0;
            std::string mi = to_upper_ascii(find_minimizer(std::string(fwd ? f : c, k), m, &which));
            size_t j = which / 2; bool from_rc = which & 1;
            uint32_t pos = (uint32_t)(i + (fwd ? j : k - m - j));
            uint8_t mfwd = (uint8_t)((!from_rc) == fwd);
            uint64_t h = fnv_hash_str(mi);
            int idx = (int)keys.size();
            bool fresh = tab.insert(h, idx, [&](int slot_key) { return keys[slot_key].s == mi; });
            if (fresh) keys.push_back(ReadKmer{mi, (uint32_t)si, pos, mfwd});
        }
    }
    order = tab.iter_order();
    return true;
}

struct Report {                        // FnvHashMap<usize,usize> with emulated iteration order
    std::vector<uint64_t> keys; std::vector<uint64_t> vals; HbTable tab;
    explicit Report(HbPolicy pol) : tab(pol) {}
    void bump(uint64_t key) {          // *final_report.entry(key).or_insert(0) += 1
        uint64_t h = fnv_hash_usize(key);
        int idx = (int)keys.size();
        bool fresh = tab.entry_or_insert(h, idx, [&](int sk) { return keys[sk] == key; });
        if (fresh) { keys.push_back(key); vals.push_back(1); }
        else for (size_t i = 0; i < keys.size(); i++) if (keys[i] == key) { vals[i]++; break; }
    }
    std::vector<std::pair<uint64_t, uint64_t>> iter() const {
        std::vector<std::pair<uint64_t, uint64_t>> o;
        for (int i : tab.iter_order()) o.push_back({keys[i], vals[i]});
        return o;
    }
};

static void gather_rows(const Index& ix, const std::string& km, std::vector<const uint32_t*>& slices, bool zero_is_absent) {
    // read_id_mt_pe.rs:117-125 / 76-85, batch_search_pe.rs:47-59: stop at the first absent row
    (void)zero_is_absent;  // dense model: absent == all-zero (SURVEY §8a note), same in both variants
    slices.clear();
    for (uint32_t i = 0; i < ix.H; i++) {
        uint64_t bi = bloom_bit(km, i, ix.S, ix.hv);
        if (!ix.row_present(bi)) break;
        slices.push_back(&ix.rows[bi * ix.W]);
    }
}
static void and_rows(const Index& ix, const std::vector<const uint32_t*>& slices, std::vector<uint32_t>& first) {
    first.assign(slices[0], slices[0] + ix.W);                                  // bitwise_and :41-52
    for (size_t j = 1; j < slices.size(); j++) for (uint32_t w = 0; w < ix.W; w++) first[w] &= slices[j][w];
}

// read_id_mt_pe.rs:104-165 (start_sample > 0) and :66-102 (classic, start_sample == 0)
static void search_index(const Index& ix, const std::vector<ReadKmer>& keys, const std::vector<int>& order,
                         size_t start_sample, Report& final_report) {
    std::vector<uint64_t> cand;          // `report` FnvHashSet<usize>; only membership matters
    std::vector<uint8_t> in_cand(ix.N, 0);
    std::vector<const uint32_t*> slices;
    std::vector<uint32_t> first;
    size_t counter = 0;
    for (int ki : order) {
        gather_rows(ix, keys[ki].s, slices, start_sample == 0);
        if (slices.size() < ix.H) { final_report.bump(ix.N); break; }
        and_rows(ix, slices, first);
        if (start_sample == 0 || counter < start_sample) {
            for (uint32_t c = 0; c < ix.N; c++) if ((first[c / 32] >> (c % 32)) & 1) {
                if (start_sample != 0 && !in_cand[c]) { in_cand[c] = 1; cand.push_back(c); }
                final_report.bump(c);
            }
        } else {
            for (uint64_t c : cand) if ((first[c / 32] >> (c % 32)) & 1) final_report.bump(c);
        }
        counter++;
    }
}

enum ClassKind { CLS_TOO_SHORT = 0, CLS_NO_HITS = 1, CLS_NO_SIG = 2, CLS_ACCEPT = 3, CLS_REJECT_MULTI = 4, CLS_PANIC = 5 };

struct Classification { int kind; uint64_t hits, n_set, n_top; std::vector<uint64_t> top; };

// read_id_mt_pe.rs:187-251
static Classification kmer_poll_plus(const std::vector<std::pair<uint64_t, uint64_t>>& report_iter, uint64_t kmer_length,
                                     const std::vector<double>& child_fp, uint64_t no_hits_num, double fp_correct) {
    std::vector<std::pair<uint64_t, uint64_t>> cv = report_iter;
    std::stable_sort(cv.begin(), cv.end(), [](auto& a, auto& b) { return a.second > b.second; });  // sort_by b.1.cmp(a.1)
    Classification c{CLS_NO_HITS, 0, kmer_length, 0, {}};
    if (cv[0].first == no_hits_num && cv.size() == 1) return c;
    std::vector<std::pair<uint64_t, uint64_t>> sig;
    for (auto& t : cv) {
        if (t.first == no_hits_num) continue;
        if (not_fp_significant(kmer_length, child_fp[t.first], fp_correct, t.second)) continue;
        sig.push_back(t);
    }
    if (sig.empty()) { c.kind = CLS_NO_SIG; return c; }
    for (auto& h : sig) if (h.second == sig[0].second) c.top.push_back(h.first);
    c.hits = sig[0].second;
    c.n_top = c.top.size();
    c.kind = c.top.size() == 1 ? CLS_ACCEPT : CLS_REJECT_MULTI;
    return c;
}

}  // namespace orc

// ================================================================== extern "C" surface (ctypes)
using namespace orc;

static std::vector<std::string> split_seqs(const char* bases, const uint64_t* offs, uint64_t a, uint64_t b) {
    std::vector<std::string> v;
    for (uint64_t i = a; i < b; i++) v.emplace_back(bases + offs[i], offs[i + 1] - offs[i]);
    return v;
}

extern "C" {

uint64_t orc_xxh3_64(const uint8_t* p, uint64_t len, uint64_t seed) { return xxh3_64_with_seed(p, len, seed); }
uint64_t orc_xxh3_64_variant(const uint8_t* p, uint64_t len, uint64_t seed, uint32_t variant) { return xxh3_64_with_seed(p, len, seed, variant); }
uint64_t orc_fnv1a_str(const uint8_t* p, uint64_t len) { return fnv_hash_str(std::string((const char*)p, len)); }
uint64_t orc_fnv1a_usize(uint64_t v) { return fnv_hash_usize(v); }
double orc_binomial_mass(uint64_t n, double p, uint64_t x) { return binomial_mass(n, p, x); }
double orc_false_prob(double m, double k, double n) { return false_prob(m, k, n); }

void orc_revcomp(const char* in, uint64_t n, char* out) {
    std::string r = revcomp(std::string(in, n));
    memcpy(out, r.data(), n);
}
int orc_has_no_n(const char* s, uint64_t n) { return has_no_n(s, n) ? 1 : 0; }
// returns output length, or -1 on reference panic
int64_t orc_qual_mask(const char* seq, uint64_t ns, const char* qual, uint64_t nq, uint8_t off, char* out) {
    std::string o;
    if (!qual_mask(std::string(seq, ns), std::string(qual, nq), off, o)) return -1;
    memcpy(out, o.data(), o.size());
    return (int64_t)o.size();
}

// FnvHashSet<String> insertion sequence -> iteration order (indices of first occurrences).
int64_t orc_hashset_str_order(const char* keys, const uint64_t* offs, uint64_t n, int group_width, int reserve_before_find,
                              int32_t* order_out, uint64_t* buckets_out) {
    HbPolicy pol; pol.group_width = group_width; pol.reserve_before_find = reserve_before_find != 0;
    HbTable tab(pol);
    std::vector<std::string> ks;
    for (uint64_t i = 0; i < n; i++) ks.emplace_back(keys + offs[i], offs[i + 1] - offs[i]);
    for (uint64_t i = 0; i < n; i++)
        tab.insert(fnv_hash_str(ks[i]), (int)i, [&](int sk) { return ks[sk] == ks[i]; });
    auto o = tab.iter_order();
    for (size_t i = 0; i < o.size(); i++) order_out[i] = o[i];
    if (buckets_out) *buckets_out = tab.buckets();
    return (int64_t)o.size();
}
// FnvHashMap<usize,_> via entry().or_insert(): iteration order (indices of first occurrences).
int64_t orc_hashmap_usize_order(const uint64_t* keys, uint64_t n, int group_width, int32_t* order_out) {
    HbPolicy pol; pol.group_width = group_width;
    HbTable tab(pol);
    for (uint64_t i = 0; i < n; i++)
        tab.entry_or_insert(fnv_hash_usize(keys[i]), (int)i, [&](int sk) { return keys[sk] == keys[i]; });
    auto o = tab.iter_order();
    for (size_t i = 0; i < o.size(); i++) order_out[i] = o[i];
    return (int64_t)o.size();
}

// ---- k-mer count maps -------------------------------------------------------------------
struct orc_kmap { KMap m; };
orc_kmap* orc_kmap_new() { return new orc_kmap(); }
void orc_kmap_free(orc_kmap* h) { delete h; }
// returns 0, or -1 where the reference would panic
int orc_kmap_add(orc_kmap* h, const char* bases, const uint64_t* offs, uint64_t nseq, uint32_t k, uint32_t d, int mode) {
    for (uint64_t i = 0; i < nseq; i++)
        if (!kmap_add_seq(h->m, std::string(bases + offs[i], offs[i + 1] - offs[i]), k, d, mode)) return -1;
    return 0;
}
uint64_t orc_kmap_len(orc_kmap* h) { return h->m.size(); }
int64_t orc_kmap_auto_cutoff(orc_kmap* h) { return auto_cutoff(h->m); }
void orc_kmap_clean(orc_kmap* h, uint64_t t) { clean_map(h->m, t); }
// keys sorted bytewise; keys_out is len*k bytes
void orc_kmap_export(orc_kmap* h, uint32_t k, char* keys_out, uint64_t* counts_out) {
    std::vector<std::pair<std::string, uint64_t>> v(h->m.begin(), h->m.end());
    std::sort(v.begin(), v.end());
    for (size_t i = 0; i < v.size(); i++) { memcpy(keys_out + i * k, v[i].first.data(), k); counts_out[i] = v[i].second; }
}
int64_t orc_auto_cutoff_histo(const uint64_t* cov, const uint64_t* num, uint64_t n) {
    std::map<uint64_t, uint64_t> h; uint64_t distinct = 0;
    for (uint64_t i = 0; i < n; i++) { h[cov[i]] += num[i]; distinct += num[i]; }
    return auto_cutoff_from_histo(h, distinct);
}

// ---- index + build (build.rs:33-130 / 132-256) -------------------------------------------
struct orc_index { Index ix; };
orc_index* orc_index_new(uint64_t S, uint32_t H, uint32_t k, uint32_t N) {
    orc_index* h = new orc_index();
    h->ix.S = S; h->ix.H = H; h->ix.k = k; h->ix.N = N; h->ix.W = (N + 31) / 32;
    h->ix.rows.assign(S * h->ix.W, 0);
    h->ix.bitsets.resize(N);
    return h;
}
void orc_index_free(orc_index* h) { delete h; }
uint32_t* orc_index_words(orc_index* h) { return h->ix.rows.data(); }
uint32_t orc_index_row_words(orc_index* h) { return h->ix.W; }

// Phase 1 for one accession: count map -> (auto_cutoff) -> clean_map -> n_ref_kmers -> Bloom bitset.
// mode 0 = FASTA (build.rs:84-98), 1 = FASTQ already quality-masked (build.rs:54-83).
// cutoff -1: FASTA -> no filter; FASTQ -> auto_cutoff.  Returns 0, -1 on reference panic.
int orc_build_accession(orc_index* h, uint32_t colour, const char* bases, const uint64_t* offs, uint64_t nseq,
                        int mode, int64_t cutoff, uint64_t* n_ref_kmers, int64_t* cutoff_used) {
    Index& ix = h->ix;
    KMap m;
    for (uint64_t i = 0; i < nseq; i++)
        kmap_add_seq(m, std::string(bases + offs[i], offs[i + 1] - offs[i]), ix.k, 1, mode == 0 ? MODE_FASTA : MODE_FASTQ);
    int64_t used = -1;
    if (mode == 0) {
        if (cutoff != -1) { used = cutoff; clean_map(m, (uint64_t)cutoff); }
    } else {
        if (cutoff == -1) { used = auto_cutoff(m); if (used < 0) return -1; }
        else used = cutoff;
        clean_map(m, (uint64_t)used);
    }
    if (n_ref_kmers) *n_ref_kmers = m.size();
    if (cutoff_used) *cutoff_used = used;
    std::vector<uint32_t>& bits = ix.bitsets[colour];
    bits.assign((ix.S + 31) / 32, 0);
    for (auto& kv : m)
        for (uint32_t i = 0; i < ix.H; i++) {
            uint64_t b = bloom_bit(kv.first, i, ix.S, ix.hv);
            bits[b / 32] |= 1u << (b % 32);
        }
    return 0;
}
// Minimizer indexes (.mxi).  variant 0 = build.rs:396-492 build_single_mini: the k-mer count map and filter of the
// plain build, then BloomFilter::insert(find_minimizer(kmer, m)) per surviving k-mer; n_ref_kmers = distinct k-mers
// (the reference records it for FASTA accessions only, :450).  variant 1 = build.rs:258-394 build_multi_mini: a
// count map of MINIMIZERS (one count per k-mer position), auto_cutoff / clean_map on those counts, Bloom insert
// of the surviving minimizers, n_ref_kmers = their number (:316,337,352).
int orc_build_accession_mini(orc_index* h, uint32_t colour, const char* bases, const uint64_t* offs, uint64_t nseq,
                             int mode, int64_t cutoff, int variant, uint64_t* n_ref_kmers, int64_t* cutoff_used) {
    Index& ix = h->ix;
    if (ix.m == 0 || ix.m > ix.k) return -1;          // find_minimizer slices seq[..m]: panics for m > k
    KMap m;
    for (uint64_t i = 0; i < nseq; i++) {
        std::string l(bases + offs[i], offs[i + 1] - offs[i]);
        if (variant == 0) kmap_add_seq(m, l, ix.k, 1, mode == 0 ? MODE_FASTA : MODE_FASTQ);
        else minimap_add_seq(m, l, ix.k, ix.m, 1, mode == 0);
    }
    int64_t used = -1;
    if (mode == 0) {
        if (cutoff != -1) { used = cutoff; clean_map(m, (uint64_t)cutoff); }
    } else {
        if (cutoff == -1) { used = auto_cutoff(m); if (used < 0) return -1; }
        else used = cutoff;
        clean_map(m, (uint64_t)used);
    }
    if (n_ref_kmers) *n_ref_kmers = m.size();
    if (cutoff_used) *cutoff_used = used;
    std::vector<uint32_t>& bits = ix.bitsets[colour];
    bits.assign((ix.S + 31) / 32, 0);
    for (auto& kv : m) {
        const std::string item = variant == 0 ? find_minimizer(kv.first, ix.m) : kv.first;
        for (uint32_t i = 0; i < ix.H; i++) {
            uint64_t b = bloom_bit(item, i, ix.S, ix.hv);
            bits[b / 32] |= 1u << (b % 32);
        }
    }
    return 0;
}
void orc_index_set_minimizer(orc_index* h, uint32_t m) { h->ix.m = m; }
void orc_index_set_hash_variant(orc_index* h, uint32_t v) { h->ix.hv = v; }
// kmer.rs:971-986; out has m bytes.  Returns -1 where the reference panics (m > len).
int orc_find_minimizer(const char* seq, uint64_t n, uint32_t m, char* out) {
    if (m > n || m == 0) return -1;
    std::string r = find_minimizer(std::string(seq, n), m);
    memcpy(out, r.data(), m);
    return 0;
}
// Minimizer count map of a list of sequences (kmer.rs:328-361 with upper = 1, :694-824 with upper = 0) into an orc_kmap.
int orc_kmap_add_minimizers(orc_kmap* h, const char* bases, const uint64_t* offs, uint64_t nseq, uint32_t k, uint32_t m,
                            uint32_t d, int upper) {
    if (m == 0 || m > k) return -1;
    for (uint64_t i = 0; i < nseq; i++)
        minimap_add_seq(h->m, std::string(bases + offs[i], offs[i + 1] - offs[i]), k, m, d, upper != 0);
    return 0;
}

// Phase 2 transposition (build.rs:116-128): row i gets bit `colour` iff that accession's bit i is set.
void orc_build_finalize(orc_index* h, int threads) {
    Index& ix = h->ix;
    if (threads < 1) threads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([&, t]() {
            uint64_t lo = ix.S * t / threads, hi = ix.S * (t + 1) / threads;
            for (uint64_t i = lo; i < hi; i++)
                for (uint32_t c = 0; c < ix.N; c++) {
                    const auto& b = ix.bitsets[c];
                    if (!b.empty() && ((b[i / 32] >> (i % 32)) & 1)) ix.rows[i * ix.W + c / 32] |= 1u << (c % 32);
                }
        });
    for (auto& x : th) x.join();
    for (auto& b : ix.bitsets) { std::vector<uint32_t>().swap(b); }
}
uint64_t orc_index_nonzero_rows(orc_index* h) {
    uint64_t n = 0;
    for (uint64_t r = 0; r < h->ix.S; r++) n += h->ix.row_present(r);
    return n;
}

// ---- search (batch_search_pe.rs:9-179) ----------------------------------------------------
// One query = sequences [query_offs[q], query_offs[q+1]).  seq_mode 0 = FASTA file, 1 = FASTQ (masked).
// Outputs per query: counts[N], num_kmers, and for the default report the unique-hit multiplicities
// summarised per accession: uniq_n, uniq_sum (exact integer sum of multiplicities), uniq_mode
// (smallest value among the most frequent; the reference's tie-break is SipHash-random).
int orc_query_counts(orc_index* h, const char* bases, const uint64_t* seq_offs, const uint64_t* query_offs, uint64_t nq,
                     int seq_mode, int gene_search, int64_t filter, uint32_t* counts, uint64_t* num_kmers,
                     uint64_t* uniq_n, uint64_t* uniq_sum, uint64_t* uniq_mode, int64_t* cutoff_used) {
    Index& ix = h->ix;
    std::vector<const uint32_t*> slices;
    std::vector<uint32_t> first;
    for (uint64_t q = 0; q < nq; q++) {
        KMap m;
        for (uint64_t i = query_offs[q]; i < query_offs[q + 1]; i++)
            kmap_add_seq(m, std::string(bases + seq_offs[i], seq_offs[i + 1] - seq_offs[i]), ix.k, 1,
                         seq_mode == 0 ? MODE_FASTA : MODE_FASTQ);
        int64_t used;
        if (seq_mode == 0 && gene_search) used = 0;                       // :112-113
        else if (filter < 0) { used = auto_cutoff(m); if (used < 0) return -1; }  // :34-36 / :114-117
        else used = filter;
        clean_map(m, (uint64_t)used);
        if (cutoff_used) cutoff_used[q] = used;
        num_kmers[q] = m.size();
        uint32_t* cnt = counts + q * ix.N;
        std::fill(cnt, cnt + ix.N, 0u);
        std::vector<std::map<uint64_t, uint64_t>> freq(ix.N);
        for (auto& kv : m) {
            gather_rows(ix, kv.first, slices, false);
            if (slices.size() < ix.H) continue;                              // :57-58
            and_rows(ix, slices, first);
            uint32_t nh = 0, last = 0;
            for (uint32_t c = 0; c < ix.N; c++) if ((first[c / 32] >> (c % 32)) & 1) { cnt[c]++; nh++; last = c; }
            if (nh == 1) freq[last][kv.second] += 1;                         // :75-82
        }
        for (uint32_t c = 0; c < ix.N; c++) {
            uint64_t n = 0, s = 0, mode = 0, best = 0;
            for (auto& fv : freq[c]) { n += fv.second; s += fv.first * fv.second; if (fv.second > best) { best = fv.second; mode = fv.first; } }
            if (uniq_n) uniq_n[q * ix.N + c] = n;
            if (uniq_sum) uniq_sum[q * ix.N + c] = s;
            if (uniq_mode) uniq_mode[q * ix.N + c] = mode;
        }
    }
    return 0;
}

// ---- perfect search (perfect_search.rs:6-60 batch_search; :62-120 batch_search_mf) --------
// mf == 0: one query = sequence group (a FASTA file) through kmerize_vector;
// mf == 1: one query = one sequence through kmerize_string (query_offs ignored, nq = nseq).
// status: 0 = AND computed (and_rows valid), 1 = "No perfect hits!" (some row absent),
//         2 = no k-mers (len==0 warning / kmerize_string None).
int orc_query_perfect(orc_index* h, const char* bases, const uint64_t* seq_offs, const uint64_t* query_offs, uint64_t nq,
                      int mf, uint32_t* and_out, uint8_t* status, uint64_t* n_kmers) {
    Index& ix = h->ix;
    for (uint64_t q = 0; q < nq; q++) {
        KMap m;
        bool none = false;
        if (mf) {
            std::string s(bases + seq_offs[q], seq_offs[q + 1] - seq_offs[q]);
            if (s.size() < ix.k) none = true; else kmap_add_seq(m, s, ix.k, 1, MODE_STRING);
        } else {
            for (uint64_t i = query_offs[q]; i < query_offs[q + 1]; i++)
                kmap_add_seq(m, std::string(bases + seq_offs[i], seq_offs[i + 1] - seq_offs[i]), ix.k, 1, MODE_FASTA);
            if (m.empty()) none = true;
        }
        n_kmers[q] = m.size();
        uint32_t* out = and_out + q * ix.W;
        std::fill(out, out + ix.W, 0u);
        if (none) { status[q] = 2; continue; }
        bool missing = false;
        std::vector<uint32_t> acc(ix.W, 0xFFFFFFFFu);
        for (auto& kv : m)
            for (uint32_t i = 0; i < ix.H; i++) {
                uint64_t bi = bloom_bit(kv.first, i, ix.S, ix.hv);
                if (!ix.row_present(bi)) { missing = true; break; }
                for (uint32_t w = 0; w < ix.W; w++) acc[w] &= ix.rows[bi * ix.W + w];
            }
        if (missing) { status[q] = 1; continue; }
        status[q] = 0;
        std::copy(acc.begin(), acc.end(), out);
    }
    return 0;
}

// ---- read_id (read_id_mt_pe.rs:282-363 parallel_vec) --------------------------------------
// One read = sequences [read_offs[r], read_offs[r+1]) (1 or 2 mates, already quality-masked).
// Per read outputs:
//   n_set, cls_kind (ClassKind), hits, n_top, top[top_cap] (colours of the top hits in report order),
//   rep_n + rep_colour/rep_count[rep_cap]: final_report in emulated FnvHashMap iteration order
//   (colour == N is the "no hit" key), order_n + order_seq/order_pos[order_cap]: the k-mer set in
//   emulated iteration order as (mate index, position of first occurrence).
int orc_read_id_batch(orc_index* h, const char* bases, const uint64_t* seq_offs, const uint64_t* read_offs, uint64_t nreads,
                      uint32_t d, uint32_t start_sample, const uint64_t* n_ref_by_colour, double fp_correct,
                      int group_width, int reserve_before_find, int threads,
                      uint32_t* n_set, int32_t* cls_kind, uint32_t* hits, uint32_t* n_top,
                      uint32_t* top, uint32_t top_cap,
                      uint32_t* rep_n, uint32_t* rep_colour, uint32_t* rep_count, uint32_t rep_cap,
                      uint32_t* order_n, uint8_t* order_seq, uint32_t* order_pos, uint32_t order_cap) {
    Index& ix = h->ix;
    HbPolicy pol; pol.group_width = group_width; pol.reserve_before_find = reserve_before_find != 0;
    HbPolicy pol_entry; pol_entry.group_width = group_width;
    std::vector<double> fp(ix.N);                                               // false_prob_map :18-38
    for (uint32_t c = 0; c < ix.N; c++) fp[c] = false_prob((double)ix.S, (double)ix.H, (double)n_ref_by_colour[c]);
    if (threads < 1) threads = 1;
    std::atomic<uint64_t> next(0);
    auto work = [&]() {
        std::vector<ReadKmer> keys; std::vector<int> order;
        for (;;) {
            uint64_t r0 = next.fetch_add(256);
            if (r0 >= nreads) break;
            for (uint64_t r = r0; r < std::min(nreads, r0 + 256); r++) {
                std::vector<std::string> seqs = split_seqs(bases, seq_offs, read_offs[r], read_offs[r + 1]);
                n_set[r] = 0; hits[r] = 0; n_top[r] = 0;
                if (rep_n) rep_n[r] = 0;
                if (order_n) order_n[r] = 0;
                if (seqs.empty() || seqs[0].size() < ix.k) { cls_kind[r] = CLS_TOO_SHORT; continue; }   // :305-313
                const bool set_ok = ix.m ? read_minimizer_set(seqs, ix.k, ix.m, d, pol, keys, order)   // :318-322 `if m == 0`
                                         : read_kmer_set(seqs, ix.k, d, pol, keys, order);
                if (!set_ok) { cls_kind[r] = CLS_PANIC; continue; }
                n_set[r] = (uint32_t)keys.size();
                if (order_n) {
                    order_n[r] = (uint32_t)std::min<size_t>(order.size(), order_cap);
                    for (uint32_t i = 0; i < order_n[r]; i++) {
                        // minimizer mode: bit 7 = the window spells the minimizer itself (1) or its reverse complement (0)
                        order_seq[r * (uint64_t)order_cap + i] = (uint8_t)(keys[order[i]].seq | (ix.m ? (keys[order[i]].fwd << 7) : 0));
                        order_pos[r * (uint64_t)order_cap + i] = (uint32_t)keys[order[i]].pos;
                    }
                }
                Report rep(pol_entry);
                search_index(ix, keys, order, start_sample, rep);
                auto it = rep.iter();
                if (rep_n) {
                    rep_n[r] = (uint32_t)std::min<size_t>(it.size(), rep_cap);
                    for (uint32_t i = 0; i < rep_n[r]; i++) {
                        rep_colour[r * (uint64_t)rep_cap + i] = (uint32_t)it[i].first;
                        rep_count[r * (uint64_t)rep_cap + i] = (uint32_t)it[i].second;
                    }
                }
                if (it.empty()) { cls_kind[r] = CLS_NO_HITS; continue; }                          // :332-340
                Classification c = kmer_poll_plus(it, keys.size(), fp, ix.N, fp_correct);
                cls_kind[r] = c.kind; hits[r] = (uint32_t)c.hits; n_top[r] = (uint32_t)c.n_top;
                for (size_t i = 0; i < c.top.size() && i < top_cap; i++) top[r * (uint64_t)top_cap + i] = (uint32_t)c.top[i];
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back(work);
    for (auto& x : th) x.join();
    return 0;
}

// Host-side classification alone (for checking the product's host/device vote logic):
// report given in FnvHashMap iteration order.
int orc_kmer_poll_plus(const uint32_t* rep_colour, const uint32_t* rep_count, uint32_t rep_n, uint64_t n_set,
                       const double* fp_by_colour, uint64_t N, double fp_correct,
                       int32_t* kind, uint32_t* hits, uint32_t* n_top, uint32_t* top, uint32_t top_cap) {
    std::vector<std::pair<uint64_t, uint64_t>> it;
    for (uint32_t i = 0; i < rep_n; i++) it.push_back({rep_colour[i], rep_count[i]});
    std::vector<double> fp(fp_by_colour, fp_by_colour + N);
    Classification c = kmer_poll_plus(it, n_set, fp, N, fp_correct);
    *kind = c.kind; *hits = (uint32_t)c.hits; *n_top = (uint32_t)c.n_top;
    for (size_t i = 0; i < c.top.size() && i < top_cap; i++) top[i] = (uint32_t)c.top[i];
    return 0;
}

// Multi-threaded build of many accessions (build.rs:167-217 par_iter over accessions), for the
// CPU baseline: accession a = sequences [acc_offs[a], acc_offs[a+1]); colour = a.
int orc_build_many(orc_index* h, const char* bases, const uint64_t* seq_offs, const uint64_t* acc_offs, uint64_t nacc,
                   int mode, int64_t cutoff, int threads, uint64_t* n_ref_kmers) {
    std::atomic<uint64_t> next(0);
    std::atomic<int> err(0);
    if (threads < 1) threads = 1;
    auto work = [&]() {
        for (;;) {
            uint64_t a = next.fetch_add(1);
            if (a >= nacc) break;
            const uint64_t* so = seq_offs + acc_offs[a];
            if (orc_build_accession(h, (uint32_t)a, bases, so, acc_offs[a + 1] - acc_offs[a], mode, cutoff,
                                    n_ref_kmers ? &n_ref_kmers[a] : nullptr, nullptr) != 0) err = 1;
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back(work);
    for (auto& x : th) x.join();
    if (err) return -1;
    orc_build_finalize(h, threads);
    return 0;
}

}  // extern "C"
