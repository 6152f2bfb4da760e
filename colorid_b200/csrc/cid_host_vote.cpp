// Host side of read_id's vote: read_id_mt_pe.rs:187-251 kmer_poll_plus, :168-181 not_fp_signicant,
// :695-698 false_prob, :18-38 false_prob_map.  The device hands over each read's final_report in
// INSERTION order; ties between equally scored accessions are reported in the iteration order of
// the reference's FnvHashMap<usize,usize> (hashbrown), which is reproduced here from that order.
//
// Floating point: f64, same operation order as the Rust source; `probability::Binomial::mass`
// (third-party, not vendored in the reference) is restated as Loader's saddle-point algorithm.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "cid_internal.h"

namespace cid {

// ---- probability::Binomial::mass ----------------------------------------------------------------
static double stirling_err(double n) {
    static const double table[16] = {
        0.0, 0.081061466795327258219670264, 0.041340695955409294093822081, 0.0276779256849983391487892927,
        0.020790672103765093111522771, 0.0166446911898211921631948653, 0.013876128823070747998745727,
        0.0118967099458917700950557241, 0.010411265261972096497478567, 0.0092554621827127329177286366,
        0.008330563433362871256469318, 0.0075736754879518407949720242, 0.006942840107209529865664152,
        0.0064089941880042070684396310, 0.005951370112758847735624416, 0.0055547335519628013710386899};
    if (n < 16.0) return table[(int)n];
    const double n2 = n * n;
    const double s0 = 1.0 / 12.0, s1 = 1.0 / 360.0, s2 = 1.0 / 1260.0, s3 = 1.0 / 1680.0, s4 = 1.0 / 1188.0;
    if (n > 500.0) return (s0 - s1 / n2) / n;
    if (n > 80.0) return (s0 - (s1 - s2 / n2) / n2) / n;
    if (n > 35.0) return (s0 - (s1 - (s2 - s3 / n2) / n2) / n2) / n;
    return (s0 - (s1 - (s2 - (s3 - s4 / n2) / n2) / n2) / n2) / n;
}
static double deviance_term(double x, double np) {
    if (std::fabs(x - np) < 0.1 * (x + np)) {
        double v = (x - np) / (x + np);
        double s = (x - np) * (x - np) / (x + np);
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
double binomial_pmf(uint64_t n_, double p, uint64_t x_) {
    if (p == 0.0) return x_ == 0 ? 1.0 : 0.0;
    if (p == 1.0) return x_ == n_ ? 1.0 : 0.0;
    const double q = 1.0 - p, n = (double)n_;
    if (x_ == 0) return std::exp(n * std::log(q));
    if (x_ == n_) return std::exp(n * std::log(p));
    const double x = (double)x_, rest = n - x;
    const double lc = stirling_err(n) - stirling_err(x) - stirling_err(rest) - deviance_term(x, n * p) - deviance_term(rest, n * q);
    return std::exp(lc) * std::sqrt(n / (2.0 * M_PI * x * rest));
}

// read_id_mt_pe.rs:695-698
double bloom_false_prob(double m, double k, double n) {
    return std::pow(1.0 - std::pow(M_E, -((k * (n + 0.5)) / (m - 1.0))), k);
}

// ---- FnvHashMap<usize,usize> iteration order from the insertion sequence --------------------------
// fnv: h = 0xcbf29ce484222325; for each of the 8 LE bytes: h ^= b; h *= 0x100000001b3.
static inline uint64_t fnv_usize(uint64_t v) {
    uint64_t h = 0xcbf29ce484222325ULL;
    for (int i = 0; i < 8; i++) { h ^= (v >> (8 * i)) & 0xFF; h *= 0x100000001b3ULL; }
    return h;
}
// entry().or_insert(): the table grows only when a NEW key arrives with no growth left.
// Returns the positions (into keys[]) in ascending bucket order.  `hash_of(key)` = FNV-1a of the key.
template <class HashOf>
static void usize_map_order(const uint32_t* keys, uint32_t n, uint32_t gw, std::vector<int32_t>& tab,
                            std::vector<int32_t>& tmp, std::vector<uint32_t>& order, HashOf hash_of) {
    uint32_t nb = 0, items = 0, growth = 0;
    auto cap_of = [](uint32_t b) { return b == 0 ? 0u : (b < 8 ? b - 1 : b / 8 * 7); };
    auto probe = [&](const std::vector<int32_t>& t, uint32_t buckets, uint64_t hash) -> uint32_t {
        const uint32_t mask = buckets - 1;
        uint32_t pos = (uint32_t)hash & mask;
        if (buckets < gw) {
            for (uint32_t b = 0; b < buckets; b++) { uint32_t s = (pos + b) & mask; if (t[s] < 0) return s; }
            return 0;
        }
        for (uint32_t stride = 0;;) {
            for (uint32_t b = 0; b < gw; b++) { uint32_t s = (pos + b) & mask; if (t[s] < 0) return s; }
            stride += gw;
            pos = (pos + stride) & mask;
        }
    };
    for (uint32_t i = 0; i < n; i++) {
        if (growth == 0) {
            uint32_t nn = nb == 0 ? 4 : nb * 2;
            tmp.assign(nn, -1);
            for (uint32_t s = 0; s < nb; s++) if (tab[s] >= 0) tmp[probe(tmp, nn, hash_of(keys[tab[s]]))] = tab[s];
            tab.swap(tmp);
            nb = nn;
            growth = cap_of(nb) - items;
        }
        tab[probe(tab, nb, hash_of(keys[i]))] = (int32_t)i;
        items++; growth--;
    }
    order.clear();
    for (uint32_t s = 0; s < nb; s++) if (tab[s] >= 0) order.push_back((uint32_t)tab[s]);
}

// Reads of one library produce the same few key sequences over and over (the members of a clade in
// ascending colour order, plus the "no hit" key): memoise sequence -> iteration order per thread.
struct OrderMemo {
    enum { SIZE = 1 << 10, MAXN = 12 };
    struct Ent { uint32_t n; uint32_t keys[MAXN]; uint8_t order[MAXN]; };
    std::vector<Ent> ent;
    OrderMemo() : ent(SIZE) { for (auto& e : ent) e.n = 0; }
};

// ---- kmer_poll_plus over a chunk of reads ---------------------------------------------------------
void vote_params_init(VoteParams& vp, uint64_t bloom_size, uint32_t num_hash, uint32_t n_colors,
                      const uint64_t* n_ref_by_colour, double fp_correct, uint32_t group_width) {
    vp.n_colors = n_colors;
    vp.fp_correct = fp_correct;
    vp.group_width = (group_width == 8 || group_width == 16) ? group_width : 16;
    vp.key_hash.resize((size_t)n_colors + 1);
    for (uint32_t c = 0; c <= n_colors; c++) vp.key_hash[c] = fnv_usize(c);
    vp.fp.resize(n_colors);      // false_prob_map, read_id_mt_pe.rs:18-38
    for (uint32_t c = 0; c < n_colors; c++)
        vp.fp[c] = bloom_false_prob((double)bloom_size, (double)num_hash, (double)n_ref_by_colour[c]);
}

// Binomial::mass is a pure function of (observations, p_false[colour], hits); reads of one library
// share a handful of n_set values, so a small per-thread direct-mapped memo removes most of the
// exp/log work without changing a single result bit.
struct PmfMemo {
    enum { SIZE = 1 << 13 };
    std::vector<uint64_t> key;
    std::vector<double> val;
    PmfMemo() : key(SIZE, ~0ull), val(SIZE, 0.0) {}
    double get(const VoteParams& vp, uint64_t obs, uint32_t colour, uint32_t x) {
        if (obs >= (1u << 20) || x >= (1u << 20) || colour >= (1u << 24) - 1) return binomial_pmf(obs, vp.fp[colour], x);   // (the all-ones key is the empty marker)
        const uint64_t k = ((uint64_t)colour << 40) | (obs << 20) | x;
        const uint64_t h = (k * 0x9E3779B97F4A7C15ull) >> (64 - 13);
        if (key[h] != k) { key[h] = k; val[h] = binomial_pmf(obs, vp.fp[colour], x); }
        return val[h];
    }
};

void classify_chunk(const VoteParams& vp, uint64_t nreads, const uint32_t* n_set, const uint32_t* flags,
                    const uint32_t* rep_n, const uint32_t* rep_colour, const uint32_t* rep_count, uint32_t rep_cap,
                    int threads, int32_t* kind, uint32_t* hits, uint32_t* n_top, uint32_t* top, uint32_t top_cap) {
    const uint32_t n_colors = vp.n_colors;
    if (threads < 1) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = (int)std::min<uint64_t>((uint64_t)threads, (nreads + 1023) / 1024);
    std::atomic<uint64_t> next(0);
    auto work = [&]() {
        std::vector<int32_t> tab, tmp;
        std::vector<uint32_t> order;
        std::vector<std::pair<uint32_t, uint32_t>> cv, sig;
        PmfMemo memo;
        OrderMemo omemo;
        auto hash_of = [&](uint32_t key) -> uint64_t { return key <= n_colors ? vp.key_hash[key] : fnv_usize(key); };
        for (;;) {
            uint64_t r0 = next.fetch_add(1024);
            if (r0 >= nreads) break;
            uint64_t r1 = std::min(nreads, r0 + 1024);
            for (uint64_t r = r0; r < r1; r++) {
                hits[r] = 0; n_top[r] = 0;
                if (flags[r] & 1u) { kind[r] = CID_CLS_TOO_SHORT; continue; }          // :305-313
                if (flags[r] & 2u) { kind[r] = CID_CLS_REF_PANIC; continue; }
                const uint32_t n = rep_n[r];
                if (n == 0) { kind[r] = CID_CLS_NO_HITS; continue; }                   // report.is_empty(), :332-340
                const uint32_t* keys = rep_colour + r * (uint64_t)rep_cap;
                const uint32_t* vals = rep_count + r * (uint64_t)rep_cap;
                cv.clear();
                if (n == 1) cv.push_back({keys[0], vals[0]});
                else {
                    const uint8_t* ord8 = nullptr;
                    OrderMemo::Ent* e = nullptr;
                    if (n <= OrderMemo::MAXN) {
                        uint64_t h = 0xcbf29ce484222325ULL ^ n;
                        for (uint32_t i = 0; i < n; i++) h = (h ^ keys[i]) * 0x100000001b3ULL;
                        e = &omemo.ent[(h ^ (h >> 29)) & (OrderMemo::SIZE - 1)];
                        if (e->n == n && !memcmp(e->keys, keys, n * 4)) ord8 = e->order;
                    }
                    if (ord8) {
                        for (uint32_t i = 0; i < n; i++) cv.push_back({keys[ord8[i]], vals[ord8[i]]});
                    } else {
                        usize_map_order(keys, n, vp.group_width, tab, tmp, order, hash_of);
                        for (uint32_t i : order) cv.push_back({keys[i], vals[i]});
                        if (e) { e->n = n; memcpy(e->keys, keys, n * 4); for (uint32_t i = 0; i < n; i++) e->order[i] = (uint8_t)order[i]; }
                    }
                    // stable sort by count, descending (read_id_mt_pe.rs:196 sort_by(|a, b| b.1.cmp(a.1)))
                    for (size_t i = 1; i < cv.size(); i++) {
                        auto x = cv[i];
                        size_t j = i;
                        while (j > 0 && cv[j - 1].second < x.second) { cv[j] = cv[j - 1]; j--; }
                        cv[j] = x;
                    }
                }
                if (cv[0].first == n_colors && cv.size() == 1) { kind[r] = CID_CLS_NO_HITS; continue; }   // :197-205
                const uint64_t observations = n_set[r];
                sig.clear();
                for (auto& t : cv) {
                    if (t.first == n_colors) continue;
                    const double p_false = vp.fp[t.first];
                    const double critical = (double)observations * p_false;
                    const double th = (double)t.second;
                    bool drop = th < critical;
                    if (!drop && th > critical) drop = memo.get(vp, observations, t.first, t.second) >= vp.fp_correct;
                    if (!drop) sig.push_back(t);
                }
                if (sig.empty()) { kind[r] = CID_CLS_NO_SIGNIFICANT; continue; }
                uint32_t nt = 0;
                for (auto& h : sig) if (h.second == sig[0].second) { if (top && nt < top_cap) top[r * (uint64_t)top_cap + nt] = h.first; nt++; }
                hits[r] = sig[0].second;
                n_top[r] = nt;
                kind[r] = nt == 1 ? CID_CLS_ACCEPT : CID_CLS_REJECT_MULTI;
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(work);
    work();
    for (auto& x : th) x.join();
}

}  // namespace cid

using namespace cid;

extern "C" {

double cid_false_prob(double bloom_size, double num_hash, double n_ref_kmers) {
    return bloom_false_prob(bloom_size, num_hash, n_ref_kmers);
}
double cid_binomial_mass(uint64_t n, double p, uint64_t x) { return binomial_pmf(n, p, x); }

int cid_classify_reads(uint64_t bloom_size, uint32_t num_hash, uint32_t n_colors, const uint64_t* n_ref_by_colour,
                       double fp_correct, uint32_t group_width, uint64_t nreads, const uint32_t* n_set,
                       const uint32_t* flags, const uint32_t* rep_n, const uint32_t* rep_colour,
                       const uint32_t* rep_count, uint32_t rep_cap, int threads, int32_t* kind, uint32_t* hits,
                       uint32_t* n_top, uint32_t* top, uint32_t top_cap) {
    if (!n_ref_by_colour || !n_set || !flags || !rep_n || !rep_colour || !rep_count || !kind || !hits || !n_top) {
        set_error("cid_classify_reads: null argument");
        return CID_E_INVALID;
    }
    VoteParams vp;
    vote_params_init(vp, bloom_size, num_hash, n_colors, n_ref_by_colour, fp_correct, group_width);
    classify_chunk(vp, nreads, n_set, flags, rep_n, rep_colour, rep_count, rep_cap, threads, kind, hits, n_top, top, top_cap);
    return CID_OK;
}

// Column-sharded read_id: every shard classifies the same reads against its accession slice with the option
// readid_report_steps set, so each report entry is colour | step << 20 (step = index, in set order, of the k-mer whose AND
// row inserted the colour into final_report, read_id_mt_pe.rs:131-136).  Colours are only ever inserted by ascending k-mer
// and, within a k-mer, by ascending accession, so sorting the union of the shards' entries by (step, global colour) restores
// the insertion order of the unsharded run; counts are per colour and need no exchange.  The first-miss position is a
// property of whole rows (row-present bitmap OR-ed across shards), so every shard reports the "no hit" key or none does.
int cid_merge_shard_reports(uint32_t n_shards, const uint32_t* shard_n_colors, const uint32_t* shard_col_offset,
                            uint64_t nreads, const uint32_t* const* rep_n, const uint32_t* const* rep_colour,
                            const uint32_t* const* rep_count, uint32_t rep_cap_in, uint32_t n_total, uint32_t* out_rep_n,
                            uint32_t* out_colour, uint32_t* out_count, uint32_t rep_cap_out, uint32_t* out_flags) {
    if (!n_shards || !shard_n_colors || !shard_col_offset || !rep_n || !rep_colour || !rep_count || !out_rep_n || !out_colour ||
        !out_count) {
        set_error("cid_merge_shard_reports: null argument");
        return CID_E_INVALID;
    }
    const uint32_t cmask = (1u << 20) - 1u;
    for (uint32_t s = 0; s < n_shards; s++)        // colours share a word with the 12-bit insertion step (colour | step << 20)
        if (shard_n_colors[s] >= cmask) { set_error("cid_merge_shard_reports: shard %u has %u colours, the tagged reports hold < 2^20 - 1", s, shard_n_colors[s]); return CID_E_INVALID; }
    struct Ent { uint32_t step, colour, count; };
    // reads are independent: host threads over contiguous ranges; the first problem found wins (1 = rep_n, 2 = colour, 3 = miss)
    std::atomic<int> problem{0};
    std::atomic<unsigned long long> problem_read{0};
    auto work = [&](uint64_t r_lo, uint64_t r_hi) {
        std::vector<Ent> ents;
        for (uint64_t r = r_lo; r < r_hi && !problem.load(std::memory_order_relaxed); r++) {
            ents.clear();
            uint32_t n_miss = 0;
            int bad = 0;
            for (uint32_t s = 0; s < n_shards && !bad; s++) {
                const uint32_t n = rep_n[s][r];
                if (n > rep_cap_in) { bad = 1; break; }
                for (uint32_t i = 0; i < n; i++) {
                    const uint32_t e = rep_colour[s][r * (uint64_t)rep_cap_in + i], c = e & cmask;
                    if (c == shard_n_colors[s]) { n_miss++; continue; }          // the "no hit" key N of that shard
                    if (c > shard_n_colors[s]) { bad = 2; break; }
                    ents.push_back({e >> 20, shard_col_offset[s] + c, rep_count[s][r * (uint64_t)rep_cap_in + i]});
                }
            }
            if (!bad && n_miss != 0 && n_miss != n_shards) bad = 3;
            if (bad) {
                int expected = 0;
                if (problem.compare_exchange_strong(expected, bad)) problem_read = r;
                return;
            }
            std::sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) { return a.step != b.step ? a.step < b.step : a.colour < b.colour; });
            const uint64_t total = ents.size() + (n_miss ? 1 : 0);
            uint32_t* oc = out_colour + r * (uint64_t)rep_cap_out;
            uint32_t* ov = out_count + r * (uint64_t)rep_cap_out;
            uint32_t w = 0;
            for (const Ent& e : ents) { if (w < rep_cap_out) { oc[w] = e.colour; ov[w] = e.count; w++; } }
            if (n_miss && w < rep_cap_out) { oc[w] = n_total; ov[w] = 1; w++; }
            out_rep_n[r] = w;
            if (out_flags && total > rep_cap_out) out_flags[r] |= 4u;              // truncated, like the kernels' flag bit 2
        }
    };
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    nt = (unsigned)std::min<uint64_t>(nt, std::max<uint64_t>(1, nreads / 2048));
    if (nt <= 1) work(0, nreads);
    else {
        std::vector<std::thread> th;
        const uint64_t per = (nreads + nt - 1) / nt;
        for (unsigned t = 0; t < nt; t++) th.emplace_back(work, std::min(nreads, t * per), std::min(nreads, (t + 1) * per));
        for (auto& x : th) x.join();
    }
    switch (problem.load()) {
        case 1: set_error("cid_merge_shard_reports: rep_n above rep_cap"); return CID_E_INVALID;
        case 2: set_error("cid_merge_shard_reports: colour out of range"); return CID_E_INVALID;
        case 3:
            set_error("cid_merge_shard_reports: shards disagree on the first absent row of read %llu (row-present bitmaps not merged?)",
                      problem_read.load());
            return CID_E_INVALID;
        default: break;
    }
    return CID_OK;
}

}  // extern "C"
