// Report writers of the host layer: reports.rs:8-62 (search reports), :98-120 (read_id counts file).
#include <cmath>

#include "cid_host.hpp"

namespace cidh {

// reports.rs:8-48.  `report` holds only accessions with at least one hit; mean / modus / specific come
// from the multiplicities of the k-mers that hit exactly this one accession (batch_search_pe.rs:75-82).
void generate_report(FILE* out, const std::string& query, const Bigsi& ix, const uint32_t* counts, const uint64_t* uniq_n,
                     const uint64_t* uniq_sum, const uint64_t* uniq_mode, uint64_t num_kmers, double cov) {
    for (auto& kv : ix.colors) {
        const uint64_t c = kv.first;
        const uint32_t v = counts[c];
        if (v == 0) continue;
        double mean = 0.0;
        uint64_t modus = 0, specific = 0;
        if (uniq_n && uniq_n[c]) {
            mean = (double)uniq_sum[c] / (double)uniq_n[c];
            modus = uniq_mode[c];
            specific = uniq_n[c];
        }
        auto it = ix.n_ref_kmers.find(kv.second);
        if (it == ix.n_ref_kmers.end()) continue;
        const double genome_cov = (double)v / (double)it->second;
        if (genome_cov > cov)
            fprintf(out, "%s\t%llu\t%s\t%.2f\t%.2f\t%llu\t%llu\n", query.c_str(), (unsigned long long)num_kmers, kv.second.c_str(),
                    genome_cov, mean, (unsigned long long)modus, (unsigned long long)specific);
    }
}

// reports.rs:50-62
void generate_report_gene(FILE* out, const std::string& query, const Bigsi& ix, const uint32_t* counts, uint64_t num_kmers,
                          double cov) {
    for (auto& kv : ix.colors) {
        const uint32_t v = counts[kv.first];
        if (v == 0) continue;
        const double gene_match = (double)v / (double)num_kmers;
        if (gene_match >= cov)
            fprintf(out, "%s\t%s\t%llu\t%.3f\n", query.c_str(), kv.second.c_str(), (unsigned long long)num_kmers, gene_match);
    }
}

// FNV-1a over the key's bytes followed by 0xFF (`impl Hash for str`), hashbrown bucket layout after
// entry(key).or_insert() of the keys in order: buckets 4, 8, 16, ..; capacity b-1 (b < 8) or b/8*7; a new
// key arriving with no growth left doubles the table, re-inserting old buckets in ascending order;
// insert slot = first empty bucket of the 16-wide group at hash & mask, else triangular probing.
static uint64_t fnv1a_str(const std::string& s) {
    uint64_t h = 0xcbf29ce484222325ULL;
    for (unsigned char c : s) { h ^= c; h *= 0x100000001b3ULL; }
    h ^= 0xFF; h *= 0x100000001b3ULL;
    return h;
}
std::vector<size_t> fnv_string_map_order(const std::vector<std::string>& keys, uint32_t gw) {
    std::vector<uint64_t> hash(keys.size());
    for (size_t i = 0; i < keys.size(); i++) hash[i] = fnv1a_str(keys[i]);
    std::vector<long> tab, tmp;
    uint64_t nb = 0, items = 0, growth = 0;
    auto cap_of = [](uint64_t b) { return b == 0 ? 0 : (b < 8 ? b - 1 : b / 8 * 7); };
    auto probe = [&](const std::vector<long>& t, uint64_t buckets, uint64_t h) -> uint64_t {
        const uint64_t mask = buckets - 1;
        uint64_t pos = h & mask;
        const uint64_t width = buckets < gw ? buckets : gw;
        for (uint64_t stride = 0;;) {
            for (uint64_t b = 0; b < width; b++) { uint64_t s = (pos + b) & mask; if (t[s] < 0) return s; }
            stride += gw;
            pos = (pos + stride) & mask;
        }
    };
    for (size_t i = 0; i < keys.size(); i++) {
        if (growth == 0) {
            const uint64_t nn = nb == 0 ? 4 : nb * 2;
            tmp.assign(nn, -1);
            for (uint64_t s = 0; s < nb; s++) if (tab[s] >= 0) tmp[probe(tmp, nn, hash[tab[s]])] = tab[s];
            tab.swap(tmp);
            nb = nn;
            growth = cap_of(nb) - items;
        }
        tab[probe(tab, nb, hash[i])] = (long)i;
        items++; growth--;
    }
    std::vector<size_t> order;
    for (uint64_t s = 0; s < nb; s++) if (tab[s] >= 0) order.push_back((size_t)tab[s]);
    return order;
}

void write_counts_five_fields(const std::string& path, const std::vector<std::string>& insertion_keys,
                              const std::map<std::string, uint64_t>& counts) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw Error("could not create outfile!");
    for (size_t i : fnv_string_map_order(insertion_keys))
        fprintf(f, "%s\t%llu\n", insertion_keys[i].c_str(), (unsigned long long)counts.at(insertion_keys[i]));
    fclose(f);
}

double false_prob(double m, double k, double n) { return std::pow(1.0 - std::pow(M_E, -((k * (n + 0.5)) / (m - 1.0))), k); }

}  // namespace cidh
