// colorid-b200: the reference's command line (main.rs: clap App "colorid" 0.1.4.3) for the subcommands on
// the BIGSI hot path -- build, search, read_id, info -- with the same flags, defaults and output files,
// running on the GPU through libcolorid_b200.so, minimizer indexes (`build -m -v M`, .mxi) included.
// `batch_id` (index uploaded once, many samples) and `read_filter` (pure file I/O) complete the read workflow.
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>

#include <zlib.h>

#include "cid_host.hpp"

namespace {
struct Flag { char s; const char* l; bool takes_value; bool multi; };
struct Parsed {
    std::map<std::string, std::vector<std::string>> v;
    bool has(const char* l) const { return v.count(l) != 0; }
    const std::string& one(const char* l) const { return v.at(l).front(); }
};
Parsed parse(int argc, char** argv, int at, const std::vector<Flag>& flags) {
    Parsed p;
    auto find = [&](const std::string& a) -> const Flag* {
        for (auto& f : flags) {
            if (a.size() == 2 && a[0] == '-' && a[1] == f.s) return &f;
            if (a.size() > 2 && a[0] == '-' && a[1] == '-' && a.substr(2) == f.l) return &f;
        }
        return nullptr;
    };
    while (at < argc) {
        const std::string a = argv[at++];
        const Flag* f = find(a);
        if (!f) throw cidh::Error("error: Found argument '" + a + "' which wasn't expected, or isn't valid in this context");
        auto& dst = p.v[f->l];
        if (!f->takes_value) { dst.push_back("1"); continue; }
        if (at >= argc) throw cidh::Error("error: The argument '--" + std::string(f->l) + "' requires a value but none was supplied");
        dst.push_back(argv[at++]);
        if (f->multi) while (at < argc && argv[at][0] != '-') dst.push_back(argv[at++]);   // clap min_values(1)
    }
    return p;
}
void require(const Parsed& p, std::initializer_list<const char*> names) {
    for (auto n : names) if (!p.has(n)) throw cidh::Error(std::string("error: The following required arguments were not provided:\n    --") + n);
}
template <class T> T num(const Parsed& p, const char* l, T dflt) {      // value_t!(..).unwrap_or(default)
    if (!p.has(l)) return dflt;
    char* end = nullptr;
    const std::string& s = p.one(l);
    if (std::is_floating_point<T>::value) { double d = strtod(s.c_str(), &end); return (*end || s.empty()) ? dflt : (T)d; }
    long long v = strtoll(s.c_str(), &end, 10);
    return (*end || s.empty()) ? dflt : (T)v;
}
int usage() {
    fprintf(stderr,
            "colorid-b200 0.1.4.3 (B200)\nBIGSI based taxonomic ID of sequence data\n\nUSAGE:\n    colorid-b200 <SUBCOMMAND>\n\nSUBCOMMANDS:\n"
            "    build      builds a bigsi            -b PREFIX -r REFS.tsv -k K -n HASHES -s BLOOM [-t T] [-Q q] [-f cutoff] [-m [-v M]]\n"
            "    search     does a bigsi search       -b IDX.bxi -q Q... [-r R...] [-f n] [-p cov] [-g] [-s] [-m] [-Q q]\n"
            "    read_id    id's reads                -b IDX.bxi|.mxi -q R1.fq.gz [R2.fq.gz] -n PREFIX [-t T] [-c batch] [-d d] [-p e] [-Q q] [-B b] [-H]\n"
            "    batch_id   id's reads of many samples -b IDX -q SAMPLES.tsv -T TAG [read_id options]\n"
            "    read_filter filters reads by class   -c PREFIX_reads.txt -f R1.fq.gz [R2.fq.gz] -t TAXON -p PREFIX [-e]\n"
            "    info       dumps index parameters    -b IDX.bxi|.mxi\n");
    return 1;
}
}  // namespace

int main(int argc, char** argv) {
    printf("\n ************** initializing logger *****************\n\n");      // main.rs:18
    if (argc < 2) return usage();
    const std::string sub = argv[1];
    int device = 0;
    if (const char* d = getenv("COLORID_B200_DEVICE")) device = atoi(d);
    try {
        if (sub == "build") {
            Parsed p = parse(argc, argv, 2, {{'b', "bigsi", true, false}, {'r', "refs", true, false}, {'k', "kmer", true, false},
                                             {'n', "num_hashes", true, false}, {'s', "bloom", true, false}, {'m', "minimizer", false, false},
                                             {'v', "value", true, false}, {'t', "threads", true, false}, {'Q', "quality", true, false},
                                             {'f', "filter", true, false}});
            require(p, {"bigsi", "refs", "kmer", "num_hashes", "bloom"});
            printf(" Ref_file : %s\n Bigsi file : %s\nK-mer size: %s\nBloom filter parameters: num hashes %s, filter size %s\n",
                   p.one("refs").c_str(), p.one("bigsi").c_str(), p.one("kmer").c_str(), p.one("num_hashes").c_str(), p.one("bloom").c_str());
            cidh::BuildOpts o;
            o.minimizer = p.has("minimizer");                              // main.rs:480-482: -v defaults to 15
            o.minimizer_value = num<uint64_t>(p, "value", 15);
            o.ref_file = p.one("refs"); o.prefix = p.one("bigsi");
            o.k = num<uint64_t>(p, "kmer", 31); o.bloom = num<uint64_t>(p, "bloom", 50000000); o.hashes = num<uint64_t>(p, "num_hashes", 4);
            o.threads = num<uint64_t>(p, "threads", 1); o.quality = num<uint8_t>(p, "quality", 15); o.filter = num<int64_t>(p, "filter", -1);
            o.device = device;
            return cidh::build(o);
        }
        if (sub == "search") {
            Parsed p = parse(argc, argv, 2, {{'b', "bigsi", true, false}, {'q', "query", true, true}, {'r', "reverse", true, true},
                                             {'f', "filter", true, false}, {'p', "p_shared", true, false}, {'g', "gene_search", false, false},
                                             {'s', "perfect_search", false, false}, {'m', "multi_fasta", false, false},
                                             {'Q', "quality", true, false}});
            require(p, {"bigsi", "query"});
            cidh::SearchOpts o;
            o.bigsi = p.one("bigsi"); o.files1 = p.v.at("query");
            if (p.has("reverse") && p.one("reverse") != "none") o.files2 = p.v.at("reverse");
            o.filter = num<int64_t>(p, "filter", -1); o.cov = num<double>(p, "p_shared", 0.35);
            o.gene_search = p.has("gene_search"); o.perfect_search = p.has("perfect_search"); o.multi_fasta = p.has("multi_fasta");
            o.quality = num<uint8_t>(p, "quality", 15); o.device = device;
            return cidh::search(o);
        }
        if (sub == "read_id") {
            Parsed p = parse(argc, argv, 2, {{'b', "bigsi", true, false}, {'q', "query", true, true}, {'c', "batch", true, false},
                                             {'t', "threads", true, false}, {'n', "prefix", true, false}, {'d', "down_sample", true, false},
                                             {'H', "high_mem_load", false, false}, {'p', "fp_correct", true, false},
                                             {'Q', "quality", true, false}, {'B', "bitvector_sample", true, false}});
            require(p, {"bigsi", "query", "prefix"});
            cidh::ReadIdOpts o;
            o.bigsi = p.one("bigsi"); o.query = p.v.at("query"); o.prefix = p.one("prefix");
            o.threads = num<uint64_t>(p, "threads", 0); o.down_sample = num<uint64_t>(p, "down_sample", 1);
            o.correct = num<double>(p, "fp_correct", 3.0); o.quality = num<uint8_t>(p, "quality", 15);
            o.batch = num<uint64_t>(p, "batch", 50000); o.high_mem_load = p.has("high_mem_load");
            o.bitvector_sample = num<uint64_t>(p, "bitvector_sample", 3); o.device = device;
            return cidh::read_id(o);
        }
        if (sub == "info") {
            Parsed p = parse(argc, argv, 2, {{'b', "bigsi", true, false}, {'c', "compressed", true, false}});
            require(p, {"bigsi"});
            return cidh::info(p.one("bigsi"));
        }
        if (sub == "_host") {
            // parity hooks for the CPU test-suite (no GPU involved): dump what the host-side readers/writers produce
            const std::string op = argc > 2 ? argv[2] : "";
            if (op == "fasta" && argc == 4) {
                for (auto& r : cidh::read_fasta(argv[3])) printf("%s\n", r.c_str());
            } else if (op == "fasta_mf" && argc == 4) {
                std::vector<std::string> l, q;
                cidh::read_fasta_mf(argv[3], l, q);
                for (auto& x : l) printf("L\t%s\n", x.c_str());
                for (auto& x : q) printf("S\t%s\n", x.c_str());
            } else if (op == "fastq" && (argc == 5 || argc == 6)) {      // Q R1 [R2]
                cidh::SeqBatch sb;
                const uint8_t q = (uint8_t)atoi(argv[3]);
                const uint64_t n = argc == 6 ? cidh::fastq_masked_pe(argv[4], argv[5], q, sb) : cidh::fastq_masked_se(argv[4], q, sb);
                printf("records\t%llu\n", (unsigned long long)n);
                for (uint64_t i = 0; i < sb.n(); i++) printf("%.*s\n", (int)(sb.offs[i + 1] - sb.offs[i]), sb.bases.data() + sb.offs[i]);
            } else if (op == "lines" && argc == 4) {            // raw lines of a (gz) file, EOLs kept: the decoder's parity hook
                cidh::LineReader lr(argv[3]);
                std::string l;
                unsigned long long n = 0, bytes = 0;
                unsigned long crc = crc32(0L, Z_NULL, 0);
                while (lr.next(l, true)) { n++; bytes += l.size(); crc = crc32(crc, (const Bytef*)l.data(), (uInt)l.size()); }
                printf("lines\t%llu\nbytes\t%llu\ncrc32\t%08lx\n", n, bytes, crc);
            } else if (op == "alines" && (argc == 4 || argc == 5)) {   // the same through AsyncLineReader (block-wise splitting); argv[4] = keep EOLs
                const bool keep = argc == 5 && atoi(argv[4]) != 0;
                cidh::AsyncLineReader lr(argv[3], keep);
                std::string_view v;
                unsigned long long n = 0, bytes = 0;
                unsigned long crc = crc32(0L, Z_NULL, 0);
                while (lr.next_view(v)) { n++; bytes += v.size(); crc = crc32(crc, (const Bytef*)v.data(), (uInt)v.size()); if (!keep) crc = crc32(crc, (const Bytef*)"\n", 1); }
                printf("lines\t%llu\nbytes\t%llu\ncrc32\t%08lx\n", n, bytes, crc);
            } else if (op == "fastq_parse" && (argc == 5 || argc == 6)) {      // <quality offset> <file1> [file2]: read_id's record loop, batches digested
                return cidh::fastq_parse_digest(std::vector<std::string>(argv + 4, argv + argc), (uint8_t)atoi(argv[3]));
            } else if (op == "acount" && argc == 4) {            // AsyncLineReader as the drivers use it, lines only counted (timing aid)
                cidh::AsyncLineReader lr(argv[3]);
                std::string_view v;
                unsigned long long n = 0, bytes = 0;
                while (lr.next_view(v)) { n++; bytes += v.size(); }
                printf("lines\t%llu\nbytes\t%llu\n", n, bytes);
            } else if (op == "lines1" && argc == 4) {            // LineReader::next without EOLs, one '\n' folded in per line
                cidh::LineReader lr(argv[3]);
                std::string l;
                unsigned long long n = 0, bytes = 0;
                unsigned long crc = crc32(0L, Z_NULL, 0);
                while (lr.next(l, false)) { n++; bytes += l.size(); crc = crc32(crc, (const Bytef*)l.data(), (uInt)l.size()); crc = crc32(crc, (const Bytef*)"\n", 1); }
                printf("lines\t%llu\nbytes\t%llu\ncrc32\t%08lx\n", n, bytes, crc);
            } else if (op == "gunzip" && argc == 4) {           // decode only (timing aid): bytes out
                unsigned long long bytes = 0;
                std::vector<char> buf(1 << 20);
                if (getenv("COLORID_B200_ZLIB")) {
                    gzFile f = gzopen(argv[3], "rb"); gzbuffer(f, 1 << 20);
                    for (int n; (n = gzread(f, buf.data(), (unsigned)buf.size())) > 0;) bytes += (unsigned long long)n;
                    gzclose(f);
                } else {
                    cidh::MappedGz mg(argv[3]);
                    for (size_t n; (n = mg.read(buf.data(), buf.size())) > 0;) bytes += n;
                }
                printf("bytes\t%llu\n", bytes);
            } else if (op == "tab" && argc == 4) {
                for (auto& kv : cidh::tab_to_map(argv[3])) { printf("%s", kv.first.c_str()); for (auto& f : kv.second) printf("\t%s", f.c_str()); printf("\n"); }
            } else if (op == "bxi_copy" && argc == 5) {
                cidh::save_bigsi(argv[4], cidh::read_bigsi(argv[3]));
            } else if (op == "mxi_copy" && argc == 5) {
                cidh::save_bigsi_mini(argv[4], cidh::read_bigsi_mini(argv[3]));
            } else if (op == "map_order" && argc >= 4) {
                std::vector<std::string> keys(argv + 3, argv + argc);
                for (size_t i : cidh::fnv_string_map_order(keys)) printf("%s\n", keys[i].c_str());
            } else return usage();
            return 0;
        }
        if (sub == "batch_id") {                                   // main.rs:330-416,869-887
            Parsed p = parse(argc, argv, 2, {{'b', "bigsi", true, false}, {'q', "query", true, true}, {'T', "tag", true, false},
                                             {'c', "batch", true, false}, {'t', "threads", true, false}, {'d', "down_sample", true, false},
                                             {'H', "high_mem_load", false, false}, {'p', "fp_correct", true, false},
                                             {'Q', "quality", true, false}, {'B', "bitvector_sample", true, false}});
            require(p, {"bigsi", "query", "tag"});
            cidh::BatchIdOpts o;
            o.bigsi = p.one("bigsi"); o.batch_samples = p.one("query"); o.tag = p.one("tag");
            o.threads = num<uint64_t>(p, "threads", 0); o.down_sample = num<uint64_t>(p, "down_sample", 1);
            o.correct = num<double>(p, "fp_correct", 3.0); o.quality = num<uint8_t>(p, "quality", 15);
            o.batch = num<uint64_t>(p, "batch", 50000); o.high_mem_load = p.has("high_mem_load");
            o.bitvector_sample = num<uint64_t>(p, "bitvector_sample", 3); o.device = device;
            if (const char* ds = getenv("COLORID_B200_DEVICES"))          // e.g. "0,1,2,3": one worker (and index replica) per GPU
                for (const char* q = ds; *q;) { char* e; long v = strtol(q, &e, 10); if (e == q) break; o.devices.push_back((int)v); q = *e ? e + 1 : e; }
            return cidh::batch_id(o);
        }
        if (sub == "read_filter") {                                // main.rs:418-465,888-900
            Parsed p = parse(argc, argv, 2, {{'c', "classification", true, false}, {'f', "files", true, true}, {'t', "taxon", true, false},
                                             {'p', "prefix", true, false}, {'e', "exclude", false, false}});
            require(p, {"classification", "files", "taxon", "prefix"});
            cidh::ReadFilterOpts o;
            o.classification = p.one("classification"); o.files = p.v.at("files"); o.taxon = p.one("taxon");
            o.prefix = p.one("prefix"); o.exclude = p.has("exclude");
            return cidh::read_filter(o);
        }
        return usage();
    } catch (const std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return 101;                                     // a Rust panic exits with 101
    }
}
