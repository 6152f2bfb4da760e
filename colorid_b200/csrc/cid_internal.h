// Internal host-side structures of libcolorid_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/colorid_b200.h"

namespace cid {
struct ModS;

// Where query_gather adds its per-accession counts: `n` destination buffers (this GPU's own and, in column-sharded
// mode, the peers' -- NVLink peer memory), rows of `stride` counters, this shard's accessions starting at column `col0`.
struct GatherOut { uint32_t* base[8]; uint32_t n; uint32_t stride; uint32_t col0; uint32_t dense; };
// dense: every query is ONE gather unit (query_front path), so each counter is written exactly once: plain 16-byte stores of
// all counters (zeros included) instead of one atomic per non-zero counter -- what peer memory over NVLink wants

void set_error(const char* fmt, ...);

#define CID_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            cid::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return CID_E_CUDA;                                                                 \
        }                                                                                      \
    } while (0)

#define CID_TRY(expr)            \
    do {                         \
        int _rc = (expr);        \
        if (_rc != CID_OK) return _rc; \
    } while (0)

// Growable device buffer (never shrinks); owned by a ctx.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
    template <class T> T* as() const { return (T*)p; }
};
// Growable pinned host buffer.
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
    template <class T> T* as() const { return (T*)p; }
};

enum { SCRATCH_SLOTS = 28 };

}  // namespace cid

struct cid_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;   // library-owned stream for the host-pointer entry points
    uint64_t launches = 0;
    const cid::GatherOut* gather_out = nullptr;   // set for the duration of cid_query_counts_sharded_dev
    bool attr_done[12] = {false};     // cudaFuncSetAttribute (dynamic smem opt-in) already applied on this context's device
    cid::DevBuf scratch[cid::SCRATCH_SLOTS];
    cid::PinBuf pinned[8];
    uint32_t* d_err = nullptr;       // device error/flag words (zeroed before each call)
    uint32_t* h_err = nullptr;       // pinned mirror
    // optional per-kernel timing (CUDA events on the launching stream)
    bool prof_on = false;
    struct ProfRec { int kernel; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[32] = {0};
    uint64_t prof_n[32] = {0};
    // read_id host pipeline (cid_readid_pipe.cu): chunked H2D / kernels / D2H / host vote overlap
    struct cid_readid_pipe* pipe = nullptr;
    uint64_t opt_readid_chunk = 0;   // reads per pipeline chunk (0 = automatic)
    int opt_host_ranks = 1;          // ranks (processes or cid_mg shards) that share this host's memory system with this context
    int opt_readid_chunk_growth = 0; // size of a pipeline chunk relative to its predecessor, in percent (0 = automatic)
    uint64_t opt_readid_chunk0 = 0;  // reads in the first chunk of the ramp (0 = 32768)
    int opt_host_threads = 0;        // host threads for the vote (0 = all cores)
    // device-pointer read_id (cid_read_id_batch_dev): chunks alternate over internal streams forked from /
    // joined to the caller's stream, so the DRAM-access-bound vote kernel of one chunk runs under the
    // issue-bound kmerize/order kernels of the next (measured on B200: no gain, kernels fill the GPU; off by default)
    int opt_readid_streams = 1;
    int opt_kmerize_ctas = 0, opt_vote_ctas = 0;   // CTAs per SM of the read_id kmerize / vote grids (0 = fill the GPU);
                                                   // smaller grids let the two kernels of different chunks share the SMs
    int opt_readid_report_steps = 0; // 1 = report colours carry their insertion step in bits 20..31 (column-sharded read_id, cid_merge_shard_reports)
    int opt_readid_vote_part = 1;    // narrow-row vote partitioned by row range (L2-resident gathers): 0 = never, 1 = when the matrix spans >= 3 windows and the chunk is large, 2 = always (tests)
    int opt_readid_part_ctas = 0;    // CTAs per SM of the partitioned vote's scan kernel (0 = as many as fit)
    int opt_readid_part_cap = 0;     // tuples per partition bucket (0 = sized from the chunk); tests force the overflow path with a tiny value
    int opt_readid_part_shift = 0;   // log2 rows per partition (0 = 64 MB windows); tests use small values to get several partitions on small matrices
    int opt_readid_serialize = 0;    // 1 = kernels of consecutive pipeline chunks never overlap (measured: 44.8M vs 46.3M pairs/s e2e, off)
    int opt_build_table_div = 0;     // read-set builds: first count table = k-mer positions / this (0 = adaptive; grown x4 when > 70 % full)
    double readset_ratio = 0;        // distinct k-mers / k-mer positions of the last read-set accession built on this context
    int opt_build_packed = 1;        // 0 = never use the packed 8-byte count table (parity aid)
    int opt_build_set = 1;           // 0 = always build through the count table (parity aid)
    int opt_query_front = 1;         // 0 = never use the shared-memory dedup front end of small queries (parity aid)
    int opt_gather_l2_64b = 0;       // 1 = query_gather row copies with the 64-byte L2 prefetch size (measured: no effect)
    int opt_query_table_div = 4;     // read-set queries: first count table = k-mer positions / this (0 = always the safe 2x-positions table)
    uint64_t opt_query_table_min = 1ull << 23;   // ... for tables of at least this many slots at the safe size (0 = always: tests)
    int opt_query_compact = 1;       // count-table compaction of large queries: 0 = never, 1 = tables of >= 2^22 slots, 2 = always (tests)
    int opt_uniq_device = 1;         // 0 = unique-hit summaries always through the host maps (parity aid)
    int opt_query_fused = 0;         // 1 = force the fused collect/hash/gather kernel (parity aid)
    bool case_aware = false;         // set while a raw-case (FASTQ) input with lower-case k-mers is redone: 16-byte count-table slots carry a case mask
    cudaStream_t aux[2] = {nullptr, nullptr};
    cudaEvent_t aux_fork = nullptr, aux_join[2] = {nullptr, nullptr};
};

struct cid_index {
    cid_ctx* ctx = nullptr;
    uint64_t S = 0;
    uint32_t H = 0, k = 0, N = 0;
    uint32_t hv = 0;     // hash variant (cid_index_set_hash_variant; 0 = stable XXH3)
    void* d_hcfg = nullptr;   // cid::HashCfg of the variant, in device memory
    uint32_t m = 0;      // minimizer length of an .mxi index (bigsi.rs:40-49 m_size); 0 = plain k-mer index
    uint32_t W = 0;      // ceil(N/32)
    uint32_t Wp = 0;     // device row stride in words
    uint32_t* rows = nullptr;      // [S][Wp]
    uint32_t* rownz = nullptr;     // bitmap, bit r = row r has any bit set
    uint64_t rownz_words = 0;
    uint32_t* bitsets = nullptr;   // build mode: [N][bs_words] per-colour Bloom bitsets
    uint64_t bs_words = 0;         // words per bitset (multiple of 32)
    bool rownz_valid = false;
    bool rownz_global = false;     // column-sharded: rownz was OR-ed across ranks; never derive presence from local words
};

namespace cid {

// error flag bits written by kernels into ctx->d_err[0]
enum {
    ERRF_LOWER_RAW = 1u << 0,      // lower-case base inside a valid window in a raw-case mode
    ERRF_STARTS_OVERFLOW = 1u << 1,
    ERRF_LIST_OVERFLOW = 1u << 2,
    ERRF_READ_TOO_LONG = 1u << 3,
    ERRF_STRING_NONACGT = 1u << 4,
    ERRF_COUNT_OVERFLOW = 1u << 6, // packed count table: a multiplicity near 2^22 (caller redoes the accession with 16-byte slots)
    ERRF_TABLE_FULL = 1u << 5,     // a count-table region overflowed (only possible with optimistic sizing: caller retries) // kmerize_string window with a byte outside ACGTacgt (cannot be 2-bit packed)
};

int check_err_flags(cid_ctx* ctx, cudaStream_t st, bool* lower_raw = nullptr);   // syncs the stream

// kernel ids for the profiling hooks (names in cid_api.cu: kKernelNames)
enum { KID_KMERIZE_INSERT = 0, KID_HISTOGRAM, KID_TO_BLOOM, KID_TRANSPOSE, KID_ROWNZ, KID_QUERY_COUNTS, KID_QUERY_UNIQ_WIDE,
       KID_QUERY_PERFECT, KID_READID_KMERIZE, KID_READID_SCHED, KID_READID_ORDER, KID_READID_VOTE, KID_READID_CLASSIFY, KID_TABLE_CLEAR, KID_OTHER, KID_QUERY_HASH, KID_QUERY_FRONT, KID_READID_BIG, KID_READID_VP_SCAN, KID_READID_VP_GATHER, KID_READID_VP_COUNT, KID_COUNT };
struct ProfScope {     // records an event pair around a launch when profiling is enabled
    cid_ctx* ctx; cudaStream_t st; int idx;
    ProfScope(cid_ctx* c, cudaStream_t s, int kernel);
    ~ProfScope();
};

inline uint32_t padded_row_words(uint32_t W) {
    if (W <= 1) return 1;
    if (W <= 2) return 2;
    return (W + 3) & ~3u;   // 16-byte aligned rows
}
inline uint64_t next_pow2(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }

// ---- launchers implemented in the .cu files (all asynchronous on `st`) ------------------------
struct GroupRegions {           // count-table regions, one per group (query / accession)
    uint64_t ngroups = 0;
    uint64_t total_slots = 0;
    std::vector<uint64_t> off, mask;   // host copies
};

// Size regions from per-group base counts; fills host vectors.
void plan_regions(const uint64_t* h_seq_offs, const uint64_t* h_group_offs, uint64_t ngroups, uint32_t k,
                  GroupRegions& gr);

int launch_kmerize_insert(cid_ctx* ctx, cudaStream_t st, const uint8_t* d_bases, const uint64_t* d_seq_offs,
                          uint64_t nseq, uint64_t base_lo, uint64_t base_hi, const uint32_t* d_seq_group,
                          const uint64_t* d_region_off, const uint64_t* d_region_mask, void* d_table, uint32_t k,
                          int seq_mode, uint32_t mini_m = 0, uint64_t packed_slots = 0);
// mini_m != 0: count each k-mer's minimizer instead; packed_slots != 0: d_table is ONE packed count table (key << 22 | count,
// keys of <= 21 bases) of that many 8-byte words for the whole launch
// set-only build (no count filter): new keys are hashed straight into the accession's Bloom bitset; distinct keys in d_err[1]
int launch_kmerize_bloom(cid_ctx* ctx, cudaStream_t st, const uint8_t* d_bases, const uint64_t* d_seq_offs, uint64_t nseq,
                         uint64_t nbases, void* d_keys, uint64_t nslots, uint32_t k, int seq_mode, uint32_t count_m, uint32_t bloom_m,
                         uint32_t H, const ModS& mods, uint32_t* d_bitset);
int launch_region_histogram(cid_ctx* ctx, cudaStream_t st, const void* d_region, uint64_t nslots, uint32_t* d_hist,
                            uint32_t hist_bins, uint32_t* d_overflow, uint32_t overflow_cap, uint32_t* d_overflow_n,
                            bool packed = false);
// k = length of the table's keys; mini_m != 0: insert find_minimizer(key, mini_m) instead of the key itself
int launch_region_to_bloom(cid_ctx* ctx, cudaStream_t st, const void* d_region, uint64_t nslots, int64_t cutoff,
                           uint32_t k, uint32_t mini_m, uint32_t H, const ModS& mods, uint32_t* d_bitset, unsigned long long* d_nref,
                           bool packed = false);
int launch_transpose(cid_ctx* ctx, cudaStream_t st, const cid_index* idx);
int launch_rownz(cid_ctx* ctx, cudaStream_t st, const cid_index* idx);

struct QueryUnits {            // work list for the gather kernels: (group, first slot, nslots)
    std::vector<uint32_t> group;
    std::vector<uint64_t> slot0;
    std::vector<uint32_t> nslots;
};
void plan_units(const GroupRegions& gr, uint32_t chunk, QueryUnits& qu);
enum { QUERY_CHUNK = 1024, QUERY_ITEM_SLOTS = 16384, MAX_HASH = 8 };

int launch_query_counts(cid_ctx* ctx, cudaStream_t st, const cid_index* idx, const void* d_table,
                        const uint32_t* d_unit_group, const uint64_t* d_unit_slot0, const uint32_t* d_unit_nslots,
                        uint64_t nunits, uint64_t total_slots, const int64_t* d_filter, uint32_t* d_counts,
                        unsigned long long* d_num_kmers, bool want_uniq, uint32_t* d_uniq_list, uint32_t uniq_cap,
                        uint32_t* d_uniq_n);
// small-query fast path (distinct k-mers only): shared-memory dedup + hash, then the streaming gather
bool query_front_fits(const uint64_t* h_seq_offs, const uint64_t* h_query_offs, uint64_t q0, uint64_t q1, uint32_t k);
int launch_query_front_gather(cid_ctx* ctx, cudaStream_t st, const cid_index* idx, const uint8_t* d_bases,
                              const uint64_t* d_seq_offs, const uint64_t* d_query_offs, const uint64_t* h_seq_offs,
                              const uint64_t* h_query_offs, uint64_t q0, uint64_t q1, int seq_mode, uint32_t* d_counts,
                              unsigned long long* d_num_kmers, uint32_t* d_and_rows = nullptr, uint32_t* d_missing = nullptr);
// d_and_rows != nullptr: perfect search -- AND of all rows into d_and_rows[q][W] (pre-set to ones), absent rows flag d_missing[q]
int launch_query_perfect(cid_ctx* ctx, cudaStream_t st, const cid_index* idx, const void* d_table,
                         const uint32_t* d_unit_group, const uint64_t* d_unit_slot0, const uint32_t* d_unit_nslots,
                         uint64_t nunits, uint32_t* d_and_rows, uint32_t* d_missing, unsigned long long* d_num_kmers);

int launch_hash_kmers(cid_ctx* ctx, cudaStream_t st, const cid_index* idx, const uint8_t* d_kmers, uint64_t n,
                      uint64_t* d_rows);
// survivors (count > filt) of one count-table region of 16-byte slots -> dense slot list (d_dense == nullptr: count only);
// *d_n must be zero before the call
int launch_region_compact(cid_ctx* ctx, cudaStream_t st, const void* d_region, uint64_t nslots, int64_t filt, void* d_dense,
                          unsigned long long* d_n, uint64_t cap);
// column-sharded default report: per dense-list index min(popcount of the AND row over this shard, 2) and the single hit's colour;
// then the triples (query, colour, multiplicity) of the k-mers whose local and cross-shard popcounts are both 1
int launch_slots_popcount(cid_ctx* ctx, cudaStream_t st, const cid_index* idx, const void* d_slots, uint64_t n, uint8_t* d_pc,
                          uint32_t* d_col);
int launch_uniq_emit(cid_ctx* ctx, cudaStream_t st, const void* d_slots, const uint64_t* d_prefix, uint32_t nq, uint64_t n,
                     const uint8_t* d_pc_local, const uint8_t* d_pc_sum, const uint32_t* d_col, uint32_t* d_list, uint32_t cap,
                     uint32_t* d_n);
// unique-hit triples (query, accession, multiplicity) -> per (query, accession) number, sum and mode of the multiplicities
// (reports.rs:20-26) through a [cells][MB] histogram; *d_ovf != 0 afterwards: a multiplicity >= MB occurred, use the host path
int launch_uniq_summaries(cid_ctx* ctx, cudaStream_t st, const uint32_t* d_list, uint32_t nu, uint32_t N, uint64_t cells, uint32_t MB,
                          uint32_t* d_hist, uint32_t* d_ovf, unsigned long long* d_n, unsigned long long* d_sum,
                          unsigned long long* d_mode);

// read_id kernels on device buffers.  Reads [r_first, r_first + nreads) of the arrays are processed
// (every per-read array is indexed by the absolute read number, so a caller that stages only a
// chunk passes pointers biased by the chunk origin).  Per-read scratch (`entries`, `order`, `nocc`)
// is indexed from 0 and must hold `scr.cap_reads` reads; larger ranges are walked in pieces.
// read source of the read_id kernels: ASCII (+ qualities) or packed planes (cid_pack_reads); see cid_readid.cu
struct ReadSrc { const uint8_t* bases; const uint8_t* quals; uint32_t maxq; const uint32_t* pk; const uint64_t* pk_offs; uint32_t pk_lower; };
struct PackedReads { const uint32_t* words; const uint64_t* word_offs; uint32_t lower; };   // device pointers; word_offs indexed by absolute read
struct ReadIdScratch {
    uint32_t* entries; uint16_t* order; uint32_t* nocc; uint64_t cap_reads;
    DevBuf* vp = nullptr;      // grown on demand for the partitioned vote (cid_readid_part.cu); nullptr = one-kernel vote only
    // general path (cid_readid_big.cu): scratch of big_ctas CTAs sized by readid_big_plan(big_bases, big_kmers)
    uint8_t* big = nullptr; uint32_t big_ctas = 0; uint32_t big_bases = 0, big_kmers = 0;
};
enum { READID_FAST_BASES = 1000 };     // longest read (all mates) of the warp-per-read kernels
void readid_scratch_bytes(const cid_index* idx, uint32_t max_read_bases, uint32_t max_kmers, uint64_t reads,
                          size_t* entries_bytes, size_t* order_bytes, size_t* nocc_bytes);
int readid_run(cid_index* idx, cudaStream_t st, const uint8_t* d_bases, const uint8_t* d_quals, const PackedReads* d_packed,
               const uint64_t* d_seq_offs, const uint64_t* d_read_offs, uint64_t r_first, uint64_t nreads,
               uint32_t max_read_bases, uint32_t max_kmers, const cid_readid_params& p, const ReadIdScratch& scr,
               uint32_t* d_n_set, uint32_t* d_flags, uint32_t* d_rep_n, uint32_t* d_rep_colour, uint32_t* d_rep_count,
               uint32_t order_cap, uint32_t* d_order_n, uint8_t* d_order_seq, uint32_t* d_order_pos);
// general path: reads listed in d_list[0 .. *d_list_n) (indices relative to r0), one CTA per read, tables in `d_scratch`
void readid_big_plan(const cid_index* idx, uint32_t max_bases, uint32_t max_kmers, size_t budget, size_t* bytes, uint32_t* ctas);
int launch_readid_big(cid_index* idx, cudaStream_t st, const ReadSrc& src,
                      const uint64_t* d_seq_offs, const uint64_t* d_read_offs, uint64_t r0, const uint32_t* d_list,
                      const uint32_t* d_list_n, uint32_t max_bases, uint32_t max_kmers, const cid_readid_params& p,
                      uint8_t* d_scratch, uint32_t ctas, uint32_t* d_n_set, uint32_t* d_flags, uint32_t* d_rep_n,
                      uint32_t* d_rep_colour, uint32_t* d_rep_count, uint32_t order_cap, uint32_t* d_order_n,
                      uint8_t* d_order_seq, uint32_t* d_order_pos);

// partitioned vote of narrow rows (cid_readid_part.cu): row gathers bucketed by row range so that they hit L2
constexpr int VP_MAXP = 16;        // partitions (row ranges) at most
constexpr uint32_t VP_MAXB = 15;   // `-B` k-mers read directly at most
constexpr int VP_MAXCAND = 8;      // candidate colours per read held as one accumulator byte per k-mer
struct VotePart {                  // device view of one chunk's scratch
    uint2* tuples; uint32_t cap;   // [P][cap] (row, accumulator slot): per-warp blocks interleaved, cap_warp tuples per warp at most
    uint32_t cap_warp; uint32_t* ntup;   // [P][W] tuples each warp of the scan pushed
    uint32_t* cursor;              // [VP_MAXP] longest per-warp total of each partition, [VP_MAXP] overflow flag
    uint32_t* direct_n; uint32_t* direct;   // reads left to the one-kernel vote (more than VP_MAXCAND candidates)
    uint32_t* acc32;               // one byte per k-mer after the first -B, read r at r << ashift
    uint32_t* info; unsigned long long* candl; uint32_t* initc;
    uint32_t ashift, pshift, P;
};
struct VotePartPlan { uint32_t P, pshift, ashift, cap, cap_warp, grid, ctas_per_sm; size_t o_ntup; size_t o_info, o_initc, o_direct, o_candl, o_acc, o_tuples, bytes; };
bool votepart_plan(const cid_index* idx, const cid_readid_params& p, int cap_bases, uint32_t maxocc, uint64_t reads, VotePartPlan* out);
int launch_readid_vote_part(cid_index* idx, cudaStream_t st, const ReadSrc& rsrc, const uint64_t* d_seq_offs,
                            const uint64_t* d_read_offs, uint64_t r0, uint64_t nr, uint32_t kitem, const ModS& mods, int cap,
                            uint32_t maxocc, const uint16_t* ord16, const uint8_t* ord8, const uint16_t* d_ent16,
                            const uint32_t* d_n_set, const cid_readid_params& p, const VotePartPlan& pl, uint8_t* d_scratch,
                            uint32_t* d_flags, uint32_t* d_rep_n, uint32_t* d_rep_colour, uint32_t* d_rep_count, VotePart* vp_out);

// device side of the vote; reads it cannot decide bit-exactly are appended to `list` for the host vote
int launch_readid_classify(cid_ctx* ctx, cudaStream_t st, uint64_t r0, uint64_t nreads, uint32_t N, uint32_t rep_cap,
                           const uint32_t* d_n_set, const uint32_t* d_flags, const uint32_t* d_rep_n,
                           const uint32_t* d_rep_colour, const uint32_t* d_rep_count, const double* d_fp, double fp_correct,
                           uint32_t group_width, int32_t* d_kind, uint32_t* d_hits, uint32_t* d_n_top, uint32_t* d_top,
                           uint32_t top_cap, uint32_t* d_list, uint32_t* d_list_cursor);

// host vote (cid_host_vote.cpp): read_id_mt_pe.rs:187-251 kmer_poll_plus over a chunk of reads
struct VoteParams { uint32_t n_colors = 0; double fp_correct = 1e-3; uint32_t group_width = 16; std::vector<double> fp; std::vector<uint64_t> key_hash; };
void vote_params_init(VoteParams& vp, uint64_t bloom_size, uint32_t num_hash, uint32_t n_colors,
                      const uint64_t* n_ref_by_colour, double fp_correct, uint32_t group_width);
void classify_chunk(const VoteParams& vp, uint64_t nreads, const uint32_t* n_set, const uint32_t* flags,
                    const uint32_t* rep_n, const uint32_t* rep_colour, const uint32_t* rep_count, uint32_t rep_cap,
                    int threads, int32_t* kind, uint32_t* hits, uint32_t* n_top, uint32_t* top, uint32_t top_cap);

// host-pointer read_id pipeline (cid_readid_pipe.cu)
void default_readid_params(cid_readid_params& p, const cid_readid_params* in, uint32_t N);
void readid_pipe_destroy(cid_ctx* ctx);

}  // namespace cid
