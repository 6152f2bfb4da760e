/* colorid_b200 — C ABI of the B200-native BIGSI hot path (build / search / read_id).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  The reference
 * (hcdenbakker/colorid, Rust) has no FFI layer of its own; each entry point below names the
 * reference `pub fn` whose inner loop it replaces (file:line relative to the reference root),
 * i.e. what a Rust `extern "C"` block in the reference would bind.
 *
 * Conventions
 *  - Every function returns 0 on success or a negative CID_E_* code; cid_last_error() gives text.
 *  - Sequences are passed as one byte array `bases` plus `seq_offs[nseq+1]` (byte offsets).
 *    Grouping arrays (`query_offs`, `read_offs`, …) index into the sequence list.
 *  - `*_dev` variants take DEVICE pointers and a CUDA stream (as void*) and never synchronise
 *    the host; the plain variants take HOST pointers, copy in/out and return when results are
 *    in host memory.
 *  - A cid_ctx is single-owner (not thread-safe).  The library owns all device memory behind
 *    the handles; the caller owns every buffer it passes in.
 *  - There is NO CPU fallback: with no usable CUDA device every call fails with CID_E_CUDA.
 */
#ifndef COLORID_B200_H
#define COLORID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cid_ctx cid_ctx;
typedef struct cid_index cid_index;

enum {
    CID_OK = 0,
    CID_E_INVALID = -1,      /* bad argument (k out of range, null pointer, …) */
    CID_E_CUDA = -2,         /* CUDA runtime error / no device */
    CID_E_NOMEM = -3,        /* device or host allocation failed */
    CID_E_UNSUPPORTED = -4,  /* input the device path does not cover yet (see message) */
    CID_E_REF_PANIC = -5,    /* the reference would panic on this input (e.g. auto_cutoff underflow) */
    CID_E_CAPACITY = -6      /* an internal table/output capacity was exceeded */
};

/* k-mer extraction semantics (kmer.rs) */
enum {
    CID_SEQ_FASTA = 0, /* kmerize_vector  kmer.rs:87-125 : has_no_n, compare raw case, then uppercase */
    CID_SEQ_FASTQ = 1, /* kmers_from_fq_qual / kmers_fq_pe_qual kmer.rs:461-510,581-655 : has_no_n, raw case */
    CID_SEQ_STRING = 2 /* kmerize_string kmer.rs:271-299 (-s -m): NO has_no_n, compare raw case, then uppercase.
                          A window holding a byte outside ACGTacgt (N, IUPAC, U) is a k-mer too: hashed byte-wise
                          for records of <= 8192 positions; longer records with such bytes: CID_E_UNSUPPORTED. */
};

int cid_version(void);
const char* cid_last_error(void);

/* Page-locked host buffers for callers that fill large batches (reads, queries): copies from them overlap
 * with kernels; pageable memory works everywhere but serialises the H2D copies. */
int cid_host_alloc(size_t bytes, void** out);
void cid_host_free(void* p);

/* ---- context ---------------------------------------------------------------------------- */
int cid_ctx_create(int device, cid_ctx** out);
void cid_ctx_destroy(cid_ctx* ctx);
int cid_ctx_device(const cid_ctx* ctx);
/* Counters of kernels launched by this library since ctx creation (for bench `gpu_launches`). */
uint64_t cid_ctx_launch_count(const cid_ctx* ctx);
/* Diagnostic device counters, read and reset: "readid_gather_rows" = matrix rows the read_id vote kernel
 * actually read (bench.py's random-access rate). */
int cid_ctx_read_counter(cid_ctx* ctx, const char* name, uint64_t* value);
/* Options by name.  None of them changes a result; unknown names return CID_E_INVALID.
 *   host_threads           host threads of the read_id vote (0 = all cores; main.rs:718 rayon pool size `-t`)
 *   host_ranks             how many ranks (processes, or shards of a cid_mg) share this host with the context (default 1): the
 *                          read_id pipeline sizes its chunks for the host -> device rate a rank can expect
 *   readid_report_steps    1: read_id report colours carry their insertion step in bits 20..31 (column-sharded read_id:
 *                          see cid_merge_shard_reports); such reports must be merged before they are classified
 *   tuning:  readid_chunk_reads / readid_chunk0_reads / readid_chunk_growth_pct (largest and first pipeline chunk of the
 *            host-pointer read_id calls and the size of a chunk relative to its predecessor in percent; 0 = automatic:
 *            32,768 -> x1.5 -> 262,144 reads for ASCII input, 65,536 -> x4 -> 2^20 for packed reads), readid_part_ctas (CTAs
 *            per SM of the partitioned vote's scan kernel),
 *            readid_streams (2 = chunks of cid_read_id_batch_dev over two internal streams), readid_kmerize_ctas /
 *            readid_vote_ctas (CTAs per SM, 0 = fill the GPU), readid_serialize, build_table_div / query_table_div (first count
 *            table of a read set = k-mer positions / this; 0 = adaptive / always the safe size), query_table_min_slots,
 *            gather_l2_64b
 *   parity aids (force the slower of two equivalent paths): build_packed, build_set, query_front, query_fused,
 *            query_compact (0 never / 1 large tables / 2 always), uniq_device,
 *            readid_vote_part (narrow-row read_id vote partitioned by row range so that its gathers hit L2: 0 never / 1 when the
 *            matrix spans >= 3 windows of 64 MB and the chunk has >= 16,384 reads / 2 always), readid_part_shift (log2 rows per
 *            window, 0 = 64 MB), readid_part_cap (tuples per window bucket, 0 = sized from the chunk) */
int cid_ctx_set_option(cid_ctx* ctx, const char* name, int64_t value);
/* Per-kernel device timing with CUDA events on the launching stream (for bench.py's roofline).
 * cid_ctx_profile(ctx, 1) resets the accumulators and enables timing, (ctx, 0) disables it;
 * cid_ctx_profile_read returns name / total ms / launch count of kernel id 0..N-1
 * (CID_E_INVALID past the last id). */
int cid_ctx_profile(cid_ctx* ctx, int enable);
int cid_ctx_profile_read(cid_ctx* ctx, int kernel, const char** name, double* total_ms, uint64_t* launches);

/* ---- index: bigsi.rs:19-27 BigsyMapNew {bloom_size, num_hash, k_size, colors, map, n_ref_kmers}
 * Device layout: dense row-major matrix rows[bloom_size][row_words] of u32, bit c of a row =
 * (word[c/32] >> (c%32)) & 1 (bit-vec_serde/src/lib.rs:465-500).  An absent row of the reference's
 * sparse map is an all-zero row here (build.rs:123-127 never stores an all-zero row). */
int cid_index_create(cid_ctx* ctx, uint64_t bloom_size, uint32_t num_hash, uint32_t k_size, uint32_t n_colors,
                     cid_index** out);
void cid_index_destroy(cid_index* idx);
uint32_t cid_index_row_words(const cid_index* idx);    /* ceil(n_colors/32) */
uint32_t cid_index_row_stride(const cid_index* idx);   /* padded words per row on the device */
/* Hash variant of the index (simple_bloom.rs:21-24 calls the un-vendored crate `xxh3 = "0.1.1"`, Cargo.toml:9, which
 * predates the XXH3 freeze; which draft a real colorid binary computes is unpinned).  0 = stable XXH3 (xxHash >= 0.8,
 * the default).  Bits select the places where the drafts differ: 1 = avalanche multiplier PRIME64_3, 2 = avalanche
 * shift 29, 4 = 128-bit product folded by +, 8 = seed enters as (len + seed) * PRIME64_1 (inputs of 17+ bytes),
 * 16 = secret read as 32-bit words.  CID_HASH_XXH3_DRAFT_071 / _070 are the combinations recalled for the drafts of
 * xxHash 0.7.1-0.7.3 / 0.7.0.  Set it before the first build / upload / query; tools/pin_from_bxi.py finds the variant
 * that reproduces a given .bxi.  Every kernel hashes through it. */
enum { CID_HASH_XXH3_STABLE = 0, CID_HASH_XXH3_DRAFT_071 = 1, CID_HASH_XXH3_DRAFT_070 = 31 };
int cid_index_set_hash_variant(cid_index* idx, uint32_t variant);
uint32_t cid_index_hash_variant(const cid_index* idx);
/* Rows from a .bxi `map` (bigsi.rs:25): row_ids[n], words[n*row_words]. Replaces read_bigsi's heap map. */
int cid_index_upload_rows(cid_index* idx, const uint64_t* row_ids, const uint32_t* words, uint64_t nrows);
int cid_index_count_nonzero_rows(cid_index* idx, uint64_t* nrows);
/* Non-zero rows in ascending row order for save_bigsi (bigsi.rs:51-57). */
int cid_index_download_nonzero_rows(cid_index* idx, uint64_t* row_ids, uint32_t* words, uint64_t cap, uint64_t* nrows);
int cid_index_download_dense(cid_index* idx, uint32_t* words /* bloom_size*row_words */);
/* Device pointers for multi-GPU glue (column-sharded mode ORs the row-present bitmap across ranks). */
int cid_index_device_ptrs(cid_index* idx, void** rows, void** rownz_bitmap, uint64_t* rownz_words);
int cid_index_refresh_rownz(cid_index* idx);   /* recompute the row-present bitmap from the matrix */
/* Column-sharded mode: declare that the bitmap now holds the OR over all shards. */
int cid_index_set_rownz_global(cid_index* idx, int is_global);

/* ---- build: build.rs:33-130 build_single / :132-256 build_multi ---------------------------
 * Phase 1 for one accession: canonical k-mer count map -> (auto_cutoff | cutoff) -> clean_map ->
 * n_ref_kmers -> Bloom insert (simple_bloom.rs:19-26) into this colour's bitset.
 * cutoff == -1: FASTA -> no filter (build.rs:86-87); FASTQ -> auto_cutoff (build.rs:56-58). */
int cid_build_accession(cid_index* idx, uint32_t colour, const char* bases, const uint64_t* seq_offs, uint64_t nseq,
                        int seq_mode, int64_t cutoff, uint64_t* n_ref_kmers, int64_t* cutoff_used);
int cid_build_accession_dev(cid_index* idx, uint32_t colour, const char* d_bases, const uint64_t* d_seq_offs,
                            uint64_t nseq, uint64_t nbases, int seq_mode, int64_t cutoff, uint64_t* n_ref_kmers,
                            int64_t* cutoff_used);
/* ---- minimizer indexes (.mxi): bigsi.rs:40-49 BigsyMapMiniNew, `colorid build -m -v M` (main.rs:485-533) -------
 * cid_index_set_minimizer(idx, M) (1 <= M <= k_size; M > k_size is CID_E_REF_PANIC: find_minimizer slices
 * seq[..m], kmer.rs:974) turns the index into a minimizer index: the items hashed into the Bloom filters are then
 * kmer.rs:971-986 find_minimizer(canonical k-mer, M) strings of M bytes instead of the k-mers.  cid_read_id_* then
 * follow kmer.rs:363-394 minimerize_vector_skip_n_set (read_id_mt_pe.rs:318-322, m != 0): per-mate `len >= k`
 * guard, lower-case accepted (minimizers are upper-cased), the set holds minimizers; n_set counts them.
 * cid_query_* return CID_E_UNSUPPORTED with the reference's message (main.rs:569-573).  The two builders of the
 * reference differ, so both are offered:
 *   CID_MINI_OF_KMERS  build.rs:396-492 build_single_mini (-t 1): the k-mer count map and cutoff of the plain build,
 *                      then insert(find_minimizer(kmer, M)) per surviving k-mer; n_ref_kmers = distinct k-mers
 *                      (the reference records it for FASTA accessions only: the caller drops it for FASTQ).
 *   CID_MINI_COUNTED   build.rs:258-394 build_multi_mini (-t > 1): a count map of minimizers (one count per k-mer
 *                      position; kmer.rs:328-361 FASTA, upper-cased after the choice; :694-824 FASTQ), auto_cutoff /
 *                      clean_map on those counts, insert of the survivors; n_ref_kmers = surviving minimizers. */
enum { CID_MINI_OF_KMERS = 0, CID_MINI_COUNTED = 1 };
int cid_index_set_minimizer(cid_index* idx, uint32_t m_size);
uint32_t cid_index_minimizer(const cid_index* idx);
int cid_build_accession_mini(cid_index* idx, uint32_t colour, const char* bases, const uint64_t* seq_offs, uint64_t nseq,
                             int seq_mode, int64_t cutoff, int variant, uint64_t* n_ref_kmers, int64_t* cutoff_used);
/* Phase 2: transposition of the per-colour bitsets into the row-major matrix (build.rs:116-128). */
int cid_build_finalize(cid_index* idx);

/* ---- search: batch_search_pe.rs:9-179 batch_search (default and -g) -----------------------
 * One query = sequences [query_offs[q], query_offs[q+1]) (one file in the reference).
 * filter < 0 -> auto_cutoff per query (ignored for FASTA with gene_search: clean_map(0), :112-113).
 * Outputs (host): counts[nq*n_colors], num_kmers[nq]; optional (may be NULL) unique-hit
 * summaries for generate_report (reports.rs:8-48): uniq_n, uniq_sum, uniq_mode [nq*n_colors];
 * cutoff_used[nq]. */
int cid_query_counts(cid_index* idx, const char* bases, const uint64_t* seq_offs, uint64_t nseq,
                     const uint64_t* query_offs, uint64_t nq, int seq_mode, int gene_search, int64_t filter,
                     uint32_t* counts, uint64_t* num_kmers, uint64_t* uniq_n, uint64_t* uniq_sum, uint64_t* uniq_mode,
                     int64_t* cutoff_used);
/* Device-resident variant for gene search (filter fixed at 0, no unique-hit summaries):
 * d_counts[nq*n_colors] u32 and d_num_kmers[nq] u64 are device buffers. */
int cid_query_counts_dev(cid_index* idx, const char* d_bases, const uint64_t* d_seq_offs, uint64_t nseq,
                         uint64_t nbases, const uint64_t* d_query_offs, const uint64_t* h_query_offs,
                         const uint64_t* h_seq_offs, uint64_t nq, int seq_mode, uint32_t* d_counts,
                         uint64_t* d_num_kmers, void* stream);

/* Column-sharded index (GPU g holds the accessions [col_offset, col_offset + n_colors) of every row): every
 * GPU gathers all k-mers from its slice and the per-query counts are disjoint column slices of the full [nq][n_total]
 * result.  Instead of a collective after the kernel, the gather kernel itself adds each count into slot
 * [query][col_offset + accession] of EVERY destination buffer -- the caller's own and its peers' (device memory of the
 * other GPUs opened with cid_ipc_open: stores travel over NVLink/NVSwitch while the kernel is still gathering).
 * Protocol per pass: every rank zeroes its own destination, barrier, this call on every rank, stream sync, barrier;
 * then each rank holds the complete [nq][n_total] counts.  cid_dev_alloc gives cudaMalloc memory (what CUDA IPC needs),
 * cid_ipc_export / cid_ipc_open move a 64-byte handle between the per-GPU processes.  n_dest <= 8. */
int cid_dev_alloc(cid_ctx* ctx, size_t bytes, void** dptr);
void cid_dev_free(cid_ctx* ctx, void* dptr);
int cid_ipc_export(cid_ctx* ctx, const void* dptr, uint8_t handle[64]);
int cid_ipc_open(cid_ctx* ctx, const uint8_t handle[64], void** dptr);
int cid_ipc_close(cid_ctx* ctx, void* dptr);
int cid_query_counts_sharded_dev(cid_index* idx, const char* d_bases, const uint64_t* d_seq_offs, uint64_t nseq,
                                 uint64_t nbases, const uint64_t* d_query_offs, const uint64_t* h_query_offs,
                                 const uint64_t* h_seq_offs, uint64_t nq, int seq_mode, void* const* d_dest_counts,
                                 uint32_t n_dest, uint32_t n_total, uint32_t col_offset, uint64_t* d_num_kmers, void* stream);

/* Column-sharded DEFAULT report.  "This k-mer hits exactly one accession" (batch_search_pe.rs:75-82, the
 * unique-hit columns of reports.rs:8-48) is a statement about the whole row, so the shards exchange one byte per k-mer:
 *   1. cid_query_survivors on ONE shard: k-mer counting, per-query filter (auto_cutoff / fixed / -g), survivors compacted
 *      into a dense list of 16-byte slots {u64 packed k-mer, u32 count, u32 pad} on the device: query q holds surv[q] slots
 *      starting at the prefix sum of surv[].  *d_slots stays valid until the next search call on that context.  The list
 *      (and surv[]) is then copied to every shard (NCCL broadcast), so that list index i means the same k-mer everywhere.
 *   2. cid_query_slots_counts_dev on every shard: d_counts[nq * n_colors] (hits per accession of this shard: disjoint
 *      column slices, gathered like cid_query_counts_sharded_dev's), d_num_kmers[nq], and per list index d_pc[i] =
 *      min(popcount of the AND row over this shard, 2) and d_col[i] = the single accession hit (or 0xFFFFFFFF).
 *   3. the shards sum their d_pc arrays (one all-reduce of a byte per k-mer) and each calls cid_query_slots_uniq_dev with its
 *      own d_pc / d_col and the sum: uniq_n / uniq_sum / uniq_mode [nq * n_colors] (host) for ITS accessions, from the
 *      k-mers whose local and global popcounts are both 1.  All device buffers of 2 and 3 are the caller's. */
int cid_query_survivors(cid_index* idx, const char* bases, const uint64_t* seq_offs, uint64_t nseq,
                        const uint64_t* query_offs, uint64_t nq, int seq_mode, int gene_search, int64_t filter,
                        void** d_slots, uint64_t* surv, int64_t* cutoff_used);
int cid_query_slots_counts_dev(cid_index* idx, const void* d_slots, const uint64_t* surv, uint64_t nq, uint32_t* d_counts,
                               uint64_t* d_num_kmers, uint8_t* d_pc, uint32_t* d_col, void* stream);
int cid_query_slots_uniq_dev(cid_index* idx, const void* d_slots, const uint64_t* surv, uint64_t nq,
                             const uint8_t* d_pc_local, const uint8_t* d_pc_sum, const uint32_t* d_col, uint64_t* uniq_n,
                             uint64_t* uniq_sum, uint64_t* uniq_mode, void* stream);

/* ---- perfect search: perfect_search.rs:6-60 batch_search ----------------------------------
 * AND of all num_hash rows of all distinct k-mers of the query (kmerize_vector semantics).
 * and_rows[nq*row_words]; status[q]: 0 = AND valid, 1 = "No perfect hits!" (a row is absent),
 * 2 = no k-mers in query; n_kmers[q] = distinct canonical k-mers. */
int cid_query_perfect(cid_index* idx, const char* bases, const uint64_t* seq_offs, uint64_t nseq,
                      const uint64_t* query_offs, uint64_t nq, uint32_t* and_rows, uint8_t* status,
                      uint64_t* n_kmers);

/* perfect_search.rs:62-120 batch_search_mf (-s -m): one query per FASTA RECORD (sequence s), k-mers by
 * kmer.rs:271-299 kmerize_string.  status[s]: 0 = AND valid, 1 = "No perfect hits!", 2 = "Warning! no
 * kmers in query" (record shorter than k, kmer.rs:275-276).  and_rows[nseq*row_words], n_kmers[nseq]. */
int cid_query_perfect_mf(cid_index* idx, const char* bases, const uint64_t* seq_offs, uint64_t nseq,
                         uint32_t* and_rows, uint8_t* status, uint64_t* n_kmers);

/* ---- read_id: read_id_mt_pe.rs:282-363 parallel_vec (m == 0) -------------------------------
 * One read = sequences [read_offs[r], read_offs[r+1]) (1 or 2 mates).  If `quals` is non-NULL it
 * has the same layout as `bases` and seq.rs:36-56 qual_mask(seq, qual, qual_offset) is applied on
 * the device; otherwise `bases` must already be masked.
 * The device computes, per read: the distinct canonical k-mer set (kmer.rs:221-243), its
 * FnvHashSet iteration order, and search_index / search_index_classic (read_id_mt_pe.rs:66-165).
 * Per read r the report is written in final_report INSERTION order:
 *   rep_n[r] entries at rep_colour/rep_count[r*rep_cap ..]; colour == n_colors is the "no hit" key.
 * flags[r]: bit0 = too_short (mate 1 shorter than k), bit1 = reference would panic (a later mate
 * shorter than k-1), bit2 = report truncated at rep_cap, bits 8..23 = number of k-mers whose rows
 * were gathered (up to and including the first miss; the algorithmic traffic of SURVEY §8d). */
typedef struct cid_readid_params {
    uint32_t downsample;        /* -d, kmer.rs:229 step_by(d) */
    uint32_t start_sample;      /* -B bitvector_sample; 0 = search_index_classic */
    uint32_t qual_offset;       /* -Q (only used when quals != NULL) */
    uint32_t group_width;       /* hashbrown SIMD group width to emulate: 16 (x86_64) or 8 */
    uint32_t reserve_before_find; /* 1 = hashbrown >= 0.14 insert rule (default), 0 = older rule */
    uint32_t rep_cap;           /* report entries kept per read */
} cid_readid_params;

int cid_read_id_batch(cid_index* idx, const char* bases, const char* quals, const uint64_t* seq_offs, uint64_t nseq,
                      const uint64_t* read_offs, uint64_t nreads, const cid_readid_params* p, uint32_t* n_set,
                      uint32_t* flags, uint32_t* rep_n, uint32_t* rep_colour, uint32_t* rep_count);
/* The whole of parallel_vec (read_id_mt_pe.rs:282-363) for one batch: reads in, one classification
 * per read out (the tuple the reference writes to PREFIX_reads.txt, :779-788).  Internally the batch
 * is cut into chunks whose H2D copy, kernels, D2H copy and host-side kmer_poll_plus overlap; results
 * are identical to cid_read_id_batch followed by cid_classify_reads.  kind/hits/n_set/n_top are
 * [nreads]; top (may be NULL) is [nreads*top_cap].  p->rep_cap is ignored (reports are never truncated). */
int cid_read_id_classify(cid_index* idx, const char* bases, const char* quals, const uint64_t* seq_offs, uint64_t nseq,
                         const uint64_t* read_offs, uint64_t nreads, const cid_readid_params* p,
                         const uint64_t* n_ref_by_colour, double fp_correct, int32_t* kind, uint32_t* hits,
                         uint32_t* n_set, uint32_t* n_top, uint32_t* top, uint32_t top_cap);
/* Device-resident variant; all pointers are device memory.  h_max_read_bases = longest read (sum
 * of mates, bases) sizes the shared-memory tile; h_max_kmers = largest number of k-mer start
 * positions of any read (0 = derive a bound from h_max_read_bases). */
int cid_read_id_batch_dev(cid_index* idx, const char* d_bases, const char* d_quals, const uint64_t* d_seq_offs,
                          uint64_t nseq, uint64_t nbases, const uint64_t* d_read_offs, uint64_t nreads,
                          uint32_t h_max_read_bases, uint32_t h_max_kmers, const cid_readid_params* p,
                          uint32_t* d_n_set, uint32_t* d_flags, uint32_t* d_rep_n, uint32_t* d_rep_colour,
                          uint32_t* d_rep_count, void* stream);
/* ---- packed reads ------------------------------------------------------------------------------------------------------
 * The reference applies seq.rs:36-56 qual_mask on the host while it parses (read_id_mt_pe.rs:733-760).  cid_pack_reads does
 * the same and emits per read (all mates concatenated, L bases) the planes the kernels work on:
 *     codes[ceil(L/16)]  2 bits per base (A0 C1 G2 T3), 16 bases per u32, first base in bits 31:30
 *     bad  [ceil(L/32)]  bit j (LSB first) = base j is not one of ACGTacgt or was masked; bits >= L are set
 *     lower[ceil(L/32)]  bit j = base j is a lower-case acgt -- present iff *pack_flags & CID_PACK_LOWER (some read has one)
 * words[word_offs[r] .. word_offs[r+1]) belongs to read r.  3 bits per base instead of the 16 of ASCII + qualities: a quarter
 * of the host->device bytes, and no masking / packing left for the kernels.  Lossless for read_id (a non-base only has to be
 * one: kmer.rs:221-243, seq.rs:59-70).  `threads` <= 0: all cores.  cid_pack_words_bound sizes `words`.
 * cid_read_id_classify_packed / cid_read_id_batch_packed_dev == cid_read_id_classify / cid_read_id_batch_dev on the reads
 * the planes spell (seq_offs / read_offs as there, in bases; p->qual_offset is ignored: the packer masked). */
enum { CID_PACK_LOWER = 1 };
uint64_t cid_pack_words_bound(const uint64_t* seq_offs, const uint64_t* read_offs, uint64_t nreads, int with_lower);
int cid_pack_reads(const char* bases, const char* quals, const uint64_t* seq_offs, uint64_t nseq, const uint64_t* read_offs,
                   uint64_t nreads, uint32_t qual_offset, int threads, uint32_t* words, uint64_t words_cap, uint64_t* word_offs,
                   uint32_t* pack_flags);
int cid_read_id_classify_packed(cid_index* idx, const uint32_t* words, const uint64_t* word_offs, uint32_t pack_flags,
                                const uint64_t* seq_offs, uint64_t nseq, const uint64_t* read_offs, uint64_t nreads,
                                const cid_readid_params* p, const uint64_t* n_ref_by_colour, double fp_correct, int32_t* kind,
                                uint32_t* hits, uint32_t* n_set, uint32_t* n_top, uint32_t* top, uint32_t top_cap);
int cid_read_id_batch_packed_dev(cid_index* idx, const uint32_t* d_words, const uint64_t* d_word_offs, uint32_t pack_flags,
                                 const uint64_t* d_seq_offs, uint64_t nseq, const uint64_t* d_read_offs, uint64_t nreads,
                                 uint32_t h_max_read_bases, uint32_t h_max_kmers, const cid_readid_params* p, uint32_t* d_n_set,
                                 uint32_t* d_flags, uint32_t* d_rep_n, uint32_t* d_rep_colour, uint32_t* d_rep_count, void* stream);

/* Debug/parity hook: the emulated FnvHashSet<String> iteration order of each read's k-mer set as
 * (mate index, position of first occurrence).  order_n[r] entries at [r*order_cap ..].
 * Minimizer index: the set holds minimizers; order_pos is the start of a window of m_size bases of that mate which
 * spells the minimizer (bit 7 of order_seq set) or its reverse complement (bit 7 clear). */
int cid_read_kmer_order(cid_index* idx, const char* bases, const uint64_t* seq_offs, uint64_t nseq,
                        const uint64_t* read_offs, uint64_t nreads, const cid_readid_params* p, uint32_t order_cap,
                        uint32_t* order_n, uint8_t* order_seq, uint16_t* order_pos);
/* The same with 32-bit positions (reads longer than 65,535 bases: contigs through stream_fasta). */
int cid_read_kmer_order32(cid_index* idx, const char* bases, const uint64_t* seq_offs, uint64_t nseq,
                          const uint64_t* read_offs, uint64_t nreads, const cid_readid_params* p, uint32_t order_cap,
                          uint32_t* order_n, uint8_t* order_seq, uint32_t* order_pos);

/* ---- read_id vote, host side: read_id_mt_pe.rs:187-251 kmer_poll_plus (+ :168-181, :695-698, :18-38)
 * Takes the per-read reports produced by cid_read_id_batch (insertion order) and classifies every
 * read with `threads` host threads (<=0: all cores).  kind[r] is one of CID_CLS_*; hits[r] and
 * n_set[r] are fields 3 and 4 of PREFIX_reads.txt; top[r*top_cap ..] lists the n_top[r] best
 * accessions in the order the reference joins them with ','. */
enum {
    CID_CLS_TOO_SHORT = 0,       /* "too_short", 0, 0, "accept", 0 */
    CID_CLS_NO_HITS = 1,         /* "no_hits", 0, n_set, "accept", 0 */
    CID_CLS_NO_SIGNIFICANT = 2,  /* "no_significant_hits", 0, n_set, "reject", 0 */
    CID_CLS_ACCEPT = 3,          /* name, hits, n_set, "accept", 1 */
    CID_CLS_REJECT_MULTI = 4,    /* "a,b,..", hits, n_set, "reject", n_top */
    CID_CLS_REF_PANIC = 5        /* the reference panics on this read (a later mate shorter than k-1) */
};
int cid_classify_reads(uint64_t bloom_size, uint32_t num_hash, uint32_t n_colors, const uint64_t* n_ref_by_colour,
                       double fp_correct, uint32_t group_width, uint64_t nreads, const uint32_t* n_set,
                       const uint32_t* flags, const uint32_t* rep_n, const uint32_t* rep_colour,
                       const uint32_t* rep_count, uint32_t rep_cap, int threads, int32_t* kind, uint32_t* hits,
                       uint32_t* n_top, uint32_t* top, uint32_t top_cap);
/* Column-sharded read_id (read_id_mt_pe.rs:104-165 search_index is per-colour once the first absent row is known):
 * each shard runs cid_read_id_batch on the same reads with cid_ctx_set_option(ctx, "readid_report_steps", 1) and the
 * row-present bitmaps merged (cid_index_set_rownz_global); its rep_colour entries are then colour | step << 20.  This
 * host call merges the n_shards reports of every read into the report of the unsharded index (global colours
 * shard_col_offset[s] + colour, the reference's insertion order, the "no hit" key = n_total last), ready for
 * cid_classify_reads with n_colors = n_total.  n_set and flags are the same on every shard.  out_flags may be NULL. */
int cid_merge_shard_reports(uint32_t n_shards, const uint32_t* shard_n_colors, const uint32_t* shard_col_offset,
                            uint64_t nreads, const uint32_t* const* rep_n, const uint32_t* const* rep_colour,
                            const uint32_t* const* rep_count, uint32_t rep_cap_in, uint32_t n_total, uint32_t* out_rep_n,
                            uint32_t* out_colour, uint32_t* out_count, uint32_t rep_cap_out, uint32_t* out_flags);
/* ---- multi-GPU: one index over several GPUs of one node (BASELINE north_star) ------------------------------
 * The reference scales with rayon threads inside one process (main.rs:718, build.rs:146-149); here a cid_mg is a set of GPUs
 * behind the same host-pointer calls, one host thread per GPU inside the library.
 *   CID_MG_REPLICATED  every GPU holds the whole matrix (it fits one GPU: C1-C4); reads / queries are dealt to the GPUs in
 *                      contiguous ranges, no data-path exchange; a build fills column slices on their owner GPUs and copies
 *                      them into every replica at cid_mg_build_finalize (peer copies);
 *   CID_MG_COLUMNS     GPU g holds the accessions [col_offset, col_offset + n_colors) of every row (whole 32-accession word
 *                      columns: C5); every GPU processes ALL k-mers against its slice.  Counts and AND rows are disjoint column
 *                      slices of the result; the default report exchanges one byte per k-mer (popcounts); read_id merges the
 *                      shards' sparse reports in insertion order; the row-present bitmaps are OR-ed after build / upload.
 * Results are identical to the single-GPU entry points of the same name in either mode.  `devices` may list a device more
 * than once (several shards on one GPU).  A cid_mg is single-owner except cid_mg_build_accession, which may be called from
 * several host threads at once (accessions of different shards build concurrently). */
typedef struct cid_mg cid_mg;
enum { CID_MG_REPLICATED = 0, CID_MG_COLUMNS = 1 };
int cid_mg_create(const int* devices, int ndev, int shard_mode, cid_mg** out);
void cid_mg_destroy(cid_mg* mg);
int cid_mg_n_shards(const cid_mg* mg);
int cid_mg_mode(const cid_mg* mg);
cid_ctx* cid_mg_ctx(cid_mg* mg, int shard);                 /* the shard's context / index for the single-GPU calls */
cid_index* cid_mg_index(cid_mg* mg, int shard);
int cid_mg_shard_columns(const cid_mg* mg, int shard, uint32_t* col_offset, uint32_t* n_colors);
int cid_mg_shard_of_colour(const cid_mg* mg, uint32_t colour);
int cid_mg_set_option(cid_mg* mg, const char* name, int64_t value);      /* cid_ctx_set_option on every shard */
uint64_t cid_mg_launch_count(const cid_mg* mg);
int cid_mg_index_create(cid_mg* mg, uint64_t bloom_size, uint32_t num_hash, uint32_t k_size, uint32_t n_colors);
int cid_mg_index_set_minimizer(cid_mg* mg, uint32_t m_size);
int cid_mg_index_set_hash_variant(cid_mg* mg, uint32_t variant);
int cid_mg_index_upload_rows(cid_mg* mg, const uint64_t* row_ids, const uint32_t* words /* nrows * ceil(N/32) */, uint64_t nrows);
int cid_mg_index_count_nonzero_rows(cid_mg* mg, uint64_t* nrows);
int cid_mg_index_download_nonzero_rows(cid_mg* mg, uint64_t* row_ids, uint32_t* words, uint64_t cap, uint64_t* nrows);
/* build.rs:33-256; mini_variant: -1 = k-mer index, else CID_MINI_OF_KMERS / CID_MINI_COUNTED (cid_build_accession_mini) */
int cid_mg_build_accession(cid_mg* mg, uint32_t colour, const char* bases, const uint64_t* seq_offs, uint64_t nseq, int seq_mode,
                           int64_t cutoff, int mini_variant, uint64_t* n_ref_kmers, int64_t* cutoff_used);
int cid_mg_build_finalize(cid_mg* mg);
/* batch_search_pe.rs:9-179, perfect_search.rs:6-120, read_id_mt_pe.rs:282-363: arguments as cid_query_counts,
 * cid_query_perfect[_mf] and cid_read_id_classify; all result arrays are full width (n_colors of the whole index) */
int cid_mg_query_counts(cid_mg* mg, const char* bases, const uint64_t* seq_offs, uint64_t nseq, const uint64_t* query_offs,
                        uint64_t nq, int seq_mode, int gene_search, int64_t filter, uint32_t* counts, uint64_t* num_kmers,
                        uint64_t* uniq_n, uint64_t* uniq_sum, uint64_t* uniq_mode, int64_t* cutoff_used);
int cid_mg_query_perfect(cid_mg* mg, const char* bases, const uint64_t* seq_offs, uint64_t nseq, const uint64_t* query_offs,
                         uint64_t nq, uint32_t* and_rows, uint8_t* status, uint64_t* n_kmers);
int cid_mg_query_perfect_mf(cid_mg* mg, const char* bases, const uint64_t* seq_offs, uint64_t nseq, uint32_t* and_rows,
                            uint8_t* status, uint64_t* n_kmers);
int cid_mg_read_id_classify(cid_mg* mg, const char* bases, const char* quals, const uint64_t* seq_offs, uint64_t nseq,
                            const uint64_t* read_offs, uint64_t nreads, const cid_readid_params* p, const uint64_t* n_ref_by_colour,
                            double fp_correct, int32_t* kind, uint32_t* hits, uint32_t* n_set, uint32_t* n_top, uint32_t* top,
                            uint32_t top_cap);

/* read_id_mt_pe.rs:695-698 false_prob and the Binomial pmf used by :168-181 (parity hooks). */
double cid_false_prob(double bloom_size, double num_hash, double n_ref_kmers);
double cid_binomial_mass(uint64_t n, double p, uint64_t x);

/* Row indices of canonical k-mers given as ASCII (k bytes each; m_size bytes each for a minimizer index):
 * out[n*num_hash] = xxh3_64(kmer, seed=i) % bloom_size  (simple_bloom.rs:21-24).  Parity hook for the hash. */
int cid_hash_kmers(cid_index* idx, const char* kmers, uint64_t n, uint64_t* row_ids);

#ifdef __cplusplus
}
#endif
#endif /* COLORID_B200_H */
