// cid_mg: one BIGSI index over several GPUs of one node, behind the same host-pointer calls as the single-GPU entry points
// (include/colorid_b200.h, "multi-GPU").  SURVEY 8e / BASELINE north_star: the signature matrix is REPLICATED when it fits one
// GPU, with reads / queries split across the GPUs and no data-path exchange; it is COLUMN-SHARDED otherwise -- GPU g holds the
// accessions [col0[g], col0[g] + ncol[g]) (whole 32-accession word columns) of every row, every GPU processes all k-mers
// against its slice, and the per-accession results are disjoint column slices of the full result:
//   * search -g / -f 0 / -s: counts and AND rows are written side by side into the caller's full-width arrays (through the
//     gather kernel's own stores into a host-mapped result where the row shape allows it: cid_query_counts_sharded_dev);
//   * default report: "this k-mer hits exactly one accession" needs the popcount of the whole row: one shard counts and
//     filters the k-mers, the survivor list is copied to the peers, each shard gathers its slice and the per-k-mer popcounts
//     (one byte each) are summed across shards (cid_query_survivors / _slots_counts_dev / _slots_uniq_dev);
//   * read_id: search_index is per colour once the first absent row is known, and the OR-ed row-present bitmap answers that
//     identically on every shard; the shards' sparse reports are merged in insertion order (cid_merge_shard_reports);
//   * build: every accession goes to the GPU that owns its column; the row-present bitmaps are OR-ed once at the end.  A
//     replicated index is built the same way and its column slices are then copied into every replica (peer copies).
// One host thread per shard runs that shard's calls; a device may be listed more than once (several shards on one GPU: how
// the single-GPU test box exercises this file).
#include <algorithm>
#include <atomic>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "cid_internal.h"

struct cid_mg {
    int mode = CID_MG_COLUMNS;
    uint32_t n = 0;                       // shards (contexts)
    std::vector<int> dev;
    std::vector<cid_ctx*> ctx;
    std::vector<cid_index*> ix;           // COLUMNS: the slices; REPLICATED: the replicas (after finalize / upload)
    std::vector<cid_index*> slice;        // REPLICATED, while building: column slices, merged at finalize
    std::vector<std::mutex> mu;           // a cid_ctx is single-owner
    uint64_t S = 0;
    uint32_t H = 0, k = 0, N = 0, W = 0, m = 0, hv = 0;
    uint32_t ns = 0;                      // shards that own columns (<= n: an index of fewer than n word columns)
    std::vector<uint32_t> col0, ncol, w0, nw;
    explicit cid_mg(size_t nn) : mu(nn) {}
};

namespace {
using namespace cid;

// fn(g) on one thread per shard; the first failure wins and its message becomes this thread's cid_last_error()
int parallel(uint32_t n, const std::function<int(uint32_t)>& fn) {
    std::vector<int> rc(n, CID_OK);
    std::vector<std::string> msg(n);
    if (n == 1) { return fn(0); }
    std::vector<std::thread> th;
    for (uint32_t g = 0; g < n; g++)
        th.emplace_back([&, g] { rc[g] = fn(g); if (rc[g] != CID_OK) msg[g] = cid_last_error(); });
    for (auto& t : th) t.join();
    for (uint32_t g = 0; g < n; g++) if (rc[g] != CID_OK) { set_error("shard %u: %s", g, msg[g].c_str()); return rc[g]; }
    return CID_OK;
}

void plan_columns(cid_mg* mg) {
    mg->W = (mg->N + 31) / 32;
    const uint32_t wpg = (mg->W + mg->n - 1) / mg->n;          // word columns per shard
    mg->col0.assign(mg->n, 0); mg->ncol.assign(mg->n, 0); mg->w0.assign(mg->n, 0); mg->nw.assign(mg->n, 0);
    mg->ns = 0;
    for (uint32_t g = 0; g < mg->n; g++) {
        const uint32_t a = std::min(mg->W, g * wpg), b = std::min(mg->W, (g + 1) * wpg);
        mg->w0[g] = a; mg->nw[g] = b - a;
        mg->col0[g] = a * 32;
        mg->ncol[g] = b > a ? std::min(mg->N, b * 32) - a * 32 : 0;
        if (mg->ncol[g]) mg->ns = g + 1;
    }
}
uint32_t owner_of(const cid_mg* mg, uint32_t colour) {
    for (uint32_t g = 0; g < mg->ns; g++) if (colour >= mg->col0[g] && colour < mg->col0[g] + mg->ncol[g]) return g;
    return mg->ns;
}
int make_index(cid_mg* mg, uint32_t g, uint32_t ncol, cid_index** out) {
    CID_TRY(cid_index_create(mg->ctx[g], mg->S, mg->H, mg->k, ncol, out));
    if (mg->m) CID_TRY(cid_index_set_minimizer(*out, mg->m));
    if (mg->hv) CID_TRY(cid_index_set_hash_variant(*out, mg->hv));
    return CID_OK;
}
// OR of the shards' row-present bitmaps, installed in every shard ("absent" is a property of whole rows)
int or_rownz(cid_mg* mg, std::vector<cid_index*>& parts, uint32_t np) {
    if (np == 0) return CID_OK;
    const uint64_t words = parts[0]->rownz_words;
    std::vector<uint32_t> all(words, 0), one(words);
    for (uint32_t g = 0; g < np; g++) {
        CID_CUDA(cudaSetDevice(parts[g]->ctx->device));
        CID_CUDA(cudaMemcpy(one.data(), parts[g]->rownz, words * 4, cudaMemcpyDeviceToHost));
        for (uint64_t i = 0; i < words; i++) all[i] |= one[i];
    }
    for (uint32_t g = 0; g < np; g++) {
        CID_CUDA(cudaSetDevice(parts[g]->ctx->device));
        CID_CUDA(cudaMemcpy(parts[g]->rownz, all.data(), words * 4, cudaMemcpyHostToDevice));
        parts[g]->rownz_global = true;
    }
    (void)mg;
    return CID_OK;
}
// contiguous ranges of `units` (queries / reads) with about equal base counts, one per shard
std::vector<uint64_t> split_units(const uint64_t* seq_offs, const uint64_t* unit_offs, uint64_t nunits, uint32_t parts) {
    std::vector<uint64_t> cuts(parts + 1, nunits);
    cuts[0] = 0;
    const uint64_t total = nunits ? seq_offs[unit_offs[nunits]] - seq_offs[unit_offs[0]] : 0;
    uint64_t u = 0;
    for (uint32_t p = 1; p < parts; p++) {
        const uint64_t want = seq_offs[unit_offs[0]] + total * p / parts;
        while (u < nunits && seq_offs[unit_offs[u]] < want) u++;
        cuts[p] = u;
    }
    return cuts;
}
// the sequences of units [u0, u1) as a self-contained batch: offsets rebased to the first base / first sequence
struct SubBatch { const char* bases; const char* quals; std::vector<uint64_t> seq_offs, unit_offs; uint64_t nseq; };
SubBatch sub_batch(const char* bases, const char* quals, const uint64_t* seq_offs, const uint64_t* unit_offs, uint64_t u0, uint64_t u1) {
    SubBatch sb;
    const uint64_t s0 = unit_offs[u0], s1 = unit_offs[u1], b0 = seq_offs[s0];
    sb.bases = bases + b0;
    sb.quals = quals ? quals + b0 : nullptr;
    sb.nseq = s1 - s0;
    sb.seq_offs.resize(sb.nseq + 1);
    for (uint64_t s = s0; s <= s1; s++) sb.seq_offs[s - s0] = seq_offs[s] - b0;
    sb.unit_offs.resize(u1 - u0 + 1);
    for (uint64_t u = u0; u <= u1; u++) sb.unit_offs[u - u0] = unit_offs[u] - s0;
    return sb;
}
}  // namespace

using namespace cid;

extern "C" {

int cid_mg_create(const int* devices, int ndev, int shard_mode, cid_mg** out) {
    if (!devices || !out || ndev < 1 || ndev > 8) { set_error("cid_mg_create: 1..8 devices"); return CID_E_INVALID; }
    if (shard_mode != CID_MG_REPLICATED && shard_mode != CID_MG_COLUMNS) { set_error("cid_mg_create: bad shard mode"); return CID_E_INVALID; }
    cid_mg* mg = new cid_mg((size_t)ndev);
    mg->mode = shard_mode;
    mg->n = (uint32_t)ndev;
    mg->dev.assign(devices, devices + ndev);
    mg->ctx.assign(ndev, nullptr);
    mg->ix.assign(ndev, nullptr);
    mg->slice.assign(ndev, nullptr);
    const int rc = parallel(mg->n, [&](uint32_t g) { return cid_ctx_create(mg->dev[g], &mg->ctx[g]); });
    if (rc != CID_OK) { cid_mg_destroy(mg); return rc; }
    for (int g = 0; g < ndev; g++) mg->ctx[g]->opt_host_ranks = ndev;      // the shards' pipelines share this host's memory system
    // peer access between distinct devices: the column copies of a replicated build and the fused count exchange
    for (int a = 0; a < ndev; a++)
        for (int b = 0; b < ndev; b++) {
            if (devices[a] == devices[b]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
            if (can) { cudaSetDevice(devices[a]); cudaDeviceEnablePeerAccess(devices[b], 0); cudaGetLastError(); }
        }
    *out = mg;
    return CID_OK;
}
void cid_mg_destroy(cid_mg* mg) {
    if (!mg) return;
    for (auto* x : mg->slice) if (x) cid_index_destroy(x);
    for (auto* x : mg->ix) if (x) cid_index_destroy(x);
    for (auto* c : mg->ctx) if (c) cid_ctx_destroy(c);
    delete mg;
}
int cid_mg_n_shards(const cid_mg* mg) { return mg ? (int)mg->n : 0; }
int cid_mg_mode(const cid_mg* mg) { return mg ? mg->mode : -1; }
cid_ctx* cid_mg_ctx(cid_mg* mg, int shard) { return (mg && shard >= 0 && (uint32_t)shard < mg->n) ? mg->ctx[shard] : nullptr; }
cid_index* cid_mg_index(cid_mg* mg, int shard) { return (mg && shard >= 0 && (uint32_t)shard < mg->n) ? mg->ix[shard] : nullptr; }
int cid_mg_shard_columns(const cid_mg* mg, int shard, uint32_t* col_offset, uint32_t* n_colors) {
    if (!mg || shard < 0 || (uint32_t)shard >= mg->n || mg->col0.empty()) { set_error("cid_mg_shard_columns: bad shard"); return CID_E_INVALID; }
    const bool cols = mg->mode == CID_MG_COLUMNS;
    if (col_offset) *col_offset = cols ? mg->col0[shard] : 0;
    if (n_colors) *n_colors = cols ? mg->ncol[shard] : mg->N;
    return CID_OK;
}
int cid_mg_set_option(cid_mg* mg, const char* name, int64_t value) {
    if (!mg) { set_error("cid_mg_set_option: null"); return CID_E_INVALID; }
    for (auto* c : mg->ctx) CID_TRY(cid_ctx_set_option(c, name, value));
    return CID_OK;
}
uint64_t cid_mg_launch_count(const cid_mg* mg) {
    uint64_t n = 0;
    if (mg) for (auto* c : mg->ctx) n += cid_ctx_launch_count(c);
    return n;
}

int cid_mg_index_create(cid_mg* mg, uint64_t bloom_size, uint32_t num_hash, uint32_t k_size, uint32_t n_colors) {
    if (!mg) { set_error("cid_mg_index_create: null"); return CID_E_INVALID; }
    if (n_colors < 1) { set_error("n_colors must be >= 1"); return CID_E_INVALID; }
    for (auto*& x : mg->ix) { if (x) cid_index_destroy(x); x = nullptr; }
    for (auto*& x : mg->slice) { if (x) cid_index_destroy(x); x = nullptr; }
    mg->S = bloom_size; mg->H = num_hash; mg->k = k_size; mg->N = n_colors; mg->m = 0; mg->hv = 0;
    plan_columns(mg);
    if (mg->mode == CID_MG_COLUMNS)
        return parallel(mg->ns, [&](uint32_t g) { return make_index(mg, g, mg->ncol[g], &mg->ix[g]); });
    return CID_OK;        // replicated: replicas are made by upload_rows / build_finalize, slices on the first build call
}
int cid_mg_index_set_minimizer(cid_mg* mg, uint32_t m_size) {
    if (m_size > mg->k) { set_error("minimizer size %u larger than k-mer size %u (the reference panics)", m_size, mg->k); return CID_E_REF_PANIC; }
    mg->m = m_size;
    for (auto* x : mg->ix) if (x) CID_TRY(cid_index_set_minimizer(x, m_size));
    for (auto* x : mg->slice) if (x) CID_TRY(cid_index_set_minimizer(x, m_size));
    return CID_OK;
}
int cid_mg_index_set_hash_variant(cid_mg* mg, uint32_t variant) {
    if (variant >= 32) { set_error("hash variant %u out of range", variant); return CID_E_INVALID; }
    mg->hv = variant;
    for (auto* x : mg->ix) if (x) CID_TRY(cid_index_set_hash_variant(x, variant));
    for (auto* x : mg->slice) if (x) CID_TRY(cid_index_set_hash_variant(x, variant));
    return CID_OK;
}

// rows of a .bxi (bigsi.rs:25): every replica gets all of them, a column shard its word columns of each row
int cid_mg_index_upload_rows(cid_mg* mg, const uint64_t* row_ids, const uint32_t* words, uint64_t nrows) {
    if (!mg || (nrows && (!row_ids || !words))) { set_error("cid_mg_index_upload_rows: null argument"); return CID_E_INVALID; }
    const uint32_t W = mg->W;
    if (mg->mode == CID_MG_REPLICATED) {
        return parallel(mg->n, [&](uint32_t g) -> int {
            if (!mg->ix[g]) CID_TRY(make_index(mg, g, mg->N, &mg->ix[g]));
            return cid_index_upload_rows(mg->ix[g], row_ids, words, nrows);
        });
    }
    CID_TRY(parallel(mg->ns, [&](uint32_t g) -> int {
        const uint32_t nw = mg->nw[g], w0 = mg->w0[g];
        std::vector<uint64_t> ids;
        std::vector<uint32_t> ws;
        for (uint64_t r = 0; r < nrows; r++) {
            const uint32_t* src = words + r * W + w0;
            uint32_t any = 0;
            for (uint32_t j = 0; j < nw; j++) any |= src[j];
            if (!any) continue;                          // nothing of this row in this shard's columns
            ids.push_back(row_ids[r]);
            ws.insert(ws.end(), src, src + nw);
        }
        return cid_index_upload_rows(mg->ix[g], ids.data(), ws.data(), ids.size());
    }));
    return or_rownz(mg, mg->ix, mg->ns);
}

// build.rs:33-130 phase 1 for one accession, on the GPU that owns its column.  Calls for accessions of different shards may
// come from different host threads.
int cid_mg_build_accession(cid_mg* mg, uint32_t colour, const char* bases, const uint64_t* seq_offs, uint64_t nseq, int seq_mode,
                           int64_t cutoff, int mini_variant, uint64_t* n_ref_kmers, int64_t* cutoff_used) {
    if (!mg) { set_error("cid_mg_build_accession: null"); return CID_E_INVALID; }
    const uint32_t g = owner_of(mg, colour);
    if (g >= mg->ns) { set_error("colour %u >= n_colors %u", colour, mg->N); return CID_E_INVALID; }
    std::lock_guard<std::mutex> lk(mg->mu[g]);
    std::vector<cid_index*>& part = mg->mode == CID_MG_COLUMNS ? mg->ix : mg->slice;
    if (!part[g]) CID_TRY(make_index(mg, g, mg->ncol[g], &part[g]));
    const uint32_t local = colour - mg->col0[g];
    if (mini_variant >= 0) return cid_build_accession_mini(part[g], local, bases, seq_offs, nseq, seq_mode, cutoff, mini_variant, n_ref_kmers, cutoff_used);
    return cid_build_accession(part[g], local, bases, seq_offs, nseq, seq_mode, cutoff, n_ref_kmers, cutoff_used);
}
int cid_mg_shard_of_colour(const cid_mg* mg, uint32_t colour) { return mg ? (int)owner_of(mg, colour) : -1; }

// build.rs:116-128 phase 2 on every shard, then the cross-shard step: OR of the row-present bitmaps (columns), or the
// column slices copied into every replica (replicated)
int cid_mg_build_finalize(cid_mg* mg) {
    if (!mg) { set_error("cid_mg_build_finalize: null"); return CID_E_INVALID; }
    std::vector<cid_index*>& part = mg->mode == CID_MG_COLUMNS ? mg->ix : mg->slice;
    CID_TRY(parallel(mg->ns, [&](uint32_t g) -> int {
        if (!part[g]) CID_TRY(make_index(mg, g, mg->ncol[g], &part[g]));      // a shard none of whose accessions was built
        return cid_build_finalize(part[g]);
    }));
    if (mg->mode == CID_MG_COLUMNS) return or_rownz(mg, mg->ix, mg->ns);
    CID_TRY(parallel(mg->n, [&](uint32_t r) -> int {
        if (!mg->ix[r]) CID_TRY(make_index(mg, r, mg->N, &mg->ix[r]));
        cid_index* dst = mg->ix[r];
        CID_CUDA(cudaSetDevice(dst->ctx->device));
        for (uint32_t g = 0; g < mg->ns; g++) {
            const cid_index* src = part[g];
            CID_CUDA(cudaMemcpy2DAsync(dst->rows + mg->w0[g], (size_t)dst->Wp * 4, src->rows, (size_t)src->Wp * 4, (size_t)mg->nw[g] * 4,
                                       mg->S, cudaMemcpyDefault, dst->ctx->stream));
        }
        CID_CUDA(cudaStreamSynchronize(dst->ctx->stream));
        return cid_index_refresh_rownz(dst);
    }));
    for (auto*& x : mg->slice) { if (x) cid_index_destroy(x); x = nullptr; }
    return CID_OK;
}

int cid_mg_index_count_nonzero_rows(cid_mg* mg, uint64_t* nrows) {
    if (!mg || !mg->ix[0]) { set_error("cid_mg_index_count_nonzero_rows: no index"); return CID_E_INVALID; }
    return cid_index_count_nonzero_rows(mg->ix[0], nrows);       // (the bitmap is global on every shard)
}
// non-zero rows in ascending order for save_bigsi (bigsi.rs:51-57): a column shard contributes its words of every stored row
int cid_mg_index_download_nonzero_rows(cid_mg* mg, uint64_t* row_ids, uint32_t* words, uint64_t cap, uint64_t* nrows) {
    if (!mg || !mg->ix[0] || !row_ids || !words || !nrows) { set_error("cid_mg_index_download_nonzero_rows: null argument"); return CID_E_INVALID; }
    if (mg->mode == CID_MG_REPLICATED) return cid_index_download_nonzero_rows(mg->ix[0], row_ids, words, cap, nrows);
    const uint32_t W = mg->W;
    std::vector<uint64_t> got(mg->ns, 0);
    CID_TRY(parallel(mg->ns, [&](uint32_t g) -> int {
        std::vector<uint64_t> ids(cap);
        std::vector<uint32_t> ws((size_t)cap * mg->nw[g]);
        CID_TRY(cid_index_download_nonzero_rows(mg->ix[g], ids.data(), ws.data(), cap, &got[g]));
        for (uint64_t r = 0; r < got[g]; r++) {
            if (g == 0) row_ids[r] = ids[r];
            memcpy(words + r * W + mg->w0[g], ws.data() + r * mg->nw[g], (size_t)mg->nw[g] * 4);
        }
        return (int)CID_OK;
    }));
    for (uint32_t g = 1; g < mg->ns; g++) if (got[g] != got[0]) { set_error("shards disagree on the stored rows"); return CID_E_INVALID; }
    *nrows = got[0];
    return CID_OK;
}

// ---- search ---------------------------------------------------------------------------------------------------------
// batch_search_pe.rs:9-179 (see cid_query_counts): replicated -> the queries are dealt to the GPUs; column-sharded -> every
// GPU gathers all k-mers from its slice
static int mg_query_counts_columns_uniq(cid_mg* mg, const char* bases, const uint64_t* seq_offs, uint64_t nseq, const uint64_t* query_offs,
                                        uint64_t nq, int seq_mode, int gene_search, int64_t filter, uint32_t* counts, uint64_t* num_kmers,
                                        uint64_t* uniq_n, uint64_t* uniq_sum, uint64_t* uniq_mode, int64_t* cutoff_used);

int cid_mg_query_counts(cid_mg* mg, const char* bases, const uint64_t* seq_offs, uint64_t nseq, const uint64_t* query_offs,
                        uint64_t nq, int seq_mode, int gene_search, int64_t filter, uint32_t* counts, uint64_t* num_kmers,
                        uint64_t* uniq_n, uint64_t* uniq_sum, uint64_t* uniq_mode, int64_t* cutoff_used) {
    if (!mg || !seq_offs || !query_offs || !counts || !num_kmers) { set_error("cid_mg_query_counts: null argument"); return CID_E_INVALID; }
    const uint32_t N = mg->N;
    if (nq == 0) return CID_OK;
    if (mg->mode == CID_MG_REPLICATED) {
        const std::vector<uint64_t> cuts = split_units(seq_offs, query_offs, nq, mg->n);
        return parallel(mg->n, [&](uint32_t g) -> int {
            const uint64_t q0 = cuts[g], q1 = cuts[g + 1];
            if (q1 == q0) return (int)CID_OK;
            const SubBatch sb = sub_batch(bases, nullptr, seq_offs, query_offs, q0, q1);
            return cid_query_counts(mg->ix[g], sb.bases, sb.seq_offs.data(), sb.nseq, sb.unit_offs.data(), q1 - q0, seq_mode, gene_search, filter,
                                    counts + q0 * N, num_kmers + q0, uniq_n ? uniq_n + q0 * N : nullptr, uniq_sum ? uniq_sum + q0 * N : nullptr,
                                    uniq_mode ? uniq_mode + q0 * N : nullptr, cutoff_used ? cutoff_used + q0 : nullptr);
        });
    }
    if (uniq_n || uniq_sum || uniq_mode)
        return mg_query_counts_columns_uniq(mg, bases, seq_offs, nseq, query_offs, nq, seq_mode, gene_search, filter, counts, num_kmers,
                                            uniq_n, uniq_sum, uniq_mode, cutoff_used);
    // counts only: every shard runs the whole search on its slice; the slices land side by side in the full-width result
    return parallel(mg->ns, [&](uint32_t g) -> int {
        const uint32_t nc = mg->ncol[g];
        std::vector<uint32_t> part((size_t)nq * nc);
        std::vector<uint64_t> nk(nq);
        std::vector<int64_t> cu(nq);
        CID_TRY(cid_query_counts(mg->ix[g], bases, seq_offs, nseq, query_offs, nq, seq_mode, gene_search, filter, part.data(), nk.data(),
                                 nullptr, nullptr, nullptr, cu.data()));
        for (uint64_t q = 0; q < nq; q++) memcpy(counts + q * N + mg->col0[g], part.data() + q * nc, (size_t)nc * 4);
        if (g == 0) {
            memcpy(num_kmers, nk.data(), nq * 8);
            if (cutoff_used) memcpy(cutoff_used, cu.data(), nq * 8);
        }
        return (int)CID_OK;
    });
}

// default report on a column-sharded index: the three-call protocol of include/colorid_b200.h around one exchange
static int mg_query_counts_columns_uniq(cid_mg* mg, const char* bases, const uint64_t* seq_offs, uint64_t nseq, const uint64_t* query_offs,
                                        uint64_t nq, int seq_mode, int gene_search, int64_t filter, uint32_t* counts, uint64_t* num_kmers,
                                        uint64_t* uniq_n, uint64_t* uniq_sum, uint64_t* uniq_mode, int64_t* cutoff_used) {
    (void)nseq;
    const uint32_t N = mg->N, ns = mg->ns;
    // batches of queries that fit one count-table pass (the survivors call takes one pass at a time)
    uint64_t q0 = 0;
    while (q0 < nq) {
        uint64_t q1 = q0, slots = 0;
        while (q1 < nq) {
            const uint64_t nb = seq_offs[query_offs[q1 + 1]] - seq_offs[query_offs[q1]];
            const uint64_t s = next_pow2(std::max<uint64_t>(64, 2 * nb));
            if (q1 > q0 && slots + s > (1ull << 27)) break;
            slots += s; q1++;
        }
        const uint64_t bq = q1 - q0;
        const SubBatch sb = sub_batch(bases, nullptr, seq_offs, query_offs, q0, q1);
        // 1. shard 0 counts, filters and compacts the survivors
        void* d_slots0 = nullptr;
        std::vector<uint64_t> surv(bq);
        std::vector<int64_t> cu(bq);
        CID_TRY(cid_query_survivors(mg->ix[0], sb.bases, sb.seq_offs.data(), sb.nseq, sb.unit_offs.data(), bq, seq_mode, gene_search, filter,
                                    &d_slots0, surv.data(), cu.data()));
        uint64_t total = 0;
        for (uint64_t q = 0; q < bq; q++) total += surv[q];
        if (cutoff_used) memcpy(cutoff_used + q0, cu.data(), bq * 8);
        // 2. the list goes to every shard; each gathers its slice and reports min(popcount, 2) + the single hit per k-mer
        std::vector<void*> d_slots(ns, nullptr), d_counts(ns, nullptr), d_nk(ns, nullptr), d_pc(ns, nullptr), d_col(ns, nullptr), d_sum(ns, nullptr);
        std::vector<std::vector<uint8_t>> pc(ns);
        auto free_all = [&]() {
            for (uint32_t g = 0; g < ns; g++)
                for (void* p : {g ? d_slots[g] : nullptr, d_counts[g], d_nk[g], d_pc[g], d_col[g], d_sum[g]}) if (p) cid_dev_free(mg->ctx[g], p);
        };
        int rc = parallel(ns, [&](uint32_t g) -> int {
            cid_ctx* c = mg->ctx[g];
            const uint32_t nc = mg->ncol[g];
            CID_CUDA(cudaSetDevice(c->device));
            if (g == 0) d_slots[0] = d_slots0;
            else if (total) {
                CID_TRY(cid_dev_alloc(c, total * 16, &d_slots[g]));
                CID_CUDA(cudaMemcpy(d_slots[g], d_slots0, total * 16, cudaMemcpyDefault));        // peer copy of 16 B per k-mer
            }
            CID_TRY(cid_dev_alloc(c, (size_t)bq * nc * 4 + 16, &d_counts[g]));
            CID_TRY(cid_dev_alloc(c, bq * 8 + 16, &d_nk[g]));
            CID_TRY(cid_dev_alloc(c, total + 16, &d_pc[g]));
            CID_TRY(cid_dev_alloc(c, total * 4 + 16, &d_col[g]));
            CID_TRY(cid_dev_alloc(c, total + 16, &d_sum[g]));
            CID_TRY(cid_query_slots_counts_dev(mg->ix[g], d_slots[g], surv.data(), bq, (uint32_t*)d_counts[g], (uint64_t*)d_nk[g],
                                               (uint8_t*)d_pc[g], (uint32_t*)d_col[g], c->stream));
            std::vector<uint32_t> part((size_t)bq * nc);
            CID_CUDA(cudaMemcpyAsync(part.data(), d_counts[g], (size_t)bq * nc * 4, cudaMemcpyDeviceToHost, c->stream));
            pc[g].resize(total);
            if (total) CID_CUDA(cudaMemcpyAsync(pc[g].data(), d_pc[g], total, cudaMemcpyDeviceToHost, c->stream));
            if (g == 0) CID_CUDA(cudaMemcpyAsync(num_kmers + q0, d_nk[0], bq * 8, cudaMemcpyDeviceToHost, c->stream));
            CID_CUDA(cudaStreamSynchronize(c->stream));
            for (uint64_t q = 0; q < bq; q++) memcpy(counts + (q0 + q) * N + mg->col0[g], part.data() + q * nc, (size_t)nc * 4);
            return (int)CID_OK;
        });
        if (rc != CID_OK) { free_all(); return rc; }
        // 3. the exchange: one byte per k-mer summed over the shards; then the unique hits of each shard's accessions
        std::vector<uint8_t> sum(total, 0);
        for (uint32_t g = 0; g < ns; g++) for (uint64_t i = 0; i < total; i++) sum[i] = (uint8_t)std::min<uint32_t>(255, sum[i] + pc[g][i]);
        rc = parallel(ns, [&](uint32_t g) -> int {
            cid_ctx* c = mg->ctx[g];
            const uint32_t nc = mg->ncol[g];
            CID_CUDA(cudaSetDevice(c->device));
            if (total) CID_CUDA(cudaMemcpyAsync(d_sum[g], sum.data(), total, cudaMemcpyHostToDevice, c->stream));
            std::vector<uint64_t> un((size_t)bq * nc), us((size_t)bq * nc), um((size_t)bq * nc);
            CID_TRY(cid_query_slots_uniq_dev(mg->ix[g], d_slots[g], surv.data(), bq, (const uint8_t*)d_pc[g], (const uint8_t*)d_sum[g],
                                             (const uint32_t*)d_col[g], un.data(), us.data(), um.data(), c->stream));
            for (uint64_t q = 0; q < bq; q++) {
                const size_t at = (q0 + q) * N + mg->col0[g];
                if (uniq_n) memcpy(uniq_n + at, un.data() + q * nc, (size_t)nc * 8);
                if (uniq_sum) memcpy(uniq_sum + at, us.data() + q * nc, (size_t)nc * 8);
                if (uniq_mode) memcpy(uniq_mode + at, um.data() + q * nc, (size_t)nc * 8);
            }
            return (int)CID_OK;
        });
        free_all();
        if (rc != CID_OK) return rc;
        q0 = q1;
    }
    return CID_OK;
}

// perfect_search.rs:6-120
static int mg_query_perfect(cid_mg* mg, const char* bases, const uint64_t* seq_offs, uint64_t nseq, const uint64_t* query_offs, uint64_t nq,
                            bool mf, uint32_t* and_rows, uint8_t* status, uint64_t* n_kmers) {
    if (!mg || !seq_offs || !and_rows || !status || !n_kmers) { set_error("cid_mg_query_perfect: null argument"); return CID_E_INVALID; }
    if (nq == 0) return CID_OK;
    const uint32_t W = mg->W;
    std::vector<uint64_t> one_per_seq;
    if (mf) { one_per_seq.resize(nseq + 1); for (uint64_t i = 0; i <= nseq; i++) one_per_seq[i] = i; query_offs = one_per_seq.data(); }
    auto call = [&](cid_index* ix, const char* b, const uint64_t* so, uint64_t ns_, const uint64_t* qo, uint64_t n, uint32_t* rows, uint8_t* st, uint64_t* nk) {
        return mf ? cid_query_perfect_mf(ix, b, so, ns_, rows, st, nk) : cid_query_perfect(ix, b, so, ns_, qo, n, rows, st, nk);
    };
    if (mg->mode == CID_MG_REPLICATED) {
        const std::vector<uint64_t> cuts = split_units(seq_offs, query_offs, nq, mg->n);
        return parallel(mg->n, [&](uint32_t g) -> int {
            const uint64_t q0 = cuts[g], q1 = cuts[g + 1];
            if (q1 == q0) return (int)CID_OK;
            const SubBatch sb = sub_batch(bases, nullptr, seq_offs, query_offs, q0, q1);
            return call(mg->ix[g], sb.bases, sb.seq_offs.data(), sb.nseq, sb.unit_offs.data(), q1 - q0, and_rows + q0 * W, status + q0, n_kmers + q0);
        });
    }
    return parallel(mg->ns, [&](uint32_t g) -> int {
        const uint32_t nw = mg->nw[g];
        std::vector<uint32_t> rows((size_t)nq * nw);
        std::vector<uint8_t> st(nq);
        std::vector<uint64_t> nk(nq);
        CID_TRY(call(mg->ix[g], bases, seq_offs, nseq, query_offs, nq, rows.data(), st.data(), nk.data()));
        for (uint64_t q = 0; q < nq; q++) memcpy(and_rows + q * W + mg->w0[g], rows.data() + q * nw, (size_t)nw * 4);
        if (g == 0) { memcpy(status, st.data(), nq); memcpy(n_kmers, nk.data(), nq * 8); }      // (absent rows: the bitmap is global)
        return (int)CID_OK;
    });
}
int cid_mg_query_perfect(cid_mg* mg, const char* bases, const uint64_t* seq_offs, uint64_t nseq, const uint64_t* query_offs, uint64_t nq,
                         uint32_t* and_rows, uint8_t* status, uint64_t* n_kmers) {
    if (!query_offs) { set_error("cid_mg_query_perfect: null argument"); return CID_E_INVALID; }
    return mg_query_perfect(mg, bases, seq_offs, nseq, query_offs, nq, false, and_rows, status, n_kmers);
}
int cid_mg_query_perfect_mf(cid_mg* mg, const char* bases, const uint64_t* seq_offs, uint64_t nseq, uint32_t* and_rows, uint8_t* status,
                            uint64_t* n_kmers) {
    return mg_query_perfect(mg, bases, seq_offs, nseq, nullptr, nseq, true, and_rows, status, n_kmers);
}

// ---- read_id: read_id_mt_pe.rs:282-363 parallel_vec (see cid_read_id_classify) ----------------------------------------
int cid_mg_read_id_classify(cid_mg* mg, const char* bases, const char* quals, const uint64_t* seq_offs, uint64_t nseq,
                            const uint64_t* read_offs, uint64_t nreads, const cid_readid_params* p, const uint64_t* n_ref_by_colour,
                            double fp_correct, int32_t* kind, uint32_t* hits, uint32_t* n_set, uint32_t* n_top, uint32_t* top, uint32_t top_cap) {
    if (!mg || !seq_offs || !read_offs || !n_ref_by_colour || !kind || !hits || !n_set || !n_top) { set_error("cid_mg_read_id_classify: null argument"); return CID_E_INVALID; }
    (void)nseq;
    if (nreads == 0) return CID_OK;
    if (mg->mode == CID_MG_REPLICATED) {
        // reads are independent units: contiguous ranges to the replicas, results land in place
        const std::vector<uint64_t> cuts = split_units(seq_offs, read_offs, nreads, mg->n);
        return parallel(mg->n, [&](uint32_t g) -> int {
            const uint64_t r0 = cuts[g], r1 = cuts[g + 1];
            if (r1 == r0) return (int)CID_OK;
            const SubBatch sb = sub_batch(bases, quals, seq_offs, read_offs, r0, r1);
            return cid_read_id_classify(mg->ix[g], sb.bases, sb.quals, sb.seq_offs.data(), sb.nseq, sb.unit_offs.data(), r1 - r0, p, n_ref_by_colour,
                                        fp_correct, kind + r0, hits + r0, n_set + r0, n_top + r0, top ? top + r0 * (size_t)top_cap : nullptr, top_cap);
        });
    }
    // column shards: every shard votes all reads on its accessions, report colours tagged with their insertion step; the
    // reports are merged into the unsharded index's report (host, a few entries per read) and classified
    cid_readid_params pp;
    default_readid_params(pp, p, mg->N);
    const uint32_t ns = mg->ns;
    std::vector<uint32_t> shard_n(mg->ncol.begin(), mg->ncol.begin() + ns), shard_off(mg->col0.begin(), mg->col0.begin() + ns);
    const uint64_t chunk = 1u << 17;
    for (uint64_t r0 = 0; r0 < nreads; r0 += chunk) {
        const uint64_t r1 = std::min(nreads, r0 + chunk), nr = r1 - r0;
        const SubBatch sb = sub_batch(bases, quals, seq_offs, read_offs, r0, r1);
        std::vector<std::vector<uint32_t>> s_nset(ns), s_flags(ns), s_repn(ns), s_rc(ns), s_rv(ns);
        uint32_t cap = 0;
        for (uint32_t g = 0; g < ns; g++) cap = std::max(cap, std::min(mg->ncol[g] + 1, 48u));
        for (;;) {
            std::atomic<int> truncated{0};
            CID_TRY(parallel(ns, [&](uint32_t g) -> int {
                cid_readid_params pg = pp;
                pg.rep_cap = cap;
                s_nset[g].assign(nr, 0); s_flags[g].assign(nr, 0); s_repn[g].assign(nr, 0);
                s_rc[g].assign((size_t)nr * cap, 0); s_rv[g].assign((size_t)nr * cap, 0);
                CID_TRY(cid_ctx_set_option(mg->ctx[g], "readid_report_steps", 1));
                const int rc = cid_read_id_batch(mg->ix[g], sb.bases, sb.quals, sb.seq_offs.data(), sb.nseq, sb.unit_offs.data(), nr, &pg,
                                                 s_nset[g].data(), s_flags[g].data(), s_repn[g].data(), s_rc[g].data(), s_rv[g].data());
                cid_ctx_set_option(mg->ctx[g], "readid_report_steps", 0);
                if (rc != CID_OK) return rc;
                for (uint64_t r = 0; r < nr; r++) if (s_flags[g][r] & 4u) { truncated = 1; break; }
                return (int)CID_OK;
            }));
            uint32_t full = 0;
            for (uint32_t g = 0; g < ns; g++) full = std::max(full, mg->ncol[g] + 1);
            if (!truncated.load() || cap >= full) break;
            cap = std::min(full, cap * 4);           // a read with more candidate colours than slots: again with room for them
        }
        const uint32_t cap_out = std::min<uint64_t>((uint64_t)mg->N + 1, (uint64_t)cap * ns);
        std::vector<uint32_t> m_repn(nr), m_rc((size_t)nr * cap_out), m_rv((size_t)nr * cap_out), m_flags(nr);
        for (uint64_t r = 0; r < nr; r++) m_flags[r] = s_flags[0][r] & ~4u;
        std::vector<const uint32_t*> p_repn(ns), p_rc(ns), p_rv(ns);
        for (uint32_t g = 0; g < ns; g++) { p_repn[g] = s_repn[g].data(); p_rc[g] = s_rc[g].data(); p_rv[g] = s_rv[g].data(); }
        CID_TRY(cid_merge_shard_reports(ns, shard_n.data(), shard_off.data(), nr, p_repn.data(), p_rc.data(), p_rv.data(), cap, mg->N,
                                        m_repn.data(), m_rc.data(), m_rv.data(), cap_out, m_flags.data()));
        memcpy(n_set + r0, s_nset[0].data(), nr * 4);
        CID_TRY(cid_classify_reads(mg->S, mg->H, mg->N, n_ref_by_colour, fp_correct, pp.group_width, nr, s_nset[0].data(), m_flags.data(),
                                   m_repn.data(), m_rc.data(), m_rv.data(), cap_out, mg->ctx[0]->opt_host_threads, kind + r0, hits + r0,
                                   n_top + r0, top ? top + r0 * (size_t)top_cap : nullptr, top ? top_cap : 0));
    }
    return CID_OK;
}

}  // extern "C"
