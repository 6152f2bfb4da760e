// C ABI of libcolorid_b200 (see include/colorid_b200.h): context, index lifecycle and the
// host-pointer / device-pointer entry points that drive the kernels.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <chrono>
#include <map>
#include <vector>

#include "cid_device.cuh"
#include "cid_internal.h"

namespace cid {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int DevBuf::ensure(size_t bytes) {
    if (bytes <= cap && p) return CID_OK;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = std::max<size_t>(bytes, 256);
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { p = nullptr; set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); cudaGetLastError(); return CID_E_NOMEM; }
    cap = want;
    return CID_OK;
}
void DevBuf::release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
int PinBuf::ensure(size_t bytes) {
    if (bytes <= cap && p) return CID_OK;
    if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
    size_t want = std::max<size_t>(bytes, 256);
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) { p = nullptr; set_error("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e)); cudaGetLastError(); return CID_E_NOMEM; }
    cap = want;
    return CID_OK;
}
void PinBuf::release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }

static const char* kKernelNames[KID_COUNT] = {"kmerize_insert", "region_histogram", "region_to_bloom", "transpose_bitsets",
                                              "rownz", "query_counts", "query_uniq_wide", "query_perfect",
                                              "readid_kmerize", "readid_sched", "readid_order", "readid_vote", "readid_classify", "table_clear", "other", "query_hash", "query_front", "readid_big", "readid_vp_scan", "readid_vp_gather", "readid_vp_count"};
static cudaEvent_t prof_event(cid_ctx* c) {
    if (!c->prof_pool.empty()) { cudaEvent_t e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
ProfScope::ProfScope(cid_ctx* c, cudaStream_t s, int kernel) : ctx(c), st(s), idx(-1) {
    if (!c->prof_on) return;
    cid_ctx::ProfRec r; r.kernel = kernel; r.a = prof_event(c); r.b = prof_event(c);
    cudaEventRecord(r.a, s);
    c->prof_recs.push_back(r);
    idx = (int)c->prof_recs.size() - 1;
}
ProfScope::~ProfScope() { if (idx >= 0) cudaEventRecord(ctx->prof_recs[idx].b, st); }
static void prof_collect(cid_ctx* c) {
    for (auto& r : c->prof_recs) {
        cudaEventSynchronize(r.b);
        float ms = 0;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { c->prof_ms[r.kernel] += ms; c->prof_n[r.kernel]++; }
        c->prof_pool.push_back(r.a); c->prof_pool.push_back(r.b);
    }
    c->prof_recs.clear();
}

// lower_raw != nullptr: a raw-case input with lower-case k-mers is not an error for this caller -- it is told so and redoes
// the work with case-aware count-table slots (ctx->case_aware)
int check_err_flags(cid_ctx* ctx, cudaStream_t st, bool* lower_raw) {
    CID_CUDA(cudaMemcpyAsync(ctx->h_err, ctx->d_err, 4, cudaMemcpyDeviceToHost, st));
    CID_CUDA(cudaStreamSynchronize(st));
    uint32_t f = ctx->h_err[0];
    if (!f) return CID_OK;
    CID_CUDA(cudaMemsetAsync(ctx->d_err, 0, 8, st));
    if ((f & ERRF_LOWER_RAW) && lower_raw) { *lower_raw = true; f &= ~(uint32_t)ERRF_LOWER_RAW; if (!f) return CID_OK; }
    if (f & ERRF_LOWER_RAW) {
        set_error("lower-case bases inside k-mers of a raw-case (FASTQ) input are not supported for minimizer indexes");
        return CID_E_UNSUPPORTED;
    }
    if (f & ERRF_STRING_NONACGT) {
        set_error("-s -m query (kmerize_string) with a byte outside ACGTacgt inside a k-mer is not supported on the device");
        return CID_E_UNSUPPORTED;
    }
    if (f & ERRF_TABLE_FULL) { set_error("k-mer count table overflow"); return CID_E_CAPACITY; }
    if (f & ERRF_READ_TOO_LONG) { set_error("a read exceeds the declared maximum read length"); return CID_E_CAPACITY; }
    set_error("internal capacity exceeded (flags 0x%x)", f);
    return CID_E_CAPACITY;
}

// kmer.rs:866-942 auto_cutoff on a dense histogram h[c] = number of distinct k-mers seen c times.
// Returns CID_E_REF_PANIC where the Rust code would index out of range / underflow.
static int auto_cutoff_dense(const std::vector<uint64_t>& h, uint64_t max_cov, int64_t* out) {
    uint64_t distinct = 0, weighted = 0;
    for (uint64_t c = 1; c <= max_cov; c++) { distinct += h[c]; weighted += c * h[c]; }
    const double total_mean = (double)weighted / (double)distinct;
    if (total_mean < 1.5) { *out = 0; return CID_OK; }
    // coverages[c-1] = h[c] for c in 1..max_cov (the maximum itself is excluded, kmer.rs:887)
    const size_t ncov = max_cov >= 1 ? (size_t)(max_cov - 1) : 0;
    if (ncov < 3) { set_error("auto_cutoff: degenerate k-mer histogram (reference panics)"); return CID_E_REF_PANIC; }
    auto cov = [&](size_t i) { return (double)h[i + 1]; };
    size_t nd1 = ncov - 2;                       // d1[j] = cov[j+1]/cov[j+2]
    size_t pos1 = 0, pos2 = 0;
    for (size_t j = 0; j < nd1; j++) if (cov(j + 1) / cov(j + 2) < 1.0) { pos1 = j + 1; break; }
    for (size_t j = 0; j + 1 < nd1; j++) {
        double a = cov(j + 1) / cov(j + 2), b = cov(j + 2) / cov(j + 3);
        if (a / b < 1.0) { pos2 = j + 1; break; }
    }
    uint64_t bigsum = 0, num = 0;
    for (size_t i = 0; i + 1 < ncov; i++) { bigsum += i * h[i + 2]; num += h[i + 2]; }
    const double mean = (double)bigsum / (double)num;
    if (pos1 > 0 && (double)pos1 < mean * 0.75) *out = (int64_t)pos1;
    else if (pos2 > 0) *out = (int64_t)pos2;
    else {
        double half = std::ceil(mean / 2.0);
        uint64_t hv = std::isnan(half) ? 0 : (uint64_t)half;
        *out = (int64_t)std::max<uint64_t>(1, hv);
    }
    return CID_OK;
}

__global__ void table_clear_kernel(Slot* t, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        Slot s; s.key = CID_EMPTY_KEY; s.count = 0; s.pad = CID_CS_UNSET;
        t[i] = s;
    }
}
static int table_clear(cid_ctx* ctx, cudaStream_t st, void* d_table, uint64_t nslots) {
    unsigned grid = (unsigned)std::min<uint64_t>((nslots + 255) / 256, (uint64_t)ctx->sm_count * 32);
    if (grid == 0) grid = 1;
    ProfScope ps(ctx, st, KID_TABLE_CLEAR);
    table_clear_kernel<<<grid, 256, 0, st>>>((Slot*)d_table, nslots);
    ctx->launches++;
    CID_CUDA(cudaGetLastError());
    return CID_OK;
}

enum { HIST_BINS = 65536, HIST_OVERFLOW_CAP = 1 << 20 };

// Histogram of one region -> auto_cutoff (host arithmetic on a few KB).
static int region_auto_cutoff(cid_ctx* ctx, cudaStream_t st, const void* d_region, uint64_t nslots, int64_t* cutoff,
                              bool packed = false, uint64_t* survivors = nullptr) {
    CID_TRY(ctx->scratch[8].ensure((size_t)HIST_BINS * 4 + 16));
    CID_TRY(ctx->scratch[9].ensure((size_t)HIST_OVERFLOW_CAP * 4));
    uint32_t* d_hist = ctx->scratch[8].as<uint32_t>();
    uint32_t* d_ovn = d_hist + HIST_BINS;
    CID_CUDA(cudaMemsetAsync(d_hist, 0, (size_t)HIST_BINS * 4 + 16, st));
    CID_TRY(launch_region_histogram(ctx, st, d_region, nslots, d_hist, HIST_BINS, ctx->scratch[9].as<uint32_t>(),
                                    HIST_OVERFLOW_CAP, d_ovn, packed));
    std::vector<uint32_t> hh(HIST_BINS + 4);
    CID_CUDA(cudaMemcpyAsync(hh.data(), d_hist, (size_t)HIST_BINS * 4 + 16, cudaMemcpyDeviceToHost, st));
    CID_CUDA(cudaStreamSynchronize(st));
    uint32_t novf = hh[HIST_BINS];
    if (novf > HIST_OVERFLOW_CAP) { set_error("auto_cutoff: too many k-mers with count >= %d", HIST_BINS); return CID_E_CAPACITY; }
    std::vector<uint32_t> ovf(novf);
    if (novf) {
        CID_CUDA(cudaMemcpy(ovf.data(), ctx->scratch[9].p, (size_t)novf * 4, cudaMemcpyDeviceToHost));
    }
    uint64_t max_cov = 0;
    for (uint32_t c = 1; c < HIST_BINS; c++) if (hh[c]) max_cov = c;
    for (uint32_t v : ovf) max_cov = std::max<uint64_t>(max_cov, v);
    if (max_cov > (1u << 28)) { set_error("auto_cutoff: k-mer multiplicity above 2^28"); return CID_E_CAPACITY; }
    std::vector<uint64_t> h(max_cov + 2, 0);
    for (uint64_t c = 1; c < std::min<uint64_t>(HIST_BINS, max_cov + 1); c++) h[c] = hh[c];
    for (uint32_t v : ovf) h[v] += 1;
    CID_TRY(auto_cutoff_dense(h, max_cov, cutoff));
    if (survivors) {          // clean_map: k-mers with count > cutoff
        *survivors = 0;
        for (uint64_t c = (uint64_t)std::max<int64_t>(*cutoff, 0) + 1; c <= max_cov; c++) *survivors += h[c];
    }
    return CID_OK;
}

static int ensure_bitsets(cid_index* idx) {
    if (idx->bitsets) return CID_OK;
    idx->bs_words = ((idx->S + 31) / 32 + 31) / 32 * 32;
    size_t bytes = (size_t)idx->N * idx->bs_words * 4;
    cudaError_t e = cudaMalloc((void**)&idx->bitsets, std::max<size_t>(bytes, 256));
    if (e != cudaSuccess) { idx->bitsets = nullptr; set_error("cudaMalloc(bitsets %zu B) failed: %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); return CID_E_NOMEM; }
    CID_CUDA(cudaMemsetAsync(idx->bitsets, 0, bytes, idx->ctx->stream));
    return CID_OK;
}

// One group worth of count table on scratch[0]; off/mask on scratch[1].
static int single_region(cid_ctx* ctx, cudaStream_t st, uint64_t nbases, uint32_t k, uint64_t* nslots_out,
                         uint64_t** d_off, uint64_t** d_mask, uint64_t slots_hint = 0, bool packed = false) {
    uint64_t npos = nbases >= k ? nbases - k + 1 : 0;
    uint64_t slots = slots_hint ? slots_hint : next_pow2(std::max<uint64_t>(64, 2 * npos));
    CID_TRY(ctx->scratch[0].ensure(slots * (packed ? 8 : sizeof(Slot))));
    CID_TRY(ctx->scratch[1].ensure(64));
    uint64_t hv[2] = {0, slots - 1};
    CID_CUDA(cudaMemcpyAsync(ctx->scratch[1].p, hv, 16, cudaMemcpyHostToDevice, st));
    if (packed) {
        ProfScope ps(ctx, st, KID_TABLE_CLEAR);
        CID_CUDA(cudaMemsetAsync(ctx->scratch[0].p, 0xFF, slots * 8, st));
    } else CID_TRY(table_clear(ctx, st, ctx->scratch[0].p, slots));
    *nslots_out = slots;
    *d_off = ctx->scratch[1].as<uint64_t>();
    *d_mask = ctx->scratch[1].as<uint64_t>() + 1;
    return CID_OK;
}

}  // namespace cid

using namespace cid;

extern "C" {

int cid_version(void) { return 100; }
const char* cid_last_error(void) { return g_err; }

int cid_host_alloc(size_t bytes, void** out) {
    if (!out) { set_error("cid_host_alloc: null out"); return CID_E_INVALID; }
    cudaError_t e = cudaHostAlloc(out, std::max<size_t>(bytes, 64), cudaHostAllocPortable);
    if (e != cudaSuccess) { *out = nullptr; set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); return CID_E_NOMEM; }
    return CID_OK;
}
void cid_host_free(void* p) { if (p) cudaFreeHost(p); }

int cid_ctx_create(int device, cid_ctx** out) {
    if (!out) { set_error("cid_ctx_create: null out"); return CID_E_INVALID; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no usable CUDA device (%s); colorid_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
        cudaGetLastError();
        return CID_E_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return CID_E_INVALID; }
    CID_CUDA(cudaSetDevice(device));
    cid_ctx* c = new cid_ctx();
    c->device = device;
    cudaDeviceProp prop;
    CID_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    // Row gathers are 8..128-byte random reads: ask L2 to fetch 32-byte sectors instead of 64/128 B.
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    cudaGetLastError();
    CID_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CID_CUDA(cudaMalloc((void**)&c->d_err, 16));
    CID_CUDA(cudaMemset(c->d_err, 0, 16));
    CID_CUDA(cudaMallocHost((void**)&c->h_err, 16));
    *out = c;
    return CID_OK;
}
void cid_ctx_destroy(cid_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    readid_pipe_destroy(c);
    prof_collect(c);
    for (auto e : c->prof_pool) cudaEventDestroy(e);
    for (auto& b : c->scratch) b.release();
    for (auto& b : c->pinned) b.release();
    if (c->d_err) cudaFree(c->d_err);
    if (c->h_err) cudaFreeHost(c->h_err);
    for (int i = 0; i < 2; i++) { if (c->aux[i]) cudaStreamDestroy(c->aux[i]); if (c->aux_join[i]) cudaEventDestroy(c->aux_join[i]); }
    if (c->aux_fork) cudaEventDestroy(c->aux_fork);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}
int cid_ctx_device(const cid_ctx* c) { return c ? c->device : -1; }
int cid_ctx_profile(cid_ctx* c, int enable) {
    prof_collect(c);
    c->prof_on = enable != 0;
    for (int i = 0; i < KID_COUNT; i++) { c->prof_ms[i] = 0; c->prof_n[i] = 0; }
    return CID_OK;
}
int cid_ctx_profile_read(cid_ctx* c, int kernel, const char** name, double* total_ms, uint64_t* launches) {
    if (kernel < 0 || kernel >= KID_COUNT) return CID_E_INVALID;
    prof_collect(c);
    if (name) *name = kKernelNames[kernel];
    if (total_ms) *total_ms = c->prof_ms[kernel];
    if (launches) *launches = c->prof_n[kernel];
    return CID_OK;
}
uint64_t cid_ctx_launch_count(const cid_ctx* c) { return c ? c->launches : 0; }
int cid_ctx_read_counter(cid_ctx* c, const char* name, uint64_t* value) {
    if (!c || !name || !value) { set_error("cid_ctx_read_counter: null argument"); return CID_E_INVALID; }
    if (strcmp(name, "readid_gather_rows")) { set_error("cid_ctx_read_counter: unknown counter '%s'", name); return CID_E_INVALID; }
    CID_CUDA(cudaSetDevice(c->device));
    CID_CUDA(cudaDeviceSynchronize());
    unsigned long long v = 0;
    CID_CUDA(cudaMemcpy(&v, c->d_err + 2, 8, cudaMemcpyDeviceToHost));
    CID_CUDA(cudaMemset(c->d_err + 2, 0, 8));
    *value = v;
    return CID_OK;
}
int cid_ctx_set_option(cid_ctx* c, const char* name, int64_t value) {
    if (!c || !name) { set_error("cid_ctx_set_option: null argument"); return CID_E_INVALID; }
    if (!strcmp(name, "readid_chunk0_reads")) { c->opt_readid_chunk0 = value > 0 ? (uint64_t)value : 0; return CID_OK; }
    if (!strcmp(name, "host_ranks")) { c->opt_host_ranks = value > 1 ? (int)value : 1; return CID_OK; }
    if (!strcmp(name, "readid_chunk_growth_pct")) { c->opt_readid_chunk_growth = value > 100 ? (int)value : 0; return CID_OK; }
    if (!strcmp(name, "readid_chunk_reads")) { c->opt_readid_chunk = value > 0 ? (uint64_t)value : 0; return CID_OK; }
    if (!strcmp(name, "readid_streams")) { c->opt_readid_streams = value >= 2 ? 2 : 1; return CID_OK; }
    if (!strcmp(name, "readid_kmerize_ctas")) { c->opt_kmerize_ctas = value > 0 ? (int)value : 0; return CID_OK; }
    if (!strcmp(name, "readid_vote_ctas")) { c->opt_vote_ctas = value > 0 ? (int)value : 0; return CID_OK; }
    if (!strcmp(name, "readid_report_steps")) { c->opt_readid_report_steps = value != 0; return CID_OK; }
    if (!strcmp(name, "readid_vote_part")) { c->opt_readid_vote_part = value < 0 ? 0 : value > 2 ? 2 : (int)value; return CID_OK; }
    if (!strcmp(name, "readid_part_ctas")) { c->opt_readid_part_ctas = value > 0 ? (int)value : 0; return CID_OK; }
    if (!strcmp(name, "readid_part_cap")) { c->opt_readid_part_cap = value > 0 ? (int)((value + 31) & ~31ll) : 0; return CID_OK; }
    if (!strcmp(name, "readid_part_shift")) { c->opt_readid_part_shift = value > 0 && value < 32 ? (int)value : 0; return CID_OK; }
    if (!strcmp(name, "readid_serialize")) { c->opt_readid_serialize = value != 0; return CID_OK; }
    if (!strcmp(name, "build_table_div")) { c->opt_build_table_div = value >= 1 ? (int)value : 0; return CID_OK; }
    if (!strcmp(name, "build_packed")) { c->opt_build_packed = value != 0; return CID_OK; }
    if (!strcmp(name, "build_set")) { c->opt_build_set = value != 0; return CID_OK; }
    if (!strcmp(name, "query_front")) { c->opt_query_front = value != 0; return CID_OK; }
    if (!strcmp(name, "gather_l2_64b")) { c->opt_gather_l2_64b = value != 0; return CID_OK; }
    if (!strcmp(name, "query_table_div")) { c->opt_query_table_div = value > 0 ? (int)value : 0; return CID_OK; }
    if (!strcmp(name, "query_table_min_slots")) { c->opt_query_table_min = value > 0 ? (uint64_t)value : 0; return CID_OK; }
    if (!strcmp(name, "query_compact")) { c->opt_query_compact = value < 0 ? 0 : value > 2 ? 2 : (int)value; return CID_OK; }
    if (!strcmp(name, "uniq_device")) { c->opt_uniq_device = value != 0; return CID_OK; }
    if (!strcmp(name, "query_fused")) { c->opt_query_fused = value != 0; return CID_OK; }
    if (!strcmp(name, "host_threads")) { c->opt_host_threads = value > 0 ? (int)value : 0; return CID_OK; }
    set_error("cid_ctx_set_option: unknown option '%s'", name);
    return CID_E_INVALID;
}

// ------------------------------------------------------------------ index
int cid_index_create(cid_ctx* ctx, uint64_t bloom_size, uint32_t num_hash, uint32_t k_size, uint32_t n_colors,
                     cid_index** out) {
    if (!ctx || !out) { set_error("cid_index_create: null argument"); return CID_E_INVALID; }
    if (k_size < 1 || k_size > 31) { set_error("k-mer size %u not supported on the device (1..31)", k_size); return CID_E_UNSUPPORTED; }
    if (num_hash < 1 || num_hash > MAX_HASH) { set_error("num_hash %u not supported (1..%d)", num_hash, MAX_HASH); return CID_E_UNSUPPORTED; }
    if (bloom_size < 1 || bloom_size >= (1ull << 32)) { set_error("bloom_size %llu not supported (1..2^32-1)", (unsigned long long)bloom_size); return CID_E_UNSUPPORTED; }
    if (n_colors < 1) { set_error("n_colors must be >= 1"); return CID_E_INVALID; }
    CID_CUDA(cudaSetDevice(ctx->device));
    cid_index* ix = new cid_index();
    ix->ctx = ctx; ix->S = bloom_size; ix->H = num_hash; ix->k = k_size; ix->N = n_colors;
    ix->W = (n_colors + 31) / 32;
    ix->Wp = padded_row_words(ix->W);
    size_t bytes = (size_t)ix->S * ix->Wp * 4;
    cudaError_t e = cudaMalloc((void**)&ix->rows, std::max<size_t>(bytes, 256));
    if (e != cudaSuccess) { set_error("cudaMalloc(matrix %zu B) failed: %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); delete ix; return CID_E_NOMEM; }
    ix->rownz_words = (ix->S + 31) / 32;
    e = cudaMalloc((void**)&ix->rownz, ix->rownz_words * 4 + 256);
    if (e != cudaSuccess) { set_error("cudaMalloc(rownz) failed"); cudaGetLastError(); cudaFree(ix->rows); delete ix; return CID_E_NOMEM; }
    {
        const HashCfg h = make_hashcfg(0);
        CID_CUDA(cudaMalloc(&ix->d_hcfg, 256));
        CID_CUDA(cudaMemcpy(ix->d_hcfg, &h, sizeof(h), cudaMemcpyHostToDevice));
    }
    CID_CUDA(cudaMemsetAsync(ix->rows, 0, bytes, ctx->stream));
    CID_CUDA(cudaMemsetAsync(ix->rownz, 0, ix->rownz_words * 4, ctx->stream));
    CID_CUDA(cudaStreamSynchronize(ctx->stream));
    ix->rownz_valid = true;
    *out = ix;
    return CID_OK;
}
void cid_index_destroy(cid_index* ix) {
    if (!ix) return;
    cudaSetDevice(ix->ctx->device);
    if (ix->rows) cudaFree(ix->rows);
    if (ix->rownz) cudaFree(ix->rownz);
    if (ix->bitsets) cudaFree(ix->bitsets);
    if (ix->d_hcfg) cudaFree(ix->d_hcfg);
    delete ix;
}
uint32_t cid_index_row_words(const cid_index* ix) { return ix ? ix->W : 0; }
uint32_t cid_index_row_stride(const cid_index* ix) { return ix ? ix->Wp : 0; }

int cid_index_refresh_rownz(cid_index* ix) {
    cid_ctx* ctx = ix->ctx;
    CID_CUDA(cudaSetDevice(ctx->device));
    CID_TRY(launch_rownz(ctx, ctx->stream, ix));
    CID_CUDA(cudaStreamSynchronize(ctx->stream));
    ix->rownz_valid = true;
    return CID_OK;
}
int cid_index_set_rownz_global(cid_index* ix, int is_global) { ix->rownz_global = is_global != 0; return CID_OK; }

// rows[id] <- words of the i-th uploaded row (any order of ids; the reference writes its map in hash order)
__global__ void scatter_rows_kernel(const uint64_t* __restrict__ ids, const uint32_t* __restrict__ words, uint64_t n, uint32_t W,
                                    uint32_t Wp, uint32_t* __restrict__ rows) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * W) return;
    const uint64_t i = t / W;
    const uint32_t w = (uint32_t)(t % W);
    rows[ids[i] * Wp + w] = words[t];
}

int cid_index_upload_rows(cid_index* ix, const uint64_t* row_ids, const uint32_t* words, uint64_t nrows) {
    cid_ctx* ctx = ix->ctx;
    cudaStream_t st = ctx->stream;
    CID_CUDA(cudaSetDevice(ctx->device));
    // chunks of rows staged through two pinned windows (the caller's arrays are pageable), scattered on the device
    const uint64_t chunk = std::max<uint64_t>(1024, (64ull << 20) / (8 + 4 * (uint64_t)ix->W));   // ~64 MB windows
    const size_t idb = chunk * 8, wb = chunk * (size_t)ix->W * 4;
    cudaEvent_t freed[2] = {nullptr, nullptr};
    for (int b = 0; b < 2; b++) {
        CID_TRY(ctx->pinned[1 + b].ensure(idb + wb));
        CID_TRY(ctx->scratch[22 + b].ensure(idb + wb));
        CID_CUDA(cudaEventCreateWithFlags(&freed[b], cudaEventDisableTiming));
    }
    int rc = CID_OK;
    for (uint64_t r0 = 0, c = 0; r0 < nrows && rc == CID_OK; r0 += chunk, c++) {
        const int b = (int)(c & 1);
        const uint64_t n = std::min(chunk, nrows - r0);
        if (cudaEventSynchronize(freed[b]) != cudaSuccess) { rc = CID_E_CUDA; break; }     // window's previous copy is done
        for (uint64_t i = r0; i < r0 + n; i++)
            if (row_ids[i] >= ix->S) { set_error("row id %llu >= bloom_size", (unsigned long long)row_ids[i]); rc = CID_E_INVALID; break; }
        if (rc != CID_OK) break;
        char* h = ctx->pinned[1 + b].as<char>();
        memcpy(h, row_ids + r0, n * 8);
        memcpy(h + idb, words + r0 * ix->W, n * (size_t)ix->W * 4);
        char* d = ctx->scratch[22 + b].as<char>();
        if (cudaMemcpyAsync(d, h, n * 8, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(d + idb, h + idb, n * (size_t)ix->W * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = CID_E_CUDA; break; }
        cudaEventRecord(freed[b], st);
        const uint64_t total = n * ix->W;
        scatter_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const uint64_t*)d, (const uint32_t*)(d + idb), n, ix->W,
                                                                            ix->Wp, ix->rows);
        ctx->launches++;
    }
    cudaStreamSynchronize(st);
    for (int b = 0; b < 2; b++) if (freed[b]) cudaEventDestroy(freed[b]);
    if (rc == CID_E_CUDA) { set_error("cid_index_upload_rows: CUDA copy failed: %s", cudaGetErrorString(cudaGetLastError())); return rc; }
    if (rc != CID_OK) return rc;
    CID_CUDA(cudaGetLastError());
    return cid_index_refresh_rownz(ix);
}

int cid_index_download_dense(cid_index* ix, uint32_t* words) {
    cid_ctx* ctx = ix->ctx;
    CID_CUDA(cudaSetDevice(ctx->device));
    CID_CUDA(cudaMemcpy2DAsync(words, (size_t)ix->W * 4, ix->rows, (size_t)ix->Wp * 4, (size_t)ix->W * 4, ix->S,
                               cudaMemcpyDeviceToHost, ctx->stream));
    CID_CUDA(cudaStreamSynchronize(ctx->stream));
    return CID_OK;
}

int cid_index_count_nonzero_rows(cid_index* ix, uint64_t* nrows) {
    cid_ctx* ctx = ix->ctx;
    CID_CUDA(cudaSetDevice(ctx->device));
    std::vector<uint32_t> bm(ix->rownz_words);
    CID_CUDA(cudaMemcpy(bm.data(), ix->rownz, ix->rownz_words * 4, cudaMemcpyDeviceToHost));
    uint64_t n = 0;
    for (uint64_t w = 0; w < ix->rownz_words; w++) n += __builtin_popcount(bm[w]);
    *nrows = n;
    return CID_OK;
}

int cid_index_download_nonzero_rows(cid_index* ix, uint64_t* row_ids, uint32_t* words, uint64_t cap, uint64_t* nrows) {
    cid_ctx* ctx = ix->ctx;
    CID_CUDA(cudaSetDevice(ctx->device));
    std::vector<uint32_t> bm(ix->rownz_words);
    CID_CUDA(cudaMemcpy(bm.data(), ix->rownz, ix->rownz_words * 4, cudaMemcpyDeviceToHost));
    // stream the dense matrix through a pinned window and keep the non-zero rows
    const uint64_t win_rows = std::max<uint64_t>(1, (64ull << 20) / ((size_t)ix->Wp * 4));
    CID_TRY(ctx->pinned[0].ensure(win_rows * ix->Wp * 4));
    uint32_t* win = ctx->pinned[0].as<uint32_t>();
    uint64_t n = 0;
    for (uint64_t r0 = 0; r0 < ix->S; r0 += win_rows) {
        uint64_t nr = std::min(win_rows, ix->S - r0);
        CID_CUDA(cudaMemcpyAsync(win, ix->rows + r0 * ix->Wp, nr * ix->Wp * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CID_CUDA(cudaStreamSynchronize(ctx->stream));
        for (uint64_t r = r0; r < r0 + nr; r++) {
            if (!((bm[r >> 5] >> (r & 31)) & 1u)) continue;
            if (n < cap) {
                row_ids[n] = r;
                memcpy(words + n * ix->W, win + (r - r0) * ix->Wp, (size_t)ix->W * 4);
            }
            n++;
        }
    }
    *nrows = n;
    if (n > cap) { set_error("download_nonzero_rows: %llu rows, capacity %llu", (unsigned long long)n, (unsigned long long)cap); return CID_E_CAPACITY; }
    return CID_OK;
}

int cid_index_device_ptrs(cid_index* ix, void** rows, void** rownz_bitmap, uint64_t* rownz_words) {
    if (rows) *rows = ix->rows;
    if (rownz_bitmap) *rownz_bitmap = ix->rownz;
    if (rownz_words) *rownz_words = ix->rownz_words;
    return CID_OK;
}

// ------------------------------------------------------------------ build
// mini_variant: -1 = plain k-mer index; CID_MINI_OF_KMERS / CID_MINI_COUNTED for .mxi indexes
static int build_accession_impl(cid_index* ix, uint32_t colour, const char* d_bases, const uint64_t* d_seq_offs,
                                uint64_t nseq, uint64_t nbases, int seq_mode, int64_t cutoff, int mini_variant,
                                uint64_t* n_ref_kmers, int64_t* cutoff_used) {
    cid_ctx* ctx = ix->ctx;
    cudaStream_t st = ctx->stream;
    if (colour >= ix->N) { set_error("colour %u >= n_colors %u", colour, ix->N); return CID_E_INVALID; }
    if ((mini_variant >= 0) != (ix->m != 0)) {
        set_error(ix->m ? "this index holds minimizers: use cid_build_accession_mini" : "cid_build_accession_mini needs cid_index_set_minimizer first");
        return CID_E_INVALID;
    }
    if (mini_variant > CID_MINI_COUNTED) { set_error("bad minimizer build variant %d", mini_variant); return CID_E_INVALID; }
    const uint32_t count_m = mini_variant == CID_MINI_COUNTED ? ix->m : 0;   // the count table holds minimizers
    const uint32_t bloom_m = mini_variant == CID_MINI_OF_KMERS ? ix->m : 0;  // ... or k-mers whose minimizer is inserted
    if (seq_mode != CID_SEQ_FASTA && seq_mode != CID_SEQ_FASTQ) { set_error("bad seq_mode"); return CID_E_INVALID; }
    if (cutoff < -1) { set_error("cutoff must be >= -1"); return CID_E_INVALID; }
    CID_CUDA(cudaSetDevice(ctx->device));
    CID_TRY(ensure_bitsets(ix));
    uint32_t* bitset = ix->bitsets + (uint64_t)colour * ix->bs_words;
    // kmers_from_fq_qual / kmers_fq_pe_qual (kmer.rs:461-510,581-655) keep the case of their k-mers.  The fast tables hold
    // 2-bit codes only: a pass that meets a lower-case k-mer reports it, and the accession is redone through the 16-byte
    // count table whose slots carry a case mask (k-mer indexes; minimizer indexes refuse such input).
    struct CaseGuard { cid_ctx* c; ~CaseGuard() { c->case_aware = false; } } case_guard{ctx};
    const bool can_case = seq_mode == CID_SEQ_FASTQ && mini_variant < 0;
    // No count filter (FASTA without -f: build.rs:86-87; or -f 0): only the SET of k-mers matters.  A key-only table
    // (8-byte slots, L2-resident for a bacterial genome) replaces the count table and new keys go straight into the bitset.
    if (ctx->opt_build_set && (cutoff == 0 || (cutoff == -1 && seq_mode == CID_SEQ_FASTA))) {
        const uint64_t npos0 = nbases >= ix->k ? nbases - ix->k + 1 : 0;
        uint64_t slots = next_pow2(std::max<uint64_t>(1024, npos0 + npos0 / 2));
        // read sets repeat every k-mer ~coverage times: start small, grow on overflow
        if (seq_mode == CID_SEQ_FASTQ && nseq >= 4096) slots = next_pow2(std::max<uint64_t>(1024, npos0 / 4));
        for (;;) {
            CID_TRY(ctx->scratch[0].ensure(slots * 8));
            CID_CUDA(cudaMemsetAsync(ctx->scratch[0].p, 0xFF, slots * 8, st));
            CID_CUDA(cudaMemsetAsync(bitset, 0, ix->bs_words * 4, st));
            CID_CUDA(cudaMemsetAsync(ctx->d_err + 1, 0, 4, st));
            CID_TRY(launch_kmerize_bloom(ctx, st, (const uint8_t*)d_bases, d_seq_offs, nseq, nbases, ctx->scratch[0].p, slots, ix->k,
                                         seq_mode, count_m, bloom_m, ix->H, make_mods(ix->S, ix->hv, (const HashCfg*)ix->d_hcfg), bitset));
            CID_CUDA(cudaMemcpyAsync(ctx->h_err, ctx->d_err, 8, cudaMemcpyDeviceToHost, st));
            CID_CUDA(cudaStreamSynchronize(st));
            const uint32_t flags = ctx->h_err[0], distinct = ctx->h_err[1];
            if ((flags & ERRF_LOWER_RAW) && can_case) {           // lower-case k-mers: through the case-aware count table
                CID_CUDA(cudaMemsetAsync(ctx->d_err, 0, 8, st));
                ctx->case_aware = true;
                break;
            }
            if ((flags & ERRF_TABLE_FULL) || (uint64_t)distinct * 10 > slots * 7) {       // too full to trust the probe limit: redo larger
                if (slots >= next_pow2(std::max<uint64_t>(1024, 2 * npos0))) { set_error("k-mer set overflow"); return CID_E_CAPACITY; }
                CID_CUDA(cudaMemsetAsync(ctx->d_err, 0, 8, st));
                slots *= 2;
                continue;
            }
            CID_TRY(check_err_flags(ctx, st));
            if (n_ref_kmers) *n_ref_kmers = distinct;
            if (cutoff_used) *cutoff_used = cutoff;
            return CID_OK;
        }
    }
    uint64_t nslots; uint64_t *d_off, *d_mask;
    // Read sets (many short sequences, deep coverage) hold far fewer DISTINCT k-mers than k-mer positions, and the
    // table is scanned three times after counting (clear, histogram, Bloom insert) while its inserts are random DRAM
    // accesses whose L2 hit rate falls with the table size (36 % at 1 GB): size it optimistically -- for the
    // distinct/positions ratio the previous read set of this context had (accessions of one build are sequenced alike;
    // positions/4 for the first one), aiming at 2/3 load -- and double it when it ends up fuller than 80 %.
    const uint64_t npos = nbases >= ix->k ? nbases - ix->k + 1 : 0;
    uint64_t hint = 0;
    if (seq_mode == CID_SEQ_FASTQ && nseq >= 4096 && npos >= (1ull << 22)) {
        if (ctx->opt_build_table_div > 0) hint = next_pow2(npos / (uint64_t)ctx->opt_build_table_div);
        else if (ctx->readset_ratio > 0) hint = next_pow2((uint64_t)(ctx->readset_ratio * (double)npos / 0.68) + 1);
        else hint = next_pow2(npos / 4);
        if (hint >= next_pow2(std::max<uint64_t>(64, 2 * npos))) hint = 0;
    }
    // keys of <= 21 bases (k, or the minimizer length of build_multi_mini) leave room for the count in the same 8-byte
    // word: half the table, one atomic per insert.  A multiplicity near 2^22 sends the accession back to 16-byte slots.
    bool packed = ctx->opt_build_packed && (count_m ? count_m : ix->k) <= 21 && !ctx->case_aware;
    for (;;) {
        CID_TRY(single_region(ctx, st, nbases, ix->k, &nslots, &d_off, &d_mask, hint, packed));
        CID_CUDA(cudaMemsetAsync(ctx->d_err + 1, 0, 4, st));
        CID_TRY(launch_kmerize_insert(ctx, st, (const uint8_t*)d_bases, d_seq_offs, nseq, 0, nbases, nullptr, d_off, d_mask,
                                      ctx->scratch[0].p, ix->k, seq_mode, count_m, packed ? nslots : 0));
        CID_CUDA(cudaMemcpyAsync(ctx->h_err, ctx->d_err, 8, cudaMemcpyDeviceToHost, st));
        CID_CUDA(cudaStreamSynchronize(st));
        const uint32_t flags = ctx->h_err[0], distinct = ctx->h_err[1];
        if ((flags & ERRF_LOWER_RAW) && can_case && !ctx->case_aware) {
            CID_CUDA(cudaMemsetAsync(ctx->d_err, 0, 8, st));
            ctx->case_aware = true;
            packed = false;
            continue;
        }
        if (flags & ERRF_COUNT_OVERFLOW) {
            CID_CUDA(cudaMemsetAsync(ctx->d_err, 0, 8, st));
            packed = false;
            continue;
        }
        if (hint && ((flags & ERRF_TABLE_FULL) || (uint64_t)distinct * 10 > nslots * 8)) {        // fuller than 80 %: redo larger
            CID_CUDA(cudaMemsetAsync(ctx->d_err, 0, 8, st));
            hint = hint * 2 >= next_pow2(std::max<uint64_t>(64, 2 * npos)) ? 0 : hint * 2;     // 0 = the safe 2x-positions size
            continue;
        }
        break;
    }
    CID_TRY(check_err_flags(ctx, st));
    if (seq_mode == CID_SEQ_FASTQ && npos) ctx->readset_ratio = (double)ctx->h_err[1] / (double)npos;
    int64_t used = cutoff;       // FASTA with -1: keep everything (count > -1)
    if (seq_mode == CID_SEQ_FASTQ && cutoff == -1)
        CID_TRY(region_auto_cutoff(ctx, st, ctx->scratch[0].p, nslots, &used, packed));
    CID_CUDA(cudaMemsetAsync(bitset, 0, ix->bs_words * 4, st));
    CID_TRY(ctx->scratch[2].ensure(16));
    unsigned long long* d_nref = ctx->scratch[2].as<unsigned long long>();
    CID_CUDA(cudaMemsetAsync(d_nref, 0, 8, st));
    CID_TRY(launch_region_to_bloom(ctx, st, ctx->scratch[0].p, nslots, used, count_m ? count_m : ix->k, bloom_m, ix->H,
                                   make_mods(ix->S, ix->hv, (const HashCfg*)ix->d_hcfg), bitset, d_nref, packed));
    unsigned long long nref = 0;
    CID_CUDA(cudaMemcpyAsync(&nref, d_nref, 8, cudaMemcpyDeviceToHost, st));
    CID_CUDA(cudaStreamSynchronize(st));
    if (n_ref_kmers) *n_ref_kmers = nref;
    if (cutoff_used) *cutoff_used = used;
    return CID_OK;
}

int cid_build_accession_dev(cid_index* ix, uint32_t colour, const char* d_bases, const uint64_t* d_seq_offs,
                            uint64_t nseq, uint64_t nbases, int seq_mode, int64_t cutoff, uint64_t* n_ref_kmers,
                            int64_t* cutoff_used) {
    return build_accession_impl(ix, colour, d_bases, d_seq_offs, nseq, nbases, seq_mode, cutoff, -1, n_ref_kmers, cutoff_used);
}

static int build_accession_host(cid_index* ix, uint32_t colour, const char* bases, const uint64_t* seq_offs, uint64_t nseq,
                                int seq_mode, int64_t cutoff, int mini_variant, uint64_t* n_ref_kmers, int64_t* cutoff_used) {
    cid_ctx* ctx = ix->ctx;
    if (!seq_offs) { set_error("null seq_offs"); return CID_E_INVALID; }
    CID_CUDA(cudaSetDevice(ctx->device));
    const uint64_t nbases = seq_offs[nseq];
    CID_TRY(ctx->scratch[4].ensure(nbases + 64));
    CID_TRY(ctx->scratch[5].ensure((nseq + 1) * 8));
    if (nbases) CID_CUDA(cudaMemcpyAsync(ctx->scratch[4].p, bases, nbases, cudaMemcpyHostToDevice, ctx->stream));
    CID_CUDA(cudaMemcpyAsync(ctx->scratch[5].p, seq_offs, (nseq + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    return build_accession_impl(ix, colour, ctx->scratch[4].as<char>(), ctx->scratch[5].as<uint64_t>(), nseq, nbases,
                                seq_mode, cutoff, mini_variant, n_ref_kmers, cutoff_used);
}
int cid_build_accession(cid_index* ix, uint32_t colour, const char* bases, const uint64_t* seq_offs, uint64_t nseq,
                        int seq_mode, int64_t cutoff, uint64_t* n_ref_kmers, int64_t* cutoff_used) {
    return build_accession_host(ix, colour, bases, seq_offs, nseq, seq_mode, cutoff, -1, n_ref_kmers, cutoff_used);
}
int cid_build_accession_mini(cid_index* ix, uint32_t colour, const char* bases, const uint64_t* seq_offs, uint64_t nseq,
                             int seq_mode, int64_t cutoff, int variant, uint64_t* n_ref_kmers, int64_t* cutoff_used) {
    if (variant != CID_MINI_OF_KMERS && variant != CID_MINI_COUNTED) { set_error("bad minimizer build variant %d", variant); return CID_E_INVALID; }
    return build_accession_host(ix, colour, bases, seq_offs, nseq, seq_mode, cutoff, variant, n_ref_kmers, cutoff_used);
}

int cid_index_set_minimizer(cid_index* ix, uint32_t m_size) {
    if (!ix) { set_error("cid_index_set_minimizer: null index"); return CID_E_INVALID; }
    // find_minimizer slices seq[..m] (kmer.rs:974): the reference panics for m > k
    if (m_size > ix->k) { set_error("minimizer size %u larger than k-mer size %u (the reference panics)", m_size, ix->k); return CID_E_REF_PANIC; }
    ix->m = m_size;
    return CID_OK;
}
uint32_t cid_index_minimizer(const cid_index* ix) { return ix ? ix->m : 0; }

int cid_index_set_hash_variant(cid_index* ix, uint32_t variant) {
    if (!ix) { set_error("cid_index_set_hash_variant: null index"); return CID_E_INVALID; }
    if (variant >= CID_HASH_VARIANTS) { set_error("hash variant %u out of range (0..%d)", variant, CID_HASH_VARIANTS - 1); return CID_E_INVALID; }
    ix->hv = variant;
    const HashCfg h = make_hashcfg(variant);
    CID_CUDA(cudaSetDevice(ix->ctx->device));
    CID_CUDA(cudaMemcpy(ix->d_hcfg, &h, sizeof(h), cudaMemcpyHostToDevice));
    return CID_OK;
}
uint32_t cid_index_hash_variant(const cid_index* ix) { return ix ? ix->hv : 0; }

int cid_build_finalize(cid_index* ix) {
    cid_ctx* ctx = ix->ctx;
    CID_CUDA(cudaSetDevice(ctx->device));
    if (!ix->bitsets) return cid_index_refresh_rownz(ix);
    CID_TRY(launch_transpose(ctx, ctx->stream, ix));
    CID_TRY(launch_rownz(ctx, ctx->stream, ix));
    CID_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ix->bitsets);
    ix->bitsets = nullptr;
    ix->rownz_valid = true;
    return CID_OK;
}

// ------------------------------------------------------------------ search
// Shared front end: table regions for a batch of queries, k-mer counting, work units.
struct QueryPlan {
    GroupRegions gr;
    QueryUnits qu;
    uint32_t* d_unit_group; uint64_t* d_unit_slot0; uint32_t* d_unit_nslots;
    Slot* d_table;
    uint64_t kmers_bound = 0;      // upper bound on the distinct k-mers in the table
};
static int query_front(cid_index* ix, cudaStream_t st, const uint8_t* d_bases, const uint64_t* d_seq_offs,
                       const uint64_t* h_seq_offs, const uint64_t* h_query_offs, uint64_t q0, uint64_t q1, int seq_mode,
                       QueryPlan& qp) {
    cid_ctx* ctx = ix->ctx;
    const uint64_t nq = q1 - q0;
    const uint64_t s_lo = h_query_offs[q0], s_hi = h_query_offs[q1];
    const uint64_t nseq = s_hi - s_lo;
    plan_regions(h_seq_offs, h_query_offs + q0, nq, ix->k, qp.gr);
    // A read-set query (one query, many short sequences, deep coverage) holds far fewer DISTINCT k-mers than k-mer positions,
    // and its table is scanned several times after counting: size it optimistically like the read-set build does (positions/4,
    // doubled while it ends up fuller than 80 %; the safe 2x-positions size is the last resort).
    const uint64_t safe_slots = qp.gr.total_slots;
    uint64_t hint = 0;
    if (ctx->opt_query_table_div > 0 && nq == 1 && seq_mode == CID_SEQ_FASTQ && safe_slots >= ctx->opt_query_table_min &&
        (nseq >= 4096 || ctx->opt_query_table_min == 0)) {
        hint = next_pow2(safe_slots / 2 / (uint64_t)ctx->opt_query_table_div);
        if (hint >= safe_slots) hint = 0;
    }
    // sequence -> group
    std::vector<uint32_t> seq_group(nseq);
    for (uint64_t q = q0; q < q1; q++)
        for (uint64_t s = h_query_offs[q]; s < h_query_offs[q + 1]; s++) seq_group[s - s_lo] = (uint32_t)(q - q0);
    CID_TRY(ctx->scratch[1].ensure(nq * 16 + 64));
    CID_TRY(ctx->scratch[3].ensure(nseq * 4 + 64));
    if (nseq) CID_CUDA(cudaMemcpyAsync(ctx->scratch[3].p, seq_group.data(), nseq * 4, cudaMemcpyHostToDevice, st));
    for (;;) {
        if (hint) { qp.gr.off[0] = 0; qp.gr.mask[0] = hint - 1; qp.gr.total_slots = hint; }
        else if (qp.gr.total_slots != safe_slots) plan_regions(h_seq_offs, h_query_offs + q0, nq, ix->k, qp.gr);
        plan_units(qp.gr, QUERY_ITEM_SLOTS, qp.qu);
        const uint64_t nunits = qp.qu.group.size();
        CID_TRY(ctx->scratch[0].ensure(qp.gr.total_slots * sizeof(Slot)));
        CID_TRY(ctx->scratch[6].ensure(nunits * 16 + 64));
        uint64_t* d_off = ctx->scratch[1].as<uint64_t>();
        uint64_t* d_mask = d_off + nq;
        CID_CUDA(cudaMemcpyAsync(d_off, qp.gr.off.data(), nq * 8, cudaMemcpyHostToDevice, st));
        CID_CUDA(cudaMemcpyAsync(d_mask, qp.gr.mask.data(), nq * 8, cudaMemcpyHostToDevice, st));
        qp.d_unit_slot0 = ctx->scratch[6].as<uint64_t>();
        qp.d_unit_group = (uint32_t*)(qp.d_unit_slot0 + nunits);
        qp.d_unit_nslots = qp.d_unit_group + nunits;
        if (nunits) {
            CID_CUDA(cudaMemcpyAsync(qp.d_unit_slot0, qp.qu.slot0.data(), nunits * 8, cudaMemcpyHostToDevice, st));
            CID_CUDA(cudaMemcpyAsync(qp.d_unit_group, qp.qu.group.data(), nunits * 4, cudaMemcpyHostToDevice, st));
            CID_CUDA(cudaMemcpyAsync(qp.d_unit_nslots, qp.qu.nslots.data(), nunits * 4, cudaMemcpyHostToDevice, st));
        }
        qp.d_table = ctx->scratch[0].as<Slot>();
        CID_TRY(table_clear(ctx, st, qp.d_table, qp.gr.total_slots));
        if (hint) CID_CUDA(cudaMemsetAsync(ctx->d_err + 1, 0, 4, st));
        // the batch's sequences are [s_lo, s_hi): bases [h_seq_offs[s_lo], h_seq_offs[s_hi])
        const uint64_t b_lo = h_seq_offs[s_lo], b_hi = h_seq_offs[s_hi];
        // kmerize works on absolute base offsets: pass the offsets slice and the base range it covers
        CID_TRY(launch_kmerize_insert(ctx, st, d_bases, d_seq_offs + s_lo, nseq, b_lo, b_hi,
                                            ctx->scratch[3].as<uint32_t>(), d_off, d_mask, qp.d_table, ix->k, seq_mode));
        qp.kmers_bound = qp.gr.total_slots / 2;       // safe size: 2 slots per k-mer position
        if (hint) {
            CID_CUDA(cudaMemcpyAsync(ctx->h_err, ctx->d_err, 8, cudaMemcpyDeviceToHost, st));
            CID_CUDA(cudaStreamSynchronize(st));
            const uint32_t flags = ctx->h_err[0], distinct = ctx->h_err[1];
            if ((flags & ERRF_TABLE_FULL) || (uint64_t)distinct * 10 > hint * 8) {      // fuller than 80 %: redo larger
                CID_CUDA(cudaMemsetAsync(ctx->d_err, 0, 4, st));        // (other flags come up again in the redo)
                hint = hint * 2 >= safe_slots ? 0 : hint * 2;
                continue;
            }
            qp.kmers_bound = distinct;
        }
        break;
    }
    // pageable host vectors above must stay alive until the copies complete
    CID_CUDA(cudaStreamSynchronize(st));
    return CID_OK;
}

// Splits [0,nq) into batches whose count tables fit `max_slots`.
static std::vector<uint64_t> query_batches(const uint64_t* h_seq_offs, const uint64_t* h_query_offs, uint64_t nq,
                                           uint32_t k, uint64_t max_slots) {
    std::vector<uint64_t> cuts{0};
    uint64_t acc = 0;
    for (uint64_t q = 0; q < nq; q++) {
        uint64_t nb = h_seq_offs[h_query_offs[q + 1]] - h_seq_offs[h_query_offs[q]];
        uint64_t npos = nb >= k ? nb - k + 1 : 0;
        uint64_t slots = next_pow2(std::max<uint64_t>(64, 2 * npos));
        if (acc && acc + slots > max_slots) { cuts.push_back(q); acc = 0; }
        acc += slots;
    }
    cuts.push_back(nq);
    return cuts;
}
static const uint64_t kMaxBatchSlots = 1ull << 28;   // 4 GiB of count table per pass

// The shared-memory front end applies when only DISTINCT k-mers matter (every query's filter is 0), no unique-hit
// summaries are wanted, the streaming gather handles this row shape and every query is small.
static bool query_front_ok(const cid_index* ix, bool want_uniq, const uint64_t* h_seq_offs, const uint64_t* h_query_offs, uint64_t nq) {
    return ix->ctx->opt_query_front && !ix->ctx->opt_query_fused && !want_uniq && ix->Wp >= 4 && ix->Wp <= 512 &&
           (ix->H == 2 || ix->H == 4) && query_front_fits(h_seq_offs, h_query_offs, 0, nq, ix->k);
}
// batches of at most 2^28 k-mer positions (row-index lists of <= 4 GiB at H = 4)
static std::vector<uint64_t> query_front_batches(const uint64_t* h_seq_offs, const uint64_t* h_query_offs, uint64_t nq) {
    std::vector<uint64_t> cuts{0};
    uint64_t acc = 0;
    for (uint64_t q = 0; q < nq; q++) {
        const uint64_t nb = h_seq_offs[h_query_offs[q + 1]] - h_seq_offs[h_query_offs[q]];
        if (acc && acc + nb > (1ull << 28)) { cuts.push_back(q); acc = 0; }
        acc += nb;
    }
    cuts.push_back(nq);
    return cuts;
}

// work units -> scratch[6] (synchronous: `qu` may be a temporary)
static int upload_units(cid_ctx* ctx, cudaStream_t st, const QueryUnits& qu, QueryPlan& qp) {
    const uint64_t nu = qu.group.size();
    CID_TRY(ctx->scratch[6].ensure(nu * 16 + 64));
    qp.d_unit_slot0 = ctx->scratch[6].as<uint64_t>();
    qp.d_unit_group = (uint32_t*)(qp.d_unit_slot0 + nu);
    qp.d_unit_nslots = qp.d_unit_group + nu;
    if (nu) {
        CID_CUDA(cudaMemcpyAsync(qp.d_unit_slot0, qu.slot0.data(), nu * 8, cudaMemcpyHostToDevice, st));
        CID_CUDA(cudaMemcpyAsync(qp.d_unit_group, qu.group.data(), nu * 4, cudaMemcpyHostToDevice, st));
        CID_CUDA(cudaMemcpyAsync(qp.d_unit_nslots, qu.nslots.data(), nu * 4, cudaMemcpyHostToDevice, st));
        CID_CUDA(cudaStreamSynchronize(st));
    }
    return CID_OK;
}

// Survivors (count > filt[q]) of every query region copied into one dense slot list on scratch[22] (query q at the prefix sum
// of surv[]), gather work units re-planned over it.  surv[q] == UINT64_MAX: not known yet (counted here).  Unless `force`,
// nothing changes when the table is less than 3/4 empty.
static int compact_survivors(cid_ctx* ctx, cudaStream_t st, QueryPlan& qp, const std::vector<int64_t>& filt, std::vector<uint64_t>& surv,
                             bool force, const void** d_slots, uint64_t* slots_total) {
    const uint64_t bq = filt.size();
    CID_TRY(ctx->scratch[21].ensure(bq * 8 + 8));
    unsigned long long* d_surv = ctx->scratch[21].as<unsigned long long>();
    CID_CUDA(cudaMemsetAsync(d_surv, 0, bq * 8, st));
    bool counted = false;
    for (uint64_t q = 0; q < bq; q++)
        if (surv[q] == UINT64_MAX) {          // no histogram at hand (fixed filter): count pass
            CID_TRY(launch_region_compact(ctx, st, qp.d_table + qp.gr.off[q], qp.gr.mask[q] + 1, filt[q], nullptr, d_surv + q, 0));
            counted = true;
        }
    if (counted) {
        std::vector<unsigned long long> hs(bq);
        CID_CUDA(cudaMemcpyAsync(hs.data(), d_surv, bq * 8, cudaMemcpyDeviceToHost, st));
        CID_CUDA(cudaStreamSynchronize(st));
        for (uint64_t q = 0; q < bq; q++) if (surv[q] == UINT64_MAX) surv[q] = hs[q];
        CID_CUDA(cudaMemsetAsync(d_surv, 0, bq * 8, st));
    }
    uint64_t dense_total = 0;
    std::vector<uint64_t> dbase(bq);
    for (uint64_t q = 0; q < bq; q++) { dbase[q] = dense_total; dense_total += surv[q]; }
    if (!force && dense_total * 4 > qp.gr.total_slots) return CID_OK;       // worth it only when the table is mostly empty
    CID_TRY(ctx->scratch[22].ensure((dense_total + 1) * sizeof(Slot)));
    Slot* d_dense = ctx->scratch[22].as<Slot>();
    QueryUnits du;
    for (uint64_t q = 0; q < bq; q++) {
        if (surv[q] == 0) continue;
        CID_TRY(launch_region_compact(ctx, st, qp.d_table + qp.gr.off[q], qp.gr.mask[q] + 1, filt[q], d_dense + dbase[q], d_surv + q, surv[q]));
        for (uint64_t at = 0; at < surv[q]; at += QUERY_ITEM_SLOTS) {
            du.group.push_back((uint32_t)q);
            du.slot0.push_back(dbase[q] + at);
            du.nslots.push_back((uint32_t)std::min<uint64_t>(QUERY_ITEM_SLOTS, surv[q] - at));
        }
    }
    CID_TRY(upload_units(ctx, st, du, qp));
    qp.qu = std::move(du);
    *d_slots = d_dense;
    *slots_total = dense_total;
    return CID_OK;
}

// reports.rs:20-26 from the unique-hit triples (query in batch, accession, multiplicity) of one batch: per (query, accession)
// their number ("specific"), the sum (mean = sum / number) and the mode of the multiplicities, written at [(q0 + q) * N + c].
static int summarize_uniq(cid_ctx* ctx, cudaStream_t st, const uint32_t* d_list, uint32_t nu, uint32_t N, uint64_t bq, uint64_t q0,
                          uint64_t* uniq_n, uint64_t* uniq_sum, uint64_t* uniq_mode) {
    // on the device when a [query x accession][multiplicity] histogram fits 64 MB (one FASTQ query: 628k triples took 13 ms
    // through the host maps below, 2/3 of the whole search)
    const uint64_t cells = bq * N;
    const uint32_t MB = cells * 1024 <= (16u << 20) ? 1024u : cells * 256 <= (16u << 20) ? 256u : 0u;
    if (MB && ctx->opt_uniq_device) {
        CID_TRY(ctx->scratch[19].ensure(cells * MB * 4 + 16));
        CID_TRY(ctx->scratch[20].ensure(cells * 24));
        uint32_t* d_hist = ctx->scratch[19].as<uint32_t>();
        uint32_t* d_ovf = d_hist + cells * MB;
        unsigned long long* d_un = ctx->scratch[20].as<unsigned long long>();
        CID_TRY(launch_uniq_summaries(ctx, st, d_list, nu, N, cells, MB, d_hist, d_ovf, d_un, d_un + cells, d_un + 2 * cells));
        uint32_t ovf = 0;
        CID_CUDA(cudaMemcpyAsync(&ovf, d_ovf, 4, cudaMemcpyDeviceToHost, st));
        CID_CUDA(cudaStreamSynchronize(st));
        if (!ovf) {
            if (uniq_n) CID_CUDA(cudaMemcpyAsync(uniq_n + q0 * N, d_un, cells * 8, cudaMemcpyDeviceToHost, st));
            if (uniq_sum) CID_CUDA(cudaMemcpyAsync(uniq_sum + q0 * N, d_un + cells, cells * 8, cudaMemcpyDeviceToHost, st));
            if (uniq_mode) CID_CUDA(cudaMemcpyAsync(uniq_mode + q0 * N, d_un + 2 * cells, cells * 8, cudaMemcpyDeviceToHost, st));
            CID_CUDA(cudaStreamSynchronize(st));
            return CID_OK;
        }
    }
    std::vector<uint32_t> ul((size_t)nu * 3);
    if (nu) CID_CUDA(cudaMemcpy(ul.data(), d_list, (size_t)nu * 12, cudaMemcpyDeviceToHost));
    // Mode ties are broken towards the smallest value (the reference's tie-break is hash-order dependent).
    std::map<std::pair<uint64_t, uint32_t>, std::map<uint32_t, uint64_t>> freq;
    for (uint32_t i = 0; i < nu; i++) freq[{q0 + ul[3 * i], ul[3 * i + 1]}][ul[3 * i + 2]] += 1;
    for (auto& kv : freq) {
        uint64_t n = 0, sum = 0, mode = 0, best = 0;
        for (auto& fv : kv.second) { n += fv.second; sum += (uint64_t)fv.first * fv.second; if (fv.second > best) { best = fv.second; mode = fv.first; } }
        size_t at = kv.first.first * N + kv.first.second;
        if (uniq_n) uniq_n[at] = n;
        if (uniq_sum) uniq_sum[at] = sum;
        if (uniq_mode) uniq_mode[at] = mode;
    }
    return CID_OK;
}

int cid_query_counts(cid_index* ix, const char* bases, const uint64_t* seq_offs, uint64_t nseq,
                     const uint64_t* query_offs, uint64_t nq, int seq_mode, int gene_search, int64_t filter,
                     uint32_t* counts, uint64_t* num_kmers, uint64_t* uniq_n, uint64_t* uniq_sum, uint64_t* uniq_mode,
                     int64_t* cutoff_used) {
    cid_ctx* ctx = ix->ctx;
    cudaStream_t st = ctx->stream;
    if (!seq_offs || !query_offs || !counts || !num_kmers) { set_error("cid_query_counts: null argument"); return CID_E_INVALID; }
    if (ix->m) { set_error("An index with minimizers (.mxi) is used, but not available for this function"); return CID_E_UNSUPPORTED; }   // main.rs:569-573
    if (seq_mode != CID_SEQ_FASTA && seq_mode != CID_SEQ_FASTQ) { set_error("bad seq_mode"); return CID_E_INVALID; }
    CID_CUDA(cudaSetDevice(ctx->device));
    const uint64_t nbases = seq_offs[nseq];
    CID_TRY(ctx->scratch[4].ensure(nbases + 64));
    CID_TRY(ctx->scratch[5].ensure((nseq + 1) * 8));
    if (nbases) CID_CUDA(cudaMemcpyAsync(ctx->scratch[4].p, bases, nbases, cudaMemcpyHostToDevice, st));
    CID_CUDA(cudaMemcpyAsync(ctx->scratch[5].p, seq_offs, (nseq + 1) * 8, cudaMemcpyHostToDevice, st));
    const uint8_t* d_bases = ctx->scratch[4].as<uint8_t>();
    const uint64_t* d_seq_offs = ctx->scratch[5].as<uint64_t>();
    const bool want_uniq = uniq_n || uniq_sum || uniq_mode;
    const uint32_t N = ix->N;
    if (uniq_n) memset(uniq_n, 0, nq * N * 8);
    if (uniq_sum) memset(uniq_sum, 0, nq * N * 8);
    if (uniq_mode) memset(uniq_mode, 0, nq * N * 8);

    // Raw-case (FASTQ) queries keep the case of their k-mers (kmer.rs:461-510,581-655); the fast tables hold 2-bit codes only.
    // A pass that meets a lower-case k-mer reports it (`lower`) and the work is redone through the count table with
    // case-aware slots (ctx->case_aware, reset when this call returns).
    struct CaseGuard { cid_ctx* c; ~CaseGuard() { c->case_aware = false; } } case_guard{ctx};
    bool lower = false;
    const bool all_distinct = (seq_mode == CID_SEQ_FASTA && gene_search) || filter == 0;
    if (all_distinct && nq && query_front_ok(ix, want_uniq, seq_offs, query_offs, nq)) {
        CID_TRY(ctx->scratch[7].ensure((nq + 1) * 8));
        CID_CUDA(cudaMemcpyAsync(ctx->scratch[7].p, query_offs, (nq + 1) * 8, cudaMemcpyHostToDevice, st));
        std::vector<uint64_t> fcuts = query_front_batches(seq_offs, query_offs, nq);
        for (size_t b = 0; b + 1 < fcuts.size() && !lower; b++) {
            const uint64_t q0 = fcuts[b], q1 = fcuts[b + 1], bq = q1 - q0;
            CID_TRY(ctx->scratch[10].ensure(bq * N * 4));
            CID_TRY(ctx->scratch[11].ensure(bq * 8));
            CID_CUDA(cudaMemsetAsync(ctx->scratch[10].p, 0, bq * N * 4, st));
            CID_CUDA(cudaMemsetAsync(ctx->scratch[11].p, 0, bq * 8, st));
            CID_TRY(launch_query_front_gather(ctx, st, ix, d_bases, d_seq_offs, ctx->scratch[7].as<uint64_t>(), seq_offs, query_offs,
                                              q0, q1, seq_mode, ctx->scratch[10].as<uint32_t>(), ctx->scratch[11].as<unsigned long long>()));
            CID_TRY(check_err_flags(ctx, st, &lower));
            if (lower) break;
            CID_CUDA(cudaMemcpyAsync(counts + q0 * N, ctx->scratch[10].p, bq * N * 4, cudaMemcpyDeviceToHost, st));
            CID_CUDA(cudaMemcpyAsync(num_kmers + q0, ctx->scratch[11].p, bq * 8, cudaMemcpyDeviceToHost, st));
            CID_CUDA(cudaStreamSynchronize(st));
            if (cutoff_used) for (uint64_t q = q0; q < q1; q++) cutoff_used[q] = 0;
        }
        if (!lower) return CID_OK;
        ctx->case_aware = true;          // every query again, through the count table
    }
    std::vector<uint64_t> cuts = query_batches(seq_offs, query_offs, nq, ix->k, kMaxBatchSlots);
    static const bool trace = getenv("CID_TRACE") != nullptr;     // host stage timings on stderr (diagnostics only)
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    for (size_t b = 0; b + 1 < cuts.size(); b++) {
        const uint64_t q0 = cuts[b], q1 = cuts[b + 1], bq = q1 - q0;
        QueryPlan qp;
        const double t_0 = now();
        CID_TRY(query_front(ix, st, d_bases, d_seq_offs, seq_offs, query_offs, q0, q1, seq_mode, qp));
        lower = false;
        CID_TRY(check_err_flags(ctx, st, ctx->case_aware ? nullptr : &lower));
        if (lower) { ctx->case_aware = true; b--; continue; }        // this batch again, case-aware
        const double t_front = now();
        // per-query filter (batch_search_pe.rs:34-39 / :112-120)
        std::vector<int64_t> filt(bq);
        std::vector<uint64_t> surv(bq, UINT64_MAX);      // k-mers with count > filter, where known
        for (uint64_t q = 0; q < bq; q++) {
            if (seq_mode == CID_SEQ_FASTA && gene_search) filt[q] = 0;
            else if (filter < 0) CID_TRY(region_auto_cutoff(ctx, st, (const void*)(qp.d_table + qp.gr.off[q]), qp.gr.mask[q] + 1, &filt[q], false, &surv[q]));
            else filt[q] = filter;
            if (cutoff_used) cutoff_used[q0 + q] = filt[q];
        }
        // Large, sparse count tables (read-set queries): survivors copied once into a dense slot list, work units over that
        const void* d_slots = qp.d_table;
        uint64_t slots_total = qp.gr.total_slots;
        if (ctx->opt_query_compact && (qp.gr.total_slots >= (1ull << 22) || ctx->opt_query_compact == 2) && bq <= 64)
            CID_TRY(compact_survivors(ctx, st, qp, filt, surv, ctx->opt_query_compact == 2, &d_slots, &slots_total));
        const double t_cut = now();
        const uint64_t npos_total = qp.kmers_bound;
        const uint32_t uniq_cap = want_uniq ? (uint32_t)std::min<uint64_t>(npos_total + 1, 0xFFFFFFF0u / 3) : 0;
        CID_TRY(ctx->scratch[7].ensure(bq * 8));
        CID_TRY(ctx->scratch[10].ensure(bq * N * 4));
        CID_TRY(ctx->scratch[11].ensure(bq * 8));
        CID_TRY(ctx->scratch[12].ensure((size_t)uniq_cap * 12 + 16));
        CID_TRY(ctx->scratch[13].ensure(16));
        CID_CUDA(cudaMemcpyAsync(ctx->scratch[7].p, filt.data(), bq * 8, cudaMemcpyHostToDevice, st));
        CID_CUDA(cudaMemsetAsync(ctx->scratch[10].p, 0, bq * N * 4, st));
        CID_CUDA(cudaMemsetAsync(ctx->scratch[11].p, 0, bq * 8, st));
        CID_CUDA(cudaMemsetAsync(ctx->scratch[13].p, 0, 16, st));
        CID_TRY(launch_query_counts(ctx, st, ix, d_slots, qp.d_unit_group, qp.d_unit_slot0, qp.d_unit_nslots,
                                    qp.qu.group.size(), slots_total, ctx->scratch[7].as<int64_t>(), ctx->scratch[10].as<uint32_t>(),
                                    ctx->scratch[11].as<unsigned long long>(), want_uniq, ctx->scratch[12].as<uint32_t>(),
                                    uniq_cap, ctx->scratch[13].as<uint32_t>()));
        CID_CUDA(cudaMemcpyAsync(counts + q0 * N, ctx->scratch[10].p, bq * N * 4, cudaMemcpyDeviceToHost, st));
        CID_CUDA(cudaMemcpyAsync(num_kmers + q0, ctx->scratch[11].p, bq * 8, cudaMemcpyDeviceToHost, st));
        uint32_t nu = 0;
        CID_CUDA(cudaMemcpyAsync(&nu, ctx->scratch[13].p, 4, cudaMemcpyDeviceToHost, st));
        CID_CUDA(cudaStreamSynchronize(st));
        const double t_counts = now();
        if (want_uniq) {
            if (nu > uniq_cap) { set_error("unique-hit list overflow"); return CID_E_CAPACITY; }
            CID_TRY(summarize_uniq(ctx, st, ctx->scratch[12].as<uint32_t>(), nu, N, bq, q0, uniq_n, uniq_sum, uniq_mode));
        }
        if (trace)
            fprintf(stderr, "[cid trace] query_counts batch %zu: %llu queries, H2D + count table + k-mer counting %.2f ms, cutoffs %.2f ms, "
                            "hash + gather %.2f ms, unique-hit summaries (%u entries) %.2f ms\n", b, (unsigned long long)bq, t_front - t_0,
                    t_cut - t_front, t_counts - t_cut, nu, now() - t_counts);
    }
    return CID_OK;
}

// ---- column-sharded default report: three calls around one cross-shard exchange --------------------------
int cid_query_survivors(cid_index* ix, const char* bases, const uint64_t* seq_offs, uint64_t nseq, const uint64_t* query_offs,
                        uint64_t nq, int seq_mode, int gene_search, int64_t filter, void** d_slots, uint64_t* surv,
                        int64_t* cutoff_used) {
    cid_ctx* ctx = ix->ctx;
    cudaStream_t st = ctx->stream;
    if (!seq_offs || !query_offs || !d_slots || !surv) { set_error("cid_query_survivors: null argument"); return CID_E_INVALID; }
    if (ix->m) { set_error("An index with minimizers (.mxi) is used, but not available for this function"); return CID_E_UNSUPPORTED; }
    if (seq_mode != CID_SEQ_FASTA && seq_mode != CID_SEQ_FASTQ) { set_error("bad seq_mode"); return CID_E_INVALID; }
    CID_CUDA(cudaSetDevice(ctx->device));
    *d_slots = nullptr;
    if (nq == 0) return CID_OK;
    if (query_batches(seq_offs, query_offs, nq, ix->k, kMaxBatchSlots).size() != 2) {
        set_error("cid_query_survivors: the queries of one call must fit one count-table pass (2^28 slots); split the call");
        return CID_E_CAPACITY;
    }
    const uint64_t nbases = seq_offs[nseq];
    CID_TRY(ctx->scratch[4].ensure(nbases + 64));
    CID_TRY(ctx->scratch[5].ensure((nseq + 1) * 8));
    if (nbases) CID_CUDA(cudaMemcpyAsync(ctx->scratch[4].p, bases, nbases, cudaMemcpyHostToDevice, st));
    CID_CUDA(cudaMemcpyAsync(ctx->scratch[5].p, seq_offs, (nseq + 1) * 8, cudaMemcpyHostToDevice, st));
    QueryPlan qp;
    struct CaseGuard { cid_ctx* c; ~CaseGuard() { c->case_aware = false; } } case_guard{ctx};
    for (;;) {
        CID_TRY(query_front(ix, st, ctx->scratch[4].as<uint8_t>(), ctx->scratch[5].as<uint64_t>(), seq_offs, query_offs, 0, nq, seq_mode, qp));
        bool lower = false;
        CID_TRY(check_err_flags(ctx, st, ctx->case_aware ? nullptr : &lower));
        if (!lower) break;
        ctx->case_aware = true;          // lower-case k-mers of a raw-case query: the slots carry their case masks
    }
    std::vector<int64_t> filt(nq);
    std::vector<uint64_t> sv(nq, UINT64_MAX);
    for (uint64_t q = 0; q < nq; q++) {          // batch_search_pe.rs:34-39 / :112-120
        if (seq_mode == CID_SEQ_FASTA && gene_search) filt[q] = 0;
        else if (filter < 0) CID_TRY(region_auto_cutoff(ctx, st, (const void*)(qp.d_table + qp.gr.off[q]), qp.gr.mask[q] + 1, &filt[q], false, &sv[q]));
        else filt[q] = filter;
        if (cutoff_used) cutoff_used[q] = filt[q];
    }
    const void* dense = nullptr;
    uint64_t total = 0;
    CID_TRY(compact_survivors(ctx, st, qp, filt, sv, true, &dense, &total));
    CID_CUDA(cudaStreamSynchronize(st));
    for (uint64_t q = 0; q < nq; q++) surv[q] = sv[q];
    *d_slots = const_cast<void*>(dense);
    return CID_OK;
}

int cid_query_slots_counts_dev(cid_index* ix, const void* d_slots, const uint64_t* surv, uint64_t nq, uint32_t* d_counts,
                               uint64_t* d_num_kmers, uint8_t* d_pc, uint32_t* d_col, void* stream) {
    cid_ctx* ctx = ix->ctx;
    cudaStream_t st = (cudaStream_t)stream;
    if (!surv || !d_counts || !d_num_kmers) { set_error("cid_query_slots_counts_dev: null argument"); return CID_E_INVALID; }
    if (ix->m) { set_error("An index with minimizers (.mxi) is used, but not available for this function"); return CID_E_UNSUPPORTED; }
    CID_CUDA(cudaSetDevice(ctx->device));
    const uint32_t N = ix->N;
    CID_CUDA(cudaMemsetAsync(d_counts, 0, nq * N * 4, st));
    CID_CUDA(cudaMemsetAsync(d_num_kmers, 0, nq * 8, st));
    QueryUnits du;
    uint64_t total = 0;
    for (uint64_t q = 0; q < nq; q++) {
        for (uint64_t at = 0; at < surv[q]; at += QUERY_ITEM_SLOTS) {
            du.group.push_back((uint32_t)q);
            du.slot0.push_back(total + at);
            du.nslots.push_back((uint32_t)std::min<uint64_t>(QUERY_ITEM_SLOTS, surv[q] - at));
        }
        total += surv[q];
    }
    if (total == 0) return CID_OK;
    if (!d_slots) { set_error("cid_query_slots_counts_dev: null slot list"); return CID_E_INVALID; }
    QueryPlan qp;
    CID_TRY(upload_units(ctx, st, du, qp));
    CID_TRY(launch_query_counts(ctx, st, ix, d_slots, qp.d_unit_group, qp.d_unit_slot0, qp.d_unit_nslots, du.group.size(), total, nullptr,
                                d_counts, (unsigned long long*)d_num_kmers, false, nullptr, 0, nullptr));
    if (d_pc && d_col) CID_TRY(launch_slots_popcount(ctx, st, ix, d_slots, total, d_pc, d_col));
    return check_err_flags(ctx, st);
}

int cid_query_slots_uniq_dev(cid_index* ix, const void* d_slots, const uint64_t* surv, uint64_t nq, const uint8_t* d_pc_local,
                             const uint8_t* d_pc_sum, const uint32_t* d_col, uint64_t* uniq_n, uint64_t* uniq_sum,
                             uint64_t* uniq_mode, void* stream) {
    cid_ctx* ctx = ix->ctx;
    cudaStream_t st = (cudaStream_t)stream;
    if (!surv || !d_pc_local || !d_pc_sum || !d_col) { set_error("cid_query_slots_uniq_dev: null argument"); return CID_E_INVALID; }
    CID_CUDA(cudaSetDevice(ctx->device));
    const uint32_t N = ix->N;
    if (uniq_n) memset(uniq_n, 0, nq * N * 8);
    if (uniq_sum) memset(uniq_sum, 0, nq * N * 8);
    if (uniq_mode) memset(uniq_mode, 0, nq * N * 8);
    std::vector<uint64_t> prefix(nq + 1, 0);
    for (uint64_t q = 0; q < nq; q++) prefix[q + 1] = prefix[q] + surv[q];
    const uint64_t total = prefix[nq];
    if (total == 0) return CID_OK;
    if (total > 0xFFFFFFF0u / 3) { set_error("cid_query_slots_uniq_dev: too many k-mers in one call"); return CID_E_CAPACITY; }
    CID_TRY(ctx->scratch[7].ensure((nq + 1) * 8));
    CID_TRY(ctx->scratch[12].ensure((size_t)total * 12 + 16));
    CID_TRY(ctx->scratch[13].ensure(16));
    CID_CUDA(cudaMemcpyAsync(ctx->scratch[7].p, prefix.data(), (nq + 1) * 8, cudaMemcpyHostToDevice, st));
    CID_CUDA(cudaMemsetAsync(ctx->scratch[13].p, 0, 16, st));
    CID_TRY(launch_uniq_emit(ctx, st, d_slots, ctx->scratch[7].as<uint64_t>(), (uint32_t)nq, total, d_pc_local, d_pc_sum, d_col,
                             ctx->scratch[12].as<uint32_t>(), (uint32_t)total, ctx->scratch[13].as<uint32_t>()));
    uint32_t nu = 0;
    CID_CUDA(cudaMemcpyAsync(&nu, ctx->scratch[13].p, 4, cudaMemcpyDeviceToHost, st));
    CID_CUDA(cudaStreamSynchronize(st));
    return summarize_uniq(ctx, st, ctx->scratch[12].as<uint32_t>(), nu, N, nq, 0, uniq_n, uniq_sum, uniq_mode);
}

int cid_query_counts_dev(cid_index* ix, const char* d_bases, const uint64_t* d_seq_offs, uint64_t nseq,
                         uint64_t nbases, const uint64_t* d_query_offs, const uint64_t* h_query_offs,
                         const uint64_t* h_seq_offs, uint64_t nq, int seq_mode, uint32_t* d_counts,
                         uint64_t* d_num_kmers, void* stream) {
    (void)nseq; (void)nbases;
    cid_ctx* ctx = ix->ctx;
    cudaStream_t st = (cudaStream_t)stream;
    if (ix->m) { set_error("An index with minimizers (.mxi) is used, but not available for this function"); return CID_E_UNSUPPORTED; }   // main.rs:569-573
    CID_CUDA(cudaSetDevice(ctx->device));
    const uint32_t N = ix->N;
    if (d_counts) CID_CUDA(cudaMemsetAsync(d_counts, 0, nq * N * 4, st));
    else if (!ctx->gather_out) { set_error("cid_query_counts_dev: null counts"); return CID_E_INVALID; }
    CID_CUDA(cudaMemsetAsync(d_num_kmers, 0, nq * 8, st));
    // column-sharded call: the destination rows of a batch start at its first query
    const GatherOut* const shared_out = ctx->gather_out;
    GatherOut batch_out{};
    auto aim = [&](uint64_t q0) {
        if (!shared_out) return;
        batch_out = *shared_out;
        for (uint32_t d = 0; d < batch_out.n; d++) batch_out.base[d] += q0 * (uint64_t)batch_out.stride;
        ctx->gather_out = &batch_out;
    };
    int rc = CID_OK;
    if (nq && d_query_offs && query_front_ok(ix, false, h_seq_offs, h_query_offs, nq)) {
        std::vector<uint64_t> fcuts = query_front_batches(h_seq_offs, h_query_offs, nq);
        for (size_t b = 0; b + 1 < fcuts.size() && rc == CID_OK; b++) {
            aim(fcuts[b]);
            rc = launch_query_front_gather(ctx, st, ix, (const uint8_t*)d_bases, d_seq_offs, d_query_offs, h_seq_offs, h_query_offs,
                                           fcuts[b], fcuts[b + 1], seq_mode, d_counts ? d_counts + fcuts[b] * N : nullptr,
                                           (unsigned long long*)d_num_kmers + fcuts[b]);
        }
        ctx->gather_out = shared_out;
        return rc;
    }
    std::vector<uint64_t> cuts = query_batches(h_seq_offs, h_query_offs, nq, ix->k, kMaxBatchSlots);
    for (size_t b = 0; b + 1 < cuts.size() && rc == CID_OK; b++) {
        const uint64_t q0 = cuts[b], q1 = cuts[b + 1];
        QueryPlan qp;
        rc = query_front(ix, st, (const uint8_t*)d_bases, d_seq_offs, h_seq_offs, h_query_offs, q0, q1, seq_mode, qp);
        if (rc != CID_OK) break;
        aim(q0);
        rc = launch_query_counts(ctx, st, ix, qp.d_table, qp.d_unit_group, qp.d_unit_slot0, qp.d_unit_nslots,
                                 qp.qu.group.size(), qp.gr.total_slots, nullptr, d_counts ? d_counts + q0 * N : nullptr,
                                 (unsigned long long*)d_num_kmers + q0, false, nullptr, 0, nullptr);
    }
    ctx->gather_out = shared_out;
    return rc;
}

// ---- column-sharded search with the count exchange fused into the gather kernel -----------------------------
int cid_dev_alloc(cid_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) { set_error("cid_dev_alloc: null argument"); return CID_E_INVALID; }
    CID_CUDA(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(out, std::max<size_t>(bytes, 256));
    if (e != cudaSuccess) { *out = nullptr; set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); return CID_E_NOMEM; }
    return CID_OK;
}
void cid_dev_free(cid_ctx* ctx, void* p) { if (ctx && p) { cudaSetDevice(ctx->device); cudaFree(p); } }
int cid_ipc_export(cid_ctx* ctx, const void* dptr, uint8_t handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    if (!ctx || !dptr || !handle) { set_error("cid_ipc_export: null argument"); return CID_E_INVALID; }
    CID_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CID_CUDA(cudaIpcGetMemHandle(&h, (void*)dptr));
    memcpy(handle, &h, 64);
    return CID_OK;
}
int cid_ipc_open(cid_ctx* ctx, const uint8_t handle[64], void** out) {
    if (!ctx || !handle || !out) { set_error("cid_ipc_open: null argument"); return CID_E_INVALID; }
    CID_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CID_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return CID_OK;
}
int cid_ipc_close(cid_ctx* ctx, void* p) {
    if (!ctx || !p) return CID_OK;
    CID_CUDA(cudaSetDevice(ctx->device));
    CID_CUDA(cudaIpcCloseMemHandle(p));
    return CID_OK;
}

int cid_query_counts_sharded_dev(cid_index* ix, const char* d_bases, const uint64_t* d_seq_offs, uint64_t nseq, uint64_t nbases,
                                 const uint64_t* d_query_offs, const uint64_t* h_query_offs, const uint64_t* h_seq_offs, uint64_t nq,
                                 int seq_mode, void* const* d_dest_counts, uint32_t n_dest, uint32_t n_total, uint32_t col_offset,
                                 uint64_t* d_num_kmers, void* stream) {
    if (!ix || !d_dest_counts || n_dest < 1 || n_dest > 8) { set_error("cid_query_counts_sharded_dev: 1..8 destination buffers"); return CID_E_INVALID; }
    if ((uint64_t)col_offset + ix->N > n_total) { set_error("cid_query_counts_sharded_dev: shard columns exceed n_total"); return CID_E_INVALID; }
    if (!(ix->Wp >= 4 && ix->Wp <= 512 && (ix->H == 2 || ix->H == 4)) || ix->ctx->opt_query_fused) {
        set_error("cid_query_counts_sharded_dev: needs the streaming gather (16-byte aligned rows of <= 2 KB, num_hash 2 or 4)");
        return CID_E_UNSUPPORTED;
    }
    GatherOut go{};
    for (uint32_t d = 0; d < n_dest; d++) go.base[d] = (uint32_t*)d_dest_counts[d];
    go.n = n_dest; go.stride = n_total; go.col0 = col_offset;
    cid_ctx* ctx = ix->ctx;
    ctx->gather_out = &go;
    // the destinations are zeroed by their owners (and a barrier passed) before this call: skip the local memset
    const int rc = cid_query_counts_dev(ix, d_bases, d_seq_offs, nseq, nbases, d_query_offs, h_query_offs, h_seq_offs, nq, seq_mode,
                                        nullptr, d_num_kmers, stream);
    ctx->gather_out = nullptr;
    return rc;
}

static int query_perfect_impl(cid_index* ix, const char* bases, const uint64_t* seq_offs, uint64_t nseq,
                              const uint64_t* query_offs, uint64_t nq, int seq_mode, uint32_t* and_rows, uint8_t* status,
                              uint64_t* n_kmers);

int cid_query_perfect(cid_index* ix, const char* bases, const uint64_t* seq_offs, uint64_t nseq,
                      const uint64_t* query_offs, uint64_t nq, uint32_t* and_rows, uint8_t* status, uint64_t* n_kmers) {
    return query_perfect_impl(ix, bases, seq_offs, nseq, query_offs, nq, CID_SEQ_FASTA, and_rows, status, n_kmers);
}

int cid_query_perfect_mf(cid_index* ix, const char* bases, const uint64_t* seq_offs, uint64_t nseq, uint32_t* and_rows,
                         uint8_t* status, uint64_t* n_kmers) {
    std::vector<uint64_t> qo(nseq + 1);          // one query per record
    for (uint64_t i = 0; i <= nseq; i++) qo[i] = i;
    return query_perfect_impl(ix, bases, seq_offs, nseq, qo.data(), nseq, CID_SEQ_STRING, and_rows, status, n_kmers);
}

static int query_perfect_impl(cid_index* ix, const char* bases, const uint64_t* seq_offs, uint64_t nseq,
                              const uint64_t* query_offs, uint64_t nq, int seq_mode, uint32_t* and_rows, uint8_t* status,
                              uint64_t* n_kmers) {
    cid_ctx* ctx = ix->ctx;
    cudaStream_t st = ctx->stream;
    if (!seq_offs || !query_offs || !and_rows || !status || !n_kmers) { set_error("cid_query_perfect: null argument"); return CID_E_INVALID; }
    if (ix->m) { set_error("An index with minimizers (.mxi) is used, but not available for this function"); return CID_E_UNSUPPORTED; }   // main.rs:569-573
    CID_CUDA(cudaSetDevice(ctx->device));
    const uint64_t nbases = seq_offs[nseq];
    CID_TRY(ctx->scratch[4].ensure(nbases + 64));
    CID_TRY(ctx->scratch[5].ensure((nseq + 1) * 8));
    if (nbases) CID_CUDA(cudaMemcpyAsync(ctx->scratch[4].p, bases, nbases, cudaMemcpyHostToDevice, st));
    CID_CUDA(cudaMemcpyAsync(ctx->scratch[5].p, seq_offs, (nseq + 1) * 8, cudaMemcpyHostToDevice, st));
    const uint32_t W = ix->W;
    // small queries: shared-memory dedup front end + streaming AND gather (no count table)
    // (any row shape: rows the streaming gather does not take go through perfect_rids_kernel)
    const bool fast = nq && ctx->opt_query_front && !ctx->opt_query_fused && ix->Wp * 4ull <= 48 * 1024 &&
                      query_front_fits(seq_offs, query_offs, 0, nq, ix->k);
    if (fast) {
        CID_TRY(ctx->scratch[7].ensure((nq + 1) * 8));
        CID_CUDA(cudaMemcpyAsync(ctx->scratch[7].p, query_offs, (nq + 1) * 8, cudaMemcpyHostToDevice, st));
    }
    std::vector<uint64_t> cuts = fast ? query_front_batches(seq_offs, query_offs, nq)
                                      : query_batches(seq_offs, query_offs, nq, ix->k, kMaxBatchSlots);
    for (size_t b = 0; b + 1 < cuts.size(); b++) {
        const uint64_t q0 = cuts[b], q1 = cuts[b + 1], bq = q1 - q0;
        CID_TRY(ctx->scratch[10].ensure(bq * W * 4));
        CID_TRY(ctx->scratch[11].ensure(bq * 8));
        CID_TRY(ctx->scratch[13].ensure(bq * 4));
        CID_CUDA(cudaMemsetAsync(ctx->scratch[10].p, 0xFF, bq * W * 4, st));
        CID_CUDA(cudaMemsetAsync(ctx->scratch[11].p, 0, bq * 8, st));
        CID_CUDA(cudaMemsetAsync(ctx->scratch[13].p, 0, bq * 4, st));
        if (fast) {
            CID_TRY(launch_query_front_gather(ctx, st, ix, ctx->scratch[4].as<uint8_t>(), ctx->scratch[5].as<uint64_t>(),
                                              ctx->scratch[7].as<uint64_t>(), seq_offs, query_offs, q0, q1, seq_mode, nullptr,
                                              ctx->scratch[11].as<unsigned long long>(), ctx->scratch[10].as<uint32_t>(),
                                              ctx->scratch[13].as<uint32_t>()));
            CID_TRY(check_err_flags(ctx, st));
        } else {
            QueryPlan qp;
            CID_TRY(query_front(ix, st, ctx->scratch[4].as<uint8_t>(), ctx->scratch[5].as<uint64_t>(), seq_offs, query_offs,
                                q0, q1, seq_mode, qp));
            CID_TRY(check_err_flags(ctx, st));
            CID_TRY(launch_query_perfect(ctx, st, ix, qp.d_table, qp.d_unit_group, qp.d_unit_slot0, qp.d_unit_nslots,
                                         qp.qu.group.size(), ctx->scratch[10].as<uint32_t>(), ctx->scratch[13].as<uint32_t>(),
                                         ctx->scratch[11].as<unsigned long long>()));
        }
        std::vector<uint32_t> missing(bq);
        CID_CUDA(cudaMemcpyAsync(and_rows + q0 * W, ctx->scratch[10].p, bq * W * 4, cudaMemcpyDeviceToHost, st));
        CID_CUDA(cudaMemcpyAsync(n_kmers + q0, ctx->scratch[11].p, bq * 8, cudaMemcpyDeviceToHost, st));
        CID_CUDA(cudaMemcpyAsync(missing.data(), ctx->scratch[13].p, bq * 4, cudaMemcpyDeviceToHost, st));
        CID_CUDA(cudaStreamSynchronize(st));
        for (uint64_t q = 0; q < bq; q++) {
            uint8_t s = n_kmers[q0 + q] == 0 ? 2 : (missing[q] ? 1 : 0);
            status[q0 + q] = s;
            if (s != 0) memset(and_rows + (q0 + q) * W, 0, (size_t)W * 4);
        }
    }
    return CID_OK;
}

int cid_hash_kmers(cid_index* ix, const char* kmers, uint64_t n, uint64_t* row_ids) {
    cid_ctx* ctx = ix->ctx;
    cudaStream_t st = ctx->stream;
    CID_CUDA(cudaSetDevice(ctx->device));
    if (n == 0) return CID_OK;
    const uint32_t len = ix->m ? ix->m : ix->k;     // the hashed item of an .mxi index is the minimizer
    CID_TRY(ctx->scratch[4].ensure(n * len + 64));
    CID_TRY(ctx->scratch[10].ensure(n * ix->H * 8));
    CID_CUDA(cudaMemcpyAsync(ctx->scratch[4].p, kmers, n * len, cudaMemcpyHostToDevice, st));
    CID_TRY(launch_hash_kmers(ctx, st, ix, ctx->scratch[4].as<uint8_t>(), n, ctx->scratch[10].as<uint64_t>()));
    CID_CUDA(cudaMemcpyAsync(row_ids, ctx->scratch[10].p, n * ix->H * 8, cudaMemcpyDeviceToHost, st));
    CID_CUDA(cudaStreamSynchronize(st));
    return CID_OK;
}

}  // extern "C"
