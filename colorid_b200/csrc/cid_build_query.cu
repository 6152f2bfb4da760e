// CUDA kernels (sm_100a) for the build and search halves of the BIGSI hot path.
//
//   kmerize_insert   kmer.rs:87-125 / 461-510 / 581-655  canonical k-mer count map (device hash table)
//   region_histogram kmer.rs:866-884                     count histogram feeding auto_cutoff
//   region_to_bloom  build.rs:62-67 + simple_bloom.rs:19-26  clean_map + Bloom insert (atomicOr)
//   transpose_bitsets build.rs:116-128                   per-colour bitsets -> row-major signature matrix
//   query_counts     batch_search_pe.rs:45-84            row gather, AND, per-accession hit counts
//   query_perfect    perfect_search.rs:26-46             AND of all rows of all k-mers
//
// All of these are HBM/L2-bound integer work; no tensor cores are involved.
#include <algorithm>
#include <cstring>

#include "cid_device.cuh"
#include "cid_internal.h"

namespace cid {

// ================================================================= kmerize_insert
// One CTA = one tile of KT consecutive k-mer start positions of the concatenated batch (plus a
// k-1 halo), so reads (150 bp) and contigs (Mbp) are handled by the same code.
constexpr int KT = 1024;              // start positions per tile
constexpr int KT_CAP = KT + 64;       // bytes staged per tile (halo <= 31, rounded to 32)
constexpr int KT_THREADS = 256;
constexpr int KT_MAXSTARTS = KT_CAP + 8;

// SETONLY (builds that keep every k-mer: FASTA without -f, or -f 0): a key-only set instead of the count table, and the
// thread that inserts a NEW key hashes it into the accession's Bloom bitset at once (simple_bloom.rs:19-26) -- no count
// table scan, no region_to_bloom launch; err[1] is then n_ref_kmers.
struct SetSink {
    unsigned long long* keys; uint64_t mask;     // one set for the whole launch
    uint32_t* bitset; uint32_t H; ModS mods;
    uint32_t bloom_m;                            // != 0: insert find_minimizer(k-mer, bloom_m) (build_single_mini)
    uint32_t packed;                             // !SETONLY: `keys` is a packed count table (key << 22 | count) of `mask`+1 words
};
template <bool MINI, bool SETONLY>   // MINI: the item is each k-mer's minimizer of length mini_m (build_multi_mini)
__global__ void __launch_bounds__(KT_THREADS)
kmerize_insert_kernel(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ seq_offs, uint64_t nseq,
                      uint64_t base_lo, uint64_t nbases, const uint32_t* __restrict__ seq_group, const uint64_t* __restrict__ region_off,
                      const uint64_t* __restrict__ region_mask, Slot* __restrict__ table, uint32_t k, uint32_t mini_m,
                      int seq_mode, uint32_t* __restrict__ err, SetSink sink, uint32_t cs_on) {
    __shared__ __align__(16) uint8_t smem[tile_smem_bytes(KT_CAP)];
    __shared__ uint32_t lut[SETONLY ? 256 : 1];
    if (SETONLY) lut4_init(lut, threadIdx.x, KT_THREADS);      // made visible by the barriers below
    __shared__ uint32_t s_starts[KT_MAXSTARTS];
    __shared__ uint64_t s_s0;
    __shared__ uint32_t s_nstarts, s_fresh;

    const int tid = threadIdx.x;
    uint32_t fresh = 0;
    const uint64_t tile_start = base_lo + (uint64_t)blockIdx.x * KT;   // [base_lo, nbases) is the slice of `bases` covered by seq_offs[0..nseq]
    if (tile_start >= nbases) return;
    const int tile_len = (int)min((uint64_t)(KT + k - 1), nbases - tile_start);

    Tile t = tile_carve(smem, KT_CAP);
    t.len = tile_len;
    for (int i = tid; i < KT_CAP / 32 + 2; i += KT_THREADS) t.start[i] = 0;
    // stage raw bytes
    const uint8_t* src = bases + tile_start;
    if ((((uintptr_t)src) & 3) == 0) {
        const uint32_t* s4 = (const uint32_t*)src;
        uint32_t* d4 = (uint32_t*)t.ascii;
        int n4 = tile_len >> 2;
        for (int i = tid; i < n4; i += KT_THREADS) d4[i] = __ldg(s4 + i);
        for (int i = (n4 << 2) + tid; i < tile_len; i += KT_THREADS) t.ascii[i] = __ldg(src + i);
    } else {
        for (int i = tid; i < tile_len; i += KT_THREADS) t.ascii[i] = __ldg(src + i);
    }
    // sequence owning the first base: largest s with seq_offs[s] <= tile_start
    if (tid == 0) {
        uint64_t lo = 0, hi = nseq;   // invariant: seq_offs[lo] <= tile_start < seq_offs[hi]
        while (hi - lo > 1) {
            uint64_t mid = (lo + hi) >> 1;
            if (__ldg(seq_offs + mid) <= tile_start) lo = mid; else hi = mid;
        }
        s_s0 = lo;
        s_nstarts = 0;
        s_fresh = 0;
    }
    __syncthreads();
    const uint64_t s0 = s_s0;
    // sequence starts strictly inside the tile, in order (list index j <-> sequence s0+1+j)
    for (uint64_t base = 0;; base += KT_THREADS) {
        uint64_t s = s0 + 1 + base + tid;
        bool in = false;
        if (s < nseq) {
            uint64_t o = __ldg(seq_offs + s);
            if (o < tile_start + (uint64_t)tile_len) {
                in = true;
                uint32_t rel = (uint32_t)(o - tile_start);
                if (base + tid < KT_MAXSTARTS) s_starts[base + tid] = rel; else atomicOr(err, ERRF_STARTS_OVERFLOW);
                atomicOr(&t.start[rel >> 5], 1u << (rel & 31));
                atomicAdd(&s_nstarts, 1u);
            }
        }
        if (!__syncthreads_or(in)) break;
    }
    tile_pack(t, KT_CAP, tid, KT_THREADS);
    __syncthreads();
    const uint32_t nstarts = min(s_nstarts, (uint32_t)KT_MAXSTARTS);

    for (int p = tid; p < KT; p += KT_THREADS) {
        uint64_t key; bool fwd, low;
        if (!tile_kmer(t, p, k, key, fwd, low)) {
            // kmerize_string (kmer.rs:279-293) has no has_no_n test: a window that lies inside one record but
            // holds a byte outside ACGTacgt is still a k-mer there, and it cannot be packed into 2 bits
            if (seq_mode == CID_SEQ_STRING && p + (int)k <= t.len && (mask_window(t.start, p, k) & ~1u) == 0u)
                atomicOr(err, ERRF_STRING_NONACGT);
            continue;
        }
        uint32_t cs = 0;
        if (low && seq_mode == CID_SEQ_FASTQ) {
            // kmer.rs:461-510,581-655 keep the case: the k-mer is (codes, case mask), which only the 16-byte count table of a
            // case-aware pass can hold; any other pass reports it and the caller redoes the work that way
            if (!cs_on || MINI || SETONLY || sink.packed) { atomicOr(err, ERRF_LOWER_RAW); continue; }
            const uint32_t lw = mask_window(t.lower, p, k);
            cs = fwd ? lw : (__brev(lw) >> (32 - k));
        }
        if (MINI) {     // build_multi_mini: the counted item is the k-mer's minimizer (kmer.rs:328-361, 694-824)
            uint32_t mpos; bool mfwd;
            key = tile_minimizer(t, p, k, mini_m, key, fwd, low, mpos, mfwd);
        }
        // owner sequence = s0 + #starts <= p
        uint32_t lo = 0, hi = (SETONLY || sink.packed) ? 0u : nstarts;
        while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (s_starts[mid] <= (uint32_t)p) lo = mid + 1; else hi = mid; }
        int rc;
        if (SETONLY) {
            rc = set_insert(sink.keys, sink.mask, key);
            if (rc > 0) {
                uint64_t item = key;
                uint32_t len = MINI ? mini_m : k;
                if (sink.bloom_m) { uint32_t which; item = minimizer_packed(key, revcomp_key(key, k), k, sink.bloom_m, which); len = sink.bloom_m; }
                const HashIn in = hashin_from_key(lut, item, len);
                for (uint32_t h = 0; h < sink.H; h++) {
                    const uint64_t bit = hash_row(in, len, h, sink.mods);
                    atomicOr(&sink.bitset[bit >> 5], 1u << (bit & 31));
                }
            }
        } else if (sink.packed) {
            bool ovf = false;
            rc = packed_insert(sink.keys, sink.mask, key, ovf);
            if (ovf) atomicOr(err, ERRF_COUNT_OVERFLOW);
        } else {
            uint64_t owner = s0 + lo;
            uint32_t g = seq_group ? __ldg(seq_group + owner) : 0u;
            rc = cs_on ? table_insert_cs(table + __ldg(region_off + g), __ldg(region_mask + g), key, cs)
                       : table_insert(table + __ldg(region_off + g), __ldg(region_mask + g), key);
        }
        if (rc < 0) atomicOr(err, ERRF_TABLE_FULL);
        fresh += rc > 0;
    }
    // distinct k-mers inserted by this launch (err[1]): lets the build size count tables optimistically
    fresh = __reduce_add_sync(0xffffffffu, fresh);
    if ((tid & 31) == 0 && fresh) atomicAdd(&s_fresh, fresh);
    __syncthreads();
    if (tid == 0 && s_fresh) atomicAdd(err + 1, s_fresh);
}

int launch_kmerize_insert(cid_ctx* ctx, cudaStream_t st, const uint8_t* d_bases, const uint64_t* d_seq_offs,
                          uint64_t nseq, uint64_t base_lo, uint64_t base_hi, const uint32_t* d_seq_group,
                          const uint64_t* d_region_off, const uint64_t* d_region_mask, void* d_table, uint32_t k,
                          int seq_mode, uint32_t mini_m, uint64_t packed_slots) {
    if (base_hi <= base_lo || nseq == 0) return CID_OK;
    uint64_t ntiles = (base_hi - base_lo + KT - 1) / KT;
    ProfScope ps(ctx, st, KID_KMERIZE_INSERT);
    SetSink none{};
    if (packed_slots) { none.keys = (unsigned long long*)d_table; none.mask = packed_slots - 1; none.packed = 1; }
    if (mini_m)
        kmerize_insert_kernel<true, false><<<(unsigned)ntiles, KT_THREADS, 0, st>>>(d_bases, d_seq_offs, nseq, base_lo, base_hi, d_seq_group,
                                                                                   d_region_off, d_region_mask, (Slot*)d_table, k, mini_m,
                                                                                   seq_mode, ctx->d_err, none, 0u);
    else
        kmerize_insert_kernel<false, false><<<(unsigned)ntiles, KT_THREADS, 0, st>>>(d_bases, d_seq_offs, nseq, base_lo, base_hi, d_seq_group,
                                                                                    d_region_off, d_region_mask, (Slot*)d_table, k, 0,
                                                                                    seq_mode, ctx->d_err, none, (ctx->case_aware && !packed_slots) ? 1u : 0u);
    ctx->launches++;
    CID_CUDA(cudaGetLastError());
    return CID_OK;
}

// Set-only build of one accession: every new canonical k-mer (or minimizer) goes straight into `d_bitset`.
// count_m != 0: the set holds minimizers (build_multi_mini); bloom_m != 0: the set holds k-mers, their minimizers are inserted.
int launch_kmerize_bloom(cid_ctx* ctx, cudaStream_t st, const uint8_t* d_bases, const uint64_t* d_seq_offs, uint64_t nseq,
                         uint64_t nbases, void* d_keys, uint64_t nslots, uint32_t k, int seq_mode, uint32_t count_m, uint32_t bloom_m,
                         uint32_t H, const ModS& mods, uint32_t* d_bitset) {
    if (nbases == 0 || nseq == 0) return CID_OK;
    const uint64_t ntiles = (nbases + KT - 1) / KT;
    SetSink sink{(unsigned long long*)d_keys, nslots - 1, d_bitset, H, mods, bloom_m, 0};
    ProfScope ps(ctx, st, KID_KMERIZE_INSERT);
    if (count_m)
        kmerize_insert_kernel<true, true><<<(unsigned)ntiles, KT_THREADS, 0, st>>>(d_bases, d_seq_offs, nseq, 0, nbases, nullptr, nullptr, nullptr,
                                                                                  nullptr, k, count_m, seq_mode, ctx->d_err, sink, 0u);
    else
        kmerize_insert_kernel<false, true><<<(unsigned)ntiles, KT_THREADS, 0, st>>>(d_bases, d_seq_offs, nseq, 0, nbases, nullptr, nullptr, nullptr,
                                                                                   nullptr, k, 0, seq_mode, ctx->d_err, sink, 0u);
    ctx->launches++;
    CID_CUDA(cudaGetLastError());
    return CID_OK;
}

void plan_regions(const uint64_t* h_seq_offs, const uint64_t* h_group_offs, uint64_t ngroups, uint32_t k,
                  GroupRegions& gr) {
    gr.ngroups = ngroups;
    gr.off.resize(ngroups);
    gr.mask.resize(ngroups);
    uint64_t total = 0;
    for (uint64_t g = 0; g < ngroups; g++) {
        uint64_t nb = h_seq_offs[h_group_offs[g + 1]] - h_seq_offs[h_group_offs[g]];
        uint64_t npos = nb >= k ? nb - k + 1 : 0;          // upper bound on k-mer occurrences
        uint64_t slots = next_pow2(std::max<uint64_t>(64, 2 * npos));
        gr.off[g] = total;
        gr.mask[g] = slots - 1;
        total += slots;
    }
    gr.total_slots = total;
}

// ================================================================= region_histogram
constexpr int HIST_SMEM_BINS = 1024;
template <bool PACKED>
__global__ void __launch_bounds__(256)
region_histogram_kernel(const void* __restrict__ region, uint64_t nslots, uint32_t* __restrict__ hist,
                        uint32_t hist_bins, uint32_t* __restrict__ overflow, uint32_t overflow_cap,
                        uint32_t* __restrict__ overflow_n) {
    __shared__ uint32_t sh[HIST_SMEM_BINS];
    for (int i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += (uint64_t)gridDim.x * blockDim.x) {
        const SlotView v = slot_read<PACKED>(region, s);
        if (!v.used) continue;
        uint32_t c = v.count;
        if (c < HIST_SMEM_BINS) atomicAdd(&sh[c], 1u);
        else if (c < hist_bins) atomicAdd(&hist[c], 1u);
        else {
            uint32_t i = atomicAdd(overflow_n, 1u);
            if (i < overflow_cap) overflow[i] = c;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x) if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
int launch_region_histogram(cid_ctx* ctx, cudaStream_t st, const void* d_region, uint64_t nslots, uint32_t* d_hist,
                            uint32_t hist_bins, uint32_t* d_overflow, uint32_t overflow_cap, uint32_t* d_overflow_n, bool packed) {
    unsigned grid = (unsigned)std::min<uint64_t>((nslots + 255) / 256, (uint64_t)ctx->sm_count * 16);
    if (grid == 0) grid = 1;
    ProfScope ps(ctx, st, KID_HISTOGRAM);
    if (packed) region_histogram_kernel<true><<<grid, 256, 0, st>>>(d_region, nslots, d_hist, hist_bins, d_overflow, overflow_cap, d_overflow_n);
    else region_histogram_kernel<false><<<grid, 256, 0, st>>>(d_region, nslots, d_hist, hist_bins, d_overflow, overflow_cap, d_overflow_n);
    ctx->launches++;
    CID_CUDA(cudaGetLastError());
    return CID_OK;
}

// ================================================================= region_to_bloom
// clean_map (count > cutoff, kmer.rs:826-837), n_ref_kmers (build.rs:62), BloomFilter::insert.
template <bool PACKED>
__global__ void __launch_bounds__(256, 6)
region_to_bloom_kernel(const void* __restrict__ region, uint64_t nslots, long long cutoff, uint32_t k, uint32_t mini_m,
                       uint32_t H, ModS mods, uint32_t* __restrict__ bitset, unsigned long long* __restrict__ nref) {
    __shared__ uint32_t lut[256];
    lut4_init(lut, threadIdx.x, blockDim.x);
    __syncthreads();
    uint32_t mine = 0;
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += (uint64_t)gridDim.x * blockDim.x) {
        const SlotView v = slot_read<PACKED>(region, s);
        if (!v.used || (long long)v.count <= cutoff) continue;
        mine++;
        uint64_t item = v.key;
        uint32_t len = k;
        if (mini_m) {   // build_single_mini (build.rs:430-432,445-447,461-463): insert find_minimizer(kmer, m) of every kept k-mer
            uint32_t which;
            item = minimizer_packed(v.key, revcomp_key(v.key, k), k, mini_m, which);
            len = mini_m;
        }
        HashIn in = hashin_from_key(lut, item, len);
        if (!mini_m) hashin_apply_case(in, v.cs, len);      // (a minimizer index never runs a case-aware pass)
        for (uint32_t i = 0; i < H; i++) {
            uint64_t bit = hash_row(in, len, i, mods);
            atomicOr(&bitset[bit >> 5], 1u << (bit & 31));
        }
    }
    // block reduction of the survivor count
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(nref, (unsigned long long)mine);
}
int launch_region_to_bloom(cid_ctx* ctx, cudaStream_t st, const void* d_region, uint64_t nslots, int64_t cutoff,
                           uint32_t k, uint32_t mini_m, uint32_t H, const ModS& mods, uint32_t* d_bitset, unsigned long long* d_nref,
                           bool packed) {
    unsigned grid = (unsigned)std::min<uint64_t>((nslots + 255) / 256, (uint64_t)ctx->sm_count * 16);
    if (grid == 0) grid = 1;
    ProfScope ps(ctx, st, KID_TO_BLOOM);
    if (packed) region_to_bloom_kernel<true><<<grid, 256, 0, st>>>(d_region, nslots, (long long)cutoff, k, mini_m, H, mods, d_bitset, d_nref);
    else region_to_bloom_kernel<false><<<grid, 256, 0, st>>>(d_region, nslots, (long long)cutoff, k, mini_m, H, mods, d_bitset, d_nref);
    ctx->launches++;
    CID_CUDA(cudaGetLastError());
    return CID_OK;
}

// ================================================================= transpose_bitsets
// bitsets[colour][word j] (bit r of word j = Bloom bit 32j+r) -> rows[32j+r][colour/32] bit colour%32  (build.rs:116-128).
// One CTA: 256 rows x up to 32 word-columns (1,024 colours).  Loads: 32 contiguous bytes (8 bitset words = 256 rows) per
// colour, whole sectors, eight passes in flight; 32x32 bit blocks are transposed in registers with a five-step shuffle butterfly
// (25 instructions per block instead of 32 ballots + selects); the transposed words go through shared memory so that every ROW leaves as one
// contiguous run -- 128 bytes per row for >= 1,024 colours, the whole row for narrower indexes -- instead of one 4-byte
// store per lane at a stride of a whole row (r1: 8 % of the HBM copy peak).
constexpr int TR_ROWS = 256;
constexpr int TR_WCOLS = 32;
// lane r holds row r (bit c = column c); returns column `lane` (bit r = row r).  Five butterfly steps; per step a lane swaps
// the half of its word that belongs to its partner lane^j: the partner's word rotated by +-j (one funnel shift; the bits that
// wrap around fall where the lane keeps its own) merged under the lane's keep mask (one LOP3).  Masks and rotations depend
// only on the lane: TrLane is set up once per thread.
struct TrLane { uint32_t keep[5], rot[5]; };
__device__ __forceinline__ TrLane transpose32_setup(int lane) {
    TrLane t;
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int j = 16 >> s;
        const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
        t.keep[s] = (lane & j) ? ~m : m;
        t.rot[s] = (lane & j) ? 32 - j : j;
    }
    return t;
}
__device__ __forceinline__ uint32_t transpose32(uint32_t x, const TrLane& t) {
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, 16 >> s);
        const uint32_t z = __funnelshift_l(y, y, t.rot[s]);
        x = (x & t.keep[s]) | (z & ~t.keep[s]);
    }
    return x;
}
__global__ void __launch_bounds__(256)
transpose_bitsets_kernel(const uint32_t* __restrict__ bitsets, uint64_t bs_words, uint32_t N, uint64_t S,
                         uint32_t* __restrict__ rows, uint32_t Wp, uint32_t W) {
    extern __shared__ __align__(16) uint8_t dsm[];
    uint32_t* sin = (uint32_t*)dsm;                      // [wcol][colour in group][bitset word 0..7], row stride 9: no bank conflicts
    uint32_t* sout = sin + TR_WCOLS * 32 * 9;            // [row in tile][wcol], row stride 33
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t word0 = (uint64_t)blockIdx.x * (TR_ROWS / 32);       // first bitset word of this row tile
    const uint32_t wc0 = blockIdx.y * TR_WCOLS;
    const uint32_t nwc = min((uint32_t)TR_WCOLS, W - wc0);
    // load: 8 lanes read the 8 words (one 32-byte sector) of one colour, 32 colours per 256-thread pass; the loads of 8 passes
    // are all in flight before the first one is stored (a pass at a time is a DRAM round trip per pass: 3x slower)
    const uint32_t j = tid & 7;
    for (uint32_t p0 = 0; p0 < nwc * 32; p0 += 32 * 8) {
        uint32_t v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t pair = p0 + 32 * u + (tid >> 3), colour = wc0 * 32 + pair;
            v[u] = 0;
            if (pair < nwc * 32 && colour < N && word0 + j < bs_words) v[u] = __ldg(bitsets + (uint64_t)colour * bs_words + word0 + j);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t pair = p0 + 32 * u + (tid >> 3);
            if (pair < nwc * 32) sin[pair * 9 + j] = v[u];
        }
    }
    __syncthreads();
    // transpose: warp w takes bitset word w of every word-column; lane = colour within the group on input, row on output
    const TrLane tl = transpose32_setup(lane);
#pragma unroll 4
    for (uint32_t wc = 0; wc < nwc; wc++)
        sout[(warp * 32 + lane) * 33 + wc] = transpose32(sin[(wc * 32 + lane) * 9 + warp], tl);
    __syncthreads();
    // store: consecutive threads write consecutive words of a row (and, for narrow indexes, of consecutive rows)
    const uint64_t row0 = word0 * 32;
    if (nwc == TR_WCOLS) {
        for (uint32_t e = tid; e < TR_ROWS * TR_WCOLS; e += 256) {
            const uint32_t r = e >> 5, w = e & 31;
            if (row0 + r < S) rows[(row0 + r) * Wp + wc0 + w] = sout[r * 33 + w];
        }
    } else {
        for (uint32_t e = tid; e < TR_ROWS * nwc; e += 256) {
            const uint32_t r = e / nwc, w = e % nwc;
            if (row0 + r < S) rows[(row0 + r) * Wp + wc0 + w] = sout[r * 33 + w];
        }
    }
}
int launch_transpose(cid_ctx* ctx, cudaStream_t st, const cid_index* idx) {
    dim3 grid((unsigned)((idx->S + TR_ROWS - 1) / TR_ROWS), (idx->W + TR_WCOLS - 1) / TR_WCOLS);
    const size_t tsmem = (size_t)(TR_WCOLS * 32 * 9 + TR_ROWS * 33) * 4;
    bool& tattr = ctx->attr_done[6];
    if (!tattr) { CID_CUDA(cudaFuncSetAttribute(transpose_bitsets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem)); tattr = true; }
    ProfScope ps(ctx, st, KID_TRANSPOSE);
    transpose_bitsets_kernel<<<grid, 256, tsmem, st>>>(idx->bitsets, idx->bs_words, idx->N, idx->S, idx->rows, idx->Wp,
                                                   idx->W);
    ctx->launches++;
    CID_CUDA(cudaGetLastError());
    return CID_OK;
}

// ================================================================= rownz bitmap
__global__ void __launch_bounds__(256)
rownz_kernel(const uint32_t* __restrict__ rows, uint64_t S, uint32_t Wp, uint32_t* __restrict__ rownz) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t any = 0;
    if (r < S) {
        const uint32_t* p = rows + r * Wp;
        for (uint32_t w = 0; w < Wp; w++) any |= __ldg(p + w);
    }
    uint32_t b = __ballot_sync(0xffffffffu, any != 0);
    if ((threadIdx.x & 31) == 0 && (r >> 5) < (S + 31) / 32) rownz[r >> 5] = b;
}
int launch_rownz(cid_ctx* ctx, cudaStream_t st, const cid_index* idx) {
    uint64_t nthreads = (idx->S + 31) / 32 * 32;
    ProfScope ps(ctx, st, KID_ROWNZ);
    rownz_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(idx->rows, idx->S, idx->Wp, idx->rownz);
    ctx->launches++;
    CID_CUDA(cudaGetLastError());
    return CID_OK;
}

// ================================================================= query work list
void plan_units(const GroupRegions& gr, uint32_t chunk, QueryUnits& qu) {
    qu.group.clear(); qu.slot0.clear(); qu.nslots.clear();
    for (uint64_t g = 0; g < gr.ngroups; g++) {
        uint64_t slots = gr.mask[g] + 1;
        for (uint64_t s = 0; s < slots; s += chunk) {
            qu.group.push_back((uint32_t)g);
            qu.slot0.push_back(gr.off[g] + s);
            qu.nslots.push_back((uint32_t)std::min<uint64_t>(chunk, slots - s));
        }
    }
}

// Shared front half of the gather kernels: compact the surviving (key,count) of a slot chunk into
// shared memory, then hash every survivor to its H row indices.
struct UnitSmem {
    uint32_t* lut;        // [256]
    uint64_t* keys;       // [QUERY_CHUNK]
    uint32_t* mult;       // [QUERY_CHUNK]
    uint32_t* cs;         // [QUERY_CHUNK] case masks of raw-case k-mers (0 everywhere else)
    uint32_t* rowid;      // [QUERY_CHUNK * H]
    uint32_t* cnt;        // [32 * 32 * 4] per-accession counters of one column block (32 lanes x VEC<=4 words)
    uint32_t* n;          // [4] list length, missing flag
};
__host__ __device__ inline size_t unit_smem_bytes(uint32_t H) {
    return 256 * 4 + (size_t)QUERY_CHUNK * 8 + (size_t)QUERY_CHUNK * 8 + (size_t)QUERY_CHUNK * H * 4 + 4096 * 4 + 16;
}
__device__ __forceinline__ UnitSmem unit_carve(uint8_t* base, uint32_t H) {
    UnitSmem u;
    u.keys = (uint64_t*)base;
    u.mult = (uint32_t*)(u.keys + QUERY_CHUNK);
    u.cs = u.mult + QUERY_CHUNK;
    u.rowid = u.cs + QUERY_CHUNK;
    u.cnt = u.rowid + (size_t)QUERY_CHUNK * H;
    u.lut = u.cnt + 4096;
    u.n = u.lut + 256;
    return u;
}
__device__ __forceinline__ uint32_t unit_collect_and_hash(const UnitSmem& u, const Slot* __restrict__ table,
                                                          uint64_t slot0, uint32_t nslots, long long filter, uint32_t k,
                                                          uint32_t H, const ModS& mods) {
    const int tid = threadIdx.x, nt = blockDim.x;
    lut4_init(u.lut, tid, nt);
    if (tid == 0) { u.n[0] = 0; u.n[1] = 0; }
    __syncthreads();
    for (uint32_t s = tid; s < nslots; s += nt) {
        Slot v = table[slot0 + s];
        if (v.key != CID_EMPTY_KEY && (long long)v.count > filter) {
            uint32_t i = atomicAdd(&u.n[0], 1u);
            u.keys[i] = v.key;
            u.mult[i] = v.count;
            u.cs[i] = slot_cs(v.pad);
        }
    }
    __syncthreads();
    const uint32_t n = u.n[0];
    for (uint32_t i = tid; i < n; i += nt) {
        HashIn in = hashin_from_key(u.lut, u.keys[i], k);
        hashin_apply_case(in, u.cs[i], k);
        for (uint32_t h = 0; h < H; h++) u.rowid[i * H + h] = (uint32_t)hash_row(in, k, h, mods);
    }
    __syncthreads();
    return n;
}

// ================================================================= query_counts  (the row-gather kernel)
// For every surviving k-mer: AND of its H rows, +1 for every set accession (batch_search_pe.rs:60-74).
// A work item is up to QUERY_ITEM_SLOTS table slots of one query, walked in QUERY_CHUNK pieces
// (collect -> hash -> gather).  A lane owns VEC consecutive 32-accession words of the row (128-bit
// loads when VEC == 4: a 128-byte row is one request of 8 lanes, so a warp gathers 4 k-mers per
// instruction).  Hits are accumulated in bit-sliced counters that live in registers for the whole
// item: eight k-mers at a time go through a Harley-Seal carry-save tree (ones/twos/fours planes),
// and only the weight-8 carry ripples into the upper planes, so counting costs < 4 logic ops per
// k-mer and word instead of one atomic per set bit.
constexpr int QC_PLANES = 12;   // a lane group sees at most QUERY_ITEM_SLOTS/8 = 2048 k-mers per item; counts <= 2^11
constexpr int QC_TREE = 8;      // k-mers per lane group and carry-save tree step

template <int VEC> struct RowVec;
template <> struct RowVec<1> { uint32_t w[1]; __device__ __forceinline__ void load(const uint32_t* p) { w[0] = __ldg(p); } };
template <> struct RowVec<2> { uint32_t w[2]; __device__ __forceinline__ void load(const uint32_t* p) { uint2 v = __ldg((const uint2*)p); w[0] = v.x; w[1] = v.y; } };
template <> struct RowVec<4> { uint32_t w[4]; __device__ __forceinline__ void load(const uint32_t* p) { uint4 v = __ldg((const uint4*)p); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; } };

// carry-save adder: (carry, sum) = a + b + c, bitwise
__device__ __forceinline__ void csa(uint32_t& carry, uint32_t& sum, uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t u = a ^ b;
    carry = (a & b) | (u & c);
    sum = u ^ c;
}

template <int VEC, bool UNIQ, int HT>   // HT: compile-time num_hash (0 = run-time)
__global__ void __launch_bounds__(256, 2)
query_counts_kernel(const uint32_t* __restrict__ rows, uint32_t Wp, uint32_t W, uint32_t N, uint32_t k, uint32_t H,
                    ModS mods, const Slot* __restrict__ table, const uint32_t* __restrict__ unit_group,
                    const uint64_t* __restrict__ unit_slot0, const uint32_t* __restrict__ unit_nslots,
                    const long long* __restrict__ filter, uint32_t* __restrict__ counts,
                    unsigned long long* __restrict__ num_kmers, uint32_t* __restrict__ uniq_list, uint32_t uniq_cap,
                    uint32_t* __restrict__ uniq_n) {
    extern __shared__ __align__(16) uint8_t dsm[];
    UnitSmem u = unit_carve(dsm, H);
    const uint32_t g = unit_group[blockIdx.x];
    const uint64_t item_slot0 = unit_slot0[blockIdx.x];
    const uint32_t item_nslots = unit_nslots[blockIdx.x];
    const long long filt = filter ? filter[g] : 0ll;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const uint32_t vpr = Wp / VEC;                      // vectors per row (Wp is a multiple of VEC)
    // lanes per k-mer: the unique-hit variant reduces over the group with xor-shuffles and needs a power of
    // two; otherwise any divisor layout works (a 160-byte row takes 10 lanes, 3 k-mers per warp pass)
    uint32_t lpk = 1;
    if (UNIQ) { while (lpk < vpr && lpk < 32) lpk <<= 1; }
    else lpk = vpr < 32 ? vpr : 32;
    const uint32_t kpw = 32 / lpk;                      // k-mers per warp pass
    const uint32_t sub = lane / lpk, colv = lane % lpk;
    const bool lane_on = sub < kpw;                     // lanes left over by a non-power-of-two layout idle
    const uint32_t ncb = (vpr + 31) / 32;               // column blocks of 32 vectors
    const uint32_t hh = HT ? (uint32_t)HT : H;
    const uint32_t stride = 8 * kpw;                    // k-mers per block pass
    uint32_t total_n = 0;
    for (uint32_t cb = 0; cb < ncb; cb++) {
        const uint32_t vcol = cb * 32 + colv;
        const bool colok = lane_on && vcol < vpr;
        uint32_t pl[VEC][QC_PLANES];
#pragma unroll
        for (int v = 0; v < VEC; v++)
#pragma unroll
            for (int p = 0; p < QC_PLANES; p++) pl[v][p] = 0;
        uint32_t seen = 0;
        const uint32_t* colbase = rows + (size_t)vcol * VEC;
        for (uint32_t c0 = 0; c0 < item_nslots; c0 += QUERY_CHUNK) {
            __syncthreads();   // previous chunk's lists are dead
            const uint32_t n = unit_collect_and_hash(u, table, item_slot0 + c0, min((uint32_t)QUERY_CHUNK, item_nslots - c0),
                                                     filt, k, H, mods);
            if (cb == 0) total_n += n;
            // QC_TREE k-mers per lane group and iteration, in parts of QC_PART k-mers whose row loads (<= 8 per
            // lane) are all issued before any is consumed
            constexpr int QC_PART = (HT == 0 || HT > 2) ? 2 : 4;
            for (uint32_t i0 = warp * kpw + sub; i0 < n + sub; i0 += stride * QC_TREE) {
                uint32_t x[QC_TREE][VEC];
#pragma unroll
                for (int half = 0; half < QC_TREE / QC_PART; half++) {
                    RowVec<VEC> r[QC_PART][HT ? HT : 1];
#pragma unroll
                    for (int un = 0; un < QC_PART; un++) {
                        const uint32_t i = i0 + (half * QC_PART + un) * stride;
                        const bool on = i < n && colok;
                        if (HT) {
                            const uint32_t* rid = u.rowid + (on ? i : 0u) * HT;
#pragma unroll
                            for (int h = 0; h < (HT ? HT : 1); h++) {
                                if (on) r[un][h].load(colbase + (size_t)rid[h] * Wp);
                                else {
#pragma unroll
                                    for (int v = 0; v < VEC; v++) r[un][h].w[v] = 0;
                                }
                            }
                        } else {
#pragma unroll
                            for (int v = 0; v < VEC; v++) r[un][0].w[v] = 0;
                            if (on) {
                                const uint32_t* rid = u.rowid + i * hh;
                                r[un][0].load(colbase + (size_t)rid[0] * Wp);
                                for (uint32_t h = 1; h < hh; h++) {
                                    RowVec<VEC> t;
                                    t.load(colbase + (size_t)rid[h] * Wp);
#pragma unroll
                                    for (int v = 0; v < VEC; v++) r[un][0].w[v] &= t.w[v];
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int un = 0; un < QC_PART; un++)
#pragma unroll
                        for (int v = 0; v < VEC; v++) {
                            uint32_t a = r[un][0].w[v];
#pragma unroll
                            for (int h = 1; h < (HT ? HT : 1); h++) a &= r[un][h].w[v];
                            x[half * QC_PART + un][v] = a;
                        }
                }
                if (UNIQ) {
#pragma unroll
                    for (int un = 0; un < QC_TREE; un++) {
                        // exactly one accession hit over the whole row (batch_search_pe.rs:75-82)
                        const uint32_t i = i0 + un * stride;
                        uint32_t pc = 0;
#pragma unroll
                        for (int v = 0; v < VEC; v++) pc += __popc(x[un][v]);
                        const uint32_t mine = pc;
                        for (uint32_t o = lpk >> 1; o > 0; o >>= 1) pc += __shfl_xor_sync(0xffffffffu, pc, o);
                        if (ncb == 1 && pc == 1 && mine == 1) {
                            uint32_t colour = 0;
#pragma unroll
                            for (int v = 0; v < VEC; v++) if (x[un][v]) colour = (vcol * VEC + v) * 32 + (__ffs(x[un][v]) - 1);
                            uint32_t e = atomicAdd(uniq_n, 1u);
                            if (e < uniq_cap) {
                                uniq_list[3 * (uint64_t)e] = g;
                                uniq_list[3 * (uint64_t)e + 1] = colour;
                                uniq_list[3 * (uint64_t)e + 2] = u.mult[i];
                            }
                        }
                    }
                }
                // Harley-Seal: eight inputs -> ones/twos/fours planes + one weight-8 carry per word
#pragma unroll
                for (int v = 0; v < VEC; v++) {
                    uint32_t t2a, t2b, t4a, t4b, c8;
                    csa(t2a, pl[v][0], pl[v][0], x[0][v], x[1][v]);
                    csa(t2b, pl[v][0], pl[v][0], x[2][v], x[3][v]);
                    csa(t4a, pl[v][1], pl[v][1], t2a, t2b);
                    csa(t2a, pl[v][0], pl[v][0], x[4][v], x[5][v]);
                    csa(t2b, pl[v][0], pl[v][0], x[6][v], x[7][v]);
                    csa(t4b, pl[v][1], pl[v][1], t2a, t2b);
                    csa(c8, pl[v][2], pl[v][2], t4a, t4b);
#pragma unroll
                    for (int p = 3; p < QC_PLANES; p++) {
                        const uint32_t t = pl[v][p] & c8;
                        pl[v][p] ^= c8;
                        c8 = t;
                    }
                }
                seen += QC_TREE;
            }
        }
        // flush: planes -> shared counters of this column block (the 8 warps and 32/lpk sub-groups hold
        // partial counts of the same columns) -> one global atomic per non-zero accession.
        __syncthreads();
        for (int i = tid; i < 1024 * VEC; i += 256) u.cnt[i] = 0;
        __syncthreads();
        if (colok && seen) {
            const int depth = 32 - __clz(seen);          // planes that can be non-zero
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                const uint32_t cbase = (colv * VEC + v) * 32;
#pragma unroll
                for (int nb = 0; nb < 8; nb++) {
                    // four accessions at a time: spread each plane's nibble into four byte lanes
                    uint32_t lo = 0, hi = 0;
#pragma unroll
                    for (int p = 0; p < QC_PLANES; p++) {
                        if (p >= depth) break;
                        const uint32_t sp = (((pl[v][p] >> (4 * nb)) & 0xFu) * 0x00204081u) & 0x01010101u;
                        if (p < 8) lo += sp << p; else hi += sp << (p - 8);
                    }
                    if (lo | hi) {
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const uint32_t val = ((lo >> (8 * q)) & 0xFFu) | (((hi >> (8 * q)) & 0xFFu) << 8);
                            if (val) atomicAdd(&u.cnt[cbase + 4 * nb + q], val);
                        }
                    }
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < 1024 * VEC; i += 256) {
            const uint32_t c = cb * 1024 * VEC + i;
            const uint32_t val = u.cnt[i];
            if (val && c < N) atomicAdd(&counts[(uint64_t)g * N + c], val);
        }
    }
    if (tid == 0 && total_n) atomicAdd(&num_kmers[g], (unsigned long long)total_n);
}

// ================================================================= query_hash + query_gather (streaming row gather)
// The fast path for 16-byte-aligned rows without unique-hit summaries (-g gene search, the column
// shards of C5).  query_hash compacts and hashes a work item's k-mers once and leaves their row
// indices in HBM (H x 4 bytes per k-mer, ~3 % of the row bytes they address); query_gather is then
// a pure streaming kernel with no block-level phase changes.  Row reads are cp.async (LDGSTS) 16-byte
// copies into a warp-private ring in shared memory -- every lane copies exactly the bytes it later
// consumes, so no cross-lane synchronisation is needed -- which keeps 64 KB of row reads in flight
// per CTA without holding registers, enough to cover DRAM latency at random-row access rates.
__global__ void __launch_bounds__(256)
query_hash_kernel(const Slot* __restrict__ table, const uint32_t* __restrict__ unit_group,
                  const uint64_t* __restrict__ unit_slot0, const uint32_t* __restrict__ unit_nslots,
                  const long long* __restrict__ filter, uint32_t k, uint32_t H, ModS mods, uint32_t* __restrict__ rid_out,
                  uint32_t* __restrict__ unit_n, unsigned long long* __restrict__ num_kmers) {
    __shared__ uint32_t lut[256];
    __shared__ unsigned long long keys[QUERY_CHUNK];
    __shared__ uint32_t kcs[QUERY_CHUNK];
    __shared__ uint32_t s_n;
    const int tid = threadIdx.x;
    lut4_init(lut, tid, 256);
    const uint32_t g = unit_group[blockIdx.x];
    const uint64_t slot0 = unit_slot0[blockIdx.x];
    const uint32_t nslots = unit_nslots[blockIdx.x];
    const long long filt = filter ? filter[g] : 0ll;
    uint32_t total = 0;
    for (uint32_t c0 = 0; c0 < nslots; c0 += QUERY_CHUNK) {
        __syncthreads();
        if (tid == 0) s_n = 0;
        __syncthreads();
        const uint32_t cn = min((uint32_t)QUERY_CHUNK, nslots - c0);
        for (uint32_t s = tid; s < cn; s += 256) {
            const Slot v = table[slot0 + c0 + s];
            if (v.key != CID_EMPTY_KEY && (long long)v.count > filt) {
                const uint32_t at = atomicAdd(&s_n, 1u);
                keys[at] = v.key; kcs[at] = slot_cs(v.pad);
            }
        }
        __syncthreads();
        const uint32_t n = s_n;
        // survivors so far <= slots scanned so far, so the compact list never leaves the item's slot range
        for (uint32_t i = tid; i < n; i += 256) {
            HashIn in = hashin_from_key(lut, keys[i], k);
            hashin_apply_case(in, kcs[i], k);
            uint32_t* out = rid_out + (slot0 + total + i) * H;
            for (uint32_t h = 0; h < H; h++) out[h] = (uint32_t)hash_row(in, k, h, mods);
        }
        total += n;
    }
    if (tid == 0) {
        unit_n[blockIdx.x] = total;
        if (total) atomicAdd(&num_kmers[g], (unsigned long long)total);
    }
}

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gptr) : "memory");
}
// Same copy with the 64-byte L2 prefetch size (option gather_l2_64b).  Tried for the 160-byte rows of a C5 column shard, which
// always straddle two 128-byte lines; measured no difference on B200 (profiles/r1_gather_l2_64b.txt), so it is off by default.
__device__ __forceinline__ void cp_async16_pf64(uint32_t saddr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPEND> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}

// 4 warps per work item: the per-thread cost of flushing the bit-sliced counters is fixed, so fewer warps per
// item amortise it over more gather steps, and 4 CTAs per SM keep three gathering while one flushes.
constexpr int QG_WARPS = 4;
constexpr int QG_RING_BYTES = 32 * 1024;            // per CTA (8 KB per warp)
constexpr int QG_PLANES = 13;                       // a lane group sees <= QUERY_ITEM_SLOTS / QG_WARPS = 4096 k-mers
template <int HT> struct QGCfg {
    static constexpr int STAGE = HT * 512;          // bytes per warp and stage: 32 lanes x 16 B x HT rows
    static constexpr int D = QG_RING_BYTES / QG_WARPS / STAGE;   // stages per warp: 8 (H=2), 4 (H=4)
};

// ANDM (perfect search, perfect_search.rs:26-46): instead of counting, every lane ANDs the rows of its column slice; a row
// index whose row is absent (row-present bitmap, checked where the indices are loaded) raises `missing`; `counts` is then
// and_rows[group][W] (pre-set to all ones by the caller).
template <int HT, bool ANDM>
__global__ void __launch_bounds__(QG_WARPS * 32, 4)
query_gather_kernel(const uint32_t* __restrict__ rows, uint32_t Wp, uint32_t N, const uint32_t* __restrict__ rid,
                    const uint32_t* __restrict__ unit_group, const uint64_t* __restrict__ unit_slot0,
                    const uint32_t* __restrict__ unit_n, uint32_t* __restrict__ counts,
                    const uint32_t* __restrict__ rownz, uint32_t* __restrict__ missing, uint32_t W, GatherOut go, uint32_t pf64) {
    using Cfg = QGCfg<HT>;
    constexpr int D = Cfg::D;
    static_assert(D >= 2 && 8 % D == 0, "ring depth must divide the tree width");
    extern __shared__ __align__(16) uint8_t dsm[];
    uint32_t* cnt = (uint32_t*)(dsm + QG_RING_BYTES);                 // [4096] per-accession counters
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t n = unit_n[blockIdx.x];
    const uint32_t g = unit_group[blockIdx.x];
    if (n == 0) {
        if (!ANDM && go.dense)                            // a query without k-mers still owns its row slice: zeros
            for (uint32_t c = tid; c < N; c += QG_WARPS * 32)
#pragma unroll
                for (uint32_t d = 0; d < 8; d++) if (d < go.n) go.base[d][(uint64_t)g * go.stride + go.col0 + c] = 0u;
        return;
    }
    const uint32_t* myrid = rid + unit_slot0[blockIdx.x] * HT;

    const uint32_t vpr = Wp >> 2;                         // 16-byte vectors per row (<= 32)
    const uint32_t kpw = 32 / vpr;                        // k-mers per warp step
    const uint32_t sub = lane / vpr, colv = lane % vpr;
    const bool lane_on = sub < kpw;
    const uint32_t IT = 32 / kpw;                         // steps served by one coalesced load of row indices
    const uint32_t BK = kpw * IT;                         // k-mers per such load (<= 32)
    // contiguous k-mer range of this warp, a multiple of kpw long
    const uint32_t per = ((n + QG_WARPS * kpw - 1) / (QG_WARPS * kpw)) * kpw;
    const uint32_t lo = min(n, warp * per), hi = min(n, lo + per);
    const uint32_t T = (hi - lo + kpw - 1) / kpw;         // steps of this warp
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(dsm) + warp * (D * Cfg::STAGE) + lane * 16;
    const char* colbase = (const char*)(rows + colv * 4);
    const uint32_t rowbytes = Wp * 4;
    // steps in which this lane's k-mer (lo + kpw*t + sub) exists
    const uint32_t myT = (lane_on && lo + sub < hi) ? (hi - lo - sub + kpw - 1) / kpw : 0u;

    uint32_t pl[4][ANDM ? 1 : QG_PLANES];
#pragma unroll
    for (int v = 0; v < 4; v++)
#pragma unroll
        for (int p = 0; p < (ANDM ? 1 : QG_PLANES); p++) pl[v][p] = ANDM ? 0xFFFFFFFFu : 0u;      // ANDM: pl[v][0] is the AND accumulator
    bool miss = false;

    // row indices of k-mer lo + BK*b + lane (batch b), double-buffered in registers
    uint32_t cur[HT], nxt[HT];
    auto load_batch = [&](uint32_t b, uint32_t* dst) {
        const uint32_t i = lo + b * BK + lane;
        const bool ok = (uint32_t)lane < BK && i < hi;
#pragma unroll
        for (int h = 0; h < HT; h++) dst[h] = 0;
        if (ok) {
            if (HT == 2) { const uint2 v = __ldg((const uint2*)(myrid + (size_t)i * 2)); dst[0] = v.x; dst[1] = v.y; }
            else if (HT == 4) { const uint4 v = __ldg((const uint4*)(myrid + (size_t)i * 4)); dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w; }
            else {
#pragma unroll
                for (int h = 0; h < HT; h++) dst[h] = __ldg(myrid + (size_t)i * HT + h);
            }
            if (ANDM) {
#pragma unroll
                for (int h = 0; h < HT; h++) if (!((__ldg(rownz + (dst[h] >> 5)) >> (dst[h] & 31)) & 1u)) miss = true;
            }
        }
    };
    load_batch(0, cur);
    load_batch(1, nxt);
    uint32_t t_issue = 0, it_issue = 0, b_issue = 0;      // next step to issue, its position in its batch, its batch
    auto issue = [&](int slot) {
        if (t_issue < T) {
            if (it_issue == IT) {
                it_issue = 0; b_issue++;
#pragma unroll
                for (int h = 0; h < HT; h++) cur[h] = nxt[h];
                load_batch(b_issue + 1, nxt);
            }
            const uint32_t src = min(kpw * it_issue + sub, 31u);
            const bool valid = t_issue < myT;
#pragma unroll
            for (int h = 0; h < HT; h++) {
                const uint32_t r = __shfl_sync(0xffffffffu, cur[h], src);
                if (valid) {
                    if (pf64) cp_async16_pf64(ring + slot * Cfg::STAGE + h * 512, colbase + (uint64_t)r * rowbytes);
                    else cp_async16(ring + slot * Cfg::STAGE + h * 512, colbase + (uint64_t)r * rowbytes);
                }
            }
            t_issue++; it_issue++;
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < D - 1; s++) issue(s);
    for (uint32_t t0 = 0; t0 < T; t0 += 8) {
        uint32_t x[8][4];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            issue((u + D - 1) % D);
            cp_async_wait<D - 1>();
            const bool valid = t0 + u < myT;
            uint4 a = make_uint4(0, 0, 0, 0);
            if (valid) {
                a = lds128(ring + (u % D) * Cfg::STAGE);
#pragma unroll
                for (int h = 1; h < HT; h++) {
                    const uint4 b = lds128(ring + (u % D) * Cfg::STAGE + h * 512);
                    a.x &= b.x; a.y &= b.y; a.z &= b.z; a.w &= b.w;
                }
            }
            x[u][0] = a.x; x[u][1] = a.y; x[u][2] = a.z; x[u][3] = a.w;
            if (ANDM && valid) { pl[0][0] &= a.x; pl[1][0] &= a.y; pl[2][0] &= a.z; pl[3][0] &= a.w; }
        }
        // Harley-Seal: eight inputs -> ones/twos/fours planes + one weight-8 carry per word
#pragma unroll
        for (int v = 0; v < (ANDM ? 0 : 4); v++) {
            uint32_t t2a, t2b, t4a, t4b, c8;
            csa(t2a, pl[v][0], pl[v][0], x[0][v], x[1][v]);
            csa(t2b, pl[v][0], pl[v][0], x[2][v], x[3][v]);
            csa(t4a, pl[v][1], pl[v][1], t2a, t2b);
            csa(t2a, pl[v][0], pl[v][0], x[4][v], x[5][v]);
            csa(t2b, pl[v][0], pl[v][0], x[6][v], x[7][v]);
            csa(t4b, pl[v][1], pl[v][1], t2a, t2b);
            csa(c8, pl[v][2], pl[v][2], t4a, t4b);
#pragma unroll
            for (int p = 3; p < QG_PLANES; p++) {
                const uint32_t t = pl[v][p] & c8;
                pl[v][p] ^= c8;
                c8 = t;
            }
        }
    }
    cp_async_wait<0>();
    if (ANDM) {
        // AND across the lane groups and warps that hold the same column slice, then into and_rows[g][W]
        __syncthreads();
        for (int i = tid; i < (int)Wp; i += QG_WARPS * 32) cnt[i] = 0xFFFFFFFFu;
        __syncthreads();
        if (lane_on && T) {
#pragma unroll
            for (int v = 0; v < 4; v++) if (pl[v][0] != 0xFFFFFFFFu) atomicAnd(&cnt[colv * 4 + v], pl[v][0]);
        }
        if (__syncthreads_or(miss) && tid == 0) atomicOr(&missing[g], 1u);
        for (int i = tid; i < (int)W; i += QG_WARPS * 32) if (cnt[i] != 0xFFFFFFFFu) atomicAnd(&counts[(uint64_t)g * W + i], cnt[i]);
        return;
    }
    // flush.  The same column word is held by QG_WARPS * kpw lane groups (partial counts over disjoint k-mers).  The
    // planes go to shared memory (the ring is free now) as 16-byte stores, laid out [plane][partial][column word]; then
    // ONE thread per column word adds the partials bit-sliced (a full adder per plane: 32 accessions per instruction)
    // and only the final sums are unpacked, one global atomic per non-zero accession.
    __syncthreads();
    uint32_t* buf = (uint32_t*)dsm;
    const uint32_t P = QG_WARPS * kpw;
    const uint32_t Tmax = (per + kpw - 1) / kpw;                       // steps of warp 0, the longest
    const int depth = min(QG_PLANES, 32 - __clz(((Tmax + 7) & ~7u)));  // planes that can be non-zero in any lane
    if (lane_on) {
        const uint32_t pi = warp * kpw + sub;
#pragma unroll
        for (int p = 0; p < QG_PLANES; p++) {
            if (p >= depth) break;
            *(uint4*)(buf + ((size_t)(p * P + pi) * Wp + colv * 4)) = make_uint4(pl[0][p], pl[1][p], pl[2][p], pl[3][p]);
        }
    }
    __syncthreads();
    if ((uint32_t)tid < Wp) {
        constexpr int SP = QG_PLANES + 2;                              // a unit holds <= 16384 k-mers: 15 planes
        uint32_t acc[SP];
#pragma unroll
        for (int p = 0; p < SP; p++) acc[p] = 0;
        for (uint32_t pi = 0; pi < P; pi++) {
            uint32_t carry = 0;
#pragma unroll
            for (int p = 0; p < SP; p++) {
                const uint32_t x = p < depth ? buf[(size_t)(p * P + pi) * Wp + tid] : 0u;
                csa(carry, acc[p], acc[p], x, carry);
            }
        }
        uint32_t nz = 0;
#pragma unroll
        for (int p = 0; p < SP; p++) nz |= acc[p];
        // go.n destinations: the caller's buffer, or in column-sharded mode the same slot [query][col0 + accession] of every
        // GPU's full-width result (peer memory over NVLink): the count exchange rides on the kernel's own stores
        const uint64_t off = (uint64_t)g * go.stride + go.col0 + (uint32_t)tid * 32;
        if (go.dense) {
            // four counters at a time (nibble spread: each plane's 4 bits into 4 byte lanes), one 16-byte store per destination
            const bool vec_ok = ((go.stride | go.col0) & 3u) == 0u;
#pragma unroll
            for (int nb = 0; nb < 8; nb++) {
                uint32_t lo8 = 0, hi8 = 0;
#pragma unroll
                for (int p = 0; p < SP; p++) {
                    const uint32_t sp = (((acc[p] >> (4 * nb)) & 0xFu) * 0x00204081u) & 0x01010101u;
                    if (p < 8) lo8 += sp << p; else hi8 += sp << (p - 8);
                }
                uint4 v;
                v.x = (lo8 & 0xFFu) | ((hi8 & 0xFFu) << 8);
                v.y = ((lo8 >> 8) & 0xFFu) | (((hi8 >> 8) & 0xFFu) << 8);
                v.z = ((lo8 >> 16) & 0xFFu) | (((hi8 >> 16) & 0xFFu) << 8);
                v.w = (lo8 >> 24) | ((hi8 >> 24) << 8);
                const uint32_t c0 = (uint32_t)tid * 32 + 4 * nb;
                if (c0 >= N) break;
                if (vec_ok && c0 + 4 <= N) {
#pragma unroll
                    for (uint32_t d = 0; d < 8; d++) if (d < go.n) *(uint4*)(go.base[d] + off + 4 * nb) = v;
                } else {
                    const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (uint32_t j = 0; j < 4; j++)
#pragma unroll
                        for (uint32_t d = 0; d < 8; d++) if (d < go.n && c0 + j < N) go.base[d][off + 4 * nb + j] = vv[j];
                }
            }
            nz = 0;
        }
        while (nz) {
            const uint32_t bb = __ffs(nz) - 1;
            nz &= nz - 1;
            uint32_t val = 0;
#pragma unroll
            for (int p = 0; p < SP; p++) val |= ((acc[p] >> bb) & 1u) << p;
            if ((uint32_t)tid * 32 + bb < N) {
#pragma unroll
                for (uint32_t d = 0; d < 8; d++) if (d < go.n) atomicAdd(go.base[d] + off + bb, val);
            }
        }
    }
}

// ================================================================= query_gather_tma (rows of 516 bytes .. 2 KB)
// Wide rows -- an unsharded index of 4,100..16,384 accessions: the full C5 index has 1,264-byte rows and fits one B200 -- are
// staged by the TMA engine: per k-mer ONE elected lane of a producer warp issues num_hash bulk copies
// (cp.async.bulk.shared.global, whole rows) into a ring of D stages in shared memory and arms the stage's mbarrier with the
// byte count; four consumer warps wait on the barrier's phase, AND the rows (every thread owns one 16-byte vector of the row
// for the whole work item, so its bit-sliced counters never meet another thread's) and release the stage through a second
// mbarrier.  No thread computes an address per 16 bytes and no register holds a row index on the consumer side; with 32 KB of
// rows in flight per CTA the row stream runs at HBM speed.  (Rows of <= 512 bytes keep the cp.async ring of query_gather: a
// bulk copy is a warp-uniform instruction, so several small rows per step would cost MORE instructions than one LDGSTS per
// lane -- TMA pays where one copy moves a kilobyte.)
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

constexpr int QT_CONS_WARPS = 4;                       // 128 threads x 16 bytes: rows of up to 2 KB (16,384 accessions)
constexpr int QT_THREADS = (QT_CONS_WARPS + 1) * 32;   // + the producer warp
constexpr int QT_PLANES = 15;                          // a work item holds <= QUERY_ITEM_SLOTS = 2^14 k-mers
constexpr int QT_RING_BYTES = 32 * 1024;
template <int HT, bool ANDM>
__global__ void __launch_bounds__(QT_THREADS)
query_gather_tma_kernel(const uint32_t* __restrict__ rows, uint32_t Wp, uint32_t N, const uint32_t* __restrict__ rid,
                        const uint32_t* __restrict__ unit_group, const uint64_t* __restrict__ unit_slot0,
                        const uint32_t* __restrict__ unit_n, uint32_t* __restrict__ counts, const uint32_t* __restrict__ rownz,
                        uint32_t* __restrict__ missing, uint32_t W, GatherOut go, uint32_t D) {
    extern __shared__ __align__(16) uint8_t dsm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t n = unit_n[blockIdx.x];
    const uint32_t g = unit_group[blockIdx.x];
    if (n == 0) {
        if (!ANDM && go.dense)
            for (uint32_t c = tid; c < N; c += QT_THREADS)
#pragma unroll
                for (uint32_t d = 0; d < 8; d++) if (d < go.n) go.base[d][(uint64_t)g * go.stride + go.col0 + c] = 0u;
        return;
    }
    const uint32_t rowbytes = Wp * 4, stage = HT * rowbytes;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(dsm);
    const uint32_t full = ring + QT_RING_BYTES, empty = full + 8 * 8;       // up to 8 stages
    if (tid == 0) {
        for (uint32_t s = 0; s < D; s++) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, QT_CONS_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t* myrid = rid + unit_slot0[blockIdx.x] * HT;
    if (warp == QT_CONS_WARPS) {
        // ---- producer warp: row indices 32 k-mers at a time (coalesced), one elected lane issues the bulk copies
        bool miss = false;
        for (uint32_t i0 = 0; i0 < n; i0 += 32) {
            uint32_t r[HT];
#pragma unroll
            for (int h = 0; h < HT; h++) r[h] = 0;
            if (i0 + lane < n) {
#pragma unroll
                for (int h = 0; h < HT; h++) r[h] = __ldg(myrid + (size_t)(i0 + lane) * HT + h);
                if (ANDM) {
#pragma unroll
                    for (int h = 0; h < HT; h++) if (!((__ldg(rownz + (r[h] >> 5)) >> (r[h] & 31)) & 1u)) miss = true;
                }
            }
            const uint32_t cnt = min(32u, n - i0);
            for (uint32_t j = 0; j < cnt; j++) {
                const uint32_t i = i0 + j, s = i % D, use = i / D;
                uint32_t rr[HT];
#pragma unroll
                for (int h = 0; h < HT; h++) rr[h] = __shfl_sync(0xffffffffu, r[h], j);
                if (lane == 0) {
                    if (use) mbar_wait(empty + 8 * s, (use - 1) & 1u);          // the consumers are done with this stage
                    mbar_expect_tx(full + 8 * s, stage);
#pragma unroll
                    for (int h = 0; h < HT; h++) bulk_g2s(ring + s * stage + h * rowbytes, rows + (size_t)rr[h] * Wp, rowbytes, full + 8 * s);
                }
            }
        }
        if (ANDM && __any_sync(0xffffffffu, miss) && lane == 0) atomicOr(&missing[g], 1u);
        return;
    }
    // ---- consumer warps: thread v owns 16-byte vector v of the row
    const uint32_t v = (uint32_t)tid, vpr = Wp >> 2;
    const bool on = v < vpr;
    uint32_t pl[4][ANDM ? 1 : QT_PLANES];
#pragma unroll
    for (int w = 0; w < 4; w++)
#pragma unroll
        for (int p = 0; p < (ANDM ? 1 : QT_PLANES); p++) pl[w][p] = ANDM ? 0xFFFFFFFFu : 0u;
    for (uint32_t i0 = 0; i0 < n; i0 += 8) {
        uint32_t x[8][4];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t i = i0 + u;
            uint4 a = make_uint4(0, 0, 0, 0);
            if (i < n) {                                  // (uniform across the CTA)
                const uint32_t s = i % D;
                mbar_wait(full + 8 * s, (i / D) & 1u);
                if (on) {
                    a = lds128(ring + s * stage + v * 16);
#pragma unroll
                    for (int h = 1; h < HT; h++) {
                        const uint4 b = lds128(ring + s * stage + h * rowbytes + v * 16);
                        a.x &= b.x; a.y &= b.y; a.z &= b.z; a.w &= b.w;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + 8 * s);
                if (ANDM && on) { pl[0][0] &= a.x; pl[1][0] &= a.y; pl[2][0] &= a.z; pl[3][0] &= a.w; }
            }
            x[u][0] = a.x; x[u][1] = a.y; x[u][2] = a.z; x[u][3] = a.w;
        }
#pragma unroll
        for (int w = 0; w < (ANDM ? 0 : 4); w++) {
            uint32_t t2a, t2b, t4a, t4b, c8;
            csa(t2a, pl[w][0], pl[w][0], x[0][w], x[1][w]);
            csa(t2b, pl[w][0], pl[w][0], x[2][w], x[3][w]);
            csa(t4a, pl[w][1], pl[w][1], t2a, t2b);
            csa(t2a, pl[w][0], pl[w][0], x[4][w], x[5][w]);
            csa(t2b, pl[w][0], pl[w][0], x[6][w], x[7][w]);
            csa(t4b, pl[w][1], pl[w][1], t2a, t2b);
            csa(c8, pl[w][2], pl[w][2], t4a, t4b);
#pragma unroll
            for (int p = 3; p < QT_PLANES; p++) {
                const uint32_t t = pl[w][p] & c8;
                pl[w][p] ^= c8;
                c8 = t;
            }
        }
    }
    if (!on) return;
    if (ANDM) {
#pragma unroll
        for (int w = 0; w < 4; w++)
            if (v * 4 + w < W && pl[w][0] != 0xFFFFFFFFu) atomicAnd(&counts[(uint64_t)g * W + v * 4 + w], pl[w][0]);
        return;
    }
    // flush: this thread alone holds the counts of its 128 accessions; four at a time (nibble spread), 16-byte stores
    const bool vec_ok = ((go.stride | go.col0) & 3u) == 0u;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        const uint32_t cbase = (v * 4 + w) * 32;
        if (cbase >= N) break;
        const uint64_t off = (uint64_t)g * go.stride + go.col0 + cbase;
#pragma unroll
        for (int nb = 0; nb < 8; nb++) {
            uint32_t lo8 = 0, hi8 = 0;
#pragma unroll
            for (int p = 0; p < QT_PLANES; p++) {
                const uint32_t sp = (((pl[w][p] >> (4 * nb)) & 0xFu) * 0x00204081u) & 0x01010101u;
                if (p < 8) lo8 += sp << p; else hi8 += sp << (p - 8);
            }
            const uint32_t vv[4] = {(lo8 & 0xFFu) | ((hi8 & 0xFFu) << 8), ((lo8 >> 8) & 0xFFu) | (((hi8 >> 8) & 0xFFu) << 8),
                                    ((lo8 >> 16) & 0xFFu) | (((hi8 >> 16) & 0xFFu) << 8), (lo8 >> 24) | ((hi8 >> 24) << 8)};
            const uint32_t c0 = cbase + 4 * nb;
            if (c0 >= N) break;
            if (go.dense) {
                if (vec_ok && c0 + 4 <= N) {
#pragma unroll
                    for (uint32_t d = 0; d < 8; d++) if (d < go.n) *(uint4*)(go.base[d] + off + 4 * nb) = make_uint4(vv[0], vv[1], vv[2], vv[3]);
                } else {
#pragma unroll
                    for (uint32_t j = 0; j < 4; j++)
#pragma unroll
                        for (uint32_t d = 0; d < 8; d++) if (d < go.n && c0 + j < N) go.base[d][off + 4 * nb + j] = vv[j];
                }
            } else {
#pragma unroll
                for (uint32_t j = 0; j < 4; j++)
                    if (vv[j] && c0 + j < N) {
#pragma unroll
                        for (uint32_t d = 0; d < 8; d++) if (d < go.n) atomicAdd(go.base[d] + off + 4 * nb + j, vv[j]);
                    }
            }
        }
    }
}

// Wide-row unique-hit pass (Wp > 32): one warp per k-mer sums popcounts over the whole row.
__global__ void __launch_bounds__(256)
query_uniq_wide_kernel(const uint32_t* __restrict__ rows, uint32_t Wp, uint32_t k, uint32_t H, ModS mods,
                       const Slot* __restrict__ table, const uint32_t* __restrict__ unit_group,
                       const uint64_t* __restrict__ unit_slot0, const uint32_t* __restrict__ unit_nslots,
                       const long long* __restrict__ filter, uint32_t* __restrict__ uniq_list, uint32_t uniq_cap,
                       uint32_t* __restrict__ uniq_n) {
    extern __shared__ __align__(16) uint8_t dsm[];
    UnitSmem u = unit_carve(dsm, H);
    const uint32_t g = unit_group[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t item_nslots = unit_nslots[blockIdx.x];
    for (uint32_t c0 = 0; c0 < item_nslots; c0 += QUERY_CHUNK) {
    __syncthreads();
    const uint32_t n = unit_collect_and_hash(u, table, unit_slot0[blockIdx.x] + c0, min((uint32_t)QUERY_CHUNK, item_nslots - c0),
                                             filter ? filter[g] : 0ll, k, H, mods);
    for (uint32_t i = warp; i < n; i += 8) {
        uint32_t pc = 0, where = 0;
        for (uint32_t col = lane; col < Wp; col += 32) {
            uint32_t x = 0xFFFFFFFFu;
            for (uint32_t h = 0; h < H; h++) x &= __ldg(rows + (uint64_t)u.rowid[i * H + h] * Wp + col);
            if (x) { pc += __popc(x); where = col * 32 + (__ffs(x) - 1); }
        }
        uint32_t tot = pc;
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (tot == 1 && pc == 1) {
            uint32_t e = atomicAdd(uniq_n, 1u);
            if (e < uniq_cap) {
                uniq_list[3 * (uint64_t)e] = g;
                uniq_list[3 * (uint64_t)e + 1] = where;
                uniq_list[3 * (uint64_t)e + 2] = u.mult[i];
            }
        }
    }
    }
}

// Perfect search over prepared row-index lists for the row shapes the streaming gather does not take (rows of 1 or 2 words,
// rows above 512 bytes, num_hash other than 2 / 4): one CTA per unit, every thread ANDs the rows of its k-mers word by word
// into a shared accumulator.  Keeps -s -m exact (windows with a byte outside ACGTacgt, hashed by query_front<STRINGM>) on
// narrow indexes and narrow column shards.
__global__ void __launch_bounds__(256)
perfect_rids_kernel(const uint32_t* __restrict__ rows, uint32_t Wp, uint32_t W, uint32_t H, const uint32_t* __restrict__ rid,
                    const uint32_t* __restrict__ unit_group, const uint64_t* __restrict__ unit_slot0,
                    const uint32_t* __restrict__ unit_n, const uint32_t* __restrict__ rownz, uint32_t* __restrict__ and_rows,
                    uint32_t* __restrict__ missing) {
    extern __shared__ __align__(16) uint32_t acc[];       // [Wp]
    const uint32_t n = unit_n[blockIdx.x], g = unit_group[blockIdx.x];
    if (n == 0) return;
    for (uint32_t w = threadIdx.x; w < Wp; w += blockDim.x) acc[w] = 0xFFFFFFFFu;
    __syncthreads();
    const uint32_t* my = rid + unit_slot0[blockIdx.x] * H;
    bool miss = false;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        for (uint32_t h = 0; h < H; h++) {
            const uint32_t r = __ldg(my + (size_t)i * H + h);
            if (!((__ldg(rownz + (r >> 5)) >> (r & 31)) & 1u)) miss = true;
        }
        for (uint32_t w = 0; w < W; w++) {
            uint32_t x = 0xFFFFFFFFu;
            for (uint32_t h = 0; h < H; h++) x &= __ldg(rows + (size_t)__ldg(my + (size_t)i * H + h) * Wp + w);
            if (x != 0xFFFFFFFFu) atomicAnd(&acc[w], x);
        }
    }
    if (__syncthreads_or(miss) && threadIdx.x == 0) atomicOr(&missing[g], 1u);
    for (uint32_t w = threadIdx.x; w < W; w += blockDim.x) if (acc[w] != 0xFFFFFFFFu) atomicAnd(&and_rows[(uint64_t)g * W + w], acc[w]);
}

// query_gather over prepared row-index lists: unit u = rid[(unit_slot0[u] + i) * H + h], i < unit_n[u] (<= 16384)
static int launch_query_gather(cid_ctx* ctx, cudaStream_t st, const cid_index* idx, const uint32_t* d_rid,
                               const uint32_t* d_unit_group, const uint64_t* d_unit_slot0, const uint32_t* d_unit_n,
                               uint64_t nunits, uint32_t* d_counts, uint32_t* d_and_rows = nullptr, uint32_t* d_missing = nullptr) {
    const size_t gsmem = QG_RING_BYTES + 4096 * 4;
    bool& gattr = ctx->attr_done[0];     // function attributes are per device: remembered per context, not per process
    if (!gattr) {
        CID_CUDA(cudaFuncSetAttribute(query_gather_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
        CID_CUDA(cudaFuncSetAttribute(query_gather_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
        CID_CUDA(cudaFuncSetAttribute(query_gather_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
        CID_CUDA(cudaFuncSetAttribute(query_gather_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
        gattr = true;
    }
    const uint32_t pf64 = ctx->opt_gather_l2_64b != 0;
    GatherOut go{};
    if (!d_and_rows && ctx->gather_out) go = *ctx->gather_out;
    else { go.base[0] = d_counts; go.n = 1; go.stride = idx->N; go.col0 = 0; }
    if (idx->Wp > 128) {
        // rows above 512 bytes: TMA-staged ring (query_gather_tma_kernel)
        const size_t tsmem = QT_RING_BYTES + 128;
        bool& tattr = ctx->attr_done[7];
        if (!tattr) {
            CID_CUDA(cudaFuncSetAttribute(query_gather_tma_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
            CID_CUDA(cudaFuncSetAttribute(query_gather_tma_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
            CID_CUDA(cudaFuncSetAttribute(query_gather_tma_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
            CID_CUDA(cudaFuncSetAttribute(query_gather_tma_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
            tattr = true;
        }
        const uint32_t D = std::min<uint32_t>(8, QT_RING_BYTES / (idx->H * idx->Wp * 4));
        ProfScope ps(ctx, st, d_and_rows ? KID_QUERY_PERFECT : KID_QUERY_COUNTS);
#define CID_QT(HT, AM, OUT)                                                                                             \
    query_gather_tma_kernel<HT, AM><<<(unsigned)nunits, QT_THREADS, tsmem, st>>>(idx->rows, idx->Wp, idx->N, d_rid, d_unit_group, \
                                                                                d_unit_slot0, d_unit_n, OUT, idx->rownz, d_missing, idx->W, go, D)
        if (d_and_rows) { if (idx->H == 2) CID_QT(2, true, d_and_rows); else CID_QT(4, true, d_and_rows); }
        else { if (idx->H == 2) CID_QT(2, false, d_counts); else CID_QT(4, false, d_counts); }
#undef CID_QT
        ctx->launches++;
        CID_CUDA(cudaGetLastError());
        return CID_OK;
    }
    {
        ProfScope ps(ctx, st, d_and_rows ? KID_QUERY_PERFECT : KID_QUERY_COUNTS);
#define CID_QG(HT, AM, OUT)                                                                                             \
    query_gather_kernel<HT, AM><<<(unsigned)nunits, QG_WARPS * 32, gsmem, st>>>(idx->rows, idx->Wp, idx->N, d_rid, d_unit_group, \
                                                                               d_unit_slot0, d_unit_n, OUT, idx->rownz, d_missing, idx->W, go, pf64)
        if (d_and_rows) { if (idx->H == 2) CID_QG(2, true, d_and_rows); else CID_QG(4, true, d_and_rows); }
        else { if (idx->H == 2) CID_QG(2, false, d_counts); else CID_QG(4, false, d_counts); }
#undef CID_QG
    }
    ctx->launches++;
    CID_CUDA(cudaGetLastError());
    return CID_OK;
}

// ================================================================= query_front (small queries, distinct k-mers only)
// Gene search (-g: clean_map(0), batch_search_pe.rs:112-113) and any search with filter 0 need the DISTINCT canonical
// k-mers of a query, not their multiplicities.  For queries of up to QF_MAX_NPOS k-mer positions (genes, plasmids) one CTA
// owns one query: it dedups in a shared-memory table (64-bit CAS), and the thread that inserts a new k-mer hashes it at
// once and appends its row indices to the query's list in HBM.  This replaces table_clear + kmerize_insert + query_hash
// (three passes over a 16-byte-per-slot count table in HBM, ~3 slots per k-mer) by one pass over the bases.
constexpr int QF_THREADS = 256;
constexpr uint32_t QF_MAX_NPOS = 8192;              // table of <= 16384 slots (128 KB), list of <= 16384 k-mers per gather unit
// kmerize_string window with a byte outside ACGTacgt (query_front_kernel<true>): builds the canonical upper-cased string;
// returns true with `key` set when that string is plain ACGT after all (the caller inserts it like any k-mer), else
// deduplicates it by string comparison in the small `dirty` table (offset of first occurrence + 1 per slot) and, when new,
// appends its byte-wise XXH3 rows to the query's list.
constexpr uint32_t QF_DIRTY_SLOTS = 2048;      // distinct non-ACGT k-mers of one record (<= 3/4 of it, else CID_E_UNSUPPORTED)
__device__ __noinline__ static bool dirty_window(const uint8_t* win, const uint8_t* qbytes, uint32_t off, uint32_t k, uint32_t H, ModS mods,
                                                 uint32_t* dirty, uint32_t* s_cnt, uint32_t* rid_q, uint32_t* err, uint64_t& key) {
    uint8_t str[32];
    if (string_kmer(win, k, str, key)) return true;
    const HashIn in = hashin_from_bytes(str, k);
    const uint64_t h0 = hash64(in, k, 0, mods);
    uint32_t slot = (uint32_t)(h0 >> 24) & (QF_DIRTY_SLOTS - 1);
    for (uint32_t probes = 0;; probes++) {
        if (probes >= QF_DIRTY_SLOTS * 3 / 4) { atomicOr(err, ERRF_STRING_NONACGT); return false; }
        const uint32_t prev = atomicCAS(&dirty[slot], 0u, off + 1u);
        if (prev == 0u) break;
        uint8_t other[32]; uint64_t okey;
        string_kmer(qbytes + (prev - 1u), k, other, okey);
        bool same = true;
        for (uint32_t j = 0; j < k; j++) same = same && other[j] == str[j];
        if (same) return false;
        slot = (slot + 1) & (QF_DIRTY_SLOTS - 1);
    }
    const uint32_t i = atomicAdd(s_cnt, 1u);
    uint32_t* out = rid_q + (size_t)i * H;
    out[0] = (uint32_t)mod_s(h0, mods);
    for (uint32_t hh = 1; hh < H; hh++) out[hh] = (uint32_t)hash_row(in, k, hh, mods);
    return false;
}

template <bool STRINGM>      // kmerize_string (-s -m): windows with bytes outside ACGTacgt are k-mers too (separate instantiation:
__global__ void __launch_bounds__(QF_THREADS)    // the byte-string path costs registers the -g / -s paths should not pay)
query_front_kernel(const uint8_t* __restrict__ bases, const uint64_t* __restrict__ seq_offs,
                   const uint64_t* __restrict__ query_offs, const uint32_t* __restrict__ qlist, uint32_t tsize, uint32_t k,
                   int seq_mode, uint32_t H, ModS mods, const uint64_t* __restrict__ rid_base, uint32_t* __restrict__ rid_out,
                   uint32_t* __restrict__ unit_n, unsigned long long* __restrict__ num_kmers, uint32_t* __restrict__ err) {
    extern __shared__ __align__(16) uint8_t dsm[];
    __shared__ uint32_t lut[256];
    __shared__ uint32_t s_cnt;
    unsigned long long* keys = (unsigned long long*)dsm;
    Tile t = tile_carve(dsm + (size_t)tsize * 8, KT_CAP);
    // kmerize_string only: k-mers holding a byte outside ACGTacgt, deduplicated by (offset of first occurrence + 1)
    uint32_t* dirty = (uint32_t*)(dsm + (size_t)tsize * 8 + tile_smem_bytes(KT_CAP));
    const int tid = threadIdx.x;
    const uint32_t q = qlist[blockIdx.x];
    const uint32_t tmask = tsize - 1;
    lut4_init(lut, tid, QF_THREADS);
    for (uint32_t i = tid; i < tsize; i += QF_THREADS) keys[i] = CID_EMPTY_KEY;
    if (STRINGM) for (uint32_t i = tid; i < QF_DIRTY_SLOTS; i += QF_THREADS) dirty[i] = 0u;
    for (int i = tid; i < KT_CAP / 32 + 2; i += QF_THREADS) t.start[i] = 0;      // one sequence per tile: no boundaries inside
    if (tid == 0) s_cnt = 0;
    const uint64_t base = rid_base[q];
    const uint64_t s_lo = __ldg(query_offs + q), s_hi = __ldg(query_offs + q + 1);
    const uint64_t qbase = __ldg(seq_offs + s_lo);
    for (uint64_t s = s_lo; s < s_hi; s++) {
        const uint64_t b0 = __ldg(seq_offs + s), L = __ldg(seq_offs + s + 1) - b0;
        if (L < k) continue;                                 // kmer.rs:94 / :477 `continue`
        for (uint64_t t0 = 0; t0 + k <= L; t0 += KT) {
            const int tile_len = (int)min((uint64_t)(KT + k - 1), L - t0);
            __syncthreads();                                 // the previous tile is consumed
            t.len = tile_len;
            const uint8_t* src = bases + b0 + t0;
            if ((((uintptr_t)src) & 3) == 0) {
                const uint32_t* s4 = (const uint32_t*)src;
                uint32_t* d4 = (uint32_t*)t.ascii;
                const int n4 = tile_len >> 2;
                for (int i = tid; i < n4; i += QF_THREADS) d4[i] = __ldg(s4 + i);
                for (int i = (n4 << 2) + tid; i < tile_len; i += QF_THREADS) t.ascii[i] = __ldg(src + i);
            } else {
                for (int i = tid; i < tile_len; i += QF_THREADS) t.ascii[i] = __ldg(src + i);
            }
            __syncthreads();
            tile_pack(t, KT_CAP, tid, QF_THREADS);
            __syncthreads();
            for (int p = tid; p < KT; p += QF_THREADS) {
                uint64_t key; bool fwd, low;
                if (!tile_kmer(t, p, k, key, fwd, low)) {
                    if (!STRINGM || p + (int)k > t.len) continue;
                    // kmerize_string (kmer.rs:279-293) has no has_no_n test: a window with a byte outside ACGTacgt is a k-mer
                    // too.  Its canonical upper-cased string either turns out to be plain ACGT after all (U -> A in the reverse
                    // complement) and joins the packed set below, or it is hashed byte-wise and deduplicated by comparing strings.
                    if (!dirty_window(t.ascii + p, bases + qbase, (uint32_t)(b0 + t0 + p - qbase), k, H, mods, dirty, &s_cnt,
                                      rid_out + base * H, err, key))
                        continue;
                }
                if (low && seq_mode == CID_SEQ_FASTQ) { atomicOr(err, ERRF_LOWER_RAW); continue; }     // (the caller redoes the call case-aware)
                uint32_t h = (uint32_t)mix64(key) & tmask;
                bool fresh = false;
                for (;;) {
                    const unsigned long long prev = atomicCAS(&keys[h], CID_EMPTY_KEY, (unsigned long long)key);
                    if (prev == CID_EMPTY_KEY) { fresh = true; break; }
                    if (prev == key) break;
                    h = (h + 1) & tmask;
                }
                if (fresh) {
                    const uint32_t i = atomicAdd(&s_cnt, 1u);
                    const HashIn in = hashin_from_key(lut, key, k);
                    uint32_t* out = rid_out + (base + i) * H;
                    for (uint32_t hh = 0; hh < H; hh++) out[hh] = (uint32_t)hash_row(in, k, hh, mods);
                }
            }
        }
    }
    __syncthreads();
    if (tid == 0) { unit_n[q] = s_cnt; num_kmers[q] = s_cnt; }
}

bool query_front_fits(const uint64_t* h_seq_offs, const uint64_t* h_query_offs, uint64_t q0, uint64_t q1, uint32_t k) {
    for (uint64_t q = q0; q < q1; q++) {
        uint64_t npos = 0;
        for (uint64_t s = h_query_offs[q]; s < h_query_offs[q + 1]; s++) {
            const uint64_t L = h_seq_offs[s + 1] - h_seq_offs[s];
            if (L >= k) npos += L - k + 1;
        }
        if (npos > QF_MAX_NPOS) return false;
    }
    return true;
}

// Queries [q0, q1) (all within QF_MAX_NPOS): d_counts / d_num_kmers point at query q0.  d_query_offs / d_seq_offs are the
// caller's whole device arrays (absolute indices).  Synchronises the stream before returning (host staging vectors).
int launch_query_front_gather(cid_ctx* ctx, cudaStream_t st, const cid_index* idx, const uint8_t* d_bases,
                              const uint64_t* d_seq_offs, const uint64_t* d_query_offs, const uint64_t* h_seq_offs,
                              const uint64_t* h_query_offs, uint64_t q0, uint64_t q1, int seq_mode, uint32_t* d_counts,
                              unsigned long long* d_num_kmers, uint32_t* d_and_rows, uint32_t* d_missing) {
    const uint64_t bq = q1 - q0;
    if (bq == 0) return CID_OK;
    const uint32_t k = idx->k, H = idx->H;
    std::vector<uint64_t> base(bq);
    std::vector<uint32_t> group(bq), qlist[4];               // table sizes 2048, 4096, 8192, 16384
    uint64_t total = 0;
    for (uint64_t q = 0; q < bq; q++) {
        uint64_t npos = 0;
        for (uint64_t s = h_query_offs[q0 + q]; s < h_query_offs[q0 + q + 1]; s++) {
            const uint64_t L = h_seq_offs[s + 1] - h_seq_offs[s];
            if (L >= k) npos += L - k + 1;
        }
        base[q] = total;
        total += npos;
        group[q] = (uint32_t)q;
        qlist[npos <= 1024 ? 0 : npos <= 2048 ? 1 : npos <= 4096 ? 2 : 3].push_back((uint32_t)q);
    }
    CID_TRY(ctx->scratch[14].ensure(total * H * 4 + 64));
    CID_TRY(ctx->scratch[15].ensure(bq * 4 + 64));
    CID_TRY(ctx->scratch[6].ensure(bq * 16 + 64));
    uint32_t* d_rid = ctx->scratch[14].as<uint32_t>();
    uint32_t* d_unit_n = ctx->scratch[15].as<uint32_t>();
    uint64_t* d_base = ctx->scratch[6].as<uint64_t>();
    uint32_t* d_group = (uint32_t*)(d_base + bq);
    uint32_t* d_qlist = d_group + bq;
    CID_CUDA(cudaMemcpyAsync(d_base, base.data(), bq * 8, cudaMemcpyHostToDevice, st));
    CID_CUDA(cudaMemcpyAsync(d_group, group.data(), bq * 4, cudaMemcpyHostToDevice, st));
    bool& attr = ctx->attr_done[1];
    if (!attr) {
        CID_CUDA(cudaFuncSetAttribute(query_front_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(16384 * 8 + tile_smem_bytes(KT_CAP))));
        CID_CUDA(cudaFuncSetAttribute(query_front_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(16384 * 8 + tile_smem_bytes(KT_CAP) + QF_DIRTY_SLOTS * 4)));
        attr = true;
    }
    uint64_t at = 0;
    for (int c = 0; c < 4; c++) {
        const uint64_t n = qlist[c].size();
        if (n == 0) continue;
        CID_CUDA(cudaMemcpyAsync(d_qlist + at, qlist[c].data(), n * 4, cudaMemcpyHostToDevice, st));
        const uint32_t tsize = 2048u << c;
        ProfScope ps(ctx, st, KID_QUERY_FRONT);
        const size_t fsmem = (size_t)tsize * 8 + tile_smem_bytes(KT_CAP) + (seq_mode == CID_SEQ_STRING ? QF_DIRTY_SLOTS * 4 : 0);
        if (seq_mode == CID_SEQ_STRING)
            query_front_kernel<true><<<(unsigned)n, QF_THREADS, fsmem, st>>>(
                d_bases, d_seq_offs, d_query_offs + q0, d_qlist + at, tsize, k, seq_mode, H, make_mods(idx->S, idx->hv, (const HashCfg*)idx->d_hcfg), d_base, d_rid, d_unit_n,
                d_num_kmers, ctx->d_err);
        else
            query_front_kernel<false><<<(unsigned)n, QF_THREADS, fsmem, st>>>(
                d_bases, d_seq_offs, d_query_offs + q0, d_qlist + at, tsize, k, seq_mode, H, make_mods(idx->S, idx->hv, (const HashCfg*)idx->d_hcfg), d_base, d_rid, d_unit_n,
                d_num_kmers, ctx->d_err);
        ctx->launches++;
        CID_CUDA(cudaGetLastError());
        at += n;
    }
    // one unit per query: a column-sharded destination may be written with plain stores
    GatherOut dense_out{};
    const GatherOut* const shared = ctx->gather_out;
    if (shared && !d_and_rows) { dense_out = *shared; dense_out.dense = 1; ctx->gather_out = &dense_out; }
    int grc;
    if (d_and_rows && !(idx->Wp >= 4 && idx->Wp <= 512 && (idx->H == 2 || idx->H == 4))) {
        ProfScope ps(ctx, st, KID_QUERY_PERFECT);
        perfect_rids_kernel<<<(unsigned)bq, 256, (size_t)idx->Wp * 4, st>>>(idx->rows, idx->Wp, idx->W, idx->H, d_rid, d_group, d_base, d_unit_n,
                                                                           idx->rownz, d_and_rows, d_missing);
        ctx->launches++;
        grc = cudaGetLastError() == cudaSuccess ? CID_OK : CID_E_CUDA;
    } else grc = launch_query_gather(ctx, st, idx, d_rid, d_group, d_base, d_unit_n, bq, d_counts, d_and_rows, d_missing);
    ctx->gather_out = shared;
    CID_TRY(grc);
    CID_CUDA(cudaStreamSynchronize(st));
    return CID_OK;
}

int launch_query_counts(cid_ctx* ctx, cudaStream_t st, const cid_index* idx, const void* d_table,
                        const uint32_t* d_unit_group, const uint64_t* d_unit_slot0, const uint32_t* d_unit_nslots,
                        uint64_t nunits, uint64_t total_slots, const int64_t* d_filter, uint32_t* d_counts,
                        unsigned long long* d_num_kmers, bool want_uniq, uint32_t* d_uniq_list, uint32_t uniq_cap,
                        uint32_t* d_uniq_n) {
    if (nunits == 0) return CID_OK;
    // streaming path: 16-byte-aligned rows of at most 512 bytes, no unique-hit summaries
    if (!want_uniq && idx->Wp >= 4 && idx->Wp <= 512 && (idx->H == 2 || idx->H == 4) && !ctx->opt_query_fused) {
        CID_TRY(ctx->scratch[14].ensure(total_slots * idx->H * 4 + 64));
        CID_TRY(ctx->scratch[15].ensure(nunits * 4 + 64));
        uint32_t* d_rid = ctx->scratch[14].as<uint32_t>();
        uint32_t* d_unit_n = ctx->scratch[15].as<uint32_t>();
        {
            ProfScope ps(ctx, st, KID_QUERY_HASH);
            query_hash_kernel<<<(unsigned)nunits, 256, 0, st>>>((const Slot*)d_table, d_unit_group, d_unit_slot0, d_unit_nslots,
                                                              (const long long*)d_filter, idx->k, idx->H, make_mods(idx->S, idx->hv, (const HashCfg*)idx->d_hcfg),
                                                              d_rid, d_unit_n, d_num_kmers);
        }
        ctx->launches++;
        CID_CUDA(cudaGetLastError());
        return launch_query_gather(ctx, st, idx, d_rid, d_unit_group, d_unit_slot0, d_unit_n, nunits, d_counts);
    }
    size_t smem = unit_smem_bytes(idx->H);
    bool& attr_set = ctx->attr_done[2];
    if (!attr_set) {
#define CID_QC_ATTR(VEC, UQ, HT) CID_CUDA(cudaFuncSetAttribute(query_counts_kernel<VEC, UQ, HT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024))
#define CID_QC_ATTR3(VEC, UQ) CID_QC_ATTR(VEC, UQ, 0); CID_QC_ATTR(VEC, UQ, 2); CID_QC_ATTR(VEC, UQ, 4)
        CID_QC_ATTR3(1, true); CID_QC_ATTR3(1, false); CID_QC_ATTR3(2, true); CID_QC_ATTR3(2, false);
        CID_QC_ATTR3(4, true); CID_QC_ATTR3(4, false);
#undef CID_QC_ATTR3
#undef CID_QC_ATTR
        CID_CUDA(cudaFuncSetAttribute(query_uniq_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
    }
    ModS mods = make_mods(idx->S, idx->hv, (const HashCfg*)idx->d_hcfg);
    const uint32_t vec = idx->Wp >= 4 ? 4 : idx->Wp;      // Wp is 1, 2 or a multiple of 4
    const bool inline_uniq = want_uniq && idx->Wp <= 32 * vec;   // one column block: the warp sees the whole row
    {
    ProfScope ps(ctx, st, KID_QUERY_COUNTS);
#define CID_QC_LAUNCH_H(VEC, UQ, HT)                                                                                   \
    query_counts_kernel<VEC, UQ, HT><<<(unsigned)nunits, 256, smem, st>>>(                                             \
        idx->rows, idx->Wp, idx->W, idx->N, idx->k, idx->H, mods, (const Slot*)d_table, d_unit_group, d_unit_slot0,    \
        d_unit_nslots, (const long long*)d_filter, d_counts, d_num_kmers, d_uniq_list, uniq_cap, d_uniq_n)
#define CID_QC_LAUNCH(VEC, UQ)                                                                                         \
    do { if (idx->H == 2) CID_QC_LAUNCH_H(VEC, UQ, 2); else if (idx->H == 4) CID_QC_LAUNCH_H(VEC, UQ, 4);              \
         else CID_QC_LAUNCH_H(VEC, UQ, 0); } while (0)
    // VEC words per lane: 128-bit loads whenever a row is 16-byte aligned (several k-mers per warp instruction)
    if (vec == 4) { if (inline_uniq) CID_QC_LAUNCH(4, true); else CID_QC_LAUNCH(4, false); }
    else if (vec == 2) { if (inline_uniq) CID_QC_LAUNCH(2, true); else CID_QC_LAUNCH(2, false); }
    else { if (inline_uniq) CID_QC_LAUNCH(1, true); else CID_QC_LAUNCH(1, false); }
#undef CID_QC_LAUNCH
#undef CID_QC_LAUNCH_H
    }
    ctx->launches++;
    CID_CUDA(cudaGetLastError());
    if (want_uniq && !inline_uniq) {
        ProfScope ps(ctx, st, KID_QUERY_UNIQ_WIDE);
        query_uniq_wide_kernel<<<(unsigned)nunits, 256, smem, st>>>(
            idx->rows, idx->Wp, idx->k, idx->H, mods, (const Slot*)d_table, d_unit_group, d_unit_slot0, d_unit_nslots,
            (const long long*)d_filter, d_uniq_list, uniq_cap, d_uniq_n);
        ctx->launches++;
        CID_CUDA(cudaGetLastError());
    }
    return CID_OK;
}

// ================================================================= query_perfect
__global__ void __launch_bounds__(256)
query_perfect_kernel(const uint32_t* __restrict__ rows, const uint32_t* __restrict__ rownz, uint32_t Wp, uint32_t W,
                     uint32_t k, uint32_t H, ModS mods, const Slot* __restrict__ table,
                     const uint32_t* __restrict__ unit_group, const uint64_t* __restrict__ unit_slot0,
                     const uint32_t* __restrict__ unit_nslots, uint32_t* __restrict__ and_rows,
                     uint32_t* __restrict__ missing, unsigned long long* __restrict__ num_kmers) {
    extern __shared__ __align__(16) uint8_t dsm[];
    UnitSmem u = unit_carve(dsm, H);
    const uint32_t g = unit_group[blockIdx.x];
    const int tid = threadIdx.x;
    const uint32_t item_nslots = unit_nslots[blockIdx.x];
    uint32_t total_n = 0;
    bool miss = false;
    // thread owns word columns col = tid%32 (+32..) and the k-mers i = tid/32 (mod 8); at most 16 columns per thread
    uint32_t acc[16];
#pragma unroll
    for (int c = 0; c < 16; c++) acc[c] = 0xFFFFFFFFu;
    for (uint32_t c0 = 0; c0 < item_nslots; c0 += QUERY_CHUNK) {
        __syncthreads();
        const uint32_t n = unit_collect_and_hash(u, table, unit_slot0[blockIdx.x] + c0, min((uint32_t)QUERY_CHUNK, item_nslots - c0),
                                                 0ll, k, H, mods);
        total_n += n;
        // any absent row -> "No perfect hits!" (perfect_search.rs:32-33,38-39)
        for (uint32_t i = tid; i < n * H; i += 256) {
            uint32_t r = u.rowid[i];
            if (!((__ldg(rownz + (r >> 5)) >> (r & 31)) & 1u)) miss = true;
        }
#pragma unroll
        for (int c = 0; c < 16; c++) {
            const uint32_t col = (tid % 32) + 32 * c;
            if (col < W)
                for (uint32_t i = tid / 32; i < n; i += 8)
                    for (uint32_t h = 0; h < H; h++) acc[c] &= __ldg(rows + (uint64_t)u.rowid[i * H + h] * Wp + col);
        }
    }
    if (tid == 0 && total_n) atomicAdd(&num_kmers[g], (unsigned long long)total_n);
    if (miss) atomicOr(&missing[g], 1u);
#pragma unroll
    for (int c = 0; c < 16; c++) {
        const uint32_t col = (tid % 32) + 32 * c;
        if (col < W && acc[c] != 0xFFFFFFFFu) atomicAnd(&and_rows[(uint64_t)g * W + col], acc[c]);
    }
}
int launch_query_perfect(cid_ctx* ctx, cudaStream_t st, const cid_index* idx, const void* d_table,
                         const uint32_t* d_unit_group, const uint64_t* d_unit_slot0, const uint32_t* d_unit_nslots,
                         uint64_t nunits, uint32_t* d_and_rows, uint32_t* d_missing, unsigned long long* d_num_kmers) {
    if (nunits == 0) return CID_OK;
    if (idx->W > 512) { set_error("perfect search: more than 16384 accessions per shard not supported"); return CID_E_UNSUPPORTED; }
    size_t smem = unit_smem_bytes(idx->H);
    bool& attr_set = ctx->attr_done[3];
    if (!attr_set) {
        CID_CUDA(cudaFuncSetAttribute(query_perfect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
    }
    ProfScope ps(ctx, st, KID_QUERY_PERFECT);
    query_perfect_kernel<<<(unsigned)nunits, 256, smem, st>>>(idx->rows, idx->rownz, idx->Wp, idx->W, idx->k, idx->H,
                                                             make_mods(idx->S, idx->hv, (const HashCfg*)idx->d_hcfg), (const Slot*)d_table, d_unit_group,
                                                             d_unit_slot0, d_unit_nslots, d_and_rows, d_missing,
                                                             d_num_kmers);
    ctx->launches++;
    CID_CUDA(cudaGetLastError());
    return CID_OK;
}

// ================================================================= hash parity hook
__global__ void hash_kmers_kernel(const uint8_t* __restrict__ kmers, uint64_t n, uint32_t k, uint32_t H, ModS mods,
                                  uint64_t* __restrict__ out) {
    __shared__ uint32_t lut[256];
    lut4_init(lut, threadIdx.x, blockDim.x);
    __syncthreads();
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = 0;
    for (uint32_t j = 0; j < k; j++) key = (key << 2) | base_code(kmers[i * k + j]);
    HashIn in = hashin_from_key(lut, key, k);
    for (uint32_t h = 0; h < H; h++) out[i * H + h] = hash_row(in, k, h, mods);
}
int launch_hash_kmers(cid_ctx* ctx, cudaStream_t st, const cid_index* idx, const uint8_t* d_kmers, uint64_t n,
                      uint64_t* d_rows) {
    if (n == 0) return CID_OK;
    ProfScope ps(ctx, st, KID_OTHER);
    hash_kmers_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_kmers, n, idx->m ? idx->m : idx->k, idx->H, make_mods(idx->S, idx->hv, (const HashCfg*)idx->d_hcfg), d_rows);
    ctx->launches++;
    CID_CUDA(cudaGetLastError());
    return CID_OK;
}

// ================================================================= unique-hit summaries (reports.rs:20-26)
// The fused gather kernel leaves one (query, accession, multiplicity) triple per k-mer that hit exactly one accession.
// generate_report wants, per (query, accession): their number ("specific"), the mean and the mode of the multiplicities.
// uniq_hist: one counter per (cell = query * N + accession, multiplicity < MB); triples with larger multiplicities raise
// *ovf (the caller then falls back to the host summary).  uniq_reduce: one warp per cell -> n, sum, mode (ties towards the
// smallest multiplicity, like the host path and the oracle).
__global__ void __launch_bounds__(256)
uniq_hist_kernel(const uint32_t* __restrict__ list, uint32_t nu, uint32_t N, uint32_t MB, uint32_t* __restrict__ hist,
                 uint32_t* __restrict__ ovf) {
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < nu; i += gridDim.x * 256) {
        const uint32_t q = __ldg(list + 3 * (uint64_t)i), c = __ldg(list + 3 * (uint64_t)i + 1), m = __ldg(list + 3 * (uint64_t)i + 2);
        if (m < MB) atomicAdd(&hist[((uint64_t)q * N + c) * MB + m], 1u);
        else atomicOr(ovf, 1u);
    }
}
__global__ void __launch_bounds__(256)
uniq_reduce_kernel(const uint32_t* __restrict__ hist, uint64_t cells, uint32_t MB, unsigned long long* __restrict__ out_n,
                   unsigned long long* __restrict__ out_sum, unsigned long long* __restrict__ out_mode) {
    const uint64_t cell = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (cell >= cells) return;
    unsigned long long n = 0, sum = 0;
    uint32_t best = 0, mode = 0;
    for (uint32_t m = lane; m < MB; m += 32) {
        const uint32_t h = __ldg(hist + cell * MB + m);
        n += h; sum += (unsigned long long)h * m;
        if (h > best) { best = h; mode = m; }            // ascending m within a lane: strict > keeps the smallest
    }
    for (int o = 16; o > 0; o >>= 1) {
        n += __shfl_xor_sync(0xffffffffu, n, o);
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const uint32_t ob = __shfl_xor_sync(0xffffffffu, best, o), om = __shfl_xor_sync(0xffffffffu, mode, o);
        if (ob > best || (ob == best && om < mode)) { best = ob; mode = om; }
    }
    if (lane == 0) { out_n[cell] = n; out_sum[cell] = sum; out_mode[cell] = n ? mode : 0; }
}
int launch_uniq_summaries(cid_ctx* ctx, cudaStream_t st, const uint32_t* d_list, uint32_t nu, uint32_t N, uint64_t cells, uint32_t MB,
                          uint32_t* d_hist, uint32_t* d_ovf, unsigned long long* d_n, unsigned long long* d_sum,
                          unsigned long long* d_mode) {
    ProfScope ps(ctx, st, KID_OTHER);
    CID_CUDA(cudaMemsetAsync(d_hist, 0, cells * MB * 4, st));
    CID_CUDA(cudaMemsetAsync(d_ovf, 0, 4, st));
    if (nu) uniq_hist_kernel<<<std::min<unsigned>((nu + 255) / 256, (unsigned)ctx->sm_count * 8), 256, 0, st>>>(d_list, nu, N, MB, d_hist, d_ovf);
    uniq_reduce_kernel<<<(unsigned)((cells + 7) / 8), 256, 0, st>>>(d_hist, cells, MB, d_n, d_sum, d_mode);
    ctx->launches += nu ? 2 : 1;
    CID_CUDA(cudaGetLastError());
    return CID_OK;
}

// ================================================================= region_compact (large queries)
// The count table of a read-set query is sized for its k-mer POSITIONS (2 slots per position) and is mostly empty once the
// frequency filter is known (30x reads: 3.3 M survivors in 268 M slots).  Instead of letting the gather kernel's work units
// scan 4 GB of slots, the survivors (count > filter) are copied once, at streaming bandwidth, into a dense slot list that the
// same kernels then walk.  COUNT_ONLY: just the number of survivors (for filters whose histogram is not at hand).
template <bool COUNT_ONLY>
__global__ void __launch_bounds__(256)
region_compact_kernel(const Slot* __restrict__ region, uint64_t nslots, long long filt, Slot* __restrict__ dense,
                      unsigned long long* __restrict__ n_out, uint64_t cap) {
    constexpr int U = 8;               // slots loaded per thread before any of them is looked at (loads in flight)
    // One global atomic per CTA and round (2,048 slots): a first version with one atomic per warp and slot row spent 3.4 ms of
    // a 4.3 GB scan on 2.8 M same-address atomics.
    __shared__ uint32_t wcnt[2][8];
    __shared__ unsigned long long bbase[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long mine = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (nslots + stride * U - 1) / (stride * U);
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t r = 0; r < rounds; r++, s += stride * U) {
        uint4 v[U];                    // {key lo, key hi, count, pad}: one 16-byte streaming load per slot
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t at = s + (uint64_t)u * stride;
            v[u] = at < nslots ? __ldcs((const uint4*)region + at) : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
        }
        uint32_t keep = 0;
#pragma unroll
        for (int u = 0; u < U; u++)
            if ((v[u].x & v[u].y) != 0xFFFFFFFFu && (long long)v[u].z > filt) keep |= 1u << u;
        const uint32_t cnt = __popc(keep);
        if (COUNT_ONLY) { mine += cnt; continue; }
        uint32_t incl = cnt;           // inclusive scan over the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const int par = (int)(r & 1);
        if (lane == 31) wcnt[par][warp] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) tot += wcnt[par][w];
            bbase[par] = tot ? atomicAdd(n_out, (unsigned long long)tot) : 0ull;
        }
        __syncthreads();
        if (keep) {
            uint64_t at = bbase[par] + (incl - cnt);
            for (int w = 0; w < warp; w++) at += wcnt[par][w];
#pragma unroll
            for (int u = 0; u < U; u++)
                if ((keep >> u) & 1u) { if (at < cap) ((uint4*)dense)[at] = v[u]; at++; }
        }
    }
    if (COUNT_ONLY) {
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if (lane == 0 && mine) atomicAdd(n_out, mine);
    }
}
int launch_region_compact(cid_ctx* ctx, cudaStream_t st, const void* d_region, uint64_t nslots, int64_t filt, void* d_dense,
                          unsigned long long* d_n, uint64_t cap) {
    unsigned grid = (unsigned)std::min<uint64_t>((nslots + 255) / 256, (uint64_t)ctx->sm_count * 8);
    if (grid == 0) grid = 1;
    ProfScope ps(ctx, st, KID_OTHER);
    if (d_dense) region_compact_kernel<false><<<grid, 256, 0, st>>>((const Slot*)d_region, nslots, (long long)filt, (Slot*)d_dense, d_n, cap);
    else region_compact_kernel<true><<<grid, 256, 0, st>>>((const Slot*)d_region, nslots, (long long)filt, nullptr, d_n, 0);
    ctx->launches++;
    CID_CUDA(cudaGetLastError());
    return CID_OK;
}

// ================================================================= column-sharded default report
// "This k-mer hits exactly one accession" (batch_search_pe.rs:75-82) is a statement about the WHOLE row, so a column shard
// can only say how many of ITS accessions a k-mer hits.  All shards walk the same dense survivor list (one rank's list,
// broadcast), slots_popcount leaves min(popcount, 2) and the single hit's colour per list index, the per-index bytes are
// summed across shards (one all-reduce of a byte per k-mer), and uniq_emit keeps the k-mers whose local and global
// popcounts are both 1 -- in the same (query, accession, multiplicity) triple format the unsharded kernels emit.
__global__ void __launch_bounds__(256)
slots_popcount_kernel(const uint32_t* __restrict__ rows, uint32_t Wp, uint32_t k, uint32_t H, ModS mods,
                      const Slot* __restrict__ slots, uint64_t n, uint8_t* __restrict__ pc_out, uint32_t* __restrict__ col_out) {
    __shared__ uint32_t lut[256];
    lut4_init(lut, threadIdx.x, 256);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // a warp takes 32 consecutive list entries: lane L hashes entry L, then the rows of one entry per step are read by all lanes
    for (uint64_t i0 = ((uint64_t)blockIdx.x * 8 + warp) * 32; i0 < n; i0 += (uint64_t)gridDim.x * 8 * 32) {
        uint32_t rid_l[MAX_HASH];
#pragma unroll
        for (int h = 0; h < MAX_HASH; h++) rid_l[h] = 0;
        if (i0 + lane < n) {
            HashIn in = hashin_from_key(lut, slots[i0 + lane].key, k);
            hashin_apply_case(in, slot_cs(slots[i0 + lane].pad), k);
#pragma unroll
            for (int h = 0; h < MAX_HASH; h++) if ((uint32_t)h < H) rid_l[h] = (uint32_t)hash_row(in, k, h, mods);
        }
        const uint32_t cnt = (uint32_t)min((uint64_t)32, n - i0);
        uint32_t my_pc = 0, my_col = 0xFFFFFFFFu;
        for (uint32_t j = 0; j < cnt; j++) {
            uint64_t rid[MAX_HASH];
#pragma unroll
            for (int h = 0; h < MAX_HASH; h++) rid[h] = (uint32_t)h < H ? (uint64_t)__shfl_sync(0xffffffffu, rid_l[h], j) : 0ull;
            uint32_t pc = 0, where = 0;
            for (uint32_t c = lane; c < Wp; c += 32) {
                uint32_t x = 0xFFFFFFFFu;
#pragma unroll
                for (int h = 0; h < MAX_HASH; h++) if ((uint32_t)h < H) x &= __ldg(rows + rid[h] * Wp + c);
                if (x) { pc += __popc(x); where = c * 32 + (__ffs(x) - 1); }
            }
            uint32_t tot = pc;
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            const uint32_t owner = __ballot_sync(0xffffffffu, pc == 1);
            const uint32_t colour = __shfl_sync(0xffffffffu, where, owner ? __ffs(owner) - 1 : 0);
            if ((uint32_t)lane == j) { my_pc = min(tot, 2u); my_col = tot == 1 ? colour : 0xFFFFFFFFu; }
        }
        if (i0 + lane < n) { pc_out[i0 + lane] = (uint8_t)my_pc; col_out[i0 + lane] = my_col; }       // coalesced
    }
}
__global__ void __launch_bounds__(256)
uniq_emit_kernel(const Slot* __restrict__ slots, const uint64_t* __restrict__ prefix, uint32_t nq, uint64_t n,
                 const uint8_t* __restrict__ pc_local, const uint8_t* __restrict__ pc_sum, const uint32_t* __restrict__ col,
                 uint32_t* __restrict__ list, uint32_t cap, uint32_t* __restrict__ n_out) {
    const int lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    const uint64_t rounds = (n + stride - 1) / stride;
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    for (uint64_t r = 0; r < rounds; r++, i += stride) {
        const bool keep = i < n && pc_local[i] == 1 && pc_sum[i] == 1;
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (!bal) continue;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(n_out, (uint32_t)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) {
            uint32_t lo = 0, hi = nq;                   // query of list index i: prefix[q] <= i < prefix[q + 1]
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(prefix + mid) <= i) lo = mid; else hi = mid; }
            const uint32_t e = base + __popc(bal & ((1u << lane) - 1));
            if (e < cap) { list[3 * (uint64_t)e] = lo; list[3 * (uint64_t)e + 1] = col[i]; list[3 * (uint64_t)e + 2] = slots[i].count; }
        }
    }
}
int launch_slots_popcount(cid_ctx* ctx, cudaStream_t st, const cid_index* idx, const void* d_slots, uint64_t n, uint8_t* d_pc,
                          uint32_t* d_col) {
    if (n == 0) return CID_OK;
    ProfScope ps(ctx, st, KID_QUERY_UNIQ_WIDE);
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 32);
    slots_popcount_kernel<<<grid, 256, 0, st>>>(idx->rows, idx->Wp, idx->k, idx->H, make_mods(idx->S, idx->hv, (const HashCfg*)idx->d_hcfg), (const Slot*)d_slots, n, d_pc, d_col);
    ctx->launches++;
    CID_CUDA(cudaGetLastError());
    return CID_OK;
}
int launch_uniq_emit(cid_ctx* ctx, cudaStream_t st, const void* d_slots, const uint64_t* d_prefix, uint32_t nq, uint64_t n,
                     const uint8_t* d_pc_local, const uint8_t* d_pc_sum, const uint32_t* d_col, uint32_t* d_list, uint32_t cap,
                     uint32_t* d_n) {
    if (n == 0) return CID_OK;
    ProfScope ps(ctx, st, KID_OTHER);
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 8);
    uniq_emit_kernel<<<grid, 256, 0, st>>>((const Slot*)d_slots, d_prefix, nq, n, d_pc_local, d_pc_sum, d_col, d_list, cap, d_n);
    ctx->launches++;
    CID_CUDA(cudaGetLastError());
    return CID_OK;
}

}  // namespace cid
