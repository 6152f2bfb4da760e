// CUDA kernels (sm_100a) for read_id: read_id_mt_pe.rs:282-363 parallel_vec with m == 0.
//
// Per read (1 or 2 mates) the reference builds an FnvHashSet<String> of canonical k-mers
// (kmer.rs:221-243), iterates it IN HASH-TABLE ORDER and votes (search_index :104-165 /
// search_index_classic :66-102), breaking at the first k-mer with an absent row.  The result
// therefore depends on hashbrown's bucket layout, which is reproduced here in three kernels:
//
//   readid_kmerize  (warp per read)    canonical k-mers, per-read dedup, FNV-1a low bits,
//                                      emitted as the sequence of HashSet::insert calls
//   readid_order    (thread per read)  hashbrown growth/probe emulation in shared memory ->
//                                      k-mers in iteration order
//   readid_vote     (warp per read)    XXH3 x num_hash, row gather, AND, first-miss cut-off,
//                                      candidate set from the first `start_sample` k-mers, counts
#include <algorithm>

#include "cid_device.cuh"
#include "cid_internal.h"
#include "cid_readid_common.cuh"

namespace cid {

constexpr uint32_t FINE_STEPS = 2;   // fine-grained (lane per hash row) steps at the start of each read's vote
// option readid_report_steps (column-sharded read_id): a report entry's colour carries, above this bit, the index of the
// k-mer (in set order) whose AND row first inserted the colour into final_report -- what a merge of per-shard reports
// needs to restore the reference's insertion order (colours < 2^20 per shard, k-mers < 2^11 per read)
constexpr uint32_t REP_STEP_SHIFT = 20;

// entry layout (u32): [15:0] fnv low16 | [25:16] tile position | [26] took_fwd | [27] fresh
#define ENT_TP(e) (((e) >> 16) & 0x3FFu)
#define ENT_FWD(e) (((e) >> 26) & 1u)
#define ENT_FRESH(e) (((e) >> 27) & 1u)


// ================================================================= readid_kmerize
// COMPACT (reads of <= 255 k-mer positions): only the FIRST occurrence of every distinct k-mer is
// emitted, split into the three arrays the later kernels want -- hp8/h9w = low 9 bits of FNV-1a
// (the order kernel's only input), ent16 = tile position | strand << 10 (the vote kernel's).
// Duplicate HashSet::insert calls matter to hashbrown only through "is there a call after the
// last fresh one" (bit 31 of nocc), see readid_order_small_kernel.
template <bool COMPACT, bool MINI>     // MINI: .mxi index, the set holds minimizers of length mini_m (a separate
__global__ void __launch_bounds__(RA_WARPS * 32)   // instantiation so the k-mer path keeps its register budget)
readid_kmerize_kernel(const ReadSrc src,
                      const uint64_t* __restrict__ seq_offs, const uint64_t* __restrict__ read_offs, uint64_t r0,
                      uint64_t nreads, uint32_t k, uint32_t mini_m, uint32_t d, int cap, uint32_t maxocc, uint32_t tsize,
                      uint32_t* __restrict__ entries, uint8_t* __restrict__ hp8, uint32_t* __restrict__ h9w,
                      uint16_t* __restrict__ ent16, uint32_t* __restrict__ nocc, uint32_t* __restrict__ nfresh,
                      uint32_t* __restrict__ flags, uint32_t* __restrict__ err, uint32_t* __restrict__ slow_list,
                      uint32_t* __restrict__ slow_n) {
    extern __shared__ __align__(16) uint8_t dsm[];
    __shared__ uint32_t lut[256];
    lut4_init(lut, threadIdx.x, blockDim.x);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = ((tile_smem_bytes(cap) + 7) & ~(size_t)7) + (size_t)tsize * 8 + (size_t)tsize * 4 +
                            (size_t)cap * 4 + (MAX_MATES + 1) * 4 + 4 + 32 + (COMPACT ? 512 : 0);
    uint8_t* base = dsm + (size_t)warp * ((per_warp + 15) & ~(size_t)15);
    Tile t = tile_carve(base, cap);
    unsigned long long* tkeys = (unsigned long long*)(base + ((tile_smem_bytes(cap) + 7) & ~(size_t)7));
    uint32_t* tmin = (uint32_t*)(tkeys + tsize);
    uint32_t* pinfo = tmin + tsize;
    uint32_t* moffs = pinfo + cap;
    uint32_t* h9s = moffs + (MAX_MATES + 1) + 1;      // COMPACT: 9th hash bit of up to 256 fresh k-mers
    uint16_t* fslot = (uint16_t*)(h9s + 8);           // COMPACT: dedup-table slot of the i-th distinct k-mer
    const uint32_t tmask = tsize - 1;
    const uint32_t hwords = (maxocc + 31) / 32;

    for (uint64_t rl = (uint64_t)blockIdx.x * RA_WARPS + warp; rl < nreads; rl += (uint64_t)gridDim.x * RA_WARPS) {
        const uint64_t r = r0 + rl;
        const uint64_t s_begin = __ldg(read_offs + r), s_end = __ldg(read_offs + r + 1);
        {   // pull the NEXT read of this warp towards L1 while this one is processed (its bytes are read once, at the top)
            const uint64_t rn = rl + (uint64_t)gridDim.x * RA_WARPS;
            if (rn < nreads) {
                const uint64_t nb0 = __ldg(seq_offs + __ldg(read_offs + r0 + rn)), nb1 = __ldg(seq_offs + __ldg(read_offs + r0 + rn + 1));
                const uint64_t off = nb0 + (uint64_t)(lane & 15) * 128;
                if (!src.pk && nb1 > nb0 && off < nb1 + 127) {            // every 128-byte line the read touches
                    const uint8_t* pa = (lane < 16 ? src.bases : src.quals);
                    if (pa) asm volatile("prefetch.global.L1 [%0];" ::"l"(pa + min(off, nb1 - 1)));
                }
            }
        }
        uint32_t fl = 0, emitted = 0, nfr = 0;
        ReadGeom g;
        __syncwarp();
        bool ok = s_end > s_begin;
        if (ok) {
            // read_id_mt_pe.rs:305 `r.1[0].len() < k` -> too_short
            uint64_t l0 = __ldg(seq_offs + s_begin + 1) - __ldg(seq_offs + s_begin);
            if (l0 < k) ok = false;
        }
        if (!ok) fl |= 1u;
        // Reads this kernel cannot hold -- longer than the shared-memory tile, or (below) with a lower-case base inside a
        // k-mer, which 2-bit keys cannot spell -- go to the general path (cid_readid_big.cu) through a device list.
        bool slow = false;
        if (ok && !warp_load_read(t, cap, src, r, seq_offs, s_begin, s_end, moffs, lane, g)) { slow = true; ok = false; }
        if (ok && !MINI) {
            // kmer.rs:229 `0..l.len()-k+1` wraps for a later mate shorter than k-1 -> slice panic
            // (minimerize_vector_skip_n_set has a `length_l < k` guard instead, kmer.rs:372: the mate is skipped)
            for (int m = 1; m < g.nm; m++) if (moffs[m + 1] - moffs[m] + 1 < k) { fl |= 2u; ok = false; }
        }
        if (ok) {
            for (uint32_t i = lane; i < tsize; i += 32) { tkeys[i] = CID_EMPTY_KEY; tmin[i] = 0xFFFFFFFFu; }
            __syncwarp();
            // pass 1: every position -> (valid, dedup slot, strand)
            bool lowseen = false;
            for (int tp0 = 0; tp0 < g.len; tp0 += 32) {
                int tp = tp0 + lane;
                uint32_t info = 0;
                if (tp < g.len) {
                    bool take = true;
                    if (d != 1) {                       // kmer.rs:229 step_by(d): position inside the mate
                        int m = 0;
                        for (int j = 1; j < g.nm; j++) if ((int)moffs[j] <= tp) m = j;
                        take = ((uint32_t)tp - moffs[m]) % d == 0;
                    }
                    uint64_t key; bool fwd, low;
                    if (take && tile_kmer(t, tp, k, key, fwd, low)) {
                        uint32_t ipos = (uint32_t)tp;      // window that spells the set's item (k-mer, or its minimizer)
                        if (MINI) {                        // kmer.rs:363-394: the set holds upper-cased minimizers
                            bool mfwd;
                            key = tile_minimizer(t, tp, k, mini_m, key, fwd, low, ipos, mfwd);
                            fwd = mfwd;
                        } else if (low) lowseen = true;
                        uint32_t s = (uint32_t)mix64(key) & tmask;
                        for (;;) {
                            unsigned long long prev = atomicCAS(&tkeys[s], CID_EMPTY_KEY, (unsigned long long)key);
                            if (prev == CID_EMPTY_KEY || prev == key) break;
                            s = (s + 1) & tmask;
                        }
                        atomicMin(&tmin[s], (uint32_t)tp);
                        info = 0x80000000u | (fwd ? 0x40000000u : 0u) | (ipos << 16) | s;     // s < tsize <= 2^16
                    }
                }
                if (tp < cap) pinfo[tp] = info;
            }
            __syncwarp();
            if (__any_sync(0xffffffffu, lowseen)) { slow = true; ok = false; }
        }
        if (ok) {
            // pass 2: emit the insert-call sequence in sequence order (kmer.rs:225-240)
            uint32_t* out = entries + rl * (uint64_t)maxocc;
            uint32_t last_fresh = 1;
            if (COMPACT) { if (lane < 8) h9s[lane] = 0; __syncwarp(); }
            for (int tp0 = 0; tp0 < g.len; tp0 += 32) {
                int tp = tp0 + lane;
                uint32_t info = tp < g.len ? pinfo[tp] : 0u;
                bool valid = info >> 31;
                uint32_t ent = 0, f = 0;
                bool fresh = false;
                if (valid) {
                    uint32_t s = info & 0xFFFFu;
                    fresh = tmin[s] == (uint32_t)tp;
                    f = (!COMPACT && fresh) ? (fnv1a_low32_key_lut(lut, tkeys[s], MINI ? mini_m : k) & 0xFFFFu) : 0u;
                    ent = f | (info & 0x03FF0000u) | (((info >> 30) & 1u) << 26) | ((fresh ? 1u : 0u) << 27);
                }
                const uint32_t bal = __ballot_sync(0xffffffffu, valid), balf = __ballot_sync(0xffffffffu, fresh);
                if (COMPACT) {
                    const uint32_t fi = nfr + __popc(balf & ((1u << lane) - 1));
                    if (fresh && fi < maxocc) {
                        ent16[rl * (uint64_t)maxocc + fi] = (uint16_t)(((info >> 16) & 0x3FFu) | (((info >> 30) & 1u) << 10));
                        fslot[fi] = (uint16_t)(info & 0xFFFFu);
                    }
                    if (bal) last_fresh = (balf >> (31 - __clz(bal))) & 1u;
                } else {
                    const uint32_t rank = __popc(bal & ((1u << lane) - 1));
                    if (valid && emitted + rank < maxocc) out[emitted + rank] = ent;
                }
                emitted += __popc(bal);
                nfr += __popc(balf);
            }
            if (COMPACT) {
                // FNV-1a of the distinct k-mers, 32 at a time with every lane busy (in the loop over the read's positions only
                // the lanes of FIRST occurrences hashed: about a third of them, at the full cost of the 31-step chain per step)
                __syncwarp();
                const uint32_t nfd = min(nfr, maxocc);
                for (uint32_t f0 = 0; f0 < nfd; f0 += 32) {
                    const uint32_t fi = f0 + lane;
                    const bool act = fi < nfd;
                    const uint32_t f = act ? (fnv1a_low32_key_lut(lut, (uint64_t)tkeys[fslot[fi]], MINI ? mini_m : k) & 0xFFFFu) : 0u;
                    if (act) hp8[rl * (uint64_t)maxocc + fi] = (uint8_t)f;
                    const uint32_t b9 = __ballot_sync(0xffffffffu, (f >> 8) & 1u);
                    if (lane == 0) h9s[f0 >> 5] = b9;
                }
            }
            if (emitted > maxocc) { if (lane == 0) atomicOr(err, ERRF_LIST_OVERFLOW); emitted = maxocc; }
            if (COMPACT) {
                __syncwarp();
                if ((uint32_t)lane < hwords) h9w[rl * (uint64_t)hwords + lane] = h9s[lane];
                if (emitted && !last_fresh) emitted |= 0x80000000u;     // a duplicate insert call follows the last fresh one
            }
        } else if (COMPACT) {
            if ((uint32_t)lane < hwords) h9w[rl * (uint64_t)hwords + lane] = 0;
        }
        if (lane == 0) {
            nocc[rl] = emitted; nfresh[rl] = nfr; flags[r] = fl;
            if (slow) slow_list[atomicAdd(slow_n, 1u)] = (uint32_t)rl;
        }
    }
}

// ================================================================= readid_order
// hashbrown RawTable emulation: buckets 4,8,16,..; capacity b-1 (b<8) or
// b/8*7; a HashSet::insert with growth_left == 0 resizes (before the duplicate check when
// `rbf`, i.e. hashbrown >= 0.14); resize re-inserts the old table in ascending bucket order;
// insert slot = first EMPTY in the `gw`-wide group at hash&mask, else triangular probing.
// Occupancy bitmaps (bit s = bucket s is full) make the insert-slot search a constant-time
// funnel-shift + ffs instead of a byte-by-byte scan whose length (and the warp's, which pays the
// maximum over its 32 reads) explodes as the load factor approaches 7/8.
__device__ __forceinline__ uint32_t hb_window(const uint32_t* bm, uint32_t nb, uint32_t pos) {
    // 32 occupancy bits starting at bucket `pos`, cyclic in the nb-bucket table
    if (nb >= 32) {
        const uint32_t nw = nb >> 5, w = pos >> 5;
        return __funnelshift_r(bm[w], bm[(w + 1) & (nw - 1)], pos & 31);
    }
    const uint32_t x = bm[0];                 // tables below 32 buckets keep the pattern replicated (hb_set)
    return __funnelshift_r(x, x, pos);
}
__device__ __forceinline__ uint32_t hb_probe(const uint32_t* bm, uint32_t nb, uint32_t pos, uint32_t gw) {
    const uint32_t mask = nb - 1;
    pos &= mask;
    const uint32_t width = nb < gw ? nb : gw;       // small tables: cyclic linear probe over all buckets
    const uint32_t wmask = (1u << width) - 1;
    uint32_t stride = 0;
    for (;;) {
        const uint32_t free_ = ~hb_window(bm, nb, pos) & wmask;
        if (free_) return (pos + (uint32_t)__ffs(free_) - 1) & mask;
        stride += gw;
        pos = (pos + stride) & mask;
    }
}
__device__ __forceinline__ void hb_set(uint32_t* bm, uint32_t nb, uint32_t slot) {
    if (nb >= 32) bm[slot >> 5] |= 1u << (slot & 31);
    else bm[0] |= (nb == 4 ? 0x11111111u : nb == 8 ? 0x01010101u : 0x00010001u) << slot;
}
__device__ __forceinline__ uint32_t hb_cap(uint32_t nb) { return nb == 0 ? 0 : (nb < 8 ? nb - 1 : nb / 8 * 7); }

// Per thread in shared memory: the two ping-pong tables (TB + TB/2 entries of E = index of the
// insert call), their occupancy bitmaps and the low hash bits of every call, so the dependent
// probe chain never leaves the SM.  Control flow is a two-state machine (REHASH one old bucket |
// consume one insert call) so that the 32 reads of a warp stay converged even though they
// resize at different times.
template <typename E>
__global__ void __launch_bounds__(32)
readid_order_kernel(const uint32_t* __restrict__ entries, const uint32_t* __restrict__ nocc,
                    const uint32_t* __restrict__ perm, uint64_t nreads, uint32_t maxocc, uint32_t TB, uint32_t gw, uint32_t rbf, uint16_t* __restrict__ order,
                    uint32_t* __restrict__ n_set_out, uint64_t r0) {
    extern __shared__ __align__(16) uint8_t dsm[];
    const uint64_t ridx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = ridx < nreads;
    const uint64_t rl = live ? perm[ridx] : 0;
    // u8 tables (TB <= 512): 9 hash bits as hp (8) + h9 (1); u16 tables: 16 hash bits in hp
    constexpr bool kSmall = sizeof(E) == 1;
    const uint32_t hwords = kSmall ? (maxocc + 31) / 32 : 0;
    const uint32_t bwX = TB >= 32 ? TB / 32 : 1, bwY = TB >= 64 ? TB / 64 : 1;
    const size_t tab_bytes = ((size_t)(TB + TB / 2) * sizeof(E) + (size_t)maxocc * sizeof(E) + 3) & ~(size_t)3;
    const size_t per_thread = tab_bytes + (size_t)(hwords + bwX + bwY) * 4;
    uint8_t* mine = dsm + (size_t)threadIdx.x * per_thread;
    E* X = (E*)mine;
    E* Y = X + TB;
    E* hp = Y + TB / 2;
    uint32_t* h9 = (uint32_t*)(mine + tab_bytes);
    uint32_t* bmX = h9 + hwords;
    uint32_t* bmY = bmX + bwX;
    for (uint32_t i = 0; i < hwords + bwX + bwY; i++) h9[i] = 0;
    const uint32_t* row = entries + (live ? rl : 0) * (uint64_t)maxocc;
    const uint32_t n = live ? nocc[rl] : 0;
    uint32_t nb = 0, items = 0, growth = 0;
    uint32_t nb_old = 0, s = 0;          // rehash scan state: active while s < nb_old
    E* cur = X; uint32_t* bcur = bmX;
    E* old = X; uint32_t* bold = bmX;
    // a table of size sz lives in X when log2(TB/sz) is even, else in Y
    auto in_y = [&](uint32_t sz) -> bool { return ((__ffs(TB) - __ffs(sz)) & 1) != 0; };
    auto hash_of = [&](uint32_t j) -> uint32_t {
        if (kSmall) return (uint32_t)hp[j] | (((h9[j >> 5] >> (j & 31)) & 1u) << 8);
        return (uint32_t)hp[j];
    };
    uint32_t j = 0;
    uint4 buf = n ? __ldg((const uint4*)row) : make_uint4(0, 0, 0, 0);          // entries j&~3 .. +3
    uint4 nxt = n > 4 ? __ldg((const uint4*)(row + 4)) : make_uint4(0, 0, 0, 0);  // prefetched next group
    // Each iteration picks at most one item to place (an old bucket during a rehash, otherwise the
    // next insert call) and then runs ONE shared probe+store, so lanes in different phases do not
    // serialise two copies of the expensive part.
    while (j < n || s < nb_old) {
        uint32_t item = 0, hash = 0;
        bool place = false;
        if (s < nb_old) {
            // REHASH: next full old bucket (resize re-inserts in ascending bucket order)
            const uint32_t w = bold[s >> 5] >> (s & 31);     // old tables below 32 buckets: replicated, harmless
            if (w == 0) s = (s | 31) + 1;
            else {
                s += __ffs(w) - 1;
                if (s < nb_old) {            // (a replica bit of a small table can point past its end)
                    item = old[s];
                    hash = hash_of(item);
                    place = true;
                    s++;
                }
            }
        } else {
            const uint32_t q = j & 3;
            const uint32_t e = q == 0 ? buf.x : q == 1 ? buf.y : q == 2 ? buf.z : buf.w;
            const bool fresh = ENT_FRESH(e);
            if (growth == 0 && (rbf || fresh)) {
                // reserve_rehash -> resize(capacity_to_buckets(max(items+1, cap+1))): buckets double;
                // the same insert call is retried once the rehash has drained
                old = cur; bold = bcur; nb_old = nb; s = 0;
                nb = nb == 0 ? 4 : nb * 2;
                const bool y = in_y(nb);
                cur = y ? Y : X; bcur = y ? bmY : bmX;
                for (uint32_t i = 0; i < (nb >= 32 ? nb / 32 : 1u); i++) bcur[i] = 0;
                growth = hb_cap(nb) - items;
            } else {
                if (fresh) {
                    hp[j] = (E)e;
                    if (kSmall) h9[j >> 5] |= ((e >> 8) & 1u) << (j & 31);
                    item = j; hash = e & 0xFFFFu; place = true;
                    items++; growth--;
                }
                j++;
                if ((j & 3) == 0) {
                    buf = nxt;
                    if (j + 4 < n) nxt = __ldg((const uint4*)(row + j + 4));
                }
            }
        }
        if (place) {
            const uint32_t slot = hb_probe(bcur, nb, hash, gw);
            cur[slot] = (E)item;
            hb_set(bcur, nb, slot);
        }
    }
    if (!live) return;
    uint16_t* out = order + rl * (uint64_t)maxocc;
    uint32_t c = 0;
    for (uint32_t t = 0; t < nb; t++)
        if ((bcur[t >> 5] >> (t & 31)) & 1u) out[c++] = (uint16_t)((__ldg(row + cur[t]) >> 16) & 0x7FFu);
    n_set_out[r0 + rl] = c;
}

// ================================================================= read schedule (counting sort by set size)
// The 32 reads of a warp walk the same growth ladder only if they hold about the same number of
// distinct k-mers; quality masking and overlapping mates spread that number widely, so reads are
// dealt to the order kernel in descending order of their set size (a counting sort: histogram,
// one-block scan, scatter).  Pure scheduling: results do not depend on the permutation.
constexpr int SCHED_BINS = 1024;
__global__ void __launch_bounds__(256)
sched_hist_kernel(const uint32_t* __restrict__ key, uint64_t n, uint32_t* __restrict__ bins) {
    __shared__ uint32_t sh[SCHED_BINS];
    for (int i = threadIdx.x; i < SCHED_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&sh[min(key[i], (uint32_t)SCHED_BINS - 1)], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < SCHED_BINS; i += blockDim.x) if (sh[i]) atomicAdd(&bins[i], sh[i]);
}
__global__ void __launch_bounds__(SCHED_BINS)
sched_scan_kernel(uint32_t* __restrict__ bins) {     // bins[b] <- number of reads with a LARGER key (descending order)
    __shared__ uint32_t sh[SCHED_BINS];
    const int t = threadIdx.x;
    sh[t] = bins[SCHED_BINS - 1 - t];
    __syncthreads();
    for (int o = 1; o < SCHED_BINS; o <<= 1) {
        uint32_t v = t >= o ? sh[t - o] : 0u;
        __syncthreads();
        sh[t] += v;
        __syncthreads();
    }
    bins[SCHED_BINS - 1 - t] = t ? sh[t - 1] : 0u;
}
__global__ void __launch_bounds__(256)
sched_scatter_kernel(const uint32_t* __restrict__ key, uint64_t n, uint32_t* __restrict__ cursor, uint32_t* __restrict__ perm) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    perm[atomicAdd(&cursor[min(key[i], (uint32_t)SCHED_BINS - 1)], 1u)] = (uint32_t)i;
}

// ================================================================= readid_order, small reads (<= 255 k-mer positions)
// Same emulation as above, re-engineered for throughput (it was 43 % of a read_id step):
//  * input is the compact list of DISTINCT k-mers (9 hash bits each).  Duplicate insert calls only
//    matter through hashbrown >= 0.14's reserve-before-find: a duplicate arriving when the table is
//    exactly full triggers a resize that a fresh key would have triggered anyway -- unless no fresh
//    key follows, so one flag ("a call follows the last fresh one") carries all of it;
//  * every per-read array is interleaved across the warp at 4-byte granularity (byte i of lane t
//    lives at (i/4)*128 + 4t + i%4), so each data-dependent access hits bank == lane: no conflicts;
//  * shared memory is addressed with explicit 32-bit shared-window addresses;
//  * the item to place is fetched one step ahead (next full bucket of the old table during a
//    resize, else the next fresh k-mer) and its hash loads are consumed a step later, so the fetch
//    chain and the probe chain overlap; the occupancy window loaded by the probe is reused for the
//    bitmap update;
//  * reads arrive sorted by set size and are dealt to three instantiations (128/256/512 buckets
//    at most), so a warp's reads walk the same growth ladder and small sets leave room for more
//    resident warps;
//  * output is the index of the distinct k-mer per occupied bucket (u8); the vote kernel looks its
//    position up in ent16.
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// byte i / word w of a lane-interleaved array whose lane-0 element 0 sits at shared address `a`
__device__ __forceinline__ uint32_t il_b(uint32_t a, uint32_t i) { return a + i + (i >> 2) * 124u; }
__device__ __forceinline__ uint32_t il_w(uint32_t a, uint32_t w) { return a + w * 128u; }

template <int TB> struct OrderSmallLayout {       // words per lane
    static constexpr int X = TB / 4, Y = TB / 8, HP = TB > 256 ? 64 : TB / 32 * 7, H9 = TB > 256 ? 8 : 0, BX = TB / 32, BY = TB / 64;
    static constexpr int oX = 0, oY = oX + X, oHP = oY + Y, oH9 = oHP + HP, oBX = oH9 + H9, oBY = oBX + BX, TOTAL = oBY + BY;
    static constexpr int MAXF = TB > 256 ? 255 : TB / 8 * 7 - 1;      // largest set that never outgrows TB buckets
};

template <int TB>
__global__ void __launch_bounds__(32)
readid_order_small_kernel(const uint8_t* __restrict__ hp8, const uint32_t* __restrict__ h9w,
                          const uint32_t* __restrict__ nocc, const uint32_t* __restrict__ nfresh,
                          const uint32_t* __restrict__ perm, uint64_t nreads, uint32_t maxocc, uint32_t gw, uint32_t rbf,
                          uint8_t* __restrict__ order8, uint32_t* __restrict__ n_set_out, uint64_t r0) {
    using Lay = OrderSmallLayout<TB>;
    extern __shared__ __align__(16) uint32_t sm32[];
    const uint32_t lane = threadIdx.x;
    const uint64_t idx0 = (uint64_t)blockIdx.x * 32;
    // reads are sorted by set size (descending): the block's first read decides which instantiation owns it
    {
        const uint32_t fmax = nfresh[perm[idx0]];
        const int cls = fmax <= (uint32_t)OrderSmallLayout<128>::MAXF ? 128 : fmax <= (uint32_t)OrderSmallLayout<256>::MAXF ? 256 : 512;
        if (cls != TB) return;
    }
    const bool live = idx0 + lane < nreads;
    const uint64_t rl = live ? perm[idx0 + lane] : 0;
    const uint32_t F = live ? min(nfresh[rl], maxocc) : 0;
    const bool tail_dup = live && (nocc[rl] >> 31);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm32) + lane * 4u;
    const uint32_t aX = il_w(sbase, Lay::oX), aY = il_w(sbase, Lay::oY), aHP = il_w(sbase, Lay::oHP),
                   aH9 = il_w(sbase, Lay::oH9), aBX = il_w(sbase, Lay::oBX), aBY = il_w(sbase, Lay::oBY);
    // stage the hash bits of this lane's read
    {
        const uint32_t* src = (const uint32_t*)(hp8 + rl * (uint64_t)maxocc);     // maxocc is a multiple of 4
        const uint32_t nwords = (F + 3) / 4;
#pragma unroll 4
        for (uint32_t w = 0; w < nwords; w++) sts32(il_w(aHP, w), __ldg(src + w));
        if (TB > 256) {
            const uint32_t hwords = (maxocc + 31) / 32;
            for (uint32_t w = 0; w < 8; w++) sts32(il_w(aH9, w), w < hwords ? __ldg(h9w + rl * (uint64_t)hwords + w) : 0u);
        }
    }
    uint32_t nb = 0, growth = 0, f = 0;
    uint32_t nb_old = 0, s = 0, wcur = 0;          // resize scan: buckets [s, s+32) of the old table left in wcur
    bool scanning = false, final_done = false;
    uint32_t aCur = aX, aBcur = aBX, aOld = aX, aBold = aBX;
    bool p_valid = false;                          // pending placement, fetched in the previous step
    uint32_t p_item = 0, p_h8 = 0, p_h9 = 0, p_tab = aX, p_bm = aBX, p_nb = 4;
    for (;;) {
        // ---- fetch the next item to place ----------------------------------------------------
        bool f_valid = false, more = true;
        uint32_t f_item = 0;
        if (scanning) {
            if (wcur) {
                const uint32_t b = __ffs(wcur) - 1;
                wcur &= wcur - 1;
                f_item = lds8(il_b(aOld, s + b));
                f_valid = true;
            } else {
                s += 32;
                if (s < nb_old) {
                    wcur = lds32(il_w(aBold, s >> 5));
                    if (nb_old < 32) wcur &= (1u << nb_old) - 1;      // small tables keep their pattern replicated
                } else scanning = false;
            }
        } else if (f < F ? growth == 0 : (rbf && tail_dup && growth == 0 && F > 0 && !final_done)) {
            // reserve_rehash -> resize: buckets double; the insert call is retried after the old table drained
            final_done = f >= F;
            aOld = aCur; aBold = aBcur; nb_old = nb; s = 0;
            nb = nb == 0 ? 4 : nb * 2;
            // sizes TB, TB/4, .. live in X, the others in Y
            const bool y = ((__ffs(TB) - __ffs(nb)) & 1) != 0;
            aCur = y ? aY : aX; aBcur = y ? aBY : aBX;
            for (uint32_t i = 0; i < (nb >= 32 ? nb / 32 : 1u); i++) sts32(il_w(aBcur, i), 0u);
            growth = (nb < 8 ? nb - 1 : nb / 8 * 7) - f;
            // the pending item below still goes into the old table: its bitmap is read from the next step on
            if (nb_old) { scanning = true; wcur = 0; s = 0u - 32u; }
        } else if (f < F) {
            f_item = f++; growth--;
            f_valid = true;
        } else more = false;
        uint32_t f_h8 = 0, f_h9 = 0;
        if (f_valid) {                                   // loads only; consumed when the item is placed
            f_h8 = lds8(il_b(aHP, f_item));
            if (TB > 256) f_h9 = lds32(il_w(aH9, f_item >> 5));
        }
        // ---- place the item fetched one step earlier (into the table that was current then) ----
        if (p_valid) {
            uint32_t hash = p_h8;
            if (TB > 256) hash |= ((p_h9 >> (p_item & 31)) & 1u) << 8;
            const uint32_t mask = p_nb - 1, nw = p_nb >= 32 ? p_nb >> 5 : 1;
            const uint32_t width = p_nb < gw ? p_nb : gw, wmask = (1u << width) - 1;
            uint32_t pos = hash & mask, stride = 0, slot, lo, hi, w0;
            for (;;) {
                w0 = pos >> 5;
                lo = lds32(il_w(p_bm, w0));
                hi = lds32(il_w(p_bm, (w0 + 1) & (nw - 1)));
                const uint32_t free_ = ~__funnelshift_r(lo, hi, pos & 31) & wmask;
                if (free_) { slot = (pos + (uint32_t)__ffs(free_) - 1) & mask; break; }
                stride += gw;
                pos = (pos + stride) & mask;
            }
            sts8(il_b(p_tab, slot), p_item);
            if (p_nb >= 32) {
                const uint32_t ws = slot >> 5;
                sts32(il_w(p_bm, ws), (ws == w0 ? lo : hi) | (1u << (slot & 31)));
            } else {
                sts32(p_bm, lo | ((p_nb == 4 ? 0x11111111u : p_nb == 8 ? 0x01010101u : 0x00010001u) << slot));
            }
        }
        p_valid = f_valid; p_item = f_item; p_h8 = f_h8; p_h9 = f_h9; p_tab = aCur; p_bm = aBcur; p_nb = nb;
        if (!more && !p_valid) break;
    }
    if (!live) return;
    // occupied buckets in ascending order = HashSet iteration order
    uint8_t* out = order8 + rl * (uint64_t)maxocc;
    uint32_t c = 0, pack = 0;
    for (uint32_t t0 = 0; t0 < nb; t0 += 32) {
        uint32_t w = lds32(il_w(aBcur, t0 >> 5));
        if (nb < 32) w &= (1u << nb) - 1;
        while (w) {
            const uint32_t t = t0 + __ffs(w) - 1;
            w &= w - 1;
            pack |= lds8(il_b(aCur, t)) << (8 * (c & 3));
            c++;
            if ((c & 3) == 0) { *(uint32_t*)(out + c - 4) = pack; pack = 0; }
        }
    }
    if (c & 3) *(uint32_t*)(out + (c & ~3u)) = pack;
    n_set_out[r0 + rl] = c;
}

// ================================================================= readid_vote (rows of <= 64 accessions)
template <int WP, bool STEPS>      // STEPS: report colours carry their insertion step (column-sharded read_id); a separate
__global__ void __launch_bounds__(RA_WARPS * 32, 9)   // instantiation so that the replicated-index kernel stays as it was
readid_vote_narrow_kernel(const ReadSrc src,
                          const uint64_t* __restrict__ seq_offs, const uint64_t* __restrict__ read_offs, uint64_t r0,
                          uint64_t nreads, uint32_t k, uint32_t H, ModS mods, const uint32_t* __restrict__ rows,
                          const uint32_t* __restrict__ rownz, const uint32_t* __restrict__ rownz_all, uint32_t N, int cap,
                          uint32_t maxocc, const uint16_t* __restrict__ order, const uint8_t* __restrict__ order8,
                          const uint16_t* __restrict__ ent16, const uint32_t* __restrict__ n_set,
                          uint32_t start_sample, uint32_t rep_cap, uint32_t* __restrict__ flags,
                          uint32_t* __restrict__ rep_n, uint32_t* __restrict__ rep_colour,
                          uint32_t* __restrict__ rep_count, unsigned long long* __restrict__ gather_counter,
                          const uint32_t* __restrict__ sel_list, const uint32_t* __restrict__ sel_n,
                          const uint32_t* __restrict__ sel_all) {
    // sel_list != nullptr (after the partitioned vote, cid_readid_part.cu): only the reads it left over -- sel_list[0 .. *sel_n)
    // -- or, when one of its buckets overflowed (*sel_all != 0), every read of the chunk
    extern __shared__ __align__(16) uint8_t dsm[];
    __shared__ uint32_t lut[256];
    lut4_init(lut, threadIdx.x, blockDim.x);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t my_gather_rows = 0;      // matrix rows this warp read (diagnostic: bench's random-access rate)
    const size_t per_warp = ((tile_smem_bytes(cap) + 15) & ~(size_t)15) + 64 + 128 + (MAX_MATES + 1) * 4 + 12;
    uint8_t* base = dsm + (size_t)warp * ((per_warp + 15) & ~(size_t)15);
    Tile t = tile_carve(base, cap);
    uint8_t* ord = base + ((tile_smem_bytes(cap) + 15) & ~(size_t)15);
    uint16_t* ordstep = (uint16_t*)(ord + 64);      // with_steps: index of the k-mer that inserted the colour (REP_STEP_SHIFT)
    uint32_t* moffs = (uint32_t*)(ord + 64 + 128);
    const bool classic = start_sample == 0;
    const bool listed = sel_list != nullptr && *sel_all == 0u;
    const uint64_t nwork = listed ? (uint64_t)*sel_n : nreads;

    for (uint64_t wi = (uint64_t)blockIdx.x * RA_WARPS + warp; wi < nwork; wi += (uint64_t)gridDim.x * RA_WARPS) {
        const uint64_t rl = listed ? (uint64_t)sel_list[wi] : wi;
        const uint64_t r = r0 + rl;
        const uint32_t n = n_set[r];
        __syncwarp();
        if (n == 0) { if (lane == 0) rep_n[r] = 0; continue; }
        ReadGeom g;
        warp_load_read(t, cap, src, r, seq_offs, __ldg(read_offs + r), __ldg(read_offs + r + 1), moffs, lane, g);
        const uint16_t* ordrow = order + rl * (uint64_t)maxocc;
        const uint8_t* ord8row = order8 + rl * (uint64_t)maxocc;
        const uint16_t* entrow = ent16 + rl * (uint64_t)maxocc;
        uint32_t cand0 = 0, cand1 = 0, cnt0 = 0, cnt1 = 0, nrep = 0, nproc = 0;
        bool miss = false;
        // The first FINE_STEPS steps take only 32/H k-mers with one lane per (k-mer, hash row): most
        // reads with a sequencing error or off-target origin hit an absent row within a few k-mers,
        // and every gather is a 64-byte DRAM fetch for an 8-byte row, so gathers past the first
        // miss are pure waste.  Later steps take 32 k-mers, one lane each.
        const bool can_fine = (H == 1 || H == 2 || H == 4 || H == 8);
        uint32_t fine_left = can_fine ? FINE_STEPS : 0;
        for (uint32_t c0 = 0; c0 < n && !miss;) {
            // Once the candidate set is final (c0 >= start_sample) and EMPTY, no colour can ever be counted
            // (read_id_mt_pe.rs:155-161): all that is left to learn is where the first absent row is, and the
            // row-present bitmap (S/8 bytes, L2 resident) answers that without touching the matrix in DRAM.
            const bool presence_only = !classic && c0 >= start_sample && (cand0 | cand1) == 0u;
            const bool fine = fine_left > 0 && !presence_only;
            const uint32_t csize = fine ? 32u / H : 32u;
            if (fine_left) fine_left--;
            const uint32_t idx = c0 + lane;
            const bool active = idx < n && (uint32_t)lane < csize;
            uint32_t x0 = 0, x1 = 0;
            bool m = false, skipped_rows = false;
            HashIn in;
            in.w0 = in.w1 = in.w2 = in.w3 = 0;
            if (active) {
                // k-mer `idx` of the iteration order: tile position + strand
                uint32_t e = order8 ? (uint32_t)__ldg(entrow + __ldcs(ord8row + idx)) : (uint32_t)__ldcs(ordrow + idx);
                uint32_t tp = e & 0x3FFu;
                uint64_t f = codes_window(t.codes, (int)tp, k);
                uint64_t key = ((e >> 10) & 1u) ? f : revcomp_key(f, k);
                in = hashin_from_key(lut, key, k);
            }
            if (fine) {
                // lane L serves k-mer L/H, hash L%H
                const int src = lane / (int)H;
                const uint32_t h = (uint32_t)lane % H;
                HashIn mine;
                mine.w0 = __shfl_sync(0xffffffffu, in.w0, src);
                mine.w1 = __shfl_sync(0xffffffffu, in.w1, src);
                mine.w2 = __shfl_sync(0xffffffffu, in.w2, src);
                mine.w3 = __shfl_sync(0xffffffffu, in.w3, src);
                const bool act2 = c0 + (uint32_t)src < n;
                uint32_t a = 0xFFFFFFFFu, b = WP == 2 ? 0xFFFFFFFFu : 0u;
                bool mm = false;
                if (act2) {
                    uint64_t rid = hash_row(mine, k, h, mods);
                    if (WP == 2) { uint2 v = __ldg((const uint2*)(rows + rid * 2)); a = v.x; b = v.y; }
                    else { a = __ldg(rows + rid); b = 0; }
                    mm = rownz ? !((__ldg(rownz + (rid >> 5)) >> (rid & 31)) & 1u) : ((a | b) == 0u);
                }
                // AND / OR across the H lanes of a k-mer
                for (uint32_t o = 1; o < H; o <<= 1) {
                    a &= __shfl_xor_sync(0xffffffffu, a, o);
                    b &= __shfl_xor_sync(0xffffffffu, b, o);
                    mm = __shfl_xor_sync(0xffffffffu, (int)mm, o) || mm;
                }
                // compact: lane j takes k-mer j's result from lane j*H
                const int from = (lane * (int)H) & 31;
                uint32_t ca = __shfl_sync(0xffffffffu, a, from), cb2 = __shfl_sync(0xffffffffu, b, from);
                int cm = __shfl_sync(0xffffffffu, (int)mm, from);
                if (active) { x0 = ca; x1 = cb2; m = cm != 0; }
            } else if (active && presence_only) {
                for (uint32_t h = 0; h < H; h++) {
                    const uint64_t rid = hash_row(in, k, h, mods);
                    if (!((__ldg(rownz_all + (rid >> 5)) >> (rid & 31)) & 1u)) { m = true; break; }
                }
            } else if (active) {
                // Row 0 first; the other H-1 rows are fetched (all in flight together) only while some candidate
                // colour can still be counted for this k-mer.  Once the candidate set is final, a k-mer whose
                // partial AND has lost every candidate bit contributes nothing (read_id_mt_pe.rs:155-161) and only
                // the PRESENCE of its remaining rows matters, which the L2-resident bitmap answers.
                const bool cand_final = !classic && c0 >= start_sample;
                uint32_t rid[MAX_HASH];
#pragma unroll
                for (int h = 0; h < MAX_HASH; h++) rid[h] = (uint32_t)h < H ? (uint32_t)hash_row(in, k, h, mods) : 0u;
                {
                    uint32_t a, b = 0;
                    if (WP == 2) { uint2 v = __ldg((const uint2*)(rows + (size_t)rid[0] * 2)); a = v.x; b = v.y; }
                    else a = __ldg(rows + rid[0]);
                    const bool present = rownz ? ((__ldg(rownz + (rid[0] >> 5)) >> (rid[0] & 31)) & 1u) : ((a | b) != 0u);
                    if (!present) m = true;     // read_id_mt_pe.rs:121-123 `None => break`
                    x0 = a; x1 = b;
                }
                if (cand_final) { x0 &= cand0; x1 &= cand1; }
                if (!cand_final || (x0 | x1) != 0u) {
                    uint32_t ra[MAX_HASH], rb[MAX_HASH];
#pragma unroll
                    for (int h = 1; h < MAX_HASH; h++) {
                        ra[h] = 0xFFFFFFFFu; rb[h] = 0xFFFFFFFFu;
                        if ((uint32_t)h < H) {
                            if (WP == 2) { uint2 v = __ldg((const uint2*)(rows + (size_t)rid[h] * 2)); ra[h] = v.x; rb[h] = v.y; }
                            else { ra[h] = __ldg(rows + rid[h]); rb[h] = 0; }
                        }
                    }
#pragma unroll
                    for (int h = 1; h < MAX_HASH; h++) {
                        if ((uint32_t)h < H) {
                            const bool present = rownz ? ((__ldg(rownz + (rid[h] >> 5)) >> (rid[h] & 31)) & 1u) : ((ra[h] | rb[h]) != 0u);
                            if (!present) m = true;
                            x0 &= ra[h]; x1 &= rb[h];
                        }
                    }
                    if (WP == 1) x1 = 0;
                } else {
#pragma unroll
                    for (int h = 1; h < MAX_HASH; h++)
                        if ((uint32_t)h < H && !((__ldg(rownz_all + (rid[h] >> 5)) >> (rid[h] & 31)) & 1u)) m = true;
                    skipped_rows = true;
                }
            }
            const uint32_t missmask = __ballot_sync(0xffffffffu, active && m);
            const uint32_t p_local = missmask ? (uint32_t)(__ffs(missmask) - 1) : csize;
            const bool valid = active && (uint32_t)lane < p_local;
            // candidate colours (and report insertion order) from the first start_sample k-mers
            uint32_t lim = classic ? csize : (start_sample > c0 ? min(csize, start_sample - c0) : 0u);
            lim = min(lim, min(p_local, n - c0));
            for (uint32_t jj = 0; jj < lim; jj++) {
                uint32_t y0 = __shfl_sync(0xffffffffu, x0, jj), y1 = __shfl_sync(0xffffffffu, x1, jj);
                uint32_t nw0 = y0 & ~cand0, nw1 = y1 & ~cand1;
                while (nw0) { uint32_t b = __ffs(nw0) - 1; nw0 &= nw0 - 1; if (lane == 0 && nrep < 64) { ord[nrep] = (uint8_t)b; if (STEPS) ordstep[nrep] = (uint16_t)(c0 + jj); } nrep++; }
                while (nw1) { uint32_t b = __ffs(nw1) - 1; nw1 &= nw1 - 1; if (lane == 0 && nrep < 64) { ord[nrep] = (uint8_t)(32 + b); if (STEPS) ordstep[nrep] = (uint16_t)(c0 + jj); } nrep++; }
                cand0 |= y0; cand1 |= y1;
            }
            // per-colour counts over the k-mers before the first miss (colours outside cand never count)
            const uint32_t z0 = valid ? (x0 & cand0) : 0u, z1 = valid ? (x1 & cand1) : 0u;
            for (uint32_t cm = cand0; cm; cm &= cm - 1) {
                uint32_t b = __ffs(cm) - 1;
                uint32_t v = __popc(__ballot_sync(0xffffffffu, (z0 >> b) & 1u));
                if ((uint32_t)lane == b) cnt0 += v;
            }
            if (WP == 2)
                for (uint32_t cm = cand1; cm; cm &= cm - 1) {
                    uint32_t b = __ffs(cm) - 1;
                    uint32_t v = __popc(__ballot_sync(0xffffffffu, (z1 >> b) & 1u));
                    if ((uint32_t)lane == b) cnt1 += v;
                }
            const uint32_t took = missmask ? p_local + 1 : min(csize, n - c0);
            nproc += took;
            // diagnostic: row reads in units of 1/H k-mer (a k-mer that stopped after row 0 costs one row)
            if (!presence_only) {
                const uint32_t act = __popc(__ballot_sync(0xffffffffu, active)), skp = __popc(__ballot_sync(0xffffffffu, skipped_rows));
                my_gather_rows += (act - skp) * H + skp;
            }
            if (missmask) miss = true;
            c0 += csize;
        }
        __syncwarp();
        // report in final_report insertion order; the "no hit" key N goes last (inserted at the break)
        const uint32_t total = nrep + (miss ? 1u : 0u);
        uint32_t* rc = rep_colour + r * (uint64_t)rep_cap;
        uint32_t* rv = rep_count + r * (uint64_t)rep_cap;
        for (uint32_t i0 = 0; i0 < nrep; i0 += 32) {
            uint32_t i = i0 + lane;
            uint32_t colour = i < nrep ? ord[i] : 0u;
            uint32_t v0 = __shfl_sync(0xffffffffu, cnt0, colour & 31), v1 = __shfl_sync(0xffffffffu, cnt1, colour & 31);
            if (i < nrep && i < rep_cap) { rc[i] = STEPS ? (colour | ((uint32_t)ordstep[i] << REP_STEP_SHIFT)) : colour; rv[i] = colour < 32 ? v0 : v1; }
        }
        if (lane == 0) {
            if (miss && nrep < rep_cap) { rc[nrep] = N; rv[nrep] = 1; }
            rep_n[r] = min(total, rep_cap);
            // bits 8..23: k-mers whose rows were needed (up to and including the first miss)
            flags[r] |= (total > rep_cap ? 4u : 0u) | (min(nproc, 0xFFFFu) << 8);
        }
    }
    if (lane == 0 && my_gather_rows && gather_counter) atomicAdd(gather_counter, (unsigned long long)my_gather_rows);
}

// ================================================================= readid_vote (any row width)
// One k-mer at a time per warp: lanes cover the row's words (coalesced), counts live in
// bit-sliced registers per lane; used when a row has more than 64 accessions.
constexpr int RV_MAXWPL = 4;       // words per lane: rows up to 128 words (4,096 accessions per shard)
constexpr int RV_PLANES = 11;      // counts up to 2047 k-mers per read
// carry-save adder: (carry, sum) = a + b + c, bitwise (32 accessions per instruction)
__device__ __forceinline__ void csa3(uint32_t& carry, uint32_t& sum, uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t u = a ^ b;
    carry = (a & b) | (u & c);
    sum = u ^ c;
}
template <int WPL, int HT>         // words per lane (1, 2 or 4): 1,024 / 2,048 / 4,096 accessions; sizes the bit-sliced counters.
__global__ void __launch_bounds__(RA_WARPS * 32, (WPL == 1 && HT != 0) ? 8 : 1)      // HT: compile-time num_hash (2 or 4; 0 = run-time, predicated up to MAX_HASH)
readid_vote_wide_kernel(const ReadSrc src,
                        const uint64_t* __restrict__ seq_offs, const uint64_t* __restrict__ read_offs, uint64_t r0,
                        uint64_t nreads, uint32_t k, uint32_t H, ModS mods, const uint32_t* __restrict__ rows,
                        const uint32_t* __restrict__ rownz, uint32_t N, uint32_t Wp, int cap, uint32_t maxocc,
                        const uint16_t* __restrict__ order, const uint8_t* __restrict__ order8,
                        const uint16_t* __restrict__ ent16, const uint32_t* __restrict__ n_set, uint32_t start_sample,
                        uint32_t rep_cap, uint32_t* __restrict__ flags, uint32_t* __restrict__ rep_n,
                        uint32_t* __restrict__ rep_colour, uint32_t* __restrict__ rep_count, uint32_t with_steps) {
    extern __shared__ __align__(16) uint8_t dsm[];
    __shared__ uint32_t lut[256];
    lut4_init(lut, threadIdx.x, blockDim.x);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = ((tile_smem_bytes(cap) + 15) & ~(size_t)15) + (MAX_MATES + 1) * 4 + 12;
    uint8_t* base = dsm + (size_t)warp * ((per_warp + 15) & ~(size_t)15);
    Tile t = tile_carve(base, cap);
    uint32_t* moffs = (uint32_t*)(base + ((tile_smem_bytes(cap) + 15) & ~(size_t)15));
    const bool classic = start_sample == 0;
    constexpr int U = WPL == 4 ? 2 : 4;      // k-mers whose rows are in flight together (a step of one k-mer is a dependent DRAM round trip)
    constexpr int NH = HT ? HT : MAX_HASH;   // hash slots carried per k-mer

    for (uint64_t rl = (uint64_t)blockIdx.x * RA_WARPS + warp; rl < nreads; rl += (uint64_t)gridDim.x * RA_WARPS) {
        const uint64_t r = r0 + rl;
        const uint32_t n = n_set[r];
        __syncwarp();
        if (n == 0) { if (lane == 0) rep_n[r] = 0; continue; }
        ReadGeom g;
        warp_load_read(t, cap, src, r, seq_offs, __ldg(read_offs + r), __ldg(read_offs + r + 1), moffs, lane, g);
        const uint16_t* ordrow = order + rl * (uint64_t)maxocc;
        const uint8_t* ord8row = order8 + rl * (uint64_t)maxocc;
        const uint16_t* entrow = ent16 + rl * (uint64_t)maxocc;
        uint32_t* rc = rep_colour + r * (uint64_t)rep_cap;
        uint32_t* rv = rep_count + r * (uint64_t)rep_cap;
        uint32_t cand[WPL], pl[WPL][RV_PLANES];
#pragma unroll
        for (int w = 0; w < WPL; w++) { cand[w] = 0; for (int p = 0; p < RV_PLANES; p++) pl[w][p] = 0; }
        uint32_t nrep = 0, nproc = 0;
        bool miss = false;
        // 32 k-mers of the set order at a time: lane L hashes k-mer c0 + L (H row indices, row-present bits from the
        // L2-resident bitmap), a ballot finds the first absent row, and the rows of the k-mers before it are then gathered one
        // k-mer per step with every lane on its own words.  (Before: all 32 lanes hashed the same k-mer, ~600 redundant
        // instructions per k-mer, which bound the kernel.)
        // lane L's k-mer of the batch starting at `c`: row indices; the row-present words are only LOADED here (`pres`), they
        // are looked at after the gather steps of the batch before, so that their latency hides under those
        auto hash_batch = [&](uint32_t c, uint32_t (&rid_o)[NH], uint32_t (&pres)[NH]) {
#pragma unroll
            for (int h = 0; h < NH; h++) { rid_o[h] = 0; pres[h] = 0xFFFFFFFFu; }
            const uint32_t idx = c + lane;
            if (idx < n) {
                const uint32_t e = order8 ? (uint32_t)__ldg(entrow + __ldg(ord8row + idx)) : (uint32_t)__ldg(ordrow + idx);
                const uint64_t f = codes_window(t.codes, (int)(e & 0x3FFu), k);
                const uint64_t key = ((e >> 10) & 1u) ? f : revcomp_key(f, k);
                const HashIn in = hashin_from_key(lut, key, k);
#pragma unroll
                for (int h = 0; h < NH; h++)
                    if (HT || (uint32_t)h < H) {
                        rid_o[h] = (uint32_t)hash_row(in, k, h, mods);
                        pres[h] = __ldg(rownz + (rid_o[h] >> 5)) >> (rid_o[h] & 31);
                    }
            }
        };
        uint32_t rid_l[NH], pres_l[NH];
        hash_batch(0, rid_l, pres_l);
        for (uint32_t c0 = 0; c0 < n && !miss; c0 += 32) {
            const uint32_t batch = min(32u, n - c0);
            bool absent = false;
#pragma unroll
            for (int h = 0; h < NH; h++) if (!(pres_l[h] & 1u)) absent = true;
            uint32_t rid_n[NH], pres_n[NH];
            hash_batch(c0 + 32, rid_n, pres_n);           // (a no-op past the end of the set)
            const uint32_t missmask = __ballot_sync(0xffffffffu, absent);
            const uint32_t p_local = missmask ? (uint32_t)(__ffs(missmask) - 1) : batch;
            nproc += missmask ? p_local + 1 : batch;
            if (missmask) miss = true;
            for (uint32_t jj0 = 0; jj0 < p_local; jj0 += U) {
                // the AND rows of U k-mers: all their loads are issued before the first one is consumed
                uint32_t xs[U][WPL];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const bool on = jj0 + u < p_local;
                    uint64_t rid[NH];
#pragma unroll
                    for (int h = 0; h < NH; h++)
                        rid[h] = (HT || (uint32_t)h < H) ? (uint64_t)__shfl_sync(0xffffffffu, rid_l[h], min(jj0 + u, 31u)) : 0ull;
#pragma unroll
                    for (int w = 0; w < WPL; w++) {
                        const uint32_t col = w * 32 + lane;
                        uint32_t x = 0;
                        if (on && col < Wp) {
                            x = 0xFFFFFFFFu;
#pragma unroll
                            for (int h = 0; h < NH; h++)                  // (static indices: rid[] stays in registers)
                                if (HT || (uint32_t)h < H) x &= __ldg(rows + rid[h] * Wp + col);
                        }
                        xs[u][w] = x;
                    }
                }
                if (!classic && c0 + jj0 >= start_sample && jj0 + U <= p_local) {
                    // the candidate set is final and all U k-mers count: one carry-save tree per word instead of U ripple adds
#pragma unroll
                    for (int w = 0; w < WPL; w++) {
                        uint32_t carry;
                        if (U == 4) {
                            uint32_t t2a, t2b;
                            csa3(t2a, pl[w][0], pl[w][0], xs[0][w] & cand[w], xs[1][w] & cand[w]);
                            csa3(t2b, pl[w][0], pl[w][0], xs[2][w] & cand[w], xs[U - 1][w] & cand[w]);
                            csa3(carry, pl[w][1], pl[w][1], t2a, t2b);
                        } else {
                            csa3(carry, pl[w][0], pl[w][0], xs[0][w] & cand[w], xs[U - 1][w] & cand[w]);
                        }
#pragma unroll
                        for (int p = (U == 4 ? 2 : 1); p < RV_PLANES; p++) {
                            const uint32_t t2 = pl[w][p] & carry;
                            pl[w][p] ^= carry;
                            carry = t2;
                        }
                    }
                    continue;
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (jj0 + u >= p_local) break;
                    const uint32_t j = c0 + jj0 + u;
                    const bool seeding = classic || j < start_sample;
#pragma unroll
                    for (int w = 0; w < WPL; w++) {
                        const uint32_t x = xs[u][w];
                        if (seeding) {
                            // new colours in ascending order across the whole row: lanes take turns per word
                            uint32_t nw = x & ~cand[w];
                            for (int l = 0; l < 32; l++) {
                                uint32_t y = __shfl_sync(0xffffffffu, nw, l);
                                while (y) {
                                    uint32_t b = __ffs(y) - 1; y &= y - 1;
                                    if (lane == 0 && nrep < rep_cap) rc[nrep] = ((w * 32 + l) * 32 + b) | (with_steps ? (j << REP_STEP_SHIFT) : 0u);
                                    nrep++;
                                }
                            }
                            cand[w] |= x;
                        }
                        uint32_t carry = x & cand[w];
#pragma unroll
                        for (int p = 0; p < RV_PLANES; p++) {
                            uint32_t t2 = pl[w][p] & carry;
                            pl[w][p] ^= carry;
                            carry = t2;
                        }
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < NH; h++) { rid_l[h] = rid_n[h]; pres_l[h] = pres_n[h]; }
        }
        __syncwarp();
        // counts for the reported colours
        const uint32_t kept = min(nrep, rep_cap);
        for (uint32_t i = 0; i < kept; i++) {
            uint32_t colour = rc[i];             // written by lane 0 above
            colour = __shfl_sync(0xffffffffu, colour, 0) & ((1u << REP_STEP_SHIFT) - 1u);
            uint32_t word = colour >> 5, b = colour & 31, w = word >> 5, owner = word & 31;
            uint32_t v = 0;
#pragma unroll
            for (int ww = 0; ww < WPL; ww++)
                if ((uint32_t)ww == w)
                    for (int p = 0; p < RV_PLANES; p++) v |= ((pl[ww][p] >> b) & 1u) << p;
            v = __shfl_sync(0xffffffffu, v, owner);
            if (lane == 0) rv[i] = v;
        }
        if (lane == 0) {
            uint32_t total = nrep + (miss ? 1u : 0u);
            if (miss && nrep < rep_cap) { rc[nrep] = N; rv[nrep] = 1; }
            rep_n[r] = min(total, rep_cap);
            flags[r] |= (total > rep_cap ? 4u : 0u) | (min(nproc, 0xFFFFu) << 8);
        }
    }
}

// ================================================================= readid_classify (device side of kmer_poll_plus)
// read_id_mt_pe.rs:187-251 on the per-read reports, one thread per read.  Every comparison of the
// reference is reproduced in f64 (`hits < n*p`, `hits > n*p` are single IEEE multiplies, identical
// on host and device).  The Binomial pmf goes through device log/exp, which may differ from the
// host's libm in the last ulps, so a read whose pmf lies within 1e-6 (relative) of the threshold,
// and any read with a tie at the top (the reference joins the tied names in FnvHashMap iteration
// order), is NOT decided here: its report is appended to `list` and the host vote
// (cid_host_vote.cpp, the same code cid_classify_reads runs) classifies it.
__device__ __forceinline__ double dev_stirling_err(double n) {
    const double table[16] = {
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
__device__ __forceinline__ double dev_deviance_term(double x, double np) {
    if (fabs(x - np) < 0.1 * (x + np)) {
        double v = (x - np) / (x + np);
        double s = (x - np) * (x - np) / (x + np);
        double ej = 2.0 * x * v;
        for (int j = 1; j < 1000; j++) {
            ej *= v * v;
            double s1 = s + ej / (double)(2 * j + 1);
            if (s1 == s) return s1;
            s = s1;
        }
        return s;
    }
    return x * log(x / np) + np - x;
}
__device__ __forceinline__ double dev_binomial_pmf(double n, double p, double x) {
    if (p == 0.0) return x == 0.0 ? 1.0 : 0.0;
    if (p == 1.0) return x == n ? 1.0 : 0.0;
    const double q = 1.0 - p;
    if (x == 0.0) return exp(n * log(q));
    if (x == n) return exp(n * log(p));
    const double rest = n - x;
    const double lc = dev_stirling_err(n) - dev_stirling_err(x) - dev_stirling_err(rest) - dev_deviance_term(x, n * p) -
                      dev_deviance_term(rest, n * q);
    return exp(lc) * sqrt(n / (2.0 * 3.14159265358979323846 * x * rest));
}

// FnvHashMap<usize,usize> (read_id_mt_pe.rs:195 `report.iter()`) for up to 56 keys: bucket layout
// after `entry(k).or_insert()` of keys[0..n) in order (the table grows only when a NEW key arrives
// with no growth left; at most 64 buckets).  Returns, per bucket in ascending order, the key index.
__device__ __forceinline__ uint32_t dev_probe64(uint64_t occ, uint32_t nb, uint32_t hash, uint32_t gw) {
    const uint32_t mask = nb - 1;
    uint32_t pos = hash & mask;
    const uint32_t width = nb < gw ? nb : gw;
    const uint64_t wmask = (1ull << width) - 1;
    for (uint32_t stride = 0;;) {
        // occupancy bits starting at bucket `pos`, cyclic in the nb-bucket table
        uint64_t rot = occ >> pos;
        if (pos) rot |= occ << (nb - pos);
        const uint64_t free_ = ~rot & wmask;
        if (free_) return (pos + (uint32_t)__ffsll((long long)free_) - 1) & mask;
        stride += gw;
        pos = (pos + stride) & mask;
    }
}
__device__ __forceinline__ uint32_t dev_fnv_usize_low32(uint32_t key) {
    uint32_t h = 0x84222325u;                  // low half of the FNV offset basis; low 32 bits are closed under h*0x100000001b3
#pragma unroll
    for (int b = 0; b < 8; b++) h = (h ^ (b < 4 ? ((key >> (8 * b)) & 0xFFu) : 0u)) * 0x1b3u;
    return h;
}
struct DevMapOrder { uint8_t tab[2][64]; uint32_t nb; int cur; };
__device__ __forceinline__ void dev_usize_map_order(const uint32_t* __restrict__ keys, uint32_t n, uint32_t gw, DevMapOrder& m) {
    uint32_t nb = 0, items = 0, growth = 0;
    uint64_t occ = 0;
    int cur = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (growth == 0) {
            const uint32_t nn = nb == 0 ? 4 : nb * 2;
            uint64_t occ2 = 0;
            for (uint32_t s = 0; s < nb; s++)
                if ((occ >> s) & 1ull) {
                    const uint32_t it = m.tab[cur][s];
                    const uint32_t slot = dev_probe64(occ2, nn, dev_fnv_usize_low32(keys[it]), gw);
                    m.tab[cur ^ 1][slot] = (uint8_t)it;
                    occ2 |= 1ull << slot;
                }
            occ = occ2; cur ^= 1; nb = nn;
            growth = (nb < 8 ? nb - 1 : nb / 8 * 7) - items;
        }
        const uint32_t slot = dev_probe64(occ, nb, dev_fnv_usize_low32(keys[i]), gw);
        m.tab[cur][slot] = (uint8_t)i;
        occ |= 1ull << slot;
        items++; growth--;
    }
    // mark empty buckets
    for (uint32_t s = 0; s < nb; s++) if (!((occ >> s) & 1ull)) m.tab[cur][s] = 0xFF;
    m.nb = nb; m.cur = cur;
}

__global__ void __launch_bounds__(128)
readid_classify_kernel(uint64_t r0, uint64_t nreads, uint32_t N, uint32_t rep_cap, const uint32_t* __restrict__ n_set,
                       const uint32_t* __restrict__ flags, const uint32_t* __restrict__ rep_n,
                       const uint32_t* __restrict__ rep_colour, const uint32_t* __restrict__ rep_count,
                       const double* __restrict__ fp, double fp_correct, uint32_t gw, int32_t* __restrict__ kind,
                       uint32_t* __restrict__ hits, uint32_t* __restrict__ n_top, uint32_t* __restrict__ top,
                       uint32_t top_cap, uint32_t* __restrict__ list, uint32_t* __restrict__ list_cursor) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nreads) return;
    const uint64_t r = r0 + i;
    const uint32_t fl = flags[r], n = rep_n[r];
    int32_t kd;
    uint32_t best = 0, nbest = 0, bestc = 0;
    for (uint32_t t = 0; t < top_cap; t++) top[r * (uint64_t)top_cap + t] = 0;
    if (fl & 1u) kd = CID_CLS_TOO_SHORT;                     // read_id_mt_pe.rs:305-313
    else if (fl & 2u) kd = CID_CLS_REF_PANIC;
    else if (n == 0) kd = CID_CLS_NO_HITS;                   // report.is_empty(), :332-340
    else {
        const uint32_t* rc = rep_colour + r * (uint64_t)rep_cap;
        const uint32_t* rv = rep_count + r * (uint64_t)rep_cap;
        const double obs = (double)n_set[r];
        bool uncertain = false, real = false;
        uint32_t nsig = 0;
        uint64_t sigmask = 0;                                // significant entries (first 64 of the report)
        for (uint32_t e = 0; e < n; e++) {
            const uint32_t c = rc[e], v = rv[e];
            if (c >= N) continue;                            // the "no hit" key never votes (:209)
            real = true;
            const double p = fp[c];
            const double critical = __dmul_rn(obs, p), th = (double)v;
            bool drop = th < critical;                       // not_fp_signicant :168-181
            if (!drop && th > critical) {
                const double pmf = dev_binomial_pmf(obs, p, th);
                if (fabs(pmf - fp_correct) <= 1e-6 * fp_correct || !(pmf == pmf)) uncertain = true;
                drop = pmf >= fp_correct;
            }
            if (!drop) {
                nsig++;
                if (e < 64) sigmask |= 1ull << e;
                if (v > best) { best = v; nbest = 1; bestc = c; }
                else if (v == best) nbest++;
            }
        }
        if (!real) kd = CID_CLS_NO_HITS;                     // only the "no hit" key, :197-205
        else if (!uncertain && nbest > 1 && n <= 56) {
            // several accessions share the top count: "reject", names joined in the iteration order of
            // the report map (:236-249) = ascending bucket of the emulated FnvHashMap<usize,usize>
            DevMapOrder m;
            dev_usize_map_order(rc, n, gw, m);
            uint32_t nt = 0;
            for (uint32_t sl = 0; sl < m.nb; sl++) {
                const uint32_t e = m.tab[m.cur][sl];
                if (e == 0xFFu || !((sigmask >> e) & 1ull) || rv[e] != best) continue;
                if (nt < top_cap) top[r * (uint64_t)top_cap + nt] = rc[e];
                nt++;
            }
            kd = CID_CLS_REJECT_MULTI;
            kind[r] = kd; hits[r] = best; n_top[r] = nt;
            return;
        } else if (uncertain || nbest > 1) {
            kd = -1;                                         // host decides
            const uint32_t at = atomicAdd(list_cursor, 2u + 2u * n);
            list[at] = (uint32_t)i;
            list[at + 1] = n;
            for (uint32_t e = 0; e < n; e++) { list[at + 2 + 2 * e] = rc[e]; list[at + 3 + 2 * e] = rv[e]; }
        } else if (nsig == 0) kd = CID_CLS_NO_SIGNIFICANT;
        else kd = CID_CLS_ACCEPT;
    }
    const bool acc = kd == CID_CLS_ACCEPT;
    kind[r] = kd;
    hits[r] = acc ? best : 0u;
    n_top[r] = acc ? 1u : 0u;
    if (acc && top_cap) top[r * (uint64_t)top_cap] = bestc;
}
int launch_readid_classify(cid_ctx* ctx, cudaStream_t st, uint64_t r0, uint64_t nreads, uint32_t N, uint32_t rep_cap,
                           const uint32_t* d_n_set, const uint32_t* d_flags, const uint32_t* d_rep_n,
                           const uint32_t* d_rep_colour, const uint32_t* d_rep_count, const double* d_fp, double fp_correct,
                           uint32_t group_width, int32_t* d_kind, uint32_t* d_hits, uint32_t* d_n_top, uint32_t* d_top,
                           uint32_t top_cap, uint32_t* d_list, uint32_t* d_list_cursor) {
    if (nreads == 0) return CID_OK;
    ProfScope ps(ctx, st, KID_READID_CLASSIFY);
    readid_classify_kernel<<<(unsigned)((nreads + 127) / 128), 128, 0, st>>>(r0, nreads, N, rep_cap, d_n_set, d_flags, d_rep_n,
                                                                           d_rep_colour, d_rep_count, d_fp, fp_correct, group_width,
                                                                           d_kind, d_hits, d_n_top, d_top, top_cap, d_list, d_list_cursor);
    ctx->launches++;
    CID_CUDA(cudaGetLastError());
    return CID_OK;
}

// ================================================================= order export (parity hook)
__global__ void order_export_kernel(const uint16_t* __restrict__ order, const uint8_t* __restrict__ order8,
                                    const uint16_t* __restrict__ ent16, const uint32_t* __restrict__ n_set,
                                    const uint64_t* __restrict__ seq_offs, const uint64_t* __restrict__ read_offs,
                                    uint64_t r0, uint64_t nreads, uint32_t maxocc, uint32_t order_cap, int mini,
                                    uint32_t* __restrict__ order_n, uint8_t* __restrict__ order_seq,
                                    uint32_t* __restrict__ order_pos) {
    uint64_t rl = blockIdx.x;
    if (rl >= nreads) return;
    uint64_t r = r0 + rl;
    uint32_t n = min(n_set[r], order_cap);
    uint64_t s_begin = read_offs[r], s_end = read_offs[r + 1];
    uint64_t b0 = seq_offs[s_begin];
    if (threadIdx.x == 0) order_n[r] = n;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t e = order8 ? (uint32_t)ent16[rl * (uint64_t)maxocc + order8[rl * (uint64_t)maxocc + i]]
                                  : (uint32_t)order[rl * (uint64_t)maxocc + i];
        const uint32_t tp = e & 0x3FFu;
        uint32_t m = 0;
        for (uint64_t s = s_begin + 1; s < s_end; s++) if (seq_offs[s] - b0 <= tp) m = (uint32_t)(s - s_begin);
        // minimizer sets: bit 7 = the window spells the minimizer (1) or its reverse complement (0)
        order_seq[r * (uint64_t)order_cap + i] = (uint8_t)(m | (mini ? ((e >> 10) & 1u) << 7 : 0u));
        order_pos[r * (uint64_t)order_cap + i] = (uint32_t)(tp - (seq_offs[s_begin + m] - b0));
    }
}

// ================================================================= host orchestration
static uint32_t cap_to_buckets(uint32_t cap) {
    if (cap < 8) return cap < 4 ? 4 : 8;
    uint32_t adj = cap * 8 / 7, b = 1;
    while (b < adj) b <<= 1;
    return b;
}

// Sizes of the warp-per-read kernels: they hold reads of up to READID_FAST_BASES bases (all mates); longer reads (and reads
// with lower-case k-mers) take the general path of cid_readid_big.cu, so the caller's maxima are clamped here.
static void readid_dims(const cid_index* idx, uint32_t max_read_bases, uint32_t max_kmers, int* cap, uint32_t* bound,
                        uint32_t* maxocc) {
    max_read_bases = std::min<uint32_t>(max_read_bases, READID_FAST_BASES);
    *cap = (int)((max_read_bases + 31) / 32 * 32 + 32);
    uint32_t b = max_kmers ? max_kmers : (max_read_bases >= idx->k ? max_read_bases - idx->k + 1 : 1);
    b = std::min<uint32_t>(b, std::max<uint32_t>(max_read_bases, 1));      // k-mer positions <= bases
    if (b < 1) b = 1;
    *bound = b;
    *maxocc = (b + 3) & ~3u;
}
void readid_scratch_bytes(const cid_index* idx, uint32_t max_read_bases, uint32_t max_kmers, uint64_t reads,
                          size_t* entries_bytes, size_t* order_bytes, size_t* nocc_bytes) {
    int cap; uint32_t bound, maxocc;
    readid_dims(idx, max_read_bases, max_kmers, &cap, &bound, &maxocc);
    *entries_bytes = (size_t)reads * (maxocc * 4 + ((maxocc + 31) / 32) * 4);
    *order_bytes = (size_t)reads * maxocc * 2;
    *nocc_bytes = (size_t)reads * 16 + (size_t)SCHED_BINS * 4 + 16;    // nocc, nfresh, perm, schedule bins, general-path list
}

int readid_run(cid_index* idx, cudaStream_t st, const uint8_t* d_bases, const uint8_t* d_quals, const PackedReads* d_packed,
               const uint64_t* d_seq_offs, const uint64_t* d_read_offs, uint64_t r_first, uint64_t nreads,
               uint32_t max_read_bases, uint32_t max_kmers, const cid_readid_params& p, const ReadIdScratch& scr,
               uint32_t* d_n_set, uint32_t* d_flags, uint32_t* d_rep_n, uint32_t* d_rep_colour, uint32_t* d_rep_count,
               uint32_t order_cap, uint32_t* d_order_n, uint8_t* d_order_seq, uint32_t* d_order_pos) {
    cid_ctx* ctx = idx->ctx;
    if (nreads == 0) return CID_OK;
    if (idx->k == 0) return CID_E_INVALID;
    if (!scr.big || !scr.big_ctas) { set_error("read_id: no general-path scratch"); return CID_E_INVALID; }
    if (p.downsample == 0 || (p.group_width != 16 && p.group_width != 8)) { set_error("read_id: bad params"); return CID_E_INVALID; }
    if (idx->Wp > 32 * RV_MAXWPL) { set_error("read_id: more than %d accessions per shard not supported", 32 * 32 * RV_MAXWPL); return CID_E_UNSUPPORTED; }
    int cap; uint32_t bound, maxocc;
    readid_dims(idx, max_read_bases, max_kmers, &cap, &bound, &maxocc);
    const uint32_t tsize = (uint32_t)next_pow2(2ull * bound);
    const uint32_t TB = cap_to_buckets(bound + 1);
    const bool small = maxocc <= 255;
    const uint32_t maxq = d_quals && p.qual_offset ? p.qual_offset + 33 : 0;
    const uint8_t* quals = maxq ? d_quals : nullptr;
    ReadSrc rsrc{d_bases, quals, maxq, nullptr, nullptr, 0};
    if (d_packed) { rsrc.pk = d_packed->words; rsrc.pk_offs = d_packed->word_offs; rsrc.pk_lower = d_packed->lower; rsrc.bases = rsrc.quals = nullptr; rsrc.maxq = 0; }

    const uint64_t sub = std::min<uint64_t>(nreads, scr.cap_reads);
    if (sub == 0) { set_error("read_id: no scratch"); return CID_E_INVALID; }
    uint32_t* d_entries = scr.entries;
    // compact layout of the same buffer (small reads): hash bytes | positions | 9th hash bits
    const uint32_t hwords = (maxocc + 31) / 32;
    uint8_t* d_hp8 = (uint8_t*)scr.entries;
    uint16_t* d_ent16 = (uint16_t*)(d_hp8 + sub * maxocc);
    uint32_t* d_h9w = (uint32_t*)(d_ent16 + sub * maxocc);
    uint16_t* d_order = scr.order;
    uint32_t* d_nocc = scr.nocc;
    uint32_t* d_nfresh = d_nocc + scr.cap_reads;
    uint32_t* d_perm = d_nfresh + scr.cap_reads;
    uint32_t* d_bins = d_perm + scr.cap_reads;
    uint32_t* d_slow_n = d_bins + SCHED_BINS;
    uint32_t* d_slow = d_slow_n + 4;

    // shared memory budgets
    size_t a_warp = ((tile_smem_bytes(cap) + 7) & ~(size_t)7) + (size_t)tsize * 12 + (size_t)cap * 4 + (MAX_MATES + 1) * 4 + 4 + 32 + 512;
    size_t a_smem = RA_WARPS * ((a_warp + 15) & ~(size_t)15);
    const size_t b_esz = small ? 1 : 2;
    const size_t b_thread = ((((size_t)(TB + TB / 2) + maxocc) * b_esz + 3) & ~(size_t)3) +
                            4 * ((small ? (maxocc + 31) / 32 : 0) + (TB >= 32 ? TB / 32 : 1) + (TB >= 64 ? TB / 64 : 1));
    // reads per CTA of the general order kernel: 32, fewer when a read's tables are large (1000-base reads: 8.5 KB each)
    unsigned order_threads = 32;
    while (order_threads > 1 && order_threads * b_thread > 200 * 1024) order_threads >>= 1;
    size_t b_smem = order_threads * b_thread;
    if (small) b_smem = 0;      // readid_order_small_kernel<TB> sizes its own shared memory (< 48 KB)
    size_t cn_warp = ((tile_smem_bytes(cap) + 15) & ~(size_t)15) + 64 + 128 + (MAX_MATES + 1) * 4 + 12;
    const uint32_t with_steps = ctx->opt_readid_report_steps ? 1u : 0u;
    size_t cn_smem = RA_WARPS * ((cn_warp + 15) & ~(size_t)15);
    size_t cw_warp = ((tile_smem_bytes(cap) + 15) & ~(size_t)15) + (MAX_MATES + 1) * 4 + 12;
    size_t cw_smem = RA_WARPS * ((cw_warp + 15) & ~(size_t)15);
    if (a_smem > 200 * 1024 || b_smem > 200 * 1024) { set_error("read_id: read too long for the shared-memory plan"); return CID_E_UNSUPPORTED; }
    // The limit is a property of the function on the device, shared by every context (batch_id runs one worker thread and
    // context per listed device, possibly several on one GPU): always the same constant, never a per-call size, so that one
    // worker cannot lower it under another's launch.
    bool& rattr = ctx->attr_done[4];
    if (!rattr) {
        const int lim = 200 * 1024;
        CID_CUDA(cudaFuncSetAttribute(readid_kmerize_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        CID_CUDA(cudaFuncSetAttribute(readid_kmerize_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        CID_CUDA(cudaFuncSetAttribute(readid_kmerize_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        CID_CUDA(cudaFuncSetAttribute(readid_kmerize_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        CID_CUDA(cudaFuncSetAttribute(readid_order_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        rattr = true;
    }

    const ModS mods = make_mods(idx->S, idx->hv, (const HashCfg*)idx->d_hcfg);
    const uint32_t kitem = idx->m ? idx->m : idx->k;    // length of the hashed item: k-mer, or minimizer of an .mxi index
    for (uint64_t r0 = r_first; r0 < r_first + nreads; r0 += sub) {
        const uint64_t nr = std::min(sub, r_first + nreads - r0);
        unsigned gridA = (unsigned)std::min<uint64_t>((nr + RA_WARPS - 1) / RA_WARPS, (uint64_t)ctx->sm_count * 32);
        const unsigned gridK = ctx->opt_kmerize_ctas ? std::min<unsigned>(gridA, (unsigned)(ctx->sm_count * ctx->opt_kmerize_ctas)) : gridA;
        const unsigned gridV = ctx->opt_vote_ctas ? std::min<unsigned>(gridA, (unsigned)(ctx->sm_count * ctx->opt_vote_ctas)) : gridA;
        CID_CUDA(cudaMemsetAsync(d_slow_n, 0, 4, st));
        {
        ProfScope ps(ctx, st, KID_READID_KMERIZE);
#define CID_KMERIZE(C, M)                                                                                              \
    readid_kmerize_kernel<C, M><<<gridK, RA_WARPS * 32, a_smem, st>>>(rsrc, d_seq_offs, d_read_offs, r0, nr, idx->k, \
                                                                     idx->m, p.downsample, cap, maxocc, tsize, d_entries, d_hp8,   \
                                                                     d_h9w, d_ent16, d_nocc, d_nfresh, d_flags, ctx->d_err, d_slow, d_slow_n)
        if (small) { if (idx->m) CID_KMERIZE(true, true); else CID_KMERIZE(true, false); }
        else { if (idx->m) CID_KMERIZE(false, true); else CID_KMERIZE(false, false); }
#undef CID_KMERIZE
        }
        ctx->launches++;
        CID_CUDA(cudaGetLastError());
        // schedule: reads in descending order of their set size (counting sort)
        {
            ProfScope ps(ctx, st, KID_READID_SCHED);
            CID_CUDA(cudaMemsetAsync(d_bins, 0, SCHED_BINS * 4, st));
            unsigned gridH = (unsigned)std::min<uint64_t>((nr + 255) / 256, (uint64_t)ctx->sm_count * 8);
            sched_hist_kernel<<<gridH, 256, 0, st>>>(d_nfresh, nr, d_bins);
            sched_scan_kernel<<<1, SCHED_BINS, 0, st>>>(d_bins);
            sched_scatter_kernel<<<(unsigned)((nr + 255) / 256), 256, 0, st>>>(d_nfresh, nr, d_bins, d_perm);
        }
        ctx->launches += 3;
        CID_CUDA(cudaGetLastError());
        unsigned gridB = (unsigned)((nr + 31) / 32);
        {
        ProfScope ps(ctx, st, KID_READID_ORDER);
        if (small) {
            // reads are sorted by set size; every block belongs to exactly one of the three instantiations
#define CID_ORDER_SMALL(TBV)                                                                                          \
    readid_order_small_kernel<TBV><<<gridB, 32, OrderSmallLayout<TBV>::TOTAL * 128, st>>>(                            \
        d_hp8, d_h9w, d_nocc, d_nfresh, d_perm, nr, maxocc, p.group_width, p.reserve_before_find, (uint8_t*)d_order, d_n_set, r0)
            CID_ORDER_SMALL(512); CID_ORDER_SMALL(256); CID_ORDER_SMALL(128);
#undef CID_ORDER_SMALL
            ctx->launches += 2;
        }
        else
            readid_order_kernel<uint16_t><<<(unsigned)((nr + order_threads - 1) / order_threads), order_threads, b_smem, st>>>(d_entries, d_nocc, d_perm, nr, maxocc, TB, p.group_width,
                                                                    p.reserve_before_find, d_order, d_n_set, r0);
        }
        ctx->launches++;
        CID_CUDA(cudaGetLastError());
        const uint16_t* ord16 = small ? nullptr : d_order;
        const uint8_t* ord8 = small ? (const uint8_t*)d_order : nullptr;
        if (d_order_n) {
            order_export_kernel<<<(unsigned)nr, 64, 0, st>>>(ord16, ord8, d_ent16, d_n_set, d_seq_offs, d_read_offs, r0, nr, maxocc,
                                                            order_cap, idx->m != 0, d_order_n, d_order_seq, d_order_pos);
            ctx->launches++;
            CID_CUDA(cudaGetLastError());
        }
        if (d_rep_n) {
            // narrow rows in a matrix of several L2-sized windows: scan / per-window gather / count (cid_readid_part.cu);
            // the one-kernel vote below then only sees the reads that path left over
            VotePart vpd{};
            bool parted = false;
            VotePartPlan vpl;
            if (scr.vp && votepart_plan(idx, p, cap, maxocc, nr, &vpl) && scr.vp->ensure(vpl.bytes) == CID_OK) {
                CID_TRY(launch_readid_vote_part(idx, st, rsrc, d_seq_offs, d_read_offs, r0, nr, kitem, mods, cap, maxocc, ord16, ord8,
                                                d_ent16, d_n_set, p, vpl, scr.vp->as<uint8_t>(), d_flags, d_rep_n, d_rep_colour,
                                                d_rep_count, &vpd));
                parted = true;
            }
            const uint32_t* sel_list = parted ? vpd.direct : nullptr;
            const uint32_t* sel_n = parted ? vpd.direct_n : nullptr;
            const uint32_t* sel_all = parted ? vpd.cursor + VP_MAXP : nullptr;
            ProfScope ps(ctx, st, KID_READID_VOTE);
            const uint32_t* rownz = idx->rownz;
            const unsigned gridN = gridV;      // (a smaller grid for the leftover reads of the partitioned vote measured no faster)
            if (idx->Wp <= 2) {
                if (!idx->rownz_global) rownz = nullptr;   // presence == any word set, already in registers
#define CID_VOTE_NARROW(WPV, STV)                                                                                      \
    readid_vote_narrow_kernel<WPV, STV><<<gridN, RA_WARPS * 32, cn_smem, st>>>(                                        \
        rsrc, d_seq_offs, d_read_offs, r0, nr, kitem, idx->H, mods, idx->rows, rownz, idx->rownz, idx->N, cap,  \
        maxocc, ord16, ord8, d_ent16, d_n_set, p.start_sample, p.rep_cap, d_flags, d_rep_n, d_rep_colour, d_rep_count,  \
        (unsigned long long*)(ctx->d_err + 2), sel_list, sel_n, sel_all)
                if (idx->Wp == 1) { if (with_steps) CID_VOTE_NARROW(1, true); else CID_VOTE_NARROW(1, false); }
                else { if (with_steps) CID_VOTE_NARROW(2, true); else CID_VOTE_NARROW(2, false); }
#undef CID_VOTE_NARROW
            } else {
#define CID_VOTE_WIDE_H(WPLV, HTV)                                                                                     \
    readid_vote_wide_kernel<WPLV, HTV><<<gridV, RA_WARPS * 32, cw_smem, st>>>(                                         \
        rsrc, d_seq_offs, d_read_offs, r0, nr, kitem, idx->H, mods, idx->rows, rownz, idx->N, idx->Wp, cap, maxocc,  \
        ord16, ord8, d_ent16, d_n_set, p.start_sample, p.rep_cap, d_flags, d_rep_n, d_rep_colour, d_rep_count, with_steps)
#define CID_VOTE_WIDE(WPLV) do { if (idx->H == 4) CID_VOTE_WIDE_H(WPLV, 4); else if (idx->H == 2) CID_VOTE_WIDE_H(WPLV, 2); else CID_VOTE_WIDE_H(WPLV, 0); } while (0)
                if (idx->Wp <= 32) CID_VOTE_WIDE(1); else if (idx->Wp <= 64) CID_VOTE_WIDE(2); else CID_VOTE_WIDE(4);
#undef CID_VOTE_WIDE
#undef CID_VOTE_WIDE_H
            }
            ctx->launches++;
            CID_CUDA(cudaGetLastError());
        }
        // the reads the kernels above handed over (too long for their tiles, lower-case k-mers): general path, CTA per read
        CID_TRY(launch_readid_big(idx, st, rsrc, d_seq_offs, d_read_offs, r0, d_slow, d_slow_n, scr.big_bases,
                                  scr.big_kmers, p, scr.big, scr.big_ctas, d_n_set, d_flags, d_rep_n, d_rep_colour, d_rep_count,
                                  order_cap, d_order_n, d_order_seq, d_order_pos));
    }
    return CID_OK;
}

}  // namespace cid
