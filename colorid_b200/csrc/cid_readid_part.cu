// read_id vote for narrow rows (<= 64 accessions), partitioned by row range so that the row gathers hit L2.
//
// search_index (read_id_mt_pe.rs:104-165) walks a read's k-mers in hash-set order, ANDs the num_hash rows of each and stops at
// the first k-mer with an absent row.  With 8-byte rows in a 400 MB matrix every row read is a DRAM access of its own, and the
// one-kernel vote (readid_vote_narrow_kernel) runs at the DRAM random-access rate (~54 G rows/s on B200) however it is tuned.
// The same rows read through a window of the matrix that fits L2 go five times faster (tools/l2_window_probe.cu: 275 G rows/s
// for windows <= 64 MB), and nothing in the vote needs the rows in read order:
//   * where the walk stops is a property of the row-PRESENT bitmap alone (S/8 bytes, L2 resident);
//   * the candidate set comes from the first `-B` k-mers (a handful of row reads per read, done directly);
//   * after that, colour c counts the k-mers whose H rows all have bit c: an AND per k-mer and a sum per colour, both
//     order-independent.
// So the vote becomes three kernels:
//   scan    (warp per read)  first -B k-mers gathered directly -> candidate colours (at most 8, else the read goes to the
//                            one-kernel vote); every later k-mer hashed lane-parallel, row-present bits -> first miss; one
//                            (row, accumulator slot) tuple per (k-mer, hash) before the miss, appended to the bucket of the
//                            row's partition (a warp's position in a bucket is arithmetic: block i of warp w is block i * W + w);
//   gather  (per partition)  streams the bucket, reads each row from the L2-resident window, squeezes it to one bit per
//                            candidate colour and ANDs that byte into the k-mer's accumulator (red.and; skipped when every
//                            candidate bit is set, the common case for a read that comes from an indexed genome);
//   count   (warp per read)  per candidate: count over the accumulator bytes + the first -B k-mers -> the read's report,
//                            in the layout and insertion order the one-kernel vote writes.
// Reports are bit-identical to readid_vote_narrow_kernel's (tests/test_gpu_parity.py::test_read_id_partitioned_vote*).
// Measured (B200, C2, one million read pairs): scan 5.29 + six gathers 2.17 + count 0.28 ms against 8.9 ms for the one-kernel vote,
// 15.2 instead of 43.9 GB of DRAM traffic; the stage is bound by the scan's instructions (profiles/r2_readid_part_history.txt).
#include <algorithm>

#include "cid_device.cuh"
#include "cid_internal.h"
#include "cid_readid_common.cuh"

namespace cid {

constexpr int VP_BLK_LOG2 = 7, VP_BLK = 1 << VP_BLK_LOG2;   // tuples per block of a warp's share of a bucket
constexpr uint32_t VP_NULL = 0xFFFFFFFFu;   // slot of a tuple that was never written (the gather kernel substitutes it)
enum { VPM_NONE = 0, VPM_PART = 1, VPM_DIRECT = 2 };
// info word per read: [11:0] k-mers walked (incl. the one that missed) | [23:12] accumulator bytes | [24] miss | [28:25] candidates | [30:29] mode

// The H row indices of one k-mer.  Out of line on purpose: fully inlined and unrolled (hashing at two places, eight copies of
// the tuple push) the scan kernel was 70 KB of SASS and stalled on instruction fetch (no_instruction was its largest stall reason).
template <int HT>
static __device__ __noinline__ void hash_rows_ool(uint64_t w0, uint64_t w1, uint64_t w2, uint64_t w3, uint32_t k, uint32_t H,
                                                  const ModS mods, uint32_t* __restrict__ out) {
    HashIn in; in.w0 = w0; in.w1 = w1; in.w2 = w2; in.w3 = w3;
    constexpr int NH = HT ? HT : MAX_HASH;
#pragma unroll
    for (int h = 0; h < NH; h++) if (HT || (uint32_t)h < H) out[h] = (uint32_t)hash_row(in, k, h, mods);
}
// (stable XXH3 and a bloom size above 1 -- every real index: the checks of the variant and of the degenerate modulus are
// made once per k-mer here instead of once per hash)
__device__ __forceinline__ uint32_t row_stable(const HashIn& in, uint32_t k, uint64_t seed, const ModS& m) {
    const uint64_t h = xxh3_kmer_stable(in, k, seed);
    const uint64_t r = h - __umul64hi(h, m.M) * m.S;
    return (uint32_t)(r >= m.S ? r - m.S : r);
}
static __device__ __noinline__ uint4 hash_rows_4(uint64_t w0, uint64_t w1, uint64_t w2, uint64_t w3, uint32_t k, const ModS mods) {
    HashIn in; in.w0 = w0; in.w1 = w1; in.w2 = w2; in.w3 = w3;
    if (mods.var == 0u && mods.S > 1) return make_uint4(row_stable(in, k, 0, mods), row_stable(in, k, 1, mods), row_stable(in, k, 2, mods), row_stable(in, k, 3, mods));
    return make_uint4((uint32_t)hash_row(in, k, 0, mods), (uint32_t)hash_row(in, k, 1, mods), (uint32_t)hash_row(in, k, 2, mods),
                      (uint32_t)hash_row(in, k, 3, mods));
}
static __device__ __noinline__ uint2 hash_rows_2(uint64_t w0, uint64_t w1, uint64_t w2, uint64_t w3, uint32_t k, const ModS mods) {
    HashIn in; in.w0 = w0; in.w1 = w1; in.w2 = w2; in.w3 = w3;
    if (mods.var == 0u && mods.S > 1) return make_uint2(row_stable(in, k, 0, mods), row_stable(in, k, 1, mods));
    return make_uint2((uint32_t)hash_row(in, k, 0, mods), (uint32_t)hash_row(in, k, 1, mods));
}
template <int HT, int NH>
__device__ __forceinline__ void hash_rows(const HashIn& in, uint32_t k, uint32_t H, const ModS& mods, uint32_t (&rid)[NH]) {
    if (HT == 4) { const uint4 r = hash_rows_4(in.w0, in.w1, in.w2, in.w3, k, mods); rid[0] = r.x; rid[1] = r.y; rid[2] = r.z; rid[3] = r.w; }
    else if (HT == 2) { const uint2 r = hash_rows_2(in.w0, in.w1, in.w2, in.w3, k, mods); rid[0] = r.x; rid[1] = r.y; }
    else hash_rows_ool<HT>(in.w0, in.w1, in.w2, in.w3, k, H, mods, rid);
}

template <int WP, int HT>          // WP: words per row (1 or 2); HT: compile-time num_hash (0 = run-time, up to MAX_HASH)
__global__ void __launch_bounds__(RA_WARPS * 32, 8)
readid_vp_scan_kernel(const ReadSrc src, const uint64_t* __restrict__ seq_offs, const uint64_t* __restrict__ read_offs,
                      uint64_t r0, uint64_t nreads, uint32_t k, uint32_t H, ModS mods, const uint32_t* __restrict__ rows,
                      const uint32_t* __restrict__ rownz, int cap, uint32_t maxocc, const uint16_t* __restrict__ order,
                      const uint8_t* __restrict__ order8, const uint16_t* __restrict__ ent16,
                      const uint32_t* __restrict__ n_set, uint32_t B, VotePart vp, unsigned long long* __restrict__ gather_counter) {
    extern __shared__ __align__(16) uint8_t dsm[];
    __shared__ uint32_t lut[256];
    lut4_init(lut, threadIdx.x, blockDim.x);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t tile_b = (tile_smem_bytes(cap) + 15) & ~(size_t)15;
    const size_t per_warp = tile_b + 64 + (order8 ? 768 : 0);
    uint8_t* base = dsm + (size_t)warp * per_warp;
    Tile t = tile_carve(base, cap);
    uint32_t* moffs = (uint32_t*)(base + tile_b);             // MAX_MATES + 1 words
    uint32_t* s_ord = (uint32_t*)(base + tile_b + 64);        // order8 row (<= 256 bytes)
    uint32_t* s_ent = s_ord + 64;                             // ent16 row (<= 256 entries)
    const uint32_t lt = (1u << lane) - 1;
    unsigned long long my_rows = 0;
    constexpr int NH = HT ? HT : MAX_HASH;

    // Bucket space needs no atomics: block i of warp w in partition p's bucket is block i * W + w (W warps in the grid), so a
    // warp's position in a bucket is a function of how many tuples it has pushed there (lane p counts partition p's).  The
    // gather kernel gets the per-warp totals and skips what was not written.  (Reserving blocks with atomicAdd -- 32 tuples
    // per atomic, then 128 with the next block requested ahead -- left the scan waiting on the cursor words for 35 % of its
    // stall samples: ptxas parks the returned base in a temporary and copies it at once.)
    const uint32_t W = gridDim.x * RA_WARPS, wid = blockIdx.x * RA_WARPS + warp;
    uint32_t tot = 0;                            // lane p: tuples this warp pushed into partition p
    const uint32_t pb0 = (lane & 1) ? 0xFFFFFFFFu : 0u, pb1 = (lane & 2) ? 0xFFFFFFFFu : 0u, pb2 = (lane & 4) ? 0xFFFFFFFFu : 0u,
                   pb3 = (lane & 8) ? 0xFFFFFFFFu : 0u;
    // one tuple per lane (or none) straight into its partition's bucket: the lanes of a partition take consecutive slots
    auto emit = [&](bool v, uint32_t rid, uint32_t slot) {
        const uint32_t part = rid >> vp.pshift;                       // < P <= 16
        const uint32_t bv = __ballot_sync(0xffffffffu, v);
        const uint32_t b0 = __ballot_sync(0xffffffffu, part & 1u), b1 = __ballot_sync(0xffffffffu, part & 2u),
                       b2 = __ballot_sync(0xffffffffu, part & 4u), b3 = __ballot_sync(0xffffffffu, part & 8u);
        // lanes that push into MY tuple's partition / into the partition this lane counts
        const uint32_t same = bv & ~(b0 ^ ((part & 1u) ? 0xFFFFFFFFu : 0u)) & ~(b1 ^ ((part & 2u) ? 0xFFFFFFFFu : 0u)) &
                              ~(b2 ^ ((part & 4u) ? 0xFFFFFFFFu : 0u)) & ~(b3 ^ ((part & 8u) ? 0xFFFFFFFFu : 0u));
        const uint32_t mine = lane < VP_MAXP ? (bv & ~(b0 ^ pb0) & ~(b1 ^ pb1) & ~(b2 ^ pb2) & ~(b3 ^ pb3)) : 0u;
        const uint32_t pos = __shfl_sync(0xffffffffu, tot, part) + __popc(same & lt);
        if (v) {
            const uint32_t at = ((pos >> VP_BLK_LOG2) * W + wid) * (uint32_t)VP_BLK + (pos & (uint32_t)(VP_BLK - 1));
            if (pos < vp.cap_warp) vp.tuples[(size_t)part * vp.cap + at] = make_uint2(rid, slot);
            else vp.cursor[VP_MAXP] = 1u;        // bucket share used up: every read of the chunk is redone by the one-kernel vote
        }
        tot += __popc(mine);
    };

    for (uint64_t rl = (uint64_t)blockIdx.x * RA_WARPS + warp; rl < nreads; rl += (uint64_t)gridDim.x * RA_WARPS) {
        const uint64_t r = r0 + rl;
        const uint32_t n = n_set[r];
        __syncwarp();
        if (n == 0) { if (lane == 0) vp.info[rl] = (uint32_t)VPM_NONE << 29; continue; }
        const uint16_t* ordrow = order + rl * (uint64_t)maxocc;
        // The set's iteration order (order8: index of the distinct k-mer per occupied bucket; ent16: its tile position and
        // strand) is copied to shared memory with a few coalesced loads: looked up from global memory, entry = ent16[order8[i]]
        // is two dependent loads per k-mer, and the warp sat on the second one for 12 % of its stall samples.
        if (order8) {
            const uint32_t* o32 = (const uint32_t*)(order8 + rl * (uint64_t)maxocc);      // maxocc is a multiple of 4
            const uint32_t* e32 = (const uint32_t*)(ent16 + rl * (uint64_t)maxocc);
            for (uint32_t w = lane; w < (n + 3) / 4; w += 32) s_ord[w] = __ldcs(o32 + w);
            for (uint32_t w = lane; w < (n + 1) / 2; w += 32) s_ent[w] = __ldcs(e32 + w);
        }
        ReadGeom g;
        warp_load_read(t, cap, src, r, seq_offs, __ldg(read_offs + r), __ldg(read_offs + r + 1), moffs, lane, g);   // (syncs the warp)
        // k-mer `idx` of the set's iteration order -> the words XXH3 reads
        auto load_entry = [&](uint32_t idx) -> uint32_t {
            if (order8) return (uint32_t)((const uint16_t*)s_ent)[((const uint8_t*)s_ord)[idx]];
            return (uint32_t)__ldg(ordrow + idx);
        };
        auto entry_in = [&](uint32_t e) -> HashIn {
            const uint64_t f = codes_window(t.codes, (int)(e & 0x3FFu), k);
            const uint64_t key = ((e >> 10) & 1u) ? f : revcomp_key(f, k);
            return hashin_from_key(lut, key, k);
        };
        const uint32_t nb = min(B, n);
        const uint32_t slot0 = (uint32_t)rl << vp.ashift;
        uint32_t mode = VPM_PART, ncand = 0, initc = 0, p_first = n;
        unsigned long long cl = ~0ull;                          // up to 8 colour ids in insertion order, 0xFF = unused
        bool miss = false;
        // Batches of 32 k-mers, one lane each: hash (H row indices), row-present words; in the first batch the lanes of the
        // first B k-mers also read their rows -> candidate colours.  The tuples of a batch go out one round later, while the
        // next batch's loads are in flight (one more round drains the last batch).
        uint32_t rid_p[NH], slot_p = 0;
        bool valid_p = false;
#pragma unroll
        for (int h = 0; h < NH; h++) rid_p[h] = 0;
        uint32_t e_cur = (uint32_t)lane < n ? load_entry((uint32_t)lane) : 0u;
        for (uint32_t c0 = 0;; c0 += 32) {
            const bool more = c0 < n && !miss && mode == VPM_PART;
            const uint32_t idx = c0 + lane;
            const bool active = more && idx < n;
            const bool first = c0 == 0;
            uint32_t e_nxt = 0;
            uint32_t rid[NH], pz[NH], x0 = 0xFFFFFFFFu, x1 = WP == 2 ? 0xFFFFFFFFu : 0u;
#pragma unroll
            for (int h = 0; h < NH; h++) { rid[h] = 0; pz[h] = 1u; }
            if (more) {
                if (idx + 32 < n) e_nxt = load_entry(idx + 32);
                if (active) {
                    const HashIn in = entry_in(e_cur);
                    hash_rows<HT, NH>(in, k, H, mods, rid);
#pragma unroll
                    for (int h = 0; h < NH; h++) if (HT || (uint32_t)h < H) pz[h] = __ldg(rownz + (rid[h] >> 5)) >> (rid[h] & 31);
                    if (first && (uint32_t)lane < nb) {        // the first B k-mers: rows read directly
                        uint32_t ra[NH], rb[NH];
#pragma unroll
                        for (int h = 0; h < NH; h++) {
                            ra[h] = 0xFFFFFFFFu; rb[h] = 0xFFFFFFFFu;
                            if (HT || (uint32_t)h < H) {
                                if (WP == 2) { const uint2 v = __ldg((const uint2*)(rows + (size_t)rid[h] * 2)); ra[h] = v.x; rb[h] = v.y; }
                                else ra[h] = __ldg(rows + rid[h]);
                            }
                        }
#pragma unroll
                        for (int h = 0; h < NH; h++) { x0 &= ra[h]; if (WP == 2) x1 &= rb[h]; }
                    }
                }
            }
            if (__any_sync(0xffffffffu, valid_p)) {
#pragma unroll 1
                for (uint32_t h = 0; h < H; h++) {
                    uint32_t rr = rid_p[0];
#pragma unroll
                    for (int q = 1; q < NH; q++) if (h == (uint32_t)q) rr = rid_p[q];
                    emit(valid_p, rr, slot_p);
                }
                valid_p = false;
            }
            if (!more) break;
            bool absent = false;
#pragma unroll
            for (int h = 0; h < NH; h++) if (!(pz[h] & 1u)) absent = true;
            const uint32_t missmask = __ballot_sync(0xffffffffu, active && absent);
            const uint32_t p_local = missmask ? (uint32_t)(__ffs(missmask) - 1) : 32u;
            if (missmask) { miss = true; p_first = c0 + p_local; }
            if (first) {
                // candidate colours from the k-mers before min(B, first miss), in final_report insertion order
                const uint32_t pA = min(nb, p_local);
                my_rows += (unsigned long long)nb * H;
                uint32_t cand0 = 0, cand1 = 0;
                for (uint32_t jj = 0; jj < pA; jj++) {
                    const uint32_t y0 = __shfl_sync(0xffffffffu, x0, jj), y1 = __shfl_sync(0xffffffffu, x1, jj);
                    uint32_t nw0 = y0 & ~cand0, nw1 = y1 & ~cand1;
                    while (nw0) {
                        const uint32_t b = __ffs(nw0) - 1; nw0 &= nw0 - 1;
                        if (ncand < VP_MAXCAND) cl = (cl & ~(0xFFull << (8 * ncand))) | ((unsigned long long)b << (8 * ncand));
                        ncand++;
                    }
                    while (nw1) {
                        const uint32_t b = __ffs(nw1) - 1; nw1 &= nw1 - 1;
                        if (ncand < VP_MAXCAND) cl = (cl & ~(0xFFull << (8 * ncand))) | ((unsigned long long)(32u + b) << (8 * ncand));
                        ncand++;
                    }
                    cand0 |= y0; cand1 |= y1;
                }
                for (uint32_t j = 0; j < min(ncand, (uint32_t)VP_MAXCAND); j++) {       // counts over those k-mers, 4 bits each
                    const uint32_t id = (uint32_t)(cl >> (8 * j)) & 0xFFu;
                    const uint32_t w = (id & 32u) ? x1 : x0;
                    const uint32_t c = __popc(__ballot_sync(0xffffffffu, (uint32_t)lane < pA && ((w >> (id & 31u)) & 1u)));
                    initc |= c << (4 * j);
                }
                if (ncand > (uint32_t)VP_MAXCAND) {         // more candidates than an accumulator byte has bits: one-kernel vote
                    mode = VPM_DIRECT;
                    if (lane == 0) vp.direct[atomicAdd(vp.direct_n, 1u)] = (uint32_t)rl;
                }
            }
            // tuples of the k-mers after the first B and before the miss (an empty candidate set never counts anything,
            // read_id_mt_pe.rs:155-161: only the position of the miss is needed)
            if (mode == VPM_PART && ncand) {
                valid_p = active && (uint32_t)lane < p_local && idx >= nb;
                slot_p = slot0 + (idx - nb);
#pragma unroll
                for (int h = 0; h < NH; h++) rid_p[h] = rid[h];
                my_rows += (unsigned long long)__popc(__ballot_sync(0xffffffffu, valid_p)) * H;
            }
            e_cur = e_nxt;
        }
        const uint32_t nproc = miss ? p_first + 1 : n;
        const uint32_t nacc = (mode == VPM_PART && ncand && p_first > nb) ? p_first - nb : 0u;
        if (nacc) {         // accumulator bytes of the k-mers walked: all ones (bytes past them in the last word: zero)
            uint32_t* a = vp.acc32 + ((size_t)rl << (vp.ashift - 2));
            const uint32_t nw = (nacc + 3) >> 2;
            for (uint32_t w = lane; w < nw; w += 32)
                a[w] = (w == nw - 1 && (nacc & 3u)) ? (0xFFFFFFFFu >> (8 * (4 - (nacc & 3u)))) : 0xFFFFFFFFu;
        }
        if (lane == 0) {
            vp.info[rl] = min(nproc, 0xFFFu) | (nacc << 12) | ((miss ? 1u : 0u) << 24) | (min(ncand, 15u) << 25) | (mode << 29);
            vp.candl[rl] = cl;
            vp.initc[rl] = initc;
        }
    }
    // per-warp totals (what the gather kernel may read) and the longest one (how far it has to look)
    if ((uint32_t)lane < vp.P) {
        const uint32_t t2 = min(tot, vp.cap_warp);
        vp.ntup[(size_t)lane * W + wid] = t2;
        if (t2) atomicMax(vp.cursor + lane, t2);
    }
    if (lane == 0 && my_rows && gather_counter) atomicAdd(gather_counter, my_rows);
}

// One partition's bucket: row from the L2-resident window -> one bit per candidate colour -> AND into the accumulator byte.
// The bucket is W interleaved per-warp block sequences (see the scan kernel): block b belongs to warp b % W and is its
// (b / W)-th; ntup[w] says how many tuples that warp wrote, maxtup the largest of them.
template <int WP>
__global__ void __launch_bounds__(256)
readid_vp_gather_kernel(const uint32_t* __restrict__ rows, const uint2* __restrict__ tuples, const uint32_t* __restrict__ ntup,
                        const uint32_t* __restrict__ maxtup, uint32_t W, uint64_t Wmagic, const unsigned long long* __restrict__ candl,
                        uint32_t* __restrict__ acc32, uint32_t ashift) {
    constexpr int ILP = 4;      // (32 registers: full occupancy; 8 in flight per thread at 56 registers measured 2.45 against 2.17 ms)
    const uint64_t ntot = (uint64_t)((*maxtup + VP_BLK - 1) >> VP_BLK_LOG2) * W * VP_BLK;      // < 2^32 (the plan caps the bucket)
    for (uint64_t base = (uint64_t)blockIdx.x * 256 * ILP; base < ntot; base += (uint64_t)gridDim.x * 256 * ILP) {
        uint2 tp[ILP], v[ILP];
        unsigned long long cl[ILP];
#pragma unroll
        for (int u = 0; u < ILP; u++) {
            const uint64_t t = base + (uint64_t)u * 256 + threadIdx.x;
            const uint32_t blk = (uint32_t)(t >> VP_BLK_LOG2);
            const uint32_t i = (uint32_t)(((uint64_t)blk * Wmagic) >> 40), w = blk - i * W;      // blk / W, blk % W (exact: see the launcher)
            const bool ok = t < ntot && (i << VP_BLK_LOG2) + ((uint32_t)t & (uint32_t)(VP_BLK - 1)) < __ldg(ntup + w);
            tp[u] = ok ? __ldcs(tuples + t) : make_uint2(0u, VP_NULL);          // streamed once: do not displace the window
        }
#pragma unroll
        for (int u = 0; u < ILP; u++) {
            if (WP == 2) v[u] = __ldg((const uint2*)rows + tp[u].x);
            else { v[u].x = __ldg(rows + tp[u].x); v[u].y = 0u; }
            cl[u] = tp[u].y == VP_NULL ? ~0ull : __ldg(candl + (tp[u].y >> ashift));
        }
#pragma unroll
        for (int u = 0; u < ILP; u++) {
            uint32_t b = 0;
#pragma unroll
            for (int j = 0; j < VP_MAXCAND; j++) {
                const uint32_t id = (uint32_t)(cl[u] >> (8 * j)) & 0xFFu;
                const uint32_t w = (WP == 2 && (id & 32u)) ? v[u].y : v[u].x;
                b |= (((w >> (id & 31u)) & 1u) | (id >> 7)) << j;
            }
            if (b != 0xFFu) {
                const uint32_t mask = ~((b ^ 0xFFu) << (8 * (tp[u].y & 3u)));
                asm volatile("red.global.and.b32 [%0], %1;" ::"l"(acc32 + (tp[u].y >> 2)), "r"(mask) : "memory");
            }
        }
    }
}

// Per read: candidate counts = first-B counts + set bits of the accumulator bytes; report in final_report insertion order,
// the "no hit" key N last (inserted at the break), exactly what readid_vote_narrow_kernel writes.
__global__ void __launch_bounds__(256)
readid_vp_count_kernel(uint64_t r0, uint64_t nreads, uint32_t N, uint32_t rep_cap, VotePart vp, uint32_t* __restrict__ flags,
                       uint32_t* __restrict__ rep_n, uint32_t* __restrict__ rep_colour, uint32_t* __restrict__ rep_count) {
    if (vp.cursor[VP_MAXP]) return;             // a bucket overflowed: the one-kernel vote redoes the whole chunk
    const uint64_t rl = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (rl >= nreads) return;
    const uint64_t r = r0 + rl;
    const uint32_t info = vp.info[rl], mode = (info >> 29) & 3u;
    if (mode == VPM_DIRECT) return;
    if (mode == VPM_NONE) { if (lane == 0) rep_n[r] = 0; return; }
    const uint32_t nproc = info & 0xFFFu, nacc = (info >> 12) & 0xFFFu, miss = (info >> 24) & 1u, ncand = (info >> 25) & 15u;
    const unsigned long long cl = vp.candl[rl];
    const uint32_t ic = vp.initc[rl];
    uint32_t cnt[VP_MAXCAND];
#pragma unroll
    for (int j = 0; j < VP_MAXCAND; j++) cnt[j] = 0;
    const uint32_t* a = vp.acc32 + ((size_t)rl << (vp.ashift - 2));
    for (uint32_t w = lane; w < (nacc + 3) >> 2; w += 32) {
        const uint32_t v = a[w];
#pragma unroll
        for (int j = 0; j < VP_MAXCAND; j++) cnt[j] += __popc(v & (0x01010101u << j));
    }
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < VP_MAXCAND; j++) {
        const uint32_t s = __reduce_add_sync(0xffffffffu, cnt[j]);
        if (lane == (uint32_t)j) mine = s + ((ic >> (4 * j)) & 15u);
    }
    uint32_t* rc = rep_colour + r * (uint64_t)rep_cap;
    uint32_t* rv = rep_count + r * (uint64_t)rep_cap;
    if (lane < ncand && lane < rep_cap) { rc[lane] = (uint32_t)(cl >> (8 * lane)) & 0xFFu; rv[lane] = mine; }
    if (lane == 0) {
        const uint32_t total = ncand + miss;
        if (miss && ncand < rep_cap) { rc[ncand] = N; rv[ncand] = 1; }
        rep_n[r] = min(total, rep_cap);
        flags[r] |= (total > rep_cap ? 4u : 0u) | (min(nproc, 0xFFFFu) << 8);
    }
}

// ================================================================= host side
// Partition and scratch plan of one chunk; false = the partitioned vote does not apply (the caller keeps the one-kernel vote).
bool votepart_plan(const cid_index* idx, const cid_readid_params& p, int cap_bases, uint32_t maxocc, uint64_t reads, VotePartPlan* out) {
    const cid_ctx* ctx = idx->ctx;
    VotePartPlan pl{};
    if (ctx->opt_readid_vote_part == 0 || idx->Wp > 2 || !idx->rownz || ctx->opt_readid_report_steps) return false;
    if (p.start_sample == 0 || p.start_sample > VP_MAXB || maxocc > 2047 || reads == 0 || idx->H > MAX_HASH) return false;
    const bool forced = ctx->opt_readid_vote_part >= 2;
    // windows of <= 64 MB: 275 G rows/s on B200 against 54 G/s from DRAM (profiles/r2_l2_window_probe.txt)
    uint32_t pshift = ctx->opt_readid_part_shift ? (uint32_t)ctx->opt_readid_part_shift : (idx->Wp == 2 ? 23u : 24u);
    while (((idx->S - 1) >> pshift) + 1 > (uint64_t)VP_MAXP) pshift++;
    const uint32_t P = (uint32_t)(((idx->S - 1) >> pshift) + 1);
    if (!forced && (P < 3 || reads < 16384)) return false;     // a matrix of two windows mostly sits in L2 anyway
    uint32_t ashift = 2;
    while ((1u << ashift) < maxocc) ashift++;
    if ((reads << ashift) >= 0xFFFFFFFFull) return false;
    // the scan kernel runs as resident CTAs of 4 warps striding over the reads; warp w owns every W-th block of each bucket
    const size_t smem = RA_WARPS * (((tile_smem_bytes(cap_bases) + 15) & ~(size_t)15) + 64 + 768);
    pl.ctas_per_sm = (uint32_t)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / (smem + 1024)));
    if (ctx->opt_readid_part_ctas) pl.ctas_per_sm = std::min<uint32_t>(pl.ctas_per_sm, (uint32_t)ctx->opt_readid_part_ctas);
    pl.grid = (uint32_t)std::min<uint64_t>((reads + RA_WARPS - 1) / RA_WARPS, (uint64_t)ctx->sm_count * pl.ctas_per_sm);
    const uint64_t W = (uint64_t)pl.grid * RA_WARPS;
    // a warp's share of a bucket: its reads' worst case split evenly over the partitions, 30 % slack, two blocks on top
    const uint64_t per_warp = ((reads + W - 1) / W) * (uint64_t)maxocc * idx->H / P;
    uint64_t cap_warp = (((per_warp + per_warp * 3 / 10) >> VP_BLK_LOG2) + 3) << VP_BLK_LOG2;
    if (ctx->opt_readid_part_cap) cap_warp = (uint64_t)ctx->opt_readid_part_cap;
    const uint64_t cap = W * (((cap_warp + VP_BLK - 1) >> VP_BLK_LOG2) << VP_BLK_LOG2);
    if (cap >= 0xFFFFFF00ull || W >= 8192) return false;
    pl.cap_warp = (uint32_t)cap_warp;
    pl.P = P; pl.pshift = pshift; pl.ashift = ashift; pl.cap = (uint32_t)cap;
    size_t o = 256;                                       // cursors, overflow flag, direct-list length
    pl.o_info = o; o += (reads * 4 + 255) & ~(size_t)255;
    pl.o_initc = o; o += (reads * 4 + 255) & ~(size_t)255;
    pl.o_direct = o; o += (reads * 4 + 255) & ~(size_t)255;
    pl.o_candl = o; o += (reads * 8 + 255) & ~(size_t)255;
    pl.o_acc = o; o += ((size_t)reads << ashift) + 256;
    pl.o_ntup = o; o += (size_t)VP_MAXP * W * 4 + 256;
    pl.o_tuples = o; o += (size_t)P * cap * 8;
    pl.bytes = o;
    if (!forced && pl.bytes > (24ull << 30)) return false;
    *out = pl;
    return true;
}

int launch_readid_vote_part(cid_index* idx, cudaStream_t st, const ReadSrc& rsrc, const uint64_t* d_seq_offs,
                            const uint64_t* d_read_offs, uint64_t r0, uint64_t nr, uint32_t kitem, const ModS& mods, int cap,
                            uint32_t maxocc, const uint16_t* ord16, const uint8_t* ord8, const uint16_t* d_ent16,
                            const uint32_t* d_n_set, const cid_readid_params& p, const VotePartPlan& pl, uint8_t* d_scratch,
                            uint32_t* d_flags, uint32_t* d_rep_n, uint32_t* d_rep_colour, uint32_t* d_rep_count, VotePart* vp_out) {
    cid_ctx* ctx = idx->ctx;
    VotePart vp;
    vp.cursor = (uint32_t*)d_scratch;
    vp.direct_n = vp.cursor + VP_MAXP + 1;
    vp.info = (uint32_t*)(d_scratch + pl.o_info);
    vp.initc = (uint32_t*)(d_scratch + pl.o_initc);
    vp.direct = (uint32_t*)(d_scratch + pl.o_direct);
    vp.candl = (unsigned long long*)(d_scratch + pl.o_candl);
    vp.acc32 = (uint32_t*)(d_scratch + pl.o_acc);
    vp.tuples = (uint2*)(d_scratch + pl.o_tuples);
    vp.ntup = (uint32_t*)(d_scratch + pl.o_ntup);
    vp.cap = pl.cap; vp.cap_warp = pl.cap_warp; vp.ashift = pl.ashift; vp.pshift = pl.pshift; vp.P = pl.P;
    *vp_out = vp;
    CID_CUDA(cudaMemsetAsync(d_scratch, 0, 256, st));
    const size_t tile_b = (tile_smem_bytes(cap) + 15) & ~(size_t)15;
    const size_t smem = RA_WARPS * (tile_b + 64 + (ord8 ? 768 : 0));
    if (smem > 200 * 1024) { set_error("read_id: partitioned vote does not fit shared memory"); return CID_E_UNSUPPORTED; }
    bool& attr = ctx->attr_done[8];
    if (!attr) {
        const int lim = 200 * 1024;
        CID_CUDA(cudaFuncSetAttribute(readid_vp_scan_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        CID_CUDA(cudaFuncSetAttribute(readid_vp_scan_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        CID_CUDA(cudaFuncSetAttribute(readid_vp_scan_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        CID_CUDA(cudaFuncSetAttribute(readid_vp_scan_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        CID_CUDA(cudaFuncSetAttribute(readid_vp_scan_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        CID_CUDA(cudaFuncSetAttribute(readid_vp_scan_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        attr = true;
    }
    const unsigned gridA = pl.grid;              // (the buckets are laid out for exactly this many warps)
    {
        ProfScope ps(ctx, st, KID_READID_VP_SCAN);
#define CID_VP_SCAN(WPV, HTV)                                                                                          \
    readid_vp_scan_kernel<WPV, HTV><<<gridA, RA_WARPS * 32, smem, st>>>(rsrc, d_seq_offs, d_read_offs, r0, nr, kitem, idx->H, mods, \
        idx->rows, idx->rownz, cap, maxocc, ord16, ord8, d_ent16, d_n_set, p.start_sample, vp, (unsigned long long*)(ctx->d_err + 2))
        if (idx->Wp == 1) { if (idx->H == 4) CID_VP_SCAN(1, 4); else if (idx->H == 2) CID_VP_SCAN(1, 2); else CID_VP_SCAN(1, 0); }
        else { if (idx->H == 4) CID_VP_SCAN(2, 4); else if (idx->H == 2) CID_VP_SCAN(2, 2); else CID_VP_SCAN(2, 0); }
#undef CID_VP_SCAN
    }
    ctx->launches++;
    CID_CUDA(cudaGetLastError());
    {
        ProfScope ps(ctx, st, KID_READID_VP_GATHER);
        const unsigned gridG = (unsigned)ctx->sm_count * 32;
        const uint32_t W = pl.grid * RA_WARPS;
        // blk / W as (blk * Wmagic) >> 40 with Wmagic = floor(2^40 / W) + 1: exact for blk < 2^26 and W < 2^13
        const uint64_t Wmagic = (1ull << 40) / W + 1;
        for (uint32_t q = 0; q < pl.P; q++) {
            if (idx->Wp == 1)
                readid_vp_gather_kernel<1><<<gridG, 256, 0, st>>>(idx->rows, vp.tuples + (size_t)q * vp.cap, vp.ntup + (size_t)q * W, vp.cursor + q, W, Wmagic, vp.candl, vp.acc32, vp.ashift);
            else
                readid_vp_gather_kernel<2><<<gridG, 256, 0, st>>>(idx->rows, vp.tuples + (size_t)q * vp.cap, vp.ntup + (size_t)q * W, vp.cursor + q, W, Wmagic, vp.candl, vp.acc32, vp.ashift);
        }
    }
    ctx->launches += pl.P;
    CID_CUDA(cudaGetLastError());
    {
        ProfScope ps(ctx, st, KID_READID_VP_COUNT);
        readid_vp_count_kernel<<<(unsigned)((nr * 32 + 255) / 256), 256, 0, st>>>(r0, nr, idx->N, p.rep_cap, vp, d_flags, d_rep_n, d_rep_colour, d_rep_count);
    }
    ctx->launches++;
    CID_CUDA(cudaGetLastError());
    return CID_OK;
}

}  // namespace cid
