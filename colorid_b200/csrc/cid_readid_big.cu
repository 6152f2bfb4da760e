// read_id, general path (sm_100a): one CTA per read, per-read tables in global memory.
//
// The warp-per-read kernels of cid_readid.cu keep a read's state in a slice of shared memory and 2-bit packed keys,
// which covers reads of up to 1,000 bases whose k-mers are upper case.  Everything else is classified here, with the
// same semantics and no length limit short of 8 Mbp per read:
//   * reads longer than that -- read_id_mt_pe.rs:450-569 stream_fasta classifies whole contigs, long-read FASTQ;
//   * reads with a lower-case base inside a k-mer: kmer.rs:221-243 kmerize_vector_skip_n_set never upper-cases, so such
//     a k-mer is hashed (FNV-1a for the set order, XXH3 for the rows) with its raw bytes.  The dedup key is therefore
//     (2-bit codes, case mask of the canonical orientation).
// The fast kmerize kernel appends those reads to a device list; this kernel walks the list.  Phases per read, all
// threads of the CTA in step:
//   1. masked copy of the read (seq.rs:36-56 qual_mask) into scratch;
//   2. canonical k-mers tile by tile (same Tile machinery as the other kernels) into a global open-addressing table;
//      first occurrence per distinct k-mer by atomicMin;
//   3. the sequence of fresh HashSet::insert calls in read order (block scan);
//   4. hashbrown's bucket layout epoch by epoch: within one table size, element i ends up in the first
//      slot of ITS probe sequence that no element inserted before it holds -- a fixed point that does not depend on
//      the order in which the slots are claimed, so all elements are inserted concurrently with atomicMin on their
//      insertion rank, a displaced (later) element resuming its own probe sequence; a resize re-inserts the old table
//      in ascending bucket order (block compaction) followed by the next fresh keys;
//   5. search_index / search_index_classic (read_id_mt_pe.rs:66-165) over the set order, 1024 k-mers per step:
//      row indices + row-present bits per thread, first absent row by block reduction, then one warp per k-mer ANDs
//      the rows and counts per colour in shared memory; report in final_report insertion order (bitonic sort by
//      (first step, colour)).
#include <algorithm>

#include "cid_device.cuh"
#include "cid_internal.h"

namespace cid {

constexpr int BG_THREADS = 1024;
constexpr int BG_TILE = 1024;                  // k-mer start positions per tile
constexpr int BG_TCAP = BG_TILE + 64;          // bytes staged per tile (halo <= 31)
constexpr int BG_MAX_MATES = 8;
constexpr uint32_t BG_MAX_COLOURS = 4096;      // per-colour counters + first steps + sort keys live in shared memory
constexpr uint32_t BG_UNSET = 0xFFFFFFFFu;
constexpr uint32_t BG_STEP_SHIFT = 20;         // readid_report_steps: colour | step << 20 (cid_merge_shard_reports)

struct __align__(16) BigSlot { unsigned long long key; uint32_t cs; uint32_t first; };

// byte offsets of one CTA's scratch arrays
struct BigLayout {
    size_t seq, tab, pinfo, flist, fhash, plb, total;
    uint32_t tslots;       // dedup table slots (power of two)
    uint32_t hbcap;        // buckets reserved for the hashbrown table
    uint32_t lcap, kcap;   // bases / k-mer positions the layout was sized for
};
static uint32_t big_buckets_for(uint32_t items) {      // hashbrown capacity_to_buckets
    if (items < 8) return items < 4 ? 4 : 8;
    uint64_t adj = (uint64_t)items * 8 / 7, b = 1;
    while (b < adj) b <<= 1;
    return (uint32_t)b;
}
static BigLayout big_layout(uint32_t lcap, uint32_t kcap) {
    BigLayout L{};
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    L.lcap = lcap; L.kcap = kcap;
    L.tslots = (uint32_t)next_pow2(std::max<uint64_t>(64, 2ull * kcap));
    size_t off = 0;
    L.seq = off; off += up((size_t)lcap + 128);
    // the dedup table is dead once the insert calls are listed: the hashbrown table (u32 per bucket, <= 2.3 kcap buckets
    // after the tail resize) and one priority list live in the same bytes
    L.hbcap = big_buckets_for(kcap + 2) * 2;
    const size_t hb_bytes = (size_t)L.hbcap * 4 + (size_t)kcap * 4 + 64;
    L.tab = off; off += up(std::max((size_t)L.tslots * sizeof(BigSlot), hb_bytes));
    L.pinfo = off; off += up((size_t)lcap * 4 + 64);
    L.flist = off; off += up((size_t)kcap * 4 + 64);
    L.fhash = off; off += up((size_t)kcap * 4 + 64);
    L.plb = off; off += up((size_t)kcap * 4 + 64);
    L.total = off;
    return L;
}

struct BigArgs {
    ReadSrc src;
    const uint64_t* seq_offs; const uint64_t* read_offs; uint64_t r0;
    const uint32_t* list; const uint32_t* list_n;
    uint32_t k, mini_m, d, H; ModS mods;
    const uint32_t* rows; const uint32_t* rownz; uint32_t N, Wp;
    uint32_t start_sample, rep_cap, with_steps, gw, rbf;
    uint8_t* scratch; BigLayout lay;
    uint32_t* n_set; uint32_t* flags; uint32_t* rep_n; uint32_t* rep_colour; uint32_t* rep_count;
    uint32_t order_cap; uint32_t* order_n; uint8_t* order_seq; uint32_t* order_pos;
    uint32_t* err;
};

// ---- block-level helpers (BG_THREADS threads) ----------------------------------------------------------------------
// exclusive prefix of `flag` over the block + block total; `wsum` is a 33-word shared array
__device__ __forceinline__ uint32_t block_scan_flag(bool flag, uint32_t* wsum, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
        uint32_t v = wsum[lane], x = v;
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        wsum[lane] = x - v;
        if (lane == 31) wsum[32] = x;
    }
    __syncthreads();
    const uint32_t pre = wsum[warp] + __popc(bal & ((1u << lane) - 1));
    total = wsum[32];
    __syncthreads();
    return pre;
}

// hashbrown probe sequence of a key with hash h in a table of nb buckets: `width` consecutive
// buckets from pos (the SIMD group; small tables: the whole table, cyclic), then triangular strides of gw
struct HbSeq {
    uint32_t pos, stride, i, mask, width, gw;
    __device__ __forceinline__ void start(uint32_t h, uint32_t nb, uint32_t g) {
        mask = nb - 1; gw = g; width = nb < g ? nb : g; pos = h & mask; stride = 0; i = 0;
    }
    __device__ __forceinline__ uint32_t slot() const { return (pos + i) & mask; }
    __device__ __forceinline__ void next() {
        if (++i == width) { i = 0; stride += gw; pos = (pos + stride) & mask; }
    }
};

// canonical item of the set at window [ipos, ipos+len) of the masked read: the bytes as the reference hashes them
// (raw case for k-mers, upper case for minimizers); fwd = the window spells the item, else its reverse complement does
__device__ __forceinline__ uint32_t item_byte(const uint8_t* seq, uint32_t ipos, uint32_t len, bool fwd, bool upper, uint32_t j) {
    uint32_t c = fwd ? (uint32_t)seq[ipos + j] : comp_base(seq[ipos + len - 1 - j]);
    if (upper) c &= 0xDFu;
    return c;
}

__global__ void __launch_bounds__(BG_THREADS, 1)
readid_big_kernel(const BigArgs a) {
    extern __shared__ __align__(16) uint8_t dsm[];
    __shared__ uint32_t s_wsum[33];
    __shared__ uint32_t s_moffs[BG_MAX_MATES + 1];
    __shared__ uint32_t s_u[8];
    // dynamic shared memory: tile | cnt[N] | fstep[N] | cand[Wp] | rid[1024*H] (the sort keys reuse rid)
    Tile t = tile_carve(dsm, BG_TCAP);
    uint32_t* cnt = (uint32_t*)(dsm + ((tile_smem_bytes(BG_TCAP) + 15) & ~(size_t)15));
    uint32_t* fstep = cnt + BG_MAX_COLOURS;
    uint32_t* cand = fstep + BG_MAX_COLOURS;
    uint32_t* rid_s = cand + 128;
    unsigned long long* skeys = (unsigned long long*)rid_s;     // [BG_MAX_COLOURS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t k = a.k, m = a.mini_m, len_item = m ? m : k;
    const bool mini = m != 0;

    uint8_t* my = a.scratch + (size_t)blockIdx.x * a.lay.total;
    uint8_t* seq = my + a.lay.seq;
    BigSlot* tab = (BigSlot*)(my + a.lay.tab);
    uint32_t* pinfo = (uint32_t*)(my + a.lay.pinfo);
    uint32_t* flist = (uint32_t*)(my + a.lay.flist);
    uint32_t* fhash = (uint32_t*)(my + a.lay.fhash);
    uint32_t* plB = (uint32_t*)(my + a.lay.plb);
    // after phase 3 the dedup table's bytes hold the hashbrown table and the other priority list
    uint32_t* hb = (uint32_t*)(my + a.lay.tab);
    const uint32_t nlist = *a.list_n;

    for (uint32_t li = blockIdx.x; li < nlist; li += gridDim.x) {
        const uint64_t rl = a.list[li], r = a.r0 + rl;
        const uint64_t s_begin = a.read_offs[r], s_end = a.read_offs[r + 1];
        const uint32_t nm = (uint32_t)min(s_end - s_begin, (uint64_t)(BG_MAX_MATES + 1));
        const uint64_t b0 = a.seq_offs[s_begin], b1 = a.seq_offs[s_end];
        const uint64_t L64 = b1 - b0;
        __syncthreads();
        if (tid <= (int)nm && tid <= BG_MAX_MATES) s_moffs[tid] = (uint32_t)(a.seq_offs[s_begin + tid] - b0);
        __syncthreads();
        uint32_t fl = 16u;                              // bit 4: classified by the general path
        bool ok = s_end > s_begin;
        if (ok && s_moffs[1] - s_moffs[0] < k) ok = false;                   // read_id_mt_pe.rs:305 too_short
        if (!ok) fl |= 1u;
        if (ok && (s_end - s_begin > BG_MAX_MATES || L64 > (uint64_t)a.lay.lcap)) {
            if (tid == 0) atomicOr(a.err, ERRF_READ_TOO_LONG);
            fl |= 8u; ok = false;
        }
        if (ok && !mini)                               // kmer.rs:229 `0..l.len()-k+1` underflows for a later mate shorter than k-1
            for (uint32_t j = 1; j < nm; j++) if (s_moffs[j + 1] - s_moffs[j] + 1 < k) { fl |= 2u; ok = false; }
        const uint32_t L = ok ? (uint32_t)L64 : 0u;
        uint32_t F = 0, nproc = 0, total_rep = 0;
        bool miss = false;
        if (ok) {
            // ---- 1. masked copy -------------------------------------------------------------------------------------
            const uint32_t* pw = a.src.pk ? a.src.pk + a.src.pk_offs[r] : nullptr;
            const uint32_t ncw = (L + 15) >> 4, nbw = (L + 31) >> 5;
            for (uint32_t i = tid; i < L + 64; i += BG_THREADS) {
                uint32_t c = 0;
                if (i < L) {
                    if (pw) {        // packed planes: the bytes are spelled back ('N' for any non-base; the case from the lower plane)
                        c = code_ascii((pw[i >> 4] >> (30 - 2 * (i & 15))) & 3u);
                        if (a.src.pk_lower && ((pw[ncw + nbw + (i >> 5)] >> (i & 31)) & 1u)) c |= 0x20u;
                        if ((pw[ncw + (i >> 5)] >> (i & 31)) & 1u) c = 'N';
                    } else {
                        c = a.src.bases[b0 + i];
                        if (a.src.quals && (uint32_t)a.src.quals[b0 + i] < a.src.maxq) c = 'N';
                    }
                }
                seq[i] = (uint8_t)c;
            }
            {   // dedup table: key all ones, case unset, first occurrence +inf
                uint4* t4 = (uint4*)tab;
                for (uint32_t i = tid; i < a.lay.tslots; i += BG_THREADS) t4[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
            }
            __syncthreads();
            // ---- 2. canonical k-mers -> dedup table ---------------------------------------------------------------
            const uint32_t tmask = a.lay.tslots - 1;
            for (uint32_t t0 = 0; t0 < L; t0 += BG_TILE) {
                const int tile_len = (int)min((uint32_t)(BG_TILE + k - 1), L - t0);
                __syncthreads();
                t.len = tile_len;
                for (int i = tid; i < BG_TCAP / 32 + 2; i += BG_THREADS) t.start[i] = 0;
                for (int i = tid; i < BG_TCAP; i += BG_THREADS) t.ascii[i] = i < tile_len ? seq[t0 + i] : (uint8_t)0;
                __syncthreads();
                if (tid >= 1 && tid < (int)nm) {        // mate starts inside the tile: windows may not cross them
                    const uint32_t o = s_moffs[tid];
                    if (o > t0 && o < t0 + (uint32_t)tile_len) atomicOr(&t.start[(o - t0) >> 5], 1u << ((o - t0) & 31));
                }
                tile_pack(t, BG_TCAP, tid, BG_THREADS);
                __syncthreads();
                const int p = tid;
                const uint32_t gp = t0 + (uint32_t)p;
                if (gp < L) {
                    uint32_t info = 0;
                    bool take = true;
                    if (a.d != 1) {                     // kmer.rs:229 step_by(d): position inside the mate
                        uint32_t mt = 0;
                        for (uint32_t j = 1; j < nm; j++) if (s_moffs[j] <= gp) mt = j;
                        take = (gp - s_moffs[mt]) % a.d == 0;
                    }
                    uint64_t key; bool fwd, low;
                    if (take && tile_kmer(t, p, k, key, fwd, low)) {
                        uint32_t cs = 0, delta = 0;
                        if (mini) {                     // kmer.rs:363-394: the set holds upper-cased minimizers
                            uint32_t ipos; bool mfwd;
                            key = tile_minimizer(t, p, k, m, key, fwd, low, ipos, mfwd);
                            fwd = mfwd; delta = ipos - (uint32_t)p;
                        } else if (low) {               // case mask in the orientation of the canonical string
                            const uint32_t lw = mask_window(t.lower, p, k);
                            cs = fwd ? lw : (__brev(lw) >> (32 - k));
                        }
                        uint32_t s = (uint32_t)mix64(key ^ ((uint64_t)cs << 31) ^ ((uint64_t)cs << 47)) & tmask;
                        for (;;) {
                            const unsigned long long prev = atomicCAS(&tab[s].key, CID_EMPTY_KEY, (unsigned long long)key);
                            if (prev == CID_EMPTY_KEY || prev == key) {
                                // the slot's codes are ours: whoever arrives first also settles its case mask
                                const uint32_t c = atomicCAS(&tab[s].cs, BG_UNSET, cs);
                                if (c == BG_UNSET || c == cs) break;
                            }
                            s = (s + 1) & tmask;
                        }
                        atomicMin(&tab[s].first, gp);
                        info = 0x80000000u | (fwd ? 0x40000000u : 0u) | (delta << 24) | s;
                    }
                    pinfo[gp] = info;
                }
            }
            __syncthreads();
            // ---- 3. fresh insert calls in read order ---------------------------------------------------------------
            if (tid == 0) s_u[0] = 0;                    // (position + 1) << 1 | fresh of the last insert call
            __syncthreads();
            for (uint32_t c0 = 0; c0 < L; c0 += BG_THREADS) {
                const uint32_t gp = c0 + tid;
                const uint32_t info = gp < L ? pinfo[gp] : 0u;
                const bool valid = info >> 31;
                const bool fresh = valid && tab[info & 0xFFFFFFu].first == gp;
                uint32_t tot;
                const uint32_t fi = F + block_scan_flag(fresh, s_wsum, tot);
                if (valid) atomicMax(&s_u[0], ((gp + 1) << 1) | (fresh ? 1u : 0u));
                if (fresh) {
                    const bool fwd = (info >> 30) & 1u;
                    const uint32_t ipos = gp + ((info >> 24) & 31u);
                    uint32_t h = 0x84222325u;           // FNV-1a, low 32 bits (closed under the multiply), `impl Hash for str`
                    for (uint32_t j = 0; j < len_item; j++) h = (h ^ item_byte(seq, ipos, len_item, fwd, mini, j)) * 0x1b3u;
                    h = (h ^ 0xFFu) * 0x1b3u;
                    flist[fi] = ipos | (fwd ? 0x80000000u : 0u);
                    fhash[fi] = h;
                }
                F += tot;
            }
            __syncthreads();
            const bool tail_dup = s_u[0] != 0 && !(s_u[0] & 1u);
            __syncthreads();
            // ---- 4. hashbrown layout, one table size at a time ------------------------------------------------------
            uint32_t* plcur = hb + a.lay.hbcap;               // behind the largest table
            uint32_t* plnext = plB;
            uint32_t nb = 0, f = 0;
            bool extra = a.rbf && tail_dup && F > 0;     // a duplicate insert call after the last fresh key resizes a full table
            while (f < F || extra) {
                const uint32_t nbn = nb ? nb * 2 : 4;
                const uint32_t capn = nbn < 8 ? nbn - 1 : nbn / 8 * 7;
                if (f == F) {
                    // only the tail duplicate is left: it resizes iff the table is exactly full (growth_left == 0)
                    extra = false;
                    const uint32_t capc = nb < 8 ? (nb ? nb - 1 : 0) : nb / 8 * 7;
                    if (F != capc) break;
                }
                const uint32_t hi = min(F, capn);
                for (uint32_t i = tid; i < nbn; i += BG_THREADS) hb[i] = BG_UNSET;
                __syncthreads();
                for (uint32_t rk = tid; rk < hi; rk += BG_THREADS) {
                    // rank rk of the insertion order: old buckets in ascending order, then the fresh keys f..hi
                    uint32_t cur = rk;
                    HbSeq ps;
                    ps.start(fhash[rk < f ? plcur[rk] : rk], nbn, a.gw);
                    for (;;) {
                        const uint32_t sl = ps.slot();
                        const uint32_t old = atomicMin(&hb[sl], cur);
                        if (old == BG_UNSET) break;
                        if (old > cur) {                 // a later element sat here: it moves on along its own sequence
                            cur = old;
                            ps.start(fhash[cur < f ? plcur[cur] : cur], nbn, a.gw);
                            while (ps.slot() != sl) ps.next();
                        }
                        ps.next();
                    }
                }
                __syncthreads();
                // occupied buckets in ascending order -> next priority list (element ids)
                uint32_t outn = 0;
                for (uint32_t c0 = 0; c0 < nbn; c0 += BG_THREADS) {
                    const uint32_t sidx = c0 + tid;
                    const uint32_t v = sidx < nbn ? hb[sidx] : BG_UNSET;
                    uint32_t tot;
                    const uint32_t o = outn + block_scan_flag(v != BG_UNSET, s_wsum, tot);
                    if (v != BG_UNSET) plnext[o] = v < f ? plcur[v] : v;
                    outn += tot;
                }
                __syncthreads();
                uint32_t* tmp = plcur; plcur = plnext; plnext = tmp;
                nb = nbn; f = hi;
            }
            const uint32_t* ord = plcur;                 // fresh-key index per occupied bucket = HashSet iteration order
            // ---- order export (parity hook) -----------------------------------------------------------------------
            if (a.order_n) {
                const uint32_t n = min(F, a.order_cap);
                for (uint32_t i = tid; i < n; i += BG_THREADS) {
                    const uint32_t e = flist[ord[i]], ipos = e & 0x7FFFFFFFu;
                    uint32_t mt = 0;
                    for (uint32_t j = 1; j < nm; j++) if (s_moffs[j] <= ipos) mt = j;
                    a.order_seq[r * (uint64_t)a.order_cap + i] = (uint8_t)(mt | (mini ? (e >> 31) << 7 : 0u));
                    a.order_pos[r * (uint64_t)a.order_cap + i] = ipos - s_moffs[mt];
                }
                if (tid == 0) a.order_n[r] = n;
            }
            // ---- 5. vote ------------------------------------------------------------------------------------------
            if (a.rep_n && F) {
                const uint32_t H = a.H, N = a.N, Wp = a.Wp, B = a.start_sample;
                const bool classic = B == 0;
                for (uint32_t i = tid; i < N; i += BG_THREADS) { cnt[i] = 0; fstep[i] = BG_UNSET; }
                for (uint32_t i = tid; i < Wp; i += BG_THREADS) cand[i] = 0;
                __syncthreads();
                for (uint32_t c0 = 0; c0 < F && !miss; c0 += BG_THREADS) {
                    const uint32_t batch = min((uint32_t)BG_THREADS, F - c0);
                    if (tid == 0) s_u[1] = batch;
                    __syncthreads();
                    if ((uint32_t)tid < batch) {
                        const uint32_t e = flist[ord[c0 + tid]], ipos = e & 0x7FFFFFFFu;
                        const bool fwd = e >> 31;
                        uint8_t str[32];
                        for (uint32_t j = 0; j < len_item; j++) str[j] = (uint8_t)item_byte(seq, ipos, len_item, fwd, mini, j);
                        const HashIn in = hashin_from_bytes(str, len_item);
                        bool absent = false;
                        for (uint32_t h = 0; h < H; h++) {
                            const uint32_t rid = (uint32_t)hash_row(in, len_item, h, a.mods);
                            rid_s[tid * H + h] = rid;
                            if (!((a.rownz[rid >> 5] >> (rid & 31)) & 1u)) absent = true;     // read_id_mt_pe.rs:121-128 `None => break`
                        }
                        if (absent) atomicMin(&s_u[1], (uint32_t)tid);
                    }
                    __syncthreads();
                    const uint32_t p_local = s_u[1];
                    if (p_local < batch) miss = true;
                    nproc += p_local < batch ? p_local + 1 : batch;
                    // the k-mers before the first absent row: seeding steps first (they extend the candidate set and the
                    // report), then the steps that only count candidate colours (:155-161)
                    for (int pass = 0; pass < 2; pass++) {
                        for (uint32_t i = warp; i < p_local; i += BG_THREADS / 32) {
                            const uint32_t j = c0 + i;
                            const bool seeding = classic || j < B;
                            if (seeding != (pass == 0)) continue;
                            for (uint32_t w = lane; w < Wp; w += 32) {
                                uint32_t x = 0xFFFFFFFFu;
                                for (uint32_t h = 0; h < H; h++) x &= __ldg(a.rows + (uint64_t)rid_s[i * H + h] * Wp + w);
                                if (seeding) { if (x) atomicOr(&cand[w], x); }
                                else x &= cand[w];
                                while (x) {
                                    const uint32_t c = w * 32 + (__ffs(x) - 1);
                                    x &= x - 1;
                                    if (c < N) {
                                        atomicAdd(&cnt[c], 1u);
                                        if (seeding) atomicMin(&fstep[c], j);
                                    }
                                }
                            }
                        }
                        __syncthreads();
                    }
                }
                // report in final_report insertion order: (first step, colour) ascending; the "no hit" key N goes last
                if (tid == 0) s_u[2] = 0;
                __syncthreads();
                for (uint32_t c = tid; c < N; c += BG_THREADS)
                    if (fstep[c] != BG_UNSET) skeys[atomicAdd(&s_u[2], 1u)] = ((unsigned long long)fstep[c] << 32) | c;
                __syncthreads();
                const uint32_t R = s_u[2];
                uint32_t P2 = 1;
                while (P2 < R) P2 <<= 1;
                for (uint32_t i = R + tid; i < P2; i += BG_THREADS) skeys[i] = ~0ull;
                __syncthreads();
                for (uint32_t sz = 2; sz <= P2; sz <<= 1)
                    for (uint32_t st = sz >> 1; st > 0; st >>= 1) {
                        for (uint32_t i = tid; i < P2; i += BG_THREADS) {
                            const uint32_t jx = i ^ st;
                            if (jx > i) {
                                const unsigned long long x = skeys[i], y = skeys[jx];
                                const bool up = (i & sz) == 0;
                                if ((x > y) == up) { skeys[i] = y; skeys[jx] = x; }
                            }
                        }
                        __syncthreads();
                    }
                uint32_t* rc = a.rep_colour + r * (uint64_t)a.rep_cap;
                uint32_t* rv = a.rep_count + r * (uint64_t)a.rep_cap;
                for (uint32_t i = tid; i < R && i < a.rep_cap; i += BG_THREADS) {
                    const uint32_t c = (uint32_t)skeys[i], step = (uint32_t)(skeys[i] >> 32);
                    if (a.with_steps && step >= (1u << (32 - BG_STEP_SHIFT))) atomicOr(a.err, ERRF_LIST_OVERFLOW);
                    rc[i] = a.with_steps ? (c | (step << BG_STEP_SHIFT)) : c;
                    rv[i] = cnt[c];
                }
                total_rep = R + (miss ? 1u : 0u);
                if (tid == 0 && miss && R < a.rep_cap) { rc[R] = N; rv[R] = 1; }
                __syncthreads();
            }
        }
        if (tid == 0) {
            a.n_set[r] = F;
            if (a.rep_n) a.rep_n[r] = min(total_rep, a.rep_cap);
            a.flags[r] = fl | (total_rep > a.rep_cap ? 4u : 0u) | (min(nproc, 0xFFFFu) << 8);
            if (a.order_n && !ok) a.order_n[r] = 0;
        }
    }
}

size_t readid_big_smem(uint32_t H) {
    return ((tile_smem_bytes(BG_TCAP) + 15) & ~(size_t)15) + (size_t)BG_MAX_COLOURS * 8 + 128 * 4 +
           std::max((size_t)BG_THREADS * H * 4, (size_t)BG_MAX_COLOURS * 8);
}

// bytes of general-path scratch for reads of up to max_bases bases / max_kmers k-mer positions, and the CTAs that share it
void readid_big_plan(const cid_index* idx, uint32_t max_bases, uint32_t max_kmers, size_t budget, size_t* bytes, uint32_t* ctas) {
    const BigLayout L = big_layout((max_bases + 63) & ~63u, std::max(max_kmers, 1u));
    uint64_t g = std::max<uint64_t>(1, budget / L.total);
    g = std::min<uint64_t>(g, (uint64_t)idx->ctx->sm_count);
    *ctas = (uint32_t)g;
    *bytes = L.total * g;
}

int launch_readid_big(cid_index* idx, cudaStream_t st, const ReadSrc& src,
                      const uint64_t* d_seq_offs, const uint64_t* d_read_offs, uint64_t r0, const uint32_t* d_list,
                      const uint32_t* d_list_n, uint32_t max_bases, uint32_t max_kmers, const cid_readid_params& p,
                      uint8_t* d_scratch, uint32_t ctas, uint32_t* d_n_set, uint32_t* d_flags, uint32_t* d_rep_n,
                      uint32_t* d_rep_colour, uint32_t* d_rep_count, uint32_t order_cap, uint32_t* d_order_n,
                      uint8_t* d_order_seq, uint32_t* d_order_pos) {
    cid_ctx* ctx = idx->ctx;
    if (max_bases >= (1u << 23)) { set_error("read_id: reads of 2^23 bases or more are not supported"); return CID_E_UNSUPPORTED; }
    if (d_rep_n && idx->N > BG_MAX_COLOURS) { set_error("read_id: more than %u accessions per shard not supported", BG_MAX_COLOURS); return CID_E_UNSUPPORTED; }
    BigArgs a{};
    a.src = src;
    a.seq_offs = d_seq_offs; a.read_offs = d_read_offs; a.r0 = r0;
    a.list = d_list; a.list_n = d_list_n;
    a.k = idx->k; a.mini_m = idx->m; a.d = p.downsample; a.H = idx->H; a.mods = make_mods(idx->S, idx->hv, (const HashCfg*)idx->d_hcfg);
    a.rows = idx->rows; a.rownz = idx->rownz; a.N = idx->N; a.Wp = idx->Wp;
    a.start_sample = p.start_sample; a.rep_cap = p.rep_cap; a.with_steps = ctx->opt_readid_report_steps ? 1u : 0u;
    a.gw = p.group_width; a.rbf = p.reserve_before_find;
    a.scratch = d_scratch; a.lay = big_layout((max_bases + 63) & ~63u, std::max(max_kmers, 1u));
    a.n_set = d_n_set; a.flags = d_flags; a.rep_n = d_rep_n; a.rep_colour = d_rep_colour; a.rep_count = d_rep_count;
    a.order_cap = order_cap; a.order_n = d_order_n; a.order_seq = d_order_seq; a.order_pos = d_order_pos;
    a.err = ctx->d_err;
    const size_t smem = readid_big_smem(idx->H);
    bool& attr = ctx->attr_done[5];
    if (!attr) {
        CID_CUDA(cudaFuncSetAttribute(readid_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    ProfScope ps(ctx, st, KID_READID_BIG);
    readid_big_kernel<<<ctas, BG_THREADS, smem, st>>>(a);
    ctx->launches++;
    CID_CUDA(cudaGetLastError());
    return CID_OK;
}

}  // namespace cid
