// Shared by the read_id kernels (cid_readid.cu, cid_readid_part.cu): the warp-private read tile loader.
#pragma once
#include "cid_device.cuh"
#include "cid_internal.h"

namespace cid {

constexpr int RA_WARPS = 4;
constexpr int MAX_MATES = 8;

struct ReadGeom { uint64_t b0; int len; int nm; };

// Where a read's bases come from: ASCII bases (+ qualities for seq.rs:36-56 qual_mask on the device), or the planes a host
// packer produced (cid_pack_reads: per read 2-bit codes, "not ACGTacgt / masked" bits and optionally lower-case bits, in the
// tile's own layout, so loading is a copy and the per-kernel mask + pack work disappears along with 3/4 of the H2D bytes).

// Loads one read (all mates, contiguous in `bases`) into a warp-private tile, applying
// seq.rs:36-56 qual_mask when quals != nullptr.  Returns false if the read does not fit.
__device__ __forceinline__ bool warp_load_read(Tile& t, int cap, const ReadSrc& src_, uint64_t r,
                                               const uint64_t* __restrict__ seq_offs, uint64_t s_begin, uint64_t s_end,
                                               uint32_t* moffs, int lane, ReadGeom& g) {
    const uint8_t* __restrict__ bases = src_.bases;
    const uint8_t* __restrict__ quals = src_.quals;
    const uint32_t maxq = src_.maxq;
    g.b0 = __ldg(seq_offs + s_begin);
    uint64_t b1 = __ldg(seq_offs + s_end);
    g.nm = (int)(s_end - s_begin);
    g.len = (int)min(b1 - g.b0, (uint64_t)0x7fffffff);
    if (b1 - g.b0 > (uint64_t)cap - 32 || g.nm > MAX_MATES) return false;
    t.len = g.len;
    for (int i = lane; i < cap / 32 + 2; i += 32) t.start[i] = 0;
    __syncwarp();
    if (lane <= g.nm && lane < MAX_MATES + 1) {
        uint32_t o = (uint32_t)(__ldg(seq_offs + s_begin + lane) - g.b0);
        moffs[lane] = o;
        if (lane > 0 && lane < g.nm) atomicOr(&t.start[o >> 5], 1u << (o & 31));
    }
    if (src_.pk) {
        // packed planes: codes | bad | [lower], one copy each (bits past the read's end are "bad" by construction)
        const uint32_t* __restrict__ pw = src_.pk + __ldg(src_.pk_offs + r);
        const int ncw = (g.len + 15) >> 4, nbw = (g.len + 31) >> 5;
        for (int w = lane; w < cap / 16 + 3; w += 32) t.codes[w] = w < ncw ? __ldcs(pw + w) : 0u;
        for (int w = lane; w < cap / 32 + 2; w += 32) {
            t.bad[w] = w < nbw ? __ldcs(pw + ncw + w) : 0xFFFFFFFFu;
            t.lower[w] = (src_.pk_lower && w < nbw) ? __ldcs(pw + ncw + nbw + w) : 0u;
        }
        __syncwarp();
        if (src_.pk_lower) {      // (rare) raw-case consumers read the bytes: spell them from the planes ('N' for any non-base)
            for (int i = lane; i < g.len; i += 32) {
                uint32_t c = code_ascii((t.codes[i >> 4] >> (30 - 2 * (i & 15))) & 3u);
                if ((t.lower[i >> 5] >> (i & 31)) & 1u) c |= 0x20u;
                if ((t.bad[i >> 5] >> (i & 31)) & 1u) c = 'N';
                t.ascii[i] = (uint8_t)c;
            }
            __syncwarp();
        }
        return true;
    }
    const uint8_t* src = bases + g.b0;
    const uint8_t* qsrc = quals ? quals + g.b0 : nullptr;
    // streamed once (__ldcs): do not displace matrix rows in L2.  Word loads when both streams are 4-byte aligned.
    int done = 0;
    if (((uintptr_t)src & 3) == 0 && (!qsrc || (((uintptr_t)qsrc & 3) == 0 && maxq <= 255u))) {
        const int nw = g.len >> 2;
        const uint32_t maxq4 = maxq * 0x01010101u;
        for (int w = lane; w < nw; w += 32) {
            uint32_t c4 = __ldcs((const uint32_t*)src + w);
            if (qsrc) {
                const uint32_t low = __vcmpltu4(__ldcs((const uint32_t*)qsrc + w), maxq4);   // 0xFF where qual < maxq
                c4 = (c4 & ~low) | (0x4E4E4E4Eu & low);                                      // 'N'
            }
            ((uint32_t*)t.ascii)[w] = c4;
        }
        done = nw << 2;
    }
    for (int i = done + lane; i < g.len; i += 32) {
        uint8_t c = __ldcs(src + i);
        if (qsrc && (uint32_t)__ldcs(qsrc + i) < maxq) c = 'N';
        t.ascii[i] = c;
    }
    __syncwarp();
    tile_pack(t, cap, lane, 32);
    __syncwarp();
    return true;
}

}  // namespace cid
