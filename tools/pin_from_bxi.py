#!/usr/bin/env python3
"""Which hash variant reproduces a given .bxi?

The reference hashes k-mers with the un-vendored crate `xxh3 = "0.1.1"` (Cargo.toml:9, simple_bloom.rs:21-24), which
predates the XXH3 freeze; the library, the oracle and every kernel therefore take a `hash variant` (include/colorid_b200.h:
0 = stable XXH3, bits for the places where the published drafts differ).  Given a .bxi written by ANY colorid build
(the original binary, or colorid-b200) together with the FASTA accessions it was built from, this tool rebuilds the set
of stored rows under each of the 32 variants and reports the one(s) that reproduce the file -- e.g. for the reference's
own test index:

    colorid build -s 750000 -n 4 -k 27 -b phage -r test_data/ref_file.txt        # (real binary, elsewhere)
    python tools/pin_from_bxi.py phage.bxi test_data/ref_file.txt

Self-contained numpy restatement (kmer.rs:10-45 read_fasta, :87-125 kmerize_vector, simple_bloom.rs:19-26): it shares no
code with the oracle or the product, so a match is a third independent confirmation.  FASTA accessions only.
"""
import argparse
import gzip
import struct
import sys

import numpy as np

M64 = (1 << 64) - 1
P64_1, P64_2, P64_3 = 0x9E3779B185EBCA87, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9
P_MX1, P_MX2 = 0x165667919E3779F9, 0x9FB21C651E98DF25
SECRET = bytes.fromhex(
    "b8fe6c3923a44bbe7c01812cf721ad1cded46de9839097db7240a4a4b7b3671fcb79e64eccc0e578825ad07dccff7221"
    "b8084674f743248ee03590e6813a264c")
VARIANT_NAMES = {0: "stable XXH3 (xxHash >= 0.8)", 1: "XXH3 draft of xxHash 0.7.1-0.7.3 (as recalled)",
                 31: "XXH3 draft of xxHash 0.7.0 (as recalled)"}
u64 = np.uint64


def _sec(off, variant):
    if variant & 16:      # the secret's bytes as an array of u32 constants
        lo, hi = struct.unpack_from(">II", SECRET, off)
        return u64((hi << 32) | lo)
    return u64(struct.unpack_from("<Q", SECRET, off)[0])


def _mul128(a, b):
    """(lo, hi) of the 128-bit products of two uint64 arrays."""
    m32 = u64(0xFFFFFFFF)
    a0, a1, b0, b1 = a & m32, a >> u64(32), b & m32, b >> u64(32)
    p00, p01, p10, p11 = a0 * b0, a0 * b1, a1 * b0, a1 * b1
    mid = (p00 >> u64(32)) + (p01 & m32) + (p10 & m32)
    lo = (p00 & m32) | (mid << u64(32))
    hi = p11 + (p01 >> u64(32)) + (p10 >> u64(32)) + (mid >> u64(32))
    return lo, hi


def xxh3_64(kmers, seed, variant):
    """kmers: uint8 [n, k] (k in 1..32) -> uint64 [n]: xxh3::hash64_with_seed(kmer, seed) under `variant`."""
    n, k = kmers.shape
    seed = u64(seed)
    with np.errstate(over="ignore"):
        def le(cols, width):
            v = np.zeros(n, u64)
            for i in range(width):
                v |= kmers[:, cols + i].astype(u64) << u64(8 * i)
            return v

        def fold(a, b):
            lo, hi = _mul128(a, b)
            return lo + hi if variant & 4 else lo ^ hi

        def aval(h):
            h = h ^ (h >> u64(29 if variant & 2 else 37))
            h = h * u64(P64_3 if variant & 1 else P_MX1)
            return h ^ (h >> u64(32))
        if k >= 17:
            sd = u64(0) if variant & 8 else seed
            acc = np.full(n, (u64(k) + (seed if variant & 8 else u64(0))) * u64(P64_1), u64)
            acc = acc + fold(le(0, 8) ^ (_sec(0, variant) + sd), le(8, 8) ^ (_sec(8, variant) - sd))
            acc = acc + fold(le(k - 16, 8) ^ (_sec(16, variant) + sd), le(k - 8, 8) ^ (_sec(24, variant) - sd))
            return aval(acc)
        if k >= 9:
            lo = le(0, 8) ^ ((_sec(24, variant) ^ _sec(32, variant)) + seed)
            hi = le(k - 8, 8) ^ ((_sec(40, variant) ^ _sec(48, variant)) - seed)
            return aval(u64(k) + lo.byteswap() + hi + fold(lo, hi))
        if k >= 4:
            s32 = int(seed) & 0xFFFFFFFF
            sd = u64(int(seed) ^ (int.from_bytes(s32.to_bytes(4, "little"), "big") << 32))
            h = (le(k - 4, 4) + (le(0, 4) << u64(32))) ^ ((_sec(8, variant) ^ _sec(16, variant)) - sd)
            rot = lambda x, r: (x << u64(r)) | (x >> u64(64 - r))
            h = h ^ rot(h, 49) ^ rot(h, 24)
            h = h * u64(P_MX2)
            h = h ^ ((h >> u64(35)) + u64(k))
            h = h * u64(P_MX2)
            return h ^ (h >> u64(28))
        s0 = int(_sec(0, variant))
        comb = (kmers[:, 0].astype(u64) << u64(16)) | (kmers[:, k >> 1].astype(u64) << u64(24)) | kmers[:, k - 1].astype(u64) | u64(k << 8)
        h = comb ^ (u64((s0 & 0xFFFFFFFF) ^ (s0 >> 32)) + seed)
        h = (h ^ (h >> u64(33))) * u64(P64_2)
        h = (h ^ (h >> u64(29))) * u64(P64_3)
        return h ^ (h >> u64(32))


def read_fasta(path):
    """kmer.rs:10-45: a header is any line containing '>'; the lines of a contig are concatenated."""
    data = (gzip.open(path, "rb") if str(path).endswith(".gz") else open(path, "rb")).read()
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    contigs, cur = [], []
    for line in lines:
        if line.endswith(b"\r"):
            line = line[:-1]
        if b">" in line:
            if any(cur):
                contigs.append(b"".join(cur))
            cur = []
        else:
            cur.append(line)
    if any(cur):
        contigs.append(b"".join(cur))
    return contigs


_COMP = np.full(256, ord("N"), np.uint8)
for _a, _b in zip(b"ACGTacgtUuNn", b"TGCAtgcaAaNn"):     # kmer.rs:847-863 switch_base
    _COMP[_a] = _b
_GOOD = np.zeros(256, bool)
_GOOD[list(b"ACGTacgt")] = True


def canonical_kmers(contigs, k):
    """kmer.rs:87-125 kmerize_vector: distinct canonical k-mers (raw-case compare, then upper case) as uint8 [n, k]."""
    out = []
    for c in contigs:
        a = np.frombuffer(c, np.uint8)
        if len(a) < k:
            continue
        win = np.lib.stride_tricks.sliding_window_view(a, k)
        ok = np.lib.stride_tricks.sliding_window_view(_GOOD[a], k).all(axis=1)            # seq.rs:66-70 has_no_n
        fwd = win[ok]
        rc = _COMP[fwd][:, ::-1]
        diff = fwd != rc
        first = diff.argmax(axis=1)
        rows = np.arange(len(fwd))
        take_fwd = diff.any(axis=1) & (fwd[rows, first] < rc[rows, first])               # `fwd < rc`, ties take rc
        km = np.where(take_fwd[:, None], fwd, rc)
        km = np.where((km >= 97) & (km <= 122), km - 32, km).astype(np.uint8)             # to_uppercase
        out.append(km)
    if not out:
        return np.zeros((0, k), np.uint8)
    return np.unique(np.concatenate(out), axis=0)


def read_bxi_rows(path):
    """bigsi.rs:19-27 in bincode 1.x: (bloom_size, num_hash, k_size, {colour: name}, row ids, words[n, W])."""
    d = open(path, "rb").read()
    S, H, k, ncol = struct.unpack_from("<QQQQ", d, 0)
    at = 32
    colors = {}
    for _ in range(ncol):
        c, ln = struct.unpack_from("<QQ", d, at)
        colors[c] = d[at + 16:at + 16 + ln].decode()
        at += 16 + ln
    W = (ncol + 31) // 32
    (nrows,) = struct.unpack_from("<Q", d, at)
    at += 8
    rec = np.dtype([("row", "<u8"), ("nw", "<u8"), ("w", "<u4", (W,)), ("nbits", "<u8")])
    rows = np.frombuffer(d, rec, nrows, at)
    return S, H, k, colors, rows["row"].copy(), rows["w"].reshape(nrows, W).copy()


def rows_for_variant(kmers_by_colour, S, H, variant, W):
    """{row id: words} of the index the builder would store (build.rs:116-128) under `variant`, as (ids, words)."""
    ids, cols = [], []
    for c, km in kmers_by_colour.items():
        for seed in range(H):
            r = np.unique(xxh3_64(km, seed, variant) % u64(S))
            ids.append(r)
            cols.append(np.full(len(r), c, np.uint32))
    ids, cols = np.concatenate(ids), np.concatenate(cols)
    uniq, inv = np.unique(ids, return_inverse=True)
    words = np.zeros((len(uniq), W), np.uint32)
    np.bitwise_or.at(words, (inv, cols // 32), (np.uint32(1) << (cols % 32).astype(np.uint32)))
    return uniq, words


def pin(bxi_path, ref_file, variants=range(32), log=None):
    S, H, k, colors, row_ids, words = read_bxi_rows(bxi_path)
    order = np.argsort(row_ids)
    row_ids, words = row_ids[order], words[order]
    files = {}
    for line in open(ref_file):
        f = line.rstrip("\n").split("\t")
        if len(f) >= 2:
            files[f[0]] = f[1]
    by_colour = {}
    for c, name in colors.items():
        if name not in files:
            raise SystemExit(f"accession {name} of the index is not in {ref_file}")
        by_colour[c] = canonical_kmers(read_fasta(files[name]), k)
    matches = []
    for v in variants:
        ids, w = rows_for_variant(by_colour, S, H, v, words.shape[1])
        same = len(ids) == len(row_ids) and np.array_equal(ids, row_ids.astype(ids.dtype)) and np.array_equal(w, words)
        if log:
            shared = len(np.intersect1d(ids, row_ids))
            print(f"variant {v:2d}: {len(ids)} rows, {shared} shared with the file's {len(row_ids)}" + ("  <-- reproduces the file" if same else ""), file=log)
        if same:
            matches.append(v)
    return matches


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("bxi")
    ap.add_argument("ref_file", help="the tab-delimited accession list the index was built from (FASTA accessions)")
    ap.add_argument("-q", "--quiet", action="store_true")
    a = ap.parse_args()
    m = pin(a.bxi, a.ref_file, log=None if a.quiet else sys.stderr)
    if not m:
        print("no variant reproduces this index: the hash of its builder is outside the known family")
        return 1
    for v in m:
        print(f"hash variant {v} ({VARIANT_NAMES.get(v, 'combination of draft features')}): set COLORID_B200_HASH_VARIANT={v} / cid_index_set_hash_variant(idx, {v})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
