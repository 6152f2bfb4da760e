"""ctypes binding of the CPU oracle (oracle/liboracle.so).

ORACLE — TEST INFRASTRUCTURE ONLY.  Import this from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never from colorid_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_xxh3_64.restype = C.c_uint64
        L.orc_xxh3_64.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64]
        L.orc_xxh3_64_variant.restype = C.c_uint64
        L.orc_xxh3_64_variant.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint32]
        L.orc_fnv1a_str.restype = C.c_uint64
        L.orc_fnv1a_str.argtypes = [C.c_char_p, C.c_uint64]
        L.orc_fnv1a_usize.restype = C.c_uint64
        L.orc_fnv1a_usize.argtypes = [C.c_uint64]
        L.orc_binomial_mass.restype = C.c_double
        L.orc_binomial_mass.argtypes = [C.c_uint64, C.c_double, C.c_uint64]
        L.orc_false_prob.restype = C.c_double
        L.orc_false_prob.argtypes = [C.c_double, C.c_double, C.c_double]
        L.orc_qual_mask.restype = C.c_int64
        L.orc_hashset_str_order.restype = C.c_int64
        L.orc_hashmap_usize_order.restype = C.c_int64
        L.orc_kmap_new.restype = C.c_void_p
        L.orc_kmap_len.restype = C.c_uint64
        L.orc_kmap_auto_cutoff.restype = C.c_int64
        L.orc_auto_cutoff_histo.restype = C.c_int64
        L.orc_index_new.restype = C.c_void_p
        L.orc_index_new.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_index_words.restype = u32p
        L.orc_index_row_words.restype = C.c_uint32
        L.orc_index_nonzero_rows.restype = C.c_uint64
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t)


def pack_seqs(seqs):
    """list of bytes -> (uint8 array of concatenated bases, uint64 offsets[n+1])."""
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if seqs else np.zeros(0, np.uint8)
    if bases.size == 0:
        bases = np.zeros(1, np.uint8)
    return bases, offs


def xxh3_64(b, seed=0, variant=0):
    """variant: see oracle/xxh3_ref.hpp (0 = stable XXH3)."""
    if variant:
        return lib().orc_xxh3_64_variant(bytes(b), len(b), seed, variant)
    return lib().orc_xxh3_64(bytes(b), len(b), seed)


def fnv1a_str(b):
    return lib().orc_fnv1a_str(bytes(b), len(b))


def fnv1a_usize(v):
    return lib().orc_fnv1a_usize(v)


def binomial_mass(n, p, x):
    return lib().orc_binomial_mass(n, p, x)


def false_prob(m, k, n):
    return lib().orc_false_prob(float(m), float(k), float(n))


def revcomp(s):
    out = C.create_string_buffer(len(s))
    lib().orc_revcomp(C.c_char_p(bytes(s)), C.c_uint64(len(s)), out)
    return out.raw


def has_no_n(s):
    return bool(lib().orc_has_no_n(C.c_char_p(bytes(s)), C.c_uint64(len(s))))


def qual_mask(seq, qual, off):
    out = C.create_string_buffer(max(len(qual), len(seq), 1))
    n = lib().orc_qual_mask(C.c_char_p(bytes(seq)), C.c_uint64(len(seq)), C.c_char_p(bytes(qual)),
                            C.c_uint64(len(qual)), C.c_uint8(off), out)
    if n < 0:
        return None
    return out.raw[:n]


def hashset_str_order(keys, group_width=16, reserve_before_find=True):
    """Insertion sequence of byte strings -> (order of first-occurrence indices, final buckets)."""
    bases, offs = pack_seqs(list(keys))
    order = np.zeros(len(keys) + 1, dtype=np.int32)
    buckets = C.c_uint64(0)
    n = lib().orc_hashset_str_order(_p(bases, C.c_char_p), _p(offs, u64p), C.c_uint64(len(keys)),
                                    C.c_int(group_width), C.c_int(int(reserve_before_find)), _p(order, i32p),
                                    C.byref(buckets))
    return order[:n].copy(), buckets.value


def hashmap_usize_order(keys, group_width=16):
    k = np.asarray(keys, dtype=np.uint64)
    order = np.zeros(len(k) + 1, dtype=np.int32)
    n = lib().orc_hashmap_usize_order(_p(k, u64p), C.c_uint64(len(k)), C.c_int(group_width), _p(order, i32p))
    return order[:n].copy()


MODE_FASTA, MODE_FASTQ, MODE_STRING, MODE_READSET = 0, 1, 2, 3


class KMap:
    """Count map of canonical k-mers (kmer.rs kmerize_* semantics selected by `mode`)."""

    def __init__(self, k):
        self.k = k
        self.h = C.c_void_p(lib().orc_kmap_new())

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_kmap_free(self.h)
            self.h = None

    def add(self, seqs, mode, d=1):
        bases, offs = pack_seqs(list(seqs))
        return lib().orc_kmap_add(self.h, _p(bases, C.c_char_p), _p(offs, u64p), C.c_uint64(len(seqs)),
                                  C.c_uint32(self.k), C.c_uint32(d), C.c_int(mode))

    def __len__(self):
        return lib().orc_kmap_len(self.h)

    def auto_cutoff(self):
        return lib().orc_kmap_auto_cutoff(self.h)

    def clean(self, t):
        lib().orc_kmap_clean(self.h, C.c_uint64(t))

    def export(self):
        n = len(self)
        keys = np.zeros(max(n * self.k, 1), dtype=np.uint8)
        counts = np.zeros(max(n, 1), dtype=np.uint64)
        lib().orc_kmap_export(self.h, C.c_uint32(self.k), _p(keys, C.c_char_p), _p(counts, u64p))
        return keys[: n * self.k].reshape(n, self.k), counts[:n]


def find_minimizer(seq, m):
    """kmer.rs:971-986 find_minimizer(seq, m) -> bytes (None where the reference panics: m > len(seq))."""
    out = C.create_string_buffer(max(m, 1))
    rc = lib().orc_find_minimizer(C.c_char_p(bytes(seq)), C.c_uint64(len(seq)), C.c_uint32(m), out)
    return None if rc != 0 else out.raw[:m]


def minimizer_map(seqs, k, m, d=1, upper=True):
    """Minimizer count map (kmer.rs:328-361 upper=True / :694-824 upper=False) -> (keys [n, m] uint8, counts)."""
    km = KMap(m)
    bases, offs = pack_seqs(list(seqs))
    rc = lib().orc_kmap_add_minimizers(km.h, _p(bases, C.c_char_p), _p(offs, u64p), C.c_uint64(len(seqs)),
                                       C.c_uint32(k), C.c_uint32(m), C.c_uint32(d), C.c_int(int(upper)))
    if rc != 0:
        raise RuntimeError("reference would panic (m > k)")
    return km


def auto_cutoff_histo(histo):
    cov = np.array(list(histo.keys()), dtype=np.uint64)
    num = np.array(list(histo.values()), dtype=np.uint64)
    return lib().orc_auto_cutoff_histo(_p(cov, u64p), _p(num, u64p), C.c_uint64(len(cov)))


class Index:
    """Dense BIGSI index model on the CPU (bigsi.rs:19-27 semantics; absent row == zero row)."""

    def __init__(self, bloom_size, num_hash, k, n_colours, m=0, hash_variant=0):
        self.S, self.H, self.k, self.N, self.m = bloom_size, num_hash, k, n_colours, m
        self.hash_variant = hash_variant
        self.W = (n_colours + 31) // 32
        self.h = C.c_void_p(lib().orc_index_new(bloom_size, num_hash, k, n_colours))
        self.n_ref = np.zeros(n_colours, dtype=np.uint64)
        if m:
            lib().orc_index_set_minimizer(self.h, C.c_uint32(m))     # .mxi: bigsi.rs:40-49 m_size
        if hash_variant:
            lib().orc_index_set_hash_variant(self.h, C.c_uint32(hash_variant))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_index_free(self.h)
            self.h = None

    def words(self):
        """numpy view [S, W] uint32 of the dense matrix (valid while the index lives)."""
        ptr = lib().orc_index_words(self.h)
        return np.ctypeslib.as_array(ptr, shape=(self.S, self.W))

    def build_accession(self, colour, seqs, mode, cutoff=-1):
        bases, offs = pack_seqs(list(seqs))
        n_ref = C.c_uint64(0)
        used = C.c_int64(0)
        rc = lib().orc_build_accession(self.h, C.c_uint32(colour), _p(bases, C.c_char_p), _p(offs, u64p),
                                       C.c_uint64(len(seqs)), C.c_int(mode), C.c_int64(cutoff), C.byref(n_ref),
                                       C.byref(used))
        if rc != 0:
            raise RuntimeError("reference would panic (auto_cutoff on a degenerate histogram)")
        self.n_ref[colour] = n_ref.value
        return n_ref.value, used.value

    def build_accession_mini(self, colour, seqs, mode, cutoff=-1, variant=0):
        """variant 0 = build_single_mini (build.rs:396-492), 1 = build_multi_mini (build.rs:258-394)."""
        bases, offs = pack_seqs(list(seqs))
        n_ref = C.c_uint64(0)
        used = C.c_int64(0)
        rc = lib().orc_build_accession_mini(self.h, C.c_uint32(colour), _p(bases, C.c_char_p), _p(offs, u64p),
                                            C.c_uint64(len(seqs)), C.c_int(mode), C.c_int64(cutoff), C.c_int(variant),
                                            C.byref(n_ref), C.byref(used))
        if rc != 0:
            raise RuntimeError("reference would panic")
        self.n_ref[colour] = n_ref.value
        return n_ref.value, used.value

    def finalize(self, threads=1):
        lib().orc_build_finalize(self.h, C.c_int(threads))

    def build_many(self, accessions, mode, cutoff=-1, threads=1):
        """accessions: list of lists of bytes.  Colour = list position."""
        flat = [s for acc in accessions for s in acc]
        bases, offs = pack_seqs(flat)
        acc_offs = np.zeros(len(accessions) + 1, dtype=np.uint64)
        acc_offs[1:] = np.cumsum([len(a) for a in accessions], dtype=np.uint64)
        rc = lib().orc_build_many(self.h, _p(bases, C.c_char_p), _p(offs, u64p), _p(acc_offs, u64p),
                                  C.c_uint64(len(accessions)), C.c_int(mode), C.c_int64(cutoff), C.c_int(threads),
                                  _p(self.n_ref, u64p))
        if rc != 0:
            raise RuntimeError("reference would panic")

    def nonzero_rows(self):
        return lib().orc_index_nonzero_rows(self.h)

    def query_counts(self, queries, seq_mode=MODE_FASTA, gene_search=False, filt=-1):
        """queries: list of lists of bytes (one query = one file's sequences)."""
        flat = [s for q in queries for s in q]
        bases, offs = pack_seqs(flat)
        qoffs = np.zeros(len(queries) + 1, dtype=np.uint64)
        qoffs[1:] = np.cumsum([len(q) for q in queries], dtype=np.uint64)
        nq = len(queries)
        counts = np.zeros((nq, self.N), dtype=np.uint32)
        num_kmers = np.zeros(nq, dtype=np.uint64)
        un = np.zeros((nq, self.N), dtype=np.uint64)
        us = np.zeros((nq, self.N), dtype=np.uint64)
        um = np.zeros((nq, self.N), dtype=np.uint64)
        used = np.zeros(nq, dtype=np.int64)
        rc = lib().orc_query_counts(self.h, _p(bases, C.c_char_p), _p(offs, u64p), _p(qoffs, u64p), C.c_uint64(nq),
                                    C.c_int(seq_mode), C.c_int(int(gene_search)), C.c_int64(filt), _p(counts, u32p),
                                    _p(num_kmers, u64p), _p(un, u64p), _p(us, u64p), _p(um, u64p), _p(used, i64p))
        if rc != 0:
            raise RuntimeError("reference would panic")
        return dict(counts=counts, num_kmers=num_kmers, uniq_n=un, uniq_sum=us, uniq_mode=um, cutoff=used)

    def query_perfect(self, queries, mf=False):
        """mf=False: queries is a list of lists of bytes; mf=True: a list of bytes (one record each)."""
        if mf:
            flat = list(queries)
            qoffs = np.arange(len(flat) + 1, dtype=np.uint64)
        else:
            flat = [s for q in queries for s in q]
            qoffs = np.zeros(len(queries) + 1, dtype=np.uint64)
            qoffs[1:] = np.cumsum([len(q) for q in queries], dtype=np.uint64)
        bases, offs = pack_seqs(flat)
        nq = len(queries)
        and_rows = np.zeros((nq, self.W), dtype=np.uint32)
        status = np.zeros(nq, dtype=np.uint8)
        n_kmers = np.zeros(nq, dtype=np.uint64)
        lib().orc_query_perfect(self.h, _p(bases, C.c_char_p), _p(offs, u64p), _p(qoffs, u64p), C.c_uint64(nq),
                                C.c_int(int(mf)), _p(and_rows, u32p), _p(status, u8p), _p(n_kmers, u64p))
        return dict(and_rows=and_rows, status=status, n_kmers=n_kmers)

    def read_id_batch(self, reads, d=1, start_sample=3, fp_correct=1e-3, group_width=16, reserve_before_find=True,
                      threads=1, n_ref=None, top_cap=8, rep_cap=None, order_cap=0):
        """reads: list of lists of bytes (1 or 2 mates per read, already quality-masked)."""
        flat = [s for r in reads for s in r]
        bases, offs = pack_seqs(flat)
        roffs = np.zeros(len(reads) + 1, dtype=np.uint64)
        roffs[1:] = np.cumsum([len(r) for r in reads], dtype=np.uint64)
        nr = len(reads)
        if rep_cap is None:
            rep_cap = self.N + 1
        n_ref = self.n_ref if n_ref is None else np.asarray(n_ref, dtype=np.uint64)
        n_set = np.zeros(nr, np.uint32)
        kind = np.zeros(nr, np.int32)
        hits = np.zeros(nr, np.uint32)
        n_top = np.zeros(nr, np.uint32)
        top = np.zeros((nr, top_cap), np.uint32)
        rep_n = np.zeros(nr, np.uint32)
        rep_c = np.zeros((nr, rep_cap), np.uint32)
        rep_v = np.zeros((nr, rep_cap), np.uint32)
        ord_n = np.zeros(nr, np.uint32)
        ord_s = np.zeros((nr, max(order_cap, 1)), np.uint8)
        ord_p = np.zeros((nr, max(order_cap, 1)), np.uint32)
        lib().orc_read_id_batch(
            self.h, _p(bases, C.c_char_p), _p(offs, u64p), _p(roffs, u64p), C.c_uint64(nr), C.c_uint32(d),
            C.c_uint32(start_sample), _p(n_ref, u64p), C.c_double(fp_correct), C.c_int(group_width),
            C.c_int(int(reserve_before_find)), C.c_int(threads), _p(n_set, u32p), _p(kind, i32p), _p(hits, u32p),
            _p(n_top, u32p), _p(top, u32p), C.c_uint32(top_cap), _p(rep_n, u32p), _p(rep_c, u32p), _p(rep_v, u32p),
            C.c_uint32(rep_cap), _p(ord_n, u32p) if order_cap else None, _p(ord_s, u8p), _p(ord_p, u32p),
            C.c_uint32(order_cap))
        return dict(n_set=n_set, kind=kind, hits=hits, n_top=n_top, top=top, rep_n=rep_n, rep_colour=rep_c,
                    rep_count=rep_v, order_n=ord_n, order_seq=ord_s, order_pos=ord_p)


CLS_TOO_SHORT, CLS_NO_HITS, CLS_NO_SIG, CLS_ACCEPT, CLS_REJECT_MULTI, CLS_PANIC = range(6)


def kmer_poll_plus(rep_colour, rep_count, n_set, fp_by_colour, fp_correct=1e-3, top_cap=8):
    rc = np.asarray(rep_colour, np.uint32)
    rv = np.asarray(rep_count, np.uint32)
    fp = np.asarray(fp_by_colour, np.float64)
    kind = C.c_int32(0)
    hits = C.c_uint32(0)
    n_top = C.c_uint32(0)
    top = np.zeros(top_cap, np.uint32)
    lib().orc_kmer_poll_plus(_p(rc, u32p), _p(rv, u32p), C.c_uint32(len(rc)), C.c_uint64(n_set), _p(fp, f64p),
                             C.c_uint64(len(fp)), C.c_double(fp_correct), C.byref(kind), C.byref(hits),
                             C.byref(n_top), _p(top, u32p), C.c_uint32(top_cap))
    return kind.value, hits.value, n_top.value, top[: min(n_top.value, top_cap)].copy()


def read_fasta(path):
    """kmer.rs:10-45 read_fasta: header = any line containing '>'; contigs concatenated."""
    with open(path, "rb") as f:
        contents = f.read()
    lines = contents.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()                      # str::lines() has no trailing empty item
    lines = [l[:-1] if l.endswith(b"\r") else l for l in lines]
    vec, sub = [], []                    # (sub: the pieces of the current contig; joined once, bytes += is quadratic)
    n = len(lines)
    for i, line in enumerate(lines, start=1):
        if b">" in line:
            if any(sub):
                vec.append(b"".join(sub))
            sub = []
        elif i == n:
            sub.append(line)
            if any(sub):
                vec.append(b"".join(sub))
        else:
            sub.append(line)
    return vec


def read_fasta_mf(path):
    """kmer.rs:47-84 read_fasta_mf: (labels, sequences); label = header line minus its first byte."""
    with open(path, "rb") as f:
        contents = f.read()
    lines = contents.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    lines = [l[:-1] if l.endswith(b"\r") else l for l in lines]
    labels, vec, sub = [], [], b""
    n = len(lines)
    for i, line in enumerate(lines, start=1):
        if b">" in line:
            labels.append(line[1:])
            if sub:
                vec.append(sub)
            sub = b""
        elif i == n:
            sub += line
            if sub:
                vec.append(sub)
        else:
            sub += line
    return labels, vec
