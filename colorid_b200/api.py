"""Thin numpy-facing wrappers over the C ABI (include/colorid_b200.h).

These exist for the tests and bench.py; the product is the shared library.  Naming follows the
reference's domain: accessions, colours, queries, reads, rows.
"""
import ctypes as C

import numpy as np

from . import lib as L


def _p(a, t=L.vp):
    if a is None:
        return None
    return a.ctypes.data_as(t)


def pack_seqs(seqs):
    """list of bytes -> (uint8 bases, uint64 seq_offs[n+1])."""
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if seqs else np.zeros(0, np.uint8)
    if bases.size == 0:
        bases = np.zeros(1, np.uint8)
    return bases, offs


def group_offsets(groups):
    o = np.zeros(len(groups) + 1, dtype=np.uint64)
    if groups:
        o[1:] = np.cumsum([len(g) for g in groups], dtype=np.uint64)
    return o


class Context:
    def __init__(self, device=0):
        self.lib = L.load()
        h = L.vp()
        L.check(self.lib.cid_ctx_create(device, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.cid_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_option(self, name, value):
        L.check(self.lib.cid_ctx_set_option(self.h, name.encode(), int(value)))

    @property
    def launches(self):
        return self.lib.cid_ctx_launch_count(self.h)

    def read_counter(self, name):
        v = C.c_uint64(0)
        L.check(self.lib.cid_ctx_read_counter(self.h, name.encode(), C.byref(v)))
        return v.value

    def profile(self, enable=True):
        L.check(self.lib.cid_ctx_profile(self.h, int(enable)))

    def profile_read(self):
        """{kernel name: (total ms, launches)} accumulated since profile(True)."""
        out = {}
        k = 0
        while True:
            name, ms, n = C.c_char_p(), C.c_double(0), C.c_uint64(0)
            if self.lib.cid_ctx_profile_read(self.h, k, C.byref(name), C.byref(ms), C.byref(n)) != 0:
                break
            if n.value:
                out[name.value.decode()] = (ms.value, n.value)
            k += 1
        return out


CLS_NAMES = ["too_short", "no_hits", "no_significant_hits", "accept", "reject_multi", "ref_panic"]


def classify_reads(index_params, n_ref, rep, fp_correct=1e-3, group_width=16, threads=0, top_cap=8):
    """kmer_poll_plus over a batch (host threads). index_params = (bloom_size, num_hash, n_colours);
    rep = dict from Index.read_id_batch."""
    lib = L.load()
    S, H, N = index_params
    nr = len(rep["n_set"])
    n_ref = np.ascontiguousarray(n_ref, dtype=np.uint64)
    kind = np.zeros(max(nr, 1), np.int32)
    hits = np.zeros(max(nr, 1), np.uint32)
    n_top = np.zeros(max(nr, 1), np.uint32)
    top = np.zeros((max(nr, 1), top_cap), np.uint32)
    rc = np.ascontiguousarray(rep["rep_colour"])
    rv = np.ascontiguousarray(rep["rep_count"])
    L.check(lib.cid_classify_reads(S, H, N, _p(n_ref, L.u64p), fp_correct, group_width, nr,
                                   _p(np.ascontiguousarray(rep["n_set"]), L.u32p),
                                   _p(np.ascontiguousarray(rep["flags"]), L.u32p),
                                   _p(np.ascontiguousarray(rep["rep_n"]), L.u32p), _p(rc, L.u32p), _p(rv, L.u32p),
                                   rc.shape[1] if rc.ndim == 2 else 1, threads, _p(kind, C.POINTER(C.c_int32)),
                                   _p(hits, L.u32p), _p(n_top, L.u32p), _p(top, L.u32p), top_cap))
    return dict(kind=kind[:nr], hits=hits[:nr], n_top=n_top[:nr], top=top[:nr])


def merge_shard_reports(shard_reports, shards, n_total, rep_cap=None):
    """Column-sharded read_id: per-shard reports (dicts from Index.read_id_batch run with the context option
    readid_report_steps = 1, in shard order) -> the report of the unsharded index (cid_merge_shard_reports).
    shards = [(c_lo, c_hi)] accession ranges (sharding.column_shards)."""
    lib = L.load()
    ns = len(shard_reports)
    nr = len(shard_reports[0]["n_set"])
    cap_in = max(r["rep_colour"].shape[1] for r in shard_reports)      # shards of different widths: pad to one row stride
    cap_out = rep_cap if rep_cap else n_total + 1
    ncol = np.array([hi - lo for lo, hi in shards], np.uint32)
    coff = np.array([lo for lo, _ in shards], np.uint32)

    def padded(a):
        a = np.ascontiguousarray(a, dtype=np.uint32)
        if a.ndim == 2 and a.shape[1] < cap_in:
            a = np.ascontiguousarray(np.pad(a, ((0, 0), (0, cap_in - a.shape[1]))))
        return a
    keep = [[padded(r[k]) for r in shard_reports] for k in ("rep_n", "rep_colour", "rep_count")]
    arrs = [(L.u32p * ns)(*[_p(a, L.u32p) for a in ks]) for ks in keep]
    out_n = np.zeros(max(nr, 1), np.uint32)
    out_c = np.zeros((max(nr, 1), cap_out), np.uint32)
    out_v = np.zeros((max(nr, 1), cap_out), np.uint32)
    flags = np.ascontiguousarray(shard_reports[0]["flags"], dtype=np.uint32).copy()
    for r in shard_reports[1:]:
        # too_short / panic (bits 0, 1) and the set size are properties of the read: identical on every shard
        if not (np.array_equal(np.asarray(r["flags"]) & 3, flags & 3) and np.array_equal(r["n_set"], shard_reports[0]["n_set"])):
            raise ValueError("merge_shard_reports: the shards disagree on n_set / flags of a read (different reads or parameters?)")
    # a shard whose own report was cut at its rep_cap has lost colours: the merged report stays marked truncated (bit 2);
    # otherwise truncation is decided again for the merged report
    lost = np.zeros_like(flags)
    for r in shard_reports:
        lost |= np.asarray(r["flags"], dtype=np.uint32) & np.uint32(4)
    flags = (flags & ~np.uint32(4)) | lost
    if nr == 0:
        flags = np.zeros(1, np.uint32)
    L.check(lib.cid_merge_shard_reports(ns, _p(ncol, L.u32p), _p(coff, L.u32p), nr, arrs[0], arrs[1], arrs[2], cap_in, n_total,
                                        _p(out_n, L.u32p), _p(out_c, L.u32p), _p(out_v, L.u32p), cap_out, _p(flags, L.u32p)))
    return dict(n_set=shard_reports[0]["n_set"], flags=flags[:nr], rep_n=out_n[:nr], rep_colour=out_c[:nr], rep_count=out_v[:nr])


class Index:
    """Device-resident BIGSI index (bigsi.rs:19-27 BigsyMapNew)."""

    def __init__(self, ctx, bloom_size, num_hash, k, n_colours, m=0, hash_variant=0):
        self.ctx, self.lib = ctx, ctx.lib
        self.S, self.H, self.k, self.N, self.m = bloom_size, num_hash, k, n_colours, m
        h = L.vp()
        L.check(self.lib.cid_index_create(ctx.h, bloom_size, num_hash, k, n_colours, C.byref(h)))
        self.h = h
        if m:       # minimizer index (.mxi, bigsi.rs:40-49)
            L.check(self.lib.cid_index_set_minimizer(h, m))
        if hash_variant:   # which draft of XXH3 the rows were / are hashed with (include/colorid_b200.h; 0 = stable)
            L.check(self.lib.cid_index_set_hash_variant(h, hash_variant))
        self.W = self.lib.cid_index_row_words(h)
        self.n_ref = np.zeros(n_colours, dtype=np.uint64)

    def close(self):
        if getattr(self, "h", None):
            self.lib.cid_index_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    # ---- build (build.rs:33-130) ----
    def build_accession(self, colour, seqs, mode=L.CID_SEQ_FASTA, cutoff=-1):
        bases, offs = pack_seqs(list(seqs))
        n_ref, used = C.c_uint64(0), C.c_int64(0)
        L.check(self.lib.cid_build_accession(self.h, colour, _p(bases), _p(offs, L.u64p), len(seqs), mode, cutoff,
                                             C.byref(n_ref), C.byref(used)))
        self.n_ref[colour] = n_ref.value
        return n_ref.value, used.value

    def build_accession_mini(self, colour, seqs, mode=L.CID_SEQ_FASTA, cutoff=-1, variant=L.CID_MINI_OF_KMERS):
        """build.rs:396-492 build_single_mini (CID_MINI_OF_KMERS) / :258-394 build_multi_mini (CID_MINI_COUNTED)."""
        bases, offs = pack_seqs(list(seqs))
        n_ref, used = C.c_uint64(0), C.c_int64(0)
        L.check(self.lib.cid_build_accession_mini(self.h, colour, _p(bases), _p(offs, L.u64p), len(seqs), mode, cutoff,
                                                  variant, C.byref(n_ref), C.byref(used)))
        self.n_ref[colour] = n_ref.value
        return n_ref.value, used.value

    def build_accession_dev(self, colour, d_bases_ptr, d_offs_ptr, nseq, nbases, mode=L.CID_SEQ_FASTA, cutoff=-1):
        n_ref, used = C.c_uint64(0), C.c_int64(0)
        L.check(self.lib.cid_build_accession_dev(self.h, colour, d_bases_ptr, d_offs_ptr, nseq, nbases, mode, cutoff,
                                                 C.byref(n_ref), C.byref(used)))
        self.n_ref[colour] = n_ref.value
        return n_ref.value, used.value

    def finalize(self):
        L.check(self.lib.cid_build_finalize(self.h))

    # ---- index I/O (bigsi.rs:51-69) ----
    def upload_rows(self, row_ids, words):
        row_ids = np.ascontiguousarray(row_ids, dtype=np.uint64)
        words = np.ascontiguousarray(words, dtype=np.uint32)
        L.check(self.lib.cid_index_upload_rows(self.h, _p(row_ids, L.u64p), _p(words, L.u32p), len(row_ids)))

    def nonzero_rows(self):
        n = C.c_uint64(0)
        L.check(self.lib.cid_index_count_nonzero_rows(self.h, C.byref(n)))
        return n.value

    def download_nonzero_rows(self):
        n = self.nonzero_rows()
        ids = np.zeros(max(n, 1), dtype=np.uint64)
        words = np.zeros((max(n, 1), self.W), dtype=np.uint32)
        got = C.c_uint64(0)
        L.check(self.lib.cid_index_download_nonzero_rows(self.h, _p(ids, L.u64p), _p(words, L.u32p), n, C.byref(got)))
        return ids[:n], words[:n]

    def download_dense(self):
        words = np.zeros((self.S, self.W), dtype=np.uint32)
        L.check(self.lib.cid_index_download_dense(self.h, _p(words, L.u32p)))
        return words

    def device_ptrs(self):
        rows, bm, nw = L.vp(), L.vp(), C.c_uint64(0)
        L.check(self.lib.cid_index_device_ptrs(self.h, C.byref(rows), C.byref(bm), C.byref(nw)))
        return rows.value, bm.value, nw.value

    def set_rownz_global(self, flag=True):
        L.check(self.lib.cid_index_set_rownz_global(self.h, int(flag)))

    # ---- search (batch_search_pe.rs:9-179) ----
    def query_counts(self, queries, seq_mode=L.CID_SEQ_FASTA, gene_search=False, filt=-1, want_uniq=True):
        flat = [s for q in queries for s in q]
        bases, offs = pack_seqs(flat)
        qoffs = group_offsets(queries)
        nq = len(queries)
        counts = np.zeros((nq, self.N), dtype=np.uint32)
        num_kmers = np.zeros(nq, dtype=np.uint64)
        un = np.zeros((nq, self.N), dtype=np.uint64) if want_uniq else None
        us = np.zeros((nq, self.N), dtype=np.uint64) if want_uniq else None
        um = np.zeros((nq, self.N), dtype=np.uint64) if want_uniq else None
        used = np.zeros(max(nq, 1), dtype=np.int64)
        L.check(self.lib.cid_query_counts(self.h, _p(bases), _p(offs, L.u64p), len(flat), _p(qoffs, L.u64p), nq, seq_mode,
                                          int(gene_search), filt, _p(counts, L.u32p), _p(num_kmers, L.u64p),
                                          _p(un, L.u64p), _p(us, L.u64p), _p(um, L.u64p), _p(used, L.i64p)))
        return dict(counts=counts, num_kmers=num_kmers, uniq_n=un, uniq_sum=us, uniq_mode=um, cutoff=used[:nq])

    # ---- column-sharded default report (three calls around one exchange; see include/colorid_b200.h) ----
    def query_survivors(self, queries, seq_mode=L.CID_SEQ_FASTA, gene_search=False, filt=-1, packed=None):
        """-> (device pointer of the dense survivor list, surv[nq], cutoff[nq]); the list is 16 bytes per k-mer.
        packed = (bases u8[], seq_offs u64[nseq + 1], query_offs u64[nq + 1]) replaces `queries` (lists of bytes)."""
        if packed is not None:
            bases, offs, qoffs = packed
            nflat, nq = len(offs) - 1, len(qoffs) - 1
        else:
            flat = [s for q in queries for s in q]
            bases, offs = pack_seqs(flat)
            qoffs = group_offsets(queries)
            nflat, nq = len(flat), len(queries)
        surv = np.zeros(max(nq, 1), dtype=np.uint64)
        used = np.zeros(max(nq, 1), dtype=np.int64)
        ptr = L.vp()
        L.check(self.lib.cid_query_survivors(self.h, _p(bases), _p(offs, L.u64p), nflat, _p(qoffs, L.u64p), nq, seq_mode,
                                             int(gene_search), filt, C.byref(ptr), _p(surv, L.u64p), _p(used, L.i64p)))
        return ptr.value or 0, surv[:nq], used[:nq]

    def slots_counts_dev(self, d_slots, surv, d_counts, d_num_kmers, d_pc, d_col, stream=0):
        surv = np.ascontiguousarray(surv, dtype=np.uint64)
        L.check(self.lib.cid_query_slots_counts_dev(self.h, d_slots, _p(surv, L.u64p), len(surv), d_counts, d_num_kmers, d_pc, d_col,
                                                    stream))

    def slots_uniq_dev(self, d_slots, surv, d_pc_local, d_pc_sum, d_col, stream=0):
        surv = np.ascontiguousarray(surv, dtype=np.uint64)
        nq = len(surv)
        un, us, um = (np.zeros((nq, self.N), dtype=np.uint64) for _ in range(3))
        L.check(self.lib.cid_query_slots_uniq_dev(self.h, d_slots, _p(surv, L.u64p), nq, d_pc_local, d_pc_sum, d_col,
                                                  _p(un, L.u64p), _p(us, L.u64p), _p(um, L.u64p), stream))
        return un, us, um

    # ---- perfect search (perfect_search.rs:6-60) ----
    def query_perfect(self, queries):
        flat = [s for q in queries for s in q]
        bases, offs = pack_seqs(flat)
        qoffs = group_offsets(queries)
        nq = len(queries)
        and_rows = np.zeros((nq, self.W), dtype=np.uint32)
        status = np.zeros(max(nq, 1), dtype=np.uint8)
        n_kmers = np.zeros(max(nq, 1), dtype=np.uint64)
        L.check(self.lib.cid_query_perfect(self.h, _p(bases), _p(offs, L.u64p), len(flat), _p(qoffs, L.u64p), nq,
                                           _p(and_rows, L.u32p), _p(status, L.u8p), _p(n_kmers, L.u64p)))
        return dict(and_rows=and_rows, status=status[:nq], n_kmers=n_kmers[:nq])

    def query_perfect_mf(self, records):
        """perfect_search.rs:62-120 batch_search_mf (-s -m): one query per FASTA record (list of bytes)."""
        bases, offs = pack_seqs(list(records))
        nq = len(records)
        and_rows = np.zeros((nq, self.W), dtype=np.uint32)
        status = np.zeros(max(nq, 1), dtype=np.uint8)
        n_kmers = np.zeros(max(nq, 1), dtype=np.uint64)
        L.check(self.lib.cid_query_perfect_mf(self.h, _p(bases), _p(offs, L.u64p), nq, _p(and_rows, L.u32p),
                                              _p(status, L.u8p), _p(n_kmers, L.u64p)))
        return dict(and_rows=and_rows, status=status[:nq], n_kmers=n_kmers[:nq])

    # ---- read_id (read_id_mt_pe.rs:282-363) ----
    def _params(self, d, start_sample, qual_offset, group_width, reserve_before_find, rep_cap):
        return L.ReadIdParams(d, start_sample, qual_offset, group_width, int(reserve_before_find),
                              rep_cap if rep_cap else self.N + 1)

    def read_id_batch(self, reads, quals=None, d=1, start_sample=3, qual_offset=0, group_width=16,
                      reserve_before_find=True, rep_cap=None):
        flat = [s for r in reads for s in r]
        bases, offs = pack_seqs(flat)
        roffs = group_offsets(reads)
        qarr = None
        if quals is not None:
            qarr, _ = pack_seqs([s for r in quals for s in r])
        nr = len(reads)
        p = self._params(d, start_sample, qual_offset, group_width, reserve_before_find, rep_cap)
        cap = p.rep_cap
        n_set = np.zeros(max(nr, 1), np.uint32)
        flags = np.zeros(max(nr, 1), np.uint32)
        rep_n = np.zeros(max(nr, 1), np.uint32)
        rep_c = np.zeros((max(nr, 1), cap), np.uint32)
        rep_v = np.zeros((max(nr, 1), cap), np.uint32)
        L.check(self.lib.cid_read_id_batch(self.h, _p(bases), _p(qarr), _p(offs, L.u64p), len(flat), _p(roffs, L.u64p), nr,
                                           C.byref(p), _p(n_set, L.u32p), _p(flags, L.u32p), _p(rep_n, L.u32p),
                                           _p(rep_c, L.u32p), _p(rep_v, L.u32p)))
        return dict(n_set=n_set[:nr], flags=flags[:nr], rep_n=rep_n[:nr], rep_colour=rep_c[:nr], rep_count=rep_v[:nr])

    def read_id_classify(self, reads, quals=None, d=1, start_sample=3, qual_offset=0, group_width=16,
                         reserve_before_find=True, fp_correct=1e-3, top_cap=8):
        """parallel_vec end to end: reads in, (kind, hits, n_set, n_top, top) per read out."""
        flat = [s for r in reads for s in r]
        bases, offs = pack_seqs(flat)
        roffs = group_offsets(reads)
        qarr = None
        if quals is not None:
            qarr, _ = pack_seqs([s for r in quals for s in r])
        nr = len(reads)
        p = self._params(d, start_sample, qual_offset, group_width, reserve_before_find, None)
        m = max(nr, 1)
        kind = np.zeros(m, np.int32)
        hits, n_set, n_top = np.zeros(m, np.uint32), np.zeros(m, np.uint32), np.zeros(m, np.uint32)
        top = np.zeros((m, top_cap), np.uint32)
        n_ref = np.ascontiguousarray(self.n_ref, dtype=np.uint64)
        L.check(self.lib.cid_read_id_classify(self.h, _p(bases), _p(qarr), _p(offs, L.u64p), len(flat), _p(roffs, L.u64p),
                                              nr, C.byref(p), _p(n_ref, L.u64p), fp_correct, _p(kind, L.i32p),
                                              _p(hits, L.u32p), _p(n_set, L.u32p), _p(n_top, L.u32p), _p(top, L.u32p),
                                              top_cap))
        return dict(kind=kind[:nr], hits=hits[:nr], n_set=n_set[:nr], n_top=n_top[:nr], top=top[:nr])

    def read_id_classify_packed(self, packed, d=1, start_sample=3, group_width=16, reserve_before_find=True, fp_correct=1e-3, top_cap=8):
        """cid_read_id_classify_packed on the output of pack_reads()."""
        nr = len(packed["read_offs"]) - 1
        p = self._params(d, start_sample, 0, group_width, reserve_before_find, None)
        m = max(nr, 1)
        kind = np.zeros(m, np.int32)
        hits, n_set, n_top = np.zeros(m, np.uint32), np.zeros(m, np.uint32), np.zeros(m, np.uint32)
        top = np.zeros((m, top_cap), np.uint32)
        n_ref = np.ascontiguousarray(self.n_ref, dtype=np.uint64)
        L.check(self.lib.cid_read_id_classify_packed(self.h, _p(packed["words"], L.u32p), _p(packed["word_offs"], L.u64p), packed["flags"],
                                                     _p(packed["seq_offs"], L.u64p), packed["nseq"], _p(packed["read_offs"], L.u64p), nr,
                                                     C.byref(p), _p(n_ref, L.u64p), fp_correct, _p(kind, L.i32p), _p(hits, L.u32p),
                                                     _p(n_set, L.u32p), _p(n_top, L.u32p), _p(top, L.u32p), top_cap))
        return dict(kind=kind[:nr], hits=hits[:nr], n_set=n_set[:nr], n_top=n_top[:nr], top=top[:nr])

    def read_kmer_order(self, reads, d=1, group_width=16, reserve_before_find=True, order_cap=512):
        flat = [s for r in reads for s in r]
        bases, offs = pack_seqs(flat)
        roffs = group_offsets(reads)
        nr = len(reads)
        p = self._params(d, 3, 0, group_width, reserve_before_find, None)
        on = np.zeros(max(nr, 1), np.uint32)
        osq = np.zeros((max(nr, 1), order_cap), np.uint8)
        opos = np.zeros((max(nr, 1), order_cap), np.uint32)
        L.check(self.lib.cid_read_kmer_order32(self.h, _p(bases), _p(offs, L.u64p), len(flat), _p(roffs, L.u64p), nr,
                                               C.byref(p), order_cap, _p(on, L.u32p), _p(osq, L.u8p), _p(opos, L.u32p)))
        return on[:nr], osq[:nr], opos[:nr]

    def hash_kmers(self, kmers):
        """kmers: list of k-byte ASCII strings -> row ids [n, H]."""
        arr = np.frombuffer(b"".join(kmers), dtype=np.uint8).copy()
        out = np.zeros((len(kmers), self.H), dtype=np.uint64)
        L.check(self.lib.cid_hash_kmers(self.h, _p(arr), len(kmers), _p(out, L.u64p)))
        return out


def pack_reads(reads, quals=None, qual_offset=0, threads=0):
    """cid_pack_reads: lists of mates (bytes) -> dict(words u32[], word_offs u64[n+1], flags, seq_offs, read_offs, nseq)."""
    lib = L.load()
    flat = [s for r in reads for s in r]
    bases, offs = pack_seqs(flat)
    roffs = group_offsets(reads)
    qarr = None
    if quals is not None:
        qarr, _ = pack_seqs([s for r in quals for s in r])
    nr = len(reads)
    cap = int(lib.cid_pack_words_bound(_p(offs, L.u64p), _p(roffs, L.u64p), nr, 1))
    words = np.zeros(max(cap, 1), dtype=np.uint32)
    woffs = np.zeros(nr + 1, dtype=np.uint64)
    flags = C.c_uint32(0)
    L.check(lib.cid_pack_reads(_p(bases), _p(qarr), _p(offs, L.u64p), len(flat), _p(roffs, L.u64p), nr, qual_offset, threads,
                               _p(words, L.u32p), cap, _p(woffs, L.u64p), C.byref(flags)))
    return dict(words=words[:max(int(woffs[nr]), 1)], word_offs=woffs, flags=flags.value, seq_offs=offs, read_offs=roffs, nseq=len(flat))


class MultiIndex:
    """One index over several GPUs of this node (include/colorid_b200.h `cid_mg`): replicated (reads / queries dealt to the
    GPUs) or column-sharded (every GPU gathers all k-mers from its accession slice).  Same results as `Index`.
    devices: list of CUDA device ids; an id may repeat (several shards on one GPU)."""

    def __init__(self, devices, mode, bloom_size, num_hash, k, n_colours, m=0, hash_variant=0):
        self.lib = L.load()
        self.S, self.H, self.k, self.N, self.m = bloom_size, num_hash, k, n_colours, m
        self.W = (n_colours + 31) // 32
        devs = (C.c_int * len(devices))(*devices)
        h = L.vp()
        L.check(self.lib.cid_mg_create(devs, len(devices), mode, C.byref(h)))
        self.h = h
        L.check(self.lib.cid_mg_index_create(h, bloom_size, num_hash, k, n_colours))
        if m:
            L.check(self.lib.cid_mg_index_set_minimizer(h, m))
        if hash_variant:
            L.check(self.lib.cid_mg_index_set_hash_variant(h, hash_variant))
        self.n_ref = np.zeros(n_colours, dtype=np.uint64)

    def close(self):
        if getattr(self, "h", None):
            self.lib.cid_mg_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @property
    def launches(self):
        return int(self.lib.cid_mg_launch_count(self.h))

    def set_option(self, name, value):
        L.check(self.lib.cid_mg_set_option(self.h, name.encode(), int(value)))

    def shard_columns(self):
        out = []
        for g in range(self.lib.cid_mg_n_shards(self.h)):
            a, b = C.c_uint32(0), C.c_uint32(0)
            L.check(self.lib.cid_mg_shard_columns(self.h, g, C.byref(a), C.byref(b)))
            out.append((a.value, b.value))
        return out

    def build_accession(self, colour, seqs, mode=L.CID_SEQ_FASTA, cutoff=-1, mini_variant=-1):
        bases, offs = pack_seqs(list(seqs))
        n_ref, used = C.c_uint64(0), C.c_int64(0)
        L.check(self.lib.cid_mg_build_accession(self.h, colour, _p(bases), _p(offs, L.u64p), len(seqs), mode, cutoff, mini_variant,
                                                C.byref(n_ref), C.byref(used)))
        self.n_ref[colour] = n_ref.value
        return n_ref.value, used.value

    def finalize(self):
        L.check(self.lib.cid_mg_build_finalize(self.h))

    def upload_rows(self, row_ids, words):
        row_ids = np.ascontiguousarray(row_ids, dtype=np.uint64)
        words = np.ascontiguousarray(words, dtype=np.uint32)
        L.check(self.lib.cid_mg_index_upload_rows(self.h, _p(row_ids, L.u64p), _p(words, L.u32p), len(row_ids)))

    def nonzero_rows(self):
        n = C.c_uint64(0)
        L.check(self.lib.cid_mg_index_count_nonzero_rows(self.h, C.byref(n)))
        return n.value

    def download_nonzero_rows(self):
        n = self.nonzero_rows()
        ids = np.zeros(max(n, 1), dtype=np.uint64)
        words = np.zeros((max(n, 1), self.W), dtype=np.uint32)
        got = C.c_uint64(0)
        L.check(self.lib.cid_mg_index_download_nonzero_rows(self.h, _p(ids, L.u64p), _p(words, L.u32p), n, C.byref(got)))
        return ids[:n], words[:n]

    def query_counts(self, queries, seq_mode=L.CID_SEQ_FASTA, gene_search=False, filt=-1, want_uniq=True):
        flat = [s for q in queries for s in q]
        bases, offs = pack_seqs(flat)
        qoffs = group_offsets(queries)
        nq = len(queries)
        counts = np.zeros((nq, self.N), dtype=np.uint32)
        num_kmers = np.zeros(nq, dtype=np.uint64)
        un, us, um = ((np.zeros((nq, self.N), dtype=np.uint64) if want_uniq else None) for _ in range(3))
        used = np.zeros(max(nq, 1), dtype=np.int64)
        L.check(self.lib.cid_mg_query_counts(self.h, _p(bases), _p(offs, L.u64p), len(flat), _p(qoffs, L.u64p), nq, seq_mode,
                                             int(gene_search), filt, _p(counts, L.u32p), _p(num_kmers, L.u64p),
                                             _p(un, L.u64p), _p(us, L.u64p), _p(um, L.u64p), _p(used, L.i64p)))
        return dict(counts=counts, num_kmers=num_kmers, uniq_n=un, uniq_sum=us, uniq_mode=um, cutoff=used[:nq])

    def query_perfect(self, queries):
        flat = [s for q in queries for s in q]
        bases, offs = pack_seqs(flat)
        qoffs = group_offsets(queries)
        nq = len(queries)
        and_rows = np.zeros((nq, self.W), dtype=np.uint32)
        status = np.zeros(max(nq, 1), dtype=np.uint8)
        n_kmers = np.zeros(max(nq, 1), dtype=np.uint64)
        L.check(self.lib.cid_mg_query_perfect(self.h, _p(bases), _p(offs, L.u64p), len(flat), _p(qoffs, L.u64p), nq,
                                              _p(and_rows, L.u32p), _p(status, L.u8p), _p(n_kmers, L.u64p)))
        return dict(and_rows=and_rows, status=status[:nq], n_kmers=n_kmers[:nq])

    def query_perfect_mf(self, records):
        bases, offs = pack_seqs(list(records))
        nq = len(records)
        and_rows = np.zeros((nq, self.W), dtype=np.uint32)
        status = np.zeros(max(nq, 1), dtype=np.uint8)
        n_kmers = np.zeros(max(nq, 1), dtype=np.uint64)
        L.check(self.lib.cid_mg_query_perfect_mf(self.h, _p(bases), _p(offs, L.u64p), nq, _p(and_rows, L.u32p),
                                                 _p(status, L.u8p), _p(n_kmers, L.u64p)))
        return dict(and_rows=and_rows, status=status[:nq], n_kmers=n_kmers[:nq])

    def read_id_classify(self, reads, quals=None, d=1, start_sample=3, qual_offset=0, group_width=16, reserve_before_find=True,
                         fp_correct=1e-3, top_cap=8):
        flat = [s for r in reads for s in r]
        bases, offs = pack_seqs(flat)
        roffs = group_offsets(reads)
        qarr = None
        if quals is not None:
            qarr, _ = pack_seqs([s for r in quals for s in r])
        nr = len(reads)
        p = L.ReadIdParams(d, start_sample, qual_offset, group_width, int(reserve_before_find), 0)
        mm = max(nr, 1)
        kind = np.zeros(mm, np.int32)
        hits, n_set, n_top = np.zeros(mm, np.uint32), np.zeros(mm, np.uint32), np.zeros(mm, np.uint32)
        top = np.zeros((mm, top_cap), np.uint32)
        n_ref = np.ascontiguousarray(self.n_ref, dtype=np.uint64)
        L.check(self.lib.cid_mg_read_id_classify(self.h, _p(bases), _p(qarr), _p(offs, L.u64p), len(flat), _p(roffs, L.u64p), nr,
                                                 C.byref(p), _p(n_ref, L.u64p), fp_correct, _p(kind, L.i32p), _p(hits, L.u32p),
                                                 _p(n_set, L.u32p), _p(n_top, L.u32p), _p(top, L.u32p), top_cap))
        return dict(kind=kind[:nr], hits=hits[:nr], n_set=n_set[:nr], n_top=n_top[:nr], top=top[:nr])
