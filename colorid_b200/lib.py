"""ctypes loader for libcolorid_b200.so (the CUDA library behind the C ABI in include/colorid_b200.h).

There is no CPU fallback: if the shared library is missing this raises, and every entry point
fails with CID_E_CUDA when no CUDA device is usable.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcolorid_b200.so")

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)
vp = C.c_void_p

CID_OK, CID_E_INVALID, CID_E_CUDA, CID_E_NOMEM, CID_E_UNSUPPORTED, CID_E_REF_PANIC, CID_E_CAPACITY = 0, -1, -2, -3, -4, -5, -6
CID_SEQ_FASTA, CID_SEQ_FASTQ, CID_SEQ_STRING = 0, 1, 2
CID_MG_REPLICATED, CID_MG_COLUMNS = 0, 1
CID_PACK_LOWER = 1
CID_MINI_OF_KMERS, CID_MINI_COUNTED = 0, 1


class ReadIdParams(C.Structure):
    _fields_ = [("downsample", C.c_uint32), ("start_sample", C.c_uint32), ("qual_offset", C.c_uint32),
                ("group_width", C.c_uint32), ("reserve_before_find", C.c_uint32), ("rep_cap", C.c_uint32)]


# name -> (restype, argtypes); must list every symbol declared in include/colorid_b200.h
SIGNATURES = {
    "cid_version": (C.c_int, []),
    "cid_last_error": (C.c_char_p, []),
    "cid_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
    "cid_host_free": (None, [vp]),
    "cid_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "cid_ctx_destroy": (None, [vp]),
    "cid_ctx_device": (C.c_int, [vp]),
    "cid_ctx_launch_count": (C.c_uint64, [vp]),
    "cid_index_create": (C.c_int, [vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp)]),
    "cid_index_destroy": (None, [vp]),
    "cid_index_row_words": (C.c_uint32, [vp]),
    "cid_index_row_stride": (C.c_uint32, [vp]),
    "cid_index_upload_rows": (C.c_int, [vp, u64p, u32p, C.c_uint64]),
    "cid_index_count_nonzero_rows": (C.c_int, [vp, u64p]),
    "cid_index_download_nonzero_rows": (C.c_int, [vp, u64p, u32p, C.c_uint64, u64p]),
    "cid_index_download_dense": (C.c_int, [vp, u32p]),
    "cid_index_device_ptrs": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), u64p]),
    "cid_index_refresh_rownz": (C.c_int, [vp]),
    "cid_index_set_rownz_global": (C.c_int, [vp, C.c_int]),
    "cid_build_accession": (C.c_int, [vp, C.c_uint32, vp, u64p, C.c_uint64, C.c_int, C.c_int64, u64p, i64p]),
    "cid_build_accession_dev": (C.c_int, [vp, C.c_uint32, vp, vp, C.c_uint64, C.c_uint64, C.c_int, C.c_int64, u64p, i64p]),
    "cid_index_set_minimizer": (C.c_int, [vp, C.c_uint32]),
    "cid_index_set_hash_variant": (C.c_int, [vp, C.c_uint32]),
    "cid_pack_words_bound": (C.c_uint64, [u64p, u64p, C.c_uint64, C.c_int]),
    "cid_pack_reads": (C.c_int, [vp, vp, u64p, C.c_uint64, u64p, C.c_uint64, C.c_uint32, C.c_int, u32p, C.c_uint64, u64p, u32p]),
    "cid_read_id_classify_packed": (C.c_int, [vp, vp, u64p, C.c_uint32, u64p, C.c_uint64, u64p, C.c_uint64, C.POINTER(ReadIdParams), u64p,
                                              C.c_double, C.POINTER(C.c_int32), u32p, u32p, u32p, u32p, C.c_uint32]),
    "cid_read_id_batch_packed_dev": (C.c_int, [vp, vp, vp, C.c_uint32, vp, C.c_uint64, vp, C.c_uint64, C.c_uint32, C.c_uint32,
                                               C.POINTER(ReadIdParams), vp, vp, vp, vp, vp, vp]),
    "cid_mg_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(vp)]),
    "cid_mg_destroy": (None, [vp]),
    "cid_mg_n_shards": (C.c_int, [vp]),
    "cid_mg_mode": (C.c_int, [vp]),
    "cid_mg_ctx": (vp, [vp, C.c_int]),
    "cid_mg_index": (vp, [vp, C.c_int]),
    "cid_mg_shard_columns": (C.c_int, [vp, C.c_int, u32p, u32p]),
    "cid_mg_shard_of_colour": (C.c_int, [vp, C.c_uint32]),
    "cid_mg_set_option": (C.c_int, [vp, C.c_char_p, C.c_int64]),
    "cid_mg_launch_count": (C.c_uint64, [vp]),
    "cid_mg_index_create": (C.c_int, [vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]),
    "cid_mg_index_set_minimizer": (C.c_int, [vp, C.c_uint32]),
    "cid_mg_index_set_hash_variant": (C.c_int, [vp, C.c_uint32]),
    "cid_mg_index_upload_rows": (C.c_int, [vp, u64p, u32p, C.c_uint64]),
    "cid_mg_index_count_nonzero_rows": (C.c_int, [vp, u64p]),
    "cid_mg_index_download_nonzero_rows": (C.c_int, [vp, u64p, u32p, C.c_uint64, u64p]),
    "cid_mg_build_accession": (C.c_int, [vp, C.c_uint32, vp, u64p, C.c_uint64, C.c_int, C.c_int64, C.c_int, u64p, C.POINTER(C.c_int64)]),
    "cid_mg_build_finalize": (C.c_int, [vp]),
    "cid_mg_query_counts": (C.c_int, [vp, vp, u64p, C.c_uint64, u64p, C.c_uint64, C.c_int, C.c_int, C.c_int64, u32p, u64p, u64p, u64p, u64p,
                                      C.POINTER(C.c_int64)]),
    "cid_mg_query_perfect": (C.c_int, [vp, vp, u64p, C.c_uint64, u64p, C.c_uint64, u32p, u8p, u64p]),
    "cid_mg_query_perfect_mf": (C.c_int, [vp, vp, u64p, C.c_uint64, u32p, u8p, u64p]),
    "cid_mg_read_id_classify": (C.c_int, [vp, vp, vp, u64p, C.c_uint64, u64p, C.c_uint64, C.POINTER(ReadIdParams), u64p,
                                          C.c_double, C.POINTER(C.c_int32), u32p, u32p, u32p, u32p, C.c_uint32]),
    "cid_index_hash_variant": (C.c_uint32, [vp]),
    "cid_index_minimizer": (C.c_uint32, [vp]),
    "cid_build_accession_mini": (C.c_int, [vp, C.c_uint32, vp, u64p, C.c_uint64, C.c_int, C.c_int64, C.c_int, u64p, i64p]),
    "cid_build_finalize": (C.c_int, [vp]),
    "cid_query_counts": (C.c_int, [vp, vp, u64p, C.c_uint64, u64p, C.c_uint64, C.c_int, C.c_int, C.c_int64, u32p, u64p,
                                   u64p, u64p, u64p, i64p]),
    "cid_query_counts_dev": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_uint64, vp, u64p, u64p, C.c_uint64, C.c_int, vp, vp, vp]),
    "cid_dev_alloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    "cid_dev_free": (None, [vp, vp]),
    "cid_ipc_export": (C.c_int, [vp, vp, u8p]),
    "cid_ipc_open": (C.c_int, [vp, u8p, C.POINTER(vp)]),
    "cid_ipc_close": (C.c_int, [vp, vp]),
    "cid_query_counts_sharded_dev": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_uint64, vp, u64p, u64p, C.c_uint64, C.c_int,
                                               C.POINTER(vp), C.c_uint32, C.c_uint32, C.c_uint32, vp, vp]),
    "cid_query_survivors": (C.c_int, [vp, vp, u64p, C.c_uint64, u64p, C.c_uint64, C.c_int, C.c_int, C.c_int64, C.POINTER(vp), u64p,
                                      i64p]),
    "cid_query_slots_counts_dev": (C.c_int, [vp, vp, u64p, C.c_uint64, vp, vp, vp, vp, vp]),
    "cid_query_slots_uniq_dev": (C.c_int, [vp, vp, u64p, C.c_uint64, vp, vp, vp, u64p, u64p, u64p, vp]),
    "cid_query_perfect": (C.c_int, [vp, vp, u64p, C.c_uint64, u64p, C.c_uint64, u32p, u8p, u64p]),
    "cid_query_perfect_mf": (C.c_int, [vp, vp, u64p, C.c_uint64, u32p, u8p, u64p]),
    "cid_read_id_batch": (C.c_int, [vp, vp, vp, u64p, C.c_uint64, u64p, C.c_uint64, C.POINTER(ReadIdParams), u32p, u32p,
                                    u32p, u32p, u32p]),
    "cid_read_id_batch_dev": (C.c_int, [vp, vp, vp, vp, C.c_uint64, C.c_uint64, vp, C.c_uint64, C.c_uint32, C.c_uint32,
                                        C.POINTER(ReadIdParams), vp, vp, vp, vp, vp, vp]),
    "cid_read_kmer_order": (C.c_int, [vp, vp, u64p, C.c_uint64, u64p, C.c_uint64, C.POINTER(ReadIdParams), C.c_uint32,
                                      u32p, u8p, u16p]),
    "cid_read_kmer_order32": (C.c_int, [vp, vp, u64p, C.c_uint64, u64p, C.c_uint64, C.POINTER(ReadIdParams), C.c_uint32,
                                        u32p, u8p, u32p]),
    "cid_read_id_classify": (C.c_int, [vp, vp, vp, u64p, C.c_uint64, u64p, C.c_uint64, C.POINTER(ReadIdParams), u64p,
                                       C.c_double, C.POINTER(C.c_int32), u32p, u32p, u32p, u32p, C.c_uint32]),
    "cid_ctx_set_option": (C.c_int, [vp, C.c_char_p, C.c_int64]),
    "cid_ctx_read_counter": (C.c_int, [vp, C.c_char_p, u64p]),
    "cid_hash_kmers": (C.c_int, [vp, vp, C.c_uint64, u64p]),
    "cid_ctx_profile": (C.c_int, [vp, C.c_int]),
    "cid_ctx_profile_read": (C.c_int, [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), u64p]),
    "cid_classify_reads": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, u64p, C.c_double, C.c_uint32, C.c_uint64, u32p,
                                     u32p, u32p, u32p, u32p, C.c_uint32, C.c_int, C.POINTER(C.c_int32), u32p, u32p, u32p,
                                     C.c_uint32]),
    "cid_merge_shard_reports": (C.c_int, [C.c_uint32, u32p, u32p, C.c_uint64, C.POINTER(u32p), C.POINTER(u32p), C.POINTER(u32p),
                                          C.c_uint32, C.c_uint32, u32p, u32p, u32p, C.c_uint32, u32p]),
    "cid_false_prob": (C.c_double, [C.c_double, C.c_double, C.c_double]),
    "cid_binomial_mass": (C.c_double, [C.c_uint64, C.c_double, C.c_uint64]),
}

_LIB = None


def load():
    """Load the CUDA library; raises if it has not been built (run __graft_entry__.build())."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(colorid_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


class CidError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"colorid_b200 error {code}: {msg}")
        self.code = code


def check(rc):
    if rc != CID_OK:
        raise CidError(rc, load().cid_last_error().decode("utf-8", "replace"))
