"""colorid_b200 — B200-native BIGSI hot path (build / search / read_id) of hcdenbakker/colorid.

The product is the CUDA shared library `libcolorid_b200.so` behind the C ABI declared in
include/colorid_b200.h; this package only loads it (ctypes) and offers numpy-facing wrappers
for tests and benchmarks.  There is no CPU fallback.
"""
from . import lib  # noqa: F401
from .api import Context, Index, MultiIndex, pack_seqs, group_offsets, pack_reads  # noqa: F401
from .lib import CID_SEQ_FASTA, CID_SEQ_FASTQ, CID_SEQ_STRING, CID_MINI_OF_KMERS, CID_MINI_COUNTED, CID_MG_REPLICATED, CID_MG_COLUMNS, CidError  # noqa: F401
