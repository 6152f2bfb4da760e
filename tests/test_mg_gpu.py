"""cid_mg (include/colorid_b200.h "multi-GPU"): one index over several GPUs, replicated or column-sharded, through the C ABI,
against the oracle on the whole index and against the single-GPU entry points.  Devices are taken round-robin from the
visible GPUs, so on a one-GPU box the shards share the device (same code path: threads, exchanges, merges); with
`gpurun --gpus N` they spread over N GPUs and the copies are peer copies."""
import numpy as np
import pytest

import colorid_b200 as cb
from tests import synth
from tests.test_gpu_parity import _readid_compare, _rng  # noqa: F401

pytestmark = pytest.mark.gpu


def _devices(n):
    import ctypes
    nd = ctypes.c_int(0)
    ctypes.CDLL("libcudart.so").cudaGetDeviceCount(ctypes.byref(nd))
    return [g % max(nd.value, 1) for g in range(n)]


def _build(oracle, shards, mode, N, k, S, H, rng, glen=5000, fastq=False):
    genomes = synth.clade_genomes(rng, N, glen, n_clades=max(2, N // 10), div=0.01)
    oix = oracle.Index(S, H, k, N)
    mix = cb.MultiIndex(_devices(shards), mode, S, H, k, N)
    for c, g in enumerate(genomes):
        if fastq and c % 7 == 3:
            reads = synth.reads_from(rng, [g], 500, read_len=100, insert=200, err=0.006, frac_random=0.0)
            acc, m, om = [x for r in reads for x in r], cb.CID_SEQ_FASTQ, oracle.MODE_FASTQ
        else:
            acc, m, om = [g[:glen // 2], g[glen // 2:]], cb.CID_SEQ_FASTA, oracle.MODE_FASTA
        assert mix.build_accession(c, acc, m) == oix.build_accession(c, acc, om), c
    oix.finalize(threads=4)
    mix.finalize()
    return genomes, oix, mix


@pytest.mark.parametrize("mode", [cb.CID_MG_COLUMNS, cb.CID_MG_REPLICATED])
@pytest.mark.parametrize("shards,N,k,S,H", [(2, 150, 21, 300_007, 2), (3, 200, 31, 200_003, 4), (4, 70, 27, 100_003, 3)])
def test_mg_build_search_read_id_equal_the_oracle(oracle, mode, shards, N, k, S, H):
    rng = _rng(7000 + N + shards)
    genomes, oix, mix = _build(oracle, shards, mode, N, k, S, H, rng, fastq=True)
    if mode == cb.CID_MG_COLUMNS:
        cols = mix.shard_columns()
        assert sum(n for _, n in cols) == N and all(a % 32 == 0 for a, _ in cols)
    # .bxi rows: ascending, full width, identical to the oracle's matrix
    ids, words = mix.download_nonzero_rows()
    dense = oix.words()
    nz = np.flatnonzero(dense.any(axis=1))
    assert np.array_equal(ids, nz.astype(np.uint64)) and np.array_equal(words, dense[nz])
    # search: -g (streaming gather), default report with unique-hit summaries (FASTA and a FASTQ read set), -s, -s -m
    queries = [[genomes[int(rng.integers(0, N))][200:200 + int(rng.integers(300, 2500))]] for _ in range(23)]
    queries += [[synth.rand_seq(rng, 900)], [b"ACGT"], [genomes[1][:800], genomes[2][100:700]]]
    for gene, filt, uq in ((True, -1, False), (False, 0, True), (False, 1, True)):
        o = oix.query_counts(queries, oracle.MODE_FASTA, gene, filt)
        g = mix.query_counts(queries, cb.CID_SEQ_FASTA, gene, filt, want_uniq=uq)
        assert np.array_equal(g["num_kmers"], o["num_kmers"]) and np.array_equal(g["counts"], o["counts"]), (gene, filt)
        if uq:
            for key in ("uniq_n", "uniq_sum", "uniq_mode"):
                assert np.array_equal(g[key], o[key]), key
    rs = synth.reads_from(rng, [genomes[5]], 900, read_len=100, insert=220, err=0.005, frac_random=0.0)
    fq = [[m for r in rs for m in r], [m for r in rs[:200] for m in r]]
    o = oix.query_counts(fq, oracle.MODE_FASTQ, False, -1)
    g = mix.query_counts(fq, cb.CID_SEQ_FASTQ, False, -1)
    for key in ("counts", "num_kmers", "cutoff", "uniq_n", "uniq_sum", "uniq_mode"):
        assert np.array_equal(g[key], o[key]), key
    op, gp = oix.query_perfect(queries), mix.query_perfect(queries)
    assert np.array_equal(gp["status"], op["status"]) and np.array_equal(gp["and_rows"], op["and_rows"]) and np.array_equal(gp["n_kmers"], op["n_kmers"])
    recs = [q[0] for q in queries] + [genomes[3][:50] + b"N" + genomes[3][51:400]]
    om, gm = oix.query_perfect(recs, mf=True), mix.query_perfect_mf(recs)
    assert np.array_equal(gm["status"], om["status"]) and np.array_equal(gm["and_rows"], om["and_rows"])
    # read_id: classifications of parallel_vec, short and long and lower-case reads, ties included (clade members share k-mers)
    mix.n_ref[:] = oix.n_ref
    reads = synth.reads_from(rng, genomes, 500, read_len=150, insert=320, err=0.003, frac_random=0.2, n_rate=0.002)
    reads += synth.reads_from(rng, genomes, 6, read_len=1800, insert=1900, err=0.001, frac_random=0.0, paired=False)
    reads += [[genomes[2][:150].lower()], [b"ACGT", genomes[0][:150]], [b"N" * 150]]
    for kw in (dict(), dict(start_sample=0), dict(d=2, start_sample=1)):
        o = oix.read_id_batch(reads, top_cap=16, **kw)
        g = mix.read_id_classify(reads, top_cap=16, **kw)
        for key in ("kind", "hits", "n_set", "n_top"):
            assert np.array_equal(g[key], o[key]), (kw, key)
        for r in range(len(reads)):
            nt = min(int(o["n_top"][r]), 16)
            assert g["top"][r, :nt].tolist() == o["top"][r, :nt].tolist(), (kw, r)
    assert (o["kind"] == oracle.CLS_REJECT_MULTI).sum() > 5 and (o["kind"] == oracle.CLS_ACCEPT).sum() > 50
    mix.close()


def test_mg_upload_rows_and_more_shards_than_word_columns(oracle):
    """A .bxi loaded into a column-sharded / replicated index (cid_mg_index_upload_rows), and an index of 40 accessions
    (2 word columns) on 4 shards: two shards stay empty."""
    rng = _rng(7100)
    N, k, S, H = 40, 21, 100_003, 2
    genomes, oix, mix = _build(oracle, 4, cb.CID_MG_COLUMNS, N, k, S, H, rng)
    assert [n for _, n in mix.shard_columns()] == [32, 8, 0, 0]
    ids, words = mix.download_nonzero_rows()
    queries = [[g[100:1500]] for g in genomes[::5]]
    o = oix.query_counts(queries, oracle.MODE_FASTA, True, 0)
    for mode, shards in ((cb.CID_MG_COLUMNS, 2), (cb.CID_MG_REPLICATED, 3), (cb.CID_MG_COLUMNS, 1)):
        m2 = cb.MultiIndex(_devices(shards), mode, S, H, k, N)
        m2.upload_rows(ids, words)
        g = m2.query_counts(queries, cb.CID_SEQ_FASTA, True, 0, want_uniq=False)
        assert np.array_equal(g["counts"], o["counts"])
        i2, w2 = m2.download_nonzero_rows()
        assert np.array_equal(i2, ids) and np.array_equal(w2, words)
        m2.n_ref[:] = oix.n_ref
        reads = synth.reads_from(rng, genomes, 100, read_len=120, insert=300, err=0.003, frac_random=0.2)
        oo, gg = oix.read_id_batch(reads), m2.read_id_classify(reads)
        for key in ("kind", "hits", "n_set", "n_top"):
            assert np.array_equal(gg[key], oo[key]), key
        m2.close()
    mix.close()
