"""GPU parity tests for minimizer indexes (.mxi): build_single_mini / build_multi_mini (build.rs:258-492) and
read_id over minimizer sets (kmer.rs:363-394), through the C ABI against the CPU oracle, bit-exact."""
import numpy as np
import pytest

import colorid_b200 as cb
from tests import synth
from tests.test_gpu_parity import _readid_compare

pytestmark = pytest.mark.gpu


def _rng(seed):
    return np.random.default_rng(0xC0101D00 + 7000 + seed)


def _messy_fasta_accessions(rng, N, glen=6000):
    genomes = synth.clade_genomes(rng, N, glen, n_clades=3, div=0.02)
    accs = []
    for i, g in enumerate(genomes):
        contigs = [g[:2500], g[2500:2510], g[2510:], g[100:900]]      # a contig shorter than k is skipped; a repeat
        if i % 3 == 0:
            contigs[0] = synth.sprinkle(rng, contigs[0], b"NRYn", 0.01)   # has_no_n
        if i % 4 == 1:
            contigs[2] = contigs[2].lower()                           # all lower case: same order as upper case
        if i % 4 == 2:
            contigs[2] = synth.sprinkle(rng, contigs[2], b"acgt", 0.3)    # mixed case: raw-byte minimizer choice
        accs.append(contigs)
    return genomes, accs


def _build_both_mini(oracle, ctx, accs, S, H, k, m, mode, cutoff, variant):
    N = len(accs)
    oix = oracle.Index(S, H, k, N, m=m)
    gix = cb.Index(ctx, S, H, k, N, m=m)
    for c, acc in enumerate(accs):
        o = oix.build_accession_mini(c, acc, mode, cutoff, variant)
        g = gix.build_accession_mini(c, acc, mode, cutoff, variant)
        assert g == o, f"accession {c}: (n_ref, cutoff) {g} vs {o}"
    oix.finalize(threads=4)
    gix.finalize()
    return oix, gix


@pytest.mark.parametrize("k,m,S,H,N", [(31, 15, 300_007, 4, 9), (27, 17, 200_003, 2, 40), (21, 8, 65_536, 3, 5),
                                       (21, 21, 100_003, 2, 4), (16, 3, 50_021, 2, 3), (31, 9, 150_001, 4, 70)])
@pytest.mark.parametrize("variant", [cb.CID_MINI_OF_KMERS, cb.CID_MINI_COUNTED])
def test_build_fasta_minimizer_index_bit_exact(oracle, ctx, k, m, S, H, N, variant):
    rng = _rng(k * 100 + m)
    _, accs = _messy_fasta_accessions(rng, N)
    for cutoff in (-1, 1):
        oix, gix = _build_both_mini(oracle, ctx, accs, S, H, k, m, cb.CID_SEQ_FASTA, cutoff, variant)
        assert np.array_equal(gix.download_dense(), oix.words())
        assert gix.nonzero_rows() == oix.nonzero_rows() > 0


@pytest.mark.parametrize("variant", [cb.CID_MINI_OF_KMERS, cb.CID_MINI_COUNTED])
def test_build_fastq_minimizer_index_with_cutoffs(oracle, ctx, variant):
    rng = _rng(11)
    k, m, S, H = 21, 11, 400_009, 2
    genomes = synth.clade_genomes(rng, 3, 5000, n_clades=3, div=0.05)
    accs = []
    for g in genomes:
        reads = synth.reads_from(rng, [g], 1200, read_len=100, insert=200, err=0.01, frac_random=0.0, n_rate=0.002)
        accs.append([mate for r in reads for mate in r])
    for cutoff in (-1, 0, 2):
        oix, gix = _build_both_mini(oracle, ctx, accs, S, H, k, m, cb.CID_SEQ_FASTQ, cutoff, variant)
        assert np.array_equal(gix.download_dense(), oix.words())


def test_minimizer_hash_rows_match_oracle(oracle, ctx):
    rng = _rng(12)
    for m in (3, 8, 9, 15, 16, 17, 31):
        gix = cb.Index(ctx, 999_983, 4, 31, 1, m=m)
        items = [synth.rand_seq(rng, m) for _ in range(500)]
        exp = np.array([[oracle.xxh3_64(it, s) % 999_983 for s in range(4)] for it in items], dtype=np.uint64)
        assert np.array_equal(gix.hash_kmers(items), exp)


def _mini_index_pair(oracle, ctx, rng, N, k, m, S, H, variant=cb.CID_MINI_COUNTED, glen=6000):
    genomes = synth.clade_genomes(rng, N, glen, n_clades=3, div=0.02)
    oix, gix = _build_both_mini(oracle, ctx, [[g] for g in genomes], S, H, k, m, cb.CID_SEQ_FASTA, -1, variant)
    return genomes, oix, gix


@pytest.mark.parametrize("N,k,m,S,H", [(4, 27, 15, 750_000, 4), (46, 31, 15, 1_000_003, 4), (40, 21, 9, 100_003, 2),
                                       (150, 21, 12, 300_007, 2)])
def test_read_id_over_minimizer_sets(oracle, ctx, N, k, m, S, H):
    rng = _rng(500 + N)
    genomes, oix, gix = _mini_index_pair(oracle, ctx, rng, N, k, m, S, H, glen=4000 if N > 100 else 6000)
    reads = synth.reads_from(rng, genomes, 250, read_len=150, insert=320, err=0.004, frac_random=0.25, n_rate=0.002)
    reads += synth.reads_from(rng, genomes, 40, read_len=150, insert=200, err=0.0, frac_random=0.0)   # overlapping mates
    reads += synth.reads_from(rng, genomes, 30, read_len=100, insert=300, err=0.0, frac_random=0.0, paired=False)
    reads += [[a.lower(), b] for a, b in synth.reads_from(rng, genomes, 10, err=0.0, frac_random=0.0)]          # lower case
    reads += [[synth.sprinkle(rng, a, b"acgt", 0.3), synth.sprinkle(rng, b, b"acgt", 0.1)]
              for a, b in synth.reads_from(rng, genomes, 20, err=0.0, frac_random=0.0)]                         # mixed case
    reads.append([b"ACGT", genomes[0][:150]])                         # mate 1 shorter than k -> too_short
    reads.append([genomes[0][:150], genomes[0][200:200 + k - 1]])     # mate 2 shorter than k: skipped (kmer.rs:372)
    reads.append([genomes[0][:150], b"ACGTAC"])                       # ... and no panic, unlike the k-mer path
    reads.append([b"N" * 150, b"N" * 150])                            # no valid k-mers: empty set
    reads.append([genomes[1][:150], genomes[1][:150]])                # identical mates
    for kw in (dict(), dict(start_sample=0), dict(d=3), dict(group_width=8)):
        o, g = _readid_compare(oracle, oix, gix, reads, **kw)
        assert (o["kind"] != oracle.CLS_PANIC).all()
    if N <= 64:
        # the vote partitioned by row window (cid_readid_part.cu) on minimizer sets: items of m bases (XXH3's 9..16-byte path)
        ctx.set_option("readid_vote_part", 2)
        ctx.set_option("readid_part_shift", 13)
        try:
            for kw in (dict(), dict(d=3)):
                _readid_compare(oracle, oix, gix, reads, **kw)
        finally:
            ctx.set_option("readid_vote_part", 1)
            ctx.set_option("readid_part_shift", 0)
    # the set really holds minimizers: far fewer items than k-mers
    assert 0 < int(o["n_set"].max()) < 2 * (150 - k + 1) // 2


def test_read_id_classify_pipeline_on_minimizer_index(oracle, ctx):
    rng = _rng(77)
    genomes, oix, gix = _mini_index_pair(oracle, ctx, rng, 12, 31, 15, 1_000_003, 4, variant=cb.CID_MINI_OF_KMERS)
    reads = synth.reads_from(rng, genomes, 3000, read_len=150, insert=320, err=0.004, frac_random=0.25, n_rate=0.002)
    o = oix.read_id_batch(reads, threads=4)
    g = gix.read_id_classify(reads)
    assert np.array_equal(g["kind"], o["kind"])
    assert np.array_equal(g["hits"], o["hits"])
    assert np.array_equal(g["n_set"], o["n_set"])
    assert np.array_equal(g["n_top"], o["n_top"])
    for r in np.flatnonzero(o["n_top"] > 0):
        n = min(int(o["n_top"][r]), 8)
        assert g["top"][r, :n].tolist() == o["top"][r, :n].tolist()
    assert (o["kind"] == oracle.CLS_ACCEPT).sum() > 500


def test_minimizer_index_refuses_search_and_bad_sizes(ctx):
    gix = cb.Index(ctx, 50_021, 2, 21, 3, m=11)
    rng = _rng(5)
    g = synth.rand_seq(rng, 500)
    gix.build_accession_mini(0, [g], cb.CID_SEQ_FASTA, -1, cb.CID_MINI_COUNTED)
    with pytest.raises(cb.lib.CidError) as ei:       # a minimizer index must go through the minimizer builder
        gix.build_accession(1, [g])
    assert ei.value.code == cb.lib.CID_E_INVALID
    gix.finalize()
    for call in (lambda: gix.query_counts([[g[:100]]], gene_search=True, filt=0), lambda: gix.query_perfect([[g[:100]]]),
                 lambda: gix.query_perfect_mf([g[:100]])):
        with pytest.raises(cb.lib.CidError) as ei:   # main.rs:569-573
            call()
        assert ei.value.code == cb.lib.CID_E_UNSUPPORTED and "minimizers" in str(ei.value)
    with pytest.raises(cb.lib.CidError) as ei:       # find_minimizer slices seq[..m]: m > k panics in the reference
        cb.Index(ctx, 50_021, 2, 21, 3, m=22)
    assert ei.value.code == cb.lib.CID_E_REF_PANIC
    plain = cb.Index(ctx, 50_021, 2, 21, 3)
    with pytest.raises(cb.lib.CidError) as ei:
        plain.build_accession_mini(0, [g], cb.CID_SEQ_FASTA, -1, cb.CID_MINI_COUNTED)
    assert ei.value.code == cb.lib.CID_E_INVALID
