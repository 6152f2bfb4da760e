"""Seeded differential fuzzing of the C ABI against the oracle: random k / bloom_size / num_hash / accession counts,
sequences with N, IUPAC codes, lower and mixed case, very short contigs and reads, all three builders, the three
search modes and read_id (k-mer and minimizer sets).  Everything is compared bit for bit."""
import numpy as np
import pytest

import colorid_b200 as cb
from tests import synth

pytestmark = pytest.mark.gpu


def _seqs(rng, genomes, n, lo, hi, fastq=False):
    out = []
    for _ in range(n):
        g = genomes[int(rng.integers(0, len(genomes)))]
        L = int(rng.integers(lo, hi + 1))
        s = int(rng.integers(0, max(1, len(g) - L)))
        q = g[s:s + L]
        r = rng.random()
        if r < 0.25:
            q = synth.mutate(rng, q, 0.03)
        elif r < 0.35:
            q = synth.rand_seq(rng, L)
        if rng.random() < 0.3:
            q = synth.sprinkle(rng, q, b"NNRYKM", 0.02)
        if not fastq:
            c = rng.random()
            if c < 0.15:
                q = q.lower()
            elif c < 0.3:
                q = synth.sprinkle(rng, q, b"acgtn", 0.2)
        if rng.random() < 0.1:
            q = synth.revcomp(q.upper()) if fastq else q
        out.append(q)
    return out


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_build_search_read_id(oracle, ctx, seed):
    rng = np.random.default_rng(0xF0220000 + seed)
    k = int(rng.choice([1, 2, 3, 4, 5, 8, 9, 11, 15, 16, 17, 19, 21, 25, 27, 31]))
    H = int(rng.choice([1, 2, 3, 4, 5, 8]))
    N = int(rng.choice([1, 2, 5, 31, 32, 33, 63, 64, 65, 100, 129, 257]))
    S = int(rng.choice([1, 2, 7, 64, 1021, 4096, 65_537, 300_007]))
    glen = int(rng.choice([40, 200, 1500]))
    genomes = synth.clade_genomes(rng, N, glen, n_clades=min(N, 4), div=0.03)
    fasta = rng.random() < 0.7
    mode = cb.CID_SEQ_FASTA if fasta else cb.CID_SEQ_FASTQ
    omode = oracle.MODE_FASTA if fasta else oracle.MODE_FASTQ
    cutoff = int(rng.choice([-1, 0, 1])) if fasta else int(rng.choice([0, 1, 2]))   # auto_cutoff panics on tiny histograms
    oix, gix = oracle.Index(S, H, k, N), cb.Index(ctx, S, H, k, N)
    for c, g in enumerate(genomes):
        if fasta:
            acc = [g[:glen // 2], g[glen // 2:glen // 2 + int(rng.integers(0, k + 2))], g[glen // 2:]]
            if c % 3 == 1:
                acc = [synth.sprinkle(rng, a, b"Nnacgt", 0.05) for a in acc]
            if c % 5 == 2:
                acc += acc[:1]                                   # every k-mer of the first contig twice
        else:
            acc = _seqs(rng, [g], 30, max(1, k - 2), min(glen, k + 40), fastq=True) * 2
        assert gix.build_accession(c, acc, mode, cutoff) == oix.build_accession(c, acc, omode, cutoff)
    oix.finalize(threads=2)
    gix.finalize()
    assert np.array_equal(gix.download_dense(), oix.words()), (k, H, N, S)
    # search: gene mode, default report (unique hits), explicit filter; perfect search on both entry points
    queries = [[q] if i % 3 else [q, q[: len(q) // 2]] for i, q in enumerate(_seqs(rng, genomes, 25, 1, min(glen, 300)))]
    queries += [[], [b""], [b"N" * (k + 3)]]
    orng = np.random.default_rng(0xF0230000 + seed)          # path choices (a separate stream: the inputs stay as they were)
    compact, uniq_device = int(orng.choice([1, 2])), int(orng.choice([0, 1]))
    ctx.set_option("query_compact", compact)                 # 2: gather units over the compacted survivor list
    ctx.set_option("uniq_device", uniq_device)               # unique-hit summaries on the device / through the host maps
    try:
        for gene, filt, uniq in ((True, 0, False), (False, 0, True), (False, 1, True)):
            o = oix.query_counts(queries, oracle.MODE_FASTA, gene, filt)
            g = gix.query_counts(queries, cb.CID_SEQ_FASTA, gene, filt, want_uniq=uniq)
            assert np.array_equal(g["num_kmers"], o["num_kmers"]) and np.array_equal(g["counts"], o["counts"]), (k, H, N, S, gene, filt, compact)
            if uniq:
                for key in ("uniq_n", "uniq_sum", "uniq_mode"):
                    assert np.array_equal(g[key], o[key]), (key, compact, uniq_device)
    finally:
        ctx.set_option("query_compact", 1)
        ctx.set_option("uniq_device", 1)
    o, g = oix.query_perfect(queries), gix.query_perfect(queries)
    assert np.array_equal(g["status"], o["status"]) and np.array_equal(g["n_kmers"], o["n_kmers"])
    assert np.array_equal(g["and_rows"], o["and_rows"])
    recs = [q[0] for q in queries if q and all(c in b"ACGTacgt" for c in q[0])]
    if recs:
        o, g = oix.query_perfect(recs, mf=True), gix.query_perfect_mf(recs)
        assert np.array_equal(g["status"], o["status"]) and np.array_equal(g["and_rows"], o["and_rows"])
    # read_id: reports and classifications (raw-case path: upper-case reads only)
    reads = []
    for i in range(60):
        m = _seqs(rng, genomes, 2, max(1, k - 1), min(glen, 160), fastq=True)
        reads.append(m if i % 4 else m[:1])
    B = int(rng.choice([0, 1, 3]))
    d = int(rng.choice([1, 1, 2, 5]))
    o = oix.read_id_batch(reads, d=d, start_sample=B, threads=2)
    ok = o["kind"] != oracle.CLS_PANIC
    g = gix.read_id_batch(reads, d=d, start_sample=B)
    assert np.array_equal(g["n_set"][ok], o["n_set"][ok])
    assert np.array_equal((g["flags"] & 2) != 0, ~ok)
    for r in np.flatnonzero(ok):
        gd = dict(zip(g["rep_colour"][r, :g["rep_n"][r]].tolist(), g["rep_count"][r, :g["rep_n"][r]].tolist()))
        od = dict(zip(o["rep_colour"][r, :o["rep_n"][r]].tolist(), o["rep_count"][r, :o["rep_n"][r]].tolist()))
        assert gd == od, (r, k, H, N, S, B, d)
    if ok.all():
        c = gix.read_id_classify(reads, d=d, start_sample=B)
        assert np.array_equal(c["kind"], o["kind"]) and np.array_equal(c["hits"], o["hits"]) and np.array_equal(c["n_top"], o["n_top"])


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_minimizer_index(oracle, ctx, seed):
    rng = np.random.default_rng(0xF0330000 + seed)
    k = int(rng.choice([5, 9, 16, 17, 21, 27, 31]))
    m = int(rng.integers(1, k + 1))
    H = int(rng.choice([1, 2, 4, 7]))
    N = int(rng.choice([1, 3, 33, 70]))
    S = int(rng.choice([3, 509, 65_537, 300_007]))
    variant = int(rng.integers(0, 2))
    genomes = synth.clade_genomes(rng, N, 600, n_clades=min(N, 3), div=0.03)
    oix, gix = oracle.Index(S, H, k, N, m=m), cb.Index(ctx, S, H, k, N, m=m)
    cutoff = int(rng.choice([-1, 0, 1]))
    for c, g in enumerate(genomes):
        acc = [g[:300], g[300:]]
        if c % 2:
            acc = [synth.sprinkle(rng, a, b"Nnacgt", 0.08) for a in acc]
        assert gix.build_accession_mini(c, acc, cb.CID_SEQ_FASTA, cutoff, variant) == \
            oix.build_accession_mini(c, acc, oracle.MODE_FASTA, cutoff, variant)
    oix.finalize(threads=2)
    gix.finalize()
    assert np.array_equal(gix.download_dense(), oix.words()), (k, m, H, N, S, variant)
    reads = []
    for i in range(50):
        mates = _seqs(rng, genomes, 2, max(1, k - 1), 160)          # lower / mixed case allowed: minimizers are upper-cased
        reads.append(mates if i % 3 else mates[:1])
    d = int(rng.choice([1, 2, 4]))
    o = oix.read_id_batch(reads, d=d, threads=2)
    g = gix.read_id_batch(reads, d=d)
    assert np.array_equal(g["n_set"], o["n_set"])
    for r in range(len(reads)):
        gd = dict(zip(g["rep_colour"][r, :g["rep_n"][r]].tolist(), g["rep_count"][r, :g["rep_n"][r]].tolist()))
        od = dict(zip(o["rep_colour"][r, :o["rep_n"][r]].tolist(), o["rep_count"][r, :o["rep_n"][r]].tolist()))
        assert gd == od, (r, k, m, H, N, S)
