"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, bit-exact."""
import numpy as np
import pytest

import colorid_b200 as cb
from tests import synth

pytestmark = pytest.mark.gpu


def _rng(seed):
    return np.random.default_rng(0xC0101D00 + seed)


def build_both(oracle, ctx, accessions, S, H, k, mode, cutoff=-1):
    N = len(accessions)
    oix = oracle.Index(S, H, k, N)
    gix = cb.Index(ctx, S, H, k, N)
    for c, acc in enumerate(accessions):
        o_n, o_used = oix.build_accession(c, acc, mode, cutoff)
        g_n, g_used = gix.build_accession(c, acc, mode, cutoff)
        assert (g_n, g_used) == (o_n, o_used), f"accession {c}: n_ref/cutoff differ"
    oix.finalize(threads=4)
    gix.finalize()
    return oix, gix


@pytest.mark.parametrize("k", [4, 8, 9, 16, 17, 21, 27, 31])
def test_hash_rows_match_oracle(oracle, ctx, k):
    rng = _rng(k)
    H = 4
    gix2 = cb.Index(ctx, 999_983, H, k, 1)   # small matrix; only bloom_size/num_hash/k matter here
    kmers = [synth.rand_seq(rng, k) for _ in range(2000)]
    got = gix2.hash_kmers(kmers)
    exp = np.array([[oracle.xxh3_64(km, s) % 999_983 for s in range(H)] for km in kmers], dtype=np.uint64)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("k", [3, 8, 12, 21, 31])
def test_hash_variants_match_oracle(oracle, ctx, k):
    """cid_index_set_hash_variant: the 32 combinations of the places where the XXH3 drafts differ (the reference's
    `xxh3 = "0.1.1"` crate predates the freeze); every variant against the oracle's restatement."""
    rng = _rng(900 + k)
    H, S = 4, 999_983
    kmers = [synth.rand_seq(rng, k) for _ in range(300)]
    seen = set()
    for v in range(32):
        got = cb.Index(ctx, S, H, k, 1, hash_variant=v).hash_kmers(kmers)
        exp = np.array([[oracle.xxh3_64(km, s, v) % S for s in range(H)] for km in kmers], dtype=np.uint64)
        assert np.array_equal(got, exp), v
        seen.add(got.tobytes())
    if k >= 17:
        assert len(seen) == 32
    with pytest.raises(cb.CidError):
        cb.Index(ctx, S, H, k, 1, hash_variant=32)


@pytest.mark.parametrize("variant", [1, 31])
def test_build_search_read_id_under_a_hash_variant(oracle, ctx, variant):
    """Every kernel hashes through the index's variant: build (set, count table), the three search front ends, read_id on
    both paths."""
    rng = _rng(910 + variant)
    N, k, S, H = 40, 27, 300_007, 4
    genomes = synth.clade_genomes(rng, N, 5000, n_clades=5, div=0.01)
    oix, gix = oracle.Index(S, H, k, N, hash_variant=variant), cb.Index(ctx, S, H, k, N, hash_variant=variant)
    for c, g in enumerate(genomes):
        assert gix.build_accession(c, [g]) == oix.build_accession(c, [g], oracle.MODE_FASTA)
    oix.finalize(); gix.finalize()
    assert np.array_equal(gix.download_dense(), oix.words())
    stable = cb.Index(ctx, S, H, k, N)
    for c, g in enumerate(genomes):
        stable.build_accession(c, [g])
    stable.finalize()
    assert not np.array_equal(stable.download_dense(), oix.words())
    queries = [[genomes[i][200:1500]] for i in range(0, N, 5)] + [[synth.rand_seq(rng, 700)]]
    for gene, uq in ((True, False), (False, True)):
        o = oix.query_counts(queries, oracle.MODE_FASTA, gene, 0)
        g = gix.query_counts(queries, cb.CID_SEQ_FASTA, gene, 0, want_uniq=uq)
        assert np.array_equal(g["counts"], o["counts"]) and np.array_equal(g["num_kmers"], o["num_kmers"])
    op, gp = oix.query_perfect(queries), gix.query_perfect(queries)
    assert np.array_equal(gp["status"], op["status"]) and np.array_equal(gp["and_rows"], op["and_rows"])
    reads = synth.reads_from(rng, genomes, 80, read_len=150, insert=320, err=0.003, frac_random=0.2)
    reads += synth.reads_from(rng, genomes, 4, read_len=1500, insert=1600, err=0.001, frac_random=0.0, paired=False)
    reads.append([genomes[3][:150].lower()])
    _readid_compare(oracle, oix, gix, reads, order_cap=1600)
    # FASTQ read-set build through the count table (auto_cutoff)
    accs = []
    for gnm in genomes[:2]:
        rs = synth.reads_from(rng, [gnm], 600, read_len=100, insert=200, err=0.008, frac_random=0.0)
        accs.append([m for r in rs for m in r])
    o2, g2 = oracle.Index(S, H, k, 2, hash_variant=variant), cb.Index(ctx, S, H, k, 2, hash_variant=variant)
    for c, a in enumerate(accs):
        assert g2.build_accession(c, a, cb.CID_SEQ_FASTQ, -1) == o2.build_accession(c, a, oracle.MODE_FASTQ, -1)
    o2.finalize(); g2.finalize()
    assert np.array_equal(g2.download_dense(), o2.words())


@pytest.mark.parametrize("k,S,H,N", [(27, 750_000, 4, 4), (31, 300_007, 4, 46), (21, 200_003, 2, 70), (15, 65_536, 3, 33)])
def test_build_fasta_matrix_bit_exact(oracle, ctx, k, S, H, N):
    rng = _rng(N)
    genomes = synth.clade_genomes(rng, N, 6000, n_clades=4, div=0.02)
    accs = []
    for i, g in enumerate(genomes):
        contigs = [g[:2500], g[2500:2510], g[2510:]]          # a contig shorter than k is skipped
        if i % 5 == 0:
            contigs[0] = synth.sprinkle(rng, contigs[0], b"NRYn", 0.01)   # has_no_n
        if i % 7 == 0:
            contigs[2] = contigs[2].lower()                   # raw-case compare, then upper-case
        if i % 11 == 0:
            contigs[2] = synth.sprinkle(rng, contigs[2], b"acgt", 0.3)    # mixed case
        accs.append(contigs)
    oix, gix = build_both(oracle, ctx, accs, S, H, k, cb.CID_SEQ_FASTA)      # no filter: key-only set + fused Bloom insert
    assert np.array_equal(gix.download_dense(), oix.words())
    ctx.set_option("build_set", 0)                                           # the count-table path gives the same bits
    try:
        _, gix2 = build_both(oracle, ctx, accs, S, H, k, cb.CID_SEQ_FASTA)
    finally:
        ctx.set_option("build_set", 1)
    assert np.array_equal(gix2.download_dense(), oix.words())
    assert gix.nonzero_rows() == oix.nonzero_rows()
    ids, words = gix.download_nonzero_rows()
    dense = oix.words()
    nz = np.flatnonzero(dense.any(axis=1))
    assert np.array_equal(ids, nz.astype(np.uint64)) and np.array_equal(words, dense[nz])


def test_build_fastq_autocutoff(oracle, ctx):
    rng = _rng(7)
    k, S, H = 21, 400_009, 2
    genomes = synth.clade_genomes(rng, 3, 5000, n_clades=3, div=0.05)
    accs = []
    for g in genomes:
        reads = synth.reads_from(rng, [g], 1200, read_len=100, insert=200, err=0.01, frac_random=0.0, n_rate=0.002)
        accs.append([m for r in reads for m in r])
    for cutoff in (-1, 0, 2):
        oix, gix = build_both(oracle, ctx, accs, S, H, k, cb.CID_SEQ_FASTQ, cutoff)      # k <= 21: packed 8-byte count table
        assert np.array_equal(gix.download_dense(), oix.words())
    ctx.set_option("build_packed", 0)                                                    # 16-byte slots: same bits
    try:
        for cutoff in (-1, 2):
            oix, gix = build_both(oracle, ctx, accs, S, H, k, cb.CID_SEQ_FASTQ, cutoff)
            assert np.array_equal(gix.download_dense(), oix.words())
    finally:
        ctx.set_option("build_packed", 1)


def test_build_packed_count_table_falls_back_on_huge_multiplicity(oracle, ctx):
    # the packed count table keeps 22 bits of multiplicity next to a 42-bit key: a k-mer seen > 3.1 M times (a poly-A
    # run) must send the accession through the 16-byte table, with identical results
    rng = _rng(8)
    k, S, H = 15, 100_003, 2
    g = synth.rand_seq(rng, 3000)
    acc = [b"A" * 3_300_000, g, g, g[:1500]]
    oix, gix = oracle.Index(S, H, k, 1), cb.Index(ctx, S, H, k, 1)
    for cutoff in (1, 2, 3_200_000):
        assert gix.build_accession(0, acc, cb.CID_SEQ_FASTQ, cutoff) == oix.build_accession(0, acc, oracle.MODE_FASTQ, cutoff)
    oix.finalize()
    gix.finalize()
    assert np.array_equal(gix.download_dense(), oix.words()) and gix.nonzero_rows() == H


def _fast_reads(rng, genome, n, rl):
    """n single-end reads of length rl as a list of bytes (vectorised: these sets hold > 4M bases)."""
    g = np.frombuffer(genome, dtype=np.uint8)
    pos = rng.integers(0, len(g) - rl, size=n)
    arr = g[pos[:, None] + np.arange(rl)[None, :]].copy()
    err = rng.random(arr.shape) < 0.004
    arr[err] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(err.sum()))]
    return [bytes(r) for r in arr]


def test_build_fastq_optimistic_count_table(oracle, ctx):
    """Read sets of >= 4096 sequences / 4M k-mer positions get a count table sized for positions/2 (deep coverage:
    few distinct k-mers); a low-coverage set overflows it and must fall back to the safe size.  Same bits either way."""
    rng = _rng(8)
    k, S, H = 21, 1_000_003, 2
    deep = _fast_reads(rng, synth.rand_seq(rng, 150_000), 32_000, 150)        # 32x of 150 kb: ~0.7M distinct k-mers
    shallow = _fast_reads(rng, synth.rand_seq(rng, 6_000_000), 30_000, 150)   # 0.75x of 6 Mb: ~3.9M distinct of 3.9M positions
    for cutoff in (-1, 0):
        oix, gix = build_both(oracle, ctx, [deep, shallow], S, H, k, cb.CID_SEQ_FASTQ, cutoff)
        assert np.array_equal(gix.download_dense(), oix.words())


def _index_pair(oracle, ctx, rng, N, k, S, H, glen=8000):
    genomes = synth.clade_genomes(rng, N, glen, n_clades=max(2, N // 8), div=0.01)
    oix, gix = build_both(oracle, ctx, [[g] for g in genomes], S, H, k, cb.CID_SEQ_FASTA)
    return genomes, oix, gix


@pytest.mark.parametrize("N,k,S,H", [(4, 27, 750_000, 4), (46, 31, 500_009, 4), (100, 21, 300_007, 2), (1100, 21, 60_013, 2)])
def test_search_counts_and_unique_hits(oracle, ctx, N, k, S, H):
    rng = _rng(100 + N)
    genomes, oix, gix = _index_pair(oracle, ctx, rng, N, k, S, H, glen=3000 if N > 200 else 8000)
    queries = []
    for i in range(40):
        g = genomes[int(rng.integers(0, N))]
        s = int(rng.integers(0, len(g) - 2500))
        q = g[s:s + int(rng.integers(k, 2400))]
        if i % 3 == 1:
            q = synth.mutate(rng, q, 0.02)
        if i % 3 == 2:
            q = synth.rand_seq(rng, 1500)
        queries.append([q] if i % 4 else [q[:len(q) // 2], q[len(q) // 2:], b"ACGT"])
    queries.append([b"ACG"])               # shorter than k: zero k-mers
    queries.append([genomes[0]])           # a whole genome (several work units)
    try:
        for compact in (1, 2):             # 2: every query's survivors go through the dense slot list (42 regions)
            ctx.set_option("query_compact", compact)
            for gene in (True, False):
                o = oix.query_counts(queries, oracle.MODE_FASTA, gene, 0)
                g = gix.query_counts(queries, cb.CID_SEQ_FASTA, gene, 0)
                assert np.array_equal(g["num_kmers"], o["num_kmers"])
                assert np.array_equal(g["counts"], o["counts"])
                assert np.array_equal(g["uniq_n"], o["uniq_n"])
                assert np.array_equal(g["uniq_sum"], o["uniq_sum"])
                assert np.array_equal(g["uniq_mode"], o["uniq_mode"])
    finally:
        ctx.set_option("query_compact", 1)


@pytest.mark.parametrize("N,k,S,H", [(70, 21, 300_007, 2), (300, 21, 200_003, 4), (1000, 21, 100_003, 2),
                                     (1250, 31, 80_021, 4), (4000, 21, 20_011, 2)])
def test_gene_search_streaming_gather(oracle, ctx, N, k, S, H):
    """-g without unique-hit summaries takes query_hash + query_gather (cp.async ring, 1..32 lanes per row):
    row strides of 4, 12, 32, 40 and 128 words; must equal the oracle and the fused kernel."""
    rng = _rng(150 + N)
    genomes, oix, gix = _index_pair(oracle, ctx, rng, N, k, S, H, glen=2500)
    queries = []
    for i in range(60):
        g = genomes[int(rng.integers(0, N))]
        s = int(rng.integers(0, len(g) - 2100))
        q = g[s:s + int(rng.integers(k, 2000))]
        if i % 3 == 1:
            q = synth.mutate(rng, q, 0.02)
        if i % 5 == 4:
            q = synth.rand_seq(rng, 900)
        queries.append([q] if i % 4 else [q[:len(q) // 2], q[len(q) // 2:]])
    queries.append([b"ACG"])                                   # zero k-mers
    queries.append([genomes[1][:k]])                           # exactly one k-mer
    queries.append([b"".join(genomes[:12])])                   # several work items, > 8K k-mers
    o = oix.query_counts(queries, oracle.MODE_FASTA, True, 0)
    g = gix.query_counts(queries, cb.CID_SEQ_FASTA, True, 0, want_uniq=False)
    assert np.array_equal(g["num_kmers"], o["num_kmers"])
    assert np.array_equal(g["counts"], o["counts"])
    ctx.set_option("query_fused", 1)
    try:
        f = gix.query_counts(queries, cb.CID_SEQ_FASTA, True, 0, want_uniq=False)
    finally:
        ctx.set_option("query_fused", 0)
    assert np.array_equal(f["num_kmers"], o["num_kmers"])
    assert np.array_equal(f["counts"], o["counts"])
    assert int(o["counts"].max()) > 500


@pytest.mark.parametrize("N,k,S,H", [(70, 21, 300_007, 2), (1000, 21, 100_003, 2), (1250, 31, 80_021, 4), (130, 9, 50_021, 4)])
def test_gene_search_small_query_front(oracle, ctx, N, k, S, H):
    """Queries of <= 8192 k-mer positions with filter 0 take query_front (shared-memory dedup + hash, one CTA per query)
    in front of query_gather; must equal the oracle and the count-table path (option query_front = 0)."""
    rng = _rng(170 + N)
    genomes, oix, gix = _index_pair(oracle, ctx, rng, N, k, S, H, glen=2500)
    queries = []
    for i in range(90):
        g = genomes[int(rng.integers(0, N))]
        s = int(rng.integers(0, 400))
        q = g[s:s + int(rng.integers(k, 2100))]
        if i % 3 == 1:
            q = synth.mutate(rng, q, 0.02)
        if i % 5 == 4:
            q = synth.rand_seq(rng, 900)
        if i % 7 == 3:
            q = synth.sprinkle(rng, q, b"NnRacgt", 0.02)       # has_no_n, mixed case (raw compare, then upper-case)
        if i % 11 == 5:
            q = q.lower()
        queries.append([q] if i % 4 else [q[:len(q) // 2], q[len(q) // 2:], q[:k - 1], q[5:5 + k]])
    queries.append([b"ACG"])                                   # zero k-mers
    queries.append([])                                         # no sequences at all
    queries.append([genomes[1][:k]])                           # exactly one k-mer
    queries.append([genomes[2][:1500], genomes[2][:1500], synth.revcomp(genomes[2][:1500])])   # every k-mer three times
    queries.append([genomes[3][:2400], genomes[4][:2400], genomes[5][:2400], genomes[6][:900]])  # 8,000+ positions: largest table
    o = oix.query_counts(queries, oracle.MODE_FASTA, True, 0)
    launches0 = ctx.launches
    g = gix.query_counts(queries, cb.CID_SEQ_FASTA, True, 0, want_uniq=False)
    assert ctx.launches - launches0 <= 6                       # <= 4 query_front size classes + 1 gather, no count table
    assert np.array_equal(g["num_kmers"], o["num_kmers"])
    assert np.array_equal(g["counts"], o["counts"])
    ctx.set_option("query_front", 0)
    try:
        t = gix.query_counts(queries, cb.CID_SEQ_FASTA, True, 0, want_uniq=False)
    finally:
        ctx.set_option("query_front", 1)
    assert np.array_equal(t["num_kmers"], o["num_kmers"]) and np.array_equal(t["counts"], o["counts"])
    # FASTQ-mode queries (raw case kept) with -f 0 take the same path
    fq = [[synth.sprinkle(rng, q[0].upper(), b"N", 0.01)] for q in queries[:30] if q and len(q[0]) >= k]
    o = oix.query_counts(fq, oracle.MODE_FASTQ, False, 0)
    g = gix.query_counts(fq, cb.CID_SEQ_FASTQ, False, 0, want_uniq=False)
    assert np.array_equal(g["num_kmers"], o["num_kmers"]) and np.array_equal(g["counts"], o["counts"])
    assert int(o["counts"].max()) > 100


def test_search_fastq_query_with_filters(oracle, ctx):
    rng = _rng(300)
    genomes, oix, gix = _index_pair(oracle, ctx, rng, 12, 27, 400_009, 4)
    reads = synth.reads_from(rng, genomes[:2], 1500, read_len=120, insert=250, err=0.01, frac_random=0.05, n_rate=0.001)
    q = [[m for r in reads for m in r]]
    # a read repeated 1,100 times: unique-hit multiplicities beyond the device histogram's 1,024 bins -> host summary
    q_rep = [[m for r in reads[:50] for m in r] + [genomes[0][500:620]] * 1100]
    try:
        # unique-hit summaries (reports.rs:20-26) on the device / through the host maps; gather units over the sparse count
        # table / over the compacted survivor list (forced: these tables are far below the 2^22-slot threshold)
        # optimistic count table of a read-set query (forced on these small inputs): positions/4, and positions/64, which
        # overflows and is redone larger
        for uniq_device, compact, table_div in ((1, 1, 0), (0, 1, 0), (1, 2, 4), (0, 2, 64), (1, 1, 64)):
            ctx.set_option("uniq_device", uniq_device)
            ctx.set_option("query_compact", compact)
            ctx.set_option("query_table_min_slots", 0 if table_div else 1 << 23)
            ctx.set_option("query_table_div", table_div if table_div else 4)
            for qq, filters in ((q, (-1, 0, 1, 3)), (q_rep, (0, 2))):
                for filt in filters:
                    o = oix.query_counts(qq, oracle.MODE_FASTQ, False, filt)
                    g = gix.query_counts(qq, cb.CID_SEQ_FASTQ, False, filt)
                    assert np.array_equal(g["cutoff"], o["cutoff"])
                    assert np.array_equal(g["num_kmers"], o["num_kmers"])
                    assert np.array_equal(g["counts"], o["counts"])
                    assert np.array_equal(g["uniq_n"], o["uniq_n"]) and np.array_equal(g["uniq_sum"], o["uniq_sum"])
                    assert np.array_equal(g["uniq_mode"], o["uniq_mode"])
        assert o["uniq_sum"].max() >= 10 * 1100          # the repeated read did hit one accession only
    finally:
        ctx.set_option("uniq_device", 1)
        ctx.set_option("query_compact", 1)
        ctx.set_option("query_table_min_slots", 1 << 23)
        ctx.set_option("query_table_div", 4)


@pytest.mark.parametrize("N,k,S,H", [(4, 27, 750_000, 4), (70, 21, 300_007, 2), (1100, 21, 60_013, 2), (1250, 31, 80_021, 4)])
def test_perfect_search(oracle, ctx, N, k, S, H):
    rng = _rng(400 + N)
    genomes, oix, gix = _index_pair(oracle, ctx, rng, N, k, S, H, glen=3000 if N > 200 else 8000)
    queries = []
    for i in range(30):
        g = genomes[int(rng.integers(0, N))]
        s = int(rng.integers(0, len(g) - 1500))
        q = g[s:s + int(rng.integers(k, 1400))]
        if i % 3 == 2:
            q = synth.mutate(rng, q, 0.01)
        queries.append([q])
    queries.append([b"AC"])
    queries.append([synth.rand_seq(rng, 500)])
    queries.append([genomes[0][:700], genomes[0][650:1300], b"ACGTACG", genomes[0][:700]])   # several contigs, one too short
    o = oix.query_perfect(queries)
    # small queries: query_front + streaming AND gather (rows of >= 4 words, H in {2,4}); else / option off: count table
    for front in (1, 0):
        ctx.set_option("query_front", front)
        try:
            g = gix.query_perfect(queries)
        finally:
            ctx.set_option("query_front", 1)
        assert np.array_equal(g["status"], o["status"])
        assert np.array_equal(g["n_kmers"], o["n_kmers"])
        assert np.array_equal(g["and_rows"], o["and_rows"])
    # one query beyond 8192 k-mer positions sends the whole call through the count-table path
    big = queries + [[b"".join(genomes[:4])]]
    ob, gb = oix.query_perfect(big), gix.query_perfect(big)
    assert np.array_equal(gb["status"], ob["status"]) and np.array_equal(gb["and_rows"], ob["and_rows"])
    assert (o["status"] == 0).sum() >= 10


@pytest.mark.parametrize("N,H,k", [(4200, 2, 21), (5000, 4, 31), (16_384, 2, 21)])
def test_wide_rows_through_the_tma_gather(oracle, ctx, N, H, k):
    """Rows above 512 bytes (an unsharded index of thousands of accessions: 1,264-byte rows at C5) are staged by TMA bulk
    copies + mbarriers (query_gather_tma_kernel): -g counts, -s AND rows and -s -m on such an index equal the oracle's."""
    rng = _rng(470 + N)
    S = 20_011
    roots = [synth.rand_seq(rng, 1500) for _ in range(24)]
    oix, gix = oracle.Index(S, H, k, N), cb.Index(ctx, S, H, k, N)
    for c in range(N):
        g = roots[c % 24] if c % 97 else synth.mutate(rng, roots[c % 24], 0.02)
        assert gix.build_accession(c, [g]) == oix.build_accession(c, [g], oracle.MODE_FASTA)
    oix.finalize(threads=4); gix.finalize()
    queries = [[roots[i % 24][50:50 + int(rng.integers(k, 1200))]] for i in range(40)] + [[synth.rand_seq(rng, 400)], [b"ACGT"]]
    queries.append([roots[0][:600], roots[1][:600]])
    o = oix.query_counts(queries, oracle.MODE_FASTA, True, 0)
    for front in (1, 0):                           # shared-memory front end / count table + query_hash, both into the TMA gather
        ctx.set_option("query_front", front)
        try:
            g = gix.query_counts(queries, cb.CID_SEQ_FASTA, True, 0, want_uniq=False)
            gp = gix.query_perfect(queries)
        finally:
            ctx.set_option("query_front", 1)
        assert np.array_equal(g["counts"], o["counts"]) and np.array_equal(g["num_kmers"], o["num_kmers"]), front
        op = oix.query_perfect(queries)
        assert np.array_equal(gp["status"], op["status"]) and np.array_equal(gp["and_rows"], op["and_rows"]), front
    recs = [q[0] for q in queries] + [roots[2][:100] + b"N" + roots[2][101:500]]
    om, gm = oix.query_perfect(recs, mf=True), gix.query_perfect_mf(recs)
    assert np.array_equal(gm["status"], om["status"]) and np.array_equal(gm["and_rows"], om["and_rows"])
    # a multi-unit query (more than 16,384 distinct k-mers: counts are added with atomics across its work units)
    big = [[b"".join(roots)]]
    ob, gb = oix.query_counts(big, oracle.MODE_FASTA, True, 0), gix.query_counts(big, cb.CID_SEQ_FASTA, True, 0, want_uniq=False)
    assert np.array_equal(gb["counts"], ob["counts"]) and int(ob["num_kmers"][0]) > 16_384


@pytest.mark.parametrize("N,H", [(4, 4), (46, 4), (40, 3), (150, 3), (4200, 2)])
def test_perfect_search_multifasta_non_acgt_on_any_row_shape(oracle, ctx, N, H):
    """kmerize_string (kmer.rs:271-299) has no has_no_n test: -s -m records with N / IUPAC / U bytes are exact on every row
    shape -- rows of 1 or 2 words (the C1 index, narrow column shards), num_hash 3, rows above 512 bytes -- through
    query_front<STRINGM> + perfect_rids_kernel where the streaming gather does not apply."""
    rng = _rng(460 + N)
    k, S = 21, 100_003
    genomes, oix, gix = _index_pair(oracle, ctx, rng, N, k, S, H, glen=1500 if N > 1000 else 4000)
    recs = []
    for i in range(24):
        g = genomes[int(rng.integers(0, N))]
        q = g[100:100 + int(rng.integers(k, 1200))]
        if i % 3 == 0:
            q = synth.sprinkle(rng, q, b"NnRYUu", 0.01)
        recs.append(q)
    recs += [b"ACGTN", b"N" * 40, genomes[0][:300]]
    o = oix.query_perfect(recs, mf=True)
    g = gix.query_perfect_mf(recs)
    assert np.array_equal(g["status"], o["status"]) and np.array_equal(g["n_kmers"], o["n_kmers"]) and np.array_equal(g["and_rows"], o["and_rows"])
    assert (o["status"] == 0).sum() >= 5
    qs = [[r] for r in recs if b"N" not in r.upper() and b"R" not in r and b"Y" not in r and b"U" not in r.upper()]
    o2, g2 = oix.query_perfect(qs), gix.query_perfect(qs)
    assert np.array_equal(g2["status"], o2["status"]) and np.array_equal(g2["and_rows"], o2["and_rows"])


def test_perfect_search_multifasta_records(oracle, ctx):
    """-s -m: perfect_search.rs:62-120 batch_search_mf over kmer.rs:271-299 kmerize_string (one query per record)."""
    rng = _rng(451)
    N, k, S, H = 70, 21, 300_007, 2
    genomes, oix, gix = _index_pair(oracle, ctx, rng, N, k, S, H, glen=8000)
    recs = []
    for i in range(40):
        g = genomes[int(rng.integers(0, N))]
        s = int(rng.integers(0, len(g) - 1500))
        q = g[s:s + int(rng.integers(k, 1400))]
        if i % 4 == 3:
            q = synth.mutate(rng, q, 0.01)
        if i % 5 == 1:
            q = q.lower()                      # raw-case compare, then to_uppercase (kmer.rs:280-283)
        recs.append(q)
    recs.append(b"ACGTAC")                     # shorter than k -> None -> "Warning! no kmers in query"
    recs.append(genomes[3][100:100 + k])       # exactly one k-mer
    recs.append(synth.rand_seq(rng, 700))
    o = oix.query_perfect(recs, mf=True)
    for front in (1, 0):
        ctx.set_option("query_front", front)
        try:
            g = gix.query_perfect_mf(recs)
        finally:
            ctx.set_option("query_front", 1)
        assert np.array_equal(g["status"], o["status"])
        assert np.array_equal(g["n_kmers"], o["n_kmers"])
        assert np.array_equal(g["and_rows"], o["and_rows"])
    assert (o["status"] == 0).sum() >= 10 and (o["status"] == 2).sum() == 1
    # kmerize_string has no has_no_n test: windows holding N, IUPAC codes, U/u ... are k-mers too (hashed byte-wise,
    # upper-cased; U -> A in the reverse complement can even give a plain ACGT string).  Records of <= 8192 positions
    # take query_front, which handles them; the count-table path cannot represent them and refuses loudly.
    g0 = genomes[0]
    dirty = [g0[:200] + b"N" + g0[201:400], g0[:300].replace(b"T", b"U"), g0[100:400].lower().replace(b"t", b"u"),
             g0[:150] + b"RYKMnn-*" + g0[150:300], b"N" * (k + 5), g0[:k - 1] + b"N", g0[:k] + b"N" + g0[:k], b"U" * 60,
             synth.sprinkle(rng, g0[:600], b"NnUuRacgt", 0.05), g0[:500]]
    o = oix.query_perfect(dirty, mf=True)
    g = gix.query_perfect_mf(dirty)
    assert np.array_equal(g["n_kmers"], o["n_kmers"]) and np.array_equal(g["status"], o["status"])
    assert np.array_equal(g["and_rows"], o["and_rows"])
    assert o["status"][-1] == 0 and (o["status"][:-1] == 1).sum() >= 5          # absent rows: "No perfect hits!"
    dense_o, dense_g = oracle.Index(101, H, k, N), cb.Index(ctx, 101, H, k, N)   # 101 rows, all present: the AND is computed
    for c in range(N):
        assert dense_g.build_accession(c, [genomes[c]]) == dense_o.build_accession(c, [genomes[c]], oracle.MODE_FASTA)
    dense_o.finalize()
    dense_g.finalize()
    o, g = dense_o.query_perfect(dirty, mf=True), dense_g.query_perfect_mf(dirty)
    assert np.array_equal(g["n_kmers"], o["n_kmers"]) and np.array_equal(g["status"], o["status"])
    assert np.array_equal(g["and_rows"], o["and_rows"]) and (o["status"] == 0).sum() >= 8
    ctx.set_option("query_front", 0)
    try:
        with pytest.raises(cb.lib.CidError) as ei:
            gix.query_perfect_mf(dirty[:1])
        assert ei.value.code == cb.lib.CID_E_UNSUPPORTED
    finally:
        ctx.set_option("query_front", 1)


def _readid_compare(oracle, oix, gix, reads, order_cap=600, **kw):
    okw = dict(kw)
    o = oix.read_id_batch(reads, order_cap=order_cap, **okw)
    on, osq, opos = gix.read_kmer_order(reads, d=kw.get("d", 1), group_width=kw.get("group_width", 16),
                                        reserve_before_find=kw.get("reserve_before_find", True), order_cap=order_cap)
    ok_reads = o["kind"] != oracle.CLS_PANIC
    assert np.array_equal(on[ok_reads], o["order_n"][ok_reads])
    for r in np.flatnonzero(ok_reads):
        n = on[r]
        assert np.array_equal(osq[r, :n], o["order_seq"][r, :n]), f"read {r}: k-mer order (mate) differs"
        assert np.array_equal(opos[r, :n], o["order_pos"][r, :n]), f"read {r}: k-mer order (pos) differs"
    g = gix.read_id_batch(reads, d=kw.get("d", 1), start_sample=kw.get("start_sample", 3),
                          group_width=kw.get("group_width", 16), reserve_before_find=kw.get("reserve_before_find", True))
    assert np.array_equal(g["n_set"][ok_reads], o["n_set"][ok_reads])
    too_short = (g["flags"] & 1) != 0
    assert np.array_equal(too_short, o["kind"] == oracle.CLS_TOO_SHORT)
    assert np.array_equal((g["flags"] & 2) != 0, o["kind"] == oracle.CLS_PANIC)
    # the device reports final_report in insertion order; the oracle in FnvHashMap iteration order:
    # compare as sets here, ordering of ties is checked in test_host_vote
    for r in np.flatnonzero(ok_reads):
        gd = dict(zip(g["rep_colour"][r, :g["rep_n"][r]].tolist(), g["rep_count"][r, :g["rep_n"][r]].tolist()))
        od = dict(zip(o["rep_colour"][r, :o["rep_n"][r]].tolist(), o["rep_count"][r, :o["rep_n"][r]].tolist()))
        assert gd == od, f"read {r}: report differs {gd} vs {od}"
        # insertion order -> emulated map iteration order must reproduce the oracle's sequence
        keys = g["rep_colour"][r, :g["rep_n"][r]].astype(np.uint64)
        order = oracle.hashmap_usize_order(keys, kw.get("group_width", 16))
        assert keys[order].tolist() == o["rep_colour"][r, :o["rep_n"][r]].tolist()
    return o, g


@pytest.mark.parametrize("N,k,S,H", [(4, 27, 750_000, 4), (46, 31, 2_000_003, 4), (40, 21, 100_003, 2)])
def test_read_id_narrow_rows(oracle, ctx, N, k, S, H):
    rng = _rng(500 + N)
    genomes, oix, gix = _index_pair(oracle, ctx, rng, N, k, S, H)
    reads = synth.reads_from(rng, genomes, 300, read_len=150, insert=320, err=0.004, frac_random=0.25, n_rate=0.002)
    reads += synth.reads_from(rng, genomes, 60, read_len=150, insert=200, err=0.0, frac_random=0.0)   # overlapping mates
    reads += synth.reads_from(rng, genomes, 40, read_len=100, insert=300, err=0.0, frac_random=0.0, paired=False)
    reads.append([b"ACGT", genomes[0][:150]])                         # mate 1 shorter than k -> too_short
    reads.append([genomes[0][:150], genomes[0][200:200 + k - 1]])     # mate 2 of length k-1: no k-mers, no panic
    reads.append([genomes[0][:150], b"ACGTAC"])                       # mate 2 shorter than k-1: reference panics
    reads.append([b"N" * 150, b"N" * 150])                            # no valid k-mers: empty report
    reads.append([genomes[1][:150], genomes[1][:150]])                # identical mates: every k-mer duplicated
    for kw in (dict(), dict(start_sample=0), dict(start_sample=1), dict(d=3), dict(reserve_before_find=False),
               dict(group_width=8)):
        _readid_compare(oracle, oix, gix, reads, **kw)


def test_read_id_reads_up_to_1000_bases(oracle, ctx):
    """Reads (all mates together) of up to 1,000 bases: the general order kernel with u16 tables of up to 2,048 buckets, fewer
    reads per CTA as the tables grow (longer reads: test_read_id_long_reads)."""
    rng = _rng(950)
    genomes, oix, gix = _index_pair(oracle, ctx, rng, 20, 31, 1_000_003, 4, glen=6000)
    reads = []
    for L_ in (300, 520, 700, 901, 990, 1000):
        reads += synth.reads_from(rng, genomes, 12, read_len=L_, insert=L_ + 50, err=0.002, frac_random=0.2, paired=False)
    reads += synth.reads_from(rng, genomes, 12, read_len=480, insert=700, err=0.002, frac_random=0.1)       # 2 x 480
    reads += synth.reads_from(rng, genomes, 20, read_len=150, insert=320, err=0.004, frac_random=0.2)
    for kw in (dict(d=4), dict(start_sample=0), dict()):
        o, g = _readid_compare(oracle, oix, gix, reads, **kw)
    assert int(o["n_set"].max()) > 900
    assert not (g["flags"] & 16).any()          # none of these took the general path


def test_read_id_long_reads(oracle, ctx):
    """Reads above 1,000 bases (read_id_mt_pe.rs:450-569 stream_fasta hands whole contigs to parallel_vec; long-read FASTQ)
    go through the CTA-per-read general path: the same k-mer set order (hashbrown tables of up to 65,536 buckets here),
    reports and flags as the oracle, mixed with short reads in one batch."""
    rng = _rng(951)
    genomes, oix, gix = _index_pair(oracle, ctx, rng, 20, 31, 1_000_003, 4, glen=30_000)
    gix.n_ref[:] = oix.n_ref
    reads = []
    for L_ in (1001, 1023, 1500, 2048, 4000, 12_000):
        reads += synth.reads_from(rng, genomes, 4, read_len=L_, insert=L_ + 50, err=0.002, frac_random=0.25, paired=False)
    reads += synth.reads_from(rng, genomes, 4, read_len=800, insert=2000, err=0.001, frac_random=0.0)            # 2 x 800
    reads += synth.reads_from(rng, genomes, 4, read_len=700, insert=900, err=0.0, frac_random=0.0)               # overlapping mates
    reads += synth.reads_from(rng, genomes, 30, read_len=150, insert=320, err=0.004, frac_random=0.2)            # fast path
    reads.append([genomes[3]])                                           # a whole accession: no absent row, counts of ~30k
    reads.append([genomes[5][:9000], synth.revcomp(genomes[5][4000:15_000])])
    reads.append([synth.sprinkle(rng, genomes[7][:5000], b"NRn", 0.01)])
    reads.append([genomes[2][:3000] + genomes[2][:3000]])                # every k-mer twice: duplicates at the tail
    reads.append([b"ACGT", genomes[0][:2000]])                           # too_short
    reads.append([genomes[0][:2000], b"ACGTAC"])                         # reference panics
    reads.append([b"N" * 1500])
    reads.append([synth.rand_seq(rng, 1800)])
    for kw in (dict(), dict(start_sample=0), dict(d=3, start_sample=1), dict(reserve_before_find=False), dict(group_width=8)):
        o, g = _readid_compare(oracle, oix, gix, reads, order_cap=31_000, **kw)
    longish = np.array([sum(len(m) for m in r) > 1000 and len(r[0]) >= 31 for r in reads])      # (too_short never leaves the fast kernel)
    assert np.array_equal((g["flags"] & 16) != 0, longish)
    assert int(o["n_set"].max()) > 29_000
    # the fused pipeline (device vote + host re-vote) on the same batch, one chunk and several
    oc = oix.read_id_batch(reads)
    try:
        for chunk in (0, 9):
            ctx.set_option("readid_chunk_reads", chunk)
            gc = gix.read_id_classify(reads)
            for key in ("kind", "hits", "n_set", "n_top"):
                assert np.array_equal(gc[key], oc[key]), f"chunk={chunk}: {key}"
            for r in range(len(reads)):
                nt = min(int(oc["n_top"][r]), 8)
                assert gc["top"][r, :nt].tolist() == oc["top"][r, :nt].tolist()
    finally:
        ctx.set_option("readid_chunk_reads", 0)


def test_read_id_table_exactly_full_at_the_tail(oracle, ctx):
    """hashbrown >= 0.14 reserves before it looks the key up: a duplicate insert arriving when the table is exactly full
    resizes it.  Sets of exactly 7/8 * 2^n k-mers followed by a duplicate, on both read_id paths and both growth rules."""
    rng = _rng(952)
    k = 21
    genomes, oix, gix = _index_pair(oracle, ctx, rng, 6, k, 300_007, 3, glen=9000)
    reads = []
    for nk in (14, 28, 56, 112, 224, 448, 896, 1792, 3584, 7168):
        g = genomes[int(rng.integers(0, 6))]
        first = g[:nk + k - 1]                                       # nk distinct k-mers
        reads.append([first, first[:k + 3]])                         # + duplicates at the tail
        reads.append([first])                                        # no tail duplicate
        reads.append([first + g[nk + k - 1:nk + k + 1]])             # two more fresh k-mers
    for kw in (dict(), dict(reserve_before_find=False)):
        _readid_compare(oracle, oix, gix, reads, order_cap=8000, **kw)


def test_read_id_lowercase_kmers(oracle, ctx):
    """kmer.rs:221-243 never upper-cases: a soft-masked (lower-case) k-mer of a read is hashed with its raw bytes, for the set
    order and for the rows.  Such reads take the general path; upper-case reads of the same batch stay on the fast one."""
    rng = _rng(953)
    genomes, oix, gix = _index_pair(oracle, ctx, rng, 12, 27, 500_009, 4, glen=8000)
    gix.n_ref[:] = oix.n_ref
    base = synth.reads_from(rng, genomes, 60, read_len=150, insert=320, err=0.002, frac_random=0.1)
    reads = []
    for i, r in enumerate(base):
        if i % 4 == 0: r = [m.lower() for m in r]                                   # all lower case
        elif i % 4 == 1: r = [r[0][:70] + r[0][70:110].lower() + r[0][110:], r[1]]  # a soft-masked stretch
        elif i % 4 == 2: r = [synth.sprinkle(rng, m, b"acgtn", 0.02) for m in r]    # scattered lower-case bases
        reads.append(r)
    reads.append([genomes[1][:4000].lower()])                                       # long and lower case
    reads.append([genomes[1][:40] + b"n" + genomes[1][41:150]])                     # 'n' is not a base: no lower-case k-mer
    def has_lower_kmer(read, k, d):
        for m in read:
            for i in range(0, len(m) - k + 1, d):
                w = m[i:i + k]
                if all(c in b"ACGTacgt" for c in w) and any(c in b"acgt" for c in w):
                    return True
        return False
    for kw in (dict(), dict(start_sample=0), dict(d=2)):
        o, g = _readid_compare(oracle, oix, gix, reads, order_cap=4200, **kw)
        general = np.array([has_lower_kmer(r, 27, kw.get("d", 1)) or sum(len(m) for m in r) > 1000 for r in reads])
        assert np.array_equal((g["flags"] & 16) != 0, general)
        assert general.sum() > 30 and not general[-1]
    oc = oix.read_id_batch(reads)
    gc = gix.read_id_classify(reads)
    for key in ("kind", "hits", "n_set", "n_top"):
        assert np.array_equal(gc[key], oc[key]), key


def test_read_id_wide_rows(oracle, ctx):
    rng = _rng(900)
    genomes, oix, gix = _index_pair(oracle, ctx, rng, 150, 21, 300_007, 2, glen=4000)
    reads = synth.reads_from(rng, genomes, 200, read_len=150, insert=320, err=0.003, frac_random=0.2)
    for kw in (dict(), dict(start_sample=0)):
        _readid_compare(oracle, oix, gix, reads, **kw)


def test_read_id_quality_mask_on_device(oracle, ctx):
    rng = _rng(950)
    genomes, oix, gix = _index_pair(oracle, ctx, rng, 6, 27, 750_000, 4)
    reads = synth.reads_from(rng, genomes, 120, read_len=150, insert=320, err=0.002, frac_random=0.1)
    quals = [[bytes(rng.choice(np.frombuffer(b"#+5?I", dtype=np.uint8), size=len(m), p=[.03, .03, .04, .3, .6]).tolist())
              for m in r] for r in reads]
    masked = [[oracle.qual_mask(m, q, 15) for m, q in zip(r, qr)] for r, qr in zip(reads, quals)]
    o = oix.read_id_batch(masked)
    g = gix.read_id_batch(reads, quals=quals, qual_offset=15)
    assert np.array_equal(g["n_set"], o["n_set"])
    for r in range(len(reads)):
        gd = dict(zip(g["rep_colour"][r, :g["rep_n"][r]].tolist(), g["rep_count"][r, :g["rep_n"][r]].tolist()))
        od = dict(zip(o["rep_colour"][r, :o["rep_n"][r]].tolist(), o["rep_count"][r, :o["rep_n"][r]].tolist()))
        assert gd == od


def test_read_id_classify_pipeline_equals_oracle(oracle, ctx):
    """cid_read_id_classify (chunked H2D / kernels / D2H / host vote pipeline) == the oracle's parallel_vec,
    for one chunk and for many small chunks, and == cid_read_id_batch + cid_classify_reads."""
    rng = _rng(1234)
    genomes, oix, gix = _index_pair(oracle, ctx, rng, 46, 31, 2_000_003, 4)
    gix.n_ref[:] = oix.n_ref
    reads = synth.reads_from(rng, genomes, 700, read_len=150, insert=330, err=0.004, frac_random=0.25, n_rate=0.002)
    reads.append([b"ACGT", genomes[0][:150]])
    reads.append([b"N" * 150, b"N" * 150])
    reads += synth.reads_from(rng, genomes, 45, read_len=120, insert=300, err=0.0, frac_random=0.0, paired=False)
    o = oix.read_id_batch(reads)
    rep = gix.read_id_batch(reads)
    two_step = cb.api.classify_reads((gix.S, gix.H, gix.N), gix.n_ref, rep)
    try:
        for chunk in (0, 64, 97):
            ctx.set_option("readid_chunk_reads", chunk)
            g = gix.read_id_classify(reads)
            for key in ("kind", "hits", "n_set", "n_top"):
                assert np.array_equal(g[key], o[key]), f"chunk={chunk}: {key} differs from the oracle"
            for r in range(len(reads)):
                nt = min(int(o["n_top"][r]), 8)
                assert g["top"][r, :nt].tolist() == o["top"][r, :nt].tolist(), f"chunk={chunk}: read {r} top differs"
            for key in ("kind", "hits", "n_top"):
                assert np.array_equal(g[key], two_step[key])
            # the chunked raw-report path must equal the single-chunk one
            rep2 = gix.read_id_batch(reads)
            for key in ("n_set", "flags", "rep_n"):
                assert np.array_equal(rep2[key], rep[key])
            for r in range(len(reads)):
                n = rep["rep_n"][r]
                assert np.array_equal(rep2["rep_colour"][r, :n], rep["rep_colour"][r, :n])
                assert np.array_equal(rep2["rep_count"][r, :n], rep["rep_count"][r, :n])
    finally:
        ctx.set_option("readid_chunk_reads", 0)


def test_read_id_classify_ties_go_through_host_vote(oracle, ctx):
    """Identical accessions: every classified read ties at the top, so the device defers all of them to the host
    vote (names joined in FnvHashMap iteration order)."""
    rng = _rng(1236)
    g0, g1 = synth.rand_seq(rng, 6000), synth.rand_seq(rng, 6000)
    genomes = [g0, g0, g1, g0, g1, synth.mutate(rng, g1, 0.01)]
    oix, gix = build_both(oracle, ctx, [[g] for g in genomes], 400_009, 3, 25, cb.CID_SEQ_FASTA)
    gix.n_ref[:] = oix.n_ref
    reads = synth.reads_from(rng, genomes, 400, read_len=120, insert=300, err=0.002, frac_random=0.1)
    o = oix.read_id_batch(reads)
    assert (o["kind"] == oracle.CLS_REJECT_MULTI).sum() > 100
    try:
        for chunk in (0, 90):
            ctx.set_option("readid_chunk_reads", chunk)
            g = gix.read_id_classify(reads)
            for key in ("kind", "hits", "n_set", "n_top"):
                assert np.array_equal(g[key], o[key]), key
            for r in range(len(reads)):
                nt = min(int(o["n_top"][r]), 8)
                assert g["top"][r, :nt].tolist() == o["top"][r, :nt].tolist()
    finally:
        ctx.set_option("readid_chunk_reads", 0)


@pytest.mark.parametrize("N,k,S,H", [(46, 31, 2_000_003, 4), (150, 21, 300_007, 2)])
def test_read_id_packed_input(oracle, ctx, N, k, S, H):
    """cid_pack_reads + cid_read_id_classify_packed: quality masking and 2-bit packing on the host (where the reference masks:
    seq.rs:36-56 from read_id_mt_pe.rs:733-760), planes to the device.  Same classifications as the oracle on the masked
    reads: narrow and wide rows, N / IUPAC bytes, soft-masked reads (lower plane), long reads, too-short and panic reads,
    one chunk and several."""
    rng = _rng(1300 + N)
    genomes, oix, gix = _index_pair(oracle, ctx, rng, N, k, S, H, glen=6000)
    gix.n_ref[:] = oix.n_ref
    reads = synth.reads_from(rng, genomes, 500, read_len=150, insert=320, err=0.003, frac_random=0.2, n_rate=0.003)
    reads += synth.reads_from(rng, genomes, 40, read_len=97, insert=250, err=0.002, frac_random=0.1, paired=False)
    reads += synth.reads_from(rng, genomes, 5, read_len=2100, insert=2200, err=0.001, frac_random=0.0, paired=False)
    reads += [[b"ACGT", genomes[0][:150]], [genomes[0][:150], b"ACGTAC"], [b"N" * 150, b"N" * 150], [genomes[1][:k]], [genomes[1][:150], b""]]
    quals = [[bytes(rng.choice(np.frombuffer(b"#+5?I", dtype=np.uint8), size=len(m), p=[.02, .02, .03, .33, .6]).tolist()) for m in r] for r in reads]
    for soft in (False, True):
        rr = list(reads)
        if soft:
            rr = _softmask(rng, rr, frac=0.2)
        masked = [[oracle.qual_mask(m, q, 15) for m, q in zip(r, qr)] for r, qr in zip(rr, quals)]
        o = oix.read_id_batch(masked)
        pk = cb.pack_reads(rr, quals, 15)
        assert bool(pk["flags"] & 1) == soft
        try:
            for chunk in (0, 101):
                ctx.set_option("readid_chunk_reads", chunk)
                for kw in (dict(), dict(start_sample=0, d=2)):
                    oo = o if not kw else oix.read_id_batch(masked, **kw)
                    g = gix.read_id_classify_packed(pk, **kw)
                    for key in ("kind", "hits", "n_set", "n_top"):
                        assert np.array_equal(g[key], oo[key]), (soft, chunk, kw, key)
                    for r in range(len(rr)):
                        nt = min(int(oo["n_top"][r]), 8)
                        assert g["top"][r, :nt].tolist() == oo["top"][r, :nt].tolist()
        finally:
            ctx.set_option("readid_chunk_reads", 0)


def test_read_id_classify_with_quals(oracle, ctx):
    rng = _rng(1235)
    genomes, oix, gix = _index_pair(oracle, ctx, rng, 6, 27, 750_000, 4)
    gix.n_ref[:] = oix.n_ref
    reads = synth.reads_from(rng, genomes, 300, read_len=150, insert=320, err=0.002, frac_random=0.1)
    quals = [[bytes(rng.choice(np.frombuffer(b"#+5?I", dtype=np.uint8), size=len(m), p=[.03, .03, .04, .3, .6]).tolist())
              for m in r] for r in reads]
    masked = [[oracle.qual_mask(m, q, 15) for m, q in zip(r, qr)] for r, qr in zip(reads, quals)]
    o = oix.read_id_batch(masked)
    try:
        ctx.set_option("readid_chunk_reads", 70)
        g = gix.read_id_classify(reads, quals=quals, qual_offset=15)
    finally:
        ctx.set_option("readid_chunk_reads", 0)
    for key in ("kind", "hits", "n_set", "n_top"):
        assert np.array_equal(g[key], o[key]), key


def test_upload_rows_roundtrip(oracle, ctx):
    rng = _rng(77)
    genomes, oix, gix = _index_pair(oracle, ctx, rng, 40, 21, 100_003, 2)
    ids, words = gix.download_nonzero_rows()
    g2 = cb.Index(ctx, 100_003, 2, 21, 40)
    g2.upload_rows(ids, words)
    assert np.array_equal(g2.download_dense(), oix.words())
    q = [[genomes[3][100:900]]]
    assert np.array_equal(g2.query_counts(q, gene_search=True, filt=0)["counts"], oix.query_counts(q, 0, True, 0)["counts"])


def _softmask(rng, reads, frac=0.4):
    """Lower-case stretches, whole lower-case mates and scattered lower-case bases in a fraction of the reads."""
    out = []
    for r in reads:
        u = rng.random()
        if u < frac / 3:
            r = [m.lower() for m in r]
        elif u < 2 * frac / 3:
            a = int(rng.integers(0, max(1, len(r[0]) - 40)))
            r = [r[0][:a] + r[0][a:a + 40].lower() + r[0][a + 40:]] + list(r[1:])
        elif u < frac:
            r = [synth.sprinkle(rng, m, b"acgt", 0.03) for m in r]
        out.append(r)
    return out


@pytest.mark.parametrize("k", [15, 21, 27, 31])
def test_build_and_search_fastq_lowercase_kmers(oracle, ctx, k):
    """kmers_from_fq_qual / kmers_fq_pe_qual (kmer.rs:461-510,581-655) never upper-case: a lower-case k-mer of a read set is
    counted, filtered and hashed with its raw bytes, as a k-mer of its own.  Builds (auto_cutoff, -f 0, -f 2; packed 8-byte
    table for k <= 21, key-only set, 16-byte slots) and FASTQ searches (shared-memory front end, count table, compacted
    survivors, unique-hit summaries) all redo such inputs through the case-aware count table."""
    rng = _rng(40 + k)
    S, H = 400_009, 3
    genomes = synth.clade_genomes(rng, 4, 4000, n_clades=2, div=0.03)
    accs = []
    for i, g in enumerate(genomes):
        reads = synth.reads_from(rng, [g], 700, read_len=100, insert=200, err=0.008, frac_random=0.0, n_rate=0.002)
        if i != 1:
            reads = _softmask(rng, reads)
        accs.append([m for r in reads for m in r])
    for cutoff in (-1, 0, 2):
        oix, gix = build_both(oracle, ctx, accs, S, H, k, cb.CID_SEQ_FASTQ, cutoff)
        assert np.array_equal(gix.download_dense(), oix.words()), f"cutoff {cutoff}"
    # the same read sets as FASTQ queries against the last index (cutoff 2): default report with unique-hit summaries
    queries = [accs[0], accs[1], accs[2][:300], [accs[3][0].lower()]]
    for filt in (-1, 0, 1):
        o = oix.query_counts(queries, oracle.MODE_FASTQ, False, filt)
        for opts in (dict(), dict(query_compact=2), dict(query_fused=1)):
            for name, v in opts.items():
                ctx.set_option(name, v)
            try:
                for uq in (True, False):
                    g = gix.query_counts(queries, cb.CID_SEQ_FASTQ, False, filt, want_uniq=uq)
                    assert np.array_equal(g["num_kmers"], o["num_kmers"]), (filt, opts, uq)
                    assert np.array_equal(g["counts"], o["counts"]), (filt, opts, uq)
                    assert np.array_equal(g["cutoff"], o["cutoff"])
                    if uq:
                        for key in ("uniq_n", "uniq_sum", "uniq_mode"):
                            assert np.array_equal(g[key], o[key]), (filt, opts, key)
            finally:
                ctx.set_option("query_compact", 1)
                ctx.set_option("query_fused", 0)
    # lower-case reads hit lower-case k-mers of an index built from soft-masked read sets
    gix.n_ref[:] = oix.n_ref
    reads = [[accs[0][i], accs[0][i + 1]] for i in range(0, 120, 2)] + [[accs[2][i]] for i in range(0, 60)]
    o, g = _readid_compare(oracle, oix, gix, reads)
    assert (g["flags"] & 16).sum() > 10 and int(o["rep_n"].max()) >= 1
    low_hits = [r for r in range(len(reads)) if (g["flags"][r] & 16) and any(c < 4 for c in o["rep_colour"][r, :o["rep_n"][r]])]
    assert len(low_hits) > 3, "no lower-case read with a hit: the test does not exercise raw-case row hashing"


def test_empty_and_degenerate_inputs(oracle, ctx):
    """Empty batches, empty sequences and queries without k-mers go through every entry point without touching memory
    they should not (run under compute-sanitizer when changing kernels) and agree with the oracle."""
    rng = _rng(77)
    N, k, S, H = 70, 21, 100_003, 2
    genomes, oix, gix = _index_pair(oracle, ctx, rng, N, k, S, H, glen=3000)
    # no queries / no reads at all
    g = gix.query_counts([], cb.CID_SEQ_FASTA, True, 0, want_uniq=False)
    assert g["counts"].shape == (0, N)
    assert gix.query_perfect([])["and_rows"].shape[0] == 0
    assert gix.query_perfect_mf([])["status"].shape[0] == 0
    assert len(gix.read_id_classify([])["kind"]) == 0
    assert len(gix.read_id_batch([])["rep_n"]) == 0
    # queries made only of empty / too-short / all-N sequences, mixed with a real one
    queries = [[b""], [b"", b"ACGT"], [b"N" * 100], [genomes[3][10:400]], [b"", genomes[4][:k], b""]]
    o = oix.query_counts(queries, oracle.MODE_FASTA, True, 0)
    for uq in (True, False):
        g = gix.query_counts(queries, cb.CID_SEQ_FASTA, True, 0, want_uniq=uq)
        assert np.array_equal(g["num_kmers"], o["num_kmers"]) and np.array_equal(g["counts"], o["counts"])
    assert o["num_kmers"].tolist()[:3] == [0, 0, 0] and o["num_kmers"][4] == 1
    op, gp = oix.query_perfect(queries), gix.query_perfect(queries)
    assert np.array_equal(gp["status"], op["status"]) and np.array_equal(gp["and_rows"], op["and_rows"])
    # reads: empty mates, one-base mates, a read whose only k-mer sits at the very end
    reads = [[genomes[0][:150], b""], [b"A"], [b"N" * 40 + genomes[2][:k]], [genomes[1][:k]]]
    _readid_compare(oracle, oix, gix, reads)
    # an accession without any k-mer leaves its column empty and n_ref_kmers == 0
    gi2 = cb.Index(ctx, S, H, k, 3)
    oi2 = oracle.Index(S, H, k, 3)
    for c, seqs in enumerate([[b""], [genomes[0]], [b"ACGTN" * 3]]):
        assert gi2.build_accession(c, seqs, cb.CID_SEQ_FASTA) == oi2.build_accession(c, seqs, oracle.MODE_FASTA)
    gi2.finalize()
    oi2.finalize()
    assert np.array_equal(gi2.download_dense(), oi2.words())


def _reports_equal(a, b):
    for key in ("n_set", "flags", "rep_n"):
        assert np.array_equal(a[key], b[key]), key
    for r in range(len(a["rep_n"])):
        n = a["rep_n"][r]
        assert np.array_equal(a["rep_colour"][r, :n], b["rep_colour"][r, :n]), f"read {r}: report order"
        assert np.array_equal(a["rep_count"][r, :n], b["rep_count"][r, :n]), f"read {r}: report counts"


@pytest.mark.parametrize("N,k,S,H,clades", [(46, 31, 2_000_003, 4, 6), (64, 21, 300_007, 2, 3), (20, 27, 750_000, 3, 10), (30, 31, 1_000_003, 4, 2)])
def test_read_id_partitioned_vote(oracle, ctx, N, k, S, H, clades):
    """cid_readid_part.cu (scan / per-window gather / count) against the oracle and, entry by entry (insertion order, flags),
    against the one-kernel vote: reads with 0, 1..8 and more than 8 candidate colours (clades of up to 21 near-identical
    genomes), misses at every position, sets shorter than -B, -B 1/3/7/15, -d 3, 4- and 8-byte rows, run-time num_hash."""
    rng = _rng(9100 + N)
    genomes = synth.clade_genomes(rng, N, 8000, n_clades=clades, div=0.004)
    oix, gix = build_both(oracle, ctx, [[g] for g in genomes], S, H, k, cb.CID_SEQ_FASTA)
    reads = synth.reads_from(rng, genomes, 500, read_len=150, insert=320, err=0.003, frac_random=0.2, n_rate=0.002)
    reads += synth.reads_from(rng, genomes, 100, read_len=150, insert=200, err=0.0, frac_random=0.0)
    reads += synth.reads_from(rng, genomes, 60, read_len=100, insert=300, err=0.02, frac_random=0.0, paired=False)
    reads.append([b"ACGT", genomes[0][:150]])
    reads.append([genomes[0][:k + 1]])                                 # two k-mers: fewer than -B
    reads.append([genomes[0][:k]])
    reads.append([b"N" * 150, b"N" * 150])
    reads.append([genomes[1][:150], genomes[1][:150]])
    for kw in (dict(), dict(start_sample=1), dict(start_sample=7), dict(start_sample=15), dict(d=3), dict(start_sample=0)):
        res = {}
        for part in (0, 2):
            ctx.set_option("readid_vote_part", part)
            ctx.set_option("readid_part_shift", 14)                    # 16,384-row windows (the plan widens them until P <= 16)
            try:
                if part == 2:
                    _readid_compare(oracle, oix, gix, reads, **kw)
                res[part] = gix.read_id_batch(reads, d=kw.get("d", 1), start_sample=kw.get("start_sample", 3))
            finally:
                ctx.set_option("readid_vote_part", 1)
                ctx.set_option("readid_part_shift", 0)
        _reports_equal(res[0], res[2])
        if kw.get("start_sample", 3) == 3 and clades <= 3:
            # the clades are wide enough for reads with more than 8 candidates (left to the one-kernel vote)
            assert (res[2]["rep_n"] > 9).sum() > 10
    # a bucket that overflows: the whole chunk is redone by the one-kernel vote, same reports
    ctx.set_option("readid_vote_part", 2)
    ctx.set_option("readid_part_shift", 14)
    ctx.set_option("readid_part_cap", 64)
    try:
        over = gix.read_id_batch(reads)
    finally:
        ctx.set_option("readid_vote_part", 1)
        ctx.set_option("readid_part_shift", 0)
        ctx.set_option("readid_part_cap", 0)
    ctx.set_option("readid_vote_part", 0)
    try:
        base = gix.read_id_batch(reads)
    finally:
        ctx.set_option("readid_vote_part", 1)
    _reports_equal(base, over)


def test_read_id_partitioned_vote_classify_and_packed(oracle, ctx):
    """The same path under cid_read_id_classify (ASCII + qualities masked on the device) and cid_read_id_classify_packed (planes
    from the host packer), several pipeline chunks in flight: classifications and top accessions against the oracle."""
    rng = _rng(9200)
    N, k, S, H = 46, 31, 2_000_003, 4
    genomes, oix, gix = _index_pair(oracle, ctx, rng, N, k, S, H)
    gix.n_ref[:] = oix.n_ref
    reads = synth.reads_from(rng, genomes, 3000, read_len=150, insert=320, err=0.004, frac_random=0.2, n_rate=0.002)
    quals = [[bytes(rng.choice(np.frombuffer(b"#+5?I", dtype=np.uint8), size=len(m), p=[.01, .01, .02, .36, .6]).tolist())
              for m in r] for r in reads]
    masked = [[oracle.qual_mask(m, q, 15) for m, q in zip(r, qr)] for r, qr in zip(reads, quals)]
    o = oix.read_id_batch(masked)
    pk = cb.pack_reads(reads, quals, 15)
    ctx.set_option("readid_vote_part", 2)
    ctx.set_option("readid_part_shift", 15)
    ctx.set_option("readid_chunk_reads", 700)
    try:
        g = gix.read_id_classify(reads, quals=quals, qual_offset=15)
        gp = gix.read_id_classify_packed(pk)
    finally:
        ctx.set_option("readid_vote_part", 1)
        ctx.set_option("readid_part_shift", 0)
        ctx.set_option("readid_chunk_reads", 0)
    for res in (g, gp):
        for key in ("kind", "hits", "n_set", "n_top"):
            assert np.array_equal(res[key], o[key]), key
        for r in range(len(reads)):
            nt = min(int(o["n_top"][r]), 8)
            assert res["top"][r, :nt].tolist() == o["top"][r, :nt].tolist()


def test_read_id_partitioned_vote_long_sets(oracle, ctx):
    """The partitioned vote on reads of 2 x 250 and 2 x 480 bases: more than 255 distinct k-mers per read, i.e. the general
    order kernel (16-bit entries, no compact arrays) feeds the scan kernel; plus a few reads beyond the fast path (general
    CTA-per-read path in the same batch)."""
    rng = _rng(9300)
    N, k, S, H = 40, 31, 1_500_007, 4
    genomes = synth.clade_genomes(rng, N, 9000, n_clades=5, div=0.005)
    oix, gix = build_both(oracle, ctx, [[g] for g in genomes], S, H, k, cb.CID_SEQ_FASTA)
    reads = synth.reads_from(rng, genomes, 200, read_len=250, insert=520, err=0.003, frac_random=0.2, n_rate=0.002)
    reads += synth.reads_from(rng, genomes, 60, read_len=480, insert=1000, err=0.002, frac_random=0.1)
    reads += synth.reads_from(rng, genomes, 4, read_len=1500, insert=3200, err=0.001, frac_random=0.0, paired=False)
    res = {}
    for part in (0, 2):
        ctx.set_option("readid_vote_part", part)
        ctx.set_option("readid_part_shift", 15)
        try:
            if part == 2:
                _readid_compare(oracle, oix, gix, reads, order_cap=3200)
            res[part] = gix.read_id_batch(reads)
        finally:
            ctx.set_option("readid_vote_part", 1)
            ctx.set_option("readid_part_shift", 0)
    _reports_equal(res[0], res[2])
    assert int(res[2]["n_set"].max()) > 255
