"""CPU: the oracle against the golden vectors (third-party arithmetic pins + SURVEY Appendix D)."""
import json
import os

import numpy as np
import pytest

from tests.conftest import REFERENCE, needs_reference

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HV = json.load(open(os.path.join(GOLD, "hash_vectors.json")))
AD = json.load(open(os.path.join(GOLD, "appendix_d.json")))


def test_xxh3_matches_libxxhash_vectors(oracle):
    for v in HV["vectors"]:
        assert oracle.xxh3_64(bytes.fromhex(v["hex"]), v["seed"]) == v["xxh3_64"]
    for k, exp in HV["named"].items():
        assert [oracle.xxh3_64(k.encode(), s) for s in range(4)] == exp
    # the 31-mer's Bloom bits at S = 50M
    k31 = b"ACGTACGTACGTACGTACGTACGTACGTACG"
    assert [oracle.xxh3_64(k31, s) % 50_000_000 for s in range(4)] == [42825530, 37371628, 5173840, 42473127]


def test_xxh3_matches_python_xxhash_live(oracle):
    xxhash = pytest.importorskip("xxhash")
    rng = np.random.default_rng(1)
    for n in range(0, 241, 3):
        b = bytes(rng.integers(0, 256, size=n, dtype=np.uint8).tolist())
        for seed in (0, 3, 2**64 - 1):
            assert oracle.xxh3_64(b, seed) == xxhash.xxh3_64_intdigest(b, seed=seed)


def test_fnv1a_vectors(oracle):
    for k, v in HV["fnv1a_str"].items():
        assert oracle.fnv1a_str(k.encode()) == v
    for k, v in HV["fnv1a_usize"].items():
        assert oracle.fnv1a_usize(int(k)) == v


def test_seq_rs_unit_tests(oracle):
    # seq.rs:72-76 can_detect_no_n
    assert oracle.has_no_n(b"AAGT") and not oracle.has_no_n(b"NAGT")
    assert oracle.has_no_n(b"acgtACGT") and not oracle.has_no_n(b"ACGU")
    # kmer.rs:847-863 switch_base
    assert oracle.revcomp(b"ACGTacgtUuNnXx-") == b"NNNnNaAacgtACGT"[::1]


def test_revcomp_table(oracle):
    assert oracle.revcomp(b"AACG") == b"CGTT"
    assert oracle.revcomp(b"u") == b"a" and oracle.revcomp(b"U") == b"A" and oracle.revcomp(b"R") == b"N"


def test_qual_mask(oracle):
    # seq.rs:36-56: Q==0 copies; else base -> 'N' where qual < Q+33; output length == len(qual)
    assert oracle.qual_mask(b"ACGT", b"IIII", 0) == b"ACGT"
    assert oracle.qual_mask(b"ACGT", b"I#I0", 15) == b"ANGT"      # '#'=35 < 48 ; '0'=48 is kept
    assert oracle.qual_mask(b"ACGT", b"I/", 15) == b"AN"          # shorter qual truncates
    assert oracle.qual_mask(b"AC", b"IIII", 15) is None           # reference panics


def test_simple_bloom_unit_test(oracle):
    # simple_bloom.rs:58-66 use_filter on a 250M-bit / 4-hash filter: insert ATGC, contains ATGC, not ATGT
    S, H = 250_000_000, 4
    bits = {oracle.xxh3_64(b"ATGC", i) % S for i in range(H)}
    assert all((oracle.xxh3_64(b"ATGC", i) % S) in bits for i in range(H))
    assert not all((oracle.xxh3_64(b"ATGT", i) % S) in bits for i in range(H))


def test_auto_cutoff_appendix_d3(oracle):
    for histo, exp in AD["D3"]:
        assert oracle.auto_cutoff_histo({int(k): v for k, v in histo.items()}) == exp
    assert oracle.auto_cutoff_histo({2: 10}) == -1                 # coverages.len()-1 underflow -> panic
    assert oracle.auto_cutoff_histo({1: 1, 3: 10}) == -1           # d1.len()-1 underflow -> panic


def test_bitvec_layout(oracle):
    # bit-vec_serde/src/lib.rs:465-500: bit i lives in storage[i/32] at mask 1 << (i%32)
    ix = oracle.Index(64, 1, 5, 40)
    ix.build_accession(33, [b"ACGTACGTAC"], oracle.MODE_FASTA)
    ix.finalize()
    w = ix.words()
    assert w.shape == (64, 2)
    assert w[:, 0].sum() == 0 and set(np.unique(w[:, 1])) <= {0, 2}


def test_hashbrown_growth_ladder(oracle):
    # Appendix C: buckets/capacity 4/3, 8/7, 16/14, 32/28, 64/56, 128/112, 256/224, 512/448
    keys = [b"K%05d" % i for i in range(449)]
    exp = {1: 4, 3: 4, 4: 8, 7: 8, 8: 16, 14: 16, 15: 32, 28: 32, 29: 64, 56: 64, 57: 128, 112: 128, 113: 256, 224: 256,
           225: 512, 448: 512, 449: 1024}
    for n, b in exp.items():
        order, buckets = oracle.hashset_str_order(keys[:n])
        assert buckets == b and sorted(order.tolist()) == list(range(n))
    # a duplicate arriving at a full table: modern rule grows, old rule does not
    _, b_new = oracle.hashset_str_order(keys[:224] + [keys[0]], reserve_before_find=True)
    _, b_old = oracle.hashset_str_order(keys[:224] + [keys[0]], reserve_before_find=False)
    assert (b_new, b_old) == (512, 256)


def test_hashbrown_small_table_is_cyclic_linear_probe(oracle):
    # 4 buckets, usize keys: FNV(1) & 3 == 0, FNV(2) & 3 == 3 (SURVEY D.5 tie order: key 1 before key 2)
    assert oracle.fnv1a_usize(1) & 3 == 0 and oracle.fnv1a_usize(2) & 3 == 3
    assert oracle.hashmap_usize_order([2, 1]).tolist() == [1, 0]
    assert oracle.hashmap_usize_order([1, 4, 2]).tolist() == [0, 1, 2]


def test_read_pair_order_and_votes_from_golden_without_reference(oracle):
    # the k-mer iteration order only needs k (no index rows): pin it for all four pairs x both growth rules
    ix = oracle.Index(1024, 4, 27, 4)
    for name, g in AD["D45"].items():
        rbf = name.endswith("rbf=1")
        res = ix.read_id_batch([[m.encode() for m in g["mates"]]], order_cap=512, reserve_before_find=rbf)
        n = int(res["order_n"][0])
        assert n == g["n_set"] == int(res["n_set"][0])
        assert res["order_seq"][0][:n].tolist() == g["order_seq"]
        assert res["order_pos"][0][:n].tolist() == g["order_pos"]


@needs_reference
def test_appendix_d1_n_ref_kmers(oracle):
    for row in AD["D1"]:
        seqs = oracle.read_fasta(os.path.join(REFERENCE, row["file"]))
        assert (len(seqs), sum(map(len, seqs))) == (row["contigs"], row["bases"])
        m = oracle.KMap(row["k"])
        m.add(seqs, oracle.MODE_FASTA)
        assert len(m) == row["n_ref_kmers"]


@needs_reference
def test_appendix_d2_phage_index_and_d5_reads(oracle):
    d2 = AD["D2"]
    seqs = [oracle.read_fasta(os.path.join(REFERENCE, "test_data/refs", n + ".fasta")) for n in d2["colours"]]
    one = oracle.Index(750000, 4, 27, 4)
    for c, s in enumerate(seqs):
        one.build_accession(c, s, oracle.MODE_FASTA)
    one.finalize(threads=1)
    multi = oracle.Index(750000, 4, 27, 4)
    multi.build_many(seqs, oracle.MODE_FASTA, threads=2)          # test.sh:3 vs test.sh:11 (-t 2): same index
    assert np.array_equal(one.words(), multi.words())
    assert one.n_ref.tolist() == d2["n_ref_kmers"] == [31297, 27583, 32634, 27491]
    assert one.nonzero_rows() == d2["nonzero_rows"] == 324869
    r = one.query_counts(seqs, oracle.MODE_FASTA, False, 0)
    assert r["counts"].tolist() == d2["self_query_counts"]
    assert r["uniq_n"].sum(axis=1).tolist() == d2["self_query_unique"] == [27148, 19794, 26585, 20675]
    # own accession always scores count == n_ref_kmers (no false negatives in a Bloom filter)
    assert np.array_equal(np.diag(r["counts"]), one.n_ref)
    # -s on a sub-sequence of an indexed genome lists that accession
    p = one.query_perfect([[seqs[1][0][100:600]]])
    assert p["status"][0] == 0 and (p["and_rows"][0, 0] >> 1) & 1
    for name, g in AD["D45"].items():
        rbf = name.endswith("rbf=1")
        res = one.read_id_batch([[m.encode() for m in g["mates"]]], reserve_before_find=rbf)
        rep = [[int(c), int(v)] for c, v in zip(res["rep_colour"][0][:res["rep_n"][0]], res["rep_count"][0][:res["rep_n"][0]])]
        assert rep == g["report"]
        assert (int(res["kind"][0]), int(res["hits"][0]), int(res["n_top"][0])) == (g["kind"], g["hits"], g["n_top"])
    # SURVEY D.5 table
    assert AD["D45"]["clean/rbf=1"]["report"] == [[1, 248], [2, 248]] and AD["D45"]["clean/rbf=1"]["top"] == [1, 2]
    assert sorted(AD["D45"]["mutated/rbf=1"]["report"]) == [[1, 1], [2, 1], [4, 1]] and AD["D45"]["mutated/rbf=1"]["kind"] == 2
    assert AD["D45"]["random/rbf=1"]["report"] == [[4, 1]] and AD["D45"]["random/rbf=1"]["kind"] == 1


@needs_reference
def test_fasta_reader_quirks(oracle, tmp_path):
    # kmer.rs:10-45: header = any line containing '>'; contigs concatenated; last line handling
    p = tmp_path / "x.fa"
    p.write_bytes(b">a\nACGT\nAC\n>b desc\nGG\r\nTT")
    assert oracle.read_fasta(str(p)) == [b"ACGTAC", b"GGTT"]
    p.write_bytes(b"ACGT\n>late header\n")
    assert oracle.read_fasta(str(p)) == [b"ACGT"]
    labels, seqs = oracle.read_fasta_mf(str(tmp_path / "x.fa"))
    assert labels == [b"late header"] and seqs == [b"ACGT"]


# ---- minimizer (.mxi) path: fixtures from an independent pure-Python restatement (tests/golden/make_golden_minimizer.py)
MZ = json.load(open(os.path.join(GOLD, "minimizer.json")))


def test_find_minimizer_vectors(oracle):
    for v in MZ["find_minimizer"]:
        assert oracle.find_minimizer(v["seq"].encode(), v["m"]) == v["min"].encode(), v
    # kmer.rs:974-985: the reverse complement of the FIRST m-mer is never a candidate
    assert oracle.find_minimizer(b"TTTTC", 4) == b"GAAA"       # candidates TTTT, TTTC, GAAA (not AAAA)
    assert oracle.find_minimizer(b"ACGT", 4) == b"ACGT"        # m == k: the k-mer itself
    assert oracle.find_minimizer(b"ACG", 4) is None            # m > len: the reference panics


def test_read_minimizer_sets(oracle):
    for g in MZ["read_sets"]:
        ix = oracle.Index(1024, 2, g["k"], 2, m=g["m"])
        mates = [x.encode() for x in g["mates"]]
        if len(mates[0]) < g["k"]:
            continue                                            # parallel_vec: too_short before the set is built
        res = ix.read_id_batch([mates], d=g["d"], order_cap=512)
        n = int(res["order_n"][0])
        assert n == int(res["n_set"][0]) == len(g["set"])
        got = set()
        for i in range(n):
            sq, pos = int(res["order_seq"][0][i]), int(res["order_pos"][0][i])
            w = mates[sq & 0x7F][pos:pos + g["m"]]
            got.add((w if sq & 0x80 else oracle.revcomp(w)).upper().decode())
        assert sorted(got) == g["set"]


@needs_reference
def test_phage_minimizer_index_both_builders(oracle):
    import hashlib
    ph = MZ["phage"]
    P = ph["params"]
    seqs = [oracle.read_fasta(os.path.join(REFERENCE, "test_data/refs", n + ".fasta")) for n in ph["colours"]]
    for variant, key in ((0, "single"), (1, "multi")):
        ix = oracle.Index(P["S"], P["H"], P["k"], 4, m=P["m"])
        n_ref = [ix.build_accession_mini(c, s, oracle.MODE_FASTA, -1, variant)[0] for c, s in enumerate(seqs)]
        ix.finalize(threads=2)
        assert n_ref == ph[key]["n_ref_kmers"]
        w = ix.words()
        nz = np.flatnonzero(w.any(axis=1))
        assert len(nz) == ph[key]["nonzero_rows"]
        h = hashlib.sha256()
        for r in nz:
            h.update(int(r).to_bytes(8, "little") + int(w[r, 0]).to_bytes(4, "little"))
        assert h.hexdigest() == ph[key]["sha256_rows"]
    # minimizer counts feed clean_map in build_multi_mini: cutoff 1 drops singletons
    km = oracle.minimizer_map(seqs[0], P["k"], P["m"])
    total = len(km)
    km.clean(1)
    ix = oracle.Index(P["S"], P["H"], P["k"], 4, m=P["m"])
    assert ix.build_accession_mini(0, seqs[0], oracle.MODE_FASTA, 1, 1) == (len(km), 1) and 0 < len(km) < total
