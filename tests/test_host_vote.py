"""CPU: the product's host-side vote (cid_classify_reads: kmer_poll_plus, false_prob, Binomial pmf,
FnvHashMap tie order) against the oracle.  Host logic only: no CUDA call is made."""
import numpy as np
import pytest

import colorid_b200.lib as L
from colorid_b200.api import classify_reads


def test_false_prob_and_binomial_equal_oracle(oracle):
    lib = L.load()
    rng = np.random.default_rng(3)
    for _ in range(300):
        m, k, n = int(rng.integers(1000, 10**8)), int(rng.integers(1, 8)), int(rng.integers(0, 10**7))
        assert lib.cid_false_prob(m, k, n) == oracle.false_prob(m, k, n)
    for _ in range(3000):
        n = int(rng.integers(1, 600))
        x = int(rng.integers(0, n + 1))
        p = float(10 ** rng.uniform(-6, -0.01))
        assert lib.cid_binomial_mass(n, p, x) == oracle.binomial_mass(n, p, x)
    assert lib.cid_binomial_mass(10, 0.0, 0) == 1.0 and lib.cid_binomial_mass(10, 1.0, 10) == 1.0
    # sanity: it is a pmf
    tot = sum(lib.cid_binomial_mass(40, 0.3, x) for x in range(41))
    assert abs(tot - 1.0) < 1e-12


def test_classify_reads_equals_oracle_kmer_poll_plus(oracle):
    rng = np.random.default_rng(11)
    S, H, N = 50_000_000, 4, 46
    n_ref = rng.integers(1_500_000, 5_000_000, size=N).astype(np.uint64)
    fp = np.array([oracle.false_prob(S, H, int(v)) for v in n_ref])
    nr, cap = 4000, N + 1
    n_set = rng.integers(1, 241, size=nr).astype(np.uint32)
    flags = np.zeros(nr, np.uint32)
    rep_n = np.zeros(nr, np.uint32)
    rc = np.zeros((nr, cap), np.uint32)
    rv = np.zeros((nr, cap), np.uint32)
    for r in range(nr):
        kind = rng.integers(0, 6)
        if kind == 0:
            cols = []
        elif kind == 1:
            cols = [N]
        else:
            cols = rng.permutation(N)[: rng.integers(1, 9)].tolist()
            if rng.random() < 0.5:
                cols.append(N)
        rep_n[r] = len(cols)
        for i, c in enumerate(cols):
            rc[r, i] = c
            # many ties and a few tiny counts to hit every branch
            rv[r, i] = 1 if c == N else int(rng.choice([1, 2, n_set[r], max(1, n_set[r] // 2), max(1, n_set[r] // 2)]))
            rv[r, i] = min(rv[r, i], n_set[r])
    flags[:5] = 1
    flags[5:8] = 2
    got = classify_reads((S, H, N), n_ref, dict(n_set=n_set, flags=flags, rep_n=rep_n, rep_colour=rc, rep_count=rv),
                         1e-3, 16, threads=3, top_cap=16)
    for r in range(nr):
        if flags[r] & 1:
            assert got["kind"][r] == L.load().cid_version() * 0 + 0        # CID_CLS_TOO_SHORT
            continue
        if flags[r] & 2:
            assert got["kind"][r] == 5
            continue
        if rep_n[r] == 0:
            assert got["kind"][r] == 1
            continue
        keys = rc[r, :rep_n[r]].astype(np.uint64)
        order = oracle.hashmap_usize_order(keys)          # insertion order -> FnvHashMap iteration order
        kind, hits, n_top, top = oracle.kmer_poll_plus(rc[r, :rep_n[r]][order], rv[r, :rep_n[r]][order], int(n_set[r]), fp,
                                                       1e-3, top_cap=16)
        assert (got["kind"][r], got["hits"][r], got["n_top"][r]) == (kind, hits, n_top), r
        assert got["top"][r, :n_top].tolist() == top.tolist(), r


def test_merge_shard_reports_restores_insertion_order():
    """cid_merge_shard_reports (host only): per-shard read_id reports with insertion steps -> the unsharded report."""
    import colorid_b200 as cb
    from colorid_b200.api import merge_shard_reports
    rng = np.random.default_rng(4242)
    n_total, shards = 300, [(0, 96), (96, 224), (224, 300)]
    nreads, cap = 200, 40
    glob_c = np.zeros((nreads, n_total + 1), np.uint32)
    glob_v = np.zeros((nreads, n_total + 1), np.uint32)
    glob_n = np.zeros(nreads, np.uint32)
    reps = [dict(n_set=np.full(nreads, 50, np.uint32), flags=np.zeros(nreads, np.uint32), rep_n=np.zeros(nreads, np.uint32),
                 rep_colour=np.zeros((nreads, cap), np.uint32), rep_count=np.zeros((nreads, cap), np.uint32)) for _ in shards]
    for r in range(nreads):
        ncol = int(rng.integers(0, 25))
        cols = rng.choice(n_total, size=ncol, replace=False)
        steps = np.sort(rng.integers(0, 3, size=ncol))                 # colours arrive at k-mers 0..2, ascending within a k-mer
        seq = sorted(zip(steps.tolist(), cols.tolist()))
        miss = bool(rng.integers(0, 2))
        for i, (st, c) in enumerate(seq):
            glob_c[r, i], glob_v[r, i] = c, int(rng.integers(1, 200))
        n = len(seq)
        if miss:
            glob_c[r, n], glob_v[r, n] = n_total, 1
            n += 1
        glob_n[r] = n
        for s, (lo, hi) in enumerate(shards):
            k = 0
            for i, (st, c) in enumerate(seq):
                if lo <= c < hi:
                    reps[s]["rep_colour"][r, k] = (c - lo) | (st << 20)
                    reps[s]["rep_count"][r, k] = glob_v[r, i]
                    k += 1
            if miss:
                reps[s]["rep_colour"][r, k], reps[s]["rep_count"][r, k] = hi - lo, 1
                k += 1
            reps[s]["rep_n"][r] = k
    m = merge_shard_reports(reps, shards, n_total)
    assert np.array_equal(m["rep_n"], glob_n)
    for r in range(nreads):
        n = glob_n[r]
        assert np.array_equal(m["rep_colour"][r, :n], glob_c[r, :n]) and np.array_equal(m["rep_count"][r, :n], glob_v[r, :n])
    # shards that disagree on the first absent row (row-present bitmaps not merged) are refused
    reps[1]["rep_n"][:] = 0
    r_miss = int(np.argmax(glob_c[np.arange(nreads), np.maximum(glob_n, 1) - 1] == n_total))
    if glob_n[r_miss] and glob_c[r_miss, glob_n[r_miss] - 1] == n_total:
        with pytest.raises(cb.CidError):
            merge_shard_reports(reps, shards, n_total)
    # a small output capacity truncates and says so
    m2 = merge_shard_reports([dict(x, rep_n=x["rep_n"].copy()) for x in reps[:1]], shards[:1], 96, rep_cap=2)
    assert (m2["rep_n"] <= 2).all() and ((m2["flags"] & 4) != 0).sum() == (reps[0]["rep_n"] > 2).sum()
