"""CPU: the product's host-side vote (cid_classify_reads: kmer_poll_plus, false_prob, Binomial pmf,
FnvHashMap tie order) against the oracle.  Host logic only: no CUDA call is made."""
import numpy as np

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
