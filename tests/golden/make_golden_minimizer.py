#!/usr/bin/env python
"""Generates tests/golden/minimizer.json: known answers for the minimizer (.mxi) path.

INDEPENDENT of the C++ oracle: a pure-Python restatement of kmer.rs:971-986 find_minimizer, kmer.rs:328-394
minimerize_vector_skip_n[_set] and build.rs:258-492 build_single_mini / build_multi_mini (hash = python-xxhash), run on
random sequences and on the four phage genomes of the reference checkout (test.sh:3 parameters plus -m -v 15).  The
Rust reference cannot be executed here, so these are not outputs of the colorid binary; the test compares the oracle
with them (two restatements agreeing).  Run:  python tests/golden/make_golden_minimizer.py
"""
import hashlib
import json
import os
import random

import xxhash

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
COMP = {65: 84, 67: 71, 71: 67, 84: 65, 97: 116, 99: 103, 103: 99, 116: 97, 85: 65, 117: 97, 78: 78, 110: 110}


def revcomp(s):                        # kmer.rs:839-863
    return bytes(COMP.get(c, 78) for c in reversed(s))


def find_minimizer(seq, m):            # kmer.rs:971-986
    r = revcomp(seq)
    n = len(seq)
    best = seq[:m]
    for i in range(1, n - m + 1):
        f, c = seq[i:i + m], r[n - (i + m):n - i]
        if f < best:
            best = f
        if c < best:
            best = c
    return best


GOOD = set(b"ACGTacgt")


def canonical_kmers(l, k, d=1):        # the loop shared by kmer.rs:328-394 (len guard, has_no_n, raw-byte compare)
    if len(l) < k:
        return
    r = revcomp(l)
    n = len(l)
    for i in range(0, n - k + 1, d):
        f = l[i:i + k]
        if not all(c in GOOD for c in f):
            continue
        c = r[n - (i + k):n - i]
        yield f if f < c else c


def read_fasta(path):                  # kmer.rs:10-45
    lines = open(path, "rb").read().split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    vec, sub = [], b""
    for i, line in enumerate(lines, 1):
        line = line.rstrip(b"\r")
        if b">" in line:
            if sub:
                vec.append(sub)
            sub = b""
        elif i == len(lines):
            sub += line
            if sub:
                vec.append(sub)
        else:
            sub += line
    return vec


def bits_of(items, S, H):
    return {xxhash.xxh3_64_intdigest(it, seed=s) % S for it in items for s in range(H)}


def main():
    rnd = random.Random(20261017)
    vectors = []
    for _ in range(300):
        k = rnd.randint(3, 31)
        m = rnd.randint(1, k)
        alphabet = rnd.choice([b"ACGT", b"ACGT", b"ACGTacgt", b"acgt"])
        s = bytes(rnd.choice(alphabet) for _ in range(k))
        vectors.append({"seq": s.decode(), "m": m, "min": find_minimizer(s, m).decode()})
    # mixed-case read: minimizer set (kmer.rs:363-394) as sorted upper-cased strings
    sets = []
    for _ in range(20):
        k, m, d = rnd.choice([(31, 15, 1), (27, 15, 1), (21, 9, 3), (21, 21, 1)])
        mates = [bytes(rnd.choice(b"ACGTACGTACGTacgtN") for _ in range(rnd.choice([150, 100, 40, k, k - 1]))) for _ in range(2)]
        st = set()
        for l in mates:
            for km in canonical_kmers(l, k, d):
                st.add(find_minimizer(km, m).upper())
        sets.append({"k": k, "m": m, "d": d, "mates": [x.decode() for x in mates], "set": sorted(x.decode() for x in st)})

    phage = None
    names = ["Listeria_phage_B021", "Listeria_phage_B051", "Listeria_phage_B056", "Listeria_phage_B545"]
    if os.path.isdir(REF):
        k, m, S, H = 27, 15, 750000, 4
        single_rows, multi_rows = {}, {}
        n_single, n_multi = [], []
        for c, nm in enumerate(names):
            contigs = read_fasta(os.path.join(REF, "test_data/refs", nm + ".fasta"))
            kmers, minis = set(), {}
            for l in contigs:
                for km in canonical_kmers(l, k):
                    kmers.add(km.upper())                                         # kmerize_vector: then uppercase
                    mi = find_minimizer(km, m).upper()                            # minimerize_vector_skip_n
                    minis[mi] = minis.get(mi, 0) + 1
            n_single.append(len(kmers))
            n_multi.append(len(minis))
            for b in bits_of({find_minimizer(km, m) for km in kmers}, S, H):      # build_single_mini
                single_rows[b] = single_rows.get(b, 0) | (1 << c)
            for b in bits_of(minis.keys(), S, H):                                 # build_multi_mini
                multi_rows[b] = multi_rows.get(b, 0) | (1 << c)

        def digest(rows):
            h = hashlib.sha256()
            for r in sorted(rows):
                h.update(r.to_bytes(8, "little") + rows[r].to_bytes(4, "little"))
            return h.hexdigest()
        phage = {"params": {"k": k, "m": m, "S": S, "H": H}, "colours": names,
                 "single": {"n_ref_kmers": n_single, "nonzero_rows": len(single_rows), "sha256_rows": digest(single_rows)},
                 "multi": {"n_ref_kmers": n_multi, "nonzero_rows": len(multi_rows), "sha256_rows": digest(multi_rows)}}
    json.dump({"source": "pure-Python restatement + python-xxhash %s (tests/golden/make_golden_minimizer.py)" % xxhash.VERSION,
               "find_minimizer": vectors, "read_sets": sets, "phage": phage},
              open(os.path.join(OUT, "minimizer.json"), "w"), indent=0)
    print("wrote minimizer.json", phage)


if __name__ == "__main__":
    main()
