#!/usr/bin/env python
"""Generates tests/golden/*.json from the reference checkout (/root/reference) with the CPU oracle.

The Rust reference cannot be executed here, so these are NOT outputs of the colorid binary: they are
(a) the known-answer values of SURVEY.md Appendix D, which were derived by an independent throw-away
Python restatement, re-derived here with the C++ oracle (two restatements agreeing), and
(b) vectors of third-party arithmetic (XXH3 via python-xxhash, FNV-1a) pinned to published algorithms.
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np
import xxhash

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20261017)
    # ---- hash vectors: python-xxhash (libxxhash 0.8.2) is the pin for XXH3_64bits_withSeed
    vec = []
    for n in list(range(0, 34)) + [48, 64, 65, 96, 97, 128, 129, 200, 240]:
        for seed in (0, 1, 2, 3, 0x9E3779B185EBCA87):
            b = bytes(rng.integers(65, 85, size=n, dtype=np.uint8).tolist())
            vec.append({"hex": b.hex(), "seed": seed, "xxh3_64": xxhash.xxh3_64_intdigest(b, seed=seed)})
    kmers = ["ATGC", "ACGTACGTACGTACGTACGTACGTACGTACG"]
    named = {k: [xxhash.xxh3_64_intdigest(k.encode(), seed=s) for s in range(4)] for k in kmers}
    json.dump({"source": "python-xxhash %s / libxxhash %s" % (xxhash.VERSION, xxhash.XXHASH_VERSION), "vectors": vec,
               "named": named,
               "fnv1a_str": {"ATGC": 0xecb8d39dd5df523f, "ACGTACGTACGTACGTACGTACGTACGTACG": 0x08a89ae8cf8f0d74},
               "fnv1a_usize": {"0": 0xa8c7f832281a39c5, "1": 0x89cd31291d2aefa4, "45": 0x265f6c2beafed408}},
              open(os.path.join(OUT, "hash_vectors.json"), "w"), indent=0)

    # ---- Appendix D.1: n_ref_kmers of bundled FASTAs (hash independent)
    d1 = []
    for f, k in [("test_data/refs/Listeria_phage_B021.fasta", 27), ("test_data/refs/Listeria_phage_B051.fasta", 27),
                 ("test_data/refs/Listeria_phage_B056.fasta", 27), ("test_data/refs/Listeria_phage_B545.fasta", 27),
                 ("refs/Campylobacter_jejuni_CJ677CC524.fasta", 31), ("refs/Listeria_marthii_S4_120.fasta", 31),
                 ("refs/Staphylococcus_aureus_NCTC8532.fasta", 31), ("refs/SRR2167842_ST0.fasta", 31),
                 ("refs/LmonoEGDe.fasta", 31)]:
        seqs = O.read_fasta(os.path.join(REF, f))
        m = O.KMap(k)
        m.add(seqs, O.MODE_FASTA)
        d1.append({"file": f, "k": k, "contigs": len(seqs), "bases": sum(map(len, seqs)), "n_ref_kmers": len(m)})

    # ---- Appendix D.2: phage index of test.sh:3 (-k 27 -s 750000 -n 4)
    names = ["Listeria_phage_B021", "Listeria_phage_B051", "Listeria_phage_B056", "Listeria_phage_B545"]
    seqs = [O.read_fasta(os.path.join(REF, "test_data/refs", n + ".fasta")) for n in names]
    ix = O.Index(750000, 4, 27, 4)
    ix.build_many(seqs, O.MODE_FASTA, threads=4)
    r = ix.query_counts(seqs, O.MODE_FASTA, False, 0)
    words = ix.words()
    d2 = {"params": {"k": 27, "S": 750000, "H": 4}, "colours": names, "n_ref_kmers": ix.n_ref.tolist(),
          "nonzero_rows": int(ix.nonzero_rows()), "self_query_counts": r["counts"].tolist(),
          "self_query_unique": r["uniq_n"].sum(axis=1).tolist(),
          "matrix_checksum_xxh3": xxhash.xxh3_64_intdigest(words.tobytes())}

    # ---- Appendix D.4 / D.5: read pairs cut from contig 1 of phage B056
    c1 = seqs[2][0]
    m1, m2 = c1[0:150], O.revcomp(c1[200:350])
    m2b = O.revcomp(c1[100:250])
    a = bytearray(m1); a[106] = ord("C")
    b = bytearray(m2); b[46] = ord("T")
    rnd = [bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 150).tolist()) for _ in range(2)]
    pairs = {"clean": [m1, m2], "overlap": [m1, m2b], "mutated": [bytes(a), bytes(b)], "random": rnd}
    d45 = {}
    for name, pr in pairs.items():
        for rbf in (True, False):
            res = ix.read_id_batch([pr], order_cap=512, reserve_before_find=rbf)
            n = int(res["order_n"][0])
            d45["%s/rbf=%d" % (name, rbf)] = {
                "mates": [pr[0].decode(), pr[1].decode()], "n_set": int(res["n_set"][0]),
                "order_seq": res["order_seq"][0][:n].tolist(), "order_pos": res["order_pos"][0][:n].tolist(),
                "report": [[int(c), int(v)] for c, v in zip(res["rep_colour"][0][:res["rep_n"][0]], res["rep_count"][0][:res["rep_n"][0]])],
                "kind": int(res["kind"][0]), "hits": int(res["hits"][0]), "n_top": int(res["n_top"][0]),
                "top": res["top"][0][:res["n_top"][0]].tolist()}
    # the rows of the index those four pairs touch are not shipped; GPU tests rebuild the index from
    # synthetic data instead.  These vectors pin the ORACLE (CPU) on the machine that has the reference.
    json.dump({"D1": d1, "D2": d2, "D45": d45,
               "D3": [[{"1": 1000, "2": 100, "3": 20, "4": 10, "5": 8, "6": 9, "7": 12, "8": 20, "9": 30, "10": 40, "11": 45,
                        "12": 40, "13": 30, "14": 20, "15": 10, "16": 5, "17": 1}, 4],
                      [{"1": 50, "2": 40, "3": 30, "4": 20, "5": 10, "6": 5}, 1],
                      [{"1": 1000, "2": 100}, 0],
                      [{"1": 500, "2": 50, "3": 5, "6": 2, "10": 40, "11": 60, "12": 40, "13": 3}, 4]]},
              open(os.path.join(OUT, "appendix_d.json"), "w"), indent=0)
    print("wrote golden vectors")


if __name__ == "__main__":
    main()
