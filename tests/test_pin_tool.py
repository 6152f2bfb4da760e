"""tools/pin_from_bxi.py (a third, numpy-only restatement of read_fasta / kmerize_vector / BloomFilter::insert with the 32
hash variants) must name the variant an index was built with.  Here the index is the oracle's, written by the
independent Python .bxi writer, over the reference's own test.sh accessions (tests/golden/phage)."""
import gzip
import importlib.util
import os

import numpy as np
import pytest

from tests import bxi_py

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("pin_from_bxi", os.path.join(ROOT, "tools", "pin_from_bxi.py"))
pin = importlib.util.module_from_spec(spec)
spec.loader.exec_module(pin)
PHAGES = ["Listeria_phage_B021", "Listeria_phage_B051", "Listeria_phage_B056", "Listeria_phage_B545"]


def stage_phages(d):
    lines = []
    for n in PHAGES:
        with gzip.open(os.path.join(ROOT, "tests", "golden", "phage", n + ".fasta.gz"), "rb") as f:
            (d / (n + ".fasta")).write_bytes(f.read())
        lines.append(f"{n}\t{d}/{n}.fasta")
    (d / "ref_file.txt").write_text("\n".join(lines) + "\n")
    return d / "ref_file.txt"


def test_numpy_xxh3_equals_oracle_for_every_variant(oracle):
    rng = np.random.default_rng(5)
    for k in (1, 3, 4, 8, 9, 15, 16, 17, 21, 27, 31, 32):
        km = rng.integers(65, 123, size=(20, k)).astype(np.uint8)
        for v in range(32):
            for seed in (0, 3):
                got = pin.xxh3_64(km, seed, v)
                exp = np.array([oracle.xxh3_64(bytes(r), seed, v) for r in km], dtype=np.uint64)
                assert np.array_equal(got, exp), (k, v, seed)
    # the variants are pairwise different functions on k-mers of 17+ bytes
    km = rng.integers(65, 85, size=(4, 27)).astype(np.uint8)
    assert len({tuple(pin.xxh3_64(km, 1, v).tolist()) for v in range(32)}) == 32


@pytest.mark.parametrize("variant", [0, 1, 31, 6])
def test_pin_tool_names_the_variant_of_an_oracle_built_index(oracle, tmp_path, variant):
    ref_file = stage_phages(tmp_path)
    seqs = [oracle.read_fasta(str(tmp_path / (n + ".fasta"))) for n in PHAGES]
    oix = oracle.Index(750_000, 4, 27, 4, hash_variant=variant)
    for c, s in enumerate(seqs):
        oix.build_accession(c, s, oracle.MODE_FASTA)
    oix.finalize(threads=2)
    dense = oix.words()
    nz = np.flatnonzero(dense.any(axis=1))
    if variant == 0:
        assert len(nz) == 324_869            # SURVEY D.2 (stable XXH3)
    bxi_py.write_bxi(tmp_path / "phage.bxi", 750_000, 4, 27, dict(enumerate(PHAGES)), nz, dense[nz],
                     {n: int(oix.n_ref[c]) for c, n in enumerate(PHAGES)}, row_order=np.random.default_rng(1).permutation(len(nz)))
    assert pin.pin(tmp_path / "phage.bxi", ref_file) == [variant]
