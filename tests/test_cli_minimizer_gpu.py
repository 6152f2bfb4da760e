"""GPU tests of the CLI on minimizer indexes: `build -m -v M` (-t 1 = build_single_mini, -t 2 = build_multi_mini,
build.rs:258-492) writing PREFIX.mxi (bigsi.rs:40-49,71-77), `info` and `read_id` on the .mxi, `search` refusing it."""
import gzip

import numpy as np
import pytest

from tests import bxi_py, synth
from tests.test_cli_gpu import CLS_NAMES, _expected_read_lines, fastq_bytes, read_pairs, run, wrap_fasta

pytestmark = pytest.mark.gpu

K, M, S, H = 27, 11, 300_007, 3


@pytest.fixture(scope="module")
def world(tmp_path_factory, oracle):
    d = tmp_path_factory.mktemp("cli_mini")
    rng = np.random.default_rng(0xC0101D99)
    genomes = synth.clade_genomes(rng, 5, 6000, n_clades=2, div=0.02)
    acc, refs = {}, []
    for i in range(4):
        g = genomes[i]
        if i == 1:
            g = synth.sprinkle(rng, g, b"acgt", 0.2)           # mixed case: raw-byte minimizer choice in build_multi_mini
        name = f"Mini_{chr(ord('D') - i)}"
        contigs = [g[:3000], g[3000:]]
        wrap_fasta(d / f"{name}.fasta", [(f"{name}_c{j}", c) for j, c in enumerate(contigs)])
        acc[name] = dict(mode=oracle.MODE_FASTA, seqs=contigs)
        refs.append(f"{name}\t{d}/{name}.fasta")
    (d / "refs_fasta.tsv").write_text("\n".join(refs) + "\n")
    n, s1, s2, q1, q2 = read_pairs(rng, [genomes[4]], 900, "pe", read_len=120, insert=260, err=0.004, frac_random=0.0)
    (d / "pe_1.fastq.gz").write_bytes(gzip.compress(fastq_bytes(n, s1, q1)))
    (d / "pe_2.fastq.gz").write_bytes(gzip.compress(fastq_bytes(n, s2, q2)))
    masked = []
    for a, b, qa, qb in zip(s1, s2, q1, q2):
        masked += [oracle.qual_mask(a, qa.encode(), 20), oracle.qual_mask(b, qb.encode(), 20)]
    acc["Reads_PE"] = dict(mode=oracle.MODE_FASTQ, seqs=masked)
    refs.append(f"Reads_PE\t{d}/pe_1.fastq.gz\t{d}/pe_2.fastq.gz")
    (d / "refs_all.tsv").write_text("\n".join(refs) + "\n")
    return dict(dir=d, genomes=genomes, acc=acc, rng=rng)


def _oracle_index(oracle, acc, names, variant):
    oix = oracle.Index(S, H, K, len(names), m=M)
    for c, name in enumerate(names):
        oix.build_accession_mini(c, acc[name]["seqs"], acc[name]["mode"], -1, variant)
    oix.finalize()
    return oix


@pytest.mark.parametrize("threads,variant", [(1, 0), (2, 1)])
def test_build_writes_the_reference_mxi(world, oracle, threads, variant):
    d, acc = world["dir"], world["acc"]
    names = sorted(acc)
    oix = _oracle_index(oracle, acc, names, variant)
    body, _ = run("build", "-b", d / f"all{threads}", "-r", d / "refs_all.tsv", "-k", K, "-n", H, "-s", S, "-m", "-v", M,
                  "-t", threads, "-Q", 20)
    assert body[-2:] == [f"Build with minimizers, minimizer size: {M}", "Saving BIGSI to file."]
    got = bxi_py.read_bxi(d / f"all{threads}.mxi", mini=True)
    assert (got["bloom_size"], got["num_hash"], got["k_size"], got["m_size"]) == (S, H, K, M)
    assert got["colors"] == dict(enumerate(names))
    exp_ref = {n: int(oix.n_ref[c]) for c, n in enumerate(names)}
    if variant == 0:
        del exp_ref["Reads_PE"]             # build_single_mini records n_ref_kmers for FASTA accessions only (build.rs:450)
    assert got["n_ref"] == exp_ref
    dense = oix.words()
    nz = np.flatnonzero(dense.any(axis=1))
    o = np.argsort(got["row_ids"])
    assert np.array_equal(got["row_ids"][o], nz.astype(np.uint64))
    assert np.array_equal(got["words"][o], dense[nz])
    if variant == 1:
        body, _ = run("info", "-b", d / f"all{threads}.mxi")
        assert body[:6] == ["BIGSI parameters:", f"Bloomfilter-size: {S}", f"Number of hashes: {H}", f"K-mer size: {K}",
                            f" minimizer size: {M}", ""]
        assert body[6] == f"Number of accessions in index: {len(names)}"
        assert body[7:] == [f"{n} {exp_ref[n]} {oracle.false_prob(S, H, exp_ref[n]):.3f}" for n in names]


@pytest.mark.parametrize("threads,variant", [(1, 0), (2, 1)])
def test_read_id_and_search_on_mxi(world, oracle, threads, variant):
    d, rng, genomes, acc = world["dir"], world["rng"], world["genomes"], world["acc"]
    names = sorted(n for n in acc if n != "Reads_PE")
    oix = _oracle_index(oracle, acc, names, variant)
    run("build", "-b", d / f"fa{threads}", "-r", d / "refs_fasta.tsv", "-k", K, "-n", H, "-s", S, "-m", "-v", M, "-t", threads)
    n, s1, s2, q1, q2 = read_pairs(rng, genomes[:4], 2500, "r", read_len=100, insert=180, err=0.006, frac_random=0.2)
    s1[5], q1[5] = s1[5][:15], q1[5][:15]                      # mate 1 shorter than k -> too_short
    s2[7], q2[7] = s2[7][:20], q2[7][:20]                      # mate 2 shorter than k: skipped, no panic (kmer.rs:372)
    (d / "r_1.fastq.gz").write_bytes(gzip.compress(fastq_bytes(n, s1, q1)))
    (d / "r_2.fastq.gz").write_bytes(gzip.compress(fastq_bytes(n, s2, q2)))
    ids = ["@" + x for x in n]
    reads = [[oracle.qual_mask(a, qa.encode(), 15), oracle.qual_mask(b, qb.encode(), 15)] for a, b, qa, qb in zip(s1, s2, q1, q2)]
    w = dict(oix=oix, names=names)
    exp_lines, exp_counts = _expected_read_lines(w, oracle, ids, reads)
    run("read_id", "-b", d / f"fa{threads}.mxi", "-q", d / "r_1.fastq.gz", d / "r_2.fastq.gz", "-n", d / f"pe{threads}")
    assert (d / f"pe{threads}_reads.txt").read_text().split("\n")[:-1] == exp_lines
    assert (d / f"pe{threads}_counts.txt").read_text().split("\n")[:-1] == exp_counts
    classes = {l.split("\t")[1] for l in exp_lines}
    assert "too_short" in classes and len(classes - CLS_NAMES) >= 2
    # single end, -B 0, -d 2
    reads = [[oracle.qual_mask(a, qa.encode(), 15)] for a, qa in zip(s1, q1)]
    exp_lines, exp_counts = _expected_read_lines(w, oracle, ids, reads, d=2, start_sample=0)
    run("read_id", "-b", d / f"fa{threads}.mxi", "-q", d / "r_1.fastq.gz", "-n", d / f"se{threads}", "-B", 0, "-d", 2)
    assert (d / f"se{threads}_reads.txt").read_text().split("\n")[:-1] == exp_lines
    assert (d / f"se{threads}_counts.txt").read_text().split("\n")[:-1] == exp_counts
    # search refuses a minimizer index exactly like main.rs:569-573: message on stderr, nothing on stdout, exit 0
    body, r = run("search", "-b", d / f"fa{threads}.mxi", "-q", d / "Mini_A.fasta")
    assert body == [] and "An index with minimizers (.mxi) is used, but not available for this function" in r.stderr
