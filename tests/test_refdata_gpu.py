"""GPU + CLI parity on the reference's OWN bundled inputs (staged by tests/golden/make_refdata.py; /root/reference is not
read here):

  * the build half of the reference's test.sh (:3 one thread, :11 `-t 2`): `build -s 750000 -n 4 -k 27` over
    test_data/ref_file.txt (four phage assemblies) through the colorid-b200 CLI -- .bxi contents equal to the oracle's
    index, 324,869 stored rows, SURVEY Appendix D.1/D.2 numbers, self-queries through `search`;
  * the 16 real genomes of refs/ (k = 31, s = 50 M, n = 4, the C1 parameters): the all-lower-case
    Staphylococcus_aureus_NCTC8532.fasta, the N-rich Listeria assemblies, the 31-contig SRR2167842_ST0 -- n_ref_kmers and
    matrix bits against the oracle, and every genome queried against the index must score hits == n_ref_kmers on itself;
  * whole assemblies as read_id "reads" (read_id_mt_pe.rs:450-569 stream_fasta) through the CLI.
"""
import gzip
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import colorid_b200 as cb
from tests import bxi_py

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "colorid_b200", "colorid-b200")
PHAGE = os.path.join(ROOT, "tests", "golden", "phage")
GENOMES = os.path.join(ROOT, "oracle", "_ref", "refs")
AD = json.load(open(os.path.join(ROOT, "tests", "golden", "appendix_d.json")))
PHAGES = ["Listeria_phage_B021", "Listeria_phage_B051", "Listeria_phage_B056", "Listeria_phage_B545"]


def run(*args):
    r = subprocess.run([CLI, *map(str, args)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.split("\n")[3:]
    return [l for l in lines if l], r


def gunzip_to(src, dst):
    with gzip.open(src, "rb") as f, open(dst, "wb") as g:
        g.write(f.read())


@pytest.fixture(scope="module")
def phage(tmp_path_factory, oracle):
    d = tmp_path_factory.mktemp("phage")
    os.makedirs(d / "test_data" / "refs")
    lines = []
    for n in PHAGES:
        gunzip_to(os.path.join(PHAGE, n + ".fasta.gz"), d / "test_data" / "refs" / (n + ".fasta"))
        lines.append(f"{n}\t{d}/test_data/refs/{n}.fasta")
    (d / "test_data" / "ref_file.txt").write_text("\n".join(lines) + "\n")
    seqs = [oracle.read_fasta(str(d / "test_data" / "refs" / (n + ".fasta"))) for n in PHAGES]
    oix = oracle.Index(750_000, 4, 27, 4)
    for c, s in enumerate(seqs):
        oix.build_accession(c, s, oracle.MODE_FASTA)
    oix.finalize(threads=2)
    return dict(dir=d, seqs=seqs, oix=oix)


@pytest.mark.parametrize("threads", [1, 2])
def test_test_sh_build_half(phage, threads):
    d, oix = phage["dir"], phage["oix"]
    args = ["build", "-s", 750000, "-n", 4, "-k", 27, "-b", d / f"phage_t{threads}", "-r", d / "test_data" / "ref_file.txt"]
    if threads > 1:
        args += ["-t", threads]                                  # test.sh:11
    run(*args)
    got = bxi_py.read_bxi(d / f"phage_t{threads}.bxi")
    assert (got["bloom_size"], got["num_hash"], got["k_size"]) == (750_000, 4, 27)
    assert got["colors"] == dict(enumerate(PHAGES))
    assert got["n_ref"] == dict(zip(PHAGES, [31297, 27583, 32634, 27491]))
    assert len(got["row_ids"]) == 324_869 == AD["D2"]["nonzero_rows"]
    dense = oix.words()
    nz = np.flatnonzero(dense.any(axis=1))
    o = np.argsort(got["row_ids"])
    assert np.array_equal(got["row_ids"][o], nz.astype(np.uint64))
    assert np.array_equal(got["words"][o], dense[nz])
    # the file is byte-identical to the oracle's index written by the independent Python writer in the same row order
    bxi_py.write_bxi(d / "oracle.bxi", 750_000, 4, 27, got["colors"], got["row_ids"], dense[got["row_ids"].astype(np.int64)],
                     dict(zip(got["n_ref"].keys(), [int(oix.n_ref[PHAGES.index(n)]) for n in got["n_ref"]])))
    assert hashlib.sha256(open(d / "oracle.bxi", "rb").read()).digest() == hashlib.sha256(open(d / f"phage_t{threads}.bxi", "rb").read()).digest()


def test_phage_self_queries_through_search(phage, oracle):
    d, oix = phage["dir"], phage["oix"]
    if not os.path.exists(d / "phage_t1.bxi"):
        run("build", "-s", 750000, "-n", 4, "-k", 27, "-b", d / "phage_t1", "-r", d / "test_data" / "ref_file.txt")
    files = [d / "test_data" / "refs" / (n + ".fasta") for n in PHAGES]
    # -g: gene_match = hits / num_kmers per accession (reports.rs:50-62); the own accession scores 1.000
    body, _ = run("search", "-b", d / "phage_t1.bxi", "-g", "-p", 0, "-q", *files)
    o = oix.query_counts(phage["seqs"], oracle.MODE_FASTA, True, -1)
    assert o["counts"].tolist() == AD["D2"]["self_query_counts"]
    exp = sorted(f"{files[q]}\t{PHAGES[c]}\t{int(o['num_kmers'][q])}\t{o['counts'][q, c] / o['num_kmers'][q]:.3f}"
                 for q in range(4) for c in range(4))
    assert sorted(body) == exp
    assert all(f"{files[q]}\t{PHAGES[q]}\t{int(oix.n_ref[q])}\t1.000" in body for q in range(4))
    # -s perfect search of sub-sequences cut from each phage lists that phage (perfect_search.rs:6-60)
    for q, n in enumerate(PHAGES):
        sub = d / f"sub_{q}.fasta"
        sub.write_bytes(b">sub\n" + phage["seqs"][q][0][200:1200] + b"\n")
        body, _ = run("search", "-b", d / "phage_t1.bxi", "-s", "-q", sub)
        assert any(l.split("\t")[1] == n for l in body), body


def test_cli_hash_variant_is_found_by_the_pin_tool(phage):
    """COLORID_B200_HASH_VARIANT selects the XXH3 draft of the CLI; tools/pin_from_bxi.py (numpy only) names it from the .bxi."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("pin_from_bxi", os.path.join(ROOT, "tools", "pin_from_bxi.py"))
    pin = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pin)
    d = phage["dir"]
    for v in (0, 31):
        env = dict(os.environ, COLORID_B200_HASH_VARIANT=str(v))
        r = subprocess.run([CLI, "build", "-s", "750000", "-n", "4", "-k", "27", "-b", str(d / f"hv{v}"), "-r", str(d / "test_data" / "ref_file.txt")],
                           capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        assert pin.pin(d / f"hv{v}.bxi", d / "test_data" / "ref_file.txt") == [v]
        # searching the variant's index needs the variant too: with it the own accession scores 1.000, without it nothing matches
        f = d / "test_data" / "refs" / "Listeria_phage_B056.fasta"
        r = subprocess.run([CLI, "search", "-b", str(d / f"hv{v}.bxi"), "-g", "-q", str(f)], capture_output=True, text=True, env=env)
        assert f"{f}\tListeria_phage_B056\t32634\t1.000" in r.stdout
    r = subprocess.run([CLI, "search", "-b", str(d / "hv31.bxi"), "-g", "-q", str(f)], capture_output=True, text=True)
    assert "Listeria_phage_B056\t32634\t1.000" not in r.stdout and "hash variant" in r.stderr


# ------------------------------------------------------------------------------------------------ the 16 real genomes
have_genomes = pytest.mark.skipif(not os.path.isdir(GENOMES), reason="oracle/_ref/refs not staged (run __graft_entry__.build() where /root/reference exists)")


@pytest.fixture(scope="module")
def genomes(tmp_path_factory, oracle, ctx):
    d = tmp_path_factory.mktemp("refs")
    names = sorted(f[:-len(".fasta.gz")] for f in os.listdir(GENOMES) if f.endswith(".fasta.gz"))
    seqs = []
    for n in names:
        gunzip_to(os.path.join(GENOMES, n + ".fasta.gz"), d / (n + ".fasta"))
        seqs.append(oracle.read_fasta(str(d / (n + ".fasta"))))
    k, S, H = 31, 50_000_000, 4
    oix = oracle.Index(S, H, k, len(names))
    oix.build_many(seqs, oracle.MODE_FASTA, threads=os.cpu_count() or 4)
    gix = cb.Index(ctx, S, H, k, len(names))
    g_nref = [gix.build_accession(c, s, cb.CID_SEQ_FASTA)[0] for c, s in enumerate(seqs)]
    gix.finalize()
    gix.n_ref[:] = g_nref
    return dict(dir=d, names=names, seqs=seqs, oix=oix, gix=gix, g_nref=g_nref)


@have_genomes
def test_real_genomes_index_bits_and_n_ref_kmers(genomes):
    names, oix, gix = genomes["names"], genomes["oix"], genomes["gix"]
    assert len(names) == 16
    assert genomes["g_nref"] == oix.n_ref.tolist()
    d1 = {os.path.basename(r["file"])[:-len(".fasta")]: r["n_ref_kmers"] for r in AD["D1"] if r["k"] == 31}
    for n, v in d1.items():                                      # lower-case Staph, N-rich Listeria, 31 contigs
        assert genomes["g_nref"][names.index(n)] == v, n
    assert d1["Staphylococcus_aureus_NCTC8532"] == 2_733_652 and d1["Listeria_marthii_S4_120"] == 2_699_707
    assert np.array_equal(gix.download_dense(), oix.words())
    assert gix.nonzero_rows() == oix.nonzero_rows()


@have_genomes
def test_real_genomes_self_queries(genomes, oracle):
    names, oix, gix, seqs = genomes["names"], genomes["oix"], genomes["gix"], genomes["seqs"]
    n_ref = np.array(genomes["g_nref"], dtype=np.uint64)
    for gene, uq in ((True, False), (False, True)):              # -g (streaming gather) and the default report (fused kernel)
        g = gix.query_counts(seqs, cb.CID_SEQ_FASTA, gene, 0 if not gene else -1, want_uniq=uq)
        assert np.array_equal(g["num_kmers"], n_ref)
        assert np.array_equal(np.diag(g["counts"]), n_ref.astype(np.uint32)), "hits != n_ref_kmers on the own accession"
    sub = [0, 5, 14]                                             # (the oracle's single-threaded search is slow: three genomes)
    o = oix.query_counts([seqs[i] for i in sub], oracle.MODE_FASTA, False, 0)
    assert np.array_equal(g["counts"][sub], o["counts"])
    for key in ("uniq_n", "uniq_sum", "uniq_mode"):
        assert np.array_equal(g[key][sub], o[key]), key
    # -s: a 5 kb piece of every genome has that genome's bit in the AND row
    cut = [[s[0][1000:6000]] for s in seqs]
    p = gix.query_perfect(cut)
    for i in range(len(names)):
        assert p["status"][i] == 0 and (p["and_rows"][i, 0] >> i) & 1, names[i]


@have_genomes
def test_real_genomes_cli_build_and_contig_read_id(genomes, oracle):
    """Four of the assemblies through the CLI (-t 2): the lower-case Staph, its upper-case sibling, an N-rich Listeria and the
    31-contig SRR2167842_ST0; then SRR2167842_ST0.fasta itself as read_id input (stream_fasta: its 31 contigs are the reads)."""
    d, names, seqs = genomes["dir"], genomes["names"], genomes["seqs"]
    pick = ["Staphylococcus_aureus_NCTC8532", "Staphylococcus_aureus_NCTC8532_2", "Listeria_marthii_S4_120", "SRR2167842_ST0"]
    (d / "four.tsv").write_text("".join(f"{n}\t{d}/{n}.fasta\n" for n in pick))
    k, S, H = 31, 20_000_003, 4
    run("build", "-s", S, "-n", H, "-k", k, "-b", d / "four", "-r", d / "four.tsv", "-t", 2)
    order = sorted(pick)
    oix = oracle.Index(S, H, k, 4)
    oix.build_many([seqs[names.index(n)] for n in order], oracle.MODE_FASTA, threads=4)
    got = bxi_py.read_bxi(d / "four.bxi")
    assert got["colors"] == dict(enumerate(order))
    assert got["n_ref"] == {n: int(oix.n_ref[c]) for c, n in enumerate(order)}
    dense = oix.words()
    nz = np.flatnonzero(dense.any(axis=1))
    o = np.argsort(got["row_ids"])
    assert np.array_equal(got["row_ids"][o], nz.astype(np.uint64)) and np.array_equal(got["words"][o], dense[nz])
    # contigs as reads
    fa = d / "SRR2167842_ST0.fasta"
    run("read_id", "-b", d / "four.bxi", "-q", fa, "-n", d / "ctg")
    ids, reads, sub = [], [], b""
    for line in open(fa, "rb"):                                  # read_id_mt_pe.rs:477-493: the sequence keeps its line feeds
        if b">" in line:
            if sub:
                reads.append([sub]); sub = b""
            ids.append(line[:-1].decode())
        else:
            sub += line
    reads.append([sub])
    res = oix.read_id_batch(reads, top_cap=4)
    lines = (d / "ctg_reads.txt").read_text().split("\n")[:-1]
    assert len(lines) == len(reads) == 31
    for r, line in enumerate(lines):
        f = line.split("\t")
        assert f[0] == ids[r] and int(f[3]) == int(res["n_set"][r]), line
        kind = int(res["kind"][r])
        if kind == oracle.CLS_ACCEPT:
            assert (f[1], int(f[2]), f[4], f[5]) == (order[int(res["top"][r, 0])], int(res["hits"][r]), "accept", "1"), line
        elif kind == oracle.CLS_REJECT_MULTI:
            nt = int(res["n_top"][r])
            assert f[1] == ",".join(order[int(c)] for c in res["top"][r, :nt]) and f[4] == "reject", line
        else:
            assert f[1] == {0: "too_short", 1: "no_hits", 2: "no_significant_hits"}[kind], line
    assert sum(l.split("\t")[1] == "SRR2167842_ST0" for l in lines) >= 20
