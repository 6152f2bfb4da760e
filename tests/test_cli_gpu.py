"""GPU tests of the host layer end to end: the colorid-b200 CLI (build / search / read_id on files) against the
oracle's restatement of the reference, byte for byte on the .bxi contents and the output lines (lines whose
order is hash-random in the reference are compared as sorted lists)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from tests import bxi_py, synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "colorid_b200", "colorid-b200")
K, S, H = 21, 300_007, 3


def run(*args, ok=True, env=None):
    r = subprocess.run([CLI, *map(str, args)], capture_output=True, text=True, env=None if env is None else dict(os.environ, **env))
    if ok:
        assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.split("\n")
    body = lines[3:]
    if body and body[-1] == "":
        body = body[:-1]
    return body, r


def wrap_fasta(path, records, width=70):
    with open(path, "w") as f:
        for name, seq in records:
            f.write(f">{name}\n")
            s = seq.decode()
            for i in range(0, len(s), width):
                f.write(s[i:i + width] + "\n")


def fastq_bytes(names, seqs, quals):
    return "".join(f"@{n}\n{s.decode()}\n+\n{q}\n" for n, s, q in zip(names, seqs, quals)).encode()


def rand_quals(rng, n, low=0.03):
    q = np.full(n, 73, np.uint8)
    q[rng.random(n) < low] = 35
    return q.tobytes().decode()


def read_pairs(rng, genomes, n, tag, **kw):
    reads = synth.reads_from(rng, genomes, n, **kw)
    names = [f"{tag}.{i} {i}/1" for i in range(n)]
    q1 = [rand_quals(rng, len(r[0])) for r in reads]
    q2 = [rand_quals(rng, len(r[1])) for r in reads]
    return names, [r[0] for r in reads], [r[1] for r in reads], q1, q2


@pytest.fixture(scope="module")
def world(tmp_path_factory, oracle):
    """5 FASTA accessions (2 contigs each, one lower-case), one paired-end and one single-end FASTQ.gz accession."""
    d = tmp_path_factory.mktemp("cli")
    rng = np.random.default_rng(0xC0101D77)
    genomes = synth.clade_genomes(rng, 7, 6000, n_clades=3, div=0.02)
    acc = {}
    refs = []
    for i in range(5):
        g = genomes[i]
        if i == 2:
            g = g.lower()
        name = f"Acc_{chr(ord('E') - i)}"                      # file order differs from sorted (colour) order
        contigs = [g[:3500], g[3500:]]
        wrap_fasta(d / f"{name}.fasta", [(f"{name}_c{j}", c) for j, c in enumerate(contigs)])
        acc[name] = dict(mode=oracle.MODE_FASTA, seqs=contigs)
        refs.append(f"{name}\t{d}/{name}.fasta")
    n, s1, s2, q1, q2 = read_pairs(rng, [genomes[5]], 900, "pe", read_len=120, insert=260, err=0.004, frac_random=0.0)
    (d / "pe_1.fastq.gz").write_bytes(gzip.compress(fastq_bytes(n, s1, q1)))
    (d / "pe_2.fastq.gz").write_bytes(gzip.compress(fastq_bytes(n, s2, q2)))
    masked = []
    for a, b, qa, qb in zip(s1, s2, q1, q2):
        masked += [oracle.qual_mask(a, qa.encode(), 15), oracle.qual_mask(b, qb.encode(), 15)]
    acc["Reads_PE"] = dict(mode=oracle.MODE_FASTQ, seqs=masked)
    refs.append(f"Reads_PE\t{d}/pe_1.fastq.gz\t{d}/pe_2.fastq.gz")
    n, s1, _, q1, _ = read_pairs(rng, [genomes[6]], 1500, "se", read_len=120, insert=260, err=0.004, frac_random=0.0)
    (d / "se.fastq.gz").write_bytes(gzip.compress(fastq_bytes(n, s1, q1)))
    acc["Reads_SE"] = dict(mode=oracle.MODE_FASTQ, seqs=[oracle.qual_mask(a, qa.encode(), 15) for a, qa in zip(s1, q1)])
    refs.append(f"Reads_SE\t{d}/se.fastq.gz")
    (d / "refs.tsv").write_text("\n".join(refs) + "\n")
    names = sorted(acc)                                         # colour = rank in byte-wise sorted order
    oix = oracle.Index(S, H, K, len(names))
    for c, name in enumerate(names):
        oix.build_accession(c, acc[name]["seqs"], acc[name]["mode"], -1)
    oix.finalize()
    body, r = run("build", "-b", d / "idx", "-r", d / "refs.tsv", "-k", K, "-n", H, "-s", S)
    assert body[-1] == "Saving BIGSI to file."
    return dict(dir=d, genomes=genomes, names=names, oix=oix, rng=rng)


def test_build_writes_the_reference_index(world, oracle):
    got = bxi_py.read_bxi(world["dir"] / "idx.bxi")
    oix, names = world["oix"], world["names"]
    assert (got["bloom_size"], got["num_hash"], got["k_size"]) == (S, H, K)
    assert got["colors"] == dict(enumerate(names))
    assert got["n_ref"] == {n: int(oix.n_ref[c]) for c, n in enumerate(names)}
    dense = oix.words()
    nz = np.flatnonzero(dense.any(axis=1))
    o = np.argsort(got["row_ids"])
    assert np.array_equal(got["row_ids"][o], nz.astype(np.uint64))
    assert np.array_equal(got["words"][o], dense[nz])


def _gene_files(world, n=12):
    d, rng, genomes = world["dir"], world["rng"], world["genomes"]
    files, seqs = [], []
    for i in range(n):
        g = genomes[int(rng.integers(0, 5))]
        s = int(rng.integers(0, len(g) - 1600))
        q = g[s:s + int(rng.integers(200, 1500))]
        if i % 3 == 1:
            q = synth.mutate(rng, q, 0.03)
        if i % 5 == 4:
            q = synth.rand_seq(rng, 600)
        p = d / f"gene_{i}.fasta"
        wrap_fasta(p, [(f"gene_{i}", q)])
        files.append(str(p))
        seqs.append([q])
    return files, seqs


def test_search_gene_mode(world, oracle):
    files, seqs = _gene_files(world)
    o = world["oix"].query_counts(seqs, oracle.MODE_FASTA, True, -1)
    exp = []
    for q, f in enumerate(files):
        nk = int(o["num_kmers"][q])
        for c, name in enumerate(world["names"]):
            v = int(o["counts"][q, c])
            if v and v / nk >= 0.35:
                exp.append(f"{f}\t{name}\t{nk}\t{v / nk:.3f}")
    body, _ = run("search", "-b", world["dir"] / "idx.bxi", "-g", "-q", *files)
    assert len(exp) >= 8
    assert sorted(body) == sorted(exp)


def test_search_default_report_fasta_and_fastq(world, oracle):
    d, rng, names, oix = world["dir"], world["rng"], world["names"], world["oix"]
    files, seqs = _gene_files(world, 4)

    def expected(fname, o, q, cov):
        out = []
        nk = int(o["num_kmers"][q])
        for c, name in enumerate(names):
            v = int(o["counts"][q, c])
            if not v:
                continue
            gc = v / int(oix.n_ref[c])
            un = int(o["uniq_n"][q, c])
            mean = int(o["uniq_sum"][q, c]) / un if un else 0.0
            if gc > cov:
                out.append(f"{fname}\t{nk}\t{name}\t{gc:.2f}\t{mean:.2f}\t{int(o['uniq_mode'][q, c])}\t{un}")
        return out

    o = oix.query_counts(seqs, oracle.MODE_FASTA, False, 0)
    exp = [l for q, f in enumerate(files) for l in expected(f, o, q, 0.01)]
    body, _ = run("search", "-b", d / "idx.bxi", "-f", 0, "-p", 0.01, "-q", *files)
    assert len(exp) >= 3 and sorted(body) == sorted(exp)
    # paired FASTQ sample: quality masking + auto_cutoff on the query (batch_search_pe.rs:24-41)
    n, s1, s2, q1, q2 = read_pairs(rng, [world["genomes"][1]], 800, "q", read_len=120, insert=260, err=0.004, frac_random=0.02)
    (d / "q_1.fastq.gz").write_bytes(gzip.compress(fastq_bytes(n, s1, q1)))
    (d / "q_2.fastq.gz").write_bytes(gzip.compress(fastq_bytes(n, s2, q2)))
    masked = []
    for a, b, qa, qb in zip(s1, s2, q1, q2):
        masked += [oracle.qual_mask(a, qa.encode(), 20), oracle.qual_mask(b, qb.encode(), 20)]
    o = oix.query_counts([masked], oracle.MODE_FASTQ, False, -1)
    exp = expected(str(d / "q_1.fastq.gz"), o, 0, 0.35)
    body, _ = run("search", "-b", d / "idx.bxi", "-Q", 20, "-q", d / "q_1.fastq.gz", "-r", d / "q_2.fastq.gz")
    assert len(exp) >= 1 and sorted(body) == sorted(exp)


def test_search_perfect_and_multifasta(world, oracle):
    d, rng, names, oix, genomes = world["dir"], world["rng"], world["names"], world["oix"], world["genomes"]
    files, seqs = [], []
    for i in range(6):
        g = genomes[i % 5]
        s = int(rng.integers(0, 3000))
        q = g[s:s + 300] if i != 4 else synth.rand_seq(rng, 300)
        p = d / f"perf_{i}.fasta"
        wrap_fasta(p, [(f"p{i}", q)])
        files.append(str(p))
        seqs.append([q])
    o = oix.query_perfect(seqs)
    exp = []
    for q, f in enumerate(files):
        if o["status"][q] == 0:
            for c, name in enumerate(names):
                if (o["and_rows"][q, c // 32] >> (c % 32)) & 1:
                    exp.append(f"{f}\t{name}\t{int(o['n_kmers'][q])}\t1.00")
    body, r = run("search", "-b", d / "idx.bxi", "-s", "-q", *files)
    assert len(exp) >= 5 and body == exp                      # ascending colour order inside a query, queries in order
    assert "No perfect hits!" in r.stderr
    # -s -m: one query per record, labels from the headers; a short record gives the stdout warning
    recs = [(f"rec{i} note", s[0]) for i, s in enumerate(seqs)] + [("tiny", b"ACGTACGT")]
    wrap_fasta(d / "multi.fasta", recs)
    o = oix.query_perfect([s for _, s in recs], mf=True)
    exp = []
    for q, (label, _) in enumerate(recs):
        if o["status"][q] == 2:
            exp.append(f"Warning! no kmers in query '{label}'; maybe your kmer length is larger than your query length?")
        elif o["status"][q] == 0:
            for c, name in enumerate(names):
                if (o["and_rows"][q, c // 32] >> (c % 32)) & 1:
                    exp.append(f"{label}\t{name}\t{int(o['n_kmers'][q])}\t1.00")
    body, _ = run("search", "-b", d / "idx.bxi", "-s", "-m", "-q", d / "multi.fasta")
    assert body == exp


CLS = {0: ("too_short", "accept"), 1: ("no_hits", "accept"), 2: ("no_significant_hits", "reject")}


def _expected_read_lines(world, oracle, ids, reads, **kw):
    oix, names = world["oix"], world["names"]
    o = oix.read_id_batch(reads, top_cap=len(names), **kw)
    lines, counts, order = [], {}, []
    for r, rid in enumerate(ids):
        kind = int(o["kind"][r])
        ns = int(o["n_set"][r])
        if kind in CLS:
            cls, verdict = CLS[kind]
            h, nt = 0, 0
            if kind == 0:
                ns = 0
        elif kind == oracle.CLS_ACCEPT:
            cls, verdict, h, nt = names[int(o["top"][r, 0])], "accept", int(o["hits"][r]), 1
        else:
            assert kind == oracle.CLS_REJECT_MULTI
            nt = int(o["n_top"][r])
            cls, verdict, h = ",".join(names[int(c)] for c in o["top"][r, :nt]), "reject", int(o["hits"][r])
        lines.append(f"{rid}\t{cls}\t{h}\t{ns}\t{verdict}\t{nt}")
        key = cls if verdict == "accept" else "reject"
        if key not in counts:
            order.append(key)
        counts[key] = counts.get(key, 0) + 1
    it, _ = oracle.hashset_str_order([k.encode() for k in order], 16, True)
    return lines, [f"{order[i]}\t{counts[order[i]]}" for i in it]


def test_read_id_paired_single_and_fasta(world, oracle):
    d, rng, genomes = world["dir"], world["rng"], world["genomes"]
    n, s1, s2, q1, q2 = read_pairs(rng, genomes[:5], 3000, "r", read_len=100, insert=180, err=0.006, frac_random=0.2)
    s1[5], q1[5] = s1[5][:15], q1[5][:15]                      # mate 1 shorter than k -> too_short
    (d / "r_1.fastq.gz").write_bytes(gzip.compress(fastq_bytes(n, s1, q1)))
    (d / "r_2.fastq.gz").write_bytes(gzip.compress(fastq_bytes(n, s2, q2)))
    ids = ["@" + x for x in n]
    reads = [[oracle.qual_mask(a, qa.encode(), 15), oracle.qual_mask(b, qb.encode(), 15)] for a, b, qa, qb in zip(s1, s2, q1, q2)]
    exp_lines, exp_counts = _expected_read_lines(world, oracle, ids, reads)
    run("read_id", "-b", d / "idx.bxi", "-q", d / "r_1.fastq.gz", d / "r_2.fastq.gz", "-n", d / "pe_out")
    assert (d / "pe_out_reads.txt").read_text().split("\n")[:-1] == exp_lines
    assert (d / "pe_out_counts.txt").read_text().split("\n")[:-1] == exp_counts
    classes = {l.split("\t")[1] for l in exp_lines}
    assert "too_short" in classes and len(classes - CLS_NAMES) >= 3 and len(classes & CLS_NAMES) >= 2
    # the host layer's threaded paths on this small input (several inflating threads per file, bases copied and report lines
    # formatted by four threads each), and the line-by-line record loop with the sequential decoder
    for name, env in (("par", {"COLORID_B200_PAR_MIN": "1", "COLORID_B200_GZ_THREADS": "3", "COLORID_B200_GZ_SPAN": "8192"}),
                      ("seq", {"COLORID_B200_PARSE_SLOW": "1", "COLORID_B200_GZ_THREADS": "1"})):
        run("read_id", "-b", d / "idx.bxi", "-q", d / "r_1.fastq.gz", d / "r_2.fastq.gz", "-n", d / f"pe_{name}", env=env)
        assert (d / f"pe_{name}_reads.txt").read_text().split("\n")[:-1] == exp_lines, name
        assert (d / f"pe_{name}_counts.txt").read_text().split("\n")[:-1] == exp_counts, name
    # single end, -B 0 (search_index_classic), -d 2, -p 2, -Q 25
    reads = [[oracle.qual_mask(a, qa.encode(), 25)] for a, qa in zip(s1, q1)]
    exp_lines, exp_counts = _expected_read_lines(world, oracle, ids, reads, d=2, start_sample=0, fp_correct=10.0 ** -2.0)
    run("read_id", "-b", d / "idx.bxi", "-q", d / "r_1.fastq.gz", "-n", d / "se_out", "-B", 0, "-d", 2, "-p", 2, "-Q", 25, "-t", 2)
    assert (d / "se_out_reads.txt").read_text().split("\n")[:-1] == exp_lines
    assert (d / "se_out_counts.txt").read_text().split("\n")[:-1] == exp_counts
    # FASTA reads (stream_fasta keeps the line feeds inside the sequence, read_id_mt_pe.rs:477-493)
    recs = [(f"fa{i}", s1[i] + s2[i]) for i in range(0, 400)]
    wrap_fasta(d / "reads.fasta", recs, width=60)
    ids, reads = [], []
    for name, s in recs:
        ids.append(">" + name)
        t = s.decode()
        reads.append([("".join(t[i:i + 60] + "\n" for i in range(0, len(t), 60))).encode()])
    exp_lines, exp_counts = _expected_read_lines(world, oracle, ids, reads)
    run("read_id", "-b", d / "idx.bxi", "-q", d / "reads.fasta", "-n", d / "fa_out")
    assert (d / "fa_out_reads.txt").read_text().split("\n")[:-1] == exp_lines
    assert (d / "fa_out_counts.txt").read_text().split("\n")[:-1] == exp_counts


CLS_NAMES = {"too_short", "no_hits", "no_significant_hits"}


@pytest.mark.parametrize("shard", ["columns", "replicated"])
def test_multi_gpu_cli_outputs_equal_the_single_gpu_outputs(world, oracle, shard):
    """COLORID_B200_DEVICES=.. (one index over several GPUs: cid_mg) with COLORID_B200_SHARD=columns|replicated: `build` writes
    the same .bxi bytes, `search` (default report, -g, -s, -s -m) the same lines and `read_id` the same files as on one GPU.
    The device list repeats GPU 0 when the box has a single GPU (several shards on one device: the same host code path)."""
    import ctypes
    nd = ctypes.c_int(0)
    ctypes.CDLL("libcudart.so").cudaGetDeviceCount(ctypes.byref(nd))
    devs = ",".join(str(g % max(nd.value, 1)) for g in range(3))
    env = dict(os.environ, COLORID_B200_DEVICES=devs, COLORID_B200_SHARD=shard)
    d = world["dir"]

    def both(*args, out_files=()):
        one = subprocess.run([CLI, *map(str, args)], capture_output=True, text=True)
        keep = {f: open(f, "rb").read() for f in out_files}
        many = subprocess.run([CLI, *map(str, args)], capture_output=True, text=True, env=env)
        assert one.returncode == 0 and many.returncode == 0, many.stderr[-2000:]
        assert "3 GPUs, index " + ("column-sharded" if shard == "columns" else "replicated") in many.stderr
        assert sorted(one.stdout.split("\n")) == sorted(many.stdout.split("\n"))
        for f in out_files:
            assert open(f, "rb").read() == keep[f], f
    both("build", "-b", d / "mgidx", "-r", d / "refs.tsv", "-k", K, "-n", H, "-s", S, out_files=[str(d / "mgidx.bxi")])
    assert open(d / "mgidx.bxi", "rb").read() == open(d / "idx.bxi", "rb").read()
    files, _ = _gene_files(world, n=7)
    both("search", "-b", d / "idx.bxi", "-g", "-q", *files)
    both("search", "-b", d / "idx.bxi", "-f", 0, "-p", 0.05, "-q", *files)
    both("search", "-b", d / "idx.bxi", "-q", d / "pe_1.fastq.gz", "-r", d / "pe_2.fastq.gz")
    both("search", "-b", d / "idx.bxi", "-s", "-q", *files)
    n, s1, s2, q1, q2 = read_pairs(world["rng"], world["genomes"][:5], 1200, "mg", read_len=100, insert=180, err=0.006, frac_random=0.2)
    (d / "mg_1.fastq.gz").write_bytes(gzip.compress(fastq_bytes(n, s1, q1)))
    (d / "mg_2.fastq.gz").write_bytes(gzip.compress(fastq_bytes(n, s2, q2)))
    both("read_id", "-b", d / "idx.bxi", "-q", d / "mg_1.fastq.gz", d / "mg_2.fastq.gz", "-n", d / "mg_out",
         out_files=[str(d / "mg_out_reads.txt"), str(d / "mg_out_counts.txt")])


def test_read_id_contigs_and_softmasked_reads(world, oracle):
    """read_id_mt_pe.rs:450-569 stream_fasta on real assemblies: records are whole contigs (single-line and wrapped, one all
    lower-case like refs/Staphylococcus_aureus_NCTC8532.fasta), far above the 1,000 bases of the warp-per-read kernels; and
    a FASTQ whose reads carry soft-masked (lower-case) stretches, which kmer.rs:221-243 hashes as they are.  One odd read must
    never abort the run: every line of PREFIX_reads.txt equals the oracle's."""
    d, rng, genomes = world["dir"], world["rng"], world["genomes"]
    recs = [("contig_wrapped", genomes[0][:6000]), ("contig_lower", genomes[1][:5000].lower()),
            ("contig_mixed", synth.sprinkle(rng, genomes[2][1000:4200], b"acgtN", 0.03)), ("short", genomes[3][:200]),
            ("random", synth.rand_seq(rng, 2500))]
    wrap_fasta(d / "contigs.fasta", recs, width=80)
    ids, reads = [], []
    for name, s in recs:
        ids.append(">" + name)
        t = s.decode()
        reads.append([("".join(t[i:i + 80] + "\n" for i in range(0, len(t), 80))).encode()])
    exp_lines, exp_counts = _expected_read_lines(world, oracle, ids, reads)
    run("read_id", "-b", d / "idx.bxi", "-q", d / "contigs.fasta", "-n", d / "ctg_out")
    assert (d / "ctg_out_reads.txt").read_text().split("\n")[:-1] == exp_lines
    assert (d / "ctg_out_counts.txt").read_text().split("\n")[:-1] == exp_counts
    # single-line records: every k-mer of a contig is valid -> sets of thousands of k-mers, hits in the thousands
    with open(d / "contigs1.fasta", "wb") as f:
        for name, s in recs:
            f.write(b">" + name.encode() + b"\n" + s + b"\n")
    reads1 = [[s + b"\n"] for _, s in recs]
    exp_lines, exp_counts = _expected_read_lines(world, oracle, ids, reads1)
    run("read_id", "-b", d / "idx.bxi", "-q", d / "contigs1.fasta", "-n", d / "ctg1_out", "-B", 0)
    exp0, expc0 = _expected_read_lines(world, oracle, ids, reads1, start_sample=0)
    assert (d / "ctg1_out_reads.txt").read_text().split("\n")[:-1] == exp0
    run("read_id", "-b", d / "idx.bxi", "-q", d / "contigs1.fasta", "-n", d / "ctg1b_out")
    got = (d / "ctg1b_out_reads.txt").read_text().split("\n")[:-1]
    assert got == exp_lines and int(got[0].split("\t")[3]) > 5000
    # soft-masked FASTQ, single end
    n, s1, s2, q1, q2 = read_pairs(rng, genomes[:5], 500, "m", read_len=120, insert=200, err=0.004, frac_random=0.1)
    for i in range(0, 500, 3):
        s1[i] = s1[i][:30] + s1[i][30:75].lower() + s1[i][75:]
    (d / "m_1.fastq.gz").write_bytes(gzip.compress(fastq_bytes(n, s1, q1)))
    reads = [[oracle.qual_mask(a, qa.encode(), 15)] for a, qa in zip(s1, q1)]
    exp_lines, exp_counts = _expected_read_lines(world, oracle, ["@" + x for x in n], reads)
    run("read_id", "-b", d / "idx.bxi", "-q", d / "m_1.fastq.gz", "-n", d / "m_out")
    assert (d / "m_out_reads.txt").read_text().split("\n")[:-1] == exp_lines
    assert (d / "m_out_counts.txt").read_text().split("\n")[:-1] == exp_counts


def test_batch_id_classifies_every_sample_with_one_index_load(world, oracle):
    # read_id_batch.rs:7-181: SAMPLE_TAG_reads.txt / SAMPLE_TAG_counts.txt per line of the sample list
    d, rng, genomes = world["dir"], world["rng"], world["genomes"]
    lines, expect = [], {}
    for si, paired in enumerate([True, False, True]):
        n, s1, s2, q1, q2 = read_pairs(rng, genomes[:5], 400 + 50 * si, f"b{si}", read_len=100, insert=180, err=0.006, frac_random=0.2)
        (d / f"b{si}_1.fastq.gz").write_bytes(gzip.compress(fastq_bytes(n, s1, q1)))
        (d / f"b{si}_2.fastq.gz").write_bytes(gzip.compress(fastq_bytes(n, s2, q2)))
        ids = ["@" + x for x in n]
        if paired:
            reads = [[oracle.qual_mask(a, qa.encode(), 15), oracle.qual_mask(b, qb.encode(), 15)] for a, b, qa, qb in zip(s1, s2, q1, q2)]
            lines.append(f"{d}/sample{si}\t{d}/b{si}_1.fastq.gz\t{d}/b{si}_2.fastq.gz")
        else:
            reads = [[oracle.qual_mask(a, qa.encode(), 15)] for a, qa in zip(s1, q1)]
            lines.append(f"{d}/sample{si}\t{d}/b{si}_1.fastq.gz")
        expect[si] = _expected_read_lines(world, oracle, ids, reads, d=2)
    (d / "samples.tsv").write_text("\n".join(lines) + "\n")
    _, r = run("batch_id", "-b", d / "idx.bxi", "-q", d / "samples.tsv", "-T", "run7", "-d", 2)
    assert [l for l in r.stderr.split("\n") if l.startswith("Classifying")] == [f"Classifying {d}/sample{i}" for i in range(3)]
    for si, (exp_lines, exp_counts) in expect.items():
        assert (d / f"sample{si}_run7_reads.txt").read_text().split("\n")[:-1] == exp_lines
        assert (d / f"sample{si}_run7_counts.txt").read_text().split("\n")[:-1] == exp_counts
    # samples dealt to one worker per GPU (replicated index): same files whatever the number of devices
    import torch
    ndev = torch.cuda.device_count()
    devs = ",".join(str(i % ndev) for i in range(2))          # two workers (on one GPU if the box has a single one)
    r = subprocess.run([CLI, "batch_id", "-b", str(d / "idx.bxi"), "-q", str(d / "samples.tsv"), "-T", "multi", "-d", "2"],
                       capture_output=True, text=True, env=dict(os.environ, COLORID_B200_DEVICES=devs))
    assert r.returncode == 0, r.stderr[-2000:]
    for si, (exp_lines, exp_counts) in expect.items():
        assert (d / f"sample{si}_multi_reads.txt").read_text().split("\n")[:-1] == exp_lines
        assert (d / f"sample{si}_multi_counts.txt").read_text().split("\n")[:-1] == exp_counts
