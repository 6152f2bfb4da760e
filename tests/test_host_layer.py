"""CPU tests of the C++ host layer above the C ABI (colorid_b200/host): FASTA/FASTQ(.gz) readers, qual_mask,
tab_to_map, the .bxi reader/writer and the counts-file ordering, each against the oracle's restatement of the
reference (oracle/pyoracle.py) or an independent Python implementation.  No GPU is touched."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as O
from tests import bxi_py, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "colorid_b200", "colorid-b200")


def host(*args):
    r = subprocess.run([CLI, "_host", *map(str, args)], capture_output=True, text=True, check=True)
    lines = r.stdout.split("\n")
    assert "initializing logger" in lines[1]
    return lines[3:-1] if lines[-1] == "" else lines[3:]


def cli(*args):
    return subprocess.run([CLI, *map(str, args)], capture_output=True, text=True)


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "colorid_b200", "host")])


FASTA_CASES = [
    ">a desc\nACGT\nTTGA\n>b\nGG\n",                       # plain
    ">a\nACGT\nTTGA",                                       # no final newline
    ">a\r\nACGT\r\nTT\r\n>b\r\nCC\r\n",                     # CRLF
    "ACGT\n>a\nCC\n\nGG\n>b\n>c\nTT\n",                     # sequence before the first header, blank line, empty record
    ">a\nAC>GT\nTT\n",                                      # '>' inside a sequence line counts as a header (kmer.rs:26)
    ">a\nACGT\n>tail\n",                                    # header on the last line
    "",
]


@pytest.mark.parametrize("i", range(len(FASTA_CASES)))
def test_read_fasta_matches_reference_quirks(tmp_path, i):
    p = tmp_path / "x.fa"
    p.write_bytes(FASTA_CASES[i].encode())
    assert [s.encode() for s in host("fasta", p)] == O.read_fasta(str(p))
    labels, seqs = O.read_fasta_mf(str(p))
    got = host("fasta_mf", p)
    assert [x[2:].encode() for x in got if x.startswith("L\t")] == labels
    assert [x[2:].encode() for x in got if x.startswith("S\t")] == seqs


def _fastq(records):
    return "".join(f"@{i}\n{s}\n+\n{q}\n" for i, (s, q) in enumerate(records)).encode()


def test_fastq_readers_and_qual_mask(tmp_path):
    rng = np.random.default_rng(5)
    recs1, recs2 = [], []
    for i in range(50):
        n = int(rng.integers(1, 60))
        s = synth.rand_seq(rng, n).decode()
        q = "".join(chr(int(x)) for x in rng.integers(33, 75, n))
        recs1.append((s, q if i % 7 else q[:max(1, n // 2)]))            # a shorter quality string truncates (seq.rs:45)
        s2 = synth.rand_seq(rng, n).decode()
        recs2.append((s2, "".join(chr(int(x)) for x in rng.integers(33, 75, n))))
    p1, p2 = tmp_path / "r1.fq.gz", tmp_path / "r2.fq.gz"
    # two gzip members in one file: MultiGzDecoder reads through both
    p1.write_bytes(gzip.compress(_fastq(recs1[:20])) + gzip.compress(_fastq(recs1[20:])))
    p2.write_bytes(gzip.compress(_fastq(recs2[:40])))                     # the shorter mate file ends the stream
    for Q in (0, 15, 30):
        got = host("fastq", Q, p1)
        assert got[0] == "records\t50"
        assert [g.encode() for g in got[1:]] == [O.qual_mask(s.encode(), q.encode(), Q) for s, q in recs1]
        got = host("fastq", Q, p1, p2)
        assert got[0] == "records\t40"
        exp = []
        for (s1, q1), (s2, q2) in zip(recs1[:40], recs2[:40]):
            exp += [O.qual_mask(s1.encode(), q1.encode(), Q), O.qual_mask(s2.encode(), q2.encode(), Q)]
        assert [g.encode() for g in got[1:]] == exp
    plain = tmp_path / "plain.fq"
    plain.write_bytes(_fastq(recs2[:5]))
    assert host("fastq", 0, plain)[0] == "records\t5"


def test_qual_longer_than_sequence_fails_like_the_reference(tmp_path):
    p = tmp_path / "bad.fq"
    p.write_bytes(b"@r\nACG\n+\nIIIII\n")
    r = cli("_host", "fastq", 15, p)
    assert r.returncode == 101 and "could not get the next nt" in r.stderr


def test_tab_to_map(tmp_path):
    p = tmp_path / "refs.tsv"
    p.write_text("b\tb.fa\nA\tr1.gz\tr2.gz\na\ta.fa\nb\tb2.fa\n")
    # byte-wise sorted accessions (colour order, build.rs:102-113); the later duplicate wins (HashMap::insert)
    assert host("tab", p) == ["A\tr1.gz\tr2.gz", "a\ta.fa", "b\tb2.fa"]


@pytest.mark.parametrize("N", [1, 33, 70])
def test_bxi_roundtrip(tmp_path, N):
    rng = np.random.default_rng(N)
    W = (N + 31) // 32
    S = 5000
    row_ids = np.sort(rng.choice(S, size=700, replace=False)).astype(np.uint64)
    words = rng.integers(0, 2**32, size=(700, W), dtype=np.uint64).astype(np.uint32)
    if N % 32:
        words[:, -1] &= np.uint32((1 << (N % 32)) - 1)
    colors = {c: f"acc_{c:03d}" for c in range(N)}
    n_ref = {f"acc_{c:03d}": int(rng.integers(1, 10**7)) for c in range(N)}
    a, b = tmp_path / "a.bxi", tmp_path / "b.bxi"
    # the reference writes map entries in hash-table order: readers must accept any order
    bxi_py.write_bxi(a, S, 3, 21, colors, row_ids, words, n_ref, row_order=rng.permutation(700))
    assert cli("_host", "bxi_copy", a, b).returncode == 0
    got = bxi_py.read_bxi(b)
    assert (got["bloom_size"], got["num_hash"], got["k_size"]) == (S, 3, 21)
    assert got["colors"] == colors and got["n_ref"] == n_ref
    o = np.argsort(got["row_ids"])
    assert np.array_equal(got["row_ids"][o], row_ids)
    assert np.array_equal(got["words"][o], words)
    # `info` (main.rs:630-703)
    r = cli("info", "-b", a)
    lines = r.stdout.split("\n")
    assert lines[3:8] == ["BIGSI parameters:", f"Bloomfilter-size: {S}", "Number of hashes: 3", "K-mer size: 21",
                          f"Number of accessions in index: {N}"]
    for c in range(N):
        name = colors[c]
        assert lines[8 + c] == f"{name} {n_ref[name]} {O.false_prob(S, 3, n_ref[name]):.3f}"


def test_mxi_roundtrip_and_info(tmp_path):
    # bigsi.rs:40-49,71-89: BigsyMapMiniNew = BigsyMapNew with m_size after k_size
    rng = np.random.default_rng(5)
    N, S = 40, 3000
    row_ids = np.sort(rng.choice(S, size=300, replace=False)).astype(np.uint64)
    words = rng.integers(0, 2**32, size=(300, 2), dtype=np.uint64).astype(np.uint32)
    words[:, -1] &= np.uint32((1 << (N % 32)) - 1)
    colors = {c: f"acc_{c:03d}" for c in range(N)}
    n_ref = {f"acc_{c:03d}": int(rng.integers(1, 10**6)) for c in range(N)}
    a, b = tmp_path / "a.mxi", tmp_path / "b.mxi"
    bxi_py.write_bxi(a, S, 4, 31, colors, row_ids, words, n_ref, row_order=rng.permutation(300), m_size=15)
    assert cli("_host", "mxi_copy", a, b).returncode == 0
    got = bxi_py.read_bxi(b, mini=True)
    assert (got["bloom_size"], got["num_hash"], got["k_size"], got["m_size"]) == (S, 4, 31, 15)
    assert got["colors"] == colors and got["n_ref"] == n_ref
    o = np.argsort(got["row_ids"])
    assert np.array_equal(got["row_ids"][o], row_ids) and np.array_equal(got["words"][o], words)
    lines = cli("info", "-b", a).stdout.split("\n")          # main.rs:633-668: suffix decides the struct
    assert lines[3:10] == ["BIGSI parameters:", f"Bloomfilter-size: {S}", "Number of hashes: 4", "K-mer size: 31",
                           " minimizer size: 15", "", f"Number of accessions in index: {N}"]
    assert lines[10] == f"acc_000 {n_ref['acc_000']} {O.false_prob(S, 4, n_ref['acc_000']):.3f}"
    # an .mxi read as a .bxi (or the reverse) does not deserialize
    c = tmp_path / "c.bxi"
    c.write_bytes(a.read_bytes())
    assert cli("info", "-b", c).returncode == 101


def test_read_filter_matches_reference_semantics(tmp_path):
    # read_filter.rs:10-191: class column CONTAINS the taxon (accept/reject is not looked at); read id = header up to
    # the first blank; paired input stops at the shorter file; -e inverts
    rng = np.random.default_rng(3)
    n = 60
    hdr = [f"@read.{i} {i}/1" for i in range(n)]
    cls = [["Listeria_mono", "Listeria_mono,Listeria_innocua", "Salmonella", "no_hits"][int(x)] for x in rng.integers(0, 4, n)]
    (tmp_path / "x_reads.txt").write_text("".join(f"{h}\t{c}\t5\t120\t{'reject' if ',' in c else 'accept'}\t1\n" for h, c in zip(hdr, cls)))
    recs1 = [(h, "ACGT" * 5 + str(i % 7) * 0, "I" * 20) for i, h in enumerate(hdr)]
    recs2 = [(h.replace("/1", "/2"), "TTGCA" * 4, "H" * 20) for h in hdr[:50]]               # mate file is shorter
    fq = lambda recs: "".join(f"{h}\n{s}\n+\n{q}\n" for h, s, q in recs).encode()
    (tmp_path / "r1.fq.gz").write_bytes(gzip.compress(fq(recs1)))
    (tmp_path / "r2.fq.gz").write_bytes(gzip.compress(fq(recs2)))
    want = [i for i in range(n) if "Listeria_mono" in cls[i]]
    r = cli("read_filter", "-c", tmp_path / "x_reads.txt", "-f", tmp_path / "r1.fq.gz", tmp_path / "r2.fq.gz", "-t", "Listeria_mono",
            "-p", tmp_path / "out")
    assert r.returncode == 0, r.stderr
    pe = [i for i in want if i < 50]
    assert gzip.decompress((tmp_path / "out_Listeria_mono_R1.fq.gz").read_bytes()) == fq([recs1[i] for i in pe])
    assert gzip.decompress((tmp_path / "out_Listeria_mono_R2.fq.gz").read_bytes()) == fq([recs2[i] for i in pe])
    assert f"Wrote {len(pe)} read-pairs with classification containing 'Listeria_mono' to output files" in r.stderr
    r = cli("read_filter", "-c", tmp_path / "x_reads.txt", "-f", tmp_path / "r1.fq.gz", "-t", "Listeria_mono", "-p", tmp_path / "ex", "-e")
    assert r.returncode == 0, r.stderr
    rest = [i for i in range(n) if i not in want]
    assert gzip.decompress((tmp_path / "ex_Listeria_mono.fq.gz").read_bytes()) == fq([recs1[i] for i in rest])
    assert f"Excluded {len(rest)} read pairs  with classification containing 'Listeria_mono' from output files" in r.stderr


def test_truncated_bxi_is_rejected(tmp_path):
    a = tmp_path / "a.bxi"
    bxi_py.write_bxi(a, 100, 2, 5, {0: "x"}, [3, 7], [[1], [1]], {"x": 4})
    data = a.read_bytes()
    a.write_bytes(data[:-5])
    r = cli("info", "-b", a)
    assert r.returncode == 101 and "deserialize" in r.stderr


def test_counts_file_order_is_fnv_hashmap_order():
    rng = np.random.default_rng(11)
    for n in (1, 3, 4, 7, 8, 15, 29, 60):
        keys = ["reject"] + [f"Genus_species_{int(x)}" for x in rng.choice(10**6, size=n - 1, replace=False)]
        order, _ = O.hashset_str_order([k.encode() for k in keys], 16, True)
        assert host("map_order", *keys) == [keys[i] for i in order]


def test_cli_flag_errors():
    r = cli("build", "-b", "x")
    assert r.returncode == 101 and "required arguments" in r.stderr
    r = cli("search", "-b", "x.bxi", "-q", "q.fa", "--bogus")
    assert r.returncode == 101 and "wasn't expected" in r.stderr


# ----------------------------------------------------------------------------- the host layer's own gzip decoder
def _lines_report(path, zlib_only=False, gz_threads=None, gz_span=None):
    env = dict(os.environ)
    if zlib_only:
        env["COLORID_B200_ZLIB"] = "1"
    if gz_threads is not None:
        env["COLORID_B200_GZ_THREADS"] = str(gz_threads)
    if gz_span is not None:
        env["COLORID_B200_GZ_SPAN"] = str(gz_span)
    r = subprocess.run([CLI, "_host", "lines", str(path)], capture_output=True, text=True, env=env)
    if r.returncode != 0:
        return r.returncode, r.stderr
    d = dict(l.split("\t") for l in r.stdout.split("\n") if l.count("\t") == 1 and l.split("\t")[0] in ("lines", "bytes", "crc32"))
    return 0, (int(d["lines"]), int(d["bytes"]), int(d["crc32"], 16))


def _fastq_blob(rng, n, rl=150):
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    seq = lut[rng.integers(0, 4, (n, rl))]
    q = np.full((n, rl), 73, np.uint8)
    q[rng.random((n, rl)) < 0.05] = 50
    ids = np.frombuffer(b"".join(b"@read%07d/1\n" % i for i in range(n)), dtype=np.uint8).reshape(n, -1)
    nl = np.full((n, 1), 10, np.uint8)
    plus = np.tile(np.frombuffer(b"+\n", dtype=np.uint8), (n, 1))
    return np.concatenate([ids, seq, nl, plus, q, nl], axis=1).tobytes()


def test_gzip_decoder_equals_zlib(tmp_path):
    """fast_inflate.cpp (the mapped-file gzip decoder behind LineReader) against zlib on every block type and framing:
    dynamic / fixed / stored blocks, Huffman-only streams, long runs (distance 1), periods 2..7, tiny blocks, several members
    (an empty one among them), trailing garbage; and through COLORID_B200_ZLIB=1 the gzread path gives the same lines."""
    import zlib
    rng = np.random.default_rng(0xC0101D05)
    fq = _fastq_blob(rng, 6000)
    rb = rng.integers(0, 256, 200_000, dtype=np.uint8).tobytes()

    def z(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, memlevel=8):
        co = zlib.compressobj(level, zlib.DEFLATED, 31, memlevel, strategy)
        return co.compress(data) + co.flush()

    periodic = b"".join((bytes([65 + i]) * 1 + b"CG"[:p - 1]) * 400 + b"\n" for i, p in enumerate([2, 3, 4, 5, 6, 7, 2, 3]))
    cases = {
        "l1": (z(fq, 1), fq), "l6": (z(fq, 6), fq), "l9": (z(fq[:400_000], 9), fq[:400_000]),
        "stored": (z(rb, 0), rb), "incompressible": (z(rb, 6), rb),
        "fixed": (z(fq[:300_000], 6, zlib.Z_FIXED), fq[:300_000]), "huffman_only": (z(fq[:300_000], 6, zlib.Z_HUFFMAN_ONLY), fq[:300_000]),
        "runs": (z(b"I" * 3_000_000 + b"\n" + b"\0" * 70_000, 9), b"I" * 3_000_000 + b"\n" + b"\0" * 70_000),
        "periods": (z(periodic, 9), periodic), "tiny_blocks": (z(fq, 1, memlevel=1), fq),
        "members": (z(fq[:50_000], 6) + z(b"", 6) + z(fq[50_000:90_000], 1) + z(b"no newline at the end", 9), fq[:90_000] + b"no newline at the end"),
        "garbage_after": (z(b"hello\nworld\n", 6) + b"\0\0\0\0 not gzip", b"hello\nworld\n"),
        "big": (z(fq * 9, 1), fq * 9),            # several 4 MB output chunks: history carried across them
    }
    # empty stored blocks in mid-stream (Z_SYNC_FLUSH / Z_FULL_FLUSH: pigz, bgzip-like writers) and a header with every
    # optional field (FEXTRA as BGZF writes it, FNAME, FCOMMENT, FHCRC)
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    parts = [co.compress(fq[:100_000]), co.flush(zlib.Z_SYNC_FLUSH), co.compress(fq[100_000:250_000]), co.flush(zlib.Z_FULL_FLUSH),
             co.compress(fq[250_000:300_000]), co.flush()]
    cases["sync_flushes"] = (b"".join(parts), fq[:300_000])
    raw = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = raw.compress(fq[:120_000]) + raw.flush()
    hdr = bytes([0x1f, 0x8b, 8, 0x02 | 0x04 | 0x08 | 0x10, 0, 0, 0, 0, 0, 3]) + (6).to_bytes(2, "little") + b"BC\x02\x00\x00\x00" + b"name.fq\0" + b"a comment\0"
    hdr += (zlib.crc32(hdr) & 0xFFFF).to_bytes(2, "little")
    cases["all_header_fields"] = (hdr + body + (zlib.crc32(fq[:120_000]) & 0xFFFFFFFF).to_bytes(4, "little") + (120_000).to_bytes(4, "little"), fq[:120_000])
    for name, (gz, data) in cases.items():
        p = tmp_path / f"{name}.gz"
        p.write_bytes(gz)
        want = (data.count(b"\n") + (1 if data and not data.endswith(b"\n") else 0), len(data), zlib.crc32(data) & 0xFFFFFFFF)
        assert _lines_report(p) == (0, want), name
        assert _lines_report(p, zlib_only=True) == (0, want), name


def test_gzip_decoder_fails_loudly_on_damage(tmp_path):
    """A flipped byte in the middle, a truncated file and a wrong CRC in the trailer all end in an error, never in wrong data."""
    import zlib
    rng = np.random.default_rng(0xC0101D06)
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    gz = co.compress(_fastq_blob(rng, 3000)) + co.flush()
    damaged = {"flip": bytearray(gz), "trunc": bytearray(gz[:len(gz) // 2]), "crc": bytearray(gz)}
    damaged["flip"][len(gz) // 2] ^= 0x55
    damaged["crc"][-6] ^= 1
    for name, blob in damaged.items():
        p = tmp_path / f"{name}.gz"
        p.write_bytes(bytes(blob))
        rc, msg = _lines_report(p)
        assert rc != 0 and "gzip stream" in msg, (name, rc, msg)


def test_gzip_decoder_fuzz_against_zlib(tmp_path):
    """Seeded differential fuzzing of the decoder: texts of mixed entropy (random bytes, small alphabets, repeats, runs),
    every compression level / strategy / window size / memory level zlib offers, members concatenated at random."""
    import zlib
    rng = np.random.default_rng(0xC0101D07)

    def blob():
        parts = []
        for _ in range(int(rng.integers(1, 6))):
            kind = int(rng.integers(0, 5))
            n = int(rng.integers(1, 60_000))
            if kind == 0:
                parts.append(rng.integers(0, 256, n, dtype=np.uint8).tobytes())
            elif kind == 1:
                parts.append(np.frombuffer(b"ACGT\n", dtype=np.uint8)[rng.integers(0, 5, n)].tobytes())
            elif kind == 2:
                unit = rng.integers(33, 80, int(rng.integers(1, 40)), dtype=np.uint8).tobytes()
                parts.append(unit * (n // len(unit) + 1))
            elif kind == 3:
                parts.append(bytes([int(rng.integers(0, 256))]) * n)
            else:
                parts.append(_fastq_blob(rng, max(1, n // 330)))
        return b"".join(parts)

    for it in range(40):
        members, data = [], b""
        for _ in range(int(rng.integers(1, 4))):
            d = blob() if rng.random() > 0.1 else b""
            co = zlib.compressobj(int(rng.integers(0, 10)), zlib.DEFLATED, 16 + int(rng.integers(9, 16)), int(rng.integers(1, 10)),
                                  int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED])))
            cut = int(rng.integers(0, len(d) + 1))
            members.append(co.compress(d[:cut]) + (co.flush(zlib.Z_SYNC_FLUSH) if rng.random() < 0.5 else b"") + co.compress(d[cut:]) + co.flush())
            data += d
        p = tmp_path / f"f{it}.gz"
        p.write_bytes(b"".join(members))
        if len(p.read_bytes()) < 18:
            continue
        want = (data.count(b"\n") + (1 if data and not data.endswith(b"\n") else 0), len(data), zlib.crc32(data) & 0xFFFFFFFF)
        assert _lines_report(p) == (0, want), it
        assert _lines_report(p, gz_threads=3, gz_span=4096) == (0, want), ("parallel", it)      # (files of 16 KB and more)


def _gz(data, level=6, strategy=None, memlevel=8):
    import zlib
    co = zlib.compressobj(level, zlib.DEFLATED, 31, memlevel, zlib.Z_DEFAULT_STRATEGY if strategy is None else strategy)
    return co.compress(data) + co.flush()


def test_parallel_gzip_decoder_equals_zlib(tmp_path):
    """par_inflate.cpp (one gzip stream on several threads: block headers found by search and confirmed by the span before,
    unknown window bytes carried as 16-bit markers) against zlib.  Spans of 4 KB .. 64 KB make small files take every path:
    chains of 2 .. 5 spans (tiny blocks), spans without a block header (blocks larger than a span, stored and fixed blocks),
    a member younger than one window when its second span is resolved, members ending inside a round, runs of small members
    (handed to the sequential decoder), a block that inflates beyond the span buffers (handed over mid-stream), bytes after
    the last member.  COLORID_B200_GZ_THREADS=1 is the sequential decoder on the same files."""
    import zlib
    rng = np.random.default_rng(0xC0101D08)
    fq = _fastq_blob(rng, 12000)
    rb = rng.integers(0, 256, 600_000, dtype=np.uint8).tobytes()
    runs = b"I" * 30_000_000 + b"\n" + fq[:200_000] + b"\0" * 7_000_000
    mixed = fq[:700_000] + rb + fq[700_000:] + b"A" * 500_000 + rb[:1000]
    cases = {
        "l1": (_gz(fq, 1), fq), "l6": (_gz(fq, 6), fq), "l9": (_gz(fq, 9), fq),
        "stored": (_gz(rb, 0), rb), "incompressible": (_gz(rb, 6), rb),
        "fixed": (_gz(fq, 6, zlib.Z_FIXED), fq), "huffman_only": (_gz(fq, 6, zlib.Z_HUFFMAN_ONLY), fq),
        "runs": (_gz(runs, 9), runs), "tiny_blocks": (_gz(fq, 1, memlevel=1), fq),
        "members_big": (_gz(fq, 6) + _gz(fq[:1_000_000], 1) + _gz(b"", 6) + _gz(fq[5:], 9), fq + fq[:1_000_000] + fq[5:]),
        "members_small": (b"".join(_gz(fq[i:i + 20_000], 6) for i in range(0, len(fq), 20_000)), fq),
        "mixed": (_gz(mixed, 6), mixed), "garbage_after": (_gz(fq, 6) + b"\0\0\0 trailing", fq),
    }
    for name, (gz, data) in cases.items():
        p = tmp_path / f"{name}.gz"
        p.write_bytes(gz)
        want = (data.count(b"\n") + (1 if data and not data.endswith(b"\n") else 0), len(data), zlib.crc32(data) & 0xFFFFFFFF)
        for threads, span in ((2, 4096), (3, 16384), (5, 65536), (1, None)):
            assert _lines_report(p, gz_threads=threads, gz_span=span) == (0, want), (name, threads, span)


def test_parallel_gzip_decoder_fails_loudly_on_damage(tmp_path):
    """Single flipped bits anywhere in a stream, truncation and a wrong trailer: an error, never different data -- a span that
    started on a false block header is dropped by the chain check, real damage is reported by the sequential decoder that
    takes over at the last verified block."""
    rng = np.random.default_rng(0xC0101D09)
    gz = _gz(_fastq_blob(rng, 12000), 1, memlevel=2)
    for trial in range(24):
        b = bytearray(gz)
        pos = int(rng.integers(20, len(b) - 8))
        b[pos] ^= 1 << int(rng.integers(0, 8))
        p = tmp_path / "flip.gz"
        p.write_bytes(bytes(b))
        rc, msg = _lines_report(p, gz_threads=3, gz_span=16384)
        assert rc != 0 and "gzip stream" in msg, (trial, pos, rc, msg)
    for cut in (len(gz) // 3, len(gz) * 2 // 3, len(gz) - 9, len(gz) - 4):
        p = tmp_path / "trunc.gz"
        p.write_bytes(gz[:cut])
        rc, msg = _lines_report(p, gz_threads=3, gz_span=16384)
        assert rc != 0 and "gzip stream" in msg, (cut, rc, msg)
    b = bytearray(gz)
    b[-6] ^= 1
    p = tmp_path / "crc.gz"
    p.write_bytes(bytes(b))
    rc, msg = _lines_report(p, gz_threads=3, gz_span=16384)
    assert rc != 0 and "CRC mismatch" in msg, (rc, msg)


def _parse_digest(files, quality, slow, **env_extra):
    env = dict(os.environ, **{k: str(v) for k, v in env_extra.items()})
    env["COLORID_B200_PARSE_SLOW"] = "1" if slow else "0"
    r = subprocess.run([CLI, "_host", "fastq_parse", str(quality)] + [str(f) for f in files], capture_output=True, text=True, env=env)
    if r.returncode != 0:
        return r.returncode, r.stderr.strip().split("\n")[-1]
    d = dict(l.split("\t") for l in r.stdout.split("\n") if l.count("\t") == 1 and l.split("\t")[0] in ("reads", "batches", "crc32"))
    return 0, (int(d["reads"]), int(d["batches"]), d["crc32"])


def test_read_id_record_loop_blockwise_equals_linewise(tmp_path):
    """read_id's FASTQ record loop (drivers.cpp parse_fastq_records) takes whole stretches of records out of the readers' line
    blocks and copies bases and qualities on several threads; COLORID_B200_PARSE_SLOW=1 is the line-by-line loop of
    per_read_stream_pe / _se (read_id_mt_pe.rs:701-951).  `_host fastq_parse` digests every batch either way (bases,
    qualities, offsets, ids, batch boundaries): equal on paired and single files, quality masking on and off, several line
    blocks, a second file that ends early (at a record boundary and inside a record), a shorter first file, records whose
    quality line is shorter than the sequence (qual_mask on the host), CRLF files, reads long enough to fill a batch by
    bytes, and .gz input through the multi-threaded decoder."""
    import zlib
    rng = np.random.default_rng(0xC0101D0A)

    def fastq(n, rl, tag, crlf=False, short_qual_every=0):
        lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
        out = []
        for i in range(n):
            L = int(rl if isinstance(rl, int) else rng.integers(rl[0], rl[1]))
            seq = lut[rng.integers(0, 5, L)].tobytes()
            q = rng.integers(35, 74, L, dtype=np.uint8).tobytes()
            if short_qual_every and i % short_qual_every == short_qual_every - 1:
                q = q[: max(1, L - 7)]
            out.append(b"@%s%07d\n" % (tag, i) + seq + b"\n+\n" + q + b"\n")
        blob = b"".join(out)
        return blob.replace(b"\n", b"\r\n") if crlf else blob

    n = 40_000                                   # 2.4 line blocks of 65,536 lines
    r1, r2 = fastq(n, (30, 200), b"a"), fastq(n, (30, 200), b"b")
    cases = {}

    def put(name, files, quality=15, **env):
        paths = []
        for j, blob in enumerate(files):
            p = tmp_path / f"{name}_{j}.fastq"
            p.write_bytes(blob)
            paths.append(p)
        cases[name] = (paths, quality, env)

    put("pe", [r1, r2])
    put("pe_q0", [r1, r2], quality=0)
    put("se", [r1])
    lines2 = r2.split(b"\n")
    for cut in (4 * 17_000, 4 * 17_000 + 1, 4 * 17_000 + 2, 4 * 17_000 + 3, 4 * 16_384, 4 * 16_384 + 2):
        put(f"pe_short2_{cut}", [r1, b"\n".join(lines2[:cut]) + b"\n"])
    put("pe_short1", [b"\n".join(r1.split(b"\n")[: 4 * 20_001 + 2]) + b"\n", r2])
    put("pe_qual_mask", [fastq(n, (30, 200), b"a", short_qual_every=1000), fastq(n, (30, 200), b"b", short_qual_every=777)])
    put("pe_crlf", [fastq(20_000, 100, b"a", crlf=True), fastq(20_000, 100, b"b", crlf=True)])
    put("se_no_final_newline", [r1[:-1]])
    put("empty", [b"", b""])
    for name, (paths, quality, env) in cases.items():
        fast, slow = _parse_digest(paths, quality, False, **env), _parse_digest(paths, quality, True, **env)
        assert fast[0] == 0 and fast == slow, (name, fast, slow)
        assert _parse_digest(paths, quality, False, COLORID_B200_PAR_MIN=1, **env) == slow, (name, "copy threads on every stretch")
    assert _parse_digest(cases["pe"][0], 15, False)[1][0] == n and _parse_digest(cases["pe_short2_%d" % (4 * 17_000 + 2)][0], 15, False)[1][0] == 17_000
    # batches that fill up by bytes (96 MB of bases): 2 x 1,900 reads of 30 kb
    big = [fastq(1900, 30_000, t) for t in (b"a", b"b")]
    put("pe_long", big)
    fast, slow = _parse_digest(cases["pe_long"][0], 15, False), _parse_digest(cases["pe_long"][0], 15, True)
    assert fast[0] == 0 and fast == slow and fast[1][1] == 2, (fast, slow)
    # .gz through the multi-threaded decoder
    gz = []
    for j, blob in enumerate((r1, r2)):
        p = tmp_path / f"pe_{j}.fastq.gz"
        p.write_bytes(_gz(blob, 1, memlevel=2))
        gz.append(p)
    want = _parse_digest(cases["pe"][0], 15, True)
    for threads in (1, 3):
        assert _parse_digest(gz, 15, False, COLORID_B200_GZ_THREADS=threads, COLORID_B200_GZ_SPAN=65536) == want, threads


def test_block_line_splitter_equals_line_reader(tmp_path):
    """AsyncLineReader splits whole blocks with memchr (LineReader::next_lines) where LineReader::next walks line by line:
    same lines on CRLF files, blank lines, a last line without newline, lines longer than the 1 MB read buffer, lone CRs,
    more than one block of 65,536 lines -- with the EOLs stripped (BufRead::lines) and kept (stream_fasta)."""
    import zlib
    rng = np.random.default_rng(0xC0101D08)
    texts = {
        "crlf": b"".join(b"line %d\r\n" % i for i in range(1000)),
        "mixed": b"a\n\r\nb\r\n\n\nc\rd\ne\r",                       # blank lines, CR inside a line, a last line ending in CR
        "no_final_newline": b"x\ny\nzzz",
        "empty": b"",
        "only_newlines": b"\n" * 70_000,
        "long_lines": b"".join(bytes([65 + i % 26]) * n + b"\n" for i, n in enumerate([3_000_000, 1, 1_048_576, 1_048_575, 0, 2_500_000])),
        "many_blocks": b"".join(b"%d\n" % i for i in range(200_000)),
        "fastq": _fastq_blob(rng, 40_000),
    }
    for name, data in texts.items():
        for gz in (False, True):
            p = tmp_path / (name + (".gz" if gz else ".txt"))
            p.write_bytes(zlib.compress(data, 6, 31) if gz else data)
            for op_a, op_b in ((("alines", p, 1), ("lines", p)), (("alines", p, 0), ("lines1", p))):
                ra = subprocess.run([CLI, "_host", *map(str, op_a)], capture_output=True, text=True)
                rb = subprocess.run([CLI, "_host", *map(str, op_b)], capture_output=True, text=True)
                assert ra.returncode == 0 and rb.returncode == 0, (name, gz, ra.stderr, rb.stderr)
                pick = lambda out: [l for l in out.split("\n") if l.split("\t")[0] in ("lines", "bytes", "crc32")]
                assert pick(ra.stdout) == pick(rb.stdout) and len(pick(ra.stdout)) == 3, (name, gz, op_a[0], op_a[-1])
