#!/usr/bin/env python
"""CLI end-to-end timing (SURVEY §8d timed region 3: gz files in -> output files out) on synthetic C1/C2-shaped data:
46 genomes x 3.3 Mbp as FASTA -> `colorid-b200 build`; N read pairs as R1/R2 .fastq.gz -> `read_id` and `search`.
Prints one JSON line.  Profiling aid, not the bench contract."""
import json, os, subprocess, sys, tempfile, time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "colorid_b200", "colorid-b200")
n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_acc, glen, rl = 46, 3_300_000, 150
rng = np.random.default_rng(0xC0101D01)
lut = np.frombuffer(b"ACGT", dtype=np.uint8)
d = tempfile.mkdtemp(prefix="cid_cli_")
roots = rng.integers(0, 4, size=(8, glen), dtype=np.uint8)
genomes = []
refs = []
for a in range(n_acc):
    g = roots[a % 8].copy()
    m = rng.random(glen) < 0.01
    g[m] = rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)
    genomes.append(g)
    asc = lut[g]
    with open(f"{d}/acc{a:02d}.fasta", "wb") as f:
        f.write(b">acc%02d\n" % a)
        pad = (-glen) % 80
        body = np.concatenate([asc, np.full(pad, ord("A"), np.uint8)]).reshape(-1, 80)
        lines = np.concatenate([body, np.full((body.shape[0], 1), 10, np.uint8)], axis=1).reshape(-1)
        f.write(lines.tobytes()[: -(pad + 1)] + b"\n" if pad else lines.tobytes())
    refs.append(f"acc{a:02d}\t{d}/acc{a:02d}.fasta")
open(f"{d}/refs.tsv", "w").write("\n".join(refs) + "\n")


def write_fastq_gz(path, seqs, tag):
    n = seqs.shape[0]
    ids = np.char.add(np.char.add("@" + tag, np.char.zfill(np.arange(n).astype(str), 9)), "\n")
    idb = np.frombuffer("".join(ids.tolist()).encode(), dtype=np.uint8).reshape(n, -1)
    q = np.full((n, rl), 73, np.uint8)
    q[rng.random((n, rl)) < 0.01] = 35
    nl = np.full((n, 1), 10, np.uint8)
    plus = np.tile(np.frombuffer(b"+\n", dtype=np.uint8), (n, 1))
    rec = np.concatenate([idb, lut[seqs], nl, plus, q, nl], axis=1)
    p = subprocess.Popen(["gzip", "-1", "-c"], stdin=subprocess.PIPE, stdout=open(path, "wb"))
    p.stdin.write(rec.tobytes())
    p.stdin.close()
    p.wait()
    return rec.nbytes


gi = rng.integers(0, n_acc, n_pairs)
pos = rng.integers(0, glen - 400, n_pairs)
ins = rng.integers(300, 401, n_pairs)
G = np.stack(genomes)
ar = np.arange(rl)
m1 = G[gi[:, None], pos[:, None] + ar[None, :]]
m2 = 3 - G[gi[:, None], (pos + ins - 1)[:, None] - ar[None, :]]
rnd = rng.random(n_pairs) < 0.3
m1[rnd] = rng.integers(0, 4, size=(int(rnd.sum()), rl), dtype=np.uint8)
m2[rnd] = rng.integers(0, 4, size=(int(rnd.sum()), rl), dtype=np.uint8)
for m in (m1, m2):
    e = rng.random(m.shape) < 0.005
    m[e] = rng.integers(0, 4, size=int(e.sum()), dtype=np.uint8)
raw = write_fastq_gz(f"{d}/r_1.fastq.gz", m1, "r") + write_fastq_gz(f"{d}/r_2.fastq.gz", m2, "r")


def timed(*args):
    t = time.perf_counter()
    r = subprocess.run([CLI, *args], capture_output=True, text=True, env=dict(os.environ, COLORID_B200_TRACE="1"))
    dt = time.perf_counter() - t
    assert r.returncode == 0, r.stderr[-1000:]
    print(args[0], "\n".join(l for l in r.stderr.split("\n") if l.startswith("[trace]")), file=sys.stderr)
    return dt, r


t_build, _ = timed("build", "-b", f"{d}/idx", "-r", f"{d}/refs.tsv", "-k", "31", "-n", "4", "-s", "50000000")
t_read, r = timed("read_id", "-b", f"{d}/idx.bxi", "-q", f"{d}/r_1.fastq.gz", f"{d}/r_2.fastq.gz", "-n", f"{d}/out")
n_lines = sum(1 for _ in open(f"{d}/out_reads.txt"))
t_search, rs = timed("search", "-b", f"{d}/idx.bxi", "-q", f"{d}/r_1.fastq.gz", "-r", f"{d}/r_2.fastq.gz")
t_info, _ = timed("info", "-b", f"{d}/idx.bxi")
# batch_id: four samples (the same pair of files under four names), index read and uploaded once
n_samples = 4
with open(f"{d}/samples.tsv", "w") as f:
    for i in range(n_samples):
        f.write(f"{d}/sample{i}\t{d}/r_1.fastq.gz\t{d}/r_2.fastq.gz\n")
t_batch, _ = timed("batch_id", "-b", f"{d}/idx.bxi", "-q", f"{d}/samples.tsv", "-T", "b")
same = all(open(f"{d}/sample{i}_b_reads.txt").read() == open(f"{d}/out_reads.txt").read() for i in range(n_samples))
# minimizer index of the same references (-m -v 15) and read_id on it
t_build_mxi, _ = timed("build", "-b", f"{d}/idx", "-r", f"{d}/refs.tsv", "-k", "31", "-n", "4", "-s", "50000000", "-m")
t_read_mxi, _ = timed("read_id", "-b", f"{d}/idx.mxi", "-q", f"{d}/r_1.fastq.gz", f"{d}/r_2.fastq.gz", "-n", f"{d}/outm")
print(json.dumps({"batch_id_samples": n_samples, "batch_id_seconds": t_batch, "batch_id_pairs_per_s": n_samples * n_pairs / t_batch,
                  "batch_id_outputs_equal_read_id": same, "build_mxi_seconds": t_build_mxi, "mxi_bytes": os.path.getsize(f"{d}/idx.mxi"),
                  "read_id_mxi_seconds": t_read_mxi,
                  "workload": f"CLI end to end, {n_acc} x {glen} bp FASTA index (k=31 S=50M H=4), {n_pairs} read pairs of 2x{rl} bp as .fastq.gz",
                  "build_seconds": t_build, "build_gbp_per_s": n_acc * glen / t_build / 1e9,
                  "bxi_bytes": os.path.getsize(f"{d}/idx.bxi"), "index_load_seconds_(info)": t_info,
                  "read_id_seconds": t_read, "read_id_pairs_per_s_incl_index_load": n_pairs / t_read,
                  "read_id_pairs_per_s_excl_index_load": n_pairs / max(t_read - t_info, 1e-9), "read_id_lines": n_lines,
                  "search_fastq_seconds": t_search, "search_report_lines": len(rs.stdout.strip().split("\n")) - 2,
                  "fastq_raw_bytes": raw, "host_cores": os.cpu_count()}))
