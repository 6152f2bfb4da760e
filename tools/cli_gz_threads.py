#!/usr/bin/env python
"""CLI `batch_id` on .fastq.gz pairs against the number of inflating threads per file (COLORID_B200_GZ_THREADS; 0 = default), and the bare
decoder (`_host gunzip`) beside it.  Same synthetic data as tools/cli_e2e.py (46 x 3.3 Mbp index, 2 x 150 bp pairs, gzip -1).
Prints one JSON line; the CLI's [trace] lines go to stderr.  Profiling aid, not the bench contract."""
import hashlib, json, os, re, subprocess, sys, tempfile, time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "colorid_b200", "colorid-b200")
n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
threads = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,4,6").split(",")]
n_acc, glen, rl = 46, 3_300_000, 150
rng = np.random.default_rng(0xC0101D01)
lut = np.frombuffer(b"ACGT", dtype=np.uint8)
d = tempfile.mkdtemp(prefix="cid_gz_")
t0 = time.perf_counter()
roots = rng.integers(0, 4, size=(8, glen), dtype=np.uint8)
genomes, refs = [], []
for a in range(n_acc):
    g = roots[a % 8].copy()
    m = rng.random(glen) < 0.01
    g[m] = rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)
    genomes.append(g)
    with open(f"{d}/acc{a:02d}.fasta", "wb") as f:
        f.write(b">acc%02d\n" % a + lut[g].tobytes() + b"\n")
    refs.append(f"acc{a:02d}\t{d}/acc{a:02d}.fasta")
open(f"{d}/refs.tsv", "w").write("\n".join(refs) + "\n")
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
procs, raw = [], 0
for path, seqs in ((f"{d}/r_1.fastq.gz", m1), (f"{d}/r_2.fastq.gz", m2)):
    n = seqs.shape[0]
    idb = np.frombuffer(b"".join(b"@r%09d\n" % i for i in range(n)), dtype=np.uint8).reshape(n, -1)
    q = np.full((n, rl), 73, np.uint8)
    q[rng.random((n, rl)) < 0.01] = 35
    nl = np.full((n, 1), 10, np.uint8)
    plus = np.tile(np.frombuffer(b"+\n", dtype=np.uint8), (n, 1))
    rec = np.concatenate([idb, lut[seqs], nl, plus, q, nl], axis=1)
    raw += rec.nbytes
    open(path[:-3], "wb").write(rec.tobytes())
    procs.append(subprocess.Popen(["gzip", "-1", "-f", path[:-3]]))
for p in procs:
    p.wait()
t_gen = time.perf_counter() - t0


def cli(env_extra, *args):
    t = time.perf_counter()
    r = subprocess.run([CLI, *args], capture_output=True, text=True, env=dict(os.environ, COLORID_B200_TRACE="1", **env_extra))
    dt = time.perf_counter() - t
    assert r.returncode == 0, r.stderr[-1000:]
    return dt, r


t_build, _ = cli({}, "build", "-b", f"{d}/idx", "-r", f"{d}/refs.tsv", "-k", "31", "-n", "4", "-s", "50000000")
n_samples = 4
with open(f"{d}/samples.tsv", "w") as f:
    for i in range(n_samples):
        f.write(f"{d}/sample{i}\t{d}/r_1.fastq.gz\t{d}/r_2.fastq.gz\n")
out = {"workload": f"CLI batch_id, {n_samples} samples of {n_pairs} read pairs (2x{rl} bp, .fastq.gz, gzip -1), index {n_acc} x {glen} bp k=31 S=50M H=4",
       "host_cores": os.cpu_count(), "fastq_raw_bytes": raw, "gz_bytes": os.path.getsize(f"{d}/r_1.fastq.gz") + os.path.getsize(f"{d}/r_2.fastq.gz"),
       "generate_seconds": t_gen, "build_seconds": t_build, "batch_id": {}, "gunzip_mb_per_s": {}}
digests = set()
for T in threads:
    env = {"COLORID_B200_GZ_THREADS": str(T)} if T else {}            # 0 = the CLI's default
    dt, r = cli(env, "batch_id", "-b", f"{d}/idx.bxi", "-q", f"{d}/samples.tsv", "-T", "b")
    per = [float(m.group(1)) for m in re.finditer(r"\[trace\] parse \+ classify \+ write\s+([0-9.]+) s", r.stderr)]
    print(f"== COLORID_B200_GZ_THREADS={T}\n" + "\n".join(l for l in r.stderr.split("\n") if l.startswith("[trace]")), file=sys.stderr)
    h = hashlib.sha256()
    for i in range(n_samples):
        h.update(open(f"{d}/sample{i}_b_reads.txt", "rb").read())
        h.update(open(f"{d}/sample{i}_b_counts.txt", "rb").read())
    digests.add(h.hexdigest())
    steady = sorted(per[1:])[len(per[1:]) // 2] if len(per) > 1 else None
    out["batch_id"][str(T)] = {"seconds": dt, "per_sample_seconds": per, "steady_pairs_per_s": n_pairs / steady if steady else None}
    sys.stderr.flush()
out["outputs_identical_across_thread_counts"] = len(digests) == 1
for T in sorted(set(t for t in threads if t) | {2, 8}):
    best = None
    for _ in range(2):
        t = time.perf_counter()
        subprocess.run([CLI, "_host", "gunzip", f"{d}/r_1.fastq.gz"], capture_output=True, env=dict(os.environ, COLORID_B200_GZ_THREADS=str(T)))
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    out["gunzip_mb_per_s"][str(T)] = raw / 2 / best / 1e6
t = time.perf_counter()
subprocess.run([CLI, "_host", "gunzip", f"{d}/r_1.fastq.gz"], capture_output=True, env=dict(os.environ, COLORID_B200_ZLIB="1"))
out["gunzip_mb_per_s"]["zlib"] = raw / 2 / (time.perf_counter() - t) / 1e6
print(json.dumps(out))
