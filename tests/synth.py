"""Seeded synthetic sequence data for the parity tests (numpy only)."""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.zeros(256, dtype=np.uint8)
COMP[:] = ord("N")
for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
    COMP[a] = b


def rand_seq(rng, n):
    return ACGT[rng.integers(0, 4, size=n)].tobytes()


def mutate(rng, seq, rate):
    a = np.frombuffer(seq, dtype=np.uint8).copy()
    m = rng.random(a.size) < rate
    a[m] = ACGT[rng.integers(0, 4, size=int(m.sum()))]
    return a.tobytes()


def revcomp(seq):
    return COMP[np.frombuffer(seq, dtype=np.uint8)][::-1].tobytes()


def clade_genomes(rng, n_genomes, length, n_clades=3, div=0.01):
    roots = [rand_seq(rng, length) for _ in range(n_clades)]
    return [mutate(rng, roots[i % n_clades], div) for i in range(n_genomes)]


def sprinkle(rng, seq, chars, rate):
    a = np.frombuffer(seq, dtype=np.uint8).copy()
    m = rng.random(a.size) < rate
    pool = np.frombuffer(chars, dtype=np.uint8)
    a[m] = pool[rng.integers(0, pool.size, size=int(m.sum()))]
    return a.tobytes()


def reads_from(rng, genomes, n_reads, read_len=150, insert=350, err=0.005, frac_random=0.3, paired=True, n_rate=0.0):
    reads = []
    for _ in range(n_reads):
        if rng.random() < frac_random:
            frag = rand_seq(rng, insert)
        else:
            g = genomes[rng.integers(0, len(genomes))]
            s = int(rng.integers(0, max(1, len(g) - insert)))
            frag = g[s:s + insert]
        if rng.random() < 0.5:
            frag = revcomp(frag)
        m1 = mutate(rng, frag[:read_len], err)
        m2 = mutate(rng, revcomp(frag)[:read_len], err)
        if n_rate:
            m1 = sprinkle(rng, m1, b"N", n_rate)
            m2 = sprinkle(rng, m2, b"N", n_rate)
        reads.append([m1, m2] if paired else [m1])
    return reads
