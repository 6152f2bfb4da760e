"""cid_pack_reads (host code of the library: AVX2 and scalar paths) against a numpy restatement of seq.rs:36-56 qual_mask +
the plane layout of include/colorid_b200.h.  No GPU needed."""
import numpy as np
import pytest

import colorid_b200 as cb
from tests import synth


def planes_ref(read, quals, qual_offset):
    s = b"".join(read)
    a = np.frombuffer(s, np.uint8).copy()
    if quals is not None and qual_offset:
        q = np.frombuffer(b"".join(quals), np.uint8)
        a[q < qual_offset + 33] = ord("N")
    L = len(a)
    u = a & 0xDF
    good = (u == 65) | (u == 67) | (u == 71) | (u == 84)
    code = (((a >> 1) ^ (a >> 2)) & 3).astype(np.uint32) * good
    codes = np.zeros((L + 15) // 16, np.uint32)
    for i in range(L):
        codes[i >> 4] |= code[i] << np.uint32(30 - 2 * (i & 15))
    nb = (L + 31) // 32
    bad = np.zeros(nb, np.uint32)
    low = np.zeros(nb, np.uint32)
    for i in range(nb * 32):
        if i >= L or not good[i]:
            bad[i >> 5] |= np.uint32(1) << np.uint32(i & 31)
        elif a[i] & 0x20:
            low[i >> 5] |= np.uint32(1) << np.uint32(i & 31)
    return codes, bad, low


@pytest.mark.parametrize("with_lower", [False, True])
@pytest.mark.parametrize("threads", [1, 3])
def test_pack_reads_planes(with_lower, threads):
    rng = np.random.default_rng(77)
    reads, quals = [], []
    for i in range(300):
        n1, n2 = int(rng.integers(0, 200)), int(rng.integers(0, 200))
        r = [synth.sprinkle(rng, synth.rand_seq(rng, n1), b"NRYn-", 0.03), synth.rand_seq(rng, n2)]
        if with_lower and i % 9 == 0:
            r[1] = synth.sprinkle(rng, r[1], b"acgt", 0.2)
        if i % 50 == 7:
            r = [r[0]]
        reads.append(r)
        quals.append([bytes(rng.choice(np.frombuffer(b"#+5?I", np.uint8), size=len(m)).tolist()) for m in r])
    reads += [[b""], [b"A"], [synth.rand_seq(rng, 32)], [synth.rand_seq(rng, 64), synth.rand_seq(rng, 33)], [synth.rand_seq(rng, 5000)]]
    quals += [[b""], [b"I"], [b"I" * 32], [b"#" * 64, b"I" * 33], [b"I" * 5000]]
    for qo, qs in ((0, None), (15, quals), (40, quals)):
        p = cb.pack_reads(reads, qs, qo, threads)
        assert bool(p["flags"] & cb.lib.CID_PACK_LOWER) == with_lower
        for r, read in enumerate(reads):
            codes, bad, low = planes_ref(read, qs[r] if qs else None, qo)
            w = p["words"][int(p["word_offs"][r]):int(p["word_offs"][r + 1])]
            exp = np.concatenate([codes, bad, low]) if with_lower else np.concatenate([codes, bad])
            assert np.array_equal(w, exp), (r, qo)
