"""Independent Python reader/writer of the reference's .bxi format (bigsi.rs:19-27,51-69 in bincode 1.x
defaults) -- test infrastructure used to check the C++ host layer's bxi.cpp."""
import struct

import numpy as np


def write_bxi(path, bloom_size, num_hash, k_size, colors, row_ids, words, n_ref, row_order=None, m_size=None):
    """colors: {colour: name}; row_ids: [n]; words: [n, W] uint32; n_ref: {name: count}; m_size: write an .mxi."""
    N = len(colors)
    W = (N + 31) // 32
    words = np.asarray(words, dtype=np.uint32).reshape(len(row_ids), W)
    with open(path, "wb") as f:
        f.write(struct.pack("<QQQ", bloom_size, num_hash, k_size))
        if m_size is not None:
            f.write(struct.pack("<Q", m_size))
        f.write(struct.pack("<Q", N))
        for c, name in colors.items():
            b = name.encode()
            f.write(struct.pack("<QQ", c, len(b)) + b)
        f.write(struct.pack("<Q", len(row_ids)))
        order = range(len(row_ids)) if row_order is None else row_order
        for i in order:
            f.write(struct.pack("<QQ", int(row_ids[i]), W) + words[i].astype("<u4").tobytes() + struct.pack("<Q", N))
        f.write(struct.pack("<Q", len(n_ref)))
        for name, n in n_ref.items():
            b = name.encode()
            f.write(struct.pack("<Q", len(b)) + b + struct.pack("<Q", n))


def read_bxi(path, mini=False):
    """mini=True: .mxi (bigsi.rs:40-49 BigsyMapMiniNew: m_size follows k_size)."""
    d = open(path, "rb").read()
    at = 0

    def u64():
        nonlocal at
        v = struct.unpack_from("<Q", d, at)[0]
        at += 8
        return v

    def s():
        nonlocal at
        n = u64()
        v = d[at:at + n].decode()
        at += n
        return v

    out = dict(bloom_size=u64(), num_hash=u64(), k_size=u64())
    if mini:
        out["m_size"] = u64()
    colors = {}
    for _ in range(u64()):
        c = u64()
        colors[c] = s()
    N = len(colors)
    W = (N + 31) // 32
    nrows = u64()
    rec = np.dtype([("row", "<u8"), ("nw", "<u8"), ("w", "<u4", (W,)), ("nbits", "<u8")])
    rows = np.frombuffer(d, dtype=rec, count=nrows, offset=at)
    at += nrows * rec.itemsize
    assert nrows == 0 or (np.all(rows["nw"] == W) and np.all(rows["nbits"] == N))
    n_ref = {}
    for _ in range(u64()):
        name = s()
        n_ref[name] = u64()
    assert at == len(d), "trailing bytes"
    out.update(colors=colors, row_ids=rows["row"].copy(), words=rows["w"].reshape(nrows, W).copy(), n_ref=n_ref)
    return out
