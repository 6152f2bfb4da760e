"""GPU, two processes (one per GPU; both on GPU 0 when the box has a single one): a column-sharded index whose per-query
counts are exchanged by the gather kernel itself -- each rank stores its column slice into every rank's full-width
result through CUDA-IPC peer memory (cid_query_counts_sharded_dev) -- against the oracle's counts on the whole index."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import synth

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ndev, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)       # plumbing only: handles and barriers
    import ctypes as C
    import colorid_b200 as cb
    from colorid_b200 import lib as L
    from colorid_b200 import sharding
    from oracle import pyoracle as O
    devno = rank % ndev
    torch.cuda.set_device(devno)
    dev = torch.device(f"cuda:{devno}")
    rng = np.random.default_rng(777)                                   # same data on every rank
    N, k, S, H = 300, 21, 200_003, 2                                   # 300 accessions -> 160 + 140: rows of 8 and 8 words
    genomes = synth.clade_genomes(rng, N, 2500, n_clades=6, div=0.01)
    queries = [[genomes[int(rng.integers(0, N))][100:100 + int(rng.integers(k, 1500))]] for _ in range(40)]
    queries += [[synth.rand_seq(rng, 700)], [b"ACG"], [], [genomes[5][:900], genomes[200][:900]]]
    queries.append([b"".join(genomes[:5])])                            # > 8192 positions: count-table path, several units (atomics)
    shards = sharding.column_shards(N, world)
    lo, hi = shards[rank]
    ctx = cb.Context(devno)
    gix = cb.Index(ctx, S, H, k, hi - lo)
    for c in range(lo, hi):
        gix.build_accession(c - lo, [genomes[c]])
    gix.finalize()
    flat = [s for q in queries for s in q]
    bases, offs = cb.pack_seqs(flat)
    qoffs = cb.group_offsets(queries)
    nq = len(queries)
    d_bases = torch.from_numpy(bases).to(dev)
    d_offs = torch.from_numpy(offs.view(np.int64)).to(dev)
    d_qoffs = torch.from_numpy(qoffs.view(np.int64)).to(dev)
    d_nk = torch.zeros(nq, device=dev, dtype=torch.int64)
    P = lambda a: a.ctypes.data_as(L.u64p)
    results = []
    for subset in (slice(0, nq - 1), slice(0, nq)):                     # all small (plain stores), then with the large query (atomics)
        n = len(range(*subset.indices(nq)))
        pc = sharding.PeerCounts(ctx, n, N, dev)
        for _ in range(2):                                              # a second pass over the same buffers
            pc.begin_pass()
            L.check(ctx.lib.cid_query_counts_sharded_dev(gix.h, d_bases.data_ptr(), d_offs.data_ptr(), len(flat), int(offs[-1]),
                                                         d_qoffs.data_ptr(), P(qoffs), P(offs), n, 0, pc.dest, world, N, lo,
                                                         d_nk.data_ptr(), torch.cuda.current_stream().cuda_stream))
            full = pc.end_pass()
        results.append((full.cpu().numpy().astype(np.uint32).copy(), d_nk.cpu().numpy()[:n].astype(np.uint64).copy()))
        pc.close()
    if rank == 0:
        whole = O.Index(S, H, k, N)
        whole.build_many([[g] for g in genomes], O.MODE_FASTA, threads=2)
        o = whole.query_counts(queries, O.MODE_FASTA, True, 0)
        ok = all(np.array_equal(c, o["counts"][: len(c)]) and np.array_equal(nk, o["num_kmers"][: len(nk)]) for c, nk in results)
        open(os.path.join(out_dir, "ok"), "w").write("1" if ok and int(o["counts"].max()) > 300 else "0")
    both = torch.tensor(results[1][0].astype(np.int64))
    parts = [torch.empty_like(both) for _ in range(world)]
    dist.all_gather(parts, both)
    assert all(torch.equal(p, both) for p in parts), "ranks disagree on the exchanged counts"
    dist.destroy_process_group()


def test_column_sharded_counts_exchanged_through_peer_memory(tmp_path):
    ndev = torch.cuda.device_count()
    assert ndev >= 1
    mp.spawn(_worker, args=(2, _free_port(), ndev, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok").read() == "1"
