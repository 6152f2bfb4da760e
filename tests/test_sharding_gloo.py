"""CPU, world_size 2 over gloo: the multi-GPU host logic (unit slices, column shards, count gather,
row-present OR).  The per-rank compute stand-in is the oracle; on GPUs the same glue moves device tensors."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from colorid_b200 import sharding
from tests import synth


def test_slices_and_shards():
    assert sharding.unit_slices(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert sharding.unit_slices(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    sh = sharding.column_shards(10000, 8)
    assert sh[0] == (0, 1280) and sh[-1][1] == 10000 and all(lo % 32 == 0 for lo, _ in sh)
    assert sum(hi - lo for lo, hi in sh) == 10000
    assert sharding.column_shards(46, 2) == [(0, 32), (32, 46)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as O
    rng = np.random.default_rng(4242)                      # same data on every rank
    N, k, S, H = 70, 21, 100_003, 2
    genomes = synth.clade_genomes(rng, N, 3000, n_clades=5, div=0.01)
    queries = [[genomes[i][100:900]] for i in (0, 33, 69)] + [[synth.rand_seq(rng, 500)]]
    reads = synth.reads_from(rng, genomes, 40, read_len=100, insert=220)
    full = O.Index(S, H, k, N)
    full.build_many([[g] for g in genomes], O.MODE_FASTA, threads=2)

    # ---- column-sharded: this rank indexes only its accession range
    shards = sharding.column_shards(N, world)
    lo, hi = shards[rank]
    mine = O.Index(S, H, k, hi - lo)
    mine.build_many([[g] for g in genomes[lo:hi]], O.MODE_FASTA, threads=2)
    local = torch.from_numpy(mine.query_counts(queries, O.MODE_FASTA, True, 0)["counts"].astype(np.int64))
    got = sharding.gather_counts(local, shards).numpy()
    exp = full.query_counts(queries, O.MODE_FASTA, True, 0)["counts"]
    assert np.array_equal(got, exp), "gathered counts differ from the unsharded index"
    # row-present bitmap: OR over shards == bitmap of the full matrix
    present = torch.from_numpy(np.packbits(mine.words().any(axis=1), bitorder="little").astype(np.int32))
    sharding.or_reduce_bitmap(present)
    exp_present = np.packbits(full.words().any(axis=1), bitorder="little").astype(np.int32)
    assert np.array_equal(present.numpy(), exp_present)
    # perfect search: AND-rows are word-aligned slices; "a row is absent" needs the global bitmap
    p_local = mine.query_perfect(queries)
    words = sharding.gather_and_rows(torch.from_numpy(p_local["and_rows"].astype(np.int64)), shards).numpy()
    p_full = full.query_perfect(queries)
    ok = p_full["status"] == 0
    assert np.array_equal(words[ok], p_full["and_rows"][ok].astype(np.int64))

    # ---- column-sharded read_id, the exchange: every rank holds the entries of ITS colours of each read's report, tagged with
    # the step at which the colour entered the report; merge_read_reports (sparse all-gather + cid_merge_shard_reports)
    # must give back the whole report in order on every rank, and the host vote on it the whole index's classification.
    # (The per-rank compute is synthetic here: the oracle's report of the whole index, its sequence taken as the insertion
    # order, split by colour range.  The GPU tests run the real kernels per shard.)
    from colorid_b200.api import classify_reads
    whole_rep = full.read_id_batch(reads)
    nr, cap = len(reads), hi - lo + 1
    mine_rep = dict(n_set=whole_rep["n_set"], flags=np.zeros(nr, np.uint32), rep_n=np.zeros(nr, np.uint32),
                    rep_colour=np.zeros((nr, cap), np.uint32), rep_count=np.zeros((nr, cap), np.uint32))
    for r in range(nr):
        n_loc = 0
        for i in range(whole_rep["rep_n"][r]):
            c, v = int(whole_rep["rep_colour"][r, i]), int(whole_rep["rep_count"][r, i])
            if c == N:                                        # the "no hit" key: every shard reports it (as its own N)
                mine_rep["rep_colour"][r, n_loc], mine_rep["rep_count"][r, n_loc] = hi - lo, 1
                n_loc += 1
            elif lo <= c < hi:
                mine_rep["rep_colour"][r, n_loc], mine_rep["rep_count"][r, n_loc] = (c - lo) | (i << 20), v
                n_loc += 1
        mine_rep["rep_n"][r] = n_loc
    merged = sharding.merge_read_reports(mine_rep, shards, N)
    for r in range(nr):
        n_all = int(whole_rep["rep_n"][r])
        assert int(merged["rep_n"][r]) == n_all
        exp_c, exp_v = whole_rep["rep_colour"][r, :n_all].tolist(), whole_rep["rep_count"][r, :n_all].tolist()
        if N in exp_c:                                        # the merge puts the "no hit" key last, like the kernels
            j = exp_c.index(N)
            exp_c, exp_v = exp_c[:j] + exp_c[j + 1:] + [N], exp_v[:j] + exp_v[j + 1:] + [exp_v[j]]
        assert merged["rep_colour"][r, :n_all].tolist() == exp_c and merged["rep_count"][r, :n_all].tolist() == exp_v
    cls = classify_reads((S, H, N), full.n_ref, merged)
    assert np.array_equal(cls["kind"], whole_rep["kind"]) and np.array_equal(cls["hits"], whole_rep["hits"])
    assert np.array_equal(cls["n_top"], whole_rep["n_top"])

    # ---- replicated: reads split across ranks, results concatenated in input order
    sl = sharding.unit_slices(len(reads), world)[rank]
    part = full.read_id_batch(reads[sl[0]:sl[1]])
    kinds = sharding.concat_in_order(part["kind"])
    hits = sharding.concat_in_order(part["hits"])
    whole = full.read_id_batch(reads)
    assert np.array_equal(kinds, whole["kind"]) and np.array_equal(hits, whole["hits"])
    dist.barrier()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
