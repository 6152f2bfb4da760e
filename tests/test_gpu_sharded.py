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
    # ---- column-sharded default report: survivors broadcast from rank 0, one byte per k-mer all-reduced
    fq = synth.reads_from(rng, genomes[:3], 600, read_len=120, insert=250, err=0.01, frac_random=0.05, n_rate=0.001)
    dq = [[m for r in fq for m in r], [genomes[7][:1200]], [b"ACG"], [genomes[250][100:900], genomes[3][:500]]]
    # (auto_cutoff on a query without k-mers panics in the reference: the empty query only goes with the fixed filters)
    reports = [(filt, dq[:2] if filt < 0 else dq, sharding.sharded_default_report(gix, dq[:2] if filt < 0 else dq, shards, dev,
                                                                                   seq_mode=L.CID_SEQ_FASTQ, filt=filt)) for filt in (-1, 0, 2)]
    if rank == 0:
        same = True
        for filt, qs, got in reports:
            o = whole.query_counts(qs, O.MODE_FASTQ, False, filt)
            same = same and all(np.array_equal(got[key], o[key]) for key in ("counts", "num_kmers", "cutoff", "uniq_n", "uniq_sum", "uniq_mode"))
        open(os.path.join(out_dir, "ok_report"), "w").write("1" if same and int(o["uniq_n"].sum()) > 50 else "0")
    # ---- column-sharded read_id: row-present bitmaps OR-ed across ranks, per-shard reports with insertion steps,
    # all-gathered and merged on every rank (sharding.merge_read_reports)
    from colorid_b200.api import classify_reads
    _, bm_ptr, bm_words = gix.device_ptrs()
    view = sharding.device_view(bm_ptr, bm_words, dev)
    bm = view.cpu()
    sharding.or_reduce_bitmap(bm)
    view.copy_(bm.to(dev))
    torch.cuda.synchronize()
    gix.set_rownz_global(True)
    ctx.set_option("readid_report_steps", 1)
    reads = synth.reads_from(rng, genomes, 150, read_len=150, insert=320, err=0.004, frac_random=0.2)
    merged = sharding.merge_read_reports(gix.read_id_batch(reads), shards, N)
    n_ref_parts = [None] * world
    dist.all_gather_object(n_ref_parts, gix.n_ref)
    cls = classify_reads((S, H, N), np.concatenate(n_ref_parts), merged)
    if rank == 0:
        o = whole.read_id_batch(reads, n_ref=np.concatenate(n_ref_parts))
        same = bool(np.array_equal(merged["n_set"], o["n_set"]) and np.array_equal(cls["kind"], o["kind"]) and
                    np.array_equal(cls["hits"], o["hits"]) and np.array_equal(cls["n_top"], o["n_top"]))
        for r in range(len(reads)):
            gd = dict(zip(merged["rep_colour"][r, :merged["rep_n"][r]].tolist(), merged["rep_count"][r, :merged["rep_n"][r]].tolist()))
            od = dict(zip(o["rep_colour"][r, :o["rep_n"][r]].tolist(), o["rep_count"][r, :o["rep_n"][r]].tolist()))
            same = same and gd == od
        open(os.path.join(out_dir, "ok_readid"), "w").write("1" if same and int(o["rep_n"].max()) >= 2 else "0")
    dist.destroy_process_group()


def test_column_sharded_counts_exchanged_through_peer_memory(tmp_path):
    ndev = torch.cuda.device_count()
    assert ndev >= 1
    mp.spawn(_worker, args=(2, _free_port(), ndev, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok").read() == "1"
    assert open(tmp_path / "ok_report").read() == "1", "column-sharded default report differs from the oracle on the whole index"
    assert open(tmp_path / "ok_readid").read() == "1", "column-sharded read_id: merged reports / classification differ from the oracle"


@pytest.mark.parametrize("N,k,S,H,world", [(100, 21, 200_003, 2, 2), (300, 21, 200_003, 2, 2), (200, 27, 300_007, 4, 3),
                                           (2500, 31, 20_000_003, 2, 2)])
def test_column_sharded_read_id_merges_to_the_unsharded_report(N, k, S, H, world):
    """Column-sharded read_id in one process: `world` shard indices (64 + 36 accessions: the narrow-row vote kernel; 160 + 140
    and 96 + 64 + 40: the wide one) with their row-present bitmaps OR-ed, per-shard reports carrying insertion steps
    (option readid_report_steps), merged by cid_merge_shard_reports -> the exact report sequence, and classification, of the
    whole index on the GPU, which test_gpu_parity holds to the oracle."""
    import colorid_b200 as cb
    from colorid_b200 import sharding
    from colorid_b200.api import classify_reads, merge_shard_reports
    from oracle import pyoracle as O
    rng = np.random.default_rng(0xC0101D00 + 1300 + N)
    genomes = synth.clade_genomes(rng, N, 3000 if N < 1000 else 1500, n_clades=max(2, N // 10), div=0.01)
    reads = synth.reads_from(rng, genomes, 300, read_len=150, insert=320, err=0.004, frac_random=0.2, n_rate=0.002)
    reads += synth.reads_from(rng, genomes, 40, read_len=100, insert=300, err=0.0, frac_random=0.0, paired=False)
    reads.append([b"ACGT", genomes[0][:150]])                         # too_short
    reads.append([b"N" * 150, b"N" * 150])                            # empty set
    dev = torch.device("cuda:0")
    ctx = cb.Context(0)
    sctx = cb.Context(0)
    sctx.set_option("readid_report_steps", 1)
    full = cb.Index(ctx, S, H, k, N)
    for c in range(N):
        full.build_accession(c, [genomes[c]])
    full.finalize()
    shards = sharding.column_shards(N, world)
    parts, views = [], []
    for lo, hi in shards:
        ix = cb.Index(sctx, S, H, k, hi - lo)
        for c in range(lo, hi):
            ix.build_accession(c - lo, [genomes[c]])
        ix.finalize()
        _, bm_ptr, bm_words = ix.device_ptrs()
        views.append(sharding.device_view(bm_ptr, bm_words, dev))
        parts.append(ix)
    torch.cuda.synchronize()
    merged_bm = views[0].clone()
    for v in views[1:]:
        merged_bm |= v
    for ix, v in zip(parts, views):                                   # what or_reduce_bitmap does across ranks
        v.copy_(merged_bm)
        ix.set_rownz_global(True)
    torch.cuda.synchronize()
    n_ref = np.concatenate([ix.n_ref for ix in parts])
    assert np.array_equal(n_ref, full.n_ref)
    for kw in (dict(), dict(start_sample=0), dict(start_sample=1), dict(d=2)):
        want = full.read_id_batch(reads, **kw)
        got = merge_shard_reports([ix.read_id_batch(reads, **kw) for ix in parts], shards, N)
        assert np.array_equal(got["n_set"], want["n_set"]) and np.array_equal(got["rep_n"], want["rep_n"])
        assert np.array_equal(got["flags"] & 7, want["flags"] & 7)
        for r in range(len(reads)):
            n = want["rep_n"][r]
            assert got["rep_colour"][r, :n].tolist() == want["rep_colour"][r, :n].tolist(), (kw, r)
            assert got["rep_count"][r, :n].tolist() == want["rep_count"][r, :n].tolist(), (kw, r)
        a, b = classify_reads((S, H, N), n_ref, got), classify_reads((S, H, N), n_ref, want)
        for key in ("kind", "hits", "n_top", "top"):
            assert np.array_equal(a[key], b[key])
    assert int(want["rep_n"].max()) >= 3 and int((want["rep_n"] == 0).sum()) >= 1
    if N > 400:          # (the C5 shard shape, 1,280 + 1,220 accessions: GPU against GPU only)
        for ix in parts + [full]:
            ix.close()
        sctx.close()
        ctx.close()
        return
    # the oracle on the whole index agrees with the merged report (as a map; order is checked against the GPU above)
    whole = O.Index(S, H, k, N)
    whole.build_many([[g] for g in genomes], O.MODE_FASTA, threads=2)
    o = whole.read_id_batch(reads)
    got = merge_shard_reports([ix.read_id_batch(reads) for ix in parts], shards, N)
    for r in range(len(reads)):
        gd = dict(zip(got["rep_colour"][r, :got["rep_n"][r]].tolist(), got["rep_count"][r, :got["rep_n"][r]].tolist()))
        od = dict(zip(o["rep_colour"][r, :o["rep_n"][r]].tolist(), o["rep_count"][r, :o["rep_n"][r]].tolist()))
        assert gd == od, r
    for ix in parts + [full]:
        ix.close()
    sctx.close()
    ctx.close()


@pytest.mark.parametrize("N,k,S,H,world", [(100, 21, 200_003, 2, 2), (300, 27, 300_007, 4, 3)])
def test_column_sharded_default_report_one_process(N, k, S, H, world):
    """The three calls of the column-sharded default report with the exchange done by hand (sum of the per-shard popcount
    bytes): counts, num_kmers, cutoffs and the unique-hit summaries of every shard, side by side, equal the oracle's on the
    whole index -- FASTQ read-set query with auto_cutoff and fixed filters, FASTA queries, an empty query."""
    import colorid_b200 as cb
    from colorid_b200 import lib as L
    from colorid_b200 import sharding
    from oracle import pyoracle as O
    rng = np.random.default_rng(0xC0101D00 + 1700 + N)
    genomes = synth.clade_genomes(rng, N, 3000, n_clades=max(2, N // 10), div=0.01)
    whole = O.Index(S, H, k, N)
    whole.build_many([[g] for g in genomes], O.MODE_FASTA, threads=2)
    dev = torch.device("cuda:0")
    ctx = cb.Context(0)
    shards = sharding.column_shards(N, world)
    parts = []
    for lo, hi in shards:
        ix = cb.Index(ctx, S, H, k, hi - lo)
        for c in range(lo, hi):
            ix.build_accession(c - lo, [genomes[c]])
        ix.finalize()
        parts.append(ix)
    fq = synth.reads_from(rng, genomes[:3], 800, read_len=120, insert=250, err=0.01, frac_random=0.05, n_rate=0.001)
    # (auto_cutoff on a query without k-mers panics in the reference: the empty query only goes with the fixed filters)
    cases = [(L.CID_SEQ_FASTQ, O.MODE_FASTQ, [[m for r in fq for m in r], [genomes[N - 1][:900]]], (-1,)),
             (L.CID_SEQ_FASTQ, O.MODE_FASTQ, [[m for r in fq for m in r], [genomes[N - 1][:900]], [b"ACG"]], (0, 2)),
             (L.CID_SEQ_FASTA, O.MODE_FASTA, [[genomes[5]], [genomes[N // 2][100:1500], genomes[1][:700]], [synth.rand_seq(rng, 900)]], (0, 1))]
    checked = 0
    for seq_mode, omode, queries, filters in cases:
        nq = len(queries)
        for filt in filters:
            o = whole.query_counts(queries, omode, False, filt)
            d_slots, surv, cutoff = parts[0].query_survivors(queries, seq_mode, False, filt)
            total = int(surv.sum())
            assert np.array_equal(cutoff, o["cutoff"]) and np.array_equal(surv, o["num_kmers"])
            slots = sharding.device_view(d_slots, max(total, 1) * 2, dev, "<i8").clone() if total else torch.zeros(2, dtype=torch.int64, device=dev)
            per = []
            for ix in parts:
                counts = torch.zeros((nq, ix.N), dtype=torch.int32, device=dev)
                nk = torch.zeros(nq, dtype=torch.int64, device=dev)
                pc = torch.zeros(max(total, 1), dtype=torch.uint8, device=dev)
                col = torch.zeros(max(total, 1), dtype=torch.int32, device=dev)
                ix.slots_counts_dev(slots.data_ptr(), surv, counts.data_ptr(), nk.data_ptr(), pc.data_ptr(), col.data_ptr())
                per.append((counts, nk, pc, col))
            torch.cuda.synchronize()
            pc_sum = sum(p[2].to(torch.int32) for p in per).to(torch.uint8)
            un, us, um = zip(*[ix.slots_uniq_dev(slots.data_ptr(), surv, p[2].data_ptr(), pc_sum.data_ptr(), p[3].data_ptr())
                               for ix, p in zip(parts, per)])
            assert np.array_equal(np.concatenate([p[0].cpu().numpy() for p in per], axis=1).astype(np.uint32), o["counts"])
            for p in per:
                assert np.array_equal(p[1].cpu().numpy().astype(np.uint64), o["num_kmers"])
            assert np.array_equal(np.concatenate(un, axis=1), o["uniq_n"])
            assert np.array_equal(np.concatenate(us, axis=1), o["uniq_sum"])
            assert np.array_equal(np.concatenate(um, axis=1), o["uniq_mode"])
            checked += int(o["uniq_n"].sum())
    assert checked > 100
    for ix in parts:
        ix.close()
    ctx.close()
