"""Multi-GPU glue (one process per GPU, torch.distributed): how units are split across ranks.

Replicated index (configs C1-C4): every rank holds the whole signature matrix; reads / queries are
independent units dealt in contiguous slices, results concatenated in input order.  No data-path
collective is needed.

Column-sharded index (C5, or whenever the matrix exceeds one GPU): rank g owns the accessions
[c_lo, c_hi) (whole 32-accession word columns) of EVERY row.  Every rank processes all k-mers against
its slice; per-query counts / AND-rows are disjoint column slices, so one all_gather reassembles
them (NCCL over NVLink on GPUs, gloo in the CPU tests).  Row presence ("is this Bloom row absent?",
perfect_search.rs:32 / read_id_mt_pe.rs:121) is a property of the WHOLE row, so the per-rank
row-present bitmaps are OR-reduced once after the build.
"""
import numpy as np
import torch
import torch.distributed as dist


def unit_slices(n_units, world):
    """Contiguous, near-equal slices of [0, n_units) for ranks 0..world-1."""
    base, rem = divmod(n_units, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def column_shards(n_colours, world):
    """Accession ranges per rank, aligned to 32-accession words (bit c of a row = word c//32, bit c%32)."""
    words = (n_colours + 31) // 32
    out = []
    for lo, hi in unit_slices(words, world):
        out.append((min(lo * 32, n_colours), min(hi * 32, n_colours)))
    return out


def gather_counts(local_counts, shards, group=None):
    """local_counts [nq, n_local] (this rank's accession slice) -> [nq, N] on every rank."""
    world = dist.get_world_size(group)
    width = max(hi - lo for lo, hi in shards)
    nq = local_counts.shape[0]
    pad = torch.zeros((nq, width), dtype=local_counts.dtype, device=local_counts.device)
    pad[:, : local_counts.shape[1]] = local_counts
    if pad.is_cuda and dist.get_backend(group) == "gloo":          # gloo has no all_gather of device tensors: stage through the host
        host = pad.cpu()
        parts = [torch.empty_like(host) for _ in range(world)]
        dist.all_gather(parts, host, group=group)
        parts = [p.to(pad.device) for p in parts]
    else:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:, : hi - lo] for p, (lo, hi) in zip(parts, shards)], dim=1)


def gather_and_rows(local_words, shards, group=None):
    """Perfect search: local AND-rows [nq, W_local] -> [nq, W]; shards are word-aligned so words concatenate."""
    world = dist.get_world_size(group)
    wshards = [((lo + 31) // 32, (hi + 31) // 32) for lo, hi in shards]
    width = max(hi - lo for lo, hi in wshards)
    nq = local_words.shape[0]
    pad = torch.zeros((nq, width), dtype=local_words.dtype, device=local_words.device)
    pad[:, : local_words.shape[1]] = local_words
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:, : hi - lo] for p, (lo, hi) in zip(parts, wshards)], dim=1)


def or_reduce_bitmap(bitmap, group=None):
    """In-place OR over ranks of a row-present bitmap (int32/int64 tensor).  NCCL has no bitwise reductions,
    so on GPUs the bitmaps (S/8 bytes each) are all-gathered over NVLink and OR-ed locally."""
    if dist.get_backend(group) == "nccl":
        parts = [torch.empty_like(bitmap) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, bitmap, group=group)
        acc = parts[0]
        for p in parts[1:]:
            acc = torch.bitwise_or(acc, p)
        bitmap.copy_(acc)
    else:
        dist.all_reduce(bitmap, op=dist.ReduceOp.BOR, group=group)
    return bitmap


class _DevArray:
    """__cuda_array_interface__ view of a raw device pointer owned by libcolorid_b200 (no copy)."""

    def __init__(self, ptr, n, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def device_view(ptr, n, device, typestr="<i4"):
    """torch tensor aliasing n elements at device pointer `ptr` (e.g. cid_index_device_ptrs' row-present bitmap)."""
    return torch.as_tensor(_DevArray(ptr, n, typestr), device=device)


def any_reduce_flags(flags, group=None):
    """Per-query 'some row is absent' flags: OR over ranks (uint8/int32 tensor), in place."""
    dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
    return flags


def concat_in_order(local_rows, group=None):
    """Replicated mode: per-rank result arrays (numpy, first axis = this rank's units) -> global order."""
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, local_rows, group=group)
    return np.concatenate(parts, axis=0)


def sharded_default_report(gix, queries, shards, device, seq_mode=0, filt=-1, group=None, root=0, packed=None):
    """Column-sharded `search` with the default 7-column report (batch_search_pe.rs:9-179, reports.rs:8-48): counts are
    per-accession and only need gathering, but "this k-mer hits exactly one accession" is about the whole row.  Rank `root`
    counts the queries' k-mers, applies the filter and broadcasts the dense survivor list, so that a list index means the
    same k-mer on every rank; every rank gathers its column slice for all of them and leaves one byte per k-mer
    (min(local popcount, 2)); the bytes are summed across ranks (all_reduce) and each rank summarises the k-mers whose
    local and global popcounts are both 1 for ITS accessions.  Returns the full-width result on every rank, like
    api.Index.query_counts on an unsharded index."""
    rank = dist.get_rank(group)
    nq = len(packed[2]) - 1 if packed is not None else len(queries)
    meta = [None]
    d_slots = 0
    if rank == root:
        d_slots, surv, cutoff = gix.query_survivors(queries, seq_mode, False, filt, packed=packed)
        meta = [(surv, cutoff)]
    dist.broadcast_object_list(meta, src=root, group=group)
    surv, cutoff = meta[0]
    total = int(surv.sum())
    slots = torch.empty(max(total, 1) * 2, dtype=torch.int64, device=device)          # 16 bytes per k-mer
    if rank == root and total:
        slots[: total * 2].copy_(device_view(d_slots, total * 2, device, "<i8"))
    dist.broadcast(slots, src=root, group=group)
    n_local = gix.N
    counts = torch.zeros((nq, n_local), dtype=torch.int32, device=device)
    nk = torch.zeros(max(nq, 1), dtype=torch.int64, device=device)
    pc = torch.zeros(max(total, 1), dtype=torch.uint8, device=device)
    col = torch.zeros(max(total, 1), dtype=torch.int32, device=device)
    stream = torch.cuda.current_stream(device).cuda_stream if device.type == "cuda" else 0
    gix.slots_counts_dev(slots.data_ptr(), surv, counts.data_ptr(), nk.data_ptr(), pc.data_ptr(), col.data_ptr(), stream)
    pc_sum = pc.clone()
    dist.all_reduce(pc_sum, group=group)                      # <= 2 per rank: no overflow of a byte below 128 ranks
    un, us, um = gix.slots_uniq_dev(slots.data_ptr(), surv, pc.data_ptr(), pc_sum.data_ptr(), col.data_ptr(), stream)
    out = dict(counts=gather_counts(counts, shards, group).cpu().numpy().astype(np.uint32), num_kmers=nk[:nq].cpu().numpy().astype(np.uint64),
               cutoff=np.asarray(cutoff))
    for name, a in (("uniq_n", un), ("uniq_sum", us), ("uniq_mode", um)):
        t = torch.from_numpy(a.astype(np.int64)).to(device)
        out[name] = gather_counts(t, shards, group).cpu().numpy().astype(np.uint64)
    return out


def merge_read_reports(local_report, shards, n_total, group=None, rep_cap=None):
    """Column-sharded read_id (read_id_mt_pe.rs:104-165 is per-colour once the first absent row is known, and that is a
    property of whole rows: or_reduce_bitmap + Index.set_rownz_global): every rank classifies ALL reads against its
    accession slice with the context option readid_report_steps = 1; the sparse per-read reports are all-gathered and
    merged into the unsharded report on every rank (api.merge_shard_reports), ready for api.classify_reads."""
    from .api import merge_shard_reports
    world = dist.get_world_size(group)
    # only the rep_n[r] used entries of every read travel (a report row has n_local + 1 slots, a read fills a handful)
    n = np.ascontiguousarray(local_report["rep_n"], dtype=np.uint32)
    rc, rv = np.asarray(local_report["rep_colour"]), np.asarray(local_report["rep_count"])
    used = np.arange(rc.shape[1], dtype=np.uint32)[None, :] < n[:, None]
    mine = dict(n_set=np.ascontiguousarray(local_report["n_set"]), flags=np.ascontiguousarray(local_report["flags"]), rep_n=n,
                colour=rc[used], count=rv[used])
    parts = [None] * world
    dist.all_gather_object(parts, mine, group=group)
    cap = max(1, max(int(p["rep_n"].max()) if len(p["rep_n"]) else 0 for p in parts))
    reports = []
    for p in parts:
        nr = len(p["rep_n"])
        sel = np.arange(cap, dtype=np.uint32)[None, :] < p["rep_n"][:, None]
        c, v = np.zeros((nr, cap), np.uint32), np.zeros((nr, cap), np.uint32)
        c[sel], v[sel] = p["colour"], p["count"]
        reports.append(dict(n_set=p["n_set"], flags=p["flags"], rep_n=p["rep_n"], rep_colour=c, rep_count=v))
    return merge_shard_reports(reports, shards, n_total, rep_cap=rep_cap)


class PeerCounts:
    """Full-width [nq, n_total] count buffers, one per rank, each opened on every other rank through CUDA IPC so that
    cid_query_counts_sharded_dev can add its column slice straight into all of them over NVLink (no collective)."""

    def __init__(self, ctx, nq, n_total, device, group=None):
        import ctypes as C
        from . import lib as L
        self.ctx, self.lib, self.group = ctx, ctx.lib, group
        self.nq, self.n_total = nq, n_total
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise ValueError("at most 8 destination buffers (one NVSwitch domain)")
        own = L.vp()
        L.check(self.lib.cid_dev_alloc(ctx.h, nq * n_total * 4, C.byref(own)))
        self.own = own.value
        handle = (C.c_uint8 * 64)()
        L.check(self.lib.cid_ipc_export(ctx.h, own, handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.ptrs, self._opened = [], []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.own)
                continue
            p = L.vp()
            L.check(self.lib.cid_ipc_open(ctx.h, (C.c_uint8 * 64).from_buffer_copy(h), C.byref(p)))
            self.ptrs.append(p.value)
            self._opened.append(p.value)
        self.dest = (L.vp * self.world)(*self.ptrs)
        self.view = device_view(self.own, nq * n_total, device).view(nq, n_total)     # this rank's complete result

    def begin_pass(self):
        """zero the own buffer; no peer may write before every rank has done so"""
        self.view.zero_()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def end_pass(self):
        """own stores (local and remote) are complete after the stream sync; the barrier says the same of every peer"""
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        return self.view

    def close(self):
        for p in self._opened:
            self.lib.cid_ipc_close(self.ctx.h, p)
        self._opened = []
        dist.barrier(group=self.group)          # nobody frees memory a peer still has mapped
        if self.own:
            self.lib.cid_dev_free(self.ctx.h, self.own)
            self.own = None
