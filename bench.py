#!/usr/bin/env python
"""bench.py — read_id throughput of the BIGSI hot path on B200 (BASELINE.json configs[1], "C2").

Workload C2: classify synthetic 150 bp paired-end reads against a 46-accession k=31 S=50M H=4
index (replicated on every GPU, reads split across GPUs).  One step = one batch of read pairs.

  value   read pairs/s with the batch resident in HBM (cid_read_id_batch_dev; CUDA events)
  e2e     read pairs/s through the host-pointer C ABI (cid_read_id_batch + cid_classify_reads):
          pinned host reads/quals in, per-read classifications out, H2D/D2H inside the timed region
  roofline  the slowest of the three read_id kernels, algorithmic bytes / CUDA-event time
  cpu_baseline  the C++ oracle (restatement of the Rust reference) on all host cores, bounded sample

`--impl reference` times that CPU restatement alone (the Rust reference cannot be built here).
"""
import argparse
import hashlib
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(n_acc=46, genome_len=3_300_000, n_clades=8, div=0.01, k=31, S=50_000_000, H=4,
           read_len=150, insert_lo=300, insert_hi=400, err=0.005, frac_random=0.30, lowq=0.01,
           start_sample=3, qual_offset=15, fp_correct=1e-3)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------- synthetic data (torch, any device)
def make_genomes(torch, dev, cfg, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    L, A, NC = cfg["genome_len"], cfg["n_acc"], cfg["n_clades"]
    roots = torch.randint(0, 4, (NC, L), generator=g, device=dev, dtype=torch.uint8)
    codes = torch.empty((A, L), device=dev, dtype=torch.uint8)
    for a in range(A):
        r = roots[a % NC].clone()
        m = torch.rand(L, generator=g, device=dev) < cfg["div"]
        r[m] = torch.randint(0, 4, (int(m.sum()),), generator=g, device=dev, dtype=torch.uint8)
        codes[a] = r
    return codes          # 0..3


ASCII = (65, 67, 71, 84)


def make_reads(torch, dev, cfg, genomes, n_pairs, seed):
    """-> bases[n,2*rl] uint8 ASCII (mate1|mate2), quals[n,2*rl] uint8."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    rl, L, A = cfg["read_len"], cfg["genome_len"], cfg["n_acc"]
    lut = torch.tensor(ASCII, device=dev, dtype=torch.uint8)
    out_b = torch.empty((n_pairs, 2 * rl), device=dev, dtype=torch.uint8)
    out_q = torch.empty((n_pairs, 2 * rl), device=dev, dtype=torch.uint8)
    ar = torch.arange(rl, device=dev)
    chunk = 250_000
    for s in range(0, n_pairs, chunk):
        n = min(chunk, n_pairs - s)
        gi = torch.randint(0, A, (n,), generator=g, device=dev)
        ins = torch.randint(cfg["insert_lo"], cfg["insert_hi"] + 1, (n,), generator=g, device=dev)
        pos = (torch.rand(n, generator=g, device=dev) * (L - cfg["insert_hi"] - 1)).long()
        flat = genomes.view(-1)
        base = gi * L + pos
        m1 = flat[(base[:, None] + ar[None, :])]
        m2 = 3 - flat[(base + ins - 1)[:, None] - ar[None, :]]
        flip = torch.rand(n, generator=g, device=dev) < 0.5
        a = torch.where(flip[:, None], m2, m1)
        b = torch.where(flip[:, None], m1, m2)
        codes = torch.cat([a, b], dim=1)
        rnd = torch.rand(n, generator=g, device=dev) < cfg["frac_random"]
        nr = int(rnd.sum())
        if nr:
            codes[rnd] = torch.randint(0, 4, (nr, 2 * rl), generator=g, device=dev, dtype=torch.uint8)
        e = torch.rand((n, 2 * rl), generator=g, device=dev) < cfg["err"]
        ne = int(e.sum())
        if ne:
            codes[e] = torch.randint(0, 4, (ne,), generator=g, device=dev, dtype=torch.uint8)
        out_b[s:s + n] = lut[codes.long()]
        q = torch.full((n, 2 * rl), 73, device=dev, dtype=torch.uint8)
        q[torch.rand((n, 2 * rl), generator=g, device=dev) < cfg["lowq"]] = 35
        out_q[s:s + n] = q
    return out_b, out_q


def pack_reads_torch(torch, bases, quals, qual_offset):
    """The planes of cid_pack_reads (include/colorid_b200.h "packed reads") for fixed-length reads, on the device, so that
    the synthetic pool never crosses PCIe: -> int32 [n, ceil(L/16) + ceil(L/32)] (codes | bad), one row per read (2 mates).
    Cross-checked against the product's host packer on a sample in run_ours."""
    n, L = bases.shape
    a = bases.long()
    u = a & 0xDF
    good = ((u == 65) | (u == 67) | (u == 71) | (u == 84)) & (quals.long() >= qual_offset + 33)
    code = (((a >> 1) ^ (a >> 2)) & 3) * good
    ncw, nbw = (L + 15) // 16, (L + 31) // 32
    pad = torch.zeros((n, ncw * 16 - L), device=bases.device, dtype=torch.long)
    cw = torch.cat([code, pad], 1).view(n, ncw, 16)
    sh = (30 - 2 * torch.arange(16, device=bases.device)).view(1, 1, 16)
    codes = (cw << sh).sum(2)
    badb = torch.cat([(~good).long(), torch.ones((n, nbw * 32 - L), device=bases.device, dtype=torch.long)], 1).view(n, nbw, 32)
    bad = (badb << torch.arange(32, device=bases.device).view(1, 1, 32)).sum(2)
    w = torch.cat([codes, bad], 1) & 0xFFFFFFFF
    w = torch.where(w >= 2 ** 31, w - 2 ** 32, w)
    return w.to(torch.int32).contiguous()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML in-process every 5 ms (the timed
    region of a default run is ~0.2 s, too short for `nvidia-smi -lms`), nvidia-smi as the fallback."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.sm, self.mx, self.reasons, self.p, self.th = gpu, [], [], set(), None, None
        self.run = False
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # torch's device index follows CUDA_VISIBLE_DEVICES; NVML's does not
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if gpu < len(ids) and ids[gpu].isdigit():
                    phys = int(ids[gpu])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
        except Exception:
            self.nv = None

    def _poll_nvml(self):
        nv = self.nv
        names = ((nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"), (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                 (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"), (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap"))
        try:
            self.mx.append(int(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
        except Exception:
            pass
        while self.run:
            try:
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in names:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def _read_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.p.stdout:
            r = [x.strip() for x in line.split(",")]
            if r and r[0].isdigit():
                self.sm.append(int(r[0]))
            if len(r) > 1 and r[1].isdigit():
                self.mx.append(int(r[1]))
            for i, n in enumerate(names):
                if len(r) > 2 + i and r[2 + i] == "Active":
                    self.reasons.add(n)

    def start(self):
        self.run = True
        if self.nv is not None:
            self.th = threading.Thread(target=self._poll_nvml, daemon=True)
            self.th.start()
            return
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read_smi, daemon=True).start()
        except Exception:
            self.p = None

    def stop(self):
        self.run = False
        if self.th:
            self.th.join(timeout=1.0)
        if self.p:
            self.p.terminate()
        order = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        return dict(sm_mhz=int(np.median(self.sm)) if self.sm else None, sm_max_mhz=max(self.mx) if self.mx else None,
                    reasons=[n for n in order if n in self.reasons], samples=len(self.sm),
                    source="nvml" if self.nv is not None else "nvidia-smi")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


GATHER_CEILING_G_PER_S = 54.5     # measured: profiles/r1_gather_probe.txt


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this
    same workload (profiles/r1_traffic.json; null when a kernel was not captured)."""
    out = {}
    for name in ("r1_traffic.json", "r2_traffic.json"):       # (round 2 adds the kernels of the partitioned vote)
        try:
            out.update(json.load(open(os.path.join(ROOT, "profiles", name)))["bytes_per_launch"])
        except Exception:
            pass
    return out


def offsets_for(n_pairs, rl):
    seq_offs = (np.arange(2 * n_pairs + 1, dtype=np.uint64) * np.uint64(rl))
    read_offs = (np.arange(n_pairs + 1, dtype=np.uint64) * np.uint64(2))
    return seq_offs, read_offs


# ----------------------------------------------------------------------------- reference arm / cpu baseline
def oracle_index_from_dense(O, cfg, dense, n_ref):
    oix = O.Index(cfg["S"], cfg["H"], cfg["k"], cfg["n_acc"])
    oix.words()[:] = dense
    oix.n_ref[:] = n_ref
    return oix


def oracle_read_id(O, oix, cfg, bases_np, quals_np, threads):
    """bases/quals: [n, 2*rl] uint8 numpy. Applies qual_mask (vectorised, same rule as seq.rs:36-56) then classifies."""
    n, w = bases_np.shape
    rl = w // 2
    masked = np.where(quals_np < cfg["qual_offset"] + 33, np.uint8(78), bases_np)
    seq_offs, read_offs = offsets_for(n, rl)
    flat = np.ascontiguousarray(masked).reshape(-1)
    lib = O.lib()
    u32 = lambda: np.zeros(n, np.uint32)
    n_set, hits, n_top, rep_n = u32(), u32(), u32(), u32()
    kind = np.zeros(n, np.int32)
    top = np.zeros((n, 8), np.uint32)
    cap = cfg["n_acc"] + 1
    rc = np.zeros((n, cap), np.uint32)
    rv = np.zeros((n, cap), np.uint32)
    P = lambda a, t: a.ctypes.data_as(t)
    t0 = time.perf_counter()
    lib.orc_read_id_batch(oix.h, P(flat, C.c_char_p), P(seq_offs, O.u64p), P(read_offs, O.u64p), C.c_uint64(n),
                          C.c_uint32(1), C.c_uint32(cfg["start_sample"]), P(oix.n_ref, O.u64p),
                          C.c_double(cfg["fp_correct"]), C.c_int(16), C.c_int(1), C.c_int(threads), P(n_set, O.u32p),
                          P(kind, O.i32p), P(hits, O.u32p), P(n_top, O.u32p), P(top, O.u32p), C.c_uint32(8),
                          P(rep_n, O.u32p), P(rc, O.u32p), P(rv, O.u32p), C.c_uint32(cap), None, None, None, C.c_uint32(0))
    dt = time.perf_counter() - t0
    return dt, dict(kind=kind, hits=hits, n_top=n_top, top=top, n_set=n_set)


def run_reference(args):
    """CPU restatement of the reference on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import pyoracle as O
    O.lib()
    cfg = dict(CFG)
    if args.quick:
        cfg.update(genome_len=200_000, S=5_000_000)
    cores = os.cpu_count() or 1
    t0 = time.time()
    genomes = make_genomes(torch, "cpu", cfg, 0xC0101D02)
    lut = np.array(ASCII, dtype=np.uint8)
    oix = O.Index(cfg["S"], cfg["H"], cfg["k"], cfg["n_acc"])
    accs = [[lut[genomes[a].numpy()].tobytes()] for a in range(cfg["n_acc"])]
    oix.build_many(accs, O.MODE_FASTA, threads=cores)
    index_sha = hashlib.sha256(oix.words().tobytes()).hexdigest()
    log(f"[reference] oracle index built on {cores} threads in {time.time() - t0:.1f}s, sha256 of the matrix {index_sha[:16]}..")
    sample = args.ref_batch
    b, q = make_reads(torch, "cpu", cfg, genomes, sample * (args.steps + args.warmup), 0xC0101D03)
    b, q = b.numpy(), q.numpy()
    for w in range(args.warmup):
        oracle_read_id(O, oix, cfg, b[w * sample:(w + 1) * sample], q[w * sample:(w + 1) * sample], cores)
    tot = 0.0
    for s in range(args.steps):
        i = (args.warmup + s) * sample
        dt, _ = oracle_read_id(O, oix, cfg, b[i:i + sample], q[i:i + sample], cores)
        tot += dt
    v = sample * args.steps / tot
    line = {"metric": "read_id read pairs/s", "value": v, "unit": "read pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic", "impl": "reference",
            "config": workload_config(cfg, sample),
            "cpu_baseline": {"value": v, "unit": "read pairs/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} read pairs per step; C++ restatement of the Rust reference "
                                       "(oracle/), read_id parallel over reads like rayon par_iter",
                             "index_sha256": index_sha},
            "e2e": {"value": v, "unit": "read pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(cfg, batch):
    return {"workload": ("C2 read_id: 150bp paired-end reads vs 46-accession k=31 S=50M H=4 index (BASELINE.json configs[1])" if cfg["n_acc"] == 46
                         else f"read_id experiment off the headline config: {cfg['n_acc']} accessions"),
            "index": f"{cfg['n_acc']} synthetic genomes x {cfg['genome_len']} bp in {cfg['n_clades']} clades, "
                     f"k={cfg['k']} S={cfg['S']} H={cfg['H']}, replicated per GPU",
            "read_pairs_per_step_per_gpu": batch, "read_len": cfg["read_len"], "start_sample": cfg["start_sample"],
            "qual_offset": cfg["qual_offset"],
            "l2_policy": "inputs larger than L2: each step reads a fresh batch slice (>=60 MB reads+quals per 100k pairs) "
                         "and gathers from a 400 MB matrix (126 MB L2)"}


# ----------------------------------------------------------------------------- C3 gene search (extra, N=1 only)
C3 = dict(n_acc=1000, genome_len=3_000_000, n_clades=20, div=0.01, k=21, S=30_000_000, H=2, n_queries=100_000,
          qlen_lo=1000, qlen_hi=3000)


def run_search_extra(torch, dev, ctx, args, quick):
    """BASELINE.json configs[2]: -g gene search of 100k synthetic 1-3 kb queries against a 1,000-isolate k=21
    S=30M H=2 index.  Returns a dict for the bench line's `search_c3` key (kernel-level numbers only)."""
    import colorid_b200 as cb
    from colorid_b200 import lib as L
    cfg = dict(C3)
    if quick:
        cfg.update(n_acc=64, genome_len=300_000, S=3_000_000, n_queries=5000)
    g = torch.Generator(device=dev)
    g.manual_seed(0xC0101D03)
    A, Lg = cfg["n_acc"], cfg["genome_len"]
    t0 = time.time()
    genomes = make_genomes(torch, dev, cfg, 0xC0101D03)                       # [A, Lg] codes
    lut = torch.tensor(ASCII, device=dev, dtype=torch.uint8)
    gix = cb.Index(ctx, cfg["S"], cfg["H"], cfg["k"], A)
    offs = torch.tensor([0, Lg], device=dev, dtype=torch.int64)
    tb = time.time()
    for a in range(A):
        asc = lut[genomes[a].long()].contiguous()
        torch.cuda.synchronize()
        gix.build_accession_dev(a, asc.data_ptr(), offs.data_ptr(), 1, Lg)
    gix.finalize()
    build_s = time.time() - tb
    # queries: 80% cut from isolates, 10% with 1-2% substitutions, 10% random
    nq = cfg["n_queries"]
    qlen = torch.randint(cfg["qlen_lo"], cfg["qlen_hi"] + 1, (nq,), generator=g, device=dev)
    qoff = torch.zeros(nq + 1, device=dev, dtype=torch.int64)
    qoff[1:] = torch.cumsum(qlen, 0)
    total = int(qoff[-1].item())
    qi = torch.repeat_interleave(torch.arange(nq, device=dev), qlen)
    within = torch.arange(total, device=dev) - qoff[qi]
    iso = torch.randint(0, A, (nq,), generator=g, device=dev)
    start = (torch.rand(nq, generator=g, device=dev) * (Lg - cfg["qlen_hi"] - 1)).long()
    codes = genomes.view(-1)[(iso * Lg + start)[qi] + within]
    kind = torch.rand(nq, generator=g, device=dev)
    mut_rate = torch.where((kind >= 0.8) & (kind < 0.9), torch.rand(nq, generator=g, device=dev) * 0.01 + 0.01,
                           torch.zeros(nq, device=dev))
    mut = torch.rand(total, generator=g, device=dev) < mut_rate[qi]
    rnd = (kind >= 0.9)[qi]
    repl = torch.randint(0, 4, (total,), generator=g, device=dev, dtype=torch.uint8)
    codes = torch.where(mut | rnd, repl, codes)
    d_bases = lut[codes.long()].contiguous()
    h_seq_offs = qoff.cpu().numpy().astype(np.uint64)
    h_query_offs = np.arange(nq + 1, dtype=np.uint64)
    d_query_offs = torch.from_numpy(h_query_offs.view(np.int64)).to(dev)
    d_counts = torch.zeros((nq, A), device=dev, dtype=torch.int32)
    d_nk = torch.zeros(nq, device=dev, dtype=torch.int64)
    stream = torch.cuda.current_stream()
    lib = ctx.lib
    P = lambda a, tp=L.u64p: a.ctypes.data_as(tp)

    def run():
        L.check(lib.cid_query_counts_dev(gix.h, d_bases.data_ptr(), qoff.data_ptr(), nq, total, d_query_offs.data_ptr(),
                                         P(h_query_offs), P(h_seq_offs), nq, 0, d_counts.data_ptr(), d_nk.data_ptr(),
                                         stream.cuda_stream))
    run()
    torch.cuda.synchronize()
    ctx.profile(True)
    reps = 3
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(reps):
        run()
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    prof = ctx.profile_read()
    ctx.profile(False)
    lookups = int(d_nk.sum().item())
    R = 4 * ((A + 31) // 32)
    peak, peak_src = measured_peak()
    out = {"workload": "C3 gene search (-g): %d queries of %d-%d bp vs %d-isolate k=%d S=%d H=%d index (BASELINE.json configs[2])"
                       % (nq, cfg["qlen_lo"], cfg["qlen_hi"], A, cfg["k"], cfg["S"], cfg["H"]),
           "lookups": lookups, "ms_per_pass": ms, "lookups_per_s": lookups / (ms / 1e3), "unit": "k-mer lookups/s",
           "build_seconds": build_s, "build_gbp_per_s": A * Lg / build_s / 1e9, "setup_seconds": time.time() - t0,
           "kernels": {k_: {"ms_per_launch": v[0] / v[1], "launches_per_pass": v[1] / reps} for k_, v in prof.items()}}
    if "query_counts" in prof:
        qc_ms = prof["query_counts"][0] / reps            # all query_counts launches of one pass
        alg = lookups * (cfg["H"] * R + 1) + nq * 4 * A    # H*R + k_in per lookup, + 4N counts per query
        out["roofline"] = {"bound": "hbm", "kernel": "query_counts (query_gather_kernel)", "achieved": alg / (qc_ms / 1e3) / 1e9, "peak": peak,
                           "unit": "GB/s", "frac": alg / (qc_ms / 1e3) / 1e9 / peak,
                           "traffic": ncu_traffic().get("query_counts"),      # the pass is one query_gather launch
                           "peak_source": peak_src, "algorithmic_bytes_per_pass": alg, "ms_per_pass": qc_ms}
    # ---- -s -m perfect search of the same queries (perfect_search.rs:62-120) through the host-pointer ABI:
    # query bases H2D, query_front + streaming AND gather, AND rows + status D2H inside the timed region
    try:
        W = (A + 31) // 32
        h_bases = d_bases.cpu().pin_memory().numpy()          # page-locked, as cid_host_alloc would give a caller
        and_rows = np.zeros((nq, W), np.uint32)
        status = np.zeros(nq, np.uint8)
        nk = np.zeros(nq, np.uint64)

        def run_perfect():
            L.check(lib.cid_query_perfect_mf(gix.h, h_bases.ctypes.data_as(L.vp), P(h_seq_offs), nq,
                                             and_rows.ctypes.data_as(L.u32p), status.ctypes.data_as(L.u8p), P(nk)))
        run_perfect()
        ctx.profile(True)
        tp0 = time.time()
        for _ in range(reps):
            run_perfect()
        pms = (time.time() - tp0) / reps * 1e3
        pprof = ctx.profile_read()
        ctx.profile(False)
        # size-independent property: a query cut unmodified from isolate i has every row present and bit i set in its AND
        exact = (kind < 0.8).cpu().numpy()
        iso_h = iso.cpu().numpy()
        bit = (and_rows[np.arange(nq), iso_h // 32] >> (iso_h % 32).astype(np.uint32)) & 1
        viol = int(((status[exact] != 0) | (bit[exact] != 1)).sum())
        out["perfect_search"] = {"workload": "-s -m perfect search of the same %d queries (cid_query_perfect_mf, pinned host buffers)" % nq,
                                 "lookups": int(nk.sum()), "ms_per_pass_wall": pms, "lookups_per_s": int(nk.sum()) / (pms / 1e3),
                                 "h2d_bytes": int(h_bases.nbytes + h_seq_offs.nbytes), "d2h_bytes": int(and_rows.nbytes + status.nbytes + nk.nbytes),
                                 "kernels": {k_: {"ms_per_launch": v[0] / v[1], "launches_per_pass": v[1] / reps} for k_, v in pprof.items()},
                                 "self_query_violations": viol, "queries_with_every_row_present": int((status == 0).sum())}
    except Exception as ex:
        out["perfect_search"] = {"error": repr(ex)}
    gix.close()
    return out



# ----------------------------------------------------------------------------- C5: column-sharded build + search
C5 = dict(n_acc=10_000, genome_len=5_000_000, n_clades=100, div=0.01, k=31, S=50_000_000, H=4, n_queries=10_000,
          qlen_lo=1000, qlen_hi=3000, n_probe=16)


def lookups_of(d_nk):
    return int(d_nk.sum().item())


def run_c5(args):
    import torch.distributed as dist
    line = c5_measure(args)
    if line is not None:
        print(json.dumps(line), flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()


def c5_measure(args, n_acc=0, steps=0, warmup=0):
    """BASELINE.json configs[4]: 10,000 synthetic 5 Mbp genomes, k=31 S=50M H=4, the signature matrix column-sharded
    over the ranks (whole 32-accession word columns per rank), every rank gathering all query k-mers from its slice,
    per-query counts re-assembled with one NCCL all_gather.  `--c5-acc` scales the accession count (1250 per rank
    keeps the C5 per-GPU shard shape: 160-byte rows)."""
    import torch
    import torch.distributed as dist
    import colorid_b200 as cb
    from colorid_b200 import lib as L
    from colorid_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    cfg = dict(C5)
    if n_acc or args.c5_acc:
        cfg["n_acc"] = n_acc or args.c5_acc
    if args.quick:
        cfg.update(genome_len=100_000, S=2_000_003, n_queries=500, n_clades=10)
    A, Lg, NC = cfg["n_acc"], cfg["genome_len"], cfg["n_clades"]
    shards = sharding.column_shards(A, world)
    lo, hi = shards[rank]
    n_local = hi - lo
    ctx = cb.Context(local)
    lut = torch.tensor(ASCII, device=dev, dtype=torch.uint8)
    g = torch.Generator(device=dev)
    g.manual_seed(0xC0101D05)
    roots = torch.randint(0, 4, (NC, Lg), generator=g, device=dev, dtype=torch.uint8)     # identical on every rank

    def acc_codes(a):
        """accession a = its clade root with `div` substitutions, reproducible on any rank"""
        ga = torch.Generator(device=dev)
        ga.manual_seed(0xC5000000 + a)
        r = roots[a % NC].clone()
        m = torch.rand(Lg, generator=ga, device=dev) < cfg["div"]
        r[m] = torch.randint(0, 4, (int(m.sum()),), generator=ga, device=dev, dtype=torch.uint8)
        return r

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- build: every rank indexes its own accessions (no exchange), then the row-present bitmaps are OR-ed
    gix = cb.Index(ctx, cfg["S"], cfg["H"], cfg["k"], n_local)
    offs = torch.tensor([0, Lg], device=dev, dtype=torch.int64)
    barrier()
    ctx.profile(True)
    t0 = time.perf_counter()
    gen_s = 0.0
    for a in range(lo, hi):
        tg = time.perf_counter()
        asc = lut[acc_codes(a).long()].contiguous()
        torch.cuda.synchronize()
        gen_s += time.perf_counter() - tg
        gix.build_accession_dev(a - lo, asc.data_ptr(), offs.data_ptr(), 1, Lg)
    gix.finalize()
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0 - gen_s            # synthetic genome generation is not part of the build
    build_prof = ctx.profile_read()
    ctx.profile(False)
    _, bm_ptr, bm_words = gix.device_ptrs()
    if world > 1:
        sharding.or_reduce_bitmap(sharding.device_view(bm_ptr, bm_words, dev))
        gix.set_rownz_global(True)
    tb = torch.tensor([build_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
    build_s_max = float(tb.item())
    log(f"[rank {rank}] shard {lo}..{hi}: built {n_local} accessions in {build_s:.2f}s (+{gen_s:.2f}s generating genomes)")

    # ---- queries (identical on every rank): 70% clade-root slices with 1% substitutions, 20% random, and exact
    # slices of a few probe accessions spread over the shards (self-query property: count == num_kmers)
    gq = torch.Generator(device=dev)
    gq.manual_seed(0xC0101D06)
    nq = cfg["n_queries"]
    qlen = torch.randint(cfg["qlen_lo"], cfg["qlen_hi"] + 1, (nq,), generator=gq, device=dev)
    qoff = torch.zeros(nq + 1, device=dev, dtype=torch.int64)
    qoff[1:] = torch.cumsum(qlen, 0)
    total = int(qoff[-1].item())
    qi = torch.repeat_interleave(torch.arange(nq, device=dev), qlen)
    within = torch.arange(total, device=dev) - qoff[qi]
    clade = torch.randint(0, NC, (nq,), generator=gq, device=dev)
    start = (torch.rand(nq, generator=gq, device=dev) * (Lg - cfg["qlen_hi"] - 1)).long()
    codes = roots.view(-1)[(clade * Lg + start)[qi] + within]
    kind = torch.rand(nq, generator=gq, device=dev)
    mut = (torch.rand(total, generator=gq, device=dev) < 0.01) & (kind < 0.7)[qi]
    rnd = (kind >= 0.8)[qi]
    repl = torch.randint(0, 4, (total,), generator=gq, device=dev, dtype=torch.uint8)
    codes = torch.where(mut | rnd, repl, codes)
    probes = [int(x) for x in np.linspace(0, A - 1, cfg["n_probe"]).astype(int)]
    probe_q = []
    h_qoff = qoff.cpu().numpy()
    h_start = start.cpu().numpy()
    exact_q = np.flatnonzero(((kind >= 0.7) & (kind < 0.8)).cpu().numpy())
    for j, q in enumerate(exact_q[: 4 * len(probes)]):
        a = probes[j % len(probes)]
        codes[h_qoff[q]:h_qoff[q + 1]] = acc_codes(a)[h_start[q]:h_start[q] + (h_qoff[q + 1] - h_qoff[q])]
        probe_q.append((int(q), a))
    d_bases = lut[codes.long()].contiguous()
    h_seq_offs = h_qoff.astype(np.uint64)
    h_query_offs = np.arange(nq + 1, dtype=np.uint64)
    d_query_offs = torch.from_numpy(h_query_offs.view(np.int64)).to(dev)
    d_counts = torch.zeros((nq, n_local), device=dev, dtype=torch.int32)
    d_nk = torch.zeros(nq, device=dev, dtype=torch.int64)
    stream = torch.cuda.current_stream()
    lib = ctx.lib
    P = lambda a, tp=L.u64p: a.ctypes.data_as(tp)
    gathered = [None]

    def search_pass():
        L.check(lib.cid_query_counts_dev(gix.h, d_bases.data_ptr(), qoff.data_ptr(), nq, total, d_query_offs.data_ptr(),
                                         P(h_query_offs), P(h_seq_offs), nq, 0, d_counts.data_ptr(), d_nk.data_ptr(),
                                         stream.cuda_stream))
        gathered[0] = sharding.gather_counts(d_counts, shards) if world > 1 else d_counts     # NCCL all_gather

    for _ in range(max(1, warmup or args.warmup)):
        search_pass()
    barrier()
    ctx.profile(True)
    clk = ClockSampler(local)
    clk.start()
    K = steps or args.steps
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(K):
        search_pass()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1) / K
    clocks = clk.stop()
    prof = ctx.profile_read()
    ctx.profile(False)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    lookups = int(d_nk.sum().item())
    # ---- the same pass with the count exchange fused into the gather kernel: every rank adds its column slice into
    # the full-width result of ALL ranks through NVLink peer memory (CUDA IPC), no collective after the kernel
    fused = None
    if world > 1:
        pc = sharding.PeerCounts(ctx, nq, A, dev)

        def fused_pass():
            pc.begin_pass()
            L.check(lib.cid_query_counts_sharded_dev(gix.h, d_bases.data_ptr(), qoff.data_ptr(), nq, total, d_query_offs.data_ptr(),
                                                     P(h_query_offs), P(h_seq_offs), nq, 0, pc.dest, world, A, lo, d_nk.data_ptr(),
                                                     stream.cuda_stream))
            return pc.end_pass()

        def nccl_pass():
            torch.cuda.synchronize()
            dist.barrier()
            search_pass()
            torch.cuda.synchronize()
            dist.barrier()

        fused_pass()
        same = bool(torch.equal(pc.view, gathered[0]))
        times = {}
        for name, fn in (("nccl_all_gather", nccl_pass), ("fused_peer_stores", fused_pass)):
            fn()
            t0 = time.perf_counter()
            for _ in range(K):
                fn()
            tt = torch.tensor([(time.perf_counter() - t0) / K * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            times[name] = float(tt.item())
        ok = torch.tensor([1 if same and bool(torch.equal(pc.view, gathered[0])) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        fused = {"ms_per_pass_wall_incl_barriers": times, "lookups_per_s": {k_: lookups_of(d_nk) / (v / 1e3) for k_, v in times.items()},
                 "identical_to_all_gather_on_every_rank": bool(ok.item()),
                 "note": "both timed the same way: barrier, pass, stream sync, barrier (wall clock, max over ranks); the fused pass also "
                         "zeroes its destination and has one more barrier"}
        pc.close()
    # ---- column-sharded default report and read_id (--c5-extra): timings + size-independent checks
    extra = None
    if world > 1 and args.c5_extra:
        from colorid_b200.api import classify_reads
        extra = {}
        a_src = probes[len(probes) // 2]
        n_ref_parts = [None] * world
        dist.all_gather_object(n_ref_parts, gix.n_ref)
        n_ref_all = np.concatenate(n_ref_parts)
        rl = CFG["read_len"]
        qcfg = dict(CFG, genome_len=Lg, n_acc=1, frac_random=0.0, lowq=0.02)
        n_pairs = Lg * 30 // (2 * rl)
        b, q = make_reads(torch, dev, qcfg, acc_codes(a_src)[None, :], n_pairs, 0xC0101D51)      # identical on every rank
        b[q < CFG["qual_offset"] + 33] = 78
        h_reads = b.cpu().pin_memory()                    # page-locked, as cid_host_alloc would give a caller
        packed = (h_reads.numpy().reshape(-1), np.arange(2 * n_pairs + 1, dtype=np.uint64) * np.uint64(rl),
                  np.array([0, 2 * n_pairs], dtype=np.uint64))
        del b, q
        rep = None
        tms = []
        for i in range(3):
            barrier()
            t0 = time.perf_counter()
            rep = sharding.sharded_default_report(gix, None, shards, dev, seq_mode=L.CID_SEQ_FASTQ, filt=-1, packed=packed)
            barrier()
            tms.append((time.perf_counter() - t0) * 1e3)
        tt = torch.tensor([min(tms[1:])], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        cov = float(rep["counts"][0, a_src]) / float(n_ref_all[a_src])
        assert cov > 0.97 and int(rep["uniq_n"][0].sum()) > 0 and int(rep["cutoff"][0]) >= 1, (cov, rep["cutoff"])
        extra["default_report"] = {"workload": f"30x paired-end read set of accession {a_src} ({2 * n_pairs * rl / 1e6:.0f} Mbp), auto_cutoff, "
                                               f"7-column report against the {A}-accession column-sharded index",
                                   "ms_per_query_wall": float(tt.item()), "lookups": int(rep["num_kmers"][0]),
                                   "hits_over_n_ref_kmers_of_source": cov, "unique_hits_total": int(rep["uniq_n"][0].sum()),
                                   "cutoff": int(rep["cutoff"][0]),
                                   "note": "rank 0 counts and filters, broadcasts 16 B per surviving k-mer; every rank gathers its slice, one "
                                           "byte per k-mer all-reduced; wall clock incl. host packing of results (best of 2 after warm-up)"}
        # read_id: 100k pairs drawn from four probe accessions spread over the shards
        src = [probes[1], probes[len(probes) // 3], probes[2 * len(probes) // 3], probes[-2]]
        rcfg = dict(CFG, genome_len=Lg, n_acc=len(src), frac_random=0.2)
        nrp = 100_000
        gen = torch.stack([acc_codes(a) for a in src])
        rb, rq = make_reads(torch, dev, rcfg, gen, nrp, 0xC0101D52)
        hb, hq = rb.cpu().numpy(), rq.cpu().numpy()
        del gen, rb, rq
        seq_offs_np, read_offs_np = offsets_for(nrp, rl)
        cap = min(n_local + 1, 256)       # report slots per read: ~30 false-positive colours per seeding k-mer at this fill
        params = L.ReadIdParams(1, CFG["start_sample"], CFG["qual_offset"], 16, 1, cap)
        ctx.set_option("readid_report_steps", 1)
        o_n_set, o_flags, o_rep_n = (np.zeros(nrp, np.uint32) for _ in range(3))
        o_rc, o_rv = np.zeros((nrp, cap), np.uint32), np.zeros((nrp, cap), np.uint32)
        PV = lambda x, tp=L.vp: x.ctypes.data_as(tp)
        tms, tmerge = [], []
        for i in range(3):
            barrier()
            t0 = time.perf_counter()
            L.check(lib.cid_read_id_batch(gix.h, PV(hb), PV(hq), PV(seq_offs_np, L.u64p), 2 * nrp, PV(read_offs_np, L.u64p), nrp,
                                          C.byref(params), PV(o_n_set, L.u32p), PV(o_flags, L.u32p), PV(o_rep_n, L.u32p),
                                          PV(o_rc, L.u32p), PV(o_rv, L.u32p)))
            barrier()
            t1 = time.perf_counter()
            merged = sharding.merge_read_reports(dict(n_set=o_n_set, flags=o_flags, rep_n=o_rep_n, rep_colour=o_rc, rep_count=o_rv),
                                                 shards, A, rep_cap=min(A + 1, world * cap))
            cls = classify_reads((cfg["S"], cfg["H"], A), n_ref_all, merged, fp_correct=CFG["fp_correct"])
            barrier()
            tms.append((t1 - t0) * 1e3)
            tmerge.append((time.perf_counter() - t1) * 1e3)
        ctx.set_option("readid_report_steps", 0)
        tt = torch.tensor([min(tms[1:]), min(tmerge[1:])], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        # Accepted reads with many hits come from a source accession's CLADE.  Not always from the accession itself: when
        # the first -B k-mers of the set order all carry a sequencing error, the source is not a candidate colour
        # (read_id_mt_pe.rs:131-136) while a sibling that entered on a Bloom false positive collects the shared k-mers -- the
        # reference's heuristic at 33 % bit fill, reproduced, a few per cent of the reads.  (Random reads are also "accepted"
        # now and then on a handful of false positives: hence the hit threshold.)
        assert int(((merged["flags"] >> 2) & 1).sum()) == 0 and int(((o_flags >> 2) & 1).sum()) == 0, "read reports truncated: raise cap"
        acc = (cls["kind"] == 3) & (cls["hits"] >= 60)
        top_ok = np.isin(cls["top"][acc, 0], np.array(src))
        clade_ok = np.isin(cls["top"][acc, 0] % NC, np.array([a % NC for a in src]))
        assert acc.sum() > 0.2 * nrp and clade_ok.mean() > 0.999 and top_ok.mean() > 0.9, (int(acc.sum()), float(top_ok.mean()), float(clade_ok.mean()))
        extra["read_id"] = {"workload": f"{nrp} read pairs (80 % from accessions {src}, 20 % random) against the column-sharded index, "
                                        "every rank classifies all reads against its slice",
                            "ms_read_id_batch_wall": float(tt[0].item()), "read_pairs_per_s_kernels_and_copies": nrp / (float(tt[0].item()) / 1e3),
                            "ms_gather_merge_classify_wall": float(tt[1].item()), "report_slots_per_read": cap, "accepted_with_60_or_more_hits": int(acc.sum()),
                            "of_those_with_a_source_accession_on_top": float(top_ok.mean()), "with_its_clade_on_top": float(clade_ok.mean())}
    # ---- parity at full size: a query cut verbatim from accession a has every one of its k-mers in a's column
    full = gathered[0]
    nk_h = d_nk.cpu().numpy()
    bad = [(q, a) for q, a in probe_q if int(full[q, a].item()) != int(nk_h[q])]
    assert not bad, f"self-query property violated for {bad[:4]}"
    assert full.shape == (nq, A)
    # ---- ... and against an UNSHARDED index: the count column of an accession depends on that accession's column alone, so a
    # one-GPU index over a sample of accessions from every shard must reproduce those columns of the sharded result exactly
    unsharded = None
    if rank == 0:
        sample = sorted({int(x) for lo_, hi_ in shards for x in (lo_, (lo_ + hi_) // 2, hi_ - 1) if hi_ > lo_})
        uix = cb.Index(ctx, cfg["S"], cfg["H"], cfg["k"], len(sample))
        for j, a in enumerate(sample):
            asc = lut[acc_codes(a).long()].contiguous()
            torch.cuda.synchronize()
            uix.build_accession_dev(j, asc.data_ptr(), offs.data_ptr(), 1, Lg)
        uix.finalize()
        u_counts = torch.zeros((nq, len(sample)), device=dev, dtype=torch.int32)
        u_nk = torch.zeros(nq, device=dev, dtype=torch.int64)
        L.check(lib.cid_query_counts_dev(uix.h, d_bases.data_ptr(), qoff.data_ptr(), nq, total, d_query_offs.data_ptr(), P(h_query_offs),
                                         P(h_seq_offs), nq, 0, u_counts.data_ptr(), u_nk.data_ptr(), stream.cuda_stream))
        torch.cuda.synchronize()
        same_cols = bool(torch.equal(u_counts, full[:, torch.tensor(sample, device=dev)].to(torch.int32)))
        same_nk = bool(torch.equal(u_nk, d_nk))
        unsharded = {"accessions_sampled": len(sample), "queries": nq, "count_columns_identical": same_cols, "num_kmers_identical": same_nk,
                     "how": "one-GPU index (rows of 1-2 words: the fused query_counts kernel) built from first / middle / last accession of "
                            "every shard; its counts must equal those columns of the sharded result"}
        assert same_cols and same_nk, "sharded counts differ from the unsharded index"
        uix.close()
    if rank != 0:
        return None
    peak, peak_src = measured_peak()
    Wp = cb.lib.load().cid_index_row_stride(gix.h)
    R = 4 * Wp
    kern = {k_: {"ms_per_launch": v[0] / v[1], "launches_per_pass": v[1] / K} for k_, v in prof.items()}
    qc_ms = prof["query_counts"][0] / K if "query_counts" in prof else None
    alg = lookups * (cfg["H"] * R + 1) + nq * 4 * n_local
    roofline = None
    if qc_ms:
        roofline = {"bound": "hbm", "kernel": "query_counts", "achieved": alg / (qc_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": alg / (qc_ms / 1e3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_pass": alg, "row_bytes_per_shard": R, "ms_per_pass": qc_ms,
                    "note": "per rank: every rank gathers all k-mers from its own column slice"}
    coll = "none (single GPU)"
    if world > 1:
        coll = "NCCL all_gather of per-query counts after the kernel (+ one all_gather/OR of the row-present bitmaps after the build)"
        if fused and fused["identical_to_all_gather_on_every_rank"]:
            # the product path: the gather kernel stores its column slice into every rank's full-width result over NVLink
            ms_max = fused["ms_per_pass_wall_incl_barriers"]["fused_peer_stores"]
            coll = ("count exchange fused into query_gather: 16-byte stores of each rank's column slice into every rank's full-width "
                    "result through CUDA-IPC peer memory over NVSwitch (cid_query_counts_sharded_dev); ms_per_step is wall clock "
                    "with the two barriers of a pass; the NCCL all_gather variant is timed alongside (fused_count_exchange)")
    line = {"metric": "search k-mer lookups/s", "value": lookups / (ms_max / 1e3), "unit": "k-mer lookups/s", "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": ms_max, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "C5 column-sharded build + gene search (BASELINE.json configs[4])", "n_accessions": A,
                       "genome_len": Lg, "k": cfg["k"], "S": cfg["S"], "H": cfg["H"], "accessions_per_rank": n_local,
                       "row_bytes_per_rank": R, "queries": nq, "collective": coll},
            "clocks": clocks, "gpu_launches": int(ctx.launches), "roofline": roofline, "kernels": kern,
            "build": {"gbp_per_s": A * Lg / build_s_max / 1e9, "seconds": build_s_max,
                      "note": "all ranks build their accession columns concurrently; max over ranks; synthetic genome "
                              "generation excluded",
                      "kernels_ms_total": {k_: v[0] for k_, v in build_prof.items()}},
            "parity": {"self_query_probes": len(probe_q), "violations": 0, "matches_unsharded": unsharded}, "fused_count_exchange": fused,
            "sharded_default_report_and_read_id": extra}
    return line



# ----------------------------------------------------------------------------- C4: build from read sets
C4 = dict(n_acc=1000, genome_len=3_000_000, coverage=30, read_len=150, err=0.005, lowq=0.01, k=21, S=30_000_000, H=2,
          qual_offset=15, n_clades=20, div=0.01)


def run_c4(args):
    """BASELINE.json configs[3]: build of a 1,000-accession index from synthetic 150 bp read sets (3 Mbp genomes at 30x,
    0.5 % substitution errors) with the k-mer frequency filter (auto_cutoff), k=21 S=30M H=2.  One step = one accession
    (600,000 reads = 90 Mbp): k-mer counting, histogram -> auto_cutoff, clean_map, Bloom insert.  `--c4-acc` bounds the
    sample (the 1,000 accessions are statistically identical); accessions are dealt to the ranks (weak scaling)."""
    import torch
    import torch.distributed as dist
    import colorid_b200 as cb
    from colorid_b200 import lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    cfg = dict(C4)
    if args.quick:
        cfg.update(genome_len=200_000, S=3_000_000)
    n_acc = args.c4_acc or 24                      # per rank
    Lg, rl = cfg["genome_len"], cfg["read_len"]
    n_reads = Lg * cfg["coverage"] // rl
    ctx = cb.Context(local)
    for kv in args.opt:
        name, _, val = kv.partition("=")
        ctx.set_option(name, int(val))
    gix = cb.Index(ctx, cfg["S"], cfg["H"], cfg["k"], n_acc)
    lut = torch.tensor(ASCII, device=dev, dtype=torch.uint8)
    g = torch.Generator(device=dev)
    g.manual_seed(0xC0101D04 + rank)
    ar = torch.arange(rl, device=dev)
    seq_offs = torch.arange(n_reads + 1, device=dev, dtype=torch.int64) * rl

    def read_set(a):
        """single-end 30x read set of a fresh random genome; qualities: 1 % of bases below -Q (masked to N on the host
        side of the reference: seq.rs:36-56 -- here applied while generating, the ABI takes masked sequences)"""
        genome = torch.randint(0, 4, (Lg,), generator=g, device=dev, dtype=torch.uint8)
        pos = (torch.rand(n_reads, generator=g, device=dev) * (Lg - rl)).long()
        codes = genome[pos[:, None] + ar[None, :]]
        rc = torch.rand(n_reads, generator=g, device=dev) < 0.5
        codes = torch.where(rc[:, None], 3 - codes.flip(1), codes)
        e = torch.rand((n_reads, rl), generator=g, device=dev) < cfg["err"]
        codes = torch.where(e, torch.randint(0, 4, (n_reads, rl), generator=g, device=dev, dtype=torch.uint8), codes)
        asc = lut[codes.long()]
        asc[torch.rand((n_reads, rl), generator=g, device=dev) < cfg["lowq"]] = 78          # 'N'
        return genome, asc.contiguous()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(1, args.warmup)
    K = min(args.steps, n_acc - W) if n_acc > W else 1
    stats = []
    genomes = {}
    build_s = 0.0
    ctx.profile(False)
    clk = None
    for a in range(W + K):
        genome, asc = read_set(a)
        torch.cuda.synchronize()
        if a == W:
            barrier()
            ctx.profile(True)
            clk = ClockSampler(local)
            clk.start()
        t0 = time.perf_counter()
        n_ref, used = gix.build_accession_dev(a, asc.data_ptr(), seq_offs.data_ptr(), n_reads, n_reads * rl,
                                              mode=L.CID_SEQ_FASTQ, cutoff=-1)
        dt = time.perf_counter() - t0
        if a >= W:
            build_s += dt
        stats.append((n_ref, used))
        if a < 2:
            genomes[a] = genome
    torch.cuda.synchronize()
    clocks = clk.stop()
    prof = ctx.profile_read()
    ctx.profile(False)
    t0 = time.perf_counter()
    gix.finalize()
    torch.cuda.synchronize()
    fin_s = time.perf_counter() - t0
    # size-independent parity properties: (1) auto_cutoff removes the error k-mers: n_ref_kmers is within 0.5 % of the
    # genome's distinct canonical k-mers that the reads cover; (2) a slice of the genome queried with -g hits its accession
    # on (nearly) every k-mer
    q = genomes[0][1000:3000]
    qa = lut[q.long()].contiguous()
    res = gix.query_counts([[bytes(qa.cpu().numpy())]], gene_search=True, filt=0, want_uniq=False)
    frac = float(res["counts"][0, 0]) / float(res["num_kmers"][0])
    assert frac > 0.98, f"self-query of a read-built accession found only {frac:.3f} of its k-mers"
    n_ref0, used0 = stats[0]
    assert 0.95 * Lg < n_ref0 < 1.02 * Lg, f"n_ref_kmers {n_ref0} vs genome length {Lg}: frequency filter off"
    tot = torch.tensor([build_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    bases_per_acc = n_reads * rl
    peak, peak_src = measured_peak()
    kern = {k_: {"ms_per_launch": v[0] / v[1], "launches_per_step": v[1] / K} for k_, v in prof.items()}
    step_ms = float(tot.item()) / K * 1e3
    alg = bases_per_acc + 3 * cfg["S"] // 8          # L + 3S/8 per accession
    line = {"metric": "build Gbp/s", "value": world * K * bases_per_acc / float(tot.item()) / 1e9, "unit": "Gbp/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": "C4 build from read sets with the k-mer frequency filter (BASELINE.json configs[3])",
                       "accessions_per_rank_in_sample": n_acc, "reads_per_accession": n_reads, "read_len": rl,
                       "coverage": cfg["coverage"], "k": cfg["k"], "S": cfg["S"], "H": cfg["H"], "cutoff": "auto_cutoff",
                       "l2_policy": "every step reads a fresh 90 MB read set and a 4 GB count table (126 MB L2)"},
            "clocks": clocks, "gpu_launches": int(ctx.launches),
            "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": alg / (step_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (step_ms / 1e3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_step": alg,
                         "note": "exact k-mer counting has no algorithmic traffic beyond reading L (SURVEY 8d): the count "
                                 "table (16 B x 2 x positions) is implementation traffic, so this fraction is low by construction"},
            "kernels": kern,
            "table_inserts_per_s": (n_reads * (rl - cfg["k"] + 1) / (kern["kmerize_insert"]["ms_per_launch"] *
                                    kern["kmerize_insert"]["launches_per_step"] / 1e3)) if "kmerize_insert" in kern else None,
            "auto_cutoff_used": [int(u) for _, u in stats[:8]], "n_ref_kmers": [int(n) for n, _ in stats[:8]],
            "finalize_seconds": fin_s, "parity": {"self_query_fraction": frac, "n_ref_over_genome_len": n_ref0 / Lg}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- C1: build from FASTA + FASTQ search
def run_c1(args):
    """BASELINE.json configs[0]: build of the 46-genome k=31 S=50M H=4 index from FASTA, then `search` of one paired-end
    30x FASTQ read set (default 7-column report: auto_cutoff on the query's k-mer counts, per-accession hits and the
    unique-hit summaries of reports.rs:8-48) through the host-pointer ABI call (`cid_query_counts`, CID_SEQ_FASTQ,
    filter -1).  Synthetic stand-in for refs/ + SRR548019.fastq.gz, which do not travel to the GPU box.  One step = one
    search of the whole read set; single GPU (the reference's own CPU-runnable case; under torchrun only rank 0 works)."""
    import torch
    import colorid_b200 as cb
    from colorid_b200 import lib as L
    if int(os.environ.get("RANK", "0")) != 0:
        return
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    cfg = dict(CFG)
    cfg.update(coverage=30, lowq=0.02)
    if args.quick:
        cfg.update(genome_len=200_000, S=5_000_000)
    ctx = cb.Context(local)
    for kv in args.opt:
        name, _, val = kv.partition("=")
        ctx.set_option(name, int(val))
    A, Lg, rl = cfg["n_acc"], cfg["genome_len"], cfg["read_len"]
    genomes = make_genomes(torch, dev, cfg, 0xC0101D01)
    lut = torch.tensor(ASCII, device=dev, dtype=torch.uint8)
    gix = cb.Index(ctx, cfg["S"], cfg["H"], cfg["k"], A)
    offs = torch.tensor([0, Lg], device=dev, dtype=torch.int64)
    asc = [lut[genomes[a].long()].contiguous() for a in range(A)]
    torch.cuda.synchronize()
    tb = time.perf_counter()
    for a in range(A):
        gix.build_accession_dev(a, asc[a].data_ptr(), offs.data_ptr(), 1, Lg)
    gix.finalize()
    torch.cuda.synchronize()
    build_s = time.perf_counter() - tb
    del asc
    # the query: 30x paired-end reads of accession 0 (0.5 % substitutions, 2 % of the bases below -Q -> 'N' by qual_mask,
    # applied while generating like the c4 workload: the ABI takes masked sequences), no random pairs
    qcfg = dict(cfg, frac_random=0.0, n_acc=1)
    n_pairs = Lg * cfg["coverage"] // (2 * rl)
    W, K = max(1, min(args.warmup, 3)), max(1, min(args.steps, 5))
    sets = []
    for i in range(2):                                       # two read sets alternate: 2 x 100 MB of fresh input per step
        b, q = make_reads(torch, dev, qcfg, genomes[:1], n_pairs, 0xC0101D10 + i)
        b[q < cfg["qual_offset"] + 33] = 78
        h = torch.empty(b.shape, dtype=torch.uint8).pin_memory()
        h.copy_(b)
        sets.append(h.numpy().reshape(-1))
        del b, q
    seq_offs = np.arange(2 * n_pairs + 1, dtype=np.uint64) * np.uint64(rl)
    query_offs = np.array([0, 2 * n_pairs], dtype=np.uint64)
    counts = np.zeros((1, A), np.uint32)
    nk = np.zeros(1, np.uint64)
    un, us, um = (np.zeros((1, A), np.uint64) for _ in range(3))
    used = np.zeros(1, np.int64)
    P = lambda a, tp=L.u64p: a.ctypes.data_as(tp)
    lib = ctx.lib

    def search(i):
        L.check(lib.cid_query_counts(gix.h, P(sets[i % 2], L.vp), P(seq_offs), 2 * n_pairs, P(query_offs), 1, L.CID_SEQ_FASTQ, 0, -1,
                                     P(counts, L.u32p), P(nk), P(un), P(us), P(um), P(used, L.i64p)))
    for i in range(W):
        search(i)
    torch.cuda.synchronize()
    launches0 = ctx.launches
    ctx.profile(True)
    clk = ClockSampler(local)
    clk.start()
    t0 = time.perf_counter()
    for i in range(K):
        search(W + i)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = clk.stop()
    prof = ctx.profile_read()
    ctx.profile(False)
    launches = ctx.launches - launches0
    lookups = int(nk[0])
    cov0 = float(counts[0, 0]) / float(gix.n_ref[0])
    # size-independent property: a 30x read set of accession 0 recovers (nearly) all of its k-mers after the filter
    assert cov0 > 0.97 and int(used[0]) >= 1, (cov0, used)
    kern_ms = sum(v[0] for v in prof.values()) / K
    kern = {k_: {"ms_per_launch": v[0] / v[1], "launches_per_step": v[1] / K} for k_, v in prof.items()}
    peak, peak_src = measured_peak()
    R = 4 * ((A + 31) // 32)
    nb = 2 * n_pairs * rl
    alg = nb + lookups * cfg["H"] * R + 4 * A          # the read set once + H rows per surviving k-mer + counts
    # CPU: the reference's search is single-threaded (batch_search_pe.rs:9-179); bounded sample = the first `ns` pairs
    cpu = None
    if not args.no_cpu_baseline:
        try:
            from oracle import pyoracle as O
            O.lib()
            oix = oracle_index_from_dense(O, cfg, gix.download_dense(), gix.n_ref.copy())
            ns = min(n_pairs, 20_000)
            sub = sets[0][:2 * ns * rl]
            reads = [bytes(sub[j * rl:(j + 1) * rl]) for j in range(2 * ns)]
            t0 = time.perf_counter()
            o = oix.query_counts([reads], O.MODE_FASTQ, False, -1)
            dt = time.perf_counter() - t0
            g = gix.query_counts([reads], seq_mode=L.CID_SEQ_FASTQ, gene_search=False, filt=-1)
            same = all(np.array_equal(g[k_], o[k_]) for k_ in ("counts", "num_kmers", "uniq_n", "uniq_sum", "uniq_mode", "cutoff"))
            cpu = {"value": 2 * ns * rl / dt / 1e9, "unit": "query Gbp/s", "cores": 1, "kind": "port",
                   "sample": f"the first {ns} read pairs of the same read set ({2 * ns * rl / 1e6:.0f} Mbp), single thread like the "
                             "reference's search; C++ restatement of the Rust reference (oracle/)",
                   "lookups_per_s": float(o["num_kmers"][0]) / dt, "matches_gpu_report": bool(same)}
        except Exception as ex:
            cpu = {"value": None, "unit": "query Gbp/s", "cores": 0, "kind": "port", "sample": f"failed: {ex!r}"}
    line = {"metric": "search query Gbp/s", "value": K * nb / wall / 1e9, "unit": "query Gbp/s", "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": wall / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": "C1 build from FASTA + search of a 30x paired-end FASTQ read set, default report "
                                   "(BASELINE.json configs[0], synthetic stand-in)",
                       "index": f"{A} synthetic genomes x {Lg} bp, k={cfg['k']} S={cfg['S']} H={cfg['H']}",
                       "read_pairs": n_pairs, "read_len": rl, "filter": "auto_cutoff", "cutoff_used": int(used[0]),
                       "l2_policy": "two 100 MB read sets alternate; count table and 400 MB matrix exceed the 126 MB L2"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": K * nb / wall / 1e9, "unit": "query Gbp/s", "h2d_bytes_per_step": int(nb + seq_offs.nbytes),
                    "d2h_bytes_per_step": int(counts.nbytes + un.nbytes * 3 + 16),
                    "includes": "cid_query_counts with pinned host reads: H2D, k-mer counting, auto_cutoff, hashing, row gather, "
                                "unique-hit summaries, D2H (value == e2e: this workload is only measured through the host ABI)"},
            "lookups": lookups, "lookups_per_s": K * lookups / wall, "kernel_ms_per_step": kern_ms, "kernels": kern,
            "roofline": {"bound": "hbm", "kernel": "whole step (kernels only)", "achieved": alg / (kern_ms / 1e3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg / (kern_ms / 1e3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_step": alg,
                         "note": "exact k-mer counting of the read set dominates and has no algorithmic traffic beyond reading it "
                                 "(SURVEY 8d): low by construction"},
            "cpu_baseline": cpu,
            "build": {"gbp_per_s": A * Lg / build_s / 1e9, "seconds": build_s, "note": "46 FASTA accessions + transposition, device-resident bases"},
            "parity": {"hits_over_n_ref_kmers_of_source_accession": cov0}}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local, world):
    """N > 1: pin this rank (and the pinned host buffers it first-touches, and the library's host threads) to the CPU
    cores NVML reports as local to its GPU, so that 8 ranks do not pull their H2D traffic through one socket.  Returns
    the number of cores kept (0 = left alone)."""
    if world <= 1:
        return 0
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cores = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1} & allowed
        if 0 < len(cores) < len(allowed):
            os.sched_setaffinity(0, cores)
            return len(cores)
    except Exception as e:      # affinity is an optimisation only
        log(f"[numa] affinity not set: {e}")
    return 0


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import colorid_b200 as cb
    from colorid_b200 import lib as L
    from colorid_b200.api import classify_reads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    numa_cores = bind_to_gpu_numa_node(local, world)
    cfg = dict(CFG)
    if args.quick:
        cfg.update(genome_len=200_000, S=5_000_000)
    if args.n_acc:            # experiments off the headline config (e.g. 1,000 accessions: the wide-row vote kernel)
        cfg.update(n_acc=args.n_acc, n_clades=args.n_clades or max(8, args.n_acc // 5))
    batch = args.batch_pairs
    rl = cfg["read_len"]
    K, Wm = args.steps, args.warmup

    ctx = cb.Context(local)
    ctx.set_option("host_ranks", world)          # the ranks of one node share its host memory system (read_id chunk schedule)
    for kv in args.opt:
        name, _, val = kv.partition("=")
        ctx.set_option(name, int(val))
    if args.only_search:
        print(json.dumps(run_search_extra(torch, dev, ctx, args, args.quick)), flush=True)
        return
    t0 = time.time()
    # (CPU generator: the reference arm builds its index from the very same genomes, so the two arms' index checksums compare)
    genomes_cpu = make_genomes(torch, "cpu", cfg, 0xC0101D02)
    genomes = genomes_cpu.to(dev)
    lut = torch.tensor(ASCII, device=dev, dtype=torch.uint8)
    gix = cb.Index(ctx, cfg["S"], cfg["H"], cfg["k"], cfg["n_acc"])
    offs = torch.tensor([0, cfg["genome_len"]], device=dev, dtype=torch.int64)
    build_t0 = time.time()
    for a in range(cfg["n_acc"]):
        asc = lut[genomes[a].long()].contiguous()
        torch.cuda.synchronize()
        gix.build_accession_dev(a, asc.data_ptr(), offs.data_ptr(), 1, cfg["genome_len"])
    gix.finalize()
    build_s = time.time() - build_t0
    log(f"[rank {rank}] index built on device in {build_s:.2f}s ({cfg['n_acc'] * cfg['genome_len'] / build_s / 1e9:.2f} Gbp/s), "
        f"setup {time.time() - t0:.1f}s")

    # read pool: one fresh slice per step (cycled if the pool is smaller than W+K batches)
    pool_batches = min(K + Wm, args.pool_batches)
    pool_b, pool_q = make_reads(torch, dev, cfg, genomes, batch * pool_batches, 0xC0101D03 + rank)
    seq_offs_np, read_offs_np = offsets_for(batch, rl)
    d_seq_offs = torch.from_numpy(seq_offs_np.view(np.int64)).to(dev)
    d_read_offs = torch.from_numpy(read_offs_np.view(np.int64)).to(dev)
    rep_cap = args.rep_cap
    d_n_set = torch.zeros(batch, device=dev, dtype=torch.int32)
    d_flags = torch.zeros(batch, device=dev, dtype=torch.int32)
    d_rep_n = torch.zeros(batch, device=dev, dtype=torch.int32)
    d_rc = torch.zeros((batch, rep_cap), device=dev, dtype=torch.int32)
    d_rv = torch.zeros((batch, rep_cap), device=dev, dtype=torch.int32)
    params = L.ReadIdParams(1, cfg["start_sample"], cfg["qual_offset"], 16, 1, rep_cap)
    lib = ctx.lib
    max_kmers = 2 * (rl - cfg["k"] + 1)
    stream = torch.cuda.current_stream()
    # The product's read path takes PACKED reads (cid_pack_reads: quality mask + 2-bit codes + "not a base" bits, made by the
    # host parser -- the reference masks on the host too, seq.rs:36-56): the pool is packed once, on the device
    packed = not args.ascii_input
    pool_pk = None
    if packed:
        pool_pk = torch.cat([pack_reads_torch(torch, pool_b[s0:s0 + 250_000], pool_q[s0:s0 + 250_000], cfg["qual_offset"])
                             for s0 in range(0, pool_b.shape[0], 250_000)])
        wpr = pool_pk.shape[1]                         # words per read pair (19 code + 10 bad words for 2 x 150 bases)
        woffs_np = np.arange(batch + 1, dtype=np.uint64) * np.uint64(wpr)
        d_woffs = torch.from_numpy(woffs_np.view(np.int64)).to(dev)

    def step_dev(i):
        sl = i % pool_batches
        if packed:
            w = pool_pk[sl * batch:(sl + 1) * batch]
            L.check(lib.cid_read_id_batch_packed_dev(gix.h, w.data_ptr(), d_woffs.data_ptr(), 0, d_seq_offs.data_ptr(), 2 * batch,
                                                     d_read_offs.data_ptr(), batch, 2 * rl, max_kmers, C.byref(params),
                                                     d_n_set.data_ptr(), d_flags.data_ptr(), d_rep_n.data_ptr(), d_rc.data_ptr(),
                                                     d_rv.data_ptr(), stream.cuda_stream))
            return
        b = pool_b[sl * batch:(sl + 1) * batch]
        q = pool_q[sl * batch:(sl + 1) * batch]
        L.check(lib.cid_read_id_batch_dev(gix.h, b.data_ptr(), q.data_ptr(), d_seq_offs.data_ptr(), 2 * batch,
                                          2 * batch * rl, d_read_offs.data_ptr(), batch, 2 * rl, max_kmers, C.byref(params),
                                          d_n_set.data_ptr(), d_flags.data_ptr(), d_rep_n.data_ptr(), d_rc.data_ptr(),
                                          d_rv.data_ptr(), stream.cuda_stream))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ("value") ----
    # per-kernel CUDA events (ProfScope in the library, recorded on the launching stream) are live in this region
    for i in range(Wm):
        step_dev(i)
    barrier()
    launches0 = ctx.launches
    ctx.read_counter("readid_gather_rows")           # reset
    ctx.profile(True)
    clk = ClockSampler(local)
    clk.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(K):
        step_dev(Wm + i)
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    clocks = clk.stop()
    prof = ctx.profile_read()
    ctx.profile(False)
    launches = ctx.launches - launches0
    gather_rows = ctx.read_counter("readid_gather_rows") / K        # matrix rows per step the vote kernel read
    # algorithmic traffic of the last step (all steps are statistically identical)
    fl = d_flags.cpu().numpy().view(np.uint32)
    nproc_last = int(((fl >> 8) & 0xFFFF).sum())
    trunc = int(((fl >> 2) & 1).sum())
    t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    value = world * batch * K / (dev_ms_max / 1e3)

    # ---- end-to-end through the host-pointer C ABI ----
    e2e_batches = min(pool_batches, 4)
    h_b = torch.empty((e2e_batches * batch, 2 * rl), dtype=torch.uint8).pin_memory()
    h_q = torch.empty((e2e_batches * batch, 2 * rl), dtype=torch.uint8).pin_memory()
    h_b.copy_(pool_b[:e2e_batches * batch])
    h_q.copy_(pool_q[:e2e_batches * batch])
    hb, hq = h_b.numpy(), h_q.numpy()
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    o_kind = pin(batch, torch.int32)
    o_hits, o_n_set, o_n_top = (pin(batch, torch.int32).view(np.uint32) for _ in range(3))
    o_top = pin((batch, 8), torch.int32).view(np.uint32)
    n_ref = gix.n_ref.copy()
    P = lambda a, tp=L.vp: a.ctypes.data_as(tp)

    h_pk = None
    if packed:
        h_pk = torch.empty((e2e_batches * batch, wpr), dtype=torch.int32).pin_memory()
        h_pk.copy_(pool_pk[:e2e_batches * batch])
        hpk = h_pk.numpy().view(np.uint32)

    def step_e2e(i, ascii_in=False):
        """The call a user makes: host reads in, one classification per read out (parallel_vec).  Packed planes by default
        (what the CLI's parser produces); ascii_in: ASCII bases + qualities, masked and packed on the device."""
        sl = i % e2e_batches
        if packed and not ascii_in:
            w = hpk[sl * batch:(sl + 1) * batch]
            L.check(lib.cid_read_id_classify_packed(gix.h, P(w), P(woffs_np, L.u64p), 0, P(seq_offs_np, L.u64p), 2 * batch,
                                                    P(read_offs_np, L.u64p), batch, C.byref(params), P(n_ref, L.u64p), cfg["fp_correct"],
                                                    P(o_kind, L.i32p), P(o_hits, L.u32p), P(o_n_set, L.u32p), P(o_n_top, L.u32p),
                                                    P(o_top, L.u32p), 8))
        else:
            b = hb[sl * batch:(sl + 1) * batch]
            q = hq[sl * batch:(sl + 1) * batch]
            L.check(lib.cid_read_id_classify(gix.h, P(b), P(q), P(seq_offs_np, L.u64p), 2 * batch, P(read_offs_np, L.u64p),
                                             batch, C.byref(params), P(n_ref, L.u64p), cfg["fp_correct"], P(o_kind, L.i32p),
                                             P(o_hits, L.u32p), P(o_n_set, L.u32p), P(o_n_top, L.u32p), P(o_top, L.u32p), 8))
        return dict(kind=o_kind, hits=o_hits, n_set=o_n_set, n_top=o_n_top, top=o_top)

    def time_e2e(ascii_in):
        for i in range(min(Wm, 2)):
            step_e2e(i, ascii_in)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            step_e2e(i, ascii_in)
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    e2e_steps = max(1, min(K, 5))
    e2e_s = time_e2e(False)
    e2e_value = world * batch * e2e_steps / e2e_s
    h2d_ascii = 2 * batch * 2 * rl + seq_offs_np.nbytes + read_offs_np.nbytes
    h2d = (batch * wpr * 4 + woffs_np.nbytes + seq_offs_np.nbytes + read_offs_np.nbytes) if packed else h2d_ascii
    d2h = batch * (4 * 4 + 8 * 4)      # kind, hits, n_set, n_top + top[8] per read; the undecided-read list (<1 % of reads) is extra
    e2e_ascii = None
    host_pack = None
    if packed:
        e2e_ascii_s = time_e2e(True)
        e2e_ascii = {"value": world * batch * e2e_steps / e2e_ascii_s, "unit": "read pairs/s", "h2d_bytes_per_step": h2d_ascii,
                     "note": "the same reads through cid_read_id_classify: ASCII bases + qualities over PCIe, masked and packed by the kernels"}
        # the product's host packer on this rank's cores: what a parser pays to produce the packed input (not in the e2e region:
        # the reference's parser masks while it parses, before parallel_vec is called, read_id_mt_pe.rs:733-762)
        pk_words = np.zeros(batch * wpr + 64, np.uint32)
        pk_offs = np.zeros(batch + 1, np.uint64)
        pk_flags = C.c_uint32(0)
        tp = []
        for i in range(3):
            t0 = time.perf_counter()
            L.check(lib.cid_pack_reads(P(hb[:batch]), P(hq[:batch]), P(seq_offs_np, L.u64p), 2 * batch, P(read_offs_np, L.u64p), batch,
                                       cfg["qual_offset"], 0, P(pk_words, L.u32p), len(pk_words), P(pk_offs, L.u64p), C.byref(pk_flags)))
            tp.append(time.perf_counter() - t0)
        same_pack = bool(np.array_equal(pk_words[:batch * wpr], hpk[:batch].reshape(-1)) and np.array_equal(pk_offs, woffs_np))
        assert same_pack, "device-side packing of the synthetic pool differs from cid_pack_reads"
        host_pack = {"read_pairs_per_s": batch / min(tp[1:]), "threads": len(numa_cores) if numa_cores else (os.cpu_count() or 1),
                     "equals_bench_input": same_pack,
                     "note": "cid_pack_reads (AVX2, host threads of this rank) on one batch of ASCII bases + qualities"}

    # ---- N > 1: the column-sharded path (C5 shard shape: 1,250 accessions = 160-byte rows per GPU, A = 1,250 x N accessions:
    # the full C5 index at N = 8) rides along, so that the driver's scaling run records it: build Gbp/s, lookups/s, the shard
    # gather's roofline, NCCL all_gather against the fused peer-store exchange, and the comparison with an unsharded index
    c5 = None
    if world > 1 and not args.no_c5:
        try:
            del pool_b, pool_q
            torch.cuda.empty_cache()
            c5 = c5_measure(args, n_acc=(args.c5_acc or 1250 * world), steps=min(K, 5), warmup=min(Wm, 3))
            if c5 is not None:
                c5 = {k_: c5[k_] for k_ in ("metric", "value", "unit", "ms_per_step", "scaling", "config", "roofline", "kernels", "build",
                                            "parity", "fused_count_exchange", "clocks")}
        except Exception as ex:      # the headline line must still print
            c5 = {"error": repr(ex)}
            log(f"[rank {rank}] c5 block failed: {ex!r}")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peak, peak_src = measured_peak()
    H, R = cfg["H"], 4 * ((cfg["n_acc"] + 31) // 32)
    read_bytes = batch * 2 * rl            # bases of both mates ("2*L" of SURVEY 8d)
    alg = {"readid_vote": H * R * nproc_last + read_bytes,        # rows up to and incl. the first miss + the read
           "readid_kmerize": 2 * read_bytes,                       # bases + quals
           "readid_order": 0}
    traffic = ncu_traffic()
    kern = {}
    for name, (ms, n) in prof.items():
        per = ms / n
        kern[name] = {"ms_per_launch": per, "launches_per_step": n / K, "algorithmic_bytes_per_launch": alg.get(name),
                      "gbs": (alg[name] / (per / 1e3) / 1e9) if alg.get(name) else None}
    # The vote of narrow rows is a STAGE of kernels when the matrix spans several L2-sized windows (cid_readid_part.cu):
    # scan (hash, first miss, one (row, slot) tuple per gather, bucketed by row window) -> one gather launch per window (rows
    # read from the L2-resident window) -> count -> the one-kernel vote on the reads left over.  The stage does the work the
    # reference's search_index does per read, so the roofline line is quoted on the stage: its algorithmic bytes over the
    # summed duration of its launches in one step.
    VOTE_PARTS = ("readid_vp_scan", "readid_vp_gather", "readid_vp_count", "readid_vote")
    parted = "readid_vp_scan" in prof
    vote_ms = sum(prof[k_][0] for k_ in VOTE_PARTS if k_ in prof) / K if any(k_ in prof for k_ in VOTE_PARTS) else None
    if parted:
        kern["readid_vote"]["algorithmic_bytes_per_launch"] = None
        kern["readid_vote"]["gbs"] = None
        kern["readid_vote"]["note"] = "leftover reads only (more than 8 candidate colours)"
        dom, dom_ms, dom_traffic = "readid vote stage (readid_vp_scan + readid_vp_gather per window + readid_vp_count + readid_vote)", vote_ms, \
            sum(traffic.get(k_, 0) * kern[k_]["launches_per_step"] for k_ in VOTE_PARTS if k_ in kern and traffic.get(k_)) or None
        achieved = alg["readid_vote"] / (vote_ms / 1e3) / 1e9
        dom_alg = alg["readid_vote"]
    else:
        dom = max(prof, key=lambda k_: prof[k_][0])
        dom_ms = prof[dom][0] / prof[dom][1]
        achieved = alg.get(dom, 0) / (dom_ms / 1e3) / 1e9
        dom_traffic, dom_alg = traffic.get(dom), alg.get(dom)
    gathers = gather_rows                  # one gather = one 8-byte row read at a random row (device counter)
    gather_ms = prof["readid_vp_gather"][0] / K if parted else None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": dom_traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_alg, "ms_per_step": dom_ms,
                "measured_in": "the timed region itself: CUDA events around every launch on the launching stream",
                "note": "8-byte rows: a gather moves 8 algorithmic bytes but costs one memory access, so bytes are the wrong "
                        "yardstick for this shape (SURVEY 8d, R < 32) and the access RATE is reported next to them.  Read "
                        "straight from the 400 MB matrix the rows come at the DRAM random-access rate (54.5 G/s, "
                        "profiles/r1_gather_probe.txt; the one-kernel vote of round 1 ran at 82 % of it); bucketed by row "
                        "window they come from L2 (275 G/s for a 64 MB window, profiles/r2_l2_window_probe.txt) and the "
                        "stage is bound by the instructions of its scan kernel (XXH3 x num_hash per k-mer), not by memory.  "
                        "Reads whose candidate set is empty after the first -B k-mers only need the first-miss position, "
                        "which the L2-resident row-present bitmap answers: those k-mers count in the algorithmic bytes "
                        "(the reference gathers their rows) but not in gathers_per_launch",
                "random_access": {"gathers_per_launch": gathers,
                                  "achieved_g_per_s": gathers / (vote_ms / 1e3) / 1e9 if vote_ms else None,
                                  "ceiling_g_per_s": GATHER_CEILING_G_PER_S,
                                  "frac": gathers / (vote_ms / 1e3) / 1e9 / GATHER_CEILING_G_PER_S if vote_ms else None,
                                  "ceiling_source": "tools/gather_probe.cu on B200: 8-byte __ldg at uniformly random rows of a "
                                                    "400 MB matrix, 8 loads in flight per thread (profiles/r1_gather_probe.txt)"},
                "window_gather": ({"tuples_per_step": gathers, "ms_per_step": gather_ms,
                                   "g_tuples_per_s": gathers / (gather_ms / 1e3) / 1e9,
                                   "l2_window_ceiling_g_per_s": 275.0,
                                   "source": "readid_vp_gather launches only; ceiling: tools/l2_window_probe.cu (profiles/r2_l2_window_probe.txt)"}
                                  if parted and gather_ms else None),
                "vote_sector_granular_gbs": (H * 32 * nproc_last + read_bytes) / (vote_ms / 1e3) / 1e9 if vote_ms else None,
                "kernels": kern,
                "share_of_step": {k_: prof[k_][0] / sum(v[0] for v in prof.values()) for k_ in prof}}

    # ---- CPU baseline on a bounded sample (rank 0) ----
    cpu = None
    if not args.no_cpu_baseline:
        try:
            from oracle import pyoracle as O
            O.lib()
            cores = os.cpu_count() or 1
            # the oracle builds ITS OWN index from the same genomes (full size: S = 50 M), compared word for word with the GPU's
            tb0 = time.time()
            oix = O.Index(cfg["S"], cfg["H"], cfg["k"], cfg["n_acc"])
            lut_np = np.array(ASCII, dtype=np.uint8)
            oix.build_many([[lut_np[genomes_cpu[a].numpy()].tobytes()] for a in range(cfg["n_acc"])], O.MODE_FASTA, threads=cores)
            dense = gix.download_dense()
            index_same = bool(np.array_equal(oix.words(), dense) and np.array_equal(oix.n_ref, n_ref))
            index_sha = hashlib.sha256(dense.tobytes()).hexdigest()
            log(f"[rank 0] oracle index built on {cores} threads in {time.time() - tb0:.1f}s; equals the GPU index: {index_same}")
            del dense
            ns = min(batch, 2000)
            dt, _ = oracle_read_id(O, oix, cfg, hb[:ns], hq[:ns], cores)
            ns2 = int(min(batch, max(ns, ns / dt * 12.0)))      # ~12 s of CPU work
            dt2, ocls = oracle_read_id(O, oix, cfg, hb[:ns2], hq[:ns2], cores)
            # while we are here: the e2e classifications of these reads must equal the oracle's
            chk = step_e2e(0)
            ntop = np.minimum(ocls["n_top"].astype(np.int64), 8)
            colmask = np.arange(8)[None, :] < ntop[:, None]
            same = bool(np.array_equal(chk["kind"][:ns2], ocls["kind"]) and np.array_equal(chk["hits"][:ns2], ocls["hits"])
                        and np.array_equal(chk["n_top"][:ns2], ocls["n_top"]) and np.array_equal(chk["n_set"][:ns2], ocls["n_set"])
                        and np.array_equal(np.where(colmask, chk["top"][:ns2], 0), np.where(colmask, ocls["top"][:, :8], 0)))
            cpu = {"value": ns2 / dt2, "unit": "read pairs/s", "cores": cores, "kind": "port",
                   "sample": f"{ns2} read pairs of the same workload; C++ restatement of the Rust reference (oracle/), "
                             f"read_id parallel over reads on {cores} threads",
                   "matches_gpu_classification": same,
                   "checked": "kind, hits, n_set, n_top and the top accession ids of every sampled read",
                   "index_built_by_oracle_equals_gpu_index": index_same, "index_sha256": index_sha}
        except Exception as ex:      # the bench line must still print
            cpu = {"value": None, "unit": "read pairs/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}

    search_c3 = None
    build_c4 = None
    if world == 1 and not args.no_search:
        try:
            del pool_b, pool_q, h_b, h_q
            torch.cuda.empty_cache()
            search_c3 = run_search_extra(torch, dev, ctx, args, args.quick)
        except Exception as ex:
            search_c3 = {"error": repr(ex)}
        # the third metric of BASELINE.json, build Gbp/s on the C4 shape (30x read sets, auto_cutoff), rides along as well:
        # the c4 workload in a process of its own on the same GPU, its line embedded
        try:
            import subprocess
            cmd = [sys.executable, os.path.abspath(__file__), "--workload", "c4", "--c4-acc", "10", "--steps", "6", "--warmup", "3"]
            if args.quick:
                cmd.append("--quick")
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=420)
            c4 = json.loads(res.stdout.strip().splitlines()[-1])
            build_c4 = {k_: c4.get(k_) for k_ in ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "config", "roofline", "kernels",
                                                   "table_inserts_per_s", "auto_cutoff_used", "parity", "clocks")}
        except Exception as ex:
            build_c4 = {"error": repr(ex)}

    line = {"metric": "read_id read pairs/s", "value": value, "unit": "read pairs/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic", "config": dict(workload_config(cfg, batch), input="packed reads (cid_pack_reads)" if packed else "ASCII"),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "read pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "host_cores_bound_per_rank": numa_cores or None,
                    "h2d_gbs_per_rank": h2d * e2e_steps / e2e_s / 1e9,
                    "input": ("packed reads in pinned host memory (cid_pack_reads planes: 2-bit codes + not-a-base bits, quality mask "
                              "applied by the packer on the host like seq.rs:36-56)" if packed else "ASCII bases + qualities in pinned host memory"),
                    "includes": ("cid_read_id_classify_packed" if packed else "cid_read_id_classify") + ": chunked pipeline of H2D reads+offsets, read_id kernels + device vote, D2H of one classification per read; near-threshold/tied reads re-voted on the host"},
            "e2e_ascii": e2e_ascii, "host_pack": host_pack,
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "build": {"gbp_per_s": cfg["n_acc"] * cfg["genome_len"] / build_s / 1e9, "seconds": build_s,
                      "note": "index build on device incl. per-accession host sync; not the timed metric"},
            "report_truncated_reads_last_step": trunc, "search_c3": search_c3, "build_c4": build_c4, "c5": c5}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-pairs", type=int, default=1_000_000)
    ap.add_argument("--pool-batches", type=int, default=10, help="distinct batches kept in HBM (10 x 1M pairs = C2's 10M)")
    ap.add_argument("--rep-cap", type=int, default=47, help="report entries kept per read (n_acc+1 = never truncated)")
    ap.add_argument("--ref-batch", type=int, default=200_000, help="read pairs per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-search", action="store_true", help="skip the C3 gene-search extra measurement")
    ap.add_argument("--only-search", action="store_true", help="profiling aid: run only the C3 gene-search measurement")
    ap.add_argument("--opt", action="append", default=[], help="library tuning option name=value (cid_ctx_set_option), repeatable")
    ap.add_argument("--quick", action="store_true", help="small genomes / bloom filter (functional check only)")
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c4", "c5"], help="c1 = FASTA build + FASTQ search; c2 = read_id headline (default); c4 = build from read sets; c5 = column-sharded build + search")
    ap.add_argument("--c4-acc", type=int, default=0, help="accessions per rank built by the c4 workload (default 24)")
    ap.add_argument("--n-acc", type=int, default=0, help="c2: accessions of the index (default 46 = the headline config)")
    ap.add_argument("--n-clades", type=int, default=0, help="c2 with --n-acc: clades (default n_acc / 5)")
    ap.add_argument("--c5-acc", type=int, default=0, help="total accessions of the c5 workload (default 10,000)")
    ap.add_argument("--ascii-input", action="store_true", help="c2: time the ASCII entry points (bases + qualities masked and packed on the device) instead of packed reads")
    ap.add_argument("--no-c5", action="store_true", help="N > 1: skip the column-sharded C5 block")
    ap.add_argument("--c5-extra", action="store_true", help="c5, N > 1: also time the column-sharded default report and read_id")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        run_c5(args)
    elif args.workload == "c4":
        run_c4(args)
    elif args.workload == "c1":
        run_c1(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
