#!/usr/bin/env python3
"""Benchmark of the index-and-scan hot path (BASELINE.json metric:
SA+BWT+LCP Gbases/s and FM backward-search queries/s at 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload = BASELINE.json configs[4]: a synthetic 24-contig human-scale genome (hg38 contig
lengths, 3.09 Gb), every contig indexed (C array, suffix array, BWT + Occ checkpoints, LCP)
and Tier 1 + Tier 2 scanned (Tier 1 tandems, strict adjacency u <= 1000, LCP plateaus, period
scan).  Contigs are dealt to the N ranks by longest-processing-time bin packing (strong
scaling: the genome is fixed, every rank gets its share) and the int32 record rows of all
ranks are gathered to rank 0 over NCCL inside the timed region.

A step = one pass over the whole genome.
  value : texts already resident in HBM, rows stay on the device, gather included.
  e2e   : the same pass through bwt_algorithm_b200.genome.GenomeScanner from PINNED HOST
          buffers: every contig's text is uploaded (1 B/base), the index stays in HBM, every
          row is downloaded to pinned host memory (32 B/row) -- uploads, downloads and the
          NCCL gather all inside the timed region.
Inputs are far larger than the 126 MB L2 (each contig touches >= 2 GB), so no flush is needed
between genome steps; the single-contig sections below flush L2 (256 MB write) between steps.

Explaining sections (rank 0): per-kernel profile + roofline of the dominant kernel of the
index build on a chr21-sized contig (configs[2]); the full-index download path (SA + BWT + LCP
to the host, 9 B/base) with its PCIe rate; FM backward search on the chr1-sized contig
(configs[3]: all 1..10 bp motifs and 2^26 random 10-mers) with a sector roofline; the
detectors one by one; a CPU baseline (the oracle port, and the Python reference itself when
baseline/_ref/bwt.py travelled with the repo) timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONTIG_N = 46_709_983          # hg38 chr21 length (BASELINE.json configs[2])
HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
        133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
        58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
BYTES_PER_BASE_COMPULSORY = 20.2   # SA 5 + BWT/C/Occ 6.19 + LCP 9
SCAN_BYTES_PER_BASE = 4.0          # Tier 1, strict, plateaus (+SA/LCP reads are counted in the index figure), period: 1 B/base each
WORKLOAD = "24 synthetic contigs with hg38 lengths, 3.09 Gb (BASELINE configs[4]): index + Tier 1 + Tier 2 per contig"


def gen_contig(n, seed, sub_rate=0.03):
    """SURVEY.md Appendix B (host generator; used for the CPU samples)."""
    rng = np.random.default_rng(seed)
    A = np.frombuffer(b"ACGT", dtype=np.uint8)
    s = A[rng.integers(0, 4, n)]
    pos = 500
    while pos < n - 400:
        k = int(rng.integers(1, 10))
        m = A[rng.integers(0, 4, k)]
        c = int(rng.integers(4, 25))
        arr = np.tile(m, c)
        mut = rng.random(arr.size) < sub_rate
        arr[mut] = A[rng.integers(0, 4, int(mut.sum()))]
        s[pos:pos + arr.size] = arr
        pos += arr.size + int(rng.integers(300, 1500))
    return s


def device_contig(torch, n, seed, dev):
    """Appendix B's generator vectorised on the device (random ACGT + a planted imperfect tandem
    array every ~900 bp, 3 % substitutions) + '$'.  Setup only: the 3.09 Gb genome would take
    minutes in the host generator's Python loop."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    text = lut[torch.randint(0, 4, (n,), generator=g, device=dev)]
    K = max(1, n // 970)
    gaps = torch.randint(300, 1500, (K,), generator=g, device=dev)
    k = torch.randint(1, 10, (K,), generator=g, device=dev)
    c = torch.randint(4, 25, (K,), generator=g, device=dev)
    lens = k * c
    starts = 500 + torch.cumsum(gaps + lens, 0) - lens
    keep = starts + lens < n - 400
    starts, k, lens = starts[keep], k[keep], lens[keep]
    if starts.numel():
        motifs = torch.randint(0, 4, (starts.numel(), 9), generator=g, device=dev)
        arr = torch.repeat_interleave(torch.arange(starts.numel(), device=dev), lens)
        first = torch.cumsum(lens, 0) - lens
        j = torch.arange(arr.numel(), device=dev) - first[arr]
        val = motifs[arr, j % k[arr]]
        mut = torch.rand(arr.numel(), generator=g, device=dev) < 0.03
        val = torch.where(mut, torch.randint(0, 4, (arr.numel(),), generator=g, device=dev), val)
        text[starts[arr] + j] = lut[val]
        del motifs, arr, first, j, val, mut
    return torch.cat([text, torch.tensor([36], dtype=torch.uint8, device=dev)])


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from this round's committed ncu capture of this same
    command (profiles/r2_ncu_traffic.json, written by tools/ncu_traffic.py); None when the
    kernel has no capture."""
    for name in ("r2_ncu_traffic.json",):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                d = json.load(f)
            if kernel in d:
                return d[kernel]["traffic_bytes_per_launch"], name
        except Exception:
            pass
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- ours
class IndexStep:
    """Pre-allocated buffers + the C-ABI call of one index build (configs[2] section)."""

    def __init__(self, torch, L, lib, n, device):
        self.torch, self.L, self.lib, self.n, self.dev = torch, L, lib, n, device
        self.sa = torch.empty(n, dtype=torch.int32, device=device)
        self.bwt = torch.empty(n, dtype=torch.uint8, device=device)
        self.lcp = torch.empty(n, dtype=torch.int32, device=device)
        self.nrows = 5
        self.ncp = n // 128 + 1 + (1 if n % 128 else 0)
        self.occ = torch.zeros((self.nrows, self.ncp), dtype=torch.int32, device=device)
        self.ws = torch.empty(int(L.bwtk_index_workspace_bytes(n)), dtype=torch.uint8, device=device)
        self.totals = np.zeros(256, np.int64)
        self.stats = np.zeros(8, np.int64)
        self.row = np.full(256, -1, np.int32)

    def run(self, d_text):
        """One fused call: C array + SA + BWT + Occ + LCP (bwtk_index_build)."""
        L, lib, n = self.L, self.lib, self.n
        lib.check(L.bwtk_index_build(d_text.data_ptr(), n, 128, self.sa.data_ptr(), None,
                                     self.bwt.data_ptr(), self.occ.data_ptr(), self.nrows, self.lcp.data_ptr(),
                                     self.totals.ctypes.data, self.row.ctypes.data, self.stats.ctypes.data,
                                     self.ws.data_ptr(), self.ws.numel(), lib.stream_ptr()), "index_build")


def profile_report(L):
    buf = C.create_string_buffer(1 << 16)
    L.bwtk_profile_report(buf, 1 << 16)
    rows = []
    for ln in buf.value.decode().splitlines():
        nm, cnt, ms, by = ln.split("\t")
        rows.append({"kernel": nm, "launches": int(cnt), "ms": float(ms), "algo_bytes": int(by)})
    return rows


def fm_section(torch, L, lib, dev, n, bwt, occ, ncp, totals, row, flush, args):
    """configs[3]: every 1..10 bp motif and 2^26 random 10-mers against the index of the
    chr1-sized contig (BWT + Occ resident in HBM; the packed search index is built from them)."""
    from bwt_algorithm_b200.device_index import FMReplica

    rep = FMReplica(bwt[:n], occ.view(-1)[: int((row >= 0).sum()) * ncp].view(-1, ncp), totals, row, n, 128, device=dev)
    out = {"index_bases": n - 1}

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    t0 = time.perf_counter()
    rep.ensure_packed()
    torch.cuda.synchronize()
    out["packed_index"] = {"bytes": int(rep.packed_bytes), "build_ms": round((time.perf_counter() - t0) * 1e3, 3),
                           "layout": "64-B blocks: 4 x u32 counts + 192 x 2-bit symbols, exceptions ($, N, IUPAC) listed"}
    kmax = 10
    total_m = (4 ** (kmax + 1) - 4) // 3
    ms = timed(lambda: rep.motif_sweep(kmax))
    out["motif_sweep_1_10"] = {"queries": total_m, "ms": round(ms, 4), "queries_per_s": round(total_m / (ms * 1e-3), 1),
                               "sectors_per_query_model": 2, "note": "trie sweep: one LF step (2 ranks) per motif"}
    nq = 1 << args.fm_log2_queries
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    pats = torch.empty((nq, 10), dtype=torch.uint8, device=dev)
    chunk = 1 << 22
    for a in range(0, nq, chunk):       # uniform random 10-mers over ACGT, generated on the device
        b = min(nq, a + chunk)
        pats[a:b] = lut[torch.randint(0, 4, (b - a, 10), generator=g, device=dev)]
    lens = torch.full((nq,), 10, dtype=torch.int32, device=dev)
    sp = ep = None

    def run_batch():
        nonlocal sp, ep
        sp, ep = rep.search_device(pats, 10, lens, nq)

    ms = timed(run_batch)
    found = int((sp >= 0).sum().item())
    ktab = int(rep.ftab_k or 0)          # patterns start from the interval of their last ktab characters
    sect = 2 * (10 - ktab)               # two ranks per remaining character, one 64-B block (two 32-B sectors) each
    bytes_q = sect * 64 + (8 if ktab else 0) + 10 + 8
    peak, _ = measured_peak()
    out["random_10mers"] = {"queries": nq, "ms": round(ms, 4), "queries_per_s": round(nq / (ms * 1e-3), 1),
                            "found": found, "kmer_table": ktab,
                            "roofline": {"bound": "hbm-sectors", "unit": "GB/s",
                                         "blocks_per_query": sect, "bytes_per_query": bytes_q,
                                         "achieved": round(nq * bytes_q / (ms * 1e-3) / 1e9, 1), "peak": peak,
                                         "frac": round(nq * bytes_q / (ms * 1e-3) / 1e9 / peak, 4),
                                         "note": "packed index (62 MB of blocks for chr1) is L2-resident: frac may exceed 1; "
                                                 "see profiles/ for lts__t_sectors and the L2 hit rate"}}
    want_sp, want_ep = sp.clone(), ep.clone()
    variants = {}
    fx = rep.ensure_packed()
    fx.ftab_k = 0
    vms = timed(run_batch, reps=2)
    fx.ftab_k = ktab
    variants["no_kmer_table"] = {"ms": round(vms, 4), "queries_per_s": round(nq / (vms * 1e-3), 1),
                                 "same_answers": bool(torch.equal(sp, want_sp) and torch.equal(ep, want_ep)),
                                 "note": "all 10 LF steps (20 rank blocks per query) through the packed index"}
    for name, attrs in (("thread_per_query", {"thread_per_query": True}), ("no_l2_window", {"l2_persist": False}),
                        ("byte_bwt_occ_rows", {"use_packed": False})):
        for k, v in attrs.items():
            setattr(rep, k, v)
        vms = timed(run_batch, reps=2)
        same = bool(torch.equal(sp, want_sp) and torch.equal(ep, want_ep))
        variants[name] = {"ms": round(vms, 4), "queries_per_s": round(nq / (vms * 1e-3), 1), "same_answers": same}
        for k in attrs:
            setattr(rep, k, getattr(type(rep), k))
    out["random_10mers"]["variants"] = variants
    del pats, lens, sp, ep, want_sp, want_ep
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200 import _lib as lib
    from bwt_algorithm_b200 import sharding
    from bwt_algorithm_b200.genome import KINDS, GenomeScanner

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = lib.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the genome: generated on the device, parked in pinned host memory (setup, untimed) ----
    lengths = [max(2000, int(x * args.scale)) for x in HG38]
    total_bases = sum(lengths)
    bins = sharding.lpt_partition(lengths, world)
    mine = bins[rank]                       # longest first
    dev_texts, host_texts = [], []
    for i in mine:
        t = device_contig(torch, lengths[i], 1000 + i, dev)
        dev_texts.append(t)
        h = torch.empty(t.numel(), dtype=torch.uint8, pin_memory=True)
        h.copy_(t)
        host_texts.append(h)
    torch.cuda.synchronize()
    my_bases = sum(lengths[i] for i in mine)
    max_n = max([lengths[i] for i in mine] + [2000]) + 1
    arena_rows = my_bases // 14 + 262144 * max(len(mine), 1)   # small contigs: the period scan's raw rows dominate
    scanner = GenomeScanner(max_n, arena_rows, device=dev, lanes=args.lanes)
    all_rows_cap = total_bases // 14 + 262144 * len(lengths)
    gather_buf = torch.empty((all_rows_cap, 8), dtype=torch.int32, device=dev) if (rank == 0 and world > 1) else None
    kind_id = {k: j for j, k in enumerate(KINDS)}

    def genome_step(texts, download):
        res = scanner.scan(texts, ids=mine, download=download)
        table = [(r.contig, kind_id[k], a, c) for r in res for k, (a, c) in r.span.items()]
        if world > 1:
            rows, merged = sharding.gather_rows_to_rank0(scanner.arena[: scanner.used], table, out=gather_buf)
        else:
            rows, merged = scanner.arena[: scanner.used], table
        return res, rows, merged

    def timed_steps(texts, download, steps):
        barrier()
        launches0 = L.bwtk_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        scan_ms = 0.0
        for _ in range(steps):
            res, rows, merged = genome_step(texts, download)
            scan_ms += scanner.elapsed_ms
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = L.bwtk_launch_count() - launches0
        barrier()
        return ms, scan_ms, launches, res, rows, merged

    for _ in range(args.warmup):            # warm-up: NCCL channels, first-call kernel attributes, both paths
        genome_step(host_texts, True)
    genome_step(dev_texts, False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    scanner.profile_phases = True          # one untimed pass with events around every phase (rank-local breakdown)
    genome_step(dev_texts, False)
    phase_ms = dict(scanner.phase_ms)
    scanner.profile_phases = False
    dev_ms, dev_scan_ms, launches, res, rows, merged = timed_steps(dev_texts, False, args.steps)
    e2e_ms, e2e_scan_ms, e2e_launches, res, rows, merged = timed_steps(host_texts, True, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    h2d_b, d2h_b = scanner.h2d_bytes, scanner.d2h_bytes
    my_rows = scanner.used
    per_kind = {k: sum(r.span[k][1] for r in res if k in r.span) for k in KINDS}
    sa_stats_big = res[0].sa_stats if res else np.zeros(8, np.int64)

    t = torch.tensor([dev_ms, e2e_ms, dev_scan_ms, e2e_scan_ms, float(h2d_b), float(d2h_b), float(my_rows),
                      float(my_bases), float(launches)] + [float(per_kind[k]) for k in KINDS],
                     dtype=torch.float64, device=dev)
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
    else:
        allt = [t]
    allt = [x.cpu().numpy() for x in allt]
    gathered_rows = int(rows.shape[0]) if rank == 0 else 0
    del gather_buf
    # free the genome before the single-contig sections
    del dev_texts
    torch.cuda.empty_cache()

    extras = {}
    roof = None
    peak, peak_src = measured_peak()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # ---- full-index download path (SA + BWT + LCP to the host), all ranks at once --------------
    if not args.skip_extras:
        from bwt_algorithm_b200.streaming import IndexPipeline
        n21 = min(CONTIG_N, max(lengths)) + 1
        text21 = device_contig(torch, n21 - 1, 21 + rank, dev)
        pinned21 = torch.empty(n21, dtype=torch.uint8, pin_memory=True)
        pinned21.copy_(text21)
        pipe = IndexPipeline(n21, occ_rate=128, device=dev, slots=2, want_lcp=True)
        pl_h2d, pl_d2h = pipe.bytes_per_contig(n21)
        for _ in range(2):
            r = pipe.result(pipe.submit(pinned21))
        assert int(r["sa"][0]) == n21 - 1 and r["lcp"].shape[0] == n21     # '$' sorts first
        pipe.drain()
        barrier()
        first = pipe.slots[pipe.tickets % len(pipe.slots)]
        p0 = torch.cuda.Event(enable_timing=True)
        p0.record(first.stream)
        prev = None
        psteps = 6
        for _ in range(psteps):
            tk = pipe.submit(pinned21)
            if prev is not None:
                r = pipe.result(prev)
                _ = int(r["sa"][0])       # the host consumes the previous contig while this one is in flight
            prev = tk
        r = pipe.result(prev)
        pipe.drain()
        pl_ms = max(p0.elapsed_time(sl.done) for sl in pipe.slots)
        pt = torch.tensor([pl_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(pt, op=dist.ReduceOp.MAX)
        pl_ms = float(pt.item())
        extras["index_arrays_to_host"] = {
            "what": "chr21-sized contig per rank: text up, SA + BWT + LCP (9 B/base) down to pinned host memory, "
                    "two contigs in flight (streaming.IndexPipeline); all ranks at once",
            "ms_per_contig": round(pl_ms / psteps, 4),
            "gbases_per_s": round(world * (n21 - 1) * psteps / (pl_ms * 1e-3) / 1e9, 4),
            "h2d_bytes": pl_h2d, "d2h_bytes": pl_d2h,
            "d2h_gb_per_s_per_link": round(pl_d2h * psteps / (pl_ms * 1e-3) / 1e9, 2),
            "d2h_gb_per_s_all_links": round(world * pl_d2h * psteps / (pl_ms * 1e-3) / 1e9, 2),
            "ceiling": "PCIe Gen5 x16 per GPU (~55 GB/s measured D2H on one link); with 8 ranks the host's "
                       "memory system caps the sum (all GPUs of this box sit on NUMA node 0)"}
        del pipe, pinned21
        torch.cuda.empty_cache()

    if rank == 0 and not args.skip_extras:
        # ---- configs[2]: per-kernel profile of one index build (events around every launch) ----
        d_text = text21
        step = IndexStep(torch, L, lib, n21, dev)
        for _ in range(3):
            step.run(d_text)
        torch.cuda.synchronize()
        times = []
        for _ in range(5):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step.run(d_text)
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        idx_ms = statistics.median(times)
        L.bwtk_profile_enable(1)
        step.run(d_text)
        prof = profile_report(L)
        L.bwtk_profile_enable(0)
        tot_ms = sum(r["ms"] for r in prof) or 1.0
        for r in prof:
            r["share"] = round(r["ms"] / tot_ms, 4)
            r["gbs"] = round(r["algo_bytes"] / (r["ms"] * 1e-3) / 1e9, 1) if r["ms"] > 0 else None
            r["ms"] = round(r["ms"], 4)
        top = max(prof, key=lambda r: r["ms"])
        achieved = top["algo_bytes"] / (top["ms"] * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(top["kernel"])
        roof = {"bound": "hbm", "kernel": top["kernel"], "achieved": round(achieved, 1), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src,
                "launches_per_step": top["launches"], "avg_launch_ms": round(top["ms"] / top["launches"], 4),
                "algo_bytes_per_launch": top["algo_bytes"] // max(top["launches"], 1),
                "workload": "index build of one chr21-sized contig (BASELINE configs[2]), L2 flushed between steps"}
        algo_total = sum(r["algo_bytes"] for r in prof)
        extras["index_build_chr21"] = {
            "ms": round(idx_ms, 4), "gbases_per_s": round((n21 - 1) / (idx_ms * 1e-3) / 1e9, 4),
            "compulsory_fraction": round(BYTES_PER_BASE_COMPULSORY * (n21 - 1) / (idx_ms * 1e-3) / 1e9 / peak, 5),
            "algorithmic_fraction": round(algo_total / (tot_ms * 1e-3) / 1e9 / peak, 4),
            "sa_stats": {"rounds": int(step.stats[0]), "bits_per_symbol": int(step.stats[1]),
                         "symbols_per_key": int(step.stats[2]), "active_after_round0": int(step.stats[3]),
                         "sum_active": int(step.stats[4]), "radix_passes": int(step.stats[5])}}
        extras["kernels"] = prof

        # ---- detectors one by one on the same contig (device-timed) ----------------------------
        from bwt_algorithm_b200 import detect
        scan = {}
        for name, fn in (
            ("tier1", lambda: detect.tier1_rows(d_text)),
            ("strict_u1_1000", lambda: detect.strict_rows(d_text, 1, 1000, 0, 3)),
            ("lcp_plateaus", lambda: detect.plateau_rows(d_text, step.sa, step.lcp, 1, 1000, 3)[0]),
            ("period_scan", lambda: detect.period_scan_rows(d_text)[0]),
        ):
            fn()                      # warm-up: allocator blocks, first-call kernel attributes
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rws = fn()
            b.record()
            torch.cuda.synchronize()
            dt = a.elapsed_time(b) * 1e-3
            scan[name] = {"records": int(len(rws)), "ms": round(dt * 1e3, 3),
                          "gbases_per_s": round((n21 - 1) / dt / 1e9, 4),
                          "timing": "CUDA events, second call, rows downloaded to pinned host memory inside"}
        extras["scan"] = scan
        del step, d_text, text21
        torch.cuda.empty_cache()

        # ---- configs[3]: FM search on the chr1-sized contig ------------------------------------
        big = max(range(len(lengths)), key=lambda i: lengths[i])
        if big in mine:
            ht = host_texts[mine.index(big)]
        else:
            ht = device_contig(torch, lengths[big], 1000 + big, dev).cpu()
        scanner.kinds = ()
        scanner.scan([ht], download=False)          # index only; BWT + Occ stay in the scanner's buffers
        nbig = int(ht.numel())
        ncp = nbig // 128 + 1 + (1 if nbig % 128 else 0)
        extras["fm_search_chr1"] = fm_section(torch, L, lib, dev, nbig, scanner.bwt, scanner.occ, ncp,
                                              scanner.totals.copy(), scanner.row.copy(), flush, args)

    if rank == 0:
        steps = args.steps
        dev_ms_max = max(float(x[0]) for x in allt)
        e2e_ms_max = max(float(x[1]) for x in allt)
        value = total_bases * steps / (dev_ms_max * 1e-3) / 1e9
        e2e_val = total_bases * steps / (e2e_ms_max * 1e-3) / 1e9
        h2d_total = int(sum(float(x[4]) for x in allt))
        d2h_total = int(sum(float(x[5]) for x in allt))
        rows_total = int(sum(float(x[6]) for x in allt))
        out = {
            "metric": "SA+BWT+LCP Gbases/s", "value": round(value, 4), "unit": "Gbases/s", "n_gpus": world,
            "steps": steps, "warmup": args.warmup, "ms_per_step": round(dev_ms_max / steps, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8/int32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "bases": total_bases, "contigs": len(lengths), "scale": args.scale,
                       "generator": "SURVEY Appendix B density, vectorised on the device, seed 1000+i; parked in pinned host memory",
                       "detectors": "tier1(9,3,6,1.0) strict(u 1..1000, mm 0, copies 3) lcp_plateaus(1,1000,3) "
                                    "period_scan(1..1000, masked by Tier 1)",
                       "occ_rate": 128, "contigs_in_flight_per_gpu": args.lanes, "l2": "inputs larger than L2: every contig touches >= 2 GB (no flush between genome steps)",
                       "parallelism": f"contigs dealt to {world} rank(s) by LPT bin packing; rows gathered to rank 0 over NCCL",
                       "per_rank_bases": [int(float(x[7])) for x in allt]},
            "e2e": {"value": round(e2e_val, 4), "unit": "Gbases/s", "h2d_bytes_per_step": h2d_total,
                    "d2h_bytes_per_step": d2h_total, "ms_per_step": round(e2e_ms_max / steps, 3),
                    "api": "bwt_algorithm_b200.genome.GenomeScanner.scan(pinned host texts) + sharding.gather_rows_to_rank0 (NCCL)",
                    "rows_per_step": rows_total, "rows_gathered_on_rank0": gathered_rows,
                    "per_rank_ms": [round(float(x[1]) / steps, 3) for x in allt],
                    "per_rank_scan_ms": [round(float(x[3]) / steps, 3) for x in allt],
                    "note": "every rank uploads its contigs and downloads its own rows over its own PCIe link; "
                            "the NCCL gather leaves all rows in rank 0's HBM; index arrays stay on the device"},
            "per_rank_ms": [round(float(x[0]) / steps, 3) for x in allt],
            "per_rank_scan_ms": [round(float(x[2]) / steps, 3) for x in allt],
            "rows": {k: int(sum(float(x[9 + j]) for x in allt)) for j, k in enumerate(KINDS)},
            "phase_ms_rank0": {k: round(v, 3) for k, v in phase_ms.items()},
            "gpu_launches": int(sum(float(x[8]) for x in allt)),
            "clocks": clocks,
            "roofline": roof,
            "compulsory_fraction_of_aggregate_hbm": round(
                (BYTES_PER_BASE_COMPULSORY + SCAN_BYTES_PER_BASE) * value / (peak * world), 5),
            "sa_stats_largest_contig": {"rounds": int(sa_stats_big[0]), "active_after_round0": int(sa_stats_big[3]),
                                        "sum_active": int(sa_stats_big[4]), "radix_passes": int(sa_stats_big[5])},
        }
        out.update(extras)
        if not args.no_cpu and world == 1:
            out["cpu_baseline"] = cpu_baseline(args, bounded_s=25.0)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------- CPU
WINDOW = 20_000     # detectors on 20 kb windows treated as standalone contigs
REF_WINDOW = 10_000 # per step of the reference arm (the driver asks for 25 steps: ~6 s of CPU work each)


def cpu_window_once(seed_window):
    """The genome workload on one window treated as a standalone contig, oracle port
    (oracle/oracle.c): index (SA, BWT, Occ, LCP) + Tier 1 + strict adjacency u <= 1000 + LCP
    plateaus + period scan."""
    from oracle import oracle as orc

    seed, WINDOW = seed_window
    text = gen_contig(WINDOW, seed).tobytes() + b"$"
    sa = orc.suffix_array(text)
    bwt = orc.bwt_array(text, sa)
    orc.occ_checkpoints(bwt, 128)
    lcp = orc.kasai_lcp(text, sa)
    t1 = orc.tier1_scan(text)
    mask = np.zeros(WINDOW, np.uint8)
    for s, e in t1[:, :2].tolist():
        mask[s:e] = 1
    rows = orc.strict_scan(text, 1, 1000, 0, 3)
    orc.lcp_plateaus(text, sa, lcp, 1, 1000, 3)
    orc.period_scan(text, tier1_mask=mask)
    return len(rows)


def cpu_index_once(n_seed):
    from oracle import oracle as orc

    n, seed = n_seed
    text = gen_contig(n, seed).tobytes() + b"$"
    t0 = time.perf_counter()
    sa = orc.suffix_array(text)
    bwt = orc.bwt_array(text, sa)
    orc.occ_checkpoints(bwt, 128)
    orc.kasai_lcp(text, sa)
    return time.perf_counter() - t0


def _pool(cores):
    import multiprocessing as mp

    return mp.get_context("fork").Pool(cores)


def python_reference_sample():
    """The unmodified Python reference (baseline/_ref/bwt.py, copied there by __graft_entry__.build()
    when /root/reference exists; git-ignored) on one host core: BWTCore + LCP of a 1 Mb contig,
    the strict scan on a 2 kb window.  None when the file did not travel."""
    ref = os.path.join(ROOT, "baseline", "_ref", "bwt.py")
    if not os.path.exists(ref):
        return None
    code = r"""
import sys, time, json, os
sys.path.insert(0, {ref_dir!r}); sys.path.insert(0, {root!r})
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_bench")
import importlib.util
spec = importlib.util.spec_from_file_location("ref_bwt", {ref!r})
m = importlib.util.module_from_spec(spec); sys.modules["ref_bwt"] = m; spec.loader.exec_module(m)
from bench import gen_contig
seq = gen_contig(1_000_000, 21).tobytes().decode()
t0 = time.perf_counter(); core = m.BWTCore(seq + "$"); t1 = time.perf_counter()
t2f = m.Tier2LCPFinder(core); lcp = t2f._compute_lcp_array(); t2 = time.perf_counter()
w = gen_contig(2000, 5).tobytes().decode()
c2 = m.BWTCore(w + "$"); f2 = m.Tier2LCPFinder(c2); f2.min_copies = 3
t3 = time.perf_counter(); r = f2.find_long_unit_repeats_strict("w", min_unit_len=1, max_unit_len=666, max_mismatch=0, min_copies=3); t4 = time.perf_counter()
print(json.dumps({{"index_1mb_s": t1 - t0, "lcp_1mb_s": t2 - t1, "strict_2kb_s": t4 - t3, "strict_rows": len(r), "numba": bool(m.HAVE_NUMBA)}}))
""".format(ref=ref, ref_dir=os.path.dirname(ref), root=ROOT)
    try:
        res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=240,
                             env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1", CUDA_VISIBLE_DEVICES=""))
        d = json.loads(res.stdout.strip().splitlines()[-1])
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {exc}"[:200]}
    return {"kind": "reference", "cores": 1,
            "index_lcp_gbases_per_s": round(1e6 / (d["index_1mb_s"] + d["lcp_1mb_s"]) / 1e9, 8),
            "strict_scan_bases_per_s": round(2000 / d["strict_2kb_s"], 1), "numba": d["numba"],
            "sample": f"BWTCore + _compute_lcp_array on a 1 Mb planted contig ({d['index_1mb_s'] + d['lcp_1mb_s']:.1f} s); "
                      f"find_long_unit_repeats_strict on a 2 kb window ({d['strict_2kb_s']:.1f} s, {d['strict_rows']} records)"}


def cpu_baseline(args, bounded_s: float = 25.0):
    """The genome workload (index + Tier 1 + Tier 2) on this box's host cores, oracle port, one
    20 kb window per core under a process pool; plus the index alone on whole 2 Mb contigs."""
    from oracle import oracle as orc

    orc.build()
    cores = min(os.cpu_count() or 1, args.cpu_cores)
    with _pool(cores) as pool:
        t0 = time.perf_counter()
        rows = pool.map(cpu_window_once, [(900 + i, WINDOW) for i in range(cores)])
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        pool.map(cpu_index_once, [(2_000_000, 21 + i) for i in range(cores)])
        dti = time.perf_counter() - t1
    t2 = time.perf_counter()
    oi = orc.OracleIndex(gen_contig(2_000_000, 21).tobytes() + b"$")
    q = np.frombuffer(b"ACGT", np.uint8)[np.random.default_rng(7).integers(0, 4, (100_000, 10))]
    t3 = time.perf_counter()
    oi.backward_search_batch(np.ascontiguousarray(q), np.full(100_000, 10, np.int32))
    dq = time.perf_counter() - t3
    out = {"value": round(cores * WINDOW / dt / 1e9, 9), "unit": "Gbases/s", "cores": cores, "kind": "port",
           "sample": f"{cores} processes x one {WINDOW}-base window each: index + Tier 1 + strict(u<=1000) + plateaus + "
                     f"period scan, {dt:.1f} s ({sum(rows)} strict rows); the strict scan is ~90 % of it",
           "host_cores": os.cpu_count(),
           "index_only": {"value": round(cores * 2_000_000 / dti / 1e9, 6), "unit": "Gbases/s", "cores": cores,
                          "sample": f"{cores} processes x one whole 2 Mb contig: SA + BWT + Occ + LCP, {dti:.1f} s"},
           "backward_search": {"value": round(100_000 / dq, 1), "unit": "queries/s", "cores": 1,
                               "sample": f"100 000 random 10-mers on a 2 Mb index, {dq:.2f} s (index build {t3 - t2:.1f} s untimed)"}}
    ref = python_reference_sample()
    if ref is not None:
        out["python_reference"] = ref
    return out


def run_reference(args):
    """Reference arm: the reference's CPU algorithm for the same workload on every host core.
    The reference is pure Python (230-350 bases/s/core in its strict scan): a step is a bounded
    sample -- one 20 kb window of the genome workload per core -- through the oracle port
    (kind: port, ~5x faster per core than bwt.py, so the ratio the driver computes is a floor)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc

    orc.build()
    cores = min(os.cpu_count() or 1, args.cpu_cores)
    times = []
    WINDOW = args.ref_window
    with _pool(cores) as pool:
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(cpu_window_once, [(900 + it * cores + i, WINDOW) for i in range(cores)])
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = cores * WINDOW * len(times) / total / 1e9
    world = int(os.environ.get("WORLD_SIZE", "1"))
    base = {"value": round(value, 9), "unit": "Gbases/s", "cores": cores, "kind": "port",
            "sample": f"per step: {cores} processes x one {WINDOW}-base window (index + Tier 1 + Tier 2, oracle port)"}
    ref = None if args.no_python_ref else python_reference_sample()
    if ref is not None:
        base["python_reference"] = ref
    out = {
        "impl": "reference", "metric": "SA+BWT+LCP Gbases/s", "value": round(value, 9), "unit": "Gbases/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(total / len(times) * 1e3, 3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "bases": sum(HG38), "contigs": len(HG38),
                   "sample_per_step": f"{cores} windows x {WINDOW} bases (bounded sample of the workload)"},
        "cpu_baseline": base,
        "e2e": {"value": round(value, 9), "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every contig of the genome (smoke runs)")
    ap.add_argument("--fm-log2-queries", type=int, default=26, help="random 10-mers of the FM section (2^k)")
    ap.add_argument("--cpu-cores", type=int, default=64, help="upper bound on host processes of the CPU legs")
    ap.add_argument("--lanes", type=int, default=4, help="contigs in flight per GPU (GenomeScanner lanes)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--ref-window", type=int, default=REF_WINDOW,
                    help="bases per window (one per core and step) of the reference arm")
    ap.add_argument("--no-python-ref", action="store_true",
                    help="skip the 1 Mb sample of the unmodified Python reference (baseline/_ref/bwt.py)")
    ap.add_argument("--skip-extras", action="store_true",
                    help="genome steps only (no per-kernel profile, FM search or scan sections); used under ncu")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
