#!/usr/bin/env python3
"""Benchmark of the index-and-scan hot path (BASELINE.json metric:
SA+BWT+LCP Gbases/s and FM backward-search queries/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one full index build (byte histogram/C array, suffix array, BWT + Occ
checkpoints, LCP) over one synthetic chr21-sized contig (46 709 983 bases + '$',
SURVEY.md Appendix B generator, BASELINE.json configs[2]).  `value` times the
step with the text already resident in HBM; `e2e` times the same build through
host buffers (pinned H2D of the text, D2H of SA + BWT + LCP).  With N > 1 every
rank builds its own contig (weak scaling, no data-path collective); the only
collective is the gather of per-rank record counts.  Extra keys report the FM
search rates, the Tier 1+2 scan rate, the per-kernel roofline of the dominant
kernel and a CPU baseline (the oracle port) timed on this box's host cores.
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
BYTES_PER_BASE_COMPULSORY = 20.2  # SA 5 + BWT/C/Occ 6.19 + LCP 9


def gen_contig(n, seed, sub_rate=0.03):
    """SURVEY.md Appendix B."""
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


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/r1_ncu_traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")) as f:
            return json.load(f)[kernel]["traffic_bytes_per_launch"]
    except Exception:
        return None


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
    """Pre-allocated buffers + the C-ABI calls of one index build."""

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
        """One fused call: C array + SA (+ISA) + BWT + Occ + LCP (bwtk_index_build)."""
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


def run_ours(args):
    import torch
    import torch.distributed as dist

    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200 import _lib as lib

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
    n = args.n + 1
    text_h = np.concatenate([gen_contig(args.n, 21 + rank), np.frombuffer(b"$", np.uint8)])
    pinned = torch.from_numpy(text_h).pin_memory()
    d_text = pinned.to(dev)
    step = IndexStep(torch, L, lib, n, dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step.run(d_text)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.bwtk_launch_count()
    times = []
    for _ in range(args.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step.run(d_text)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    launches = L.bwtk_launch_count() - launches0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = float(sum(times))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    sa_stats = step.stats.copy()

    # ---- e2e: host text in, SA + BWT + LCP out --------------------------------
    # Through the package's public streaming API (bwt_algorithm_b200.streaming.IndexPipeline ->
    # bwtk_index_build): every step uploads the contig from pinned host memory and downloads
    # SA, BWT and LCP into pinned host buffers.  Two contigs are in flight (the multi-contig
    # pipeline's steady state), so transfers of one overlap the kernels of the next; the serial
    # one-contig-at-a-time figure is reported next to it.
    from bwt_algorithm_b200.streaming import IndexPipeline
    del step.ws
    torch.cuda.empty_cache()
    pipe = IndexPipeline(n, occ_rate=128, device=dev, slots=2, want_lcp=True)
    h2d_b, d2h_b = pipe.bytes_per_contig(n)
    for _ in range(max(args.warmup, 2)):
        r = pipe.result(pipe.submit(pinned))
    assert int(r["sa"][0]) == n - 1 and r["lcp"].shape[0] == n     # '$' sorts first
    # serial: one contig at a time
    serial = []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = pipe.result(pipe.submit(pinned))
        serial.append((time.perf_counter() - t0) * 1e3)
    e2e_serial_ms = float(sum(serial))
    # pipelined: K contigs back to back, timed with events from the first upload to the last download
    pipe.drain()
    torch.cuda.synchronize()
    barrier()
    first = pipe.slots[pipe.tickets % len(pipe.slots)]
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record(first.stream)
    prev = None
    for _ in range(args.steps):
        tk = pipe.submit(pinned)
        if prev is not None:
            r = pipe.result(prev)
            _ = int(r["sa"][0])       # the host consumes the previous contig while this one is in flight
        prev = tk
    r = pipe.result(prev)
    pipe.drain()
    e2e_ms = max(e0.elapsed_time(sl.done) for sl in pipe.slots)
    if world > 1:
        t = torch.tensor([e2e_ms, e2e_serial_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms, e2e_serial_ms = float(t[0].item()), float(t[1].item())
    del pipe
    torch.cuda.empty_cache()
    step.ws = torch.empty(int(L.bwtk_index_workspace_bytes(n)), dtype=torch.uint8, device=dev)

    extras = {}
    roof = None
    if world > 1 and not args.skip_extras:
        # ---- one contig, N GPUs: rank 0's FM index replicated over NVLink, the 10-mer batch split ----
        # (SURVEY 8e / BASELINE "motif batches per GPU"): broadcast of BWT + Occ, then every rank answers
        # its block of the batch and one all_gather returns all (sp, ep); device-timed, max over ranks.
        from types import SimpleNamespace

        from bwt_algorithm_b200 import sharding
        from bwt_algorithm_b200.device_index import FMReplica

        built = SimpleNamespace(n=n, occ_rate=128, bwt=step.bwt, occ=step.occ, totals=step.totals,
                                row_of_code=step.row) if rank == 0 else None
        for _ in range(2):          # the first broadcast also sets up NCCL's broadcast channels
            barrier()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record()
            fm = sharding.broadcast_fm(built, src=0, device=dev)
            b1.record()
            torch.cuda.synchronize()
            bcast_ms = b0.elapsed_time(b1)
        replica = FMReplica(device=dev, **fm)
        nq_all = (1 << 22) * world
        q_host = np.frombuffer(b"ACGT", np.uint8)[np.random.default_rng(7).integers(0, 4, (nq_all, 10), dtype=np.uint8)]
        q_pats = torch.from_numpy(q_host).to(dev)
        q_lens = torch.full((nq_all,), 10, dtype=torch.int32, device=dev)
        sharding.sharded_search(replica.search_block, q_pats, q_lens)       # warm-up
        qs = []
        for _ in range(3):
            flush.fill_(1)
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            q_sp, q_ep = sharding.sharded_search(replica.search_block, q_pats, q_lens)
            a1.record()
            torch.cuda.synchronize()
            qs.append(a0.elapsed_time(a1))
        t = torch.tensor([statistics.median(qs), bcast_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q_ms, bcast_ms = float(t[0].item()), float(t[1].item())
        found = int((q_sp >= 0).sum().item())
        extras["fm_sharded_10mers"] = {
            "queries": nq_all, "ms": round(q_ms, 4), "queries_per_s": round(nq_all / (q_ms * 1e-3), 1),
            "found": found, "index_broadcast_ms": round(bcast_ms, 4),
            "index_broadcast_bytes": int(fm["bwt"].numel() + fm["occ"].numel() * 4),
            "note": "rank 0's chr21-sized index broadcast over NCCL, batch split in equal blocks, "
                    "all_gather of (sp, ep) inside the timed region"}
        del replica, fm, q_pats, q_lens, q_sp, q_ep
    if rank == 0 and args.skip_extras:
        peak, peak_src = measured_peak()
    if rank == 0 and not args.skip_extras:
        # ---- per-kernel profile (one extra step, events around every launch) ---
        L.bwtk_profile_enable(1)
        step.run(d_text)
        prof = profile_report(L)
        L.bwtk_profile_enable(0)
        peak, peak_src = measured_peak()
        tot_ms = sum(r["ms"] for r in prof) or 1.0
        for r in prof:
            r["share"] = round(r["ms"] / tot_ms, 4)
            r["gbs"] = round(r["algo_bytes"] / (r["ms"] * 1e-3) / 1e9, 1) if r["ms"] > 0 else None
            r["ms"] = round(r["ms"], 4)
        top = max(prof, key=lambda r: r["ms"])
        achieved = top["algo_bytes"] / (top["ms"] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": top["kernel"], "achieved": round(achieved, 1), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": ncu_traffic(top["kernel"]),
                "peak_source": peak_src,
                "launches_per_step": top["launches"], "avg_launch_ms": round(top["ms"] / top["launches"], 4),
                "algo_bytes_per_launch": top["algo_bytes"] // max(top["launches"], 1)}
        extras["kernels"] = prof

        # ---- FM backward search rates ------------------------------------------
        d_C = torch.zeros(256, dtype=torch.int64, device=dev)
        d_tot = torch.from_numpy(step.totals.copy()).to(dev)
        cum = np.zeros(256, np.int64)
        c = 0
        for b in range(256):
            cum[b] = c
            c += int(step.totals[b])
        d_C.copy_(torch.from_numpy(cum))
        d_row = torch.from_numpy(step.row).to(dev)
        kmax = 10
        total_m = (4 ** (kmax + 1) - 4) // 3
        sp = torch.empty(total_m, dtype=torch.int32, device=dev)
        ep = torch.empty(total_m, dtype=torch.int32, device=dev)

        def timed(fn, reps=5):
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

        st = lib.stream_ptr()
        ms = timed(lambda: lib.check(L.bwtk_bsearch_motif_sweep(
            step.bwt.data_ptr(), step.occ.data_ptr(), step.ncp, 128, d_C.data_ptr(), d_tot.data_ptr(),
            d_row.data_ptr(), n, kmax, sp.data_ptr(), ep.data_ptr(), st), "sweep"))
        extras["fm_motif_sweep"] = {"queries": total_m, "ms": round(ms, 4), "queries_per_s": round(total_m / (ms * 1e-3), 1),
                                    "algo_bytes_per_query": 144}
        nq = 1 << 22
        pats = torch.from_numpy(np.frombuffer(b"ACGT", np.uint8)[np.random.default_rng(7).integers(0, 4, (nq, 10))]).to(dev)
        lens = torch.full((nq,), 10, dtype=torch.int32, device=dev)
        sp2 = torch.empty(nq, dtype=torch.int32, device=dev)
        ep2 = torch.empty(nq, dtype=torch.int32, device=dev)
        ms = timed(lambda: lib.check(L.bwtk_bsearch_batch(
            step.bwt.data_ptr(), step.occ.data_ptr(), step.ncp, 128, d_C.data_ptr(), d_tot.data_ptr(),
            d_row.data_ptr(), n, pats.data_ptr(), 10, lens.data_ptr(), nq, sp2.data_ptr(), ep2.data_ptr(), st), "batch"))
        extras["fm_random_10mers"] = {"queries": nq, "ms": round(ms, 4), "queries_per_s": round(nq / (ms * 1e-3), 1),
                                      "algo_bytes_per_query": 1232}

        # ---- Tier 1 + Tier 2 scans on the same contig ------------------------------
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
            t0 = time.perf_counter()
            rows = fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            scan[name] = {"records": int(len(rows)), "ms": round(dt * 1e3, 2),
                          "gbases_per_s": round(args.n / dt / 1e9, 4), "timing": "wall clock, second call, rows on host"}
        extras["scan"] = scan

    counts = [0]
    if world > 1:
        # the only collective of the path: gather of per-rank record counts
        mine = torch.tensor([int(sa_stats[0])], dtype=torch.int64, device=dev)
        got = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(got, mine)
        counts = [int(x.item()) for x in got]

    if rank == 0:
        value = world * args.n * args.steps / (total_ms * 1e-3) / 1e9
        # one or two contigs in flight per GPU, whichever moved more bases: with many GPUs on one host the
        # host's memory system (not PCIe) limits the downloads and the second stream only adds contention
        e2e_best_ms = min(e2e_ms, e2e_serial_ms)
        e2e_val = world * args.n * args.steps / (e2e_best_ms * 1e-3) / 1e9
        peak, peak_src = measured_peak()
        out = {
            "metric": "SA+BWT+LCP Gbases/s", "value": round(value, 4), "unit": "Gbases/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32",
            "data": "synthetic",
            "config": {"workload": "chr21-sized planted contig per GPU (BASELINE configs[2])", "bases_per_gpu": args.n,
                       "generator": "SURVEY Appendix B gen_contig(seed=21+rank)", "occ_rate": 128,
                       "l2": "256 MB write between timed steps (flush)", "parallelism": f"contig-per-gpu x{world}"},
            "e2e": {"value": round(e2e_val, 4), "unit": "Gbases/s", "h2d_bytes_per_step": h2d_b,
                    "d2h_bytes_per_step": d2h_b, "ms_per_step": round(e2e_best_ms / args.steps, 4),
                    "in_flight": 2 if e2e_ms <= e2e_serial_ms else 1,
                    "api": "bwt_algorithm_b200.streaming.IndexPipeline -> bwtk_index_build",
                    "pipelined_ms_per_step": round(e2e_ms / args.steps, 4),
                    "pipelined_value": round(world * args.n * args.steps / (e2e_ms * 1e-3) / 1e9, 4),
                    "serial_ms_per_step": round(e2e_serial_ms / args.steps, 4),
                    "serial_value": round(world * args.n * args.steps / (e2e_serial_ms * 1e-3) / 1e9, 4),
                    "l2": "no flush between pipelined steps: each contig touches ~2 GB >> 126 MB L2"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "compulsory_fraction": round(BYTES_PER_BASE_COMPULSORY * value / world / peak, 5),
            "sa_stats": {"rounds": int(sa_stats[0]), "bits_per_symbol": int(sa_stats[1]),
                         "symbols_per_key": int(sa_stats[2]), "active_after_round0": int(sa_stats[3]),
                         "sum_active": int(sa_stats[4]), "radix_passes": int(sa_stats[5])},
        }
        out.update(extras)
        if not args.no_cpu and world == 1:
            out["cpu_baseline"] = cpu_baseline_port(args.cpu_sample)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- CPU
def cpu_index_once(text: bytes):
    from oracle import oracle as orc

    sa = orc.suffix_array(text)
    bwt = orc.bwt_array(text, sa)
    orc.occ_checkpoints(bwt, 128)
    orc.kasai_lcp(text, sa)


def cpu_baseline_port(sample: int):
    """Oracle port (oracle/oracle.c: prefix doubling + Kasai, the reference's
    algorithm) on one host core over a bounded sample of the same contig."""
    from oracle import oracle as orc

    orc.build()
    text = gen_contig(sample, 21).tobytes() + b"$"
    t0 = time.perf_counter()
    cpu_index_once(text)
    dt = time.perf_counter() - t0
    return {"value": round(sample / dt / 1e9, 6), "unit": "Gbases/s", "cores": 1, "kind": "port",
            "sample": f"first-principles rebuild of a {sample}-base planted contig (seed 21), SA+BWT+Occ+LCP, {dt:.2f} s",
            "host_cores": os.cpu_count()}


def run_reference(args):
    """Reference arm: the reference is pure Python and cannot travel to the GPU
    box, so its CPU algorithm is timed through the oracle port (kind: port) with
    every host thread busy on its own bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as orc

    orc.build()
    cores = min(os.cpu_count() or 1, 32)
    sample = args.cpu_sample // 2
    texts = [gen_contig(sample, 21 + i).tobytes() + b"$" for i in range(cores)]
    times = []
    with ThreadPoolExecutor(cores) as pool:
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            list(pool.map(cpu_index_once, texts))
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = cores * sample * len(times) / total / 1e9
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = {
        "impl": "reference", "metric": "SA+BWT+LCP Gbases/s", "value": round(value, 6), "unit": "Gbases/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(total / len(times) * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
        "config": {"workload": "chr21-sized planted contig per GPU (BASELINE configs[2])", "bases_per_gpu": args.n,
                   "sample_per_step": f"{cores} contigs x {sample} bases (bounded sample of the workload)"},
        "cpu_baseline": {"value": round(value, 6), "unit": "Gbases/s", "cores": cores, "kind": "port",
                         "sample": f"{cores} threads x {sample}-base planted contigs per step"},
        "e2e": {"value": round(value, 6), "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=CONTIG_N, help="bases per GPU (default: chr21-sized)")
    ap.add_argument("--cpu-sample", type=int, default=10_000_000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--skip-extras", action="store_true",
                    help="index build only (no per-kernel profile, FM search or scan sections); used under ncu")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
