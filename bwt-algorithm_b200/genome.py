"""Whole-genome index-and-scan on one GPU: the row-level form of the per-contig worker.

The reference handles a genome one contig at a time (``_process_chromosome_worker``,
bwt.py:3040-3141; dispatch and merge in ``find_tandem_repeats_parallel``, bwt.py:3863-3899):
build the contig's index, run the detectors, hand the records back.  ``GenomeScanner`` is
that unit for a rank that owns one B200 and a list of contigs:

    host text (pinned) --H2D--> text slot --bwtk_index_build--> SA, BWT, Occ, LCP   (stay in HBM)
        --bwtk_tier1_scan (a12) --bwtk_strict_scan (a13) --bwtk_lcp_plateaus (a15)
        --bwtk_period_scan (a16, masked by the Tier 1 calls)--> int32[R, 8] rows in a device arena
        --D2H--> pinned host arena (optional)

Three streams: the next contig's text is uploaded while the current one is processed, and the
rows of the previous contig travel to the host while the next index is built.  The index arrays
never leave the device (the detectors read them there); only text comes in (1 B/base) and rows
go out.  Nothing is allocated inside ``scan``: all buffers are sized for the largest contig when
the scanner is created.  ``find_short_imperfect_repeats`` (a14) is not part of this path: it
returns [] above 1 Mb by design (bwt.py:2048) and is served by ``seeded.py`` below that.

There is no CPU fallback: without libbwtk.so or a CUDA device the constructor raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib, detect

REC_W = _lib.REC_W
KINDS = ("tier1", "strict", "plateaus", "period")


@dataclass
class ContigRows:
    """Rows of one contig: ``span[kind] = (first_row, count)`` into the scanner's arenas."""
    contig: int
    n: int                                    # symbols incl. '$'
    span: Dict[str, tuple] = field(default_factory=dict)
    threshold: int = -1                       # plateau threshold (-1: no scan)
    period_iterations: int = 0
    sa_stats: Optional[np.ndarray] = None

    def count(self) -> int:
        return sum(c for _, c in self.span.values())


class GenomeScanner:
    """Index + Tier 1 + Tier 2 rows for a list of contigs on one device (see module docstring)."""

    def __init__(self, max_n: int, arena_rows: int, device=None, occ_rate: int = 128, host_arena: bool = True,
                 kinds: Sequence[str] = KINDS, min_copies: int = 3, max_motif_len: int = 9,
                 min_array_len: int = 6, min_entropy: float = 1.0, max_period: int = 1000,
                 max_unit_len: int = 120, allow_mismatches: bool = True):
        torch = _lib.require_cuda()
        self.torch = torch
        self.L = L = _lib.lib()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.max_n = max_n = int(max_n)
        self.occ_rate = int(occ_rate)
        self.kinds = tuple(kinds)
        self.min_copies, self.max_motif_len = int(min_copies), int(max_motif_len)
        self.min_array_len, self.min_entropy = int(min_array_len), float(min_entropy)
        self.max_period, self.max_unit_len = int(max_period), int(max_unit_len)
        self.allow_mismatches = bool(allow_mismatches)
        self.arena_rows = int(arena_rows)
        dev = self.device
        with torch.cuda.device(dev):
            self.text_slots = [torch.empty(max_n, dtype=torch.uint8, device=dev) for _ in range(2)]
            self.sa = torch.empty(max_n, dtype=torch.int32, device=dev)
            self.bwt = torch.empty(max_n, dtype=torch.uint8, device=dev)
            self.lcp = torch.empty(max_n, dtype=torch.int32, device=dev)
            self.occ_rows = 8
            self.occ = torch.zeros(self.occ_rows * (max_n // self.occ_rate + 2), dtype=torch.int32, device=dev)
            self.seen = torch.zeros(max_n, dtype=torch.uint8, device=dev)
            # the index build and the detectors run one after the other on one stream: one workspace
            wsb = max(int(L.bwtk_index_workspace_bytes(max_n)), int(L.bwtk_tier1_workspace_bytes(max_n)),
                      int(L.bwtk_strict_workspace_bytes(max_n, 1000)), int(L.bwtk_plateau_workspace_bytes(max_n)))
            self.ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            self.arena = torch.empty((self.arena_rows, REC_W), dtype=torch.int32, device=dev)
            self.plogp_dim = 202
            self.plogp = torch.from_numpy(detect.plogp_table(self.plogp_dim).reshape(-1)).to(dev)
            self.main = torch.cuda.Stream(device=dev)
            self.up = torch.cuda.Stream(device=dev)
            self.dl = torch.cuda.Stream(device=dev)
            self.up_done = [torch.cuda.Event() for _ in range(2)]
            self.slot_free = [torch.cuda.Event() for _ in range(2)]
            self.rows_ready = torch.cuda.Event()
            self.t_start = torch.cuda.Event(enable_timing=True)
            self.t_end = torch.cuda.Event(enable_timing=True)
        self.host = torch.empty((self.arena_rows, REC_W), dtype=torch.int32, pin_memory=True) if host_arena else None
        self.totals = np.zeros(256, np.int64)
        self.row = np.full(256, -1, np.int32)
        self.stats = np.zeros(8, np.int64)
        self.used = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.profile_phases = False     # record events around the index build and every detector
        self._phase_events = []
        self.phase_ms: Dict[str, float] = {}

    def _mark(self, name: str) -> None:
        if self.profile_phases:
            ev = self.torch.cuda.Event(enable_timing=True)
            ev.record(self.main)
            self._phase_events.append((name, ev))

    # ------------------------------------------------------------------ one contig
    def _rows_call(self, what: str, call) -> int:
        """Runs one detector into the arena at ``self.used``; returns the row count."""
        cnt = C.c_int64(0)
        cap = self.arena_rows - self.used
        rc = call(self.arena.data_ptr() + self.used * REC_W * 4, cap, C.addressof(cnt))
        if rc == _lib.E_OVERFLOW:
            raise _lib.BwtkError(f"{what}: {cnt.value} rows do not fit the arena ({cap} rows left of "
                                 f"{self.arena_rows}); create the scanner with a larger arena_rows")
        _lib.check(rc, what)
        return int(cnt.value)

    def _process(self, d_text, n: int, res: ContigRows) -> None:
        """Index build + the four detectors for one device-resident contig, on ``self.main``."""
        L, st = self.L, self.main.cuda_stream
        self._mark("start")
        for _attempt in range(2):
            rc = L.bwtk_index_build(d_text.data_ptr(), n, self.occ_rate, self.sa.data_ptr(), None,
                                    self.bwt.data_ptr(), self.occ.data_ptr(), self.occ_rows, self.lcp.data_ptr(),
                                    self.totals.ctypes.data, self.row.ctypes.data, self.stats.ctypes.data,
                                    self.ws.data_ptr(), self.ws.numel(), st)
            if rc == _lib.E_OVERFLOW and self.stats[7] > self.occ_rows:
                # more than 8 distinct byte values (IUPAC codes, soft-masked text): a taller Occ matrix
                self.occ_rows = int(self.stats[7])
                self.occ = self.torch.zeros(self.occ_rows * (self.max_n // self.occ_rate + 2), dtype=self.torch.int32,
                                            device=self.device)
                continue
            break
        _lib.check(rc, "index_build")
        self._mark("index")
        res.sa_stats = self.stats.copy()
        wsp, wsb = self.ws.data_ptr(), self.ws.numel()
        n_seq = n - 1                      # callers append exactly one '$' (bwt.py:3053)
        want_mask = "period" in self.kinds and "tier1" in self.kinds
        if "tier1" in self.kinds:
            first = self.used
            c = self._rows_call("tier1_scan", lambda rec, cap, cnt: L.bwtk_tier1_scan(
                d_text.data_ptr(), n, self.max_motif_len, self.min_copies, self.min_array_len, self.min_entropy,
                rec, cap, cnt, self.seen.data_ptr() if want_mask else None, wsp, wsb, st))
            res.span["tier1"] = (first, c)
            self.used += c
            self._mark("tier1")
        if "strict" in self.kinds:
            first = self.used
            unit_cap = max(self.max_unit_len, min(n_seq // self.min_copies, 1000))   # bwt.py:3088-3096
            c = self._rows_call("strict_scan", lambda rec, cap, cnt: L.bwtk_strict_scan(
                d_text.data_ptr(), n, 1, unit_cap, 0, self.min_copies, rec, cap, cnt, wsp, wsb, st))
            res.span["strict"] = (first, c)
            self.used += c
            self._mark("strict")
        if "plateaus" in self.kinds:
            first = self.used
            thr = C.c_int64(-1)
            c = self._rows_call("lcp_plateaus", lambda rec, cap, cnt: L.bwtk_lcp_plateaus(
                d_text.data_ptr(), n, self.sa.data_ptr(), self.lcp.data_ptr(), n, 1, self.max_period,
                self.min_copies, rec, cap, cnt, C.addressof(thr), wsp, wsb, st))
            res.span["plateaus"] = (first, c)
            res.threshold = int(thr.value)
            self.used += c
            self._mark("plateaus")
        if "period" in self.kinds:
            first = self.used
            it = C.c_int64(0)
            c = self._rows_call("period_scan", lambda rec, cap, cnt: L.bwtk_period_scan(
                d_text.data_ptr(), n, 1, self.max_period, int(self.allow_mismatches), self.min_copies,
                self.min_array_len, self.min_entropy, self.seen.data_ptr() if want_mask else None,
                self.plogp.data_ptr(), self.plogp_dim, rec, cap, cnt, C.addressof(it), st))
            res.span["period"] = (first, c)
            res.period_iterations = int(it.value)
            self.used += c
            self._mark("period")

    # ------------------------------------------------------------------ a list of contigs
    def scan(self, contigs: Sequence, ids: Optional[Sequence[int]] = None, download: bool = True) -> List[ContigRows]:
        """Processes the contigs in order.  Each is a uint8 tensor with its '$': a CPU tensor (pinned for
        full-speed uploads) or a tensor already on this device.  Rows land in ``self.arena`` (device) and,
        with ``download``, in ``self.host`` (pinned), contig after contig, kind after kind, from row 0.
        Returns after everything (including the downloads) has completed; ``self.elapsed_ms`` is the
        device time from the first upload to the last download."""
        torch = self.torch
        if download and self.host is None:
            raise _lib.BwtkError("scanner was created without a host arena")
        ids = list(range(len(contigs))) if ids is None else list(ids)
        out: List[ContigRows] = []
        self.used = 0
        self.h2d_bytes = self.d2h_bytes = 0
        self._phase_events = []
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream()
            self.main.wait_stream(cur)
            self.up.wait_stream(cur)
            self.dl.wait_stream(cur)
            self.t_start.record(self.main)
            self.up.wait_event(self.t_start)

            def stage(i: int):
                t = contigs[i]
                n = int(t.numel())
                if n < 1 or n > self.max_n:
                    raise _lib.BwtkError(f"contig of {n} symbols does not fit the scanner (max_n={self.max_n})")
                if t.is_cuda:
                    return t
                slot = self.text_slots[i % 2]
                with torch.cuda.stream(self.up):
                    self.up.wait_event(self.slot_free[i % 2])
                    slot[:n].copy_(t, non_blocking=True)
                    self.up_done[i % 2].record(self.up)
                self.h2d_bytes += n
                return slot[:n]

            for k in range(2):
                self.slot_free[k].record(self.main)
            nxt = stage(0) if contigs else None
            for i in range(len(contigs)):
                d_text = nxt
                n = int(d_text.numel())
                if not contigs[i].is_cuda:
                    self.main.wait_event(self.up_done[i % 2])
                if i + 1 < len(contigs):
                    nxt = stage(i + 1)       # its slot was released when contig i-1 finished
                res = ContigRows(contig=ids[i], n=n)
                first = self.used
                with torch.cuda.stream(self.main):
                    self._process(d_text, n, res)
                    self.slot_free[i % 2].record(self.main)
                    self.rows_ready.record(self.main)
                out.append(res)
                if download and self.used > first:
                    with torch.cuda.stream(self.dl):
                        self.dl.wait_event(self.rows_ready)
                        self.host[first:self.used].copy_(self.arena[first:self.used], non_blocking=True)
                    self.d2h_bytes += (self.used - first) * REC_W * 4
            self.main.wait_stream(self.dl)
            self.main.wait_stream(self.up)
            self.t_end.record(self.main)
            self.t_end.synchronize()
            cur.wait_stream(self.main)
            self.elapsed_ms = self.t_start.elapsed_time(self.t_end)
            if self.profile_phases:
                self.phase_ms = {}
                for (_, a), (name, b) in zip(self._phase_events, self._phase_events[1:]):
                    if name != "start":
                        self.phase_ms[name] = self.phase_ms.get(name, 0.0) + a.elapsed_time(b)
        return out

    # ------------------------------------------------------------------ views
    def device_rows(self, res: ContigRows, kind: str):
        a, c = res.span[kind]
        return self.arena[a:a + c]

    def host_rows(self, res: ContigRows, kind: str) -> np.ndarray:
        a, c = res.span[kind]
        return self.host[a:a + c].numpy()
