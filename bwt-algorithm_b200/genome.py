"""Whole-genome index-and-scan on one GPU: the row-level form of the per-contig worker.

The reference handles a genome one contig at a time (``_process_chromosome_worker``,
bwt.py:3040-3141; dispatch and merge in ``find_tandem_repeats_parallel``, bwt.py:3863-3899):
build the contig's index, run the detectors, hand the records back.  ``GenomeScanner`` is
that unit for a rank that owns one B200 and a list of contigs:

    host text (pinned) --H2D--> text slot --bwtk_index_build--> SA, BWT, Occ, LCP   (stay in HBM)
        --bwtk_tier1_scan (a12) --bwtk_strict_scan (a13) --bwtk_lcp_plateaus (a15)
        --bwtk_period_scan (a16, masked by the Tier 1 calls)--> int32[R, 8] rows in the lane's scratch
        --D2D--> the scanner's row arena --D2H--> pinned host arena (optional)

``lanes`` contigs are in flight at once (default 4; 70 bytes of device memory per symbol of the longest contig
and lane), each on its own host thread, CUDA stream and set of
device buffers: the suffix-sort rounds, the nine Tier 1 passes and the period scan are chains of small
launches with host read-backs in between, and a second contig's kernels fill the SMs they leave idle
(the C-ABI calls release the GIL and only ever synchronise the stream they are given).  The lanes take
contigs from one queue in the order given (pass the longest first); a finished contig reserves its rows
in the shared arena under a lock, so the arena is dense but its order is the completion order --
``ContigRows.span`` says where each kind of each contig lives.  The index arrays never leave the device
(the detectors read them there); only text comes in (1 B/base) and rows go out.  Nothing is allocated
inside ``scan``.  ``find_short_imperfect_repeats`` (a14) is not part of this path: it returns [] above
1 Mb by design (bwt.py:2048) and is served by ``seeded.py`` below that.

There is no CPU fallback: without libbwtk.so or a CUDA device the constructor raises.
"""
from __future__ import annotations

import ctypes as C
import threading
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


class _Lane:
    """Device buffers, stream and host-side outputs of one contig in flight."""

    def __init__(self, torch, L, max_n: int, occ_rate: int, scratch_rows: int, dev):
        self.text = [torch.empty(max_n, dtype=torch.uint8, device=dev) for _ in range(2)]   # current + prefetched
        self.sa = torch.empty(max_n, dtype=torch.int32, device=dev)
        self.bwt = torch.empty(max_n, dtype=torch.uint8, device=dev)
        self.lcp = torch.empty(max_n, dtype=torch.int32, device=dev)
        self.occ_rows = 8
        self.occ = torch.zeros(self.occ_rows * (max_n // occ_rate + 2), dtype=torch.int32, device=dev)
        self.seen = torch.zeros(max_n, dtype=torch.uint8, device=dev)
        self.hint = torch.zeros(int(L.bwtk_repeat_hint_bytes(max_n)) // 4 + 4, dtype=torch.int32, device=dev)
        # the index build and the detectors run one after the other on one stream: one workspace
        wsb = max(int(L.bwtk_index_workspace_bytes(max_n)), int(L.bwtk_tier1_workspace_bytes(max_n)),
                  int(L.bwtk_strict_workspace_bytes(max_n, 1000)), int(L.bwtk_plateau_workspace_bytes(max_n)))
        self.ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        self.scratch = torch.empty((scratch_rows, REC_W), dtype=torch.int32, device=dev)
        self.scratch_rows = scratch_rows
        self.stream = torch.cuda.Stream(device=dev)
        self.dl = torch.cuda.Stream(device=dev)      # row downloads: the lane goes on computing meanwhile
        self.up = torch.cuda.Stream(device=dev)      # text uploads: the lane's next contig arrives during this one
        self.rows_ready = torch.cuda.Event()
        self.up_done = [torch.cuda.Event() for _ in range(2)]
        self.totals = np.zeros(256, np.int64)
        self.row = np.full(256, -1, np.int32)
        self.stats = np.zeros(8, np.int64)
        self.events = []          # (name, event) marks of the phase profile
        self.h2d = 0
        self.d2h = 0


class GenomeScanner:
    """Index + Tier 1 + Tier 2 rows for a list of contigs on one device (see module docstring)."""

    def __init__(self, max_n: int, arena_rows: int, device=None, occ_rate: int = 128, host_arena: bool = True,
                 kinds: Sequence[str] = KINDS, min_copies: int = 3, max_motif_len: int = 9,
                 min_array_len: int = 6, min_entropy: float = 1.0, max_period: int = 1000,
                 max_unit_len: int = 120, allow_mismatches: bool = True, lanes: int = 4):
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
        # rows of ONE contig: the strict scan's run rows (~n/19 on planted sequence) + the period scan's raw rows
        scratch_rows = min(self.arena_rows, max_n // 8 + 262144)
        dev = self.device
        with torch.cuda.device(dev):
            self.lanes = [_Lane(torch, L, max_n, self.occ_rate, scratch_rows, dev) for _ in range(max(1, int(lanes)))]
            self.arena = torch.empty((self.arena_rows, REC_W), dtype=torch.int32, device=dev)
            self.plogp_dim = 202
            self.plogp = torch.from_numpy(detect.plogp_table(self.plogp_dim).reshape(-1)).to(dev)
            self.t_start = torch.cuda.Event(enable_timing=True)
            self.t_end = torch.cuda.Event(enable_timing=True)
        self.host = torch.empty((self.arena_rows, REC_W), dtype=torch.int32, pin_memory=True) if host_arena else None
        self.used = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.elapsed_ms = 0.0
        self.profile_phases = False     # record events around the index build and every detector
        self.phase_ms: Dict[str, float] = {}
        self._lock = threading.Lock()

    # the index arrays of the contig lane 0 processed last (a one-contig scan always runs on lane 0)
    @property
    def sa(self):
        return self.lanes[0].sa

    @property
    def bwt(self):
        return self.lanes[0].bwt

    @property
    def lcp(self):
        return self.lanes[0].lcp

    @property
    def occ(self):
        return self.lanes[0].occ

    @property
    def totals(self):
        return self.lanes[0].totals

    @property
    def row(self):
        return self.lanes[0].row

    def _mark(self, lane: _Lane, name: str) -> None:
        if self.profile_phases:
            ev = self.torch.cuda.Event(enable_timing=True)
            ev.record(lane.stream)
            lane.events.append((name, ev))

    # ------------------------------------------------------------------ one contig
    def _rows_call(self, lane: _Lane, used: int, what: str, call) -> int:
        """Runs one detector into the lane's scratch at row ``used``; returns the row count."""
        cnt = C.c_int64(0)
        cap = lane.scratch_rows - used
        rc = call(lane.scratch.data_ptr() + used * REC_W * 4, cap, C.addressof(cnt))
        if rc == _lib.E_OVERFLOW:
            raise _lib.BwtkError(f"{what}: {cnt.value} rows do not fit the arena ({cap} rows left for this contig of "
                                 f"{lane.scratch_rows}); create the scanner with a larger arena_rows")
        _lib.check(rc, what)
        return int(cnt.value)

    def _process(self, lane: _Lane, d_text, n: int, res: ContigRows) -> Dict[str, tuple]:
        """Index build + the four detectors for one device-resident contig, on the lane's stream.
        Returns the spans inside the lane's scratch."""
        L, st = self.L, lane.stream.cuda_stream
        self._mark(lane, "start")
        for _attempt in range(2):
            rc = L.bwtk_index_build(d_text.data_ptr(), n, self.occ_rate, lane.sa.data_ptr(), None,
                                    lane.bwt.data_ptr(), lane.occ.data_ptr(), lane.occ_rows, lane.lcp.data_ptr(),
                                    lane.totals.ctypes.data, lane.row.ctypes.data, lane.stats.ctypes.data,
                                    lane.ws.data_ptr(), lane.ws.numel(), st)
            if rc == _lib.E_OVERFLOW and lane.stats[7] > lane.occ_rows:
                # more than 8 distinct byte values (IUPAC codes, soft-masked text): a taller Occ matrix
                lane.occ_rows = int(lane.stats[7])
                lane.occ = self.torch.zeros(lane.occ_rows * (self.max_n // self.occ_rate + 2),
                                            dtype=self.torch.int32, device=self.device)
                continue
            break
        _lib.check(rc, "index_build")
        self._mark(lane, "index")
        res.sa_stats = lane.stats.copy()
        wsp, wsb = lane.ws.data_ptr(), lane.ws.numel()
        n_seq = n - 1                      # callers append exactly one '$' (bwt.py:3053)
        want_mask = "period" in self.kinds and "tier1" in self.kinds
        local: Dict[str, tuple] = {}
        used = 0
        if "tier1" in self.kinds:
            c = self._rows_call(lane, used, "tier1_scan", lambda rec, cap, cnt: L.bwtk_tier1_scan(
                d_text.data_ptr(), n, self.max_motif_len, self.min_copies, self.min_array_len, self.min_entropy,
                rec, cap, cnt, lane.seen.data_ptr() if want_mask else None, wsp, wsb, st))
            local["tier1"] = (used, c)
            used += c
            self._mark(lane, "tier1")
        if "strict" in self.kinds:
            unit_cap = max(self.max_unit_len, min(n_seq // self.min_copies, 1000))   # bwt.py:3088-3096
            # the index is at hand: suffixes whose 16-mer occurs twice (LCP >= 16 with a neighbour) are the only places
            # where a unit length >= 16 can repeat
            _lib.check(L.bwtk_repeat_hint(lane.sa.data_ptr(), lane.lcp.data_ptr(), n, 16, lane.hint.data_ptr(), st),
                       "repeat_hint")
            c = self._rows_call(lane, used, "strict_scan", lambda rec, cap, cnt: L.bwtk_strict_scan_hinted(
                d_text.data_ptr(), n, 1, unit_cap, 0, self.min_copies, rec, cap, cnt, lane.hint.data_ptr(), 16,
                wsp, wsb, st))
            local["strict"] = (used, c)
            used += c
            self._mark(lane, "strict")
        if "plateaus" in self.kinds:
            thr = C.c_int64(-1)
            c = self._rows_call(lane, used, "lcp_plateaus", lambda rec, cap, cnt: L.bwtk_lcp_plateaus(
                d_text.data_ptr(), n, lane.sa.data_ptr(), lane.lcp.data_ptr(), n, 1, self.max_period,
                self.min_copies, rec, cap, cnt, C.addressof(thr), wsp, wsb, st))
            local["plateaus"] = (used, c)
            res.threshold = int(thr.value)
            used += c
            self._mark(lane, "plateaus")
        if "period" in self.kinds:
            it = C.c_int64(0)
            c = self._rows_call(lane, used, "period_scan", lambda rec, cap, cnt: L.bwtk_period_scan(
                d_text.data_ptr(), n, 1, self.max_period, int(self.allow_mismatches), self.min_copies,
                self.min_array_len, self.min_entropy, lane.seen.data_ptr() if want_mask else None,
                self.plogp.data_ptr(), self.plogp_dim, rec, cap, cnt, C.addressof(it), st))
            local["period"] = (used, c)
            res.period_iterations = int(it.value)
            used += c
            self._mark(lane, "period")
        return local

    def _lane_loop(self, lane: _Lane, contigs, ids, out, queue, download: bool, errors) -> None:
        torch = self.torch
        try:
            with torch.cuda.device(self.device), torch.cuda.stream(lane.stream):
                def claim(slot: int):
                    """Takes the next contig off the queue and starts its upload into text slot `slot`."""
                    with self._lock:
                        if not queue or errors:
                            return None
                        j = queue.pop(0)
                    t = contigs[j]
                    n = int(t.numel())
                    if n < 1 or n > self.max_n:
                        raise _lib.BwtkError(f"contig of {n} symbols does not fit the scanner (max_n={self.max_n})")
                    if t.is_cuda:
                        return j, t, n
                    d = lane.text[slot][:n]
                    with torch.cuda.stream(lane.up):
                        d.copy_(t, non_blocking=True)
                        lane.up_done[slot].record(lane.up)
                    lane.h2d += n
                    return j, d, n

                slot = 0
                nxt = claim(slot)
                while nxt is not None and not errors:
                    i, d_text, n = nxt
                    if not contigs[i].is_cuda:
                        lane.stream.wait_event(lane.up_done[slot])
                    # the other slot's contig finished on this stream before the host got here (the detector
                    # calls synchronise it): its text can be overwritten by the next upload
                    slot ^= 1
                    nxt = claim(slot)
                    res = ContigRows(contig=ids[i], n=n)
                    local = self._process(lane, d_text, n, res)
                    total = sum(c for _, c in local.values())
                    with self._lock:
                        first = self.used
                        if first + total > self.arena_rows:
                            raise _lib.BwtkError(f"{total} rows of contig {ids[i]} do not fit the arena "
                                                 f"({self.arena_rows - first} rows left of {self.arena_rows}); create the "
                                                 f"scanner with a larger arena_rows")
                        self.used = first + total
                    for k, (a, c) in local.items():
                        res.span[k] = (first + a, c)
                    if total:
                        self.arena[first:first + total].copy_(lane.scratch[:total], non_blocking=True)
                        if download:
                            lane.rows_ready.record(lane.stream)
                            with torch.cuda.stream(lane.dl):
                                lane.dl.wait_event(lane.rows_ready)
                                self.host[first:first + total].copy_(self.arena[first:first + total], non_blocking=True)
                            lane.d2h += total * REC_W * 4
                    out[i] = res
                lane.stream.wait_stream(lane.dl)
                lane.stream.wait_stream(lane.up)
                lane.stream.synchronize()
        except BaseException as exc:   # noqa: BLE001 -- re-raised by scan() on the caller's thread
            errors.append(exc)

    # ------------------------------------------------------------------ a list of contigs
    def scan(self, contigs: Sequence, ids: Optional[Sequence[int]] = None, download: bool = True) -> List[ContigRows]:
        """Processes the contigs, ``lanes`` at a time, taking them in the order given.  Each is a uint8 tensor
        with its '$': a CPU tensor (pinned for full-speed uploads) or a tensor already on this device.  Rows land
        in ``self.arena`` (device) and, with ``download``, at the same rows of ``self.host`` (pinned), densely from
        row 0 in completion order; ``ContigRows.span`` locates them.  Returns after everything (including the
        downloads) has completed; ``self.elapsed_ms`` is the device time from the first upload to the last
        download."""
        torch = self.torch
        if download and self.host is None:
            raise _lib.BwtkError("scanner was created without a host arena")
        ids = list(range(len(contigs))) if ids is None else list(ids)
        out: List[Optional[ContigRows]] = [None] * len(contigs)
        self.used = 0
        queue = list(range(len(contigs)))
        errors: list = []
        for lane in self.lanes:
            lane.events, lane.h2d, lane.d2h = [], 0, 0
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream()
            self.t_start.record(cur)
            active = self.lanes[: max(1, min(len(self.lanes), len(contigs)))]
            for lane in active:
                lane.stream.wait_event(self.t_start)
                lane.dl.wait_event(self.t_start)
                lane.up.wait_event(self.t_start)
            if len(active) == 1:
                self._lane_loop(active[0], contigs, ids, out, queue, download, errors)
            else:
                threads = [threading.Thread(target=self._lane_loop, args=(lane, contigs, ids, out, queue, download, errors),
                                            daemon=True) for lane in active]
                for th in threads:
                    th.start()
                for th in threads:
                    th.join()
            for lane in active:
                cur.wait_stream(lane.stream)
            self.t_end.record(cur)
            self.t_end.synchronize()
            if errors:
                raise errors[0]
            self.elapsed_ms = self.t_start.elapsed_time(self.t_end)
            self.h2d_bytes = sum(lane.h2d for lane in self.lanes)
            self.d2h_bytes = sum(lane.d2h for lane in self.lanes)
            if self.profile_phases:
                self.phase_ms = {}
                for lane in active:
                    for (_, a), (name, b) in zip(lane.events, lane.events[1:]):
                        if name != "start":
                            self.phase_ms[name] = self.phase_ms.get(name, 0.0) + a.elapsed_time(b)
        return out  # type: ignore[return-value]

    # ------------------------------------------------------------------ views
    def device_rows(self, res: ContigRows, kind: str):
        a, c = res.span[kind]
        return self.arena[a:a + c]

    def host_rows(self, res: ContigRows, kind: str) -> np.ndarray:
        a, c = res.span[kind]
        return self.host[a:a + c].numpy()
