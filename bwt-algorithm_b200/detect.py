"""Row-level launchers of the detector kernels (include/bwtk.h).

Every function takes a text that is either a device tensor (uint8, CUDA) or a
host array (uploaded once), launches the kernels through the C ABI and returns
integer record rows ``int32[R, 8]`` = (start, end, period, copies, total_mm,
max_mm_per_copy, aux0, aux1) on the host.  The float fields of a
``TandemRepeat`` are derived from these rows in finders.py.
"""
from __future__ import annotations

import ctypes as C
import threading
import weakref
from typing import List, Optional, Tuple

import numpy as np

from . import _lib

REC_W = _lib.REC_W


def device_text(text_arr, device=None):
    """uint8 CUDA tensor for a host array / bytes / str; device tensors pass through."""
    torch = _lib.require_cuda()
    if isinstance(text_arr, torch.Tensor):
        if not text_arr.is_cuda:
            return text_arr.to(device or "cuda").contiguous()
        return text_arr.contiguous()
    if isinstance(text_arr, str):
        text_arr = text_arr.encode("utf-8")
    if isinstance(text_arr, (bytes, bytearray, memoryview)):
        host = np.frombuffer(bytes(text_arr), dtype=np.uint8)
    else:
        host = np.ascontiguousarray(text_arr, dtype=np.uint8)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if host.size == 0:
        return torch.empty(0, dtype=torch.uint8, device=dev)
    return torch.from_numpy(host.copy()).to(dev)


def _run_rows(call, ws_bytes: int, device, cap0: int) -> np.ndarray:
    """Runs `call(d_rec_ptr, cap, count_ptr, ws_ptr, ws_bytes)`; grows the record
    buffer once when the kernel reports BWTK_EOVERFLOW (results are deterministic)."""
    torch = _lib.require_cuda()
    with torch.cuda.device(device):
        ws = torch.empty(max(int(ws_bytes), 256), dtype=torch.uint8, device=device)
        cap = max(int(cap0), 1024)
        for _ in range(3):
            rec = torch.empty((cap, REC_W), dtype=torch.int32, device=device)
            cnt = C.c_int64(0)
            rc = call(rec.data_ptr(), cap, C.addressof(cnt), ws.data_ptr(), ws.numel())
            if rc == _lib.E_OVERFLOW and cnt.value > cap:
                cap = int(cnt.value)
                continue
            _lib.check(rc, "detector kernel")
            return _rows_to_host(torch, rec[: cnt.value])
    raise _lib.BwtkError("record buffer kept overflowing")


class _PinnedPool:
    """Pinned host blocks for detector rows, lent out as NumPy arrays and taken back when the
    array (and every view of it) has been garbage-collected.

    torch's caching host allocator reuses a block only inside its power-of-two size class, and
    the first block of a class costs a cudaHostAlloc of 100+ ms: the strict scans of a genome's
    contigs (425, 292, 267, 174, 80 MB of rows ...) hit four classes.  Here any free block that
    is large enough serves a request, so one block per rows-in-flight is allocated for the
    largest contig and then reused by every smaller one."""

    KEEP_FREE = 3

    def __init__(self):
        self._free: List = []
        self._lock = threading.Lock()

    def _give(self, block) -> None:
        with self._lock:
            self._free.append(block)
            if len(self._free) > self.KEEP_FREE:
                self._free.remove(min(self._free, key=lambda b: b.numel()))

    def lend(self, torch, d_rows) -> np.ndarray:
        nbytes = d_rows.numel() * d_rows.element_size()
        with self._lock:
            fits = [b for b in self._free if b.numel() >= nbytes]
            block = min(fits, key=lambda b: b.numel()) if fits else None
            if block is not None:
                self._free.remove(block)
        if block is None:
            block = torch.empty(1 << max(nbytes - 1, 1).bit_length(), dtype=torch.uint8, pin_memory=True)
        h = block[:nbytes].view(d_rows.dtype).view(d_rows.shape)
        h.copy_(d_rows, non_blocking=True)
        torch.cuda.current_stream(d_rows.device).synchronize()
        out = h.numpy()               # views of `out` keep `out` alive, `out` keeps the storage alive
        weakref.finalize(out, self._give, block)
        return out


_pinned_rows = _PinnedPool()


def _rows_to_host(torch, d_rows) -> np.ndarray:
    """Device rows -> NumPy.  Large results (a chr21-sized strict scan returns 80 MB) land in pinned
    memory (``_PinnedPool``): the download runs at PCIe speed instead of the ~2 GB/s of a copy into
    freshly allocated pageable memory, and the array owns its block until it is collected."""
    if d_rows.numel() * d_rows.element_size() >= (8 << 20):
        try:
            return _pinned_rows.lend(torch, d_rows)
        except RuntimeError:
            pass   # pinned memory exhausted: fall through to the pageable copy
    return d_rows.cpu().numpy()


def tier1_rows(text, max_motif_len: int = 9, min_copies: int = 3, min_array_len: int = 6,
               min_entropy: float = 1.0, want_seen: bool = False):
    """bwtk_tier1_scan -> rows (start,end,motif_len,copies,...) in emission order."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    d = device_text(text)
    n = int(d.numel())
    if n == 0:
        return (np.zeros((0, REC_W), np.int32), np.zeros(0, np.uint8)) if want_seen else np.zeros((0, REC_W), np.int32)
    seen = torch.zeros(n, dtype=torch.uint8, device=d.device) if want_seen else None

    def call(rec, cap, cnt, ws, wsb):
        return L.bwtk_tier1_scan(d.data_ptr(), n, max_motif_len, min_copies, min_array_len, float(min_entropy),
                                 rec, cap, cnt, _lib.ptr(seen), ws, wsb, _lib.stream_ptr())

    rows = _run_rows(call, L.bwtk_tier1_workspace_bytes(n), d.device, n // 16 + 1024)
    return (rows, seen.cpu().numpy()) if want_seen else rows


def repeat_hint(sa, lcp, n: int, min_len: int = 16):
    """bwtk_repeat_hint -> int32 bitmap tensor: bit i set iff suffix i shares >= min_len symbols with a neighbour in
    suffix order (its min_len-mer occurs more than once).  `sa`, `lcp`: device int32 tensors of the text's index."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    with torch.cuda.device(sa.device):
        bits = torch.empty(int(L.bwtk_repeat_hint_bytes(n)) // 4 + 4, dtype=torch.int32, device=sa.device)
        _lib.check(L.bwtk_repeat_hint(sa.data_ptr(), lcp.data_ptr(), n, int(min_len), bits.data_ptr(), _lib.stream_ptr()),
                   "repeat_hint")
    return bits


def strict_rows(text, min_unit_len: int = 20, max_unit_len: int = 120, max_mismatch: int = 2,
                min_copies: int = 3, hint=None, hint_len: int = 16) -> np.ndarray:
    """bwtk_strict_scan -> rows (start,end,primitive_period,copies,0,0,unit_len,0).  `hint`: the bitmap of
    ``repeat_hint`` for the same text (bwtk_strict_scan_hinted: same rows, far fewer compares)."""
    L = _lib.lib()
    d = device_text(text)
    n = int(d.numel())
    if n == 0:
        return np.zeros((0, REC_W), np.int32)

    def call(rec, cap, cnt, ws, wsb):
        if hint is not None:
            return L.bwtk_strict_scan_hinted(d.data_ptr(), n, int(min_unit_len), int(max_unit_len), int(max_mismatch),
                                             int(min_copies), rec, cap, cnt, hint.data_ptr(), int(hint_len), ws, wsb,
                                             _lib.stream_ptr())
        return L.bwtk_strict_scan(d.data_ptr(), n, int(min_unit_len), int(max_unit_len), int(max_mismatch),
                                  int(min_copies), rec, cap, cnt, ws, wsb, _lib.stream_ptr())

    return _run_rows(call, L.bwtk_strict_workspace_bytes(n, int(max_unit_len)), d.device, n // 8 + 4096)


def plateau_rows(text, sa, lcp, min_period: int = 1, max_period: int = 1000,
                 min_copies: int = 3) -> Tuple[np.ndarray, int]:
    """bwtk_lcp_plateaus -> (rows (start,end,period,copies,...), threshold or -1)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    d = device_text(text)
    dev = d.device
    d_sa = sa if isinstance(sa, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(sa, np.int32)).to(dev)
    d_lcp = lcp if isinstance(lcp, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(lcp, np.int32)).to(dev)
    n = int(d_lcp.numel())
    thr = C.c_int64(-1)
    if n == 0:
        return np.zeros((0, REC_W), np.int32), -1

    def call(rec, cap, cnt, ws, wsb):
        return L.bwtk_lcp_plateaus(d.data_ptr(), int(d.numel()), d_sa.data_ptr(), d_lcp.data_ptr(), n,
                                   int(min_period), int(max_period), int(min_copies), rec, cap, cnt,
                                   C.addressof(thr), ws, wsb, _lib.stream_ptr())

    rows = _run_rows(call, L.bwtk_plateau_workspace_bytes(n), dev, n // 16 + 1024)
    return rows, int(thr.value)


def consensus_batch(text, starts, periods, copies):
    """bwtk_consensus_batch -> (list of consensus uint8 arrays, mm int32[m,4])."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    d = device_text(text)
    dev = d.device
    starts = np.ascontiguousarray(starts, np.int32)
    periods = np.ascontiguousarray(periods, np.int32)
    copies = np.ascontiguousarray(copies, np.int32)
    m = starts.size
    if m == 0:
        return [], np.zeros((0, 4), np.int32)
    offs = np.zeros(m + 1, np.int64)
    np.cumsum(periods.astype(np.int64), out=offs[1:])
    with torch.cuda.device(dev):
        d_s = torch.from_numpy(starts).to(dev)
        d_p = torch.from_numpy(periods).to(dev)
        d_c = torch.from_numpy(copies).to(dev)
        d_o = torch.from_numpy(offs).to(dev)
        d_cons = torch.zeros(max(int(offs[-1]), 1), dtype=torch.uint8, device=dev)
        d_mm = torch.zeros((m, 4), dtype=torch.int32, device=dev)
        _lib.check(L.bwtk_consensus_batch(d.data_ptr(), int(d.numel()), d_s.data_ptr(), d_p.data_ptr(),
                                          d_c.data_ptr(), d_o.data_ptr(), m, d_cons.data_ptr(), d_mm.data_ptr(),
                                          _lib.stream_ptr()), "consensus_batch")
        cons = d_cons.cpu().numpy()
        mm = d_mm.cpu().numpy()
    return [cons[offs[i]:offs[i + 1]].copy() for i in range(m)], mm


def consensus_one(text, start: int, motif_len: int, n_copies: int):
    """(consensus, total_mm, max_mm_per_copy, copies_used) for one array."""
    cons, mm = consensus_batch(text, [start], [motif_len], [n_copies])
    return cons[0], int(mm[0, 0]), int(mm[0, 1]), int(mm[0, 2])


def extend_batch(text, n: int, seeds, periods, flags=None, mode: int = 0) -> np.ndarray:
    """bwtk_extend_batch -> int32[m, 8] rows (see include/bwtk.h)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    d = device_text(text)
    dev = d.device
    seeds = np.ascontiguousarray(seeds, np.int32)
    periods = np.ascontiguousarray(periods, np.int32)
    m = seeds.size
    if m == 0:
        return np.zeros((0, 8), np.int32)
    flags = np.ones(m, np.int32) if flags is None else np.ascontiguousarray(flags, np.int32)
    with torch.cuda.device(dev):
        d_s = torch.from_numpy(seeds).to(dev)
        d_p = torch.from_numpy(periods).to(dev)
        d_f = torch.from_numpy(flags).to(dev)
        out = torch.zeros((m, 8), dtype=torch.int32, device=dev)
        _lib.check(L.bwtk_extend_batch(d.data_ptr(), int(n), d_s.data_ptr(), d_p.data_ptr(), d_f.data_ptr(), m,
                                       int(mode), out.data_ptr(), _lib.stream_ptr()), "extend_batch")
        return out.cpu().numpy()


def plogp_table(dim: int) -> np.ndarray:
    """plogp[c*dim + L] = (c/L)*np.log2(c/L): the entropy terms exactly as the
    host (numpy) evaluates them, handed to kernels that gate on entropy."""
    t = np.zeros((dim, dim), np.float64)
    for L in range(1, dim):
        c = np.arange(1, L + 1, dtype=np.float64)
        p = c / float(L)
        t[1:L + 1, L] = p * np.log2(p)
    return t


def period_scan_rows(text, min_period: int = 1, max_period: int = 1000, allow_mismatches: bool = True,
                     min_copies: int = 3, min_array_len: int = 6, min_entropy: float = 1.0,
                     tier1_mask: Optional[np.ndarray] = None):
    """bwtk_period_scan -> (raw rows before the host (start,end,canonical) dedup, iterations)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    d = device_text(text)
    dev = d.device
    n = int(d.numel())
    if n == 0:
        return np.zeros((0, REC_W), np.int32), 0
    dim = 202
    with torch.cuda.device(dev):
        d_tab = torch.from_numpy(plogp_table(dim).reshape(-1)).to(dev)
        d_mask = None
        if tier1_mask is not None:
            d_mask = torch.from_numpy(np.ascontiguousarray(tier1_mask, np.uint8)).to(dev)
        cap = 1 << 14
        for _ in range(3):
            rec = torch.empty((cap, REC_W), dtype=torch.int32, device=dev)
            cnt = C.c_int64(0)
            it = C.c_int64(0)
            rc = L.bwtk_period_scan(d.data_ptr(), n, int(min_period), int(max_period), int(bool(allow_mismatches)),
                                    int(min_copies), int(min_array_len), float(min_entropy), _lib.ptr(d_mask),
                                    d_tab.data_ptr(), dim, rec.data_ptr(), cap, C.addressof(cnt),
                                    C.addressof(it), _lib.stream_ptr())
            if rc == _lib.E_OVERFLOW and cnt.value > cap:
                cap = int(cnt.value)
                continue
            _lib.check(rc, "period_scan")
            return rec[: cnt.value].cpu().numpy(), int(it.value)
    raise _lib.BwtkError("record buffer kept overflowing")
