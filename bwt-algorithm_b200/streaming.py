"""Index builds of a stream of host-resident contigs with transfers overlapped.

A chr21-sized build spends 3.7 ms in kernels but 8.5 ms on PCIe (46.7 MB of text up,
SA + BWT + LCP = 9 bytes per base down), so a per-contig loop is transfer-bound.
``IndexPipeline`` keeps ``slots`` contigs in flight, each on its own CUDA stream with
its own device buffers and pinned host buffers: while contig i's arrays travel to the
host, contig i+1 is uploaded and built.  ``bwtk_index_build`` only synchronises the
stream it is given, so one host thread drives all slots.

This is the host side of BWTCore.__init__ (bwt.py:106-136) for callers that process
many contigs (TandemRepeatFinder.build_indices, bwt.py:3229-3262); there is no CPU
fallback.
"""
from __future__ import annotations

from typing import List

import numpy as np

from . import _lib


class _Slot:
    def __init__(self, torch, L, cap: int, occ_rate: int, device, want_lcp: bool):
        self.cap = cap
        self.stream = torch.cuda.Stream(device=device)
        self.d_text = torch.empty(cap, dtype=torch.uint8, device=device)
        self.sa = torch.empty(cap, dtype=torch.int32, device=device)
        self.bwt = torch.empty(cap, dtype=torch.uint8, device=device)
        self.lcp = torch.empty(cap, dtype=torch.int32, device=device) if want_lcp else None
        self.rows_cap = 8
        self.ncp_cap = cap // occ_rate + 2
        self.occ = torch.zeros((self.rows_cap, self.ncp_cap), dtype=torch.int32, device=device)
        self.ws = torch.empty(int(L.bwtk_index_workspace_bytes(cap)), dtype=torch.uint8, device=device)
        self.h_sa = torch.empty(cap, dtype=torch.int32).pin_memory()
        self.h_bwt = torch.empty(cap, dtype=torch.uint8).pin_memory()
        self.h_lcp = torch.empty(cap, dtype=torch.int32).pin_memory() if want_lcp else None
        self.totals = np.zeros(256, np.int64)
        self.row = np.full(256, -1, np.int32)
        self.stats = np.zeros(8, np.int64)
        self.start = torch.cuda.Event(enable_timing=True)
        self.done = torch.cuda.Event(enable_timing=True)
        self.built = torch.cuda.Event()
        self.n = 0
        self.busy = False


class IndexPipeline:
    """``submit(host_text)`` enqueues upload + fused index build + download on the next slot and
    returns a ticket; ``result(ticket)`` waits for that contig and returns NumPy views of the pinned
    result buffers (valid until the slot is reused, i.e. until ``slots`` further submits)."""

    def __init__(self, max_n: int, occ_rate: int = 128, device=None, slots: int = 2, want_lcp: bool = True):
        torch = _lib.require_cuda()
        self.torch = torch
        self.L = _lib.lib()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.occ_rate = int(occ_rate)
        self.max_n = int(max_n)
        with torch.cuda.device(self.device):
            self.slots: List[_Slot] = [_Slot(torch, self.L, self.max_n, self.occ_rate, self.device, want_lcp)
                                       for _ in range(max(1, int(slots)))]
        self.tickets = 0
        with torch.cuda.device(self.device):
            self.dl_stream = torch.cuda.Stream(device=self.device)
        self.kernel_copies = False   # True: transfers by SM kernels (bwtk_upload_text / bwtk_download)

    def submit(self, host_text) -> int:
        """host_text: uint8 NumPy array or (ideally pinned) CPU tensor, at most max_n symbols."""
        torch, L = self.torch, self.L
        t = host_text if isinstance(host_text, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(host_text, np.uint8))
        n = int(t.numel())
        if n == 0 or n > self.max_n:
            raise _lib.BwtkError(f"contig of {n} symbols does not fit the pipeline (max_n={self.max_n})")
        ticket = self.tickets
        slot = self.slots[ticket % len(self.slots)]
        if slot.busy:
            slot.done.synchronize()     # its previous contig (and the D2H into its pinned buffers) is complete
        slot.n = n
        slot.busy = True
        with torch.cuda.device(self.device), torch.cuda.stream(slot.stream):
            slot.start.record()
            if t.is_pinned() and self.kernel_copies:
                # kernel upload: a copy engine may be busy (or queued) with another slot's download
                _lib.check(L.bwtk_upload_text(t.data_ptr(), slot.d_text.data_ptr(), n, slot.stream.cuda_stream),
                           "upload_text")
            else:
                slot.d_text[:n].copy_(t, non_blocking=True)
            for _attempt in range(2):
                rc = L.bwtk_index_build(slot.d_text.data_ptr(), n, self.occ_rate, slot.sa.data_ptr(), None,
                                        slot.bwt.data_ptr(), slot.occ.data_ptr(), slot.rows_cap, _lib.ptr(slot.lcp),
                                        slot.totals.ctypes.data, slot.row.ctypes.data, slot.stats.ctypes.data,
                                        slot.ws.data_ptr(), slot.ws.numel(), slot.stream.cuda_stream)
                if rc == _lib.E_OVERFLOW and slot.stats[7] > slot.rows_cap:
                    # more than 8 distinct byte values (IUPAC codes, soft-masked text): a taller Occ matrix
                    slot.rows_cap = int(slot.stats[7])
                    slot.occ = torch.zeros((slot.rows_cap, slot.ncp_cap), dtype=torch.int32, device=self.device)
                    continue
                break
            _lib.check(rc, "index_build")
            slot.built.record()
        # All downloads go through ONE stream, in order.  A download enqueued on its own stream
        # while another stream's download occupies the copy engine gets parked in hardware, and
        # every kernel submitted afterwards -- the next contig's build -- waits for it (measured:
        # every second build started 7 ms late).  Stream order keeps the queue in software.
        with torch.cuda.device(self.device), torch.cuda.stream(self.dl_stream):
            self.dl_stream.wait_event(slot.built)
            if self.kernel_copies:
                for dst, src in ((slot.h_sa, slot.sa), (slot.h_bwt, slot.bwt), (slot.h_lcp, slot.lcp)):
                    if src is not None:
                        _lib.check(L.bwtk_download(src.data_ptr(), dst.data_ptr(), n * src.element_size(),
                                                   self.dl_stream.cuda_stream), "download")
            else:
                slot.h_sa[:n].copy_(slot.sa[:n], non_blocking=True)
                slot.h_bwt[:n].copy_(slot.bwt[:n], non_blocking=True)
                if slot.lcp is not None:
                    slot.h_lcp[:n].copy_(slot.lcp[:n], non_blocking=True)
            slot.done.record()
        self.tickets += 1
        return ticket

    def result(self, ticket: int) -> dict:
        if not 0 <= ticket < self.tickets or ticket < self.tickets - len(self.slots):
            raise _lib.BwtkError("that ticket's slot has been reused")
        slot = self.slots[ticket % len(self.slots)]
        slot.done.synchronize()
        n = slot.n
        ncp = n // self.occ_rate + 1 + (1 if n % self.occ_rate else 0)
        nrows = int(slot.stats[7])
        # bwtk_index_build writes Occ rows with stride ncp(n), not with the slot's capacity stride, and the
        # slot's buffer is overwritten by the next submit: hand out a right-shaped copy (device tensor)
        occ = slot.occ.view(-1)[: nrows * ncp].view(nrows, ncp).clone()
        out = {"n": n, "sa": slot.h_sa[:n].numpy(), "bwt": slot.h_bwt[:n].numpy(),
               "totals": slot.totals.copy(), "row_of_code": slot.row.copy(), "stats": slot.stats.copy(),
               "occ": occ, "ncp": ncp}
        if slot.h_lcp is not None:
            out["lcp"] = slot.h_lcp[:n].numpy()
        return out

    def drain(self) -> None:
        for s in self.slots:
            if s.busy:
                s.done.synchronize()

    def bytes_per_contig(self, n: int) -> tuple:
        """(host->device, device->host) bytes moved for a contig of n symbols."""
        return n, n * (4 + 1 + (4 if self.slots[0].lcp is not None else 0))
