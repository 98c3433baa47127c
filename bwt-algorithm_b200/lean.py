"""The CLI's start-up path: detection without `import torch`.

A fresh `python bwt.py contig.fa` spends seconds importing torch before the first kernel can run, for a detector
call that needs three device buffers.  When torch has not been imported by anyone yet, the per-contig worker
(pipeline._detect_rows, bwt.py:3040-3106) takes its buffers from the library itself (bwtk_dev_alloc /
bwtk_copy_to_device / bwtk_copy_to_host, include/bwtk.h) and calls the same `bwtk_strict_scan`: same kernels, same
rows.  Everything else -- the library API, GenomeScanner, the multi-GPU workers, any process in which torch is
already loaded -- keeps torch's allocator and streams.  Set BWTK_LEAN=0 to switch this path off.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

from . import _lib

REC_W = _lib.REC_W


def enabled() -> bool:
    """True while nobody in this process has imported torch (and BWTK_LEAN is not 0)."""
    return "torch" not in sys.modules and os.environ.get("BWTK_LEAN", "1") != "0"


def device_count() -> int:
    """Visible CUDA devices according to the library (0 without a driver, or when the library is missing)."""
    try:
        L = _lib.lib()
    except _lib.BwtkError:
        return 0
    n = C.c_int32(0)
    if L.bwtk_device_count(C.addressof(n)) != 0:
        return 0
    return int(n.value)


class _DeviceBuffer:
    def __init__(self, L, nbytes: int):
        self._L = L
        p = C.c_void_p(0)
        _lib.check(L.bwtk_dev_alloc(C.addressof(p), int(nbytes)), "device allocation")
        self.ptr = p.value

    def free(self) -> None:
        if self.ptr:
            ptr, self.ptr = self.ptr, None
            _lib.check(self._L.bwtk_dev_free(ptr), "device free")


def strict_rows(text_arr: np.ndarray, min_unit_len: int, max_unit_len: int, max_mismatch: int,
                min_copies: int) -> np.ndarray:
    """detect.strict_rows for a host text (uint8, with its '$') on library-owned buffers: rows
    (start, end, primitive_period, copies, 0, 0, unit_len, 0) in the reference's order."""
    L = _lib.lib()
    text_arr = np.ascontiguousarray(text_arr, np.uint8)
    n = int(text_arr.size)
    if n == 0:
        return np.zeros((0, REC_W), np.int32)
    held = []
    try:
        d_text = _DeviceBuffer(L, n + 64)
        held.append(d_text)
        _lib.check(L.bwtk_copy_to_device(d_text.ptr, text_arr.ctypes.data, n, None), "text upload")
        ws_bytes = max(int(L.bwtk_strict_workspace_bytes(n, int(max_unit_len))), 256)
        ws = _DeviceBuffer(L, ws_bytes)
        held.append(ws)
        cap = max(n // 8 + 4096, 1024)
        for _ in range(3):
            rec = _DeviceBuffer(L, cap * REC_W * 4)
            held.append(rec)
            cnt = C.c_int64(0)
            rc = L.bwtk_strict_scan(d_text.ptr, n, int(min_unit_len), int(max_unit_len), int(max_mismatch),
                                    int(min_copies), rec.ptr, cap, C.addressof(cnt), ws.ptr, ws_bytes, None)
            if rc == _lib.E_OVERFLOW and cnt.value > cap:
                cap = int(cnt.value)
                rec.free()
                continue
            _lib.check(rc, "strict scan")
            rows = np.empty((int(cnt.value), REC_W), np.int32)
            _lib.check(L.bwtk_copy_to_host(rows.ctypes.data, rec.ptr, rows.nbytes, None), "row download")
            return rows
        raise _lib.BwtkError("record buffer kept overflowing")
    finally:
        for buf in held:
            try:
                buf.free()
            except _lib.BwtkError:
                pass
