"""Drop-in module surface of the reference's ``bwt.py``.

Everything the reference module exports is importable from here under the same
name; the index-and-scan hot path underneath runs on the B200 through
libbwtk.so (include/bwtk.h).  There is no CPU fallback: without the shared
library or a CUDA device the hot-path calls raise.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .core import BWTCore, KmerTable
from .finders import Tier1STRFinder, Tier2LCPFinder
from .motifs import MotifUtils
from .pipeline import TandemRepeatFinder, _process_chromosome_worker, main
from .records import AlignmentResult, RepeatAlignmentSummary, TandemRepeat, _natural_sort_key

HAVE_NUMBA = False  # the reference's optional JIT seam (bwt.py:38-44) has no role here


def _count_equal_range(arr: np.ndarray, start: int, end: int, code: int) -> int:
    """# of `code` in arr[start:end] (bwt.py:46-53), counted by the byte
    histogram kernel."""
    import ctypes as C

    from . import detect

    start, end = max(0, int(start)), min(int(end), int(arr.size))
    if end <= start or not (0 <= int(code) <= 255):
        return 0
    d = detect.device_text(np.ascontiguousarray(arr[start:end], np.uint8))
    totals = np.zeros(256, np.int64)
    _lib.check(_lib.lib().bwtk_byte_histogram(d.data_ptr(), int(d.numel()), totals.ctypes.data, _lib.stream_ptr()),
               "byte_histogram")
    return int(totals[int(code)])


def _kasai_lcp_uint8(text_codes: np.ndarray, sa: np.ndarray) -> np.ndarray:
    """LCP array of (text, suffix array) (bwt.py:55-72) on the GPU."""
    torch = _lib.require_cuda()
    from . import detect

    n = int(text_codes.size)
    if n == 0:
        return np.zeros(0, dtype=np.int32)
    L = _lib.lib()
    d_text = detect.device_text(text_codes)
    d_sa = torch.from_numpy(np.ascontiguousarray(sa, np.int32)).to(d_text.device)
    lcp = torch.empty(n, dtype=torch.int32, device=d_text.device)
    wsb = int(L.bwtk_lcp_workspace_bytes(n))
    ws = torch.empty(wsb, dtype=torch.uint8, device=d_text.device)
    _lib.check(L.bwtk_lcp_build(d_text.data_ptr(), d_sa.data_ptr(), n, lcp.data_ptr(), ws.data_ptr(), wsb,
                                _lib.stream_ptr()), "lcp_build")
    return lcp.cpu().numpy()


class Tier3LongReadFinder:
    """Tier 3 (bwt.py:2828-3036) is outside the accelerated path:
    the class exists so imports keep working; using it raises."""

    def __init__(self, bwt_core, show_progress: bool = False):
        self.bwt = bwt_core
        self.min_read_length = 1000
        self.min_span_length = 100
        self.show_progress = show_progress

    def find_very_long_repeats(self, long_reads, chromosome):
        raise NotImplementedError("Tier 3 long-read analysis is out of scope of the B200 hot path")


__all__ = [
    "BWTCore", "KmerTable", "TandemRepeat", "AlignmentResult", "RepeatAlignmentSummary", "MotifUtils",
    "Tier1STRFinder", "Tier2LCPFinder", "Tier3LongReadFinder", "TandemRepeatFinder",
    "_process_chromosome_worker", "_natural_sort_key", "_count_equal_range", "_kasai_lcp_uint8", "HAVE_NUMBA",
    "main",
]

if __name__ == "__main__":
    main()
