"""ctypes binding of libbwtk.so (include/bwtk.h).

The product path has NO CPU fallback: if the shared library is missing or the
process has no CUDA device, every entry point raises.  PyTorch is used only to
own device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("BWTK_LIB") or os.path.join(_HERE, "libbwtk.so")  # BWTK_LIB: tuning variants
CSRC = os.path.join(_HERE, "csrc")

REC_W = 8
E_OVERFLOW = -4


class BwtkError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libbwtk.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1))]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise BwtkError("building libbwtk.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return SO_PATH


_lib = None
_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32

_SIGS = {
    "bwtk_version": (_i32, []),
    "bwtk_last_error": (_i32, [C.c_char_p, _i32]),
    "bwtk_launch_count": (_i64, []),
    "bwtk_profile_enable": (_i32, [_i32]),
    "bwtk_profile_report": (_i32, [C.c_char_p, _i32]),
    "bwtk_upload_text": (_i32, [_p, _p, _i64, _p]),
    "bwtk_download": (_i32, [_p, _p, _i64, _p]),
    "bwtk_device_count": (_i32, [_p]),
    "bwtk_dev_alloc": (_i32, [_p, _i64]),
    "bwtk_dev_free": (_i32, [_p]),
    "bwtk_copy_to_device": (_i32, [_p, _p, _i64, _p]),
    "bwtk_copy_to_host": (_i32, [_p, _p, _i64, _p]),
    "bwtk_fasta_index": (_i32, [_p, _i64, _p, _i64, _p, _p]),
    "bwtk_fasta_sequence": (_i64, [_p, _i64, _i64, _i64, _i64, _i32, _p]),
    "bwtk_suppress_nested": (_i32, [_p, _p, _p, _p, _i64, C.c_double, _p]),
    "bwtk_align_repeat_region": (_i32, [C.c_char_p, _i64, _i64, _i64, C.c_char_p, _i32, _i32, _i32, _i32, _p, _p, _p,
                                        _i64]),
    "bwtk_byte_histogram": (_i32, [_p, _i64, _p, _p]),
    "bwtk_sa_workspace_bytes": (_i64, [_i64]),
    "bwtk_sa_build": (_i32, [_p, _i64, _p, _p, _p, _i64, _p, _p]),
    "bwtk_index_workspace_bytes": (_i64, [_i64]),
    "bwtk_index_build": (_i32, [_p, _i64, _i32, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _i64, _p]),
    "bwtk_bwt_occ_workspace_bytes": (_i64, [_i64, _i32, _i32]),
    "bwtk_bwt_occ": (_i32, [_p, _p, _i64, _i32, _p, _i32, _p, _p, _p, _i64, _p]),
    "bwtk_lcp_workspace_bytes": (_i64, [_i64]),
    "bwtk_lcp_build": (_i32, [_p, _p, _i64, _p, _p, _i64, _p]),
    "bwtk_kmer8_workspace_bytes": (_i64, [_i64]),
    "bwtk_kmer8_index": (_i32, [_p, _i64, _p, _p, _p, _p, _i64, _p]),
    "bwtk_bsearch_batch": (_i32, [_p, _p, _i64, _i32, _p, _p, _p, _i64, _p, _i64, _p, _i64, _p, _p, _p]),
    "bwtk_bsearch_motif_sweep": (_i32, [_p, _p, _i64, _i32, _p, _p, _p, _i64, _i32, _p, _p, _p]),
    "bwtk_rank_batch": (_i32, [_p, _p, _i64, _i32, _p, _i64, _p, _p, _i64, _p, _p]),
    "bwtk_fm_pack_bytes": (_i64, [_i64]),
    "bwtk_fm_pack_workspace_bytes": (_i64, [_i64, _i64]),
    "bwtk_fm_pack": (_i32, [_p, _i64, _p, _p, _p, _p, _i64, _p, _p, _p, _i64, _p]),
    "bwtk_fm_search_batch": (_i32, [_p, _p, _i64, _p, _i64, _p, _p, _i32, _p]),
    "bwtk_fm_rank_batch": (_i32, [_p, _p, _p, _i64, _p, _p]),
    "bwtk_fm_motif_sweep": (_i32, [_p, _i32, _p, _p, _i32, _p]),
    "bwtk_tier1_workspace_bytes": (_i64, [_i64]),
    "bwtk_tier1_scan": (_i32, [_p, _i64, _i32, _i32, _i32, C.c_double, _p, _i64, _p, _p, _p, _i64, _p]),
    "bwtk_strict_workspace_bytes": (_i64, [_i64, _i64]),
    "bwtk_strict_scan": (_i32, [_p, _i64, _i64, _i64, _i64, _i64, _p, _i64, _p, _p, _i64, _p]),
    "bwtk_repeat_hint_bytes": (_i64, [_i64]),
    "bwtk_repeat_hint": (_i32, [_p, _p, _i64, _i32, _p, _p]),
    "bwtk_strict_scan_hinted": (_i32, [_p, _i64, _i64, _i64, _i64, _i64, _p, _i64, _p, _p, _i32, _p, _i64, _p]),
    "bwtk_plateau_workspace_bytes": (_i64, [_i64]),
    "bwtk_lcp_plateaus": (_i32, [_p, _i64, _p, _p, _i64, _i64, _i64, _i64, _p, _i64, _p, _p, _p, _i64, _p]),
    "bwtk_extend_batch": (_i32, [_p, _i64, _p, _p, _p, _i64, _i32, _p, _p]),
    "bwtk_consensus_batch": (_i32, [_p, _i64, _p, _p, _p, _p, _i64, _p, _p, _p]),
    "bwtk_period_scan": (_i32, [_p, _i64, _i64, _i64, _i32, _i64, _i64, C.c_double, _p, _p, _i64, _p, _i64,
                                _p, _p, _p]),
}

EXPORTS = tuple(_SIGS.keys())

FM_L2_PERSIST = 1
FM_THREAD_PER_QUERY = 2


class FmIndex(C.Structure):
    """``bwtk_fm_index`` of include/bwtk.h (device pointers as integers)."""
    _fields_ = [("d_blocks", _p), ("n", _i64), ("d_exc_pos", _p), ("d_exc_by_code", _p), ("n_exc", _i64),
                ("d_code_off", _p), ("d_C", _p), ("d_tot", _p), ("acgt_C", _i64 * 4), ("acgt_tot", _i64 * 4),
                ("d_ftab_sp", _p), ("d_ftab_ep", _p), ("ftab_k", _i32)]


def lib():
    """Load libbwtk.so; raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise BwtkError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    lib().bwtk_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise BwtkError(f"{what} failed (code {rc}): {last_error()}")


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise BwtkError("no CUDA device: the bwt hot path runs only on the GPU (no CPU fallback)")
    return torch


def ptr(t) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream
