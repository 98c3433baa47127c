"""Device-resident FM index of one contig: the arrays behind ``BWTCore``.

Layout in HBM (all torch tensors on one CUDA device):
  text   uint8[n]      the text bytes (with the caller's '$')
  sa     int32[n]      suffix array            (bwt.py:212-264)
  isa    int32[n]      inverse suffix array    (by-product of prefix doubling)
  bwt    uint8[n]      BWT bytes               (bwt.py:266-274)
  occ    int32[σ][ncp] checkpoints every occ_rate symbols (bwt.py:288-326)
  C/tot  int64[256]    FM "C" array / totals   (bwt.py:276-286)
  lcp    int32[n]      LCP array, built on demand (bwt.py:55-72)
  kmer   int32[65537] bucket offsets + int32[R] positions (bwt.py:138-171)
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import numpy as np

from . import _lib


def _torch():
    return _lib.require_cuda()


class _FMSearch:
    """Rank / backward search over (bwt, occ, C) held on one device; shared by the full
    ``DeviceIndex`` and the search-only ``FMReplica``."""

    # ------------------------------------------------------------------ packed search index
    #: searches run over the 64-byte rank blocks (bwtk_fm_*); False = the byte BWT + Occ rows
    use_packed = True
    #: length of the k-mer interval table that seeds longer patterns (0 = none; None = 8 from 2^20 symbols on:
    #: the 65 536 intervals are level 8 of the motif sweep, 0.1 ms to build, and save 8 of a pattern's LF steps)
    ftab_k = None
    #: keep the rank blocks resident in L2 while a batch streams through
    l2_persist = True
    #: one thread per query instead of four lanes sharing every block load (kept for A/B measurements)
    thread_per_query = False

    def ensure_packed(self):
        """Builds the packed search index (bwtk_fm_pack) from the BWT bytes on first use."""
        if getattr(self, "_fm", None) is not None:
            return self._fm
        torch, L, n = self.torch, _lib.lib(), self.n
        if n < 1:
            raise _lib.BwtkError("empty text has no search index")
        tot = np.ascontiguousarray(self.totals, np.int64)
        n_exc = int(n - sum(int(tot[c]) for c in b"ACGT"))
        with torch.cuda.device(self.device):
            blocks = torch.empty(int(L.bwtk_fm_pack_bytes(n)), dtype=torch.uint8, device=self.device)
            exc_pos = torch.empty(max(n_exc, 1), dtype=torch.int32, device=self.device)
            exc_by_code = torch.empty(max(n_exc, 1), dtype=torch.int32, device=self.device)
            code_off = torch.zeros(257, dtype=torch.int64, device=self.device)
            wsb = int(L.bwtk_fm_pack_workspace_bytes(n, n_exc))
            ws = torch.empty(wsb, dtype=torch.uint8, device=self.device)
            cnt = C.c_int64(0)
            _lib.check(L.bwtk_fm_pack(self.bwt.data_ptr(), n, tot.ctypes.data, blocks.data_ptr(), exc_pos.data_ptr(),
                                      exc_by_code.data_ptr(), max(n_exc, 1), code_off.data_ptr(), C.addressof(cnt),
                                      ws.data_ptr(), wsb, _lib.stream_ptr()), "fm_pack")
        fx = _lib.FmIndex()
        fx.d_blocks, fx.n = blocks.data_ptr(), n
        fx.d_exc_pos, fx.d_exc_by_code, fx.n_exc = exc_pos.data_ptr(), exc_by_code.data_ptr(), int(cnt.value)
        fx.d_code_off, fx.d_C, fx.d_tot = code_off.data_ptr(), self.d_C.data_ptr(), self.d_tot.data_ptr()
        counts = np.concatenate(([0], np.cumsum(tot)[:-1]))
        for k, c in enumerate(b"ACGT"):
            fx.acgt_C[k] = int(counts[c])
            fx.acgt_tot[k] = int(tot[c])
        fx.d_ftab_sp = fx.d_ftab_ep = None
        fx.ftab_k = 0
        self._fm = fx
        self._fm_keep = (blocks, exc_pos, exc_by_code, code_off)      # the struct holds raw pointers
        self.packed_bytes = blocks.numel() + 8 * int(cnt.value)
        k = self.ftab_k if self.ftab_k is not None else (8 if n >= (1 << 20) else 0)
        if k:
            self.build_ftab(k)
        return fx

    def build_ftab(self, k: int):
        """Interval table of every ACGT k-mer (level k of the motif sweep): patterns of length >= k
        whose last k characters are ACGT start from it and do k fewer LF steps."""
        fx = self.ensure_packed() if getattr(self, "_fm", None) is None else self._fm
        fx.ftab_k = 0
        sp, ep = self.motif_sweep(k)
        off = (4 ** k - 4) // 3
        self._ftab = (sp[off:off + 4 ** k].contiguous(), ep[off:off + 4 ** k].contiguous())
        fx.d_ftab_sp, fx.d_ftab_ep, fx.ftab_k = self._ftab[0].data_ptr(), self._ftab[1].data_ptr(), int(k)
        self.ftab_k = int(k)

    def _flags(self) -> int:
        return (_lib.FM_L2_PERSIST if self.l2_persist else 0) | (_lib.FM_THREAD_PER_QUERY if self.thread_per_query else 0)

    # ------------------------------------------------------------------ a8 / a9
    def search_device(self, d_pats, stride: int, d_lens, nq: int):
        torch, L = self.torch, _lib.lib()
        with torch.cuda.device(self.device):
            sp = torch.empty(nq, dtype=torch.int32, device=self.device)
            ep = torch.empty(nq, dtype=torch.int32, device=self.device)
            if self.use_packed and self.n >= 1:
                fx = self.ensure_packed()
                _lib.check(L.bwtk_fm_search_batch(C.addressof(fx), _lib.ptr(d_pats), stride, d_lens.data_ptr(), nq,
                                                  sp.data_ptr(), ep.data_ptr(), self._flags(), _lib.stream_ptr()),
                           "fm_search_batch")
            else:
                _lib.check(L.bwtk_bsearch_batch(self.bwt.data_ptr(), self.occ.data_ptr(), self.occ.shape[1],
                                                self.occ_rate, self.d_C.data_ptr(), self.d_tot.data_ptr(),
                                                self.d_row.data_ptr(), self.n, _lib.ptr(d_pats), stride,
                                                d_lens.data_ptr(), nq, sp.data_ptr(), ep.data_ptr(),
                                                _lib.stream_ptr()), "bsearch_batch")
        return sp, ep

    def search_block(self, d_pats, d_lens):
        """(sp, ep) device tensors for a block of a padded query matrix (uint8[nq, stride],
        int32[nq]); the callable ``sharding.sharded_search`` expects."""
        nq = int(d_lens.shape[0])
        d_pats = d_pats.contiguous()
        stride = int(d_pats.shape[1]) if d_pats.dim() == 2 and nq else 1
        return self.search_device(d_pats, stride, d_lens.contiguous(), nq)

    def backward_search_batch(self, patterns: Sequence) -> Tuple[np.ndarray, np.ndarray]:
        """Inclusive (sp, ep) for every pattern (bytes/str); (-1,-1) when absent."""
        torch = self.torch
        nq = len(patterns)
        if nq == 0:
            return np.zeros(0, np.int32), np.zeros(0, np.int32)
        bs = [p.encode("utf-8") if isinstance(p, str) else bytes(p) for p in patterns]
        stride = max(1, max(len(b) for b in bs))
        mat = np.zeros((nq, stride), np.uint8)
        lens = np.zeros(nq, np.int32)
        for i, b in enumerate(bs):
            lens[i] = len(b)
            if b:
                mat[i, : len(b)] = np.frombuffer(b, np.uint8)
        d_p = torch.from_numpy(mat).to(self.device)
        d_l = torch.from_numpy(lens).to(self.device)
        sp, ep = self.search_device(d_p, stride, d_l, nq)
        return sp.cpu().numpy(), ep.cpu().numpy()

    def motif_sweep(self, kmax: int = 10):
        """(sp, ep) device tensors for every ACGT motif of length 1..kmax, index
        (4^k-4)/3 + base-4 value (first character most significant)."""
        torch, L = self.torch, _lib.lib()
        total = (4 ** (kmax + 1) - 4) // 3
        with torch.cuda.device(self.device):
            sp = torch.empty(total, dtype=torch.int32, device=self.device)
            ep = torch.empty(total, dtype=torch.int32, device=self.device)
            if self.use_packed and self.n >= 1:
                fx = self.ensure_packed()
                _lib.check(L.bwtk_fm_motif_sweep(C.addressof(fx), kmax, sp.data_ptr(), ep.data_ptr(), self._flags(),
                                                 _lib.stream_ptr()), "fm_motif_sweep")
                return sp, ep
            _lib.check(L.bwtk_bsearch_motif_sweep(self.bwt.data_ptr(), self.occ.data_ptr(), self.occ.shape[1],
                                                  self.occ_rate, self.d_C.data_ptr(), self.d_tot.data_ptr(),
                                                  self.d_row.data_ptr(), self.n, kmax, sp.data_ptr(),
                                                  ep.data_ptr(), _lib.stream_ptr()), "motif_sweep")
        return sp, ep

    def rank_batch(self, codes: Sequence[int], positions: Sequence[int]) -> np.ndarray:
        torch, L = self.torch, _lib.lib()
        nq = len(codes)
        if nq == 0:
            return np.zeros(0, np.int64)
        with torch.cuda.device(self.device):
            d_c = torch.tensor(list(codes), dtype=torch.int32, device=self.device)
            d_p = torch.tensor(list(positions), dtype=torch.int64, device=self.device)
            out = torch.empty(nq, dtype=torch.int64, device=self.device)
            if self.use_packed and self.n >= 1:
                fx = self.ensure_packed()
                _lib.check(L.bwtk_fm_rank_batch(C.addressof(fx), d_c.data_ptr(), d_p.data_ptr(), nq, out.data_ptr(),
                                                _lib.stream_ptr()), "fm_rank_batch")
                return out.cpu().numpy()
            _lib.check(L.bwtk_rank_batch(self.bwt.data_ptr(), self.occ.data_ptr(), self.occ.shape[1],
                                         self.occ_rate, self.d_row.data_ptr(), self.n, d_c.data_ptr(),
                                         d_p.data_ptr(), nq, out.data_ptr(), _lib.stream_ptr()), "rank_batch")
        return out.cpu().numpy()


class DeviceIndex(_FMSearch):
    def __init__(self, text, occ_rate: int = 128, device=None, build_kmer: bool = True,
                 text_is_device: bool = False, build_lcp: bool = False, build_isa: bool = False):
        torch = _torch()
        self.torch = torch
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.occ_rate = int(occ_rate)
        L = _lib.lib()
        with torch.cuda.device(self.device):
            if text_is_device:
                self.text = text
            else:
                if isinstance(text, str):
                    text = text.encode("utf-8")
                if isinstance(text, (bytes, bytearray, memoryview)):
                    host = np.frombuffer(text, dtype=np.uint8)
                else:
                    host = np.ascontiguousarray(text, dtype=np.uint8)
                self.text = torch.from_numpy(host.copy() if host.size else np.zeros(0, np.uint8)).to(self.device)
            n = int(self.text.numel())
            self.n = n
            st = _lib.stream_ptr()
            # a3 + a4 + a5 + a6 + a10 in one fused call (one histogram, one packed text)
            totals = np.zeros(256, np.int64)
            row = np.full(256, -1, np.int32)
            self.sa = torch.empty(n, dtype=torch.int32, device=self.device)
            self.isa = torch.empty(n, dtype=torch.int32, device=self.device) if build_isa else None
            self.bwt = torch.empty(n, dtype=torch.uint8, device=self.device)
            self.sa_stats = np.zeros(8, np.int64)
            self.ncp = (n // self.occ_rate + 1 + (1 if n % self.occ_rate else 0)) if n else 0
            self._lcp = torch.empty(n, dtype=torch.int32, device=self.device) if build_lcp else None
            rows_cap = 8
            self.occ = torch.zeros((rows_cap, max(self.ncp, 1)), dtype=torch.int32, device=self.device)
            if n:
                wsb = int(L.bwtk_index_workspace_bytes(n))
                ws = torch.empty(wsb, dtype=torch.uint8, device=self.device)
                for _attempt in range(2):
                    rc = L.bwtk_index_build(self.text.data_ptr(), n, self.occ_rate, self.sa.data_ptr(),
                                            _lib.ptr(self.isa), self.bwt.data_ptr(), self.occ.data_ptr(), rows_cap,
                                            _lib.ptr(self._lcp), totals.ctypes.data, row.ctypes.data,
                                            self.sa_stats.ctypes.data, ws.data_ptr(), wsb, st)
                    if rc == _lib.E_OVERFLOW and self.sa_stats[7] > rows_cap:
                        rows_cap = int(self.sa_stats[7])      # more than 8 distinct bytes: re-allocate Occ
                        self.occ = torch.zeros((rows_cap, max(self.ncp, 1)), dtype=torch.int32, device=self.device)
                        continue
                    _lib.check(rc, "index_build")
                    break
                del ws
            self._set_alphabet(totals, row)
            # a2: 8-mer index
            self.kmer_off = None
            self.kmer_pos = None
            self.kmer_count = 0
            if build_kmer:
                self.build_kmer()

    def _set_alphabet(self, totals: np.ndarray, row: np.ndarray) -> None:
        """Byte totals and byte->Occ-row map -> C array, code list and their device copies."""
        torch = self.torch
        self.totals = totals
        counts = np.zeros(256, np.int64)
        cum = 0
        for b in range(256):
            counts[b] = cum
            cum += int(totals[b])
        self.counts = counts
        codes = [b for b in range(256) if totals[b] > 0]
        self.codes = codes
        self.row_of_code = row
        self.occ = self.occ[: max(len(codes), 1)]
        self.d_C = torch.from_numpy(counts).to(self.device)
        self.d_tot = torch.from_numpy(totals).to(self.device)
        self.d_row = torch.from_numpy(row).to(self.device)

    # ------------------------------------------------------------------ persistence
    _FORMAT = 1

    def save(self, path, with_lcp: bool = True) -> None:
        """Writes the index arrays to one uncompressed ``.npz`` (text, SA, BWT, Occ, byte totals,
        row map and, unless declined, LCP) so that later scans of the same contig skip
        construction.  The reference has no counterpart; the arrays are the ones of bwt.py:106-136."""
        arrays = dict(format=np.int64(self._FORMAT), occ_rate=np.int64(self.occ_rate),
                      text=self.text.cpu().numpy(), sa=self.sa.cpu().numpy(), bwt=self.bwt.cpu().numpy(),
                      occ=self.occ.cpu().numpy(), totals=self.totals, row=self.row_of_code)
        if with_lcp:
            arrays["lcp"] = self.lcp.cpu().numpy()
        with open(path, "wb") as fh:
            np.savez(fh, **arrays)

    @classmethod
    def load(cls, path, device=None, build_kmer: bool = True) -> "DeviceIndex":
        """The inverse of ``save``: uploads the stored arrays (no suffix sorting); the 8-mer table
        is rebuilt from the text (one counting pass) and ISA is not kept."""
        torch = _torch()
        with np.load(path) as z:
            if int(z["format"]) != cls._FORMAT:
                raise ValueError(f"{path}: index format {int(z['format'])}, expected {cls._FORMAT}")
            host = {k: z[k] for k in z.files}
        self = cls.__new__(cls)
        self.torch = torch
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.occ_rate = int(host["occ_rate"])
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        self.text, self.sa, self.bwt, self.occ = up(host["text"]), up(host["sa"]), up(host["bwt"]), up(host["occ"])
        n = self.n = int(self.text.numel())
        if not (self.sa.numel() == n and self.bwt.numel() == n and self.sa.dtype == torch.int32):
            raise ValueError(f"{path}: array sizes do not agree")
        self.isa = None
        self._lcp = up(host["lcp"]) if "lcp" in host else None
        self.sa_stats = np.zeros(8, np.int64)
        self.ncp = (n // self.occ_rate + 1 + (1 if n % self.occ_rate else 0)) if n else 0
        row = host["row"].astype(np.int32)
        nrows = int((row >= 0).sum())
        if n and (self.occ.dtype != torch.int32 or tuple(self.occ.shape) != (nrows, self.ncp)
                  or self.bwt.dtype != torch.uint8 or self.text.dtype != torch.uint8
                  or int(host["totals"].astype(np.int64).sum()) != n):
            raise ValueError(f"{path}: Occ table {tuple(self.occ.shape)}/{self.occ.dtype} does not match "
                             f"({nrows}, {self.ncp}) int32 for n={n}, occ_rate={self.occ_rate}")
        self._set_alphabet(host["totals"].astype(np.int64), row)
        self.kmer_off = self.kmer_pos = None
        self.kmer_count = 0
        if build_kmer:
            self.build_kmer()
        return self

    # ------------------------------------------------------------------ a2
    def build_kmer(self):
        torch, L, n = self.torch, _lib.lib(), self.n
        with torch.cuda.device(self.device):
            self.kmer_off = torch.zeros(65537, dtype=torch.int32, device=self.device)
            self.kmer_pos = torch.empty(max(n, 1), dtype=torch.int32, device=self.device)
            cnt = C.c_int64(0)
            if n >= 8:
                wsb = int(L.bwtk_kmer8_workspace_bytes(n))
                ws = torch.empty(wsb, dtype=torch.uint8, device=self.device)
                _lib.check(L.bwtk_kmer8_index(self.text.data_ptr(), n, self.kmer_off.data_ptr(),
                                              self.kmer_pos.data_ptr(), C.addressof(cnt), ws.data_ptr(), wsb,
                                              _lib.stream_ptr()), "kmer8_index")
            self.kmer_count = int(cnt.value)
            self.kmer_pos = self.kmer_pos[: self.kmer_count]
            self._kmer_off_host = None

    def kmer_bucket(self, code: int) -> np.ndarray:
        if self._kmer_off_host is None:
            self._kmer_off_host = self.kmer_off.cpu().numpy()
        lo, hi = int(self._kmer_off_host[code]), int(self._kmer_off_host[code + 1])
        if hi <= lo:
            return np.zeros(0, np.int32)
        return self.kmer_pos[lo:hi].cpu().numpy()

    # ------------------------------------------------------------------ a10
    @property
    def lcp(self):
        if self._lcp is None:
            torch, L, n = self.torch, _lib.lib(), self.n
            with torch.cuda.device(self.device):
                self._lcp = torch.empty(n, dtype=torch.int32, device=self.device)
                if n:
                    wsb = int(L.bwtk_lcp_workspace_bytes(n))
                    ws = torch.empty(wsb, dtype=torch.uint8, device=self.device)
                    _lib.check(L.bwtk_lcp_build(self.text.data_ptr(), self.sa.data_ptr(), n, self._lcp.data_ptr(),
                                                ws.data_ptr(), wsb, _lib.stream_ptr()), "lcp_build")
        return self._lcp


class FMReplica(_FMSearch):
    """Search-only copy of a contig's FM index (BWT bytes, Occ checkpoints, C array) on
    another GPU: what ``sharding.broadcast_fm`` hands to the ranks that did not build the
    index, so that a motif batch can be split across the GPUs of a box."""

    def __init__(self, bwt, occ, totals, row, n: int, occ_rate: int, device=None):
        torch = _torch()
        self.torch = torch
        self.device = torch.device(device if device is not None else bwt.device)
        self.n, self.occ_rate = int(n), int(occ_rate)
        self.bwt = bwt.to(self.device)
        self.occ = occ.to(self.device)
        totals = np.ascontiguousarray(totals, np.int64)
        self.totals = totals
        self.counts = np.concatenate(([0], np.cumsum(totals)[:-1])).astype(np.int64)
        self.row_of_code = np.ascontiguousarray(row, np.int32)
        self.codes = [b for b in range(256) if totals[b] > 0]
        self.d_C = torch.from_numpy(self.counts).to(self.device)
        self.d_tot = torch.from_numpy(totals).to(self.device)
        self.d_row = torch.from_numpy(self.row_of_code).to(self.device)
