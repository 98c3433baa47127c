"""ctypes front-end of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by the product package.
Each wrapper names the reference lines (bwt.py) restated by the C function it
calls.  Parity of this oracle with the reference itself is pinned by
tests/test_oracle_golden.py against tests/golden/*.json (made by
oracle/gen_golden.py from the imported Python reference).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
REC_W = 8


def build(force: bool = False) -> str:
    """Compile oracle.c with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        p = C.c_void_p
        i64 = C.c_int64
        L.orc_suffix_array.argtypes = [p, i64, p]
        L.orc_suffix_array.restype = C.c_int
        L.orc_bwt.argtypes = [p, p, i64, p]
        L.orc_char_counts.argtypes = [p, i64, p, p]
        L.orc_occ_checkpoints.argtypes = [p, i64, i64, C.c_int, p]
        L.orc_occ_checkpoints.restype = i64
        L.orc_backward_search_batch.argtypes = [p, i64, i64, p, i64, p, p, p, i64, p, i64, p, p]
        L.orc_kasai_lcp.argtypes = [p, p, i64, p]
        L.orc_kasai_lcp.restype = C.c_int
        L.orc_kmer8_pairs.argtypes = [p, i64, p, p]
        L.orc_kmer8_pairs.restype = i64
        L.orc_entropy.argtypes = [p, i64]
        L.orc_entropy.restype = C.c_double
        L.orc_smallest_period.argtypes = [p, i64]
        L.orc_smallest_period.restype = i64
        L.orc_mm_budget.argtypes = [i64, i64]
        L.orc_mm_budget.restype = i64
        L.orc_tier1_scan.argtypes = [p, i64, C.c_int, C.c_int, C.c_int, C.c_double, p, i64]
        L.orc_tier1_scan.restype = i64
        L.orc_strict_scan.argtypes = [p, i64, i64, i64, i64, i64, p, i64]
        L.orc_strict_scan.restype = i64
        L.orc_lcp_plateaus.argtypes = [p, i64, p, p, i64, i64, i64, i64, p, i64, p]
        L.orc_lcp_plateaus.restype = i64
        L.orc_consensus.argtypes = [p, i64, i64, i64, i64, p, p, p]
        L.orc_consensus.restype = i64
        L.orc_extend_with_mismatches.argtypes = [p, i64, i64, i64, C.c_int, p]
        L.orc_extend_tandem_fm.argtypes = [p, i64, i64, i64, p]
        L.orc_period_scan.argtypes = [p, i64, i64, i64, C.c_int, i64, i64, C.c_double, p, p, i64, p]
        L.orc_period_scan.restype = i64
        _lib = L
    return _lib


def _u8(text) -> np.ndarray:
    if isinstance(text, str):
        text = text.encode("utf-8")
    if isinstance(text, (bytes, bytearray)):
        return np.frombuffer(bytes(text), dtype=np.uint8).copy()
    return np.ascontiguousarray(text, dtype=np.uint8)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- index (a2-a10)
def suffix_array(text) -> np.ndarray:
    """bwt.py:212-264 (prefix doubling, byte order)."""
    t = _u8(text)
    sa = np.empty(t.size, dtype=np.int32)
    if lib().orc_suffix_array(_ptr(t), t.size, _ptr(sa)) != 0:
        raise MemoryError("oracle SA")
    return sa


def bwt_array(text, sa: np.ndarray) -> np.ndarray:
    """bwt.py:266-274."""
    t = _u8(text)
    out = np.empty(t.size, dtype=np.uint8)
    lib().orc_bwt(_ptr(t), _ptr(np.ascontiguousarray(sa, np.int32)), t.size, _ptr(out))
    return out


def char_counts(text) -> Tuple[np.ndarray, np.ndarray]:
    """bwt.py:276-286 -> (totals[256], C[256]) int64."""
    t = _u8(text)
    totals = np.zeros(256, np.int64)
    counts = np.zeros(256, np.int64)
    lib().orc_char_counts(_ptr(t), t.size, _ptr(totals), _ptr(counts))
    return totals, counts


def occ_checkpoints(bwt: np.ndarray, k: int = 128) -> Dict[int, np.ndarray]:
    """bwt.py:288-326 -> {code: int32[ncp]} for every code present."""
    b = np.ascontiguousarray(bwt, np.uint8)
    n = b.size
    if n == 0:
        return {}
    ncp = n // k + 1 + (1 if n % k else 0)
    out = {}
    for code in np.unique(b).tolist():
        cp = np.empty(ncp, np.int32)
        m = lib().orc_occ_checkpoints(_ptr(b), n, k, int(code), _ptr(cp))
        assert m == ncp
        out[int(code)] = cp
    return out


def kasai_lcp(text, sa: np.ndarray) -> np.ndarray:
    """bwt.py:55-72."""
    t = _u8(text)
    lcp = np.empty(t.size, np.int32)
    lib().orc_kasai_lcp(_ptr(t), _ptr(np.ascontiguousarray(sa, np.int32)), t.size, _ptr(lcp))
    return lcp


def kmer8_pairs(text) -> Tuple[np.ndarray, np.ndarray]:
    """bwt.py:138-171 -> (codes, positions) in recording order."""
    t = _u8(text)
    codes = np.empty(max(t.size, 1), np.int32)
    pos = np.empty(max(t.size, 1), np.int32)
    m = lib().orc_kmer8_pairs(_ptr(t), t.size, _ptr(codes), _ptr(pos))
    return codes[:m].copy(), pos[:m].copy()


def kmer8_table(text) -> Dict[int, List[int]]:
    codes, pos = kmer8_pairs(text)
    order = np.argsort(codes, kind="stable")
    cs, ps = codes[order], pos[order]
    out: Dict[int, List[int]] = {}
    if cs.size:
        cuts = np.flatnonzero(np.diff(cs)) + 1
        starts = np.concatenate(([0], cuts))
        ends = np.concatenate((cuts, [cs.size]))
        for s, e in zip(starts.tolist(), ends.tolist()):
            out[int(cs[s])] = ps[s:e].tolist()
    return out


class OracleIndex:
    """All index arrays of BWTCore.__init__ (bwt.py:106-136) for one text."""

    def __init__(self, text, occ_rate: int = 128):
        self.text = _u8(text)
        self.n = int(self.text.size)
        self.k = occ_rate
        self.sa = suffix_array(self.text)
        self.bwt = bwt_array(self.text, self.sa)
        self.totals, self.counts = char_counts(self.text)
        self.occ = occ_checkpoints(self.bwt, occ_rate)
        self.ncp = (self.n // occ_rate + 1 + (1 if self.n % occ_rate else 0)) if self.n else 0
        self._occ_mat = None

    def lcp(self) -> np.ndarray:
        return kasai_lcp(self.text, self.sa)

    def occ_matrix(self) -> np.ndarray:
        if self._occ_mat is None:
            m = np.zeros((256, max(self.ncp, 1)), np.int32)
            for code, cp in self.occ.items():
                m[code, : cp.size] = cp
            self._occ_mat = m
        return self._occ_mat

    def backward_search_batch(self, pats: np.ndarray, lens: np.ndarray):
        """bwt.py:359-389 for a [nq, stride] uint8 pattern matrix."""
        pats = np.ascontiguousarray(pats, np.uint8)
        lens = np.ascontiguousarray(lens, np.int32)
        nq = lens.size
        stride = pats.shape[1] if pats.ndim == 2 else 0
        sp = np.empty(nq, np.int64)
        ep = np.empty(nq, np.int64)
        occ = self.occ_matrix()
        lib().orc_backward_search_batch(
            _ptr(self.bwt), self.n, self.k, _ptr(occ), occ.shape[1], _ptr(self.totals),
            _ptr(self.counts), _ptr(pats), stride, _ptr(lens), nq, _ptr(sp), _ptr(ep))
        return sp, ep

    def backward_search(self, pattern) -> Tuple[int, int]:
        p = _u8(pattern)
        sp, ep = self.backward_search_batch(p.reshape(1, -1) if p.size else np.zeros((1, 1), np.uint8),
                                            np.array([p.size], np.int32))
        return int(sp[0]), int(ep[0])


# ---------------------------------------------------------------- helpers
def entropy(seq) -> float:
    s = _u8(seq)
    return float(lib().orc_entropy(_ptr(s), s.size))


def smallest_period(seq) -> int:
    s = _u8(seq)
    return int(lib().orc_smallest_period(_ptr(s), s.size))


def mm_budget(motif_len: int, n_copies: int) -> int:
    return int(lib().orc_mm_budget(motif_len, n_copies))


def _records(fn, cap0: int = 1 << 16):
    cap = cap0
    while True:
        rec = np.zeros((cap, REC_W), np.int32)
        m = fn(rec, cap)
        if m <= cap:
            return rec[:m].copy()
        cap = int(m)


# ---------------------------------------------------------------- detectors
def tier1_scan(text_arr, max_motif_length: int = 9, min_copies: int = 3,
               min_array_length: int = 6, min_entropy: float = 1.0) -> np.ndarray:
    """bwt.py:1426-1531 -> rows (start,end,motif_len,copies,...)."""
    t = _u8(text_arr)
    return _records(lambda rec, cap: lib().orc_tier1_scan(
        _ptr(t), t.size, max_motif_length, min_copies, min_array_length, min_entropy, _ptr(rec), cap))


def strict_scan(text_arr, min_unit_len: int = 20, max_unit_len: int = 120, max_mismatch: int = 2,
                min_copies: int = 3) -> np.ndarray:
    """bwt.py:1891-2001 -> rows (start,end,primitive_period,copies,0,0,unit_len,0)."""
    t = _u8(text_arr)
    return _records(lambda rec, cap: lib().orc_strict_scan(
        _ptr(t), t.size, min_unit_len, max_unit_len, max_mismatch, min_copies, _ptr(rec), cap))


def lcp_plateaus(text_arr, sa, lcp, min_period: int = 1, max_period: int = 1000,
                 min_copies: int = 3):
    """bwt.py:2118-2145 + 2500-2560 -> (rows, threshold)."""
    t = _u8(text_arr)
    sa = np.ascontiguousarray(sa, np.int32)
    lcp = np.ascontiguousarray(lcp, np.int32)
    thr = C.c_int64(-1)
    rows = _records(lambda rec, cap: lib().orc_lcp_plateaus(
        _ptr(t), t.size, _ptr(sa), _ptr(lcp), lcp.size, min_period, max_period, min_copies,
        _ptr(rec), cap, C.byref(thr)))
    return rows, int(thr.value)


def consensus(text_arr, start: int, motif_len: int, n_copies: int):
    """bwt.py:1207-1256 -> (consensus uint8[], mismatch_rate, max_mm_per_copy)."""
    t = _u8(text_arr)
    cons = np.zeros(max(motif_len, 1), np.uint8)
    tmm = C.c_int64(0)
    mmm = C.c_int64(0)
    used = lib().orc_consensus(_ptr(t), t.size, start, motif_len, n_copies, _ptr(cons),
                               C.byref(tmm), C.byref(mmm))
    if used == 0:
        return np.array([], np.uint8), 0.0, 0
    total = used * motif_len
    return cons[:motif_len].copy(), (tmm.value / total if total > 0 else 0.0), int(mmm.value)


def extend_with_mismatches(s_arr, start_pos: int, period: int, n: int, allow_mismatches: bool = True):
    """bwt.py:2392-2498 -> (array_start, array_end, copies, full_start, full_end)."""
    t = _u8(s_arr)
    out = np.zeros(5, np.int64)
    lib().orc_extend_with_mismatches(_ptr(t), start_pos, period, n, int(allow_mismatches), _ptr(out))
    return tuple(int(x) for x in out)


def extend_tandem_fm(text_arr, seed_pos: int, motif_len: int):
    """bwt.py:2697-2805 -> (start, end, copies); n = text_arr.size."""
    t = _u8(text_arr)
    out = np.zeros(3, np.int64)
    lib().orc_extend_tandem_fm(_ptr(t), t.size, seed_pos, motif_len, _ptr(out))
    return tuple(int(x) for x in out)


def period_scan(text_arr, min_period: int = 1, max_period: int = 1000, allow_mismatches: bool = True,
                min_copies: int = 3, min_array_length: int = 6, min_entropy: float = 1.0,
                tier1_mask: Optional[np.ndarray] = None):
    """bwt.py:2177-2390 with the clock frozen -> (raw rows before the
    (start,end,canonical) dedup, iterations)."""
    t = _u8(text_arr)
    mask = None if tier1_mask is None else np.ascontiguousarray(tier1_mask, np.uint8)
    it = C.c_int64(0)
    rows = _records(lambda rec, cap: lib().orc_period_scan(
        _ptr(t), t.size, min_period, max_period, int(allow_mismatches), min_copies,
        min_array_length, min_entropy, _ptr(mask), _ptr(rec), cap, C.byref(it)))
    return rows, int(it.value)


# ---------------------------------------------------------------- a14 (sequential restatement)
_COMP = {"A": "T", "T": "A", "C": "G", "G": "C", "N": "N"}
_BITS = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 0}


def _canonical(m: str) -> str:
    return min(m[i:] + m[:i] for i in range(len(m))) if m else m


def _primitive(m: str) -> bool:
    n = len(m)
    return not any(n % p == 0 and m[:p] * (n // p) == m for p in range(1, n))


def enumerate_motifs(k: int):
    """bwt.py:1368-1381: canonical primitive motifs in lexicographic order."""
    import itertools

    for tup in itertools.product("ACGT", repeat=k):
        m = "".join(tup)
        if _canonical(m) == m and _primitive(m):
            yield m


def short_imperfect_arrays(text, tier1_seen=(), min_period: int = 1, max_short_motif: int = 9,
                           min_copies: int = 3, min_array_length: int = 6, min_entropy: float = 1.0):
    """bwt.py:2027-2095 + 2562-2695, walked seed by seed exactly as written.
    Returns accepted arrays (start, end, motif_len, copies, consensus, total_mm, max_mm, used)."""
    t = _u8(text)
    n = int(t.size)
    if n > 1_000_000:
        return []
    ix = OracleIndex(t)
    table = kmer8_table(t)
    seen = set(tuple(x) for x in tier1_seen)
    out = []

    def locate(pat: str):
        sp, ep = ix.backward_search(pat)
        return [] if sp == -1 else sorted(int(x) for x in ix.sa[sp:ep + 1])

    def kmer_positions(kmer: str):
        if len(kmer) > 8 or not table:
            return locate(kmer)
        w = 0
        for b in kmer.upper():
            if b not in _BITS:
                return []
            w = (w << 2) | _BITS[b]
        return table.get(w, [])

    for k in range(min_period, min(max_short_motif + 1, 10)):
        motifs = [m for m in enumerate_motifs(k) if not (entropy(m) < min_entropy)]
        rot_sets = []
        for motif in motifs:
            rc = "".join(_COMP.get(b, b) for b in reversed(motif))
            rot_sets.append(sorted(set([motif[i:] + motif[:i] for i in range(k)] + [rc[i:] + rc[:i] for i in range(k)])))
        fm_hits = {}
        if k > 8 and motifs:
            # the same backward searches the reference issues one by one (bwt.py:2078-2079), batched
            flat = [r for rs in rot_sets for r in rs]
            mat = np.frombuffer("".join(flat).encode(), np.uint8).reshape(len(flat), k)
            bsp, bep = ix.backward_search_batch(mat, np.full(len(flat), k, np.int32))
            for r, a, b in zip(flat, bsp.tolist(), bep.tolist()):
                if a != -1:
                    fm_hits[r] = (a, b)
        for motif, rots in zip(motifs, rot_sets):
            allpos = []
            for r in rots:
                if k <= 8:
                    allpos.extend(kmer_positions(r))
                elif r in fm_hits:
                    a, b = fm_hits[r]
                    allpos.extend(int(x) for x in ix.sa[a:b + 1])
            positions = sorted(set(allpos))
            if len(positions) < min_copies:
                continue
            motif_len = k
            for seed in positions:
                if any(a <= seed < b for a, b in seen):
                    continue
                if seed + motif_len > n:
                    continue
                best = None
                for shift in range(min(motif_len, seed + 1)):
                    cand = seed - shift
                    if cand < 0 or cand + motif_len > n:
                        continue
                    if any(a <= cand < b for a, b in seen):
                        continue
                    s, e, c = extend_tandem_fm(t, cand, motif_len)
                    if not (s <= seed < e):
                        continue
                    if best is None or c > best[2] or (c == best[2] and s < best[0]):
                        best = (s, e, c)
                if best is None:
                    continue
                start, end, copies = best
                if not (copies >= min_copies and end - start >= min_array_length):
                    continue
                cons, rate, mx = consensus(t, start, motif_len, copies)
                if cons.size == 0:
                    continue
                prim = smallest_period(cons)
                if prim < cons.size:
                    motif_len = prim
                    copies = max(1, (end - start) // motif_len)
                    end = start + copies * motif_len
                    cons, rate, mx = consensus(t, start, motif_len, copies)
                    if cons.size == 0:
                        continue
                if start > 0 and t[start - 1] == cons[motif_len - 1]:
                    continue
                if end < n and t[end] == cons[0]:
                    continue
                if (1.0 - rate) * 100.0 < (90.0 if motif_len <= 6 else 85.0):
                    continue
                used = copies
                total_mm = int(round(rate * used * motif_len))
                out.append((start, end, motif_len, copies, cons, total_mm, mx, used))
                seen.add((start, end))
    return out
