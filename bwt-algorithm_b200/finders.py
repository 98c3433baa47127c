"""Tier 1 / Tier 2 detectors with the reference's class surface.

``Tier1STRFinder`` (bwt.py:1387-1538) and ``Tier2LCPFinder`` (bwt.py:1865-2825)
keep their constructors, tunable attributes and method names; the scans run in
the CUDA kernels (detect.py -> libbwtk.so) and return integer rows, which are
turned into ``TandemRepeat`` records here with the reference's float formulas.
"""
from __future__ import annotations

import gc
from typing import Dict, List, Optional, Set, Tuple

import numpy as np

from . import detect
from .core import BWTCore
from .motifs import MotifUtils
from .records import TandemRepeat


def _budget(motif_len: int, n_copies: int) -> int:
    """Mismatch budget of a whole array (bwt.py:1402-1424 == 2003-2025)."""
    if motif_len == 1:
        return 0
    frac = 0.05 if motif_len <= 6 else 0.08
    return max(1, int(np.ceil(frac * (motif_len * n_copies))))


def _decode(arr: np.ndarray, a: int, b: int) -> str:
    return arr[a:b].tobytes().decode("ascii", errors="replace")


def _device_text_of(bwt_core):
    """Device text of a BWTCore when its host text is still the one it was built
    from; otherwise the host array (uploaded by the launcher)."""
    dev = getattr(bwt_core, "_dev", None)
    if dev is not None and dev.n == int(bwt_core.text_arr.size):
        return dev.text
    return bwt_core.text_arr


# ---------------------------------------------------------------------------
# integer rows (kernels) -> TandemRepeat records (host floats, bwt.py formulas)
# ---------------------------------------------------------------------------
def tier1_records(text_arr: np.ndarray, rows: np.ndarray, chromosome: str) -> List[TandemRepeat]:
    """Rows (start,end,motif_len,copies) -> Tier 1 records (bwt.py:1493-1522)."""
    out: List[TandemRepeat] = []
    for start, end, m, copies in rows[:, :4].tolist():
        motif = _decode(text_arr, start, start + m)
        pm, pi, score, comp, ent, _ = MotifUtils.calculate_trf_statistics(text_arr, start, end, motif, copies, 0.0)
        out.append(TandemRepeat(
            chrom=chromosome, start=start, end=end, motif=motif, copies=float(copies), length=end - start,
            tier=1, confidence=1.0, consensus_motif=motif, mismatch_rate=0.0, max_mismatches_per_copy=0,
            n_copies_evaluated=copies, strand="+", percent_matches=pm, percent_indels=pi, score=score,
            composition=comp, entropy=ent, actual_sequence=_decode(text_arr, start, end), variations=None))
    return out


class StrictRecordMaker:
    """(start, end, primitive period, copies) -> the record ``find_long_unit_repeats_strict`` builds for
    an exact array (bwt.py:1951-1996).  ``MotifUtils.calculate_trf_statistics(text_arr, start, end, motif,
    count, 0.0)`` is unrolled with its per-motif parts (composition, entropy) computed once per distinct
    motif and its mismatch-free arithmetic written out: a chr21-sized contig yields millions of calls."""

    def __init__(self, text_arr: np.ndarray, chromosome: str, max_mismatch: int, text_bytes: Optional[bytes] = None):
        self.text = text_bytes if text_bytes is not None else text_arr.tobytes()
        self.size = len(self.text)
        self.chromosome = chromosome
        self.per_motif: Dict[bytes, tuple] = {}
        self.pm = (1.0 - 0.0) * 100.0
        self.mm_per_copy = 0 if self.pm >= 99.9 else max_mismatch

    def __call__(self, start: int, end: int, prim: int, count: int) -> TandemRepeat:
        text = self.text
        raw = text[start:start + prim]
        known = self.per_motif.get(raw)
        if known is None:
            motif = raw.decode("ascii", errors="replace")
            known = self.per_motif[raw] = (motif, MotifUtils.calculate_composition(motif),
                                           MotifUtils.calculate_entropy(motif))
        motif, comp, ent = known
        length = end - start
        actual = text[start:end].decode("ascii", errors="replace") if end <= self.size else motif * int(count)
        score = max(0, int((length * (1.0 - 0.0) * 2) - (length * 0.0 * 7)))
        # positional, in the field order of records.TandemRepeat (bwt.py:429-452)
        return TandemRepeat(self.chromosome, start, end, motif, float(count), length, 2, 0.95, motif, 0.0,
                            self.mm_per_copy, count, "+", self.pm, 0.0, score, dict(comp), ent, actual, None)


def strict_records(text_arr: np.ndarray, rows: np.ndarray, chromosome: str, max_mismatch: int,
                   text_bytes: Optional[bytes] = None) -> List[TandemRepeat]:
    """Rows (start,end,primitive_period,copies) -> strict-scan records (bwt.py:1951-1996).
    `text_bytes`: the bytes of `text_arr` when the caller already holds them."""
    if len(rows) == 0:
        return []
    make = StrictRecordMaker(text_arr, chromosome, max_mismatch, text_bytes)
    # millions of records, each owning a dict: the cyclic collector would rescan the growing list
    # again and again (it more than doubles the loop's time), and nothing here can form a cycle
    gc_was_on = gc.isenabled()
    gc.disable()
    try:
        return [make(start, end, prim, count) for start, end, prim, count in rows[:, :4].tolist()]
    finally:
        if gc_was_on:
            gc.enable()


def plateau_records(text_arr: np.ndarray, rows: np.ndarray, chromosome: str) -> List[TandemRepeat]:
    """Rows (start,end,period,copies) -> LCP-plateau records (bwt.py:2536-2547);
    coordinates are numpy int32 scalars as in the reference."""
    out: List[TandemRepeat] = []
    for start, end, period, copies in rows[:, :4].tolist():
        out.append(TandemRepeat(chrom=chromosome, start=np.int32(start), end=np.int32(end),
                                motif=_decode(text_arr, start, start + period), copies=copies,
                                length=copies * period, tier=2, confidence=0.9))
    return out


def period_records(s_arr: np.ndarray, rows: np.ndarray, cons_list, chromosome: str) -> List[TandemRepeat]:
    """Period-scan rows + consensus bytes -> records, with the (start,end,canonical)
    dedup of bwt.py:2343-2385."""
    out: List[TandemRepeat] = []
    seen: Set[Tuple[int, int, str]] = set()
    for row, cons_arr in zip(rows.tolist(), cons_list):
        a_start, a_end, p_eff, copies_full, total_mm, max_mm, _cons_start, used = row
        cons = cons_arr.tobytes().decode("ascii", errors="replace")
        cells = used * p_eff
        mm_rate = total_mm / cells if cells > 0 else 0.0
        canonical, strand = MotifUtils.get_canonical_motif_stranded(cons)
        key = (a_start, a_end, canonical)
        if key in seen:
            continue
        seen.add(key)
        pm, pi, score, comp, ent, actual = MotifUtils.calculate_trf_statistics(
            s_arr, a_start, a_end, cons, copies_full, mm_rate)
        notes = MotifUtils.summarize_variations_array(s_arr, a_start, a_end, p_eff, cons_arr)
        out.append(TandemRepeat(
            chrom=chromosome, start=a_start, end=a_end, motif=cons, copies=float(copies_full),
            length=a_end - a_start, tier=2, confidence=max(0.5, 0.95 - mm_rate), consensus_motif=cons,
            mismatch_rate=mm_rate, max_mismatches_per_copy=max_mm, n_copies_evaluated=copies_full,
            strand=strand, percent_matches=pm, percent_indels=pi, score=score, composition=comp, entropy=ent,
            actual_sequence=actual, variations=notes if notes else None))
    return out


class Tier1STRFinder:
    """Perfect short tandem repeats, motif length 9..1 (bwt.py:1387-1538)."""

    def __init__(self, text_arr: np.ndarray, max_motif_length: int = 9, show_progress: bool = False):
        self.text_arr = text_arr
        self.max_motif_length = max_motif_length
        self.min_copies = 3
        self.min_array_length = 6
        self.min_entropy = 1.0
        self.show_progress = show_progress

    def _get_max_mismatches_for_array(self, motif_len: int, n_copies: int) -> int:
        return _budget(motif_len, n_copies)

    def _find_simple_tandems_kmer(self, chromosome: str) -> List[TandemRepeat]:
        text_arr = self.text_arr
        n = int(text_arr.size)
        if self.show_progress and n > 10_000_000:
            print(f"  [{chromosome}] Large sequence ({n:,} bp) - using fast sampling mode (step=50)")
        rows = detect.tier1_rows(text_arr, self.max_motif_length, self.min_copies, self.min_array_length,
                                 self.min_entropy)
        return tier1_records(text_arr, rows, chromosome)

    def find_strs(self, chromosome: str) -> List[TandemRepeat]:
        return self._find_simple_tandems_kmer(chromosome)


class Tier2LCPFinder:
    """FM-index / LCP based detectors for all motif lengths (bwt.py:1865-2825)."""

    def __init__(self, bwt_core: BWTCore, min_period: int = 1, max_period: int = 1000,
                 max_short_motif: int = 9, allow_mismatches: bool = True, show_progress: bool = False):
        self.bwt = bwt_core
        self.min_period = min_period
        self.max_period = max_period
        self.max_short_motif = max_short_motif
        self.min_copies = 3
        self.min_array_length = 6
        self.min_entropy = 1.0
        self.allow_mismatches = allow_mismatches
        self.show_progress = show_progress
        self.period_step = 1

    # ---- small helpers kept for API compatibility ---------------------------
    def _hamming_distance(self, arr1: np.ndarray, arr2: np.ndarray) -> int:
        return int(np.sum(arr1 != arr2))

    def _get_max_mismatches_for_array(self, motif_len: int, n_copies: int) -> int:
        return _budget(motif_len, n_copies)

    def _smallest_period(self, s: str) -> int:
        """Smallest p dividing len(s) with s == s[:p]*(len/p) (bwt.py:2147-2159)."""
        return MotifUtils.smallest_period_str(s) if s else 0

    def _smallest_period_codes(self, arr: np.ndarray) -> int:
        n = int(arr.size)
        if n == 0:
            return 0
        for p in range(1, n // 2 + 1):
            if n % p == 0 and np.array_equal(arr[p:], arr[:-p]):
                return p
        return n

    def _validate_periodicity_arr(self, text_arr: np.ndarray, motif_arr: np.ndarray, period: int) -> bool:
        m = text_arr.size
        if m < 2 * period:
            return False
        same = np.count_nonzero(text_arr == motif_arr[np.arange(m, dtype=np.int32) % period])
        return bool((same / m if m > 0 else 0.0) >= 0.8)

    # ---- a13: strict adjacency (the CLI's detector) ---------------------------
    def find_long_unit_repeats_strict(self, chromosome: str, min_unit_len: int = 20, max_unit_len: int = 120,
                                      max_mismatch: int = 2, min_copies: int = 3) -> List[TandemRepeat]:
        text_arr = self.bwt.text_arr
        rows = detect.strict_rows(_device_text_of(self.bwt), min_unit_len, max_unit_len, max_mismatch, min_copies)
        return strict_records(text_arr, rows, chromosome, max_mismatch)

    # ---- a10 / a15: LCP array and plateaus --------------------------------------
    def _compute_lcp_array(self) -> np.ndarray:
        if self.bwt.n == 0:
            return np.zeros(0, dtype=np.int32)
        return self.bwt.device_index.lcp.cpu().numpy()

    def _detect_lcp_plateaus(self, lcp_array: np.ndarray, chromosome: str) -> List[TandemRepeat]:
        if len(lcp_array) == 0:
            return []
        text_arr = self.bwt.text_arr
        dev = getattr(self.bwt, "_dev", None)
        sa = dev.sa if dev is not None else self.bwt.suffix_array
        lcp = np.ascontiguousarray(lcp_array, np.int32)  # the caller's array is authoritative
        rows, _thr = detect.plateau_rows(_device_text_of(self.bwt), sa, lcp, self.min_period, self.max_period,
                                         self.min_copies)
        return plateau_records(text_arr, rows, chromosome)

    def _analyze_sa_interval_for_tandems(self, start_idx: int, end_idx: int, period: int,
                                         chromosome: str) -> List[TandemRepeat]:
        """One SA interval (bwt.py:2500-2549); the batched form is _detect_lcp_plateaus."""
        text_arr = self.bwt.text_arr
        pos = np.sort(np.asarray(self.bwt.suffix_array[start_idx:end_idx]).astype(np.int64))
        out: List[TandemRepeat] = []
        for a in range(pos.size):
            copies, start = 1, int(pos[a])
            b = a + 1
            while b < pos.size and pos[b] == start + copies * period:
                copies += 1
                b += 1
            if copies >= self.min_copies and start + period <= int(text_arr.size):
                total = copies * period
                if self._validate_periodicity_arr(text_arr[start:start + total], text_arr[start:start + period], period):
                    out.append(TandemRepeat(chrom=chromosome, start=np.int32(start), end=np.int32(start + total),
                                            motif=_decode(text_arr, start, start + period), copies=copies,
                                            length=total, tier=2, confidence=0.9))
        return out

    # ---- a16: period scan ---------------------------------------------------------
    def find_long_repeats(self, chromosome: str,
                          tier1_seen: Optional[Set[Tuple[int, int]]] = None) -> List[TandemRepeat]:
        return self._find_repeats_simple(chromosome, tier1_seen or set())

    def _extend_with_mismatches(self, s_arr: np.ndarray, start_pos: int, period: int, n: int,
                                allow_mismatches: bool = True) -> Tuple[int, int, int, int, int]:
        row = detect.extend_batch(s_arr, n, [start_pos], [period], [1 if allow_mismatches else 0], mode=0)[0]
        return tuple(int(x) for x in row[:5])

    def _find_repeats_simple(self, chromosome: str, tier1_seen: Set[Tuple[int, int]]) -> List[TandemRepeat]:
        s_arr = self.bwt.text_arr
        n = int(s_arr.size)
        if n > 0 and s_arr[n - 1] == 36:
            n -= 1
        mask = None
        if tier1_seen:
            mask = np.zeros(n, dtype=np.uint8)
            for a, b in tier1_seen:
                mask[a:min(b, n)] = 1
        text_dev = _device_text_of(self.bwt)
        rows, _iters = detect.period_scan_rows(text_dev, self.min_period, self.max_period, self.allow_mismatches,
                                               self.min_copies, self.min_array_length, self.min_entropy, mask)
        if len(rows) == 0:
            return []
        cons_list, _mm = detect.consensus_batch(text_dev, rows[:, 6], rows[:, 2], rows[:, 3])
        return period_records(s_arr, rows, cons_list, chromosome)

    # ---- a14: FM / k-mer seeded seed-and-extend ------------------------------------
    def _extend_tandem_fm(self, text_arr: np.ndarray, seed_pos: int, motif: str, motif_len: int,
                          max_mismatches: int) -> Tuple[int, int, int]:
        row = detect.extend_batch(text_arr, int(text_arr.size), [seed_pos], [motif_len], None, mode=1)[0]
        return int(row[0]), int(row[1]), int(row[2])

    def _is_maximal_fm(self, start: int, end: int, consensus: np.ndarray, motif_len: int, max_mm: int) -> bool:
        text_arr = self.bwt.text_arr
        if start > 0 and text_arr[start - 1] == consensus[motif_len - 1]:
            return False
        if end < text_arr.size and text_arr[end] == consensus[0]:
            return False
        return True

    def find_short_imperfect_repeats(self, chromosome: str,
                                     tier1_seen: Set[Tuple[int, int]]) -> List[TandemRepeat]:
        from . import seeded

        return seeded.find_short_imperfect_repeats(self, chromosome, tier1_seen)

    def _find_tandems_fm_with_mismatches(self, positions: List[int], motif: str, chromosome: str, motif_len: int,
                                         seen_regions: Set[Tuple[int, int]]) -> List[TandemRepeat]:
        from . import seeded

        return seeded.find_tandems_from_seeds(self, positions, motif, chromosome, motif_len, seen_regions)
