"""``MotifUtils``: motif helpers of the ``bwt`` module surface.

Host-side, per-record arithmetic (strings of a few bases, float64 derived
fields).  The array-level hot parts of the detectors -- Hamming/transversion
counting, majority-vote consensus, run detection -- live in the CUDA kernels;
what stays here is what the reference also evaluates once per record, written
with the same expression order so derived floats match bit for bit
(SURVEY.md A.4-A.9).  ``build_consensus_motif_array`` runs on the GPU
(``bwtk_consensus_batch``).
"""
from __future__ import annotations

import functools
import math
import operator
import threading
from collections import Counter
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

from .records import AlignmentResult, RepeatAlignmentSummary

_COMPLEMENT = {"A": "T", "T": "A", "C": "G", "G": "C", "N": "N"}
_COUNT_OF = operator.itemgetter(1)
_TRANSITIONS = {("A", "G"), ("G", "A"), ("C", "T"), ("T", "C")}


def _rotations(s: str) -> List[str]:
    return [s[i:] + s[:i] for i in range(len(s))]


def _canonical_stranded(motif) -> Tuple[str, str]:
    """bwt.py:700-727: the smaller of the least rotation of the motif and of its reverse complement."""
    fwd = min(_rotations(motif))
    rev = min(_rotations("".join(_COMPLEMENT.get(b, b) for b in reversed(motif))))
    return (fwd, "+") if fwd <= rev else (rev, "-")


_canonical_stranded_cached = functools.lru_cache(maxsize=1 << 16)(_canonical_stranded)


class MotifUtils:
    """Static helpers; names and results follow bwt.py:675-1381."""

    # ---- canonical forms (bwt.py:678-727) --------------------------------
    @staticmethod
    def get_canonical_motif(motif: str) -> str:
        return min(_rotations(motif)) if motif else motif

    @staticmethod
    def reverse_complement(seq: str) -> str:
        return "".join(_COMPLEMENT.get(b, b) for b in reversed(seq))

    @staticmethod
    def get_canonical_motif_stranded(motif: str) -> Tuple[str, str]:
        if not motif:
            return motif, "+"
        if type(motif) is str:          # the post-processing chain asks for the same few motifs millions of times
            return _canonical_stranded_cached(motif)
        return _canonical_stranded(motif)

    @staticmethod
    def is_primitive_motif(motif: str) -> bool:
        n = len(motif)
        return not any(n % p == 0 and motif[:p] * (n // p) == motif for p in range(1, n))

    # ---- entropy / distances (bwt.py:729-800) ------------------------------
    @staticmethod
    def calculate_entropy(seq: str) -> float:
        """Shannon entropy in bits; symbols are accumulated in first-seen order
        with numpy's log2, as the reference does (bwt.py:729-745)."""
        if not seq:
            return 0.0
        n = len(seq)
        h = 0.0
        for c in Counter(seq).values():
            if c > 0:
                p = c / n
                h -= p * np.log2(p)
        return h

    @staticmethod
    def is_transition(base1: str, base2: str) -> bool:
        return base1 == base2 or (base1, base2) in _TRANSITIONS

    @staticmethod
    def hamming_distance(s1: str, s2: str) -> int:
        if len(s1) != len(s2):
            return max(len(s1), len(s2))
        return sum(a != b for a, b in zip(s1, s2))

    @staticmethod
    def hamming_distance_array(arr1: np.ndarray, arr2: np.ndarray) -> int:
        if arr1.size != arr2.size:
            return max(arr1.size, arr2.size)
        return int(np.count_nonzero(arr1 != arr2))

    @staticmethod
    def count_transversions_array(arr1: np.ndarray, arr2: np.ndarray) -> int:
        """Mismatches that are not A<->G / C<->T; bytes outside 65..84 count as 'N'
        (bwt.py:780-800)."""
        if arr1.size != arr2.size:
            return max(arr1.size, arr2.size)
        a = np.asarray(arr1, np.uint8)
        b = np.asarray(arr2, np.uint8)
        a = np.where((a >= 65) & (a <= 84), a, 78)
        b = np.where((b >= 65) & (b <= 84), b, 78)
        differ = np.asarray(arr1) != np.asarray(arr2)
        ts = ((a == 65) & (b == 71)) | ((a == 71) & (b == 65)) | ((a == 67) & (b == 84)) | ((a == 84) & (b == 67))
        return int(np.count_nonzero(differ & ~ts & (a != b)))

    @staticmethod
    def edit_distance(a: str, b: str) -> int:
        if not a:
            return len(b)
        if not b:
            return len(a)
        row = list(range(len(b) + 1))
        for i, ca in enumerate(a, 1):
            nxt = [i]
            for j, cb in enumerate(b, 1):
                nxt.append(min(row[j] + 1, nxt[j - 1] + 1, row[j - 1] + (ca != cb)))
            row = nxt
        return row[-1]

    # ---- per-copy alignment (bwt.py:828-1102) ------------------------------
    _EXACT_COPY_FAST_PATH = True   # tests switch it off to compare with the DP-only walk

    @staticmethod
    def _align_unit_to_window(motif: str, window: str, max_indel: int,
                              mismatch_tolerance: int) -> Optional[AlignmentResult]:
        """Banded unit-cost alignment of `motif` against a prefix of `window`.

        Tie order substitution > deletion > insertion (strict <); the end column
        is the first minimum in [m-max_indel, m+max_indel]; rejected when
        substitutions exceed the tolerance or an indel total exceeds max_indel.

        A pure function of its arguments; short motifs (the merge stage asks about the same few
        homopolymer / dinucleotide windows hundreds of thousands of times per contig) are memoised.
        The result is shared between callers and must not be mutated.
        """
        if len(motif) <= 12:
            return _align_unit_cached(motif, window, max_indel, mismatch_tolerance)
        return MotifUtils._align_unit_to_window_dp(motif, window, max_indel, mismatch_tolerance)

    @staticmethod
    def _align_unit_to_window_dp(motif: str, window: str, max_indel: int,
                                 mismatch_tolerance: int) -> Optional[AlignmentResult]:
        m, n = len(motif), len(window)
        if m == 0 or n == 0:
            return None
        max_indel = max(0, max_indel)
        mismatch_tolerance = max(0, mismatch_tolerance)
        lo, hi = max(0, m - max_indel), min(n, m + max_indel)
        if lo > hi:
            return None
        big = m + n + 10
        cost = [[big] * (n + 1) for _ in range(m + 1)]
        move = [[0] * (n + 1) for _ in range(m + 1)]  # 1 match/sub, 2 deletion, 3 insertion
        for j in range(n + 1):
            cost[0][j] = j
            move[0][j] = 3
        for i in range(m + 1):
            cost[i][0] = i
            move[i][0] = 2
        move[0][0] = 0
        band = max_indel + 2
        for i in range(1, m + 1):
            ci, cp, mi = cost[i], cost[i - 1], move[i]
            a = motif[i - 1]
            for j in range(max(1, i - band), min(n, i + band) + 1):
                best, mv = cp[j - 1] + (a != window[j - 1]), 1
                d = cp[j] + 1
                if d < best:
                    best, mv = d, 2
                d = ci[j - 1] + 1
                if d < best:
                    best, mv = d, 3
                ci[j], mi[j] = best, mv
        end_j, end_cost = -1, big
        for j in range(lo, hi + 1):
            if cost[m][j] < end_cost:
                end_cost, end_j = cost[m][j], j
        if end_j <= 0 or end_cost >= big:
            return None
        # traceback -> aligned column pairs (motif base or None, window base or None)
        cols = []
        i, j = m, end_j
        while i > 0 or j > 0:
            mv = move[i][j]
            if mv == 1:
                cols.append((motif[i - 1], window[j - 1]))
                i, j = i - 1, j - 1
            elif mv == 2:
                cols.append((motif[i - 1], None))
                i -= 1
            elif mv == 3:
                cols.append((None, window[j - 1]))
                j -= 1
            else:
                break
        cols.reverse()
        ops: List[Tuple] = []
        seen: List[Tuple[int, str]] = []
        subs = ins_total = del_total = 0
        ref_pos = 0
        ins_buf: List[str] = []
        ins_at = 0
        del_run = 0
        del_at = 0
        for r, q in cols:
            if r is None:
                if not ins_buf:
                    ins_at = ref_pos
                ins_buf.append(q)
                continue
            if ins_buf:
                ops.append(("ins", ins_at, "".join(ins_buf)))
                ins_total += len(ins_buf)
                ins_buf, ins_at = [], 0
            ref_pos += 1
            if q is None:
                if del_run == 0:
                    del_at = ref_pos
                del_run += 1
                continue
            if del_run:
                ops.append(("del", del_at, del_run))
                del_total += del_run
                del_run = 0
            seen.append((ref_pos - 1, q))
            if r != q:
                ops.append(("sub", ref_pos, r, q))
                subs += 1
        if ins_buf:
            ops.append(("ins", ins_at, "".join(ins_buf)))
            ins_total += len(ins_buf)
        if del_run:
            ops.append(("del", del_at, del_run))
            del_total += del_run
        if subs > mismatch_tolerance or ins_total > max_indel or del_total > max_indel:
            return None
        return AlignmentResult(consumed=end_j, unit_sequence=window[:end_j], mismatch_count=subs,
                               insertion_length=ins_total, deletion_length=del_total, operations=ops,
                               observed_bases=seen, edit_distance=end_cost)

    @staticmethod
    def _consensus_from_counts(counts: List[Counter], fallback: str) -> str:
        out = []
        for idx, tally in enumerate(counts):
            if tally:
                # == tally.most_common(1)[0][0] (heapq.nlargest(1) is max(): the first maximum in
                # insertion order), without the detour through heapq
                out.append(max(tally.items(), key=_COUNT_OF)[0])
            else:
                out.append(fallback[idx] if idx < len(fallback) else "N")
        return "".join(out)

    @staticmethod
    def align_repeat_region(sequence: str, start: int, end: int, motif_template: str,
                            mismatch_fraction: float = 0.1, max_indel: Optional[int] = None,
                            min_copies: int = 3) -> Optional[RepeatAlignmentSummary]:
        """Walk copy by copy from `start`, re-deriving the consensus after every
        copy (bwt.py:997-1102)."""
        if not motif_template or not sequence:
            return None
        total = len(sequence)
        start = max(0, start)
        end = min(total, end if end > start else total)
        k = len(motif_template)
        tol = max(1, int(math.floor(k * mismatch_fraction)))
        if max_indel is None:
            max_indel = max(1, min(10, k // 2 if k >= 4 else 1))
        else:
            max_indel = max(0, max_indel)
        tallies: List[Counter] = [Counter() for _ in range(k)]
        units: List[str] = []
        unit_ops: List[List[Tuple]] = []
        errors: List[int] = []
        n_ins = n_del = 0
        template = motif_template
        pos = start
        stop = min(total, max(end, start + k * min_copies) + max(k * 3, max_indel * 4))
        fast = MotifUtils._EXACT_COPY_FAST_PATH
        while pos < stop:
            if fast and sequence.startswith(template, pos):
                # Exact copies: the banded DP would return the all-diagonal alignment (cost 0 at
                # column k, every shorter end column costs >= 1, ties prefer the diagonal), and
                # adding the template's own bases keeps every column's first-maximum tally, so
                # the consensus does not move.  A run of exact copies is therefore taken in one
                # step; most copies of a merged block are exact, which makes this the difference
                # between O(k^2) and a string compare per copy.
                run, nxt = 1, pos + k
                while nxt < stop and sequence.startswith(template, nxt):
                    run += 1
                    nxt += k
                units.extend([template] * run)
                unit_ops.extend([] for _ in range(run))
                errors.extend([0] * run)
                for col, base in enumerate(template):
                    tallies[col][base] += run
                pos = nxt
                continue
            window = sequence[pos:min(total, pos + k + max_indel)]
            if len(window) < k - max_indel:
                break
            res = MotifUtils._align_unit_to_window(template, window, max_indel, tol)
            if res is None or res.consumed == 0:
                break
            units.append(res.unit_sequence)
            unit_ops.append(res.operations)
            errors.append(res.error_count)
            n_ins += res.insertion_length
            n_del += res.deletion_length
            for col, base in res.observed_bases:
                if 0 <= col < k:
                    tallies[col][base] += 1
            pos += res.consumed
            template = MotifUtils._consensus_from_counts(tallies, template)
        if len(units) < min_copies or pos - start <= 0:
            return None
        consensus = MotifUtils._consensus_from_counts(tallies, template)
        denom = len(units) * k
        notes: List[str] = []
        for copy_no, ops in enumerate(unit_ops, 1):
            for op in ops:
                if not op:
                    continue
                if op[0] == "sub":
                    notes.append(f"{copy_no}:{op[1]}:{op[2]}>{op[3]}")
                elif op[0] == "ins" and op[2]:
                    notes.append(f"{copy_no}:{op[1]}:ins({op[2]})")
                elif op[0] == "del" and op[2] > 0:
                    notes.append(f"{copy_no}:{op[1]}:del({op[2]})")
        return RepeatAlignmentSummary(
            consensus=consensus, motif_len=k, copies=len(units), consumed_length=pos - start,
            mismatch_rate=(sum(errors) / denom if denom > 0 else 0.0),
            max_errors_per_copy=(max(errors) if errors else 0), variations=notes, copy_sequences=units,
            total_insertions=n_ins, total_deletions=n_del, error_counts=errors)

    # ---- small string predicates (bwt.py:1104-1167) -------------------------
    @staticmethod
    def is_insertion_variant(candidate: str, consensus: str) -> bool:
        if len(candidate) != len(consensus) + 1:
            return False
        return any(candidate[:i] + candidate[i + 1:] == consensus for i in range(len(candidate)))

    @staticmethod
    def is_deletion_variant(candidate: str, consensus: str) -> bool:
        if len(candidate) + 1 != len(consensus):
            return False
        return any(consensus[:i] + consensus[i + 1:] == candidate for i in range(len(consensus)))

    @staticmethod
    def smallest_period_str(s: str) -> int:
        n = len(s)
        for p in range(1, n + 1):
            if n % p == 0 and s[:p] * (n // p) == s:
                return p
        return n

    @staticmethod
    def normalize_variant(candidate: str, consensus: str) -> str:
        if not candidate:
            return candidate
        best, best_cost = candidate, MotifUtils.edit_distance(candidate, consensus)
        for rot in _rotations(candidate)[1:]:
            c = MotifUtils.edit_distance(rot, consensus)
            if c < best_cost or (c == best_cost and rot < best):
                best, best_cost = rot, c
        return best

    @staticmethod
    def rotate_deletion_variant(candidate: str, consensus: str) -> str:
        if not candidate:
            return candidate
        rot = candidate
        for _ in range(len(candidate)):
            if rot[0] != consensus[0]:
                break
            rot = rot[1:] + rot[:1]
        return rot

    # ---- consensus (bwt.py:1170-1256) ----------------------------------------
    @staticmethod
    def build_consensus_motif(sequences: List[str]) -> Tuple[str, float]:
        if not sequences:
            return "", 0.0
        if len(sequences) == 1:
            return sequences[0], 0.0
        k = len(sequences[0])
        cons, mism = [], 0
        for col in range(k):
            column = [s[col] for s in sequences if col < len(s)]
            if not column:
                cons.append("N")
                continue
            tally = Counter(column)
            top = tally.most_common(1)[0][0]
            cons.append(top)
            mism += len(column) - tally[top]
        cells = len(sequences) * k
        return "".join(cons), (mism / cells if cells > 0 else 0.0)

    @staticmethod
    def build_consensus_motif_array(text_arr: np.ndarray, start: int, motif_len: int,
                                    n_copies: int) -> Tuple[np.ndarray, float, int]:
        """Majority vote per column, ties to the smallest byte; runs on the GPU
        (bwtk_consensus_batch).  Returns (consensus, mismatch_rate, max per copy)."""
        if n_copies == 0 or motif_len == 0:
            return np.array([], dtype=np.uint8), 0.0, 0
        from . import detect

        cons, total_mm, max_mm, used = detect.consensus_one(text_arr, int(start), int(motif_len), int(n_copies))
        if used == 0:
            return np.array([], dtype=np.uint8), 0.0, 0
        cells = used * motif_len
        return cons, (total_mm / cells if cells > 0 else 0.0), max_mm

    @staticmethod
    def summarize_variations_array(text_arr: np.ndarray, start: int, end: int, motif_len: int,
                                   consensus_arr: np.ndarray) -> List[str]:
        if text_arr.size == 0 or motif_len <= 0:
            return []
        sequence = text_arr.tobytes().decode("ascii", errors="replace")
        start = max(0, start)
        end = min(len(sequence), end if end > start else len(sequence))
        if end <= start:
            return []
        if consensus_arr.size > 0:
            template = consensus_arr.tobytes().decode("ascii", errors="replace")
        else:
            template = sequence[start:start + motif_len]
        summary = MotifUtils.align_repeat_region(sequence, start, end, template, mismatch_fraction=0.1,
                                                 min_copies=1)
        return summary.variations if summary else []

    # ---- TRF-style derived fields (bwt.py:1289-1366) ---------------------------
    @staticmethod
    def calculate_composition(sequence: str) -> Dict[str, float]:
        if not sequence:
            return {"A": 0.0, "C": 0.0, "G": 0.0, "T": 0.0}
        tally = Counter(sequence.upper())
        n = len(sequence)
        return {b: (tally.get(b, 0) / n) * 100.0 for b in "ACGT"}

    @staticmethod
    def calculate_trf_score(consensus: str, copies: int, mismatch_rate: float, length: int) -> int:
        good = length * (1.0 - mismatch_rate)
        bad = length * mismatch_rate
        return max(0, int((good * 2) - (bad * 7)))

    @staticmethod
    def calculate_trf_statistics(text_arr: np.ndarray, start: int, end: int, consensus_motif: str,
                                 copies: int, mismatch_rate: float):
        if end <= text_arr.size:
            actual = text_arr[start:end].tobytes().decode("ascii", errors="replace")
        else:
            actual = consensus_motif * int(copies)
        pct = (1.0 - mismatch_rate) * 100.0
        return (pct, 0.0, MotifUtils.calculate_trf_score(consensus_motif, copies, mismatch_rate, end - start),
                MotifUtils.calculate_composition(consensus_motif), MotifUtils.calculate_entropy(consensus_motif),
                actual)

    # ---- motif enumeration (bwt.py:1368-1381) -----------------------------------
    @staticmethod
    def enumerate_motifs(k: int, alphabet: str = "ACGT") -> Iterator[str]:
        """Canonical (smallest rotation) primitive motifs of length k in
        lexicographic order, i.e. the Lyndon words of length k."""
        def grow(prefix: str):
            if len(prefix) == k:
                if MotifUtils.get_canonical_motif(prefix) == prefix and MotifUtils.is_primitive_motif(prefix):
                    yield prefix
                return
            for ch in alphabet:
                yield from grow(prefix + ch)

        yield from grow("")


@functools.lru_cache(maxsize=1 << 17)
def _align_unit_cached(motif: str, window: str, max_indel: int, mismatch_tolerance: int):
    return MotifUtils._align_unit_to_window_dp(motif, window, max_indel, mismatch_tolerance)


# ---- native copy-by-copy walk (csrc/rowchain.cu: bwtk_align_repeat_region) ---------------------------------
_NATIVE_WALK = None      # False once the library turned out to be unavailable


class _WalkBuffers(threading.local):
    """Output buffers of the native walk, one set per thread (ctypes releases the GIL during the call)."""

    def __init__(self):
        import ctypes as C

        self.notes = C.create_string_buffer(1 << 12)
        self.cons = C.create_string_buffer(1 << 10)
        self.out = (C.c_int64 * 8)()


_WALK_BUFFERS = None


def align_region_native(seq_bytes: bytes, start: int, end: int, template: bytes, mismatch_fraction: float = 0.1,
                        max_indel: Optional[int] = None, min_copies: int = 3):
    """``MotifUtils.align_repeat_region`` on the ASCII bytes of a contig, in native code.

    Returns the fields ``_recompute_repeat`` reads -- (consensus str, copies, consumed_length, mismatch_rate,
    max_errors_per_copy, variations, total_insertions, total_deletions) --, ``None`` where the reference
    returns None, or ``NotImplemented`` when the native walk does not apply (library missing, a consensus
    column with more than 8 distinct symbols): the caller then runs the Python walk, which defines the
    behaviour (tests/test_host_glue_cpu.py holds the two equal on random and planted sequences)."""
    global _NATIVE_WALK, _WALK_BUFFERS
    import ctypes as C

    if _NATIVE_WALK is False:
        return NotImplemented
    if _NATIVE_WALK is None:
        from . import _lib

        try:
            walk = _lib.lib().bwtk_align_repeat_region
        except Exception:
            _NATIVE_WALK = False
            return NotImplemented
        _WALK_BUFFERS = _WalkBuffers()
        _NATIVE_WALK = walk
    k = len(template)
    if k == 0 or not seq_bytes:
        return None
    tol = max(1, int(math.floor(k * mismatch_fraction)))
    if max_indel is None:
        max_indel = max(1, min(10, k // 2 if k >= 4 else 1))
    else:
        max_indel = max(0, max_indel)
    buf = _WALK_BUFFERS
    if k > len(buf.cons):
        buf.cons = C.create_string_buffer(2 * k)
    cons, out = buf.cons, buf.out
    while True:
        rc = _NATIVE_WALK(seq_bytes, len(seq_bytes), int(start), int(end), template, k, tol, max_indel, int(min_copies),
                          cons, out, buf.notes, len(buf.notes))
        if rc == -4:                                     # BWTK_EOVERFLOW: the notes need a larger buffer
            buf.notes = C.create_string_buffer(int(out[7]) * 2)
            continue
        break
    if rc == 0:
        return None
    if rc != 1:
        return NotImplemented
    copies, consumed, errs, worst, n_ins, n_del, n_notes, note_bytes = out
    denom = copies * k
    notes = buf.notes.raw[:note_bytes].decode("ascii").split("\n") if n_notes else []
    return (cons.raw[:k].decode("ascii"), copies, consumed, (errs / denom if denom > 0 else 0.0), worst, notes,
            n_ins, n_del)
