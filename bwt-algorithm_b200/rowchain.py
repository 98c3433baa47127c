"""Row-oriented post-processing chain for the CLI's detector.

The reference turns every strict-adjacency call into a ``TandemRepeat`` and runs the chain
suppress -> dedup -> merge -> refine -> restore -> collapse -> final filter over Python objects
(bwt.py:3827-3846, 3189-3614).  A chr21-sized contig yields 2.5 M calls of which ~7 % survive, so
building the objects up front is the CLI's bottleneck (≈1 s per Mb).  Here the calls stay the integer
rows the kernels emit -- (start, end, primitive period, copies, 0, 0, unit, 0) -- for as long as the
chain's decisions depend on those integers only:

  * nested-call suppression: native, on the rows (csrc/rowchain.cu, bwtk_suppress_nested);
  * dedup and the sorts: NumPy on (start, end, period);
  * merge and collapse are sequential passes in which a call only interacts with its neighbour when a
    cheap NECESSARY condition holds (same canonical motif and a small gap for a merge; a positive
    overlap for a collapse).  Rows are skipped in bulk while the condition is false; where it holds, the
    rows involved are turned into records and the reference's own predicates and re-alignment run on
    them (``TandemRepeatFinder._should_merge_repeats`` etc. -- the same code the object chain uses);
  * the final filter is an integer test for untouched rows; only the survivors become records.

Every decision is made by the same predicate on the same values as in the object chain, so the output is
the same list, field for field (tests/test_host_glue_cpu.py holds the two chains equal on the reference's
fixtures, on planted contigs and on the 60 kb golden chain).  Works on ONE contig; the caller
concatenates contigs and sorts (every stage of the chain acts inside one contig).
"""
from __future__ import annotations

import copy
from typing import List, Optional, Tuple

import numpy as np

from . import _lib
from .finders import StrictRecordMaker, strict_records
from .motifs import MotifUtils
from .records import TandemRepeat

FAST_K1 = True     # tests switch it off to compare with the generic merge path
STATS = {"requeued": 0}   # settled homopolymer events that a merge chain reached after all (tests look at it)

_AT = np.zeros(256, np.int8)
_AT[:] = -1                      # canonical class of a 1-bp motif: A/T -> 0, C/G -> 1, anything else: ask MotifUtils
for _c, _v in ((65, 0), (84, 0), (67, 1), (71, 1)):
    _AT[_c] = _v


class _RowCall:
    """What the merge predicates read of an exact strict-scan call (TandemRepeatFinder._should_merge_repeats /
    _merge_repeats: chrom, motif, span, tier, mismatch rate), without the record around it: four out of five
    rows that reach a merge test are absorbed by the union and never become records."""
    __slots__ = ("chrom", "start", "end", "motif", "consensus_motif", "tier", "mismatch_rate")

    def __init__(self, chrom, start, end, motif):
        self.chrom, self.start, self.end = chrom, start, end
        self.motif = self.consensus_motif = motif
        self.tier, self.mismatch_rate = 2, 0.0


def suppress_rows(rows: np.ndarray, overlap_threshold: float = 0.5) -> np.ndarray:
    """Boolean keep mask of ``_suppress_nested_short_calls`` for the exact calls of one contig."""
    n = int(rows.shape[0])
    keep = np.ones(n, np.uint8)
    if n == 0:
        return keep.astype(bool)
    L = _lib.lib()
    start = np.ascontiguousarray(rows[:, 0], np.int32)
    end = np.ascontiguousarray(rows[:, 1], np.int32)
    k = np.ascontiguousarray(rows[:, 2], np.int32)
    _lib.check(L.bwtk_suppress_nested(start.ctypes.data, end.ctypes.data, k.ctypes.data, None, n,
                                      float(overlap_threshold), keep.ctypes.data), "suppress_nested")
    return keep.astype(bool)


def finish_rows(finder, chrom: str, text_arr: np.ndarray, rows: np.ndarray,
                info: Optional[dict] = None) -> Tuple[List[TandemRepeat], int, int, int]:
    """The whole chain for the exact (mismatch-free) strict-scan rows of one contig.

    `finder` is a ``TandemRepeatFinder`` that knows this contig (sequences / full_sequences /
    trim_offsets); returns (final calls, raw, after suppression, after dedup) like
    ``TandemRepeatFinder._postprocess_counts``.  `info` (optional) receives ``reach``: the first position at which a
    later call could no longer be merged into, or collapse with, anything this chain formed beyond its rows (merged
    and re-aligned calls: end + motif length + 1) -- what ``finish_rows_segmented`` checks its cuts against."""
    if info is not None:
        info["reach"] = -1
    n_raw = int(rows.shape[0])
    if n_raw == 0:
        return [], 0, 0, 0
    rows = np.ascontiguousarray(rows, np.int32)
    min_copies = finder.min_copies

    # ---- 1. nested-call suppression; survivors sorted by (start, end), ties: longer motif first, then
    #         arrival order (the reference's two stable sorts)
    keep = suppress_rows(rows)
    idx = np.flatnonzero(keep)
    n_kept = int(idx.size)
    S = rows[idx, 0].astype(np.int64)
    E = rows[idx, 1].astype(np.int64)
    K = rows[idx, 2].astype(np.int64)
    order = np.lexsort((idx, -K, E, S))
    idx, S, E, K = idx[order], S[order], E[order], K[order]

    # ---- 2. dedup on (chrom, start, end, motif): equal spans with equal primitive period carry the same
    #         motif; all fields of exact calls tie, so the first one stays
    if idx.size > 1:
        same = (S[1:] == S[:-1]) & (E[1:] == E[:-1]) & (K[1:] == K[:-1])
        # equal (start, end) rows are adjacent and ordered by -K: duplicates of a (start, end, K) are adjacent
        first = np.concatenate(([True], ~same))
        idx, S, E, K = idx[first], S[first], E[first], K[first]
    n_unique = int(idx.size)
    m = int(idx.size)
    Cn = rows[idx, 3].astype(np.int64)

    # the contig's bytes, kept on the finder while it is finishing pieces of the same text
    held = finder.__dict__.get("_chain_text")
    if held is None or held[0] is not text_arr:
        held = finder.__dict__["_chain_text"] = (text_arr, text_arr.tobytes())
    text_bytes = held[1]
    maker = StrictRecordMaker(text_arr, chrom, 0, text_bytes)
    # the contig as the chain's re-alignments see it (finder.sequences: no sentinel); every slice below ends at or
    # before seq_len, so the text's own bytes serve
    seq_len = len(finder.sequences.get(chrom) or "")
    seq_bytes = text_bytes
    fast_k1 = FAST_K1 and seq_len > 0 and min_copies >= 1

    def make_row(i: int) -> TandemRepeat:
        return maker(int(S[i]), int(E[i]), int(K[i]), int(Cn[i]))

    def light_row(i: int) -> _RowCall:
        s0 = int(S[i])
        return _RowCall(chrom, s0, int(E[i]), text_bytes[s0:s0 + int(K[i])].decode("ascii", errors="replace"))

    # ---- 3. merge adjacent.  link[i]: rows i and i+1 pass the cheap (necessary) tests of _should_merge_repeats
    link = np.zeros(m, bool)
    if m > 1:
        gap = np.maximum(0, S[1:] - E[:-1])
        cand = (K[1:] == K[:-1]) & (gap <= np.minimum(K[1:], K[:-1]) + 1)
        ci = np.flatnonzero(cand)
        if ci.size:
            one = K[ci] == 1
            a = _AT[text_arr[S[ci]]]
            b = _AT[text_arr[S[ci + 1]]]
            known = one & (a >= 0) & (b >= 0)
            link[ci[known & (a == b)]] = True
            for i in ci[~known].tolist():
                k = int(K[i])
                m1 = text_bytes[S[i]:S[i] + k].decode("ascii", errors="replace")
                m2 = text_bytes[S[i + 1]:S[i + 1] + k].decode("ascii", errors="replace")
                if MotifUtils.get_canonical_motif_stranded(m1)[0] == MotifUtils.get_canonical_motif_stranded(m2)[0]:
                    link[i] = True
    alive = np.ones(m, bool)             # False: the row was merged into another item
    vrun = np.zeros(m, np.int64)         # > 0: the row stands for a merged homopolymer call [S, S + vrun)
    vemax = np.zeros(m, np.int64)        #      ... whose union was asked for up to this end
    todo = link.copy()
    if fast_k1 and link.any():
        # Two neighbouring homopolymer calls, isolated from other merge candidates.  With a 1-bp motif the
        # per-copy aligner can only take exact copies (on a mismatching base "delete the motif base" -- end
        # column 0 -- is the first minimum and the walk stops), so the union the reference computes is the
        # leading run of the first call's base, capped by the walk's stop column; the second call is
        # absorbed.  The merged call then meets the next row, which must fail the gap test for the event
        # to be settled here.  These events are 6 % of a genome's rows; they stay integers ("virtual"
        # items) and become records -- through the real _recompute_repeat -- only if they survive.
        prev_l = np.zeros(m, bool)
        prev_l[1:] = link[:-1]
        next_l = np.zeros(m, bool)
        next_l[:-1] = link[1:]
        iso = np.flatnonzero(link & (K == 1) & ~prev_l & ~next_l)
        if iso.size:
            s0 = S[iso]
            e_max = np.maximum(E[iso], E[iso + 1])
            stop = np.minimum(seq_len, np.maximum(e_max, s0 + max(1, min_copies)) + 4)
            # starts of the maximal runs, inside the window these events can see (a piece of a contig only looks at
            # its own stretch); a run that leaves the window ends beyond every stop, which caps it
            w0, w1 = int(s0.min()), min(seq_len, int(stop.max()) + 1)
            window = text_arr[w0:w1]
            bounds = np.flatnonzero(window[1:] != window[:-1]) + 1 + w0
            bi = np.searchsorted(bounds, s0, side="right")
            run_end = np.where(bi < bounds.size, bounds[np.minimum(bi, max(bounds.size - 1, 0))], w1) \
                if bounds.size else np.full(iso.size, w1, np.int64)
            run = np.minimum(run_end, stop) - s0
            nxt = iso + 2
            has = nxt < m
            nx = np.minimum(nxt, m - 1)
            settled = ~has | (np.maximum(0, S[nx] - (s0 + run)) > np.minimum(1, K[nx]) + 1)
            good = (run >= max(1, min_copies)) & settled & (s0 + 1 <= seq_len)
            simple = iso[good]
            vrun[simple] = run[good]
            vemax[simple] = e_max[good]
            alive[simple + 1] = False
            todo[simple] = False
    # the reference's loop from every remaining candidate: on integers while only homopolymer calls over
    # A/C/G/T are involved (same closed form, step by step), on records otherwise
    merged: List[Tuple[int, TandemRepeat]] = []         # (list slot = position of the first row, record)
    free_from = 0
    mc1 = max(1, min_copies)
    seq_arr = text_arr[:seq_len]
    cands = np.flatnonzero(todo).tolist()
    p_c, n_c = 0, len(cands)
    pending = -1            # a settled event that a chain reached and undid: it is a candidate again, and the next one
    while True:
        while p_c < n_c and cands[p_c] < free_from:
            p_c += 1
        if pending >= 0 and (p_c >= n_c or pending <= cands[p_c]):
            e, pending = pending, -1
            if e < free_from:
                continue
        elif p_c < n_c:
            e = cands[p_c]
            p_c += 1
        else:
            break
        if fast_k1 and K[e] == 1 and _AT[seq_arr[S[e]]] >= 0:
            s0 = int(S[e])
            cls = int(_AT[seq_arr[s0]])
            base = seq_bytes[s0:s0 + 1]
            cur_end = int(E[e])
            j = e + 1
            emax_last = 0
            clean = True
            while j < m:
                if K[j] != 1 or max(0, int(S[j]) - cur_end) > 2:
                    break                        # motif lengths differ / gap test fails: no merge
                cj = int(_AT[seq_arr[S[j]]])
                if cj < 0:
                    clean = False                # a non-ACGT homopolymer: its canonical form is MotifUtils' business
                    break
                if cj != cls:
                    break
                if vrun[j] > 0:                  # a settled event is reached after all: it never happened
                    vrun[j] = 0
                    alive[j + 1] = True
                e_max = max(cur_end, int(E[j]))
                stop = min(seq_len, max(e_max, s0 + mc1) + 4)
                lim = seq_bytes[s0:stop]
                run = len(lim) - len(lim.lstrip(base))
                if run < mc1:
                    clean = False
                    break
                cur_end = s0 + run
                emax_last = e_max
                j += 1
            if clean:
                if j > e + 1:
                    alive[e + 1:j] = False
                    vrun[e] = cur_end - s0
                    vemax[e] = emax_last
                    free_from = j
                else:
                    free_from = e + 1
                if j < m and vrun[j] == 0 and link[j] and not todo[j]:
                    pending = j                  # an undone event is current again
                    todo[j] = True
                    STATS["requeued"] += 1
                continue
            # fall through: redo this candidate on records (rows released above stay released)
        cur = light_row(e)
        j = e + 1
        touched = False
        while j < m:
            k_cur = len(cur.consensus_motif or cur.motif)
            if max(0, int(S[j]) - cur.end) > min(k_cur, int(K[j])) + 1:
                break                                    # fails the gap test: no merge (cheap exit)
            if vrun[j] > 0:                              # a settled homopolymer event is reached after all:
                vrun[j] = 0                              # it never happened -- its rows are ordinary rows again
                alive[j + 1] = True
            nxt_rec = light_row(j)
            if finder._should_merge_repeats(cur, nxt_rec):
                cur = finder._merge_repeats(cur, nxt_rec)
                touched = True
                j += 1
            else:
                break
        if touched:
            alive[e:j] = False
            merged.append((e, cur))
            free_from = j
        else:
            free_from = e + 1
        if j < m and vrun[j] == 0 and link[j] and not todo[j]:
            pending = j                          # an undone event is current again
            todo[j] = True
            STATS["requeued"] += 1

    # ---- 4. refine (records with mismatches only; rows and virtual items are exact) + stable sort by (start, end)
    rec_list: List[TandemRepeat] = []
    rec_slot: List[int] = []
    for slot, r in merged:
        if r.mismatch_rate != 0.0:
            k = len(r.consensus_motif or r.motif)
            if k <= 0:
                k = max(1, r.length // max(1, int(round(r.copies)) or 1))
            r = finder._recompute_repeat(r.chrom, r.start, r.end, k, tier_hint=r.tier)
        rec_list.append(r)
        rec_slot.append(slot)
    if info is not None:
        reach = -1
        for _, r in merged:                               # as merged (before the refinement moved their ends)
            reach = max(reach, r.end + len(r.consensus_motif or r.motif) + 1)
        for r in rec_list:
            reach = max(reach, r.end + len(r.consensus_motif or r.motif) + 1)
        virt_i = np.flatnonzero(alive & (vrun > 0))
        if virt_i.size:
            reach = max(reach, int((S[virt_i] + vrun[virt_i]).max()) + 2)
        info["reach"] = reach
    rpos = np.flatnonzero(alive)                          # rows and virtual items, by list slot
    n_rows_items = int(rpos.size)
    ref = np.concatenate((rpos, -1 - np.arange(len(rec_list), dtype=np.int64)))     # >= 0: position i; < 0: record
    slot = np.concatenate((rpos, np.array(rec_slot, np.int64)))
    kS = np.concatenate((S[rpos], np.array([r.start for r in rec_list], np.int64)))
    kE = np.concatenate((np.where(vrun[rpos] > 0, S[rpos] + vrun[rpos], E[rpos]),
                         np.array([r.end for r in rec_list], np.int64)))
    # == np.lexsort((slot, kE, kS)): the items are brought into slot order first (two ascending runs), then sorted
    # stably by one combined key -- both inputs are nearly sorted, which the stable sort (timsort) turns into a
    # couple of merges where a three-key lexsort pays for three full sorts of millions of items
    if kS.size and int(kS.min()) >= 0 and int(kE.min()) >= 0 and int(kE.max()) < (1 << 31) and int(kS.max()) < (1 << 31):
        by_slot = np.argsort(slot, kind="stable")
        order = by_slot[np.argsort((kS[by_slot] << 31) | kE[by_slot], kind="stable")]
    else:
        order = np.lexsort((slot, kE, kS))
    ref, slot, kS, kE = ref[order], slot[order], kS[order], kE[order]
    is_rec = ref < 0

    # ---- 5. restore reference coordinates (records now; rows when they become records)
    shift = finder.trim_offsets.get(chrom, 0)
    full = finder.full_sequences.get(chrom)

    def restore(r: TandemRepeat) -> TandemRepeat:
        r.start += shift
        r.end += shift
        r.length = r.end - r.start
        if full:
            r.actual_sequence = full[r.start:r.end]
        return r

    for r in rec_list:
        restore(r)
    kS = kS + shift
    kE = kE + shift
    k1_cache = {}

    def make_virtual(i: int) -> TandemRepeat:
        """The merged homopolymer call of a settled event: the real union for the first (base, run), copies of
        it at other offsets (its fields depend on nothing else)."""
        s0, run = int(S[i]), int(vrun[i])
        key = (seq_bytes[s0:s0 + 1], run)
        tmpl = k1_cache.get(key)
        if tmpl is None:
            u = finder._recompute_repeat(chrom, s0, int(vemax[i]), 1, tier_hint=2)
            if not (u.start == s0 and u.end == s0 + run and u.mismatch_rate == 0.0 and u.variations is None):
                raise _lib.BwtkError("homopolymer union differs from its closed form")      # never: see tests
            k1_cache[key] = u
            return copy.copy(u)
        r = copy.copy(tmpl)
        r.start, r.end = s0, s0 + run
        r.composition = dict(tmpl.composition)
        return r

    made_cache = {}

    def make_item(q: int) -> TandemRepeat:
        """Item q of the sorted list as a record in reference coordinates (one object per item)."""
        got = made_cache.get(q)
        if got is None:
            if is_rec[q]:
                got = rec_list[-1 - int(ref[q])]
            else:
                i = int(ref[q])
                got = restore(make_virtual(i) if vrun[i] > 0 else make_row(i))
            made_cache[q] = got
        return got

    # ---- 6. collapse overlapping: a call can only collapse with its predecessor when they overlap
    t = int(ref.size)
    gone = np.zeros(t, bool)                 # items collapsed away
    winners = {}                             # slot position q -> the record that holds it after collapsing
    if t > 1:
        ov = (np.minimum(kE[:-1], kE[1:]) - np.maximum(kS[:-1], kS[1:])) > 0
        q_free = 0
        for q in np.flatnonzero(ov).tolist():
            if q < q_free:
                continue
            cur = make_item(q)
            p = q + 1
            while p < t:
                if min(cur.end, int(kE[p])) - max(cur.start, int(kS[p])) <= 0:
                    break
                nxt_rec = make_item(p)
                if finder._should_collapse_duplicates(cur, nxt_rec):
                    cur = finder._prefer_repeat_entry(cur, nxt_rec)
                    gone[p] = True
                    p += 1
                else:
                    break
            if p > q + 1:
                winners[q] = cur
                q_free = p
            else:
                q_free = q + 1

    # ---- 7. final filter, then records for the survivors only
    ri = np.maximum(ref, 0)
    virt = ~is_rec & (vrun[ri] > 0)
    copies = np.where(virt, vrun[ri], Cn[ri])
    length = kE - kS
    ok = ~gone & ~is_rec & (copies >= min_copies) & (length >= 6)
    for q, r in winners.items():
        ok[q] = False
    out: List[Tuple[Tuple[int, int, int], TandemRepeat]] = []
    for q in np.flatnonzero(is_rec & ~gone).tolist():
        if q in winners:
            continue
        r = rec_list[-1 - int(ref[q])]
        if r.copies >= min_copies and r.length >= 6:
            out.append(((r.start, r.end, q), r))
    for q, r in winners.items():
        if r.copies >= min_copies and r.length >= 6:
            out.append(((r.start, r.end, q), r))
    qs = np.flatnonzero(ok)
    plain = qs[~virt[qs]]
    made: List[TandemRepeat] = []
    if plain.size:
        have = [made_cache.get(int(q)) for q in plain.tolist()] if made_cache else None
        made = strict_records(text_arr, rows[idx[ref[plain]]], chrom, 0, text_bytes)
        for n_, r in enumerate(made):
            if have is not None and have[n_] is not None:
                made[n_] = have[n_]
            else:
                restore(r)
    for q in qs[virt[qs]].tolist():
        r = make_item(q)
        out.append(((r.start, r.end, q), r))
    if not out:
        return made, n_raw, n_kept, n_unique
    # sort by (start, end, list position): the survivors made from rows are in that order already, the others
    # (merged, collapsed and virtual calls) are filed among them
    o_keys = np.array([k for k, _ in out], np.int64).reshape(-1, 3)
    s_all = np.concatenate((kS[plain], o_keys[:, 0]))
    e_all = np.concatenate((kE[plain], o_keys[:, 1]))
    q_all = np.concatenate((plain, o_keys[:, 2]))
    every = made + [r for _, r in out]
    return [every[i] for i in np.lexsort((q_all, e_all, s_all)).tolist()], n_raw, n_kept, n_unique


# ---- one contig, several processes -------------------------------------------------------------------
# Every stage of the chain is local: a call interacts with another only by overlapping it (suppression, collapse)
# or by starting within motif length + 1 of its end (merge), and what a merge forms extends forwards only.  So
# the rows, sorted by start, can be cut wherever every row before the cut ends more than its motif length + 1
# before the next row starts; the pieces are finished independently (in worker processes) and concatenated.  The
# one thing a cut cannot know beforehand is how far a merged, re-aligned call reaches beyond its rows: every piece
# reports that (``info["reach"]``), and if any piece reached the first row of the next one the contig is finished
# in one piece instead.
SEGMENT_MIN_ROWS = 200_000        # below this the chain is faster than starting workers; tests lower it


def segment_bounds(rows: np.ndarray, order: np.ndarray, pieces: int, min_rows: int) -> List[int]:
    """Boundaries (indices into `order`, the rows' stable order by start) of at most `pieces` independent
    pieces of at least `min_rows` rows: [0, ..., len(rows)]."""
    n = int(order.size)
    if pieces < 2 or n < 2 * max(1, min_rows):
        return [0, n]
    S = rows[order, 0].astype(np.int64)
    reach = np.maximum.accumulate(rows[order, 1].astype(np.int64) + rows[order, 2].astype(np.int64) + 1)
    room = S[1:] - reach[:-1]                                # > 0: a piece may start at row i + 1
    pieces = min(pieces, max(1, n // max(1, min_rows)))
    look = max(1, min(4096, min_rows // 4))                  # rows around the even split among which a cut is chosen
    bounds = [0]
    for t in range(1, pieces):
        want = n * t // pieces
        lo, hi = max(1, want - look), min(n - 1, want + look)
        if lo >= hi:
            continue
        # the widest gap nearby: what a merge forms beyond its rows (the re-alignment walks a few motif lengths
        # past them) must stay clear of the next piece's first row, or the contig is finished in one piece
        at = int(np.argmax(room[lo - 1:hi - 1])) + lo
        if room[at - 1] > 0 and at - bounds[-1] >= min_rows and n - at >= min_rows:
            bounds.append(at)
    bounds.append(n)
    return bounds


def finish_segment(finder, chrom: str, text_arr: np.ndarray, rows: np.ndarray):
    """One piece: (final calls, raw, kept, unique, reach)."""
    info: dict = {}
    final, a, b, c = finish_rows(finder, chrom, text_arr, rows, info)
    return final, a, b, c, info["reach"]


def finish_rows_segmented(finder, chrom: str, text_arr: np.ndarray, rows: np.ndarray, run_pieces, pieces: int,
                          min_rows: Optional[int] = None) -> Tuple[List[TandemRepeat], int, int, int]:
    """``finish_rows`` with the contig's rows cut into independent pieces.  `run_pieces(list of row arrays)` returns
    ``finish_segment``'s result for every piece, in order (the caller decides where they run: a process pool
    in the CLI, a plain loop in the tests).  Same calls, same order, same counts as ``finish_rows``."""
    rows = np.ascontiguousarray(rows, np.int32)
    order = np.argsort(rows[:, 0], kind="stable")
    bounds = segment_bounds(rows, order, pieces, SEGMENT_MIN_ROWS if min_rows is None else min_rows)
    if len(bounds) <= 2:
        return finish_rows(finder, chrom, text_arr, rows)
    # a piece keeps the arrival order of its rows (the chain's tie rules look at it)
    parts = [rows[np.sort(order[bounds[i]:bounds[i + 1]])] for i in range(len(bounds) - 1)]
    firsts = [int(rows[order[b], 0]) for b in bounds[1:-1]]
    results = run_pieces(parts)
    for (_, _, _, _, reach), first in zip(results[:-1], firsts):
        if reach >= first:                       # a merged call of this piece could have met the next piece's rows
            return finish_rows(finder, chrom, text_arr, rows)
    final: List[TandemRepeat] = []
    n_raw = n_kept = n_unique = 0
    for part, a, b, c, _ in results:
        final.extend(part)
        n_raw, n_kept, n_unique = n_raw + a, n_kept + b, n_unique + c
    return final, n_raw, n_kept, n_unique


def _finish_rows_slow(finder, chrom: str, text_arr: np.ndarray, rows: np.ndarray):
    """The same chain with the homopolymer shortcut off (every merge candidate goes through the records)."""
    global FAST_K1
    old = FAST_K1
    FAST_K1 = False
    try:
        return finish_rows(finder, chrom, text_arr, rows)
    finally:
        FAST_K1 = old
