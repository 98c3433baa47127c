"""FM / k-mer seeded seed-and-extend for short imperfect repeats
(Tier2LCPFinder.find_short_imperfect_repeats, reference bwt.py:2027-2095, and
_find_tandems_fm_with_mismatches, bwt.py:2562-2695).

What runs where:
  * seeds -- every rotation (forward and reverse complement) of every canonical
    primitive motif is looked up in ONE batch: 8-mer table buckets for k <= 8
    (with the reference's unpadded-code aliasing), one batched
    FM backward search + SA-interval read for k = 9;
  * extension -- `_extend_tandem_fm` does not depend on the seen-region state,
    so it is evaluated once per (position, motif length) by bwtk_extend_batch
    (mode 1) for all positions, not once per seed and shift;
  * the reference's ordered walk over seeds (seen regions grow, `motif_len`
    persists after a primitive reduction, bwt.py:2635) is replayed on the host
    over those tables; consensus / mismatch statistics of accepted arrays come
    from bwtk_consensus_batch.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Set, Tuple

import numpy as np

from . import detect
from .motifs import MotifUtils
from .records import TandemRepeat

_BITS = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 0}


def _expand_ranges(lo: np.ndarray, hi: np.ndarray, owner: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Concatenated arange(lo[i], hi[i]) for all i plus the owner id of each element."""
    lens = np.maximum(hi - lo, 0).astype(np.int64)
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    starts = np.repeat(lo.astype(np.int64) - (np.cumsum(lens) - lens), lens)
    return starts + np.arange(total, dtype=np.int64), np.repeat(owner.astype(np.int64), lens)


def rotations_of(motif: str) -> List[str]:
    """Distinct rotations of the motif and of its reverse complement (bwt.py:2063-2071)."""
    rc = MotifUtils.reverse_complement(motif)
    both = [motif[i:] + motif[:i] for i in range(len(motif))] + [rc[i:] + rc[:i] for i in range(len(rc))]
    return list(set(both))


def seed_positions(finder, k: int, motifs: Sequence[str]) -> List[np.ndarray]:
    """Sorted unique seed positions of every motif of length k (bwt.py:2073-2084)."""
    core = finder.bwt
    dev = core.device_index
    pats: List[str] = []
    owner: List[int] = []
    for mi, m in enumerate(motifs):
        for r in rotations_of(m):
            pats.append(r)
            owner.append(mi)
    owner_arr = np.asarray(owner, np.int64)
    use_table = k <= 8 and bool(core.kmer_hash)
    if use_table:
        off = dev.kmer_off.cpu().numpy().astype(np.int64)
        codes = np.zeros(len(pats), np.int64)
        for i, p in enumerate(pats):
            w = 0
            for ch in p:
                w = (w << 2) | _BITS[ch]
            codes[i] = w
        idx, own = _expand_ranges(off[codes], off[codes + 1], owner_arr)
        pos = dev.kmer_pos.cpu().numpy().astype(np.int64)[idx] if idx.size else idx
    else:
        sp, ep = dev.backward_search_batch([p.encode("ascii") for p in pats])
        sp = sp.astype(np.int64)
        ep = ep.astype(np.int64)
        idx, own = _expand_ranges(np.where(sp >= 0, sp, 0), np.where(sp >= 0, ep + 1, 0), owner_arr)
        pos = dev.sa.cpu().numpy().astype(np.int64)[idx] if idx.size else idx
    out: List[np.ndarray] = [np.zeros(0, np.int64) for _ in motifs]
    if pos.size:
        order = np.lexsort((pos, own))
        own, pos = own[order], pos[order]
        keep = np.ones(pos.size, bool)
        keep[1:] = (own[1:] != own[:-1]) | (pos[1:] != pos[:-1])
        own, pos = own[keep], pos[keep]
        cuts = np.searchsorted(own, np.arange(len(motifs) + 1))
        for mi in range(len(motifs)):
            out[mi] = pos[cuts[mi]:cuts[mi + 1]]
    return out


class ExtensionTables:
    """ext[L][p] = _extend_tandem_fm(text, p, L) for every position p, built on demand per L."""

    def __init__(self, text_dev, n_total: int):
        self.text_dev = text_dev
        self.n = n_total
        self._tab: Dict[int, np.ndarray] = {}

    def get(self, L: int) -> np.ndarray:
        t = self._tab.get(L)
        if t is None:
            seeds = np.arange(self.n, dtype=np.int32)
            t = detect.extend_batch(self.text_dev, self.n, seeds, np.full(self.n, L, np.int32), None, mode=1)
            self._tab[L] = t
        return t


def replay_seeds(finder, positions: np.ndarray, motif_len: int, covered: np.ndarray, tables: ExtensionTables,
                 text_dev, text_arr: np.ndarray) -> List[tuple]:
    """The ordered seed walk of bwt.py:2574-2693.  Returns accepted arrays as
    (start, end, motif_len, copies, consensus bytes, total_mm, max_mm, used)."""
    out: List[tuple] = []
    size = int(text_arr.size)
    for seed in positions.tolist():
        if covered[seed]:
            continue
        if seed + motif_len > size:
            continue
        ext = tables.get(motif_len)
        best = None
        for shift in range(min(motif_len, seed + 1)):
            cand = seed - shift
            if cand < 0 or cand + motif_len > size or covered[cand]:
                continue
            s, e, c = int(ext[cand, 0]), int(ext[cand, 1]), int(ext[cand, 2])
            if not (s <= seed < e):
                continue
            if best is None or c > best[2] or (c == best[2] and s < best[0]):
                best = (s, e, c)
        if best is None:
            continue
        start, end, copies = best
        if not (copies >= finder.min_copies and end - start >= finder.min_array_length):
            continue
        cons, total_mm, max_mm, used = detect.consensus_one(text_dev, start, motif_len, copies)
        if used == 0:
            continue
        cons_str = cons.tobytes().decode("ascii", errors="replace")
        prim = MotifUtils.smallest_period_str(cons_str)
        if prim < len(cons_str):
            motif_len = prim                      # persists for the later seeds of this call (bwt.py:2635)
            copies = max(1, (end - start) // motif_len)
            end = start + copies * motif_len
            cons, total_mm, max_mm, used = detect.consensus_one(text_dev, start, motif_len, copies)
            if used == 0:
                continue
        if not finder._is_maximal_fm(start, end, cons, motif_len, 0):
            continue
        cells = used * motif_len
        mm_rate = total_mm / cells if cells > 0 else 0.0
        pct = (1.0 - mm_rate) * 100.0
        if pct < (90.0 if motif_len <= 6 else 85.0):
            continue
        out.append((start, end, motif_len, copies, cons, total_mm, max_mm, used))
        covered[start:end] = True
    return out


def records_from_arrays(text_arr: np.ndarray, arrays: List[tuple], chromosome: str) -> List[TandemRepeat]:
    """Accepted arrays -> records (bwt.py:2645-2692)."""
    out: List[TandemRepeat] = []
    for start, end, motif_len, copies, cons, total_mm, max_mm, used in arrays:
        cons_str = cons.tobytes().decode("ascii", errors="replace")
        cells = used * motif_len
        mm_rate = total_mm / cells if cells > 0 else 0.0
        _canon, strand = MotifUtils.get_canonical_motif_stranded(cons_str)
        pm, pi, score, comp, ent, actual = MotifUtils.calculate_trf_statistics(text_arr, start, end, cons_str, copies,
                                                                              mm_rate)
        notes = MotifUtils.summarize_variations_array(text_arr, start, end, motif_len, cons)
        out.append(TandemRepeat(
            chrom=chromosome, start=start, end=end, motif=cons_str, copies=copies, length=end - start, tier=2,
            confidence=max(0.5, 1.0 - mm_rate), consensus_motif=cons_str, mismatch_rate=mm_rate,
            max_mismatches_per_copy=max_mm, n_copies_evaluated=copies, strand=strand, percent_matches=pm,
            percent_indels=pi, score=score, composition=comp, entropy=ent, actual_sequence=actual,
            variations=notes if notes else None))
    return out


def _covered_from(seen_regions: Set[Tuple[int, int]], size: int) -> np.ndarray:
    covered = np.zeros(size + 1, bool)
    for a, b in seen_regions:
        covered[max(0, a):max(0, min(b, size))] = True
    return covered


def find_tandems_from_seeds(finder, positions: List[int], motif: str, chromosome: str, motif_len: int,
                            seen_regions: Set[Tuple[int, int]]) -> List[TandemRepeat]:
    """Drop-in for _find_tandems_fm_with_mismatches (one motif)."""
    from .finders import _device_text_of

    if not positions:
        return []
    text_arr = finder.bwt.text_arr
    text_dev = detect.device_text(_device_text_of(finder.bwt))
    covered = _covered_from(seen_regions, int(text_arr.size))
    tables = ExtensionTables(text_dev, int(text_arr.size))
    arrays = replay_seeds(finder, np.asarray(sorted(positions), np.int64), motif_len, covered, tables, text_dev,
                          text_arr)
    for a in arrays:
        seen_regions.add((a[0], a[1]))
    return records_from_arrays(text_arr, arrays, chromosome)


def short_imperfect_arrays(finder, tier1_seen: Set[Tuple[int, int]]) -> List[tuple]:
    """All accepted arrays of find_short_imperfect_repeats in the reference's order."""
    from .finders import _device_text_of

    text_arr = finder.bwt.text_arr
    n = int(text_arr.size)
    if n > 1_000_000:
        return []
    if not finder.allow_mismatches:
        return []                                  # the reference only extends when mismatches are allowed
    text_dev = detect.device_text(_device_text_of(finder.bwt))
    covered = _covered_from(tier1_seen, n)
    tables = ExtensionTables(text_dev, n)
    out: List[tuple] = []
    for k in range(finder.min_period, min(finder.max_short_motif + 1, 10)):
        motifs = [m for m in MotifUtils.enumerate_motifs(k) if not (MotifUtils.calculate_entropy(m) < finder.min_entropy)]
        if not motifs:
            continue
        seeds = seed_positions(finder, k, motifs)
        for pos in seeds:
            if pos.size >= finder.min_copies:
                out.extend(replay_seeds(finder, pos, k, covered, tables, text_dev, text_arr))
    return out


def find_short_imperfect_repeats(finder, chromosome: str, tier1_seen: Set[Tuple[int, int]]) -> List[TandemRepeat]:
    if int(finder.bwt.text_arr.size) > 1_000_000:
        if finder.show_progress:
            print(f"  [{chromosome}] Tier 2 short imperfect repeats: SKIPPED (>{int(finder.bwt.text_arr.size):,} bp, "
                  "too expensive)")
        return []
    arrays = short_imperfect_arrays(finder, tier1_seen)
    return records_from_arrays(finder.bwt.text_arr, arrays, chromosome)
