"""Host-side glue that was restructured for scale must keep the reference's results."""
import numpy as np

import bwt_algorithm_b200  # noqa: F401
from bwt_algorithm_b200 import finders
from bwt_algorithm_b200.motifs import MotifUtils
from bwt_algorithm_b200.pipeline import TandemRepeatFinder
from bwt_algorithm_b200.records import TandemRepeat, _natural_sort_key, _natural_sort_key_uncached


def _quadratic_suppression(repeats, overlap_threshold=0.5):
    """bwt.py:3402-3497 as written: every call against every kept span."""
    by_chrom = {}
    for r in repeats:
        by_chrom.setdefault(r.chrom, []).append(r)
    survivors = []
    for chrom_repeats in by_chrom.values():
        ordered = sorted(chrom_repeats, key=lambda r: (r.mismatch_rate > 0, -len(r.motif)))
        spans = []
        for r in ordered:
            span = r.end - r.start
            k = len(r.motif)
            nested = False
            for o_start, o_end, o_k in spans:
                if o_k <= k:
                    continue
                ov = max(0, min(r.end, o_end) - max(r.start, o_start))
                if ov == 0:
                    continue
                if k == 1 and o_k > 1 and ov / span >= 0.8:
                    nested = True
                    break
                ratio = o_k / k
                cut = 0.1 if ratio >= 10 else 0.3 if ratio >= 5 else overlap_threshold
                if ov / span >= cut:
                    nested = True
                    break
            if not nested:
                survivors.append(r)
                spans.append((r.start, r.end, k))
    survivors.sort(key=TandemRepeatFinder._repeat_sort_key)
    return survivors


def test_bucketed_suppression_equals_quadratic_reference():
    rng = np.random.default_rng(4)
    for trial in range(6):
        recs = []
        for i in range(1500):
            k = int(rng.choice([1, 1, 2, 3, 4, 6, 10, 12, 30, 60, 200]))
            span = int(k * rng.integers(3, 40)) if rng.random() < 0.9 else int(rng.integers(2000, 9000))
            start = int(rng.integers(0, 40_000))
            mm = 0.0 if rng.random() < 0.7 else 0.05
            motif = "".join("ACGT"[x] for x in rng.integers(0, 4, k))
            recs.append(TandemRepeat(chrom=f"chr{i % 3}", start=start, end=start + span, motif=motif,
                                     copies=span / k, length=span, tier=2, confidence=0.9, mismatch_rate=mm))
        finder = TandemRepeatFinder("/dev/null")
        got = finder._suppress_nested_short_calls(list(recs), overlap_threshold=0.5)
        want = _quadratic_suppression(list(recs), 0.5)
        assert [id(r) for r in got] == [id(r) for r in want], f"trial {trial}"
        assert 0 < len(got) < len(recs)


def test_strict_records_match_the_per_record_formula():
    rng = np.random.default_rng(8)
    text = np.frombuffer(("".join("ACGTN"[x] for x in rng.integers(0, 5, 5000)) + "$").encode(), np.uint8)
    rows = []
    for _ in range(400):
        prim = int(rng.integers(1, 12))
        count = int(rng.integers(3, 9))
        start = int(rng.integers(0, 4800))
        rows.append([start, start + prim * count, prim, count, 0, 0, prim, 0])
    rows = np.array(rows, np.int32)
    got = finders.strict_records(text, rows, "c", 0)
    for rec, (start, end, prim, count) in zip(got, rows[:, :4].tolist()):
        motif = text[start:start + prim].tobytes().decode("ascii", errors="replace")
        pm, pi, score, comp, ent, actual = MotifUtils.calculate_trf_statistics(text, start, end, motif, count, 0.0)
        assert (rec.motif, rec.consensus_motif, rec.percent_matches, rec.percent_indels, rec.score) == \
            (motif, motif, pm, pi, score)
        assert rec.composition == comp and rec.entropy == ent and rec.actual_sequence == actual
        assert (rec.start, rec.end, rec.copies, rec.length, rec.max_mismatches_per_copy) == \
            (start, end, float(count), end - start, 0)
    # records own their composition dict
    got[0].composition["A"] = -1.0
    assert all(r.composition["A"] >= 0.0 for r in got[1:])


def test_natural_sort_key_memo():
    for v in ("chr10", "chr2", "Chr2_random", None, "7", "scaffold_12_3"):
        assert _natural_sort_key(v) == _natural_sort_key_uncached(v)
    assert sorted(["chr10", "chr2", "chr1"], key=_natural_sort_key) == ["chr1", "chr2", "chr10"]
