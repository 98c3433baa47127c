"""Host-side glue that was restructured for scale must keep the reference's results."""
import numpy as np
import pytest

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


def _mutated_tandem(rng, k, copies, rate):
    motif = "".join("ACGT"[x] for x in rng.integers(0, 4, k))
    out = []
    for _ in range(copies):
        unit = list(motif)
        r = rng.random()
        if r < rate:
            unit[int(rng.integers(0, k))] = "ACGT"[int(rng.integers(0, 4))]
        elif r < 1.5 * rate:
            unit.insert(int(rng.integers(0, k + 1)), "ACGT"[int(rng.integers(0, 4))])
        elif r < 2 * rate and k > 1:
            del unit[int(rng.integers(0, k))]
        out.append("".join(unit))
    return motif, "".join(out)


def test_exact_copy_fast_path_equals_dp_walk():
    """align_repeat_region skips the banded DP for copies equal to the running consensus; the
    summary (consensus, variations, error counts, consumed span) must not change."""
    rng = np.random.default_rng(17)
    checked = hits = 0
    for trial in range(400):
        k = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 9, 12, 20, 33]))
        motif, body = _mutated_tandem(rng, k, int(rng.integers(2, 30)), float(rng.choice([0.0, 0.1, 0.3])))
        left = "".join("ACGTN"[x] for x in rng.integers(0, 5, int(rng.integers(0, 40))))
        right = "".join("ACGT"[x] for x in rng.integers(0, 4, int(rng.integers(0, 40))))
        seq = left + body + right
        start = len(left) + int(rng.integers(0, 3))
        end = len(left) + len(body) - int(rng.integers(0, 3))
        template = seq[start:start + k] if rng.random() < 0.8 else motif
        for min_copies in (3, 1):
            MotifUtils._EXACT_COPY_FAST_PATH = False
            try:
                want = MotifUtils.align_repeat_region(seq, start, end, template, 0.1, min_copies=min_copies)
            finally:
                MotifUtils._EXACT_COPY_FAST_PATH = True
            got = MotifUtils.align_repeat_region(seq, start, end, template, 0.1, min_copies=min_copies)
            assert got == want, (trial, seq, start, end, template)
            checked += 1
            hits += bool(want and want.copies > 3 and 0 in want.error_counts and any(want.error_counts))
    assert checked == 800 and hits > 50   # mixed exact / inexact walks were exercised


def _walk_fields(summary):
    if summary is None:
        return None
    return (summary.consensus, summary.copies, summary.consumed_length, summary.mismatch_rate,
            summary.max_errors_per_copy, summary.variations, summary.total_insertions, summary.total_deletions)


def test_native_walk_equals_python_walk():
    """bwtk_align_repeat_region (csrc/rowchain.cu) is MotifUtils.align_repeat_region answer for answer: consensus,
    copies, consumed span, mismatch rate, worst copy, variation notes, inserted / deleted bases -- on mutated
    tandem arrays over several alphabets, with templates, spans, indel bands and copy floors of every kind."""
    import random

    from bwt_algorithm_b200.motifs import align_region_native

    rng = random.Random(41)

    def mk(alpha, n):
        return "".join(rng.choice(alpha) for _ in range(n))

    cases = summaries = with_sub = with_ins = with_del = 0
    for _ in range(1200):
        alpha = rng.choice(["ACGT", "ACGT", "AC", "ACGTN", "ACGTNRYK"])
        k = rng.choice([1, 1, 2, 3, 4, 5, 6, 7, 9, 12, 13, 17, 30, 64, 150])
        motif = mk(alpha, k)
        parts = [mk(alpha, rng.randrange(0, 40))]
        for _ in range(rng.randrange(1, 4)):
            arr = list(motif * rng.randrange(1, 12))
            for _ in range(rng.randrange(0, max(1, len(arr) // 8) + 1)):
                p, r = rng.randrange(len(arr)), rng.random()
                if r < 0.5:
                    arr[p] = rng.choice(alpha)
                elif r < 0.75:
                    arr.insert(p, rng.choice(alpha))
                elif len(arr) > 1:
                    del arr[p]
            parts += ["".join(arr), mk(alpha, rng.randrange(0, 30))]
        seq = "".join(parts)
        seq_bytes = seq.encode("ascii")
        for _ in range(5):
            start = rng.randrange(-2, len(seq))
            end = rng.choice([0, start, start + rng.randrange(1, 200), len(seq) + 5])
            kk = rng.choice([k, k, max(1, k - 1), k + 1])
            template = seq[max(0, start):max(0, start) + kk] if rng.random() < 0.8 else mk(alpha, kk)
            if not template:
                continue
            min_copies = rng.choice([1, 2, 3, 3, 5])
            max_indel = rng.choice([None, None, 0, 1, 3])
            frac = rng.choice([0.1, 0.1, 0.0, 0.25])
            want = _walk_fields(MotifUtils.align_repeat_region(seq, start, end, template, frac, max_indel, min_copies))
            got = align_region_native(seq_bytes, start, end, template.encode("ascii"), frac, max_indel, min_copies)
            assert got is not NotImplemented
            assert got == want, (seq, start, end, template, frac, max_indel, min_copies)
            cases += 1
            if want is not None:
                summaries += 1
                with_sub += any(">" in v for v in want[5])
                with_ins += any("ins(" in v for v in want[5])
                with_del += any("del(" in v for v in want[5])
    assert cases > 5000 and summaries > 1500 and min(with_sub, with_ins, with_del) > 40


def test_both_walks_equal_the_reference_on_its_own_answers():
    """tests/golden/align_walk.json (oracle/gen_align_golden.py): the UNMODIFIED reference's
    MotifUtils.align_repeat_region on 2 400 seeded cases.  The Python port and the native walk give the same
    consensus, copy count, consumed span, mismatch rate, worst copy, variation notes and indel totals (or None)."""
    import json
    import os

    from bwt_algorithm_b200.motifs import align_region_native
    from tests.align_cases import cases

    with open(os.path.join(os.path.dirname(__file__), "golden", "align_walk.json")) as fh:
        g = json.load(fh)
    n = notes = 0
    for (seq, start, end, template, frac, max_indel, min_copies), want in zip(cases(), g["answers"]):
        want = None if want is None else tuple(want)
        port = _walk_fields(MotifUtils.align_repeat_region(seq, start, end, template, frac, max_indel, min_copies))
        assert port == want, (seq, start, end, template, frac, max_indel, min_copies)
        native = align_region_native(seq.encode("ascii"), start, end, template.encode("ascii"), frac, max_indel,
                                     min_copies)
        assert native is not NotImplemented and native == want, (seq, start, end, template, frac, max_indel, min_copies)
        n += 1
        notes += bool(want and want[5])
    assert n == g["cases"] == 2400 and notes > 300


def test_native_walk_declines_what_it_cannot_represent():
    """More than eight distinct symbols in one consensus column: the native walk says NotImplemented and
    _recompute_repeat takes the Python walk -- same record either way; a non-ASCII contig never reaches it."""
    from bwt_algorithm_b200.motifs import align_region_native

    column = "ABCDEFGHIJ"
    seq = "".join(c + "CGT" for c in column) + "ACGT" * 3
    assert align_region_native(seq.encode(), 0, len(seq), b"ACGT", 0.25, None, 1) is NotImplemented
    for text in (seq, "ACGÅ" * 6 + "TTTT"):
        recs = []
        for native in (True, False):
            f = TandemRepeatFinder("/dev/null")
            f._NATIVE_WALK = native
            f.sequences = {"c": text}
            recs.append(f._recompute_repeat("c", 0, len(text), 4))
        assert recs[0] == recs[1]


def test_recompute_repeat_native_and_python_walks_agree():
    rng = np.random.default_rng(29)
    pieces = []
    spans = []
    pos = 0
    for i in range(400):
        gap = "".join("ACGTN"[x] for x in rng.integers(0, 5, int(rng.integers(0, 30))))
        k = int(rng.choice([1, 2, 3, 4, 5, 6, 8, 11, 16, 40]))
        _, body = _mutated_tandem(rng, k, int(rng.integers(2, 25)), float(rng.choice([0.0, 0.05, 0.2])))
        pieces += [gap, body]
        pos += len(gap)
        spans.append((pos, pos + len(body), k))
        pos += len(body)
    seq = "".join(pieces)
    on, off = TandemRepeatFinder("/dev/null"), TandemRepeatFinder("/dev/null")
    off._NATIVE_WALK = False
    on.sequences = off.sequences = {"c": seq}
    imperfect = 0
    for a, b, k in spans:
        for kk in (k, max(1, k - 1), 2 * k):
            got, want = on._recompute_repeat("c", a, b, kk, tier_hint=2), off._recompute_repeat("c", a, b, kk, tier_hint=2)
            assert got == want, (a, b, kk)
            imperfect += bool(want.variations)
    assert imperfect > 100


def test_merge_hands_over_the_union_record():
    """_merge_adjacent_repeats re-uses the union derived by _should_merge_repeats when the
    arguments coincide; same merged records as two independent derivations."""
    rng = np.random.default_rng(23)
    pieces, recs, pos = [], [], 0
    for i in range(300):
        gap = "".join("ACGT"[x] for x in rng.integers(0, 4, int(rng.integers(0, 6)) if i % 2 else 40))
        k = int(rng.integers(1, 8))
        motif, body = _mutated_tandem(rng, k, int(rng.integers(3, 12)), 0.0)
        if i % 2 and recs:                       # a second block of the previous motif, after a short gap
            motif = recs[-1].motif
            k = len(motif)
            body = motif * int(rng.integers(3, 9))
        pieces.append(gap)
        pos += len(gap)
        recs.append(TandemRepeat(chrom="c", start=pos, end=pos + len(body), motif=motif, copies=len(body) / k,
                                 length=len(body), tier=2, confidence=0.95, consensus_motif=motif))
        pieces.append(body)
        pos += len(body)
    seq = "".join(pieces)

    def finder():
        f = TandemRepeatFinder("/dev/null")
        f.sequences = {"c": seq}
        return f

    fast = finder()._merge_adjacent_repeats(list(recs))
    plain = finder()
    plain._merge_repeats = lambda r1, r2: plain._recompute_repeat(
        r1.chrom, min(r1.start, r2.start), max(r1.end, r2.end), len(r1.consensus_motif or r1.motif),
        tier_hint=min(r1.tier, r2.tier))
    slow = plain._merge_adjacent_repeats(list(recs))
    assert fast == slow and len(fast) < len(recs)
    assert len({id(r) for r in fast}) == len(fast)


def test_canonical_motif_memo_equals_direct_form():
    from bwt_algorithm_b200.motifs import _canonical_stranded

    rng = np.random.default_rng(2)
    motifs = ["", "A", "T", "AC", "GT", "ACGT", "TTAGGG", "CCCTAA", "NNA", "acgt", "AXT"]
    motifs += ["".join("ACGTN"[x] for x in rng.integers(0, 5, int(rng.integers(1, 40)))) for _ in range(300)]
    for m in motifs + motifs:                      # second round is served from the cache
        got = MotifUtils.get_canonical_motif_stranded(m)
        if m:
            assert got == _canonical_stranded(m), m
            fwd = min(m[i:] + m[:i] for i in range(len(m)))
            rc = MotifUtils.reverse_complement(m)
            rev = min(rc[i:] + rc[:i] for i in range(len(rc)))
            assert got == ((fwd, "+") if fwd <= rev else (rev, "-")), m
        else:
            assert got == ("", "+")


@pytest.mark.parametrize("fixture", ["chain_60k.json", "chain_indel_30k.json"])
def test_postprocessing_chain_equals_reference_golden(fixture):
    """tests/golden/chain_60k.json, chain_indel_30k.json (oracle/gen_chain_golden.py): the unmodified reference's
    chain on the strict-scan calls of a 60 kb planted contig (3 166 calls) and of a 30 kb contig whose arrays carry
    insertions and deletions and lie close together (1 749 calls, 28 final calls with ins/del notes) -- stage
    counts, every field of every surviving record (digest of the dataclass tuples, before and after the final
    filter) and the BED lines."""
    import dataclasses
    import hashlib
    import json
    import os

    from tests import util

    with open(os.path.join(os.path.dirname(__file__), "golden", fixture)) as fh:
        g = json.load(fh)
    full = getattr(util, g["generator"])(g["n"], g["seed"]).tobytes().decode()
    seq = full[g["flank"]:len(full) - g["flank"]]
    rows = np.zeros((len(g["rows"]), 8), np.int32)
    rows[:, :4] = np.array(g["rows"], np.int32)
    raw = finders.strict_records(np.frombuffer(seq.encode(), np.uint8), rows, g["chrom"], 0)
    f = TandemRepeatFinder("/dev/null")
    f.sequences, f.full_sequences, f.trim_offsets = {g["chrom"]: seq}, {g["chrom"]: full}, {g["chrom"]: g["flank"]}
    kept = f._suppress_nested_short_calls(raw, overlap_threshold=0.5)
    unique = f._deduplicate_repeats(kept)
    merged = f._merge_adjacent_repeats(unique)
    refined = f._refine_repeats(merged)
    f._restore_reference_coordinates(refined)
    refined = f._collapse_overlapping_repeats(refined)
    final = [r for r in refined if r.copies >= f.min_copies and r.length >= 6]
    final.sort(key=f._repeat_sort_key)
    assert [len(raw), len(kept), len(unique), len(merged), len(refined), len(final)] == g["stage_counts"]
    digest = lambda recs: hashlib.sha256("\n".join(repr(dataclasses.astuple(r)) for r in recs).encode()).hexdigest()
    assert [r.to_bed() for r in final] == g["bed"]
    assert digest(refined) == g["digest_before_filter"]
    assert digest(final) == g["digest"]
    # and the packaged entry point gives the same final list
    again, dups = f._postprocess(finders.strict_records(np.frombuffer(seq.encode(), np.uint8), rows, g["chrom"], 0),
                                 lambda msg: None)
    assert digest(again) == g["digest"] and dups == g["stage_counts"][0] - g["stage_counts"][2]
    # the row-oriented chain (rowchain.finish_rows: records only for what survives) gives the same list
    from bwt_algorithm_b200 import rowchain

    rfinal, n_raw, n_kept, n_unique = rowchain.finish_rows(f, g["chrom"], np.frombuffer(seq.encode(), np.uint8), rows)
    assert [n_raw, n_kept, n_unique] == g["stage_counts"][:3]
    assert digest(rfinal) == g["digest"] and [r.to_bed() for r in rfinal] == g["bed"]
    # ... also when the contig's rows are cut into independent pieces (what the CLI's CPU workers are handed) ...
    text = np.frombuffer(seq.encode(), np.uint8)
    for pieces, min_rows in ((4, 300), (40, 20), (1000, 3)):
        ran = []

        def run_pieces(parts, ran=ran):
            ran.append(len(parts))
            return [rowchain.finish_segment(f, g["chrom"], text, p) for p in parts]

        sfinal, a, b, c = rowchain.finish_rows_segmented(f, g["chrom"], text, rows, run_pieces, pieces, min_rows)
        assert [a, b, c] == g["stage_counts"][:3] and ran and ran[0] > 1
        assert digest(sfinal) == g["digest"]
    # ... and with the re-alignment walk in Python instead of native code (the default above)
    slow = TandemRepeatFinder("/dev/null")
    slow._NATIVE_WALK = False
    slow.sequences, slow.full_sequences, slow.trim_offsets = f.sequences, f.full_sequences, f.trim_offsets
    assert digest(rowchain.finish_rows(slow, g["chrom"], text, rows)[0]) == g["digest"]


def test_per_contig_finish_equals_global_chain(oracle, tmp_path):
    """The multi-process path lets every worker run the chain on its own contigs
    (pipeline._finish_contig) and sorts the concatenation; same calls, same counts as the chain on the
    mixed list, on the reference's 12-contig fixture and on planted contigs with flank trimming."""
    import dataclasses
    import os
    import pickle

    from bwt_algorithm_b200 import pipeline
    from tests.util import gen_contig

    planted = tmp_path / "planted.fa"
    with open(planted, "w") as fh:
        for i, n in enumerate((12_000, 5_000, 50, 8_000)):
            fh.write(f">ctg{10 - i} planted\n{gen_contig(n, 40 + i).tobytes().decode()}\n")
    fixture = os.path.join(os.path.dirname(__file__), "golden", "cli", "test2.fa")
    for path, trim in ((fixture, 30), (str(planted), 30), (str(planted), 0)):
        finder = TandemRepeatFinder(path, flank_trim=trim)
        seqs = finder.load_reference()
        assert len(seqs) >= 4 and pipeline._distinct_sort_names(seqs)
        cfg = finder._config(True, True)
        raw_of = {}
        for chrom, seq in seqs.items():
            text = (seq + "$").encode()
            rows = oracle.strict_scan(text, 1, max(120, min(len(seq) // 3, 1000)), 0, 3)
            raw_of[chrom] = lambda text=text, rows=rows, chrom=chrom: finders.strict_records(
                np.frombuffer(text, np.uint8), rows, chrom, 0)
        mixed = [r for chrom in seqs for r in raw_of[chrom]()]
        want, n_raw, n_kept, n_unique = finder._postprocess_counts(mixed)
        got, sums = [], [0, 0, 0]
        tasks = finder._finish_tasks([(chrom, seq, cfg) for chrom, seq in seqs.items()])
        for chrom, seq, c, left, right in reversed(tasks):          # workers finish in any order
            assert left + seq + right == finder.full_sequences[chrom] and len(left) == finder.trim_offsets[chrom]
            part, a, b, c3 = pipeline._finish_contig(chrom, seq, c, left, right, raw_of[chrom]())
            got.extend(part)
            sums = [sums[0] + a, sums[1] + b, sums[2] + c3]
        got.sort(key=finder._repeat_sort_key)
        assert sums == [n_raw, n_kept, n_unique]
        assert [dataclasses.astuple(r) for r in got] == [dataclasses.astuple(r) for r in want]
        assert len(want) > 0
        # ... and so does the row-oriented chain, which never builds the records that do not survive
        from bwt_algorithm_b200 import rowchain

        got2, sums2 = [], [0, 0, 0]
        for chrom, seq, c, left, right in tasks:
            text = (seq + "$").encode()
            rows = oracle.strict_scan(text, 1, max(120, min(len(seq) // 3, 1000)), 0, 3)
            part, a, b, c3 = rowchain.finish_rows(pipeline._contig_finder(chrom, seq, c, left, right), chrom,
                                                  np.frombuffer(text, np.uint8), rows)
            got2.extend(part)
            sums2 = [sums2[0] + a, sums2[1] + b, sums2[2] + c3]
        got2.sort(key=finder._repeat_sort_key)
        assert sums2 == [n_raw, n_kept, n_unique]
        assert [dataclasses.astuple(r) for r in got2] == [dataclasses.astuple(r) for r in want]
    pickle.dumps(pipeline._process_and_finish_contig)                # spawn-able: top level, importable
    assert not pipeline._distinct_sort_names(["chr1", "Chr1"])       # equal sort keys: the global chain is used


def test_drivers_use_the_row_chain_and_match_the_object_chain(oracle, monkeypatch, capsys):
    """find_tandem_repeats / find_tandem_repeats_parallel (detection served by the oracle) finish every contig
    at row level; the calls and the summary lines are those of the object chain on the mixed list."""
    import dataclasses
    import os

    from bwt_algorithm_b200 import pipeline, sharding

    def fake_rows(chrom, seq, config):
        text = (seq + "$").encode()
        rows = oracle.strict_scan(text, 1, max(config["max_unit_len"], min(len(seq) // config["min_copies"], 1000)),
                                  0, config["min_copies"])
        return rows, np.frombuffer(text, np.uint8)

    monkeypatch.setattr(pipeline, "_detect_rows", fake_rows)
    fixture = os.path.join(os.path.dirname(__file__), "golden", "cli", "test2.fa")
    ref = TandemRepeatFinder(fixture)
    seqs = ref.load_reference()
    mixed = []
    for chrom, seq in seqs.items():
        rows, text_arr = fake_rows(chrom, seq, ref._config(False, True))
        mixed.extend(finders.strict_records(text_arr, rows, chrom, 0))
    want, n_raw, n_kept, n_unique = ref._postprocess_counts(mixed)
    assert n_kept < n_raw and len(want) > 0
    want_t = [dataclasses.astuple(r) for r in want]
    for procs in (1, 2):
        monkeypatch.setattr(sharding, "worker_processes", lambda n_tasks, n_jobs, procs=procs: procs)
        monkeypatch.setattr(sharding, "run_tasks", lambda worker, tasks, n_jobs: (worker(t) for t in tasks))
        finder = TandemRepeatFinder(fixture)
        finder.load_reference()
        final = finder.find_tandem_repeats_parallel(enable_tier1=False, enable_tier2=True, n_jobs=4)
        lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith(("Nested call", "Analysis complete"))]
        assert [dataclasses.astuple(r) for r in final] == want_t
        assert lines == [f"Nested call suppression: {n_raw} -> {n_kept} repeats",
                         f"Analysis complete! Found {len(want)} unique repeats"
                         + (f" (deduplicated {n_raw - n_unique})." if n_raw - n_unique > 0 else ".")]
    finder = TandemRepeatFinder(fixture)
    finder.load_reference()
    final = finder.find_tandem_repeats(enable_tier1=False, enable_tier2=True)
    out = capsys.readouterr().out
    assert [dataclasses.astuple(r) for r in final] == want_t
    assert f"Nested call suppression: {n_raw} -> {n_kept} repeats" in out
    assert f"Analysis complete! Found {len(want)} total repeats." in out
    # the worker of the reference API still returns records (bwt.py:3040-3141)
    recs = pipeline._process_chromosome_worker(("c", "ACGT" * 3 + "TTTTTTT" + "ACG" * 5, ref._config(False, True)))
    assert recs and all(isinstance(r, finders.TandemRepeat) for r in recs)


def test_chain_pool_finishes_contigs_in_cpu_workers(oracle, monkeypatch, capsys):
    """One GPU, several contigs, --jobs > 1: detection stays in the parent, every contig's chain runs in a spawned
    CPU worker (pipeline._run_with_chain_pool).  Same calls and summary lines as the in-process path."""
    import dataclasses
    import os

    from bwt_algorithm_b200 import pipeline, sharding

    def fake_rows(chrom, seq, config):
        text = (seq + "$").encode()
        rows = oracle.strict_scan(text, 1, max(config["max_unit_len"], min(len(seq) // config["min_copies"], 1000)),
                                  0, config["min_copies"])
        return rows, np.frombuffer(text, np.uint8)

    monkeypatch.setattr(pipeline, "_detect_rows", fake_rows)
    monkeypatch.setattr(sharding, "worker_processes", lambda n_tasks, n_jobs: 1)
    fixture = os.path.join(os.path.dirname(__file__), "golden", "cli", "test2.fa")
    outs = []
    for min_bases, jobs in ((1 << 40, 3), (0, 3), (0, 1)):
        monkeypatch.setattr(pipeline, "CHAIN_POOL_MIN_BASES", min_bases)
        finder = TandemRepeatFinder(fixture)
        seqs = finder.load_reference()
        assert len(seqs) >= 2
        tasks = finder._finish_tasks([(c, q, finder._config(False, True)) for c, q in seqs.items()])
        assert pipeline._chain_pool_size(tasks, jobs) == (min(jobs, len(seqs)) if min_bases == 0 and jobs > 1 else 0)
        final = finder.find_tandem_repeats_parallel(enable_tier1=False, enable_tier2=True, n_jobs=jobs)
        lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith(("Nested call", "Analysis complete"))]
        outs.append(([dataclasses.astuple(r) for r in final], lines))
    assert outs[0][0] and outs[0] == outs[1] == outs[2]
    # a contig whose detection fails is reported and skipped, the others are finished (bwt.py:3134-3141)
    monkeypatch.setattr(pipeline, "CHAIN_POOL_MIN_BASES", 0)
    first = next(iter(seqs))

    def flaky(chrom, seq, config):
        if chrom == first:
            raise RuntimeError("boom")
        return fake_rows(chrom, seq, config)

    monkeypatch.setattr(pipeline, "_detect_rows", flaky)
    finder = TandemRepeatFinder(fixture)
    finder.load_reference()
    final = finder.find_tandem_repeats_parallel(enable_tier1=False, enable_tier2=True, n_jobs=3)
    assert f"ERROR processing chromosome {first}: boom" in capsys.readouterr().out
    assert [dataclasses.astuple(r) for r in final] == [t for t in outs[0][0] if t[0] != first]


def test_row_chain_homopolymer_shortcut_equals_generic_path():
    """rowchain.finish_rows settles isolated homopolymer merge events on integers (FAST_K1); with the shortcut
    off every candidate goes through _should_merge_repeats / _recompute_repeat.  Same calls either way, on
    run-length rows of random and of low-complexity text (dense events, chains of events, contig ends)."""
    import dataclasses

    from bwt_algorithm_b200 import pipeline, rowchain

    cfg = {"sa_sample_rate": 32, "show_progress": False, "allow_mismatches": True, "max_motif_length": 9,
           "min_period": 10, "max_period": 1000, "min_copies": 3, "min_entropy": 1.0, "max_unit_len": 120}
    rng = np.random.default_rng(11)
    texts = {
        "random_300k": np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 300_000)],
        "at_rich": np.frombuffer(b"AATTAC", np.uint8)[rng.integers(0, 6, 120_000)],
        "two_letters": np.frombuffer(b"AT", np.uint8)[rng.integers(0, 2, 60_000)],
        "with_N": np.frombuffer(b"ACGTNN", np.uint8)[rng.integers(0, 6, 80_000)],
        "ends_in_runs": np.concatenate([np.full(7, 65, np.uint8), np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 5000)],
                                        np.full(4, 84, np.uint8), np.full(5, 65, np.uint8)]),
    }
    total_events = 0
    for name, s in texts.items():
        for trim in (0, 30):
            n = s.size
            d = np.flatnonzero(np.diff(s) != 0) + 1
            starts, ends = np.concatenate(([0], d)), np.concatenate((d, [n]))
            sel = (ends - starts >= 3) & (starts >= trim) & (ends <= n - trim)
            rows = np.zeros((int(sel.sum()), 8), np.int32)
            rows[:, 0], rows[:, 1], rows[:, 2] = starts[sel] - trim, ends[sel] - trim, 1
            rows[:, 3], rows[:, 6] = (ends - starts)[sel], 1
            # a few longer-unit rows on top (they take part in suppression and collapse)
            extra = rows[:: max(1, len(rows) // 50)].copy()
            extra[:, 1] = np.minimum(extra[:, 0] + 12, n - 2 * trim)
            extra[:, 2], extra[:, 3], extra[:, 6] = 4, 3, 4
            extra = extra[extra[:, 1] - extra[:, 0] == 12]
            rows = np.concatenate([rows, extra])
            seq = s.tobytes().decode()
            body = seq[trim:n - trim] if trim else seq
            text = np.frombuffer((body + "$").encode(), np.uint8)
            f = pipeline._contig_finder("c1", body, cfg, seq[:trim], seq[n - trim:] if trim else "")
            fast = rowchain.finish_rows(f, "c1", text, rows)
            slow = rowchain._finish_rows_slow(pipeline._contig_finder("c1", body, cfg, seq[:trim], seq[n - trim:] if trim else ""),
                                              "c1", text, rows)
            assert fast[1:] == slow[1:], name
            assert [dataclasses.astuple(r) for r in fast[0]] == [dataclasses.astuple(r) for r in slow[0]], name
            total_events += len(slow[0])
    assert total_events > 1000


def test_row_chain_requeues_an_undone_homopolymer_event():
    """A merge chain that reaches a settled homopolymer event undoes it; the event's own link is then a candidate
    again.  finish_rows picks it up next (it used to redo the whole contig on records, 8 s per chr21-sized contig).
    Rows: the oracle's strict scan of A-rich text, three seeds that each hit the case once."""
    import dataclasses

    from bwt_algorithm_b200 import pipeline, rowchain
    from oracle import oracle as orc

    cfg = {"sa_sample_rate": 32, "show_progress": False, "allow_mismatches": True, "max_motif_length": 9,
           "min_period": 10, "max_period": 1000, "min_copies": 3, "min_entropy": 1.0, "max_unit_len": 120}
    hits = 0
    for seed in (1, 3, 5):
        rng = np.random.default_rng(seed)
        seq = np.frombuffer(b"AAAT", np.uint8)[rng.integers(0, 4, 30_000)].tobytes().decode()
        text = np.frombuffer((seq + "$").encode(), np.uint8)
        rows = orc.strict_scan(text.tobytes(), 1, 12, 0, 3)
        rowchain.STATS["requeued"] = 0
        fast = rowchain.finish_rows(pipeline._contig_finder("c1", seq, cfg, "", ""), "c1", text, rows)
        hits += rowchain.STATS["requeued"]
        slow = rowchain._finish_rows_slow(pipeline._contig_finder("c1", seq, cfg, "", ""), "c1", text, rows)
        assert fast[1:] == slow[1:]
        assert [dataclasses.astuple(r) for r in fast[0]] == [dataclasses.astuple(r) for r in slow[0]]
    assert hits >= 3


def _strict_rows_numpy(s: np.ndarray, max_unit: int, min_copies: int = 3) -> np.ndarray:
    """Strict-scan rows (bwt.py:1891-2001 semantics for max_mismatch 0) by run lengths in NumPy -- fast enough for
    long texts with small unit caps (the oracle's loop is quadratic in the unit length).  Checked against the oracle
    in test_segmented_chain_equals_the_whole_chain."""
    out = []
    n = s.size
    for u in range(min(max_unit, n // min_copies), 0, -1):
        eq = s[:-u] == s[u:]
        d = np.diff(np.concatenate([[0], eq.view(np.int8), [0]]))
        a, b = np.flatnonzero(d == 1), np.flatnonzero(d == -1)
        keep = (b - a) >= (min_copies - 1) * u
        prev_end = -1
        for x, y in zip(a[keep].tolist(), b[keep].tolist()):
            e = x if x > prev_end else prev_end
            if e <= y - (min_copies - 1) * u:
                cnt = 1 + (y - e) // u
                prev_end = e + cnt * u
                unit = s[e:e + u].tobytes()
                prim = next((p for p in range(1, u // 2 + 1) if u % p == 0 and unit[p:] == unit[:-p]), u)
                out.append((e, prev_end, prim, (prev_end - e) // prim if prim < u else cnt, 0, 0, u, 0))
    return np.array(out, np.int32).reshape(-1, 8)


def test_segmented_chain_equals_the_whole_chain(oracle):
    """rowchain.finish_rows_segmented cuts a contig's rows where no row reaches the next one, finishes the pieces
    independently and checks afterwards that nothing a piece formed reached its neighbour.  Same calls, order and
    counts as finish_rows -- with pieces of a handful of rows, so that cuts fall next to merge chains, nested calls,
    collapsing calls and contig ends all the time (and the one-piece fallback is exercised too)."""
    import dataclasses

    from bwt_algorithm_b200 import pipeline, rowchain
    from tests.util import gen_contig

    cfg = {"sa_sample_rate": 32, "show_progress": False, "allow_mismatches": True, "max_motif_length": 9,
           "min_period": 10, "max_period": 1000, "min_copies": 3, "min_entropy": 1.0, "max_unit_len": 120}
    rng = np.random.default_rng(19)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    planted = gen_contig(60_000, 5)
    dense = gen_contig(30_000, 6, sub_rate=0.08)
    lowc = np.frombuffer(b"AATTAC", np.uint8)[rng.integers(0, 6, 40_000)]
    two = np.frombuffer(b"AT", np.uint8)[rng.integers(0, 2, 20_000)]
    drift = np.tile(np.frombuffer(b"AACAACAAGAAG", np.uint8), 1500)            # one long chain of merges (unit 3)
    drift[rng.integers(0, drift.size, 150)] = acgt[rng.integers(0, 4, 150)]
    texts = {"planted": (planted, 40), "dense": (dense, 40), "low_complexity": (lowc, 12), "two_letters": (two, 12),
             "drift": (drift, 24), "drift_short_units": (drift, 9)}
    assert np.array_equal(_strict_rows_numpy(planted[:6000], 40), oracle.strict_scan(planted[:6000].tobytes() + b"$", 1, 40, 0, 3))
    fallbacks = pieces_run = 0
    for name, (s, cap) in texts.items():
        rows = _strict_rows_numpy(s, cap)
        seq = s.tobytes().decode()
        text = np.frombuffer((seq + "$").encode(), np.uint8)
        whole = rowchain.finish_rows(pipeline._contig_finder("c1", seq, cfg, "", ""), "c1", text, rows)
        want = [dataclasses.astuple(r) for r in whole[0]]
        assert len(want) > 10, name
        for min_rows, pieces in ((3, 1000), (25, 64), (400, 7)):
            finder = pipeline._contig_finder("c1", seq, cfg, "", "")
            ran = []

            def run_pieces(parts, finder=finder, ran=ran):
                ran.append(len(parts))
                return [rowchain.finish_segment(finder, "c1", text, p) for p in parts]

            got = rowchain.finish_rows_segmented(finder, "c1", text, rows, run_pieces, pieces, min_rows)
            assert got[1:] == whole[1:], (name, min_rows)
            assert [dataclasses.astuple(r) for r in got[0]] == want, (name, min_rows)
            pieces_run += sum(ran)
            # the fallback is visible as a reach beyond a neighbour's first row; count it through a dry run
            order = np.argsort(rows[:, 0], kind="stable")
            bounds = rowchain.segment_bounds(rows, order, pieces, min_rows)
            if len(bounds) > 2:
                parts = [rows[np.sort(order[bounds[i]:bounds[i + 1]])] for i in range(len(bounds) - 1)]
                firsts = [int(rows[order[b], 0]) for b in bounds[1:-1]]
                reaches = [rowchain.finish_segment(finder, "c1", text, p)[4] for p in parts[:-1]]
                fallbacks += any(r >= f for r, f in zip(reaches, firsts))
    assert pieces_run > 500 and fallbacks >= 1


def test_segment_pool_finishes_one_contig_in_pieces(oracle, monkeypatch, capsys, tmp_path):
    """One large contig on one GPU with --jobs > 1: the contig's chain is cut into pieces for spawned CPU workers
    (pipeline._run_with_segment_pool -> _finish_rows_in_pool -> _finish_piece).  Same calls and summary lines as the
    in-process chain; the text reaches the workers through a temporary file that is gone afterwards."""
    import dataclasses
    import glob
    import os
    import tempfile

    from bwt_algorithm_b200 import pipeline, rowchain, sharding
    from tests.util import gen_contig

    s = gen_contig(90_000, 31)
    fa = tmp_path / "one.fa"
    fa.write_text(">big one contig\n" + s.tobytes().decode() + "\n")

    def fake_rows(chrom, seq, config):
        text = np.frombuffer((seq + "$").encode(), np.uint8)
        return _strict_rows_numpy(text[:-1], 60, config["min_copies"]), text

    monkeypatch.setattr(pipeline, "_detect_rows", fake_rows)
    monkeypatch.setattr(sharding, "worker_processes", lambda n_tasks, n_jobs: 1)
    outs = []
    for min_bases, min_rows, jobs in ((1 << 40, 1 << 40, 3), (0, 200, 3)):
        monkeypatch.setattr(pipeline, "SEGMENT_POOL_MIN_BASES", min_bases)
        monkeypatch.setattr(rowchain, "SEGMENT_MIN_ROWS", min_rows)
        used = []
        real = pipeline._finish_rows_in_pool
        monkeypatch.setattr(pipeline, "_finish_rows_in_pool", lambda *a, **k: used.append(1) or real(*a, **k))
        finder = TandemRepeatFinder(str(fa))
        finder.load_reference()
        final = finder.find_tandem_repeats_parallel(enable_tier1=False, enable_tier2=True, n_jobs=jobs)
        lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith(("Nested call", "Analysis complete"))]
        outs.append(([dataclasses.astuple(r) for r in final], lines))
        assert bool(used) == (min_bases == 0)
        monkeypatch.setattr(pipeline, "_finish_rows_in_pool", real)
    assert len(outs[0][0]) > 100 and outs[0] == outs[1]
    assert not glob.glob(os.path.join(tempfile.gettempdir(), "bwtk_chain_*"))
