#!/usr/bin/env python3
"""Generate tests/golden/align_walk.json: the UNMODIFIED reference's MotifUtils.align_repeat_region
(bwt.py:997-1102, with _align_unit_to_window 828-995) on the seeded cases of tests/align_cases.py -- per case the
fields TandemRepeatFinder._recompute_repeat reads (consensus, copies, consumed length, mismatch rate, worst copy,
variation notes, inserted / deleted bases) or null.  Build-container only (needs /root/reference); the test
regenerates the inputs from the same seed and holds the Python port and the native walk
(bwtk_align_repeat_region) to these answers."""
import importlib.util
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
sys.dont_write_bytecode = True
REF_DIR = os.environ.get("BWT_REFERENCE_DIR", "/root/reference")


def main():
    from tests.align_cases import N_CASES, cases

    spec = importlib.util.spec_from_file_location("refbwt", os.path.join(REF_DIR, "bwt.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = []
    for seq, start, end, template, frac, max_indel, min_copies in cases():
        s = ref.MotifUtils.align_repeat_region(seq, start, end, template, frac, max_indel, min_copies)
        out.append(None if s is None else [s.consensus, s.copies, s.consumed_length, s.mismatch_rate,
                                           s.max_errors_per_copy, list(s.variations), s.total_insertions,
                                           s.total_deletions])
    assert len(out) == N_CASES
    path = os.path.join(ROOT, "tests", "golden", "align_walk.json")
    with open(path, "w") as fh:
        json.dump({"cases": N_CASES, "summaries": sum(o is not None for o in out), "answers": out}, fh,
                  separators=(",", ":"))
    print(f"wrote {path}: {N_CASES} cases, {sum(o is not None for o in out)} summaries, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
