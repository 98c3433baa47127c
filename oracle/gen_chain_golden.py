#!/usr/bin/env python3
"""Generate tests/golden/chain_60k.json: the UNMODIFIED reference's post-processing chain
(bwt.py:3189-3614 -- nested-call suppression, dedup, adjacent merge with per-copy DP re-alignment,
refine, coordinate restore, overlap collapse, final filter) run on the strict-scan calls of a
60 kb planted contig.  Build-container only (needs /root/reference).  The raw calls are the oracle's
strict-scan rows (pinned to the reference elsewhere); they are stored in the fixture so that the test
neither needs the reference nor the slow CPU scan."""
import dataclasses
import hashlib
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

N, SEED, FLANK = 60_000, 33, 30
# (file, generator, n, seed): the Appendix B contig, and an indel-rich one with arrays close together
FIXTURES = [("chain_60k.json", "gen_contig", 60_000, 33), ("chain_indel_30k.json", "gen_contig_indel", 30_000, 7)]


def digest(records) -> str:
    return hashlib.sha256("\n".join(repr(dataclasses.astuple(r)) for r in records).encode()).hexdigest()


def main():
    for name, gen, n, seed in FIXTURES:
        if len(sys.argv) < 2 or sys.argv[1] == name:
            one(name, gen, n, seed)


def one(name, gen, N, SEED):
    import numpy as np

    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200 import finders
    from oracle import oracle as orc
    from tests import util

    spec = importlib.util.spec_from_file_location("refbwt", os.path.join(REF_DIR, "bwt.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    arr = getattr(util, gen)(N, SEED)
    full = arr.tobytes().decode()
    seq = full[FLANK:len(full) - FLANK]                      # what load_reference hands to the worker
    trimmed = np.frombuffer(seq.encode(), np.uint8)
    rows = orc.strict_scan(seq.encode() + b"$", 1, 1000, 0, 3)
    raw = [ref.TandemRepeat(**dataclasses.asdict(r)) for r in finders.strict_records(trimmed, rows, "synC", 0)]
    f = ref.TandemRepeatFinder("unused.fa")
    f.sequences, f.full_sequences, f.trim_offsets = {"synC": seq}, {"synC": full}, {"synC": FLANK}
    kept = f._suppress_nested_short_calls(raw, overlap_threshold=0.5)
    unique = f._deduplicate_repeats(kept)
    merged = f._merge_adjacent_repeats(unique)
    refined = f._refine_repeats(merged)
    f._restore_reference_coordinates(refined)
    refined = f._collapse_overlapping_repeats(refined)
    final = [r for r in refined if r.copies >= f.min_copies and r.length >= 6]
    final.sort(key=f._repeat_sort_key)
    out = {"n": N, "seed": SEED, "flank": FLANK, "chrom": "synC", "generator": gen,
           "rows": rows[:, :4].tolist(),
           "stage_counts": [len(raw), len(kept), len(unique), len(merged), len(refined), len(final)],
           "digest": digest(final), "digest_before_filter": digest(refined),
           "bed": [r.to_bed() for r in final],
           "imperfect": sum(1 for r in final if r.mismatch_rate > 0),
           "with_indel_notes": sum(1 for r in final if r.variations and any("ins(" in v or "del(" in v for v in r.variations))}
    path = os.path.join(ROOT, "tests", "golden", name)
    with open(path, "w") as fh:
        json.dump(out, fh)
    print(path, os.path.getsize(path), out["stage_counts"], out["imperfect"], out["with_indel_notes"], out["digest"][:16])


if __name__ == "__main__":
    main()
