"""Shared test helpers: the synthetic contig generator of SURVEY.md Appendix B
and the adversarial text list."""
import numpy as np


def gen_contig(n, seed, sub_rate=0.03):
    rng = np.random.default_rng(seed)
    A = np.frombuffer(b"ACGT", dtype=np.uint8)
    s = A[rng.integers(0, 4, n)]
    pos = 500
    while pos < n - 400:
        k = int(rng.integers(1, 10))
        m = A[rng.integers(0, 4, k)]
        c = int(rng.integers(4, 25))
        arr = np.tile(m, c)
        mut = rng.random(arr.size) < sub_rate
        arr[mut] = A[rng.integers(0, 4, int(mut.sum()))]
        s[pos:pos + arr.size] = arr
        pos += arr.size + int(rng.integers(300, 1500))
    return s


def gen_contig_indel(n, seed, sub_rate=0.05, indel_rate=0.02, gap=(20, 400)):
    """Planted arrays of 1-12 bp motifs with substitutions AND single-base insertions / deletions, close together:
    dense merge chains, re-alignments with ins/del notes, collapsing calls."""
    rng = np.random.default_rng(seed)
    A = np.frombuffer(b"ACGT", dtype=np.uint8)
    s = A[rng.integers(0, 4, n)]
    pos = 200
    while pos < n - 400:
        k = int(rng.integers(1, 13))
        m = A[rng.integers(0, 4, k)]
        arr = np.tile(m, int(rng.integers(4, 30)))
        mut = rng.random(arr.size) < sub_rate
        arr[mut] = A[rng.integers(0, 4, int(mut.sum()))]
        for _ in range(int(rng.binomial(arr.size, indel_rate))):
            at = int(rng.integers(1, max(2, arr.size - 1)))
            arr = np.insert(arr, at, A[rng.integers(0, 4)]) if rng.random() < 0.5 else np.delete(arr, at)
        end = min(n, pos + arr.size)
        s[pos:end] = arr[:end - pos]
        pos = end + int(rng.integers(gap[0], gap[1]))
    return s


def text_cases():
    """(name, text bytes) -- texts as BWTCore receives them (caller appends '$')."""
    rng = np.random.default_rng(99)
    rnd = lambda n: bytes(b"ACGT"[x] for x in rng.integers(0, 4, n))
    cs = [
        ("one_sentinel", b"$"),
        ("one_base", b"G$"),
        ("n7", b"ACGTAC$"),
        ("n8", b"ACGTACG$"),
        ("n17", rnd(16) + b"$"),
        ("n33", rnd(32) + b"$"),
        ("n128", rnd(127) + b"$"),
        ("n129", rnd(128) + b"$"),
        ("n4097", rnd(4096) + b"$"),
        ("allA_2000", b"A" * 2000 + b"$"),
        ("AC_period2", b"AC" * 700 + b"A$"),
        ("CAG_runs", b"CAG" * 120 + b"T" + b"CAG" * 60 + b"$"),
        ("N_runs", b"ACGT" * 10 + b"N" * 50 + b"ACGTTGCA" * 8 + b"N" * 9 + b"GATTACA" * 6 + b"$"),
        ("iupac", b"ACGTRYACGTACGTKMACGTACGTACGTSWACGTNNACGTACGTACGTACGT" * 3 + b"$"),
        ("lower", b"acgtacgtnnacgtACGTacgg" * 5 + b"$"),
        ("no_sentinel", rnd(500)),
        ("no_sentinel_allA", b"A" * 300),
        ("two_symbols", b"AB" * 100 + b"BBA" * 30),
        ("binaryish", bytes(int(x) for x in rng.integers(0, 256, 3000))),
        ("sentinel_in_middle", b"ACGT$ACGT$"),
        ("planted_3000", gen_contig(3000, 1).tobytes() + b"$"),
        ("planted_20000_N", (lambda s: (s.__setitem__(slice(7000, 9000), ord("N")), s)[1])(gen_contig(20000, 5)).tobytes() + b"$"),
        ("long_repeat", rnd(700) * 6 + b"$"),
    ]
    return cs
