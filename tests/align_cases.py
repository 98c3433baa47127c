"""Seeded inputs for the per-copy re-alignment walk (MotifUtils.align_repeat_region, bwt.py:997-1102): mutated tandem
arrays over several alphabets with start / end / template / indel band / copy floor of every kind.  Shared by
oracle/gen_align_golden.py (which runs the UNMODIFIED reference on them) and tests/test_host_glue_cpu.py (which runs
the Python port and the native walk on the same cases)."""
import random

N_CASES = 2400


def cases(seed: int = 20261019, count: int = N_CASES):
    """Yields (sequence, start, end, template, mismatch_fraction, max_indel, min_copies)."""
    rng = random.Random(seed)

    def mk(alpha, n):
        return "".join(rng.choice(alpha) for _ in range(n))

    made = 0
    while made < count:
        alpha = rng.choice(["ACGT", "ACGT", "AC", "ACGTN", "ACGTNRYK"])
        k = rng.choice([1, 1, 2, 3, 4, 5, 6, 7, 9, 12, 13, 17, 30, 64])
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
        for _ in range(4):
            start = rng.randrange(-2, len(seq))
            end = rng.choice([0, start, start + rng.randrange(1, 200), len(seq) + 5])
            kk = rng.choice([k, k, max(1, k - 1), k + 1])
            template = seq[max(0, start):max(0, start) + kk] if rng.random() < 0.8 else mk(alpha, kk)
            if not template:
                continue
            yield (seq, start, end, template, rng.choice([0.1, 0.1, 0.0, 0.25]), rng.choice([None, None, 0, 1, 3]),
                   rng.choice([1, 2, 3, 3, 5]))
            made += 1
            if made >= count:
                return
