"""GPU-vs-GPU consistency checks that need no CPU oracle (seconds instead of minutes): the strict scan and Tier 1 on
text views at odd addresses against the 16-byte aligned view, and the hinted strict scan against the plain one, on the
texts of tests/test_gpu_detect.py.  The aligned / plain paths themselves are pinned to the oracle by the test suite;
this is the quick look after a kernel change.  Prints one JSON line, exits 1 on a difference."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200 import detect
    from bwt_algorithm_b200.device_index import DeviceIndex
    from tests.util import gen_contig

    ok = {}
    text = gen_contig(120_000, 11).tobytes() + b"$"
    buf = torch.zeros(len(text) + 64, dtype=torch.uint8, device="cuda")
    base = None
    for off in (0, 1, 7, 16):
        view = buf[off:off + len(text)]
        view.copy_(torch.from_numpy(np.frombuffer(text, np.uint8).copy()))
        got = (detect.strict_rows(view, 1, 1000, 0, 3), detect.tier1_rows(view))
        if base is None:
            base = got
        ok[f"offset_{off}"] = bool(np.array_equal(got[0], base[0]) and np.array_equal(got[1], base[1]) and len(got[0]) > 1000)
    rng = np.random.default_rng(3)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    s2 = gen_contig(60_000, 8)
    s2[10_000:14_000] = ord("N")
    s2[30_000:36_000] = np.tile(acgt[rng.integers(0, 4, 300)], 20)
    s2[40_000:41_000] = np.tile(acgt[rng.integers(0, 4, 17)], 59)[:1000]
    s4 = acgt[rng.integers(0, 4, 40_000)].copy()
    s4[5_000:5_000 + 3 * 997] = np.tile(acgt[rng.integers(0, 4, 997)], 3)
    for name, s in (("planted_150k", gen_contig(150_000, 42)), ("N_and_long_units", s2),
                    ("period2", np.tile(np.frombuffer(b"AC", np.uint8), 6000)), ("unit_997", s4)):
        t = s.tobytes() + b"$"
        ix = DeviceIndex(t, build_lcp=True)
        plain = detect.strict_rows(ix.text, 1, 1000, 0, 3)
        for hl in (16, 12, 20):
            hint = detect.repeat_hint(ix.sa, ix.lcp, len(t), hl)
            got = detect.strict_rows(ix.text, 1, 1000, 0, 3, hint=hint, hint_len=hl)
            ok[f"hint_{name}_{hl}"] = bool(np.array_equal(got, plain))
    print(json.dumps({"all_ok": all(ok.values()), "checks": ok}))
    sys.exit(0 if all(ok.values()) else 1)


if __name__ == "__main__":
    main()
